"""ctypes wrapper of the reference-backed checker libraries under oracle/_ref/ (built by build_oracle.py from
/root/reference in place).  TEST INFRASTRUCTURE ONLY.  `host()` needs no GPU; `gpu()` runs the reference's own
sm_100a-compiled kernels and therefore needs the B200."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_REF = os.path.join(HERE, "_ref")


def available(kind="host"):
    return os.path.exists(os.path.join(_REF, "libref_%s.so" % kind))


_libs = {}


def _lib(kind):
    if kind not in _libs:
        lib = C.CDLL(os.path.join(_REF, "libref_%s.so" % kind))
        lib.ref_grid_create.restype = C.c_void_p
        _libs[kind] = lib
    return _libs[kind]


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


class RefGrid:
    """NanoVDB grid(s) built by the reference's GridBuilder: 1 channel = NanoGrid<float>, 3k = k NanoGrid<Vec3f>."""

    def __init__(self, reso, channels=1, active=None, kind="host"):
        self.kind = kind
        self.lib = _lib(kind)
        self.reso = tuple(int(r) for r in reso)
        self.channels = channels
        a = None if active is None else np.ascontiguousarray(np.asarray(active).astype(np.uint8))
        self.h = C.c_void_p(self.lib.ref_grid_create(self.reso[0], self.reso[1], self.reso[2], channels, _p(a)))

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.ref_grid_destroy(self.h)
            self.h = None

    @property
    def n_leaf(self):
        return self.lib.ref_grid_leaf_count(self.h)

    def leaf_origins(self):
        out = np.zeros((self.n_leaf, 3), np.int32)
        self.lib.ref_grid_leaf_origins(self.h, _p(out))
        return out

    def leaf_masks(self):
        out = np.zeros((self.n_leaf, 8), np.uint64)
        self.lib.ref_grid_leaf_masks(self.h, _p(out))
        return out

    # host-side access through the reference accessor (GPU build: download first / upload after)
    def copy_from_dense_host(self, dense):
        self.lib.ref_grid_download(self.h)
        self.lib.ref_grid_copy_from_dense_host(self.h, _p(_f32(dense)))
        self.lib.ref_grid_upload(self.h)

    def to_dense(self):
        self.lib.ref_grid_download(self.h)
        out = np.zeros(self.reso + (self.channels,), np.float32)
        self.lib.ref_grid_copy_to_dense_host(self.h, _p(out))
        return out

    def probe_corners(self, x, y, z):
        x, y, z = _f32(x), _f32(y), _f32(z)
        cl, co = np.zeros((x.size, 8), np.int32), np.zeros((x.size, 8), np.int32)
        self.lib.ref_probe_corners(self.h, _p(x), _p(y), _p(z), C.c_int64(x.size), _p(cl), _p(co))
        return cl, co

    def host_forward(self, x, y, z, threads=1):
        x, y, z = _f32(x), _f32(y), _f32(z)
        out = np.zeros((x.size, self.channels), np.float32)
        self.lib.ref_host_forward(self.h, _p(x), _p(y), _p(z), C.c_int64(x.size), _p(out), threads)
        return out

    def host_backward(self, x, y, z, g, threads=1):
        x, y, z, g = _f32(x), _f32(y), _f32(z), _f32(g)
        self.lib.ref_host_backward(self.h, _p(x), _p(y), _p(z), _p(g), C.c_int64(x.size), threads)

    # the reference's own CUDA kernels (kind == "gpu")
    def gpu_copy_from_dense(self, dense):
        self.lib.ref_gpu_copy_from_dense(self.h, _p(_f32(dense)))

    def gpu_copy_from_dense_dev(self, dense_cuda_tensor):
        """dense_cuda_tensor: contiguous float32 CUDA tensor [rx, ry, rz(, channels)] on the current device."""
        self.lib.ref_gpu_copy_from_dense_dev(self.h, C.c_void_p(dense_cuda_tensor.data_ptr()))

    def gpu_set_on_by_mask(self, mask, val):
        m = np.ascontiguousarray(np.asarray(mask).astype(np.uint8))
        self.lib.ref_gpu_set_on_by_mask(self.h, _p(m), C.c_float(val))

    def gpu_forward(self, x, y, z):
        x, y, z = _f32(x), _f32(y), _f32(z)
        out = np.zeros((x.size, self.channels), np.float32)
        self.lib.ref_gpu_forward(self.h, _p(x), _p(y), _p(z), int(x.size), _p(out))
        return out

    def gpu_backward(self, x, y, z, g):
        x, y, z, g = _f32(x), _f32(y), _f32(z), _f32(g)
        self.lib.ref_gpu_backward(self.h, _p(x), _p(y), _p(z), _p(g), int(x.size))

    def gpu_zero_grad(self):
        self.lib.ref_gpu_zero_grad(self.h)


def gpu_adam(p, g, m, v, mode, stepsz, eps, b0, b1, perlr=None):
    p.lib.ref_gpu_adam(p.h, g.h, m.h, v.h, mode, C.c_float(stepsz), C.c_float(eps), C.c_float(b0), C.c_float(b1),
                       perlr.h if perlr is not None else None)


def gpu_render(idx_grid, dendata, coldata, mlp, reso, K, xyz_min, xyz_max, near, stepdist, act_shift, interval, thres, bg,
               inverse_y, H, W, c2w):
    """render_an_image_cuda of the reference (renderer.cu:370-424). Returns (rgb[H*W,3], n_samples[H*W], seconds)."""
    lib = idx_grid.lib
    w0, b0, w1, b1, w2, b2 = [_f32(t) for t in mlp]
    dd, cd = _f32(dendata), _f32(coldata)
    out = np.zeros((H * W, 3), np.float32)
    ns = np.zeros(H * W, np.int32)
    sec = C.c_float(0)
    r = np.ascontiguousarray(reso, np.int32)
    lib.ref_gpu_render(idx_grid.h, _p(dd), _p(cd), int(dd.size), int(cd.shape[1]), _p(w0), _p(b0), _p(w1), _p(b1), _p(w2), _p(b2),
                       _p(r), _p(_f32(K).reshape(-1)), _p(_f32(xyz_min)), _p(_f32(xyz_max)), C.c_float(near), C.c_float(stepdist),
                       C.c_float(act_shift), C.c_float(interval), C.c_float(thres), C.c_float(bg), int(bool(inverse_y)), int(H),
                       int(W), _p(_f32(c2w).reshape(-1)), _p(out), _p(ns), C.byref(sec))
    return out, ns, float(sec.value)


def torch_ext(name):
    """Import oracle/_ref/<name>.so (the reference's torch extension compiled for sm_100a) or return None."""
    path = os.path.join(_REF, name + ".so")
    if not os.path.exists(path):
        return None
    import importlib.util
    import torch  # noqa: F401  (must be loaded before the extension)
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod
