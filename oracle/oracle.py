"""ctypes wrapper of the CPU oracle (oracle/liboracle.so).  TEST INFRASTRUCTURE ONLY — see plenvdb_oracle.h."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liboracle.so")


def _load():
    if not os.path.exists(LIB):
        from . import build_oracle
        build_oracle.build_oracle()
    return C.CDLL(LIB)


lib = _load()
lib.orc_grid_create.restype = C.c_void_p
lib.orc_adam_stepsize.restype = C.c_float
lib.orc_adam_stepsize.argtypes = [C.c_float, C.c_float, C.c_float, C.c_int]
lib.orc_sample_pts_on_rays.restype = C.c_int64
lib.orc_merge.restype = C.c_int64

NET_N = 22019


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


class Grid:
    """orc_grid handle: tree + payload of `channels` floats per voxel."""

    def __init__(self, reso, channels=1, active=None):
        self.reso = tuple(int(r) for r in reso)
        self.channels = channels
        a = None if active is None else np.ascontiguousarray(np.asarray(active).astype(np.uint8))
        self._keep = a
        self.h = C.c_void_p(lib.orc_grid_create(self.reso[0], self.reso[1], self.reso[2], channels, _p(a)))

    def __del__(self):
        if getattr(self, "h", None):
            lib.orc_grid_destroy(self.h)
            self.h = None

    @property
    def n_leaf(self):
        return lib.orc_grid_leaf_count(self.h)

    def leaf_origins(self):
        out = np.zeros((self.n_leaf, 3), np.int32)
        lib.orc_grid_leaf_origins(self.h, _p(out))
        return out

    def leaf_masks(self):
        out = np.zeros((self.n_leaf, 8), np.uint64)
        lib.orc_grid_leaf_masks(self.h, _p(out))
        return out

    def copy_from_dense(self, dense):
        d = _f32(dense)
        assert d.size == np.prod(self.reso) * self.channels
        lib.orc_grid_copy_from_dense(self.h, _p(d))

    def to_dense(self):
        out = np.zeros(self.reso + (self.channels,), np.float32)
        lib.orc_grid_copy_to_dense(self.h, _p(out))
        return out

    def set_on_by_mask(self, mask, val):
        m = np.ascontiguousarray(np.asarray(mask).astype(np.uint8))
        lib.orc_grid_set_on_by_mask(self.h, _p(m), C.c_float(val))

    def fill(self, v):
        lib.orc_grid_fill(self.h, C.c_float(v))

    def set_values(self, plane):
        """Leaf-major payload [n_leaf, 512, channels] (the caller has checked the leaf order against leaf_origins())."""
        p = _f32(plane)
        assert p.size == self.n_leaf * 512 * self.channels
        lib.orc_grid_set_values(self.h, _p(p))

    def get_values(self):
        out = np.zeros((self.n_leaf, 512, self.channels), np.float32)
        lib.orc_grid_get_values(self.h, _p(out))
        return out

    def forward(self, x, y, z, corners=False, threads=1):
        x, y, z = _f32(x), _f32(y), _f32(z)
        n = x.size
        out = np.zeros((n, self.channels), np.float32)
        cl = np.zeros((n, 8), np.int32) if corners else None
        co = np.zeros((n, 8), np.int32) if corners else None
        lib.orc_sample_forward(self.h, _p(x), _p(y), _p(z), C.c_int64(n), _p(out), _p(cl), _p(co), threads)
        return (out, cl, co) if corners else out

    def backward(self, x, y, z, g, threads=1):
        x, y, z, g = _f32(x), _f32(y), _f32(z), _f32(g)
        lib.orc_sample_backward(self.h, _p(x), _p(y), _p(z), _p(g), C.c_int64(x.size), threads)

    def zero_grad(self):
        lib.orc_zero_grad(self.h)


def adam_stepsize(lr, b0, b1, step):
    return float(lib.orc_adam_stepsize(lr, b0, b1, step))


def adam_step(p, g, m, v, mode, stepsz, eps, b0, b1, perlr=None):
    lib.orc_adam_step(p.h, g.h, m.h, v.h, mode, C.c_float(stepsz), C.c_float(eps), C.c_float(b0), C.c_float(b1),
                      perlr.h if perlr is not None else None)


# ---- B2 ops
def infer_t_minmax(ro, rd, mn, mx, near, far):
    ro, rd, mn, mx = _f32(ro), _f32(rd), _f32(mn), _f32(mx)
    n = ro.shape[0]
    tmin, tmax = np.zeros(n, np.float32), np.zeros(n, np.float32)
    lib.orc_infer_t_minmax(_p(ro), _p(rd), _p(mn), _p(mx), C.c_float(near), C.c_float(far), n, _p(tmin), _p(tmax))
    return tmin, tmax


def infer_n_samples(rd, tmin, tmax, stepdist):
    rd, tmin, tmax = _f32(rd), _f32(tmin), _f32(tmax)
    out = np.zeros(tmin.size, np.int64)
    lib.orc_infer_n_samples(_p(rd), _p(tmin), _p(tmax), C.c_float(stepdist), tmin.size, _p(out))
    return out


def infer_ray_start_dir(ro, rd, tmin):
    ro, rd, tmin = _f32(ro), _f32(rd), _f32(tmin)
    s, d = np.zeros_like(ro), np.zeros_like(ro)
    lib.orc_infer_ray_start_dir(_p(ro), _p(rd), _p(tmin), ro.shape[0], _p(s), _p(d))
    return s, d


def sample_pts_on_rays(ro, rd, mn, mx, near, far, stepdist):
    ro, rd, mn, mx = _f32(ro), _f32(rd), _f32(mn), _f32(mx)
    n = ro.shape[0]
    args = [_p(ro), _p(rd), _p(mn), _p(mx), C.c_float(near), C.c_float(far), C.c_float(stepdist), n]
    total = lib.orc_sample_pts_on_rays(*args, None, None, None, None, None, None, None)
    pts = np.zeros((total, 3), np.float32)
    mob = np.zeros(total, np.uint8)
    rid, sid = np.zeros(total, np.int64), np.zeros(total, np.int64)
    ns = np.zeros(n, np.int64)
    tmin, tmax = np.zeros(n, np.float32), np.zeros(n, np.float32)
    lib.orc_sample_pts_on_rays(*args, _p(pts), _p(mob), _p(rid), _p(sid), _p(ns), _p(tmin), _p(tmax))
    return pts, mob.astype(bool), rid, sid, ns, tmin, tmax


def maskcache_lookup(world, xyz, scale, shift):
    w = np.ascontiguousarray(np.asarray(world).astype(np.uint8))
    xyz, scale, shift = _f32(xyz), _f32(scale), _f32(shift)
    out = np.zeros(xyz.shape[0], np.uint8)
    lib.orc_maskcache_lookup(_p(w), _p(xyz), _p(out), _p(scale), _p(shift), w.shape[0], w.shape[1], w.shape[2],
                             C.c_int64(xyz.shape[0]))
    return out.astype(bool)


def raw2alpha(density, shift, interval):
    d = _f32(density)
    e, a = np.zeros_like(d), np.zeros_like(d)
    lib.orc_raw2alpha(_p(d), C.c_float(shift), C.c_float(interval), C.c_int64(d.size), _p(e), _p(a))
    return e, a


def raw2alpha_backward(exp_d, gback, interval):
    e, g = _f32(exp_d), _f32(gback)
    out = np.zeros_like(e)
    lib.orc_raw2alpha_backward(_p(e), _p(g), C.c_float(interval), C.c_int64(e.size), _p(out))
    return out


def alpha2weight(alpha, ray_id, n_rays):
    a = _f32(alpha)
    rid = np.ascontiguousarray(ray_id, dtype=np.int64)
    w, T = np.zeros_like(a), np.ones_like(a)
    ail = np.ones(n_rays, np.float32)
    i_s, i_e = np.zeros(n_rays, np.int64), np.zeros(n_rays, np.int64)
    lib.orc_alpha2weight(_p(a), _p(rid), C.c_int64(a.size), n_rays, _p(w), _p(T), _p(ail), _p(i_s), _p(i_e))
    return w, T, ail, i_s, i_e


def alpha2weight_backward(alpha, weight, T, ail, i_s, i_e, n_rays, gw, glast):
    grad = np.zeros_like(_f32(alpha))
    lib.orc_alpha2weight_backward(_p(_f32(alpha)), _p(_f32(weight)), _p(_f32(T)), _p(_f32(ail)),
                                  _p(np.ascontiguousarray(i_s, np.int64)), _p(np.ascontiguousarray(i_e, np.int64)), n_rays,
                                  _p(_f32(gw)), _p(_f32(glast)), _p(grad))
    return grad


def dense_adam(p, g, m, v, perlr, mode, step, beta1, beta2, lr, eps):
    """In place on float32 numpy arrays."""
    lib.orc_dense_adam(_p(p), _p(_f32(g)), _p(m), _p(v), _p(perlr), C.c_int64(p.size), mode, step, C.c_float(beta1),
                       C.c_float(beta2), C.c_float(lr), C.c_float(eps))


# ---- training step
class TrainCfg(C.Structure):
    _fields_ = [("xyz_min", C.c_float * 3), ("xyz_max", C.c_float * 3), ("reso", C.c_int32 * 3),
                ("near", C.c_float), ("far", C.c_float), ("stepdist", C.c_float), ("act_shift", C.c_float),
                ("interval", C.c_float), ("fast_color_thres", C.c_float), ("bg", C.c_float),
                ("weight_main", C.c_float), ("weight_entropy_last", C.c_float), ("weight_rgbper", C.c_float),
                ("lr_density", C.c_float), ("lr_k0", C.c_float), ("lr_net", C.c_float), ("eps", C.c_float),
                ("beta0", C.c_float), ("beta1", C.c_float),
                ("den_mode", C.c_int32), ("k0_mode", C.c_int32), ("step", C.c_int32), ("n_rays_global", C.c_int32),
                ("do_update", C.c_int32), ("threads", C.c_int32)]


class TrainOut(C.Structure):
    _fields_ = [("n_steps", C.c_void_p), ("cnt_inbbox", C.c_void_p), ("cnt_mask", C.c_void_p),
                ("cnt_alpha_full", C.c_void_p), ("cnt_alpha", C.c_void_p), ("cnt_keep", C.c_void_p),
                ("alphainv_last", C.c_void_p), ("rgb_marched", C.c_void_p),
                ("loss", C.c_float * 4),
                ("M0", C.c_int64), ("M0_in", C.c_int64), ("M1", C.c_int64), ("M2", C.c_int64), ("M2_trim", C.c_int64),
                ("M3", C.c_int64),
                ("cap_keep", C.c_int64), ("keep_ray", C.c_void_p), ("keep_step", C.c_void_p), ("keep_weight", C.c_void_p),
                ("keep_rgb", C.c_void_p), ("keep_feat", C.c_void_p), ("keep_leaf", C.c_void_p), ("keep_off", C.c_void_p),
                ("net_grad", C.c_void_p),
                ("V_mask", C.c_int64), ("V_den", C.c_int64), ("V_den_grad", C.c_int64), ("V_k0", C.c_int64)]


def train_step(cfg, den, den_grad, den_m, den_v, k0, k0_grad, k0_m, k0_v, mask, net, net_m, net_v, rays_o, rays_d, viewdirs,
               target, cap_keep=0, den_perlr=None):
    """Runs orc_train_step. `cfg` is a dict of TrainCfg fields. net/net_m/net_v: float32[22019], updated in place.
    Returns a dict of numpy outputs."""
    c = TrainCfg()
    for k, v in cfg.items():
        if k in ("xyz_min", "xyz_max"):
            setattr(c, k, (C.c_float * 3)(*[float(t) for t in v]))
        elif k == "reso":
            c.reso = (C.c_int32 * 3)(*[int(t) for t in v])
        else:
            setattr(c, k, v)
    ro, rd, vd, tg = _f32(rays_o), _f32(rays_d), _f32(viewdirs), _f32(target)
    n = ro.shape[0]
    m = np.ascontiguousarray(np.asarray(mask).astype(np.uint8))
    o = TrainOut()
    res = {
        "n_steps": np.zeros(n, np.int64), "cnt_inbbox": np.zeros(n, np.int32), "cnt_mask": np.zeros(n, np.int32),
        "cnt_alpha_full": np.zeros(n, np.int32), "cnt_alpha": np.zeros(n, np.int32), "cnt_keep": np.zeros(n, np.int32),
        "alphainv_last": np.zeros(n, np.float32), "rgb_marched": np.zeros((n, 3), np.float32),
        "net_grad": np.zeros(NET_N, np.float32),
    }
    for k, a in res.items():
        setattr(o, k, a.ctypes.data)
    o.cap_keep = cap_keep
    if cap_keep:
        keep = {"keep_ray": np.zeros(cap_keep, np.int32), "keep_step": np.zeros(cap_keep, np.int32),
                "keep_weight": np.zeros(cap_keep, np.float32), "keep_rgb": np.zeros((cap_keep, 3), np.float32),
                "keep_feat": np.zeros((cap_keep, 12), np.float32), "keep_leaf": np.zeros((cap_keep, 8), np.int32),
                "keep_off": np.zeros((cap_keep, 8), np.int32)}
        for k, a in keep.items():
            setattr(o, k, a.ctypes.data)
        res.update(keep)
    lib.orc_train_step_perlr(C.byref(c), den.h, den_grad.h, den_m.h, den_v.h, k0.h, k0_grad.h, k0_m.h, k0_v.h,
                             den_perlr.h if den_perlr is not None else None, _p(m), _p(net), _p(net_m), _p(net_v), _p(ro), _p(rd), _p(vd),
                             _p(tg), n, C.byref(o))
    res["loss"] = np.array(list(o.loss), np.float32)
    for k in ("M0", "M0_in", "M1", "M2", "M2_trim", "M3", "V_mask", "V_den", "V_den_grad", "V_k0"):
        res[k] = int(getattr(o, k))
    if cap_keep:
        for k in list(keep):
            res[k] = res[k][: min(cap_keep, res["M3"])]
    return res


# ---- merged renderer
def merge(den, k0, mask):
    m = np.ascontiguousarray(np.asarray(mask).astype(np.uint8))
    n = lib.orc_merge(den.h, k0.h, _p(m), None, None, None)
    dend = np.zeros(n + 1, np.float32)
    cold = np.zeros((n + 1, k0.channels), np.float32)
    idx = np.zeros(den.reso, np.float32)
    lib.orc_merge(den.h, k0.h, _p(m), _p(dend), _p(cold), _p(idx))
    return dend, cold, idx


class RenderCfg(C.Structure):
    _fields_ = [("reso", C.c_int32 * 3), ("K", C.c_float * 9), ("xyz_min", C.c_float * 3), ("xyz_max", C.c_float * 3),
                ("near", C.c_float), ("stepdist", C.c_float), ("act_shift", C.c_float), ("interval", C.c_float),
                ("fast_color_thres", C.c_float), ("bg", C.c_float),
                ("inverse_y", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("threads", C.c_int32)]


def render(cfg, idx_grid, dendata, coldata, mlp, c2w, row_begin=0, row_end=None):
    """mlp = (w0[39,128], b0, w1[128,128], b1, w2[128,3], b2) transposed layout (run.py:98-104)."""
    c = RenderCfg()
    c.reso = (C.c_int32 * 3)(*cfg["reso"])
    c.K = (C.c_float * 9)(*[float(t) for t in np.asarray(cfg["K"]).reshape(-1)])
    c.xyz_min = (C.c_float * 3)(*[float(t) for t in cfg["xyz_min"]])
    c.xyz_max = (C.c_float * 3)(*[float(t) for t in cfg["xyz_max"]])
    for k in ("near", "stepdist", "act_shift", "interval", "fast_color_thres", "bg", "inverse_y", "H", "W", "threads"):
        setattr(c, k, cfg[k])
    row_end = cfg["H"] if row_end is None else row_end
    npix = (row_end - row_begin) * cfg["W"]
    out = np.zeros((npix, 3), np.float32)
    ns = np.zeros(npix, np.int32)
    bad = C.c_int32(0)
    w0, b0, w1, b1, w2, b2 = [_f32(t) for t in mlp]
    dd, cd, cw = _f32(dendata), _f32(coldata), _f32(c2w)
    lib.orc_render(C.byref(c), idx_grid.h, _p(dd), _p(cd), cd.shape[1], _p(w0), _p(b0), _p(w1), _p(b1), _p(w2), _p(b2),
                   _p(cw), row_begin, row_end, _p(out), _p(ns), C.byref(bad))
    return out, ns, bad.value


def march_check(cfg, idx_grid, dendata, c2w, skip_k=16, lanes=8, slot=64):
    """Exactness of the fast march of csrc/renderer.cu (empty-space skipping, lane-parallel evaluation, simulated second march)
    against the reference's step-by-step marches, pixel by pixel, on the CPU (orc_march_check)."""
    c = RenderCfg()
    c.reso = (C.c_int32 * 3)(*cfg["reso"])
    c.K = (C.c_float * 9)(*[float(t) for t in np.asarray(cfg["K"]).reshape(-1)])
    c.xyz_min = (C.c_float * 3)(*[float(t) for t in cfg["xyz_min"]])
    c.xyz_max = (C.c_float * 3)(*[float(t) for t in cfg["xyz_max"]])
    for k in ("near", "stepdist", "act_shift", "interval", "fast_color_thres", "bg", "inverse_y", "H", "W", "threads"):
        setattr(c, k, cfg[k])
    dd, cw = _f32(dendata), _f32(c2w)
    stats = np.zeros(10, np.int64)
    lib.orc_march_check(C.byref(c), idx_grid.h, _p(dd), _p(cw), int(skip_k), int(lanes), int(slot), _p(stats))
    keys = ("pixels", "with_samples", "handed_over", "not_handed_t_chain", "not_handed_count", "not_handed_slot", "mismatches",
            "steps_skipped", "steps_total", "reference_inconsistent")
    return dict(zip(keys, [int(v) for v in stats]))
