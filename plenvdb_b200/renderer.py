"""``MGRenderer`` — mirror of the reference's merged-VDB renderer class (plenvdb/lib/vdb/plenvdb.h:933-1068,
bound in plenvdb.cpp:135-170; driven by run.py:70-127, 157-167), plus the merge step of
plenvdb/vdb_compression.py:19-59 done on the device instead of through pyopenvdb.

Same method names and numpy contract; ``render_rows_torch`` / ``render_torch`` are the zero-copy additions used by
the tile-sharded multi-GPU path.
"""
import ctypes as C
import os
import time

import numpy as np
import torch

from . import _lib
from .tree import Topology


def merge_grids(density, k0, mask):
    """vdb_compression.py:19-59 on the device.  density: DensityVDB, k0: ColorVDB, mask: bool [reso] (mask_cache.mask).
    Returns (den[N+1] f32, col[N+1,C] f32 — both rounded through fp16 —, idx_dense int32 [reso] 1-based, N)."""
    dev = density.device
    if torch.is_tensor(mask):
        m = mask.to(dev).bool().contiguous()
    else:
        m = torch.as_tensor(np.ascontiguousarray(np.asarray(mask).astype(np.bool_))).to(dev)
    reso = tuple(m.shape)
    flat = m.reshape(-1)
    row = torch.cumsum(flat.to(torch.int32), 0, dtype=torch.int32) * flat.to(torch.int32)   # idxs = arange(1, N+1) in C order
    n = int(flat.sum().item())
    cdim = k0.ndim
    den = torch.zeros(n + 1, dtype=torch.float32, device=dev)
    col = torch.zeros((n + 1, cdim), dtype=torch.float32, device=dev)
    assert k0.topo is density.topo or k0.topo.n_leaf == density.topo.n_leaf
    _lib.call("pvdb_merge_gather", density.topo.ref, _lib.ptr(density.grid), _lib.ptr(k0.grid), cdim, _lib.ptr(row), reso[0],
              reso[1], reso[2], _lib.ptr(den), _lib.ptr(col), _lib.current_stream())
    return den, col, row.reshape(reso), n


def save_merged(basepath, den, col, idx_dense):
    """mergeddata.npz with fp16 `den`, `col` (vdb_compression.py:56-59) + mergedidxs.vdb, a FloatGrid of 1-based row ids."""
    np.savez_compressed(os.path.join(basepath, "mergeddata"), den=den.cpu().numpy().astype(np.float16),
                        col=col.cpu().numpy().astype(np.float16))
    from . import vdbio
    vdbio.save_dense_as_vdb(os.path.join(basepath, "mergedidxs.vdb"), idx_dense.cpu().numpy().astype(np.float32), name="density")


class MGRenderer:
    """MGRenderer(dcol, dpe, dhid, dout) (plenvdb.h:936-945)."""

    def __init__(self, dcol=12, dpe=27, dhid=128, dout=3, device="cuda", use_tensor_cores=True, skip_empty=True, px_entries=64,
                 cap_per_pixel=6):
        self.dcol, self.dpe, self.dhid, self.dout = int(dcol), int(dpe), int(dhid), int(dout)
        self.dev = torch.device(device)
        self.flags = [False] * 5   # load_data, load_params, setScene, setKwargs, input_a_c2w (plenvdb.h:1056)
        self.timer = 0.0
        self.cfg = _lib.pvdb_render_cfg()
        self.cfg.dcol, self.cfg.dpe, self.cfg.dhid, self.cfg.dout = self.dcol, self.dpe, self.dhid, self.dout
        self.cfg.use_tensor_cores = int(bool(use_tensor_cores))   # tcgen05 3xTF32 MLP (fp32 CUDA-core MLP when 0)
        self._scratch_rows = -1
        self.skip_empty = bool(skip_empty)   # skip runs of march steps that cannot touch a leaf (bit-identical results)
        self._skip_key, self._skip_bits, self._topo_version = None, None, 0
        # pass 1 hands its first px_entries kept samples per pixel over to the sample list (0: every pixel is marched twice)
        self.px_entries = int(px_entries)
        self.cap_per_pixel = int(cap_per_pixel)    # capacity of the frame's sample list, in kept samples per pixel (overflow is counted)
        self.c2w = torch.zeros(16, dtype=torch.float32, device=self.dev)
        self.out = None

    # ---- setup (plenvdb.h:959-1020)
    def load_data_dense(self, den, col, idx_dense):
        """den [N+1], col [N+1,dcol] (fp32 values), idx_dense [reso] int (1-based row ids, 0 = inactive)."""
        idx = np.ascontiguousarray(np.asarray(idx_dense.cpu() if torch.is_tensor(idx_dense) else idx_dense)).astype(np.int32)
        self.idx_topo = Topology.from_mask(idx != 0, device=self.dev)       # copyFromArray: active <=> value != background
        plane = self.idx_topo.new_plane(1)
        # leaf planes from the dense ids
        dense_f = torch.from_numpy(idx.astype(np.float32)).to(self.dev).contiguous()
        _lib.call("pvdb_copy_from_dense", self.idx_topo.ref, _lib.ptr(plane), 1, _lib.ptr(dense_f), idx.shape[0], idx.shape[1],
                  idx.shape[2], _lib.current_stream())
        self.idx_plane = plane.reshape(-1).to(torch.int32).contiguous()   # int(acc.getValue()) (renderer.cu:202-209)
        self.dendata = torch.as_tensor(np.asarray(den.cpu() if torch.is_tensor(den) else den, np.float32)).to(self.dev).contiguous()
        self.coldata = torch.as_tensor(np.asarray(col.cpu() if torch.is_tensor(col) else col, np.float32)).to(self.dev).reshape(-1, self.dcol).contiguous()
        assert self.dendata.numel() == self.coldata.shape[0]
        self._topo_version += 1
        self._scratch_rows = -1          # the buffer struct holds pointers into the data just replaced
        self.flags[0] = True

    def load_data(self, den, col, vdb_path, N):
        """Reference signature (plenvdb.h:959-983): flat den [N], col [N*dcol], path of mergedidxs.vdb, N = rows."""
        from . import vdbio
        if os.path.exists(vdb_path + ".npy"):            # round-1 interim files
            idx = np.load(vdb_path + ".npy")
        else:
            idx = vdbio.load_vdb_as_dense(vdb_path, tuple(int(r) for r in self.cfg.reso) if self.flags[2] else None)
        self.load_data_dense(np.asarray(den, np.float32).reshape(-1)[:N], np.asarray(col, np.float32).reshape(N, self.dcol), idx)

    def load_params(self, w0, b0, w1, b1, w2, b2):
        """Transposed weights exactly as run.py:98-104 passes them: w0 [39*128], w1 [128*128], w2 [128*3]."""
        up = lambda a: torch.as_tensor(np.ascontiguousarray(np.asarray(a, np.float32)).reshape(-1)).to(self.dev)
        self.w0, self.b0, self.w1, self.b1, self.w2, self.b2 = up(w0), up(b0), up(w1), up(b1), up(w2), up(b2)
        assert self.w0.numel() == (self.dcol + self.dpe) * self.dhid and self.w2.numel() == self.dhid * self.dout
        self._scratch_rows = -1
        self.flags[1] = True

    def setScene(self, reso, K, xyz_min, xyz_max):
        self.cfg.reso = (C.c_int32 * 3)(*[int(r) for r in reso])
        self.cfg.K = (C.c_float * 9)(*[float(v) for v in np.asarray(K, np.float32).reshape(-1)])
        self.cfg.xyz_min = (C.c_float * 3)(*[float(v) for v in xyz_min])
        self.cfg.xyz_max = (C.c_float * 3)(*[float(v) for v in xyz_max])
        self.flags[2] = True

    def setKwargs(self, near, far, stepdist, act_shift, interval, fast_color_thres, bg, inverse_y, h, w):
        c = self.cfg
        c.near, c.far = float(near), 1e9      # `far` is ignored like plenvdb.h:1008
        c.stepdist, c.act_shift, c.interval = float(stepdist), float(act_shift), float(interval)
        c.fast_color_thres, c.bg, c.inverse_y = float(fast_color_thres), float(bg), int(bool(inverse_y))
        c.H, c.W = int(h), int(w)
        self.flags[3] = True

    def input_a_c2w(self, c2w):
        self.c2w.copy_(torch.as_tensor(np.asarray(c2w, np.float32).reshape(-1)[:16]), non_blocking=True)
        self.flags[4] = True

    # ---- execution
    def _ensure_skip_bits(self):
        """Dilated block map of the index tree for the empty-space skipping; rebuilt when the tree or the resolution changes."""
        key = (self._topo_version, tuple(int(r) for r in self.cfg.reso)) if self.skip_empty else None
        if key != self._skip_key:
            self._skip_key, self._skip_bits = key, None
            if key is not None:
                rx, ry, rz = key[1]
                words = int(_lib.lib.pvdb_render_block_bits_words(rx, ry, rz))
                self._skip_bits = torch.zeros(max(words, 1), dtype=torch.int32, device=self.dev)
                _lib.call("pvdb_render_block_bits", self.idx_topo.ref, rx, ry, rz, _lib.ptr(self._skip_bits), _lib.current_stream())
        if getattr(self, "bufs", None) is not None:
            self.bufs.skip_bits = self._skip_bits.data_ptr() if self._skip_bits is not None else None

    def _ensure_scratch(self, rows):
        if rows == self._scratch_rows:
            self._ensure_skip_bits()
            return
        npix = rows * self.cfg.W
        i32 = dict(dtype=torch.int32, device=self.dev)
        f32 = dict(dtype=torch.float32, device=self.dev)
        cap = max(int(npix * self.cap_per_pixel), 4096)
        self.s = dict(n_samples=torch.zeros(npix, **i32), i_starts=torch.zeros(npix + 1, **i32), tmins=torch.zeros(npix, **f32),
                      tmaxs=torch.zeros(npix, **f32), scan_tmp=torch.zeros(npix // 4096 + 3, **i32), s_ray=torch.zeros(cap, **i32),
                      s_weight=torch.zeros(cap, **f32), s_feat=torch.zeros(cap, 12, **f32), s_rgb=torch.zeros(cap, 3, **f32),
                      counters=torch.zeros(8, **i32), w_img=torch.zeros(256 * 1024 // 4, **i32),
                      active_list=torch.zeros(npix, **i32))
        b = _lib.pvdb_render_bufs()
        b.idx_tree = C.pointer(self.idx_topo.c)
        b.idx_plane, b.dendata, b.coldata = self.idx_plane.data_ptr(), self.dendata.data_ptr(), self.coldata.data_ptr()
        for k in ("w0", "b0", "w1", "b1", "w2", "b2"):
            setattr(b, k, getattr(self, k).data_ptr())
        for k, t in self.s.items():
            setattr(b, k, t.data_ptr())
        b.cap_samples = cap
        if self.px_entries > 0:
            self.s["px_scratch"] = torch.empty(npix * self.px_entries * 2, **f32)
            self.s["fallback_list"] = torch.zeros(npix, **i32)
            b.px_scratch, b.fallback_list = self.s["px_scratch"].data_ptr(), self.s["fallback_list"].data_ptr()
            b.px_entries = self.px_entries
        self.bufs = b
        self._scratch_rows = rows
        self._ensure_skip_bits()

    def render_rows_torch(self, c2w_dev, row_begin, row_end, out=None):
        """Render rows [row_begin,row_end) for a device-resident c2w (float32[16]); returns a CUDA tensor [rows, W, 3]."""
        assert all(self.flags[:4]), "load_data, load_params, setScene and setKwargs must be called first"
        rows = row_end - row_begin
        self._ensure_scratch(rows)
        if out is None:
            out = torch.empty((rows, self.cfg.W, 3), dtype=torch.float32, device=self.dev)
        _lib.call("pvdb_render_rows", C.byref(self.cfg), C.byref(self.bufs), _lib.ptr(c2w_dev), row_begin, row_end, _lib.ptr(out),
                  _lib.current_stream())
        return out

    def interleaved_rows(self, band_rows, rank, world):
        return int(_lib.lib.pvdb_interleaved_rows(self.cfg.H, int(band_rows), int(rank), int(world)))

    def render_interleaved_torch(self, c2w_dev, band_rows, rank, world, frame_out=None, out=None):
        """Render the rows that fall to `rank` when groups of `band_rows` rows are dealt round-robin to `world` ranks; returns the
        band [rows, W, 3] (local rows in ascending image order).  frame_out: optional full [H, W, 3] CUDA tensor (or a raw device
        address, possibly another GPU's memory) that also receives every pixel of the band at its place in the frame."""
        assert all(self.flags[:4]), "load_data, load_params, setScene and setKwargs must be called first"
        rows = self.interleaved_rows(band_rows, rank, world)
        self._ensure_scratch(rows)
        if out is None:
            out = torch.empty((rows, self.cfg.W, 3), dtype=torch.float32, device=self.dev)
        fo = C.c_void_p(frame_out) if isinstance(frame_out, int) else _lib.ptr(frame_out)
        _lib.call("pvdb_render_rows_interleaved", C.byref(self.cfg), C.byref(self.bufs), _lib.ptr(c2w_dev), int(band_rows), int(rank),
                  int(world), _lib.ptr(out), fo, _lib.current_stream())
        return out

    def render_frame_sharded(self, peers, c2w_dev, band_rows, frame_no):
        """This rank's part of pvdb_render_frame_sharded (see dist.PeerFrame): interleaved rows, stored by the composite kernel
        straight into root's frame buffer over NVLink, one signal per rank instead of a gather."""
        assert all(self.flags[:4]), "load_data, load_params, setScene and setKwargs must be called first"
        rows = max(self.interleaved_rows(band_rows, peers.rank, peers.world), 1)
        self._ensure_scratch(rows)
        if getattr(self, "_band", None) is None or self._band.shape[0] != rows:
            self._band = torch.empty((rows, self.cfg.W, 3), dtype=torch.float32, device=self.dev)
        _lib.call("pvdb_render_frame_sharded", C.byref(self.cfg), C.byref(self.bufs), C.byref(peers), _lib.ptr(c2w_dev),
                  int(band_rows), int(frame_no), _lib.ptr(self._band), _lib.current_stream())

    def render_an_image(self):
        """plenvdb.h:1027-1036: silently does nothing until all five setup calls happened."""
        if not all(self.flags):
            return
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        self.out = self.render_rows_torch(self.c2w, 0, self.cfg.H)
        torch.cuda.synchronize()
        self.timer += time.perf_counter() - t1

    def output_an_image(self):
        """-> float32 numpy [H*W*3] (plenvdb.h:1038-1046)."""
        if not all(self.flags) or self.out is None:
            return None
        return self.out.reshape(-1).cpu().numpy()

    def resetTimer(self):
        self.timer = 0.0

    def getTimer(self):
        return self.timer

    def counters(self):
        c = self.s["counters"].cpu().numpy()
        return dict(total=int(c[0]), overflow=int(c[1]), inconsistent=int(c[2]), remarched=int(c[4]))

    def set_px_entries(self, n):
        """Change the hand-over slot size (0 = march every pixel twice like the reference); takes effect on the next call."""
        self.px_entries = int(n)
        self._scratch_rows = -1

    def launches_last_call(self):
        return int(_lib.lib.pvdb_last_launch_count())
