"""The opaque-handle C entry points (include/plenvdb_b200.h, csrc/handles.cu) driven through ctypes with plain
host buffers — no torch tensor, no Python-side state — the way a C++ / pybind11 binder of the reference's `plenvdb` module
(plenvdb/lib/vdb/plenvdb.cpp:3-172) would: DensityVDB / ColorVDB (forward, backward, copyFromDense, get_dense_grid,
setValuesOn_bymask), DensityOpt / ColorOpt (zero_grad, step in the three modes, update_lr, per-voxel lr) and MGRenderer, against
the CPU oracle."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.fixture(scope="module")
def scene():
    from plenvdb_b200 import synth
    return synth.make_scene(64, "sparse")


def _points(scene, n, seed):
    rng = np.random.default_rng(seed)
    occ = np.argwhere(scene["mask"])
    idx = occ[rng.integers(0, len(occ), n)] + rng.uniform(-1.5, 1.5, (n, 3))
    idx = np.clip(idx, 0, np.array(scene["reso"]) - 1.001).astype(np.float32)
    return [np.ascontiguousarray(idx[:, a]) for a in range(3)]


@pytest.mark.parametrize("channels", [1, 12])
def test_grid_and_optimiser_handles_match_the_oracle(scene, channels):
    from oracle import oracle as orc
    from plenvdb_b200 import _lib
    L = _lib.lib
    R = scene["reso"]
    act = np.ascontiguousarray(scene["active"].astype(np.uint8))
    dense = scene["density"] if channels == 1 else scene["k0"]
    g = L.pvdb_grid_create(R[0], R[1], R[2], channels, _p(act))
    assert g, _lib.last_error()
    try:
        reso, ch, nl = (C.c_int32 * 3)(), C.c_int32(), C.c_int32()
        _lib.call("pvdb_grid_info", g, reso, C.byref(ch), C.byref(nl))
        og, ograd, om, ov = (orc.Grid(R, channels, scene["active"]) for _ in range(4))
        assert list(reso) == list(R) and ch.value == channels and nl.value == og.n_leaf
        d = np.ascontiguousarray(dense, np.float32)
        _lib.call("pvdb_grid_copy_from_dense", g, _p(d))
        og.copy_from_dense(d)
        back = np.zeros(tuple(R) + (channels,), np.float32)
        _lib.call("pvdb_grid_copy_to_dense", g, _p(back))
        assert np.array_equal(back, og.to_dense())
        opt = L.pvdb_opt_create(g, 0.1, 1e-8, 0.9, 0.99)
        assert opt, _lib.last_error()
        rng = np.random.default_rng(3)
        for it, mode in enumerate((1, 0, 1, 2) if channels == 1 else (1, 0, 1), 1):
            x, y, z = _points(scene, 3000, 20 + it)
            out = np.zeros((3000, channels), np.float32)
            _lib.call("pvdb_grid_forward", g, _p(x), _p(y), _p(z), 3000, _p(out))
            want = og.forward(x, y, z)
            if it == 1:
                assert np.array_equal(out, want.reshape(out.shape))
            else:
                np.testing.assert_allclose(out, want.reshape(out.shape), rtol=1e-5, atol=1e-6)
            _lib.call("pvdb_opt_zero_grad", opt)
            ograd.fill(0.0)
            gout = rng.standard_normal((3000, channels)).astype(np.float32)
            _lib.call("pvdb_grid_backward", g, _p(x), _p(y), _p(z), _p(gout), 3000)
            ograd.backward(x, y, z, gout)
            if mode == 2:
                lrmap = rng.random(R).astype(np.float32)
                _lib.call("pvdb_opt_set_pervoxel_lr", opt, _p(lrmap))
                operlr = orc.Grid(R, 1, scene["active"])
                operlr.copy_from_dense(lrmap)
            _lib.call("pvdb_opt_step", opt, mode)
            step, lr = C.c_int32(), C.c_float()
            _lib.call("pvdb_opt_get", opt, C.byref(step), C.byref(lr), None, None, None)
            assert step.value == it
            orc.adam_step(og, ograd, om, ov, mode, orc.adam_stepsize(lr.value, 0.9, 0.99, it), 1e-8, 0.9, 0.99, operlr if mode == 2 else None)
            _lib.call("pvdb_grid_copy_to_dense", g, _p(back))
            want_d = og.to_dense()
            err = np.abs(back - want_d)
            off = err > 1e-4 * np.abs(want_d) + 2e-4          # gradients agree to 1e-5; Adam turns a cancelled one into +-lr (counted)
            assert off.mean() < 1e-4 and err.max() <= 2.0 * it * 0.1 + 1e-3, (mode, int(off.sum()), float(err.max()))
            if it == 2:
                _lib.call("pvdb_opt_update_lr", opt, 0.5)
        if channels == 1:
            _lib.call("pvdb_grid_set_values_on_by_mask", g, _p(np.ascontiguousarray(scene["mask"].astype(np.uint8))), -5.0)
            _lib.call("pvdb_grid_copy_to_dense", g, _p(back))
            inside = scene["mask"] & scene["active"]
            assert np.all(back[..., 0][inside] == -5.0)
        L.pvdb_opt_destroy(opt)
    finally:
        L.pvdb_grid_destroy(g)


def test_renderer_handle_matches_the_oracle():
    from oracle import oracle as orc
    from plenvdb_b200 import _lib, synth
    L = _lib.lib
    scene = synth.make_scene(64, "dense")
    R = scene["reso"]
    oden, ok0 = orc.Grid(R, 1), orc.Grid(R, 12)
    oden.copy_from_dense(scene["density"])
    ok0.copy_from_dense(scene["k0"])
    dend, cold, idx = orc.merge(oden, ok0, scene["mask"])
    net = synth.rgbnet_init()
    w0, b0, w1, b1, w2, b2 = synth.unpack_net(net)
    mlp = [np.ascontiguousarray(a, np.float32) for a in (w0.T, b0, w1.T, b1, w2.T, b2)]
    H, W = 96, 112
    K = synth.intrinsics(H, W)
    r = L.pvdb_renderer_create(12, 27, 128, 3)
    assert r, _lib.last_error()
    try:
        img = np.zeros((H, W, 3), np.float32)
        done = C.c_int(7)
        _lib.call("pvdb_renderer_render", r, _p(img), C.byref(done))
        assert done.value == 0                                              # silently nothing before the five setup calls (plenvdb.h:1027-1030)
        idx_i = np.ascontiguousarray(idx.astype(np.int32))
        _lib.call("pvdb_renderer_load_data", r, _p(np.ascontiguousarray(dend)), _p(np.ascontiguousarray(cold)), len(dend), _p(idx_i), R[0], R[1], R[2])
        _lib.call("pvdb_renderer_load_params", r, *[_p(a) for a in mlp])
        _lib.call("pvdb_renderer_set_scene", r, (C.c_int32 * 3)(*R), _p(np.ascontiguousarray(K.reshape(-1))), _p(scene["xyz_min"]), _p(scene["xyz_max"]))
        _lib.call("pvdb_renderer_set_kwargs", r, scene["near"], 6.0, scene["stepdist"], scene["act_shift"], scene["interval"], scene["fast_color_thres"],
                  scene["bg"], 0, H, W)
        oidx = orc.Grid(R, 1, idx != 0)
        oidx.copy_from_dense(idx)
        cfg = dict(reso=R, K=K, xyz_min=scene["xyz_min"], xyz_max=scene["xyz_max"], near=scene["near"], stepdist=scene["stepdist"],
                   act_shift=scene["act_shift"], interval=scene["interval"], fast_color_thres=scene["fast_color_thres"], bg=scene["bg"],
                   inverse_y=0, H=H, W=W, threads=8)
        for cam in (1, 5):
            c2w = np.ascontiguousarray(synth.render_cameras(8)[cam].reshape(-1), np.float32)
            _lib.call("pvdb_renderer_input_c2w", r, _p(c2w))
            _lib.call("pvdb_renderer_render", r, _p(img), C.byref(done))
            assert done.value == 1
            want, wns, bad = orc.render(cfg, oidx, dend, cold, mlp, c2w.reshape(4, 4))
            cnt = (C.c_int32 * 8)()
            _lib.call("pvdb_renderer_counters", r, cnt)
            assert cnt[0] == int(wns.sum()) > 1000 and cnt[1] == 0
            np.testing.assert_allclose(img.reshape(-1, 3), want, rtol=1e-5, atol=3e-6)
    finally:
        L.pvdb_renderer_destroy(r)
