"""GPU parity of the render_utils operator set (SURVEY.md §8a rows S1,S2,A1,A2,N3) through the C-ABI mirror
`plenvdb_b200.render_utils_cuda`: vs the CPU oracle, and vs the reference's own torch extension compiled for
sm_100a (oracle/_ref/render_utils_ref.so) when it travelled with the repo.  Integer outputs bit-exact."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _scene(n_rays=512, seed=11):
    from plenvdb_b200 import synth
    P = synth.scene_params(160)
    ro, rd, vd, tg = synth.ray_batch(n_rays, seed=seed)
    # edge cases: a zero direction component, a ray that misses the box, a ray starting inside
    rd[0, 1] = 0.0
    ro[1] = np.array([10, 10, 10], np.float32); rd[1] = np.array([1, 0, 0], np.float32)
    ro[2] = np.array([0.1, 0.2, -0.3], np.float32)
    return P, ro, rd, vd, tg


def _cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _ref_ext(name):
    from oracle import ref
    m = ref.torch_ext(name)
    if m is None:
        pytest.skip("oracle/_ref/%s.so not present" % name)
    return m


def test_sample_pts_on_rays_bit_exact_vs_oracle():
    from oracle import oracle as orc
    from plenvdb_b200 import render_utils_cuda as ru
    P, ro, rd, _, _ = _scene()
    want = orc.sample_pts_on_rays(ro, rd, P["xyz_min"], P["xyz_max"], P["near"], P["far"], P["stepdist"])
    got = ru.sample_pts_on_rays(_cu(ro), _cu(rd), _cu(P["xyz_min"]), _cu(P["xyz_max"]), P["near"], P["far"], P["stepdist"])
    names = ["rays_pts", "mask_outbbox", "ray_id", "step_id", "N_steps", "t_min", "t_max"]
    for n, w, g in zip(names, want, got):
        g = g.cpu().numpy()
        assert g.shape == w.shape, n
        assert np.array_equal(g, w), "%s differs" % n
    tmin, tmax = ru.infer_t_minmax(_cu(ro), _cu(rd), _cu(P["xyz_min"]), _cu(P["xyz_max"]), P["near"], P["far"])
    assert np.array_equal(tmin.cpu().numpy(), want[5]) and np.array_equal(tmax.cpu().numpy(), want[6])
    ns = ru.infer_n_samples(_cu(rd), tmin, tmax, P["stepdist"])
    assert ns.dtype == torch.int64 and np.array_equal(ns.cpu().numpy(), want[4])
    s, d = ru.infer_ray_start_dir(_cu(ro), _cu(rd), tmin)
    ws, wd = orc.infer_ray_start_dir(ro, rd, want[5])
    assert np.array_equal(s.cpu().numpy(), ws) and np.array_equal(d.cpu().numpy(), wd)


def test_sample_pts_on_rays_bit_exact_vs_reference_ext():
    ext = _ref_ext("render_utils_ref")
    from plenvdb_b200 import render_utils_cuda as ru
    P, ro, rd, _, _ = _scene(1024, seed=12)
    args = (_cu(ro), _cu(rd), _cu(P["xyz_min"]), _cu(P["xyz_max"]), P["near"], P["far"], P["stepdist"])
    want = ext.sample_pts_on_rays(*args)
    got = ru.sample_pts_on_rays(*args)
    for w, g in zip(want, got):
        assert w.dtype == g.dtype and w.shape == g.shape
        assert torch.equal(w, g)


def test_maskcache_lookup():
    from oracle import oracle as orc
    from plenvdb_b200 import render_utils_cuda as ru
    from plenvdb_b200.synth import mask_scale_shift
    rng = np.random.default_rng(3)
    shape = (33, 40, 17)
    world = rng.random(shape) < 0.3
    mn, mx = np.array([-1.3, -1.0, -0.5], np.float32), np.array([1.3, 1.2, 0.9], np.float32)
    sc, sh = mask_scale_shift(shape, mn, mx)
    xyz = (rng.random((50000, 3)) * 3.2 - 1.6).astype(np.float32)
    # exact voxel-boundary (x.5) cases for the half-away-from-zero rounding
    xyz[:100, 0] = ((np.arange(100) + 0.5) / sc[0] - sh[0] / sc[0]).astype(np.float32)
    want = orc.maskcache_lookup(world, xyz, sc, sh)
    got = ru.maskcache_lookup(_cu(world), _cu(xyz), _cu(sc), _cu(sh))
    assert got.dtype == torch.bool and np.array_equal(got.cpu().numpy(), want)
    ext = None
    from oracle import ref
    ext = ref.torch_ext("render_utils_ref")
    if ext is not None:
        assert torch.equal(ext.maskcache_lookup(_cu(world), _cu(xyz), _cu(sc), _cu(sh)), got)
    assert ru.maskcache_lookup(_cu(world), _cu(xyz[:0]), _cu(sc), _cu(sh)).numel() == 0


def test_raw2alpha_and_backward():
    from oracle import oracle as orc
    from plenvdb_b200 import render_utils_cuda as ru
    rng = np.random.default_rng(4)
    d = np.concatenate([rng.normal(0, 6, 100000), [-100, 100, 88.8, 0]]).astype(np.float32)
    shift, interval = -4.59512, 0.5
    e, a = ru.raw2alpha(_cu(d), shift, interval)
    we, wa = orc.raw2alpha(d, shift, interval)
    # libm vs libdevice differ in the last ulps of expf/powf: 1e-5 relative (north_star), bitwise vs the reference ext
    np.testing.assert_allclose(a.cpu().numpy(), wa, rtol=1e-5, atol=1e-7)
    gb = rng.standard_normal(d.size).astype(np.float32)
    g = ru.raw2alpha_backward(e, _cu(gb), interval)
    wg = orc.raw2alpha_backward(e.cpu().numpy(), gb, interval)
    np.testing.assert_allclose(g.cpu().numpy(), wg, rtol=1e-5, atol=1e-12)
    from oracle import ref
    ext = ref.torch_ext("render_utils_ref")
    if ext is not None:
        re_, ra = ext.raw2alpha(_cu(d), shift, interval)
        assert torch.equal(re_, e) and torch.equal(ra, a)
        assert torch.equal(ext.raw2alpha_backward(e, _cu(gb), interval), g)


def _ragged_alpha(rng, n_rays=300):
    lens = rng.integers(0, 40, n_rays)
    lens[:3] = [0, 1, 39]
    ray_id = np.repeat(np.arange(n_rays), lens).astype(np.int64)
    alpha = rng.random(ray_id.size).astype(np.float32) ** 2
    alpha[rng.random(alpha.size) < 0.05] = 0.999   # drives T below 1e-3 -> early stop
    return alpha, ray_id, n_rays


def test_alpha2weight_forward_backward():
    from oracle import oracle as orc
    from plenvdb_b200 import render_utils_cuda as ru
    rng = np.random.default_rng(5)
    alpha, ray_id, n = _ragged_alpha(rng)
    got = ru.alpha2weight(_cu(alpha), _cu(ray_id), n)
    want = orc.alpha2weight(alpha, ray_id, n)
    for g, w in zip(got, want):
        assert np.array_equal(g.cpu().numpy(), w)          # identical arithmetic -> bit exact incl. i_start / i_end
    gw = rng.standard_normal(alpha.size).astype(np.float32)
    gl = rng.standard_normal(n).astype(np.float32)
    grad = ru.alpha2weight_backward(_cu(alpha), *got, n, _cu(gw), _cu(gl))
    wgrad = orc.alpha2weight_backward(alpha, *want, n, gw, gl)
    assert np.array_equal(grad.cpu().numpy(), wgrad)
    from oracle import ref
    ext = ref.torch_ext("render_utils_ref")
    if ext is not None:
        rw = ext.alpha2weight(_cu(alpha), _cu(ray_id), n)
        for g, w in zip(got, rw):
            assert torch.equal(g, w)
        assert torch.equal(ext.alpha2weight_backward(_cu(alpha), *rw, n, _cu(gw), _cu(gl)), grad)
    # empty input keeps the initial values (render_utils_kernel.cu:630-632)
    e = ru.alpha2weight(_cu(alpha[:0]), _cu(ray_id[:0]), 5)
    assert e[2].cpu().numpy().tolist() == [1.0] * 5 and e[3].sum().item() == 0


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_dense_adam(mode):
    from oracle import oracle as orc
    from plenvdb_b200 import render_utils_cuda as ru
    rng = np.random.default_rng(6)
    n = 22019
    p = rng.standard_normal(n).astype(np.float32)
    m, v = np.zeros(n, np.float32), np.zeros(n, np.float32)
    perlr = rng.random(n).astype(np.float32)
    tp, tm, tv, tl = _cu(p), _cu(m), _cu(v), _cu(perlr)
    ext = None
    from oracle import ref
    ext = ref.torch_ext("adam_upd_ref")
    ep, em, ev = (tp.clone(), tm.clone(), tv.clone()) if ext is not None else (None, None, None)
    for step in range(1, 4):
        g = rng.standard_normal(n).astype(np.float32)
        g[rng.random(n) < 0.4] = 0
        tg = _cu(g)
        if mode == 0:
            ru.adam_upd(tp, tg, tm, tv, step, 0.9, 0.99, 1e-3, 1e-8)
        elif mode == 1:
            ru.masked_adam_upd(tp, tg, tm, tv, step, 0.9, 0.99, 1e-3, 1e-8)
        else:
            ru.adam_upd_with_perlr(tp, tg, tm, tv, tl, step, 0.9, 0.99, 1e-3, 1e-8)
        orc.dense_adam(p, g, m, v, perlr if mode == 2 else None, mode, step, 0.9, 0.99, 1e-3, 1e-8)
        if ext is not None:
            if mode == 0:
                ext.adam_upd(ep, tg, em, ev, step, 0.9, 0.99, 1e-3, 1e-8)
            elif mode == 1:
                ext.masked_adam_upd(ep, tg, em, ev, step, 0.9, 0.99, 1e-3, 1e-8)
            else:
                ext.adam_upd_with_perlr(ep, tg, em, ev, tl, step, 0.9, 0.99, 1e-3, 1e-8)
    assert np.array_equal(tp.cpu().numpy(), p) and np.array_equal(tm.cpu().numpy(), m) and np.array_equal(tv.cpu().numpy(), v)
    if ext is not None:
        assert torch.equal(ep, tp) and torch.equal(ev, tv)
