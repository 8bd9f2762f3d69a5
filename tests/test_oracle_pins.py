"""Pins the CPU oracle (oracle/plenvdb_oracle.cpp) to the reference:

  * tests/golden/*.npz — outputs of the reference's OWN CUDA kernels and torch extensions compiled for sm_100a and run on
    a B200 by tests/golden/make_golden.py (the reference ships no tests or golden vectors for this path);
  * oracle/_ref/libref_host.so — the same kernel bodies executed through the reference's own NanoVDB ReadAccessor /
    GridBuilder on the host, when /root/reference was available at build time.

Integer outputs must match bit for bit.  Float outputs are bit-exact where the arithmetic is +,-,*,/,sqrt,fma in the
reference's compiled order, and within a few ulp where libm's expf/powf stand in for libdevice's.
"""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _gold(name):
    p = os.path.join(GOLD, name)
    if not os.path.exists(p):
        pytest.skip("golden file %s not generated yet (tests/golden/make_golden.py on the GPU box)" % name)
    return np.load(p)


# ------------------------------------------------------------------------------------------- golden vectors
def test_grid_ops_against_reference_kernels():
    from oracle import oracle as orc
    g = _gold("grid_ops.npz")
    R, active, pts = tuple(int(v) for v in g["R"]), g["active"], g["pts"]
    for C in (1, 12):
        og = orc.Grid(R, C, active)
        og.copy_from_dense(g["dense%d" % C])
        assert np.array_equal(og.leaf_origins(), g["origins"])
        out, cl, co = og.forward(*pts, corners=True)
        assert np.array_equal(cl, g["corner_leaf"]) and np.array_equal(co, g["corner_off"])
        assert np.array_equal(out, g["fwd%d" % C]), "forward differs from the reference kernel by %g" % np.abs(out - g["fwd%d" % C]).max()
        gr = orc.Grid(R, C, active)
        gr.backward(*pts, g["gout%d" % C])
        want = g["bwd%d" % C]
        assert np.abs(gr.to_dense() - want).max() <= 1e-5 * np.abs(want).max()   # float atomics: order differs
        for mode in (0, 1):
            p, gg, m, v = (orc.Grid(R, C, active) for _ in range(4))
            p.copy_from_dense(g["dense%d" % C])
            gg.copy_from_dense(g["adam%d_m%d_g" % (C, mode)])
            orc.adam_step(p, gg, m, v, mode, float(g["adam%d_m%d_stepsz" % (C, mode)]), 1e-8, 0.9, 0.99)
            assert np.array_equal(p.to_dense(), g["adam%d_m%d_p" % (C, mode)])
            assert np.array_equal(v.to_dense(), g["adam%d_m%d_v" % (C, mode)])


def test_render_utils_against_reference_extension():
    from oracle import oracle as orc
    from plenvdb_b200 import synth
    g = _gold("render_utils.npz")
    P = synth.scene_params(64)
    got = orc.sample_pts_on_rays(g["rays_o"], g["rays_d"], P["xyz_min"], P["xyz_max"], P["near"], P["far"], float(g["stepdist"]))
    for n, a in zip(["rays_pts", "mask_outbbox", "ray_id", "step_id", "N_steps", "t_min", "t_max"], got):
        assert np.array_equal(a, g[n]), n
    assert np.array_equal(orc.maskcache_lookup(g["world"], g["mxyz"], g["mscale"], g["mshift"]), g["mask_out"])
    e, a = orc.raw2alpha(g["density"], -4.59512, 0.5)
    np.testing.assert_allclose(a, g["alpha"], rtol=2e-6, atol=1e-7)        # libm vs libdevice expf/powf
    np.testing.assert_allclose(e, g["exp_d"], rtol=2e-6)
    np.testing.assert_allclose(orc.raw2alpha_backward(g["exp_d"], g["gback"], 0.5), g["r2a_grad"], rtol=2e-6, atol=1e-30)
    w = orc.alpha2weight(g["a2w_alpha"], g["a2w_ray_id"], 60)
    for a, n in zip(w, ["a2w_weight", "a2w_T", "a2w_last", "a2w_i_start", "a2w_i_end"]):
        assert np.array_equal(a, g[n]), n
    ga = orc.alpha2weight_backward(g["a2w_alpha"], *w, 60, g["a2w_gw"], g["a2w_gl"])
    assert np.array_equal(ga, g["a2w_grad"])
    for mode in (0, 1):
        p, m, v = g["dadam_p0"].copy(), np.zeros(500, np.float32), np.zeros(500, np.float32)
        orc.dense_adam(p, g["dadam_g"], m, v, None, mode, 3, 0.9, 0.99, 1e-3, 1e-8)
        assert np.array_equal(p, g["dadam%d_p" % mode]) and np.array_equal(v, g["dadam%d_v" % mode])


def test_renderer_against_reference_kernels():
    from oracle import oracle as orc
    from plenvdb_b200 import synth
    g = _gold("renderer.npz")
    scene = synth.make_scene(int(g["reso"]), "dense")
    oden, ok0 = orc.Grid(scene["reso"], 1), orc.Grid(scene["reso"], 12)
    oden.copy_from_dense(scene["density"])
    ok0.copy_from_dense(scene["k0"])
    wd, wc, widx = orc.merge(oden, ok0, scene["mask"])
    og = orc.Grid(scene["reso"], 1, widx != 0)
    og.copy_from_dense(widx)
    w0, b0, w1, b1, w2, b2 = synth.unpack_net(synth.rgbnet_init())
    mlp = (np.ascontiguousarray(w0.T), b0, np.ascontiguousarray(w1.T), b1, np.ascontiguousarray(w2.T), b2)
    H, W = int(g["H"]), int(g["W"])
    cfg = dict(reso=scene["reso"], K=synth.intrinsics(H, W), xyz_min=scene["xyz_min"], xyz_max=scene["xyz_max"], near=scene["near"],
               stepdist=scene["stepdist"], act_shift=scene["act_shift"], interval=scene["interval"],
               fast_color_thres=scene["fast_color_thres"], bg=scene["bg"], inverse_y=0, H=H, W=W, threads=4)
    rgb, ns, bad = orc.render(cfg, og, wd, wc, mlp, g["c2w"])
    assert g["n_samples"].sum() > 100
    same = ns == g["n_samples"]
    assert same.mean() > 0.995          # libm vs libdevice can flip a threshold on isolated samples
    np.testing.assert_allclose(rgb[same], g["rgb"][same], rtol=1e-5, atol=3e-6)


# ------------------------------------------------------------------------------------------- reference NanoVDB on the host
def _ref():
    from oracle import ref
    if not ref.available("host"):
        pytest.skip("oracle/_ref/libref_host.so not built (needs /root/reference)")
    return ref


@pytest.mark.parametrize("kind", ["dense", "mask"])
def test_tree_semantics_against_reference_nanovdb(kind):
    from oracle import oracle as orc
    ref = _ref()
    rng = np.random.default_rng(0)
    R = (40, 24, 33)
    active = None if kind == "dense" else rng.random(R) < 0.02
    for C in (1, 12):
        og, rg = orc.Grid(R, C, active), ref.RefGrid(R, C, active)
        assert og.n_leaf == rg.n_leaf
        assert np.array_equal(og.leaf_origins(), rg.leaf_origins())      # OpenToNanoVDB / GridBuilder leaf order
        assert np.array_equal(og.leaf_masks(), rg.leaf_masks())
        dense = rng.standard_normal(R + (C,)).astype(np.float32)
        og.copy_from_dense(dense)
        rg.copy_from_dense_host(dense)
        assert np.array_equal(og.to_dense(), rg.to_dense())
        pts = (rng.random((3, 5000)) * np.array(R)[:, None] * 1.2 - 2).astype(np.float32)
        out, cl, co = og.forward(*pts, corners=True)
        rcl, rco = rg.probe_corners(*pts)
        assert np.array_equal(cl, rcl) and np.array_equal(co, rco)       # isCached<Leaf> semantics incl. negative coords
        r = rg.host_forward(*pts)                                        # host build has no FMA contraction: <= 1 ulp apart
        np.testing.assert_allclose(out, r, rtol=0, atol=4e-7 * max(1.0, np.abs(r).max()))
        g = rng.standard_normal((5000, C)).astype(np.float32)
        og2, rg2 = orc.Grid(R, C, active), ref.RefGrid(R, C, active)
        og2.backward(*pts, g)
        rg2.host_backward(*pts, g)
        assert np.array_equal(og2.to_dense(), rg2.to_dense())            # same order, same products: exact


def test_nanovdb_known_answers():
    """Known-answer cases of the substrate's own unit tests, re-expressed against the oracle tree
    (openvdb/nanovdb/nanovdb/unittest/TestNanoVDB.cc:4277-4332 trilinear of a linear field, :1291-1437 leaf offsets)."""
    from oracle import oracle as orc
    R = (32, 32, 32)
    og = orc.Grid(R, 1)
    X, Y, Z = np.meshgrid(*[np.arange(r, dtype=np.float32) for r in R], indexing="ij")
    og.copy_from_dense((0.34 + 1.6 * X + 6.7 * Y - 3.5 * Z).astype(np.float32))
    rng = np.random.default_rng(2)
    p = (rng.random((3, 2000)) * 30).astype(np.float32)
    want = 0.34 + 1.6 * p[0].astype(np.float64) + 6.7 * p[1] - 3.5 * p[2]
    np.testing.assert_allclose(og.forward(*p)[:, 0], want, rtol=1e-5, atol=1e-4)
    _, cl, co = og.forward(np.float32([9.5]), np.float32([2.25]), np.float32([7.75]), corners=True)
    assert co[0].tolist() == [(1 << 6) | (2 << 3) | 7, (1 << 6) | (2 << 3) | 0, (1 << 6) | (3 << 3) | 0, (1 << 6) | (3 << 3) | 7,
                              (2 << 6) | (3 << 3) | 7, (2 << 6) | (3 << 3) | 0, (2 << 6) | (2 << 3) | 0, (2 << 6) | (2 << 3) | 7]
    assert len(set(cl[0].tolist())) == 2      # the z = 8 corners live in the next leaf
