"""GPU parity of the merged-VDB renderer: device merge vs the oracle's restatement of
vdb_compression.py, and MGRenderer vs (a) the oracle and (b) the reference's own render_an_image_cuda compiled for
sm_100a.  Per-pixel sample counts and segment offsets bit-exact; RGB within 1e-5 (rays whose two reference passes
disagree, SURVEY App. A.9b, are excluded and counted)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def merged():
    from oracle import oracle as orc
    from plenvdb_b200 import synth
    from plenvdb_b200.fused import build_scene_grids
    from plenvdb_b200.renderer import merge_grids
    scene = synth.make_scene(96, "dense")
    den, k0 = build_scene_grids(scene)
    dend, cold, idx, n = merge_grids(den, k0, scene["mask"])
    oden, ok0 = orc.Grid(scene["reso"], 1), orc.Grid(scene["reso"], 12)
    oden.copy_from_dense(scene["density"])
    ok0.copy_from_dense(scene["k0"])
    wd, wc, widx = orc.merge(oden, ok0, scene["mask"])
    return scene, (dend, cold, idx, n), (wd, wc, widx)


def test_merge_matches_oracle(merged):
    scene, (dend, cold, idx, n), (wd, wc, widx) = merged
    assert n == int(scene["mask"].sum()) == wd.size - 1
    assert np.array_equal(idx.cpu().numpy(), widx.astype(np.int32))
    assert np.array_equal(dend.cpu().numpy(), wd)        # fp16 rounding is exact arithmetic
    assert np.array_equal(cold.cpu().numpy(), wc)
    assert float(dend[0]) == 0.0 and float(cold[0].abs().sum()) == 0.0


def _renderer(scene, dend, cold, idx, H, W, inverse_y=False, use_tc=True):
    from plenvdb_b200 import synth
    from plenvdb_b200.plenvdb import MGRenderer
    net = synth.rgbnet_init()
    w0, b0, w1, b1, w2, b2 = synth.unpack_net(net)
    mlp = (np.ascontiguousarray(w0.T), b0, np.ascontiguousarray(w1.T), b1, np.ascontiguousarray(w2.T), b2)   # run.py:98-104
    K = synth.intrinsics(H, W)
    r = MGRenderer(12, 27, 128, 3, use_tensor_cores=use_tc)
    r.load_data_dense(dend, cold, idx)
    r.load_params(mlp[0].reshape(-1), mlp[1], mlp[2].reshape(-1), mlp[3], mlp[4].reshape(-1), mlp[5])
    r.setScene(list(scene["reso"]), K.reshape(-1), scene["xyz_min"], scene["xyz_max"])
    r.setKwargs(scene["near"], 6.0, scene["stepdist"], scene["act_shift"], scene["interval"], scene["fast_color_thres"], scene["bg"],
                inverse_y, H, W)
    return r, mlp, K


@pytest.mark.parametrize("inverse_y", [False, True])
def test_render_matches_oracle(merged, inverse_y):
    from oracle import oracle as orc
    from plenvdb_b200 import synth
    scene, (dend, cold, idx, n), (wd, wc, widx) = merged
    H = W = 120
    r, mlp, K = _renderer(scene, dend, cold, idx, H, W, inverse_y)
    assert r.output_an_image() is None          # nothing happens before input_a_c2w (plenvdb.h:1029)
    c2w = synth.render_cameras(8)[3].copy()
    if inverse_y:
        c2w[:3, 1:3] *= -1      # OpenCV-style camera (y down, z forward) so the inverse_y rays still look at the object
    r.input_a_c2w(c2w.reshape(-1))
    r.render_an_image()
    img = r.output_an_image().reshape(H * W, 3)
    og = orc.Grid(scene["reso"], 1, widx != 0)
    og.copy_from_dense(widx)
    cfg = dict(reso=scene["reso"], K=K, xyz_min=scene["xyz_min"], xyz_max=scene["xyz_max"], near=scene["near"],
               stepdist=scene["stepdist"], act_shift=scene["act_shift"], interval=scene["interval"],
               fast_color_thres=scene["fast_color_thres"], bg=scene["bg"], inverse_y=int(inverse_y), H=H, W=W, threads=8)
    want, wns, bad = orc.render(cfg, og, wd, wc, mlp, c2w)
    ns = r.s["n_samples"].cpu().numpy()
    assert ns.sum() > 2000, "degenerate view"
    # libm vs libdevice expf/powf may flip a threshold on a handful of samples; the GPU reference test is the bit-exact one
    assert (ns != wns).mean() < 2e-3
    same = ns == wns
    np.testing.assert_allclose(img[same], want[same], rtol=1e-5, atol=3e-6)
    assert np.array_equal(r.s["i_starts"].cpu().numpy()[:-1], np.concatenate([[0], np.cumsum(ns)[:-1]]))
    # row-band rendering (tile sharding) reproduces the full frame exactly
    top = r.render_rows_torch(r.c2w, 0, 50).reshape(-1, 3).cpu().numpy()
    bot = r.render_rows_torch(r.c2w, 50, H).reshape(-1, 3).cpu().numpy()
    assert np.array_equal(np.concatenate([top, bot]), img)


@pytest.mark.parametrize("use_tc", [True, False])
def test_render_matches_reference_kernels(merged, use_tc):
    from oracle import ref
    from plenvdb_b200 import synth
    if not ref.available("gpu"):
        pytest.skip("oracle/_ref/libref_gpu.so not present")
    scene, (dend, cold, idx, n), (wd, wc, widx) = merged
    H = W = 200
    r, mlp, K = _renderer(scene, dend, cold, idx, H, W, use_tc=use_tc)
    rg = ref.RefGrid(scene["reso"], 1, widx != 0, kind="gpu")
    rg.gpu_copy_from_dense(widx)
    for cam in (0, 5):
        c2w = synth.render_cameras(8)[cam]
        r.input_a_c2w(c2w.reshape(-1))
        r.render_an_image()
        img = r.output_an_image().reshape(H * W, 3)
        ns = r.s["n_samples"].cpu().numpy()
        want, wns, _ = ref.gpu_render(rg, wd, wc, mlp, scene["reso"], K, scene["xyz_min"], scene["xyz_max"], scene["near"],
                                      scene["stepdist"], scene["act_shift"], scene["interval"], scene["fast_color_thres"],
                                      scene["bg"], False, H, W, c2w)
        assert np.array_equal(ns, wns), "per-pixel sample counts differ from the reference kernel on %d pixels" % (ns != wns).sum()
        bad = r.counters()["inconsistent"]
        ok = np.ones(H * W, bool)
        if bad:   # the reference overruns its segments on these rays; exclude the rays and their successors' garbage
            ok = np.abs(img - want).max(1) < 1e-3
            assert (~ok).sum() <= 4 * bad
        np.testing.assert_allclose(img[ok], want[ok], rtol=1e-5, atol=3e-6)


@pytest.mark.parametrize("world,band_rows", [(1, 4), (3, 4), (8, 4), (4, 5), (2, 64)])
def test_interleaved_row_groups_assemble_the_exact_frame(merged, world, band_rows):
    """Tile sharding for N GPUs, exercised rank by rank on one GPU: groups of band_rows rows dealt round-robin.
    Each rank's band equals its rows of the full frame bit for bit, and the frame_out stores assemble the whole frame."""
    from plenvdb_b200 import synth
    from plenvdb_b200 import dist as pdist
    scene, (dend, cold, idx, n), _ = merged
    H, W = 122, 120      # H is not a multiple of band_rows * world: the last group is ragged
    r, mlp, K = _renderer(scene, dend, cold, idx, H, W)
    c2w = torch.from_numpy(synth.render_cameras(8)[2].reshape(-1).copy()).cuda()
    full = r.render_rows_torch(c2w, 0, H).clone()
    assert int(r.s["n_samples"].sum()) > 2000, "degenerate view"
    frame = torch.full((H, W, 3), -1.0, device="cuda")
    total = 0
    for rank in range(world):
        rows = pdist.interleaved_rows_of(H, band_rows, rank, world)
        assert len(rows) == r.interleaved_rows(band_rows, rank, world)
        total += len(rows)
        if not rows:
            continue
        band = r.render_interleaved_torch(c2w, band_rows, rank, world, frame_out=frame)
        assert torch.equal(band, full[torch.tensor(rows, device="cuda")]), "rank %d" % rank
        assert r.counters()["overflow"] == 0
    assert total == H
    assert torch.equal(frame, full)
    # without frame_out nothing but the band is written, and the band is the same
    rows = pdist.interleaved_rows_of(H, band_rows, 0, world)
    assert torch.equal(r.render_interleaved_torch(c2w, band_rows, 0, world), full[torch.tensor(rows, device="cuda")])


@pytest.mark.parametrize("variant", ["dense", "sparse"])
def test_empty_space_skipping_is_bit_identical(variant):
    """Runs of march steps that cannot touch a leaf are skipped (only their `t += steplen` chain is evaluated).  Everything
    the frame produces — per-pixel counts, tightened t ranges, offsets, samples, RGB — equals the step-by-step march bit for bit."""
    from plenvdb_b200 import synth
    from plenvdb_b200.fused import build_scene_grids
    from plenvdb_b200.renderer import merge_grids
    scene = synth.make_scene(96, variant)
    den, k0 = build_scene_grids(scene)
    dend, cold, idx, n = merge_grids(den, k0, scene["mask"])
    H, W = 150, 170
    out = {}
    for skip in (False, True):
        r, mlp, K = _renderer(scene, dend, cold, idx, H, W)
        r.skip_empty = skip
        res = []
        for cam in (1, 4, 6):
            c2w = torch.from_numpy(synth.render_cameras(8)[cam].reshape(-1).copy()).cuda()
            img = r.render_rows_torch(c2w, 0, H).clone()
            tot = r.counters()["total"]
            res.append((img, r.s["n_samples"].clone(), r.s["i_starts"].clone(), r.s["tmins"].clone(), r.s["tmaxs"].clone(),
                        r.s["s_weight"][:tot].clone(), r.s["s_ray"][:tot].clone(), r.s["s_feat"][:tot].clone(), r.counters()))
        assert (r.bufs.skip_bits is not None) == skip
        out[skip] = res
    for a, b in zip(out[False], out[True]):
        assert a[8] == b[8] and a[8]["total"] > 2000 and a[8]["overflow"] == 0
        for x, y in zip(a[:8], b[:8]):
            assert torch.equal(x, y)


def test_sample_hand_over_equals_the_second_march(merged):
    """Pass 1 hands its kept samples over to the sample list (px_entries per pixel); pixels with more samples, or whose
    restarted t chain would differ, are marched a second time like in the reference.  Whatever the slot size — 0 (every pixel
    marched twice), 2 (most pixels fall back) or 32 — the sample list and the frame are the same bit for bit."""
    from plenvdb_b200 import synth
    scene, (dend, cold, idx, n), _ = merged
    H, W = 150, 170
    out = {}
    for P in (0, 2, 32):
        r, mlp, K = _renderer(scene, dend, cold, idx, H, W)
        r.set_px_entries(P)
        res = []
        for cam in (1, 4, 6):
            c2w = torch.from_numpy(synth.render_cameras(8)[cam].reshape(-1).copy()).cuda()
            img = r.render_rows_torch(c2w, 0, H).clone()
            c = r.counters()
            tot = c["total"]
            res.append((img, r.s["n_samples"].clone(), r.s["i_starts"].clone(), r.s["s_weight"][:tot].clone(), r.s["s_ray"][:tot].clone(),
                        r.s["s_feat"][:tot].clone(), c))
        out[P] = res
    for P in (2, 32):
        for a, b in zip(out[0], out[P]):
            assert a[6]["total"] == b[6]["total"] > 2000 and b[6]["overflow"] == 0 and a[6]["inconsistent"] == b[6]["inconsistent"]
            for x, y in zip(a[:6], b[:6]):
                assert torch.equal(x, y)
    active = int((out[0][0][1] > 0).sum())
    big = int((out[0][0][1] > 2).sum())
    assert out[0][0][6]["remarched"] == 0                       # nothing is handed over, nothing falls back
    assert big <= out[2][0][6]["remarched"] <= active and big > 100
    assert out[32][0][6]["remarched"] <= active // 20           # the restarted chain almost always reproduces pass 1's


@pytest.mark.parametrize("px_entries", [2, 64])
def test_lane_parallel_first_pass_is_bit_identical(merged, px_entries):
    """The two forms of the first pass — one pixel per thread (k_render_pass1) and probe + 8 lanes per hit pixel
    (k_render_probe, k_render_march_lanes) — give the same per-pixel counts, tightened ranges, hand-over decisions, sample list
    and frame, bit for bit, for the whole frame and for a row band, with a slot so small that most pixels fall back to the
    second march and with the default slot.  (The library picks one by the share of the frame a call renders.)"""
    from plenvdb_b200 import _lib, synth
    scene, (dend, cold, idx, n), _ = merged
    H, W = 160, 168
    res = {}
    try:
        for lanes in (0, 1):
            _lib.lib.pvdb_debug_set_render_lanes(lanes)
            r, mlp, K = _renderer(scene, dend, cold, idx, H, W)
            r.set_px_entries(px_entries)
            out = []
            for cam, (lo, hi) in ((1, (0, H)), (4, (40, 104)), (6, (0, H))):
                c2w = torch.from_numpy(synth.render_cameras(8)[cam].reshape(-1).copy()).cuda()
                img = r.render_rows_torch(c2w, lo, hi).clone()
                c = r.counters()
                tot = c["total"]
                out.append((img, r.s["n_samples"].clone(), r.s["i_starts"].clone(), r.s["tmins"].clone(), r.s["tmaxs"].clone(),
                            r.s["s_weight"][:tot].clone(), r.s["s_ray"][:tot].clone(), r.s["s_feat"][:tot].clone(), c))
            res[lanes] = out
    finally:
        _lib.lib.pvdb_debug_set_render_lanes(-1)
    for a, b in zip(res[0], res[1]):
        assert a[8] == b[8] and a[8]["total"] > 1000 and a[8]["overflow"] == 0
        hit = a[1] > 0                      # tmins / tmaxs of pixels without samples are never read (the probe leaves its resume point there)
        assert torch.equal(a[3][hit], b[3][hit]) and torch.equal(a[4][hit], b[4][hit])
        for k in (0, 1, 2, 5, 6, 7):
            assert torch.equal(a[k], b[k]), k
