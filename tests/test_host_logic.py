"""Host-side logic that needs no GPU: topology builder vs the oracle's tree, synthetic inputs, sharding helpers,
the plane container used by save_to/load_from."""
import numpy as np
import pytest


@pytest.mark.parametrize("R", [(40, 24, 33), (160, 160, 160), (130, 8, 257), (7, 7, 7)])
@pytest.mark.parametrize("kind", ["dense", "mask"])
def test_topology_builder_matches_oracle_tree(R, kind):
    from oracle import oracle as orc
    from plenvdb_b200.tree import Topology
    rng = np.random.default_rng(0)
    active = None if kind == "dense" else rng.random(R) < 0.004
    og = orc.Grid(R, 1, active)
    tp = Topology.dense(R, device="cpu") if active is None else Topology.from_mask(active, device="cpu")
    assert tp.n_leaf == og.n_leaf
    assert np.array_equal(tp.h_leaf_origin[: tp.n_leaf], og.leaf_origins())
    assert np.array_equal(tp.h_leaf_mask[: tp.n_leaf], og.leaf_masks())
    n_active = int(np.prod(R)) if active is None else int(active.sum())
    assert tp.active_voxel_count() == n_active
    # every leaf is reachable through upper -> lower tables at its own origin
    for leaf in range(0, tp.n_leaf, max(1, tp.n_leaf // 50)):
        x, y, z = tp.h_leaf_origin[leaf]
        low = tp.h_upper[((x >> 7) << 10) | ((y >> 7) << 5) | (z >> 7)]
        assert low >= 0
        assert tp.h_lower[low * 4096 + ((((x & 127) >> 3) << 8) | (((y & 127) >> 3) << 4) | ((z & 127) >> 3))] == leaf


def test_dense_160_layout_facts():
    """160^3 dense = 8000 leaves, 8 lower nodes, 1 upper; leaf 20 = (0,8,32), leaf 400 = (8,72,0)."""
    from plenvdb_b200.tree import Topology
    tp = Topology.dense((160, 160, 160), device="cpu")
    assert (tp.n_upper, tp.n_lower, tp.n_leaf) == (1, 8, 8000)
    assert tp.h_leaf_origin[1].tolist() == [0, 0, 8]
    assert tp.h_leaf_origin[20].tolist() == [0, 8, 32]
    assert tp.h_leaf_origin[400].tolist() == [8, 72, 0]


def test_partial_leaves_of_a_20_cube():
    """denseFill of 20^3 -> 27 leaves, 8000 active voxels, voxel (20,20,20) inactive in leaf 26."""
    from plenvdb_b200.tree import Topology
    tp = Topology.dense((20, 20, 20), device="cpu")
    assert tp.n_leaf == 27 and tp.active_voxel_count() == 8000
    off = ((20 & 7) << 6) | ((20 & 7) << 3) | (20 & 7)
    assert off == 292 and not (int(tp.h_leaf_mask[26][off >> 6]) >> (off & 63)) & 1


def test_synthetic_inputs_are_deterministic():
    from plenvdb_b200 import synth
    a, b = synth.make_scene(48, "sparse"), synth.make_scene(48, "sparse")
    for k in ("density", "k0", "mask", "active"):
        assert np.array_equal(a[k], b[k])
    assert a["mask"].sum() > a["occ"].sum() > 0
    assert np.all(a["density"][~a["active"]] == 0)
    r1, r2 = synth.ray_batch(64, H=40, W=40, K=synth.intrinsics(40, 40)), synth.ray_batch(64, H=40, W=40, K=synth.intrinsics(40, 40))
    for x, y in zip(r1, r2):
        assert np.array_equal(x, y)
    assert np.allclose(np.linalg.norm(r1[2], axis=1), 1, atol=1e-6)
    net = synth.rgbnet_init()
    assert net.size == 22019 and np.all(net[-3:] == 0)
    assert abs(synth.act_shift_of(1e-2) - (-4.59512)) < 1e-5


def test_shard_ranges_partition_exactly():
    from plenvdb_b200.dist import shard_range
    for n in (800, 8192, 7, 1):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_mask_scale_shift_matches_torch():
    import torch
    from plenvdb_b200.synth import mask_scale_shift
    mn, mx = torch.tensor([-1.3, -1.0, -0.7]), torch.tensor([1.3, 1.1, 0.9])
    shape = torch.tensor([160.0, 33.0, 75.0])
    scale = (shape - 1) / (mx - mn)          # grid.py:229-231
    shift = -mn * scale
    s, t = mask_scale_shift((160, 33, 75), mn.numpy(), mx.numpy())
    assert np.array_equal(s, scale.numpy()) and np.array_equal(t, shift.numpy())


def test_plane_container_round_trip(tmp_path):
    import torch
    from plenvdb_b200 import vdbio
    from plenvdb_b200.tree import Topology
    rng = np.random.default_rng(1)
    R = (24, 40, 17)
    active = rng.random(R) < 0.05
    tp = Topology.from_mask(active, device="cpu")
    plane = torch.zeros((tp.n_leaf, 512, 3))
    xyz, leaf, off = vdbio._active_coords(tp)
    assert xyz.shape[0] == active.sum()
    vals = rng.standard_normal((xyz.shape[0], 3)).astype(np.float32)
    plane[leaf, off] = torch.from_numpy(vals)
    p = str(tmp_path / "finecolor.vdb")
    vdbio.save_planes(p, tp, plane, R, ["color0"])
    tp2, plane2, reso2 = vdbio.load_planes(p, 3, "cpu")
    assert tp2.n_leaf == tp.n_leaf and np.array_equal(tp2.h_leaf_mask[: tp.n_leaf], tp.h_leaf_mask[: tp.n_leaf])
    assert torch.equal(plane2, plane)
    assert all(r2 <= r for r2, r in zip(reso2, R))
    with pytest.raises(ValueError):
        vdbio.load_planes(p, 12, "cpu")


def test_interleaved_row_groups_partition_exactly():
    """Tile sharding of the renderer: row r belongs to rank (r // band_rows) % world; the C-ABI's row count per
    rank agrees with the host enumeration and the ranks' rows partition the image, ragged last group included."""
    from plenvdb_b200 import _lib
    from plenvdb_b200 import dist as pdist
    for H, B, world in [(800, 4, 8), (800, 4, 1), (801, 4, 8), (122, 4, 2), (122, 5, 4), (7, 4, 2), (3, 4, 2), (1, 1, 8)]:
        seen = []
        for rank in range(world):
            rows = pdist.interleaved_rows_of(H, B, rank, world)
            assert len(rows) == _lib.lib.pvdb_interleaved_rows(H, B, rank, world)
            assert rows == sorted(rows)
            # local row lr -> image row, the mapping the kernels use
            want = [rank * B + (lr // B) * world * B + lr % B for lr in range(len(rows))]
            assert rows == want
            seen += rows
        assert sorted(seen) == list(range(H))
    assert _lib.lib.pvdb_interleaved_rows(0, 4, 0, 2) == 0 and _lib.lib.pvdb_interleaved_rows(8, 0, 0, 2) == 0
    assert _lib.lib.pvdb_frame_symm_bytes(800, 800) == 1024 + 2 * 800 * 800 * 3 * 4


def test_rays_of_a_view_match_the_pixel_formula():
    """fused.get_rays_of_a_view (the torch expressions of dvgo.py:470-499) against the per-pixel numpy formula the ray batches
    of every other test come from, for both y conventions; flips mirror the image."""
    import torch
    from plenvdb_b200 import synth
    from plenvdb_b200.fused import get_rays_of_a_view
    H, W = 37, 52
    K = synth.intrinsics(H, W)
    c2w = synth.render_cameras(8)[3]
    py, px = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    for inverse_y in (False, True):
        ro, rd, vd = get_rays_of_a_view(H, W, K, c2w, inverse_y=inverse_y)
        wo, wd, wv = synth.rays_of_pixels(K, c2w, px.reshape(-1), py.reshape(-1), inverse_y)
        assert ro.shape == (H, W, 3) and ro.dtype == torch.float32
        assert np.array_equal(ro.reshape(-1, 3).numpy(), wo)
        np.testing.assert_allclose(rd.reshape(-1, 3).numpy(), wd, rtol=2e-7, atol=1e-7)      # same formula, torch.sum vs numpy sum order
        np.testing.assert_allclose(vd.reshape(-1, 3).numpy(), wv, rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(np.linalg.norm(vd.numpy(), axis=-1), 1.0, rtol=1e-6)
    a = get_rays_of_a_view(H, W, K, c2w)[1]
    assert torch.equal(get_rays_of_a_view(H, W, K, c2w, flip_x=True)[1], a.flip((1,)))
    assert torch.equal(get_rays_of_a_view(H, W, K, c2w, flip_y=True)[1], a.flip((0,)))
