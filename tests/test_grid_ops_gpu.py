"""GPU parity of the grid operators (SURVEY.md §8a rows T1,T2,D1,D2,C1,C2,O1,O2,O3) through the C-ABI:
CUDA path vs the CPU oracle on the same seeded inputs, and vs the reference's own CUDA kernels
(oracle/_ref/libref_gpu.so, compiled unmodified for sm_100a) when that library travelled with the repo.
Bit-exact for indices and for every value whose arithmetic order is deterministic; gradients (float atomics
in the reference) within 1e-5 relative."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

RTOL = 1e-5   # north_star tolerance for fp32 values


def _setup(active_kind, R=(40, 24, 33), seed=0):
    rng = np.random.default_rng(seed)
    active = None if active_kind == "dense" else (rng.random(R) < 0.02)
    return rng, R, active


def _points(rng, R, n):
    pts = (rng.random((3, n)) * np.array(R)[:, None] * 1.2 - 2.0).astype(np.float32)
    pts[:, :8] = np.array([[0, 0, 0], [R[0] - 1, R[1] - 1, R[2] - 1], [7.999, 7.5, 8.0], [-0.5, 3, 3], [R[0], 1, 1],
                           [8, 8, 8], [15.5, 15.5, 15.5], [7, 7, 7]], np.float32).T
    return pts


def _product_grid(R, C, active):
    from plenvdb_b200.plenvdb import ColorVDB, DensityVDB
    from plenvdb_b200.tree import Topology
    g = DensityVDB(list(R), 1) if C == 1 else ColorVDB(list(R), C)
    if active is not None:
        g._set_topology(Topology.from_mask(active))
    return g


def _ref_gpu():
    from oracle import ref
    if not ref.available("gpu"):
        pytest.skip("oracle/_ref/libref_gpu.so not present (built only where /root/reference exists)")
    return ref


@pytest.mark.parametrize("active_kind", ["dense", "sparse"])
@pytest.mark.parametrize("C", [1, 12, 3])
def test_forward_matches_oracle_and_corner_indices(active_kind, C):
    from oracle import oracle as orc
    rng, R, active = _setup(active_kind)
    dense = rng.standard_normal(R + (C,)).astype(np.float32)
    og = orc.Grid(R, C, active)
    og.copy_from_dense(dense)
    pg = _product_grid(R, C, active)
    assert pg.topo.n_leaf == og.n_leaf
    assert (pg.topo.h_leaf_origin[: og.n_leaf] == og.leaf_origins()).all()
    assert (pg.topo.h_leaf_mask[: og.n_leaf] == og.leaf_masks()).all()
    pg.copyFromDense(dense.reshape(-1))
    # dense round trip (O3): tree view of every coordinate
    assert (pg.get_dense_grid().reshape(R + (C,)) == og.to_dense()).all()
    pts = _points(rng, R, 20000)
    want, wl, wo = og.forward(*pts, corners=True)
    cl = torch.zeros((pts.shape[1], 8), dtype=torch.int32, device="cuda")
    co = torch.zeros_like(cl)
    got = pg.forward_torch(torch.from_numpy(pts).cuda(), corner_out=(cl, co)).cpu().numpy()
    assert (cl.cpu().numpy() == wl).all(), "leaf index of a corner differs"
    assert (co.cpu().numpy() == wo).all(), "voxel offset of a corner differs"
    assert np.array_equal(got, want), "forward differs from the oracle (max %g)" % np.abs(got - want).max()
    # host (numpy) contract of the reference API
    got_h = pg.forward(pts[0], pts[1], pts[2]).reshape(-1, C)
    assert np.array_equal(got_h, want)


@pytest.mark.parametrize("active_kind", ["dense", "sparse"])
@pytest.mark.parametrize("C", [1, 12])
def test_forward_backward_match_reference_kernels(active_kind, C):
    ref = _ref_gpu()
    rng, R, active = _setup(active_kind, seed=1)
    dense = rng.standard_normal(R + (C,)).astype(np.float32)
    rg = ref.RefGrid(R, C, active, kind="gpu")
    rg.gpu_copy_from_dense(dense)
    pg = _product_grid(R, C, active)
    pg.copyFromDense(dense.reshape(-1))
    assert (pg.topo.h_leaf_origin[: rg.n_leaf] == rg.leaf_origins()).all()
    pts = _points(rng, R, 30000)
    want = rg.gpu_forward(*pts)
    got = pg.forward_torch(torch.from_numpy(pts).cuda()).cpu().numpy()
    assert np.array_equal(got, want), "forward not bit-identical to the reference kernel (max %g)" % np.abs(got - want).max()
    # backward into fresh grad grids
    g = rng.standard_normal((pts.shape[1], C)).astype(np.float32)
    rgrad = ref.RefGrid(R, C, active, kind="gpu")
    rgrad.gpu_backward(*pts, g)
    pg.backward_torch(torch.from_numpy(pts).cuda(), torch.from_numpy(g).cuda())
    want_g = rgrad.to_dense()
    got_g = pg.get_dense_grid_torch(pg.grad).cpu().numpy()
    scale = np.abs(want_g).max()
    assert np.abs(got_g - want_g).max() <= RTOL * scale
    # dense dumps of the value grids agree too (copyFromDense kernels)
    assert np.array_equal(pg.get_dense_grid().reshape(R + (C,)), rg.to_dense())


@pytest.mark.parametrize("C", [1, 12])
def test_backward_matches_oracle(C):
    from oracle import oracle as orc
    rng, R, active = _setup("sparse", seed=2)
    pts = _points(rng, R, 20000)
    g = rng.standard_normal((pts.shape[1], C)).astype(np.float32)
    og = orc.Grid(R, C, active)
    og.backward(*pts, g)
    pg = _product_grid(R, C, active)
    pg.backward(pts[0], pts[1], pts[2], g.reshape(-1))     # host contract
    want, got = og.to_dense(), pg.get_dense_grid_torch(pg.grad).cpu().numpy()
    assert np.abs(got - want).max() <= RTOL * np.abs(want).max()
    # gradients land only where a leaf exists, including inactive slots of partial leaves
    assert ((got != 0) == (want != 0)).mean() > 0.9999


@pytest.mark.parametrize("C", [1, 12])
@pytest.mark.parametrize("mode", [0, 1, 2])
def test_adam_modes_bit_exact(C, mode):
    from oracle import oracle as orc
    from plenvdb_b200.plenvdb import ColorOpt, DensityOpt
    if C == 12 and mode == 2:
        pytest.skip("per-voxel lr is only wired for density (masked_adam.py:43-46)")
    rng, R, active = _setup("sparse", seed=3)
    shape = R + (C,)
    p0 = rng.standard_normal(shape).astype(np.float32)
    pg = _product_grid(R, C, active)
    pg.copyFromDense(p0.reshape(-1))
    opt = (DensityOpt if C == 1 else ColorOpt)(pg, 0.1, 1e-8, 0.9, 0.99)
    op, ogr, om, ov = (orc.Grid(R, C, active) for _ in range(4))
    op.copy_from_dense(p0)
    operlr = None
    if mode == 2:
        cnt = rng.random(R).astype(np.float32)
        opt.set_pervoxel_lr(cnt.reshape(-1))
        operlr = orc.Grid(R, 1, active)
        operlr.copy_from_dense(cnt)
    for step in range(1, 4):
        g = rng.standard_normal(shape).astype(np.float32)
        g[rng.random(R) < 0.5] = 0.0          # exercise the skip-zero-grad rule (Vec3: all three comps)
        if C == 12:
            g[..., 1] = np.where(rng.random(R) < 0.3, 0.0, g[..., 1])
        opt.set_grad(g.reshape(-1))
        ogr.copy_from_dense(g)
        opt.step(mode)
        orc.adam_step(op, ogr, om, ov, mode, orc.adam_stepsize(0.1, 0.9, 0.99, step), 1e-8, 0.9, 0.99, operlr)
        assert abs(opt.stepsize() - orc.adam_stepsize(0.1, 0.9, 0.99, step)) == 0
    assert np.array_equal(pg.get_dense_grid().reshape(shape), op.to_dense())
    assert np.array_equal(pg.get_dense_grid_torch(opt.exp_avg).cpu().numpy(), om.to_dense())
    assert np.array_equal(pg.get_dense_grid_torch(opt.exp_avg_sq).cpu().numpy(), ov.to_dense())
    # zero_grad clears active voxels only
    opt.zero_grad()
    ogr.zero_grad()
    assert np.array_equal(pg.get_dense_grid_torch(pg.grad).cpu().numpy(), ogr.to_dense())


@pytest.mark.parametrize("mode", [0, 1])
def test_adam_matches_reference_kernels(mode):
    ref = _ref_gpu()
    from plenvdb_b200.plenvdb import ColorOpt
    rng, R, active = _setup("sparse", seed=4)
    C = 12
    shape = R + (C,)
    p0 = rng.standard_normal(shape).astype(np.float32)
    g = rng.standard_normal(shape).astype(np.float32)
    g[rng.random(R) < 0.5] = 0.0
    pg = _product_grid(R, C, active)
    pg.copyFromDense(p0.reshape(-1))
    opt = ColorOpt(pg, 0.1, 1e-8, 0.9, 0.99)
    opt.set_grad(g.reshape(-1))
    opt.step(mode)
    rp, rg, rm, rv = (ref.RefGrid(R, C, active, kind="gpu") for _ in range(4))
    rp.gpu_copy_from_dense(p0)
    rg.gpu_copy_from_dense(g)
    ref.gpu_adam(rp, rg, rm, rv, mode, opt.stepsize(), 1e-8, 0.9, 0.99)
    assert np.array_equal(pg.get_dense_grid().reshape(shape), rp.to_dense())
    assert np.array_equal(pg.get_dense_grid_torch(opt.exp_avg_sq).cpu().numpy(), rv.to_dense())


def test_set_values_on_by_mask_and_nearest():
    from oracle import oracle as orc
    rng, R, active = _setup("sparse", seed=5)
    dense = rng.standard_normal(R + (1,)).astype(np.float32)
    m = rng.random(R) < 0.3
    og = orc.Grid(R, 1, active)
    og.copy_from_dense(dense)
    og.set_on_by_mask(m, -100.0)
    pg = _product_grid(R, 1, active)
    pg.copyFromDense(dense.reshape(-1))
    pg.setValuesOn_bymask(m.reshape(-1), -100.0)
    want = og.to_dense()
    assert np.array_equal(pg.get_dense_grid().reshape(R + (1,)), want)
    ijk = rng.integers(-2, 45, (3, 500)).astype(np.int32)
    got = pg.forward_single(ijk[0], ijk[1], ijk[2])
    inside = (ijk[0] >= 0) & (ijk[0] < R[0]) & (ijk[1] >= 0) & (ijk[1] < R[1]) & (ijk[2] >= 0) & (ijk[2] < R[2])
    exp = np.zeros(500, np.float32)
    exp[inside] = want[ijk[0][inside], ijk[1][inside], ijk[2][inside], 0]
    assert np.array_equal(got, exp)


def test_empty_inputs_and_errors():
    from plenvdb_b200 import _lib
    pg = _product_grid((16, 16, 16), 1, None)
    assert pg.forward(np.zeros(0, np.float32), np.zeros(0, np.float32), np.zeros(0, np.float32)).size == 0
    pg.backward(np.zeros(0, np.float32), np.zeros(0, np.float32), np.zeros(0, np.float32), np.zeros(0, np.float32))
    with pytest.raises(_lib.PvdbError):
        _lib.call("pvdb_sample_forward", pg.topo.ref, _lib.ptr(pg.grid), 5, None, None, None, 4, _lib.ptr(pg.grid), None, None,
                  _lib.current_stream())


def test_save_load_round_trip(tmp_path):
    rng, R, active = _setup("sparse", seed=6)
    dense = rng.standard_normal(R + (12,)).astype(np.float32)
    pg = _product_grid(R, 12, active)
    pg.copyFromDense(dense.reshape(-1))
    path = str(tmp_path / "finecolor.vdb")
    pg.save_to(path)
    from plenvdb_b200.plenvdb import ColorVDB
    q = ColorVDB([2, 2, 2], 12)
    q.load_from(path)
    a = pg.get_dense_grid().reshape(R + (12,))
    rr = q.reso
    b = q.get_dense_grid().reshape(tuple(rr) + (12,))
    assert np.array_equal(a[: rr[0], : rr[1], : rr[2]] * active[: rr[0], : rr[1], : rr[2], None], b)
