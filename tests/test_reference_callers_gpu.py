"""The drop-in exercised by the REFERENCE'S OWN CALLERS.

plenvdb/lib/grid.py (QueryVerticalInVDB :40-60, VDBGrid :65-137) and plenvdb/lib/masked_adam.py (VDBAdam :18-95) are the
Python code that sits on top of the boundary in the reference.  oracle/build_oracle.py byte-compiles them from where they lie
into oracle/_ref/ref_caller_*.pycode (CPython byte code; compiled artefacts of the reference like the .so files next to them: no source enters the
repo, and /root/reference is not read here).  This test loads those modules with

    sys.modules["plenvdb"]                 = plenvdb_b200.plenvdb        (what `from plenvdb import DensityVDB, ...` finds)
    torch.utils.cpp_extension.load(name=)  = plenvdb_b200.render_utils_cuda for 'render_utils_cuda' / 'adam_upd_cuda'

and drives VDBGrid.forward / autograd backward / VDBAdam.zero_grad + step through them, against the CPU oracle.
"""
import importlib.machinery
import importlib.util
import os
import sys
import types

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")


def _load_reference_caller(name):
    path = os.path.join(REF, "ref_caller_%s.pycode" % name)
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/ref_caller_%s.pycode not present (built by oracle/build_oracle.py where the reference tree exists)" % name)
    from torch.utils import cpp_extension
    from plenvdb_b200 import plenvdb as ours
    from plenvdb_b200 import render_utils_cuda as ru
    asked = []

    def fake_load(name, sources=None, **kw):      # the reference JIT-compiles its extensions at import time
        asked.append(name)
        if name in ("render_utils_cuda", "adam_upd_cuda"):
            return ru
        return types.SimpleNamespace()             # total_variation_cuda: used by DenseGrid only (VDBGrid's TV returns early, grid.py:103)

    saved_mod, saved_load = sys.modules.get("plenvdb"), cpp_extension.load
    sys.modules["plenvdb"] = ours
    cpp_extension.load = fake_load
    try:
        loader = importlib.machinery.SourcelessFileLoader("ref_caller_" + name, path)
        spec = importlib.util.spec_from_loader("ref_caller_" + name, loader)
        mod = importlib.util.module_from_spec(spec)
        loader.exec_module(mod)
    finally:
        cpp_extension.load = saved_load
        if saved_mod is None:
            sys.modules.pop("plenvdb", None)
        else:
            sys.modules["plenvdb"] = saved_mod
    return mod, asked


@pytest.fixture(scope="module")
def scene():
    from plenvdb_b200 import synth
    return synth.make_scene(64, "dense")


def _points(scene, n, seed):
    """World-space query points around the occupied voxels (most corners exist), plus a few outside the grid."""
    rng = np.random.default_rng(seed)
    R = np.array(scene["reso"], np.float32)
    occ = np.argwhere(scene["mask"])
    idx = occ[rng.integers(0, len(occ), n)] + rng.uniform(-1.5, 1.5, (n, 3))
    idx = np.clip(idx, 0, R - 1.001)
    mn, mx = scene["xyz_min"], scene["xyz_max"]
    return (idx / (R - 1) * (mx - mn) + mn).astype(np.float32)


def test_vdbgrid_and_vdbadam_of_the_reference_run_on_the_drop_in(scene):
    from oracle import oracle as orc
    grid_mod, asked = _load_reference_caller("grid")
    adam_mod, asked2 = _load_reference_caller("masked_adam")
    assert "render_utils_cuda" in asked and "adam_upd_cuda" in asked2
    R = scene["reso"]
    ws = torch.tensor(R)
    mn, mx = torch.tensor(scene["xyz_min"]), torch.tensor(scene["xyz_max"])
    # ---- the reference's modules, built the way dvgo.py:72-79 builds them
    dgrid = grid_mod.create_grid("VDBGrid", channels=1, world_size=ws, xyz_min=mn, xyz_max=mx)
    cgrid = grid_mod.create_grid("VDBGrid", channels=12, world_size=ws, xyz_min=mn, xyz_max=mx)
    assert type(dgrid).__name__ == "VDBGrid" and type(dgrid.grid).__module__ == "plenvdb_b200.plenvdb"
    dgrid.grid.copyFromDense(scene["density"].reshape(-1))
    cgrid.grid.copyFromDense(scene["k0"].reshape(-1))
    # get_dense_grid (grid.py:113-115) round trip
    np.testing.assert_array_equal(dgrid.get_dense_grid()[0, 0].numpy(), scene["density"])
    # ---- oracle twins
    oden, ok0 = orc.Grid(R, 1), orc.Grid(R, 12)
    oden.copy_from_dense(scene["density"])
    ok0.copy_from_dense(scene["k0"])
    oaux = {1: [orc.Grid(R, 1) for _ in range(3)], 12: [orc.Grid(R, 12) for _ in range(3)]}      # grad, m, v
    # ---- VDBAdam (masked_adam.py:18-45) on the two modules, skip_zero_grad like the fine stage (run.py:386-388)
    opt = adam_mod.VDBAdam([{"params": dgrid, "lr": 0.1, "skip_zero_grad": True}, {"params": cgrid, "lr": 0.1, "skip_zero_grad": True}],
                           betas=(0.9, 0.99))
    assert type(opt.densityOpt).__module__ == "plenvdb_b200.plenvdb"
    rng = np.random.default_rng(2)
    for it in range(1, 4):
        opt.zero_grad()
        for g in oaux[1][:1] + oaux[12][:1]:
            g.fill(0.0)
        xyz = _points(scene, 4000, 10 + it)
        x = torch.from_numpy(xyz)
        # ---- forward through VDBGrid.forward -> QueryVerticalInVDB.apply (grid.py:80-90, 40-52)
        dout = dgrid(x)
        cout = cgrid(x)
        assert dout.shape == (4000,) and cout.shape == (4000, 12)
        idx = ((x - mn) / (mx - mn) * (ws - 1)).float().numpy()            # grid.py:77-78, same torch expression
        wd = oden.forward(idx[:, 0], idx[:, 1], idx[:, 2]).reshape(-1)
        wc = ok0.forward(idx[:, 0], idx[:, 1], idx[:, 2])
        if it == 1:
            assert np.array_equal(dout.detach().cpu().numpy(), wd) and np.array_equal(cout.detach().cpu().numpy(), wc)     # bit-exact
        else:
            np.testing.assert_allclose(dout.detach().cpu().numpy(), wd, rtol=1e-5, atol=1e-6)
            np.testing.assert_allclose(cout.detach().cpu().numpy(), wc, rtol=1e-5, atol=1e-6)
        # ---- backward through autograd -> QueryVerticalInVDB.backward -> grid.backward (grid.py:53-60)
        gd = rng.standard_normal(4000).astype(np.float32)
        gc = rng.standard_normal((4000, 12)).astype(np.float32)
        (dout * torch.from_numpy(gd).to(dout.device)).sum().backward()
        (cout * torch.from_numpy(gc).to(cout.device)).sum().backward()
        oaux[1][0].backward(idx[:, 0], idx[:, 1], idx[:, 2], gd)
        oaux[12][0].backward(idx[:, 0], idx[:, 1], idx[:, 2], gc)
        got_gd = dgrid.grid.grad.cpu().numpy().reshape(-1)
        want_gd = oaux[1][0].get_values().reshape(-1)
        np.testing.assert_allclose(got_gd, want_gd, rtol=1e-5, atol=1e-5 * np.abs(want_gd).max())
        got_gc = cgrid.grid.grad.cpu().numpy().reshape(-1)
        want_gc = oaux[12][0].get_values().reshape(-1)
        np.testing.assert_allclose(got_gc, want_gc, rtol=1e-5, atol=1e-5 * np.abs(want_gc).max())
        # ---- VDBAdam.step (masked_adam.py:56-68): stepmode 1 on both grids
        opt.step()
        for o, (p, (g, m, v)) in ((opt.densityOpt, (oden, oaux[1])), (opt.colorOpt, (ok0, oaux[12]))):
            orc.adam_step(p, g, m, v, 1, orc.adam_stepsize(o.getLr(), o.getBeta0(), o.getBeta1(), o.getStep()), o.getEps(), o.getBeta0(), o.getBeta1())
        # the gradients agree to 1e-5 (float atomics on one side, a serial sum on the other); Adam's g / (sqrt(g^2) + eps) turns a
        # gradient that cancels to ~0 into a step of either sign, so a handful of parameters may land 2 lr apart: counted, bounded
        for got, want, lr in ((dgrid.grid.grid, oden, opt.densityOpt.getLr()), (cgrid.grid.grid, ok0, opt.colorOpt.getLr())):
            got, want = got.cpu().numpy().reshape(-1), want.get_values().reshape(-1)
            err = np.abs(got - want)
            off = err > 1e-4 * np.abs(want) + 2e-4
            assert off.mean() < 1e-4 and err.max() <= 2.0 * 3 * lr, (int(off.sum()), float(err.max()))
        if it == 2:
            opt.update_lr(0.5)                                              # masked_adam.py:47-49
    assert opt.densityOpt.getStep() == 3 and opt.colorOpt.getStep() == 3
    # setValuesOn_bymask through VDBGrid (grid.py:108-110)
    m = torch.from_numpy(scene["mask"])
    dgrid.setValuesOn_bymask(m, -7.0)
    dense = dgrid.get_dense_grid()[0, 0].numpy()
    assert np.all(dense[scene["mask"]] == -7.0)
