"""The tcgen05 (3xTF32, TMEM-resident activations) rgbnet against the fp32 CUDA-core rgbnet and the oracle:
tensor-core products are error-compensated, so the fused step must stay inside the 1e-5 parity tolerance."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.fixture(scope="module")
def setup():
    from plenvdb_b200 import synth
    scene = synth.make_scene(96, "dense")
    net = synth.rgbnet_init()
    # scale the last layer up so the logits are not tiny (default init gives |logit| << 1)
    rays = synth.ray_batch(2048, H=200, W=200, K=synth.intrinsics(200, 200), seed=777)
    return scene, net, rays


def _run(scene, net, rays, use_tc):
    from plenvdb_b200.fused import FusedTrainer, build_scene_grids
    den, k0 = build_scene_grids(scene)
    tr = FusedTrainer(scene, den, k0, scene["mask"], net, rays[0].shape[0], use_tensor_cores=use_tc)
    tr.forward_backward(*[_cu(a) for a in rays])
    torch.cuda.synchronize()
    return tr, den, k0


def test_tc_forward_matches_fp32(setup):
    scene, net, rays = setup
    a, *_ = _run(scene, net, rays, False)
    b, *_ = _run(scene, net, rays, True)
    M = a.counters()["M_keep"]
    assert M == b.counters()["M_keep"] and M > 1000
    def rows(tr, k):   # the tensor-core path keeps activations chunk-major [tile][8 chunks][feature][16 samples]
        t = tr.t[k]
        if tr.use_tc and k in ("k_h0", "k_h1"):
            t = t.reshape(-1, 8, 128, 16).permute(0, 1, 3, 2).reshape(-1, 128)
        return t[:M].cpu().numpy()
    for k in ("k_feat", "k_h0", "k_h1"):
        x, y = rows(a, k), rows(b, k)
        scale = np.abs(x).max()
        assert np.abs(x - y).max() <= 2e-6 * max(scale, 1.0), "%s: max err %g (scale %g)" % (k, np.abs(x - y).max(), scale)
    np.testing.assert_allclose(b.t["rgb_marched"].cpu().numpy(), a.t["rgb_marched"].cpu().numpy(), rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(b.t["loss"].cpu().numpy(), a.t["loss"].cpu().numpy(), rtol=1e-5)
    # tcgen05 backward (activation-gradient chain + weight-gradient GEMM) against the fp32 backward
    ga, gb = a.net_grad.cpu().numpy(), b.net_grad.cpu().numpy()
    assert np.abs(ga - gb).max() <= 2e-5 * np.abs(ga).max(), "net_grad: max err %g of %g" % (np.abs(ga - gb).max(), np.abs(ga).max())
    ka, kb = a.k0.grad.cpu().numpy(), b.k0.grad.cpu().numpy()
    assert np.abs(ka - kb).max() <= 2e-5 * np.abs(ka).max(), "k0 grad: max err %g of %g" % (np.abs(ka - kb).max(), np.abs(ka).max())
    assert np.array_equal(a.t["k0_touched"].cpu().numpy(), b.t["k0_touched"].cpu().numpy())


def test_tc_forward_with_large_weights(setup):
    """Random weights an order of magnitude above the default init: logits of O(10), exercises hi/lo cancellation."""
    scene, net, rays = setup
    rng = np.random.default_rng(0)
    big = (net * 6 + rng.standard_normal(net.size).astype(np.float32) * 0.05).astype(np.float32)
    a, *_ = _run(scene, big, rays, False)
    b, *_ = _run(scene, big, rays, True)
    M = a.counters()["M_keep"]
    x = a.t["k_h1"][:M].cpu().numpy()
    y = b.t["k_h1"].reshape(-1, 8, 128, 16).permute(0, 1, 3, 2).reshape(-1, 128)[:M].cpu().numpy()
    assert np.abs(x - y).max() <= 4e-6 * np.abs(x).max()
    np.testing.assert_allclose(b.t["rgb_marched"].cpu().numpy(), a.t["rgb_marched"].cpu().numpy(), rtol=1e-5, atol=3e-6)
    ga, gb = a.net_grad.cpu().numpy(), b.net_grad.cpu().numpy()
    assert np.abs(ga - gb).max() <= 3e-5 * np.abs(ga).max()
    ka, kb = a.k0.grad.cpu().numpy(), b.k0.grad.cpu().numpy()
    assert np.abs(ka - kb).max() <= 3e-5 * np.abs(ka).max()


def test_per_ray_embedding_table_gives_the_same_bits_as_the_per_sample_evaluation(setup):
    """pvdb_train_bufs.ray_pe is optional: with it the forward reads the view-direction embedding of a kept sample from the row
    k_ray_pe wrote for its ray, without it the producers evaluate the 24 sinf / cosf per sample (dvgo.py:354-357) — same
    expressions, so the rgbnet inputs, activations, colours and every gradient must be the same bits."""
    scene, net, rays = setup
    from plenvdb_b200.fused import FusedTrainer, build_scene_grids
    out = []
    for with_table in (True, False):
        den, k0 = build_scene_grids(scene)
        tr = FusedTrainer(scene, den, k0, scene["mask"], net, rays[0].shape[0], use_tensor_cores=True)
        if not with_table:
            tr._bufs.ray_pe = None
        tr.forward_backward(*[_cu(a) for a in rays])
        torch.cuda.synchronize()
        M = tr.counters()["M_keep"]
        out.append({k: tr.t[k].cpu().numpy().copy() for k in ("k_x", "k_h0", "k_h1", "k_mask", "rgb_marched", "loss")})
        out[-1]["net_grad"] = tr.net_grad.cpu().numpy().copy()
        out[-1]["M"] = M
    a, b = out
    assert a["M"] == b["M"] > 1000
    rows = (a["M"] + 127) // 128 * 128
    for k in ("k_x", "k_h0", "k_h1", "k_mask"):
        x, y = a[k].reshape(-1)[: rows * a[k].shape[-1]], b[k].reshape(-1)[: rows * b[k].shape[-1]]
        assert np.array_equal(x, y), k
    assert np.array_equal(a["rgb_marched"], b["rgb_marched"])
    np.testing.assert_allclose(a["loss"], b["loss"], rtol=1e-6)      # the loss terms are float atomics over the CTAs: order, not values
    assert np.array_equal(a["net_grad"], b["net_grad"])      # the weight-gradient sums are deterministic (partials in CTA order)


def test_odd_batch_whose_sample_capacity_is_not_a_whole_tile(setup):
    """cap_keep = 64 * n_rays is not a multiple of the 128-sample tile for an odd batch: the tile-major activation tensors are
    padded to whole tiles, and the step agrees with the fp32 path as for any other batch."""
    scene, net, rays = setup
    n = 1001
    sub = [a[:n] for a in rays]
    a, *_ = _run(scene, net, sub, False)
    b, *_ = _run(scene, net, sub, True)
    assert a.counters()["M_keep"] == b.counters()["M_keep"] > 500 and b.cap_keep % 128 != 0
    np.testing.assert_allclose(b.t["rgb_marched"].cpu().numpy(), a.t["rgb_marched"].cpu().numpy(), rtol=1e-5, atol=2e-6)
    ga, gb = a.net_grad.cpu().numpy(), b.net_grad.cpu().numpy()
    assert np.abs(ga - gb).max() <= 2e-5 * np.abs(ga).max()
    ka, kb = a.k0.grad.cpu().numpy(), b.k0.grad.cpu().numpy()
    assert np.abs(ka - kb).max() <= 2e-5 * np.abs(ka).max()
