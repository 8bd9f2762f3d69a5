"""Generates tests/golden/*.npz: outputs of the REFERENCE's own code for small seeded inputs.

Needs a B200 and the reference-backed libraries under oracle/_ref/ (built in the dev container from
/root/reference by oracle/build_oracle.py; they travel with the repo snapshot):
    libref_gpu.so         the reference's unmodified densityvdb.cu / colorvdb.cu / renderer.cu for sm_100a
    render_utils_ref.so   the reference's torch extension render_utils_cuda
    adam_upd_ref.so       the reference's torch extension adam_upd_cuda
Run on the GPU box:  python tests/golden/make_golden.py gpurun_out/golden   (then copy the files into tests/golden/).
The CPU oracle is pinned against these files by tests/test_oracle_pins.py.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref  # noqa: E402
from plenvdb_b200 import synth  # noqa: E402


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def grid_ops(out):
    rng = np.random.default_rng(100)
    R = (24, 20, 17)
    active = rng.random(R) < 0.03
    pts = (rng.random((3, 400)) * np.array(R)[:, None] * 1.2 - 2).astype(np.float32)
    res = dict(R=np.array(R), active=active, pts=pts)
    for C in (1, 12):
        dense = rng.standard_normal(R + (C,)).astype(np.float32)
        g = rng.standard_normal((400, C)).astype(np.float32)
        rg = ref.RefGrid(R, C, active, kind="gpu")
        rg.gpu_copy_from_dense(dense)
        res["dense%d" % C], res["gout%d" % C] = dense, g
        res["fwd%d" % C] = rg.gpu_forward(*pts)
        res["origins"] = rg.leaf_origins()
        cl, co = rg.probe_corners(*pts)
        res["corner_leaf"], res["corner_off"] = cl, co
        gr = ref.RefGrid(R, C, active, kind="gpu")
        gr.gpu_backward(*pts, g)
        res["bwd%d" % C] = gr.to_dense()
        for mode in (0, 1):
            p, gg, m, v = (ref.RefGrid(R, C, active, kind="gpu") for _ in range(4))
            p.gpu_copy_from_dense(dense)
            gd = rng.standard_normal(R + (C,)).astype(np.float32)
            gd[rng.random(R) < 0.5] = 0
            gg.gpu_copy_from_dense(gd)
            stepsz = np.float32(0.1 * np.sqrt(np.float32(1 - np.float32(0.99))) / np.float32(1 - np.float32(0.9)))
            ref.gpu_adam(p, gg, m, v, mode, float(stepsz), 1e-8, 0.9, 0.99)
            res["adam%d_m%d_g" % (C, mode)] = gd
            res["adam%d_m%d_stepsz" % (C, mode)] = np.float32(stepsz)
            res["adam%d_m%d_p" % (C, mode)] = p.to_dense()
            res["adam%d_m%d_v" % (C, mode)] = v.to_dense()
    np.savez_compressed(os.path.join(out, "grid_ops.npz"), **res)


def render_utils(out):
    ext = ref.torch_ext("render_utils_ref")
    P = synth.scene_params(64)
    ro, rd, vd, tg = synth.ray_batch(48, H=64, W=64, K=synth.intrinsics(64, 64), seed=9)
    rd[0, 1] = 0.0
    got = ext.sample_pts_on_rays(cu(ro), cu(rd), cu(P["xyz_min"]), cu(P["xyz_max"]), P["near"], P["far"], P["stepdist"])
    res = dict(rays_o=ro, rays_d=rd, stepdist=np.float32(P["stepdist"]))
    for n, t in zip(["rays_pts", "mask_outbbox", "ray_id", "step_id", "N_steps", "t_min", "t_max"], got):
        res[n] = t.cpu().numpy()
    rng = np.random.default_rng(101)
    world = rng.random((20, 24, 17)) < 0.3
    from plenvdb_b200.synth import mask_scale_shift
    sc, sh = mask_scale_shift(world.shape, P["xyz_min"], P["xyz_max"])
    xyz = (rng.random((3000, 3)) * 3.0 - 1.5).astype(np.float32)
    res.update(world=world, mxyz=xyz, mscale=sc, mshift=sh,
               mask_out=ext.maskcache_lookup(cu(world), cu(xyz), cu(sc), cu(sh)).cpu().numpy())
    d = rng.normal(0, 6, 2000).astype(np.float32)
    e, a = ext.raw2alpha(cu(d), -4.59512, 0.5)
    gb = rng.standard_normal(2000).astype(np.float32)
    res.update(density=d, exp_d=e.cpu().numpy(), alpha=a.cpu().numpy(), gback=gb,
               r2a_grad=ext.raw2alpha_backward(e, cu(gb), 0.5).cpu().numpy())
    lens = rng.integers(0, 30, 60)
    ray_id = np.repeat(np.arange(60), lens).astype(np.int64)
    alpha = (rng.random(ray_id.size).astype(np.float32)) ** 2
    alpha[rng.random(alpha.size) < 0.05] = 0.999
    w = ext.alpha2weight(cu(alpha), cu(ray_id), 60)
    gw, gl = rng.standard_normal(alpha.size).astype(np.float32), rng.standard_normal(60).astype(np.float32)
    ga = ext.alpha2weight_backward(cu(alpha), *w, 60, cu(gw), cu(gl))
    res.update(a2w_alpha=alpha, a2w_ray_id=ray_id, a2w_weight=w[0].cpu().numpy(), a2w_T=w[1].cpu().numpy(),
               a2w_last=w[2].cpu().numpy(), a2w_i_start=w[3].cpu().numpy(), a2w_i_end=w[4].cpu().numpy(), a2w_gw=gw, a2w_gl=gl,
               a2w_grad=ga.cpu().numpy())
    adam = ref.torch_ext("adam_upd_ref")
    p = rng.standard_normal(500).astype(np.float32)
    g = rng.standard_normal(500).astype(np.float32)
    g[rng.random(500) < 0.4] = 0
    for mode, fn in ((0, adam.adam_upd), (1, adam.masked_adam_upd)):
        tp, tm, tv = cu(p), cu(np.zeros(500, np.float32)), cu(np.zeros(500, np.float32))
        fn(tp, cu(g), tm, tv, 3, 0.9, 0.99, 1e-3, 1e-8)
        res["dadam%d_p" % mode], res["dadam%d_v" % mode] = tp.cpu().numpy(), tv.cpu().numpy()
    res.update(dadam_p0=p, dadam_g=g)
    np.savez_compressed(os.path.join(out, "render_utils.npz"), **res)


def renderer(out):
    from oracle import oracle as orc
    scene = synth.make_scene(48, "dense")
    oden, ok0 = orc.Grid(scene["reso"], 1), orc.Grid(scene["reso"], 12)
    oden.copy_from_dense(scene["density"])
    ok0.copy_from_dense(scene["k0"])
    wd, wc, widx = orc.merge(oden, ok0, scene["mask"])   # exact fp16 rounding; inputs of the renderer only
    net = synth.rgbnet_init()
    w0, b0, w1, b1, w2, b2 = synth.unpack_net(net)
    mlp = (np.ascontiguousarray(w0.T), b0, np.ascontiguousarray(w1.T), b1, np.ascontiguousarray(w2.T), b2)
    H = W = 48
    K = synth.intrinsics(H, W)
    c2w = synth.render_cameras(8)[3]
    rg = ref.RefGrid(scene["reso"], 1, widx != 0, kind="gpu")
    rg.gpu_copy_from_dense(widx)
    rgb, ns, _ = ref.gpu_render(rg, wd, wc, mlp, scene["reso"], K, scene["xyz_min"], scene["xyz_max"], scene["near"],
                                scene["stepdist"], scene["act_shift"], scene["interval"], scene["fast_color_thres"], scene["bg"],
                                False, H, W, c2w)
    np.savez_compressed(os.path.join(out, "renderer.npz"), rgb=rgb, n_samples=ns, c2w=c2w, H=H, W=W, reso=48)


if __name__ == "__main__":
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden")
    os.makedirs(out, exist_ok=True)
    grid_ops(out)
    render_utils(out)
    renderer(out)
    print("golden vectors written to", out)
