"""CPU proof of the exactness claims behind the renderer's fast march (plenvdb_b200/csrc/renderer.cu): the oracle restates the
fast march — runs of steps whose index box cannot touch a leaf replaced by their `t += steplen` chain, steps evaluated several
at a time before the ordered transmittance bookkeeping, the reference's second march simulated on the first march's corner
values — and compares it, pixel by pixel and bit for bit, with the reference's own two step-by-step marches
(renderer.cu:222-268, 312-367).  The -m gpu tests check the kernels; this checks the argument on many more rays."""
import pytest


def _setup(variant, reso=64):
    from oracle import oracle as orc
    from plenvdb_b200 import synth
    scene = synth.make_scene(reso, variant)
    den, k0 = orc.Grid(scene["reso"], 1, scene["active"]), orc.Grid(scene["reso"], 12, scene["active"])
    den.copy_from_dense(scene["density"])
    k0.copy_from_dense(scene["k0"])
    wd, wc, widx = orc.merge(den, k0, scene["mask"])
    og = orc.Grid(scene["reso"], 1, widx != 0)
    og.copy_from_dense(widx)
    return scene, og, wd


@pytest.mark.parametrize("variant", ["dense", "sparse"])
def test_fast_march_equals_the_reference_marches(variant):
    from oracle import oracle as orc
    from plenvdb_b200 import synth
    scene, og, wd = _setup(variant)
    H, W = 72, 88
    base = dict(reso=scene["reso"], K=synth.intrinsics(H, W), xyz_min=scene["xyz_min"], xyz_max=scene["xyz_max"], near=scene["near"],
                stepdist=scene["stepdist"], act_shift=scene["act_shift"], interval=scene["interval"],
                fast_color_thres=scene["fast_color_thres"], bg=scene["bg"], H=H, W=W, threads=8)
    seen = dict(handed=0, samples=0, skipped=0, total=0, fallback=0)
    for cam, inverse_y, skip_k, lanes, slot in [(0, 0, 16, 8, 64), (3, 0, 16, 1, 64), (5, 0, 8, 4, 6), (2, 1, 16, 8, 64), (6, 0, 32, 8, 64)]:
        c2w = synth.render_cameras(8)[cam].copy()
        if inverse_y:
            c2w[:3, 1:3] *= -1
        st = orc.march_check(dict(base, inverse_y=inverse_y), og, wd, c2w, skip_k=skip_k, lanes=lanes, slot=slot)
        assert st["pixels"] == H * W
        assert st["mismatches"] == 0, st
        not_handed = st["not_handed_t_chain"] + st["not_handed_count"] + st["not_handed_slot"]
        assert st["handed_over"] + not_handed == st["with_samples"]
        if slot >= 64:
            assert st["not_handed_slot"] == 0
        # a pixel whose simulated count differs from pass 1's is exactly what the reference calls an inconsistent ray
        assert st["not_handed_count"] <= st["reference_inconsistent"] + st["not_handed_t_chain"] + 5
        seen["handed"] += st["handed_over"]; seen["samples"] += st["with_samples"]; seen["fallback"] += not_handed
        seen["skipped"] += st["steps_skipped"]; seen["total"] += st["steps_total"]
    assert seen["samples"] > 1500, "degenerate views"
    assert seen["handed"] > 0.8 * seen["samples"]              # the hand-over is the rule, the second march the exception
    assert seen["skipped"] > 0.5 * seen["total"]               # most of a ray is empty space
