// render_lanes_draft.cu — DRAFT, NOT PART OF THE LIBRARY AND NEVER RUN ON A GPU.
//
// Next step for the renderer's first pass: `k_render_pass1` is bounded by the serial instruction
// stream of the warps that cover the object (profiles/prof_render_pass1_r01f.md).  This file splits it in two:
//
//   k_render_probe        thread per pixel, empty-space skipping only: finds the `t` in front of the first run of steps that
//                         may touch a leaf.  ~93 % of the pixels of the bench frame end here (no samples, background colour).
//   k_render_march_lanes  LANES = 8 lanes per hit pixel (4 pixels per warp): the steps of a run are evaluated 8 at a time, one
//                         per lane (position, active test, 8 corner densities, both alphas), then the ordered bookkeeping —
//                         transmittance, thresholds, early stop, the simulated second march — runs identically in all 8 lanes
//                         on values exchanged by shuffles.  Values of steps behind an early stop are computed and ignored.
//
// The ORDER (values of `lanes` steps first, bookkeeping second) is proven bit-identical to the reference's step-by-step marches
// on the CPU by oracle `orc_march_check(lanes = 8)` (tests/test_fast_march_cpu.py).  What is NOT verified is this CUDA
// translation: sub-warp `__shfl_sync` / `__ballot_sync` masks, the per-lane `t` chain, the list handling.  It compiles
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -Xptxas -v -x cu -c scratch/render_lanes_draft.cu -o /dev/null
// and is kept out of csrc/ until it has passed tests/test_renderer_gpu.py on a B200 (bit-identity against px_entries = 0).
//
// Wiring it in (render_impl): probe instead of pass 1 -> march_lanes -> scans -> emit -> pass 2 (fallback list) -> gather ->
// MLP -> composite.  `active_list` then holds every HIT pixel; march_lanes rewrites each entry with PX_FALLBACK_BIT (march
// again) or PX_EMPTY_BIT (no samples after all), k_render_emit must skip PX_EMPTY_BIT entries and mask both bits.
#include "../plenvdb_b200/csrc/renderer.cu"

namespace {

constexpr int LANES = 8;
constexpr int PX_EMPTY_BIT = 0x40000000;

__global__ void __launch_bounds__(256) k_render_probe(RenderConst C, const float* __restrict__ c2w, int row_begin, int rows,
                                                      int32_t* __restrict__ n_samples, float* __restrict__ tmins,
                                                      float* __restrict__ tmaxs, int32_t* __restrict__ hit_list,
                                                      int32_t* __restrict__ counters, float* __restrict__ out_rgb) {
    pvdb_pdl_wait();
    if (blockIdx.x == 0 && threadIdx.x == 0) counters[RC_FALLBACK] = 0;
    int local;
    const bool in_image = pixel_of_thread(C.W, rows, local) >= 0;
    bool hit = false;
    if (in_image) {
        Ray R;
        ray_setup(C, c2w, render_gpix(C, row_begin, local), R);
        float t = R.tmin;
        while (t < R.tmax) {
            float t1, tk;
            run_chain(t, R.steplen, R.tmax, t1, tk);
            if (!run_is_empty(C, R, t1, tk)) { hit = true; break; }
            t = tk;
        }
        // what pass 1 leaves behind for a pixel without samples (:324-329); a hit pixel's entries are rewritten by the march,
        // which resumes at tmins[local]: every run before it was empty, so nothing has happened to the ray yet
        n_samples[local] = 0;
        tmins[local] = hit ? t : R.tmin;
        tmaxs[local] = R.tmax;
        out_rgb[local * 3] = C.bg; out_rgb[local * 3 + 1] = C.bg; out_rgb[local * 3 + 2] = C.bg;
    }
    const unsigned act = __ballot_sync(0xffffffffu, hit);
    if (act) {
        const int lane = threadIdx.x & 31;
        int base = 0;
        if (lane == 0) base = atomicAdd(counters + RC_ACTIVE, __popc(act));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (hit) hit_list[base + __popc(act & ((1u << lane) - 1))] = local;
    }
}

__global__ void __launch_bounds__(256) k_render_march_lanes(RenderConst C, const float* __restrict__ c2w, int row_begin,
                                                            int32_t* __restrict__ n_samples, float* __restrict__ tmins,
                                                            float* __restrict__ tmaxs, int32_t* __restrict__ hit_list,
                                                            const int32_t* __restrict__ counters, float* __restrict__ out_rgb,
                                                            float2* __restrict__ px_scratch, int P) {
    pvdb_pdl_wait();
    const int n_hit = counters[RC_ACTIVE];
    const int lane = threadIdx.x & 31, sub = lane & (LANES - 1), gbase = lane & ~(LANES - 1);
    const unsigned gmask = ((1u << LANES) - 1u) << gbase;          // the 8 lanes that share this pixel
    const int groups = (gridDim.x * blockDim.x) / LANES;
    // every lane of a group runs the same control flow on replicated state, so the group-masked shuffles below are always
    // reached by all 8 lanes together; different groups of a warp may diverge from each other (independent thread scheduling)
    for (int slot = (blockIdx.x * blockDim.x + threadIdx.x) / LANES; slot < n_hit; slot += groups) {
        const int local = hit_list[slot];
        Ray R;
        ray_setup(C, c2w, render_gpix(C, row_begin, local), R);
        MarchState S;
        PvdbLeafCache vcache;
        float2* px_slot = px_scratch ? px_scratch + (size_t)local * P : nullptr;
        float T_cum = 1.0f, T2 = 1.0f, t = tmins[local], tmin_out = R.tmin, tmax_out = R.tmax;
        const float tmax0 = R.tmax;
        bool update_tmin = false, done = false, sim = false;
        int ns = 0, r2 = 0;
        while (!done && t < tmax0) {
            float t1, tk;
            const int k = run_chain(t, R.steplen, tmax0, t1, tk);
            if (run_is_empty(C, R, t1, tk)) { t = tk; continue; }
            for (int base = 0; base < k && !done; base += LANES) {
                const int m = min(LANES, k - base);
                // this lane's step of the round: t advanced (sub + 1) times, through the same roundings as the serial chain
                float tq = t;
#pragma unroll
                for (int q = 0; q < LANES; ++q)
                    if (q <= sub && q < m) tq = __fadd_rn(tq, R.steplen);
                bool act = false;
                float a1 = 0.f, a2 = 0.f;
                if (sub < m) {
                    float xyz[3];
                    int leaf;
                    if (step_active(C, R, S, tq, xyz, leaf)) {
                        act = true;
                        const int i = (int)xyz[0], j = (int)xyz[1], kk = (int)xyz[2];
                        const float u = __fsub_rn(xyz[0], (float)i), v = __fsub_rn(xyz[1], (float)j), w = __fsub_rn(xyz[2], (float)kk);
                        float den[8], sc[8];
#pragma unroll
                        for (int q = 0; q < 8; ++q)
                            den[q] = __ldg(C.dendata + idx_at(C, vcache, i + PVDB_CORNER[q][0], j + PVDB_CORNER[q][1], kk + PVDB_CORNER[q][2]));
                        float res = 0.f;      // trigetDensity (:191-220)
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            const int dx = PVDB_CORNER[q][0], dy = PVDB_CORNER[q][1], dz = PVDB_CORNER[q][2];
                            const float f0 = dx ? u : __fsub_rn(1.f, u), f1 = dy ? v : __fsub_rn(1.f, v), f2 = dz ? w : __fsub_rn(1.f, w);
                            res = __fmaf_rn(f2, __fmul_rn(f1, __fmul_rn(f0, den[q])), res);
                            sc[q] = __fmul_rn(__fmul_rn(f0, f1), f2);
                        }
                        float vden = __fmaf_rn(den[0], sc[0], __fmul_rn(den[1], sc[1]));      // trigetDensity2 (:271-300)
#pragma unroll
                        for (int q = 2; q < 8; ++q) vden = __fmaf_rn(den[q], sc[q], vden);
                        a1 = render_alpha(res, C.act_shift, C.interval);
                        a2 = render_alpha(vden, C.act_shift, C.interval);
                    }
                }
                // ordered bookkeeping over the steps of the round that can have an effect, identical in all lanes of the group
                unsigned todo = (__ballot_sync(gmask, act && (a1 > C.thres || a2 > C.thres)) >> gbase) & ((1u << LANES) - 1u);
                while (todo) {
                    const int q = __ffs(todo) - 1;
                    todo &= todo - 1;
                    const float tb = __shfl_sync(gmask, tq, gbase + q);
                    const float a1b = __shfl_sync(gmask, a1, gbase + q), a2b = __shfl_sync(gmask, a2, gbase + q);
                    bool kept = false;
                    if (a1b > C.thres) {
                        const float weight = __fmul_rn(T_cum, a1b);
                        T_cum = __fmul_rn(T_cum, __fsub_rn(1.f, a1b));
                        kept = weight > C.thres;
                    }
                    if (kept) {
                        ++ns;
                        if (!update_tmin) {
                            tmin_out = __fsub_rn(tb, R.steplen);
                            update_tmin = true;
                            sim = px_slot != nullptr && __fadd_rn(tmin_out, R.steplen) == tb;
                        }
                    }
                    if (sim && a2b > C.thres) {
                        const float w2 = __fmul_rn(T2, a2b);
                        T2 = __fmul_rn(T2, __fsub_rn(1.f, a2b));
                        if (w2 > C.thres) {
                            if (r2 < P && sub == 0) px_slot[r2] = make_float2(tb, w2);
                            ++r2;
                        }
                    }
                    if (kept && (double)T_cum < 1e-3) { tmax_out = tb; done = true; break; }
                }
                t = __shfl_sync(gmask, tq, gbase + m - 1);      // t after the m steps of the round
            }
        }
        if (sub == 0) {
            const bool handed = sim && r2 == ns && ns <= P;
            n_samples[local] = ns;
            tmins[local] = tmin_out;
            tmaxs[local] = tmax_out;
            if (ns > 0 && handed) {
                const float last = __fmul_rn(T2, C.bg);      // :364-365; the composite adds the samples
                out_rgb[local * 3] = last; out_rgb[local * 3 + 1] = last; out_rgb[local * 3 + 2] = last;
            }
            hit_list[slot] = ns == 0 ? (local | PX_EMPTY_BIT) : handed ? local : (local | PX_FALLBACK_BIT);
        }
    }
}

}  // namespace
