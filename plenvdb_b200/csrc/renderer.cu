// renderer.cu — merged-VDB renderer (R1/R2): MGRenderer::render_an_image of the reference
// (plenvdb/lib/vdb/renderer.cu:370-424, plenvdb.h:933-1068) for a band of image rows.
//
//   P1 k_render_pass1     thread per pixel (8x4 pixel tile per warp): ray from (pixel, K, c2w), unit-box t range,
//                         fixed-step march with the reference's `t += steplen` recurrence, active test on the index
//                         grid, density trilinear through the index -> data indirection, alpha, thresholds;
//                         counts kept samples and tightens [tmin, tmax]                       (renderer.cu:222-268).
//                         Runs of steps that cannot touch a leaf are skipped (only their t chain is evaluated), and the
//                         reference's second march is simulated on the same corner values so that the pixel's samples
//                         can be handed over instead of marched again — both bit-identical, see the comments below
//   SC k_scan_*           exclusive scan of the per-pixel counts in pixel order -> i_starts    (renderer.cu:401-402)
//   EM k_render_emit      the (t, weight) slots of the pixels handed over -> their segments of the sample list
//   P2 k_render_pass2     the reference's second march (renderer.cu:312-367) for the pixels that were not handed over
//   GA k_render_gather    thread per sample: the 12 colour features (trigetColor, renderer.cu:303-310)
//   ML k_render_mlp       64-sample tiles: view PE + MLP(39->128->128->3), weight*sigmoid       (renderer.cu:83-119)
//                         (tcgen05 version: rgbnet_tc.cu)
//   CO k_render_composite per-pixel ordered sum (deterministic; the reference uses float atomics); optionally stores the
//                         pixel into a full frame that may live on another GPU (tile-sharded rendering, k_frame_*)
//
// Differences from the reference by design: no per-frame cudaMalloc/cudaFree or host sync, no race on rays_o
// (renderer.cu:128-132), pass 2 cannot overrun its segment (such rays are counted), the
// per-ray PE contribution is folded into the layer-0 tile GEMM instead of a separate HW x 128 GEMM.
#include <cuda_fp16.h>
#include <stdlib.h>
#include "common.cuh"
#include "mlp_tile.cuh"
#include "render_ray.cuh"
#include "peer_sync.cuh"

namespace {

constexpr int RC_TOTAL = 0, RC_OVERFLOW = 1, RC_INCONSISTENT = 2, RC_ACTIVE = 3;   // RC_ACTIVE: pixels with samples (pass 1 -> pass 2)
constexpr int RC_FALLBACK = 4;   // pixels whose samples pass 1 could not hand over (k_render_emit -> pass 2)
constexpr int PX_FALLBACK_BIT = (int)0x80000000;

// alpha = 1 - pow(1 + exp(v + shift), -interval) exactly as nvcc compiles renderer.cu:256 / :350: the final scale
// multiply of expf is contracted with the "+ 1" into one fma, so this expression is deliberately left to the
// compiler (checked in the PTX of this file: fma.rn.f32 ..., 0f3F800000 after ex2.approx).
__device__ __forceinline__ float render_alpha(float v_den, float act_shift, float interval) {
    return 1 - powf(1 + expf(v_den + act_shift), -interval);
}

struct MarchState {
    PvdbLeafCache cache;
};

// One march step: position, bbox test, active test.  Returns false when the step is skipped.
__device__ __forceinline__ bool step_active(const RenderConst& C, const Ray& R, MarchState& S, float t, float* xyz, int& leaf_r) {
    const float p0 = __fmaf_rn(R.rd[0], t, R.ro[0]), p1 = __fmaf_rn(R.rd[1], t, R.ro[1]), p2 = __fmaf_rn(R.rd[2], t, R.ro[2]);
    if ((0.f > p0) | (0.f > p1) | (0.f > p2) | (1.f < p0) | (1.f < p1) | (1.f < p2)) return false;
    xyz[0] = __fmul_rn(p0, C.wld[0]); xyz[1] = __fmul_rn(p1, C.wld[1]); xyz[2] = __fmul_rn(p2, C.wld[2]);
    // acc.isActive(Round<Coord>(xyz)): rintf, leaf must exist and the voxel bit be set (NanoVDB.h:773-779, 3035-3045)
    const int i = __float2int_rn(xyz[0]), j = __float2int_rn(xyz[1]), k = __float2int_rn(xyz[2]);
    const int leaf = S.cache.find(C.tree, i, j, k);
    leaf_r = leaf;
    if (leaf < 0) return false;
    return pvdb_mask_bit(C.tree.leaf_mask, leaf, pvdb_leaf_off(i, j, k));
}

__device__ __forceinline__ int idx_at(const RenderConst& C, PvdbLeafCache& cache, int x, int y, int z) {
    const int leaf = cache.find(C.tree, x, y, z);
    return leaf >= 0 ? __ldg(C.idx_plane + (size_t)leaf * 512 + pvdb_leaf_off(x, y, z)) : 0;
}

__device__ __forceinline__ int pixel_of_thread(int W, int rows, int& local) {
    // 8x4 pixel tile per warp
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int tiles_x = (W + 7) >> 3;
    const int tx = warp % tiles_x, ty = warp / tiles_x;
    const int col = tx * 8 + (lane & 7), row = ty * 4 + (lane >> 3);
    if (col >= W || row >= rows) { local = -1; return -1; }
    local = row * W + col;
    return local;
}

// Empty-space skipping.  The reference marches every step; a step has an effect only when its voxel is active, so a run of
// steps none of which can be active may be replaced by its `t += steplen` chain alone (the chain itself is kept: t must go
// through the same roundings).  For steps t_a <= t <= t_b every coordinate p(t) = fma(rd, t, ro) lies between p(t_a) and
// p(t_b) — one rounding of a linear function is monotone — and so do p * wld and rint(.): the voxels of the run lie in the
// index box spanned by the two ends (clamped to the unit box, outside of which a step is skipped anyway).  When that box
// covers at most two 8^3 blocks per axis, one bit of the dilated block map says whether any of them holds a leaf.
constexpr int SKIP_K = 16;     // steps per run: 16 x 0.5 voxel <= 8 voxels, i.e. at most two blocks per axis
__device__ __forceinline__ bool run_is_empty(const RenderConst& C, const Ray& R, float ta, float tb) {
    int b[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const float pa = __fmaf_rn(R.rd[a], ta, R.ro[a]), pb = __fmaf_rn(R.rd[a], tb, R.ro[a]);
        if (pa != pa || pb != pb) return false;                       // NaN: leave it to the step-by-step code
        const float lo = fmaxf(fminf(pa, pb), 0.f), hi = fminf(fmaxf(pa, pb), 1.f);
        if (lo > hi) return true;                                     // every step of the run is outside the box on this axis
        const int ilo = __float2int_rn(__fmul_rn(lo, C.wld[a])) >> 3, ihi = __float2int_rn(__fmul_rn(hi, C.wld[a])) >> 3;
        if (ihi - ilo > 1 || ilo < 0 || ihi >= C.nb[a]) return false;
        b[a] = ilo;
    }
    const int bit = (b[0] * C.nb[1] + b[1]) * C.nb[2] + b[2];
    return !((__ldg(C.skip_bits + (bit >> 5)) >> (bit & 31)) & 1u);
}
// Chain of up to SKIP_K steps from t (t < tmax): returns the number of steps, t1 = t after the first, tk = after the last.
__device__ __forceinline__ int run_chain(float t, float steplen, float tmax, float& t1, float& tk) {
    t1 = __fadd_rn(t, steplen);
    tk = t1;
    int k = 1;
#pragma unroll
    for (int q = 1; q < SKIP_K; ++q)
        if (tk < tmax) { tk = __fadd_rn(tk, steplen); ++k; }
    return k;
}

// Pass 1 of one pixel: counts the kept samples and tightens [tmin, tmax] (renderer.cu:222-268).
// With a per-pixel slot (px_slot, P entries of (t, weight)) pass 1 also SIMULATES pass 2 for the pixel and hands its samples
// over, so that the pixel need not be marched a second time.  The reference's second pass (renderer.cu:312-367) is not a replay
// of the first: it restarts the `t` chain from the tightened tmin = t_first - steplen and interpolates the density in another
// association order (trigetDensity2, :271-300, vs trigetDensity, :191-220), so its alphas, weights and — at a threshold — even
// its sample count can differ in the last bit.  The simulation therefore runs pass 2's own arithmetic on the corner values
// pass 1 has already loaded, from the first kept sample on, with its own transmittance and its own thresholds.  It is exact
// when the restarted chain reproduces this one ((t_first - steplen) + steplen == t_first in float); the pixel is handed over
// when in addition the simulated count equals pass 1's (a "consistent" ray) and fits the slot.  Every other pixel is marched
// again by k_render_pass2, exactly like the reference does.
__device__ __forceinline__ int pass1_march(const RenderConst& C, const float* __restrict__ c2w, int row_begin, int local,
                                           float* __restrict__ tmins, float* __restrict__ tmaxs, float2* __restrict__ px_slot, int P,
                                           bool& handed, float& T_last) {
    const int n = render_gpix(C, row_begin, local);
    Ray R;
    ray_setup(C, c2w, n, R);
    MarchState S;
    PvdbLeafCache vcache;
    float T_cum = 1.0f, t = R.tmin, tmin_out = R.tmin, tmax_out = R.tmax;
    const float tmax0 = R.tmax;
    bool update_tmin = false, done = false, sim = false;
    float T2 = 1.0f;       // pass 2's transmittance
    int ns = 0, r2 = 0;    // kept samples of pass 1 / of the simulated pass 2
    while (!done && t < tmax0) {
        int nslow = 0x7fffffff;
        if (C.skip_bits) {
            float t1, tk;
            nslow = run_chain(t, R.steplen, tmax0, t1, tk);
            if (run_is_empty(C, R, t1, tk)) { t = tk; continue; }
        }
        for (int s = 0; s < nslow && t < tmax0; ++s) {
            t = __fadd_rn(t, R.steplen);
            float xyz[3];
            int leaf;
            if (!step_active(C, R, S, t, xyz, leaf)) continue;
            const int i = (int)xyz[0], j = (int)xyz[1], k = (int)xyz[2];
            const float u = __fsub_rn(xyz[0], (float)i), v = __fsub_rn(xyz[1], (float)j), w = __fsub_rn(xyz[2], (float)k);
            float den[8];
#pragma unroll
            for (int q = 0; q < 8; ++q)
                den[q] = __ldg(C.dendata + idx_at(C, vcache, i + PVDB_CORNER[q][0], j + PVDB_CORNER[q][1], k + PVDB_CORNER[q][2]));
            // trigetDensity (:191-220): int() truncation, res += d*f0*f1*f2 -> fma(f2, f1*(f0*d), res)
            float res = 0.f;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int dx = PVDB_CORNER[q][0], dy = PVDB_CORNER[q][1], dz = PVDB_CORNER[q][2];
                const float f0 = dx ? u : __fsub_rn(1.f, u), f1 = dy ? v : __fsub_rn(1.f, v), f2 = dz ? w : __fsub_rn(1.f, w);
                res = __fmaf_rn(f2, __fmul_rn(f1, __fmul_rn(f0, den[q])), res);
            }
            const float alpha = render_alpha(res, C.act_shift, C.interval);
            bool kept = false;
            if (alpha > C.thres) {
                const float weight = __fmul_rn(T_cum, alpha);
                T_cum = __fmul_rn(T_cum, __fsub_rn(1.f, alpha));
                kept = weight > C.thres;
            }
            if (kept) {
                ++ns;
                if (!update_tmin) {
                    tmin_out = __fsub_rn(t, R.steplen);
                    update_tmin = true;
                    sim = px_slot != nullptr && __fadd_rn(tmin_out, R.steplen) == t;
                }
            }
            if (sim) {
                // pass 2 on this step: trigetDensity2 (:271-300), d0*s0 + d1*s1 + ... -> fma(d7,s7, ... fma(d2,s2, fma(d0,s0, d1*s1)))
                float sc[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const int dx = PVDB_CORNER[q][0], dy = PVDB_CORNER[q][1], dz = PVDB_CORNER[q][2];
                    const float f0 = dx ? u : __fsub_rn(1.f, u), f1 = dy ? v : __fsub_rn(1.f, v), f2 = dz ? w : __fsub_rn(1.f, w);
                    sc[q] = __fmul_rn(__fmul_rn(f0, f1), f2);
                }
                float vden = __fmaf_rn(den[0], sc[0], __fmul_rn(den[1], sc[1]));
#pragma unroll
                for (int q = 2; q < 8; ++q) vden = __fmaf_rn(den[q], sc[q], vden);
                const float alpha2 = render_alpha(vden, C.act_shift, C.interval);
                if (alpha2 > C.thres) {
                    const float w2 = __fmul_rn(T2, alpha2);
                    T2 = __fmul_rn(T2, __fsub_rn(1.f, alpha2));
                    if (w2 > C.thres) {
                        if (r2 < P) px_slot[r2] = make_float2(t, w2);
                        ++r2;
                    }
                }
            }
            if (kept && (double)T_cum < 1e-3) { tmax_out = t; done = true; break; }
        }
    }
    tmins[local] = tmin_out;
    tmaxs[local] = tmax_out;
    handed = sim && r2 == ns && ns <= P;
    T_last = T2;
    return ns;
}

__global__ void __launch_bounds__(256) k_render_pass1(RenderConst C, const float* __restrict__ c2w, int row_begin, int rows,
                                                      int32_t* __restrict__ n_samples, float* __restrict__ tmins,
                                                      float* __restrict__ tmaxs, int32_t* __restrict__ active_list,
                                                      int32_t* __restrict__ counters, float* __restrict__ out_rgb,
                                                      float2* __restrict__ px_scratch, int P) {
    pvdb_pdl_wait();
    if (blockIdx.x == 0 && threadIdx.x == 0) counters[RC_FALLBACK] = 0;   // k_render_emit of this frame counts from 0; stays readable after the frame
    int local;
    const bool in_image = pixel_of_thread(C.W, rows, local) >= 0;
    int ns = 0;
    bool handed = false;
    float T_last = 1.f;
    if (in_image) ns = pass1_march(C, c2w, row_begin, local, tmins, tmaxs, px_scratch ? px_scratch + (size_t)local * P : nullptr, P, handed, T_last);
    if (in_image) {
        n_samples[local] = ns;
        // :324-329 for a pixel without samples; :364-365 (T * bg, the composite adds the samples) for one handed over
        const float last = ns == 0 ? C.bg : __fmul_rn(T_last, C.bg);
        if (ns == 0 || handed) { out_rgb[local * 3] = last; out_rgb[local * 3 + 1] = last; out_rgb[local * 3 + 2] = last; }
    }
    // pixels that have samples are appended to the work list of pass 2 (one atomic per warp; a warp's pixels stay together)
    const unsigned act = __ballot_sync(0xffffffffu, ns > 0);
    if (act) {
        const int lane = threadIdx.x & 31;
        int base = 0;
        if (lane == 0) base = atomicAdd(counters + RC_ACTIVE, __popc(act));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (ns > 0) active_list[base + __popc(act & ((1u << lane) - 1))] = (px_scratch && !handed) ? (local | PX_FALLBACK_BIT) : local;
    }
}

// ---- lane-parallel first pass -------------------------------------------------------------------------------------------
// k_render_pass1 is bounded by the serial instruction stream of the warps that cover the object (one thread marches one pixel:
// ~250 occupied steps x ~80 dependent instructions for the longest).  The same work split in two:
//   k_render_probe        thread per pixel, empty-space skipping only: finds the `t` in front of the first run of steps that may
//                         touch a leaf.  ~93 % of the pixels of the bench frame end here (no samples, background colour).
//   k_render_march_lanes  LANES = 8 lanes per hit pixel (4 pixels per warp): the steps of a run are evaluated 8 at a time, one
//                         per lane (position, active test, 8 corner densities, both alphas), then the ordered bookkeeping —
//                         transmittance, thresholds, early stop, the simulated second march — runs identically in all 8 lanes
//                         on values exchanged by shuffles.  Values of steps behind an early stop are computed and ignored.
// The order (values of `lanes` steps first, bookkeeping second) is bit-identical to the reference's step-by-step marches:
// proven on the CPU by oracle `orc_march_check(lanes = 8)` (tests/test_fast_march_cpu.py) and on the GPU against
// k_render_pass1 / px_entries = 0 by tests/test_renderer_gpu.py.  `active_list` then holds every HIT pixel; the march rewrites
// each entry with PX_FALLBACK_BIT (march again) or PX_EMPTY_BIT (no samples after all).
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


// Hand-over of pass 1's samples: the pixel's (t, weight) slot becomes its segment of the sample list — position as step_active
// computes it, in the first three floats of the feature row; the pixels pass 1 could not hand over go to the fallback list,
// which pass 2 marches like the reference does.  A warp takes 32 listed pixels: every lane sets up one pixel's ray, then the
// warp walks the 32 pixels together, lane r copying sample r (and r + 32): coalesced slot reads and list writes instead of
// one thread looping over its pixel's samples.
__global__ void __launch_bounds__(256) k_render_emit(RenderConst C, const float* __restrict__ c2w, int row_begin,
                                                     const int32_t* __restrict__ n_samples, const int32_t* __restrict__ i_starts,
                                                     const int32_t* __restrict__ active_list, const float2* __restrict__ px_scratch, int P,
                                                     int32_t* __restrict__ s_ray, float* __restrict__ s_weight, float* __restrict__ s_feat,
                                                     int64_t cap, int32_t* __restrict__ fallback_list, int32_t* __restrict__ counters) {
    pvdb_pdl_wait();
    const int n_active = counters[RC_ACTIVE];
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int base = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 32; base < n_active; base += warps * 32) {
        // lane j: metadata and ray of listed pixel base + j
        const int slot = base + lane;
        int e = 0, ns = 0;
        int64_t i0 = 0;
        Ray R;
#pragma unroll
        for (int a = 0; a < 3; ++a) { R.ro[a] = 0.f; R.rd[a] = 0.f; }
        if (slot < n_active) {
            e = active_list[slot];
            const int local = e & ~(PX_FALLBACK_BIT | PX_EMPTY_BIT);
            if (e & PX_EMPTY_BIT) {
                // a hit pixel of the lane-parallel march that has no samples after all: nothing to emit
            } else if (e & PX_FALLBACK_BIT) {
                fallback_list[atomicAdd(counters + RC_FALLBACK, 1)] = local;
            } else {
                ns = n_samples[local];
                i0 = i_starts[local];
                ray_setup(C, c2w, render_gpix(C, row_begin, local), R);
            }
        }
        const int cnt = min(32, n_active - base);
#pragma unroll 4
        for (int j = 0; j < cnt; ++j) {
            const int nsj = __shfl_sync(0xffffffffu, ns, j);
            if (nsj == 0) continue;                 // marched again by pass 2 (or past the end of the list)
            const int local = __shfl_sync(0xffffffffu, e, j) & ~(PX_FALLBACK_BIT | PX_EMPTY_BIT);
            const int64_t i0j = __shfl_sync(0xffffffffu, i0, j);
            float ro[3], rd[3];
#pragma unroll
            for (int a = 0; a < 3; ++a) { ro[a] = __shfl_sync(0xffffffffu, R.ro[a], j); rd[a] = __shfl_sync(0xffffffffu, R.rd[a], j); }
            const float2* src = px_scratch + (size_t)local * P;
            for (int r = lane; r < nsj; r += 32) {
                if (i0j + r >= cap) break;
                const float2 v = src[r];
                float* dst = s_feat + (i0j + r) * 12;      // position now, colour features after k_render_gather
#pragma unroll
                for (int a = 0; a < 3; ++a) dst[a] = __fmul_rn(__fmaf_rn(rd[a], v.x, ro[a]), C.wld[a]);
                s_weight[i0j + r] = v.y;
                s_ray[i0j + r] = local;
            }
        }
    }
}

// ---- exclusive scan over npix ints: 4096 items per CTA, then the block sums, then the offsets
__global__ void __launch_bounds__(1024) k_scan_blocks(const int32_t* __restrict__ in, int32_t* __restrict__ out, int n,
                                                      int32_t* __restrict__ block_sums) {
    pvdb_pdl_wait();
    __shared__ int wsum[32];
    const int base = blockIdx.x * 4096 + threadIdx.x * 4;
    int v[4], s = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) { v[i] = base + i < n ? in[base + i] : 0; s += v[i]; }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int incl = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += u; }
    if (lane == 31) wsum[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        int w = wsum[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += u; }
        wsum[lane] = w;
    }
    __syncthreads();
    int excl = incl - s + (wid ? wsum[wid - 1] : 0);
#pragma unroll
    for (int i = 0; i < 4; ++i) { if (base + i < n) out[base + i] = excl; excl += v[i]; }
    if (threadIdx.x == 1023) block_sums[blockIdx.x] = excl;
}
__global__ void __launch_bounds__(1024) k_scan_tops(int32_t* __restrict__ block_sums, int nb) {
    pvdb_pdl_wait();
    __shared__ int wsum[32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int own = threadIdx.x < nb ? block_sums[threadIdx.x] : 0;
    int incl = own;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += u; }
    if (lane == 31) wsum[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        int w = wsum[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += u; }
        wsum[lane] = w;
    }
    __syncthreads();
    const int excl = incl - own + (wid ? wsum[wid - 1] : 0);
    if (threadIdx.x < nb) block_sums[threadIdx.x] = excl;
    if (threadIdx.x == nb - 1) block_sums[nb] = excl + own;   // grand total
}
__global__ void __launch_bounds__(1024) k_scan_add(int32_t* __restrict__ out, int n, const int32_t* __restrict__ block_sums, int nb,
                                                   int32_t* __restrict__ counters, int64_t cap) {
    pvdb_pdl_wait();
    const int base = blockIdx.x * 4096 + threadIdx.x * 4;
    const int off = block_sums[blockIdx.x];
#pragma unroll
    for (int i = 0; i < 4; ++i)
        if (base + i < n) out[base + i] += off;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        const int total = block_sums[nb];
        out[n] = total;
        counters[RC_TOTAL] = total;
        counters[RC_OVERFLOW] = total > cap ? 1 : 0;
        counters[RC_INCONSISTENT] = 0;
    }
}

// Pass 2 (renderer.cu:312-367) is split in two.  The march itself stays one thread per listed pixel, but it only FINDS the kept
// samples (density, alpha, the sequential transmittance): per sample it stores weight, pixel and the index-space position in
// the first three floats of the sample's feature row.  The 8 x 12-channel colour gather (trigetColor, :303-310) — the bulk of
// the dependent loads, serial per pixel in the reference and the floor of a row band's latency — runs afterwards with one
// thread per SAMPLE (k_render_gather), same arithmetic in the same order.
__global__ void __launch_bounds__(256) k_render_pass2(RenderConst C, const float* __restrict__ c2w, int row_begin, int rows,
                                                      const int32_t* __restrict__ n_samples, const int32_t* __restrict__ i_starts,
                                                      const float* __restrict__ tmins, const float* __restrict__ tmaxs,
                                                      int32_t* __restrict__ s_ray, float* __restrict__ s_weight,
                                                      float* __restrict__ s_feat, int64_t cap, const int32_t* __restrict__ active_list,
                                                      float* __restrict__ out_rgb, int32_t* __restrict__ counters, int count_slot) {
    pvdb_pdl_wait();
    // one thread per listed pixel (every pixel with samples, or just those pass 1 could not hand over): full warps instead of
    // the ~15 live lanes of a pixel tile
    const int n_active = counters[count_slot];
  for (int slot = blockIdx.x * blockDim.x + threadIdx.x; slot < n_active; slot += gridDim.x * blockDim.x) {
    const int local = active_list[slot];
    const int ns = n_samples[local];
    const int n = render_gpix(C, row_begin, local);
    Ray R;
    ray_setup(C, c2w, n, R);
    MarchState S;
    PvdbLeafCache vcache;
    float T_cum = 1.0f, t = tmins[local];
    const float tmax = tmaxs[local];
    const int64_t i0 = i_starts[local];
    int r = 0;
    while (t < tmax) {
        int nslow = 0x7fffffff;
        if (C.skip_bits) {
            float t1, tk;
            nslow = run_chain(t, R.steplen, tmax, t1, tk);
            if (run_is_empty(C, R, t1, tk)) { t = tk; continue; }
        }
        for (int s = 0; s < nslow && t < tmax; ++s) {
            t = __fadd_rn(t, R.steplen);
            float xyz[3];
            int leaf;
            if (!step_active(C, R, S, t, xyz, leaf)) continue;
            // trigetDensity2 (:271-300)
            const int i = (int)xyz[0], j = (int)xyz[1], k = (int)xyz[2];
            const float u = __fsub_rn(xyz[0], (float)i), v = __fsub_rn(xyz[1], (float)j), w = __fsub_rn(xyz[2], (float)k);
            float den[8], sc[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int dx = PVDB_CORNER[q][0], dy = PVDB_CORNER[q][1], dz = PVDB_CORNER[q][2];
                den[q] = __ldg(C.dendata + idx_at(C, vcache, i + dx, j + dy, k + dz));
                const float f0 = dx ? u : __fsub_rn(1.f, u), f1 = dy ? v : __fsub_rn(1.f, v), f2 = dz ? w : __fsub_rn(1.f, w);
                sc[q] = __fmul_rn(__fmul_rn(f0, f1), f2);
            }
            // d0*s0 + d1*s1 + ... -> fma(d7,s7, ... fma(d2,s2, fma(d0,s0, d1*s1)))
            float vden = __fmaf_rn(den[0], sc[0], __fmul_rn(den[1], sc[1]));
#pragma unroll
            for (int q = 2; q < 8; ++q) vden = __fmaf_rn(den[q], sc[q], vden);
            const float alpha = render_alpha(vden, C.act_shift, C.interval);
            if (alpha <= C.thres) continue;
            const float weight = __fmul_rn(T_cum, alpha);
            T_cum = __fmul_rn(T_cum, __fsub_rn(1.f, alpha));
            if (weight <= C.thres) continue;
            if (r < ns && i0 + r < cap) {
                float* dst = s_feat + (i0 + r) * 12;      // position now, colour features after k_render_gather
                dst[0] = xyz[0]; dst[1] = xyz[1]; dst[2] = xyz[2];
                s_weight[i0 + r] = weight;
                s_ray[i0 + r] = local;
            }
            ++r;
        }
    }
    if (r != ns) {
        atomicAdd(counters + RC_INCONSISTENT, 1);
        // keep the tail of the segment well defined: zero weight contributes nothing
        for (int q = r; q < ns; ++q)
            if (i0 + q < cap) {
                s_weight[i0 + q] = 0.f; s_ray[i0 + q] = local;
                float* dst = s_feat + (i0 + q) * 12;
                dst[0] = dst[1] = dst[2] = 0.f;
            }
    }
    const float last = __fmul_rn(T_cum, C.bg);   // :364-365
    out_rgb[local * 3] = last; out_rgb[local * 3 + 1] = last; out_rgb[local * 3 + 2] = last;
  }
}

// trigetColor (:303-310) of every kept sample: one thread per sample, 8 index lookups, then 24 independent 16-byte loads and
// the reference's accumulation order per channel: c0*s0 + c1*s1 -> fma(c7,s7, ... fma(c2,s2, fma(c0,s0, c1*s1))).
__global__ void __launch_bounds__(256) k_render_gather(RenderConst C, float* __restrict__ s_feat, const int32_t* __restrict__ counters,
                                                       int64_t cap) {
    pvdb_pdl_wait();
    const int64_t total = min((int64_t)counters[RC_TOTAL], cap);
    for (int64_t sidx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; sidx < total; sidx += (int64_t)gridDim.x * blockDim.x) {
        float* row = s_feat + sidx * 12;
        const float x = row[0], y = row[1], z = row[2];
        const int i = (int)x, j = (int)y, k = (int)z;
        const float u = __fsub_rn(x, (float)i), v = __fsub_rn(y, (float)j), w = __fsub_rn(z, (float)k);
        PvdbLeafCache vcache;
        int idx[8];
        float sc[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int dx = PVDB_CORNER[q][0], dy = PVDB_CORNER[q][1], dz = PVDB_CORNER[q][2];
            idx[q] = idx_at(C, vcache, i + dx, j + dy, k + dz);
            const float f0 = dx ? u : __fsub_rn(1.f, u), f1 = dy ? v : __fsub_rn(1.f, v), f2 = dz ? w : __fsub_rn(1.f, w);
            sc[q] = __fmul_rn(__fmul_rn(f0, f1), f2);
        }
        float4 f[3];
#pragma unroll
        for (int c4 = 0; c4 < 3; ++c4) {
            const float4 a0 = __ldg(reinterpret_cast<const float4*>(C.coldata + (size_t)idx[0] * 12) + c4);
            const float4 a1 = __ldg(reinterpret_cast<const float4*>(C.coldata + (size_t)idx[1] * 12) + c4);
            f[c4].x = __fmaf_rn(a0.x, sc[0], __fmul_rn(a1.x, sc[1])); f[c4].y = __fmaf_rn(a0.y, sc[0], __fmul_rn(a1.y, sc[1]));
            f[c4].z = __fmaf_rn(a0.z, sc[0], __fmul_rn(a1.z, sc[1])); f[c4].w = __fmaf_rn(a0.w, sc[0], __fmul_rn(a1.w, sc[1]));
        }
#pragma unroll
        for (int q = 2; q < 8; ++q)
#pragma unroll
            for (int c4 = 0; c4 < 3; ++c4) {
                const float4 a = __ldg(reinterpret_cast<const float4*>(C.coldata + (size_t)idx[q] * 12) + c4);
                f[c4].x = __fmaf_rn(a.x, sc[q], f[c4].x); f[c4].y = __fmaf_rn(a.y, sc[q], f[c4].y);
                f[c4].z = __fmaf_rn(a.z, sc[q], f[c4].z); f[c4].w = __fmaf_rn(a.w, sc[q], f[c4].w);
            }
        float4* dst = reinterpret_cast<float4*>(row);
        dst[0] = f[0]; dst[1] = f[1]; dst[2] = f[2];
    }
}

// Dilated block map of the index tree for run_is_empty: bit b is set when block b or one of its +1 neighbours holds a leaf.
__global__ void __launch_bounds__(256) k_block_bits(pvdb_tree t, int nbx, int nby, int nbz, uint32_t* __restrict__ bits) {
    const int leaf = blockIdx.x * blockDim.x + threadIdx.x;
    if (leaf >= t.n_leaf) return;
    const int bx = t.leaf_origin[leaf * 3] >> 3, by = t.leaf_origin[leaf * 3 + 1] >> 3, bz = t.leaf_origin[leaf * 3 + 2] >> 3;
    for (int d = 0; d < 8; ++d) {
        const int x = bx - (d & 1), y = by - ((d >> 1) & 1), z = bz - (d >> 2);
        if (x < 0 || y < 0 || z < 0 || x >= nbx || y >= nby || z >= nbz) continue;
        const int bit = (x * nby + y) * nbz + z;
        atomicOr(bits + (bit >> 5), 1u << (bit & 31));
    }
}

__global__ void __launch_bounds__(NT, 1) k_render_mlp(RenderMlpArgs A) {
    extern __shared__ __align__(16) float smem[];
    float* sW0t = smem;                    // [KX][128]  (w0 is already k-major: w0[i][j], run.py:98)
    float* sW1t = sW0t + KX * W;           // [128][128]
    float* sW2 = sW1t + W * W;             // [3][128]   sW2[j][i] = w2[i][j]
    float* sb0 = sW2 + 3 * W;
    float* sb1 = sb0 + W;
    float* sb2 = sb1 + W;
    float* sX = sb2 + 4;
    float* sH0 = sX + TS * LDX;
    float* sH1 = sH0 + TS * LDH;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    for (int e = tid; e < KX * W; e += NT) sW0t[e] = e < DIN * W ? __ldg(A.w0 + e) : 0.f;
    for (int e = tid; e < W * W; e += NT) sW1t[e] = __ldg(A.w1 + e);
    for (int e = tid; e < 3 * W; e += NT) { const int j = e / W, i = e % W; sW2[e] = __ldg(A.w2 + i * 3 + j); }
    if (tid < W) { sb0[tid] = __ldg(A.b0 + tid); sb1[tid] = __ldg(A.b1 + tid); }
    if (tid < 3) sb2[tid] = __ldg(A.b2 + tid);
    const int64_t M = min((int64_t)A.counters[RC_TOTAL], A.cap);
    const int64_t n_tiles = (M + TS - 1) / TS;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t s0 = tile * TS;
        __syncthreads();
        if (tid < 192) {
            const int s = tid / 3, c4 = tid % 3;
            float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
            if (s0 + s < M) f = __ldg(reinterpret_cast<const float4*>(A.s_feat + (s0 + s) * 12) + c4);
            *reinterpret_cast<float4*>(sX + s * LDX + c4 * 4) = f;
        } else {
            const int s = tid - 192;
            float pe[27];
#pragma unroll
            for (int i = 0; i < 27; ++i) pe[i] = 0.f;
            if (s0 + s < M) {
                Ray R;
                ray_setup(A.C, A.c2w, render_gpix(A.C, A.row_begin, A.s_ray[s0 + s]), R);
                pe[0] = R.vd[0]; pe[1] = R.vd[1]; pe[2] = R.vd[2];
                // pefeat (:153-166): sin/cos(viewdir * pebase), pebase = 1,2,4,8 as int -> float
#pragma unroll
                for (int k = 0; k < 4; ++k)
#pragma unroll
                    for (int a = 0; a < 3; ++a) {
                        const float x = __fmul_rn(R.vd[a], (float)(1 << k));
                        pe[3 + k + 4 * a] = sinf(x);
                        pe[15 + k + 4 * a] = cosf(x);
                    }
            }
#pragma unroll
            for (int i = 0; i < 27; ++i) sX[s * LDX + 12 + i] = pe[i];
            sX[s * LDX + 39] = 0.f;
        }
        __syncthreads();
        float acc[4][8];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int c = 0; c < 8; ++c) acc[i][c] = sb0[(c < 4 ? 0 : 64) + tx * 4 + (c & 3)];
        gemm_4x8<KX, LDX, W>(sX, sW0t, ty, tx, acc);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            *reinterpret_cast<float4*>(sH0 + (ty * 4 + i) * LDH + tx * 4) =
                make_float4(fmaxf(acc[i][0], 0.f), fmaxf(acc[i][1], 0.f), fmaxf(acc[i][2], 0.f), fmaxf(acc[i][3], 0.f));
            *reinterpret_cast<float4*>(sH0 + (ty * 4 + i) * LDH + 64 + tx * 4) =
                make_float4(fmaxf(acc[i][4], 0.f), fmaxf(acc[i][5], 0.f), fmaxf(acc[i][6], 0.f), fmaxf(acc[i][7], 0.f));
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int c = 0; c < 8; ++c) acc[i][c] = sb1[(c < 4 ? 0 : 64) + tx * 4 + (c & 3)];
        gemm_4x8<W, LDH, W>(sH0, sW1t, ty, tx, acc);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            *reinterpret_cast<float4*>(sH1 + (ty * 4 + i) * LDH + tx * 4) =
                make_float4(fmaxf(acc[i][0], 0.f), fmaxf(acc[i][1], 0.f), fmaxf(acc[i][2], 0.f), fmaxf(acc[i][3], 0.f));
            *reinterpret_cast<float4*>(sH1 + (ty * 4 + i) * LDH + 64 + tx * 4) =
                make_float4(fmaxf(acc[i][4], 0.f), fmaxf(acc[i][5], 0.f), fmaxf(acc[i][6], 0.f), fmaxf(acc[i][7], 0.f));
        }
        __syncthreads();
        if (tid < 192) {
            const int s = tid / 3, j = tid % 3;
            float a = sb2[j];
#pragma unroll 8
            for (int i = 0; i < W; ++i) a = fmaf(sH1[s * LDH + i], sW2[j * W + i], a);
            // final_render (:115-117): weight / (1 + exp(-raw))
            if (s0 + s < M) A.s_rgb[(s0 + s) * 3 + j] = A.s_weight[s0 + s] / (1 + expf(-a));
        }
    }
}

// Local pixel -> image pixel of the call's row layout (see RenderConst::band_rows).
struct RowMap { int W, row_begin, band_rows, band_stride; };

__global__ void __launch_bounds__(256) k_render_composite(const int32_t* __restrict__ n_samples, const int32_t* __restrict__ i_starts,
                                                          const float* __restrict__ s_rgb, int npix, int64_t cap,
                                                          float* __restrict__ out_rgb, int32_t* __restrict__ counters, RowMap M,
                                                          float* __restrict__ frame_out) {
    pvdb_pdl_wait();
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p == 0) counters[RC_ACTIVE] = 0;      // this frame is done with it; pass 1 of the next frame counts from 0
    if (p >= npix) return;
    const int ns = n_samples[p];
    if (ns == 0 && !frame_out) return;
    float r = out_rgb[p * 3], g = out_rgb[p * 3 + 1], b = out_rgb[p * 3 + 2];
    if (ns) {
        const int64_t i0 = i_starts[p];
        for (int64_t s = i0; s < i0 + ns && s < cap; ++s) {
            r = __fadd_rn(r, s_rgb[s * 3]); g = __fadd_rn(g, s_rgb[s * 3 + 1]); b = __fadd_rn(b, s_rgb[s * 3 + 2]);
        }
        out_rgb[p * 3] = r; out_rgb[p * 3 + 1] = g; out_rgb[p * 3 + 2] = b;
    }
    if (frame_out) {
        // Every pixel of the band goes to its place in the full frame.  frame_out may be another GPU's memory (NVLink peer
        // mapping): plain stores, a warp covers 384 contiguous bytes; the frame-done signal that follows publishes them.
        int gp = M.row_begin * M.W + p;
        if (M.band_stride) {
            const int lr = p / M.W, col = p - lr * M.W, k = lr / M.band_rows;
            gp = (M.row_begin + k * M.band_stride + (lr - k * M.band_rows)) * M.W + col;
        }
        float* dst = frame_out + (size_t)gp * 3;
        dst[0] = r; dst[1] = g; dst[2] = b;
    }
}

// ---- frame assembly over NVLink peer memory (no counterpart in the reference, which is single-GPU) ---------------------------
// One symmetric block per rank (pvdb_dp_symm_alloc + CUDA IPC):  [0,32) done[8] u32 | [32,36) free u32 | [36,40) err i32 |
// [1024,..) frame[2][H*W*3] f32.  Frame f is assembled in ROOT's frame[f & 1]: every rank's composite kernel stores its rows
// straight into it, then signals done[rank] = f + 1 in root's block; root waits for all of them.  Before it starts frame f, root
// writes free = f + 1 into every peer's block (frame f - 2, the previous user of that buffer, has been consumed in stream order);
// a peer waits for it only in front of its composite kernel, so its march overlaps root's wait.
struct FrameBlk { uint32_t* done; uint32_t* free_; int32_t* err; float* frame[2]; };
__host__ __device__ inline FrameBlk frame_view(void* base, int H, int W) {
    char* p = static_cast<char*>(base);
    FrameBlk b;
    b.done = reinterpret_cast<uint32_t*>(p);
    b.free_ = reinterpret_cast<uint32_t*>(p + 32);
    b.err = reinterpret_cast<int32_t*>(p + 36);
    b.frame[0] = reinterpret_cast<float*>(p + 1024);
    b.frame[1] = b.frame[0] + (size_t)H * W * 3;
    return b;
}
// root, first kernel of frame f: buffer f & 1 may be written
__global__ void k_frame_release(pvdb_frame_peers P, uint32_t epoch) {
    const int r = threadIdx.x;
    if (r < P.world && r != P.root) st_release_sys(frame_view(P.base[r], P.H, P.W).free_, epoch);
}
// peer, in front of its composite kernel
__global__ void k_frame_wait_free(pvdb_frame_peers P, uint32_t epoch) {
    const FrameBlk me = frame_view(P.base[P.rank], P.H, P.W);
    wait_epoch(me.free_, epoch, me.err, 1);
}
// after the composite kernel: a peer publishes its rows (the kernel boundary orders the composite's stores before this thread,
// the release is cumulative over them), root collects every peer's signal
__global__ void k_frame_done(pvdb_frame_peers P, uint32_t epoch) {
    if (P.rank != P.root) {
        if (threadIdx.x == 0) {
            __threadfence_system();
            st_release_sys(frame_view(P.base[P.root], P.H, P.W).done + P.rank, epoch);
        }
        return;
    }
    const int r = threadIdx.x;
    const FrameBlk me = frame_view(P.base[P.rank], P.H, P.W);
    if (r < P.world && r != P.root) wait_epoch(me.done + r, epoch, me.err, 1);
}

constexpr size_t RENDER_MLP_SMEM = (size_t)(KX * W + W * W + 3 * W + W + W + 4 + TS * LDX + 2 * TS * LDH) * sizeof(float);

// vdb_compression.py:36-48: dendata[row], coldata[row][:] of every masked voxel, rounded through fp16 (:56-57)
__global__ void __launch_bounds__(256) k_merge_gather(pvdb_tree t, const float* __restrict__ den, const float* __restrict__ k0, int cdim,
                                                      const int32_t* __restrict__ row_of_voxel, int rx, int ry, int rz,
                                                      float* __restrict__ dendata, float* __restrict__ coldata) {
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= (int64_t)rx * ry * rz) return;
    const int row = row_of_voxel[v];
    if (row <= 0) return;
    const int z = (int)(v % rz), y = (int)((v / rz) % ry), x = (int)(v / ((int64_t)rz * ry));
    const int leaf = pvdb_find_leaf(t, x, y, z);
    const int off = pvdb_leaf_off(x, y, z);
    dendata[row] = __half2float(__float2half_rn(leaf >= 0 ? den[(size_t)leaf * 512 + off] : 0.f));
    for (int c = 0; c < cdim; ++c)
        coldata[(size_t)row * cdim + c] = __half2float(__float2half_rn(leaf >= 0 ? k0[((size_t)leaf * 512 + off) * cdim + c] : 0.f));
}

}  // namespace

int pvdb_render_mlp_tc(const void* render_mlp_args, cudaStream_t st);   // rgbnet_tc.cu

// Rows of an H-row image that fall to `rank` when groups of band_rows rows are dealt round-robin to `world` ranks.
extern "C" int pvdb_interleaved_rows(int H, int band_rows, int rank, int world) {
    if (H <= 0 || band_rows <= 0 || world <= 0 || rank < 0 || rank >= world) return 0;
    int rows = 0;
    for (int r0 = rank * band_rows; r0 < H; r0 += world * band_rows) rows += min(band_rows, H - r0);
    return rows;
}

// The frame as a fixed sequence of kernels on one stream.  Local row lr of the band buffer out_rgb is image row
// row_begin + lr (band_stride == 0) or row_begin + (lr / band_rows) * band_stride + lr % band_rows.  frame_out (optional):
// the full H x W x 3 frame, possibly peer memory, that also receives every pixel of the band.  fp (optional): the peer
// protocol of pvdb_render_frame_sharded around the composite kernel.
// First pass of the renderer: 1 probe + lane-parallel march, 0 k_render_pass1 (thread per pixel), -1 chosen per call by the share
// of the frame (see render_impl), -2 not read from the environment yet.  Bit-identical results either way.
static int g_render_lanes = -2;
extern "C" void pvdb_debug_set_render_lanes(int on) { g_render_lanes = on < 0 ? -1 : (on ? 1 : 0); }

static int render_impl(const pvdb_render_cfg* cfg, const pvdb_render_bufs* b, const float* c2w, int row_begin, int rows, int band_rows,
                       int band_stride, float* out_rgb, float* frame_out, const pvdb_frame_peers* fp, uint32_t epoch, void* stream) {
    PVDB_CHECK_ARG(cfg && b && b->idx_tree && c2w && out_rgb, "null pointer");
    PVDB_CHECK_ARG(cfg->dcol == 12 && cfg->dpe == 27 && cfg->dhid == 128 && cfg->dout == 3,
                   "the merged renderer is specialised for MGRenderer(12, 27, 128, 3) (run.py:77-82)");
    cudaStream_t st = (cudaStream_t)stream;
    RenderConst C;
    C.tree = *b->idx_tree; C.idx_plane = b->idx_plane; C.dendata = b->dendata; C.coldata = b->coldata;
    for (int i = 0; i < 9; ++i) C.K[i] = cfg->K[i];
    for (int a = 0; a < 3; ++a) {
        C.xyz_min[a] = cfg->xyz_min[a];
        C.ext[a] = cfg->xyz_max[a] - cfg->xyz_min[a];
        C.wld[a] = (float)(cfg->reso[a] - 1);
    }
    C.near = cfg->near; C.stepdist = cfg->stepdist; C.act_shift = cfg->act_shift; C.interval = cfg->interval;
    C.thres = cfg->fast_color_thres; C.bg = cfg->bg; C.inverse_y = cfg->inverse_y; C.H = cfg->H; C.W = cfg->W;
    C.band_rows = band_stride ? band_rows : rows; C.band_stride = band_stride;
    C.skip_bits = b->skip_bits;
    for (int a = 0; a < 3; ++a) C.nb[a] = (cfg->reso[a] + 7) / 8;
    const int npix = rows * cfg->W;
    const int tiles = ((cfg->W + 7) / 8) * ((rows + 3) / 4);
    const int pgrid = pvdb_grid_for((int64_t)tiles * 32, 256);
    PVDB_CHECK_ARG(b->active_list, "active_list scratch missing");
    const bool hand_over = b->px_scratch && b->fallback_list && b->px_entries > 0;
    float2* px = hand_over ? static_cast<float2*>(b->px_scratch) : nullptr;
    // First pass: measured on B200 on the 800x800 bench frame (profiles/render_share_timing_r02.json, us for a full frame and for
    // a 1/2, 1/4, 1/8 share of its rows): thread per pixel 243 / 201 / 169 / 165 — bounded below by its longest warps —, probe +
    // 8-lane march 305 / 166 / 102 / 76 — more instructions (replicated bookkeeping, shuffles), but they scale with the share.
    // So the whole frame on one GPU takes the thread-per-pixel pass and a row share (tile-sharded rendering) the lane-parallel
    // one.  g_render_lanes: -1 that rule, 0 / 1 force one of them (environment PVDB_RENDER_LANES, pvdb_debug_set_render_lanes).
    if (g_render_lanes == -2) { const char* e = getenv("PVDB_RENDER_LANES"); g_render_lanes = e ? (atoi(e) != 0) : -1; }
    const bool lanes = g_render_lanes < 0 ? rows * 2 <= cfg->H : g_render_lanes != 0;
    if (lanes && hand_over && C.skip_bits) {
        PVDB_CUDA(pvdb_launch_pdl(k_render_probe, dim3(pgrid), dim3(256), 0, st, C, c2w, row_begin, rows, b->n_samples, b->tmins, b->tmaxs, b->active_list,
                                  b->counters, out_rgb));
        PVDB_LAUNCH_CHECK();
        pvdb_prof_mark("render_probe", st);
        PVDB_CUDA(pvdb_launch_pdl(k_render_march_lanes, dim3(PVDB_SMS * 8), dim3(256), 0, st, C, c2w, row_begin, b->n_samples, b->tmins, b->tmaxs,
                                  b->active_list, (const int32_t*)b->counters, out_rgb, px, (int)b->px_entries));
        PVDB_LAUNCH_CHECK();
        pvdb_prof_mark("render_march_lanes", st);
    } else {
        PVDB_CUDA(pvdb_launch_pdl(k_render_pass1, dim3(pgrid), dim3(256), 0, st, C, c2w, row_begin, rows, b->n_samples, b->tmins, b->tmaxs, b->active_list,
                                  b->counters, out_rgb, px, (int)b->px_entries));
        PVDB_LAUNCH_CHECK();
        pvdb_prof_mark("render_pass1", st);
    }
    const int nb = (npix + 4095) / 4096;
    PVDB_CHECK_ARG(nb <= 1023, "row band too large for the scan (max 4M pixels per call)");
    PVDB_CUDA(pvdb_launch_pdl(k_scan_blocks, dim3(nb), dim3(1024), 0, st, (const int32_t*)b->n_samples, b->i_starts, npix, b->scan_tmp));
    PVDB_LAUNCH_CHECK();
    PVDB_CUDA(pvdb_launch_pdl(k_scan_tops, dim3(1), dim3(1024), 0, st, b->scan_tmp, nb));
    PVDB_LAUNCH_CHECK();
    PVDB_CUDA(pvdb_launch_pdl(k_scan_add, dim3(nb), dim3(1024), 0, st, b->i_starts, npix, (const int32_t*)b->scan_tmp, nb, b->counters, b->cap_samples));
    PVDB_LAUNCH_CHECK();
    pvdb_prof_mark("render_scan", st);
    if (hand_over) {
        PVDB_CUDA(pvdb_launch_pdl(k_render_emit, dim3(min(pgrid, PVDB_SMS * 8)), dim3(256), 0, st, C, c2w, row_begin, (const int32_t*)b->n_samples,
                                  (const int32_t*)b->i_starts, (const int32_t*)b->active_list, (const float2*)px, (int)b->px_entries, b->s_ray, b->s_weight, b->s_feat, b->cap_samples,
                                  b->fallback_list, b->counters));
        PVDB_LAUNCH_CHECK();
        pvdb_prof_mark("render_emit", st);
    }
    PVDB_CUDA(pvdb_launch_pdl(k_render_pass2, dim3(hand_over ? PVDB_SMS : min(pgrid, PVDB_SMS * 8)), dim3(256), 0, st, C, c2w, row_begin, rows,
                              (const int32_t*)b->n_samples, (const int32_t*)b->i_starts, (const float*)b->tmins, (const float*)b->tmaxs, b->s_ray, b->s_weight,
                              b->s_feat, b->cap_samples, (const int32_t*)(hand_over ? b->fallback_list : b->active_list), out_rgb, b->counters,
                              hand_over ? RC_FALLBACK : RC_ACTIVE));
    PVDB_LAUNCH_CHECK();
    pvdb_prof_mark("render_pass2", st);
    PVDB_CUDA(pvdb_launch_pdl(k_render_gather, dim3(PVDB_SMS * 8), dim3(256), 0, st, C, b->s_feat, (const int32_t*)b->counters, b->cap_samples));
    PVDB_LAUNCH_CHECK();
    pvdb_prof_mark("render_gather", st);
    static bool attr_set = false;
    if (!attr_set) {
        PVDB_CUDA(cudaFuncSetAttribute(k_render_mlp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RENDER_MLP_SMEM));
        attr_set = true;
    }
    RenderMlpArgs A;
    A.C = C; A.c2w = c2w; A.w0 = b->w0; A.b0 = b->b0; A.w1 = b->w1; A.b1 = b->b1; A.w2 = b->w2; A.b2 = b->b2;
    A.s_ray = b->s_ray; A.s_weight = b->s_weight; A.s_feat = b->s_feat; A.s_rgb = b->s_rgb; A.counters = b->counters;
    A.cap = b->cap_samples; A.row_begin = row_begin; A.img = static_cast<unsigned char*>(b->w_img);
    if (cfg->use_tensor_cores) {
        int rc = pvdb_render_mlp_tc(&A, st);
        if (rc) return rc;
    } else {
        k_render_mlp<<<PVDB_SMS, NT, RENDER_MLP_SMEM, st>>>(A);
        PVDB_LAUNCH_CHECK();
    }
    pvdb_prof_mark("render_mlp", st);
    if (fp && fp->rank != fp->root) {
        k_frame_wait_free<<<1, 32, 0, st>>>(*fp, epoch);
        PVDB_LAUNCH_CHECK();
    }
    RowMap M;
    M.W = cfg->W; M.row_begin = row_begin; M.band_rows = C.band_rows; M.band_stride = band_stride;
    PVDB_CUDA(pvdb_launch_pdl(k_render_composite, dim3(pvdb_grid_for(npix, 256)), dim3(256), 0, st, (const int32_t*)b->n_samples, (const int32_t*)b->i_starts,
                              (const float*)b->s_rgb, npix, b->cap_samples, out_rgb, b->counters, M, frame_out));
    PVDB_LAUNCH_CHECK();
    pvdb_prof_mark("render_composite", st);
    return PVDB_OK;
}

extern "C" int pvdb_render_rows(const pvdb_render_cfg* cfg, const pvdb_render_bufs* b, const float* c2w, int row_begin, int row_end,
                                float* out_rgb, void* stream) {
    PVDB_CHECK_ARG(cfg, "null pointer");
    PVDB_CHECK_ARG(0 <= row_begin && row_begin < row_end && row_end <= cfg->H, "bad row range");
    pvdb_reset_launch_count();
    pvdb_prof_begin((cudaStream_t)stream);
    return render_impl(cfg, b, c2w, row_begin, row_end - row_begin, 0, 0, out_rgb, nullptr, nullptr, 0, stream);
}

extern "C" int pvdb_render_rows_interleaved(const pvdb_render_cfg* cfg, const pvdb_render_bufs* b, const float* c2w, int band_rows, int rank,
                                            int world, float* band_out, float* frame_out, void* stream) {
    PVDB_CHECK_ARG(cfg, "null pointer");
    PVDB_CHECK_ARG(band_rows > 0 && world > 0 && rank >= 0 && rank < world, "bad band_rows / rank / world");
    const int rows = pvdb_interleaved_rows(cfg->H, band_rows, rank, world);
    PVDB_CHECK_ARG(rows > 0, "this rank has no rows (H < rank * band_rows)");
    pvdb_reset_launch_count();
    pvdb_prof_begin((cudaStream_t)stream);
    return render_impl(cfg, b, c2w, rank * band_rows, rows, band_rows, world * band_rows, band_out, frame_out, nullptr, 0, stream);
}

extern "C" size_t pvdb_frame_symm_bytes(int H, int W) {
    if (H <= 0 || W <= 0) return 0;
    return 1024 + 2 * (size_t)H * W * 3 * sizeof(float);
}

static int check_frame_peers(const pvdb_frame_peers* P) {
    PVDB_CHECK_ARG(P, "null peers");
    PVDB_CHECK_ARG(P->world >= 1 && P->world <= 8 && P->rank >= 0 && P->rank < P->world && P->root >= 0 && P->root < P->world,
                   "world must be 1..8, rank and root inside it");
    PVDB_CHECK_ARG(P->H > 0 && P->W > 0, "bad frame size");
    for (int r = 0; r < P->world; ++r) PVDB_CHECK_ARG(P->base[r], "peer block not mapped");
    return PVDB_OK;
}

extern "C" int pvdb_render_frame_sharded(const pvdb_render_cfg* cfg, const pvdb_render_bufs* b, const pvdb_frame_peers* P, const float* c2w,
                                         int band_rows, uint32_t frame_no, float* band_out, void* stream) {
    if (int rc = check_frame_peers(P)) return rc;
    PVDB_CHECK_ARG(cfg && cfg->H == P->H && cfg->W == P->W, "the peer blocks were sized for another frame");
    PVDB_CHECK_ARG(band_rows > 0, "bad band_rows");
    cudaStream_t st = (cudaStream_t)stream;
    const uint32_t epoch = frame_no + 1;      // monotone; the signal words start at 0
    pvdb_reset_launch_count();
    pvdb_prof_begin(st);
    if (P->rank == P->root && P->world > 1) {
        k_frame_release<<<1, 32, 0, st>>>(*P, epoch);
        PVDB_LAUNCH_CHECK();
    }
    const int rows = pvdb_interleaved_rows(cfg->H, band_rows, P->rank, P->world);
    if (rows > 0) {
        float* frame = frame_view(P->base[P->root], P->H, P->W).frame[frame_no & 1];
        int rc = render_impl(cfg, b, c2w, P->rank * band_rows, rows, band_rows, P->world * band_rows, band_out, frame, P, epoch, stream);
        if (rc) return rc;
    }
    if (P->world > 1) {
        k_frame_done<<<1, 32, 0, st>>>(*P, epoch);
        PVDB_LAUNCH_CHECK();
    }
    return PVDB_OK;
}

extern "C" int pvdb_frame_ptr(const pvdb_frame_peers* P, uint32_t frame_no, float** frame) {
    if (int rc = check_frame_peers(P)) return rc;
    PVDB_CHECK_ARG(frame, "null pointer");
    *frame = frame_view(P->base[P->root], P->H, P->W).frame[frame_no & 1];
    return PVDB_OK;
}

extern "C" int pvdb_frame_copy(const pvdb_frame_peers* P, uint32_t frame_no, float* dst, void* stream) {
    if (int rc = check_frame_peers(P)) return rc;
    PVDB_CHECK_ARG(dst, "null pointer");
    PVDB_CUDA(cudaMemcpyAsync(dst, frame_view(P->base[P->root], P->H, P->W).frame[frame_no & 1], (size_t)P->H * P->W * 3 * sizeof(float),
                              cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return PVDB_OK;
}

extern "C" int pvdb_frame_error(const pvdb_frame_peers* P, int32_t* err_out) {
    if (int rc = check_frame_peers(P)) return rc;
    PVDB_CHECK_ARG(err_out, "null pointer");
    PVDB_CUDA(cudaMemcpy(err_out, frame_view(P->base[P->rank], P->H, P->W).err, sizeof(int32_t), cudaMemcpyDeviceToHost));
    return PVDB_OK;
}

extern "C" size_t pvdb_render_block_bits_words(int rx, int ry, int rz) {
    if (rx <= 0 || ry <= 0 || rz <= 0) return 0;
    return ((size_t)((rx + 7) / 8) * ((ry + 7) / 8) * ((rz + 7) / 8) + 31) / 32;
}

extern "C" int pvdb_render_block_bits(const pvdb_tree* tree, int rx, int ry, int rz, uint32_t* bits, void* stream) {
    PVDB_CHECK_ARG(tree && bits && rx > 0 && ry > 0 && rz > 0, "bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    PVDB_CUDA(cudaMemsetAsync(bits, 0, pvdb_render_block_bits_words(rx, ry, rz) * sizeof(uint32_t), st));
    if (tree->n_leaf > 0) {
        k_block_bits<<<pvdb_grid_for(tree->n_leaf, 256), 256, 0, st>>>(*tree, (rx + 7) / 8, (ry + 7) / 8, (rz + 7) / 8, bits);
        PVDB_LAUNCH_CHECK();
    }
    return PVDB_OK;
}

extern "C" int pvdb_merge_gather(const pvdb_tree* tree, const float* den, const float* k0, int k0_dim, const int32_t* row_of_voxel,
                                 int rx, int ry, int rz, float* dendata, float* coldata, void* stream) {
    PVDB_CHECK_ARG(tree && den && k0 && row_of_voxel && dendata && coldata && k0_dim > 0, "bad arguments");
    const int64_t n = (int64_t)rx * ry * rz;
    k_merge_gather<<<pvdb_grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(*tree, den, k0, k0_dim, row_of_voxel, rx, ry, rz, dendata,
                                                                             coldata);
    PVDB_LAUNCH_CHECK();
    return PVDB_OK;
}
