// train_step.cu — fused fine-stage training step on one stream, no host synchronisation.
//
// Restates DirectVoxGO.forward (plenvdb/lib/dvgo.py:296-388), the losses of plenvdb/run.py:551-574,
// autograd's backward through Alphas2Weights / Raw2Alpha / QueryVerticalInVDB (dvgo.py:408-466,
// plenvdb/lib/grid.py:40-60) and the optimiser calls of run.py:585-588 as a fixed kernel sequence:
//
//   F1 k_march<COUNT>   warp per ray: t-range, N_steps, march, occupancy bits, density trilinear, alpha,
//                       sequential transmittance with the reference's early stop  -> per-ray counts
//   F2 k_scan_counts    exclusive scans -> segment offsets (ray order, deterministic)
//   F3 k_march<EMIT>    same march, writes the compacted sample lists
//   F4 k_rgbnet_fwd     k0 trilinear (12 ch) + view PE + MLP + sigmoid, 64-sample tiles
//   F5 k_composite      per-ray compositing, losses, dL/d(rgb_marched), dL/d(alphainv_last), dL/d(rgb), dL/d(w)
//   B1 k_rgbnet_bwd     MLP backward, weight-gradient partials in registers, k0 gradient scatter
//   B2 k_ray_bwd        warp per ray: reverse cumprod backward (shuffle suffix scan) + raw2alpha backward
//   B3 k_density_scatter density gradient scatter
//   U1 k_update_fused   one launch: Adam over touched leaves only (gradients cleared in the same pass) + rgbnet Adam
//
// The per-sample arithmetic is shared with the drop-in ops through ray_math.cuh / common.cuh, so sample
// counts, segment offsets and voxel indices are the reference's bit for bit.
#include <stdlib.h>
#include "common.cuh"
#include "ray_math.cuh"
#include "rgbnet.cuh"
#include "dp_exchange.cuh"
#include "leaf_local.cuh"

namespace {

constexpr int CNT_M_ALPHA = 0, CNT_M_KEEP = 1, CNT_N_TOUCHED_DEN = 2, CNT_OVERFLOW = 3, CNT_N_TOUCHED_K0 = 4, CNT_RAY_TICKET = 5, CNT_CTA_DONE = 6, CNT_MARCH_DONE = 9;

struct MarchParams {
    pvdb_tree tree;
    const float* den;
    const uint64_t* occ_fine;
    const uint64_t* occ_coarse;
    float xyz_min[3], xyz_max[3];
    float rm1[3];                 // world_size - 1 as float
    int mask_reso[3];
    int nb[3];                    // blocks of the occupancy bit grid
    float mask_scale[3], mask_shift[3];
    float near, far, stepdist, act_shift, interval, thres;
    int run_skip;                 // pass A skips runs of steps that cannot hit the mask (bit-identical; 0 = test every step)
};

__device__ __forceinline__ bool occ_test(const MarchParams& P, int i, int j, int k) {
    if (i < 0 || i >= P.mask_reso[0] || j < 0 || j >= P.mask_reso[1] || k < 0 || k >= P.mask_reso[2]) return false;
    const int b = ((i >> 3) * P.nb[1] + (j >> 3)) * P.nb[2] + (k >> 3);
    if (!((__ldg(P.occ_coarse + (b >> 6)) >> (b & 63)) & 1ull)) return false;   // empty 8^3 block
    const int n = pvdb_leaf_off(i, j, k);
    return (__ldg(P.occ_fine + (size_t)b * 8 + (n >> 6)) >> (n & 63)) & 1ull;
}

// Trilinear density at index-space (x,y,z): densityvdb.cu:101-125 arithmetic.
// rec[q] (EMIT only) receives the record id leaf*512+voxel of corner q, -1 when the corner has no leaf: the k0 grid
// shares the topology, so the rgbnet kernels reuse them instead of walking the tree again.
template <bool WITH_REC>
__device__ __forceinline__ float density_at(const MarchParams& P, float x, float y, float z, int* rec) {
    PvdbTri tri;
    tri.set(x, y, z);
    // The eight corner lookups are independent (no accessor cache threading them together), so their table and value
    // loads overlap.  The root key and the three table offsets are bit-field unions of per-axis parts (NanoVDB.h:2702-2709,
    // 3377-3385, 3893-3900), so they are formed once per axis for c and c+1 and OR-ed per corner.
    const int c3[3] = {tri.i, tri.j, tri.k};
    uint64_t kp[3][2];
    int up[3][2], lp[3][2], fp[3][2];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int d = 0; d < 2; ++d) {
            const int c = c3[a] + d;
            kp[a][d] = (uint64_t)((uint32_t)c >> 12) << (a == 0 ? 42 : a == 1 ? 21 : 0);
            up[a][d] = ((c & 4095) >> 7) << (a == 0 ? 10 : a == 1 ? 5 : 0);
            lp[a][d] = ((c & 127) >> 3) << (a == 0 ? 8 : a == 1 ? 4 : 0);
            fp[a][d] = (c & 7) << (a == 0 ? 6 : a == 1 ? 3 : 0);
        }
    int id[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const int dx = PVDB_CORNER[q][0], dy = PVDB_CORNER[q][1], dz = PVDB_CORNER[q][2];
        int leaf;
        if ((kp[0][dx] | kp[1][dy] | kp[2][dz]) == P.tree.root_key0 && P.tree.n_upper > 0) {
            const int l = __ldg(P.tree.upper_child + (up[0][dx] | up[1][dy] | up[2][dz]));
            leaf = l >= 0 ? __ldg(P.tree.lower_child + (size_t)l * 4096 + (lp[0][dx] | lp[1][dy] | lp[2][dz])) : -1;
        } else {
            leaf = pvdb_find_leaf(P.tree, c3[0] + dx, c3[1] + dy, c3[2] + dz);
        }
        id[q] = leaf >= 0 ? leaf * 512 + (fp[0][dx] | fp[1][dy] | fp[2][dz]) : -1;
    }
    float v[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) v[q] = id[q] >= 0 ? __ldg(P.den + id[q]) : 0.f;
    float acc = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const int dx = PVDB_CORNER[q][0], dy = PVDB_CORNER[q][1], dz = PVDB_CORNER[q][2];
        if (WITH_REC) rec[q] = id[q];
        acc = __fmaf_rn(tri.f(2, dz), __fmul_rn(tri.f(1, dy), __fmul_rn(tri.f(0, dx), v[q])), acc);
    }
    return acc;
}

// One step of the reference's sampler + mask lookup (sample_pts_on_rays :181-192, maskcache_lookup :385-395): is the step's
// point inside the bounding box and in an occupied mask voxel?
__device__ __forceinline__ bool step_in_mask(const MarchParams& P, const float* st, const float* dir, int step) {
    float px, py, pz;
    pvdb_ray_point(st[0], st[1], st[2], dir[0], dir[1], dir[2], P.stepdist, step, px, py, pz);
    const bool outb = (P.xyz_min[0] > px) | (P.xyz_min[1] > py) | (P.xyz_min[2] > pz) | (P.xyz_max[0] < px) |
                      (P.xyz_max[1] < py) | (P.xyz_max[2] < pz);
    if (outb) return false;
    return occ_test(P, pvdb_mask_ijk(px, P.mask_scale[0], P.mask_shift[0]), pvdb_mask_ijk(py, P.mask_scale[1], P.mask_shift[1]),
                    pvdb_mask_ijk(pz, P.mask_scale[2], P.mask_shift[2]));
}

// Can any step of the run [a, b] be in the mask?  Exact, not a heuristic: the point of step i is ONE rounding of a function
// that is monotone in i per axis (dist = fl(stepdist * i); p = fl(fma(dir, dist, start))), and the mask index is one more
// rounding of a monotone function of p (roundf(fl(fma(p, scale, shift))), scale > 0), so per axis every step of the run has its
// mask index between those of the two end steps.  A step is in the mask only if its voxel's bit is set, which requires the
// coarse bit of its 8^3 block: when no block of the index box spanned by the two ends has its coarse bit set, no step of the run
// passes occ_test, whatever the bounding-box test says.
__device__ __forceinline__ bool run_may_hit(const MarchParams& P, const float* st, const float* dir, int a, int b) {
    float pa[3], pb[3];
    pvdb_ray_point(st[0], st[1], st[2], dir[0], dir[1], dir[2], P.stepdist, a, pa[0], pa[1], pa[2]);
    pvdb_ray_point(st[0], st[1], st[2], dir[0], dir[1], dir[2], P.stepdist, b, pb[0], pb[1], pb[2]);
    int lo[3], hi[3];
#pragma unroll
    for (int ax = 0; ax < 3; ++ax) {
        const int ia = pvdb_mask_ijk(pa[ax], P.mask_scale[ax], P.mask_shift[ax]), ib = pvdb_mask_ijk(pb[ax], P.mask_scale[ax], P.mask_shift[ax]);
        lo[ax] = max(min(ia, ib), 0);
        hi[ax] = min(max(ia, ib), P.mask_reso[ax] - 1);
        if (lo[ax] > hi[ax]) return false;
        lo[ax] >>= 3; hi[ax] >>= 3;
    }
    for (int bx = lo[0]; bx <= hi[0]; ++bx)
        for (int by = lo[1]; by <= hi[1]; ++by)
            for (int bz = lo[2]; bz <= hi[2]; ++bz) {
                const int blk = (bx * P.nb[1] + by) * P.nb[2] + bz;
                if ((__ldg(P.occ_coarse + (blk >> 6)) >> (blk & 63)) & 1ull) return true;
            }
    return false;
}

struct MarchOut {
    // per ray
    float* t_min; float* t_max; int32_t* n_steps;
    int32_t *cnt_mask, *cnt_alpha, *cnt_keep, *cnt_alpha_full;
    const int32_t *off_alpha, *off_keep;
    float* alphainv_last;
    // per sample
    int64_t cap_alpha, cap_keep;
    int32_t *s_ray, *s_step; float *s_xyz, *s_density, *s_alpha, *s_T, *s_weight;
    int32_t *k_sample, *k_ray; float* k_xyz;
    int32_t* k_corner;   // [cap_keep][8] or null
    int32_t* counters;
    // per-ray scratch of the count pass: what it computed for each alpha-passing sample, so that the emit pass is a
    // compaction instead of a second march.  Entry (ray, i) = 5 x 16 bytes at ((ray * scr_cap + i) * 5):
    // {step, x, y, z} {density, alpha, T, weight} {keep index, -, -, -} {corner ids 0-3} {corner ids 4-7}
    uint4* scratch; int scr_cap;
    // data-parallel step only: this rank's touched-leaf flags in its symmetric block (dp_exchange.cu).  The sample lists
    // determine the touched leaves before any gradient exists, so the cross-GPU union runs under the rgbnet forward.
    uint8_t* dp_flags;       // one byte per leaf; plain stores of 1 (idempotent: no atomics, benign races)
    int dp_words;            // > 0: k_emit_scratch collects the leaves of its CTA in a shared bitmap of this many words first
};

// Flag the leaves of the eight corners of one alpha-list sample (every such sample feeds the density gradient, the kept ones
// the k0 gradient as well).  Plain stores of 1, skipped while a lane stays in the leaf it flagged last.
__device__ __forceinline__ void dp_flag_corners(uint8_t* __restrict__ flags, const int* rec, int& last_leaf) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const int leaf = rec[q] >> 9;      // -1 stays -1
        if (rec[q] >= 0 && leaf != last_leaf) { flags[leaf] = 1; last_leaf = leaf; }     // consecutive samples mostly share the leaf
    }
}

// One warp per ray.  MODE 0 = count, 1 = emit.  PARITY (count only): keep marching past the early stop so the
// full M1 / M2 counts of the reference are produced too.
// VAR: compile-time variant bits, so that the default kernel carries neither feature's code: 1 = pass A skips runs (P.run_skip),
// 2 = data-parallel step (O.dp_flags is set: the emit path flags the touched leaves, the count pass parks every sample's corners).
template <int MODE, bool PARITY, int VAR>
__device__ __forceinline__ void march_ray(const MarchParams& P, const MarchOut& O, const float* __restrict__ rays_o,
                                          const float* __restrict__ rays_d, const int r, const int lane) {
    float o[3] = {__ldg(rays_o + r * 3), __ldg(rays_o + r * 3 + 1), __ldg(rays_o + r * 3 + 2)};
    float d[3] = {__ldg(rays_d + r * 3), __ldg(rays_d + r * 3 + 1), __ldg(rays_d + r * 3 + 2)};
    float tmin, tmax, st[3], dir[3];
    pvdb_ray_t_minmax(o, d, P.xyz_min, P.xyz_max, P.near, P.far, tmin, tmax);
    const int nsteps = (int)pvdb_ray_n_samples(d, tmin, tmax, P.stepdist);
    pvdb_ray_start_dir(o, d, tmin, st, dir);

    float T_cum = 1.f;
    bool stopped = false;
    int dp_last_leaf = -1;
    int n_mask = 0, n_alpha = 0, n_keep = 0, n_alpha_full = 0;
    int64_t oa = 0, ok = 0;
    if (MODE == 1) { oa = O.off_alpha[r]; ok = O.off_keep[r]; }

    // Per segment of 1024 steps: pass A finds the in-mask steps (one 32-step ballot word per lane), pass B visits them 32 at a
    // time, in step order, so the expensive part (trilinear density, activation) runs with full lanes instead of the 2-3 live
    // lanes a plain 32-step sweep has.
    // Pass A is itself two-level (most steps are in empty space: 355 steps per ray, 17 of them in the mask at F160).  A0: one
    // lane per RUN of 8 consecutive steps asks whether the run can hit the mask at all — run_may_hit, exact, see there.  A1: only
    // the steps of the runs that can are tested one by one, four runs per warp iteration.  Same bits as testing every step.
    for (int seg = 0; seg < nsteps && !(stopped && !PARITY); seg += 1024) {
        unsigned myword = 0;
        if (VAR & 1) {
            const int n_runs = (min(1024, nsteps - seg) + 7) >> 3;      // <= 128
            unsigned rw[4] = {0u, 0u, 0u, 0u};
#pragma unroll
            for (int it = 0; it < 4; ++it) {
                if (it * 32 < n_runs) {
                    const int run = it * 32 + lane;
                    bool maybe = false;
                    if (run < n_runs) maybe = run_may_hit(P, st, dir, seg + run * 8, min(seg + run * 8 + 7, nsteps - 1));
                    rw[it] = __ballot_sync(0xffffffffu, maybe);
                }
            }
            const int c0 = __popc(rw[0]), c1 = c0 + __popc(rw[1]), c2 = c1 + __popc(rw[2]), n_flag = c2 + __popc(rw[3]);
            for (int f0 = 0; f0 < n_flag; f0 += 4) {
                const int f = f0 + (lane >> 3);
                int run = -1;
                if (f < n_flag) {
                    const int k = (f >= c0) + (f >= c1) + (f >= c2);
                    const int base = k == 0 ? 0 : k == 1 ? c0 : k == 2 ? c1 : c2;
                    const unsigned w = k == 0 ? rw[0] : k == 1 ? rw[1] : k == 2 ? rw[2] : rw[3];
                    run = k * 32 + (int)__fns(w, 0, f - base + 1);
                }
                const int step = seg + run * 8 + (lane & 7);
                const bool in_mask = run >= 0 && step < nsteps && step_in_mask(P, st, dir, step);
                const unsigned bits = __ballot_sync(0xffffffffu, in_mask);
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    const int rg = __shfl_sync(0xffffffffu, run, g * 8);
                    if (rg >= 0 && lane == (rg >> 2)) myword |= ((bits >> (g * 8)) & 0xffu) << ((rg & 3) * 8);
                }
            }
        } else {
            const int n_it = min(32, (nsteps - seg + 31) >> 5);
            for (int it = 0; it < n_it; ++it) {
                const int step = seg + it * 32 + lane;
                const bool in_mask = step < nsteps && step_in_mask(P, st, dir, step);
                const unsigned bits = __ballot_sync(0xffffffffu, in_mask);
                if (lane == it) myword = bits;
            }
        }
        const int cnt = __popc(myword);
        int incl = cnt;
#pragma unroll
        for (int o2 = 1; o2 < 32; o2 <<= 1) { const int u = __shfl_up_sync(0xffffffffu, incl, o2); if (lane >= o2) incl += u; }
        const int excl = incl - cnt;
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        n_mask += total;
      for (int b0 = 0; b0 < total; b0 += 32) {
        const int t = b0 + lane;
        const bool in_mask = t < total;
        const int tt = in_mask ? t : total - 1;
        // word holding the tt-th in-mask step: the largest j with excl[j] <= tt (its count is then non-zero)
        int j = 0;
#pragma unroll
        for (int sft = 16; sft >= 1; sft >>= 1) {
            const int e = __shfl_sync(0xffffffffu, excl, j + sft);
            if (e <= tt) j += sft;
        }
        const unsigned word = __shfl_sync(0xffffffffu, myword, j);
        const int ex = __shfl_sync(0xffffffffu, excl, j);
        const int step = seg + j * 32 + (int)__fns(word, 0, tt - ex + 1);
        float x = 0, y = 0, z = 0, dens = 0, alpha = 0;
        int rec[8];
        bool a_ok = false;
        if (in_mask) {
            float px, py, pz;
            pvdb_ray_point(st[0], st[1], st[2], dir[0], dir[1], dir[2], P.stepdist, step, px, py, pz);
            x = pvdb_wld2idx(px, P.xyz_min[0], P.xyz_max[0], P.rm1[0]);
            y = pvdb_wld2idx(py, P.xyz_min[1], P.xyz_max[1], P.rm1[1]);
            z = pvdb_wld2idx(pz, P.xyz_min[2], P.xyz_max[2], P.rm1[2]);
            dens = density_at<true>(P, x, y, z, rec);
            float e;
            alpha = pvdb_raw2alpha(dens, P.act_shift, P.interval, e);
            a_ok = alpha > P.thres;
        }
        unsigned abits = __ballot_sync(0xffffffffu, a_ok);
        n_alpha_full += __popc(abits);
        if (stopped) continue;   // PARITY only: counting beyond the stop
        // sequential transmittance over the set bits, identical in every lane (alpha2weight :590-601)
        float myT = 0, myW = 0;
        int my_ai = -1, my_ki = -1;
        while (abits) {
            const int b = __ffs(abits) - 1;
            abits &= abits - 1;
            const float a = __shfl_sync(0xffffffffu, alpha, b);
            const float w = __fmul_rn(T_cum, a);
            const bool keep = w > P.thres;
            if (lane == b) { myT = T_cum; myW = w; my_ai = n_alpha; my_ki = keep ? n_keep : -1; }
            T_cum = pvdb_T_update(T_cum, a);
            ++n_alpha;
            n_keep += keep;
            if ((double)T_cum < 1e-3) { stopped = true; break; }
        }
        if (MODE == 0 && O.scratch && my_ai >= 0 && my_ai < O.scr_cap) {
            uint4* e = O.scratch + ((int64_t)r * O.scr_cap + my_ai) * 5;
            e[0] = make_uint4((uint32_t)step, __float_as_uint(x), __float_as_uint(y), __float_as_uint(z));
            e[1] = make_uint4(__float_as_uint(dens), __float_as_uint(alpha), __float_as_uint(myT), __float_as_uint(myW));
            e[2] = make_uint4((uint32_t)my_ki, 0u, 0u, 0u);
            if (my_ki >= 0 || (VAR & 2)) {
                e[3] = make_uint4((uint32_t)rec[0], (uint32_t)rec[1], (uint32_t)rec[2], (uint32_t)rec[3]);
                e[4] = make_uint4((uint32_t)rec[4], (uint32_t)rec[5], (uint32_t)rec[6], (uint32_t)rec[7]);
            }
        }
        if (MODE == 1 && my_ai >= 0) {
            if (VAR & 2) dp_flag_corners(O.dp_flags, rec, dp_last_leaf);
            const int64_t ia = oa + my_ai;
            if (ia < O.cap_alpha) {
                O.s_ray[ia] = r; O.s_step[ia] = step;
                O.s_xyz[ia * 3] = x; O.s_xyz[ia * 3 + 1] = y; O.s_xyz[ia * 3 + 2] = z;
                O.s_density[ia] = dens; O.s_alpha[ia] = alpha; O.s_T[ia] = myT; O.s_weight[ia] = myW;
            }
            if (my_ki >= 0) {
                const int64_t ik = ok + my_ki;
                if (ik < O.cap_keep && ia < O.cap_alpha) {
                    O.k_sample[ik] = (int32_t)ia; O.k_ray[ik] = r;
                    O.k_xyz[ik * 3] = x; O.k_xyz[ik * 3 + 1] = y; O.k_xyz[ik * 3 + 2] = z;
                    if (O.k_corner) {
                        int4* kc = reinterpret_cast<int4*>(O.k_corner + ik * 8);
                        kc[0] = make_int4(rec[0], rec[1], rec[2], rec[3]);
                        kc[1] = make_int4(rec[4], rec[5], rec[6], rec[7]);
                    }
                }
            }
        }
        if (stopped && !PARITY) break;
      }
    }
    if (MODE == 0 && lane == 0) {
        O.t_min[r] = tmin; O.t_max[r] = tmax; O.n_steps[r] = nsteps;
        O.cnt_mask[r] = n_mask; O.cnt_alpha[r] = n_alpha; O.cnt_keep[r] = n_keep;
        if (PARITY) O.cnt_alpha_full[r] = n_alpha_full;
        O.alphainv_last[r] = T_cum;
    }
}

// ticket == nullptr: warp w marches ray w.  Otherwise a resident grid of warps draws rays from an atomic ticket counter
// (rays differ 10x in length; 8192 of them are 1.4 waves of static warps, so the tail of a static grid idles a third of
// the machine).  The counter is reset by k_scan_counts, which always follows the count pass.
// Option (off, measured slower — see where it is set): the exclusive scans that turn the per-ray counts into segment offsets
// (scan_counts_cta below) run in the LAST CTA of the count pass to finish instead of in k_scan_counts.
struct ScanTail {
    int32_t *oa, *ok; float* loss; int64_t cap_alpha, cap_keep;
    int enabled, keep_touched;
};
__device__ void scan_counts_cta(const int32_t* __restrict__ ca, const int32_t* __restrict__ ck, int32_t* __restrict__ oa, int32_t* __restrict__ ok, int n,
                                int32_t* __restrict__ counters, float* __restrict__ loss, int64_t cap_alpha, int64_t cap_keep, int keep_touched);

template <int MODE, bool PARITY, int VAR>
__global__ void __launch_bounds__(256, 4) k_march(MarchParams P, MarchOut O, const float* __restrict__ rays_o,
                                                  const float* __restrict__ rays_d, int n_rays, int32_t* __restrict__ ticket, ScanTail S) {
    pvdb_pdl_wait();
    const int lane = threadIdx.x & 31;
    if (ticket == nullptr) {
        const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
        if (r < n_rays) march_ray<MODE, PARITY, VAR>(P, O, rays_o, rays_d, r, lane);
    } else {
        // first ray of every resident warp: its own index (no atomic — thousands of warps drawing their first ticket at once
        // was 12 % of the kernel's stall samples); further rays are drawn from the counter, which numbers the rays after those
        const int n_warps = (gridDim.x * blockDim.x) >> 5;
        int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
        while (r < n_rays) {
            march_ray<MODE, PARITY, VAR>(P, O, rays_o, rays_d, r, lane);
            if (lane == 0) r = n_warps + atomicAdd(ticket, 1);
            r = __shfl_sync(0xffffffffu, r, 0);
        }
    }
    if (MODE == 0 && S.enabled) {
        __shared__ bool s_last;
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();                       // this CTA's counts before its arrival (threadFenceReduction pattern)
            s_last = atomicAdd(O.counters + CNT_MARCH_DONE, 1) == (int)gridDim.x - 1;
        }
        __syncthreads();
        if (s_last) {
            __threadfence();
            scan_counts_cta(O.cnt_alpha, O.cnt_keep, S.oa, S.ok, n_rays, O.counters, S.loss, S.cap_alpha, S.cap_keep, S.keep_touched);
        }
    }
}

// Emit pass as a compaction: one warp per ray copies the count pass's scratch entries to their final, scanned positions.
// A ray with more alpha-passing samples than the scratch holds is simply marched again.
template <int VAR>
__global__ void __launch_bounds__(256) k_emit_scratch(MarchParams P, MarchOut O, const float* __restrict__ rays_o,
                                                      const float* __restrict__ rays_d, int n_rays) {
    // Data-parallel step: the touched-leaf flags of the CTA's eight rays are collected in a shared bitmap and stored once per
    // leaf and CTA — ~10^4 byte stores per step instead of ~10^5 into the same few 32-byte sectors (the flags of the 74 touched
    // leaves of the bench batch lie in three), which cost the emit kernel 7 us.
    extern __shared__ uint32_t s_leafbits[];
    pvdb_pdl_wait();
    const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const bool bitmap = (VAR & 2) && O.dp_words > 0;
    if (bitmap) {
        for (int w = threadIdx.x; w < O.dp_words; w += blockDim.x) s_leafbits[w] = 0u;
        __syncthreads();
    }
    const int na = r < n_rays ? O.cnt_alpha[r] : 0;
    if (r < n_rays && na > O.scr_cap) {
        march_ray<1, false, VAR & 2>(P, O, rays_o, rays_d, r, lane);      // flags its leaves in global memory itself
    } else if (r < n_rays) {
        const int64_t oa = O.off_alpha[r], ok = O.off_keep[r];
        int dp_last_leaf = -1;
        for (int i = lane; i < na; i += 32) {
            const uint4* e = O.scratch + ((int64_t)r * O.scr_cap + i) * 5;
            const uint4 e0 = e[0], e1 = e[1], e2 = e[2];
            const int64_t ia = oa + i;
            const float x = __uint_as_float(e0.y), y = __uint_as_float(e0.z), z = __uint_as_float(e0.w);
            if (ia < O.cap_alpha) {
                O.s_ray[ia] = r; O.s_step[ia] = (int32_t)e0.x;
                O.s_xyz[ia * 3] = x; O.s_xyz[ia * 3 + 1] = y; O.s_xyz[ia * 3 + 2] = z;
                O.s_density[ia] = __uint_as_float(e1.x); O.s_alpha[ia] = __uint_as_float(e1.y);
                O.s_T[ia] = __uint_as_float(e1.z); O.s_weight[ia] = __uint_as_float(e1.w);
            }
            const int ki = (int32_t)e2.x;
            if (VAR & 2) {
                const uint4 c0 = e[3], c1 = e[4];
                const int rec[8] = {(int)c0.x, (int)c0.y, (int)c0.z, (int)c0.w, (int)c1.x, (int)c1.y, (int)c1.z, (int)c1.w};
                if (bitmap) {
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const int leaf = rec[q] >> 9;
                        if (rec[q] >= 0 && leaf != dp_last_leaf) { atomicOr(&s_leafbits[leaf >> 5], 1u << (leaf & 31)); dp_last_leaf = leaf; }
                    }
                } else {
                    dp_flag_corners(O.dp_flags, rec, dp_last_leaf);
                }
            }
            if (ki >= 0) {
                const int64_t ik = ok + ki;
                if (ik < O.cap_keep && ia < O.cap_alpha) {
                    O.k_sample[ik] = (int32_t)ia; O.k_ray[ik] = r;
                    O.k_xyz[ik * 3] = x; O.k_xyz[ik * 3 + 1] = y; O.k_xyz[ik * 3 + 2] = z;
                    if (O.k_corner) {
                        uint4* kc = reinterpret_cast<uint4*>(O.k_corner + ik * 8);
                        kc[0] = e[3];
                        kc[1] = e[4];
                    }
                }
            }
        }
    }
    if (bitmap) {
        __syncthreads();
        for (int w = threadIdx.x; w < O.dp_words; w += blockDim.x) {
            uint32_t bits = s_leafbits[w];
            while (bits) {
                const int bpos = __ffs(bits) - 1;
                bits &= bits - 1;
                O.dp_flags[w * 32 + bpos] = 1;
            }
        }
    }
}

// hit_coarse_geo (dvgo.py:253-270): does any in-bbox sample of the ray fall in an occupied voxel?  Feeds the
// 'in_maskcache' ray sampler (dvgo.py:583-625).  One warp per ray, occupancy bits only.
__global__ void __launch_bounds__(256) k_hit_mask(MarchParams P, const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                                                  int n_rays, uint8_t* __restrict__ hit) {
    const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= n_rays) return;
    float o[3] = {__ldg(rays_o + r * 3), __ldg(rays_o + r * 3 + 1), __ldg(rays_o + r * 3 + 2)};
    float d[3] = {__ldg(rays_d + r * 3), __ldg(rays_d + r * 3 + 1), __ldg(rays_d + r * 3 + 2)};
    float tmin, tmax, st[3], dir[3];
    pvdb_ray_t_minmax(o, d, P.xyz_min, P.xyz_max, P.near, P.far, tmin, tmax);
    const int nsteps = (int)pvdb_ray_n_samples(d, tmin, tmax, P.stepdist);
    pvdb_ray_start_dir(o, d, tmin, st, dir);
    bool any = false;
    for (int base = 0; base < nsteps && !any; base += 32) {
        const int step = base + lane;
        const bool in_mask = step < nsteps && step_in_mask(P, st, dir, step);
        any = __ballot_sync(0xffffffffu, in_mask) != 0;
    }
    if (lane == 0) hit[r] = any ? 1 : 0;
}

// Exclusive scans of the two per-ray counts -> segment offsets in ray order.  One CTA, one pass: each thread owns 8
// consecutive rays (two 16-byte loads per array), scans them in registers, then warp shuffle scan + a 32-entry block scan;
// batches of 8192 rays are chained through a running carry.  Also clears the per-step accumulators of the kernels that
// follow (loss sums, touched-leaf counts, tickets), which saves two memsets in the stream.
__device__ void scan_counts_cta(const int32_t* __restrict__ ca, const int32_t* __restrict__ ck, int32_t* __restrict__ oa,
                                int32_t* __restrict__ ok, int n, int32_t* __restrict__ counters, float* __restrict__ loss,
                                int64_t cap_alpha, int64_t cap_keep, int keep_touched) {
    __shared__ int2 wtot[32];
    __shared__ int2 carry_s;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    if (threadIdx.x == 0) carry_s = make_int2(0, 0);
    if (threadIdx.x < 4) loss[threadIdx.x] = 0.f;
    __syncthreads();
    const int per_round = blockDim.x * 8;
    for (int base = 0; base < n; base += per_round) {
        const int i0 = base + threadIdx.x * 8;
        int a[8], k[8];
        if (i0 + 8 <= n) {
            const int4 a0 = __ldcg(reinterpret_cast<const int4*>(ca + i0)), a1 = __ldcg(reinterpret_cast<const int4*>(ca + i0 + 4));
            const int4 k0 = __ldcg(reinterpret_cast<const int4*>(ck + i0)), k1 = __ldcg(reinterpret_cast<const int4*>(ck + i0 + 4));
            a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
            k[0] = k0.x; k[1] = k0.y; k[2] = k0.z; k[3] = k0.w; k[4] = k1.x; k[5] = k1.y; k[6] = k1.z; k[7] = k1.w;
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) { a[j] = i0 + j < n ? __ldcg(ca + i0 + j) : 0; k[j] = i0 + j < n ? __ldcg(ck + i0 + j) : 0; }
        }
        int2 tot = make_int2(0, 0);
#pragma unroll
        for (int j = 0; j < 8; ++j) { const int ta = a[j], tk = k[j]; a[j] = tot.x; k[j] = tot.y; tot.x += ta; tot.y += tk; }   // exclusive in-thread
        int2 v = tot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int ux = __shfl_up_sync(0xffffffffu, v.x, o), uy = __shfl_up_sync(0xffffffffu, v.y, o);
            if (lane >= o) { v.x += ux; v.y += uy; }
        }
        if (lane == 31) wtot[wid] = v;
        __syncthreads();
        const int2 carry = carry_s;
        if (wid == 0) {
            const int2 own = lane < nw ? wtot[lane] : make_int2(0, 0);
            int2 w = own;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int ux = __shfl_up_sync(0xffffffffu, w.x, o), uy = __shfl_up_sync(0xffffffffu, w.y, o);
                if (lane >= o) { w.x += ux; w.y += uy; }
            }
            if (lane < nw) wtot[lane] = make_int2(w.x - own.x, w.y - own.y);   // exclusive prefix of each warp
            if (lane == 31) carry_s = make_int2(carry.x + w.x, carry.y + w.y);
        }
        __syncthreads();
        const int2 pre = wtot[wid];
        const int px = carry.x + pre.x + v.x - tot.x, py = carry.y + pre.y + v.y - tot.y;   // exclusive prefix of this thread
        if (i0 + 8 <= n) {
            *reinterpret_cast<int4*>(oa + i0) = make_int4(px + a[0], px + a[1], px + a[2], px + a[3]);
            *reinterpret_cast<int4*>(oa + i0 + 4) = make_int4(px + a[4], px + a[5], px + a[6], px + a[7]);
            *reinterpret_cast<int4*>(ok + i0) = make_int4(py + k[0], py + k[1], py + k[2], py + k[3]);
            *reinterpret_cast<int4*>(ok + i0 + 4) = make_int4(py + k[4], py + k[5], py + k[6], py + k[7]);
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (i0 + j < n) { oa[i0 + j] = px + a[j]; ok[i0 + j] = py + k[j]; }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const int2 w = carry_s;
        oa[n] = w.x; ok[n] = w.y;
        counters[CNT_M_ALPHA] = w.x; counters[CNT_M_KEEP] = w.y;
        counters[CNT_OVERFLOW] = (w.x > cap_alpha || w.y > cap_keep) ? 1 : 0;
        // keep_touched: gradients of an earlier backward are still waiting for their update (gradient accumulation): their leaves
        // stay on the touched lists, which only the update empties
        if (!keep_touched) { counters[CNT_N_TOUCHED_DEN] = 0; counters[CNT_N_TOUCHED_K0] = 0; }
        counters[CNT_RAY_TICKET] = 0; counters[CNT_CTA_DONE] = 0;
        counters[CNT_MARCH_DONE] = 0;
    }
}
__global__ void __launch_bounds__(1024) k_scan_counts(const int32_t* __restrict__ ca, const int32_t* __restrict__ ck,
                                                      int32_t* __restrict__ oa, int32_t* __restrict__ ok, int n,
                                                      int32_t* __restrict__ counters, float* __restrict__ loss, int64_t cap_alpha,
                                                      int64_t cap_keep, int keep_touched) {
    pvdb_pdl_wait();
    scan_counts_cta(ca, ck, oa, ok, n, counters, loss, cap_alpha, cap_keep, keep_touched);
}

// ---------------------------------------------------------------------------------------------
// F5 — compositing, losses, and the first backward step (per ray; one warp per ray).
//   rgb_marched = sum_s w*rgb + alphainv_last*bg                          (dvgo.py:365-370)
//   loss = w_main*mse + w_ent*entropy_last + w_per*rgbper                 (run.py:551-574)
//   g_marched = w_main*2*(marched-target)/(3N);  g_last = sum_c g_marched*bg + w_ent*(-(log p - log(1-p)))/N
//   g_rgb = w*g_marched + w_per*2*(rgb-target)*w/N;  g_logit = g_rgb*rgb*(1-rgb);  g_w = sum_c rgb*g_marched
// ---------------------------------------------------------------------------------------------
struct CompositeParams {
    const int32_t* off_keep; const int32_t* k_sample; const float* s_weight;
    float* k_rgb;            // in: rgb[M3][3]; out (backward enabled): g_logit[M3][3]
    float* k_gw;             // out: dL/dw per kept sample
    const float* alphainv_last; const float* target;
    float* rgb_marched; float* grad_last; float* loss;
    int32_t* cta_done;       // zeroed by k_scan_counts; the last CTA to finish folds the three sums into loss[0]
    float bg, w_main, w_ent, w_per, inv_N;   // inv_N = 1/N_global
    int64_t cap_keep;
    int do_backward;
};
// Eight lanes per ray (a fine-stage ray keeps ~10 samples: with a warp per ray two thirds of the lanes idled through ~430
// instructions of per-ray scalar work, 3.5 M warp instructions per step).
constexpr int CG = 8;
__global__ void __launch_bounds__(256) k_composite(CompositeParams C, int n_rays) {
    pvdb_pdl_wait();
    const int r = (blockIdx.x * blockDim.x + threadIdx.x) / CG;
    const int lane = threadIdx.x & 31, sub = threadIdx.x & (CG - 1);
    const bool live = r < n_rays;      // whole groups; every lane of the warp takes part in the shuffles below
    float l_mse = 0, l_ent = 0, l_per = 0;
    {
        const int b = live ? C.off_keep[r] : 0, e = live ? (int)min((int64_t)C.off_keep[r + 1], C.cap_keep) : 0;
        float acc[3] = {0, 0, 0};
        for (int s = b + sub; s < e; s += CG) {
            const float w = C.s_weight[C.k_sample[s]];
#pragma unroll
            for (int c = 0; c < 3; ++c) acc[c] += w * C.k_rgb[(size_t)s * 3 + c];
        }
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int o = CG / 2; o; o >>= 1) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], o);
        const float ail = live ? C.alphainv_last[r] : 0.5f;
        float tg[3], gm[3], gsum = 0;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            tg[c] = C.target && live ? __ldg(C.target + r * 3 + c) : 0.f;
            const float m = acc[c] + ail * C.bg;
            if (sub == 0 && live) C.rgb_marched[r * 3 + c] = m;
            const float df = m - tg[c];
            l_mse += df * df;
            gm[c] = C.w_main * 2.0f * df * C.inv_N * (1.0f / 3.0f);
            gsum += gm[c];
        }
        const float p = fminf(fmaxf(ail, 1e-6f), 1.0f - 1e-6f);
        const float lp = logf(p), l1p = logf(1.0f - p);
        l_ent = -(p * lp + (1.0f - p) * l1p);
        if (C.do_backward) {
            float ge = 0.f;
            if (C.w_ent > 0.f && ail >= 1e-6f && ail <= 1.0f - 1e-6f) ge = C.w_ent * (-(logf(ail) - logf(1.0f - ail))) * C.inv_N;
            if (sub == 0 && live) C.grad_last[r] = gsum * C.bg + ge;
        }
        for (int s = b + sub; C.target && s < e; s += CG) {
            const float w = C.s_weight[C.k_sample[s]];
            float gw = 0, per = 0;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float col = C.k_rgb[(size_t)s * 3 + c];
                const float dc = col - tg[c];
                per += dc * dc;
                if (C.do_backward) {
                    const float grgb = w * gm[c] + C.w_per * 2.0f * dc * w * C.inv_N;
                    C.k_rgb[(size_t)s * 3 + c] = grgb * col * (1.0f - col);
                    gw += col * gm[c];
                }
            }
            l_per += per * w;
            if (C.do_backward) C.k_gw[s] = gw;
        }
        // warp sums: the per-ray terms count once per group (lane sub 0 of a live group), the per-sample term on every lane
        if (sub != 0 || !live) { l_mse = 0.f; l_ent = 0.f; }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            l_mse += __shfl_xor_sync(0xffffffffu, l_mse, o);
            l_ent += __shfl_xor_sync(0xffffffffu, l_ent, o);
            l_per += __shfl_xor_sync(0xffffffffu, l_per, o);
        }
    }
    // block reduction of the three loss sums -> one atomic per CTA
    __shared__ float red[3][8];
    const int wid = threadIdx.x >> 5;
    if (lane == 0) { red[0][wid] = l_mse; red[1][wid] = l_ent; red[2][wid] = l_per; }
    __syncthreads();
    if (threadIdx.x < 3 && C.target) {
        float s = 0;
        for (int w = 0; w < 8; ++w) s += red[threadIdx.x][w];
        const float scale = threadIdx.x == 0 ? C.inv_N * (1.0f / 3.0f) : C.inv_N;
        atomicAdd(C.loss + 1 + threadIdx.x, s * scale);
    }
    if (C.target) {   // last CTA out: total loss (run.py:551-574)
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0 && atomicAdd(C.cta_done, 1) == (int)gridDim.x - 1) {
            __threadfence();
            const volatile float* L = C.loss;
            C.loss[0] = C.w_main * L[1] + C.w_ent * L[2] + C.w_per * L[3];
        }
    }
}
// ---------------------------------------------------------------------------------------------
// B2 — per-ray reverse pass (alpha2weight_backward :654-677 + raw2alpha_backward :507-517).  One warp per ray: the
// reference's sequential recurrence  back_cum += gw*w  is a suffix sum, evaluated here 32 samples at a time from the far
// end of the segment with a shuffle scan (summation order differs from the reference's serial loop at the 1e-7 level).
// The gradient wrt the weight of a kept sample is fetched through its position among the ray's kept samples.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_ray_bwd(const int32_t* __restrict__ off_alpha, const int32_t* __restrict__ off_keep,
                                                 const float* __restrict__ s_alpha, const float* __restrict__ s_T,
                                                 const float* __restrict__ s_weight, const float* __restrict__ s_density,
                                                 const float* __restrict__ k_gw, const float* __restrict__ alphainv_last,
                                                 const float* __restrict__ grad_last, float* __restrict__ s_gden, int n_rays,
                                                 float thres, float act_shift, float interval, int64_t cap_alpha, int64_t cap_keep) {
    const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= n_rays) return;
    const int64_t b = off_alpha[r], e = off_alpha[r + 1];
    if (e > cap_alpha) return;
    int64_t ks_end = off_keep[r + 1];                       // kept samples are consumed from the end of the ray
    float carry = __fmul_rn(grad_last[r], alphainv_last[r]);   // back_cum before the last sample
    for (int64_t top = e; top > b; top -= 32) {
        const int64_t i = top - 1 - lane;                   // lane 0 = farthest sample of this chunk
        const bool live = i >= b;
        float w = 0.f, a = 0.f, T = 0.f, dens = 0.f;
        if (live) { w = s_weight[i]; a = s_alpha[i]; T = s_T[i]; dens = s_density[i]; }
        const bool kept = live && w > thres;
        const unsigned kb = __ballot_sync(0xffffffffu, kept);
        float gw = 0.f;
        if (kept) {
            const int64_t ks = ks_end - 1 - __popc(kb & ((1u << lane) - 1u));
            gw = ks < cap_keep ? k_gw[ks] : 0.f;
        }
        ks_end -= __popc(kb);
        // back_cum seen by sample i = carry + sum of gw*w over the samples behind it (lanes < lane)
        const float c = __fmul_rn(gw, w);
        float incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const float u = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += u;
        }
        const float back_cum = carry + (incl - c);
        if (live) {
            const float ga = pvdb_a2w_grad(gw, T, back_cum, a);
            const float ex = expf(__fadd_rn(dens, act_shift));
            s_gden[i] = pvdb_raw2alpha_bwd(ex, ga, interval);
        }
        carry += __shfl_sync(0xffffffffu, incl, 31);
    }
}

__device__ __forceinline__ void red_add(float* addr, float v) { asm volatile("red.global.add.f32 [%0], %1;" ::"l"(addr), "f"(v) : "memory"); }
__device__ __forceinline__ void red_add4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// B3 — density gradient scatter (densityvdb.cu:143-167) over the compacted alpha list; marks touched leaves.
__global__ void __launch_bounds__(256) k_density_scatter(pvdb_tree t, float* __restrict__ den_grad, const float* __restrict__ s_xyz,
                                                         const float* __restrict__ s_gden, const int32_t* __restrict__ counters,
                                                         int32_t* __restrict__ touched, int32_t* __restrict__ touched_list,
                                                         int32_t* __restrict__ counters_w, int64_t cap_alpha) {
    pvdb_pdl_wait();
    const int64_t n = min((int64_t)counters[CNT_M_ALPHA], cap_alpha);
    for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < n; s += (int64_t)gridDim.x * blockDim.x) {
        const float g = s_gden[s];
        if (g == 0.f) continue;   // adding an exact zero leaves the plane unchanged
        PvdbTri tri;
        tri.set(s_xyz[s * 3], s_xyz[s * 3 + 1], s_xyz[s * 3 + 2]);
        PvdbLeafCache cache;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int dx = PVDB_CORNER[q][0], dy = PVDB_CORNER[q][1], dz = PVDB_CORNER[q][2];
            const int x = tri.i + dx, y = tri.j + dy, z = tri.k + dz;
            const int leaf = cache.find(t, x, y, z);
            if (leaf < 0) continue;
            red_add(den_grad + (size_t)leaf * 512 + pvdb_leaf_off(x, y, z),
                    __fmul_rn(__fmul_rn(__fmul_rn(g, tri.f(0, dx)), tri.f(1, dy)), tri.f(2, dz)));
            pvdb_touch_leaf(touched, touched_list, counters_w + CNT_N_TOUCHED_DEN, leaf);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// U1 — Adam over touched leaves (mode 1 semantics: a voxel whose gradient is zero is skipped, so leaves no
// sample touched need no visit at all; densityvdb.cu:298-330, colorvdb.cu:297-331).  The same pass clears the
// gradient of the active voxels it visits and the touched flag, which replaces the full-grid zero_grad sweep
// (densityvdb.cu:353-368) of the next iteration.
// ---------------------------------------------------------------------------------------------
// One launch for the whole update: Adam over the touched leaves (lists built by the scatter kernels) + the dense Adam
// of the 22019 rgbnet parameters (adam_upd_kernel.cu:9-23) in the trailing blocks.
// Rebuild a touched-leaf list from the flags (update-only calls after a data-parallel flag all-reduce).
__global__ void __launch_bounds__(256) k_touched_compact(const int32_t* __restrict__ touched, int n_leaf, int32_t* __restrict__ list,
                                                         int32_t* __restrict__ count) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool t = i < n_leaf && touched[i] != 0;
    const unsigned bits = __ballot_sync(0xffffffffu, t);
    const int lane = threadIdx.x & 31;
    int base = 0;
    if (lane == 0 && bits) base = atomicAdd(count, __popc(bits));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (t) list[base + __popc(bits & ((1u << lane) - 1))] = i;
}

struct UpdateArgs {
    pvdb_tree tree;
    float *den, *den_g, *den_m, *den_v, *k0, *k0_g, *k0_m, *k0_v, *net, *net_g, *net_m, *net_v;
    int32_t *den_touched, *k0_touched, *den_list, *k0_list, *counters;
    float den_stepsz, k0_stepsz, net_stepsize, eps, b0, b1;
    int leaf_blocks;
    int block_offset;   // 0: leaf work starts at CTA 0; leaf_blocks: a launch of the rgbnet Adam CTAs only
    const uint32_t* dp_tiles_signal;   // data-parallel step: D[0..world) of the own block — every owner has stored its sums into
    uint32_t dp_tiles_epoch;           // this rank's gradient planes once these words reached the epoch (nullptr: no wait)
    int dp_world; int32_t* dp_err;
    const float* scalars;   // optional device array {den_stepsz, k0_stepsz, net_stepsize}: overrides the three values above (CUDA-graph replay)
    PvdbDpNetWait dp;   // dp.world > 1: the rgbnet gradient is the rank-ordered sum of the world slots the peers pushed into this block
};
__global__ void __launch_bounds__(256) k_update_fused(UpdateArgs U) {
    pvdb_pdl_wait();
    const int bid = (int)blockIdx.x + U.block_offset;
    if (bid >= U.leaf_blocks) {
        const int i = (bid - U.leaf_blocks) * blockDim.x + threadIdx.x;
        if (U.dp.world > 1) {
            if (threadIdx.x < U.dp.world) wait_epoch(U.dp.signal + threadIdx.x, U.dp.epoch, U.dp.err, 1);
            __syncthreads();
        }
        if (i < PVDB_NET_N) {
            float g = U.net_g[i];
            if (U.dp.world > 1) {
                g = __ldcg(U.dp.src + i);
                for (int r = 1; r < U.dp.world; ++r) g += __ldcg(U.dp.src + (size_t)r * PVDB_DP_NET_PAD + i);
            }
            pvdb_dense_adam_update(U.net[i], U.net_m[i], U.net_v[i], g, 1.f, false, U.scalars ? __ldg(U.scalars + 2) : U.net_stepsize, U.b0, U.b1, U.eps);
            U.net_g[i] = 0.f;
        }
        return;
    }
    if (U.dp_tiles_signal) {
        if (threadIdx.x < U.dp_world) wait_epoch(U.dp_tiles_signal + threadIdx.x, U.dp_tiles_epoch, U.dp_err, 1);
        __syncthreads();
    }
    const int nd = U.counters[CNT_N_TOUCHED_DEN], nk = U.counters[CNT_N_TOUCHED_K0];
    const float omb0 = __fsub_rn(1.0f, U.b0), omb1 = __fsub_rn(1.0f, U.b1);
    // Work items: (touched density leaf) and (touched k0 leaf, quarter); a persistent grid strides over them.  One thread per
    // (voxel, Vec3 group) with scalar loads: a variant with one thread per k0 voxel and 3 x 16-byte loads per plane needed 80
    // registers instead of 31 and was 10 % SLOWER at S512 (0.49 vs 0.45 ms for 18 k touched leaves: the kernel lives on loads
    // in flight, i.e. on occupancy), equal at F160.
    for (int w = bid; w < nd + nk * 4; w += U.leaf_blocks) {
        const bool is_den = w < nd;
        const int leaf = is_den ? U.den_list[w] : U.k0_list[(w - nd) >> 2];
        const int part = is_den ? 0 : (w - nd) & 3;
        float *p = is_den ? U.den : U.k0, *g = is_den ? U.den_g : U.k0_g, *m = is_den ? U.den_m : U.k0_m, *v = is_den ? U.den_v : U.k0_v;
        const float stepsz = U.scalars ? __ldg(U.scalars + (is_den ? 0 : 1)) : (is_den ? U.den_stepsz : U.k0_stepsz);
        const int C = is_den ? 1 : 12, G = is_den ? 1 : 3, ngrp = C / G;
        // density: 512 voxels by 256 threads (2 each); k0 quarter: 128 voxels x 4 groups = 512 items (2 each)
        for (int e = threadIdx.x; e < 512; e += blockDim.x) {
            const int off = is_den ? e : part * 128 + (e >> 2), grp = is_den ? 0 : (e & 3);
            if (!pvdb_mask_bit(U.tree.leaf_mask, leaf, off)) continue;
            const size_t base = ((size_t)leaf * 512 + off) * C + (size_t)grp * G;
            float gg[3];
            bool allzero = true;
            for (int c = 0; c < G; ++c) { gg[c] = g[base + c]; allzero = allzero && gg[c] == 0.0f; }
            if (allzero) continue;
            for (int c = 0; c < G; ++c) {
                const float nm = __fmaf_rn(omb0, gg[c], __fmul_rn(U.b0, m[base + c]));
                const float nv = __fmaf_rn(gg[c], __fmul_rn(omb1, gg[c]), __fmul_rn(U.b1, v[base + c]));
                m[base + c] = nm;
                v[base + c] = nv;
                p[base + c] = __fsub_rn(p[base + c], __fdiv_rn(__fmul_rn(stepsz, nm), __fadd_rn(U.eps, __fsqrt_rn(nv))));
                g[base + c] = 0.f;
            }
        }
        if (threadIdx.x == 0 && part == 0) (is_den ? U.den_touched : U.k0_touched)[leaf] = 0;
        (void)ngrp;
    }
}

// Occupancy bits: fine[(bx*nby+by)*nbz+bz][8] + one coarse bit per block.  One warp per block.
__global__ void __launch_bounds__(256) k_occ_build(const uint8_t* __restrict__ mask, int rx, int ry, int rz, int nbx, int nby, int nbz,
                                                   uint64_t* __restrict__ fine, unsigned long long* __restrict__ coarse) {
    const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (b >= nbx * nby * nbz) return;
    const int bz = b % nbz, by = (b / nbz) % nby, bx = b / (nbz * nby);
    bool any = false;
    for (int w = 0; w < 8; ++w) {   // word w covers dx = w, (dy,dz) = 64 bits; each lane builds 2 bits
        uint64_t bits = 0;
        for (int h = 0; h < 2; ++h) {
            const int n = lane * 2 + h;          // bit index inside the word: dy = n>>3, dz = n&7
            const int x = bx * 8 + w, y = by * 8 + (n >> 3), z = bz * 8 + (n & 7);
            if (x < rx && y < ry && z < rz && mask[((size_t)x * ry + y) * rz + z]) bits |= 1ull << n;
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) bits |= __shfl_xor_sync(0xffffffffu, bits, o);
        if (lane == 0) fine[(size_t)b * 8 + w] = bits;
        any = any || bits != 0;
    }
    if (lane == 0 && any) atomicOr(coarse + (b >> 6), 1ull << (b & 63));
}

}  // namespace

extern "C" int pvdb_occ_build(const uint8_t* mask, int rx, int ry, int rz, uint64_t* fine, uint64_t* coarse, void* stream) {
    PVDB_CHECK_ARG(mask && fine && coarse && rx > 0 && ry > 0 && rz > 0, "bad arguments");
    const int nbx = (rx + 7) / 8, nby = (ry + 7) / 8, nbz = (rz + 7) / 8;
    const int nb = nbx * nby * nbz;
    cudaStream_t st = (cudaStream_t)stream;
    PVDB_CUDA(cudaMemsetAsync(coarse, 0, (size_t)((nb + 63) / 64) * 8, st));
    k_occ_build<<<pvdb_grid_for((int64_t)nb * 32, 256), 256, 0, st>>>(mask, rx, ry, rz, nbx, nby, nbz, fine,
                                                                      (unsigned long long*)coarse);
    PVDB_LAUNCH_CHECK();
    return PVDB_OK;
}

// Debug timeline (environment PVDB_STAMPS=1): one-thread kernels write %globaltimer into a device array at the marked points of
// the fused step, on the stream they are enqueued on; pvdb_debug_stamps_fetch copies the last step's stamps to the host.  Used
// by scratch/dp_timeline.py to see where a data-parallel step waits.  Off: no kernel is launched.
__global__ void k_stamp(unsigned long long* out, int slot) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    out[slot] = t;
}
static unsigned long long* g_stamps = nullptr;
static int g_stamps_on = -1;
static void stamp(cudaStream_t st, int slot) {
    if (g_stamps_on < 0) { const char* e = getenv("PVDB_STAMPS"); g_stamps_on = e && atoi(e) != 0; }
    if (!g_stamps_on) return;
    if (!g_stamps) { cudaMalloc(&g_stamps, 64 * sizeof(unsigned long long)); cudaMemset(g_stamps, 0, 64 * sizeof(unsigned long long)); }
    k_stamp<<<1, 1, 0, st>>>(g_stamps, slot);
}
unsigned long long* pvdb_debug_stamps_ptr() {      // device array of 64 stamps, or nullptr when stamping is off (dp_exchange.cu)
    if (g_stamps_on < 0) { const char* e = getenv("PVDB_STAMPS"); g_stamps_on = e && atoi(e) != 0; }
    if (!g_stamps_on) return nullptr;
    if (!g_stamps) { cudaMalloc(&g_stamps, 64 * sizeof(unsigned long long)); cudaMemset(g_stamps, 0, 64 * sizeof(unsigned long long)); }
    return g_stamps;
}
extern "C" int pvdb_debug_stamps_fetch(unsigned long long* out64) {
    if (!g_stamps) return 1;
    return cudaMemcpy(out64, g_stamps, 64 * sizeof(unsigned long long), cudaMemcpyDeviceToHost) == cudaSuccess ? 0 : 2;
}

// test switch (pvdb_debug_set_run_skip): 0 = pass A of the march tests every step one by one
static int g_run_skip = -1;     // -1: not decided yet (environment PVDB_RUN_SKIP, else the default)
extern "C" void pvdb_debug_set_run_skip(int on) { g_run_skip = on ? 1 : 0; }

// The rgbnet Adam step size (adam_upd_kernel.cu:72) as the fused step computes it from cfg->net_lr / net_step: exported so that a
// caller who feeds pvdb_train_bufs.step_scalars (CUDA-graph replay) writes the same bits.
extern "C" float pvdb_dense_adam_stepsize_host(float lr, float beta0, float beta1, int step) { return pvdb_dense_adam_stepsize(lr, beta0, beta1, step); }

static int fill_march(const pvdb_train_cfg* cfg, const pvdb_train_bufs* b, MarchParams& P, MarchOut& O) {
    P.tree = *b->tree;
    P.den = b->den;
    P.occ_fine = b->occ_fine;
    P.occ_coarse = b->occ_coarse;
    for (int a = 0; a < 3; ++a) {
        P.xyz_min[a] = cfg->xyz_min[a]; P.xyz_max[a] = cfg->xyz_max[a];
        P.rm1[a] = (float)(cfg->reso[a] - 1);
        P.mask_reso[a] = cfg->mask_reso[a];
        P.nb[a] = (cfg->mask_reso[a] + 7) / 8;
        P.mask_scale[a] = cfg->mask_scale[a]; P.mask_shift[a] = cfg->mask_shift[a];
    }
    P.near = cfg->near; P.far = cfg->far; P.stepdist = cfg->stepdist; P.act_shift = cfg->act_shift;
    P.interval = cfg->interval; P.thres = cfg->fast_color_thres;
    // Default OFF: measured on B200 (profiles/prof_march_r02.md) the two-level pass A issues MORE instructions than testing every
    // step (19.2 M vs 17.7 M warp instructions at F160, 39 vs 36 us; 0.297 vs 0.270 ms at S512) — the 8^3 coarse blocks are not
    // selective enough for runs of 8 steps, and pass A is only a third of the kernel.  Kept (exact, tested) behind PVDB_RUN_SKIP=1.
    if (g_run_skip < 0) { const char* e = getenv("PVDB_RUN_SKIP"); g_run_skip = e ? (atoi(e) != 0) : 0; }
    P.run_skip = g_run_skip;
    O.t_min = b->t_min; O.t_max = b->t_max; O.n_steps = b->n_steps;
    O.cnt_mask = b->cnt_mask; O.cnt_alpha = b->cnt_alpha; O.cnt_keep = b->cnt_keep; O.cnt_alpha_full = b->cnt_alpha_full;
    O.off_alpha = b->off_alpha; O.off_keep = b->off_keep; O.alphainv_last = b->alphainv_last;
    O.cap_alpha = b->cap_alpha; O.cap_keep = b->cap_keep;
    O.s_ray = b->s_ray; O.s_step = b->s_step; O.s_xyz = b->s_xyz; O.s_density = b->s_density; O.s_alpha = b->s_alpha;
    O.s_T = b->s_T; O.s_weight = b->s_weight;
    O.k_sample = b->k_sample; O.k_ray = b->k_ray; O.k_xyz = b->k_xyz; O.k_corner = b->k_corner;
    O.counters = b->counters;
    O.scratch = reinterpret_cast<uint4*>(b->march_scratch);
    O.scr_cap = b->march_scratch ? b->scratch_per_ray : 0;
    O.dp_flags = nullptr; O.dp_words = 0;
    return 0;
}

// ---------------------------------------------------------------------------------------------
// Data-parallel exchange helpers.  After the ranks' touched flags were MAX-all-reduced, pack the
// gradient tiles of the union leaves (density [512] + k0 [512][12] per leaf) and the rgbnet gradients into one
// contiguous buffer for a single SUM all-reduce; unpack writes the reduced tiles back.  The union list is built on
// the device; its length goes to a pinned host word so the caller can size the collective.
// ---------------------------------------------------------------------------------------------
// One CTA; the list is in ascending leaf order so that every rank packs the same leaf into the same slot.
__global__ void __launch_bounds__(1024) k_dp_union_list(int32_t* __restrict__ den_touched, int32_t* __restrict__ k0_touched, int n_leaf,
                                                        int32_t* __restrict__ list, int32_t* __restrict__ count) {
    __shared__ int warp_cnt[32];
    __shared__ int running;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) running = 0;
    __syncthreads();
    for (int base = 0; base < n_leaf; base += 1024) {
        const int i = base + threadIdx.x;
        const bool t = i < n_leaf && (den_touched[i] | k0_touched[i]) != 0;
        if (t) { den_touched[i] = 1; k0_touched[i] = 1; }
        const unsigned bits = __ballot_sync(0xffffffffu, t);
        if (lane == 0) warp_cnt[warp] = __popc(bits);
        __syncthreads();
        int c = warp_cnt[lane], incl = c;
        #pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const int o = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += o; }
        const int warp_off = __shfl_sync(0xffffffffu, incl - c, warp);
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        const int start = running;
        if (t) list[start + warp_off + __popc(bits & ((1u << lane) - 1))] = i;
        __syncthreads();
        if (threadIdx.x == 0) running = start + total;
        __syncthreads();
    }
    if (threadIdx.x == 0) *count = running;
}
// grid: one CTA per union slot (grid-stride), 256 threads move 512*13 floats as float4
template <bool PACK>
__global__ void __launch_bounds__(256) k_dp_move(float* __restrict__ den_grad, float* __restrict__ k0_grad, float* __restrict__ net_grad,
                                                 const int32_t* __restrict__ list, const int32_t* __restrict__ count,
                                                 float* __restrict__ buf) {
    const int n = *count;
    float* net_slot = buf + (size_t)n * (512 * 13);
    for (int slot = blockIdx.x; slot <= n; slot += gridDim.x) {
        if (slot == n) {   // trailing slot: rgbnet gradients
            for (int i = threadIdx.x; i < PVDB_NET_N; i += blockDim.x) {
                if (PACK) net_slot[i] = net_grad[i]; else net_grad[i] = net_slot[i];
            }
            continue;
        }
        const int leaf = list[slot];
        float4* b4 = reinterpret_cast<float4*>(buf + (size_t)slot * (512 * 13));
        float4* d4 = reinterpret_cast<float4*>(den_grad + (size_t)leaf * 512);
        float4* k4 = reinterpret_cast<float4*>(k0_grad + (size_t)leaf * 512 * 12);
        for (int i = threadIdx.x; i < 128 + 1536; i += blockDim.x) {
            float4* g = i < 128 ? d4 + i : k4 + (i - 128);
            if (PACK) b4[i] = *g; else *g = b4[i];
        }
    }
}

extern "C" int pvdb_dp_pack(const pvdb_train_bufs* b, int32_t* union_list, int32_t* union_count_dev, int32_t* union_count_host,
                            float* buf, int64_t buf_capacity_floats, void* stream) {
    PVDB_CHECK_ARG(b && b->tree && union_list && union_count_dev && union_count_host && buf, "null pointer");
    PVDB_CHECK_ARG(buf_capacity_floats >= (int64_t)b->tree->n_leaf * 512 * 13 + PVDB_NET_N, "pack buffer smaller than the worst case");
    cudaStream_t st = (cudaStream_t)stream;
    const int n_leaf = b->tree->n_leaf;
    k_dp_union_list<<<1, 1024, 0, st>>>(b->den_touched, b->k0_touched, n_leaf, union_list, union_count_dev);
    PVDB_LAUNCH_CHECK();
    PVDB_CUDA(cudaMemcpyAsync(union_count_host, union_count_dev, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    k_dp_move<true><<<PVDB_SMS * 2, 256, 0, st>>>(b->den_grad, b->k0_grad, b->net_grad, union_list, union_count_dev, buf);
    PVDB_LAUNCH_CHECK();
    return PVDB_OK;
}
extern "C" int pvdb_dp_unpack(const pvdb_train_bufs* b, const int32_t* union_list, const int32_t* union_count_dev, float* buf,
                              void* stream) {
    PVDB_CHECK_ARG(b && b->tree && union_list && union_count_dev && buf, "null pointer");
    k_dp_move<false><<<PVDB_SMS * 2, 256, 0, (cudaStream_t)stream>>>(b->den_grad, b->k0_grad, b->net_grad, union_list, union_count_dev, buf);
    PVDB_LAUNCH_CHECK();
    return PVDB_OK;
}

// One launch that gathers a batch (rays_o, rays_d, viewdirs, target: [n][3] each, anywhere in device memory) into one staging
// buffer [4][n][3]: a captured CUDA graph of the step reads its inputs from fixed addresses.
__global__ void __launch_bounds__(256) k_stage_rays(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ c,
                                                    const float* __restrict__ d, int n3, float* __restrict__ stage) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n3) return;
    stage[i] = a[i]; stage[n3 + i] = b[i]; stage[2 * n3 + i] = c[i];
    if (d) stage[3 * n3 + i] = d[i];
}
extern "C" int pvdb_stage_rays(const float* rays_o, const float* rays_d, const float* viewdirs, const float* target, int n_rays, float* stage,
                               void* stream) {
    PVDB_CHECK_ARG(rays_o && rays_d && viewdirs && stage && n_rays > 0, "bad arguments");
    k_stage_rays<<<pvdb_grid_for((int64_t)n_rays * 3, 256), 256, 0, (cudaStream_t)stream>>>(rays_o, rays_d, viewdirs, target, n_rays * 3, stage);
    PVDB_LAUNCH_CHECK();
    return PVDB_OK;
}

extern "C" int pvdb_rays_hit_mask(const pvdb_train_cfg* cfg, const pvdb_train_bufs* b, const float* rays_o, const float* rays_d,
                                  int n_rays, uint8_t* hit, void* stream) {
    PVDB_CHECK_ARG(cfg && b && b->tree && rays_o && rays_d && hit, "null pointer");
    if (n_rays <= 0) return PVDB_OK;
    MarchParams P;
    MarchOut O;
    fill_march(cfg, b, P, O);
    k_hit_mask<<<pvdb_grid_for((int64_t)n_rays * 32, 256), 256, 0, (cudaStream_t)stream>>>(P, rays_o, rays_d, n_rays, hit);
    PVDB_LAUNCH_CHECK();
    return PVDB_OK;
}

static int train_step_impl(const pvdb_train_cfg* cfg, const pvdb_train_bufs* b, const float* rays_o, const float* rays_d,
                           const float* viewdirs, const float* target, int n_rays, int phases, void* stream,
                           const pvdb_dp_peers* peers, uint32_t dp_step) {
    PVDB_CHECK_ARG(cfg && b && b->tree, "null cfg/bufs");
    const bool direct = pvdb_direct_colour(cfg);      // coarse stage: 3 colour channels, no rgbnet (coarse.cu)
    PVDB_CHECK_ARG((cfg->k0_dim == 12 && cfg->net_width == 128) || direct,
                   "the fused step is specialised for k0_dim=12 with rgbnet_width=128 (fine stage) or k0_dim=3 without rgbnet (coarse stage)");
    PVDB_CHECK_ARG(!direct || !peers, "the data-parallel step covers the fine stage only");
    PVDB_CHECK_ARG(n_rays > 0, "n_rays must be positive");
    const bool do_fwd = phases & PVDB_PHASE_FORWARD, do_bwd = phases & PVDB_PHASE_BACKWARD, do_upd = phases & PVDB_PHASE_UPDATE;
    PVDB_CHECK_ARG(!do_bwd || (do_fwd && target), "backward needs the forward phase and target colours");
    PVDB_CHECK_ARG(!do_upd || direct || (cfg->den_mode == 1 && cfg->k0_mode == 1),
                   "the fine-stage update implements stepmode 1 (skip zero grad); stepmodes 0 / 2 are the coarse stage's (k0_dim=3)");
    PVDB_CHECK_ARG(!do_upd || !direct || cfg->den_mode != 2 || b->den_perlr, "stepmode 2 needs the per-voxel lr plane (den_perlr)");
    cudaStream_t st = (cudaStream_t)stream;
    pvdb_reset_launch_count();
    pvdb_prof_begin(st);
    MarchParams P;
    MarchOut O;
    fill_march(cfg, b, P, O);
    const int n_glob = cfg->n_rays_global > 0 ? cfg->n_rays_global : n_rays;
    const int warp_grid = pvdb_grid_for((int64_t)n_rays * 32, 256);
    // Side stream for work that is independent of the main chain (weight-image prep under the march; the density branch of
    // the backward under the rgbnet backward).  Not used while per-kernel profiling is on (serial, clean per-kernel times).
    struct Side { cudaStream_t s; cudaEvent_t fork, join, fork2, join2, fork3, join3; bool ok; };
    static thread_local Side side[16] = {};
    int dev = 0;
    PVDB_CUDA(cudaGetDevice(&dev));
    Side* sd = (dev >= 0 && dev < 16 && !pvdb_prof_active()) ? &side[dev] : nullptr;
    if (sd && !sd->ok) {
        PVDB_CUDA(cudaStreamCreateWithFlags(&sd->s, cudaStreamNonBlocking));
        PVDB_CUDA(cudaEventCreateWithFlags(&sd->fork, cudaEventDisableTiming));
        PVDB_CUDA(cudaEventCreateWithFlags(&sd->join, cudaEventDisableTiming));
        PVDB_CUDA(cudaEventCreateWithFlags(&sd->fork2, cudaEventDisableTiming));
        PVDB_CUDA(cudaEventCreateWithFlags(&sd->join2, cudaEventDisableTiming));
        PVDB_CUDA(cudaEventCreateWithFlags(&sd->fork3, cudaEventDisableTiming));
        PVDB_CUDA(cudaEventCreateWithFlags(&sd->join3, cudaEventDisableTiming));
        sd->ok = true;
    }

    // the fused sparse Adam: CTAs [0, leaf_blocks) walk the touched leaves, the rest update the rgbnet.  part 0 = both in one
    // launch, 1 = leaves only, 2 = rgbnet only (the leaf half can run on the side stream under the weight-gradient kernel)
    UpdateArgs U;
    U.tree = *b->tree;
    U.den = b->den; U.den_g = b->den_grad; U.den_m = b->den_m; U.den_v = b->den_v;
    U.k0 = b->k0; U.k0_g = b->k0_grad; U.k0_m = b->k0_m; U.k0_v = b->k0_v;
    U.net = b->net; U.net_g = b->net_grad; U.net_m = b->net_m; U.net_v = b->net_v;
    U.den_touched = b->den_touched; U.k0_touched = b->k0_touched; U.counters = b->counters;
    U.den_list = b->den_touched_list; U.k0_list = b->k0_touched_list;
    U.den_stepsz = cfg->den_stepsz; U.k0_stepsz = cfg->k0_stepsz;
    U.net_stepsize = pvdb_dense_adam_stepsize(cfg->net_lr, cfg->beta0, cfg->beta1, cfg->net_step);
    U.eps = cfg->eps; U.b0 = cfg->beta0; U.b1 = cfg->beta1;
    U.leaf_blocks = PVDB_SMS * 4;
    U.block_offset = 0;
    U.dp = PvdbDpNetWait{};
    U.dp_tiles_signal = nullptr; U.dp_tiles_epoch = 0; U.dp_world = 0; U.dp_err = nullptr;
    U.scalars = b->step_scalars;
    const int net_blocks = (PVDB_NET_N + 255) / 256;
    auto launch_update = [&](cudaStream_t s_, int part) -> int {
        UpdateArgs V = U;
        V.block_offset = part == 2 ? U.leaf_blocks : 0;
        const int grid = part == 0 ? U.leaf_blocks + net_blocks : part == 1 ? U.leaf_blocks : net_blocks;
        PVDB_CUDA(pvdb_launch_pdl(k_update_fused, dim3(grid), dim3(256), 0, s_, V));
        PVDB_LAUNCH_CHECK();
        return PVDB_OK;
    };
    {   // side-stream kernels run next to persistent tcgen05 CTAs: same shared-memory carve-out preference (see dp_exchange.cu)
        static bool attrs = false;
        if (!attrs) {
            cudaFuncSetAttribute(k_update_fused, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            cudaFuncSetAttribute(k_ray_bwd, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            cudaFuncSetAttribute(k_density_scatter, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            cudaFuncSetAttribute(k_stamp, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            attrs = true;
        }
    }
    bool update_done = false;
    // Fused data-parallel step: the emit kernel writes this rank's touched-leaf flags into its symmetric block, the union (with
    // the cross-GPU barrier that absorbs the ranks' skew) runs on the side stream under the rgbnet forward, pack / reduce-scatter /
    // unpack and the leaf Adam under the weight-gradient kernel, and the rgbnet gradients ride on the weight-gradient reduction
    // (peer stores) and the rgbnet Adam (wait + rank-ordered sum).  Needs the side stream and the tensor-core backward;
    // otherwise the stand-alone exchange runs after the backward.
    const bool dp_fused = peers && sd && cfg->use_tensor_cores && do_fwd && do_bwd && do_upd;
    if (dp_fused) {
        O.dp_flags = pvdb_dp_flags_ptr(peers, dp_step);
        const int words = (b->tree->n_leaf + 31) / 32;
        O.dp_words = words * 4 <= 40 * 1024 ? words : 0;      // larger trees flag straight into global memory
    }
    stamp(st, 0);

    bool scan_fused = false;
    if (do_fwd) {
        PVDB_CHECK_ARG(rays_o && rays_d && viewdirs, "null rays");
        if (sd) {
            PVDB_CUDA(cudaEventRecord(sd->fork2, st));
            PVDB_CUDA(cudaStreamWaitEvent(sd->s, sd->fork2, 0));
            int rc = pvdb_rgbnet_prepare(cfg, b, viewdirs, n_rays, sd->s);
            if (rc) return rc;
            PVDB_CUDA(cudaEventRecord(sd->join2, sd->s));
        }
        const int var = (P.run_skip ? 1 : 0) | (O.dp_flags ? 2 : 0);
        // PVDB_SCAN_IN_MARCH=1: the offset scans in the last CTA of the count pass instead of a kernel of their own.  Measured
        // slower (0.2037 vs 0.2008 ms per step, two runs each on one box): the 256-thread tail needs four dependent rounds where
        // k_scan_counts does one with 1024 threads, and that kernel's launch is already hidden by the programmatic dependent launch.
        static int scan_in_march = -1;
        if (scan_in_march < 0) { const char* e = getenv("PVDB_SCAN_IN_MARCH"); scan_in_march = e ? (atoi(e) != 0) : 0; }
        ScanTail S;
        S.oa = b->off_alpha; S.ok = b->off_keep; S.loss = b->loss; S.cap_alpha = b->cap_alpha; S.cap_keep = b->cap_keep;
        S.keep_touched = (phases & PVDB_PHASE_ACCUMULATE) ? 1 : 0;
        S.enabled = (scan_in_march && !pvdb_prof_active()) ? 1 : 0;       // per-kernel profiling keeps the scan visible as a kernel
        scan_fused = S.enabled != 0;
        if (cfg->parity_counts) {
            if (var & 1) k_march<0, true, 1><<<warp_grid, 256, 0, st>>>(P, O, rays_o, rays_d, n_rays, nullptr, S);
            else k_march<0, true, 0><<<warp_grid, 256, 0, st>>>(P, O, rays_o, rays_d, n_rays, nullptr, S);
            PVDB_CHECK_ARG(!(var & 2), "parity_counts is a single-GPU diagnostic (not available in the data-parallel step)");
        } else {
            static int resident = 0;   // CTAs of the count kernel that fit the device at once
            if (!resident) {
                int per_sm = 0;
                PVDB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_march<0, false, 0>, 256, 0));
                resident = PVDB_SMS * (per_sm > 0 ? per_sm : 1);
            }
            const dim3 grid(min(resident, warp_grid));
            int32_t* ticket = b->counters + CNT_RAY_TICKET;
            if (var == 0) PVDB_CUDA(pvdb_launch_pdl(k_march<0, false, 0>, grid, dim3(256), 0, st, P, O, rays_o, rays_d, n_rays, ticket, S));
            else if (var == 1) PVDB_CUDA(pvdb_launch_pdl(k_march<0, false, 1>, grid, dim3(256), 0, st, P, O, rays_o, rays_d, n_rays, ticket, S));
            else if (var == 2) PVDB_CUDA(pvdb_launch_pdl(k_march<0, false, 2>, grid, dim3(256), 0, st, P, O, rays_o, rays_d, n_rays, ticket, S));
            else PVDB_CUDA(pvdb_launch_pdl(k_march<0, false, 3>, grid, dim3(256), 0, st, P, O, rays_o, rays_d, n_rays, ticket, S));
        }
        PVDB_LAUNCH_CHECK();
        stamp(st, 8);
        pvdb_prof_mark("march_count", st);
        if (!scan_fused) {
            PVDB_CUDA(pvdb_launch_pdl(k_scan_counts, dim3(1), dim3(1024), 0, st, b->cnt_alpha, b->cnt_keep, b->off_alpha, b->off_keep, n_rays,
                                      b->counters, b->loss, b->cap_alpha, b->cap_keep, (int)((phases & PVDB_PHASE_ACCUMULATE) ? 1 : 0)));
            PVDB_LAUNCH_CHECK();
        }
        stamp(st, 9);
        pvdb_prof_mark("scan", st);
        PVDB_CHECK_ARG(!O.scratch || n_rays <= b->scratch_rays, "march_scratch holds fewer rays than this batch");
        if (O.scratch) {
            if (O.dp_flags) PVDB_CUDA(pvdb_launch_pdl(k_emit_scratch<2>, dim3(warp_grid), dim3(256), (size_t)O.dp_words * 4, st, P, O, rays_o, rays_d, n_rays));
            else PVDB_CUDA(pvdb_launch_pdl(k_emit_scratch<0>, dim3(warp_grid), dim3(256), 0, st, P, O, rays_o, rays_d, n_rays));
        } else if (O.dp_flags) {
            k_march<1, false, 2><<<warp_grid, 256, 0, st>>>(P, O, rays_o, rays_d, n_rays, nullptr, ScanTail{});
        } else {
            k_march<1, false, 0><<<warp_grid, 256, 0, st>>>(P, O, rays_o, rays_d, n_rays, nullptr, ScanTail{});
        }
        PVDB_LAUNCH_CHECK();
        pvdb_prof_mark("march_emit", st);
        stamp(st, 1);
        if (cfg->use_tensor_cores && !direct && pvdb_leaf_local_enabled(b)) {      // alternative k0 path: leaf buckets + staged gather (leaf_local.cu)
            int rcl = pvdb_leaf_local_forward(b, st);
            if (rcl) return rcl;
        }
        int rc;
        if (dp_fused) PVDB_CUDA(cudaEventRecord(sd->fork3, st));      // the emit kernel has written this rank's touched-leaf flags
        if (sd) {
            PVDB_CUDA(cudaStreamWaitEvent(st, sd->join2, 0));
        } else {
            rc = pvdb_rgbnet_prepare(cfg, b, viewdirs, n_rays, st);
            if (rc) return rc;
            pvdb_prof_mark("rgbnet_prep", st);      // per-kernel profiling runs without the side stream: keep the prep kernels out of the forward's time
        }
        rc = pvdb_rgbnet_forward(cfg, b, viewdirs, st);
        if (rc) return rc;
        stamp(st, 2);
        pvdb_prof_mark("rgbnet_fwd", st);
        if (dp_fused) {      // side stream, under the rgbnet forward (launched first, see the backward): union of the ranks' touched leaves
            PVDB_CUDA(cudaStreamWaitEvent(sd->s, sd->fork3, 0));
            stamp(sd->s, 10);
            rc = pvdb_dp_union_early(peers, b, dp_step, sd->s);
            if (rc) return rc;
            stamp(sd->s, 11);
            PVDB_CUDA(cudaEventRecord(sd->join3, sd->s));
        }
        CompositeParams C;
        C.off_keep = b->off_keep; C.k_sample = b->k_sample; C.s_weight = b->s_weight; C.k_rgb = b->k_rgb; C.k_gw = b->k_gw;
        C.alphainv_last = b->alphainv_last; C.target = target; C.rgb_marched = b->rgb_marched; C.grad_last = b->grad_last;
        C.loss = b->loss; C.cta_done = b->counters + CNT_CTA_DONE; C.bg = cfg->bg; C.w_main = cfg->weight_main; C.w_ent = cfg->weight_entropy_last;
        C.w_per = cfg->weight_rgbper; C.inv_N = 1.0f / (float)n_glob; C.cap_keep = b->cap_keep; C.do_backward = do_bwd ? 1 : 0;
        PVDB_CUDA(pvdb_launch_pdl(k_composite, dim3(pvdb_grid_for((int64_t)n_rays * CG, 256)), dim3(256), 0, st, C, n_rays));
        PVDB_LAUNCH_CHECK();
        pvdb_prof_mark("composite", st);
        stamp(st, 3);
    }
    if (do_bwd) {
        // The density branch of the backward (ray recurrence -> density scatter) depends only on the composite, not on the
        // rgbnet backward: it runs on a side stream underneath the two tcgen05 kernels (which leave most thread slots of
        // every SM free) and joins before the update.  Serial when per-kernel profiling is on.
        cudaStream_t sb = sd ? sd->s : st;
        if (sd) {
            PVDB_CUDA(cudaEventRecord(sd->fork, st));
            PVDB_CUDA(cudaStreamWaitEvent(sd->s, sd->fork, 0));
        } else {
            int rc = pvdb_rgbnet_backward(cfg, b, viewdirs, st);
            if (rc) return rc;
            pvdb_prof_mark(cfg->use_tensor_cores && !direct ? "rgbnet_bwd_wgrad" : "rgbnet_bwd", st);
        }
        k_ray_bwd<<<warp_grid, 256, 0, sb>>>(b->off_alpha, b->off_keep, b->s_alpha, b->s_T, b->s_weight, b->s_density,
                                                             b->k_gw, b->alphainv_last, b->grad_last, b->s_gden, n_rays,
                                                             cfg->fast_color_thres, cfg->act_shift, cfg->interval, b->cap_alpha,
                                                             b->cap_keep);
        PVDB_LAUNCH_CHECK();
        pvdb_prof_mark("ray_bwd", st);
        PVDB_CUDA(pvdb_launch_pdl(k_density_scatter, dim3(PVDB_SMS * 8), dim3(256), 0, sb, *b->tree, b->den_grad, (const float*)b->s_xyz,
                                  (const float*)b->s_gden, (const int32_t*)b->counters, b->den_touched, b->den_touched_list, b->counters,
                                  b->cap_alpha));
        PVDB_LAUNCH_CHECK();
        pvdb_prof_mark("density_scatter", st);
        if (sd && cfg->use_tensor_cores && !direct && (peers || do_upd)) {
            // The grid gradients are final once the activation-gradient kernel (main) and the density scatter (side) are done.
            // Everything that only needs them runs on the side stream UNDER the weight-gradient kernel: the NVLink tile
            // exchange of a data-parallel step, and the sparse Adam of the touched leaves.  After the weight-gradient kernel
            // only the 88 KB rgbnet exchange and the rgbnet Adam are left.
            // dp_fused: the union list and the touched flags (all set) must be in place before the k0 scatter of the
            // activation-gradient kernel looks at them
            if (dp_fused) PVDB_CUDA(cudaStreamWaitEvent(st, sd->join3, 0));
            int rc = pvdb_rgbnet_backward_act_tc(cfg, b, viewdirs, st);
            if (rc) return rc;
            stamp(st, 4);
            PVDB_CUDA(cudaEventRecord(sd->fork2, st));
            // The persistent weight-gradient kernel is launched FIRST: the side-stream kernels below become eligible at the same
            // moment (both wait for the activation-gradient kernel) and are small enough to fit next to its CTAs (<= 10 K
            // registers per CTA), but CTAs of theirs that got an SM first — spinning on a peer's signal — would keep that SM's
            // weight-gradient CTA from starting at all.
            PvdbDpNetPush push;
            if (dp_fused) push = pvdb_dp_net_push_args(peers, dp_step);
            // the kernel that sums the weight-gradient partials applies the rgbnet Adam to what it has summed (data parallel:
            // to the rank-ordered sum of what the ranks pushed)
            const bool adam_in_reduce = do_upd && (!peers || dp_fused);
            PvdbNetAdam nad = {};
            if (adam_in_reduce) {
                nad.on = 1; nad.net = b->net; nad.m = b->net_m; nad.v = b->net_v; nad.stepsize = U.net_stepsize; nad.b0 = cfg->beta0; nad.b1 = cfg->beta1;
                nad.eps = cfg->eps; nad.scalars = b->step_scalars;
            }
            PvdbDpNetWait nwait = {};
            if (dp_fused && adam_in_reduce) nwait = pvdb_dp_net_wait_args(peers, dp_step);
            rc = pvdb_rgbnet_backward_wgrad_tc(cfg, b, st, dp_fused ? &push : nullptr, adam_in_reduce ? &nad : nullptr,
                                               dp_fused && adam_in_reduce ? &nwait : nullptr);
            if (rc) return rc;
            stamp(st, 5);
            PVDB_CUDA(cudaStreamWaitEvent(sd->s, sd->fork2, 0));
            stamp(sd->s, 12);
            if (peers) {
                rc = dp_fused ? pvdb_dp_move_tiles(peers, b, dp_step, sd->s) : pvdb_dp_exchange_tiles(peers, b, dp_step, sd->s);
                if (rc) return rc;
                stamp(sd->s, 13);
            }
            if (do_upd) {
                if (dp_fused) {      // the leaf Adam itself waits for the owners' sums (no separate wait kernel)
                    const PvdbDpTilesWait w = pvdb_dp_tiles_wait_args(peers, dp_step);
                    U.dp_tiles_signal = w.signal; U.dp_tiles_epoch = w.epoch; U.dp_world = w.world; U.dp_err = w.err;
                }
                rc = launch_update(sd->s, 1);
                U.dp_tiles_signal = nullptr;
                if (rc) return rc;
                stamp(sd->s, 14);
            }
            PVDB_CUDA(cudaEventRecord(sd->join, sd->s));
            if (peers && !dp_fused) {
                rc = pvdb_dp_exchange_net(peers, b, dp_step, st);
                if (rc) return rc;
            }
            if (do_upd) {
                if (!adam_in_reduce) {
                    if (dp_fused) U.dp = pvdb_dp_net_wait_args(peers, dp_step);
                    rc = launch_update(st, 2);
                    U.dp = PvdbDpNetWait{};
                    if (rc) return rc;
                }
                update_done = true;
                stamp(st, 6);
            }
            PVDB_CUDA(cudaStreamWaitEvent(st, sd->join, 0));
            stamp(st, 7);
        } else {
            if (sd) {
                PVDB_CUDA(cudaEventRecord(sd->join, sd->s));
                int rc = pvdb_rgbnet_backward(cfg, b, viewdirs, st);
                if (rc) return rc;
                PVDB_CUDA(cudaStreamWaitEvent(st, sd->join, 0));
            }
            if (peers) {
                int rc = pvdb_dp_exchange_tiles(peers, b, dp_step, st);
                if (rc) return rc;
                rc = pvdb_dp_exchange_net(peers, b, dp_step, st);
                if (rc) return rc;
            }
        }
    }
    if (do_upd && direct) {
        // Coarse stage (masked_adam.py:56-68 with configs/default.py:48, 64): full-grid Adam in the configured stepmodes — 0 updates every
        // voxel whatever its gradient, 2 scales the step with the per-voxel lr —, then zero_grad of both planes (run.py:549-550 does it
        // at the start of the next iteration) and the touched-leaf bookkeeping back to empty.
        const pvdb_tree* t = b->tree;
        int rc = pvdb_adam_step(t, b->den, b->den_grad, b->den_m, b->den_v, 1, cfg->den_mode, cfg->den_stepsz, cfg->eps, cfg->beta0, cfg->beta1,
                                cfg->den_mode == 2 ? b->den_perlr : nullptr, st);
        if (rc) return rc;
        rc = pvdb_adam_step(t, b->k0, b->k0_grad, b->k0_m, b->k0_v, 3, cfg->k0_mode, cfg->k0_stepsz, cfg->eps, cfg->beta0, cfg->beta1, nullptr, st);
        if (rc) return rc;
        rc = pvdb_zero_grad(t, b->den_grad, 1, st);
        if (rc) return rc;
        rc = pvdb_zero_grad(t, b->k0_grad, 3, st);
        if (rc) return rc;
        const size_t nl = (size_t)(t->n_leaf > 0 ? t->n_leaf : 1) * sizeof(int32_t);
        PVDB_CUDA(cudaMemsetAsync(b->den_touched, 0, nl, st));
        PVDB_CUDA(cudaMemsetAsync(b->k0_touched, 0, nl, st));
        pvdb_prof_mark("update_dense", st);
        return PVDB_OK;
    }
    if (do_upd && !update_done) {
        if (!do_bwd && !(phases & PVDB_PHASE_LISTS_READY)) {   // gradients (and flags) came from elsewhere, e.g. a data-parallel all-reduce: rebuild the lists
            const int n_leaf = b->tree->n_leaf;
            PVDB_CUDA(cudaMemsetAsync(b->counters + CNT_N_TOUCHED_DEN, 0, sizeof(int32_t), st));
            PVDB_CUDA(cudaMemsetAsync(b->counters + CNT_N_TOUCHED_K0, 0, sizeof(int32_t), st));
            k_touched_compact<<<pvdb_grid_for(n_leaf, 256), 256, 0, st>>>(b->den_touched, n_leaf, b->den_touched_list,
                                                                          b->counters + CNT_N_TOUCHED_DEN);
            PVDB_LAUNCH_CHECK();
            k_touched_compact<<<pvdb_grid_for(n_leaf, 256), 256, 0, st>>>(b->k0_touched, n_leaf, b->k0_touched_list,
                                                                          b->counters + CNT_N_TOUCHED_K0);
            PVDB_LAUNCH_CHECK();
        }
        int rc = launch_update(st, 0);
        if (rc) return rc;
        pvdb_prof_mark("update_fused", st);
    }
    return PVDB_OK;
}

extern "C" int pvdb_train_step(const pvdb_train_cfg* cfg, const pvdb_train_bufs* b, const float* rays_o, const float* rays_d,
                               const float* viewdirs, const float* target, int n_rays, int phases, void* stream) {
    return train_step_impl(cfg, b, rays_o, rays_d, viewdirs, target, n_rays, phases, stream, nullptr, 0);
}

extern "C" int pvdb_train_step_dp(const pvdb_train_cfg* cfg, const pvdb_train_bufs* b, const pvdb_dp_peers* peers, uint32_t dp_step,
                                  const float* rays_o, const float* rays_d, const float* viewdirs, const float* target, int n_rays,
                                  void* stream) {
    PVDB_CHECK_ARG(peers, "null peers");
    return train_step_impl(cfg, b, rays_o, rays_d, viewdirs, target, n_rays, PVDB_PHASE_FORWARD | PVDB_PHASE_BACKWARD | PVDB_PHASE_UPDATE,
                           stream, peers, dp_step);
}
