// rgbnet_tc_bwd.cu — rgbnet backward on tcgen05 (3xTF32), two kernels (details at each section):
//
//  B1 k_rgbnet_bwd_act_tc   activation gradients, 128-sample tiles, warp specialised like the forward:
//                                      dH1 = (g_logit . W2) * [h1>0]  (K = 3, CUDA cores, in registers)
//                                      dH0 = (dH1 . W1)   * [h0>0]  (tcgen05, A = dH1 in TMEM, B = W1^T image in smem)
//                                      dX  =  dH0 . W0[:, :12]       (tcgen05, N = 16) -> k0 gradient scatter
//                           dH1, dH0 are also written to HBM (chunk-major) for B2.
//  B2 k_rgbnet_bwd_wgrad_tc weight gradients = sums over samples of outer products:  dW1 = dH1^T H0, dW0 = dH0^T X
//                           (tcgen05 SS, operands streamed by TMA into K-major SWIZZLE_64B tiles, bias gradients as a
//                           ones row of B), dW2 = G^T H1 on the CUDA cores.  Accumulators live in TMEM for the CTA's
//                           lifetime; per-CTA partial sums are stored and added by k_wgrad_reduce.
// Reference semantics: autograd through nn.Sequential(Linear, ReLU, Linear, ReLU, Linear) (dvgo.py:99-107) and
// QueryVerticalInVDB.backward -> color_backward (grid.py:53-60, colorvdb.cu:130-175).
#include <cuda.h>
#include <stdlib.h>
#include <string.h>
#include "common.cuh"
#include "ray_math.cuh"
#include "rgbnet.cuh"
#include "tc_ptx.cuh"
#include "dp_exchange.cuh"
#include "leaf_local.cuh"

namespace {

constexpr int WD = PVDB_NET_W;
constexpr int CNT_M_KEEP = 1;
constexpr uint32_t COL_AHI = 0, COL_ALO = 128;

__device__ __forceinline__ void red_add(float* addr, float v) { asm volatile("red.global.add.f32 [%0], %1;" ::"l"(addr), "f"(v) : "memory"); }
__device__ __forceinline__ void red_add4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__device__ __forceinline__ void store_a_row32(uint32_t tmem_lane, int c0, const float* v) {
#pragma unroll
    for (int c = 0; c < 32; c += 8) {
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) split_tf32(v[c + i], hi[i], lo[i]);
        tmem_st8(tmem_lane + COL_AHI + c0 + c, hi);
        tmem_st8(tmem_lane + COL_ALO + c0 + c, lo);
    }
}
// ---------------------------------------------------------------------------------------------- B1
// Same warp-specialised structure as the forward (rgbnet_tc.cu): 8 lane warps (thread = sample lane x column half), one
// issuer warp; the weight image (W1^T, W0[:, :12]^T as tf32 hi/lo in the canonical K-major layout + W2 as fp32) is built
// once per step (prep_bwd_image, run by the forward's prep kernel) and pulled into shared memory with one bulk async copy per CTA.
//   step 1  dH1 = (g . W2) * [h1 > 0] on the CUDA cores, 32 columns at a time, into TMEM as the A operand (B2 recomputes it:
//           it is never stored); the issuer starts the matching k-steps of dH0 = dH1 . W1 (into D0) chunk by chunk
//   step 2  dH0 = D0 * [h0 > 0]: both chunks back into TMEM as A first — the issuer follows with dX = dH0 . W0[:, :12]
//           (into D1) —, then to HBM (chunk-major, for B2) through the per-warp staging regions
//   step 3  dX -> k0 gradient scatter (colorvdb.cu:130-160) with the corner record ids the march saved; it is deferred
//           until the NEXT tile's step 1 has been issued, so the scatter atomics run under that tile's MMAs
constexpr int B1_BAR = B1_IMG;                          // W, D0, D1, A1_RDY[4], A0_RDY[4]; tmem slot at +88
constexpr int B1B_W = 0, B1B_D0 = 8, B1B_D1 = 16, B1B_A1 = 24, B1B_A0 = 56, B1_TMEM_SLOT = 88;
constexpr int B1_STAGE = B1_BAR + 96;                   // dH0 store staging, one region per lane warp (stage_store32)
constexpr int B1_TOTAL = B1_STAGE + 8 * STAGE_WARP_BYTES;
static_assert(B1_STAGE % 16 == 0 && B1_TOTAL <= 227 * 1024, "activation-gradient smem");
constexpr uint32_t COL_D0 = 256, COL_D1 = 384;
constexpr int B1_THREADS = 288, B1_LANES = 256;
static_assert(B1_IMG % 16 == 0, "bulk copy granularity");

struct BwdActArgs {
    const unsigned char* img;
    const float *k_glogit, *k_xyz;
    const uint32_t* k_mask;
    const int32_t* k_corner;
    float* k_dh0;
    float* k0_grad; int32_t* k0_touched; int32_t* k0_touched_list; int32_t* counters_w;
    const int32_t* counters; int64_t cap_keep;
    float* k_dx;   // non-null: store dL/dx of every kept sample here instead of scattering it (leaf_local.cu accumulates it per leaf)
};

// k-steps [ks0, ks1) of D[128 x N] (+)= A(TMEM)[128 x 128] * B(smem, K-major)[N x 128]^T, 3xTF32
__device__ __forceinline__ void issue_b1(uint32_t tmem, uint32_t d_col, uint32_t smem_hi, uint32_t smem_lo, int N, int ks0, int ks1, bool zero_first) {
    const uint32_t idesc = make_idesc(N);
    const uint64_t bhi = make_desc(smem_hi, WD), blo = make_desc(smem_lo, WD);
    uint32_t acc = zero_first ? 0u : 1u;
    for (int ks = ks0; ks < ks1; ++ks) {
        const uint64_t adv = (uint64_t)(ks * 256) >> 4;
        const uint32_t ahi = tmem + COL_AHI + ks * 8, alo = tmem + COL_ALO + ks * 8;
        umma_tf32_ts(tmem + d_col, ahi, bhi + adv, idesc, acc);
        acc = 1;
        umma_tf32_ts(tmem + d_col, alo, bhi + adv, idesc, 1);
        umma_tf32_ts(tmem + d_col, ahi, blo + adv, idesc, 1);
    }
}

__global__ void __launch_bounds__(B1_THREADS, 1) k_rgbnet_bwd_act_tc(BwdActArgs A) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, warp = tid >> 5;
    const uint32_t sbase = smem_u32(smem);
    const uint32_t bars = sbase + B1_BAR;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + B1_BAR + B1_TMEM_SLOT);
    const float* sW2 = reinterpret_cast<const float*>(smem + B1_W2);
    if (tid == 0) {
        mbar_init(bars + B1B_W, 1);
        mbar_init(bars + B1B_D0, 1);
        mbar_init(bars + B1B_D1, 1);
        for (int c = 0; c < 4; ++c) { mbar_init(bars + B1B_A1 + 8 * c, 128); mbar_init(bars + B1B_A0 + 8 * c, 128); }
        fence_async_smem();
        mbar_expect_tx(bars + B1B_W, B1_IMG);
        constexpr int CH = 32768;
        for (int o = 0; o < B1_IMG; o += CH) bulk_g2s(sbase + o, A.img + o, min(CH, B1_IMG - o), bars + B1B_W);
    }
    if (warp == 0) tmem_alloc(sbase + B1_BAR + B1_TMEM_SLOT, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    pvdb_pdl_wait();   // the prologue above (barriers, TMEM, weight image in flight) does not depend on the kernel before
    const int64_t M = min((int64_t)A.counters[CNT_M_KEEP], A.cap_keep);
    const int64_t n_tiles = (M + TM - 1) / TM;

    if (warp == 8) {
        // ---------------- issuer
        if (tid == 256 && (int64_t)blockIdx.x < n_tiles) {
            mbar_wait(bars + B1B_W, 0);
            int i = 0;
            for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++i) {
                for (int cc = 0; cc < 4; ++cc) {
                    const int c = (cc >> 1) | ((cc & 1) << 1);      // 0, 2, 1, 3: the order the two column halves deliver their chunks
                    mbar_wait(bars + B1B_A1 + 8 * c, i & 1);
                    tc_fence_after();
                    issue_b1(tmem, COL_D0, sbase + B1_W1HI, sbase + B1_W1LO, WD, c * 4, c * 4 + 4, cc == 0);
                }
                umma_commit(bars + B1B_D0);
                for (int cc = 0; cc < 4; ++cc) {
                    const int c = (cc >> 1) | ((cc & 1) << 1);
                    mbar_wait(bars + B1B_A0 + 8 * c, i & 1);
                    tc_fence_after();
                    issue_b1(tmem, COL_D1, sbase + B1_W0HI, sbase + B1_W0LO, 16, c * 4, c * 4 + 4, cc == 0);
                }
                umma_commit(bars + B1B_D1);
            }
        }
    } else if ((int64_t)blockIdx.x < n_tiles) {
        // ---------------- lane warps
        const int grp = warp >> 2;                               // column half (and corner half for the scatter)
        const int lane_s = (warp & 3) * 32 + (tid & 31);         // sample within the tile = TMEM lane
        const uint32_t lane_addr = tmem + ((uint32_t)((warp & 3) * 32) << 16);
        const int cb = grp * 64;
        // per-sample inputs of this thread's half: logit gradient, 2 + 2 mask words, 4 corner ids, position
        struct In { float g0, g1, g2, px, py, pz; uint32_t m1a, m1b, m0a, m0b; int4 rec; };
        auto fetch = [&](int64_t tile) {
            In v;
            v.g0 = v.g1 = v.g2 = v.px = v.py = v.pz = 0.f;
            v.m1a = v.m1b = v.m0a = v.m0b = 0u;
            v.rec = make_int4(-1, -1, -1, -1);
            const int64_t s = tile * TM + lane_s;
            if (tile < n_tiles && s < M) {
                v.g0 = __ldg(A.k_glogit + s * 3); v.g1 = __ldg(A.k_glogit + s * 3 + 1); v.g2 = __ldg(A.k_glogit + s * 3 + 2);
                v.px = __ldg(A.k_xyz + s * 3); v.py = __ldg(A.k_xyz + s * 3 + 1); v.pz = __ldg(A.k_xyz + s * 3 + 2);
                const uint32_t* mk = A.k_mask + (s >> 7) * (8 * 128) + (s & 127);   // [tile][8][128]: words 0-3 h0, 4-7 h1
                v.m0a = __ldg(mk + (grp * 2) * 128); v.m0b = __ldg(mk + (grp * 2 + 1) * 128);
                v.m1a = __ldg(mk + (4 + grp * 2) * 128); v.m1b = __ldg(mk + (5 + grp * 2) * 128);
                v.rec = __ldg(reinterpret_cast<const int4*>(A.k_corner + s * 8) + grp);
            }
            return v;
        };
        // step 1 of a tile: dH1 for this thread's 64 columns -> HBM + A operand, chunk by chunk
        auto step1 = [&](int64_t tile, const In& in) {
            const int64_t s = tile * TM + lane_s;
#pragma unroll
            for (int jj = 0; jj < 2; ++jj) {
                const int c = cb + jj * 32;
                const uint32_t mw = jj == 0 ? in.m1a : in.m1b;
                float d[32];
#pragma unroll
                for (int q = 0; q < 32; q += 4) {
                    const float4 wa = *reinterpret_cast<const float4*>(sW2 + c + q);
                    const float4 wb = *reinterpret_cast<const float4*>(sW2 + WD + c + q);
                    const float4 wc = *reinterpret_cast<const float4*>(sW2 + 2 * WD + c + q);
                    d[q] = (mw >> q) & 1u ? fmaf(in.g2, wc.x, fmaf(in.g1, wb.x, in.g0 * wa.x)) : 0.f;
                    d[q + 1] = (mw >> (q + 1)) & 1u ? fmaf(in.g2, wc.y, fmaf(in.g1, wb.y, in.g0 * wa.y)) : 0.f;
                    d[q + 2] = (mw >> (q + 2)) & 1u ? fmaf(in.g2, wc.z, fmaf(in.g1, wb.z, in.g0 * wa.z)) : 0.f;
                    d[q + 3] = (mw >> (q + 3)) & 1u ? fmaf(in.g2, wc.w, fmaf(in.g1, wb.w, in.g0 * wa.w)) : 0.f;
                }
                store_a_row32(lane_addr, c, d);
                tmem_st_wait();
                tc_fence_before();
                mbar_arrive(bars + B1B_A1 + 8 * (c >> 5));
                // dH1 is not stored: the weight-gradient pass recomputes it from g, W2 and the mask bits
            }
        };
        unsigned char* stage = smem + B1_STAGE + warp * STAGE_WARP_BYTES;
        uint32_t par0 = 0, par1 = 0;
        In cur = fetch(blockIdx.x);
        mbar_wait(bars + B1B_W, 0);
        step1(blockIdx.x, cur);
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const int64_t s = tile * TM + lane_s;
            const bool valid = s < M;
            const In nxt = fetch(tile + gridDim.x);
            // ---- step 2: dH0 = D0 masked by h0 > 0
            mbar_wait(bars + B1B_D0, par0); par0 ^= 1;
            tc_fence_after();
            {
                uint32_t r[2][32];
                tmem_ld32(lane_addr + COL_D0 + cb, r[0]);
                tmem_ld32(lane_addr + COL_D0 + cb + 32, r[1]);
                tmem_ld_wait();
#pragma unroll
                for (int jj = 0; jj < 2; ++jj) {      // what dX waits for first: both chunks of the A operand
                    const int c = cb + jj * 32;
                    const uint32_t mw = jj == 0 ? cur.m0a : cur.m0b;
                    float d[32];
#pragma unroll
                    for (int q = 0; q < 32; ++q) d[q] = (mw >> q) & 1u ? __uint_as_float(r[jj][q]) : 0.f;
                    store_a_row32(lane_addr, c, d);
                    tmem_st_wait();
                    tc_fence_before();
                    mbar_arrive(bars + B1B_A0 + 8 * (c >> 5));
                }
#pragma unroll
                for (int jj = 0; jj < 2; ++jj) {      // then dH0 for the weight-gradient pass, underneath the dX MMAs (rows past M: masks are 0)
                    const int c = cb + jj * 32;
                    const uint32_t mw = jj == 0 ? cur.m0a : cur.m0b;
                    float d[32];
#pragma unroll
                    for (int q = 0; q < 32; ++q) d[q] = (mw >> q) & 1u ? __uint_as_float(r[jj][q]) : 0.f;
                    stage_store32(stage, A.k_dh0, s - (tid & 31), c, WD, d, true);
                }
            }
            // every lane has drained D0 before any lane lets the issuer start the next tile's dH0 MMAs
            asm volatile("bar.sync 1, 256;" ::: "memory");
            // ---- step 3a: dX row of this sample out of D1
            mbar_wait(bars + B1B_D1, par1); par1 ^= 1;
            tc_fence_after();
            uint32_t r[16];
            tmem_ld16(lane_addr + COL_D1, r);
            tmem_ld_wait();
            // ---- next tile's step 1 goes first: its MMAs run under the scatter below (A is free: both MMAs of this tile are done)
            if (tile + gridDim.x < n_tiles) step1(tile + gridDim.x, nxt);
            // ---- step 3b: k0 gradient scatter, 4 corners per thread (group 0: corners 0-3, group 1: 4-7), 3 x red.v4 per corner
            if (valid && A.k_dx) {
                if (grp == 0) {      // both column halves hold the same dX row
                    float4* dst = reinterpret_cast<float4*>(A.k_dx + s * 12);
                    dst[0] = make_float4(__uint_as_float(r[0]), __uint_as_float(r[1]), __uint_as_float(r[2]), __uint_as_float(r[3]));
                    dst[1] = make_float4(__uint_as_float(r[4]), __uint_as_float(r[5]), __uint_as_float(r[6]), __uint_as_float(r[7]));
                    dst[2] = make_float4(__uint_as_float(r[8]), __uint_as_float(r[9]), __uint_as_float(r[10]), __uint_as_float(r[11]));
                }
            } else if (valid) {
                PvdbTri tri;
                tri.set(cur.px, cur.py, cur.pz);
                const int rec[4] = {cur.rec.x, cur.rec.y, cur.rec.z, cur.rec.w};
#pragma unroll
                for (int qq = 0; qq < 4; ++qq) {
                    if (rec[qq] < 0) continue;
                    const int q = grp * 4 + qq;
                    const int dx = PVDB_CORNER[q][0], dy = PVDB_CORNER[q][1], dz = PVDB_CORNER[q][2];
                    const float sc = __fmul_rn(__fmul_rn(tri.f(0, dx), tri.f(1, dy)), tri.f(2, dz));
                    float* dst = A.k0_grad + (size_t)rec[qq] * 12;
#pragma unroll
                    for (int c4 = 0; c4 < 3; ++c4)
                        red_add4(dst + c4 * 4, __fmul_rn(__uint_as_float(r[c4 * 4]), sc), __fmul_rn(__uint_as_float(r[c4 * 4 + 1]), sc),
                                 __fmul_rn(__uint_as_float(r[c4 * 4 + 2]), sc), __fmul_rn(__uint_as_float(r[c4 * 4 + 3]), sc));
                    pvdb_touch_leaf(A.k0_touched, A.k0_touched_list, A.counters_w + 4, rec[qq] >> 9);
                }
            }
            cur = nxt;
        }
        if ((tid & 31) == 0) bulk_wait0();      // this warp's dH0 blocks are in global memory before the kernel ends
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

// ---------------------------------------------------------------------------------------------- B2
// Weight gradients: contraction over SAMPLES, so both operands of every GEMM are activations transposed.
//   dW1[j][i] = sum_s dH1[s][j] H0[s][i]   (M = 128, N = 128 + bias column, tcgen05)
//   dW0[j][k] = sum_s dH0[s][j] X[s][k]    (M = 128, N = 39 + bias column, tcgen05)
//   dW2[c][i] = sum_s G[s][c]  H1[s][i]    (3 x 128: CUDA cores, an N = 16 MMA costs as much tensor time as N = 128)
// MN-major tf32 operands only exist in the 128B_BASE32B swizzled layout (CUTLASS sm100_common.inl), so the operands are
// K-major instead: the activations live in HBM chunk-major [chunk of 16 samples][feature][16 samples] (act_off), i.e. a
// chunk of one tensor is a row-major [features][64 bytes] block — exactly the K-major SWIZZLE_64B UMMA layout (8-row x
// 64-byte atoms) up to the XOR of the 16-byte column with row bits, which TMA applies on the way in.
// One elected thread issues four TMA tensor loads + five bulk copies per chunk; nothing else touches the load path.
// Measured on B200 (scratch/tf32_probe.cu): kind::tf32 TRUNCATES the low 13 mantissa bits, so a raw fp32 tile in shared memory
// is its own "hi" operand and only lo = x - trunc(x) has to be computed (elementwise, layout-agnostic).
//
// The kernel is bound by shared-memory bandwidth (round 2, scratch/wg_timing.py: 1.8 k cycles per 16-sample chunk with ~180 KB
// of shared-memory traffic per chunk at 128 B/clk; four more converter warps made it slower), and 84 KB of the 180 were the
// tensor core reading its operands — every operand three times for 3xTF32.  So the A operands (dH1^T, dH0^T: 128 features x
// 16 samples) do not live in shared memory at all: a converter thread owns one feature row = one TMEM lane, computes (dH1) or
// reads (dH0) its 16 samples and writes hi / lo straight into TMEM (A-from-TMEM form of tcgen05.mma, two alternating sets of
// 64 columns).  Shared-memory traffic per chunk: ~110 KB.
//
// Warp-specialised streaming pipeline, 6 stages of 16 samples:
//   warps 0-3   converters A1: dH1 = [h1 > 0] (g . W2) for feature row j, from the 16 x 3 logit gradients, W2 and the mask words
//               (dH1 never exists in HBM) -> TMEM; their share of the lo tiles of the B operands
//   warps 4-7   converters A0: dH0 row j from the landed tile -> TMEM; dW2 / db2 partial sums on the CUDA cores; lo tiles
//   warp 8      loader: wait free[st]; expect_tx + TMA loads onto full[st]   (cp.async fallback: all 32 lanes copy)
//   warp 9      issuer: wait conv[st]; 12 MMAs (2 k-steps x 3 passes x 2 GEMMs); tcgen05.commit -> free[st]
// Accumulators stay in TMEM for the CTA's lifetime; the per-CTA partial sums leave through a shared-memory tile and bulk stores.
constexpr int KC = 16;                                   // samples per chunk = 2 k-steps of 8
constexpr int N1 = 144, N0 = 48;                         // padded N (128 + bias row, 39 + bias row)
constexpr int ROWB = KC * 4;                             // 64-byte rows
constexpr int S_B1 = 0;                                  // [H0^T ; 1 ; 0..] [144][KC] raw (= hi)
constexpr int S_A0 = S_B1 + N1 * ROWB;                   // dH0^T [128][KC] landing tile (read by the A0 converters)
constexpr int S_B0 = S_A0 + WD * ROWB;                   // [X^T(39) ; 1 ; 0..] [48][KC]  (k_x row 39 holds the ones)
constexpr int S_H1 = S_B0 + N0 * ROWB;                   // H1^T [128][KC] raw, CUDA-core dW2
constexpr int S_G = S_H1 + WD * ROWB;                    // logit gradients [KC][3] (192 B), then at +256 the h1 ReLU mask words [4][KC]
constexpr int S_M = S_G + 256;
constexpr int STAGE = S_G + 512;                         // one raw stage (what TMA fills)
constexpr int NSTAGE = 6;
// lo tiles of the B operands only live between the converters and the MMAs of one chunk: two sets, alternating by chunk parity
constexpr int LO_B1 = 0, LO_B0 = LO_B1 + N1 * ROWB;
constexpr int LOSET = LO_B0 + N0 * ROWB;
constexpr int S_LOSETS = NSTAGE * STAGE;
constexpr int B2_BAR = S_LOSETS + 2 * LOSET;             // full[6], conv[6], free[6] mbarriers + tmem slot
constexpr int BAR_FULL2 = 0, BAR_CONV2 = 64, BAR_FREE2 = 128, TMEM_SLOT2 = 192;
constexpr int B2_TOTAL = B2_BAR + 256;
// TMEM: the two accumulators, then two sets of A operand columns (chunk parity): dH1 hi, dH1 lo, dH0 hi, dH0 lo, 16 columns each
constexpr uint32_t ACC1 = 0, ACC0 = 144, ASET = 192, ASET_COLS = 64, A1HI = 0, A1LO = 16, A0HI = 32, A0LO = 48;
constexpr uint32_t TX_BYTES = 3 * WD * ROWB + 40 * ROWB + KC * 3 * 4 + 4 * KC * 4;   // bytes landing per stage
constexpr int N_LO_ITEMS = (WD + 40) * 4;                // 16-byte items of the B tiles that get a lo twin: B1 rows 0..127, B0 rows 0..39
constexpr int DR_LD = 132;                               // drain staging row stride (floats): 16-byte aligned, float4 stores 4-way = optimal
static_assert(S_A0 % 512 == 0 && S_B0 % 512 == 0 && S_H1 % 512 == 0 && LO_B0 % 512 == 0 && STAGE % 512 == 0 && LOSET % 512 == 0,
              "SW64 tiles must start on the 512-byte swizzle period");
static_assert(B2_TOTAL <= 227 * 1024, "wgrad smem");
static_assert((128 * DR_LD + 128 * PVDB_NET_DIN) * 4 <= NSTAGE * STAGE, "the drain staging reuses the stage ring");

// byte offset of (row f, 16-byte column k4) inside a K-major SWIZZLE_64B tile with 64-byte rows
__device__ __forceinline__ uint32_t sw64_off(int f, int k4) { return (uint32_t)(f * ROWB + ((k4 ^ ((f >> 1) & 3)) << 4)); }
// UMMA shared-memory descriptor, K-major SWIZZLE_64B: LBO = 1 (unused), SBO = 512 B between 8-row groups, layout type 4
__device__ __forceinline__ uint64_t make_desc_sw64(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3fff) | (1ull << 16) | ((uint64_t)(512 >> 4) << 32) | (1ull << 46) | (4ull << 61);
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_arrive(uint32_t bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
// TMA tensor tile load (3-D map: sample-in-chunk, feature, chunk), completion in bytes on `bar`
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
                 "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
// 16 values of one feature row -> hi / lo columns [col, col + 16) of this thread's TMEM lane
__device__ __forceinline__ void store_a_cols16(uint32_t lane_addr, uint32_t col_hi, uint32_t col_lo, const float* v) {
#pragma unroll
    for (int c = 0; c < 16; c += 8) {
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) split_tf32(v[c + i], hi[i], lo[i]);
        tmem_st8(lane_addr + col_hi + c, hi);
        tmem_st8(lane_addr + col_lo + c, lo);
    }
}

struct BwdWgradArgs {
    const float *k_h0, *k_dh0, *k_x, *k_h1, *k_glogit;
    const uint32_t* k_mask;
    const float* w2;   // [3][128]
    float* partial;   // [gridDim.x][PART_LD] per-CTA weight-gradient partial sums (net_grad layout)
    const int32_t* counters; int64_t cap_keep;
    int use_tma;
    int l2_discard;   // after a chunk has landed in shared memory, drop its (dead) global lines from L2 without write-back
};
constexpr int PART_LD = (PVDB_NET_N + 31) & ~31;
struct WgradMaps { CUtensorMap h0, dh0, h1, x; };

#ifdef PVDB_TC_TIMING
__device__ long long g_wg_t[PVDB_SMS][16];
#define WG_T(i, v) g_wg_t[blockIdx.x][i] = (v)
#else
#define WG_T(i, v) do { } while (0)
#endif

constexpr int B2_THREADS = 320;
constexpr int B2_CONV = 256;
__global__ void __launch_bounds__(B2_THREADS, 1) k_rgbnet_bwd_wgrad_tc(BwdWgradArgs A, const __grid_constant__ WgradMaps maps) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const int tid = threadIdx.x, warp = tid >> 5;
    const long long t_start = clock64();
    long long t_wait = 0, t_lo = 0;
    const uint32_t sbase = smem_u32(smem);
    const uint32_t bars = sbase + B2_BAR;
    const uint32_t bar_full = bars + BAR_FULL2, bar_conv = bars + BAR_CONV2, bar_free = bars + BAR_FREE2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + B2_BAR + TMEM_SLOT2);
    // Constant parts of the operand tiles (nothing here depends on the kernel before: it runs ahead of the PDL wait).  Every
    // other byte of a stage / lo set is rewritten per chunk (TMA fills whole tiles, the producers zero rows past M):
    //   B1 rows 128..143: the bias "ones" row (k = all samples; rows of A are zero past M, so padding adds nothing), then zeros
    //   B0 rows 40..47: zeros (row 39 of k_x holds the ones); the same rows of the lo sets: zeros (lo(1) = 0)
    for (int e = tid; e < NSTAGE * 24 * KC; e += B2_THREADS) {
        const int st = e / (24 * KC), r = (e / KC) % 24, k = e % KC;
        unsigned char* base = smem + st * STAGE;
        if (r < 16) *reinterpret_cast<float*>(base + S_B1 + (128 + r) * ROWB + k * 4) = r == 0 ? 1.0f : 0.f;
        else *reinterpret_cast<float*>(base + S_B0 + (40 + r - 16) * ROWB + k * 4) = 0.f;
    }
    for (int e = tid; e < 2 * 24 * KC; e += B2_THREADS) {
        const int ls = e / (24 * KC), r = (e / KC) % 24, k = e % KC;
        unsigned char* base = smem + S_LOSETS + ls * LOSET;
        if (r < 16) *reinterpret_cast<float*>(base + LO_B1 + (128 + r) * ROWB + k * 4) = 0.f;
        else *reinterpret_cast<float*>(base + LO_B0 + (40 + r - 16) * ROWB + k * 4) = 0.f;
    }
    if (tid == 0)
        for (int st = 0; st < NSTAGE; ++st) {
            mbar_init(bar_full + 8 * st, A.use_tma ? 1 : 32);
            mbar_init(bar_conv + 8 * st, B2_CONV);
            mbar_init(bar_free + 8 * st, 1);
        }
    if (warp == 0) tmem_alloc(sbase + B2_BAR + TMEM_SLOT2, 512);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    pvdb_pdl_wait();
    const int64_t M = min((int64_t)A.counters[CNT_M_KEEP], A.cap_keep);
    const int64_t n_chunks = (M + KC - 1) / KC;
    // contiguous chunk range of this CTA (sequential HBM addresses per array)
    const int64_t per = (n_chunks + gridDim.x - 1) / gridDim.x;
    const int64_t ch_lo = min(n_chunks, (int64_t)blockIdx.x * per), ch_hi = min(n_chunks, ch_lo + per);
    const int n_it = (int)(ch_hi - ch_lo);
    float w2acc[3] = {0.f, 0.f, 0.f};
    float gs[3] = {0.f, 0.f, 0.f};
    if (tid == 0) { WG_T(0, t_start); WG_T(1, clock64()); WG_T(2, (long long)n_it); }

    if (warp == 9) {
        // ---------------- issuer
        if (tid == 288) {
            const uint32_t id1 = make_idesc(N1), id0 = make_idesc(N0);
            for (int it = 0; it < n_it; ++it) {
                const int st = it % NSTAGE;
                { const long long t0 = clock64(); mbar_wait(bar_conv + 8 * st, (it / NSTAGE) & 1); t_wait += clock64() - t0; }
                tc_fence_after();
                const uint32_t base = sbase + st * STAGE, lob = sbase + S_LOSETS + (it & 1) * LOSET;
                const uint32_t aset = tmem + ASET + (uint32_t)(it & 1) * ASET_COLS;
                const uint32_t acc = it > 0 ? 1u : 0u;
                const uint64_t b1h = make_desc_sw64(base + S_B1), b1l = make_desc_sw64(lob + LO_B1);
                const uint64_t b0h = make_desc_sw64(base + S_B0), b0l = make_desc_sw64(lob + LO_B0);
#pragma unroll
                for (int ks = 0; ks < KC / 8; ++ks) {
                    const uint64_t adv = (uint64_t)(ks * 32) >> 4;   // 8 tf32 = 32 bytes along the swizzled row
                    umma_tf32_ts(tmem + ACC1, aset + A1HI + ks * 8, b1h + adv, id1, ks == 0 ? acc : 1u);
                    umma_tf32_ts(tmem + ACC1, aset + A1LO + ks * 8, b1h + adv, id1, 1u);
                    umma_tf32_ts(tmem + ACC1, aset + A1HI + ks * 8, b1l + adv, id1, 1u);
                    umma_tf32_ts(tmem + ACC0, aset + A0HI + ks * 8, b0h + adv, id0, ks == 0 ? acc : 1u);
                    umma_tf32_ts(tmem + ACC0, aset + A0LO + ks * 8, b0h + adv, id0, 1u);
                    umma_tf32_ts(tmem + ACC0, aset + A0HI + ks * 8, b0l + adv, id0, 1u);
                }
                umma_commit(bar_free + 8 * st);
            }
            WG_T(3, t_wait); WG_T(4, clock64());
        }
    } else if (warp == 8) {
        // ---------------- loader
        const int t = tid - B2_CONV;
        for (int it = 0; it < n_it; ++it) {
            const int st = it % NSTAGE;
            if (it >= NSTAGE) { const long long t0 = clock64(); mbar_wait(bar_free + 8 * st, ((it / NSTAGE) - 1) & 1); t_wait += clock64() - t0; }
            const int64_t ch = ch_lo + it;
            const uint32_t base = sbase + st * STAGE;
            const uint32_t full = bar_full + 8 * st;
            if (A.use_tma) {
                if (t == 0) {
                    mbar_expect_tx(full, TX_BYTES);
                    tma_load_3d(base + S_B1, &maps.h0, 0, 0, (int)ch, full);
                    tma_load_3d(base + S_A0, &maps.dh0, 0, 0, (int)ch, full);
                    tma_load_3d(base + S_H1, &maps.h1, 0, 0, (int)ch, full);
                    tma_load_3d(base + S_B0, &maps.x, 0, 0, (int)ch, full);
                    bulk_g2s(base + S_G, A.k_glogit + ch * (KC * 3), KC * 3 * 4, full);
                    const uint32_t* mk = A.k_mask + ((ch >> 3) * 8 + 4) * 128 + (ch & 7) * KC;   // [tile][8][128], words 4-7 = h1
#pragma unroll
                    for (int j = 0; j < 4; ++j) bulk_g2s(base + S_M + j * KC * 4, mk + j * 128, KC * 4, full);
                }
            } else {
                // fallback: 16-byte async copies straight into the swizzled positions (chunk `ch` of each tensor is contiguous)
                const float* g_h0 = A.k_h0 + ch * (WD * KC);
                const float* g_dh0 = A.k_dh0 + ch * (WD * KC);
                const float* g_h1 = A.k_h1 + ch * (WD * KC);
                const float* g_x = A.k_x + ch * (40 * KC);
#pragma unroll 4
                for (int i = 0; i < 16; ++i) {
                    const int idx = t + 32 * i, f = idx >> 2, k4 = idx & 3;
                    const uint32_t o = sw64_off(f, k4);
                    const int go = f * KC + k4 * 4;
                    cp_async16(base + S_B1 + o, g_h0 + go);
                    cp_async16(base + S_A0 + o, g_dh0 + go);
                    cp_async16(base + S_H1 + o, g_h1 + go);
                }
#pragma unroll
                for (int i = 0; i < 5; ++i) {
                    const int idx = t + 32 * i, f = idx >> 2, k4 = idx & 3;
                    cp_async16(base + S_B0 + sw64_off(f, k4), g_x + f * KC + k4 * 4);
                }
                if (t < 12) cp_async16(base + S_G + t * 16, A.k_glogit + ch * (KC * 3) + t * 4);
                else if (t < 28) cp_async16(base + S_M + (t - 12) * 16, A.k_mask + ((ch >> 3) * 8 + 4 + ((t - 12) >> 2)) * 128 + (ch & 7) * KC + ((t - 12) & 3) * 4);
                cp_async_arrive(full);
            }
        }
        if (t == 0) { WG_T(5, t_wait); WG_T(6, clock64()); }
    } else {
        // ---------------- converters
        const int role = warp >> 2;                                  // 0: dH1 rows (computed), 1: dH0 rows (landed) + dW2
        const int j = (warp & 3) * 32 + (tid & 31);                  // feature row = TMEM lane (warp w: quarter w % 4)
        const uint32_t lane_addr = tmem + ((uint32_t)((warp & 3) * 32) << 16);
        const float w2j0 = __ldg(A.w2 + j), w2j1 = __ldg(A.w2 + WD + j), w2j2 = __ldg(A.w2 + 2 * WD + j);
        for (int it = 0; it < n_it; ++it) {
            const int st = it % NSTAGE;
            { const long long t0 = clock64(); mbar_wait(bar_full + 8 * st, (it / NSTAGE) & 1); t_wait += clock64() - t0; }
            unsigned char* sb = smem + st * STAGE;
            unsigned char* lo = smem + S_LOSETS + (it & 1) * LOSET;
            // The chunk is in shared memory and nobody will ever read its global copy again (h0, dH0, h1, x are rewritten by the
            // next step's forward): discard.global.L2 drops the lines — most of them still DIRTY in L2, written by the forward /
            // activation-gradient kernels — without writing them back.  Measured in situ (ncu
            // --cache-control none): without it the step moves 318 MB through DRAM, 148 MB of it these dead tensors on their way
            // out, and every line read here displaces a dirty line another CTA is about to read (113 of 148 MB re-read from
            // DRAM); see profiles/insitu_traffic_r02.md.  212 lines of 128 bytes per chunk: one per converter thread.
            if (A.l2_discard && tid < 212) {
                const int64_t ch = ch_lo + it;
                const char* line = tid < 64    ? reinterpret_cast<const char*>(A.k_h0 + ch * (WD * KC)) + tid * 128
                                   : tid < 128 ? reinterpret_cast<const char*>(A.k_dh0 + ch * (WD * KC)) + (tid - 64) * 128
                                   : tid < 192 ? reinterpret_cast<const char*>(A.k_h1 + ch * (WD * KC)) + (tid - 128) * 128
                                               : reinterpret_cast<const char*>(A.k_x + ch * (40 * KC)) + (tid - 192) * 128;
                asm volatile("discard.global.L2 [%0], 128;" ::"l"(line) : "memory");
            }
            // this lo set and this set of A columns were last read by the MMAs of chunk it-2
            if (it >= 2) { const long long t0 = clock64(); mbar_wait(bar_free + 8 * ((it - 2) % NSTAGE), ((it - 2) / NSTAGE) & 1); t_lo += clock64() - t0; tc_fence_after(); }
            const uint32_t aset = ASET + (uint32_t)(it & 1) * ASET_COLS;
            // All shared-memory loads first (a scheduler's two converter warps cannot hide the latency by themselves), then the
            // arithmetic, then the stores.  The raw B tiles are the hi operands as they are: async-copied data tracked by the
            // mbarrier needs no proxy fence.
            float4 vl[3];
#pragma unroll
            for (int q = 0; q < 3; ++q) {      // items 0..511: B1 rows 0..127; 512..671: B0 rows 0..39 (item = 16 bytes, physical position)
                const int u = tid + q * B2_CONV;
                if (u < N_LO_ITEMS) vl[q] = *reinterpret_cast<const float4*>(sb + (u < 512 ? S_B1 + u * 16 : S_B0 + (u - 512) * 16));
            }
            float4 g4[12];      // the [16][3] logit-gradient tile (broadcast reads)
#pragma unroll
            for (int q = 0; q < 12; ++q) g4[q] = *reinterpret_cast<const float4*>(sb + S_G + q * 16);
            const float* g = reinterpret_cast<const float*>(g4);
            float a[16];
            float4 h[4];
            if (role == 0) {
                // dH1[s][j] = [h1[s][j] > 0] (g[s] . W2[:, j]) for the 16 samples of the chunk
                const uint32_t* sM = reinterpret_cast<const uint32_t*>(sb + S_M) + (j >> 5) * KC;
                uint4 mw[4];
#pragma unroll
                for (int k4 = 0; k4 < 4; ++k4) mw[k4] = *reinterpret_cast<const uint4*>(sM + k4 * 4);
                const uint32_t* m = reinterpret_cast<const uint32_t*>(mw);
                const int bit = j & 31;
#pragma unroll
                for (int s = 0; s < 16; ++s) a[s] = (m[s] >> bit) & 1u ? fmaf(g[3 * s + 2], w2j2, fmaf(g[3 * s + 1], w2j1, g[3 * s] * w2j0)) : 0.f;
            } else {
                // dH0 row j of the landed tile, and the H1 row for dW2
#pragma unroll
                for (int k4 = 0; k4 < 4; ++k4) {
                    const float4 v = *reinterpret_cast<const float4*>(sb + S_A0 + sw64_off(j, k4));
                    a[4 * k4] = v.x; a[4 * k4 + 1] = v.y; a[4 * k4 + 2] = v.z; a[4 * k4 + 3] = v.w;
                    h[k4] = *reinterpret_cast<const float4*>(sb + S_H1 + sw64_off(j, k4));
                }
            }
            store_a_cols16(lane_addr, aset + (role == 0 ? A1HI : A0HI), aset + (role == 0 ? A1LO : A0LO), a);
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                const int u = tid + q * B2_CONV;
                if (u < N_LO_ITEMS) {
                    float4 l;
                    l.x = vl[q].x - __uint_as_float(__float_as_uint(vl[q].x) & 0xffffe000u);
                    l.y = vl[q].y - __uint_as_float(__float_as_uint(vl[q].y) & 0xffffe000u);
                    l.z = vl[q].z - __uint_as_float(__float_as_uint(vl[q].z) & 0xffffe000u);
                    l.w = vl[q].w - __uint_as_float(__float_as_uint(vl[q].w) & 0xffffe000u);
                    *reinterpret_cast<float4*>(lo + (u < 512 ? LO_B1 + u * 16 : LO_B0 + (u - 512) * 16)) = l;
                }
            }
            if (role == 1) {
#pragma unroll
                for (int k4 = 0; k4 < 4; ++k4) {
                    // samples 4*k4 .. 4*k4+3 occupy floats 12*k4 .. 12*k4+11 of the [16][3] gradient tile = g4[3*k4 .. 3*k4+2]
                    const float4 ga = g4[3 * k4], gb = g4[3 * k4 + 1], gc = g4[3 * k4 + 2];
                    const float4 hv = h[k4];
                    w2acc[0] = fmaf(ga.x, hv.x, w2acc[0]); w2acc[1] = fmaf(ga.y, hv.x, w2acc[1]); w2acc[2] = fmaf(ga.z, hv.x, w2acc[2]);
                    w2acc[0] = fmaf(ga.w, hv.y, w2acc[0]); w2acc[1] = fmaf(gb.x, hv.y, w2acc[1]); w2acc[2] = fmaf(gb.y, hv.y, w2acc[2]);
                    w2acc[0] = fmaf(gb.z, hv.z, w2acc[0]); w2acc[1] = fmaf(gb.w, hv.z, w2acc[1]); w2acc[2] = fmaf(gc.x, hv.z, w2acc[2]);
                    w2acc[0] = fmaf(gc.y, hv.w, w2acc[0]); w2acc[1] = fmaf(gc.z, hv.w, w2acc[1]); w2acc[2] = fmaf(gc.w, hv.w, w2acc[2]);
                }
                if (tid == 128) {   // db2: sum of the logit gradients of the valid samples
                    const int64_t s0 = (ch_lo + it) * KC;
#pragma unroll
                    for (int e = 0; e < 48; ++e)          // float e of the tile = (sample e / 3, channel e % 3)
                        if (s0 + e / 3 < M) gs[e % 3] += g[e];
                }
            }
            tmem_st_wait();
            tc_fence_before();
            fence_async_smem();
            mbar_arrive(bar_conv + 8 * st);
        }
        if (tid == 0) { WG_T(7, t_wait); WG_T(8, clock64()); WG_T(12, t_lo); }
    }
    // ---- drain and flush: each CTA stores its partial sums (net_grad layout); k_wgrad_reduce adds the gridDim.x partials.
    // 148 CTAs x 22 k red.global.add on the same addresses cost ~16 us; plain stores straight out of the TMEM lanes (one
    // 512-byte row per thread, half-used sectors, 39 scalar stores at a 156-byte stride) cost 11 k cycles; now the rows pass
    // through the (idle) stage ring and leave as one bulk store per row (dW1) / one contiguous block (dW0): 5 k.
    __syncthreads();
    if (tid == 0) WG_T(9, clock64());
    float* P = A.partial + (size_t)blockIdx.x * PART_LD;
    if (n_it == 0) {
        for (int e = tid; e < PART_LD; e += B2_THREADS) P[e] = 0.f;
    } else if (tid < B2_CONV) {
        // all MMAs are complete once the last commit of every stage in use has fired (then nothing reads the ring any more)
        for (int st = 0; st < NSTAGE && st < n_it; ++st) {
            const int last_it = ((n_it - 1 - st) / NSTAGE) * NSTAGE + st;
            mbar_wait(bar_free + 8 * st, (last_it / NSTAGE) & 1);
        }
        tc_fence_after();
        if (tid >= 128) {
#pragma unroll
            for (int c = 0; c < 3; ++c) P[PVDB_NET_OFF_W2 + c * WD + (tid - 128)] = w2acc[c];
        }
        if (tid == 128) { P[PVDB_NET_OFF_B2] = gs[0]; P[PVDB_NET_OFF_B2 + 1] = gs[1]; P[PVDB_NET_OFF_B2 + 2] = gs[2]; }
        float* dr1 = reinterpret_cast<float*>(smem);                                   // [128][DR_LD]: dW1 rows
        float* dr0 = reinterpret_cast<float*>(smem) + 128 * DR_LD;                     // [128 * 39]: dW0, already in the net_grad layout
        asm volatile("bar.sync 2, %0;" ::"n"(B2_CONV) : "memory");                      // every converter is past its last chunk
        {
            // warps w and w+4 share TMEM quarter w%4: thread (row j, half) drains 64 of the 128 dW1 columns; half 0 also db1 and
            // dW0[j][0..31], half 1 dW0[j][32..38] and db0
            const int j = (warp & 3) * 32 + (tid & 31), half = warp >> 2;
            const uint32_t lane_addr = tmem + ((uint32_t)((warp & 3) * 32) << 16);
#pragma unroll 1
            for (int c = half * 64; c < half * 64 + 64; c += 32) {
                uint32_t r[32];
                tmem_ld32(lane_addr + ACC1 + c, r);
                tmem_ld_wait();
                float4* dst = reinterpret_cast<float4*>(dr1 + j * DR_LD + c);
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    dst[i] = make_float4(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1]), __uint_as_float(r[4 * i + 2]), __uint_as_float(r[4 * i + 3]));
            }
            if (half == 0) {
                uint32_t r[32], q[8];
                tmem_ld32(lane_addr + ACC0, r);
                tmem_ld8(lane_addr + ACC1 + 128, q);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) dr0[j * PVDB_NET_DIN + i] = __uint_as_float(r[i]);
                P[PVDB_NET_OFF_B1 + j] = __uint_as_float(q[0]);
            } else {
                uint32_t q[8];
                tmem_ld8(lane_addr + ACC0 + 32, q);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 7; ++i) dr0[j * PVDB_NET_DIN + 32 + i] = __uint_as_float(q[i]);
                P[PVDB_NET_OFF_B0 + j] = __uint_as_float(q[7]);
            }
            fence_async_smem();
        }
        asm volatile("bar.sync 2, %0;" ::"n"(B2_CONV) : "memory");
        if (tid < 128) {
            bulk_s2g(P + PVDB_NET_OFF_W1 + tid * WD, smem_u32(dr1 + tid * DR_LD), WD * 4);
            if (tid == 0) bulk_s2g(P + PVDB_NET_OFF_W0, smem_u32(dr0), 128 * PVDB_NET_DIN * 4);
            bulk_commit();
            bulk_wait0();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (tid == 0) WG_T(10, clock64());
    if (warp == 0) tmem_dealloc(tmem, 512);
}

// net_grad[e] = sum over the CTAs' partials.  32 elements x 8 partial groups per CTA: every thread has all of its ~19 loads
// in flight at once (one latency), 128-byte coalesced across the 32 elements.
// DP (push.world > 1): the sums also go straight into every rank's netx[parity][this rank] over NVLink and the last CTA
// signals C — the rgbnet-gradient exchange of the data-parallel step has no kernel of its own (dp_exchange.cu).
__global__ void __launch_bounds__(256) k_wgrad_reduce(const float* __restrict__ partial, int n_part, float* __restrict__ net_grad, PvdbDpNetPush push,
                                                      PvdbNetAdam adam, PvdbDpNetWait wait) {
    __shared__ float red[8][32];
    __shared__ bool last;
    pvdb_pdl_wait();
    const int e = blockIdx.x * 32 + (threadIdx.x & 31), g = threadIdx.x >> 5;
    float a[19];
#pragma unroll
    for (int u = 0; u < 19; ++u) {
        const int c = g + 8 * u;
        a[u] = (e < PVDB_NET_N && c < n_part) ? __ldcg(partial + (size_t)c * PART_LD + e) : 0.f;
    }
    float s = 0.f;
#pragma unroll
    for (int u = 0; u < 19; ++u) s += a[u];
    for (int c = g + 8 * 19; c < n_part && e < PVDB_NET_N; c += 8) s += __ldcg(partial + (size_t)c * PART_LD + e);
    red[g][threadIdx.x & 31] = s;
    __syncthreads();
    if (g == 0 && e < PVDB_NET_N) {
        float t = 0.f;
#pragma unroll
        for (int q = 0; q < 8; ++q) t += red[q][threadIdx.x];
        if (adam.on && push.world <= 1) {   // single GPU, update phase: the rgbnet Adam right here (what k_update_fused's trailing CTAs do otherwise)
            const float g = net_grad[e] + t;
            pvdb_dense_adam_update(adam.net[e], adam.m[e], adam.v[e], g, 1.f, false, adam.scalars ? __ldg(adam.scalars + 2) : adam.stepsize, adam.b0, adam.b1,
                                   adam.eps);
            net_grad[e] = 0.f;
        } else {
            net_grad[e] += t;      // accumulating, like the grid gradients (the rgbnet Adam clears what it consumed)
        }
        if (push.world > 1) {
#pragma unroll
            for (int r = 0; r < 8; ++r)
                if (r < push.world) push.dst[r][e] = t;
        }
    }
    if (push.world > 1) {
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence_system();
            last = atomicAdd(push.done, 1u) == gridDim.x - 1;
            if (last) { *push.done = 0; __threadfence_system(); }
        }
        __syncthreads();
        if (last && threadIdx.x < push.world) st_release_sys(push.signal[threadIdx.x], push.epoch);
        if (adam.on) {
            // Data-parallel update: every rank's sums are in this rank's slots once the C words of all ranks (the own one included)
            // have reached the epoch; then the rank-ordered sum and the Adam of this CTA's 32 elements.  Every CTA pushes before it
            // waits and all CTAs of the grid are resident, so the wait cannot starve a peer.
            if (threadIdx.x < wait.world) wait_epoch(wait.signal + threadIdx.x, wait.epoch, wait.err, 1);
            __syncthreads();
            if (g == 0 && e < PVDB_NET_N) {
                float gs = __ldcg(wait.src + e);
                for (int r = 1; r < wait.world; ++r) gs += __ldcg(wait.src + (size_t)r * PVDB_DP_NET_PAD + e);
                pvdb_dense_adam_update(adam.net[e], adam.m[e], adam.v[e], gs, 1.f, false, adam.scalars ? __ldg(adam.scalars + 2) : adam.stepsize, adam.b0,
                                       adam.b1, adam.eps);
                net_grad[e] = 0.f;
            }
        }
    }
}

#ifdef PVDB_TC_TIMING
}  // namespace
extern "C" int pvdb_debug_wgrad_timing(long long* out) {
    PVDB_CUDA(cudaMemcpyFromSymbol(out, g_wg_t, sizeof(long long) * PVDB_SMS * 16));
    return PVDB_OK;
}
namespace {
#endif

// Tensor map of one chunk-major activation tensor: dims (16 samples, nf features, chunks), 64-byte inner box, SWIZZLE_64B.
// cuTensorMapEncodeTiled is resolved through the runtime (no link-time dependency on libcuda).
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int make_act_map(CUtensorMap* map, const float* base, int nf, int64_t cap_keep) {
    static EncodeTiledFn encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn)
            return 1;
        encode = reinterpret_cast<EncodeTiledFn>(fn);
    }
    const cuuint64_t n_chunks = (cuuint64_t)(cap_keep / KC);
    const cuuint64_t dims[3] = {(cuuint64_t)KC, (cuuint64_t)nf, n_chunks};
    const cuuint64_t strides[2] = {(cuuint64_t)ROWB, (cuuint64_t)nf * ROWB};
    const cuuint32_t box[3] = {(cuuint32_t)KC, (cuuint32_t)nf, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    return encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS
               ? 0
               : 2;
}

}  // namespace

static int bwd_attrs() {
    static bool attr_set = false;
    if (!attr_set) {
        PVDB_CUDA(cudaFuncSetAttribute(k_rgbnet_bwd_act_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, B1_TOTAL));
        PVDB_CUDA(cudaFuncSetAttribute(k_rgbnet_bwd_wgrad_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, B2_TOTAL));
        attr_set = true;
    }
    return PVDB_OK;
}

// B1: activation gradients + k0 gradient scatter (final k0_grad / k0_touched once it completes)
int pvdb_rgbnet_backward_act_tc(const pvdb_train_cfg* cfg, const pvdb_train_bufs* b, const float* viewdirs, cudaStream_t st) {
    PVDB_CHECK_ARG(b->k_h0 && b->k_h1 && b->k_dh0 && b->k_x && b->k_mask, "the tcgen05 backward needs k_h0, k_h1, k_dh0, k_x, k_mask");
    if (int rc = bwd_attrs()) return rc;
    PVDB_CHECK_ARG(b->net_img && b->k_corner, "net_img / k_corner scratch missing (tensor-core backward)");
    // second half of the scratch: backward image, built together with the forward image by pvdb_rgbnet_forward_tc
    const unsigned char* img = static_cast<const unsigned char*>(b->net_img) + PVDB_BWD_IMG_OFFSET;
    BwdActArgs A;
    A.img = img; A.k_glogit = b->k_rgb; A.k_mask = b->k_mask; A.k_xyz = b->k_xyz; A.k_corner = b->k_corner;
    A.k_dh0 = b->k_dh0; A.k0_grad = b->k0_grad; A.k0_touched = b->k0_touched; A.k0_touched_list = b->k0_touched_list; A.counters_w = b->counters;
    A.counters = b->counters;
    A.cap_keep = b->cap_keep;
    const bool ll = pvdb_leaf_local_enabled(b);
    A.k_dx = ll ? b->k_dx : nullptr;
    PVDB_CUDA(pvdb_launch_pdl(k_rgbnet_bwd_act_tc, dim3(PVDB_SMS), dim3(B1_THREADS), B1_TOTAL, st, A));
    PVDB_LAUNCH_CHECK();
    pvdb_prof_mark("rgbnet_bwd_act", st);
    if (ll) return pvdb_leaf_local_backward(b, st);      // the k0 gradient planes are final when this has run
    return PVDB_OK;
}

// B2: weight gradients -> net_grad (accumulated)
int pvdb_rgbnet_backward_wgrad_tc(const pvdb_train_cfg* cfg, const pvdb_train_bufs* b, cudaStream_t st, const PvdbDpNetPush* dp_push,
                                  const PvdbNetAdam* adam, const PvdbDpNetWait* dp_wait) {
    if (int rc = bwd_attrs()) return rc;
    BwdWgradArgs W;
    W.k_h0 = b->k_h0; W.k_dh0 = b->k_dh0; W.k_x = b->k_x; W.k_h1 = b->k_h1; W.k_glogit = b->k_rgb;
    W.k_mask = b->k_mask; W.w2 = b->net + PVDB_NET_OFF_W2;
    PVDB_CHECK_ARG(b->net_partial, "net_partial scratch missing (tensor-core backward)");
    W.partial = b->net_partial; W.counters = b->counters; W.cap_keep = b->cap_keep;
    // tensor maps are pure host-side encodings of (pointer, shape): rebuilt when a buffer changes
    static thread_local WgradMaps maps;
    static thread_local const void* maps_key[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    static thread_local int maps_ok = 0;
    const void* key[6] = {nullptr, b->k_h0, b->k_dh0, b->k_h1, b->k_x, (const void*)(intptr_t)b->cap_keep};
    if (memcmp(key, maps_key, sizeof(key)) != 0) {
        const int rc = make_act_map(&maps.h0, b->k_h0, WD, b->cap_keep) |
                       make_act_map(&maps.dh0, b->k_dh0, WD, b->cap_keep) | make_act_map(&maps.h1, b->k_h1, WD, b->cap_keep) |
                       make_act_map(&maps.x, b->k_x, 40, b->cap_keep);
        maps_ok = rc == 0;
        memcpy(maps_key, key, sizeof(key));
    }
    static const bool no_tma = getenv("PVDB_NO_TMA") != nullptr;   // bring-up switch: cp.async loader instead of TMA
    W.use_tma = maps_ok && !no_tma;
    // Off by default: it takes 86 MB per step off the DRAM (319 -> 233 MB in situ) but not a microsecond off the kernels, which
    // are not DRAM-bound in situ (1.2 - 3.2 TB/s); the weight-gradient kernel even loses 2 us to the 212 extra instructions per
    // chunk (profiles/insitu_traffic_r02.md).  PVDB_L2_DISCARD=1 switches it on.
    static const bool discard = getenv("PVDB_L2_DISCARD") != nullptr && atoi(getenv("PVDB_L2_DISCARD")) != 0;
    W.l2_discard = discard ? 1 : 0;
    PVDB_CUDA(pvdb_launch_pdl(k_rgbnet_bwd_wgrad_tc, dim3(PVDB_SMS), dim3(B2_THREADS), B2_TOTAL, st, W, maps));
    PVDB_LAUNCH_CHECK();
    PvdbDpNetPush push = {};
    if (dp_push) push = *dp_push;
    PvdbNetAdam ad = {};
    if (adam) ad = *adam;
    PvdbDpNetWait wt = {};
    if (dp_wait) wt = *dp_wait;
    PVDB_CHECK_ARG(!(ad.on && push.world > 1) || wt.world == push.world, "data-parallel rgbnet Adam in the reduction needs the wait arguments");
    PVDB_CUDA(pvdb_launch_pdl(k_wgrad_reduce, dim3((PVDB_NET_N + 31) / 32), dim3(256), 0, st, (const float*)b->net_partial, (int)PVDB_SMS, b->net_grad, push, ad, wt));
    PVDB_LAUNCH_CHECK();
    return PVDB_OK;
}

int pvdb_rgbnet_backward_tc(const pvdb_train_cfg* cfg, const pvdb_train_bufs* b, const float* viewdirs, cudaStream_t st) {
    if (int rc = pvdb_rgbnet_backward_act_tc(cfg, b, viewdirs, st)) return rc;
    return pvdb_rgbnet_backward_wgrad_tc(cfg, b, st, nullptr);
}
