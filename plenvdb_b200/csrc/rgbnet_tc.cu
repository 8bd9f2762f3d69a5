// rgbnet_tc.cu — the rgbnet MLP (39 -> 128 -> 128 -> 3) on the 5th-generation tensor cores (tcgen05).
//
// The only dense contraction of the hot path (dvgo.py:99-107, 354-360; renderer.cu:83-109).  One persistent CTA
// per SM, 128-sample tiles:
//   * accumulators live in TMEM (128 lanes = 128 samples, 128 fp32 columns);
//   * between two layers the activations never touch shared memory: a thread owns one sample = one TMEM lane, reads
//     its half of the accumulator row with tcgen05.ld, applies the ReLU (the biases are inside the MMAs), and writes the
//     next layer's A operand straight back into TMEM with tcgen05.st (A-from-TMEM form of tcgen05.mma); the copies the
//     backward needs leave through per-warp staging regions and bulk async stores;
//   * the weights stay resident in shared memory for the CTA's lifetime in the canonical K-major (no swizzle)
//     UMMA layout, read through shared-memory matrix descriptors;
//   * precision: 3xTF32 error-compensated products (x = hi + lo, D = Ahi*Bhi + Alo*Bhi + Ahi*Blo, fp32 accumulate)
//     -> ~2^-21 relative per product, which keeps rendered RGB within the 1e-5 parity tolerance where single TF32
//     (~2^-11) would not.
// One thread issues the MMAs; completion is signalled through tcgen05.commit -> mbarrier.
#include <cuda_fp16.h>
#include <stdlib.h>
#include "common.cuh"
#include "ray_math.cuh"
#include "rgbnet.cuh"
#include "render_ray.cuh"
#include "tc_ptx.cuh"
#include "leaf_local.cuh"

namespace {

constexpr int WD = PVDB_NET_W;     // 128
constexpr int K0P = 40;            // layer-0 K (39 padded to a multiple of 8)
constexpr int CNT_M_KEEP = 1;

// TMEM column map (512 columns allocated): A_hi [0,128) A_lo [128,256) D0 [256,384) D1 [384,512).
// Two accumulators let the tensor core run one layer while the lane warps drain the other.
constexpr uint32_t COL_AHI = 0, COL_ALO = 128, COL_D0 = 256, COL_D1 = 384;

// Shared memory map (bytes).  [0, IMG_BYTES) is the "weight image": hi/lo tf32 splits of W0, W1 in the canonical K-major
// UMMA layout, then W2 and the biases as plain fp32.  k_prep_fwd_image builds it once per launch in global memory and
// every CTA pulls it in with one bulk async copy (TMA 1-D) instead of re-splitting 22 k weights itself.
constexpr int SM_W0HI = 0;                                  // [128][40]
constexpr int SM_W0LO = SM_W0HI + WD * K0P * 4;
constexpr int SM_W1HI = SM_W0LO + WD * K0P * 4;             // [128][128]
constexpr int SM_W1LO = SM_W1HI + WD * WD * 4;
constexpr int SM_W2F = SM_W1LO + WD * WD * 4;               // [3][128] fp32 (layer 2 runs on the CUDA cores)
// The biases ride on the tensor core: b0 is column 39 of the W0 image (the K padding; the staged input row carries a 1
// there), b1 is one extra k-step of layer 1 whose A operand is a constant tile of ones (columns 0, 1) and whose B operand
// holds hi(b1) in k = 0 and lo(b1) in k = 1 — one MMA per tile instead of an LDS + FADD per activation value.
constexpr int SM_ONES = SM_W2F + 3 * WD * 4;                // A [128][8]
constexpr int SM_B1T = SM_ONES + WD * 8 * 4;                // B [128][8]
constexpr int SM_B2 = SM_B1T + WD * 8 * 4;                  // [4]
constexpr int IMG_BYTES = SM_B2 + 16;
constexpr int SM_BAR = IMG_BYTES;   // mbarriers: W, L0, L1, D0_FREE, X_RDY, A_RDY[4]; tmem base at +96
constexpr int BAR_W = 0, BAR_L0 = 8, BAR_L1 = 16, BAR_D0FREE = 24, BAR_XRDY = 56, BAR_ARDY = 64, TMEM_SLOT = 96;
constexpr int SM_RED = SM_BAR + 112;                        // [3][128] partial outputs of the second lane-warp group
constexpr int SM_STAGE = SM_RED + 3 * 128 * 4;              // activation store staging, one region per lane warp (stage_store32)
constexpr int SM_TOTAL = SM_STAGE + 8 * STAGE_WARP_BYTES;
static_assert(IMG_BYTES % 16 == 0 && SM_STAGE % 16 == 0 && SM_TOTAL <= 227 * 1024, "smem map");

// Issue the 3xTF32 MMAs of k-steps [ks0, ks1) of one layer: D[128 x N] (+)= A[128 x K] * W[N x K]^T.  Single thread.
__device__ __forceinline__ void issue_ksteps(uint32_t tmem, uint32_t d_col, uint32_t smem_base, int off_hi, int off_lo, int K, int N,
                                             int ks0, int ks1, bool zero_first) {
    const uint32_t idesc = make_idesc(N);
    const uint64_t bhi = make_desc(smem_base + off_hi, K), blo = make_desc(smem_base + off_lo, K);
    uint32_t acc = zero_first ? 0u : 1u;
    for (int ks = ks0; ks < ks1; ++ks) {
        // one k-step = 8 tf32 = two 16-byte chunks = 256 B further along the group: start address field is in 16-B units
        const uint64_t adv = (uint64_t)(ks * 256) >> 4;
        const uint32_t ahi = tmem + COL_AHI + ks * 8, alo = tmem + COL_ALO + ks * 8;
        umma_tf32_ts(tmem + d_col, ahi, bhi + adv, idesc, acc);
        acc = 1;
        umma_tf32_ts(tmem + d_col, alo, bhi + adv, idesc, 1);
        umma_tf32_ts(tmem + d_col, ahi, blo + adv, idesc, 1);
    }
}

// Write a row of activations (n values, multiple of 8, starting at column c0) as the next A operand.
__device__ __forceinline__ void store_a_row(uint32_t tmem_lane, int c0, const float* v, int n) {
#pragma unroll
    for (int c = 0; c < n; c += 8) {
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) split_tf32(v[c + i], hi[i], lo[i]);
        tmem_st8(tmem_lane + COL_AHI + c0 + c, hi);
        tmem_st8(tmem_lane + COL_ALO + c0 + c, lo);
    }
}

// Optional phase timing (build with PVDB_EXTRA_NVCC_FLAGS=-DPVDB_TC_TIMING): clock64 stamps of thread 0, first 8 tiles per CTA.
#ifdef PVDB_TC_TIMING
__device__ unsigned long long g_tc_gt[PVDB_SMS][2];      // %globaltimer at CTA entry / exit of k_rgbnet_fwd_tc
__device__ __forceinline__ unsigned long long gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__device__ long long g_tc_t[PVDB_SMS][8][16];
#define TC_T(i) do { if (tid == 0 && tile_no < 8) g_tc_t[blockIdx.x][tile_no][i] = clock64(); } while (0)
#else
#define TC_T(i) do { } while (0)
#endif

struct TcWeights {
    const float *w0, *b0, *w1, *b1, *w2, *b2;
    int w0_sn, w0_sk, w1_sn, w1_sk, w2_sn, w2_sk;   // strides (floats) of W[n][k]
};

// Build the weight image (see the smem map) in global memory.  Any grid; ~22 k elements.
// net_packed != nullptr (training): also the activation-gradient kernel's image, PVDB_BWD_IMG_OFFSET further on.
__global__ void __launch_bounds__(256) k_prep_fwd_image(TcWeights Wt, unsigned char* __restrict__ img, const float* __restrict__ net_packed) {
    pvdb_pdl_wait();
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gsz = gridDim.x * blockDim.x;
    if (net_packed) prep_bwd_image(net_packed, img + PVDB_BWD_IMG_OFFSET, gtid, gsz);
    for (int e = gtid; e < WD * K0P; e += gsz) {
        const int n = e / K0P, k = e % K0P;
        const float v = k < PVDB_NET_DIN ? __ldg(Wt.w0 + (size_t)n * Wt.w0_sn + (size_t)k * Wt.w0_sk) : __ldg(Wt.b0 + n);
        uint32_t hi, lo;
        split_tf32(v, hi, lo);
        const int o = canon_off(n, k, K0P);
        *reinterpret_cast<uint32_t*>(img + SM_W0HI + o) = hi;
        *reinterpret_cast<uint32_t*>(img + SM_W0LO + o) = lo;
    }
    for (int e = gtid; e < WD * WD; e += gsz) {
        const int n = e / WD, k = e % WD;
        const float v = __ldg(Wt.w1 + (size_t)n * Wt.w1_sn + (size_t)k * Wt.w1_sk);
        uint32_t hi, lo;
        split_tf32(v, hi, lo);
        const int o = canon_off(n, k, WD);
        *reinterpret_cast<uint32_t*>(img + SM_W1HI + o) = hi;
        *reinterpret_cast<uint32_t*>(img + SM_W1LO + o) = lo;
    }
    float* w2f = reinterpret_cast<float*>(img + SM_W2F);
    for (int e = gtid; e < 3 * WD; e += gsz) w2f[e] = __ldg(Wt.w2 + (size_t)(e / WD) * Wt.w2_sn + (size_t)(e % WD) * Wt.w2_sk);
    for (int e = gtid; e < WD * 8; e += gsz) {
        const int n = e >> 3, k = e & 7;
        uint32_t hi, lo;
        split_tf32(__ldg(Wt.b1 + n), hi, lo);
        const int o = canon_off(n, k, 8);
        *reinterpret_cast<float*>(img + SM_ONES + o) = k < 2 ? 1.0f : 0.f;
        *reinterpret_cast<uint32_t*>(img + SM_B1T + o) = k == 0 ? hi : k == 1 ? lo : 0u;
    }
    if (gtid < 4) reinterpret_cast<float*>(img + SM_B2)[gtid] = gtid < 3 ? __ldg(Wt.b2 + gtid) : 0.f;
}

// x[12..39] of every ray of the batch: [vd(3) | sin(vd_a 2^k) a-major (12) | cos (12) | 0], the same expressions as
// view_embed_tc.  One thread per ray; runs with the image prep on the side stream, underneath the march.
__global__ void __launch_bounds__(128) k_ray_pe(const float* __restrict__ viewdirs, int n_rays, float* __restrict__ pe) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rays) return;
    const float d[3] = {__ldg(viewdirs + (size_t)r * 3), __ldg(viewdirs + (size_t)r * 3 + 1), __ldg(viewdirs + (size_t)r * 3 + 2)};
    float o[28];
    o[0] = d[0]; o[1] = d[1]; o[2] = d[2]; o[27] = 0.f;
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float x = __fmul_rn(d[a], (float)(1 << k));
            o[3 + a * 4 + k] = sinf(x);
            o[15 + a * 4 + k] = cosf(x);
        }
    float4* dst = reinterpret_cast<float4*>(pe + (size_t)r * 28);
#pragma unroll
    for (int q = 0; q < 7; ++q) dst[q] = make_float4(o[q * 4], o[q * 4 + 1], o[q * 4 + 2], o[q * 4 + 3]);
}

// Common body, warp specialised (13 warps):
//   warps 0-7   "lane warps": thread = (TMEM lane = sample, column half).  Warp w and w+4 share the 32 lanes of quarter
//               w%4 (the TMEM access rule) and split every 128-column row in two.  They run the epilogues (ReLU, tf32 split
//               back into TMEM as the next A operand, activation stores through the per-warp staging regions) and the
//               128 -> 3 output layer on the CUDA cores;
//   warps 8-11  "producers": thread = sample; warp 8+q owns TMEM quarter q.  They gather the input row of the next tile
//               (k0 trilinear / feature list + view PE: the memory-latency bound part) into registers while the current
//               tile computes, and write it straight into TMEM as the layer-0 A operand once layer 1 of the current tile
//               has left those columns;
//   warp 12     "issuer": one thread feeds the tensor core.  It never touches TMEM data itself, so the lane warps are never
//               held up behind a queue of MMAs: they signal "chunk c of A is in TMEM" through mbarriers and move on.
// Per tile: epilogue 0 drains D0 in four 32-column chunks and the issuer starts layer 1's k-steps of each chunk (into D1)
// as soon as it lands; while the lane warps run epilogue 1 from D1, the next tile's layer 0 already runs into D0.
// The biases are inside the MMAs (see the smem map).
// FeatFn(s, valid, x[40]) fills the input row of sample s; ActFn(s, valid, layer, c, h[32], stage) sees each post-ReLU chunk;
// OutFn(s, raw[3]) consumes the result.
constexpr int FWD_THREADS = 416;
constexpr int N_LANE_THREADS = 256;

template <class FeatFn, class OutFn, class ActFn>
__device__ void mlp_tiles(unsigned char* smem, const unsigned char* __restrict__ img, const int32_t* __restrict__ m_ptr, int64_t m_cap,
                          FeatFn feat, OutFn out, ActFn act) {
    const int tid = threadIdx.x, warp = tid >> 5;
    int tile_no = 0;
    TC_T(0);
    const uint32_t sbase = smem_u32(smem);
    const uint32_t bars = sbase + SM_BAR;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + SM_BAR + TMEM_SLOT);
    const float* sb2 = reinterpret_cast<const float*>(smem + SM_B2);
    const float* sw2 = reinterpret_cast<const float*>(smem + SM_W2F);
    float* red = reinterpret_cast<float*>(smem + SM_RED);
    if (tid == 0) {
        mbar_init(bars + BAR_W, 1);
        mbar_init(bars + BAR_L0, 1);
        mbar_init(bars + BAR_L1, 1);
        mbar_init(bars + BAR_XRDY, 128);
        mbar_init(bars + BAR_D0FREE, N_LANE_THREADS);
        for (int c = 0; c < 4; ++c) mbar_init(bars + BAR_ARDY + 8 * c, 128);
        fence_async_smem();
        mbar_expect_tx(bars + BAR_W, IMG_BYTES);
        constexpr int CH = 32768;
        for (int o = 0; o < IMG_BYTES; o += CH) bulk_g2s(sbase + o, img + o, min(CH, IMG_BYTES - o), bars + BAR_W);
    }
    if (warp == 0) tmem_alloc(sbase + SM_BAR + TMEM_SLOT, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    // Everything above (barriers, TMEM, the weight image in flight) is independent of the kernel before this one: with a
    // programmatic dependent launch it overlaps that kernel's tail.  The sample count and the lists are read after the wait.
    pvdb_pdl_wait();
    const int64_t M = min((int64_t)*m_ptr, m_cap);
    const int64_t n_tiles = (M + TM - 1) / TM;
    TC_T(1);

    if (warp == 12) {
        // ---------------- issuer
        if (tid == 384 && (int64_t)blockIdx.x < n_tiles) {
            mbar_wait(bars + BAR_W, 0);          // weight image landed
            int i = 0;
            mbar_wait(bars + BAR_XRDY, 0);
            tc_fence_after();
            issue_ksteps(tmem, COL_D0, sbase, SM_W0HI, SM_W0LO, K0P, WD, 0, K0P / 8, true);
            umma_commit(bars + BAR_L0);
            for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++i) {
                for (int cc = 0; cc < 4; ++cc) {
                    const int c = (cc >> 1) | ((cc & 1) << 1);      // 0, 2, 1, 3: the order the two column halves deliver their chunks
                    mbar_wait(bars + BAR_ARDY + 8 * c, i & 1);
                    tc_fence_after();
                    if (cc == 0)     // D1 = 1 * hi(b1) + 1 * lo(b1): the layer's k-steps accumulate on top of the bias
                        umma_tf32_ss(tmem + COL_D1, make_desc(sbase + SM_ONES, 8), make_desc(sbase + SM_B1T, 8), make_idesc(WD), 0u);
                    issue_ksteps(tmem, COL_D1, sbase, SM_W1HI, SM_W1LO, WD, WD, c * 4, c * 4 + 4, false);
                }
                umma_commit(bars + BAR_L1);
                if (tile + gridDim.x < n_tiles) {
                    mbar_wait(bars + BAR_XRDY, (i + 1) & 1);
                    mbar_wait(bars + BAR_D0FREE, i & 1);      // the lane warps read D0 a second time for the activation stores
                    tc_fence_after();
                    issue_ksteps(tmem, COL_D0, sbase, SM_W0HI, SM_W0LO, K0P, WD, 0, K0P / 8, true);
                    umma_commit(bars + BAR_L0);
                }
            }
        }
    } else if (warp >= 8) {
        // ---------------- producers
        const int lane_s = (warp & 3) * 32 + (tid & 31);         // sample within the tile = TMEM lane (warp 8+q: quarter q)
        const uint32_t lane_addr = tmem + ((uint32_t)((warp & 3) * 32) << 16);
        int i = 0;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++i) {
            const int64_t s = tile * TM + lane_s;
            float x[K0P];
#pragma unroll
            for (int q = 0; q < K0P; ++q) x[q] = 0.f;
#ifdef PVDB_TC_TIMING
            if (tid == N_LANE_THREADS && i < 8) g_tc_t[blockIdx.x][i][12] = clock64();
#endif
            feat(s, s < M, x);
#ifdef PVDB_TC_TIMING
            if (tid == N_LANE_THREADS && i < 8) g_tc_t[blockIdx.x][i][13] = clock64();
#endif
            x[K0P - 1] = 1.0f;      // multiplies the b0 column of the W0 image
            // columns 0-39 of A still feed layer 1 of the previous tile until its MMAs have completed
            if (i > 0) { mbar_wait(bars + BAR_L1, (i - 1) & 1); tc_fence_after(); }
            store_a_row(lane_addr, 0, x, K0P);
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(bars + BAR_XRDY);
        }
    } else if ((int64_t)blockIdx.x < n_tiles) {
        // ---------------- lane warps
        const int grp = warp >> 2;                               // column half
        const int lane_s = (warp & 3) * 32 + (tid & 31);         // sample within the tile = TMEM lane
        const uint32_t lane_addr = tmem + ((uint32_t)((warp & 3) * 32) << 16);   // this warp's 32-lane quarter
        unsigned char* stage = smem + SM_STAGE + warp * STAGE_WARP_BYTES;
        uint32_t par0 = 0, par1 = 0;
        mbar_wait(bars + BAR_W, 0);      // W2 / b2 of the image are read below
        const int cb = grp * 64;         // this thread's 64 columns of every 128-column row
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tile_no) {
            const int64_t s = tile * TM + lane_s;
            const bool valid = s < M;
            float o0 = 0.f, o1 = 0.f, o2 = 0.f;
            TC_T(2);
            // Six phases, ONE copy of the code: the epilogues are ~350 instructions per phase and every launch starts with a cold
            // instruction cache — unrolled, the first tile of a launch took 23.5 k cycles against 8.5 k for the later ones, a
            // pure-ALU block 3.0 k against 0.5 k (scratch/tc_timing.py).
            //   0, 1  D0 chunk -> ReLU -> A: only what layer 1 waits for, so its MMAs start 0.4 k cycles into the tile
            //   2, 3  D0 chunk again -> ReLU -> mask + staged store of h0, underneath those MMAs; then D0 is handed back
            //   4, 5  D1 chunk -> ReLU -> mask + staged store of h1 -> 128 -> 3 output layer on the CUDA cores (the next
            //         tile's layer 0 runs underneath)
#pragma unroll 1
            for (int ph = 0; ph < 6; ++ph) {
                const int c = cb + (ph & 1) * 32;
                if (ph == 0) { mbar_wait(bars + BAR_L0, par0); par0 ^= 1; tc_fence_after(); TC_T(3); }
                if (ph == 4) { TC_T(4); mbar_wait(bars + BAR_L1, par1); par1 ^= 1; tc_fence_after(); TC_T(5); TC_T(6); }
                uint32_t r[32];
                tmem_ld32(lane_addr + (ph < 4 ? COL_D0 : COL_D1) + c, r);
                tmem_ld_wait();
                float h[32];
#pragma unroll
                for (int q = 0; q < 32; ++q) h[q] = fmaxf(__uint_as_float(r[q]), 0.f);
                if (ph < 2) {
                    store_a_row(lane_addr, c, h, 32);
                    tmem_st_wait();
                    tc_fence_before();   // (first chunk: also orders the previous tile's D1 reads before the new layer-1 MMAs)
                    mbar_arrive(bars + BAR_ARDY + 8 * (c >> 5));
                } else {
                    act(s, valid, ph >> 2, c, h, stage);
                    if (ph == 3) { tc_fence_before(); mbar_arrive(bars + BAR_D0FREE); }
                }
                if (ph >= 4) {
#pragma unroll
                    for (int q = 0; q < 32; q += 4) {
                        const float4 a = *reinterpret_cast<const float4*>(sw2 + c + q);
                        const float4 bq = *reinterpret_cast<const float4*>(sw2 + WD + c + q);
                        const float4 cq = *reinterpret_cast<const float4*>(sw2 + 2 * WD + c + q);
                        o0 = fmaf(h[q], a.x, o0); o0 = fmaf(h[q + 1], a.y, o0); o0 = fmaf(h[q + 2], a.z, o0); o0 = fmaf(h[q + 3], a.w, o0);
                        o1 = fmaf(h[q], bq.x, o1); o1 = fmaf(h[q + 1], bq.y, o1); o1 = fmaf(h[q + 2], bq.z, o1); o1 = fmaf(h[q + 3], bq.w, o1);
                        o2 = fmaf(h[q], cq.x, o2); o2 = fmaf(h[q + 1], cq.y, o2); o2 = fmaf(h[q + 2], cq.z, o2); o2 = fmaf(h[q + 3], cq.w, o2);
                    }
                }
            }
            // the two column halves meet in shared memory: group 1 hands its partial sums to group 0
            if (grp == 1) { red[lane_s] = o0; red[128 + lane_s] = o1; red[256 + lane_s] = o2; }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (grp == 0 && valid) {
                const float raw[3] = {o0 + red[lane_s] + sb2[0], o1 + red[128 + lane_s] + sb2[1], o2 + red[256 + lane_s] + sb2[2]};
                out(s, raw);
            }
            TC_T(7);
        }
        if ((tid & 31) == 0) bulk_wait0();      // this warp's activation blocks are in global memory before the kernel ends
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

// ---- training forward ---------------------------------------------------------------------------------------
__device__ __forceinline__ void view_embed_tc(const float* __restrict__ vd, float* __restrict__ out) {
    const float d[3] = {__ldg(vd), __ldg(vd + 1), __ldg(vd + 2)};
    out[0] = d[0]; out[1] = d[1]; out[2] = d[2];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float x = __fmul_rn(d[a], (float)(1 << k));
            out[3 + a * 4 + k] = sinf(x);
            out[15 + a * 4 + k] = cosf(x);
        }
}

struct TrainFwdArgs {
    pvdb_tree tree;
    const float* k0; const float* viewdirs; const int32_t* k_ray; const float* k_xyz;
    float *k_feat, *k_h0, *k_h1, *k_rgb, *k_x;
    uint32_t* k_mask;
    const int32_t* k_corner;
    const int32_t* counters; int64_t cap_keep;
    const unsigned char* img;
    int feat_ready;   // k_feat already holds the interpolated features (leaf_local.cu): read them instead of gathering
#ifdef PVDB_TC_TIMING
    int exp;
#endif
    const float* ray_pe;   // [n_rays][28]: x[12..39] of every ray (view direction, sin, cos, padding), k_ray_pe; null: computed per sample
};

__global__ void __launch_bounds__(FWD_THREADS, 1) k_rgbnet_fwd_tc(TrainFwdArgs A) {
    extern __shared__ __align__(128) unsigned char smem[];
#ifdef PVDB_TC_TIMING
    if (threadIdx.x == 0) g_tc_gt[blockIdx.x][0] = gtimer();
#endif
    auto feat = [&](int64_t s, bool valid, float* x) {
        if (!valid) {   // lanes past M in the last tile: zero input row in HBM (the weight-gradient pass reads whole tiles)
            if (A.k_x) {
                float* kx = A.k_x + act_off(s, 0, 40);
#pragma unroll
                for (int i = 0; i < 40; ++i) kx[i * 16] = 0.f;
            }
            return;
        }
        // 12-channel trilinear sample, colorvdb.cu:81-111 arithmetic and corner order.  The eight record ids were found
        // by the march (same topology as the density grid), so all 24 16-byte loads are independent and in flight at
        // once; a missing corner contributes fma(sc, 0, x) = x, bit-identical to skipping it.
        if (A.feat_ready) {
            const float4* f = reinterpret_cast<const float4*>(A.k_feat + s * 12);
#pragma unroll
            for (int c4 = 0; c4 < 3; ++c4) {
                const float4 a = __ldcg(f + c4);
                x[c4 * 4] = a.x; x[c4 * 4 + 1] = a.y; x[c4 * 4 + 2] = a.z; x[c4 * 4 + 3] = a.w;
            }
        } else {
        const float* p = A.k_xyz + s * 3;
        PvdbTri tri;
        tri.set(p[0], p[1], p[2]);
        const int4 ca = __ldg(reinterpret_cast<const int4*>(A.k_corner + s * 8));
        const int4 cb = __ldg(reinterpret_cast<const int4*>(A.k_corner + s * 8) + 1);
        const int rec[8] = {ca.x, ca.y, ca.z, ca.w, cb.x, cb.y, cb.z, cb.w};
        float4 v[8][3];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const float4* src = reinterpret_cast<const float4*>(A.k0 + (size_t)max(rec[q], 0) * 12);
#pragma unroll
            for (int c4 = 0; c4 < 3; ++c4) v[q][c4] = rec[q] >= 0 ? __ldg(src + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int dx = PVDB_CORNER[q][0], dy = PVDB_CORNER[q][1], dz = PVDB_CORNER[q][2];
            const float sc = __fmul_rn(__fmul_rn(tri.f(0, dx), tri.f(1, dy)), tri.f(2, dz));
#pragma unroll
            for (int c4 = 0; c4 < 3; ++c4) {
                const float4 a = v[q][c4];
                x[c4 * 4 + 0] = __fmaf_rn(sc, a.x, x[c4 * 4 + 0]); x[c4 * 4 + 1] = __fmaf_rn(sc, a.y, x[c4 * 4 + 1]);
                x[c4 * 4 + 2] = __fmaf_rn(sc, a.z, x[c4 * 4 + 2]); x[c4 * 4 + 3] = __fmaf_rn(sc, a.w, x[c4 * 4 + 3]);
            }
        }
        float4* kf = reinterpret_cast<float4*>(A.k_feat + s * 12);
        kf[0] = make_float4(x[0], x[1], x[2], x[3]); kf[1] = make_float4(x[4], x[5], x[6], x[7]); kf[2] = make_float4(x[8], x[9], x[10], x[11]);
        }
        if (A.ray_pe) {      // the embedding depends on the ray only (dvgo.py:354-357): ~10 kept samples share one table row
            const float4* pe = reinterpret_cast<const float4*>(A.ray_pe + (size_t)A.k_ray[s] * 28);
#pragma unroll
            for (int q = 0; q < 7; ++q) {
                const float4 a = __ldg(pe + q);
                x[12 + q * 4] = a.x; x[13 + q * 4] = a.y; x[14 + q * 4] = a.z; x[15 + q * 4] = a.w;
            }
        } else {
            view_embed_tc(A.viewdirs + (size_t)A.k_ray[s] * 3, x + 12);
        }
        if (A.k_x) {   // input row for the weight-gradient pass, chunk-major (act_off);
                       // row 39 (the K padding) carries the constant 1 that turns the bias gradient into a GEMM column
            float* kx = A.k_x + act_off(s, 0, 40);
#pragma unroll
            for (int i = 0; i < 39; ++i) kx[i * 16] = x[i];
            kx[39 * 16] = 1.0f;
        }
    };
    auto out = [&](int64_t s, const float* raw) {
#pragma unroll
        for (int j = 0; j < 3; ++j) A.k_rgb[s * 3 + j] = 1.0f / (1.0f + expf(-raw[j]));
    };
    auto act = [&](int64_t s, bool valid, int layer, int c, const float* h, unsigned char* stage) {
        if (A.k_mask) {   // ReLU sign bits for the backward: bit i of word (layer*4 + c/32) = h[c+i] > 0
            // h >= 0 here, so h > 0 <=> its bit pattern is >= 1 <=> bit 31 of (bits + 0x7fffffff): two integer ops per value
            uint32_t m = 0;
#pragma unroll
            for (int i = 31; i >= 0; --i) m = __funnelshift_l(__float_as_uint(h[i]) + 0x7fffffffu, m, 1);
            A.k_mask[(s >> 7) * (8 * 128) + (layer * 4 + (c >> 5)) * 128 + (s & 127)] = valid ? m : 0u;   // [tile][8][128]
        }
        float* dst = layer == 0 ? A.k_h0 : A.k_h1;
        if (dst) stage_store32(stage, dst, s - (threadIdx.x & 31), c, WD, h, valid);      // chunk-major (act_off)
    };
#ifdef PVDB_TC_TIMING
    if (A.exp) {      // experiment: touch the first tile's store targets ahead of the stores (1: L2 prefetch, 2: loads)
        const size_t o = (size_t)blockIdx.x * 128 * WD + (size_t)threadIdx.x * 32;
        if (threadIdx.x < 512 && A.exp == 1) {
            asm volatile("prefetch.global.L2 [%0];" ::"l"(A.k_h0 + o));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(A.k_h1 + o));
        } else if (A.exp >= 3) {      // 3: stagger the CTAs in four phases of ~2000 cycles; 4: eight phases of ~1200
            const long long t0 = clock64();
            const long long d = A.exp == 3 ? (blockIdx.x & 3) * 2000 : (blockIdx.x & 7) * 1200;
            while (clock64() - t0 < d) { }
        } else if (threadIdx.x < 32 && A.exp == 2) {
            float v = __ldcg(A.k_h0 + o) + __ldcg(A.k_h1 + o);
            if (v == 1234.5f) A.k_rgb[0] = v;
        }
    }
#endif
    mlp_tiles(smem, A.img, A.counters + CNT_M_KEEP, A.cap_keep, feat, out, act);
#ifdef PVDB_TC_TIMING
    if (threadIdx.x == 0) g_tc_gt[blockIdx.x][1] = gtimer();
#endif
}

// ---- merged renderer MLP (renderer.cu:83-119): features from the gathered list, PE from the pixel's view direction
__global__ void __launch_bounds__(FWD_THREADS, 1) k_render_mlp_tc(RenderMlpArgs A) {
    extern __shared__ __align__(128) unsigned char smem[];
    auto feat = [&](int64_t s, bool valid, float* x) {
        if (!valid) return;
        const float4* f = reinterpret_cast<const float4*>(A.s_feat + s * 12);
#pragma unroll
        for (int c4 = 0; c4 < 3; ++c4) {
            const float4 a = __ldg(f + c4);
            x[c4 * 4] = a.x; x[c4 * 4 + 1] = a.y; x[c4 * 4 + 2] = a.z; x[c4 * 4 + 3] = a.w;
        }
        Ray R;
        ray_setup(A.C, A.c2w, render_gpix(A.C, A.row_begin, A.s_ray[s]), R);
        x[12] = R.vd[0]; x[13] = R.vd[1]; x[14] = R.vd[2];
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                const float v = __fmul_rn(R.vd[a], (float)(1 << k));
                sincosf(v, &x[12 + 3 + k + 4 * a], &x[12 + 15 + k + 4 * a]);      // one range reduction for both
            }
    };
    auto out = [&](int64_t s, const float* raw) {
        const float w = A.s_weight[s];
#pragma unroll
        for (int j = 0; j < 3; ++j) A.s_rgb[s * 3 + j] = w / (1 + expf(-raw[j]));   // final_render (:115-117)
    };
    auto act = [&](int64_t, bool, int, int, const float*, unsigned char*) {};
    mlp_tiles(smem, A.img, A.counters, A.cap, feat, out, act);
}

}  // namespace

#ifdef PVDB_TC_TIMING
extern "C" int pvdb_debug_tc_gt(unsigned long long* out) {
    PVDB_CUDA(cudaMemcpyFromSymbol(out, g_tc_gt, sizeof(unsigned long long) * PVDB_SMS * 2));
    return PVDB_OK;
}
extern "C" int pvdb_debug_tc_timing(long long* out) {
    PVDB_CUDA(cudaMemcpyFromSymbol(out, g_tc_t, sizeof(long long) * PVDB_SMS * 8 * 16));
    return PVDB_OK;
}
#endif

// Weight images of the forward and of the activation-gradient kernel (independent of the march: the fused step runs it on
// a side stream underneath the count pass).
int pvdb_rgbnet_prep_tc(const pvdb_train_cfg* cfg, const pvdb_train_bufs* b, const float* viewdirs, int n_rays, cudaStream_t st) {
    PVDB_CHECK_ARG(b->net_img, "net_img scratch missing (tensor-core rgbnet)");
    const float* net = b->net;   // PyTorch layout: W[n][k] row-major
    TcWeights W;
    W.w0 = net + PVDB_NET_OFF_W0; W.w0_sn = PVDB_NET_DIN; W.w0_sk = 1;
    W.w1 = net + PVDB_NET_OFF_W1; W.w1_sn = WD; W.w1_sk = 1;
    W.w2 = net + PVDB_NET_OFF_W2; W.w2_sn = WD; W.w2_sk = 1;
    W.b0 = net + PVDB_NET_OFF_B0; W.b1 = net + PVDB_NET_OFF_B1; W.b2 = net + PVDB_NET_OFF_B2;
    k_prep_fwd_image<<<PVDB_SMS, 256, 0, st>>>(W, static_cast<unsigned char*>(b->net_img), b->net);
    PVDB_LAUNCH_CHECK();
    if (b->ray_pe && viewdirs && n_rays > 0) {
        k_ray_pe<<<(n_rays + 127) / 128, 128, 0, st>>>(viewdirs, n_rays, b->ray_pe);
        PVDB_LAUNCH_CHECK();
    }
    return PVDB_OK;
}

// Forward over the kept list; the weight images must be current (pvdb_rgbnet_prep_tc).
int pvdb_rgbnet_forward_tc(const pvdb_train_cfg* cfg, const pvdb_train_bufs* b, const float* viewdirs, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        PVDB_CUDA(cudaFuncSetAttribute(k_rgbnet_fwd_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL));
        attr_set = true;
    }
    PVDB_CHECK_ARG(b->net_img && b->k_corner, "net_img / k_corner scratch missing (tensor-core rgbnet)");
    TrainFwdArgs A;
    A.tree = *b->tree; A.k0 = b->k0; A.viewdirs = viewdirs; A.k_ray = b->k_ray; A.k_xyz = b->k_xyz; A.k_feat = b->k_feat;
    A.k_h0 = b->k_h0; A.k_h1 = b->k_h1; A.k_rgb = b->k_rgb; A.k_x = b->k_x; A.k_mask = b->k_mask; A.counters = b->counters; A.cap_keep = b->cap_keep;
    A.k_corner = b->k_corner;
    A.img = static_cast<const unsigned char*>(b->net_img);
    A.feat_ready = pvdb_leaf_local_enabled(b) ? 1 : 0;
    A.ray_pe = b->ray_pe;
#ifdef PVDB_TC_TIMING
    A.exp = getenv("PVDB_EXP") ? atoi(getenv("PVDB_EXP")) : 0;
#endif
    PVDB_CUDA(pvdb_launch_pdl(k_rgbnet_fwd_tc, dim3(PVDB_SMS), dim3(FWD_THREADS), SM_TOTAL, st, A));
    PVDB_LAUNCH_CHECK();
    return PVDB_OK;
}

int pvdb_render_mlp_tc(const void* render_mlp_args, cudaStream_t st) {
    const RenderMlpArgs& A = *static_cast<const RenderMlpArgs*>(render_mlp_args);
    static bool attr_set = false;
    if (!attr_set) {
        PVDB_CUDA(cudaFuncSetAttribute(k_render_mlp_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL));
        attr_set = true;
    }
    TcWeights W;   // MGRenderer::load_params passes transposed weights (run.py:98-104): w0[39][128], w1[128][128], w2[128][3]
    W.w0 = A.w0; W.w0_sn = 1; W.w0_sk = WD;
    W.w1 = A.w1; W.w1_sn = 1; W.w1_sk = WD;
    W.w2 = A.w2; W.w2_sn = 1; W.w2_sk = 3;
    W.b0 = A.b0; W.b1 = A.b1; W.b2 = A.b2;
    PVDB_CHECK_ARG(A.img, "w_img scratch missing (tensor-core rgbnet)");
    PVDB_CUDA(pvdb_launch_pdl(k_prep_fwd_image, dim3(PVDB_SMS), dim3(256), 0, st, W, A.img, (const float*)nullptr));
    PVDB_LAUNCH_CHECK();
    PVDB_CUDA(pvdb_launch_pdl(k_render_mlp_tc, dim3(PVDB_SMS), dim3(FWD_THREADS), SM_TOTAL, st, A));
    PVDB_LAUNCH_CHECK();
    return PVDB_OK;
}
