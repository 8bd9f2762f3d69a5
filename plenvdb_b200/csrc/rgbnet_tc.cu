// rgbnet_tc.cu — the rgbnet MLP (39 -> 128 -> 128 -> 3) on the 5th-generation tensor cores (tcgen05).
//
// The only dense contraction of the hot path (dvgo.py:99-107, 354-360; renderer.cu:83-109).  One persistent CTA
// per SM, 128-sample tiles:
//   * accumulators live in TMEM (128 lanes = 128 samples, 128 fp32 columns);
//   * the activations never touch shared memory: each of the 128 threads owns one sample = one TMEM lane, reads
//     the accumulator row with tcgen05.ld, applies bias + ReLU, and writes the next layer's A operand straight
//     back into TMEM with tcgen05.st (A-from-TMEM form of tcgen05.mma);
//   * the weights stay resident in shared memory for the CTA's lifetime in the canonical K-major (no swizzle)
//     UMMA layout, read through shared-memory matrix descriptors;
//   * precision: 3xTF32 error-compensated products (x = hi + lo, D = Ahi*Bhi + Alo*Bhi + Ahi*Blo, fp32 accumulate)
//     -> ~2^-21 relative per product, which keeps rendered RGB within the 1e-5 parity tolerance where single TF32
//     (~2^-11) would not.
// One thread issues the MMAs; completion is signalled through tcgen05.commit -> mbarrier.
#include <cuda_fp16.h>
#include "common.cuh"
#include "ray_math.cuh"
#include "rgbnet.cuh"
#include "render_ray.cuh"
#include "tc_ptx.cuh"

namespace {

constexpr int WD = PVDB_NET_W;     // 128
constexpr int K0P = 40;            // layer-0 K (39 padded to a multiple of 8)
constexpr int N2P = 16;            // layer-2 N (3 padded to the UMMA minimum for M = 128)
constexpr int CNT_M_KEEP = 1;

// TMEM column map (512 columns allocated): A_hi [0,128) A_lo [128,256) D [256,384)
constexpr uint32_t COL_AHI = 0, COL_ALO = 128, COL_D = 256;

// shared memory map (bytes)
constexpr int SM_W0HI = 0;                                  // [128][40] tf32 hi, canonical K-major
constexpr int SM_W0LO = SM_W0HI + WD * K0P * 4;
constexpr int SM_W1HI = SM_W0LO + WD * K0P * 4;             // [128][128]
constexpr int SM_W1LO = SM_W1HI + WD * WD * 4;
constexpr int SM_W2HI = SM_W1LO + WD * WD * 4;              // [16][128]
constexpr int SM_W2LO = SM_W2HI + N2P * WD * 4;
constexpr int SM_B0 = SM_W2LO + N2P * WD * 4;               // [128] floats
constexpr int SM_B1 = SM_B0 + WD * 4;
constexpr int SM_B2 = SM_B1 + WD * 4;                       // [4]
constexpr int SM_BAR = SM_B2 + 16;                          // mbarrier (8 B) + tmem base (4 B)
constexpr int SM_XBUF = SM_BAR + 16;                        // staging rows of the next tile: [128][44] floats
constexpr int SM_TOTAL = SM_XBUF + 128 * 44 * 4;

// Issue the 3xTF32 MMAs of one layer: D[128 x N] = A[128 x K] * W[N x K]^T.  Single thread.
__device__ __forceinline__ void issue_layer(uint32_t tmem, uint32_t smem_base, int off_hi, int off_lo, int K, int N, uint32_t bar) {
    const uint32_t idesc = make_idesc(N);
    const uint64_t bhi = make_desc(smem_base + off_hi, K), blo = make_desc(smem_base + off_lo, K);
    uint32_t acc = 0;
    for (int ks = 0; ks < K / 8; ++ks) {
        // one k-step = 8 tf32 = two 16-byte chunks = 256 B further along the group: start address field is in 16-B units
        const uint64_t adv = (uint64_t)(ks * 256) >> 4;
        const uint32_t ahi = tmem + COL_AHI + ks * 8, alo = tmem + COL_ALO + ks * 8;
        umma_tf32_ts(tmem + COL_D, ahi, bhi + adv, idesc, acc);
        acc = 1;
        umma_tf32_ts(tmem + COL_D, alo, bhi + adv, idesc, 1);
        umma_tf32_ts(tmem + COL_D, ahi, blo + adv, idesc, 1);
    }
    umma_commit(bar);
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

struct TcWeights {
    const float *w0, *b0, *w1, *b1, *w2, *b2;
    int w0_sn, w0_sk, w1_sn, w1_sk, w2_sn, w2_sk;   // strides (floats) of W[n][k]
};

// Common body, warp specialised.  Warps 0-3 ("lane warps", one thread per TMEM lane = sample) run the MLP; warps 4-7
// ("producers") gather the NEXT tile's input rows (k0 trilinear / feature list + view PE, the memory- and SFU-latency
// bound part) into a shared staging buffer while the lane warps are busy with the current tile.
// FeatFn(s, x[40]) fills the input row of sample s; ActFn sees each post-ReLU chunk; OutFn(s, raw[3]) consumes the result.
constexpr int FWD_THREADS = 256;
constexpr int XLD = 44;   // staging row stride in floats: 16-byte row reads/writes by 8 consecutive threads hit distinct banks
__device__ __forceinline__ void lane_bar() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

template <class FeatFn, class OutFn, class ActFn>
__device__ void mlp_tiles(unsigned char* smem, const TcWeights& Wt, int64_t M, FeatFn feat, OutFn out, ActFn act) {
    const int tid = threadIdx.x, warp = tid >> 5;
    const bool lane_warp = tid < TM;
    const uint32_t sbase = smem_u32(smem);
    const uint32_t bar = sbase + SM_BAR;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + SM_BAR + 8);
    float* sb0 = reinterpret_cast<float*>(smem + SM_B0);
    float* sb1 = reinterpret_cast<float*>(smem + SM_B1);
    float* sb2 = reinterpret_cast<float*>(smem + SM_B2);
    float* xbuf = reinterpret_cast<float*>(smem + SM_XBUF);
    load_weight(smem, SM_W0HI, SM_W0LO, Wt.w0, Wt.w0_sn, Wt.w0_sk, WD, K0P, WD, PVDB_NET_DIN);
    load_weight(smem, SM_W1HI, SM_W1LO, Wt.w1, Wt.w1_sn, Wt.w1_sk, WD, WD, WD, WD);
    load_weight(smem, SM_W2HI, SM_W2LO, Wt.w2, Wt.w2_sn, Wt.w2_sk, N2P, WD, 3, WD);
    if (tid < WD) { sb0[tid] = __ldg(Wt.b0 + tid); sb1[tid] = __ldg(Wt.b1 + tid); }
    if (tid < 3) sb2[tid] = __ldg(Wt.b2 + tid);
    if (tid == 0) mbar_init(bar, 1);
    if (warp == 0) tmem_alloc(sbase + SM_BAR + 8, 512);
    fence_async_smem();           // weights written through the generic proxy, read by the tensor core (async proxy)
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t lane_addr = tmem + ((uint32_t)((warp & 3) * 32) << 16);   // this warp's 32-lane quarter
    uint32_t parity = 0;
    const int64_t n_tiles = (M + TM - 1) / TM;
    auto produce = [&](int64_t tile) {
        const int p = tid - TM;
        const int64_t s = tile * TM + p;
        float x[K0P];
#pragma unroll
        for (int i = 0; i < K0P; ++i) x[i] = 0.f;
        if (s < M) feat(s, x);
        float4* row = reinterpret_cast<float4*>(xbuf + p * XLD);
#pragma unroll
        for (int q = 0; q < K0P / 4; ++q) row[q] = make_float4(x[q * 4], x[q * 4 + 1], x[q * 4 + 2], x[q * 4 + 3]);
    };
    if (!lane_warp && (int64_t)blockIdx.x < n_tiles) produce(blockIdx.x);
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        __syncthreads();            // staging buffer holds this tile's rows
        float x[K0P];
        if (lane_warp) {
            const float4* row = reinterpret_cast<const float4*>(xbuf + tid * XLD);
#pragma unroll
            for (int q = 0; q < K0P / 4; ++q) { const float4 v = row[q]; x[q * 4] = v.x; x[q * 4 + 1] = v.y; x[q * 4 + 2] = v.z; x[q * 4 + 3] = v.w; }
        }
        __syncthreads();            // staging buffer is free again
        if (!lane_warp) {
            if (tile + gridDim.x < n_tiles) produce(tile + gridDim.x);
            continue;
        }
        const int64_t s = tile * TM + tid;
        const bool valid = s < M;
        // ---- layer-0 input row -> TMEM
        store_a_row(lane_addr, 0, x, K0P);
        tmem_st_wait();
        tc_fence_before();
        lane_bar();
        if (tid == 0) { tc_fence_after(); issue_layer(tmem, sbase, SM_W0HI, SM_W0LO, K0P, WD, bar); }
        mbar_wait(bar, parity); parity ^= 1;
        tc_fence_after();
        // ---- hidden layers: D -> bias + ReLU -> next A
#pragma unroll 1
        for (int layer = 0; layer < 2; ++layer) {
            const float* sb = layer == 0 ? sb0 : sb1;
#pragma unroll 1
            for (int c = 0; c < WD; c += 32) {
                uint32_t r[32];
                tmem_ld32(lane_addr + COL_D + c, r);
                tmem_ld_wait();
                float h[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) h[i] = fmaxf(__uint_as_float(r[i]) + sb[c + i], 0.f);
                if (valid) act(s, layer, c, h);
                store_a_row(lane_addr, c, h, 32);
            }
            tmem_st_wait();
            tc_fence_before();
            lane_bar();
            if (tid == 0) {
                tc_fence_after();
                if (layer == 0) issue_layer(tmem, sbase, SM_W1HI, SM_W1LO, WD, WD, bar);
                else issue_layer(tmem, sbase, SM_W2HI, SM_W2LO, WD, N2P, bar);
            }
            mbar_wait(bar, parity); parity ^= 1;
            tc_fence_after();
        }
        // ---- output layer
        {
            uint32_t r[16];
            tmem_ld16(lane_addr + COL_D, r);
            tmem_ld_wait();
            if (valid) {
                const float raw[3] = {__uint_as_float(r[0]) + sb2[0], __uint_as_float(r[1]) + sb2[1], __uint_as_float(r[2]) + sb2[2]};
                out(s, raw);
            }
        }
        tc_fence_before();
        lane_bar();          // every lane has drained D before the next tile's MMAs overwrite it
        tc_fence_after();
    }
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
    const int32_t* counters; int64_t cap_keep;
    TcWeights W;
};

__global__ void __launch_bounds__(FWD_THREADS, 1) k_rgbnet_fwd_tc(TrainFwdArgs A) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int64_t M = min((int64_t)A.counters[CNT_M_KEEP], A.cap_keep);
    auto feat = [&](int64_t s, float* x) {
        // 12-channel trilinear sample, colorvdb.cu:81-111 arithmetic and corner order
        const float* p = A.k_xyz + s * 3;
        PvdbTri tri;
        tri.set(p[0], p[1], p[2]);
        PvdbLeafCache cache;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int dx = PVDB_CORNER[q][0], dy = PVDB_CORNER[q][1], dz = PVDB_CORNER[q][2];
            const int cx = tri.i + dx, cy = tri.j + dy, cz = tri.k + dz;
            const int leaf = cache.find(A.tree, cx, cy, cz);
            if (leaf < 0) continue;
            const float sc = __fmul_rn(__fmul_rn(tri.f(0, dx), tri.f(1, dy)), tri.f(2, dz));
            const float4* v = reinterpret_cast<const float4*>(A.k0 + ((size_t)leaf * 512 + pvdb_leaf_off(cx, cy, cz)) * 12);
#pragma unroll
            for (int c4 = 0; c4 < 3; ++c4) {
                const float4 a = __ldg(v + c4);
                x[c4 * 4 + 0] = __fmaf_rn(sc, a.x, x[c4 * 4 + 0]); x[c4 * 4 + 1] = __fmaf_rn(sc, a.y, x[c4 * 4 + 1]);
                x[c4 * 4 + 2] = __fmaf_rn(sc, a.z, x[c4 * 4 + 2]); x[c4 * 4 + 3] = __fmaf_rn(sc, a.w, x[c4 * 4 + 3]);
            }
        }
        float4* kf = reinterpret_cast<float4*>(A.k_feat + s * 12);
        kf[0] = make_float4(x[0], x[1], x[2], x[3]); kf[1] = make_float4(x[4], x[5], x[6], x[7]); kf[2] = make_float4(x[8], x[9], x[10], x[11]);
        view_embed_tc(A.viewdirs + (size_t)A.k_ray[s] * 3, x + 12);
        if (A.k_x) {   // input row for the weight-gradient pass, tile-transposed [tile][40][128]: warp-coalesced stores
            float* kx = A.k_x + (s >> 7) * (40 * 128) + (s & 127);
#pragma unroll
            for (int i = 0; i < 40; ++i) kx[i * 128] = x[i];
        }
    };
    auto out = [&](int64_t s, const float* raw) {
#pragma unroll
        for (int j = 0; j < 3; ++j) A.k_rgb[s * 3 + j] = 1.0f / (1.0f + expf(-raw[j]));
    };
    auto act = [&](int64_t s, int layer, int c, const float* h) {
        if (A.k_mask) {   // ReLU sign bits for the backward: bit i of word (layer*4 + c/32) = h[c+i] > 0
            uint32_t m = 0;
#pragma unroll
            for (int i = 0; i < 32; ++i) m |= (h[i] > 0.f ? 1u : 0u) << i;
            A.k_mask[(s >> 7) * (8 * 128) + (layer * 4 + (c >> 5)) * 128 + (s & 127)] = m;   // [tile][8][128]
        }
        float* dst = layer == 0 ? A.k_h0 : A.k_h1;
        if (!dst) return;
        // tile-transposed [tile][128 features][128 samples]: lane l writes word l of a 128-byte line -> one wavefront per
        // store instead of 32, and the weight-gradient pass reads 4 consecutive samples of a feature as one LDG.128
        float* g = dst + (s >> 7) * (WD * 128) + (size_t)c * 128 + (s & 127);
#pragma unroll
        for (int i = 0; i < 32; ++i) g[i * 128] = h[i];
    };
    mlp_tiles(smem, A.W, M, feat, out, act);
}

// ---- merged renderer MLP (renderer.cu:83-119): features from the gathered list, PE from the pixel's view direction
__global__ void __launch_bounds__(FWD_THREADS, 1) k_render_mlp_tc(RenderMlpArgs A, TcWeights W) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int64_t M = min((int64_t)A.counters[0], A.cap);
    auto feat = [&](int64_t s, float* x) {
        const float4* f = reinterpret_cast<const float4*>(A.s_feat + s * 12);
#pragma unroll
        for (int c4 = 0; c4 < 3; ++c4) {
            const float4 a = __ldg(f + c4);
            x[c4 * 4] = a.x; x[c4 * 4 + 1] = a.y; x[c4 * 4 + 2] = a.z; x[c4 * 4 + 3] = a.w;
        }
        Ray R;
        ray_setup(A.C, A.c2w, A.row_begin * A.C.W + A.s_ray[s], R);
        x[12] = R.vd[0]; x[13] = R.vd[1]; x[14] = R.vd[2];
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                const float v = __fmul_rn(R.vd[a], (float)(1 << k));
                x[12 + 3 + k + 4 * a] = sinf(v);
                x[12 + 15 + k + 4 * a] = cosf(v);
            }
    };
    auto out = [&](int64_t s, const float* raw) {
        const float w = A.s_weight[s];
#pragma unroll
        for (int j = 0; j < 3; ++j) A.s_rgb[s * 3 + j] = w / (1 + expf(-raw[j]));   // final_render (:115-117)
    };
    auto act = [&](int64_t, int, int, const float*) {};
    mlp_tiles(smem, W, M, feat, out, act);
}

}  // namespace

int pvdb_rgbnet_forward_tc(const pvdb_train_cfg* cfg, const pvdb_train_bufs* b, const float* viewdirs, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        PVDB_CUDA(cudaFuncSetAttribute(k_rgbnet_fwd_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL));
        attr_set = true;
    }
    TrainFwdArgs A;
    A.tree = *b->tree; A.k0 = b->k0; A.viewdirs = viewdirs; A.k_ray = b->k_ray; A.k_xyz = b->k_xyz; A.k_feat = b->k_feat;
    A.k_h0 = b->k_h0; A.k_h1 = b->k_h1; A.k_rgb = b->k_rgb; A.k_x = b->k_x; A.k_mask = b->k_mask; A.counters = b->counters; A.cap_keep = b->cap_keep;
    const float* net = b->net;   // PyTorch layout: W[n][k] row-major
    A.W.w0 = net + PVDB_NET_OFF_W0; A.W.w0_sn = PVDB_NET_DIN; A.W.w0_sk = 1;
    A.W.w1 = net + PVDB_NET_OFF_W1; A.W.w1_sn = WD; A.W.w1_sk = 1;
    A.W.w2 = net + PVDB_NET_OFF_W2; A.W.w2_sn = WD; A.W.w2_sk = 1;
    A.W.b0 = net + PVDB_NET_OFF_B0; A.W.b1 = net + PVDB_NET_OFF_B1; A.W.b2 = net + PVDB_NET_OFF_B2;
    k_rgbnet_fwd_tc<<<PVDB_SMS, FWD_THREADS, SM_TOTAL, st>>>(A);
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
    k_render_mlp_tc<<<PVDB_SMS, FWD_THREADS, SM_TOTAL, st>>>(A, W);
    PVDB_LAUNCH_CHECK();
    return PVDB_OK;
}
