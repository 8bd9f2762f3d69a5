// rgbnet.cu — k0 feature gather + rgbnet MLP (Linear(39,128) ReLU Linear(128,128) ReLU Linear(128,3), sigmoid)
// over the compacted kept-sample list, forward and backward (dvgo.py:338-360, 99-107).
//
// This file is the fp32 CUDA-core path (cfg.use_tensor_cores == 0): persistent CTAs, one 64-sample tile at a
// time, weights resident in shared memory, 4x8 register micro-tiles.  Weight gradients are accumulated in
// registers across all tiles a CTA processes and flushed once.  The k0 gradient is scattered with 16-byte
// vector reductions (12 channels = 3 x red.v4 per corner) instead of the reference's 96 scalar atomics per
// sample (colorvdb.cu:28-37, 171-172).
#include "rgbnet.cuh"
#include "common.cuh"
#include "ray_math.cuh"
#include "mlp_tile.cuh"

namespace {

constexpr int CNT_M_KEEP = 1;

// View embedding (dvgo.py:354-356): [d, sin(d_a*2^k) a-major, cos(...)]; torch computes d*freq in fp32, then sin/cos.
__device__ __forceinline__ void view_embed(const float* __restrict__ vd, float* __restrict__ out /*27*/) {
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

// One float4 chunk (channels c4*4..+3) of the 12-channel trilinear sample, colorvdb.cu:81-111 arithmetic and order.
__device__ __forceinline__ float4 k0_gather4(const pvdb_tree& t, const float* __restrict__ k0, float x, float y, float z, int c4) {
    PvdbTri tri;
    tri.set(x, y, z);
    PvdbLeafCache cache;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const int dx = PVDB_CORNER[q][0], dy = PVDB_CORNER[q][1], dz = PVDB_CORNER[q][2];
        const int cx = tri.i + dx, cy = tri.j + dy, cz = tri.k + dz;
        const int leaf = cache.find(t, cx, cy, cz);
        if (leaf < 0) continue;
        const float sc = __fmul_rn(__fmul_rn(tri.f(0, dx), tri.f(1, dy)), tri.f(2, dz));
        const float4 v = __ldg(reinterpret_cast<const float4*>(k0 + ((size_t)leaf * 512 + pvdb_leaf_off(cx, cy, cz)) * 12) + c4);
        acc.x = __fmaf_rn(sc, v.x, acc.x); acc.y = __fmaf_rn(sc, v.y, acc.y);
        acc.z = __fmaf_rn(sc, v.z, acc.z); acc.w = __fmaf_rn(sc, v.w, acc.w);
    }
    return acc;
}

struct NetFwdArgs {
    pvdb_tree tree;
    const float* k0; const float* net; const float* viewdirs;
    const int32_t* k_ray; const float* k_xyz;
    float *k_feat, *k_h0, *k_h1, *k_rgb;
    const int32_t* counters; int64_t cap_keep;
    int save_act;
};

__global__ void __launch_bounds__(NT, 1) k_rgbnet_fwd(NetFwdArgs A) {
    extern __shared__ __align__(16) float smem[];
    float* sW0t = smem;                    // [KX][128]   W0t[i][j] = w0[j][i]
    float* sW1t = sW0t + KX * W;           // [128][128]  W1t[i][j] = w1[j][i]
    float* sW2 = sW1t + W * W;             // [3][128]
    float* sb0 = sW2 + 3 * W;              // [128]
    float* sb1 = sb0 + W;                  // [128]
    float* sb2 = sb1 + W;                  // [4]
    float* sX = sb2 + 4;                   // [TS][LDX]
    float* sH0 = sX + TS * LDX;            // [TS][LDH]
    float* sH1 = sH0 + TS * LDH;           // [TS][LDH]
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const float* net = A.net;
    for (int e = tid; e < KX * W; e += NT) { const int i = e / W, j = e % W; sW0t[e] = i < DIN ? __ldg(net + PVDB_NET_OFF_W0 + j * DIN + i) : 0.f; }
    for (int e = tid; e < W * W; e += NT) { const int i = e / W, j = e % W; sW1t[e] = __ldg(net + PVDB_NET_OFF_W1 + j * W + i); }
    for (int e = tid; e < 3 * W; e += NT) sW2[e] = __ldg(net + PVDB_NET_OFF_W2 + e);
    if (tid < W) { sb0[tid] = __ldg(net + PVDB_NET_OFF_B0 + tid); sb1[tid] = __ldg(net + PVDB_NET_OFF_B1 + tid); }
    if (tid < 3) sb2[tid] = __ldg(net + PVDB_NET_OFF_B2 + tid);
    const int64_t M = min((int64_t)A.counters[CNT_M_KEEP], A.cap_keep);
    const int64_t n_tiles = (M + TS - 1) / TS;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t s0 = tile * TS;
        __syncthreads();   // previous tile done with sX/sH*; weights visible on the first pass
        // ---- features: 3 threads per sample gather k0 (one float4 each), 64 threads build the view PE
        if (tid < 192) {
            const int s = tid / 3, c4 = tid % 3;
            float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
            if (s0 + s < M) {
                const float* p = A.k_xyz + (s0 + s) * 3;
                f = k0_gather4(A.tree, A.k0, p[0], p[1], p[2], c4);
                *reinterpret_cast<float4*>(A.k_feat + (s0 + s) * 12 + c4 * 4) = f;
            }
            *reinterpret_cast<float4*>(sX + s * LDX + c4 * 4) = f;
        } else {
            const int s = tid - 192;
            float pe[27];
            if (s0 + s < M) view_embed(A.viewdirs + (size_t)A.k_ray[s0 + s] * 3, pe);
            else {
#pragma unroll
                for (int i = 0; i < 27; ++i) pe[i] = 0.f;
            }
#pragma unroll
            for (int i = 0; i < 27; ++i) sX[s * LDX + 12 + i] = pe[i];
            sX[s * LDX + 39] = 0.f;
        }
        __syncthreads();
        float acc[4][8];
        // ---- layer 0
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int c = 0; c < 8; ++c) acc[i][c] = sb0[(c < 4 ? 0 : 64) + tx * 4 + (c & 3)];
        gemm_4x8<KX, LDX, W>(sX, sW0t, ty, tx, acc);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float4 o0 = make_float4(fmaxf(acc[i][0], 0.f), fmaxf(acc[i][1], 0.f), fmaxf(acc[i][2], 0.f), fmaxf(acc[i][3], 0.f));
            float4 o1 = make_float4(fmaxf(acc[i][4], 0.f), fmaxf(acc[i][5], 0.f), fmaxf(acc[i][6], 0.f), fmaxf(acc[i][7], 0.f));
            *reinterpret_cast<float4*>(sH0 + (ty * 4 + i) * LDH + tx * 4) = o0;
            *reinterpret_cast<float4*>(sH0 + (ty * 4 + i) * LDH + 64 + tx * 4) = o1;
            if (A.save_act && s0 + ty * 4 + i < M) {
                float* g = A.k_h0 + (s0 + ty * 4 + i) * W;
                *reinterpret_cast<float4*>(g + tx * 4) = o0;
                *reinterpret_cast<float4*>(g + 64 + tx * 4) = o1;
            }
        }
        __syncthreads();
        // ---- layer 1
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int c = 0; c < 8; ++c) acc[i][c] = sb1[(c < 4 ? 0 : 64) + tx * 4 + (c & 3)];
        gemm_4x8<W, LDH, W>(sH0, sW1t, ty, tx, acc);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float4 o0 = make_float4(fmaxf(acc[i][0], 0.f), fmaxf(acc[i][1], 0.f), fmaxf(acc[i][2], 0.f), fmaxf(acc[i][3], 0.f));
            float4 o1 = make_float4(fmaxf(acc[i][4], 0.f), fmaxf(acc[i][5], 0.f), fmaxf(acc[i][6], 0.f), fmaxf(acc[i][7], 0.f));
            *reinterpret_cast<float4*>(sH1 + (ty * 4 + i) * LDH + tx * 4) = o0;
            *reinterpret_cast<float4*>(sH1 + (ty * 4 + i) * LDH + 64 + tx * 4) = o1;
            if (A.save_act && s0 + ty * 4 + i < M) {
                float* g = A.k_h1 + (s0 + ty * 4 + i) * W;
                *reinterpret_cast<float4*>(g + tx * 4) = o0;
                *reinterpret_cast<float4*>(g + 64 + tx * 4) = o1;
            }
        }
        __syncthreads();
        // ---- layer 2 + sigmoid
        if (tid < 192) {
            const int s = tid / 3, j = tid % 3;
            float a = sb2[j];
#pragma unroll 8
            for (int i = 0; i < W; ++i) a = fmaf(sH1[s * LDH + i], sW2[j * W + i], a);
            if (s0 + s < M) A.k_rgb[(s0 + s) * 3 + j] = 1.0f / (1.0f + expf(-a));
        }
    }
}

struct NetBwdArgs {
    pvdb_tree tree;
    const float* net; const float* viewdirs;
    const int32_t* k_ray; const float* k_xyz; const float* k_feat; const float* k_h0; const float* k_h1;
    const float* k_glogit;     // [M3][3]
    float* net_grad; float* k0_grad; int32_t* k0_touched; int32_t* k0_touched_list; int32_t* counters_w;
    const int32_t* counters; int64_t cap_keep;
};

__global__ void __launch_bounds__(NT, 1) k_rgbnet_bwd(NetBwdArgs A) {
    extern __shared__ __align__(16) float smem[];
    float* sW1 = smem;                     // [128 j][128 i]  natural w1[j][i]
    float* sW0 = sW1 + W * W;              // [128 j][KX]     natural w0[j][i], cols 39.. zero
    float* sW2 = sW0 + W * KX;             // [3][128]
    float* sG = sW2 + 3 * W;               // [TS][4]
    float* sX = sG + TS * 4;               // [TS][LDX]
    float* sH0 = sX + TS * LDX;            // [TS][LDH]
    float* sD1 = sH0 + TS * LDH;           // [TS][LDH]  h1, then dH1 (masked)
    float* sD0 = sD1 + TS * LDH;           // [TS][LDH]  dH0 (masked)
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const float* net = A.net;
    for (int e = tid; e < W * W; e += NT) sW1[e] = __ldg(net + PVDB_NET_OFF_W1 + e);
    for (int e = tid; e < W * KX; e += NT) { const int j = e / KX, i = e % KX; sW0[e] = i < DIN ? __ldg(net + PVDB_NET_OFF_W0 + j * DIN + i) : 0.f; }
    for (int e = tid; e < 3 * W; e += NT) sW2[e] = __ldg(net + PVDB_NET_OFF_W2 + e);

    // weight-gradient accumulators, live across all tiles of this CTA
    float gW1[8][8];    // rows j = ty*8+a, cols i = tx*4+(b&3) + 64*(b>>2)
    float gW0[20];      // row j = tid>>1, cols i = (tid&1)*20 + c
    float gW2[3] = {0, 0, 0}, gb1 = 0, gb0 = 0, gb2 = 0;   // tid<128: column j = tid of w2 rows / biases; gb2: tid<3
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = 0; b < 8; ++b) gW1[a][b] = 0.f;
#pragma unroll
    for (int c = 0; c < 20; ++c) gW0[c] = 0.f;

    const int64_t M = min((int64_t)A.counters[CNT_M_KEEP], A.cap_keep);
    const int64_t n_tiles = (M + TS - 1) / TS;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t s0 = tile * TS;
        __syncthreads();
        // ---- load tile: g_logit, x = [k0 feat, PE], h0, h1  (rows past M are zero)
        if (tid < TS) {
            const bool ok = s0 + tid < M;
            sG[tid * 4 + 0] = ok ? A.k_glogit[(s0 + tid) * 3 + 0] : 0.f;
            sG[tid * 4 + 1] = ok ? A.k_glogit[(s0 + tid) * 3 + 1] : 0.f;
            sG[tid * 4 + 2] = ok ? A.k_glogit[(s0 + tid) * 3 + 2] : 0.f;
            sG[tid * 4 + 3] = 0.f;
        } else if (tid < 2 * TS) {
            const int s = tid - TS;
            float pe[27];
            if (s0 + s < M) view_embed(A.viewdirs + (size_t)A.k_ray[s0 + s] * 3, pe);
            else {
#pragma unroll
                for (int i = 0; i < 27; ++i) pe[i] = 0.f;
            }
#pragma unroll
            for (int i = 0; i < 27; ++i) sX[s * LDX + 12 + i] = pe[i];
            sX[s * LDX + 39] = 0.f;
        } else if (tid < 2 * TS + 64) {
            // 64 threads x 3 float4 = k0 features of the tile
            const int s = tid - 2 * TS;
#pragma unroll
            for (int c4 = 0; c4 < 3; ++c4) {
                float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
                if (s0 + s < M) f = *reinterpret_cast<const float4*>(A.k_feat + (s0 + s) * 12 + c4 * 4);
                *reinterpret_cast<float4*>(sX + s * LDX + c4 * 4) = f;
            }
        }
        for (int e = tid; e < TS * (W / 4); e += NT) {
            const int s = e / (W / 4), c = (e % (W / 4)) * 4;
            float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
            if (s0 + s < M) {
                a0 = *reinterpret_cast<const float4*>(A.k_h0 + (s0 + s) * W + c);
                a1 = *reinterpret_cast<const float4*>(A.k_h1 + (s0 + s) * W + c);
            }
            *reinterpret_cast<float4*>(sH0 + s * LDH + c) = a0;
            *reinterpret_cast<float4*>(sD1 + s * LDH + c) = a1;
        }
        __syncthreads();
        // ---- dW2 / db2 (needs h1 before it is overwritten)
        if (tid < W) {
#pragma unroll 4
            for (int s = 0; s < TS; ++s) {
                const float h = sD1[s * LDH + tid];
                gW2[0] = fmaf(sG[s * 4 + 0], h, gW2[0]); gW2[1] = fmaf(sG[s * 4 + 1], h, gW2[1]); gW2[2] = fmaf(sG[s * 4 + 2], h, gW2[2]);
            }
        } else if (tid < W + 3) {
            const int c = tid - W;
            for (int s = 0; s < TS; ++s) gb2 += sG[s * 4 + c];
        }
        __syncthreads();
        // ---- dH1 = (g_logit . W2) masked by h1 > 0, in place
        for (int e = tid; e < TS * W; e += NT) {
            const int s = e / W, j = e % W;
            const float h = sD1[s * LDH + j];
            const float d = fmaf(sG[s * 4 + 2], sW2[2 * W + j], fmaf(sG[s * 4 + 1], sW2[W + j], sG[s * 4] * sW2[j]));
            sD1[s * LDH + j] = h > 0.f ? d : 0.f;
        }
        __syncthreads();
        // ---- dW1[j][i] += sum_s dH1[s][j] * h0[s][i]   (8x8 per thread, K = TS), db1
#pragma unroll 2
        for (int s = 0; s < TS; ++s) {
            const float4 a0 = *reinterpret_cast<const float4*>(sD1 + s * LDH + ty * 8);
            const float4 a1 = *reinterpret_cast<const float4*>(sD1 + s * LDH + ty * 8 + 4);
            const float4 b0 = *reinterpret_cast<const float4*>(sH0 + s * LDH + tx * 4);
            const float4 b1 = *reinterpret_cast<const float4*>(sH0 + s * LDH + 64 + tx * 4);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int a = 0; a < 8; ++a)
#pragma unroll
                for (int b = 0; b < 8; ++b) gW1[a][b] = fmaf(av[a], bv[b], gW1[a][b]);
        }
        if (tid < W) {
            for (int s = 0; s < TS; ++s) gb1 += sD1[s * LDH + tid];
        }
        // ---- dH0 = (dH1 . W1) masked by h0 > 0
        {
            float acc[4][8];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int c = 0; c < 8; ++c) acc[i][c] = 0.f;
            gemm_4x8<W, LDH, W>(sD1, sW1, ty, tx, acc);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float4 h0a = *reinterpret_cast<const float4*>(sH0 + (ty * 4 + i) * LDH + tx * 4);
                const float4 h0b = *reinterpret_cast<const float4*>(sH0 + (ty * 4 + i) * LDH + 64 + tx * 4);
                float4 o0 = make_float4(h0a.x > 0.f ? acc[i][0] : 0.f, h0a.y > 0.f ? acc[i][1] : 0.f, h0a.z > 0.f ? acc[i][2] : 0.f,
                                        h0a.w > 0.f ? acc[i][3] : 0.f);
                float4 o1 = make_float4(h0b.x > 0.f ? acc[i][4] : 0.f, h0b.y > 0.f ? acc[i][5] : 0.f, h0b.z > 0.f ? acc[i][6] : 0.f,
                                        h0b.w > 0.f ? acc[i][7] : 0.f);
                *reinterpret_cast<float4*>(sD0 + (ty * 4 + i) * LDH + tx * 4) = o0;
                *reinterpret_cast<float4*>(sD0 + (ty * 4 + i) * LDH + 64 + tx * 4) = o1;
            }
        }
        __syncthreads();
        // ---- dW0[j][i] += sum_s dH0[s][j] * x[s][i]; db0
        {
            const int j = tid >> 1, ib = (tid & 1) * 20;
#pragma unroll 2
            for (int s = 0; s < TS; ++s) {
                const float a = sD0[s * LDH + j];
                const float* xr = sX + s * LDX + ib;
#pragma unroll
                for (int c4 = 0; c4 < 5; ++c4) {
                    const float4 xv = *reinterpret_cast<const float4*>(xr + c4 * 4);
                    gW0[c4 * 4 + 0] = fmaf(a, xv.x, gW0[c4 * 4 + 0]); gW0[c4 * 4 + 1] = fmaf(a, xv.y, gW0[c4 * 4 + 1]);
                    gW0[c4 * 4 + 2] = fmaf(a, xv.z, gW0[c4 * 4 + 2]); gW0[c4 * 4 + 3] = fmaf(a, xv.w, gW0[c4 * 4 + 3]);
                }
                if ((tid & 1) == 0) gb0 += a;
            }
        }
        // ---- dX[s][0..12) = sum_j dH0[s][j] * w0[j][i]; scatter to the k0 gradient plane (colorvdb.cu:130-160)
        if (tid < 192) {
            const int s = tid / 3, c4 = tid % 3;
            float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
            for (int j = 0; j < W; ++j) {
                const float a = sD0[s * LDH + j];
                const float4 w = *reinterpret_cast<const float4*>(sW0 + j * KX + c4 * 4);
                g.x = fmaf(a, w.x, g.x); g.y = fmaf(a, w.y, g.y); g.z = fmaf(a, w.z, g.z); g.w = fmaf(a, w.w, g.w);
            }
            if (s0 + s < M) {
                const float* p = A.k_xyz + (s0 + s) * 3;
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
                    red_add4(A.k0_grad + ((size_t)leaf * 512 + pvdb_leaf_off(cx, cy, cz)) * 12 + c4 * 4, __fmul_rn(g.x, sc),
                             __fmul_rn(g.y, sc), __fmul_rn(g.z, sc), __fmul_rn(g.w, sc));
                    if (c4 == 0) pvdb_touch_leaf(A.k0_touched, A.k0_touched_list, A.counters_w + 4, leaf);
                }
            }
        }
    }
    // ---- flush weight-gradient partials
    float* G = A.net_grad;
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            const int j = ty * 8 + a, i = tx * 4 + (b & 3) + 64 * (b >> 2);
            red_add(G + PVDB_NET_OFF_W1 + j * W + i, gW1[a][b]);
        }
    {
        const int j = tid >> 1, ib = (tid & 1) * 20;
#pragma unroll
        for (int c = 0; c < 20; ++c)
            if (ib + c < DIN) red_add(G + PVDB_NET_OFF_W0 + j * DIN + ib + c, gW0[c]);
        if ((tid & 1) == 0) red_add(G + PVDB_NET_OFF_B0 + j, gb0);
    }
    if (tid < W) {
        red_add(G + PVDB_NET_OFF_W2 + 0 * W + tid, gW2[0]); red_add(G + PVDB_NET_OFF_W2 + 1 * W + tid, gW2[1]);
        red_add(G + PVDB_NET_OFF_W2 + 2 * W + tid, gW2[2]);
        red_add(G + PVDB_NET_OFF_B1 + tid, gb1);
    } else if (tid < W + 3) {
        red_add(G + PVDB_NET_OFF_B2 + (tid - W), gb2);
    }
}

constexpr size_t FWD_SMEM = (size_t)(KX * W + W * W + 3 * W + W + W + 4 + TS * LDX + 2 * TS * LDH) * sizeof(float);
constexpr size_t BWD_SMEM = (size_t)(W * W + W * KX + 3 * W + TS * 4 + TS * LDX + 3 * TS * LDH) * sizeof(float);

}  // namespace

int pvdb_rgbnet_forward_tc(const pvdb_train_cfg* cfg, const pvdb_train_bufs* b, const float* viewdirs, cudaStream_t st);
int pvdb_rgbnet_prep_tc(const pvdb_train_cfg* cfg, const pvdb_train_bufs* b, const float* viewdirs, int n_rays, cudaStream_t st);
int pvdb_rgbnet_backward_tc(const pvdb_train_cfg* cfg, const pvdb_train_bufs* b, const float* viewdirs, cudaStream_t st);

int pvdb_rgbnet_forward(const pvdb_train_cfg* cfg, const pvdb_train_bufs* b, const float* viewdirs, cudaStream_t st) {
    if (pvdb_direct_colour(cfg)) return pvdb_direct_forward(b, st);
    if (cfg->use_tensor_cores) return pvdb_rgbnet_forward_tc(cfg, b, viewdirs, st);   // after pvdb_rgbnet_prepare
    static bool attr_set = false;
    if (!attr_set) {
        PVDB_CUDA(cudaFuncSetAttribute(k_rgbnet_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FWD_SMEM));
        attr_set = true;
    }
    NetFwdArgs A;
    A.tree = *b->tree; A.k0 = b->k0; A.net = b->net; A.viewdirs = viewdirs; A.k_ray = b->k_ray; A.k_xyz = b->k_xyz;
    A.k_feat = b->k_feat; A.k_h0 = b->k_h0; A.k_h1 = b->k_h1; A.k_rgb = b->k_rgb; A.counters = b->counters; A.cap_keep = b->cap_keep;
    A.save_act = (b->k_h0 && b->k_h1) ? 1 : 0;
    k_rgbnet_fwd<<<PVDB_SMS, NT, FWD_SMEM, st>>>(A);
    PVDB_LAUNCH_CHECK();
    return PVDB_OK;
}

int pvdb_rgbnet_backward_fp32(const pvdb_train_cfg* cfg, const pvdb_train_bufs* b, const float* viewdirs, cudaStream_t st) {
    PVDB_CHECK_ARG(b->k_h0 && b->k_h1, "the fp32 rgbnet backward needs the saved activations k_h0/k_h1");
    static bool attr_set = false;
    if (!attr_set) {
        PVDB_CUDA(cudaFuncSetAttribute(k_rgbnet_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BWD_SMEM));
        attr_set = true;
    }
    NetBwdArgs A;
    A.tree = *b->tree; A.net = b->net; A.viewdirs = viewdirs; A.k_ray = b->k_ray; A.k_xyz = b->k_xyz; A.k_feat = b->k_feat;
    A.k_h0 = b->k_h0; A.k_h1 = b->k_h1; A.k_glogit = b->k_rgb; A.net_grad = b->net_grad; A.k0_grad = b->k0_grad;
    A.k0_touched = b->k0_touched; A.k0_touched_list = b->k0_touched_list; A.counters_w = b->counters; A.counters = b->counters;
    A.cap_keep = b->cap_keep;
    k_rgbnet_bwd<<<PVDB_SMS, NT, BWD_SMEM, st>>>(A);
    PVDB_LAUNCH_CHECK();
    return PVDB_OK;
}

// Per-step preparation that does not depend on the samples (tensor-core path: weight images); a no-op for fp32.
int pvdb_rgbnet_prepare(const pvdb_train_cfg* cfg, const pvdb_train_bufs* b, const float* viewdirs, int n_rays, cudaStream_t st) {
    if (pvdb_direct_colour(cfg)) return PVDB_OK;
    if (cfg->use_tensor_cores) return pvdb_rgbnet_prep_tc(cfg, b, viewdirs, n_rays, st);
    return PVDB_OK;
}

int pvdb_rgbnet_backward(const pvdb_train_cfg* cfg, const pvdb_train_bufs* b, const float* viewdirs, cudaStream_t st) {
    if (pvdb_direct_colour(cfg)) return pvdb_direct_backward(b, st);
    if (cfg->use_tensor_cores) return pvdb_rgbnet_backward_tc(cfg, b, viewdirs, st);   // partials + reduce, added to net_grad
    return pvdb_rgbnet_backward_fp32(cfg, b, viewdirs, st);                             // atomics into net_grad (accumulating)
}
