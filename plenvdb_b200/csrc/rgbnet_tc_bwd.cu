// rgbnet_tc_bwd.cu — rgbnet backward on tcgen05 (3xTF32), two kernels:
//
//  B1 k_rgbnet_bwd_act_tc   activation gradients, 128-sample tiles, one TMEM lane per sample (same structure as the
//                           forward):  dH1 = (g_logit . W2) * [h1>0]  (K = 3, CUDA cores, in registers)
//                                      dH0 = (dH1 . W1)   * [h0>0]  (tcgen05, A = dH1 in TMEM, B = W1^T in smem)
//                                      dX  =  dH0 . W0[:, :12]       (tcgen05, N = 16) -> k0 gradient scatter
//                           dH1, dH0 are also written to HBM for B2.
//  B2 k_rgbnet_bwd_wgrad_tc weight gradients = sums over samples of outer products:  dW1 = dH1^T H0, dW0 = dH0^T X,
//                           dW2 = G^T H1.  The contraction runs over SAMPLES, so both operands are activations
//                           transposed: they are transposed while being staged into shared memory (K-major canonical
//                           layout, 16-sample chunks, double buffered, hi/lo split on the fly) and consumed by
//                           tcgen05.mma SS.  Bias gradients ride along as an extra "ones"
//                           column of the B operand.  Accumulators live in TMEM for the CTA's lifetime and are
//                           flushed once with red.global.add.
// Reference semantics: autograd through nn.Sequential(Linear, ReLU, Linear, ReLU, Linear) (dvgo.py:99-107) and
// QueryVerticalInVDB.backward -> color_backward (grid.py:53-60, colorvdb.cu:130-175).
#include "common.cuh"
#include "rgbnet.cuh"
#include "tc_ptx.cuh"

namespace {

constexpr int WD = PVDB_NET_W;
constexpr int CNT_M_KEEP = 1;
constexpr uint32_t COL_AHI = 0, COL_ALO = 128, COL_D = 256;

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
// D[128 x N] = A(TMEM)[128 x K] * B(smem K-major)[N x K]^T, 3xTF32
__device__ __forceinline__ void issue_ts(uint32_t tmem, uint32_t smem_hi, uint32_t smem_lo, int K, int N, uint32_t bar) {
    const uint32_t idesc = make_idesc(N);
    const uint64_t bhi = make_desc(smem_hi, K), blo = make_desc(smem_lo, K);
    uint32_t acc = 0;
    for (int ks = 0; ks < K / 8; ++ks) {
        const uint64_t adv = (uint64_t)(ks * 256) >> 4;
        const uint32_t ahi = tmem + COL_AHI + ks * 8, alo = tmem + COL_ALO + ks * 8;
        umma_tf32_ts(tmem + COL_D, ahi, bhi + adv, idesc, acc);
        acc = 1;
        umma_tf32_ts(tmem + COL_D, alo, bhi + adv, idesc, 1);
        umma_tf32_ts(tmem + COL_D, ahi, blo + adv, idesc, 1);
    }
    umma_commit(bar);
}

// ---------------------------------------------------------------------------------------------- B1
constexpr int B1_W1HI = 0;                              // B[N=i][K=j] = w1[j][i], canonical K-major, [128][128]
constexpr int B1_W1LO = B1_W1HI + WD * WD * 4;
constexpr int B1_W0HI = B1_W1LO + WD * WD * 4;          // B[N=i<16][K=j] = w0[j][i], [16][128]
constexpr int B1_W0LO = B1_W0HI + 16 * WD * 4;
constexpr int B1_W2 = B1_W0LO + 16 * WD * 4;            // plain floats [3][128]
constexpr int B1_BAR = B1_W2 + 3 * WD * 4;
constexpr int B1_TOTAL = B1_BAR + 16;

struct BwdActArgs {
    pvdb_tree tree;
    const float* net;
    const float *k_glogit, *k_xyz;
    const uint32_t* k_mask;
    float *k_dh1, *k_dh0;
    float* k0_grad; int32_t* k0_touched; int32_t* k0_touched_list; int32_t* counters_w;
    const int32_t* counters; int64_t cap_keep;
};

__global__ void __launch_bounds__(TM, 1) k_rgbnet_bwd_act_tc(BwdActArgs A) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, warp = tid >> 5;
    const uint32_t sbase = smem_u32(smem);
    const uint32_t bar = sbase + B1_BAR;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + B1_BAR + 8);
    float* sW2 = reinterpret_cast<float*>(smem + B1_W2);
    const float* net = A.net;
    load_weight(smem, B1_W1HI, B1_W1LO, net + PVDB_NET_OFF_W1, /*sn(i)*/ 1, /*sk(j)*/ WD, WD, WD, WD, WD);
    load_weight(smem, B1_W0HI, B1_W0LO, net + PVDB_NET_OFF_W0, /*sn(i)*/ 1, /*sk(j)*/ PVDB_NET_DIN, 16, WD, 12, WD);
    for (int e = tid; e < 3 * WD; e += TM) sW2[e] = __ldg(net + PVDB_NET_OFF_W2 + e);
    if (tid == 0) mbar_init(bar, 1);
    if (warp == 0) tmem_alloc(sbase + B1_BAR + 8, 512);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);
    uint32_t parity = 0;
    const int64_t M = min((int64_t)A.counters[CNT_M_KEEP], A.cap_keep);
    const int64_t n_tiles = (M + TM - 1) / TM;
    // per-sample inputs are tiny (3 + 8 + 3 words): fetched one tile ahead so their latency hides behind the MMAs
    struct In { float g0, g1, g2, px, py, pz; uint4 m0, m1; };
    auto fetch = [&](int64_t tile) {
        In v;
        v.g0 = v.g1 = v.g2 = v.px = v.py = v.pz = 0.f;
        v.m0 = v.m1 = make_uint4(0, 0, 0, 0);
        const int64_t s = tile * TM + tid;
        if (tile < n_tiles && s < M) {
            v.g0 = __ldg(A.k_glogit + s * 3); v.g1 = __ldg(A.k_glogit + s * 3 + 1); v.g2 = __ldg(A.k_glogit + s * 3 + 2);
            v.px = __ldg(A.k_xyz + s * 3); v.py = __ldg(A.k_xyz + s * 3 + 1); v.pz = __ldg(A.k_xyz + s * 3 + 2);
            const uint32_t* mk = A.k_mask + (s >> 7) * (8 * 128) + (s & 127);   // [tile][8][128]
            v.m0 = make_uint4(__ldg(mk), __ldg(mk + 128), __ldg(mk + 256), __ldg(mk + 384));
            v.m1 = make_uint4(__ldg(mk + 512), __ldg(mk + 640), __ldg(mk + 768), __ldg(mk + 896));
        }
        return v;
    };
    In nxt = fetch(blockIdx.x);
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t s = tile * TM + tid;
        const bool valid = s < M;
        const In cur = nxt;
        nxt = fetch(tile + gridDim.x);
        const float g0 = cur.g0, g1 = cur.g1, g2 = cur.g2;
        const uint32_t m0w[4] = {cur.m0.x, cur.m0.y, cur.m0.z, cur.m0.w}, m1w[4] = {cur.m1.x, cur.m1.y, cur.m1.z, cur.m1.w};
        // ---- dH1 = (g . W2) masked by h1 > 0
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
            const int c = cc * 32;
            float d[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const int j = c + i;
                const float v = fmaf(g2, sW2[2 * WD + j], fmaf(g1, sW2[WD + j], g0 * sW2[j]));
                d[i] = (m1w[cc] >> i) & 1u ? v : 0.f;
            }
            if (valid) {   // tile-transposed [tile][128][128], warp-coalesced
                float* o = A.k_dh1 + (s >> 7) * (WD * 128) + (size_t)c * 128 + (s & 127);
#pragma unroll
                for (int i = 0; i < 32; ++i) o[i * 128] = d[i];
            }
            store_a_row32(lane_addr, c, d);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncthreads();
        if (tid == 0) { tc_fence_after(); issue_ts(tmem, sbase + B1_W1HI, sbase + B1_W1LO, WD, WD, bar); }
        mbar_wait(bar, parity); parity ^= 1;
        tc_fence_after();
        // ---- dH0 = D masked by h0 > 0
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
            const int c = cc * 32;
            uint32_t r[32];
            tmem_ld32(lane_addr + COL_D + c, r);
            tmem_ld_wait();
            float d[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) d[i] = (m0w[cc] >> i) & 1u ? __uint_as_float(r[i]) : 0.f;
            if (valid) {
                float* o = A.k_dh0 + (s >> 7) * (WD * 128) + (size_t)c * 128 + (s & 127);
#pragma unroll
                for (int i = 0; i < 32; ++i) o[i * 128] = d[i];
            }
            store_a_row32(lane_addr, c, d);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncthreads();
        if (tid == 0) { tc_fence_after(); issue_ts(tmem, sbase + B1_W0HI, sbase + B1_W0LO, WD, 16, bar); }
        mbar_wait(bar, parity); parity ^= 1;
        tc_fence_after();
        // ---- dX[0..12) -> k0 gradient scatter (colorvdb.cu:130-160), 3 x red.v4 per corner
        {
            uint32_t r[16];
            tmem_ld16(lane_addr + COL_D, r);
            tmem_ld_wait();
            if (valid) {
                PvdbTri tri;
                tri.set(cur.px, cur.py, cur.pz);
                PvdbLeafCache cache;
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const int dx = PVDB_CORNER[q][0], dy = PVDB_CORNER[q][1], dz = PVDB_CORNER[q][2];
                    const int cx = tri.i + dx, cy = tri.j + dy, cz = tri.k + dz;
                    const int leaf = cache.find(A.tree, cx, cy, cz);
                    if (leaf < 0) continue;
                    const float sc = __fmul_rn(__fmul_rn(tri.f(0, dx), tri.f(1, dy)), tri.f(2, dz));
                    float* dst = A.k0_grad + ((size_t)leaf * 512 + pvdb_leaf_off(cx, cy, cz)) * 12;
#pragma unroll
                    for (int c4 = 0; c4 < 3; ++c4)
                        red_add4(dst + c4 * 4, __fmul_rn(__uint_as_float(r[c4 * 4]), sc), __fmul_rn(__uint_as_float(r[c4 * 4 + 1]), sc),
                                 __fmul_rn(__uint_as_float(r[c4 * 4 + 2]), sc), __fmul_rn(__uint_as_float(r[c4 * 4 + 3]), sc));
                    pvdb_touch_leaf(A.k0_touched, A.k0_touched_list, A.counters_w + 4, leaf);
                }
            }
        }
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
    }
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

// ---------------------------------------------------------------------------------------------- B2
// Both operands are transposed activations (contraction over samples).  MN-major tf32 operands only exist in the
// 128B_BASE32B swizzled layout (CUTLASS sm100_common.inl: "for mn-major tf32 operands, SW128_32B is the only available
// smem layout"), so the tiles are transposed while they are staged instead: thread t owns FEATURE t, reads it for 4
// consecutive samples (each read is a coalesced 128-byte row segment across the warp) and writes one 16-byte
// K-major core-matrix row -> the same canonical no-swizzle K-major layout the forward uses, conflict free.
constexpr int KC = 16;                                   // samples per chunk = 2 k-steps of 8
constexpr int N1 = 144, N0 = 48, N2 = 16;                // padded N of the three GEMMs (128+1 bias, 39+1 bias, 3)
constexpr int OPB(int mn) { return mn * KC * 4; }        // bytes of one operand tile [mn][KC]
constexpr int S_A1 = 0;                                  // dH1^T [128][KC] hi, lo
constexpr int S_B1 = S_A1 + 2 * OPB(WD);                 // [H0^T ; 1 ; 0..] [144][KC] hi, lo
constexpr int S_A0 = S_B1 + 2 * OPB(N1);                 // dH0^T [128][KC]
constexpr int S_B0 = S_A0 + 2 * OPB(WD);                 // [X^T(39) ; 1 ; 0..] [48][KC]
constexpr int S_A2 = S_B0 + 2 * OPB(N0);                 // H1^T [128][KC]
constexpr int S_B2 = S_A2 + 2 * OPB(WD);                 // G^T [16][KC]
constexpr int STAGE = S_B2 + 2 * OPB(N2);
constexpr int B2_BAR = 2 * STAGE;                        // two mbarriers + tmem slot
constexpr int B2_TOTAL = B2_BAR + 32;
constexpr uint32_t ACC1 = 0, ACC0 = 144, ACC2 = 192;     // TMEM columns of the three accumulators (256 allocated)

// 4 consecutive samples (k4*4 .. +3) of feature `mn` -> one 16-byte row of a K-major core matrix, hi and lo tiles
__device__ __forceinline__ void put_k4(unsigned char* base, int off_hi, int bytes_tile, int mn, int k4, float4 v) {
    uint32_t h[4], l[4];
    split_tf32(v.x, h[0], l[0]); split_tf32(v.y, h[1], l[1]); split_tf32(v.z, h[2], l[2]); split_tf32(v.w, h[3], l[3]);
    const int o = canon_off(mn, k4 * 4, KC);
    *reinterpret_cast<uint4*>(base + off_hi + o) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(base + off_hi + bytes_tile + o) = make_uint4(l[0], l[1], l[2], l[3]);
}
// feature `f` of 4 consecutive samples s0..s0+3 (s0 % 4 == 0) of a tile-transposed array [tile][nf][128]: one LDG.128.
// Samples past M read as zero (their slots may hold stale data from an earlier, larger batch).
__device__ __forceinline__ float4 col4(const float* __restrict__ a, int64_t s0, int64_t M, int nf, int f) {
    if (s0 >= M) return make_float4(0.f, 0.f, 0.f, 0.f);
    float4 v = __ldg(reinterpret_cast<const float4*>(a + (s0 >> 7) * (nf * 128) + (size_t)f * 128 + (s0 & 127)));
    if (s0 + 1 >= M) v.y = 0.f;
    if (s0 + 2 >= M) v.z = 0.f;
    if (s0 + 3 >= M) v.w = 0.f;
    return v;
}
// same for the row-major [.][3] logit gradients
__device__ __forceinline__ float4 col4_rows(const float* __restrict__ a, int64_t s0, int64_t M, int ld, int f) {
    float4 v;
    v.x = s0 + 0 < M ? __ldg(a + (s0 + 0) * ld + f) : 0.f;
    v.y = s0 + 1 < M ? __ldg(a + (s0 + 1) * ld + f) : 0.f;
    v.z = s0 + 2 < M ? __ldg(a + (s0 + 2) * ld + f) : 0.f;
    v.w = s0 + 3 < M ? __ldg(a + (s0 + 3) * ld + f) : 0.f;
    return v;
}

struct BwdWgradArgs {
    const float *k_dh1, *k_h0, *k_dh0, *k_x, *k_h1, *k_glogit;
    float* net_grad;
    const int32_t* counters; int64_t cap_keep;
};

constexpr int B2_THREADS = 512;   // 16 warps stage (memory-latency bound); warps 0-3 own the TMEM lanes for the flush
__global__ void __launch_bounds__(B2_THREADS, 1) k_rgbnet_bwd_wgrad_tc(BwdWgradArgs A) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, warp = tid >> 5;
    // feature row and 4-sample group this thread stages.  Per warp: 8 features x 4 groups; the 8 threads of a quarter warp
    // take 8 different features (conflict-free 16-byte shared stores), the 4 groups of a feature read 64 contiguous bytes.
    const int f = (tid >> 5) * 8 + (tid & 7), k4 = (tid >> 3) & 3;
    const uint32_t sbase = smem_u32(smem);
    const uint32_t bar0 = sbase + B2_BAR, bar1 = sbase + B2_BAR + 8;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + B2_BAR + 16);
    const int64_t M = min((int64_t)A.counters[CNT_M_KEEP], A.cap_keep);
    const int64_t n_chunks = (M + KC - 1) / KC;
    // zero both stages once: padding rows (bias/pad columns of B, unused n) must stay zero
    for (int e = tid; e < 2 * STAGE / 16; e += B2_THREADS) reinterpret_cast<uint4*>(smem)[e] = make_uint4(0, 0, 0, 0);
    if (tid == 0) { mbar_init(bar0, 1); mbar_init(bar1, 1); }
    if (warp == 0) tmem_alloc(sbase + B2_BAR + 16, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    uint32_t par[2] = {0, 0};
    int issued[2] = {0, 0};
    float gsum = 0.f;
    int it = 0;
    struct Chunk { float4 a1, b1, a0, a2, xg; };
    auto load_chunk = [&](int64_t ch) {
        Chunk c;
        const int64_t s0 = ch * KC + k4 * 4;
        const bool live = ch < n_chunks;
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        c.a1 = live ? col4(A.k_dh1, s0, M, WD, f) : z;
        c.b1 = live ? col4(A.k_h0, s0, M, WD, f) : z;
        c.a0 = live ? col4(A.k_dh0, s0, M, WD, f) : z;
        c.a2 = live ? col4(A.k_h1, s0, M, WD, f) : z;
        c.xg = z;
        if (live && f < 39) c.xg = col4(A.k_x, s0, M, 40, f);
        else if (live && f >= 64 && f < 67) c.xg = col4_rows(A.k_glogit, s0, M, 3, f - 64);
        return c;
    };
    for (int64_t ch = blockIdx.x; ch < n_chunks; ch += gridDim.x, ++it) {
        const int st = it & 1;
        unsigned char* sb = smem + st * STAGE;
        const Chunk cur = load_chunk(ch);
        if (issued[st]) { mbar_wait(st ? bar1 : bar0, par[st]); par[st] ^= 1; }   // MMAs that read this stage are done
        // ---- stage the six operand tiles, transposed: thread = (feature, 4 samples) -> one 16-byte K-major row
        {
            const int64_t s0 = ch * KC + k4 * 4;
            put_k4(sb, S_A1, OPB(WD), f, k4, cur.a1);
            put_k4(sb, S_B1, OPB(N1), f, k4, cur.b1);
            put_k4(sb, S_A0, OPB(WD), f, k4, cur.a0);
            put_k4(sb, S_A2, OPB(WD), f, k4, cur.a2);
            const float4 ones = make_float4(s0 < M ? 1.f : 0.f, s0 + 1 < M ? 1.f : 0.f, s0 + 2 < M ? 1.f : 0.f, s0 + 3 < M ? 1.f : 0.f);
            if (f < 39) put_k4(sb, S_B0, OPB(N0), f, k4, cur.xg);
            else if (f == 39) put_k4(sb, S_B0, OPB(N0), 39, k4, ones);             // bias row of dW0
            else if (f == 40) put_k4(sb, S_B1, OPB(N1), 128, k4, ones);            // bias row of dW1
            else if (f >= 64 && f < 67) {                                           // G^T rows; db2 on the side
                gsum += (cur.xg.x + cur.xg.y) + (cur.xg.z + cur.xg.w);
                put_k4(sb, S_B2, OPB(N2), f - 64, k4, cur.xg);
            }
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            const uint32_t base = sbase + st * STAGE;
            const uint32_t acc0 = it > 0 ? 1u : 0u;
#pragma unroll
            for (int g = 0; g < 3; ++g) {
                const int offA = g == 0 ? S_A1 : g == 1 ? S_A0 : S_A2, offB = g == 0 ? S_B1 : g == 1 ? S_B0 : S_B2;
                const int nB = g == 0 ? N1 : g == 1 ? N0 : N2;
                const uint32_t dcol = tmem + (g == 0 ? ACC1 : g == 1 ? ACC0 : ACC2);
                const uint32_t idesc = make_idesc(nB);
                const uint64_t ahi = make_desc(base + offA, KC), alo = make_desc(base + offA + OPB(WD), KC);
                const uint64_t bhi = make_desc(base + offB, KC), blo = make_desc(base + offB + OPB(nB), KC);
#pragma unroll
                for (int ks = 0; ks < KC / 8; ++ks) {
                    const uint64_t adv = (uint64_t)(ks * 256) >> 4;
                    umma_tf32_ss(dcol, ahi + adv, bhi + adv, idesc, (ks == 0) ? acc0 : 1u);
                    umma_tf32_ss(dcol, alo + adv, bhi + adv, idesc, 1u);
                    umma_tf32_ss(dcol, ahi + adv, blo + adv, idesc, 1u);
                }
            }
            umma_commit(st ? bar1 : bar0);
        }
        issued[st] = 1;
    }
    // ---- drain and flush the accumulators
    for (int st = 0; st < 2; ++st)
        if (issued[st]) { mbar_wait(st ? bar1 : bar0, par[st]); par[st] ^= 1; }
    tc_fence_after();
    if (it > 0 && f >= 64 && f < 67) red_add(A.net_grad + PVDB_NET_OFF_B2 + (f - 64), gsum);
    if (it > 0 && tid < TM) {
        const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);
        float* G = A.net_grad;
        const int j = tid;   // accumulator row
        // dW1[j][0..128) and db1[j] (column 128)
#pragma unroll 1
        for (int c = 0; c < WD; c += 32) {
            uint32_t r[32];
            tmem_ld32(lane_addr + ACC1 + c, r);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; i += 4)
                red_add4(G + PVDB_NET_OFF_W1 + j * WD + c + i, __uint_as_float(r[i]), __uint_as_float(r[i + 1]), __uint_as_float(r[i + 2]),
                         __uint_as_float(r[i + 3]));
        }
        {
            uint32_t r[8];
            tmem_ld8(lane_addr + ACC1 + 128, r);
            tmem_ld_wait();
            red_add(G + PVDB_NET_OFF_B1 + j, __uint_as_float(r[0]));
        }
        // dW0[j][0..39) and db0[j] (column 39)
        {
            uint32_t r[32];
            tmem_ld32(lane_addr + ACC0, r);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) red_add(G + PVDB_NET_OFF_W0 + j * PVDB_NET_DIN + i, __uint_as_float(r[i]));
            uint32_t q[8];
            tmem_ld8(lane_addr + ACC0 + 32, q);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 7; ++i) red_add(G + PVDB_NET_OFF_W0 + j * PVDB_NET_DIN + 32 + i, __uint_as_float(q[i]));
            red_add(G + PVDB_NET_OFF_B0 + j, __uint_as_float(q[7]));
        }
        // dW2[c][i = j] (accumulator rows are i, columns c)
        {
            uint32_t r[8];
            tmem_ld8(lane_addr + ACC2, r);
            tmem_ld_wait();
#pragma unroll
            for (int c = 0; c < 3; ++c) red_add(G + PVDB_NET_OFF_W2 + c * WD + j, __uint_as_float(r[c]));
        }

    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 256);
}

}  // namespace

int pvdb_rgbnet_backward_tc(const pvdb_train_cfg* cfg, const pvdb_train_bufs* b, const float* viewdirs, cudaStream_t st) {
    PVDB_CHECK_ARG(b->k_h0 && b->k_h1 && b->k_dh0 && b->k_dh1 && b->k_x && b->k_mask, "the tcgen05 backward needs k_h0, k_h1, k_dh0, k_dh1, k_x, k_mask");
    static bool attr_set = false;
    if (!attr_set) {
        PVDB_CUDA(cudaFuncSetAttribute(k_rgbnet_bwd_act_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, B1_TOTAL));
        PVDB_CUDA(cudaFuncSetAttribute(k_rgbnet_bwd_wgrad_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, B2_TOTAL));
        attr_set = true;
    }
    BwdActArgs A;
    A.tree = *b->tree; A.net = b->net; A.k_glogit = b->k_rgb; A.k_mask = b->k_mask; A.k_xyz = b->k_xyz;
    A.k_dh1 = b->k_dh1; A.k_dh0 = b->k_dh0; A.k0_grad = b->k0_grad; A.k0_touched = b->k0_touched; A.k0_touched_list = b->k0_touched_list; A.counters_w = b->counters;
    A.counters = b->counters;
    A.cap_keep = b->cap_keep;
    k_rgbnet_bwd_act_tc<<<PVDB_SMS, TM, B1_TOTAL, st>>>(A);
    PVDB_LAUNCH_CHECK();
    pvdb_prof_mark("rgbnet_bwd_act", st);
    BwdWgradArgs W;
    W.k_dh1 = b->k_dh1; W.k_h0 = b->k_h0; W.k_dh0 = b->k_dh0; W.k_x = b->k_x; W.k_h1 = b->k_h1; W.k_glogit = b->k_rgb;
    W.net_grad = b->net_grad; W.counters = b->counters; W.cap_keep = b->cap_keep;
    k_rgbnet_bwd_wgrad_tc<<<PVDB_SMS, B2_THREADS, B2_TOTAL, st>>>(W);
    PVDB_LAUNCH_CHECK();
    return PVDB_OK;
}
