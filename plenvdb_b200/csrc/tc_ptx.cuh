// tc_ptx.cuh — inline-PTX wrappers for tcgen05 / TMEM / mbarrier and the UMMA descriptor helpers shared by the
// tensor-core rgbnet kernels (rgbnet_tc.cu forward, rgbnet_tc_bwd.cu backward).  sm_100a only.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace {

constexpr int TM = 128;            // samples per tile = UMMA M

// ---- PTX wrappers -------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    }
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// 1-D bulk async copy global -> shared (TMA engine), completion counted in bytes on `bar`.  16-byte aligned, size % 16 == 0.
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem), "l"(src),
                 "r"(bytes), "r"(bar)
                 : "memory");
}
// 1-D bulk async copy shared -> global (TMA engine), tracked by the issuing thread's bulk groups.  16-byte aligned, size % 16 == 0.
__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src_smem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }   // sources reusable
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }             // writes done

// D[tmem] (+)= A[tmem] * B[smem desc], kind::tf32, issued by one thread
__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
                 "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}

// Activation tensors exchanged between the rgbnet kernels through HBM are "chunk-major": [tile of 128 samples][chunk of 16
// samples][feature][16 samples].  A lane (= sample) writing feature f of its row lands in one of two 64-byte runs per warp
// (full 32-byte sectors), consecutive features are adjacent, and the weight-gradient pass streams a whole 16-sample chunk of
// one tensor as ONE contiguous nf*64-byte block.  act_off: float offset of (sample s, feature f) in a tensor with nf features.
__device__ __forceinline__ size_t act_off(int64_t s, int f, int nf) {
    return (size_t)(s >> 7) * (size_t)(nf * 128) + (size_t)((s >> 4) & 7) * (size_t)(nf * 16) + (size_t)f * 16 + (size_t)(s & 15);
}

// Activation stores of the tcgen05 kernels.  One lane (= sample) writing its 32 values of a column group with 32 scalar
// st.global makes the warp push 128 KB per tile through the SM's store path while it is blocked on it: the stores of the
// forward alone are 104 MB per step = the L2 slices' whole-chip ingest rate (~6300 B/clk, /opt/skills/guides/B300_MICROARCH.md)
// for 16 k cycles, and the lane warps sat through all of it (3.3 k of a 9.4 k-cycle tile, measured by leaving the stores out).
// Staged instead: each lane warp owns a [2 chunks][32 features][16 samples] region of shared memory (chunk stride padded by
// 16 floats: the two half-warps hit disjoint banks), writes its values there and lane 0 hands the two 2 KB blocks — already in
// the chunk-major global layout (act_off) — to the TMA engine, which drains them while the warp goes on.
constexpr int STAGE_CHUNK = 32 * 16 + 16;                 // floats between the two 16-sample chunks of a warp's region
constexpr int STAGE_WARP_BYTES = 2 * STAGE_CHUNK * 4;     // 4224
// h[32]: this lane's values for features c .. c+31 of sample s (lane = s & 31); g: tensor base (nf features per sample)
__device__ __forceinline__ void stage_store32(unsigned char* stage_warp, float* g, int64_t s, int c, int nf, const float* h, bool valid) {
    const int lane = threadIdx.x & 31;
    if (lane == 0) bulk_wait_read0();      // the previous blocks of this warp have left shared memory
    __syncwarp();
    float* p = reinterpret_cast<float*>(stage_warp) + (lane >> 4) * STAGE_CHUNK + (lane & 15);
    if (valid) {
#pragma unroll
        for (int i = 0; i < 32; ++i) p[i * 16] = h[i];
    } else {      // rows past M (one partial tile per launch): zeros, the weight-gradient pass reads whole tiles
#pragma unroll
        for (int i = 0; i < 32; ++i) p[i * 16] = 0.f;
    }
    fence_async_smem();
    __syncwarp();
    if (lane == 0) {
        const uint32_t src = smem_u32(stage_warp);
        bulk_s2g(g + act_off(s, c, nf), src, 32 * 16 * 4);
        bulk_s2g(g + act_off(s + 16, c, nf), src + STAGE_CHUNK * 4, 32 * 16 * 4);
        bulk_commit();
    }
}


// ---- operand helpers ----------------------------------------------------------------------------------------
// 3xTF32 split: hi keeps the 10 explicit tf32 mantissa bits, lo = x - hi is exact in fp32.
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
    hi = __float_as_uint(x) & 0xffffe000u;
    lo = __float_as_uint(x - __uint_as_float(hi));
}
// byte offset of element (n, k) of a K-major operand with K columns in the canonical no-swizzle UMMA layout:
// 8-row x 16-byte core matrices, K-chunks of one 8-row group contiguous (LBO = 128 B), groups SBO = K/4*128 B apart.
__device__ __forceinline__ int canon_off(int n, int k, int K) { return (n >> 3) * (K / 4) * 128 + (k >> 2) * 128 + (n & 7) * 16 + (k & 3) * 4; }
// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start>>4 | LBO>>4 <<16 | SBO>>4 <<32 | version 1 <<46, no swizzle
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, int K) {
    const uint64_t lbo = 128 >> 4, sbo = (uint64_t)((K / 4) * 128) >> 4;
    return (uint64_t)((saddr >> 4) & 0x3fff) | (lbo << 16) | (sbo << 32) | (1ull << 46);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D fp32, A/B tf32, both K-major, M = 128
__device__ __forceinline__ uint32_t make_idesc(int N) { return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(TM >> 4) << 24); }


// ---- weight image of the activation-gradient kernel (rgbnet_tc_bwd.cu B1), built by the forward's prep kernel ----------
constexpr int PVDB_BWD_IMG_OFFSET = 256 * 1024;         // inside pvdb_train_bufs.net_img
constexpr int B1_W1HI = 0;                              // B[N=i][K=j] = w1[j][i], canonical K-major, [128][128]
constexpr int B1_W1LO = B1_W1HI + 128 * 128 * 4;
constexpr int B1_W0HI = B1_W1LO + 128 * 128 * 4;          // B[N=i<16][K=j] = w0[j][i], [16][128]
constexpr int B1_W0LO = B1_W0HI + 16 * 128 * 4;
constexpr int B1_W2 = B1_W0LO + 16 * 128 * 4;            // plain floats [3][128]
constexpr int B1_IMG = B1_W2 + 3 * 128 * 4;
__device__ __forceinline__ void prep_bwd_image(const float* __restrict__ net, unsigned char* __restrict__ img, int gtid, int gsz) {
    constexpr int OFF_W0 = 0, OFF_W1 = 128 * 39 + 128, OFF_W2 = OFF_W1 + 128 * 128 + 128;   // rgbnet.cuh packed layout
    for (int e = gtid; e < 128 * 128; e += gsz) {         // B[n = i][k = j] = w1[j][i]
        const int n = e % 128, k = e / 128;               // coalesced read of w1[k][n]
        uint32_t hi, lo;
        split_tf32(__ldg(net + OFF_W1 + k * 128 + n), hi, lo);
        const int o = canon_off(n, k, 128);
        *reinterpret_cast<uint32_t*>(img + B1_W1HI + o) = hi;
        *reinterpret_cast<uint32_t*>(img + B1_W1LO + o) = lo;
    }
    for (int e = gtid; e < 16 * 128; e += gsz) {          // B[n = c < 16][k = j] = w0[j][c] (c < 12), zero padding rows
        const int n = e / 128, k = e % 128;
        uint32_t hi, lo;
        split_tf32(n < 12 ? __ldg(net + OFF_W0 + k * 39 + n) : 0.f, hi, lo);
        const int o = canon_off(n, k, 128);
        *reinterpret_cast<uint32_t*>(img + B1_W0HI + o) = hi;
        *reinterpret_cast<uint32_t*>(img + B1_W0LO + o) = lo;
    }
    for (int e = gtid; e < 3 * 128; e += gsz) reinterpret_cast<float*>(img + B1_W2)[e] = __ldg(net + OFF_W2 + e);
}

// tcgen05.ld of 8 columns
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], kind::tf32
__device__ __forceinline__ void umma_tf32_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
}  // namespace
