// dp_exchange.cu — data-parallel gradient exchange over NVLink peer memory (the reference is single-GPU,
// so this has no counterpart there).  Each rank owns one "symmetric block" (cudaMalloc + CUDA IPC, mapped by every peer):
//
//   [0,448)      signal words, u32, monotone epochs (step + 1) written by the peers with st.release.sys and polled by the
//                owner: A[src] flags published, B[src] gradients final, D[src] sums delivered, C[src] rgbnet gradients delivered,
//                CS[slice][src] (stand-alone rgbnet exchange)
//   [448,512)    err, last-CTA counters
//   [1024,..)    flags[2][n_leaf rounded up to 16] u8   touched-leaf flags published by the owner (double-buffered by step parity)
//   [net_off,.)  net[2][NET_PAD]  f32      stand-alone rgbnet exchange: the owner's gradients, read by the peers
//   [netx_off,.) netx[2][8][NET_PAD] f32   fused rgbnet exchange: slot [src] is WRITTEN BY rank src (peer stores)
//   [den_off,.)  den_grad[n_leaf][512]     THE gradient planes of this rank's grids (dist.py binds DensityVDB.grad /
//   [k0_off,.)   k0_grad[n_leaf][512][12]  ColorVDB.grad to them): the scatter kernels accumulate straight into peer-visible memory
//
// No host synchronisation, no NCCL call, no staging copy.  Per step and rank the NVLink traffic is O(1) in the number of ranks:
//   k_dp_union   1 CTA : cross-GPU barrier A -> OR of all peers' flag words -> ascending union list (identical everywhere)
//   k_dp_rs      grid  : signals B (this rank's planes are final), then reduce-scatter + all-gather in one kernel.  Rank r owns the union slots s with s % world == r: it
//                        reads that leaf's tile from every rank's planes (world - 1 peer loads of 1/world of the data), sums in
//                        rank order (identical bits everywhere) and stores the sum into EVERY rank's planes (world - 1 peer
//                        stores); the last CTA signals D
//   k_dp_wait    1 warp: all owners have delivered (wait for D) — stand-alone exchange; in the fused step the leaf Adam waits itself
// The rgbnet gradients (88 KB) ride on the kernels that produce and consume them: the weight-gradient reduction stores its
// sums straight into every peer's netx[parity][rank] (pvdb_dp_net_push_args) and the rgbnet Adam waits for C and adds the
// world slots of its own block in rank order (pvdb_dp_net_wait_args) — no exchange kernel at all on that path.
// In the fused step (pvdb_train_step_dp) the flag bits are set by the emit kernel (the sample lists determine the touched
// leaves before any gradient exists), so k_dp_union — and with it the barrier that absorbs the ranks' skew — runs on the side
// stream under the rgbnet forward; ready / rs / wait run under the weight-gradient kernel.
// A plane element of a union leaf has one remote reader and one remote writer, the leaf's owner, which reads all ranks' values
// before it stores the sum, element by element; a rank touches its planes again (Adam clears them) only after D from every
// owner.  The flag words and netx are double-buffered: a rank overwrites parity p two steps later, after it passed barrier A
// of the step in between, which every peer reaches only after all of its reads of the earlier step (stream order).
#include "common.cuh"
#include "rgbnet.cuh"
#include "peer_sync.cuh"
#include "dp_exchange.cuh"

namespace {

constexpr int NET_PAD = PVDB_DP_NET_PAD;
constexpr int NET_SLICES = 8;
enum { SIG_A = 0, SIG_B = 8, SIG_D = 16, SIG_C = 24, SIG_CS = 32 };   // word offsets inside the signal area (CS: [slice][8])

struct Blk {
    uint32_t* signal;
    int32_t* err;
    uint32_t* done;      // [4] last-CTA counters
    uint8_t* flags[2];   // byte per leaf, rows padded to 16 bytes
    float* net[2];
    float* netx[2];      // [8][NET_PAD] each
    float* den_grad;
    float* k0_grad;
};
__host__ __device__ inline int flag_bytes(int n_leaf) { return (n_leaf + 15) & ~15; }
__host__ __device__ inline size_t net_off(int n_leaf) { return (1024 + (size_t)2 * flag_bytes(n_leaf) + 255) & ~(size_t)255; }
__host__ __device__ inline size_t netx_off(int n_leaf) { return net_off(n_leaf) + 2 * (size_t)NET_PAD * sizeof(float); }
__host__ __device__ inline size_t den_off(int n_leaf) { return netx_off(n_leaf) + 2 * 8 * (size_t)NET_PAD * sizeof(float); }
__host__ __device__ inline size_t k0_off(int n_leaf) { return den_off(n_leaf) + (size_t)(n_leaf > 0 ? n_leaf : 1) * PVDB_LEAF_VOX * sizeof(float); }
__host__ __device__ inline Blk view(void* base, int n_leaf) {
    char* p = static_cast<char*>(base);
    Blk b;
    b.signal = reinterpret_cast<uint32_t*>(p);
    b.err = reinterpret_cast<int32_t*>(p + 448);
    b.done = reinterpret_cast<uint32_t*>(p + 452);
    b.flags[0] = reinterpret_cast<uint8_t*>(p + 1024);
    b.flags[1] = b.flags[0] + flag_bytes(n_leaf);
    b.net[0] = reinterpret_cast<float*>(p + net_off(n_leaf));
    b.net[1] = b.net[0] + NET_PAD;
    b.netx[0] = reinterpret_cast<float*>(p + netx_off(n_leaf));
    b.netx[1] = b.netx[0] + 8 * (size_t)NET_PAD;
    b.den_grad = reinterpret_cast<float*>(p + den_off(n_leaf));
    b.k0_grad = reinterpret_cast<float*>(p + k0_off(n_leaf));
    return b;
}

// thread `peer` of a CTA: tell rank `peer` that this rank reached `epoch`
__device__ __forceinline__ void signal_peer(const pvdb_dp_peers& P, int peer, int word, uint32_t epoch) {
    st_release_sys(view(P.base[peer], P.n_leaf).signal + word + P.rank, epoch);
}
// thread `peer` of a CTA: wait until rank `peer` reached `epoch`
__device__ __forceinline__ void wait_peer(const pvdb_dp_peers& P, int peer, int word, uint32_t epoch) {
    const Blk me = view(P.base[P.rank], P.n_leaf);
    wait_epoch(me.signal + word + peer, epoch, me.err, 1);   // a dead peer must not hang the GPU
}

// publish != 0 (stand-alone exchange): the flag bits are taken from den_touched | k0_touched here; publish == 0 (fused step):
// the emit kernel has already set them in flags[parity].
// 256 threads and 32 registers: the CTA must fit on an SM NEXT TO a persistent tcgen05 CTA (rgbnet forward: 416 threads x 128
// registers), or it would keep that SM's forward CTA from starting while it spins on the peers.  Every thread owns a contiguous
// run of flag words, so that all of its peer loads are in flight together (one NVLink latency) and the list comes out in
// ascending leaf order.
constexpr int UNION_T = 256;
constexpr int UNION_WORDS = 1024;                 // flag words per round: 32 768 leaves
__global__ void __launch_bounds__(UNION_T) k_dp_union(pvdb_dp_peers P, uint32_t epoch, int parity, int publish, int32_t* __restrict__ den_touched,
                                                      int32_t* __restrict__ k0_touched, int32_t* __restrict__ den_list,
                                                      int32_t* __restrict__ k0_list, int32_t* __restrict__ counters, int cnt_den, int cnt_k0,
                                                      unsigned long long* __restrict__ dbg) {
    __shared__ int warp_cnt[UNION_T / 32];
    __shared__ int running;
    __shared__ uint32_t s_word[UNION_WORDS];
    __shared__ int s_pre[UNION_WORDS];
    pvdb_pdl_wait();
    if (dbg && threadIdx.x == 0) dbg[20] = globaltimer();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const Blk me = view(P.base[P.rank], P.n_leaf);
    const int NB = flag_bytes(P.n_leaf) / 16;          // 16-byte groups of flags
    const int W = (P.n_leaf + 31) >> 5;                // union bit words
    if (publish)
        for (int i = threadIdx.x; i < P.n_leaf; i += UNION_T) me.flags[parity][i] = (den_touched[i] | k0_touched[i]) != 0;
    if (threadIdx.x == 0) running = 0;
    __syncthreads();   // the st.release.sys below is cumulative over the CTA's flag writes ordered by this barrier
    if (dbg && threadIdx.x == 0) dbg[21] = globaltimer();
    if (threadIdx.x < P.world) {
        signal_peer(P, threadIdx.x, SIG_A, epoch);
        if (dbg && threadIdx.x == 0) dbg[22] = globaltimer();
        wait_peer(P, threadIdx.x, SIG_A, epoch);
    }
    __syncthreads();
    if (dbg && threadIdx.x == 0) dbg[23] = globaltimer();
    const uint4* pf[8];
    for (int r = 0; r < 8; ++r) pf[r] = reinterpret_cast<const uint4*>(view(P.base[r < P.world ? r : 0], P.n_leaf).flags[parity]);
    uint4* other = reinterpret_cast<uint4*>(me.flags[parity ^ 1]);
    // rounds of UNION_WORDS bit words (32 768 leaves): the ranks' byte flags are OR-ed 16 at a time into bit words in shared
    // memory, then the exclusive popcount prefix of the words, then one thread per LEAF writes the int flags and the list entry
    // (ascending leaf order)
    for (int base = 0; base < W; base += UNION_WORDS) {
        const int nw = min(UNION_WORDS, W - base);
        for (int w = threadIdx.x; w < nw; w += UNION_T) s_word[w] = 0;
        __syncthreads();
        const int g0 = base * 2, ng = min(nw * 2, NB - g0);            // 16-byte groups of this round (two per word)
        for (int g = threadIdx.x; g < ng; g += UNION_T) {
            uint4 u = make_uint4(0, 0, 0, 0);
            for (int r = 0; r < P.world; ++r) {                        // independent loads: one NVLink latency
                const uint4 v = __ldcv(pf[r] + g0 + g);
                u.x |= v.x; u.y |= v.y; u.z |= v.z; u.w |= v.w;
            }
            // every peer has passed barrier A of this step, i.e. finished the union of the previous one: the other parity's
            // flags have no reader left and are cleared for the emit kernel of the next step
            other[g0 + g] = make_uint4(0, 0, 0, 0);
            uint32_t bits = 0;
            const uint32_t q[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
            for (int k = 0; k < 16; ++k) bits |= ((q[k >> 2] >> ((k & 3) * 8)) & 0xffu ? 1u : 0u) << k;
            if (bits) atomicOr(&s_word[g >> 1], bits << ((g & 1) * 16));
        }
        __syncthreads();
        // exclusive prefix of the popcounts: UNION_WORDS / UNION_T consecutive words per thread, then a block scan
        constexpr int WPT = UNION_WORDS / UNION_T;
        int cnt = 0;
#pragma unroll
        for (int k = 0; k < WPT; ++k) {
            const int w = threadIdx.x * WPT + k;
            if (w < nw) cnt += __popc(s_word[w]);
        }
        int incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const int o = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += o; }
        if (lane == 31) warp_cnt[warp] = incl;
        __syncthreads();
        int before = running;
        for (int q = 0; q < warp; ++q) before += warp_cnt[q];
        int total = 0;
        for (int q = 0; q < UNION_T / 32; ++q) total += warp_cnt[q];
        int run = before + incl - cnt;
#pragma unroll
        for (int k = 0; k < WPT; ++k) {
            const int w = threadIdx.x * WPT + k;
            if (w < nw) { s_pre[w] = run; run += __popc(s_word[w]); }
        }
        __syncthreads();
        const int leaf0 = base * 32, nleaf = min(nw * 32, P.n_leaf - leaf0);
        for (int j = threadIdx.x; j < nleaf; j += UNION_T) {
            const uint32_t u = s_word[j >> 5];
            const int t = (u >> (j & 31)) & 1;      // bits of padding leaves (>= n_leaf) are never set: their flags are never written
            den_touched[leaf0 + j] = t; k0_touched[leaf0 + j] = t;
            if (t) {
                const int slot = s_pre[j >> 5] + __popc(u & ((1u << (j & 31)) - 1u));
                den_list[slot] = leaf0 + j; k0_list[slot] = leaf0 + j;
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) running += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        counters[cnt_den] = running;
        counters[cnt_k0] = running;
        if (dbg) dbg[24] = globaltimer();
    }
}

// Every owner has stored its sums into this rank's planes.
__global__ void __launch_bounds__(32) k_dp_wait(pvdb_dp_peers P, uint32_t epoch, int word) {
    pvdb_pdl_wait();
    if (threadIdx.x < P.world) wait_peer(P, threadIdx.x, word, epoch);
}

// Reduce-scatter + all-gather: this rank sums the union leaves it owns (slot % world == rank) over all ranks' planes, in rank
// order, and stores the sums into every rank's planes.  128 threads: with eight float4 in flight per thread (~100 registers) a
// 256-thread CTA would not fit next to the persistent weight-gradient CTA it is meant to run under.
// ITEMS tile elements per thread x at most 8 / ITEMS ranks = eight float4 in flight whatever the world size: one CTA fits on
// an SM next to the weight-gradient CTA, so the working CTAs should be ONE wave — at N = 2 with one element per thread they
// were 481 CTAs = 3.3 waves of ~9 us each (peer loads, peer stores, system fence: pure latency), the whole side chain 25 us late.
template <int ITEMS>
__global__ void __launch_bounds__(128) k_dp_rs(pvdb_dp_peers P, uint32_t epoch, const int32_t* __restrict__ list, const int32_t* __restrict__ counters,
                                               int cnt_den, unsigned long long* __restrict__ dbg) {
    constexpr int WMAX = 8 / ITEMS;
    pvdb_pdl_wait();
    if (dbg && blockIdx.x == 0 && threadIdx.x == 0) dbg[25] = globaltimer();
    // this kernel is enqueued behind the kernels that scatter into this rank's planes: they are final (signal B, first CTA)
    if (blockIdx.x == 0 && threadIdx.x < P.world) signal_peer(P, threadIdx.x, SIG_B, epoch);
    const int n = counters[cnt_den];
    constexpr int T4 = PVDB_LEAF_VOX * 13 / 4;       // 128 float4 of density + 1536 of k0 per leaf
    const int mine = n > P.rank ? (n - P.rank + P.world - 1) / P.world : 0;
    const int64_t total = (int64_t)mine * T4;
    const int64_t per = (int64_t)blockDim.x * ITEMS;      // tile elements of one CTA per grid pass
    // The grid is sized for the large case (S512: ~0.5 GB of tiles, five CTAs per SM once the weight-gradient kernel has left);
    // at F160 (74 union leaves, 2 MB) only the first ~120 CTAs have a tile element: the rest leave at once, without waiting for
    // the peers, and only take part in the last-CTA count.
    if ((int64_t)blockIdx.x * per >= total && blockIdx.x != 0) {
        if (threadIdx.x == 0) {
            uint32_t* done0 = view(P.base[P.rank], P.n_leaf).done + 1;
            if (atomicAdd(done0, 1u) == gridDim.x - 1) {      // cannot be the last one unless every working CTA is done: they count after their fences
                *done0 = 0;
                __threadfence_system();
                for (int r = 0; r < P.world; ++r) signal_peer(P, r, SIG_D, epoch);
            }
        }
        return;
    }
    if (threadIdx.x < P.world) wait_peer(P, threadIdx.x, SIG_B, epoch);
    __syncthreads();
    if (dbg && blockIdx.x == 0 && threadIdx.x == 0) dbg[26] = globaltimer();
    float4* pd[WMAX];
    float4* pk[WMAX];
#pragma unroll
    for (int r = 0; r < WMAX; ++r) {
        const Blk b = view(P.base[r < P.world ? r : 0], P.n_leaf);
        pd[r] = reinterpret_cast<float4*>(b.den_grad);
        pk[r] = reinterpret_cast<float4*>(b.k0_grad);
    }
    // plain loads: the acquire + barrier above orders them after the peers' scatter kernels, and nothing on this GPU has read
    // these peer addresses since the previous step's exchange (another launch)
    for (int64_t base = (int64_t)blockIdx.x * per; base < total; base += (int64_t)gridDim.x * per) {
        float4 v[ITEMS][WMAX];
        int64_t off[ITEMS];
        bool den[ITEMS], live[ITEMS];
#pragma unroll
        for (int k = 0; k < ITEMS; ++k) {
            const int64_t idx = base + (int64_t)k * blockDim.x + threadIdx.x;
            live[k] = idx < total;
            const int j = live[k] ? (int)(idx / T4) : 0, i = live[k] ? (int)(idx - (int64_t)j * T4) : 0;
            const int leaf = list[P.rank + j * P.world];
            den[k] = i < 128;
            off[k] = den[k] ? (int64_t)leaf * 128 + i : (int64_t)leaf * 1536 + (i - 128);
#pragma unroll
            for (int r = 0; r < WMAX; ++r)
                if (r < P.world && live[k]) v[k][r] = (den[k] ? pd[r] : pk[r])[off[k]];
        }
#pragma unroll
        for (int k = 0; k < ITEMS; ++k) {
            if (!live[k]) continue;
            float4 s = v[k][0];
#pragma unroll
            for (int r = 1; r < WMAX; ++r)
                if (r < P.world) { s.x += v[k][r].x; s.y += v[k][r].y; s.z += v[k][r].z; s.w += v[k][r].w; }
#pragma unroll
            for (int r = 0; r < WMAX; ++r)
                if (r < P.world) (den[k] ? pd[r] : pk[r])[off[k]] = s;
        }
    }
    // last CTA out tells every peer that this rank's sums are in place (threadFenceReduction pattern with system-wide fences:
    // the stores it publishes are peer stores; the final st.release.sys is cumulative over what the counter made visible)
    __shared__ bool last;
    uint32_t* done = view(P.base[P.rank], P.n_leaf).done + 1;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        last = atomicAdd(done, 1u) == gridDim.x - 1;
        if (last) { *done = 0; __threadfence_system(); }
    }
    __syncthreads();
    if (dbg && blockIdx.x == 0 && threadIdx.x == 0) dbg[27] = globaltimer();
    if (dbg && last && threadIdx.x == 0) dbg[28] = globaltimer();
    if (last && threadIdx.x < P.world) signal_peer(P, threadIdx.x, SIG_D, epoch);
}

// Stand-alone rgbnet exchange (pvdb_dp_exchange_net): CTA `g` owns slice g of the 22 019 values end to end (publish, signal,
// wait, sum), so there is no grid-wide dependency and a single latency of peer loads.
__global__ void __launch_bounds__(1024) k_dp_net(pvdb_dp_peers P, uint32_t epoch, int parity, float* __restrict__ net_grad) {
    constexpr int SL = NET_PAD / NET_SLICES;       // floats per slice (multiple of 4)
    pvdb_pdl_wait();
    const int g = blockIdx.x;
    const Blk me = view(P.base[P.rank], P.n_leaf);
    const int i = g * SL + threadIdx.x * 4;
    const bool live = threadIdx.x * 4 < SL;
    float4 own = make_float4(0.f, 0.f, 0.f, 0.f);
    if (live) {
        own.x = i < PVDB_NET_N ? net_grad[i] : 0.f; own.y = i + 1 < PVDB_NET_N ? net_grad[i + 1] : 0.f;
        own.z = i + 2 < PVDB_NET_N ? net_grad[i + 2] : 0.f; own.w = i + 3 < PVDB_NET_N ? net_grad[i + 3] : 0.f;
        *reinterpret_cast<float4*>(me.net[parity] + i) = own;
    }
    __syncthreads();   // the st.release.sys below is cumulative over the CTA's stores ordered by this barrier
    if (threadIdx.x < P.world) {
        signal_peer(P, threadIdx.x, SIG_CS + g * 8, epoch);
        wait_peer(P, threadIdx.x, SIG_CS + g * 8, epoch);
    }
    __syncthreads();
    if (!live) return;
    float4 v[8];
#pragma unroll
    for (int r = 0; r < 8; ++r)
        if (r < P.world) v[r] = r == P.rank ? own : *reinterpret_cast<const float4*>(view(P.base[r], P.n_leaf).net[parity] + i);
    float4 s = v[0];
#pragma unroll
    for (int r = 1; r < 8; ++r)
        if (r < P.world) { s.x += v[r].x; s.y += v[r].y; s.z += v[r].z; s.w += v[r].w; }
    if (i < PVDB_NET_N) net_grad[i] = s.x;
    if (i + 1 < PVDB_NET_N) net_grad[i + 1] = s.y;
    if (i + 2 < PVDB_NET_N) net_grad[i + 2] = s.z;
    if (i + 3 < PVDB_NET_N) net_grad[i + 3] = s.w;
}

int check_peers(const pvdb_dp_peers* P, const pvdb_train_bufs* b) {
    PVDB_CHECK_ARG(P && b && b->tree, "null pointer");
    PVDB_CHECK_ARG(P->world >= 1 && P->world <= 8 && P->rank >= 0 && P->rank < P->world, "world must be 1..8");
    PVDB_CHECK_ARG(P->n_leaf == b->tree->n_leaf, "peers block was sized for another tree");
    for (int r = 0; r < P->world; ++r) PVDB_CHECK_ARG(P->base[r], "peer block not mapped");
    const Blk me = view(P->base[P->rank], P->n_leaf);
    PVDB_CHECK_ARG(b->den_grad == me.den_grad && b->k0_grad == me.k0_grad,
                   "the gradient planes must be the ones inside this rank's symmetric block (pvdb_dp_grad_planes; dist.PeerExchange binds them)");
    return PVDB_OK;
}

}  // namespace

extern "C" size_t pvdb_dp_symm_bytes(int n_leaf, int cap_leaves) {
    (void)cap_leaves;
    if (n_leaf < 0) return 0;
    return k0_off(n_leaf) + (size_t)(n_leaf > 0 ? n_leaf : 1) * PVDB_LEAF_VOX * 12 * sizeof(float);
}
extern "C" int pvdb_dp_grad_planes(const pvdb_dp_peers* P, float** den_grad, float** k0_grad) {
    PVDB_CHECK_ARG(P && den_grad && k0_grad && P->rank >= 0 && P->rank < 8 && P->base[P->rank], "bad peers");
    const Blk me = view(P->base[P->rank], P->n_leaf);
    *den_grad = me.den_grad;
    *k0_grad = me.k0_grad;
    return PVDB_OK;
}
extern "C" int pvdb_dp_symm_alloc(size_t bytes, void** ptr, void* handle64) {
    PVDB_CHECK_ARG(ptr && handle64 && bytes > 0, "null pointer / zero size");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    PVDB_CUDA(cudaMalloc(ptr, bytes));
    PVDB_CUDA(cudaMemset(*ptr, 0, bytes));
    PVDB_CUDA(cudaIpcGetMemHandle(static_cast<cudaIpcMemHandle_t*>(handle64), *ptr));
    PVDB_CUDA(cudaDeviceSynchronize());
    return PVDB_OK;
}
extern "C" int pvdb_dp_symm_open(const void* handle64, void** ptr) {
    PVDB_CHECK_ARG(ptr && handle64, "null pointer");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof(h));
    PVDB_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return PVDB_OK;
}
extern "C" int pvdb_dp_symm_close(void* ptr) {
    if (ptr) PVDB_CUDA(cudaIpcCloseMemHandle(ptr));
    return PVDB_OK;
}
extern "C" int pvdb_dp_symm_free(void* ptr) {
    if (ptr) PVDB_CUDA(cudaFree(ptr));
    return PVDB_OK;
}
extern "C" int pvdb_dp_symm_error(const pvdb_dp_peers* P, int32_t* err_out) {
    PVDB_CHECK_ARG(P && err_out && P->rank >= 0 && P->rank < 8 && P->base[P->rank], "bad peers");
    PVDB_CUDA(cudaMemcpy(err_out, view(P->base[P->rank], P->n_leaf).err, sizeof(int32_t), cudaMemcpyDeviceToHost));
    return PVDB_OK;
}

// Kernels meant to run on an SM NEXT TO a persistent tcgen05 CTA ask for the same shared-memory carve-out as that CTA (maximum
// shared memory): an SM is configured for one carve-out at a time, and a CTA that prefers another one waits until the SM drains.
template <typename K>
static void prefer_max_smem(K kernel) { cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared); }
static void dp_kernel_attrs() {
    static bool done = false;
    if (done) return;
    prefer_max_smem(k_dp_union); prefer_max_smem(k_dp_rs<1>); prefer_max_smem(k_dp_rs<2>); prefer_max_smem(k_dp_rs<4>); prefer_max_smem(k_dp_wait);
    done = true;
}

static int exchange_tiles(const pvdb_dp_peers* P, const pvdb_train_bufs* b, uint32_t step, cudaStream_t st, bool do_union, int publish,
                          bool do_move, bool wait_sums) {
    if (int rc = check_peers(P, b)) return rc;
    dp_kernel_attrs();
    const int parity = step & 1;
    const uint32_t epoch = step + 1;                        // monotone; the signal words start at 0
    const int CNT_DEN = 2, CNT_K0 = 4;                      // counters[] slots of pvdb_train_bufs (include/plenvdb_b200.h)
    if (do_union) {
        PVDB_CUDA(pvdb_launch_pdl(k_dp_union, dim3(1), dim3(UNION_T), 0, st, *P, epoch, parity, publish, b->den_touched, b->k0_touched,
                                  b->den_touched_list, b->k0_touched_list, b->counters, CNT_DEN, CNT_K0, pvdb_debug_stamps_ptr()));
        PVDB_LAUNCH_CHECK();
        pvdb_prof_mark("dp_union", st);
    }
    if (!do_move) return PVDB_OK;
    // One CTA fits on each SM next to the persistent weight-gradient CTA, five once it has left; CTAs without a tile element
    // leave at once: 74 union leaves = 2 MB at F160 (latency bound, one wave of working CTAs), ~0.5 GB at S512
    auto rs = P->world <= 2 ? k_dp_rs<4> : P->world <= 4 ? k_dp_rs<2> : k_dp_rs<1>;
    PVDB_CUDA(pvdb_launch_pdl(rs, dim3(PVDB_SMS * 5), dim3(128), 0, st, *P, epoch, (const int32_t*)b->den_touched_list,
                              (const int32_t*)b->counters, CNT_DEN, pvdb_debug_stamps_ptr()));
    PVDB_LAUNCH_CHECK();
    pvdb_prof_mark("dp_rs", st);
    if (wait_sums) {
        PVDB_CUDA(pvdb_launch_pdl(k_dp_wait, dim3(1), dim3(32), 0, st, *P, epoch, (int)SIG_D));
        PVDB_LAUNCH_CHECK();
        pvdb_prof_mark("dp_wait", st);
    }
    return PVDB_OK;
}

extern "C" int pvdb_dp_exchange_tiles(const pvdb_dp_peers* P, const pvdb_train_bufs* b, uint32_t step, void* stream) {
    return exchange_tiles(P, b, step, (cudaStream_t)stream, true, 1, true, true);
}
// The fused step's two halves: the union alone (flags already written by the emit kernel), and pack / reduce / unpack alone.
int pvdb_dp_union_early(const pvdb_dp_peers* P, const pvdb_train_bufs* b, uint32_t step, cudaStream_t st) {
    return exchange_tiles(P, b, step, st, true, 0, false, false);
}
int pvdb_dp_move_tiles(const pvdb_dp_peers* P, const pvdb_train_bufs* b, uint32_t step, cudaStream_t st) {
    return exchange_tiles(P, b, step, st, false, 0, true, false);   // the leaf Adam waits for the sums itself
}
uint8_t* pvdb_dp_flags_ptr(const pvdb_dp_peers* P, uint32_t step) {
    return view(P->base[P->rank], P->n_leaf).flags[step & 1];
}
// Arguments for the kernels that carry the rgbnet exchange of the fused step (dp_exchange.cuh).
PvdbDpNetPush pvdb_dp_net_push_args(const pvdb_dp_peers* P, uint32_t step) {
    PvdbDpNetPush a;
    a.world = P->world; a.rank = P->rank; a.epoch = step + 1;
    for (int r = 0; r < 8; ++r) {
        const Blk b = view(P->base[r < P->world ? r : 0], P->n_leaf);
        a.dst[r] = b.netx[step & 1] + (size_t)P->rank * NET_PAD;
        a.signal[r] = b.signal + SIG_C + P->rank;
    }
    a.done = view(P->base[P->rank], P->n_leaf).done + 2;
    return a;
}
PvdbDpTilesWait pvdb_dp_tiles_wait_args(const pvdb_dp_peers* P, uint32_t step) {
    PvdbDpTilesWait a;
    const Blk me = view(P->base[P->rank], P->n_leaf);
    a.world = P->world; a.epoch = step + 1; a.signal = me.signal + SIG_D; a.err = me.err;
    return a;
}
PvdbDpNetWait pvdb_dp_net_wait_args(const pvdb_dp_peers* P, uint32_t step) {
    PvdbDpNetWait a;
    const Blk me = view(P->base[P->rank], P->n_leaf);
    a.world = P->world; a.epoch = step + 1;
    a.src = me.netx[step & 1];
    a.signal = me.signal + SIG_C;
    a.err = me.err;
    return a;
}

extern "C" int pvdb_dp_exchange_net(const pvdb_dp_peers* P, const pvdb_train_bufs* b, uint32_t step, void* stream) {
    if (int rc = check_peers(P, b)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    PVDB_CUDA(pvdb_launch_pdl(k_dp_net, dim3(NET_SLICES), dim3(1024), 0, st, *P, (uint32_t)(step + 1), (int)(step & 1), b->net_grad));
    PVDB_LAUNCH_CHECK();
    pvdb_prof_mark("dp_net", st);
    return PVDB_OK;
}

extern "C" int pvdb_dp_exchange(const pvdb_dp_peers* P, const pvdb_train_bufs* b, uint32_t step, void* stream) {
    pvdb_reset_launch_count();
    pvdb_prof_begin((cudaStream_t)stream);
    if (int rc = pvdb_dp_exchange_tiles(P, b, step, stream)) return rc;
    return pvdb_dp_exchange_net(P, b, step, stream);
}
