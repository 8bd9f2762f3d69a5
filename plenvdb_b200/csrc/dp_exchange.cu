// dp_exchange.cu — data-parallel gradient exchange over NVLink peer memory (the reference is single-GPU,
// so this has no counterpart there).  Each rank owns one "symmetric block" (cudaMalloc + CUDA IPC, mapped by every peer):
//
//   [0,448)      signal words, u32, monotone epochs (step + 1) written by the peers with st.release.sys and polled by the
//                owner: A[src] flags published, B[src] tiles packed, C[slice][src] rgbnet-gradient slice published
//   [448,512)    err, done_ctas
//   [1024,..)    flags[2][n_leaf] i32   touched flags published by the owner (double-buffered by step parity)
//   [net_off,.)  net[2][NET_PAD]  f32   rgbnet gradients
//   [grad_off,.) grad[2][cap]     f32   packed gradient tiles of the union leaves
//
// Two phases, no host synchronisation and no NCCL call.  They are independent so that the fused step can run the tile
// phase on a side stream UNDER the weight-gradient kernel (the grid gradients are final once the activation-gradient
// kernel and the density scatter are done) and only the 88 KB rgbnet phase after it:
//   tiles:  k_dp_union   1 CTA : publish own flags -> cross-GPU barrier A -> OR of all peers' flags -> ascending union list
//           k_dp_pack    grid  : own gradient tiles of the union leaves -> own grad[parity]; the last CTA signals B
//           k_dp_reduce  grid  : wait for B, then every rank sums the peers' tiles in rank order (identical bits
//                                everywhere) straight into its gradient planes, ready for the fused sparse Adam
//   net:    k_dp_net     8 CTAs: each publishes one slice of net_grad, signals C[slice], waits for the peers' C[slice] and
//                                sums that slice in rank order — no grid-wide dependency
// Double buffering makes further barriers unnecessary: a rank overwrites parity p two steps later, after it passed barrier
// A of the step in between, which every peer reaches only after all of its reads of the earlier step (stream order).
#include "common.cuh"
#include "rgbnet.cuh"
#include "peer_sync.cuh"

namespace {

constexpr int TILE_F = PVDB_LEAF_VOX * 13;                 // density [512] + k0 [512][12]
constexpr int NET_PAD = (PVDB_NET_N + 255) & ~255;
constexpr int NET_SLICES = 8;
enum { SIG_A = 0, SIG_B = 16, SIG_C = 32 };                // word offsets inside the signal area (C: [slice][8])

struct Blk {
    uint32_t* signal;
    int32_t* err;
    uint32_t* done;
    int32_t* flags[2];
    float* net[2];
    float* grad[2];
};
__host__ __device__ inline size_t net_off(int n_leaf) { return (1024 + (size_t)8 * n_leaf + 255) & ~(size_t)255; }
__host__ __device__ inline size_t grad_off(int n_leaf) { return net_off(n_leaf) + 2 * (size_t)NET_PAD * sizeof(float); }
__host__ __device__ inline size_t cap_floats(int cap_leaves) { return (size_t)cap_leaves * TILE_F; }
__host__ __device__ inline Blk view(void* base, int n_leaf, int cap_leaves) {
    char* p = static_cast<char*>(base);
    Blk b;
    b.signal = reinterpret_cast<uint32_t*>(p);
    b.err = reinterpret_cast<int32_t*>(p + 448);
    b.done = reinterpret_cast<uint32_t*>(p + 452);
    b.flags[0] = reinterpret_cast<int32_t*>(p + 1024);
    b.flags[1] = b.flags[0] + n_leaf;
    b.net[0] = reinterpret_cast<float*>(p + net_off(n_leaf));
    b.net[1] = b.net[0] + NET_PAD;
    b.grad[0] = reinterpret_cast<float*>(p + grad_off(n_leaf));
    b.grad[1] = b.grad[0] + cap_floats(cap_leaves);
    return b;
}

// thread `peer` of a CTA: tell rank `peer` that this rank reached `epoch`
__device__ __forceinline__ void signal_peer(const pvdb_dp_peers& P, int peer, int word, uint32_t epoch) {
    st_release_sys(view(P.base[peer], P.n_leaf, P.cap_leaves).signal + word + P.rank, epoch);
}
// thread `peer` of a CTA: wait until rank `peer` reached `epoch`
__device__ __forceinline__ void wait_peer(const pvdb_dp_peers& P, int peer, int word, uint32_t epoch) {
    const Blk me = view(P.base[P.rank], P.n_leaf, P.cap_leaves);
    wait_epoch(me.signal + word + peer, epoch, me.err, 1);   // a dead peer must not hang the GPU
}

__global__ void __launch_bounds__(1024) k_dp_union(pvdb_dp_peers P, uint32_t epoch, int parity, int32_t* __restrict__ den_touched,
                                                   int32_t* __restrict__ k0_touched, int32_t* __restrict__ den_list,
                                                   int32_t* __restrict__ k0_list, int32_t* __restrict__ counters, int cnt_den, int cnt_k0) {
    __shared__ int warp_cnt[32];
    __shared__ int running;
    pvdb_pdl_wait();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const Blk me = view(P.base[P.rank], P.n_leaf, P.cap_leaves);
    for (int i = threadIdx.x; i < P.n_leaf; i += 1024) me.flags[parity][i] = (den_touched[i] | k0_touched[i]) != 0;
    if (threadIdx.x == 0) running = 0;
    __syncthreads();   // the st.release.sys below is cumulative over the CTA's flag writes ordered by this barrier
    if (threadIdx.x < P.world) {
        signal_peer(P, threadIdx.x, SIG_A, epoch);
        wait_peer(P, threadIdx.x, SIG_A, epoch);
    }
    __syncthreads();
    const int32_t* pf[8];
    for (int r = 0; r < 8; ++r) pf[r] = view(P.base[r < P.world ? r : 0], P.n_leaf, P.cap_leaves).flags[parity];
    for (int base = 0; base < P.n_leaf; base += 1024) {
        const int i = base + threadIdx.x;
        int f = 0;
        if (i < P.n_leaf)
            for (int r = 0; r < P.world; ++r) f |= __ldcv(pf[r] + i);
        const bool t = f != 0;
        if (i < P.n_leaf) { den_touched[i] = t; k0_touched[i] = t; }
        const unsigned bits = __ballot_sync(0xffffffffu, t);
        if (lane == 0) warp_cnt[warp] = __popc(bits);
        __syncthreads();
        const int c = warp_cnt[lane];
        int incl = c;
        #pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const int o = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += o; }
        const int warp_off = __shfl_sync(0xffffffffu, incl - c, warp);
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        const int start = running;
        if (t) {
            const int slot = start + warp_off + __popc(bits & ((1u << lane) - 1));
            den_list[slot] = i;
            k0_list[slot] = i;
        }
        __syncthreads();
        if (threadIdx.x == 0) running = start + total;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        int n = running;
        if (n > P.cap_leaves) { atomicExch(me.err, 2); n = P.cap_leaves; }
        counters[cnt_den] = n;
        counters[cnt_k0] = n;
    }
}

__global__ void __launch_bounds__(256) k_dp_pack(pvdb_dp_peers P, uint32_t epoch, int parity, const float* __restrict__ den_grad,
                                                 const float* __restrict__ k0_grad,
                                                 const int32_t* __restrict__ list, const int32_t* __restrict__ counters, int cnt_den) {
    pvdb_pdl_wait();
    const Blk me = view(P.base[P.rank], P.n_leaf, P.cap_leaves);
    const int n = counters[cnt_den];
    float* buf = me.grad[parity];
    constexpr int T4 = TILE_F / 4;
    const int total = n * T4;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const int slot = idx / T4, i = idx - slot * T4;
        const int leaf = list[slot];
        const float4 v = i < 128 ? reinterpret_cast<const float4*>(den_grad + (size_t)leaf * 512)[i]
                                 : reinterpret_cast<const float4*>(k0_grad + (size_t)leaf * 512 * 12)[i - 128];
        reinterpret_cast<float4*>(buf)[idx] = v;
    }
    // last CTA out tells every peer that this rank's tiles are in place (threadFenceReduction pattern; the final
    // st.release.sys is cumulative over everything the counter made visible)
    __syncthreads();
    __shared__ bool last;
    if (threadIdx.x == 0) {
        __threadfence();
        last = atomicAdd(me.done, 1u) == gridDim.x - 1;
        if (last) { *me.done = 0; __threadfence(); }
    }
    __syncthreads();
    if (last && threadIdx.x < P.world) signal_peer(P, threadIdx.x, SIG_B, epoch);
}

__global__ void __launch_bounds__(256) k_dp_reduce(pvdb_dp_peers P, uint32_t epoch, int parity, float* __restrict__ den_grad,
                                                   float* __restrict__ k0_grad,
                                                   const int32_t* __restrict__ list, const int32_t* __restrict__ counters, int cnt_den) {
    pvdb_pdl_wait();
    if (threadIdx.x < P.world) wait_peer(P, threadIdx.x, SIG_B, epoch);
    __syncthreads();
    const int n = counters[cnt_den];
    const float* pb[8];
    for (int r = 0; r < 8; ++r) pb[r] = view(P.base[r < P.world ? r : 0], P.n_leaf, P.cap_leaves).grad[parity];
    constexpr int T4 = TILE_F / 4;
    const int total = n * T4;
    // plain loads: these peer addresses were last read two steps ago in another launch, and the acquire + barrier above
    // orders them after the peers' packs
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        float4 v[8];
        #pragma unroll
        for (int r = 0; r < 8; ++r)
            if (r < P.world) v[r] = reinterpret_cast<const float4*>(pb[r])[idx];
        float4 s = v[0];
        #pragma unroll
        for (int r = 1; r < 8; ++r)
            if (r < P.world) { s.x += v[r].x; s.y += v[r].y; s.z += v[r].z; s.w += v[r].w; }
        const int slot = idx / T4, i = idx - slot * T4;
        const int leaf = list[slot];
        if (i < 128) reinterpret_cast<float4*>(den_grad + (size_t)leaf * 512)[i] = s;
        else reinterpret_cast<float4*>(k0_grad + (size_t)leaf * 512 * 12)[i - 128] = s;
    }
}

// rgbnet gradients: CTA `g` owns slice g of the 22 019 values end to end (publish, signal, wait, sum), so there is no
// grid-wide dependency and a single latency of peer loads.
__global__ void __launch_bounds__(1024) k_dp_net(pvdb_dp_peers P, uint32_t epoch, int parity, float* __restrict__ net_grad) {
    constexpr int SL = NET_PAD / NET_SLICES;       // floats per slice (multiple of 4)
    pvdb_pdl_wait();
    const int g = blockIdx.x;
    const Blk me = view(P.base[P.rank], P.n_leaf, P.cap_leaves);
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
        signal_peer(P, threadIdx.x, SIG_C + g * 8, epoch);
        wait_peer(P, threadIdx.x, SIG_C + g * 8, epoch);
    }
    __syncthreads();
    if (!live) return;
    float4 v[8];
#pragma unroll
    for (int r = 0; r < 8; ++r)
        if (r < P.world) v[r] = r == P.rank ? own : *reinterpret_cast<const float4*>(view(P.base[r], P.n_leaf, P.cap_leaves).net[parity] + i);
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
    PVDB_CHECK_ARG(P->n_leaf == b->tree->n_leaf && P->cap_leaves >= 1, "peers block was sized for another tree");
    for (int r = 0; r < P->world; ++r) PVDB_CHECK_ARG(P->base[r], "peer block not mapped");
    return PVDB_OK;
}

}  // namespace

extern "C" size_t pvdb_dp_symm_bytes(int n_leaf, int cap_leaves) {
    if (n_leaf < 0 || cap_leaves < 0) return 0;
    return grad_off(n_leaf) + 2 * cap_floats(cap_leaves) * sizeof(float);
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
    PVDB_CUDA(cudaMemcpy(err_out, view(P->base[P->rank], P->n_leaf, P->cap_leaves).err, sizeof(int32_t), cudaMemcpyDeviceToHost));
    return PVDB_OK;
}

extern "C" int pvdb_dp_exchange_tiles(const pvdb_dp_peers* P, const pvdb_train_bufs* b, uint32_t step, void* stream) {
    if (int rc = check_peers(P, b)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int parity = step & 1;
    const uint32_t epoch = step + 1;                        // monotone; the signal words start at 0
    const int CNT_DEN = 2, CNT_K0 = 4;                      // counters[] slots of pvdb_train_bufs (include/plenvdb_b200.h)
    PVDB_CUDA(pvdb_launch_pdl(k_dp_union, dim3(1), dim3(1024), 0, st, *P, epoch, parity, b->den_touched, b->k0_touched, b->den_touched_list,
                              b->k0_touched_list, b->counters, CNT_DEN, CNT_K0));
    PVDB_LAUNCH_CHECK();
    pvdb_prof_mark("dp_union", st);
    PVDB_CUDA(pvdb_launch_pdl(k_dp_pack, dim3(PVDB_SMS), dim3(256), 0, st, *P, epoch, parity, (const float*)b->den_grad, (const float*)b->k0_grad,
                              (const int32_t*)b->den_touched_list, (const int32_t*)b->counters, CNT_DEN));
    PVDB_LAUNCH_CHECK();
    pvdb_prof_mark("dp_pack", st);
    PVDB_CUDA(pvdb_launch_pdl(k_dp_reduce, dim3(PVDB_SMS * 2), dim3(256), 0, st, *P, epoch, parity, b->den_grad, b->k0_grad,
                              (const int32_t*)b->den_touched_list, (const int32_t*)b->counters, CNT_DEN));
    PVDB_LAUNCH_CHECK();
    pvdb_prof_mark("dp_reduce", st);
    return PVDB_OK;
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
