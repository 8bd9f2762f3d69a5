// leaf_local.cu — the two mechanisms north_star words for the k0 colour features, as an alternative path of the fused step:
//
//   "trilinear interpolation ... from leaf nodes staged in shared memory"      -> k_ll_gather
//   "leaf-local gradient accumulation replaces global atomics"                 -> k_ll_scatter
//
// (reference: colorvdb.cu:81-111 forward, :28-37 + :130-160 backward).  Both need the kept samples grouped by leaf, which the
// default path never does (it reaches every corner through the record id the march saved: 24 independent 16-byte loads per
// sample in the forward, 24 vector reductions at L2 in the backward).  Here every kept sample gets a HOME leaf — the leaf of
// its first existing corner — and a counting sort (count, scan, fill) lists the samples of each leaf; then one CTA per leaf
// with samples
//   forward : pulls the leaf's 512 x 12 values (24 KB, contiguous) into shared memory with one bulk async copy and evaluates
//             the trilinear sum of each home sample in the reference's corner order, taking a corner from shared memory when
//             it lies in the home leaf (all eight do for 2/3 of the samples) and from global memory otherwise — same values,
//             same order, so k_feat is bit-identical to the default path; the rgbnet forward then reads k_feat;
//   backward: accumulates dL/dk0 of its home samples into a 24 KB shared tile with shared-memory atomics (corners in a
//             neighbouring leaf: vector reductions at L2 as before), then adds the non-zero part of the tile to the leaf's
//             gradient plane with 16-byte reductions.
// Measured against the default path at F160 and S512 in profiles/prof_leaf_local_r02.md; selected with PVDB_LEAF_LOCAL=1 /
// pvdb_debug_set_leaf_local (default: off, see there).
#include "common.cuh"
#include "tc_ptx.cuh"
#include "leaf_local.cuh"

namespace {

constexpr int CNT_M_KEEP = 1, CNT_N_TOUCHED_K0 = 4, CNT_LL_LEAVES = 8;
constexpr int TILE_FLOATS = PVDB_LEAF_VOX * 12;

__device__ __forceinline__ int home_leaf(const int* rec) {
#pragma unroll
    for (int q = 0; q < 8; ++q)
        if (rec[q] >= 0) return rec[q] >> 9;
    return -1;
}
__device__ __forceinline__ void load_rec(const int32_t* __restrict__ k_corner, int64_t s, int* rec) {
    const int4 a = __ldg(reinterpret_cast<const int4*>(k_corner + s * 8)), b = __ldg(reinterpret_cast<const int4*>(k_corner + s * 8) + 1);
    rec[0] = a.x; rec[1] = a.y; rec[2] = a.z; rec[3] = a.w; rec[4] = b.x; rec[5] = b.y; rec[6] = b.z; rec[7] = b.w;
}

// MODE 0: histogram of the home leaves (a sample without any corner gets its zero feature row here); MODE 1: fill the buckets
template <int MODE>
__global__ void __launch_bounds__(256) k_ll_bucket(const int32_t* __restrict__ counters, int64_t cap_keep, const int32_t* __restrict__ k_corner,
                                                   int32_t* __restrict__ cnt, const int32_t* __restrict__ off, int32_t* __restrict__ cur,
                                                   int32_t* __restrict__ items, float* __restrict__ k_feat) {
    pvdb_pdl_wait();
    const int64_t M = min((int64_t)counters[CNT_M_KEEP], cap_keep);
    for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < M; s += (int64_t)gridDim.x * blockDim.x) {
        int rec[8];
        load_rec(k_corner, s, rec);
        const int h = home_leaf(rec);
        if (h < 0) {
            if (MODE == 0) {
                float4* kf = reinterpret_cast<float4*>(k_feat + s * 12);
                kf[0] = kf[1] = kf[2] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            continue;
        }
        if (MODE == 0) atomicAdd(cnt + h, 1);
        else items[off[h] + atomicAdd(cur + h, 1)] = (int32_t)s;
    }
}

// One CTA: exclusive scan of the per-leaf counts, ascending list of the leaves that have samples, zeroed fill cursors.
__global__ void __launch_bounds__(1024) k_ll_scan(const int32_t* __restrict__ cnt, int n_leaf, int32_t* __restrict__ off, int32_t* __restrict__ cur,
                                                  int32_t* __restrict__ list, int32_t* __restrict__ counters) {
    pvdb_pdl_wait();
    __shared__ int2 wtot[32];
    __shared__ int2 carry_s;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = make_int2(0, 0);
    __syncthreads();
    for (int base = 0; base < n_leaf; base += 1024) {
        const int i = base + threadIdx.x;
        const int c = i < n_leaf ? cnt[i] : 0;
        int2 v = make_int2(c, c > 0 ? 1 : 0);
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int ux = __shfl_up_sync(0xffffffffu, v.x, o), uy = __shfl_up_sync(0xffffffffu, v.y, o);
            if (lane >= o) { v.x += ux; v.y += uy; }
        }
        if (lane == 31) wtot[wid] = v;
        __syncthreads();
        const int2 carry = carry_s;
        if (wid == 0) {
            const int2 own = wtot[lane];
            int2 w = own;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int ux = __shfl_up_sync(0xffffffffu, w.x, o), uy = __shfl_up_sync(0xffffffffu, w.y, o);
                if (lane >= o) { w.x += ux; w.y += uy; }
            }
            wtot[lane] = make_int2(w.x - own.x, w.y - own.y);
            if (lane == 31) carry_s = make_int2(carry.x + w.x, carry.y + w.y);
        }
        __syncthreads();
        const int2 pre = wtot[wid];
        if (i < n_leaf) {
            off[i] = carry.x + pre.x + v.x - c;
            cur[i] = 0;
            if (c > 0) list[carry.y + pre.y + v.y - 1] = i;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) { off[n_leaf] = carry_s.x; counters[CNT_LL_LEAVES] = carry_s.y; }
}

// Forward: CTA per leaf with samples; the leaf's k0 values staged in shared memory by one bulk async copy.
__global__ void __launch_bounds__(256) k_ll_gather(const float* __restrict__ k0, const float* __restrict__ k_xyz, const int32_t* __restrict__ k_corner,
                                                   const int32_t* __restrict__ off, const int32_t* __restrict__ items, const int32_t* __restrict__ list,
                                                   const int32_t* __restrict__ counters, float* __restrict__ k_feat) {
    extern __shared__ __align__(128) unsigned char smem[];
    float4* tile = reinterpret_cast<float4*>(smem);
    const uint32_t bar = smem_u32(smem + TILE_FLOATS * 4);
    pvdb_pdl_wait();
    if (threadIdx.x == 0) { mbar_init(bar, 1); fence_async_smem(); }
    __syncthreads();
    const int n_list = counters[CNT_LL_LEAVES];
    uint32_t parity = 0;
    for (int li = blockIdx.x; li < n_list; li += gridDim.x) {
        const int leaf = list[li];
        if (threadIdx.x == 0) {
            mbar_expect_tx(bar, TILE_FLOATS * 4);
            bulk_g2s(smem_u32(smem), k0 + (size_t)leaf * TILE_FLOATS, TILE_FLOATS * 4, bar);
        }
        const int b = off[leaf], e = off[leaf + 1];
        // the samples' own data is loaded while the tile is in flight
        for (int base = b; base < e; base += blockDim.x) {
            const int it = base + threadIdx.x;
            int rec[8];
            int64_t s = -1;
            PvdbTri tri;
            if (it < e) {
                s = items[it];
                load_rec(k_corner, s, rec);
                const float* p = k_xyz + s * 3;
                tri.set(p[0], p[1], p[2]);
            }
            if (base == b) { mbar_wait(bar, parity); parity ^= 1; }
            if (s < 0) continue;
            float x[12];
#pragma unroll
            for (int c = 0; c < 12; ++c) x[c] = 0.f;
            // colorvdb.cu:81-111 arithmetic and corner order; a missing corner contributes fma(sc, 0, x) = x
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int dx = PVDB_CORNER[q][0], dy = PVDB_CORNER[q][1], dz = PVDB_CORNER[q][2];
                const float sc = __fmul_rn(__fmul_rn(tri.f(0, dx), tri.f(1, dy)), tri.f(2, dz));
                if (rec[q] < 0) continue;
                float4 v[3];
                if ((rec[q] >> 9) == leaf) {
                    const float4* src = tile + (rec[q] & 511) * 3;
                    v[0] = src[0]; v[1] = src[1]; v[2] = src[2];
                } else {
                    const float4* src = reinterpret_cast<const float4*>(k0 + (size_t)rec[q] * 12);
                    v[0] = __ldg(src); v[1] = __ldg(src + 1); v[2] = __ldg(src + 2);
                }
#pragma unroll
                for (int c4 = 0; c4 < 3; ++c4) {
                    x[c4 * 4 + 0] = __fmaf_rn(sc, v[c4].x, x[c4 * 4 + 0]); x[c4 * 4 + 1] = __fmaf_rn(sc, v[c4].y, x[c4 * 4 + 1]);
                    x[c4 * 4 + 2] = __fmaf_rn(sc, v[c4].z, x[c4 * 4 + 2]); x[c4 * 4 + 3] = __fmaf_rn(sc, v[c4].w, x[c4 * 4 + 3]);
                }
            }
            float4* kf = reinterpret_cast<float4*>(k_feat + s * 12);
            kf[0] = make_float4(x[0], x[1], x[2], x[3]); kf[1] = make_float4(x[4], x[5], x[6], x[7]); kf[2] = make_float4(x[8], x[9], x[10], x[11]);
        }
        if (b == e) { mbar_wait(bar, parity); parity ^= 1; }      // (cannot happen: listed leaves have samples) keep the barrier phase in step
        __syncthreads();      // every thread is done with the tile before the next leaf's copy overwrites it
    }
}

__device__ __forceinline__ void ll_red_add4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// Backward: CTA per leaf with samples; gradients of the home samples accumulate in a shared 512 x 12 tile.
__global__ void __launch_bounds__(256) k_ll_scatter(float* __restrict__ k0_grad, const float* __restrict__ k_dx, const float* __restrict__ k_xyz,
                                                    const int32_t* __restrict__ k_corner, const int32_t* __restrict__ off, const int32_t* __restrict__ items,
                                                    const int32_t* __restrict__ list, const int32_t* __restrict__ counters, int32_t* __restrict__ k0_touched,
                                                    int32_t* __restrict__ k0_touched_list, int32_t* __restrict__ counters_w) {
    extern __shared__ __align__(128) unsigned char smem[];
    float* tile = reinterpret_cast<float*>(smem);
    pvdb_pdl_wait();
    const int n_list = counters[CNT_LL_LEAVES];
    for (int li = blockIdx.x; li < n_list; li += gridDim.x) {
        const int leaf = list[li];
        for (int i = threadIdx.x; i < TILE_FLOATS / 4; i += blockDim.x) reinterpret_cast<float4*>(tile)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        __syncthreads();
        const int b = off[leaf], e = off[leaf + 1];
        // thread = (home sample, half of the corners): 2 threads per sample, 4 corners each
        for (int w = b * 2 + threadIdx.x; w < e * 2; w += blockDim.x) {
            const int64_t s = items[w >> 1];
            const int half = w & 1;
            int rec[8];
            load_rec(k_corner, s, rec);
            const float* p = k_xyz + s * 3;
            PvdbTri tri;
            tri.set(p[0], p[1], p[2]);
            const float4* g4 = reinterpret_cast<const float4*>(k_dx + s * 12);
            const float4 ga = __ldg(g4), gb = __ldg(g4 + 1), gc = __ldg(g4 + 2);
            const float g[12] = {ga.x, ga.y, ga.z, ga.w, gb.x, gb.y, gb.z, gb.w, gc.x, gc.y, gc.z, gc.w};
#pragma unroll
            for (int qq = 0; qq < 4; ++qq) {
                const int q = half * 4 + qq;
                if (rec[q] < 0) continue;
                const int dx = PVDB_CORNER[q][0], dy = PVDB_CORNER[q][1], dz = PVDB_CORNER[q][2];
                const float sc = __fmul_rn(__fmul_rn(tri.f(0, dx), tri.f(1, dy)), tri.f(2, dz));
                if ((rec[q] >> 9) == leaf) {
                    float* dst = tile + (rec[q] & 511) * 12;
#pragma unroll
                    for (int c = 0; c < 12; ++c) atomicAdd(dst + c, __fmul_rn(g[c], sc));
                } else {
                    float* dst = k0_grad + (size_t)rec[q] * 12;
#pragma unroll
                    for (int c4 = 0; c4 < 3; ++c4)
                        ll_red_add4(dst + c4 * 4, __fmul_rn(g[c4 * 4], sc), __fmul_rn(g[c4 * 4 + 1], sc), __fmul_rn(g[c4 * 4 + 2], sc), __fmul_rn(g[c4 * 4 + 3], sc));
                    pvdb_touch_leaf(k0_touched, k0_touched_list, counters_w + CNT_N_TOUCHED_K0, rec[q] >> 9);
                }
            }
        }
        __syncthreads();
        // flush: the non-zero 16-byte groups of the tile are added to the leaf's plane (neighbouring leaves' CTAs may be adding
        // their boundary corners to the same plane at the same time, hence reductions, not stores)
        float* plane = k0_grad + (size_t)leaf * TILE_FLOATS;
        for (int i = threadIdx.x; i < TILE_FLOATS / 4; i += blockDim.x) {
            const float4 v = reinterpret_cast<const float4*>(tile)[i];
            if (v.x != 0.f || v.y != 0.f || v.z != 0.f || v.w != 0.f) ll_red_add4(plane + i * 4, v.x, v.y, v.z, v.w);
        }
        if (threadIdx.x == 0) pvdb_touch_leaf(k0_touched, k0_touched_list, counters_w + CNT_N_TOUCHED_K0, leaf);
        __syncthreads();
    }
}

constexpr int GATHER_SMEM = TILE_FLOATS * 4 + 16;
constexpr int SCATTER_SMEM = TILE_FLOATS * 4;

}  // namespace

static int g_leaf_local = -1;
extern "C" void pvdb_debug_set_leaf_local(int on) { g_leaf_local = on ? 1 : 0; }
bool pvdb_leaf_local_enabled(const pvdb_train_bufs* b) {
    if (g_leaf_local < 0) { const char* e = getenv("PVDB_LEAF_LOCAL"); g_leaf_local = e && atoi(e) != 0; }
    return g_leaf_local && b->ll_cnt && b->ll_off && b->ll_cur && b->ll_items && b->ll_list && b->k_dx;
}

// After the emit kernel: group the kept samples by home leaf and gather k_feat through leaf tiles staged in shared memory.
int pvdb_leaf_local_forward(const pvdb_train_bufs* b, cudaStream_t st) {
    const int n_leaf = b->tree->n_leaf;
    PVDB_CUDA(cudaMemsetAsync(b->ll_cnt, 0, (size_t)(n_leaf > 0 ? n_leaf : 1) * sizeof(int32_t), st));
    k_ll_bucket<0><<<PVDB_SMS * 4, 256, 0, st>>>(b->counters, b->cap_keep, b->k_corner, b->ll_cnt, nullptr, nullptr, nullptr, b->k_feat);
    PVDB_LAUNCH_CHECK();
    pvdb_prof_mark("ll_count", st);
    PVDB_CUDA(pvdb_launch_pdl(k_ll_scan, dim3(1), dim3(1024), 0, st, (const int32_t*)b->ll_cnt, n_leaf, b->ll_off, b->ll_cur, b->ll_list, b->counters));
    PVDB_LAUNCH_CHECK();
    PVDB_CUDA(pvdb_launch_pdl(k_ll_bucket<1>, dim3(PVDB_SMS * 4), dim3(256), 0, st, (const int32_t*)b->counters, b->cap_keep, (const int32_t*)b->k_corner,
                              b->ll_cnt, (const int32_t*)b->ll_off, b->ll_cur, b->ll_items, b->k_feat));
    PVDB_LAUNCH_CHECK();
    pvdb_prof_mark("ll_fill", st);
    static bool attr = false;
    if (!attr) {
        PVDB_CUDA(cudaFuncSetAttribute(k_ll_gather, cudaFuncAttributeMaxDynamicSharedMemorySize, GATHER_SMEM));
        PVDB_CUDA(cudaFuncSetAttribute(k_ll_scatter, cudaFuncAttributeMaxDynamicSharedMemorySize, SCATTER_SMEM));
        attr = true;
    }
    PVDB_CUDA(pvdb_launch_pdl(k_ll_gather, dim3(PVDB_SMS * 8), dim3(256), (size_t)GATHER_SMEM, st, (const float*)b->k0, (const float*)b->k_xyz,
                              (const int32_t*)b->k_corner, (const int32_t*)b->ll_off, (const int32_t*)b->ll_items, (const int32_t*)b->ll_list,
                              (const int32_t*)b->counters, b->k_feat));
    PVDB_LAUNCH_CHECK();
    pvdb_prof_mark("ll_gather", st);
    return PVDB_OK;
}

// After the activation-gradient kernel stored dL/dx (k_dx): leaf-local accumulation into the k0 gradient planes.
int pvdb_leaf_local_backward(const pvdb_train_bufs* b, cudaStream_t st) {
    PVDB_CUDA(pvdb_launch_pdl(k_ll_scatter, dim3(PVDB_SMS * 8), dim3(256), (size_t)SCATTER_SMEM, st, b->k0_grad, (const float*)b->k_dx, (const float*)b->k_xyz,
                              (const int32_t*)b->k_corner, (const int32_t*)b->ll_off, (const int32_t*)b->ll_items, (const int32_t*)b->ll_list,
                              (const int32_t*)b->counters, b->k0_touched, b->k0_touched_list, b->counters));
    PVDB_LAUNCH_CHECK();
    pvdb_prof_mark("ll_scatter", st);
    return PVDB_OK;
}
