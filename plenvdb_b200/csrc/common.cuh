// common.cuh — device-side tree lookup, exact-rounding math helpers and error plumbing shared by all
// kernels of libplenvdb_b200.  sm_100a only.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/plenvdb_b200.h"

#define PVDB_SMS 148          // B200 SM count; grids are sized in multiples of it
#define PVDB_LEAF_VOX 512

void pvdb_set_error(const char* fmt, ...);
void pvdb_count_launch(int n = 1);
void pvdb_reset_launch_count();
// Optional per-kernel timing for bench.py's roofline line: when enabled, pvdb_prof_mark() records a CUDA event on the
// launching stream after each kernel of a fused call; pvdb_profile_fetch() returns the elapsed ms between marks.
void pvdb_prof_begin(cudaStream_t st);
void pvdb_prof_mark(const char* name, cudaStream_t st);
bool pvdb_prof_active();

#define PVDB_CHECK_ARG(cond, msg)                 \
    do {                                          \
        if (!(cond)) {                            \
            pvdb_set_error("%s: %s", __func__, msg); \
            return PVDB_ERR_ARG;                  \
        }                                         \
    } while (0)

#define PVDB_CUDA(call)                                                                  \
    do {                                                                                 \
        cudaError_t e__ = (call);                                                        \
        if (e__ != cudaSuccess) {                                                        \
            pvdb_set_error("%s: %s (%s:%d)", __func__, cudaGetErrorString(e__), __FILE__, __LINE__); \
            return PVDB_ERR_CUDA;                                                        \
        }                                                                                \
    } while (0)

// After a launch: surface configuration errors without synchronising.
#define PVDB_LAUNCH_CHECK()                                                              \
    do {                                                                                 \
        pvdb_count_launch();                                                             \
        cudaError_t e__ = cudaPeekAtLastError();                                         \
        if (e__ != cudaSuccess) {                                                        \
            pvdb_set_error("%s: launch failed: %s (%s:%d)", __func__, cudaGetErrorString(e__), __FILE__, __LINE__); \
            return PVDB_ERR_CUDA;                                                        \
        }                                                                                \
    } while (0)

// Programmatic dependent launch (PDL): the kernels of the fused step's main chain are launched with programmatic stream
// serialisation and block at pvdb_pdl_wait (first statement) until the kernel before them has completed and its memory is
// visible: the launch itself no longer waits for the predecessor's completion handshake, which takes ~1.5 us off each of
// the ten kernel boundaries of a step (measured: 0.224 -> 0.209 ms).  Letting the dependents start even earlier
// (griddepcontrol.launch_dependents at the top of every kernel: 0.233 ms; only in the two small kernels that precede a
// tensor-core kernel, so that its weight-image TMA overlaps their tail: 0.204 vs 0.2005 ms) was measured slower: the
// early-resident CTAs of the next kernel take registers and thread slots from the running one.  The wait is a no-op in a normally launched
// kernel.
__device__ __forceinline__ void pvdb_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
template <typename... KArgs, typename... Args>
static inline cudaError_t pvdb_launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

static inline int pvdb_grid_for(int64_t n, int block) {
    int64_t g = (n + block - 1) / block;
    if (g < 1) g = 1;
    return (int)g;
}

// ---------------------------------------------------------------------------------------------
// Tree lookup.  Offsets follow NanoVDB.h: root key :2702-2709, upper CoordToOffset (LOG2DIM 5, child
// TOTAL 7) :3377-3385, lower (LOG2DIM 4, child TOTAL 3), leaf :3893-3900.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t pvdb_root_key(int x, int y, int z) {
    return (uint64_t)((uint32_t)z >> 12) | ((uint64_t)((uint32_t)y >> 12) << 21) | ((uint64_t)((uint32_t)x >> 12) << 42);
}
__device__ __forceinline__ int pvdb_upper_off(int x, int y, int z) {
    return (((x & 4095) >> 7) << 10) | (((y & 4095) >> 7) << 5) | ((z & 4095) >> 7);
}
__device__ __forceinline__ int pvdb_lower_off(int x, int y, int z) {
    return (((x & 127) >> 3) << 8) | (((y & 127) >> 3) << 4) | ((z & 127) >> 3);
}
__device__ __forceinline__ int pvdb_leaf_off(int x, int y, int z) {
    return ((x & 7) << 6) | ((y & 7) << 3) | (z & 7);
}

// Leaf index containing (x,y,z) or -1.  Top-down like ReadAccessor::getValue on a cold cache
// (NanoVDB.h:2924-2948 findTile, :2998-3010, :3327-3335) but over int32 tables.
__device__ __forceinline__ int pvdb_find_leaf(const pvdb_tree& t, int x, int y, int z) {
    const uint64_t key = pvdb_root_key(x, y, z);
    int u = 0;
    if (key != t.root_key0) {
        u = -1;
        for (int i = 1; i < t.n_upper; ++i)
            if (__ldg(t.root_keys + i) == key) { u = i; break; }
        if (u < 0) return -1;
    } else if (t.n_upper == 0) {
        return -1;
    }
    const int l = __ldg(t.upper_child + (size_t)u * 32768 + pvdb_upper_off(x, y, z));
    if (l < 0) return -1;
    return __ldg(t.lower_child + (size_t)l * 4096 + pvdb_lower_off(x, y, z));
}

// One-entry leaf cache, the register-resident analogue of ReadAccessor's leaf-level key (NanoVDB.h:4514-4749).
struct PvdbLeafCache {
    int kx, ky, kz, leaf;
    __device__ __forceinline__ PvdbLeafCache() : kx(INT32_MIN), ky(0), kz(0), leaf(-1) {}
    __device__ __forceinline__ int find(const pvdb_tree& t, int x, int y, int z) {
        if ((((x ^ kx) | (y ^ ky) | (z ^ kz)) & ~7) == 0) return leaf;
        kx = x & ~7; ky = y & ~7; kz = z & ~7;
        leaf = pvdb_find_leaf(t, x, y, z);
        return leaf;
    }
};

__device__ __forceinline__ bool pvdb_mask_bit(const uint64_t* leaf_mask, int leaf, int off) {
    return (__ldg(leaf_mask + (size_t)leaf * 8 + (off >> 6)) >> (off & 63)) & 1ull;
}

// The reference's corner walk (densityvdb.cu:116-123): 000,001,011,010,110,111,101,100 as (dx,dy,dz).
__device__ __constant__ const int8_t PVDB_CORNER[8][3] = {
    {0, 0, 0}, {0, 0, 1}, {0, 1, 1}, {0, 1, 0}, {1, 1, 0}, {1, 1, 1}, {1, 0, 1}, {1, 0, 0}};

// Per-corner weight factors in the reference's multiplication order: w = f0 * f1 * f2 with
// f_axis = (d_axis ? u : 1-u).
struct PvdbTri {
    int i, j, k;
    float u[3], m[3];   // u = frac, m = 1-u (sub.f32 as compiled from `1-uvw[a]`)
    __device__ __forceinline__ void set(float x, float y, float z) {
        const float fx = floorf(x), fy = floorf(y), fz = floorf(z);
        i = (int)fx; j = (int)fy; k = (int)fz;                       // cvt.rmi then cvt.rzi.s32 (NanoVDB Floor)
        u[0] = __fsub_rn(x, (float)i); u[1] = __fsub_rn(y, (float)j); u[2] = __fsub_rn(z, (float)k);
        m[0] = __fsub_rn(1.0f, u[0]); m[1] = __fsub_rn(1.0f, u[1]); m[2] = __fsub_rn(1.0f, u[2]);
    }
    __device__ __forceinline__ float f(int axis, int d) const { return d ? u[axis] : m[axis]; }
};

// Flag a leaf as touched by a gradient scatter and append it once to the touched-leaf list.  The plain read filters
// almost every call; the exchange elects the single appender.
__device__ __forceinline__ void pvdb_touch_leaf(int32_t* __restrict__ flags, int32_t* __restrict__ list, int32_t* __restrict__ count, int leaf) {
    if (flags[leaf] == 0 && atomicExch(flags + leaf, 1) == 0) list[atomicAdd(count, 1)] = leaf;
}
