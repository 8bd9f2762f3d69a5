// grid_ops.cu — drop-in grid operators of the `plenvdb` module on the B200-native tree:
// trilinear forward / gradient scatter (D1,D2,C1,C2), sparse Adam (O1), zero_grad (O2) and
// dense<->sparse copies (O3).  Arithmetic order follows the reference kernels as compiled by nvcc
// (FMA contraction read from their PTX) so interpolated values are bit-identical and every
// threshold downstream lands on the same side.
#include "common.cuh"

// ---------------------------------------------------------------------------------------------
// D1 / C1 — forward (densityvdb.cu:101-125, colorvdb.cu:81-111, 16-26)
// ---------------------------------------------------------------------------------------------
template <int C>
__global__ void __launch_bounds__(256) k_sample_forward(pvdb_tree t, const float* __restrict__ plane,
                                                        const float* __restrict__ xs, const float* __restrict__ ys,
                                                        const float* __restrict__ zs, int64_t n,
                                                        float* __restrict__ out, int32_t* __restrict__ corner_leaf,
                                                        int32_t* __restrict__ corner_off) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    PvdbTri tri;
    tri.set(xs[s], ys[s], zs[s]);
    PvdbLeafCache cache;
    float acc[C];
#pragma unroll
    for (int c = 0; c < C; ++c) acc[c] = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const int dx = PVDB_CORNER[q][0], dy = PVDB_CORNER[q][1], dz = PVDB_CORNER[q][2];
        const int x = tri.i + dx, y = tri.j + dy, z = tri.k + dz;
        const int leaf = cache.find(t, x, y, z);
        const int off = pvdb_leaf_off(x, y, z);
        if (corner_leaf) corner_leaf[s * 8 + q] = leaf;
        if (corner_off) corner_off[s * 8 + q] = off;
        const float f0 = tri.f(0, dx), f1 = tri.f(1, dy), f2 = tri.f(2, dz);
        if (C == 1) {
            // res += v*f0*f1*f2  ->  fma(f2, f1*(f0*v), res)
            const float v = leaf >= 0 ? __ldg(plane + (size_t)leaf * 512 + off) : 0.f;
            acc[0] = __fmaf_rn(f2, __fmul_rn(f1, __fmul_rn(f0, v)), acc[0]);
        } else {
            // scale = f0*f1*f2 formed first, res[c] = fma(scale, v[c], res[c])
            const float sc = __fmul_rn(__fmul_rn(f0, f1), f2);
            if (leaf >= 0) {
                const float* pv = plane + ((size_t)leaf * 512 + off) * C;
                if constexpr (C % 4 == 0) {   // 16 B-aligned voxel records (C = 12: 48 B)
                    const float4* p = reinterpret_cast<const float4*>(pv);
#pragma unroll
                    for (int c4 = 0; c4 < C / 4; ++c4) {
                        const float4 v = __ldg(p + c4);
                        acc[c4 * 4 + 0] = __fmaf_rn(sc, v.x, acc[c4 * 4 + 0]);
                        acc[c4 * 4 + 1] = __fmaf_rn(sc, v.y, acc[c4 * 4 + 1]);
                        acc[c4 * 4 + 2] = __fmaf_rn(sc, v.z, acc[c4 * 4 + 2]);
                        acc[c4 * 4 + 3] = __fmaf_rn(sc, v.w, acc[c4 * 4 + 3]);
                    }
                } else {
#pragma unroll
                    for (int c = 0; c < C; ++c) acc[c] = __fmaf_rn(sc, __ldg(pv + c), acc[c]);
                }
            }
        }
    }
#pragma unroll
    for (int c = 0; c < C; ++c) out[s * C + c] = acc[c];
}

// ---------------------------------------------------------------------------------------------
// D2 / C2 — gradient scatter (densityvdb.cu:143-167 + :16-27, colorvdb.cu:130-160 + :28-37).
// A corner contributes iff a leaf contains it (no value-mask test), like `accumulate`.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void pvdb_red_add(float* addr, float v) {
    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ void pvdb_red_add4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

template <int C>
__global__ void __launch_bounds__(256) k_sample_backward(pvdb_tree t, float* __restrict__ gplane,
                                                         const float* __restrict__ xs, const float* __restrict__ ys,
                                                         const float* __restrict__ zs, const float* __restrict__ gout,
                                                         int64_t n) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    PvdbTri tri;
    tri.set(xs[s], ys[s], zs[s]);
    PvdbLeafCache cache;
    float g[C];
#pragma unroll
    for (int c = 0; c < C; ++c) g[c] = gout[s * C + c];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const int dx = PVDB_CORNER[q][0], dy = PVDB_CORNER[q][1], dz = PVDB_CORNER[q][2];
        const int x = tri.i + dx, y = tri.j + dy, z = tri.k + dz;
        const int leaf = cache.find(t, x, y, z);
        if (leaf < 0) continue;
        const int off = pvdb_leaf_off(x, y, z);
        const float f0 = tri.f(0, dx), f1 = tri.f(1, dy), f2 = tri.f(2, dz);
        if (C == 1) {
            // grads[n]*f0*f1*f2, left to right
            pvdb_red_add(gplane + (size_t)leaf * 512 + off, __fmul_rn(__fmul_rn(__fmul_rn(g[0], f0), f1), f2));
        } else {
            const float sc = __fmul_rn(__fmul_rn(f0, f1), f2);
            float* p = gplane + ((size_t)leaf * 512 + off) * C;
            if (C % 4 == 0) {
#pragma unroll
                for (int c4 = 0; c4 < C / 4; ++c4)
                    pvdb_red_add4(p + c4 * 4, __fmul_rn(sc, g[c4 * 4]), __fmul_rn(sc, g[c4 * 4 + 1]),
                                  __fmul_rn(sc, g[c4 * 4 + 2]), __fmul_rn(sc, g[c4 * 4 + 3]));
            } else {
#pragma unroll
                for (int c = 0; c < C; ++c) pvdb_red_add(p + c, __fmul_rn(sc, g[c]));
            }
        }
    }
}

// forward_single (densityvdb.cu:376-390, colorvdb.cu:380-399)
template <int C>
__global__ void __launch_bounds__(256) k_sample_nearest(pvdb_tree t, const float* __restrict__ plane,
                                                        const int32_t* __restrict__ is, const int32_t* __restrict__ js,
                                                        const int32_t* __restrict__ ks, int64_t n, float* __restrict__ out) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const int x = is[s], y = js[s], z = ks[s];
    const int leaf = pvdb_find_leaf(t, x, y, z);
    const int off = pvdb_leaf_off(x, y, z);
#pragma unroll
    for (int c = 0; c < C; ++c) out[s * C + c] = leaf >= 0 ? plane[((size_t)leaf * 512 + off) * C + c] : 0.f;
}

// ---------------------------------------------------------------------------------------------
// O1 — sparse Adam on leaf planes (densityvdb.cu:185-330, colorvdb.cu:179-331).
//   m' = fma(1-b0, g, b0*m);  v' = fma(g, (1-b1)*g, b1*v);  p' = p - (stepsz*m')/(eps+sqrt(v'))
//   per-lr: p' = p - (m'*(stepsz*perlr))/(eps+sqrt(v'))
// One thread per (leaf voxel, group) where a group is 1 channel (density) or the 3 comps of a Vec3.
// ---------------------------------------------------------------------------------------------
template <int G>   // G = channels per skip group: 1 or 3
__global__ void __launch_bounds__(256) k_adam(pvdb_tree t, float* __restrict__ p, const float* __restrict__ g,
                                              float* __restrict__ m, float* __restrict__ v, int C, int mode,
                                              float stepsz, float eps, float b0, float b1,
                                              const float* __restrict__ perlr) {
    const int ngrp = C / G;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t total = (int64_t)t.n_leaf * 512 * ngrp;
    if (idx >= total) return;
    const int64_t vox = idx / ngrp;            // leaf*512 + off
    const int grp = (int)(idx - vox * ngrp);
    const int leaf = (int)(vox >> 9), off = (int)(vox & 511);
    if (!pvdb_mask_bit(t.leaf_mask, leaf, off)) return;
    const size_t base = (size_t)vox * C + (size_t)grp * G;
    float gg[G];
    bool allzero = true;
#pragma unroll
    for (int c = 0; c < G; ++c) { gg[c] = g[base + c]; allzero = allzero && (gg[c] == 0.0f); }
    if (mode == 1 && allzero) return;
    const float omb0 = __fsub_rn(1.0f, b0), omb1 = __fsub_rn(1.0f, b1);
    float st = stepsz;
    if (mode == 2) st = __fmul_rn(perlr[vox], stepsz);   // (vperlr*stepsz)*m'
#pragma unroll
    for (int c = 0; c < G; ++c) {
        const float nm = __fmaf_rn(omb0, gg[c], __fmul_rn(b0, m[base + c]));
        const float nv = __fmaf_rn(gg[c], __fmul_rn(omb1, gg[c]), __fmul_rn(b1, v[base + c]));
        m[base + c] = nm;
        v[base + c] = nv;
        const float num = (mode == 2) ? __fmul_rn(nm, st) : __fmul_rn(st, nm);
        p[base + c] = __fsub_rn(p[base + c], __fdiv_rn(num, __fadd_rn(eps, __fsqrt_rn(nv))));
    }
}

// O2 — zero_grad on active voxels only (densityvdb.cu:353-363)
__global__ void __launch_bounds__(256) k_zero_grad(pvdb_tree t, float* __restrict__ g, int C) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t total = (int64_t)t.n_leaf * 512 * C;
    if (idx >= total) return;
    const int64_t vox = idx / C;
    if (pvdb_mask_bit(t.leaf_mask, (int)(vox >> 9), (int)(vox & 511))) g[idx] = 0.f;
}

// O3 — copyFromDense (densityvdb.cu:31-49, colorvdb.cu:42-63): active voxels only.
__global__ void __launch_bounds__(256) k_copy_from_dense(pvdb_tree t, float* __restrict__ plane, int C,
                                                         const float* __restrict__ dense, int rx, int ry, int rz) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t total = (int64_t)t.n_leaf * 512 * C;
    if (idx >= total) return;
    const int64_t vox = idx / C;
    const int c = (int)(idx - vox * C);
    const int leaf = (int)(vox >> 9), off = (int)(vox & 511);
    if (!pvdb_mask_bit(t.leaf_mask, leaf, off)) return;
    const int x = t.leaf_origin[leaf * 3 + 0] + (off >> 6), y = t.leaf_origin[leaf * 3 + 1] + ((off >> 3) & 7),
              z = t.leaf_origin[leaf * 3 + 2] + (off & 7);
    plane[idx] = dense[(((int64_t)x * ry + y) * rz + z) * C + c];
}

// copyToDense (plenvdb.h:158-167 via OpenVDB tools::copyToDense): stored value wherever a leaf covers the
// voxel, background 0 elsewhere.  One thread per dense element.
__global__ void __launch_bounds__(256) k_copy_to_dense(pvdb_tree t, const float* __restrict__ plane, int C,
                                                       float* __restrict__ dense, int rx, int ry, int rz) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t total = (int64_t)rx * ry * rz * C;
    if (idx >= total) return;
    const int64_t vox = idx / C;
    const int c = (int)(idx - vox * C);
    const int z = (int)(vox % rz), y = (int)((vox / rz) % ry), x = (int)(vox / ((int64_t)rz * ry));
    const int leaf = pvdb_find_leaf(t, x, y, z);
    dense[idx] = leaf >= 0 ? plane[((size_t)leaf * 512 + pvdb_leaf_off(x, y, z)) * C + c] : 0.f;
}

// setValuesOn_bymask (densityvdb.cu:64-84)
__global__ void __launch_bounds__(256) k_set_on_by_mask(pvdb_tree t, float* __restrict__ plane,
                                                        const uint8_t* __restrict__ mask, float val, int rx, int ry, int rz) {
    const int64_t vox = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (vox >= (int64_t)t.n_leaf * 512) return;
    const int leaf = (int)(vox >> 9), off = (int)(vox & 511);
    if (!pvdb_mask_bit(t.leaf_mask, leaf, off)) return;
    const int x = t.leaf_origin[leaf * 3 + 0] + (off >> 6), y = t.leaf_origin[leaf * 3 + 1] + ((off >> 3) & 7),
              z = t.leaf_origin[leaf * 3 + 2] + (off & 7);
    if (mask[((int64_t)x * ry + y) * rz + z]) plane[vox] = val;
}

// ---------------------------------------------------------------------------------------------
// C-ABI
// ---------------------------------------------------------------------------------------------
#define DISPATCH_C(CH, ...)                                        \
    switch (CH) {                                                  \
        case 1: { constexpr int C = 1; __VA_ARGS__; } break;       \
        case 3: { constexpr int C = 3; __VA_ARGS__; } break;       \
        case 6: { constexpr int C = 6; __VA_ARGS__; } break;       \
        case 9: { constexpr int C = 9; __VA_ARGS__; } break;       \
        case 12: { constexpr int C = 12; __VA_ARGS__; } break;     \
        default: pvdb_set_error("%s: unsupported channel count %d (1,3,6,9,12)", __func__, CH); return PVDB_ERR_ARG; \
    }

extern "C" int pvdb_sample_forward(const pvdb_tree* tree, const float* plane, int channels, const float* xs,
                                   const float* ys, const float* zs, int64_t n, float* out, int32_t* corner_leaf,
                                   int32_t* corner_off, void* stream) {
    PVDB_CHECK_ARG(tree && plane && out, "null pointer");
    if (n <= 0) return PVDB_OK;
    PVDB_CHECK_ARG(xs && ys && zs, "null coordinates");
    cudaStream_t st = (cudaStream_t)stream;
    DISPATCH_C(channels, (k_sample_forward<C><<<pvdb_grid_for(n, 256), 256, 0, st>>>(*tree, plane, xs, ys, zs, n, out,
                                                                                      corner_leaf, corner_off)));
    PVDB_LAUNCH_CHECK();
    return PVDB_OK;
}

extern "C" int pvdb_sample_backward(const pvdb_tree* tree, float* grad_plane, int channels, const float* xs,
                                    const float* ys, const float* zs, const float* grad_out, int64_t n, void* stream) {
    PVDB_CHECK_ARG(tree && grad_plane, "null pointer");
    if (n <= 0) return PVDB_OK;
    PVDB_CHECK_ARG(xs && ys && zs && grad_out, "null inputs");
    cudaStream_t st = (cudaStream_t)stream;
    DISPATCH_C(channels,
               (k_sample_backward<C><<<pvdb_grid_for(n, 256), 256, 0, st>>>(*tree, grad_plane, xs, ys, zs, grad_out, n)));
    PVDB_LAUNCH_CHECK();
    return PVDB_OK;
}

extern "C" int pvdb_sample_nearest(const pvdb_tree* tree, const float* plane, int channels, const int32_t* is,
                                   const int32_t* js, const int32_t* ks, int64_t n, float* out, void* stream) {
    PVDB_CHECK_ARG(tree && plane && out, "null pointer");
    if (n <= 0) return PVDB_OK;
    cudaStream_t st = (cudaStream_t)stream;
    DISPATCH_C(channels, (k_sample_nearest<C><<<pvdb_grid_for(n, 256), 256, 0, st>>>(*tree, plane, is, js, ks, n, out)));
    PVDB_LAUNCH_CHECK();
    return PVDB_OK;
}

// Stream-ordered staging for the host-pointer (numpy) contract of B1.
struct PvdbStage {
    void* ptrs[8];
    int n = 0;
    cudaStream_t st;
    explicit PvdbStage(cudaStream_t s) : st(s) {}
    void* alloc(size_t bytes) {
        void* p = nullptr;
        if (cudaMallocAsync(&p, bytes ? bytes : 4, st) != cudaSuccess) return nullptr;
        ptrs[n++] = p;
        return p;
    }
    ~PvdbStage() {
        for (int i = 0; i < n; ++i) cudaFreeAsync(ptrs[i], st);
    }
};

extern "C" int pvdb_sample_forward_host(const pvdb_tree* tree, const float* plane, int channels, const float* xs,
                                        const float* ys, const float* zs, int64_t n, float* out, void* stream) {
    PVDB_CHECK_ARG(tree && plane, "null pointer");
    if (n <= 0) return PVDB_OK;
    PVDB_CHECK_ARG(xs && ys && zs && out, "null host buffers");
    cudaStream_t st = (cudaStream_t)stream;
    PvdbStage stage(st);
    const size_t pb = (size_t)n * sizeof(float);
    float* d_xyz = (float*)stage.alloc(3 * pb);
    float* d_out = (float*)stage.alloc(pb * channels);
    if (!d_xyz || !d_out) { pvdb_set_error("%s: cudaMallocAsync failed", __func__); return PVDB_ERR_CUDA; }
    PVDB_CUDA(cudaMemcpyAsync(d_xyz, xs, pb, cudaMemcpyHostToDevice, st));
    PVDB_CUDA(cudaMemcpyAsync(d_xyz + n, ys, pb, cudaMemcpyHostToDevice, st));
    PVDB_CUDA(cudaMemcpyAsync(d_xyz + 2 * n, zs, pb, cudaMemcpyHostToDevice, st));
    int rc = pvdb_sample_forward(tree, plane, channels, d_xyz, d_xyz + n, d_xyz + 2 * n, n, d_out, nullptr, nullptr, stream);
    if (rc) return rc;
    PVDB_CUDA(cudaMemcpyAsync(out, d_out, pb * channels, cudaMemcpyDeviceToHost, st));
    PVDB_CUDA(cudaStreamSynchronize(st));
    return PVDB_OK;
}

extern "C" int pvdb_sample_backward_host(const pvdb_tree* tree, float* grad_plane, int channels, const float* xs,
                                         const float* ys, const float* zs, const float* grad_out, int64_t n,
                                         void* stream) {
    PVDB_CHECK_ARG(tree && grad_plane, "null pointer");
    if (n <= 0) return PVDB_OK;
    PVDB_CHECK_ARG(xs && ys && zs && grad_out, "null host buffers");
    cudaStream_t st = (cudaStream_t)stream;
    PvdbStage stage(st);
    const size_t pb = (size_t)n * sizeof(float);
    float* d_xyz = (float*)stage.alloc(3 * pb);
    float* d_g = (float*)stage.alloc(pb * channels);
    if (!d_xyz || !d_g) { pvdb_set_error("%s: cudaMallocAsync failed", __func__); return PVDB_ERR_CUDA; }
    PVDB_CUDA(cudaMemcpyAsync(d_xyz, xs, pb, cudaMemcpyHostToDevice, st));
    PVDB_CUDA(cudaMemcpyAsync(d_xyz + n, ys, pb, cudaMemcpyHostToDevice, st));
    PVDB_CUDA(cudaMemcpyAsync(d_xyz + 2 * n, zs, pb, cudaMemcpyHostToDevice, st));
    PVDB_CUDA(cudaMemcpyAsync(d_g, grad_out, pb * channels, cudaMemcpyHostToDevice, st));
    int rc = pvdb_sample_backward(tree, grad_plane, channels, d_xyz, d_xyz + n, d_xyz + 2 * n, d_g, n, stream);
    if (rc) return rc;
    PVDB_CUDA(cudaStreamSynchronize(st));
    return PVDB_OK;
}

extern "C" float pvdb_adam_stepsize(float lr, float beta0, float beta1, int step) {
    // plenvdb.h:753 — std::pow(float,float), std::sqrt(float), all in float
    return lr * std::sqrt(1 - std::pow(beta1, (float)step)) / (1 - std::pow(beta0, (float)step));
}

extern "C" int pvdb_adam_step(const pvdb_tree* tree, float* param, const float* grad, float* exp_avg, float* exp_avg_sq,
                              int channels, int mode, float stepsz, float eps, float beta0, float beta1,
                              const float* perlr, void* stream) {
    PVDB_CHECK_ARG(tree && param && grad && exp_avg && exp_avg_sq, "null pointer");
    PVDB_CHECK_ARG(mode >= 0 && mode <= 2, "mode must be 0, 1 or 2");
    PVDB_CHECK_ARG(mode != 2 || perlr, "mode 2 needs a per-voxel lr plane");
    PVDB_CHECK_ARG(channels == 1 || channels % 3 == 0, "channels must be 1 or a multiple of 3");
    if (tree->n_leaf == 0) return PVDB_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (channels == 1) {
        const int64_t total = (int64_t)tree->n_leaf * 512;
        k_adam<1><<<pvdb_grid_for(total, 256), 256, 0, st>>>(*tree, param, grad, exp_avg, exp_avg_sq, 1, mode, stepsz, eps,
                                                             beta0, beta1, perlr);
    } else {
        const int64_t total = (int64_t)tree->n_leaf * 512 * (channels / 3);
        k_adam<3><<<pvdb_grid_for(total, 256), 256, 0, st>>>(*tree, param, grad, exp_avg, exp_avg_sq, channels, mode, stepsz,
                                                             eps, beta0, beta1, perlr);
    }
    PVDB_LAUNCH_CHECK();
    return PVDB_OK;
}

extern "C" int pvdb_zero_grad(const pvdb_tree* tree, float* grad, int channels, void* stream) {
    PVDB_CHECK_ARG(tree && grad && channels > 0, "bad arguments");
    if (tree->n_leaf == 0) return PVDB_OK;
    const int64_t total = (int64_t)tree->n_leaf * 512 * channels;
    k_zero_grad<<<pvdb_grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(*tree, grad, channels);
    PVDB_LAUNCH_CHECK();
    return PVDB_OK;
}

extern "C" int pvdb_copy_from_dense(const pvdb_tree* tree, float* plane, int channels, const float* dense, int rx,
                                    int ry, int rz, void* stream) {
    PVDB_CHECK_ARG(tree && plane && dense && channels > 0, "bad arguments");
    if (tree->n_leaf == 0) return PVDB_OK;
    const int64_t total = (int64_t)tree->n_leaf * 512 * channels;
    k_copy_from_dense<<<pvdb_grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(*tree, plane, channels, dense, rx, ry, rz);
    PVDB_LAUNCH_CHECK();
    return PVDB_OK;
}

extern "C" int pvdb_copy_to_dense(const pvdb_tree* tree, const float* plane, int channels, float* dense, int rx, int ry,
                                  int rz, void* stream) {
    PVDB_CHECK_ARG(tree && plane && dense && channels > 0, "bad arguments");
    const int64_t total = (int64_t)rx * ry * rz * channels;
    if (total == 0) return PVDB_OK;
    k_copy_to_dense<<<pvdb_grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(*tree, plane, channels, dense, rx, ry, rz);
    PVDB_LAUNCH_CHECK();
    return PVDB_OK;
}

extern "C" int pvdb_set_values_on_by_mask(const pvdb_tree* tree, float* plane, const uint8_t* mask, float val, int rx,
                                          int ry, int rz, void* stream) {
    PVDB_CHECK_ARG(tree && plane && mask, "bad arguments");
    if (tree->n_leaf == 0) return PVDB_OK;
    const int64_t total = (int64_t)tree->n_leaf * 512;
    k_set_on_by_mask<<<pvdb_grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(*tree, plane, mask, val, rx, ry, rz);
    PVDB_LAUNCH_CHECK();
    return PVDB_OK;
}
