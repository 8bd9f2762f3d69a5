// ray_math.cuh — the per-ray / per-sample scalar expressions of the reference, pinned to the exact
// instruction sequence nvcc emits for them (read from the PTX of the reference sources compiled for
// sm_100a).  Shared by the drop-in ops (render_utils.cu) and the fused kernels so both produce the same
// integers.  Rounding intrinsics (__fmul_rn, __fmaf_rn, ...) are never re-associated or contracted.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

// infer_t_minmax (render_utils_kernel.cu:12-35): sub, IEEE div, then the min/max tree.
__device__ __forceinline__ void pvdb_ray_t_minmax(const float* __restrict__ o, const float* __restrict__ d,
                                                  const float* __restrict__ mn, const float* __restrict__ mx, float near,
                                                  float far, float& t_min, float& t_max) {
    const float vx = d[0] == 0.f ? 1e-6f : d[0];
    const float vy = d[1] == 0.f ? 1e-6f : d[1];
    const float vz = d[2] == 0.f ? 1e-6f : d[2];
    const float ax = __fdiv_rn(__fsub_rn(mx[0], o[0]), vx);
    const float ay = __fdiv_rn(__fsub_rn(mx[1], o[1]), vy);
    const float az = __fdiv_rn(__fsub_rn(mx[2], o[2]), vz);
    const float bx = __fdiv_rn(__fsub_rn(mn[0], o[0]), vx);
    const float by = __fdiv_rn(__fsub_rn(mn[1], o[1]), vy);
    const float bz = __fdiv_rn(__fsub_rn(mn[2], o[2]), vz);
    t_min = fmaxf(fminf(fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fminf(az, bz)), far), near);
    t_max = fmaxf(fminf(fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fmaxf(az, bz)), far), near);
}

// ||d||: x*x + y*y + z*z compiles to fma(z,z, fma(x,x, y*y)) (:46-49, :66-69)
__device__ __forceinline__ float pvdb_ray_norm(const float* __restrict__ d) {
    return __fsqrt_rn(__fmaf_rn(d[2], d[2], __fmaf_rn(d[0], d[0], __fmul_rn(d[1], d[1]))));
}

// infer_n_samples (:38-55): max(ceil((t_max-t_min)*rnorm/stepdist), 1.) evaluated in double after the ceil.
__device__ __forceinline__ int64_t pvdb_ray_n_samples(const float* __restrict__ d, float t_min, float t_max, float stepdist) {
    const float rnorm = pvdb_ray_norm(d);
    const float v = ceilf(__fdiv_rn(__fmul_rn(rnorm, __fsub_rn(t_max, t_min)), stepdist));
    return (int64_t)fmax((double)v, 1.0);
}

// infer_ray_start_dir (:58-79): start = fma(d, t_min, o), dir = d / ||d||
__device__ __forceinline__ void pvdb_ray_start_dir(const float* __restrict__ o, const float* __restrict__ d, float t_min,
                                                   float* __restrict__ start, float* __restrict__ dir) {
    const float rnorm = pvdb_ray_norm(d);
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        start[a] = __fmaf_rn(d[a], t_min, o[a]);
        dir[a] = __fdiv_rn(d[a], rnorm);
    }
}

// sample_pts_on_rays_cuda_kernel (:181-187): dist = stepdist*float(i_step); p = fma(dir, dist, start)
__device__ __forceinline__ void pvdb_ray_point(float sx, float sy, float sz, float dx, float dy, float dz, float stepdist,
                                               int step, float& px, float& py, float& pz) {
    const float dist = __fmul_rn(stepdist, (float)step);
    px = __fmaf_rn(dx, dist, sx);
    py = __fmaf_rn(dy, dist, sy);
    pz = __fmaf_rn(dz, dist, sz);
}

// maskcache_lookup (:385-387): round(fma(x, scale, shift)), C round = half away from zero.
__device__ __forceinline__ int pvdb_mask_ijk(float x, float scale, float shift) {
    return (int)roundf(__fmaf_rn(x, scale, shift));
}

// VDBGrid.wld2idx (plenvdb/lib/grid.py:77-78): three separate ATen kernels, sub -> div -> mul.
__device__ __forceinline__ float pvdb_wld2idx(float p, float mn, float mx, float res_minus_1) {
    return __fmul_rn(__fdiv_rn(__fsub_rn(p, mn), __fsub_rn(mx, mn)), res_minus_1);
}

// raw2alpha (:431-443): e = expf(d + shift); alpha = 1 - powf(1 + e, -interval)
__device__ __forceinline__ float pvdb_raw2alpha(float density, float shift, float interval, float& e) {
    e = expf(__fadd_rn(density, shift));
    return __fsub_rn(1.0f, powf(__fadd_rn(1.0f, e), -interval));
}

// raw2alpha_backward (:507-517): min(e, 1e10) is a double; the product chain runs in double.
__device__ __forceinline__ float pvdb_raw2alpha_bwd(float e, float gback, float interval) {
    const double m = fmin((double)e, 1e10);
    const float pw = powf(__fadd_rn(1.0f, e), __fsub_rn(-interval, 1.0f));
    return (float)(((m * (double)pw) * (double)interval) * (double)gback);
}

// alpha2weight (:596): T_cum *= (1. - alpha) in double, rounded back to float.
__device__ __forceinline__ float pvdb_T_update(float T_cum, float alpha) {
    return (float)((1.0 - (double)alpha) * (double)T_cum);
}

// alpha2weight_backward (:673): gw*T in float; 1-alpha in float; +1e-10, divide and subtract in double.
__device__ __forceinline__ float pvdb_a2w_grad(float gw, float T, float back_cum, float alpha) {
    const double num = (double)__fmul_rn(gw, T);
    const double den = (double)__fsub_rn(1.0f, alpha) + 1e-10;
    return (float)(num - (double)back_cum / den);
}

// adam_upd_kernel.cu:9-58 as compiled: m' = fma(b1, m, (1-b1)*g); v' = fma(b2, v, g*((1-b2)*g));
// p' = p - (step_size[*perlr] * m') / (eps + sqrt(v'))
__device__ __forceinline__ void pvdb_dense_adam_update(float& p, float& m, float& v, float g, float perlr, bool use_perlr,
                                                       float step_size, float beta1, float beta2, float eps) {
    const float nm = __fmaf_rn(beta1, m, __fmul_rn(__fsub_rn(1.0f, beta1), g));
    const float nv = __fmaf_rn(beta2, v, __fmul_rn(g, __fmul_rn(__fsub_rn(1.0f, beta2), g)));
    m = nm;
    v = nv;
    const float st = use_perlr ? __fmul_rn(step_size, perlr) : step_size;
    p = __fsub_rn(p, __fdiv_rn(__fmul_rn(st, nm), __fadd_rn(eps, __fsqrt_rn(nv))));
}

// adam_upd_kernel.cu:72 — host scalar in float
static inline float pvdb_dense_adam_stepsize(float lr, float beta1, float beta2, int step) {
    return lr * sqrtf(1 - powf(beta2, (float)step)) / (1 - powf(beta1, (float)step));
}
