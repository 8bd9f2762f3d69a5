// render_utils.cu — the `render_utils_cuda` / `adam_upd_cuda` operator set of the reference
// (plenvdb/lib/cuda/render_utils_kernel.cu, adam_upd_kernel.cu) on raw device pointers.
// Each expression is pinned with rounding intrinsics to the instruction sequence nvcc emits for the
// reference source (FMA contraction included), because the integer outputs (sample counts, bbox and
// occupancy masks, segment offsets) are functions of these roundings.
#include "common.cuh"
#include "ray_math.cuh"

// S1a — infer_t_minmax (:12-35)
__global__ void __launch_bounds__(256) k_infer_t_minmax(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                                                        const float* __restrict__ xyz_min, const float* __restrict__ xyz_max,
                                                        float near, float far, int n_rays, float* __restrict__ t_min,
                                                        float* __restrict__ t_max) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rays) return;
    float tmin, tmax;
    pvdb_ray_t_minmax(rays_o + r * 3, rays_d + r * 3, xyz_min, xyz_max, near, far, tmin, tmax);
    t_min[r] = tmin;
    t_max[r] = tmax;
}

// S1b — infer_n_samples (:38-55)
__global__ void __launch_bounds__(256) k_infer_n_samples(const float* __restrict__ rays_d, const float* __restrict__ t_min,
                                                         const float* __restrict__ t_max, float stepdist, int n_rays,
                                                         int64_t* __restrict__ n_samples) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rays) return;
    n_samples[r] = pvdb_ray_n_samples(rays_d + r * 3, t_min[r], t_max[r], stepdist);
}

// S1c — infer_ray_start_dir (:58-79)
__global__ void __launch_bounds__(256) k_infer_ray_start_dir(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                                                             const float* __restrict__ t_min, int n_rays,
                                                             float* __restrict__ rays_start, float* __restrict__ rays_dir) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rays) return;
    float s[3], d[3];
    pvdb_ray_start_dir(rays_o + r * 3, rays_d + r * 3, t_min[r], s, d);
#pragma unroll
    for (int a = 0; a < 3; ++a) { rays_start[r * 3 + a] = s[a]; rays_dir[r * 3 + a] = d[a]; }
}

// Fused per-ray setup for the count phase of sample_pts_on_rays (:196-242): t range, N_steps, start, dir.
__global__ void __launch_bounds__(256) k_sample_setup(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                                                      const float* __restrict__ xyz_min, const float* __restrict__ xyz_max,
                                                      float near, float far, float stepdist, int n_rays,
                                                      float* __restrict__ t_min, float* __restrict__ t_max,
                                                      int64_t* __restrict__ n_steps, float* __restrict__ rays_start,
                                                      float* __restrict__ rays_dir) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rays) return;
    float tmin, tmax, s[3], d[3];
    pvdb_ray_t_minmax(rays_o + r * 3, rays_d + r * 3, xyz_min, xyz_max, near, far, tmin, tmax);
    t_min[r] = tmin;
    t_max[r] = tmax;
    n_steps[r] = pvdb_ray_n_samples(rays_d + r * 3, tmin, tmax, stepdist);
    pvdb_ray_start_dir(rays_o + r * 3, rays_d + r * 3, tmin, s, d);
#pragma unroll
    for (int a = 0; a < 3; ++a) { rays_start[r * 3 + a] = s[a]; rays_dir[r * 3 + a] = d[a]; }
}

// Inclusive scan of int64 counts (N_steps.cumsum(0), :211).  Single CTA, 1024 threads, chunked: the input
// is one value per ray (8192 .. 65536), far below one SM's reach.
__global__ void __launch_bounds__(1024) k_cumsum_i64(const int64_t* __restrict__ in, int64_t* __restrict__ out, int n) {
    __shared__ int64_t warp_sums[32];
    __shared__ int64_t carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int base = 0; base < n; base += 1024) {
        const int i = base + threadIdx.x;
        int64_t v = i < n ? in[i] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int64_t u = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= o) v += u;
        }
        if (lane == 31) warp_sums[wid] = v;
        __syncthreads();
        if (wid == 0) {
            int64_t w = warp_sums[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int64_t u = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += u;
            }
            warp_sums[lane] = w;
        }
        __syncthreads();
        const int64_t carry = carry_s;
        const int64_t total = v + (wid ? warp_sums[wid - 1] : 0) + carry;
        if (i < n) out[i] = total;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = total;
        __syncthreads();
    }
}

// S1d — ray_id / step_id / rays_pts / mask_outbbox (:144-194).  One warp per ray: lanes stride over the
// ray's own segment, so ids are produced directly instead of by marker + cumsum.
__global__ void __launch_bounds__(256) k_sample_fill(const float* __restrict__ rays_start, const float* __restrict__ rays_dir,
                                                     const float* __restrict__ xyz_min, const float* __restrict__ xyz_max,
                                                     const int64_t* __restrict__ cumsum, float stepdist, int n_rays,
                                                     float* __restrict__ rays_pts, uint8_t* __restrict__ mask_outbbox,
                                                     int64_t* __restrict__ ray_id, int64_t* __restrict__ step_id) {
    const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= n_rays) return;
    const int64_t beg = r ? cumsum[r - 1] : 0, end = cumsum[r];
    const float sx = rays_start[r * 3], sy = rays_start[r * 3 + 1], sz = rays_start[r * 3 + 2];
    const float dx = rays_dir[r * 3], dy = rays_dir[r * 3 + 1], dz = rays_dir[r * 3 + 2];
    const float mnx = xyz_min[0], mny = xyz_min[1], mnz = xyz_min[2];
    const float mxx = xyz_max[0], mxy = xyz_max[1], mxz = xyz_max[2];
    for (int64_t i = beg + lane; i < end; i += 32) {
        const int step = (int)(i - beg);
        float px, py, pz;
        pvdb_ray_point(sx, sy, sz, dx, dy, dz, stepdist, step, px, py, pz);
        rays_pts[i * 3] = px;
        rays_pts[i * 3 + 1] = py;
        rays_pts[i * 3 + 2] = pz;
        mask_outbbox[i] = (mnx > px) | (mny > py) | (mnz > pz) | (mxx < px) | (mxy < py) | (mxz < pz);
        ray_id[i] = r;
        step_id[i] = step;
    }
}

// S2 — maskcache_lookup (:374-392)
__global__ void __launch_bounds__(256) k_maskcache_lookup(const uint8_t* __restrict__ world, const float* __restrict__ xyz,
                                                          uint8_t* __restrict__ out, const float* __restrict__ scale,
                                                          const float* __restrict__ shift, int sz_i, int sz_j, int sz_k,
                                                          int64_t n_pts) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_pts) return;
    const int i = pvdb_mask_ijk(xyz[p * 3], scale[0], shift[0]);
    const int j = pvdb_mask_ijk(xyz[p * 3 + 1], scale[1], shift[1]);
    const int k = pvdb_mask_ijk(xyz[p * 3 + 2], scale[2], shift[2]);
    if (0 <= i && i < sz_i && 0 <= j && j < sz_j && 0 <= k && k < sz_k)
        out[p] = world[((int64_t)i * sz_j + j) * sz_k + k];
    else
        out[p] = 0;   // the reference writes into a zero-initialised tensor (:405)
}

// A1 — raw2alpha (:431-443) and backward (:507-517)
__global__ void __launch_bounds__(256) k_raw2alpha(const float* __restrict__ density, float shift, float interval,
                                                   int64_t n, float* __restrict__ exp_d, float* __restrict__ alpha) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    float e;
    alpha[p] = pvdb_raw2alpha(density[p], shift, interval, e);
    exp_d[p] = e;
}
__global__ void __launch_bounds__(256) k_raw2alpha_backward(const float* __restrict__ exp_d, const float* __restrict__ gback,
                                                            float interval, int64_t n, float* __restrict__ grad) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    grad[p] = pvdb_raw2alpha_bwd(exp_d[p], gback[p], interval);
}

// A2 — segment bounds (:607-617 + host index write :635) and the per-ray sequential cumprod (:577-605)
__global__ void __launch_bounds__(256) k_segment_bounds(const int64_t* __restrict__ ray_id, int64_t n_pts,
                                                        int64_t* __restrict__ i_start, int64_t* __restrict__ i_end) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_pts) return;
    if (idx > 0 && ray_id[idx] != ray_id[idx - 1]) {
        i_start[ray_id[idx]] = idx;
        i_end[ray_id[idx - 1]] = idx;
    }
    if (idx == n_pts - 1) i_end[ray_id[idx]] = n_pts;
}
__global__ void __launch_bounds__(128) k_alpha2weight(const float* __restrict__ alpha, int n_rays, float* __restrict__ weight,
                                                      float* __restrict__ T, float* __restrict__ alphainv_last,
                                                      const int64_t* __restrict__ i_start, int64_t* __restrict__ i_end) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rays) return;
    const int64_t i_s = i_start[r], i_e_max = i_end[r];
    float T_cum = 1.f;
    int64_t i;
    for (i = i_s; i < i_e_max; ++i) {
        const float a = alpha[i];
        T[i] = T_cum;
        weight[i] = __fmul_rn(T_cum, a);
        T_cum = pvdb_T_update(T_cum, a);
        if ((double)T_cum < 1e-3) { i += 1; break; }
    }
    i_end[r] = i;
    alphainv_last[r] = T_cum;
}
__global__ void __launch_bounds__(128) k_alpha2weight_backward(const float* __restrict__ alpha, const float* __restrict__ weight,
                                                               const float* __restrict__ T, const float* __restrict__ alphainv_last,
                                                               const int64_t* __restrict__ i_start, const int64_t* __restrict__ i_end,
                                                               int n_rays, const float* __restrict__ gw,
                                                               const float* __restrict__ glast, float* __restrict__ grad) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rays) return;
    const int64_t i_s = i_start[r], i_e = i_end[r];
    float back_cum = __fmul_rn(glast[r], alphainv_last[r]);
    for (int64_t i = i_e - 1; i >= i_s; --i) {
        grad[i] = pvdb_a2w_grad(gw[i], T[i], back_cum, alpha[i]);
        back_cum = __fmaf_rn(gw[i], weight[i], back_cum);
    }
}

// adam_upd_cuda family (adam_upd_kernel.cu:9-58; step_size :72 computed in double then narrowed)
__global__ void __launch_bounds__(256) k_dense_adam(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                    float* __restrict__ v, const float* __restrict__ perlr, int64_t n, int mode,
                                                    float step_size, float beta1, float beta2, float eps) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float gi = g[i];
    if (mode == 1 && gi == 0.f) return;
    pvdb_dense_adam_update(p[i], m[i], v[i], gi, mode == 2 ? perlr[i] : 1.f, mode == 2, step_size, beta1, beta2, eps);
}

// ---------------------------------------------------------------------------------------------
// C-ABI
// ---------------------------------------------------------------------------------------------
extern "C" int pvdb_infer_t_minmax(const float* rays_o, const float* rays_d, const float* xyz_min, const float* xyz_max,
                                   float near, float far, int n_rays, float* t_min, float* t_max, void* stream) {
    if (n_rays <= 0) return PVDB_OK;
    PVDB_CHECK_ARG(rays_o && rays_d && xyz_min && xyz_max && t_min && t_max, "null pointer");
    k_infer_t_minmax<<<pvdb_grid_for(n_rays, 256), 256, 0, (cudaStream_t)stream>>>(rays_o, rays_d, xyz_min, xyz_max, near, far,
                                                                                   n_rays, t_min, t_max);
    PVDB_LAUNCH_CHECK();
    return PVDB_OK;
}
extern "C" int pvdb_infer_n_samples(const float* rays_d, const float* t_min, const float* t_max, float stepdist, int n_rays,
                                    int64_t* n_samples, void* stream) {
    if (n_rays <= 0) return PVDB_OK;
    PVDB_CHECK_ARG(rays_d && t_min && t_max && n_samples, "null pointer");
    k_infer_n_samples<<<pvdb_grid_for(n_rays, 256), 256, 0, (cudaStream_t)stream>>>(rays_d, t_min, t_max, stepdist, n_rays,
                                                                                    n_samples);
    PVDB_LAUNCH_CHECK();
    return PVDB_OK;
}
extern "C" int pvdb_infer_ray_start_dir(const float* rays_o, const float* rays_d, const float* t_min, int n_rays,
                                        float* rays_start, float* rays_dir, void* stream) {
    if (n_rays <= 0) return PVDB_OK;
    PVDB_CHECK_ARG(rays_o && rays_d && t_min && rays_start && rays_dir, "null pointer");
    k_infer_ray_start_dir<<<pvdb_grid_for(n_rays, 256), 256, 0, (cudaStream_t)stream>>>(rays_o, rays_d, t_min, n_rays,
                                                                                        rays_start, rays_dir);
    PVDB_LAUNCH_CHECK();
    return PVDB_OK;
}
extern "C" int pvdb_sample_pts_count(const float* rays_o, const float* rays_d, const float* xyz_min, const float* xyz_max,
                                     float near, float far, float stepdist, int n_rays, float* t_min, float* t_max,
                                     int64_t* n_steps, int64_t* n_steps_cumsum, float* rays_start, float* rays_dir,
                                     void* stream) {
    if (n_rays <= 0) return PVDB_OK;
    PVDB_CHECK_ARG(rays_o && rays_d && xyz_min && xyz_max && t_min && t_max && n_steps && n_steps_cumsum && rays_start &&
                       rays_dir, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    k_sample_setup<<<pvdb_grid_for(n_rays, 256), 256, 0, st>>>(rays_o, rays_d, xyz_min, xyz_max, near, far, stepdist, n_rays,
                                                               t_min, t_max, n_steps, rays_start, rays_dir);
    PVDB_LAUNCH_CHECK();
    k_cumsum_i64<<<1, 1024, 0, st>>>(n_steps, n_steps_cumsum, n_rays);
    PVDB_LAUNCH_CHECK();
    return PVDB_OK;
}
extern "C" int pvdb_sample_pts_fill(const float* rays_start, const float* rays_dir, const float* xyz_min, const float* xyz_max,
                                    const int64_t* n_steps_cumsum, float stepdist, int n_rays, int64_t total_len,
                                    float* rays_pts, uint8_t* mask_outbbox, int64_t* ray_id, int64_t* step_id, void* stream) {
    if (n_rays <= 0 || total_len <= 0) return PVDB_OK;
    PVDB_CHECK_ARG(rays_start && rays_dir && xyz_min && xyz_max && n_steps_cumsum && rays_pts && mask_outbbox && ray_id &&
                       step_id, "null pointer");
    k_sample_fill<<<pvdb_grid_for((int64_t)n_rays * 32, 256), 256, 0, (cudaStream_t)stream>>>(
        rays_start, rays_dir, xyz_min, xyz_max, n_steps_cumsum, stepdist, n_rays, rays_pts, mask_outbbox, ray_id, step_id);
    PVDB_LAUNCH_CHECK();
    return PVDB_OK;
}
extern "C" int pvdb_maskcache_lookup(const uint8_t* world, const float* xyz, uint8_t* out, const float* xyz2ijk_scale,
                                     const float* xyz2ijk_shift, int sz_i, int sz_j, int sz_k, int64_t n_pts, void* stream) {
    if (n_pts <= 0) return PVDB_OK;
    PVDB_CHECK_ARG(world && xyz && out && xyz2ijk_scale && xyz2ijk_shift, "null pointer");
    k_maskcache_lookup<<<pvdb_grid_for(n_pts, 256), 256, 0, (cudaStream_t)stream>>>(world, xyz, out, xyz2ijk_scale,
                                                                                    xyz2ijk_shift, sz_i, sz_j, sz_k, n_pts);
    PVDB_LAUNCH_CHECK();
    return PVDB_OK;
}
extern "C" int pvdb_raw2alpha(const float* density, float shift, float interval, int64_t n_pts, float* exp_d, float* alpha,
                              void* stream) {
    if (n_pts <= 0) return PVDB_OK;
    PVDB_CHECK_ARG(density && exp_d && alpha, "null pointer");
    k_raw2alpha<<<pvdb_grid_for(n_pts, 256), 256, 0, (cudaStream_t)stream>>>(density, shift, interval, n_pts, exp_d, alpha);
    PVDB_LAUNCH_CHECK();
    return PVDB_OK;
}
extern "C" int pvdb_raw2alpha_backward(const float* exp_d, const float* grad_back, float interval, int64_t n_pts, float* grad,
                                       void* stream) {
    if (n_pts <= 0) return PVDB_OK;
    PVDB_CHECK_ARG(exp_d && grad_back && grad, "null pointer");
    k_raw2alpha_backward<<<pvdb_grid_for(n_pts, 256), 256, 0, (cudaStream_t)stream>>>(exp_d, grad_back, interval, n_pts, grad);
    PVDB_LAUNCH_CHECK();
    return PVDB_OK;
}
extern "C" int pvdb_alpha2weight(const float* alpha, const int64_t* ray_id, int64_t n_pts, int n_rays, float* weight, float* T,
                                 float* alphainv_last, int64_t* i_start, int64_t* i_end, void* stream) {
    if (n_pts <= 0 || n_rays <= 0) return PVDB_OK;   // outputs keep the caller's 0/1 initialisation (:630-632)
    PVDB_CHECK_ARG(alpha && ray_id && weight && T && alphainv_last && i_start && i_end, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    k_segment_bounds<<<pvdb_grid_for(n_pts, 256), 256, 0, st>>>(ray_id, n_pts, i_start, i_end);
    PVDB_LAUNCH_CHECK();
    k_alpha2weight<<<pvdb_grid_for(n_rays, 128), 128, 0, st>>>(alpha, n_rays, weight, T, alphainv_last, i_start, i_end);
    PVDB_LAUNCH_CHECK();
    return PVDB_OK;
}
extern "C" int pvdb_alpha2weight_backward(const float* alpha, const float* weight, const float* T, const float* alphainv_last,
                                          const int64_t* i_start, const int64_t* i_end, int n_rays, const float* grad_weights,
                                          const float* grad_last, float* grad, void* stream) {
    if (n_rays <= 0) return PVDB_OK;
    PVDB_CHECK_ARG(alpha && weight && T && alphainv_last && i_start && i_end && grad_weights && grad_last && grad, "null pointer");
    k_alpha2weight_backward<<<pvdb_grid_for(n_rays, 128), 128, 0, (cudaStream_t)stream>>>(
        alpha, weight, T, alphainv_last, i_start, i_end, n_rays, grad_weights, grad_last, grad);
    PVDB_LAUNCH_CHECK();
    return PVDB_OK;
}
extern "C" int pvdb_dense_adam(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, const float* perlr, int64_t n,
                               int mode, int step, float beta1, float beta2, float lr, float eps, void* stream) {
    if (n <= 0) return PVDB_OK;
    PVDB_CHECK_ARG(param && grad && exp_avg && exp_avg_sq, "null pointer");
    PVDB_CHECK_ARG(mode >= 0 && mode <= 2 && (mode != 2 || perlr), "bad mode / missing per-lr");
    const float step_size = pvdb_dense_adam_stepsize(lr, beta1, beta2, step);
    k_dense_adam<<<pvdb_grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq, perlr, n, mode,
                                                                          step_size, beta1, beta2, eps);
    PVDB_LAUNCH_CHECK();
    return PVDB_OK;
}
