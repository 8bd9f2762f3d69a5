// render_ray.cuh — per-pixel ray construction of the merged renderer, shared by renderer.cu and the tcgen05 MLP.
#pragma once
#include "common.cuh"

namespace {

struct RenderConst {
    pvdb_tree tree;
    const int32_t* idx_plane;
    const float* dendata;
    const float* coldata;
    float K[9];
    float xyz_min[3], ext[3];     // ext = xyz_max - xyz_min (float sub)
    float wld[3];                 // reso - 1
    float near, stepdist, act_shift, interval, thres, bg;
    int inverse_y, H, W;
    // Row layout of one call: local row lr of the band buffer is image row
    //   row_begin + lr                                                      (band_stride == 0: one contiguous band)
    //   row_begin + (lr / band_rows) * band_stride + lr % band_rows         (interleaved: every band_stride-th group of band_rows rows)
    int band_rows, band_stride;
    // Empty-space skipping (optional, NULL = every step is marched): one bit per 8^3 block, set when the block or one of its
    // +1 neighbours along any axes holds a leaf of the index tree (pvdb_render_block_bits); nb = blocks per axis.
    const uint32_t* skip_bits;
    int nb[3];
};

// image pixel index (row * W + col) of local pixel `local` of the band buffer
__device__ __forceinline__ int render_gpix(const RenderConst& C, int row_begin, int local) {
    if (C.band_stride == 0) return row_begin * C.W + local;
    const int lr = local / C.W, col = local - lr * C.W;
    const int k = lr / C.band_rows;
    return (row_begin + k * C.band_stride + (lr - k * C.band_rows)) * C.W + col;
}

struct Ray {
    float ro[3], rd[3], vd[3];
    float steplen, tmin, tmax;
};

// get_rays (:122-167) + get_tminmax (:50-64), instruction order read from the reference PTX.
__device__ __forceinline__ void ray_setup(const RenderConst& C, const float* __restrict__ c2w, int n, Ray& R) {
    const float pixeli = (float)((double)(n % C.W) + 0.5), pixelj = (float)((double)(n / C.W) + 0.5);
    float dir[3];
    dir[0] = __fdiv_rn(__fsub_rn(pixeli, C.K[2]), C.K[0]);
    if (C.inverse_y) { dir[1] = __fdiv_rn(__fsub_rn(pixelj, C.K[5]), C.K[4]); dir[2] = 1.f; }
    else { dir[1] = __fdiv_rn(-__fsub_rn(pixelj, C.K[5]), C.K[4]); dir[2] = -1.f; }
    float rdw[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        // d0*c0 + d1*c1 + d2*c2  ->  fma(d2,c2, fma(d0,c0, d1*c1))
        rdw[a] = __fmaf_rn(dir[2], __ldg(c2w + a * 4 + 2), __fmaf_rn(dir[0], __ldg(c2w + a * 4), __fmul_rn(dir[1], __ldg(c2w + a * 4 + 1))));
        R.ro[a] = __fdiv_rn(__fsub_rn(__ldg(c2w + a * 4 + 3), C.xyz_min[a]), C.ext[a]);
    }
    const float len = __fsqrt_rn(__fmaf_rn(rdw[2], rdw[2], __fmaf_rn(rdw[0], rdw[0], __fmul_rn(rdw[1], rdw[1]))));
    R.steplen = __fdiv_rn(C.stepdist, len);
    const float inv = __frcp_rn(len);   // Vec3::normalize: *this *= 1/length
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        R.rd[a] = __fdiv_rn(rdw[a], C.ext[a]);
        R.vd[a] = __fmul_rn(rdw[a], inv);
    }
    const float far = 1e9f;   // setKwargs overrides far (plenvdb.h:1008)
    const float vx = R.rd[0] == 0.f ? 1e-6f : R.rd[0], vy = R.rd[1] == 0.f ? 1e-6f : R.rd[1], vz = R.rd[2] == 0.f ? 1e-6f : R.rd[2];
    const float ax = __fdiv_rn(__fsub_rn(1.f, R.ro[0]), vx), ay = __fdiv_rn(__fsub_rn(1.f, R.ro[1]), vy), az = __fdiv_rn(__fsub_rn(1.f, R.ro[2]), vz);
    const float bx = __fdiv_rn(-R.ro[0], vx), by = __fdiv_rn(-R.ro[1], vy), bz = __fdiv_rn(-R.ro[2], vz);
    R.tmin = fmaxf(fminf(fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fminf(az, bz)), far), C.near);
    R.tmax = fmaxf(fminf(fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fmaxf(az, bz)), far), C.near);
}


// Arguments of the renderer's MLP stage (fp32 tile kernel in renderer.cu, tcgen05 kernel in rgbnet_tc.cu).
struct RenderMlpArgs {
    RenderConst C;
    const float* c2w;
    const float *w0, *b0, *w1, *b1, *w2, *b2;
    const int32_t* s_ray; const float* s_weight; const float* s_feat; float* s_rgb;
    const int32_t* counters; int64_t cap; int row_begin;
    unsigned char* img;   // tcgen05 weight image scratch (pvdb_render_bufs.w_img)
};

}  // namespace
