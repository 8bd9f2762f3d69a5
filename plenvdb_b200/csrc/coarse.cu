// coarse.cu — the colour path of the COARSE stage (configs/default.py:41-66, 94: rgbnet_dim = 0): k0 has 3 channels, there is no
// rgbnet and rgb = sigmoid(k0) (dvgo.py:72-79, 344-346).  Everything else of the iteration — march, lists, compositing, losses,
// the density branch of the backward — is the fused step's (train_step.cu); the update runs the full-grid Adam of the stepmodes
// the coarse stage uses (0 for k0, 2 = per-voxel lr for the density: masked_adam.py:56-68).
#include "common.cuh"
#include "rgbnet.cuh"

namespace {

constexpr int CNT_M_KEEP = 1, CNT_N_TOUCHED_K0 = 4;

// thread per kept sample: 3-channel trilinear sample in the reference's corner order and arithmetic (colorvdb.cu:81-111 with one
// Vec3f grid), through the record ids the march saved; k_feat keeps its 12-float row stride (channels 3..11 zero).
__global__ void __launch_bounds__(256) k_direct_fwd(const float* __restrict__ k0, const float* __restrict__ k_xyz, const int32_t* __restrict__ k_corner,
                                                    const int32_t* __restrict__ counters, int64_t cap_keep, float* __restrict__ k_feat,
                                                    float* __restrict__ k_rgb) {
    pvdb_pdl_wait();
    const int64_t M = min((int64_t)counters[CNT_M_KEEP], cap_keep);
    for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < M; s += (int64_t)gridDim.x * blockDim.x) {
        PvdbTri tri;
        tri.set(k_xyz[s * 3], k_xyz[s * 3 + 1], k_xyz[s * 3 + 2]);
        const int4 ca = __ldg(reinterpret_cast<const int4*>(k_corner + s * 8)), cb = __ldg(reinterpret_cast<const int4*>(k_corner + s * 8) + 1);
        const int rec[8] = {ca.x, ca.y, ca.z, ca.w, cb.x, cb.y, cb.z, cb.w};
        float v[8][3];
#pragma unroll
        for (int q = 0; q < 8; ++q)
#pragma unroll
            for (int c = 0; c < 3; ++c) v[q][c] = rec[q] >= 0 ? __ldg(k0 + (size_t)rec[q] * 3 + c) : 0.f;
        float x[3] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int dx = PVDB_CORNER[q][0], dy = PVDB_CORNER[q][1], dz = PVDB_CORNER[q][2];
            const float sc = __fmul_rn(__fmul_rn(tri.f(0, dx), tri.f(1, dy)), tri.f(2, dz));
#pragma unroll
            for (int c = 0; c < 3; ++c) x[c] = __fmaf_rn(sc, v[q][c], x[c]);   // a missing corner contributes fma(sc, 0, x) = x
        }
        float4* kf = reinterpret_cast<float4*>(k_feat + s * 12);
        kf[0] = make_float4(x[0], x[1], x[2], 0.f); kf[1] = make_float4(0.f, 0.f, 0.f, 0.f); kf[2] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int c = 0; c < 3; ++c) k_rgb[s * 3 + c] = 1.0f / (1.0f + expf(-x[c]));   // torch.sigmoid (dvgo.py:346)
    }
}

// thread per kept sample: dL/dlogit (left in k_rgb by the composite kernel) is dL/d(k0 sample); scatter (colorvdb.cu:130-160)
__global__ void __launch_bounds__(256) k_direct_bwd(float* __restrict__ k0_grad, const float* __restrict__ k_xyz, const int32_t* __restrict__ k_corner,
                                                    const float* __restrict__ k_glogit, const int32_t* __restrict__ counters, int64_t cap_keep,
                                                    int32_t* __restrict__ k0_touched, int32_t* __restrict__ k0_touched_list, int32_t* __restrict__ counters_w) {
    pvdb_pdl_wait();
    const int64_t M = min((int64_t)counters[CNT_M_KEEP], cap_keep);
    for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < M; s += (int64_t)gridDim.x * blockDim.x) {
        PvdbTri tri;
        tri.set(k_xyz[s * 3], k_xyz[s * 3 + 1], k_xyz[s * 3 + 2]);
        const int4 ca = __ldg(reinterpret_cast<const int4*>(k_corner + s * 8)), cb = __ldg(reinterpret_cast<const int4*>(k_corner + s * 8) + 1);
        const int rec[8] = {ca.x, ca.y, ca.z, ca.w, cb.x, cb.y, cb.z, cb.w};
        const float g[3] = {k_glogit[s * 3], k_glogit[s * 3 + 1], k_glogit[s * 3 + 2]};
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            if (rec[q] < 0) continue;
            const int dx = PVDB_CORNER[q][0], dy = PVDB_CORNER[q][1], dz = PVDB_CORNER[q][2];
            const float sc = __fmul_rn(__fmul_rn(tri.f(0, dx), tri.f(1, dy)), tri.f(2, dz));
#pragma unroll
            for (int c = 0; c < 3; ++c) atomicAdd(k0_grad + (size_t)rec[q] * 3 + c, __fmul_rn(g[c], sc));
            pvdb_touch_leaf(k0_touched, k0_touched_list, counters_w + CNT_N_TOUCHED_K0, rec[q] >> 9);
        }
    }
}

}  // namespace

int pvdb_direct_forward(const pvdb_train_bufs* b, cudaStream_t st) {
    PVDB_CHECK_ARG(b->k_corner, "k_corner scratch missing");
    PVDB_CUDA(pvdb_launch_pdl(k_direct_fwd, dim3(PVDB_SMS * 8), dim3(256), 0, st, (const float*)b->k0, (const float*)b->k_xyz, (const int32_t*)b->k_corner,
                              (const int32_t*)b->counters, b->cap_keep, b->k_feat, b->k_rgb));
    PVDB_LAUNCH_CHECK();
    return PVDB_OK;
}
int pvdb_direct_backward(const pvdb_train_bufs* b, cudaStream_t st) {
    PVDB_CUDA(pvdb_launch_pdl(k_direct_bwd, dim3(PVDB_SMS * 8), dim3(256), 0, st, b->k0_grad, (const float*)b->k_xyz, (const int32_t*)b->k_corner,
                              (const float*)b->k_rgb, (const int32_t*)b->counters, b->cap_keep, b->k0_touched, b->k0_touched_list, b->counters));
    PVDB_LAUNCH_CHECK();
    return PVDB_OK;
}
