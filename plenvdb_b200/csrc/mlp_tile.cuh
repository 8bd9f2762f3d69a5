// mlp_tile.cuh — 64-sample x 128-wide fp32 tile GEMM shared by the training rgbnet (rgbnet.cu) and the merged
// renderer's MLP (renderer.cu).  256 threads as a 16x16 grid, 4x8 register micro-tile per thread; activations
// row-major in shared memory, weights k-major.
#pragma once
#include <cuda_runtime.h>
#include "rgbnet.cuh"

namespace {

constexpr int TS = 64;          // samples per tile
constexpr int NT = 256;         // threads per CTA
constexpr int W = PVDB_NET_W;   // 128
constexpr int DIN = PVDB_NET_DIN;
constexpr int KX = 40;          // DIN padded to a multiple of 4
constexpr int LDX = 44;         // smem leading dims: multiples of 4 with (4*ld) % 32 == 16
constexpr int LDH = 132;

__device__ __forceinline__ void red_add(float* addr, float v) { asm volatile("red.global.add.f32 [%0], %1;" ::"l"(addr), "f"(v) : "memory"); }
__device__ __forceinline__ void red_add4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// acc[4][8] += A[rows ty*4..+3][0..K) * B[0..K)[cols tx*4..+3 and 64+tx*4..+3]; A row-major (lda), B k-major (ldb)
template <int K, int LDA, int LDB>
__device__ __forceinline__ void gemm_4x8(const float* __restrict__ sA, const float* __restrict__ sB, int ty, int tx, float (&acc)[4][8]) {
#pragma unroll 2
    for (int k = 0; k < K; k += 4) {
        float4 a[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = *reinterpret_cast<const float4*>(sA + (ty * 4 + i) * LDA + k);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            const float4 b0 = *reinterpret_cast<const float4*>(sB + (k + kk) * LDB + tx * 4);
            const float4 b1 = *reinterpret_cast<const float4*>(sB + (k + kk) * LDB + 64 + tx * 4);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float av = kk == 0 ? a[i].x : kk == 1 ? a[i].y : kk == 2 ? a[i].z : a[i].w;
                acc[i][0] = fmaf(av, b0.x, acc[i][0]); acc[i][1] = fmaf(av, b0.y, acc[i][1]);
                acc[i][2] = fmaf(av, b0.z, acc[i][2]); acc[i][3] = fmaf(av, b0.w, acc[i][3]);
                acc[i][4] = fmaf(av, b1.x, acc[i][4]); acc[i][5] = fmaf(av, b1.y, acc[i][5]);
                acc[i][6] = fmaf(av, b1.z, acc[i][6]); acc[i][7] = fmaf(av, b1.w, acc[i][7]);
            }
        }
    }
}


}  // namespace
