// Does tcgen05.mma kind::tf32 truncate or round the low 13 mantissa bits of its fp32-container operands?
#include <cstdio>
#include <cuda_runtime.h>
#include "../plenvdb_b200/csrc/tc_ptx.cuh"
__global__ void probe(const float* xs, int n, float* out) {
    __shared__ __align__(128) unsigned char sm[128 * 8 * 4 + 16 * 8 * 4 + 64];
    float* A = reinterpret_cast<float*>(sm);
    float* B = reinterpret_cast<float*>(sm + 4096);
    const uint32_t sbase = smem_u32(sm);
    const uint32_t bar = sbase + 4096 + 512;
    uint32_t* slot = reinterpret_cast<uint32_t*>(sm + 4096 + 512 + 8);
    const int tid = threadIdx.x;
    for (int e = tid; e < 128 * 8 + 16 * 8; e += 128) A[e] = 0.f;
    __syncthreads();
    // row r of A holds xs[r] at k = 0 ; B row 0 has 1.0 at k = 0 -> D[r][0] = tf32(xs[r]) * 1
    if (tid < n) *reinterpret_cast<float*>(sm + canon_off(tid, 0, 8)) = xs[tid];
    if (tid == 0) { *reinterpret_cast<float*>(sm + 4096 + canon_off(0, 0, 8)) = 1.0f; mbar_init(bar, 1); }
    if (tid < 32) tmem_alloc(sbase + 4096 + 512 + 8, 32);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *slot;
    if (tid == 0) {
        umma_tf32_ss(tmem, make_desc(sbase, 8), make_desc(sbase + 4096, 8), make_idesc(16), 0);
        umma_commit(bar);
    }
    mbar_wait(bar, 0);
    tc_fence_after();
    uint32_t r[8];
    tmem_ld8(tmem + ((uint32_t)((tid >> 5) * 32) << 16), r);
    tmem_ld_wait();
    out[tid] = __uint_as_float(r[0]);
    tc_fence_before();
    __syncthreads();
    if (tid < 32) tmem_dealloc(tmem, 32);
}
int main() {
    float h[8] = {1.0f + 1.0f / 2048 + 1.0f / 4096, 1.0f + 1.0f / 2048, 1.0f + 1.0f / 4096, 1.0f + 1.0f / 1024 + 1.0f / 2048,
                  -(1.0f + 1.0f / 2048 + 1.0f / 4096), 3.14159265f, 1.9999999f, 1.0f + 1.0f / 1024};
    float *dx, *dout, o[128];
    cudaMalloc(&dx, sizeof(h)); cudaMalloc(&dout, 512);
    cudaMemcpy(dx, h, sizeof(h), cudaMemcpyHostToDevice);
    probe<<<1, 128>>>(dx, 8, dout);
    cudaError_t e = cudaDeviceSynchronize();
    printf("status %s\n", cudaGetErrorString(e));
    cudaMemcpy(o, dout, 512, cudaMemcpyDeviceToHost);
    for (int i = 0; i < 8; ++i) {
        unsigned b; memcpy(&b, &h[i], 4); b &= 0xffffe000u; float t; memcpy(&t, &b, 4);
        printf("x=%.9g  mma=%.9g  trunc=%.9g  %s\n", h[i], o[i], t, o[i] == t ? "TRUNC" : "other");
    }
    return 0;
}
