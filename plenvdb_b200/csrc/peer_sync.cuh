// peer_sync.cuh — signalling primitives for kernels that talk to other GPUs through NVLink peer memory (CUDA IPC mappings):
// system-scope release stores / acquire loads on monotone epoch words, with a bounded spin so that a dead peer sets an error
// word instead of hanging the GPU.  Shared by dp_exchange.cu (gradient exchange) and renderer.cu (frame assembly).
#pragma once
#include <stdint.h>

namespace {

constexpr unsigned long long PVDB_SPIN_TIMEOUT_NS = 2000000000ull;   // 2 s

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long globaltimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// Spin until the epoch word *p has reached `epoch` (wrap-safe); on timeout set *err = code and give up.
__device__ __forceinline__ void wait_epoch(const uint32_t* p, uint32_t epoch, int32_t* err, int code) {
    const unsigned long long t0 = globaltimer();
    while ((int32_t)(ld_acquire_sys(p) - epoch) < 0) {
        if (globaltimer() - t0 > PVDB_SPIN_TIMEOUT_NS) { atomicExch(err, code); break; }
        __nanosleep(64);
    }
}

}  // namespace
