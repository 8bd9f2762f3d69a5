// profile.cu — per-kernel CUDA-event timing of the fused calls (used by bench.py for the roofline line).
#include "common.cuh"
#include <string.h>

namespace {
constexpr int MAXEV = 48;
thread_local bool g_on = false;
thread_local bool g_created = false;
thread_local cudaEvent_t g_ev[MAXEV];
thread_local char g_name[MAXEV][32];
thread_local int g_n = 0;
}  // namespace

void pvdb_prof_begin(cudaStream_t st) {
    if (!g_on) return;
    if (!g_created) {
        for (int i = 0; i < MAXEV; ++i) cudaEventCreate(&g_ev[i]);
        g_created = true;
    }
    g_n = 0;
    cudaEventRecord(g_ev[0], st);
    strncpy(g_name[0], "begin", 31);
    g_n = 1;
}
void pvdb_prof_mark(const char* name, cudaStream_t st) {
    if (!g_on || g_n == 0 || g_n >= MAXEV) return;
    cudaEventRecord(g_ev[g_n], st);
    strncpy(g_name[g_n], name, 31);
    g_name[g_n][31] = 0;
    ++g_n;
}

bool pvdb_prof_active() { return g_on; }

extern "C" int pvdb_profile_enable(int on) {
    g_on = on != 0;
    if (!g_on) g_n = 0;
    return PVDB_OK;
}
// Fills ms[i] / names[i*32] with the duration of segment i (time between mark i and mark i+1) of the last fused call
// on this thread; returns the number of segments (synchronises on the last event).
extern "C" int pvdb_profile_fetch(int max_segments, float* ms, char* names) {
    if (g_n < 2) return 0;
    cudaEventSynchronize(g_ev[g_n - 1]);
    int n = 0;
    for (int i = 1; i < g_n && n < max_segments; ++i, ++n) {
        float t = 0.f;
        cudaEventElapsedTime(&t, g_ev[i - 1], g_ev[i]);
        ms[n] = t;
        memcpy(names + n * 32, g_name[i], 32);
    }
    return n;
}
