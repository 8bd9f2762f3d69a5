// dp_exchange.cuh — what the fused data-parallel step (train_step.cu, rgbnet_tc_bwd.cu) needs from dp_exchange.cu: the address of
// the touched-flag array the emit kernel writes, the two halves of the tile exchange, and the argument blocks of the two
// kernels that carry the rgbnet-gradient exchange (the weight-gradient reduction pushes, the rgbnet Adam waits and sums).
#pragma once
#include "common.cuh"
#include "peer_sync.cuh"

constexpr int PVDB_DP_NET_PAD = (22019 + 255) & ~255;

struct PvdbDpNetPush {        // producer side: k_wgrad_reduce
    int world, rank;
    uint32_t epoch;
    float* dst[8];            // rank r's netx[parity][this rank] (peer memory)
    uint32_t* signal[8];      // rank r's C[this rank]
    uint32_t* done;           // last-CTA counter in the own block
};
struct PvdbDpNetWait {        // consumer side: the rgbnet Adam CTAs of k_update_fused
    int world;
    uint32_t epoch;
    const float* src;         // own netx[parity]: [8][PVDB_DP_NET_PAD]
    const uint32_t* signal;   // own C[0..world)
    int32_t* err;
};

struct PvdbDpTilesWait {      // consumer side of the tile exchange: the leaf-Adam CTAs of k_update_fused
    int world;
    uint32_t epoch;
    const uint32_t* signal;   // own D[0..world)
    int32_t* err;
};
PvdbDpTilesWait pvdb_dp_tiles_wait_args(const pvdb_dp_peers* P, uint32_t step);
int pvdb_dp_union_early(const pvdb_dp_peers* P, const pvdb_train_bufs* b, uint32_t step, cudaStream_t st);
int pvdb_dp_move_tiles(const pvdb_dp_peers* P, const pvdb_train_bufs* b, uint32_t step, cudaStream_t st);
uint8_t* pvdb_dp_flags_ptr(const pvdb_dp_peers* P, uint32_t step);   // byte per leaf
PvdbDpNetPush pvdb_dp_net_push_args(const pvdb_dp_peers* P, uint32_t step);
PvdbDpNetWait pvdb_dp_net_wait_args(const pvdb_dp_peers* P, uint32_t step);
unsigned long long* pvdb_debug_stamps_ptr();   // train_step.cu: debug timeline (nullptr when off)
