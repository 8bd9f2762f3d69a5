// leaf_local.cuh — entry points of csrc/leaf_local.cu for the fused step.
#pragma once
#include "common.cuh"
bool pvdb_leaf_local_enabled(const pvdb_train_bufs* b);
int pvdb_leaf_local_forward(const pvdb_train_bufs* b, cudaStream_t st);    // after the emit kernel: buckets + k_feat
int pvdb_leaf_local_backward(const pvdb_train_bufs* b, cudaStream_t st);   // after the activation-gradient kernel stored k_dx
