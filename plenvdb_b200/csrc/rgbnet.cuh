// rgbnet.cuh — interface between the fused step and the rgbnet kernels (k0 gather + MLP forward/backward).
#pragma once
#include <cuda_runtime.h>
#include "../../include/plenvdb_b200.h"

// Packed parameter layout (PyTorch nn.Linear, dvgo.py:99-107): w0[128][39] b0[128] w1[128][128] b1[128] w2[3][128] b2[3]
#define PVDB_NET_DIN 39
#define PVDB_NET_W 128
#define PVDB_NET_OFF_W0 0
#define PVDB_NET_OFF_B0 (PVDB_NET_OFF_W0 + PVDB_NET_W * PVDB_NET_DIN)
#define PVDB_NET_OFF_W1 (PVDB_NET_OFF_B0 + PVDB_NET_W)
#define PVDB_NET_OFF_B1 (PVDB_NET_OFF_W1 + PVDB_NET_W * PVDB_NET_W)
#define PVDB_NET_OFF_W2 (PVDB_NET_OFF_B1 + PVDB_NET_W)
#define PVDB_NET_OFF_B2 (PVDB_NET_OFF_W2 + 3 * PVDB_NET_W)
#define PVDB_NET_N (PVDB_NET_OFF_B2 + 3)   // 22019

// Sample-independent per-step preparation (weight images of the tensor-core kernels); must precede the forward.
// viewdirs / n_rays: the batch whose forward follows (per-ray view embedding table, when b->ray_pe is there).
int pvdb_rgbnet_prepare(const pvdb_train_cfg* cfg, const pvdb_train_bufs* b, const float* viewdirs, int n_rays, cudaStream_t st);
// Forward over the kept-sample list: k_feat, k_h0, k_h1 (fp32 path), k_rgb.  Enqueues on `st`.
int pvdb_rgbnet_forward(const pvdb_train_cfg* cfg, const pvdb_train_bufs* b, const float* viewdirs, cudaStream_t st);
// Backward: consumes g_logit (in k_rgb), produces net_grad (zeroed first) and scatters the k0 gradient.
int pvdb_rgbnet_backward(const pvdb_train_cfg* cfg, const pvdb_train_bufs* b, const float* viewdirs, cudaStream_t st);

// The two halves of the tensor-core backward (cfg->use_tensor_cores only): the data-parallel step exchanges the grid
// gradients, which are final after the first half, underneath the second.
int pvdb_rgbnet_backward_act_tc(const pvdb_train_cfg* cfg, const pvdb_train_bufs* b, const float* viewdirs, cudaStream_t st);
struct PvdbDpNetPush;   // dp_exchange.cuh; nullptr outside a data-parallel step
// The dense Adam of the 22019 rgbnet parameters (adam_upd_kernel.cu:9-23) applied by the kernel that sums the per-CTA
// weight-gradient partials, to the element it has just summed (data parallel: after the ranks' sums have arrived): the step ends
// one kernel earlier.  on = 0: the sums
// are only accumulated into net_grad.
struct PvdbNetAdam { int on; float *net, *m, *v; float stepsize, b0, b1, eps; const float* scalars; };
struct PvdbDpNetWait;   // data-parallel step with the Adam in the reduction: the CTAs wait for every rank's pushed sums first
int pvdb_rgbnet_backward_wgrad_tc(const pvdb_train_cfg* cfg, const pvdb_train_bufs* b, cudaStream_t st, const PvdbDpNetPush* dp_push,
                                  const PvdbNetAdam* adam = nullptr, const PvdbDpNetWait* dp_wait = nullptr);

// Coarse stage (k0_dim == 3, no rgbnet): rgb = sigmoid(k0) (coarse.cu)
static inline bool pvdb_direct_colour(const pvdb_train_cfg* cfg) { return cfg->k0_dim == 3 && cfg->net_width == 0; }
int pvdb_direct_forward(const pvdb_train_bufs* b, cudaStream_t st);
int pvdb_direct_backward(const pvdb_train_bufs* b, cudaStream_t st);
