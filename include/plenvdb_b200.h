/*
 * plenvdb_b200.h — C-ABI of the B200-native PlenVDB hot path.
 *
 * Every entry point is `extern "C"`, takes plain pointers and sizes, returns an int status
 * (0 = ok; pvdb_last_error() holds the message otherwise) and never throws.  No torch types.
 * All work is enqueued on the caller's CUDA stream (a `cudaStream_t` passed as void*); entry points
 * whose name ends in `_host` take HOST pointers (the reference's numpy contract), stage through
 * stream-ordered scratch and return after synchronising that stream.  Everything else takes DEVICE
 * pointers, performs no allocation and no synchronisation, and is CUDA-graph capturable.
 *
 * Citations are to the reference tree (wolfball/PlenVDB), relative to its root:
 *   B1  plenvdb/lib/vdb/plenvdb.cpp:3-172   (pybind11 module `plenvdb`)
 *   B2  plenvdb/lib/cuda/render_utils.cpp:170-184 (torch ext `render_utils_cuda`)
 * The Python mirror of both lives in plenvdb_b200/plenvdb.py and plenvdb_b200/render_utils_cuda.py.
 */
#ifndef PLENVDB_B200_H
#define PLENVDB_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PVDB_OK 0
#define PVDB_ERR_ARG 1
#define PVDB_ERR_CUDA 2
#define PVDB_ERR_STATE 3

/* Thread-local message of the last failing call. */
const char* pvdb_last_error(void);
/* ABI version, bumped when a signature changes. */
int pvdb_abi_version(void);

/* ------------------------------------------------------------------------------------------------
 * T1/T2 — sparse tree.  Same 5-4-3 hierarchy, same child/voxel offset functions and the same leaf
 * order as the NanoVDB grid the reference builds (openvdb/nanovdb/nanovdb/NanoVDB.h:2185-2203,
 * 3098-3127, 3426-3445, 3893-3900; util/OpenToNanoVDB.h:521-533), re-laid out for the GPU:
 * one topology shared by value/grad/exp_avg/exp_avg_sq planes, int32 child tables, and payload
 * planes that are plain leaf-major arrays  plane[leaf][512][C]  (C = 1 density, C = 12 colour).
 * ---------------------------------------------------------------------------------------------- */
typedef struct pvdb_tree {
    int32_t n_upper;             /* root tiles that have a child (32^3 x 16^3 x 8^3 = 4096^3 voxels each) */
    int32_t n_lower;             /* 16^3-leaf internal nodes */
    int32_t n_leaf;              /* 8^3-voxel leaves */
    int32_t reserved;
    uint64_t root_key0;          /* key of root tile 0 (NanoVDB.h:2702-2709), fast path for n_upper==1 */
    const uint64_t* root_keys;   /* [n_upper]  device */
    const int32_t* upper_child;  /* [n_upper][32768] -> lower index or -1   device */
    const int32_t* lower_child;  /* [n_lower][4096]  -> leaf index or -1    device */
    const int32_t* leaf_origin;  /* [n_leaf][3]                              device */
    const uint64_t* leaf_mask;   /* [n_leaf][8] value mask, bit n = word n>>6, bit n&63 (NanoVDB.h:1902) */
} pvdb_tree;

/* Host-side topology builder (opaque).  Replaces OpenVDB denseFill + openToNanoVDB
 * (plenvdb/lib/vdb/plenvdb.h:117-125, 197-210) and copyFromDense/pruned topologies. */
typedef struct pvdb_topo pvdb_topo;
/* denseFill(bbox [0,r-1]^3, 0, active=true): every 8^3 block touching the box is a leaf. */
pvdb_topo* pvdb_topo_create_dense(int rx, int ry, int rz);
/* Leaves where any voxel of `active` (host, uint8[rx*ry*rz], C order x-major) is set; value mask = active. */
pvdb_topo* pvdb_topo_create_from_mask(const uint8_t* active, int rx, int ry, int rz);
void pvdb_topo_destroy(pvdb_topo*);
int pvdb_topo_counts(const pvdb_topo*, int32_t* n_upper, int32_t* n_lower, int32_t* n_leaf);
/* Fill caller-provided HOST arrays (sizes as in pvdb_tree). */
int pvdb_topo_export(const pvdb_topo*, uint64_t* root_keys, int32_t* upper_child, int32_t* lower_child,
                     int32_t* leaf_origin, uint64_t* leaf_mask);

/* ------------------------------------------------------------------------------------------------
 * D1/D2/C1/C2 — trilinear sample forward / gradient scatter (densityvdb.cu:101-181, colorvdb.cu:81-175).
 * Coordinates are index-space floats, SoA.  `channels` = 1 -> density arithmetic (densityvdb.cu:116-123),
 * channels % 3 == 0 -> colour arithmetic (colorvdb.cu:16-26, 100-107).  corner_leaf/corner_off are
 * optional [n][8] outputs (leaf index or -1, voxel offset) in the reference's corner order.
 * ---------------------------------------------------------------------------------------------- */
int pvdb_sample_forward(const pvdb_tree* tree, const float* plane, int channels,
                        const float* xs, const float* ys, const float* zs, int64_t n,
                        float* out, int32_t* corner_leaf, int32_t* corner_off, void* stream);
int pvdb_sample_backward(const pvdb_tree* tree, float* grad_plane, int channels,
                         const float* xs, const float* ys, const float* zs, const float* grad_out,
                         int64_t n, void* stream);
/* Host-pointer variants honouring B1's numpy contract (plenvdb.h:441-485, 535-573). */
int pvdb_sample_forward_host(const pvdb_tree* tree, const float* plane, int channels,
                             const float* xs, const float* ys, const float* zs, int64_t n,
                             float* out, void* stream);
int pvdb_sample_backward_host(const pvdb_tree* tree, float* grad_plane, int channels,
                              const float* xs, const float* ys, const float* zs, const float* grad_out,
                              int64_t n, void* stream);
/* forward_single (densityvdb.cu:376-390, colorvdb.cu:380-399; colour starts from 0, not from
 * uninitialised memory as the reference does). */
int pvdb_sample_nearest(const pvdb_tree* tree, const float* plane, int channels,
                        const int32_t* is, const int32_t* js, const int32_t* ks, int64_t n,
                        float* out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * O1/O2/O3 — sparse Adam, zero_grad, dense <-> sparse (densityvdb.cu:31-97, 185-368; colorvdb.cu:42-77,
 * 179-371; host scalars plenvdb.h:751-767).
 * mode 0 = all active voxels, 1 = skip zero grad (colour: skip when all 3 comps of a Vec3 are 0),
 * 2 = multiply by per-voxel lr (perlr plane [n_leaf][512]).
 * ---------------------------------------------------------------------------------------------- */
float pvdb_adam_stepsize(float lr, float beta0, float beta1, int step);   /* plenvdb.h:753 in float */
int pvdb_adam_step(const pvdb_tree* tree, float* param, const float* grad, float* exp_avg, float* exp_avg_sq,
                   int channels, int mode, float stepsz, float eps, float beta0, float beta1,
                   const float* perlr, void* stream);
int pvdb_zero_grad(const pvdb_tree* tree, float* grad, int channels, void* stream);
/* dense layout [rx][ry][rz][channels] C order (plenvdb.h:149-157, 241-250). */
int pvdb_copy_from_dense(const pvdb_tree* tree, float* plane, int channels, const float* dense,
                         int rx, int ry, int rz, void* stream);
int pvdb_copy_to_dense(const pvdb_tree* tree, const float* plane, int channels, float* dense,
                       int rx, int ry, int rz, void* stream);
int pvdb_set_values_on_by_mask(const pvdb_tree* tree, float* plane, const uint8_t* mask, float val,
                               int rx, int ry, int rz, void* stream);

/* ------------------------------------------------------------------------------------------------
 * B2 — render_utils ops on raw device pointers (plenvdb/lib/cuda/render_utils_kernel.cu).
 * int64 outputs keep the reference's tensor dtypes.
 * ---------------------------------------------------------------------------------------------- */
int pvdb_infer_t_minmax(const float* rays_o, const float* rays_d, const float* xyz_min, const float* xyz_max,
                        float near, float far, int n_rays, float* t_min, float* t_max, void* stream);   /* :12-35 */
int pvdb_infer_n_samples(const float* rays_d, const float* t_min, const float* t_max, float stepdist,
                         int n_rays, int64_t* n_samples, void* stream);                                  /* :38-55 */
int pvdb_infer_ray_start_dir(const float* rays_o, const float* rays_d, const float* t_min, int n_rays,
                             float* rays_start, float* rays_dir, void* stream);                          /* :58-79 */
/* sample_pts_on_rays (:196-242) in two phases because the output length is data dependent:
 * count -> (caller reads n_steps_cumsum[n_rays-1], allocates) -> fill. */
int pvdb_sample_pts_count(const float* rays_o, const float* rays_d, const float* xyz_min, const float* xyz_max,
                          float near, float far, float stepdist, int n_rays,
                          float* t_min, float* t_max, int64_t* n_steps, int64_t* n_steps_cumsum,
                          float* rays_start, float* rays_dir, void* stream);
int pvdb_sample_pts_fill(const float* rays_start, const float* rays_dir, const float* xyz_min, const float* xyz_max,
                         const int64_t* n_steps_cumsum, float stepdist, int n_rays, int64_t total_len,
                         float* rays_pts, uint8_t* mask_outbbox, int64_t* ray_id, int64_t* step_id, void* stream);
int pvdb_maskcache_lookup(const uint8_t* world, const float* xyz, uint8_t* out, const float* xyz2ijk_scale,
                          const float* xyz2ijk_shift, int sz_i, int sz_j, int sz_k, int64_t n_pts, void* stream); /* :374-424 */
int pvdb_raw2alpha(const float* density, float shift, float interval, int64_t n_pts, float* exp_d, float* alpha,
                   void* stream);                                                                                  /* :431-481 */
int pvdb_raw2alpha_backward(const float* exp_d, const float* grad_back, float interval, int64_t n_pts, float* grad,
                            void* stream);                                                                         /* :507-552 */
/* alpha2weight (:577-651): weight/T/alphainv_last/i_start/i_end must be pre-initialised by the caller to
 * 0/1/1/0/0 exactly as the reference's host wrapper does (:625-629). */
int pvdb_alpha2weight(const float* alpha, const int64_t* ray_id, int64_t n_pts, int n_rays, float* weight, float* T,
                      float* alphainv_last, int64_t* i_start, int64_t* i_end, void* stream);
int pvdb_alpha2weight_backward(const float* alpha, const float* weight, const float* T, const float* alphainv_last,
                               const int64_t* i_start, const int64_t* i_end, int n_rays, const float* grad_weights,
                               const float* grad_last, float* grad, void* stream);                                 /* :654-707 */
/* adam_upd_cuda (plenvdb/lib/cuda/adam_upd_kernel.cu:9-132): mode 0 plain, 1 masked, 2 per-lr. */
int pvdb_dense_adam(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, const float* perlr,
                    int64_t n, int mode, int step, float beta1, float beta2, float lr, float eps, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Occupancy bits for the fused paths: mask_cache.mask (plenvdb/lib/grid.py:207-246) as a two-level
 * bit hierarchy.  fine[(bx*nby+by)*nbz+bz][8] holds the 512 voxel bits of an 8^3 block in the leaf's
 * own bit order; coarse has one bit per block (block index / 64, % 64).
 * ---------------------------------------------------------------------------------------------- */
int pvdb_occ_build(const uint8_t* mask, int rx, int ry, int rz, uint64_t* fine, uint64_t* coarse, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Fused fine-stage training step (new entry point B1/B2 callers can opt into).  Restates
 * DirectVoxGO.forward (plenvdb/lib/dvgo.py:296-388) + the losses and optimiser calls of
 * plenvdb/run.py:541-588 as a fixed sequence of kernels on one stream, no host sync.
 * ---------------------------------------------------------------------------------------------- */
typedef struct pvdb_train_cfg {
    /* scene */
    float xyz_min[3], xyz_max[3];
    int32_t reso[3];                 /* world_size */
    int32_t mask_reso[3];            /* mask_cache.mask shape */
    float mask_scale[3], mask_shift[3]; /* xyz2ijk_scale / shift (grid.py:229-231), computed by the caller */
    /* render kwargs */
    float near, far, stepdist, act_shift, interval, fast_color_thres, bg;
    /* loss weights (run.py:551-574) */
    float weight_main, weight_entropy_last, weight_rgbper;
    /* optimiser scalars; *_stepsz are the plenvdb.h:753 values for this step; rgbnet follows adam_upd_kernel.cu:72 */
    float den_stepsz, k0_stepsz, eps, beta0, beta1;
    int32_t den_mode, k0_mode;       /* the fused update implements stepmode 1 (fine stage, configs/default.py:77) */
    float net_lr; int32_t net_step;  /* rgbnet MaskedAdam: lr and step (>=1) of this iteration */
    int32_t k0_dim;                  /* 12 (fine stage) or 3 (coarse stage: rgb = sigmoid(k0), dvgo.py:344-346) */
    int32_t net_width;               /* 128, or 0 with k0_dim 3 (no rgbnet) */
    int32_t use_tensor_cores;        /* 1 = tcgen05 rgbnet, 0 = fp32 CUDA-core rgbnet */
    int32_t n_rays_global;           /* N in the means of the losses (== n_rays on one GPU; world*n_rays when sharded) */
    int32_t parity_counts;           /* 1: also produce the reference's untrimmed per-ray M2 counts (cnt_alpha_full) */
} pvdb_train_cfg;

typedef struct pvdb_train_bufs {
    /* grids: one topology, congruent planes */
    const pvdb_tree* tree;
    float *den, *den_grad, *den_m, *den_v;     /* [n_leaf][512] */
    float *k0, *k0_grad, *k0_m, *k0_v;         /* [n_leaf][512][12] */
    const uint64_t *occ_fine, *occ_coarse;     /* pvdb_occ_build */
    /* rgbnet params, PyTorch nn.Linear layout packed as w0[128][39] b0[128] w1[128][128] b1[128] w2[3][128] b2[3]
     * (22019 floats); grad/m/v use the same packing. */
    float *net, *net_grad, *net_m, *net_v;
    /* per-ray scratch [n_rays] */
    float *t_min, *t_max; int32_t* n_steps;
    int32_t *cnt_mask, *cnt_alpha, *cnt_keep, *cnt_alpha_full;  /* M1 (up to the stop), trimmed M2, M3, full M2 */
    int32_t *off_alpha, *off_keep;             /* exclusive scans = segment offsets, [n_rays+1] */
    float *alphainv_last, *rgb_marched /*[n_rays][3]*/, *grad_last;
    /* alpha list: one entry per sample with alpha > thres up to the early stop, ray order; capacity cap_alpha */
    int64_t cap_alpha, cap_keep;
    int32_t *s_ray, *s_step;
    float *s_xyz /*[.][3] index-space coords*/, *s_density, *s_alpha, *s_T, *s_weight, *s_gden;
    /* kept list: entries with weight > thres; capacity cap_keep */
    int32_t *k_sample /* index into the alpha list */, *k_ray;
    float *k_xyz /*[.][3]*/, *k_feat /*[.][12]*/, *k_rgb /*[.][3] rgb, then dL/dlogit*/, *k_gw /*[.] dL/dweight*/;
    float *k_h0, *k_h1;                        /* post-ReLU activations kept for the backward: [.][128] row-major (fp32 path) or
                                                * chunk-major [tile][8 chunks][128 features][16 samples] (tensor-core path; the tile-major
                                                * tensors k_h0, k_h1, k_x, k_dh0, k_mask hold cap_keep ROUNDED UP to 128 rows: whole tiles are written) */
    float *k_x;                                /* chunk-major [tile][8][40][16] rgbnet inputs (12 k0 + 27 PE + a ones row), tensor-core path */
    float *k_dh0, *k_dh1;                      /* k_dh0: chunk-major [tile][8][128][16] masked activation gradient of layer 0 (tensor-core
                                                * backward); k_dh1 is unused (the weight-gradient pass recomputes dH1), may be NULL */
    uint32_t *k_mask;                          /* [tile][8][128] ReLU sign bits of h0 (words 0-3) and h1 (words 4-7) */
    int32_t *k_corner;                         /* [cap_keep][8] record id (leaf*512+voxel) of the 8 trilinear corners, -1 = none */
    void *net_img;                             /* >= 512 KiB scratch: tf32 hi/lo weight images of the tensor-core kernels */
    float *net_partial;                        /* [148][22048] per-CTA weight-gradient partial sums (tensor-core backward) */
    void *march_scratch;                       /* optional: 20 * scratch_rays * scratch_per_ray words; the count pass parks its
                                                * per-sample results here and the emit pass becomes a compaction (NULL: march twice) */
    int32_t scratch_rays, scratch_per_ray;
    /* touched-leaf bookkeeping, [n_leaf] each */
    int32_t *den_touched, *k0_touched, *den_touched_list, *k0_touched_list;
    int32_t *counters;                         /* [16]: 0 M_alpha, 1 M_keep, 2 n_touched_den, 3 overflow flag, 4 n_touched_k0,
                                                * 5 ray ticket of the count pass (zero between steps) */
    float *loss;                               /* [4]: total, mse, entropy_last, rgbper */
    const float *den_perlr;                    /* coarse stage, den_mode 2: per-voxel lr plane [n_leaf][512] (masked_adam.py:43-46) */
    /* optional, all six or none — the leaf-local alternative for the k0 features (csrc/leaf_local.cu; PVDB_LEAF_LOCAL=1): kept
     * samples grouped by home leaf, forward gather through leaf tiles staged in shared memory, backward accumulation in a
     * shared tile per leaf */
    int32_t *ll_cnt, *ll_off /* [n_leaf+1] */, *ll_cur, *ll_list;   /* [n_leaf] each unless noted */
    int32_t *ll_items;                         /* [cap_keep] kept-sample ids grouped by home leaf */
    float *k_dx;                               /* [cap_keep][12] dL/d(k0 features) of every kept sample */
    const float *step_scalars;                 /* optional device-readable array [4] (device memory or pinned host memory): den_stepsz, k0_stepsz, rgbnet Adam step size (lr with the
                                                * bias corrections, pvdb_dense_adam_stepsize), reserved.  When non-NULL the update
                                                * kernels read the per-iteration scalars from here instead of cfg, so that a captured
                                                * CUDA graph of the step can be replayed for every iteration */
    float *ray_pe;                             /* optional [n_rays][28] scratch (tensor-core path): view-direction embedding of every ray
                                                * (dvgo.py:354-357: viewdirs, sin, cos, one zero of padding), filled once per step
                                                * underneath the march and read by the rgbnet forward for the ~10 kept samples of a
                                                * ray; NULL: the forward evaluates the 24 sinf / cosf per sample */
} pvdb_train_bufs;

#define PVDB_PHASE_FORWARD 1    /* sample, interpolate, rgbnet, composite (+ losses when target != NULL) */
#define PVDB_PHASE_BACKWARD 2   /* gradients into den_grad / k0_grad / net_grad (accumulating, like autograd) */
#define PVDB_PHASE_UPDATE 4     /* sparse Adam on touched leaves + rgbnet Adam; clears the gradients it consumed */
#define PVDB_PHASE_ACCUMULATE 16 /* with FORWARD: gradients of an earlier BACKWARD have not been consumed by an UPDATE yet (gradient accumulation
                                  * over several batches): keep their leaves on the touched lists */
#define PVDB_PHASE_LISTS_READY 8 /* with UPDATE alone: den/k0_touched_list + counters[2],[4] are already valid (pvdb_dp_exchange) */
int pvdb_train_step(const pvdb_train_cfg* cfg, const pvdb_train_bufs* bufs,
                    const float* rays_o, const float* rays_d, const float* viewdirs, const float* target,
                    int n_rays, int phases, void* stream);
/* Gather a batch ([n][3] arrays anywhere in device memory; target may be NULL) into one staging buffer [4][n][3] with a single
 * launch, so that a captured CUDA graph of pvdb_train_step can read its inputs from fixed addresses. */
int pvdb_stage_rays(const float* rays_o, const float* rays_d, const float* viewdirs, const float* target, int n_rays, float* stage,
                    void* stream);
/* rgbnet Adam step size, lr * sqrt(1 - beta1^step) / (1 - beta0^step) in float (adam_upd_kernel.cu:72): what the fused step
 * derives from cfg->net_lr / net_step, for callers that feed pvdb_train_bufs.step_scalars[2] themselves. */
float pvdb_dense_adam_stepsize_host(float lr, float beta0, float beta1, int step);
/* Test switch for the renderer's first pass: 0 = one pixel per thread (k_render_pass1), 1 = probe every pixel and march the hit
 * ones with 8 lanes each, negative = the default rule (lane-parallel when a call renders at most half of the frame's rows).
 * Bit-identical results; environment PVDB_RENDER_LANES = 0 / 1 forces a mode at start-up. */
void pvdb_debug_set_render_lanes(int on);
/* Debug timeline (environment PVDB_STAMPS=1): %globaltimer stamps written by one-thread kernels at marked points of the fused
 * step (main stream: 0 start, 1 emit, 2 rgbnet forward, 3 composite, 4 activation gradients, 5 weight gradients + reduction,
 * 6 rgbnet Adam, 7 join; side stream: 10/11 around the union, 12 start of the tile exchange, 13 its end, 14 leaf Adam).
 * Copies the 64 stamps of the last step to the host; returns non-zero when stamping is off. */
int pvdb_debug_stamps_fetch(unsigned long long* out64);
/* Selects the leaf-local path of the k0 features (needs the ll_* / k_dx buffers): k_feat is bit-identical, gradients agree to
 * 1e-5 like any reordered float sum.  Default off (environment PVDB_LEAF_LOCAL=1 switches it on at start-up). */
void pvdb_debug_set_leaf_local(int on);
/* Test switch: 0 makes the march of pvdb_train_step test the occupancy of every step one by one instead of skipping runs of
 * steps that provably cannot hit the mask (the results are bit-identical either way; default 1). */
void pvdb_debug_set_run_skip(int on);
/* Data-parallel gradient exchange (no counterpart in the reference, which is single-GPU).  Call after the ranks' touched
 * flags (bufs->den_touched / k0_touched) were MAX-all-reduced: builds the union leaf list on the device, copies its
 * length to *union_count_host (pinned; valid after the stream is synchronised) and packs the union leaves' gradient
 * tiles [n][512*13] followed by the 22019 rgbnet gradients into `buf` for ONE sum all-reduce of
 * (n*6656 + 22019) floats.  pvdb_dp_unpack writes the reduced values back. */
int pvdb_dp_pack(const pvdb_train_bufs* bufs, int32_t* union_list, int32_t* union_count_dev, int32_t* union_count_host,
                 float* buf, int64_t buf_capacity_floats, void* stream);
int pvdb_dp_unpack(const pvdb_train_bufs* bufs, const int32_t* union_list, const int32_t* union_count_dev, float* buf,
                   void* stream);
/* The same exchange over NVLink peer memory, without NCCL and without a host synchronisation (dp_exchange.cu).  Every
 * rank allocates one symmetric block (pvdb_dp_symm_bytes / _alloc: cudaMalloc + a 64-byte CUDA IPC handle), the handles
 * are exchanged by the caller's own transport (torch.distributed all_gather_object here) and opened with
 * pvdb_dp_symm_open; base[r] is rank r's block as mapped in THIS process (base[rank] = the own allocation).
 * pvdb_dp_exchange(step) runs after the backward phase on every rank with the same monotone `step` (0,1,2,...):
 * cross-GPU barrier, union of touched leaves, reduce-scatter + all-gather over peer memory IN PLACE on the gradient planes,
 * which live inside the symmetric blocks (every rank sums the union leaves it owns over all ranks' planes in rank order and
 * stores the sums into every rank's planes: O(1) NVLink bytes per rank, no staging copy), rgbnet gradients into net_grad, and
 * leaves den/k0_touched_list + counters[2],[4] ready for pvdb_train_step(PVDB_PHASE_UPDATE|PVDB_PHASE_LISTS_READY).
 * A peer that does not arrive within 2 s sets the block's error word (pvdb_dp_symm_error: 1 timeout). */
typedef struct {
    int32_t world, rank;      /* world <= 8 (one NVSwitch domain) */
    int32_t n_leaf, cap_leaves;
    void* base[8];
} pvdb_dp_peers;
size_t pvdb_dp_symm_bytes(int n_leaf, int cap_leaves /* unused */);
/* The gradient planes INSIDE this rank's symmetric block: den_grad [n_leaf][512], k0_grad [n_leaf][512][12].  The exchange reads
 * and writes the peers' planes in place (no pack / unpack), so pvdb_train_bufs.den_grad / k0_grad must be exactly these. */
int pvdb_dp_grad_planes(const pvdb_dp_peers* peers, float** den_grad, float** k0_grad);
int pvdb_dp_symm_alloc(size_t bytes, void** ptr, void* handle64);
int pvdb_dp_symm_open(const void* handle64, void** ptr);
int pvdb_dp_symm_close(void* ptr);
int pvdb_dp_symm_free(void* ptr);
int pvdb_dp_symm_error(const pvdb_dp_peers* peers, int32_t* err_out);
int pvdb_dp_exchange(const pvdb_dp_peers* peers, const pvdb_train_bufs* bufs, uint32_t step, void* stream);
/* The two independent halves of pvdb_dp_exchange: grid-gradient tiles (needs the k0 / density scatters) and the rgbnet
 * gradients (needs the weight-gradient kernel). */
int pvdb_dp_exchange_tiles(const pvdb_dp_peers* peers, const pvdb_train_bufs* bufs, uint32_t step, void* stream);
int pvdb_dp_exchange_net(const pvdb_dp_peers* peers, const pvdb_train_bufs* bufs, uint32_t step, void* stream);
/* One data-parallel iteration on this rank's ray shard (cfg->n_rays_global = rays of all ranks).  The emit kernel writes the
 * rank's touched-leaf flags (they follow from the sample lists), so the union and its cross-GPU barrier run on an internal
 * side stream UNDER the rgbnet forward; pack / reduce / unpack and the leaf Adam run UNDER the weight-gradient kernel; the
 * rgbnet gradients are pushed to the peers by the weight-gradient reduction and summed by the rgbnet Adam (no exchange kernel
 * on the critical path).  Every rank must call it with the same dp_step (0,1,2,...) and the same step as pvdb_dp_exchange
 * would get: the two share the epoch counters of the symmetric block. */
int pvdb_train_step_dp(const pvdb_train_cfg* cfg, const pvdb_train_bufs* bufs, const pvdb_dp_peers* peers, uint32_t dp_step,
                       const float* rays_o, const float* rays_d, const float* viewdirs, const float* target, int n_rays,
                       void* stream);
/* ---- grid maintenance on the device: the reference does each through dense host arrays ----------
 * scale_volume_grid (plenvdb/lib/grid.py:91-101): trilinear resample (F.interpolate, align_corners=True) of the dense view
 * of (src_tree, src_plane) at resolution s* into the ACTIVE voxels of (dst_tree, dst_plane) at resolution d*. */
int pvdb_resample_trilinear(const pvdb_tree* src_tree, const float* src_plane, int channels, int sx, int sy, int sz,
                            const pvdb_tree* dst_tree, float* dst_plane, int dx, int dy, int dz, void* stream);
/* Re-sparsification: every voxel slot of dst takes the value stored at the same coordinates in src (0 where src has no leaf). */
int pvdb_plane_remap(const pvdb_tree* src_tree, const float* src_plane, const pvdb_tree* dst_tree, float* dst_plane, int channels,
                     void* stream);
/* update_occupancy_cache (plenvdb/lib/dvgo.py:201-210): mask &= maxpool3(raw2alpha(density(mask voxel centres))) > thres.
 * mask: device uint8 [mx*my*mz]; alpha_tmp: device float scratch of the same count; xyz_min/xyz_max: HOST float[3]. */
int pvdb_occupancy_update(const pvdb_tree* tree, const float* den_plane, int rx, int ry, int rz, const float* xyz_min,
                          const float* xyz_max, float act_shift, float interval, float thres, uint8_t* mask, int mx, int my, int mz,
                          float* alpha_tmp, void* stream);
/* total_variation_add_grad (plenvdb/lib/cuda/total_variation_kernel.cu:14-35) on the sparse planes with the dense kernel's
 * semantics (neighbours outside the tree read the background 0); dense_mode = 0 only touches voxels whose grad is non-zero. */
int pvdb_total_variation_add_grad(const pvdb_tree* tree, const float* plane, float* grad, int channels, int rx, int ry, int rz,
                                  float wx, float wy, float wz, int dense_mode, void* stream);
/* hit_coarse_geo (plenvdb/lib/dvgo.py:253-270) for the 'in_maskcache' ray sampler: hit[r] = 1 iff some in-bbox sample
 * of ray r lands in an occupied voxel.  Uses cfg's scene scalars and bufs->occ_*. */
int pvdb_rays_hit_mask(const pvdb_train_cfg* cfg, const pvdb_train_bufs* bufs, const float* rays_o, const float* rays_d,
                       int n_rays, uint8_t* hit, void* stream);
/* Per-kernel timing of the fused calls (CUDA events on the launching stream), for the benchmark's roofline line. */
int pvdb_profile_enable(int on);
int pvdb_profile_fetch(int max_segments, float* ms, char* names /* [max_segments][32] */);
/* Number of this library's kernel launches enqueued by the last pvdb_train_step / pvdb_render call on this thread. */
int pvdb_last_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * R1/R2 — merged-VDB renderer (plenvdb/lib/vdb/renderer.cu:370-424, plenvdb.h:933-1068).
 * ---------------------------------------------------------------------------------------------- */
typedef struct pvdb_render_cfg {
    int32_t reso[3];
    float K[9];
    float xyz_min[3], xyz_max[3];
    float near, far, stepdist, act_shift, interval, fast_color_thres, bg;
    int32_t inverse_y, H, W;
    int32_t dcol, dpe, dhid, dout;   /* 12, 27, 128, 3 */
    int32_t use_tensor_cores;
} pvdb_render_cfg;

typedef struct pvdb_render_bufs {
    const pvdb_tree* idx_tree;       /* merged index grid: leaves where mask_cache.mask has a voxel, value mask = mask */
    const int32_t* idx_plane;        /* [n_leaf][512] 1-based row ids (vdb_compression.py:31-33), 0 = none */
    const float* dendata;            /* [N+1], row 0 = 0 */
    const float* coldata;            /* [N+1][12], row 0 = 0 */
    const float *w0, *b0, *w1, *b1, *w2, *b2;  /* transposed layout of run.py:98-104: w0[39][128], w1[128][128], w2[128][3] */
    /* scratch for the rows of one call, npix = (row_end-row_begin)*W */
    int32_t *n_samples;              /* [npix] */
    int32_t *i_starts;               /* [npix+1] exclusive scan = segment offsets */
    float *tmins, *tmaxs;            /* [npix] tightened range of pass 1 */
    int32_t *scan_tmp;               /* [npix/4096 + 2] */
    int64_t cap_samples;
    int32_t *s_ray;                  /* [cap] pixel index local to the band */
    float *s_weight;                 /* [cap] */
    float *s_feat;                   /* [cap][12] */
    float *s_rgb;                    /* [cap][3] weight * sigmoid(rgbnet) */
    int32_t *counters;               /* [8]: 0 total samples, 1 overflow flag, 2 rays whose two passes disagree,
                                      * 3 pixels with samples (zero between frames), 4 pixels marched a second time */
    void *w_img;                     /* >= 256 KiB scratch: tf32 hi/lo weight image (use_tensor_cores) */
    int32_t *active_list;            /* [npix] pixels with samples, built by pass 1 for pass 2 */
    const uint32_t *skip_bits;       /* optional [pvdb_render_block_bits_words(reso)]: dilated block map of idx_tree built by
                                      * pvdb_render_block_bits; runs of march steps that cannot touch a leaf are skipped (their
                                      * `t += steplen` chain is still evaluated, so results are bit-identical).  NULL = march
                                      * every step like the reference */
    /* optional hand-over of pass 1's samples (all three or none): pass 1 also runs pass 2's arithmetic on the corner values it
     * has loaded and parks (t, weight) of the first px_entries kept samples of every pixel in px_scratch[npix][px_entries]
     * (8 bytes each); k_render_emit turns them into the pixel's segment of the sample list.  Only pixels with more samples,
     * or for which the simulation is not exact (restarted `t` chain, count at a threshold), are marched a second time */
    void *px_scratch;
    int32_t *fallback_list;          /* [npix] */
    int32_t px_entries, reserved;
} pvdb_render_bufs;
size_t pvdb_render_block_bits_words(int rx, int ry, int rz);
int pvdb_render_block_bits(const pvdb_tree* idx_tree, int rx, int ry, int rz, uint32_t* bits, void* stream);

/* Renders rows [row_begin,row_end) of the H x W image for camera `c2w` (device float[16], row-major 4x4) into
 * out_rgb (device float[(row_end-row_begin)*W*3]).  No allocation, no synchronisation. */
int pvdb_render_rows(const pvdb_render_cfg* cfg, const pvdb_render_bufs* bufs, const float* c2w,
                     int row_begin, int row_end, float* out_rgb, void* stream);
/* ---- tile-sharded rendering across the GPUs of one box (the reference renders on one GPU) ---------------------
 * Rows are dealt to the ranks in groups of band_rows rows, round-robin (row r belongs to rank (r / band_rows) % world), so
 * every rank gets the same share of the object whatever rows it occupies.  pvdb_interleaved_rows = rows that fall to `rank`. */
int pvdb_interleaved_rows(int H, int band_rows, int rank, int world);
/* Renders this rank's rows into band_out (device float[rows*W*3], local rows in ascending image order) and, when frame_out is
 * not NULL, also stores every pixel at its place in the full frame frame_out (device float[H*W*3]; it may be ANOTHER GPU's
 * memory mapped through CUDA IPC — plain stores over NVLink).  No allocation, no synchronisation. */
int pvdb_render_rows_interleaved(const pvdb_render_cfg* cfg, const pvdb_render_bufs* bufs, const float* c2w, int band_rows,
                                 int rank, int world, float* band_out, float* frame_out, void* stream);
/* Frame assembly without a collective: every rank owns one symmetric block of pvdb_frame_symm_bytes(H, W) bytes (allocate with
 * pvdb_dp_symm_alloc, exchange the IPC handles by any transport, map with pvdb_dp_symm_open; base[r] = rank r's block as mapped
 * in THIS process).  pvdb_render_frame_sharded(frame_no = 0,1,2,... identical on all ranks) renders this rank's interleaved rows
 * and its composite kernel writes them straight into ROOT's frame buffer (frame_no & 1) over NVLink; a release/acquire signal
 * per rank replaces the gather.  When the call's work has completed on root's stream, pvdb_frame_ptr(frame_no) on root is the
 * full frame; it stays valid until root's call for frame_no + 2 starts (consume it in stream order before that).  A rank that
 * does not arrive within 2 s sets the error word (pvdb_frame_error: 1) instead of hanging the GPU. */
typedef struct {
    int32_t world, rank, root;   /* world <= 8 (one NVSwitch domain) */
    int32_t H, W;
    int32_t reserved;
    void* base[8];
} pvdb_frame_peers;
size_t pvdb_frame_symm_bytes(int H, int W);
int pvdb_render_frame_sharded(const pvdb_render_cfg* cfg, const pvdb_render_bufs* bufs, const pvdb_frame_peers* peers,
                              const float* c2w, int band_rows, uint32_t frame_no, float* band_out, void* stream);
int pvdb_frame_ptr(const pvdb_frame_peers* peers, uint32_t frame_no, float** frame);
/* Stream-ordered copy of the assembled frame into caller memory (device float[H*W*3]) for callers that cannot alias the block. */
int pvdb_frame_copy(const pvdb_frame_peers* peers, uint32_t frame_no, float* dst, void* stream);
int pvdb_frame_error(const pvdb_frame_peers* peers, int32_t* err_out);
/* Device-side half of the merge (vdb_compression.py:36-57): gathers density / colour rows of every masked voxel and
 * rounds them through fp16.  row_of_voxel [rx*ry*rz]: 1-based row id or 0. */
int pvdb_merge_gather(const pvdb_tree* tree, const float* den, const float* k0, int k0_dim,
                      const int32_t* row_of_voxel, int rx, int ry, int rz, float* dendata, float* coldata, void* stream);

/* ------------------------------------------------------------------------------------------------
 * B1 as opaque handles (csrc/handles.cu) — one handle per class of the reference's pybind11 module `plenvdb`
 * (plenvdb/lib/vdb/plenvdb.cpp:3-172): what a C / C++ / pybind11 binder calls instead of managing topology upload, plane
 * allocation and optimiser step counting itself.  Host-buffer contract like the pybind11 module (every call is complete when
 * it returns); the device pointers behind a handle are exposed for zero-copy callers.  `.vdb` load / save is not part of the
 * C layer (plenvdb_b200/openvdb_io.py): the handles take and return dense host arrays, like copyFromDense / get_dense_grid.
 * ---------------------------------------------------------------------------------------------- */
typedef struct pvdb_grid pvdb_grid;          /* DensityVDB (channels 1) / ColorVDB (channels 12): plenvdb.h:389-429 */
typedef struct pvdb_opt pvdb_opt;            /* DensityOpt / ColorOpt: plenvdb.h:686-789 */
typedef struct pvdb_renderer pvdb_renderer;  /* MGRenderer: plenvdb.h:933-1068 */
/* active == NULL: denseFill topology (plenvdb.h:117-125); else uint8 [rx*ry*rz]: leaves where it has a voxel, value mask = active */
pvdb_grid* pvdb_grid_create(int rx, int ry, int rz, int channels, const uint8_t* active);
void pvdb_grid_destroy(pvdb_grid*);
int pvdb_grid_info(const pvdb_grid*, int32_t* reso3, int32_t* channels, int32_t* n_leaf);
const pvdb_tree* pvdb_grid_tree(const pvdb_grid*);     /* for the stateless entry points */
float* pvdb_grid_values(pvdb_grid*);                   /* device plane [n_leaf][512][channels] */
float* pvdb_grid_grad(pvdb_grid*);
int pvdb_grid_copy_from_dense(pvdb_grid*, const float* dense_host);   /* [rx][ry][rz][channels]; plenvdb.h:149-157, 241-250 */
int pvdb_grid_copy_to_dense(const pvdb_grid*, float* dense_host);     /* plenvdb.h:158-167, 251-271 */
int pvdb_grid_forward(const pvdb_grid*, const float* x, const float* y, const float* z, int64_t n, float* out_host);   /* plenvdb.cpp:10-31 */
int pvdb_grid_backward(pvdb_grid*, const float* x, const float* y, const float* z, const float* grad_host, int64_t n);  /* plenvdb.cpp:47-67 */
int pvdb_grid_set_values_on_by_mask(pvdb_grid*, const uint8_t* mask_host, float val);                                 /* plenvdb.h:487-495 */
pvdb_opt* pvdb_opt_create(pvdb_grid*, float lr, float eps, float beta0, float beta1);
void pvdb_opt_destroy(pvdb_opt*);
int pvdb_opt_zero_grad(pvdb_opt*);
int pvdb_opt_step(pvdb_opt*, int stepmode);            /* 0 plain, 1 skip zero gradients, 2 per-voxel lr (plenvdb.h:751-789) */
int pvdb_opt_update_lr(pvdb_opt*, float factor);
int pvdb_opt_set_pervoxel_lr(pvdb_opt*, const float* dense_host);     /* [rx*ry*rz]; plenvdb.h:714-722 */
int pvdb_opt_get(const pvdb_opt*, int32_t* step, float* lr, float* eps, float* beta0, float* beta1);
int pvdb_opt_set(pvdb_opt*, int32_t step, float lr, float eps, float beta0, float beta1);
float* pvdb_opt_exp_avg(pvdb_opt*);
float* pvdb_opt_exp_avg_sq(pvdb_opt*);
pvdb_renderer* pvdb_renderer_create(int dcol, int dpe, int dhid, int dout);   /* (12, 27, 128, 3) */
void pvdb_renderer_destroy(pvdb_renderer*);
/* den [n_rows], col [n_rows][12] (row 0 = zeros), idx_dense_host int32 [rx][ry][rz] of 1-based row ids (plenvdb.h:959-983) */
int pvdb_renderer_load_data(pvdb_renderer*, const float* den_host, const float* col_host, int64_t n_rows, const int32_t* idx_dense_host,
                            int rx, int ry, int rz);
int pvdb_renderer_load_params(pvdb_renderer*, const float* w0, const float* b0, const float* w1, const float* b1, const float* w2, const float* b2);
int pvdb_renderer_set_scene(pvdb_renderer*, const int32_t* reso3, const float* K9, const float* xyz_min3, const float* xyz_max3);
int pvdb_renderer_set_kwargs(pvdb_renderer*, float near, float far, float stepdist, float act_shift, float interval, float fast_color_thres,
                             float bg, int inverse_y, int H, int W);
int pvdb_renderer_input_c2w(pvdb_renderer*, const float* c2w16_host);
/* render_an_image + output_an_image: out_host float [H][W][3] (NULL: leave the frame on the device, pvdb_renderer_frame).  Like the
 * reference it does nothing until all five setup calls were made: *rendered tells. */
int pvdb_renderer_render(pvdb_renderer*, float* out_host, int* rendered);
const float* pvdb_renderer_frame(const pvdb_renderer*);
int pvdb_renderer_counters(const pvdb_renderer*, int32_t* counters8);

#ifdef __cplusplus
}
#endif
#endif /* PLENVDB_B200_H */
