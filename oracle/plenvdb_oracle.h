/*
 * plenvdb_oracle.h — C API of the CPU oracle.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a plain CPU restatement of the reference's algorithm for the hot
 * path (wolfball/PlenVDB; every function cites the reference file:line it follows).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it; the product
 * (plenvdb_b200/) never does.
 *
 * Parity pinning: the restatement is checked (tests/test_oracle_pins.py) against
 *   (a) oracle/_ref/libref_host.so — the same kernel bodies executed through the reference's OWN
 *       NanoVDB.h ReadAccessor / GridBuilder on the host (built from /root/reference in place), and
 *   (b) tests/golden/*.npz — outputs of the reference's own CUDA kernels (oracle/_ref/libref_gpu.so,
 *       compiled unmodified for sm_100a) captured on a B200 by tests/golden/make_golden.py.
 * The reference itself ships no tests or golden vectors for this path.
 */
#ifndef PLENVDB_ORACLE_H
#define PLENVDB_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_grid orc_grid;

/* Tree with `channels` payload (1 = NanoGrid<float>; 3k = k NanoGrid<Vec3f>).  active == NULL: denseFill of
 * [0,r-1]^3 (plenvdb.h:117-125); else leaves where any active voxel, value mask = active (copyFromDense /
 * pruned topologies).  All values start at background 0. */
orc_grid* orc_grid_create(int rx, int ry, int rz, int channels, const uint8_t* active);
void orc_grid_destroy(orc_grid*);
int orc_grid_leaf_count(const orc_grid*);
void orc_grid_leaf_origins(const orc_grid*, int32_t* out /*[n_leaf][3]*/);
void orc_grid_leaf_masks(const orc_grid*, uint64_t* out /*[n_leaf][8]*/);
/* dense layout [rx][ry][rz][channels] */
void orc_grid_copy_from_dense(orc_grid*, const float* dense);   /* active voxels only (densityvdb.cu:31-49) */
void orc_grid_copy_to_dense(const orc_grid*, float* dense);     /* stored values where a leaf exists, else 0 */
void orc_grid_set_on_by_mask(orc_grid*, const uint8_t* mask, float val);   /* densityvdb.cu:64-84 */
void orc_grid_fill(orc_grid*, float v);                          /* every slot of every leaf (test helper) */
/* leaf-major payload [n_leaf][512][channels] in / out, for grids too large for a dense host array (test helper) */
void orc_grid_set_values(orc_grid*, const float* plane);
void orc_grid_get_values(const orc_grid*, float* plane);

/* D1/C1, D2/C2 (densityvdb.cu:101-167, colorvdb.cu:81-160).  corner_* optional [n][8]. */
void orc_sample_forward(const orc_grid*, const float* xs, const float* ys, const float* zs, int64_t n, float* out,
                        int32_t* corner_leaf, int32_t* corner_off, int threads);
void orc_sample_backward(orc_grid* grad, const float* xs, const float* ys, const float* zs, const float* gout, int64_t n,
                         int threads);
/* O1/O2 (densityvdb.cu:185-368, colorvdb.cu:179-371; plenvdb.h:751-767) */
float orc_adam_stepsize(float lr, float beta0, float beta1, int step);
void orc_adam_step(orc_grid* param, const orc_grid* grad, orc_grid* exp_avg, orc_grid* exp_avg_sq, int mode, float stepsz,
                   float eps, float beta0, float beta1, const orc_grid* perlr);
void orc_zero_grad(orc_grid* grad);

/* B2 ops (plenvdb/lib/cuda/render_utils_kernel.cu) */
void orc_infer_t_minmax(const float* rays_o, const float* rays_d, const float* xyz_min, const float* xyz_max, float near,
                        float far, int n_rays, float* t_min, float* t_max);
void orc_infer_n_samples(const float* rays_d, const float* t_min, const float* t_max, float stepdist, int n_rays,
                         int64_t* n_samples);
void orc_infer_ray_start_dir(const float* rays_o, const float* rays_d, const float* t_min, int n_rays, float* rays_start,
                             float* rays_dir);
/* returns total_len; call once with NULL outputs to size, then again to fill */
int64_t orc_sample_pts_on_rays(const float* rays_o, const float* rays_d, const float* xyz_min, const float* xyz_max,
                               float near, float far, float stepdist, int n_rays, float* rays_pts, uint8_t* mask_outbbox,
                               int64_t* ray_id, int64_t* step_id, int64_t* n_steps, float* t_min, float* t_max);
void orc_maskcache_lookup(const uint8_t* world, const float* xyz, uint8_t* out, const float* scale, const float* shift,
                          int sz_i, int sz_j, int sz_k, int64_t n_pts);
void orc_raw2alpha(const float* density, float shift, float interval, int64_t n, float* exp_d, float* alpha);
void orc_raw2alpha_backward(const float* exp_d, const float* grad_back, float interval, int64_t n, float* grad);
void orc_alpha2weight(const float* alpha, const int64_t* ray_id, int64_t n_pts, int n_rays, float* weight, float* T,
                      float* alphainv_last, int64_t* i_start, int64_t* i_end);
void orc_alpha2weight_backward(const float* alpha, const float* weight, const float* T, const float* alphainv_last,
                               const int64_t* i_start, const int64_t* i_end, int n_rays, const float* grad_weights,
                               const float* grad_last, float* grad);
void orc_dense_adam(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, const float* perlr, int64_t n,
                    int mode, int step, float beta1, float beta2, float lr, float eps);

/* Fine-stage training step: DirectVoxGO.forward (dvgo.py:296-388) + losses and optimiser calls (run.py:541-588). */
typedef struct orc_train_cfg {
    float xyz_min[3], xyz_max[3];
    int32_t reso[3];
    float near, far, stepdist, act_shift, interval, fast_color_thres, bg;
    float weight_main, weight_entropy_last, weight_rgbper;
    float lr_density, lr_k0, lr_net, eps, beta0, beta1;
    int32_t den_mode, k0_mode, step;   /* step >= 1: the optimiser step this call performs */
    int32_t n_rays_global;             /* N in the loss means (== n_rays unless emulating a data-parallel shard) */
    int32_t do_update;                 /* 0: stop after backward (gradients only) */
    int32_t threads;
} orc_train_cfg;

typedef struct orc_train_out {
    /* per ray [n_rays] */
    int64_t* n_steps; int32_t* cnt_inbbox; int32_t* cnt_mask; int32_t* cnt_alpha_full; int32_t* cnt_alpha; int32_t* cnt_keep;
    float* alphainv_last; float* rgb_marched /*[n_rays][3]*/;
    float loss[4];                     /* total, mse, entropy_last, rgbper */
    int64_t M0, M0_in, M1, M2, M2_trim, M3;
    /* optional per kept sample (capacity given), in ray order */
    int64_t cap_keep; int32_t* keep_ray; int32_t* keep_step; float* keep_weight; float* keep_rgb /*[.][3]*/;
    float* keep_feat /*[.][12]*/; int32_t* keep_leaf /*[.][8]*/; int32_t* keep_off /*[.][8]*/;
    float* net_grad;                   /* [22019] or NULL */
    /* distinct-voxel counts for the algorithmic-bytes model */
    int64_t V_mask, V_den, V_den_grad, V_k0;
} orc_train_out;

/* net: packed PyTorch layout w0[128][39] b0[128] w1[128][128] b1[128] w2[3][128] b2[3]; net_m/net_v Adam moments. */
void orc_train_step(const orc_train_cfg* cfg, orc_grid* den, orc_grid* den_grad, orc_grid* den_m, orc_grid* den_v,
                    orc_grid* k0, orc_grid* k0_grad, orc_grid* k0_m, orc_grid* k0_v, const uint8_t* mask /*[reso]*/,
                    float* net, float* net_m, float* net_v, const float* rays_o, const float* rays_d,
                    const float* viewdirs, const float* target, int n_rays, orc_train_out* out);

/* The same with the per-voxel lr grid of stepmode 2 (masked_adam.py:43-46).  k0 with 3 channels selects the coarse stage: no
 * rgbnet, rgb = sigmoid(k0) (dvgo.py:344-346); `net*` are not touched then. */
void orc_train_step_perlr(const orc_train_cfg* cfg, orc_grid* den, orc_grid* den_grad, orc_grid* den_m, orc_grid* den_v,
                          orc_grid* k0, orc_grid* k0_grad, orc_grid* k0_m, orc_grid* k0_v, const orc_grid* den_perlr, const uint8_t* mask /*[reso]*/,
                          float* net, float* net_m, float* net_v, const float* rays_o, const float* rays_d,
                          const float* viewdirs, const float* target, int n_rays, orc_train_out* out);

/* R1: vdb_compression.py:19-59.  Returns N (active voxels); den [N+1], col [(N+1)*cdim] rounded through fp16;
 * idx_dense [reso] 1-based float ids (0 = inactive).  Pass NULL outputs to size. */
int64_t orc_merge(const orc_grid* den, const orc_grid* k0, const uint8_t* mask, float* dendata, float* coldata,
                  float* idx_dense);

/* R2: render_an_image_cuda (renderer.cu:370-424) for rows [row_begin,row_end). idx grid: FloatGrid built from idx_dense
 * (active = non-zero).  MLP in the transposed layout of run.py:98-104. n_samples_out optional [rows*W]. */
typedef struct orc_render_cfg {
    int32_t reso[3]; float K[9]; float xyz_min[3], xyz_max[3];
    float near, stepdist, act_shift, interval, fast_color_thres, bg;
    int32_t inverse_y, H, W; int32_t threads;
} orc_render_cfg;
void orc_render(const orc_render_cfg* cfg, const orc_grid* idx_grid, const float* dendata, const float* coldata, int cdim,
                const float* w0, const float* b0, const float* w1, const float* b1, const float* w2, const float* b2,
                const float* c2w, int row_begin, int row_end, float* out_rgb, int32_t* n_samples_out,
                int32_t* inconsistent_rays);

/* Exactness check of the fast march of plenvdb_b200/csrc/renderer.cu (empty-space skipping, lane-parallel evaluation,
 * simulated second march) against the step-by-step reference marches, pixel by pixel; see the .cpp.  stats: int64[10]. */
void orc_march_check(const orc_render_cfg* cfg, const orc_grid* idx_grid, const float* dendata, const float* c2w,
                     int skip_k, int lanes, int slot, int64_t* stats);

#ifdef __cplusplus
}
#endif
#endif
