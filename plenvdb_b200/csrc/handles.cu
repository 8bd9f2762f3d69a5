// handles.cu — opaque-handle form of the B1 boundary: what a C / C++ / pybind11 binder of the reference's
// `plenvdb` module (plenvdb/lib/vdb/plenvdb.cpp:3-172, classes in plenvdb.h) would call instead of re-implementing topology
// upload, plane allocation and optimiser step counting on top of the stateless entry points.  One handle per reference class:
//
//   pvdb_grid      DensityVDB / ColorVDB   (BaseVDB, plenvdb.h:389-429): tree on the device + value plane + gradient plane
//   pvdb_opt       DensityOpt / ColorOpt   (BaseOptimizer, plenvdb.h:686-789): moments congruent with the grid, step, lr, betas
//   pvdb_renderer  MGRenderer              (plenvdb.h:933-1068): merged index tree + data rows + rgbnet + scene + scratch
//
// Host-buffer contract like the pybind11 module (numpy arrays in, numpy arrays out: every call is complete when it returns);
// the device pointers behind a handle are exposed for zero-copy callers.  Everything numeric goes through the same kernels as
// the stateless API.  `.vdb` load / save stays on the Python side of this repo (plenvdb_b200/openvdb_io.py): the handles take
// and return dense host arrays, which is also what the reference's own load path ends in (copyFromDense, plenvdb.h:149-157).
#include <stdlib.h>
#include <string.h>
#include <new>
#include <vector>
#include "common.cuh"

struct pvdb_grid {
    int reso[3];
    int channels;
    pvdb_tree tree;                      // device arrays owned below
    uint64_t* d_root_keys; int32_t *d_upper, *d_lower, *d_origin; uint64_t* d_mask;
    float *values, *grad;                // [n_leaf][512][channels]
    size_t plane_floats;
};
struct pvdb_opt {
    pvdb_grid* grid;
    float lr, eps, beta0, beta1;
    int step;
    float *exp_avg, *exp_avg_sq, *per_lr;
};
struct pvdb_renderer {
    pvdb_render_cfg cfg;
    pvdb_render_bufs bufs;
    pvdb_grid* idx;                      // index tree (values unused)
    int32_t* idx_plane;
    float *dendata, *coldata, *w[6];
    size_t n_rows;
    std::vector<void*> scratch;          // per-frame scratch, sized for the whole frame
    uint32_t* skip_bits;
    float* c2w;
    float* out;                          // [H][W][3] device
    int flags;                           // bit per setup call (plenvdb.h:1056): load_data, load_params, setScene, setKwargs, input_a_c2w
    int scratch_pixels;
};

namespace {

template <typename T>
int upload(T** dst, const T* src, size_t n) {
    *dst = nullptr;
    if (cudaMalloc(dst, (n ? n : 1) * sizeof(T)) != cudaSuccess) return PVDB_ERR_CUDA;
    if (n && cudaMemcpy(*dst, src, n * sizeof(T), cudaMemcpyHostToDevice) != cudaSuccess) return PVDB_ERR_CUDA;
    return PVDB_OK;
}
int zeros(float** dst, size_t n) {
    *dst = nullptr;
    if (cudaMalloc(dst, (n ? n : 1) * sizeof(float)) != cudaSuccess) return PVDB_ERR_CUDA;
    return cudaMemset(*dst, 0, (n ? n : 1) * sizeof(float)) == cudaSuccess ? PVDB_OK : PVDB_ERR_CUDA;
}
void grid_free(pvdb_grid* g) {
    if (!g) return;
    cudaFree(g->d_root_keys); cudaFree(g->d_upper); cudaFree(g->d_lower); cudaFree(g->d_origin); cudaFree(g->d_mask);
    cudaFree(g->values); cudaFree(g->grad);
    delete g;
}
// staged dense array on the device: [rx*ry*rz*channels] floats
struct DenseStage {
    float* d = nullptr;
    ~DenseStage() { cudaFree(d); }
};

}  // namespace

extern "C" pvdb_grid* pvdb_grid_create(int rx, int ry, int rz, int channels, const uint8_t* active) {
    if (channels != 1 && (channels <= 0 || channels % 3 != 0)) { pvdb_set_error("pvdb_grid_create: channels must be 1 or a multiple of 3"); return nullptr; }
    pvdb_topo* t = active ? pvdb_topo_create_from_mask(active, rx, ry, rz) : pvdb_topo_create_dense(rx, ry, rz);
    if (!t) return nullptr;
    int32_t nu = 0, nl = 0, nf = 0;
    pvdb_topo_counts(t, &nu, &nl, &nf);
    std::vector<uint64_t> rk(nu > 0 ? nu : 1, 0), mk((size_t)(nf > 0 ? nf : 1) * 8, 0);
    std::vector<int32_t> up((size_t)(nu > 0 ? nu : 1) * 32768, -1), lo((size_t)(nl > 0 ? nl : 1) * 4096, -1), org((size_t)(nf > 0 ? nf : 1) * 3, 0);
    pvdb_topo_export(t, rk.data(), up.data(), lo.data(), org.data(), mk.data());
    pvdb_topo_destroy(t);
    pvdb_grid* g = new (std::nothrow) pvdb_grid();
    if (!g) { pvdb_set_error("pvdb_grid_create: out of memory"); return nullptr; }
    memset(g, 0, sizeof(*g));
    g->reso[0] = rx; g->reso[1] = ry; g->reso[2] = rz; g->channels = channels;
    g->plane_floats = (size_t)(nf > 0 ? nf : 1) * 512 * channels;
    int rc = upload(&g->d_root_keys, rk.data(), rk.size()) | upload(&g->d_upper, up.data(), up.size()) | upload(&g->d_lower, lo.data(), lo.size()) |
             upload(&g->d_origin, org.data(), org.size()) | upload(&g->d_mask, mk.data(), mk.size()) | zeros(&g->values, g->plane_floats) |
             zeros(&g->grad, g->plane_floats);
    if (rc) { pvdb_set_error("pvdb_grid_create: device allocation failed (%s)", cudaGetErrorString(cudaGetLastError())); grid_free(g); return nullptr; }
    g->tree.n_upper = nu; g->tree.n_lower = nl; g->tree.n_leaf = nf; g->tree.reserved = 0;
    g->tree.root_key0 = nu ? rk[0] : 0xFFFFFFFFFFFFFFFFull;
    g->tree.root_keys = g->d_root_keys; g->tree.upper_child = g->d_upper; g->tree.lower_child = g->d_lower;
    g->tree.leaf_origin = g->d_origin; g->tree.leaf_mask = g->d_mask;
    return g;
}
extern "C" void pvdb_grid_destroy(pvdb_grid* g) { grid_free(g); }
extern "C" int pvdb_grid_info(const pvdb_grid* g, int32_t* reso3, int32_t* channels, int32_t* n_leaf) {
    PVDB_CHECK_ARG(g, "null grid");
    if (reso3) { reso3[0] = g->reso[0]; reso3[1] = g->reso[1]; reso3[2] = g->reso[2]; }
    if (channels) *channels = g->channels;
    if (n_leaf) *n_leaf = g->tree.n_leaf;
    return PVDB_OK;
}
extern "C" const pvdb_tree* pvdb_grid_tree(const pvdb_grid* g) { return g ? &g->tree : nullptr; }
extern "C" float* pvdb_grid_values(pvdb_grid* g) { return g ? g->values : nullptr; }
extern "C" float* pvdb_grid_grad(pvdb_grid* g) { return g ? g->grad : nullptr; }

static int dense_in(const pvdb_grid* g, float* plane, int channels, const float* dense_host) {
    const size_t n = (size_t)g->reso[0] * g->reso[1] * g->reso[2] * channels;
    DenseStage s;
    PVDB_CUDA(cudaMalloc(&s.d, n * sizeof(float)));
    PVDB_CUDA(cudaMemcpy(s.d, dense_host, n * sizeof(float), cudaMemcpyHostToDevice));
    if (int rc = pvdb_copy_from_dense(&g->tree, plane, channels, s.d, g->reso[0], g->reso[1], g->reso[2], nullptr)) return rc;
    PVDB_CUDA(cudaDeviceSynchronize());
    return PVDB_OK;
}
// copyFromDense (plenvdb.h:149-157, 241-250): dense_host [rx][ry][rz][channels]
extern "C" int pvdb_grid_copy_from_dense(pvdb_grid* g, const float* dense_host) {
    PVDB_CHECK_ARG(g && dense_host, "null pointer");
    return dense_in(g, g->values, g->channels, dense_host);
}
// get_dense_grid (plenvdb.h:158-167, 251-271)
extern "C" int pvdb_grid_copy_to_dense(const pvdb_grid* g, float* dense_host) {
    PVDB_CHECK_ARG(g && dense_host, "null pointer");
    const size_t n = (size_t)g->reso[0] * g->reso[1] * g->reso[2] * g->channels;
    DenseStage s;
    PVDB_CUDA(cudaMalloc(&s.d, n * sizeof(float)));
    if (int rc = pvdb_copy_to_dense(&g->tree, g->values, g->channels, s.d, g->reso[0], g->reso[1], g->reso[2], nullptr)) return rc;
    PVDB_CUDA(cudaMemcpy(dense_host, s.d, n * sizeof(float), cudaMemcpyDeviceToHost));
    return PVDB_OK;
}
// forward / backward (plenvdb.cpp:10-31, 47-67): host SoA coordinates in index space
extern "C" int pvdb_grid_forward(const pvdb_grid* g, const float* x, const float* y, const float* z, int64_t n, float* out_host) {
    PVDB_CHECK_ARG(g, "null grid");
    return pvdb_sample_forward_host(&g->tree, g->values, g->channels, x, y, z, n, out_host, nullptr);
}
extern "C" int pvdb_grid_backward(pvdb_grid* g, const float* x, const float* y, const float* z, const float* grad_host, int64_t n) {
    PVDB_CHECK_ARG(g, "null grid");
    return pvdb_sample_backward_host(&g->tree, g->grad, g->channels, x, y, z, grad_host, n, nullptr);
}
// setValuesOn_bymask (plenvdb.h:487-495): mask_host uint8 [rx*ry*rz]
extern "C" int pvdb_grid_set_values_on_by_mask(pvdb_grid* g, const uint8_t* mask_host, float val) {
    PVDB_CHECK_ARG(g && mask_host && g->channels == 1, "density grids only");
    const size_t n = (size_t)g->reso[0] * g->reso[1] * g->reso[2];
    uint8_t* d = nullptr;
    PVDB_CUDA(cudaMalloc(&d, n));
    cudaError_t e = cudaMemcpy(d, mask_host, n, cudaMemcpyHostToDevice);
    int rc = e == cudaSuccess ? pvdb_set_values_on_by_mask(&g->tree, g->values, d, val, g->reso[0], g->reso[1], g->reso[2], nullptr) : PVDB_ERR_CUDA;
    cudaDeviceSynchronize();
    cudaFree(d);
    return rc;
}

// ---- optimiser (plenvdb.h:686-789)
extern "C" pvdb_opt* pvdb_opt_create(pvdb_grid* g, float lr, float eps, float beta0, float beta1) {
    if (!g) { pvdb_set_error("pvdb_opt_create: null grid"); return nullptr; }
    pvdb_opt* o = new (std::nothrow) pvdb_opt();
    if (!o) { pvdb_set_error("pvdb_opt_create: out of memory"); return nullptr; }
    o->grid = g; o->lr = lr; o->eps = eps; o->beta0 = beta0; o->beta1 = beta1; o->step = 0; o->per_lr = nullptr;
    if (zeros(&o->exp_avg, g->plane_floats) | zeros(&o->exp_avg_sq, g->plane_floats)) {
        pvdb_set_error("pvdb_opt_create: device allocation failed");
        cudaFree(o->exp_avg); cudaFree(o->exp_avg_sq); delete o;
        return nullptr;
    }
    return o;
}
extern "C" void pvdb_opt_destroy(pvdb_opt* o) {
    if (!o) return;
    cudaFree(o->exp_avg); cudaFree(o->exp_avg_sq); cudaFree(o->per_lr);
    delete o;
}
extern "C" int pvdb_opt_zero_grad(pvdb_opt* o) {
    PVDB_CHECK_ARG(o, "null optimiser");
    if (int rc = pvdb_zero_grad(&o->grid->tree, o->grid->grad, o->grid->channels, nullptr)) return rc;
    PVDB_CUDA(cudaDeviceSynchronize());
    return PVDB_OK;
}
// step_optimizer(stepmode) (plenvdb.h:751-767, 774-789): 0 plain, 1 skip zero gradients, 2 per-voxel lr
extern "C" int pvdb_opt_step(pvdb_opt* o, int stepmode) {
    PVDB_CHECK_ARG(o && stepmode >= 0 && stepmode <= 2, "stepmode must be 0, 1 or 2");
    PVDB_CHECK_ARG(stepmode != 2 || o->per_lr, "stepmode 2 needs pvdb_opt_set_pervoxel_lr");
    o->step += 1;
    const float stepsz = pvdb_adam_stepsize(o->lr, o->beta0, o->beta1, o->step);
    if (int rc = pvdb_adam_step(&o->grid->tree, o->grid->values, o->grid->grad, o->exp_avg, o->exp_avg_sq, o->grid->channels, stepmode, stepsz,
                                o->eps, o->beta0, o->beta1, stepmode == 2 ? o->per_lr : nullptr, nullptr))
        return rc;
    PVDB_CUDA(cudaDeviceSynchronize());
    return PVDB_OK;
}
extern "C" int pvdb_opt_update_lr(pvdb_opt* o, float factor) { PVDB_CHECK_ARG(o, "null optimiser"); o->lr *= factor; return PVDB_OK; }   // plenvdb.h:712
// set_pervoxel_lr (plenvdb.h:714-722): dense_host [rx*ry*rz]
extern "C" int pvdb_opt_set_pervoxel_lr(pvdb_opt* o, const float* dense_host) {
    PVDB_CHECK_ARG(o && dense_host && o->grid->channels == 1, "density optimisers only");
    if (!o->per_lr && zeros(&o->per_lr, o->grid->plane_floats)) { pvdb_set_error("pvdb_opt_set_pervoxel_lr: device allocation failed"); return PVDB_ERR_CUDA; }
    return dense_in(o->grid, o->per_lr, 1, dense_host);
}
extern "C" int pvdb_opt_get(const pvdb_opt* o, int32_t* step, float* lr, float* eps, float* beta0, float* beta1) {
    PVDB_CHECK_ARG(o, "null optimiser");
    if (step) *step = o->step;
    if (lr) *lr = o->lr;
    if (eps) *eps = o->eps;
    if (beta0) *beta0 = o->beta0;
    if (beta1) *beta1 = o->beta1;
    return PVDB_OK;
}
extern "C" int pvdb_opt_set(pvdb_opt* o, int32_t step, float lr, float eps, float beta0, float beta1) {
    PVDB_CHECK_ARG(o, "null optimiser");
    o->step = step; o->lr = lr; o->eps = eps; o->beta0 = beta0; o->beta1 = beta1;
    return PVDB_OK;
}
extern "C" float* pvdb_opt_exp_avg(pvdb_opt* o) { return o ? o->exp_avg : nullptr; }
extern "C" float* pvdb_opt_exp_avg_sq(pvdb_opt* o) { return o ? o->exp_avg_sq : nullptr; }

// ---- merged renderer (plenvdb.h:933-1068)
extern "C" pvdb_renderer* pvdb_renderer_create(int dcol, int dpe, int dhid, int dout) {
    if (dcol != 12 || dpe != 27 || dhid != 128 || dout != 3) {
        pvdb_set_error("pvdb_renderer_create: specialised for MGRenderer(12, 27, 128, 3) (run.py:77-82)");
        return nullptr;
    }
    pvdb_renderer* r = new (std::nothrow) pvdb_renderer();
    if (!r) { pvdb_set_error("pvdb_renderer_create: out of memory"); return nullptr; }
    memset(&r->cfg, 0, sizeof(r->cfg));
    memset(&r->bufs, 0, sizeof(r->bufs));
    r->cfg.dcol = dcol; r->cfg.dpe = dpe; r->cfg.dhid = dhid; r->cfg.dout = dout; r->cfg.use_tensor_cores = 1;
    r->idx = nullptr; r->idx_plane = nullptr; r->dendata = r->coldata = nullptr;
    for (auto& p : r->w) p = nullptr;
    r->n_rows = 0; r->skip_bits = nullptr; r->c2w = nullptr; r->out = nullptr; r->flags = 0; r->scratch_pixels = 0;
    if (cudaMalloc(&r->c2w, 16 * sizeof(float)) != cudaSuccess) { pvdb_set_error("pvdb_renderer_create: device allocation failed"); delete r; return nullptr; }
    return r;
}
static void renderer_free_scratch(pvdb_renderer* r) {
    for (void* p : r->scratch) cudaFree(p);
    r->scratch.clear();
    cudaFree(r->out); r->out = nullptr;
    r->scratch_pixels = 0;
}
extern "C" void pvdb_renderer_destroy(pvdb_renderer* r) {
    if (!r) return;
    renderer_free_scratch(r);
    grid_free(r->idx);
    cudaFree(r->idx_plane); cudaFree(r->dendata); cudaFree(r->coldata); cudaFree(r->skip_bits); cudaFree(r->c2w);
    for (auto p : r->w) cudaFree(p);
    delete r;
}
// load_data (plenvdb.h:959-983): den [N], col [N][12] rows (row 0 = zeros), idx_dense_host int32 [rx][ry][rz] 1-based row ids
extern "C" int pvdb_renderer_load_data(pvdb_renderer* r, const float* den_host, const float* col_host, int64_t n_rows, const int32_t* idx_dense_host,
                                       int rx, int ry, int rz) {
    PVDB_CHECK_ARG(r && den_host && col_host && idx_dense_host && n_rows > 0, "bad arguments");
    const size_t nvox = (size_t)rx * ry * rz;
    std::vector<uint8_t> active(nvox);
    std::vector<float> idx_f(nvox);
    for (size_t i = 0; i < nvox; ++i) { active[i] = idx_dense_host[i] != 0; idx_f[i] = (float)idx_dense_host[i]; }   // copyFromArray: active <=> value != background
    grid_free(r->idx);
    r->idx = pvdb_grid_create(rx, ry, rz, 1, active.data());
    if (!r->idx) return PVDB_ERR_ARG;
    if (int rc = dense_in(r->idx, r->idx->values, 1, idx_f.data())) return rc;
    // int(acc.getValue()) (renderer.cu:202-209): the float plane as int32
    std::vector<float> plane(r->idx->plane_floats);
    PVDB_CUDA(cudaMemcpy(plane.data(), r->idx->values, plane.size() * sizeof(float), cudaMemcpyDeviceToHost));
    std::vector<int32_t> iplane(plane.size());
    for (size_t i = 0; i < plane.size(); ++i) iplane[i] = (int32_t)plane[i];
    cudaFree(r->idx_plane); cudaFree(r->dendata); cudaFree(r->coldata); cudaFree(r->skip_bits);
    r->skip_bits = nullptr;
    if (upload(&r->idx_plane, iplane.data(), iplane.size()) | upload(&r->dendata, den_host, (size_t)n_rows) |
        upload(&r->coldata, col_host, (size_t)n_rows * 12)) { pvdb_set_error("pvdb_renderer_load_data: device allocation failed"); return PVDB_ERR_CUDA; }
    r->n_rows = (size_t)n_rows;
    r->flags |= 1;
    r->scratch_pixels = 0;     // the buffer struct is rebuilt by the next render
    return PVDB_OK;
}
// load_params (plenvdb.h:985-996): transposed weights as run.py:98-104 passes them
extern "C" int pvdb_renderer_load_params(pvdb_renderer* r, const float* w0, const float* b0, const float* w1, const float* b1, const float* w2, const float* b2) {
    PVDB_CHECK_ARG(r && w0 && b0 && w1 && b1 && w2 && b2, "null pointer");
    const float* src[6] = {w0, b0, w1, b1, w2, b2};
    const size_t n[6] = {39 * 128, 128, 128 * 128, 128, 128 * 3, 3};
    for (int i = 0; i < 6; ++i) {
        cudaFree(r->w[i]);
        if (upload(&r->w[i], src[i], n[i])) { pvdb_set_error("pvdb_renderer_load_params: device allocation failed"); return PVDB_ERR_CUDA; }
    }
    r->flags |= 2;
    r->scratch_pixels = 0;
    return PVDB_OK;
}
extern "C" int pvdb_renderer_set_scene(pvdb_renderer* r, const int32_t* reso3, const float* K9, const float* xyz_min3, const float* xyz_max3) {
    PVDB_CHECK_ARG(r && reso3 && K9 && xyz_min3 && xyz_max3, "null pointer");
    for (int a = 0; a < 3; ++a) { r->cfg.reso[a] = reso3[a]; r->cfg.xyz_min[a] = xyz_min3[a]; r->cfg.xyz_max[a] = xyz_max3[a]; }
    for (int i = 0; i < 9; ++i) r->cfg.K[i] = K9[i];
    cudaFree(r->skip_bits); r->skip_bits = nullptr;
    r->flags |= 4;
    return PVDB_OK;
}
extern "C" int pvdb_renderer_set_kwargs(pvdb_renderer* r, float near, float far, float stepdist, float act_shift, float interval, float fast_color_thres,
                                        float bg, int inverse_y, int H, int W) {
    PVDB_CHECK_ARG(r && H > 0 && W > 0, "bad arguments");
    (void)far;                                            // ignored like plenvdb.h:1008
    r->cfg.near = near; r->cfg.far = 1e9f; r->cfg.stepdist = stepdist; r->cfg.act_shift = act_shift; r->cfg.interval = interval;
    r->cfg.fast_color_thres = fast_color_thres; r->cfg.bg = bg; r->cfg.inverse_y = inverse_y ? 1 : 0; r->cfg.H = H; r->cfg.W = W;
    r->flags |= 8;
    return PVDB_OK;
}
extern "C" int pvdb_renderer_input_c2w(pvdb_renderer* r, const float* c2w16_host) {
    PVDB_CHECK_ARG(r && c2w16_host, "null pointer");
    PVDB_CUDA(cudaMemcpy(r->c2w, c2w16_host, 16 * sizeof(float), cudaMemcpyHostToDevice));
    r->flags |= 16;
    return PVDB_OK;
}
static int renderer_scratch(pvdb_renderer* r) {
    const int npix = r->cfg.H * r->cfg.W;
    if (r->scratch_pixels == npix && r->skip_bits) return PVDB_OK;
    renderer_free_scratch(r);
    const int64_t cap = (int64_t)npix * 6 > 4096 ? (int64_t)npix * 6 : 4096;
    const int P = 64;
    auto alloc = [&](size_t bytes) -> void* {
        void* p = nullptr;
        if (cudaMalloc(&p, bytes ? bytes : 16) != cudaSuccess) return nullptr;
        cudaMemset(p, 0, bytes ? bytes : 16);
        r->scratch.push_back(p);
        return p;
    };
    pvdb_render_bufs& b = r->bufs;
    memset(&b, 0, sizeof(b));
    b.idx_tree = &r->idx->tree; b.idx_plane = r->idx_plane; b.dendata = r->dendata; b.coldata = r->coldata;
    b.w0 = r->w[0]; b.b0 = r->w[1]; b.w1 = r->w[2]; b.b1 = r->w[3]; b.w2 = r->w[4]; b.b2 = r->w[5];
    b.n_samples = (int32_t*)alloc((size_t)npix * 4); b.i_starts = (int32_t*)alloc((size_t)(npix + 1) * 4);
    b.tmins = (float*)alloc((size_t)npix * 4); b.tmaxs = (float*)alloc((size_t)npix * 4);
    b.scan_tmp = (int32_t*)alloc((size_t)(npix / 4096 + 3) * 4);
    b.cap_samples = cap;
    b.s_ray = (int32_t*)alloc((size_t)cap * 4); b.s_weight = (float*)alloc((size_t)cap * 4);
    b.s_feat = (float*)alloc((size_t)cap * 48); b.s_rgb = (float*)alloc((size_t)cap * 12);
    b.counters = (int32_t*)alloc(8 * 4); b.w_img = alloc(256 * 1024);
    b.active_list = (int32_t*)alloc((size_t)npix * 4);
    b.px_scratch = alloc((size_t)npix * P * 8); b.fallback_list = (int32_t*)alloc((size_t)npix * 4); b.px_entries = P;
    cudaFree(r->skip_bits);
    r->skip_bits = nullptr;
    const size_t words = pvdb_render_block_bits_words(r->cfg.reso[0], r->cfg.reso[1], r->cfg.reso[2]);
    if (cudaMalloc(&r->skip_bits, (words ? words : 1) * 4) != cudaSuccess) { pvdb_set_error("pvdb_renderer: device allocation failed"); return PVDB_ERR_CUDA; }
    if (int rc = pvdb_render_block_bits(&r->idx->tree, r->cfg.reso[0], r->cfg.reso[1], r->cfg.reso[2], r->skip_bits, nullptr)) return rc;
    b.skip_bits = r->skip_bits;
    if (cudaMalloc(&r->out, (size_t)npix * 12) != cudaSuccess) { pvdb_set_error("pvdb_renderer: device allocation failed"); return PVDB_ERR_CUDA; }
    for (void* p : r->scratch)
        if (!p) { pvdb_set_error("pvdb_renderer: device allocation failed"); return PVDB_ERR_CUDA; }
    r->scratch_pixels = npix;
    return PVDB_OK;
}
// render_an_image + output_an_image (plenvdb.h:1027-1046): silently does nothing until all five setup calls happened
// (returns PVDB_OK with *rendered = 0); out_host float [H][W][3]
extern "C" int pvdb_renderer_render(pvdb_renderer* r, float* out_host, int* rendered) {
    PVDB_CHECK_ARG(r, "null renderer");
    if (rendered) *rendered = 0;
    if (r->flags != 31) return PVDB_OK;
    if (int rc = renderer_scratch(r)) return rc;
    if (int rc = pvdb_render_rows(&r->cfg, &r->bufs, r->c2w, 0, r->cfg.H, r->out, nullptr)) return rc;
    if (out_host) PVDB_CUDA(cudaMemcpy(out_host, r->out, (size_t)r->cfg.H * r->cfg.W * 12, cudaMemcpyDeviceToHost));
    else PVDB_CUDA(cudaDeviceSynchronize());
    if (rendered) *rendered = 1;
    return PVDB_OK;
}
extern "C" const float* pvdb_renderer_frame(const pvdb_renderer* r) { return r ? r->out : nullptr; }
extern "C" int pvdb_renderer_counters(const pvdb_renderer* r, int32_t* counters8) {
    PVDB_CHECK_ARG(r && counters8 && r->bufs.counters, "no frame rendered yet");
    PVDB_CUDA(cudaMemcpy(counters8, r->bufs.counters, 8 * 4, cudaMemcpyDeviceToHost));
    return PVDB_OK;
}
