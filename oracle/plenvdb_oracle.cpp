// plenvdb_oracle.cpp — CPU oracle for the PlenVDB hot path.  TEST INFRASTRUCTURE ONLY (see plenvdb_oracle.h).
//
// A from-scratch restatement of the reference's algorithm in plain C++ with no dependency on the
// reference tree.  Float expressions are written with explicit fmaf() exactly where nvcc contracts the
// reference's CUDA source into FMA (read from the PTX of the reference compiled for sm_100a), and the
// file is compiled with -ffp-contract=off so nothing else fuses.  Citations are relative to the
// reference root (wolfball/PlenVDB).
#include "plenvdb_oracle.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <thread>
#include <unordered_set>
#include <vector>

// ------------------------------------------------------------------------------------------------
// T1/T2 — tree.  Restates the NanoVDB 32.3.3 5-4-3 tree semantics the reference relies on
// (openvdb/nanovdb/nanovdb/NanoVDB.h): root keyed tiles :2702-2709, InternalNode CoordToOffset
// :3377-3385, LeafNode CoordToOffset :3893-3900, Mask bit order :1902, getValue falling back to the
// background (0 for every PlenVDB grid) when no leaf covers the coordinate :2998-3010, :3327-3335.
// Node order = OpenToNanoVDB.h:521-533 (depth-first by increasing child offset).
// ------------------------------------------------------------------------------------------------
struct orc_grid {
    int rx, ry, rz, C;
    std::vector<uint64_t> root_keys;
    std::vector<int32_t> upper_child;   // [n_upper][32768]
    std::vector<int32_t> lower_child;   // [n_lower][4096]
    std::vector<int32_t> origin;        // [n_leaf][3]
    std::vector<uint64_t> mask;         // [n_leaf][8]
    std::vector<float> values;          // [n_leaf][512][C]  (channel c = Vec3f grid c/3, component c%3)
    int n_leaf() const { return (int)(origin.size() / 3); }
};

static inline uint64_t root_key(int x, int y, int z) {
    return (uint64_t)((uint32_t)z >> 12) | ((uint64_t)((uint32_t)y >> 12) << 21) | ((uint64_t)((uint32_t)x >> 12) << 42);
}
static inline int upper_off(int x, int y, int z) { return (((x & 4095) >> 7) << 10) | (((y & 4095) >> 7) << 5) | ((z & 4095) >> 7); }
static inline int lower_off(int x, int y, int z) { return (((x & 127) >> 3) << 8) | (((y & 127) >> 3) << 4) | ((z & 127) >> 3); }
static inline int leaf_off(int x, int y, int z) { return ((x & 7) << 6) | ((y & 7) << 3) | (z & 7); }

static int find_leaf(const orc_grid* g, int x, int y, int z) {
    const uint64_t key = root_key(x, y, z);
    int u = -1;
    for (size_t i = 0; i < g->root_keys.size(); ++i)   // findTile linear search, NanoVDB.h:2924-2948
        if (g->root_keys[i] == key) { u = (int)i; break; }
    if (u < 0) return -1;
    const int l = g->upper_child[(size_t)u * 32768 + upper_off(x, y, z)];
    if (l < 0) return -1;
    return g->lower_child[(size_t)l * 4096 + lower_off(x, y, z)];
}
static inline bool leaf_active(const orc_grid* g, int leaf, int off) { return (g->mask[(size_t)leaf * 8 + (off >> 6)] >> (off & 63)) & 1ull; }

extern "C" orc_grid* orc_grid_create(int rx, int ry, int rz, int channels, const uint8_t* active) {
    orc_grid* g = new orc_grid();
    g->rx = rx; g->ry = ry; g->rz = rz; g->C = channels;
    const int nbx = (rx + 7) / 8, nby = (ry + 7) / 8, nbz = (rz + 7) / 8;
    const int nlx = (rx + 127) / 128, nly = (ry + 127) / 128, nlz = (rz + 127) / 128;
    const int nux = (rx + 4095) / 4096, nuy = (ry + 4095) / 4096, nuz = (rz + 4095) / 4096;
    for (int ux = 0; ux < nux; ++ux) for (int uy = 0; uy < nuy; ++uy) for (int uz = 0; uz < nuz; ++uz) {
        int upper = -1;
        for (int lx = ux * 32; lx < std::min(nlx, ux * 32 + 32); ++lx)
        for (int ly = uy * 32; ly < std::min(nly, uy * 32 + 32); ++ly)
        for (int lz = uz * 32; lz < std::min(nlz, uz * 32 + 32); ++lz) {
            int lower = -1;
            for (int bx = lx * 16; bx < std::min(nbx, lx * 16 + 16); ++bx)
            for (int by = ly * 16; by < std::min(nby, ly * 16 + 16); ++by)
            for (int bz = lz * 16; bz < std::min(nbz, lz * 16 + 16); ++bz) {
                uint64_t w[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                bool any = false;
                for (int dx = 0; dx < 8; ++dx) for (int dy = 0; dy < 8; ++dy) for (int dz = 0; dz < 8; ++dz) {
                    const int x = bx * 8 + dx, y = by * 8 + dy, z = bz * 8 + dz;
                    if (x >= rx || y >= ry || z >= rz) continue;
                    // denseFill(bbox, 0, active=true) activates every voxel of the box and allocates every
                    // leaf it touches (plenvdb.h:117-125); with a mask, only voxels set in it.
                    if (active && !active[((size_t)x * ry + y) * rz + z]) continue;
                    const int n = leaf_off(dx, dy, dz);
                    w[n >> 6] |= 1ull << (n & 63);
                    any = true;
                }
                if (!any) continue;
                if (upper < 0) {
                    upper = (int)g->root_keys.size();
                    g->root_keys.push_back(root_key(ux * 4096, uy * 4096, uz * 4096));
                    g->upper_child.resize((size_t)(upper + 1) * 32768, -1);
                }
                if (lower < 0) {
                    lower = (int)(g->lower_child.size() / 4096);
                    g->lower_child.resize((size_t)(lower + 1) * 4096, -1);
                    g->upper_child[(size_t)upper * 32768 + upper_off(lx * 128, ly * 128, lz * 128)] = lower;
                }
                const int leaf = g->n_leaf();
                g->lower_child[(size_t)lower * 4096 + lower_off(bx * 8, by * 8, bz * 8)] = leaf;
                g->origin.push_back(bx * 8); g->origin.push_back(by * 8); g->origin.push_back(bz * 8);
                g->mask.insert(g->mask.end(), w, w + 8);
            }
        }
    }
    g->values.assign((size_t)g->n_leaf() * 512 * channels, 0.f);
    return g;
}
extern "C" void orc_grid_destroy(orc_grid* g) { delete g; }
extern "C" int orc_grid_leaf_count(const orc_grid* g) { return g->n_leaf(); }
extern "C" void orc_grid_leaf_origins(const orc_grid* g, int32_t* out) { std::memcpy(out, g->origin.data(), g->origin.size() * 4); }
extern "C" void orc_grid_leaf_masks(const orc_grid* g, uint64_t* out) { std::memcpy(out, g->mask.data(), g->mask.size() * 8); }
extern "C" void orc_grid_fill(orc_grid* g, float v) { std::fill(g->values.begin(), g->values.end(), v); }
// Leaf-major payload [n_leaf][512][C] in and out (test helper for grids too large for a dense host array: the caller
// checks orc_grid_leaf_origins against its own leaf order first).
extern "C" void orc_grid_set_values(orc_grid* g, const float* plane) { std::memcpy(g->values.data(), plane, g->values.size() * 4); }
extern "C" void orc_grid_get_values(const orc_grid* g, float* plane) { std::memcpy(plane, g->values.data(), g->values.size() * 4); }

// densityvdb.cu:31-49 / colorvdb.cu:42-63 — every ACTIVE voxel slot takes the dense value at its coordinate
extern "C" void orc_grid_copy_from_dense(orc_grid* g, const float* dense) {
    const int C = g->C;
    for (int leaf = 0; leaf < g->n_leaf(); ++leaf)
        for (int off = 0; off < 512; ++off) {
            if (!leaf_active(g, leaf, off)) continue;
            const int x = g->origin[leaf * 3] + (off >> 6), y = g->origin[leaf * 3 + 1] + ((off >> 3) & 7), z = g->origin[leaf * 3 + 2] + (off & 7);
            for (int c = 0; c < C; ++c)
                g->values[((size_t)leaf * 512 + off) * C + c] = dense[(((size_t)x * g->ry + y) * g->rz + z) * C + c];
        }
}
// plenvdb.h:158-167 via tools::copyToDense: the tree's value at every coordinate of the box
extern "C" void orc_grid_copy_to_dense(const orc_grid* g, float* dense) {
    const int C = g->C;
    for (int x = 0; x < g->rx; ++x) for (int y = 0; y < g->ry; ++y) for (int z = 0; z < g->rz; ++z) {
        const int leaf = find_leaf(g, x, y, z);
        for (int c = 0; c < C; ++c)
            dense[(((size_t)x * g->ry + y) * g->rz + z) * C + c] = leaf >= 0 ? g->values[((size_t)leaf * 512 + leaf_off(x, y, z)) * C + c] : 0.f;
    }
}
// densityvdb.cu:64-84
extern "C" void orc_grid_set_on_by_mask(orc_grid* g, const uint8_t* m, float val) {
    for (int leaf = 0; leaf < g->n_leaf(); ++leaf)
        for (int off = 0; off < 512; ++off) {
            if (!leaf_active(g, leaf, off)) continue;
            const int x = g->origin[leaf * 3] + (off >> 6), y = g->origin[leaf * 3 + 1] + ((off >> 3) & 7), z = g->origin[leaf * 3 + 2] + (off & 7);
            if (m[((size_t)x * g->ry + y) * g->rz + z]) g->values[(size_t)leaf * 512 + off] = val;
        }
}

// ------------------------------------------------------------------------------------------------
// helpers
// ------------------------------------------------------------------------------------------------
static void parallel_for(int64_t n, int threads, const std::function<void(int64_t, int64_t)>& fn) {
    if (threads <= 1 || n < 1024) { fn(0, n); return; }
    std::vector<std::thread> ts;
    const int64_t chunk = (n + threads - 1) / threads;
    for (int t = 0; t < threads; ++t) {
        const int64_t b = t * chunk, e = std::min(n, b + chunk);
        if (b >= e) break;
        ts.emplace_back(fn, b, e);
    }
    for (auto& t : ts) t.join();
}
static inline void atomic_add_float(float* addr, float v) {   // CAS float atomic add for the threaded backward
    auto* a = reinterpret_cast<std::atomic<uint32_t>*>(addr);
    uint32_t old = a->load(std::memory_order_relaxed), neu;
    do {
        float f; std::memcpy(&f, &old, 4); f += v; std::memcpy(&neu, &f, 4);
    } while (!a->compare_exchange_weak(old, neu, std::memory_order_relaxed));
}

// Corner walk of densityvdb.cu:116-123: 000,001,011,010,110,111,101,100
static const int CORNER[8][3] = {{0, 0, 0}, {0, 0, 1}, {0, 1, 1}, {0, 1, 0}, {1, 1, 0}, {1, 1, 1}, {1, 0, 1}, {1, 0, 0}};

struct Tri {
    int i, j, k; float u[3], m[3];
    Tri(float x, float y, float z) {
        i = (int)floorf(x); j = (int)floorf(y); k = (int)floorf(z);      // Vec3::floor -> Coord, NanoVDB Floor
        u[0] = x - (float)i; u[1] = y - (float)j; u[2] = z - (float)k;   // uvw = xyz - ijk.asVec3s()
        m[0] = 1.0f - u[0]; m[1] = 1.0f - u[1]; m[2] = 1.0f - u[2];
    }
    float f(int axis, int d) const { return d ? u[axis] : m[axis]; }
};

// D1 (densityvdb.cu:101-125): gpuRes += v*f0*f1*f2 compiles to fma(f2, f1*(f0*v), res).
// C1 (colorvdb.cu:81-111, 16-26): scale = f0*f1*f2; res[c] = fma(scale, v[c], res[c]).
extern "C" void orc_sample_forward(const orc_grid* g, const float* xs, const float* ys, const float* zs, int64_t n, float* out,
                                   int32_t* corner_leaf, int32_t* corner_off, int threads) {
    const int C = g->C;
    parallel_for(n, threads, [&](int64_t b, int64_t e) {
        for (int64_t s = b; s < e; ++s) {
            Tri t(xs[s], ys[s], zs[s]);
            float acc[64];
            for (int c = 0; c < C; ++c) acc[c] = 0.f;
            for (int q = 0; q < 8; ++q) {
                const int x = t.i + CORNER[q][0], y = t.j + CORNER[q][1], z = t.k + CORNER[q][2];
                const int leaf = find_leaf(g, x, y, z), off = leaf_off(x, y, z);
                if (corner_leaf) corner_leaf[s * 8 + q] = leaf;
                if (corner_off) corner_off[s * 8 + q] = off;
                const float f0 = t.f(0, CORNER[q][0]), f1 = t.f(1, CORNER[q][1]), f2 = t.f(2, CORNER[q][2]);
                if (C == 1) {
                    const float v = leaf >= 0 ? g->values[(size_t)leaf * 512 + off] : 0.f;
                    acc[0] = fmaf(f2, f1 * (f0 * v), acc[0]);
                } else {
                    const float sc = (f0 * f1) * f2;
                    for (int c = 0; c < C; ++c) {
                        const float v = leaf >= 0 ? g->values[((size_t)leaf * 512 + off) * C + c] : 0.f;
                        acc[c] = fmaf(sc, v, acc[c]);
                    }
                }
            }
            for (int c = 0; c < C; ++c) out[s * C + c] = acc[c];
        }
    });
}

// D2 (densityvdb.cu:143-167, accumulate :16-27): add iff a leaf covers the corner; value g*f0*f1*f2 (muls only).
// C2 (colorvdb.cu:130-160, :28-37): scale = f0*f1*f2; add g[c]*scale.
extern "C" void orc_sample_backward(orc_grid* g, const float* xs, const float* ys, const float* zs, const float* gout, int64_t n,
                                    int threads) {
    const int C = g->C;
    const bool mt = threads > 1;
    parallel_for(n, threads, [&](int64_t b, int64_t e) {
        for (int64_t s = b; s < e; ++s) {
            Tri t(xs[s], ys[s], zs[s]);
            for (int q = 0; q < 8; ++q) {
                const int x = t.i + CORNER[q][0], y = t.j + CORNER[q][1], z = t.k + CORNER[q][2];
                const int leaf = find_leaf(g, x, y, z);
                if (leaf < 0) continue;
                const int off = leaf_off(x, y, z);
                const float f0 = t.f(0, CORNER[q][0]), f1 = t.f(1, CORNER[q][1]), f2 = t.f(2, CORNER[q][2]);
                if (C == 1) {
                    const float v = ((gout[s] * f0) * f1) * f2;
                    float* p = &g->values[(size_t)leaf * 512 + off];
                    if (mt) atomic_add_float(p, v); else *p += v;
                } else {
                    const float sc = (f0 * f1) * f2;
                    for (int c = 0; c < C; ++c) {
                        float* p = &g->values[((size_t)leaf * 512 + off) * C + c];
                        const float v = gout[s * C + c] * sc;
                        if (mt) atomic_add_float(p, v); else *p += v;
                    }
                }
            }
        }
    });
}

// plenvdb.h:753 / :776 — float arithmetic on the host
extern "C" float orc_adam_stepsize(float lr, float beta0, float beta1, int step) {
    return lr * std::sqrt(1 - std::pow(beta1, (float)step)) / (1 - std::pow(beta0, (float)step));
}

// O1 (densityvdb.cu:185-330, colorvdb.cu:179-331) as compiled:
//   m' = fma(1-b0, g, b0*m); v' = fma(g, (1-b1)*g, b1*v); p' = p - (stepsz*m')/(eps+sqrt(v'))
//   mode 2: p' = p - (m'*(perlr*stepsz))/(eps+sqrt(v')).  mode 1 skips g==0 (Vec3: all three ==0, colorvdb.cu:318).
extern "C" void orc_adam_step(orc_grid* p, const orc_grid* g, orc_grid* m, orc_grid* v, int mode, float stepsz, float eps,
                              float b0, float b1, const orc_grid* perlr) {
    const int C = p->C, G = (C == 1) ? 1 : 3;
    const float omb0 = 1.0f - b0, omb1 = 1.0f - b1;
    for (int leaf = 0; leaf < p->n_leaf(); ++leaf)
        for (int off = 0; off < 512; ++off) {
            if (!leaf_active(p, leaf, off)) continue;
            const size_t vox = (size_t)leaf * 512 + off;
            for (int grp = 0; grp < C / G; ++grp) {
                const size_t base = vox * C + (size_t)grp * G;
                bool allzero = true;
                for (int c = 0; c < G; ++c) allzero = allzero && g->values[base + c] == 0.0f;
                if (mode == 1 && allzero) continue;
                for (int c = 0; c < G; ++c) {
                    const float gg = g->values[base + c];
                    const float nm = fmaf(omb0, gg, b0 * m->values[base + c]);
                    const float nv = fmaf(gg, omb1 * gg, b1 * v->values[base + c]);
                    m->values[base + c] = nm;
                    v->values[base + c] = nv;
                    const float num = (mode == 2) ? nm * (perlr->values[vox] * stepsz) : stepsz * nm;
                    p->values[base + c] = p->values[base + c] - num / (eps + sqrtf(nv));
                }
            }
        }
}
// O2 (densityvdb.cu:353-363): active voxels only
extern "C" void orc_zero_grad(orc_grid* g) {
    const int C = g->C;
    for (int leaf = 0; leaf < g->n_leaf(); ++leaf)
        for (int off = 0; off < 512; ++off)
            if (leaf_active(g, leaf, off))
                for (int c = 0; c < C; ++c) g->values[((size_t)leaf * 512 + off) * C + c] = 0.f;
}

// ------------------------------------------------------------------------------------------------
// B2 — render_utils ops (plenvdb/lib/cuda/render_utils_kernel.cu)
// ------------------------------------------------------------------------------------------------
static inline void t_minmax(const float* o, const float* d, const float* mn, const float* mx, float near, float far, float& tmin, float& tmax) {
    // :12-35
    const float vx = d[0] == 0.f ? 1e-6f : d[0], vy = d[1] == 0.f ? 1e-6f : d[1], vz = d[2] == 0.f ? 1e-6f : d[2];
    const float ax = (mx[0] - o[0]) / vx, ay = (mx[1] - o[1]) / vy, az = (mx[2] - o[2]) / vz;
    const float bx = (mn[0] - o[0]) / vx, by = (mn[1] - o[1]) / vy, bz = (mn[2] - o[2]) / vz;
    tmin = fmaxf(fminf(fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fminf(az, bz)), far), near);
    tmax = fmaxf(fminf(fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fmaxf(az, bz)), far), near);
}
static inline float ray_norm(const float* d) {   // x*x+y*y+z*z -> fma(z,z, fma(x,x, y*y)) (:46-49)
    return sqrtf(fmaf(d[2], d[2], fmaf(d[0], d[0], d[1] * d[1])));
}
static inline int64_t n_samples_of(const float* d, float tmin, float tmax, float stepdist) {   // :38-55
    const float v = ceilf((ray_norm(d) * (tmax - tmin)) / stepdist);
    return (int64_t)std::fmax((double)v, 1.0);
}
static inline void start_dir(const float* o, const float* d, float tmin, float* s, float* dir) {   // :58-79
    const float rn = ray_norm(d);
    for (int a = 0; a < 3; ++a) { s[a] = fmaf(d[a], tmin, o[a]); dir[a] = d[a] / rn; }
}
static inline void ray_point(const float* s, const float* dir, float stepdist, int step, float* p) {   // :181-187
    const float dist = stepdist * (float)step;
    for (int a = 0; a < 3; ++a) p[a] = fmaf(dir[a], dist, s[a]);
}
static inline int mask_ijk(float x, float scale, float shift) { return (int)roundf(fmaf(x, scale, shift)); }   // :385-387
static inline float raw2alpha1(float d, float shift, float interval, float& e) {   // :431-443
    e = expf(d + shift);
    return 1.0f - powf(1.0f + e, -interval);
}
static inline float raw2alpha_bwd1(float e, float gb, float interval) {   // :507-517, double product chain
    const double m = std::fmin((double)e, 1e10);
    const float pw = powf(1.0f + e, -interval - 1.0f);
    return (float)(((m * (double)pw) * (double)interval) * (double)gb);
}
static inline float T_update(float T, float a) { return (float)((1.0 - (double)a) * (double)T); }   // :596
static inline float a2w_grad(float gw, float T, float back_cum, float a) {   // :673
    return (float)((double)(gw * T) - (double)back_cum / ((double)(1.0f - a) + 1e-10));
}

extern "C" void orc_infer_t_minmax(const float* ro, const float* rd, const float* mn, const float* mx, float near, float far,
                                   int n, float* tmin, float* tmax) {
    for (int r = 0; r < n; ++r) t_minmax(ro + r * 3, rd + r * 3, mn, mx, near, far, tmin[r], tmax[r]);
}
extern "C" void orc_infer_n_samples(const float* rd, const float* tmin, const float* tmax, float stepdist, int n, int64_t* out) {
    for (int r = 0; r < n; ++r) out[r] = n_samples_of(rd + r * 3, tmin[r], tmax[r], stepdist);
}
extern "C" void orc_infer_ray_start_dir(const float* ro, const float* rd, const float* tmin, int n, float* rs, float* rdir) {
    for (int r = 0; r < n; ++r) start_dir(ro + r * 3, rd + r * 3, tmin[r], rs + r * 3, rdir + r * 3);
}
// :196-242
extern "C" int64_t orc_sample_pts_on_rays(const float* ro, const float* rd, const float* mn, const float* mx, float near,
                                          float far, float stepdist, int n_rays, float* rays_pts, uint8_t* mask_outbbox,
                                          int64_t* ray_id, int64_t* step_id, int64_t* n_steps, float* t_min, float* t_max) {
    int64_t total = 0;
    for (int r = 0; r < n_rays; ++r) {
        float tmin, tmax, s[3], dir[3];
        t_minmax(ro + r * 3, rd + r * 3, mn, mx, near, far, tmin, tmax);
        const int64_t ns = n_samples_of(rd + r * 3, tmin, tmax, stepdist);
        if (n_steps) n_steps[r] = ns;
        if (t_min) t_min[r] = tmin;
        if (t_max) t_max[r] = tmax;
        if (rays_pts) {
            start_dir(ro + r * 3, rd + r * 3, tmin, s, dir);
            for (int64_t k = 0; k < ns; ++k) {
                float p[3];
                ray_point(s, dir, stepdist, (int)k, p);
                const int64_t i = total + k;
                rays_pts[i * 3] = p[0]; rays_pts[i * 3 + 1] = p[1]; rays_pts[i * 3 + 2] = p[2];
                mask_outbbox[i] = (mn[0] > p[0]) | (mn[1] > p[1]) | (mn[2] > p[2]) | (mx[0] < p[0]) | (mx[1] < p[1]) | (mx[2] < p[2]);
                ray_id[i] = r;
                step_id[i] = k;
            }
        }
        total += ns;
    }
    return total;
}
// :374-392
extern "C" void orc_maskcache_lookup(const uint8_t* world, const float* xyz, uint8_t* out, const float* scale, const float* shift,
                                     int si, int sj, int sk, int64_t n) {
    for (int64_t p = 0; p < n; ++p) {
        const int i = mask_ijk(xyz[p * 3], scale[0], shift[0]), j = mask_ijk(xyz[p * 3 + 1], scale[1], shift[1]),
                  k = mask_ijk(xyz[p * 3 + 2], scale[2], shift[2]);
        out[p] = (0 <= i && i < si && 0 <= j && j < sj && 0 <= k && k < sk) ? world[((size_t)i * sj + j) * sk + k] : 0;
    }
}
extern "C" void orc_raw2alpha(const float* d, float shift, float interval, int64_t n, float* exp_d, float* alpha) {
    for (int64_t i = 0; i < n; ++i) alpha[i] = raw2alpha1(d[i], shift, interval, exp_d[i]);
}
extern "C" void orc_raw2alpha_backward(const float* exp_d, const float* gb, float interval, int64_t n, float* grad) {
    for (int64_t i = 0; i < n; ++i) grad[i] = raw2alpha_bwd1(exp_d[i], gb[i], interval);
}
// :577-651 — outputs must be pre-initialised by the caller like the reference's host wrapper (:625-629)
extern "C" void orc_alpha2weight(const float* alpha, const int64_t* ray_id, int64_t n_pts, int n_rays, float* weight, float* T,
                                 float* alphainv_last, int64_t* i_start, int64_t* i_end) {
    if (n_pts == 0) return;
    for (int64_t idx = 1; idx < n_pts; ++idx)
        if (ray_id[idx] != ray_id[idx - 1]) { i_start[ray_id[idx]] = idx; i_end[ray_id[idx - 1]] = idx; }
    i_end[ray_id[n_pts - 1]] = n_pts;
    for (int r = 0; r < n_rays; ++r) {
        const int64_t i_s = i_start[r], i_e_max = i_end[r];
        float T_cum = 1.f;
        int64_t i;
        for (i = i_s; i < i_e_max; ++i) {
            T[i] = T_cum;
            weight[i] = T_cum * alpha[i];
            T_cum = T_update(T_cum, alpha[i]);
            if ((double)T_cum < 1e-3) { i += 1; break; }
        }
        i_end[r] = i;
        alphainv_last[r] = T_cum;
    }
}
// :654-677
extern "C" void orc_alpha2weight_backward(const float* alpha, const float* weight, const float* T, const float* ail,
                                          const int64_t* i_start, const int64_t* i_end, int n_rays, const float* gw,
                                          const float* glast, float* grad) {
    for (int r = 0; r < n_rays; ++r) {
        float back_cum = glast[r] * ail[r];
        for (int64_t i = i_end[r] - 1; i >= i_start[r]; --i) {
            grad[i] = a2w_grad(gw[i], T[i], back_cum, alpha[i]);
            back_cum = fmaf(gw[i], weight[i], back_cum);
        }
    }
}
// adam_upd_kernel.cu:9-58, :72 as compiled: m' = fma(b1, m, (1-b1)*g); v' = fma(b2, v, g*((1-b2)*g))
static inline float dense_adam_stepsize(float lr, float beta1, float beta2, int step) {
    return lr * sqrtf(1 - powf(beta2, (float)step)) / (1 - powf(beta1, (float)step));
}
extern "C" void orc_dense_adam(float* p, const float* g, float* m, float* v, const float* perlr, int64_t n, int mode, int step,
                               float beta1, float beta2, float lr, float eps) {
    const float ss = dense_adam_stepsize(lr, beta1, beta2, step);
    for (int64_t i = 0; i < n; ++i) {
        if (mode == 1 && g[i] == 0.f) continue;
        const float nm = fmaf(beta1, m[i], (1.0f - beta1) * g[i]);
        const float nv = fmaf(beta2, v[i], g[i] * ((1.0f - beta2) * g[i]));
        m[i] = nm; v[i] = nv;
        const float st = (mode == 2) ? ss * perlr[i] : ss;
        p[i] = p[i] - (st * nm) / (eps + sqrtf(nv));
    }
}

// ------------------------------------------------------------------------------------------------
// Fine-stage training step.
// ------------------------------------------------------------------------------------------------
namespace {
constexpr int K0 = 12, PE = 27, DIN = 39, W = 128;
constexpr int OFF_W0 = 0, OFF_B0 = OFF_W0 + W * DIN, OFF_W1 = OFF_B0 + W, OFF_B1 = OFF_W1 + W * W, OFF_W2 = OFF_B1 + W,
              OFF_B2 = OFF_W2 + 3 * W, NET_N = OFF_B2 + 3;   // 22019

// viewdirs_emb (dvgo.py:354-356): [d, sin(d_a * 2^k), cos(d_a * 2^k)] with (a, k) flattened a-major
void view_embed(const float* vd, float* emb) {
    emb[0] = vd[0]; emb[1] = vd[1]; emb[2] = vd[2];
    for (int a = 0; a < 3; ++a)
        for (int k = 0; k < 4; ++k) {
            const float x = vd[a] * (float)(1 << k);
            emb[3 + a * 4 + k] = sinf(x);
            emb[15 + a * 4 + k] = cosf(x);
        }
}
inline float wld2idx(float p, float mn, float mx, float rm1) { return ((p - mn) / (mx - mn)) * rm1; }   // grid.py:77-78
inline float sigmoidf(float x) { return 1.0f / (1.0f + expf(-x)); }
}  // namespace

// One iteration of run.py:541-588.  k0->C == 12: fine stage, view-dependent colour through the rgbnet (dvgo.py:347-360).
// k0->C == 3: coarse stage — no rgbnet, rgb = sigmoid(k0) (dvgo.py:344-346, configs/default.py:94 rgbnet_dim = 0), `net` unused;
// den_perlr: the per-voxel lr grid of stepmode 2 (masked_adam.py:43-46, 58-59), may be null otherwise.
static void train_step_impl(const orc_train_cfg* cfg, orc_grid* den, orc_grid* den_grad, orc_grid* den_m, orc_grid* den_v,
                            orc_grid* k0, orc_grid* k0_grad, orc_grid* k0_m, orc_grid* k0_v, const orc_grid* den_perlr, const uint8_t* mask, float* net,
                            float* net_m, float* net_v, const float* rays_o, const float* rays_d, const float* viewdirs,
                            const float* target, int n_rays, orc_train_out* out) {
    const int threads = std::max(1, cfg->threads);
    const bool direct = k0->C == 3;
    const int KC = k0->C;
    const float* mn = cfg->xyz_min; const float* mx = cfg->xyz_max;
    const int R[3] = {cfg->reso[0], cfg->reso[1], cfg->reso[2]};
    // MaskGrid buffers (grid.py:229-231) in float32 like torch: scale = (shape-1)/xyz_len; shift = -xyz_min*scale
    float mscale[3], mshift[3], rm1[3];
    for (int a = 0; a < 3; ++a) {
        mscale[a] = (float)(R[a] - 1) / (mx[a] - mn[a]);
        mshift[a] = -mn[a] * mscale[a];
        rm1[a] = (float)(R[a] - 1);
    }
    const float thres = cfg->fast_color_thres;
    const int N = n_rays;
    const float Ng = (float)(cfg->n_rays_global > 0 ? cfg->n_rays_global : N);

    struct Smp { int step; float p[3]; float density, e, alpha, T, weight; int keep; };   // one per trimmed-M2 sample
    std::vector<std::vector<Smp>> ray_smp(N);
    std::vector<float> ail(N, 1.f);
    std::vector<int64_t> c_steps(N); std::vector<int> c_in(N), c_mask(N), c_afull(N), c_a(N), c_keep(N);
    std::vector<std::unordered_set<int64_t>> vmask_t(threads), vden_t(threads), vdeng_t(threads), vk0_t(threads);

    // ---- forward, geometry part: sample -> bbox -> mask -> density -> alpha -> weights (dvgo.py:306-335)
    std::atomic<int> tid_ctr{0};
    parallel_for(N, threads, [&](int64_t b, int64_t e_) {
        const int tid = tid_ctr++;
        auto& vmask = vmask_t[tid]; auto& vden = vden_t[tid];
        for (int64_t r = b; r < e_; ++r) {
            float tmin, tmax, s[3], dir[3];
            t_minmax(rays_o + r * 3, rays_d + r * 3, mn, mx, cfg->near, cfg->far, tmin, tmax);
            const int64_t ns = n_samples_of(rays_d + r * 3, tmin, tmax, cfg->stepdist);
            start_dir(rays_o + r * 3, rays_d + r * 3, tmin, s, dir);
            c_steps[r] = ns;
            float T_cum = 1.f;
            bool stopped = false;
            int n_in = 0, n_mask = 0, n_afull = 0, n_a = 0, n_keep = 0;
            for (int64_t k = 0; k < ns; ++k) {
                float p[3];
                ray_point(s, dir, cfg->stepdist, (int)k, p);
                if ((mn[0] > p[0]) | (mn[1] > p[1]) | (mn[2] > p[2]) | (mx[0] < p[0]) | (mx[1] < p[1]) | (mx[2] < p[2])) continue;
                ++n_in;
                const int mi = mask_ijk(p[0], mscale[0], mshift[0]), mj = mask_ijk(p[1], mscale[1], mshift[1]), mk = mask_ijk(p[2], mscale[2], mshift[2]);
                if (!(0 <= mi && mi < R[0] && 0 <= mj && mj < R[1] && 0 <= mk && mk < R[2])) continue;
                if (out) vmask.insert(((int64_t)mi * R[1] + mj) * R[2] + mk);
                if (!mask[((size_t)mi * R[1] + mj) * R[2] + mk]) continue;
                ++n_mask;
                // density query (grid.py:80-89 -> densityvdb.cu:101-125)
                const float x = wld2idx(p[0], mn[0], mx[0], rm1[0]), y = wld2idx(p[1], mn[1], mx[1], rm1[1]), z = wld2idx(p[2], mn[2], mx[2], rm1[2]);
                float d;
                int32_t cl[8], co[8];
                orc_sample_forward(den, &x, &y, &z, 1, &d, cl, co, 1);
                if (out) for (int q = 0; q < 8; ++q) if (cl[q] >= 0) vden.insert((int64_t)cl[q] * 512 + co[q]);
                float ex;
                const float a = raw2alpha1(d, cfg->act_shift, cfg->interval, ex);
                if (!(a > thres)) continue;
                ++n_afull;
                if (stopped) continue;     // past the early stop: weight 0 (zeros_like, :625), dropped by the weights mask
                Smp sm; sm.step = (int)k; sm.p[0] = p[0]; sm.p[1] = p[1]; sm.p[2] = p[2]; sm.density = d; sm.e = ex; sm.alpha = a;
                sm.T = T_cum; sm.weight = T_cum * a;
                T_cum = T_update(T_cum, a);
                sm.keep = sm.weight > thres;
                n_keep += sm.keep;
                ++n_a;
                ray_smp[r].push_back(sm);
                if ((double)T_cum < 1e-3) stopped = true;
            }
            ail[r] = T_cum;
            c_in[r] = n_in; c_mask[r] = n_mask; c_afull[r] = n_afull; c_a[r] = n_a; c_keep[r] = n_keep;
        }
    });

    // ---- forward, colour part (dvgo.py:338-370): k0 query, rgbnet, sigmoid, composite
    std::vector<int64_t> keep_off(N + 1, 0);
    for (int r = 0; r < N; ++r) keep_off[r + 1] = keep_off[r] + c_keep[r];
    const int64_t M3 = keep_off[N];
    std::vector<float> feat((size_t)M3 * DIN), h0((size_t)M3 * W), h1((size_t)M3 * W), rgb((size_t)M3 * 3);
    std::vector<int> k_ray(M3), k_idx(M3);   // ray and index into ray_smp[ray]
    std::vector<float> kx(M3), ky(M3), kz(M3);
    std::vector<int32_t> k_leaf((size_t)M3 * 8), k_offs((size_t)M3 * 8);
    for (int r = 0; r < N; ++r) {
        int64_t c = keep_off[r];
        for (size_t i = 0; i < ray_smp[r].size(); ++i)
            if (ray_smp[r][i].keep) {
                const Smp& sm = ray_smp[r][i];
                k_ray[c] = r; k_idx[c] = (int)i;
                kx[c] = wld2idx(sm.p[0], mn[0], mx[0], rm1[0]); ky[c] = wld2idx(sm.p[1], mn[1], mx[1], rm1[1]); kz[c] = wld2idx(sm.p[2], mn[2], mx[2], rm1[2]);
                ++c;
            }
    }
    {
        std::vector<float> k0v((size_t)M3 * KC);
        orc_sample_forward(k0, kx.data(), ky.data(), kz.data(), M3, k0v.data(), k_leaf.data(), k_offs.data(), threads);
        parallel_for(M3, threads, [&](int64_t b, int64_t e_) {
            for (int64_t s = b; s < e_; ++s) {
                for (int c = 0; c < KC; ++c) feat[s * DIN + c] = k0v[s * KC + c];
                if (!direct) view_embed(viewdirs + (size_t)k_ray[s] * 3, &feat[s * DIN + K0]);
            }
        });
    }
    const float *w0 = net + OFF_W0, *b0 = net + OFF_B0, *w1 = net + OFF_W1, *b1 = net + OFF_B1, *w2 = net + OFF_W2, *b2 = net + OFF_B2;
    // rgbnet (dvgo.py:99-107): Linear(39,128) ReLU Linear(128,128) ReLU Linear(128,3); fp32 in, double accumulate
    if (direct) {   // dvgo.py:344-346: rgb = torch.sigmoid(k0)
        for (int64_t s = 0; s < M3; ++s)
            for (int j = 0; j < 3; ++j) rgb[s * 3 + j] = sigmoidf(feat[s * DIN + j]);
    } else
    parallel_for(M3, threads, [&](int64_t b, int64_t e_) {
        for (int64_t s = b; s < e_; ++s) {
            for (int j = 0; j < W; ++j) {
                double a = b0[j];
                for (int i = 0; i < DIN; ++i) a += (double)w0[j * DIN + i] * (double)feat[s * DIN + i];
                h0[s * W + j] = a > 0 ? (float)a : 0.f;
            }
            for (int j = 0; j < W; ++j) {
                double a = b1[j];
                for (int i = 0; i < W; ++i) a += (double)w1[j * W + i] * (double)h0[s * W + i];
                h1[s * W + j] = a > 0 ? (float)a : 0.f;
            }
            for (int j = 0; j < 3; ++j) {
                double a = b2[j];
                for (int i = 0; i < W; ++i) a += (double)w2[j * W + i] * (double)h1[s * W + i];
                rgb[s * 3 + j] = sigmoidf((float)a);
            }
        }
    });
    std::vector<float> marched((size_t)N * 3);
    for (int r = 0; r < N; ++r) {
        double acc[3] = {0, 0, 0};
        for (int64_t s = keep_off[r]; s < keep_off[r + 1]; ++s) {
            const float w = ray_smp[r][k_idx[s]].weight;
            for (int c = 0; c < 3; ++c) acc[c] += (double)(w * rgb[s * 3 + c]);   // segment_coo sum (dvgo.py:365-369)
        }
        for (int c = 0; c < 3; ++c) marched[r * 3 + c] = (float)acc[c] + ail[r] * cfg->bg;   // dvgo.py:370
    }

    // ---- losses (run.py:551-574)
    double mse = 0, ent = 0, rgbper = 0;
    for (int r = 0; r < N; ++r) {
        for (int c = 0; c < 3; ++c) { const float d = marched[r * 3 + c] - target[r * 3 + c]; mse += (double)(d * d); }
        const float p = std::min(std::max(ail[r], 1e-6f), 1.0f - 1e-6f);
        ent += -((double)(p * logf(p)) + (double)((1.0f - p) * logf(1.0f - p)));
    }
    for (int64_t s = 0; s < M3; ++s) {
        const int r = k_ray[s];
        float acc = 0;
        for (int c = 0; c < 3; ++c) { const float d = rgb[s * 3 + c] - target[r * 3 + c]; acc += d * d; }
        rgbper += (double)(acc * ray_smp[r][k_idx[s]].weight);
    }
    mse /= (double)N * 3; ent /= (double)N; rgbper /= (double)N;
    // With n_rays_global != n_rays (data-parallel shard) the local sums are divided by the global N instead.
    const double shard = (double)N / (double)Ng;
    const double l_mse = mse * shard, l_ent = ent * shard, l_per = rgbper * shard;
    const float loss = (float)(cfg->weight_main * l_mse + cfg->weight_entropy_last * l_ent + cfg->weight_rgbper * l_per);

    // ---- backward
    // vdbopt.zero_grad() (run.py:549-550)
    orc_zero_grad(den_grad);
    orc_zero_grad(k0_grad);
    std::vector<float> g_marched((size_t)N * 3), g_last(N);
    for (int r = 0; r < N; ++r) {
        float gsum = 0;
        for (int c = 0; c < 3; ++c) {
            g_marched[r * 3 + c] = cfg->weight_main * 2.0f * (marched[r * 3 + c] - target[r * 3 + c]) / (Ng * 3.0f);
            gsum += g_marched[r * 3 + c];
        }
        float ge = 0;
        if (cfg->weight_entropy_last > 0 && ail[r] >= 1e-6f && ail[r] <= 1.0f - 1e-6f)
            ge = cfg->weight_entropy_last * (-(logf(ail[r]) - logf(1.0f - ail[r]))) / Ng;
        g_last[r] = gsum * cfg->bg + ge;
    }
    // d rgb, d weight for kept samples; sigmoid backward; rgbnet backward
    std::vector<float> g_logit((size_t)M3 * 3), g_w(M3), g_feat((size_t)M3 * KC);
    for (int64_t s = 0; s < M3; ++s) {
        const int r = k_ray[s];
        const float w = ray_smp[r][k_idx[s]].weight;
        float gw = 0;
        for (int c = 0; c < 3; ++c) {
            const float col = rgb[s * 3 + c];
            const float grgb = w * g_marched[r * 3 + c] + cfg->weight_rgbper * 2.0f * (col - target[r * 3 + c]) * w / Ng;
            g_logit[s * 3 + c] = grgb * col * (1.0f - col);
            gw += col * g_marched[r * 3 + c];
        }
        g_w[s] = gw;
    }
    std::vector<std::vector<double>> gnet_t(threads, std::vector<double>(NET_N, 0.0));
    tid_ctr = 0;
    if (direct) {   // the logit IS the interpolated k0 value
        for (int64_t s = 0; s < M3; ++s)
            for (int j = 0; j < 3; ++j) g_feat[s * 3 + j] = g_logit[s * 3 + j];
    } else
    parallel_for(M3, threads, [&](int64_t b, int64_t e_) {
        std::vector<double>& gn = gnet_t[tid_ctr++];
        std::vector<float> gh1(W), gh0(W);
        for (int64_t s = b; s < e_; ++s) {
            for (int i = 0; i < W; ++i) {
                double a = 0;
                for (int j = 0; j < 3; ++j) a += (double)g_logit[s * 3 + j] * (double)w2[j * W + i];
                gh1[i] = h1[s * W + i] > 0 ? (float)a : 0.f;
            }
            for (int j = 0; j < 3; ++j) {
                gn[OFF_B2 + j] += g_logit[s * 3 + j];
                for (int i = 0; i < W; ++i) gn[OFF_W2 + j * W + i] += (double)g_logit[s * 3 + j] * (double)h1[s * W + i];
            }
            for (int i = 0; i < W; ++i) {
                double a = 0;
                for (int j = 0; j < W; ++j) a += (double)gh1[j] * (double)w1[j * W + i];
                gh0[i] = h0[s * W + i] > 0 ? (float)a : 0.f;
            }
            for (int j = 0; j < W; ++j) {
                if (gh1[j] == 0.f) continue;
                gn[OFF_B1 + j] += gh1[j];
                for (int i = 0; i < W; ++i) gn[OFF_W1 + j * W + i] += (double)gh1[j] * (double)h0[s * W + i];
            }
            for (int j = 0; j < W; ++j) {
                if (gh0[j] == 0.f) continue;
                gn[OFF_B0 + j] += gh0[j];
                for (int i = 0; i < DIN; ++i) gn[OFF_W0 + j * DIN + i] += (double)gh0[j] * (double)feat[s * DIN + i];
            }
            for (int i = 0; i < K0; ++i) {
                double a = 0;
                for (int j = 0; j < W; ++j) a += (double)gh0[j] * (double)w0[j * DIN + i];
                g_feat[s * K0 + i] = (float)a;
            }
        }
    });
    std::vector<float> gnet(NET_N);
    for (int i = 0; i < NET_N; ++i) { double a = 0; for (int t = 0; t < threads; ++t) a += gnet_t[t][i]; gnet[i] = (float)a; }
    // k0 gradient scatter (grid.py:53-60 -> colorvdb.cu:130-175)
    orc_sample_backward(k0_grad, kx.data(), ky.data(), kz.data(), g_feat.data(), M3, 1);
    // alpha2weight backward (:654-677) per ray over the trimmed segment, raw2alpha backward, density scatter
    {
        std::vector<float> dx, dy, dz, dg;
        for (int r = 0; r < N; ++r) {
            auto& v = ray_smp[r];
            float back_cum = g_last[r] * ail[r];
            // position of each kept sample inside the kept list, walking backwards
            int64_t ks = keep_off[r + 1];
            for (int i = (int)v.size() - 1; i >= 0; --i) {
                float gw = 0;
                if (v[i].keep) { --ks; gw = g_w[ks]; }
                const float ga = a2w_grad(gw, v[i].T, back_cum, v[i].alpha);
                back_cum = fmaf(gw, v[i].weight, back_cum);
                const float gd = raw2alpha_bwd1(v[i].e, ga, cfg->interval);
                dx.push_back(wld2idx(v[i].p[0], mn[0], mx[0], rm1[0]));
                dy.push_back(wld2idx(v[i].p[1], mn[1], mx[1], rm1[1]));
                dz.push_back(wld2idx(v[i].p[2], mn[2], mx[2], rm1[2]));
                dg.push_back(gd);
            }
        }
        orc_sample_backward(den_grad, dx.data(), dy.data(), dz.data(), dg.data(), (int64_t)dg.size(), 1);
        if (out) {
            std::vector<float> tmp(1); std::vector<int32_t> cl(8), co(8);
            for (size_t i = 0; i < dg.size(); ++i) {
                if (dg[i] == 0.f) continue;
                Tri t(dx[i], dy[i], dz[i]);
                for (int q = 0; q < 8; ++q) {
                    const int x = t.i + CORNER[q][0], y = t.j + CORNER[q][1], z = t.k + CORNER[q][2];
                    const int leaf = find_leaf(den_grad, x, y, z);
                    if (leaf >= 0) vdeng_t[0].insert((int64_t)leaf * 512 + leaf_off(x, y, z));
                }
            }
        }
    }

    // ---- update (run.py:585-588): MaskedAdam on rgbnet (mode 0), VDBAdam step on density / k0
    if (cfg->do_update) {
        if (!direct) orc_dense_adam(net, gnet.data(), net_m, net_v, nullptr, NET_N, 0, cfg->step, cfg->beta0, cfg->beta1, cfg->lr_net, cfg->eps);
        orc_adam_step(den, den_grad, den_m, den_v, cfg->den_mode, orc_adam_stepsize(cfg->lr_density, cfg->beta0, cfg->beta1, cfg->step),
                      cfg->eps, cfg->beta0, cfg->beta1, cfg->den_mode == 2 ? den_perlr : nullptr);
        orc_adam_step(k0, k0_grad, k0_m, k0_v, cfg->k0_mode, orc_adam_stepsize(cfg->lr_k0, cfg->beta0, cfg->beta1, cfg->step), cfg->eps,
                      cfg->beta0, cfg->beta1, nullptr);
    }

    if (!out) return;
    out->loss[0] = loss; out->loss[1] = (float)l_mse; out->loss[2] = (float)l_ent; out->loss[3] = (float)l_per;
    out->M0 = out->M0_in = out->M1 = out->M2 = out->M2_trim = 0; out->M3 = M3;
    for (int r = 0; r < N; ++r) {
        out->M0 += c_steps[r]; out->M0_in += c_in[r]; out->M1 += c_mask[r]; out->M2 += c_afull[r]; out->M2_trim += c_a[r];
        if (out->n_steps) out->n_steps[r] = c_steps[r];
        if (out->cnt_inbbox) out->cnt_inbbox[r] = c_in[r];
        if (out->cnt_mask) out->cnt_mask[r] = c_mask[r];
        if (out->cnt_alpha_full) out->cnt_alpha_full[r] = c_afull[r];
        if (out->cnt_alpha) out->cnt_alpha[r] = c_a[r];
        if (out->cnt_keep) out->cnt_keep[r] = c_keep[r];
        if (out->alphainv_last) out->alphainv_last[r] = ail[r];
        if (out->rgb_marched) for (int c = 0; c < 3; ++c) out->rgb_marched[r * 3 + c] = marched[r * 3 + c];
    }
    for (int64_t s = 0; s < std::min<int64_t>(M3, out->cap_keep); ++s) {
        const int r = k_ray[s];
        const Smp& sm = ray_smp[r][k_idx[s]];
        if (out->keep_ray) out->keep_ray[s] = r;
        if (out->keep_step) out->keep_step[s] = sm.step;
        if (out->keep_weight) out->keep_weight[s] = sm.weight;
        if (out->keep_rgb) for (int c = 0; c < 3; ++c) out->keep_rgb[s * 3 + c] = rgb[s * 3 + c];
        if (out->keep_feat) for (int c = 0; c < K0; ++c) out->keep_feat[s * K0 + c] = c < KC ? feat[s * DIN + c] : 0.f;
        if (out->keep_leaf) for (int q = 0; q < 8; ++q) out->keep_leaf[s * 8 + q] = k_leaf[s * 8 + q];
        if (out->keep_off) for (int q = 0; q < 8; ++q) out->keep_off[s * 8 + q] = k_offs[s * 8 + q];
    }
    if (out->net_grad) std::memcpy(out->net_grad, gnet.data(), NET_N * sizeof(float));
    std::unordered_set<int64_t> vm, vd, vk;
    for (int t = 0; t < threads; ++t) { vm.insert(vmask_t[t].begin(), vmask_t[t].end()); vd.insert(vden_t[t].begin(), vden_t[t].end()); }
    for (int64_t s = 0; s < M3; ++s) for (int q = 0; q < 8; ++q) if (k_leaf[s * 8 + q] >= 0) vk.insert((int64_t)k_leaf[s * 8 + q] * 512 + k_offs[s * 8 + q]);
    out->V_mask = (int64_t)vm.size(); out->V_den = (int64_t)vd.size(); out->V_den_grad = (int64_t)vdeng_t[0].size(); out->V_k0 = (int64_t)vk.size();
}

extern "C" void orc_train_step(const orc_train_cfg* cfg, orc_grid* den, orc_grid* den_grad, orc_grid* den_m, orc_grid* den_v,
                               orc_grid* k0, orc_grid* k0_grad, orc_grid* k0_m, orc_grid* k0_v, const uint8_t* mask, float* net,
                               float* net_m, float* net_v, const float* rays_o, const float* rays_d, const float* viewdirs,
                               const float* target, int n_rays, orc_train_out* out) {
    train_step_impl(cfg, den, den_grad, den_m, den_v, k0, k0_grad, k0_m, k0_v, nullptr, mask, net, net_m, net_v, rays_o, rays_d, viewdirs, target, n_rays, out);
}
extern "C" void orc_train_step_perlr(const orc_train_cfg* cfg, orc_grid* den, orc_grid* den_grad, orc_grid* den_m, orc_grid* den_v,
                                     orc_grid* k0, orc_grid* k0_grad, orc_grid* k0_m, orc_grid* k0_v, const orc_grid* den_perlr, const uint8_t* mask,
                                     float* net, float* net_m, float* net_v, const float* rays_o, const float* rays_d, const float* viewdirs,
                                     const float* target, int n_rays, orc_train_out* out) {
    train_step_impl(cfg, den, den_grad, den_m, den_v, k0, k0_grad, k0_m, k0_v, den_perlr, mask, net, net_m, net_v, rays_o, rays_d, viewdirs, target, n_rays, out);
}

// ------------------------------------------------------------------------------------------------
// R1 — merge (vdb_compression.py:19-59)
// ------------------------------------------------------------------------------------------------
static inline float through_half(float f) {   // torch .half() round-to-nearest-even, back to float
    _Float16 h = (_Float16)f;
    return (float)h;
}
extern "C" int64_t orc_merge(const orc_grid* den, const orc_grid* k0, const uint8_t* mask, float* dendata, float* coldata,
                             float* idx_dense) {
    const int rx = den->rx, ry = den->ry, rz = den->rz, C = k0->C;
    int64_t N = 0;
    if (dendata) { dendata[0] = 0.f; for (int c = 0; c < C; ++c) coldata[c] = 0.f; }
    for (int x = 0; x < rx; ++x) for (int y = 0; y < ry; ++y) for (int z = 0; z < rz; ++z) {
        const size_t v = ((size_t)x * ry + y) * rz + z;
        if (!mask[v]) { if (idx_dense) idx_dense[v] = 0.f; continue; }
        ++N;                                   // idxs = arange(1, N+1) in C order over the mask (:31-33)
        if (idx_dense) idx_dense[v] = (float)N;
        if (dendata) {
            const int ld = find_leaf(den, x, y, z), lk = find_leaf(k0, x, y, z), off = leaf_off(x, y, z);
            dendata[N] = through_half(ld >= 0 ? den->values[(size_t)ld * 512 + off] : 0.f);
            for (int c = 0; c < C; ++c) coldata[N * C + c] = through_half(lk >= 0 ? k0->values[((size_t)lk * 512 + off) * C + c] : 0.f);
        }
    }
    return N;
}

// ------------------------------------------------------------------------------------------------
// R2 — merged renderer (renderer.cu).  Per pixel: get_rays :122-167, first_look_merged :222-268 (count +
// tighten t range), ray_marching_merged :312-367 (gather), cuda_rgbnet :83-109, final_render :112-119.
// ------------------------------------------------------------------------------------------------
namespace {
struct IdxAcc {
    const orc_grid* g;
    int value(int x, int y, int z) const {   // int(acc.getValue(coord)) (:202-209)
        const int leaf = find_leaf(g, x, y, z);
        return leaf >= 0 ? (int)g->values[(size_t)leaf * 512 + leaf_off(x, y, z)] : 0;
    }
    bool active(int x, int y, int z) const {   // acc.isActive: false when no leaf (root tile without child, NanoVDB.h:3035-3045)
        const int leaf = find_leaf(g, x, y, z);
        return leaf >= 0 && leaf_active(g, leaf, leaf_off(x, y, z));
    }
};
// trigetDensity / trigetDensity2 (:191-220, :271-300).  First-look: res += d*(f0)*(f1)*(f2) -> fma(f2, f1*(f0*d), res);
// second pass: one expression d0*s0 + d1*s1 + ... -> mul then a chain of fma.
inline void tri_setup(const float* xyz, int* ijk, float* uvw) {
    for (int a = 0; a < 3; ++a) { ijk[a] = (int)xyz[a]; uvw[a] = xyz[a] - (float)ijk[a]; }   // int() truncation (:194-199)
}
// get_rays (:134-166) + get_tminmax (:50-64) of pixel n: direction in unit-box coordinates, unit view direction, step length in t
// and the t range against the unit box.
inline void pixel_ray(const orc_render_cfg* cfg, const float* c2w, const float* ro, const float* ext, int n, float* rd, float* vd,
                      float& steplen, float& tmin, float& tmax) {
    const int Wd = cfg->W;
    const float far = 1e9f;   // setKwargs ignores `far` (plenvdb.h:1008)
    const float pixeli = (float)((double)(n % Wd) + 0.5), pixelj = (float)((double)(n / Wd) + 0.5);
    float dir[3];
    if (cfg->inverse_y) { dir[0] = (pixeli - cfg->K[2]) / cfg->K[0]; dir[1] = (pixelj - cfg->K[5]) / cfg->K[4]; dir[2] = 1.f; }
    else { dir[0] = (pixeli - cfg->K[2]) / cfg->K[0]; dir[1] = -((pixelj - cfg->K[5]) / cfg->K[4]); dir[2] = -1.f; }
    float rdw[3];
    for (int a = 0; a < 3; ++a)   // d0*c0 + d1*c1 + d2*c2 compiles to fma(d2,c2, fma(d0,c0, d1*c1)) (:140-142)
        rdw[a] = fmaf(dir[2], c2w[a * 4 + 2], fmaf(dir[0], c2w[a * 4], dir[1] * c2w[a * 4 + 1]));
    const float len = sqrtf(fmaf(rdw[2], rdw[2], fmaf(rdw[0], rdw[0], rdw[1] * rdw[1])));   // Vec3::length -> fma(z,z, fma(x,x, y*y))
    steplen = cfg->stepdist / len;
    for (int a = 0; a < 3; ++a) rd[a] = rdw[a] / ext[a];
    const float inv = 1.0f / len;   // normalize(): *this *= 1/length (NanoVDB.h:1122-1123)
    for (int a = 0; a < 3; ++a) vd[a] = rdw[a] * inv;
    // get_tminmax (:50-64) against the unit box
    const float vx = rd[0] == 0 ? 1e-6f : rd[0], vy = rd[1] == 0 ? 1e-6f : rd[1], vz = rd[2] == 0 ? 1e-6f : rd[2];
    const float ax = (1 - ro[0]) / vx, ay = (1 - ro[1]) / vy, az = (1 - ro[2]) / vz;
    const float bx = -ro[0] / vx, by = -ro[1] / vy, bz = -ro[2] / vz;
    tmin = fmaxf(fminf(fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fminf(az, bz)), far), cfg->near);
    tmax = fmaxf(fminf(fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fmaxf(az, bz)), far), cfg->near);
}
}  // namespace

extern "C" void orc_render(const orc_render_cfg* cfg, const orc_grid* idx_grid, const float* dendata, const float* coldata, int cdim,
                           const float* w0, const float* b0, const float* w1, const float* b1, const float* w2, const float* b2,
                           const float* c2w, int row_begin, int row_end, float* out_rgb, int32_t* n_samples_out,
                           int32_t* inconsistent_rays) {
    const int Wd = cfg->W;
    const int64_t npix = (int64_t)(row_end - row_begin) * Wd;
    const IdxAcc acc{idx_grid};
    const float ext[3] = {cfg->xyz_max[0] - cfg->xyz_min[0], cfg->xyz_max[1] - cfg->xyz_min[1], cfg->xyz_max[2] - cfg->xyz_min[2]};
    // rays_o normalised to the unit box (:128-132)
    const float ro[3] = {(c2w[3] - cfg->xyz_min[0]) / ext[0], (c2w[7] - cfg->xyz_min[1]) / ext[1], (c2w[11] - cfg->xyz_min[2]) / ext[2]};
    const float wld[3] = {(float)(cfg->reso[0] - 1), (float)(cfg->reso[1] - 1), (float)(cfg->reso[2] - 1)};
    const float thres = cfg->fast_color_thres;
    const float far = 1e9f;   // setKwargs ignores `far` (plenvdb.h:1008)
    std::atomic<int> bad{0};
    parallel_for(npix, std::max(1, cfg->threads), [&](int64_t pb, int64_t pe) {
        std::vector<float> feats, weights, h0(W), h1(W);
        for (int64_t pp = pb; pp < pe; ++pp) {
            const int n = (int)(row_begin * (int64_t)Wd + pp);
            float rd[3], vd[3], steplen, tmin, tmax;
            pixel_ray(cfg, c2w, ro, ext, n, rd, vd, steplen, tmin, tmax);
            float pe_feat[PE];
            pe_feat[0] = vd[0]; pe_feat[1] = vd[1]; pe_feat[2] = vd[2];
            {
                int pebase = 1;
                for (int i_ = 3; i_ < 7; ++i_) {
                    for (int a = 0; a < 3; ++a) {
                        pe_feat[i_ + 4 * a] = sinf(vd[a] * (float)pebase);
                        pe_feat[i_ + 12 + 4 * a] = cosf(vd[a] * (float)pebase);
                    }
                    pebase *= 2;
                }
            }
            // pass 1 (:240-266)
            int ns = 0;
            {
                float T_cum = 1.f, t = tmin;
                bool update_tmin = false;
                const float tmax0 = tmax;
                while (t < tmax0) {
                    t += steplen;
                    float p[3], xyz[3];
                    for (int a = 0; a < 3; ++a) p[a] = fmaf(t, rd[a], ro[a]);
                    if ((0 > p[0]) | (0 > p[1]) | (0 > p[2]) | (1 < p[0]) | (1 < p[1]) | (1 < p[2])) continue;
                    for (int a = 0; a < 3; ++a) xyz[a] = p[a] * wld[a];
                    if (!acc.active((int)rintf(xyz[0]), (int)rintf(xyz[1]), (int)rintf(xyz[2]))) continue;
                    int ijk[3]; float u[3];
                    tri_setup(xyz, ijk, u);
                    float res = 0;
                    for (int q = 0; q < 8; ++q) {
                        const int link = acc.value(ijk[0] + CORNER[q][0], ijk[1] + CORNER[q][1], ijk[2] + CORNER[q][2]);
                        const float f0 = CORNER[q][0] ? u[0] : 1 - u[0], f1 = CORNER[q][1] ? u[1] : 1 - u[1], f2 = CORNER[q][2] ? u[2] : 1 - u[2];
                        res = fmaf(f2, f1 * (f0 * dendata[link]), res);
                    }
                    const float alpha = 1 - powf(1 + expf(res + cfg->act_shift), -cfg->interval);
                    if (alpha <= thres) continue;
                    const float weight = T_cum * alpha;
                    T_cum *= (1 - alpha);
                    if (weight <= thres) continue;
                    ++ns;
                    if (!update_tmin) { tmin = t - steplen; update_tmin = true; }
                    if (T_cum < 1e-3) { tmax = t; break; }
                }
            }
            if (n_samples_out) n_samples_out[pp] = ns;
            float* px = out_rgb + pp * 3;
            if (ns == 0) { px[0] = px[1] = px[2] = cfg->bg; continue; }
            // pass 2 (:333-366): restart from the stored tmin, no early break
            feats.assign((size_t)ns * cdim, 0.f); weights.assign(ns, 0.f);
            int r = 0;
            float T_cum = 1.f, t = tmin;
            while (t < tmax) {
                t += steplen;
                float p[3], xyz[3];
                for (int a = 0; a < 3; ++a) p[a] = fmaf(t, rd[a], ro[a]);
                if ((0 > p[0]) | (0 > p[1]) | (0 > p[2]) | (1 < p[0]) | (1 < p[1]) | (1 < p[2])) continue;
                for (int a = 0; a < 3; ++a) xyz[a] = p[a] * wld[a];
                if (!acc.active((int)rintf(xyz[0]), (int)rintf(xyz[1]), (int)rintf(xyz[2]))) continue;
                int ijk[3]; float u[3]; int idx[8]; float sc[8];
                tri_setup(xyz, ijk, u);
                for (int q = 0; q < 8; ++q) {
                    idx[q] = acc.value(ijk[0] + CORNER[q][0], ijk[1] + CORNER[q][1], ijk[2] + CORNER[q][2]);
                    const float f0 = CORNER[q][0] ? u[0] : 1 - u[0], f1 = CORNER[q][1] ? u[1] : 1 - u[1], f2 = CORNER[q][2] ? u[2] : 1 - u[2];
                    sc[q] = (f0 * f1) * f2;
                }
                // d0*s0 + d1*s1 + ... compiles to fma(d7,s7, ... fma(d2,s2, fma(d0,s0, d1*s1))) (:298-299)
                float vden = fmaf(dendata[idx[0]], sc[0], dendata[idx[1]] * sc[1]);
                for (int q = 2; q < 8; ++q) vden = fmaf(dendata[idx[q]], sc[q], vden);
                const float alpha = 1 - powf(1 + expf(vden + cfg->act_shift), -cfg->interval);
                if (alpha <= thres) continue;
                const float weight = T_cum * alpha;
                T_cum *= (1 - alpha);
                if (weight <= thres) continue;
                if (r >= ns) { ++r; continue; }   // the reference would overrun its segment here
                for (int i = 0; i < cdim; ++i) {
                    float v = fmaf(coldata[(size_t)idx[0] * cdim + i], sc[0], coldata[(size_t)idx[1] * cdim + i] * sc[1]);   // :305-308
                    for (int q = 2; q < 8; ++q) v = fmaf(coldata[(size_t)idx[q] * cdim + i], sc[q], v);
                    feats[(size_t)r * cdim + i] = v;
                }
                weights[r++] = weight;
            }
            if (r != ns) ++bad;
            const float last = T_cum * cfg->bg;
            double accum[3] = {last, last, last};
            // cuda_rgbnet (:83-109): W0 split into colour rows [0,cdim) and PE rows [cdim,cdim+27); weights in the
            // transposed layout w0[39][128], w1[128][128], w2[128][3] (run.py:98-104)
            std::vector<double> pe_part(W);
            for (int j = 0; j < W; ++j) {
                double a = 0;
                for (int i = 0; i < PE; ++i) a += (double)w0[(cdim + i) * W + j] * (double)pe_feat[i];
                pe_part[j] = (double)(float)a;
            }
            for (int s = 0; s < std::min(r, ns); ++s) {
                for (int j = 0; j < W; ++j) {
                    double a = 0;
                    for (int i = 0; i < cdim; ++i) a += (double)w0[i * W + j] * (double)feats[(size_t)s * cdim + i];
                    const float v = ((float)a + (float)pe_part[j]) + b0[j];
                    h0[j] = v < 0 ? 0.f : v;
                }
                for (int j = 0; j < W; ++j) {
                    double a = 0;
                    for (int i = 0; i < W; ++i) a += (double)w1[i * W + j] * (double)h0[i];
                    const float v = (float)a + b1[j];
                    h1[j] = v < 0 ? 0.f : v;
                }
                for (int j = 0; j < 3; ++j) {
                    double a = 0;
                    for (int i = 0; i < W; ++i) a += (double)w2[i * 3 + j] * (double)h1[i];
                    const float raw = (float)a + b2[j];
                    accum[j] += (double)(weights[s] / (1 + expf(-raw)));   // final_render (:115-117)
                }
            }
            for (int j = 0; j < 3; ++j) px[j] = (float)accum[j];
        }
    });
    if (inconsistent_rays) *inconsistent_rays = bad.load();
}

// ------------------------------------------------------------------------------------------------
// Exactness check of the fast march of plenvdb_b200/csrc/renderer.cu (test infrastructure, like everything here).
// For every pixel it runs (A) the reference's two marches step by step, as orc_render does, and (B) a restatement of the
// fast march: runs of `skip_k` steps whose index box cannot touch a leaf are replaced by their `t += steplen` chain alone
// (dilated block map: a block or one of its +1 neighbours holds a leaf), the remaining steps are evaluated `lanes` at a time
// BEFORE the ordered bookkeeping (the lane-parallel form: values of later steps are computed speculatively and ignored
// after an early stop), and the second march is simulated on the corner values of the first — its own interpolation order,
// transmittance and thresholds — from the first kept sample on.  Compared: sample count and tightened t range of every
// pixel; for a pixel the fast march would hand over ((t_first - steplen) + steplen == t_first, simulated count == count,
// count <= slot) the (t, weight) of every sample of the reference's second march and its final transmittance, bit for bit.
// stats: 0 pixels, 1 pixels with samples, 2 handed over, 3 not handed: t chain, 4 not handed: count differs, 5 not handed:
// slot too small, 6 MISMATCHES (must be 0), 7 steps skipped, 8 steps in total, 9 reference-inconsistent pixels (r != ns).
// ------------------------------------------------------------------------------------------------
extern "C" void orc_march_check(const orc_render_cfg* cfg, const orc_grid* idx_grid, const float* dendata, const float* c2w,
                                int skip_k, int lanes, int slot, int64_t* stats) {
    const int Wd = cfg->W, Hd = cfg->H;
    const IdxAcc acc{idx_grid};
    const float ext[3] = {cfg->xyz_max[0] - cfg->xyz_min[0], cfg->xyz_max[1] - cfg->xyz_min[1], cfg->xyz_max[2] - cfg->xyz_min[2]};
    const float ro[3] = {(c2w[3] - cfg->xyz_min[0]) / ext[0], (c2w[7] - cfg->xyz_min[1]) / ext[1], (c2w[11] - cfg->xyz_min[2]) / ext[2]};
    const float wld[3] = {(float)(cfg->reso[0] - 1), (float)(cfg->reso[1] - 1), (float)(cfg->reso[2] - 1)};
    const float thres = cfg->fast_color_thres;
    // dilated block map
    const int nb[3] = {(cfg->reso[0] + 7) / 8, (cfg->reso[1] + 7) / 8, (cfg->reso[2] + 7) / 8};
    std::vector<uint8_t> blk((size_t)nb[0] * nb[1] * nb[2], 0);
    for (int l = 0; l < idx_grid->n_leaf(); ++l) {
        const int bx = idx_grid->origin[l * 3] >> 3, by = idx_grid->origin[l * 3 + 1] >> 3, bz = idx_grid->origin[l * 3 + 2] >> 3;
        for (int d = 0; d < 8; ++d) {
            const int x = bx - (d & 1), y = by - ((d >> 1) & 1), z = bz - (d >> 2);
            if (x < 0 || y < 0 || z < 0 || x >= nb[0] || y >= nb[1] || z >= nb[2]) continue;
            blk[((size_t)x * nb[1] + y) * nb[2] + z] = 1;
        }
    }
    std::atomic<int64_t> st[10];
    for (auto& v : st) v = 0;
    struct Sample { float t, w; };
    // one step of either march: position, active test, corner densities; returns false when the step has no effect
    auto step_values = [&](const float* rd, float t, float* uvw, float* den) -> bool {
        float p[3], xyz[3];
        for (int a = 0; a < 3; ++a) p[a] = fmaf(t, rd[a], ro[a]);
        if ((0 > p[0]) | (0 > p[1]) | (0 > p[2]) | (1 < p[0]) | (1 < p[1]) | (1 < p[2])) return false;
        for (int a = 0; a < 3; ++a) xyz[a] = p[a] * wld[a];
        if (!acc.active((int)rintf(xyz[0]), (int)rintf(xyz[1]), (int)rintf(xyz[2]))) return false;
        int ijk[3];
        tri_setup(xyz, ijk, uvw);
        for (int q = 0; q < 8; ++q) den[q] = dendata[acc.value(ijk[0] + CORNER[q][0], ijk[1] + CORNER[q][1], ijk[2] + CORNER[q][2])];
        return true;
    };
    auto alpha_first = [&](const float* u, const float* den) {    // trigetDensity (:191-220)
        float res = 0;
        for (int q = 0; q < 8; ++q) {
            const float f0 = CORNER[q][0] ? u[0] : 1 - u[0], f1 = CORNER[q][1] ? u[1] : 1 - u[1], f2 = CORNER[q][2] ? u[2] : 1 - u[2];
            res = fmaf(f2, f1 * (f0 * den[q]), res);
        }
        return 1 - powf(1 + expf(res + cfg->act_shift), -cfg->interval);
    };
    auto alpha_second = [&](const float* u, const float* den) {   // trigetDensity2 (:271-300)
        float sc[8];
        for (int q = 0; q < 8; ++q) {
            const float f0 = CORNER[q][0] ? u[0] : 1 - u[0], f1 = CORNER[q][1] ? u[1] : 1 - u[1], f2 = CORNER[q][2] ? u[2] : 1 - u[2];
            sc[q] = (f0 * f1) * f2;
        }
        float vden = fmaf(den[0], sc[0], den[1] * sc[1]);
        for (int q = 2; q < 8; ++q) vden = fmaf(den[q], sc[q], vden);
        return 1 - powf(1 + expf(vden + cfg->act_shift), -cfg->interval);
    };
    parallel_for((int64_t)Hd * Wd, std::max(1, cfg->threads), [&](int64_t pb, int64_t pe) {
        std::vector<Sample> ref2, sim2;
        std::vector<float> lt(lanes), la1(lanes), la2(lanes);
        std::vector<uint8_t> lact(lanes);
        for (int64_t pp = pb; pp < pe; ++pp) {
            float rd[3], vd[3], steplen, tmin0, tmax0;
            pixel_ray(cfg, c2w, ro, ext, (int)pp, rd, vd, steplen, tmin0, tmax0);
            // ---- (A) the reference, step by step
            int ns = 0, r = 0;
            float tmin = tmin0, tmax = tmax0, T_last;
            {
                float T_cum = 1.f, t = tmin0, u[3], den[8];
                bool update_tmin = false;
                while (t < tmax0) {
                    t += steplen;
                    if (!step_values(rd, t, u, den)) continue;
                    const float alpha = alpha_first(u, den);
                    if (alpha <= thres) continue;
                    const float weight = T_cum * alpha;
                    T_cum *= (1 - alpha);
                    if (weight <= thres) continue;
                    ++ns;
                    if (!update_tmin) { tmin = t - steplen; update_tmin = true; }
                    if (T_cum < 1e-3) { tmax = t; break; }
                }
                ref2.clear();
                T_cum = 1.f; t = tmin;
                if (ns > 0)
                    while (t < tmax) {
                        t += steplen;
                        if (!step_values(rd, t, u, den)) continue;
                        const float alpha = alpha_second(u, den);
                        if (alpha <= thres) continue;
                        const float weight = T_cum * alpha;
                        T_cum *= (1 - alpha);
                        if (weight <= thres) continue;
                        ref2.push_back({t, weight});
                        ++r;
                    }
                T_last = T_cum;
            }
            // ---- (B) the fast march
            int ns_f = 0, r2 = 0;
            float tmin_f = tmin0, tmax_f = tmax0, T2 = 1.f;
            bool sim = false, chain_exact = true;
            int64_t skipped = 0, total = 0;
            {
                float T_cum = 1.f, t = tmin0;
                bool update_tmin = false, done = false;
                sim2.clear();
                while (!done && t < tmax0) {
                    // chain of up to skip_k steps
                    float t1 = t + steplen, tk = t1;
                    int k = 1;
                    for (int q = 1; q < skip_k; ++q)
                        if (tk < tmax0) { tk += steplen; ++k; }
                    total += k;
                    bool empty = true, decided = false;
                    int b[3];
                    for (int a = 0; a < 3 && !decided; ++a) {
                        const float pa = fmaf(rd[a], t1, ro[a]), pbb = fmaf(rd[a], tk, ro[a]);
                        if (pa != pa || pbb != pbb) { empty = false; decided = true; break; }
                        const float lo = fmaxf(fminf(pa, pbb), 0.f), hi = fminf(fmaxf(pa, pbb), 1.f);
                        if (lo > hi) { empty = true; decided = true; break; }
                        const int ilo = (int)rintf(lo * wld[a]) >> 3, ihi = (int)rintf(hi * wld[a]) >> 3;
                        if (ihi - ilo > 1 || ilo < 0 || ihi >= nb[a]) { empty = false; decided = true; break; }
                        b[a] = ilo;
                    }
                    if (!decided) empty = !blk[((size_t)b[0] * nb[1] + b[1]) * nb[2] + b[2]];
                    if (empty) { t = tk; skipped += k; continue; }
                    // the k steps of the run, `lanes` at a time: values first, ordered bookkeeping second
                    for (int base = 0; base < k && !done; base += lanes) {
                        const int m = std::min(lanes, k - base);
                        float tt = t;
                        for (int q = 0; q < m; ++q) {
                            tt += steplen;
                            lt[q] = tt;
                            float u[3], den[8];
                            lact[q] = step_values(rd, tt, u, den);
                            if (lact[q]) { la1[q] = alpha_first(u, den); la2[q] = alpha_second(u, den); }
                        }
                        for (int q = 0; q < m; ++q) {
                            t = lt[q];
                            if (!lact[q]) continue;
                            bool kept = false;
                            if (la1[q] > thres) {
                                const float weight = T_cum * la1[q];
                                T_cum *= (1 - la1[q]);
                                kept = weight > thres;
                            }
                            if (kept) {
                                ++ns_f;
                                if (!update_tmin) {
                                    tmin_f = t - steplen;
                                    update_tmin = true;
                                    sim = true;
                                    chain_exact = (tmin_f + steplen) == t;
                                }
                            }
                            if (sim && la2[q] > thres) {
                                const float w2 = T2 * la2[q];
                                T2 *= (1 - la2[q]);
                                if (w2 > thres) { sim2.push_back({t, w2}); ++r2; }
                            }
                            if (kept && T_cum < 1e-3) { tmax_f = t; done = true; break; }
                        }
                    }
                }
            }
            st[0]++; st[7] += skipped; st[8] += total;
            if (ns > 0) st[1]++;
            if (ns > 0 && r != ns) st[9]++;
            bool bad = ns_f != ns || std::memcmp(&tmin_f, &tmin, 4) || std::memcmp(&tmax_f, &tmax, 4);
            if (ns > 0) {
                if (!chain_exact) st[3]++;
                else if (r2 != ns) st[4]++;
                else if (ns > slot) st[5]++;
                else {
                    st[2]++;
                    bad = bad || r != ns || (int)ref2.size() != ns || (int)sim2.size() != ns || std::memcmp(&T2, &T_last, 4);
                    if (!bad) bad = std::memcmp(ref2.data(), sim2.data(), sizeof(Sample) * ns) != 0;
                }
            }
            if (bad) st[6]++;
        }
    });
    for (int i = 0; i < 10; ++i) stats[i] = st[i].load();
}
