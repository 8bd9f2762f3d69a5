// ref_harness.cu — thin C harness around the REFERENCE's own code.  TEST INFRASTRUCTURE ONLY.
//
// Built by oracle/build_oracle.py against /root/reference IN PLACE (include paths only; no reference
// source is copied into this repository), outputs under the git-ignored oracle/_ref/:
//   * libref_host.so (g++, no CUDA): grids are built with the reference's nanovdb::GridBuilder and read
//     through the reference's own nanovdb::ReadAccessor on the host.  The PlenVDB kernel bodies for
//     D1/D2/C1/C2 are executed through that accessor (they are __hostdev__-clean), which pins the
//     oracle's tree semantics: leaf order, background fall-back, isCached<Leaf>.
//   * libref_gpu.so (nvcc, -DREF_WITH_CUDA): additionally links the reference's unmodified
//     plenvdb.cu / densityvdb.cu / colorvdb.cu / renderer.cu compiled for sm_100a and calls their
//     extern host wrappers (declared in plenvdb/lib/vdb/plenvdb.h:70-86, 307-387, 606-684, 921-929,
//     re-declared here because plenvdb.h itself pulls in OpenVDB which cannot be built).
#include <nanovdb/NanoVDB.h>
#include <nanovdb/util/GridBuilder.h>
#ifdef REF_WITH_CUDA
#include <cuda_runtime.h>
#include <nanovdb/util/CudaDeviceBuffer.h>
#include "plenvdb.cuh"   // RenderKwargs / SceneInfo / MLP PODs of the reference (plenvdb/lib/vdb/plenvdb.cuh:37-79)
using BufferT = nanovdb::CudaDeviceBuffer;
#else
using BufferT = nanovdb::HostBuffer;
#endif
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>

using HandleT = nanovdb::GridHandle<BufferT>;
using Vec3f = nanovdb::Vec3f;
using Coord = nanovdb::Coord;

struct ref_grid {
    int rx, ry, rz, num;   // num == 0: NanoGrid<float>; else `num` NanoGrid<Vec3f>
    std::vector<HandleT> handles;
#ifdef REF_WITH_CUDA
    std::vector<void*> dev;             // device grid pointers
    nanovdb::NanoGrid<Vec3f>** dev_arr = nullptr;   // device array of device pointers (VectorGrid::cuda, plenvdb.h:285-295)
#endif
    int nleafCount = 0;
};

template <class ValueT>
static HandleT build_one(int rx, int ry, int rz, const uint8_t* active) {
    nanovdb::GridBuilder<ValueT> builder(ValueT(0.0f));
    auto acc = builder.getAccessor();
    // accessor setValue path: activates the voxel and allocates its leaf even for value 0,
    // i.e. the topology of denseFill(bbox, 0, true) / copyFromDense.
    for (int x = 0; x < rx; ++x)
        for (int y = 0; y < ry; ++y)
            for (int z = 0; z < rz; ++z)
                if (!active || active[((size_t)x * ry + y) * rz + z]) acc.setValue(Coord(x, y, z), ValueT(0.0f));
    builder.setStats(nanovdb::StatsMode::Disable);
    builder.setChecksum(nanovdb::ChecksumMode::Disable);
    return builder.template getHandle<nanovdb::AbsDiff, BufferT>(1.0, nanovdb::Vec3d(0.0), "ref");
}

extern "C" ref_grid* ref_grid_create(int rx, int ry, int rz, int channels, const uint8_t* active) {
    ref_grid* g = new ref_grid();
    g->rx = rx; g->ry = ry; g->rz = rz;
    if (channels == 1) {
        g->num = 0;
        g->handles.push_back(build_one<float>(rx, ry, rz, active));
        g->nleafCount = (int)(g->handles[0].grid<float>()->tree().nodeCount(0) << 9);
    } else {
        g->num = channels / 3;
        for (int d = 0; d < g->num; ++d) g->handles.push_back(build_one<Vec3f>(rx, ry, rz, active));
        g->nleafCount = (int)(g->handles[0].grid<Vec3f>()->tree().nodeCount(0) << 9);
    }
#ifdef REF_WITH_CUDA
    for (auto& h : g->handles) {
        h.deviceUpload();
        g->dev.push_back(g->num ? (void*)h.deviceGrid<Vec3f>() : (void*)h.deviceGrid<float>());
    }
    if (g->num) {
        cudaMalloc(&g->dev_arr, g->num * sizeof(void*));
        cudaMemcpy(g->dev_arr, g->dev.data(), g->num * sizeof(void*), cudaMemcpyHostToDevice);
    }
#endif
    return g;
}
extern "C" void ref_grid_destroy(ref_grid* g) {
#ifdef REF_WITH_CUDA
    if (g->dev_arr) cudaFree(g->dev_arr);
#endif
    delete g;
}
extern "C" int ref_grid_leaf_count(const ref_grid* g) { return g->nleafCount >> 9; }

// host views (after ref_grid_download on the GPU build)
template <class T> static const nanovdb::NanoGrid<T>* hgrid(const ref_grid* g, int d) { return g->handles[d].template grid<T>(); }
template <class T> static nanovdb::NanoGrid<T>* hgrid_mut(ref_grid* g, int d) { return g->handles[d].template grid<T>(); }

extern "C" void ref_grid_leaf_origins(const ref_grid* g, int32_t* out) {
    const int n = g->nleafCount >> 9;
    for (int i = 0; i < n; ++i) {
        Coord o = g->num ? (hgrid<Vec3f>(g, 0)->tree().getFirstNode<0>() + i)->origin()
                         : (hgrid<float>(g, 0)->tree().getFirstNode<0>() + i)->origin();
        out[i * 3] = o[0]; out[i * 3 + 1] = o[1]; out[i * 3 + 2] = o[2];
    }
}
extern "C" void ref_grid_leaf_masks(const ref_grid* g, uint64_t* out) {
    const int n = g->nleafCount >> 9;
    for (int i = 0; i < n; ++i)
        for (int w = 0; w < 8; ++w) {
            uint64_t bits = 0;
            for (int b = 0; b < 64; ++b) {
                const bool on = g->num ? (hgrid<Vec3f>(g, 0)->tree().getFirstNode<0>() + i)->isActive(w * 64 + b)
                                       : (hgrid<float>(g, 0)->tree().getFirstNode<0>() + i)->isActive(w * 64 + b);
                bits |= (uint64_t)on << b;
            }
            out[i * 8 + w] = bits;
        }
}

#ifdef REF_WITH_CUDA
extern "C" void ref_grid_download(ref_grid* g) { for (auto& h : g->handles) h.deviceDownload(); }
extern "C" void ref_grid_upload(ref_grid* g) { for (auto& h : g->handles) h.deviceUpload(); }
#else
extern "C" void ref_grid_download(ref_grid*) {}
extern "C" void ref_grid_upload(ref_grid*) {}
#endif

// Host copy dense -> active voxels through the reference leaf API (same effect as densityvdb.cu:31-49).
extern "C" void ref_grid_copy_from_dense_host(ref_grid* g, const float* dense) {
    const int n = g->nleafCount >> 9, C = g->num ? g->num * 3 : 1;
    for (int i = 0; i < n; ++i)
        for (int v = 0; v < 512; ++v) {
            if (g->num == 0) {
                auto* leaf = hgrid_mut<float>(g, 0)->tree().getFirstNode<0>() + i;
                if (!leaf->isActive(v)) continue;
                Coord c = leaf->offsetToGlobalCoord(v);
                leaf->setValueOnly(v, dense[((size_t)c[0] * g->ry + c[1]) * g->rz + c[2]]);
            } else {
                for (int d = 0; d < g->num; ++d) {
                    auto* leaf = hgrid_mut<Vec3f>(g, d)->tree().getFirstNode<0>() + i;
                    if (!leaf->isActive(v)) continue;
                    Coord c = leaf->offsetToGlobalCoord(v);
                    const float* p = dense + (((size_t)c[0] * g->ry + c[1]) * g->rz + c[2]) * C + 3 * d;
                    leaf->setValueOnly(v, Vec3f(p[0], p[1], p[2]));
                }
            }
        }
}
// Tree value at every coordinate of the box through the reference accessor (what copyToDense returns).
extern "C" void ref_grid_copy_to_dense_host(const ref_grid* g, float* dense) {
    const int C = g->num ? g->num * 3 : 1;
    if (g->num == 0) {
        auto acc = hgrid<float>(g, 0)->getAccessor();
        for (int x = 0; x < g->rx; ++x) for (int y = 0; y < g->ry; ++y) for (int z = 0; z < g->rz; ++z)
            dense[((size_t)x * g->ry + y) * g->rz + z] = acc.getValue(Coord(x, y, z));
    } else {
        for (int d = 0; d < g->num; ++d) {
            auto acc = hgrid<Vec3f>(g, d)->getAccessor();
            for (int x = 0; x < g->rx; ++x) for (int y = 0; y < g->ry; ++y) for (int z = 0; z < g->rz; ++z) {
                Vec3f v = acc.getValue(Coord(x, y, z));
                float* p = dense + (((size_t)x * g->ry + y) * g->rz + z) * C + 3 * d;
                p[0] = v[0]; p[1] = v[1]; p[2] = v[2];
            }
        }
    }
}

// For each sample: leaf index (pointer difference from the first leaf) or -1, and voxel offset, of the 8
// corners in the reference's walk order; leaf existence decided like `accumulate` (densityvdb.cu:22-26).
extern "C" void ref_probe_corners(const ref_grid* g, const float* xs, const float* ys, const float* zs, int64_t n,
                                  int32_t* corner_leaf, int32_t* corner_off) {
    static const int CORNER[8][3] = {{0, 0, 0}, {0, 0, 1}, {0, 1, 1}, {0, 1, 0}, {1, 1, 0}, {1, 1, 1}, {1, 0, 1}, {1, 0, 0}};
    using LeafT = nanovdb::NanoLeaf<float>;
    using LeafV = nanovdb::NanoLeaf<Vec3f>;
    for (int64_t s = 0; s < n; ++s) {
        Vec3f xyz(xs[s], ys[s], zs[s]);
        Coord ijk = xyz.floor();
        for (int q = 0; q < 8; ++q) {
            Coord c(ijk[0] + CORNER[q][0], ijk[1] + CORNER[q][1], ijk[2] + CORNER[q][2]);
            int leaf = -1;
            if (g->num == 0) {
                auto acc = hgrid<float>(g, 0)->getAccessor();
                acc.getValue(c);
                if (acc.isCached<LeafT>(c)) leaf = (int)(acc.getNode<LeafT>() - hgrid<float>(g, 0)->tree().getFirstNode<0>());
            } else {
                auto acc = hgrid<Vec3f>(g, 0)->getAccessor();
                acc.getValue(c);
                if (acc.isCached<LeafV>(c)) leaf = (int)(acc.getNode<LeafV>() - hgrid<Vec3f>(g, 0)->tree().getFirstNode<0>());
            }
            corner_leaf[s * 8 + q] = leaf;
            corner_off[s * 8 + q] = (int)LeafT::CoordToOffset(c);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Host execution of the PlenVDB kernel bodies through the reference accessor.  The statements are the
// reference's (densityvdb.cu:110-123, 154-165; colorvdb.cu:93-108, 141-156) with the thread index replaced
// by a loop; compiled with -ffp-contract=off, so unlike the nvcc build nothing is fused (values agree with
// the GPU to 1 ulp, integers exactly).  Used for CPU-side pinning and as the multi-threaded CPU baseline
// described in BASELINE.md §4.
// ------------------------------------------------------------------------------------------------
static void par(int64_t n, int threads, const std::function<void(int64_t, int64_t)>& fn) {
    if (threads <= 1) { fn(0, n); return; }
    std::vector<std::thread> ts;
    const int64_t chunk = (n + threads - 1) / threads;
    for (int t = 0; t < threads; ++t) { int64_t b = t * chunk, e = std::min(n, b + chunk); if (b < e) ts.emplace_back(fn, b, e); }
    for (auto& t : ts) t.join();
}
static inline void cas_add(float* addr, float v) {
    uint32_t* a = reinterpret_cast<uint32_t*>(addr);
    uint32_t old = __atomic_load_n(a, __ATOMIC_RELAXED), neu;
    do { float f; std::memcpy(&f, &old, 4); f += v; std::memcpy(&neu, &f, 4); }
    while (!__atomic_compare_exchange_n(a, &old, neu, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED));
}

extern "C" void ref_host_forward(const ref_grid* g, const float* xs, const float* ys, const float* zs, int64_t N, float* res, int threads) {
    if (g->num == 0) {
        const auto* grid = hgrid<float>(g, 0);
        par(N, threads, [&](int64_t b, int64_t e) {
            for (int64_t n = b; n < e; ++n) {
                auto acc = grid->getAccessor();
                Vec3f xyz(xs[n], ys[n], zs[n]);
                Coord ijk = xyz.floor();
                Vec3f uvw = xyz - ijk.asVec3s();
                float r = 0;
                r += acc.getValue(ijk) * (1 - uvw[0]) * (1 - uvw[1]) * (1 - uvw[2]); ijk[2] += 1;
                r += acc.getValue(ijk) * (1 - uvw[0]) * (1 - uvw[1]) * uvw[2]; ijk[1] += 1;
                r += acc.getValue(ijk) * (1 - uvw[0]) * uvw[1] * uvw[2]; ijk[2] -= 1;
                r += acc.getValue(ijk) * (1 - uvw[0]) * uvw[1] * (1 - uvw[2]); ijk[0] += 1;
                r += acc.getValue(ijk) * uvw[0] * uvw[1] * (1 - uvw[2]); ijk[2] += 1;
                r += acc.getValue(ijk) * uvw[0] * uvw[1] * uvw[2]; ijk[1] -= 1;
                r += acc.getValue(ijk) * uvw[0] * (1 - uvw[1]) * uvw[2]; ijk[2] -= 1;
                r += acc.getValue(ijk) * uvw[0] * (1 - uvw[1]) * (1 - uvw[2]);
                res[n] = r;
            }
        });
    } else {
        const int num = g->num, ndim = 3 * num;
        par(N, threads, [&](int64_t b, int64_t e) {
            for (int64_t n = b; n < e; ++n) {
                Vec3f xyz(xs[n], ys[n], zs[n]);
                Coord ijk = xyz.floor();
                Vec3f uvw = xyz - ijk.asVec3s();
                float* p = res + n * ndim;
                for (int d = 0; d < num; ++d) {
                    auto acc = hgrid<Vec3f>(g, d)->getAccessor();
                    p[0] = p[1] = p[2] = 0;
                    auto one = [&](float scale) { auto v = acc.getValue(ijk); p[0] += v[0] * scale; p[1] += v[1] * scale; p[2] += v[2] * scale; };
                    one((1 - uvw[0]) * (1 - uvw[1]) * (1 - uvw[2])); ijk[2] += 1;
                    one((1 - uvw[0]) * (1 - uvw[1]) * uvw[2]); ijk[1] += 1;
                    one((1 - uvw[0]) * uvw[1] * uvw[2]); ijk[2] -= 1;
                    one((1 - uvw[0]) * uvw[1] * (1 - uvw[2])); ijk[0] += 1;
                    one(uvw[0] * uvw[1] * (1 - uvw[2])); ijk[2] += 1;
                    one(uvw[0] * uvw[1] * uvw[2]); ijk[1] -= 1;
                    one(uvw[0] * (1 - uvw[1]) * uvw[2]); ijk[2] -= 1;
                    one(uvw[0] * (1 - uvw[1]) * (1 - uvw[2])); ijk[0] -= 1;
                    p += 3;
                }
            }
        });
    }
}

extern "C" void ref_host_backward(ref_grid* g, const float* xs, const float* ys, const float* zs, const float* grads, int64_t N, int threads) {
    using LeafT = nanovdb::NanoLeaf<float>;
    using LeafV = nanovdb::NanoLeaf<Vec3f>;
    if (g->num == 0) {
        auto* grid = hgrid_mut<float>(g, 0);
        par(N, threads, [&](int64_t b, int64_t e) {
            for (int64_t n = b; n < e; ++n) {
                auto acc = grid->getAccessor();
                Vec3f xyz(xs[n], ys[n], zs[n]);
                Coord ijk = xyz.floor();
                Vec3f uvw = xyz - ijk.asVec3s();
                auto accumulate = [&](float val) {
                    acc.getValue(ijk);
                    if (acc.isCached<LeafT>(ijk)) {
                        auto* leaf = const_cast<LeafT*>(acc.getNode<LeafT>());
                        cas_add(leaf->data()->mValues + LeafT::CoordToOffset(ijk), val);
                    }
                };
                accumulate(grads[n] * (1 - uvw[0]) * (1 - uvw[1]) * (1 - uvw[2])); ijk[2] += 1;
                accumulate(grads[n] * (1 - uvw[0]) * (1 - uvw[1]) * uvw[2]); ijk[1] += 1;
                accumulate(grads[n] * (1 - uvw[0]) * uvw[1] * uvw[2]); ijk[2] -= 1;
                accumulate(grads[n] * (1 - uvw[0]) * uvw[1] * (1 - uvw[2])); ijk[0] += 1;
                accumulate(grads[n] * uvw[0] * uvw[1] * (1 - uvw[2])); ijk[2] += 1;
                accumulate(grads[n] * uvw[0] * uvw[1] * uvw[2]); ijk[1] -= 1;
                accumulate(grads[n] * uvw[0] * (1 - uvw[1]) * uvw[2]); ijk[2] -= 1;
                accumulate(grads[n] * uvw[0] * (1 - uvw[1]) * (1 - uvw[2]));
            }
        });
    } else {
        const int num = g->num, ndim = 3 * num;
        for (int d = 0; d < num; ++d) {
            auto* grid = hgrid_mut<Vec3f>(g, d);
            par(N, threads, [&](int64_t b, int64_t e) {
                for (int64_t n = b; n < e; ++n) {
                    auto acc = grid->getAccessor();
                    const float* src = grads + n * ndim + 3 * d;
                    Vec3f xyz(xs[n], ys[n], zs[n]);
                    Coord ijk = xyz.floor();
                    Vec3f uvw = xyz - ijk.asVec3s();
                    auto accumulate = [&](float scale) {
                        acc.getValue(ijk);
                        if (acc.isCached<LeafV>(ijk)) {
                            auto* leaf = const_cast<LeafV*>(acc.getNode<LeafV>());
                            Vec3f* v = leaf->data()->mValues + LeafV::CoordToOffset(ijk);
                            cas_add(&((*v)[0]), src[0] * scale); cas_add(&((*v)[1]), src[1] * scale); cas_add(&((*v)[2]), src[2] * scale);
                        }
                    };
                    accumulate((1 - uvw[0]) * (1 - uvw[1]) * (1 - uvw[2])); ijk[2] += 1;
                    accumulate((1 - uvw[0]) * (1 - uvw[1]) * uvw[2]); ijk[1] += 1;
                    accumulate((1 - uvw[0]) * uvw[1] * uvw[2]); ijk[2] -= 1;
                    accumulate((1 - uvw[0]) * uvw[1] * (1 - uvw[2])); ijk[0] += 1;
                    accumulate(uvw[0] * uvw[1] * (1 - uvw[2])); ijk[2] += 1;
                    accumulate(uvw[0] * uvw[1] * uvw[2]); ijk[1] -= 1;
                    accumulate(uvw[0] * (1 - uvw[1]) * uvw[2]); ijk[2] -= 1;
                    accumulate(uvw[0] * (1 - uvw[1]) * (1 - uvw[2]));
                }
            });
        }
    }
}

#ifdef REF_WITH_CUDA
// ------------------------------------------------------------------------------------------------
// The reference's own CUDA kernels (unmodified objects).  Prototypes re-declared from plenvdb.h.
// ------------------------------------------------------------------------------------------------
using NanoFloatGridT = nanovdb::NanoGrid<float>;
using NanoVec3fGridT = nanovdb::NanoGrid<Vec3f>;
void density_copyFromDense(NanoFloatGridT*, float*, const int, const int, const int, const int);
void color_copyFromDense(NanoVec3fGridT**, float*, const int, const int, const int, const int, const int);
void setValuesOn_bymask_cuda(NanoFloatGridT*, bool*, const float, const int, const int, const int, const int);
void density_forward(float*, float*, float*, float*, NanoFloatGridT*, const int);
void density_backward(float*, float*, float*, float*, NanoFloatGridT*, const int);
void density_updateData(NanoFloatGridT*, NanoFloatGridT*, NanoFloatGridT*, NanoFloatGridT*, const float, const float, const float, const float, const int);
void density_updateDataWithPerlr(NanoFloatGridT*, NanoFloatGridT*, NanoFloatGridT*, NanoFloatGridT*, const float, const float, const float, const float, const int, NanoFloatGridT*);
void density_updateDataSkipGrad(NanoFloatGridT*, NanoFloatGridT*, NanoFloatGridT*, NanoFloatGridT*, const float, const float, const float, const float, const int);
void density_zero_grad(NanoFloatGridT*, const int);
void color_forward(float*, float*, float*, float*, NanoVec3fGridT**, const int, const int);
void color_backward(float*, float*, float*, float*, NanoVec3fGridT**, int, int);
void color_updateData(NanoVec3fGridT**, NanoVec3fGridT**, NanoVec3fGridT**, NanoVec3fGridT**, const float, const float, const float, const float, const int, const int);
void color_updateDataSkipGrad(NanoVec3fGridT**, NanoVec3fGridT**, NanoVec3fGridT**, NanoVec3fGridT**, const float, const float, const float, const float, const int, const int);
void color_zero_grad(NanoVec3fGridT**, const int, const int);
void render_an_image_cuda(RenderKwargs&, MLP&, SceneInfo&, float*, NanoFloatGridT*, float*, float*);

template <class T> static T* to_dev(const T* h, size_t n) { T* d; cudaMalloc(&d, n * sizeof(T)); cudaMemcpy(d, h, n * sizeof(T), cudaMemcpyHostToDevice); return d; }

extern "C" void ref_gpu_copy_from_dense(ref_grid* g, const float* dense) {   // plenvdb.h:149-157, 241-250
    const size_t n = (size_t)g->rx * g->ry * g->rz * (g->num ? 3 * g->num : 1);
    float* d = to_dev(dense, n);
    if (g->num == 0) density_copyFromDense((NanoFloatGridT*)g->dev[0], d, g->rx, g->ry, g->rz, g->nleafCount);
    else color_copyFromDense(g->dev_arr, d, g->rx, g->ry, g->rz, g->nleafCount, g->num);
    cudaFree(d);
}
// the same with the dense array already on the device (512^3 x 12 channels is 6.4 GB: built there by the caller)
extern "C" void ref_gpu_copy_from_dense_dev(ref_grid* g, float* d_dense) {
    if (g->num == 0) density_copyFromDense((NanoFloatGridT*)g->dev[0], d_dense, g->rx, g->ry, g->rz, g->nleafCount);
    else color_copyFromDense(g->dev_arr, d_dense, g->rx, g->ry, g->rz, g->nleafCount, g->num);
    cudaDeviceSynchronize();
}
extern "C" void ref_gpu_set_on_by_mask(ref_grid* g, const uint8_t* mask, float val) {   // plenvdb.h:487-495
    const size_t n = (size_t)g->rx * g->ry * g->rz;
    bool* d = to_dev(reinterpret_cast<const bool*>(mask), n);
    setValuesOn_bymask_cuda((NanoFloatGridT*)g->dev[0], d, val, g->rx, g->ry, g->rz, g->nleafCount);
    cudaFree(d);
}
// DensityVDB::forward / ColorVDB::forward host-marshalling path (plenvdb.h:441-466, 535-556)
extern "C" void ref_gpu_forward(const ref_grid* g, const float* xs, const float* ys, const float* zs, int n, float* out) {
    const int C = g->num ? 3 * g->num : 1;
    float *dx = to_dev(xs, n), *dy = to_dev(ys, n), *dz = to_dev(zs, n), *dres;
    cudaMalloc(&dres, (size_t)n * C * sizeof(float));
    if (g->num == 0) density_forward(dres, dx, dy, dz, (NanoFloatGridT*)g->dev[0], n);
    else color_forward(dres, dx, dy, dz, g->dev_arr, n, g->num);
    cudaMemcpy(out, dres, (size_t)n * C * sizeof(float), cudaMemcpyDeviceToHost);
    cudaFree(dx); cudaFree(dy); cudaFree(dz); cudaFree(dres);
}
extern "C" void ref_gpu_backward(ref_grid* g, const float* xs, const float* ys, const float* zs, const float* grads, int n) {
    const int C = g->num ? 3 * g->num : 1;
    float *dx = to_dev(xs, n), *dy = to_dev(ys, n), *dz = to_dev(zs, n), *dg = to_dev(grads, (size_t)n * C);
    if (g->num == 0) density_backward(dg, dx, dy, dz, (NanoFloatGridT*)g->dev[0], n);
    else color_backward(dg, dx, dy, dz, g->dev_arr, n, g->num);
    cudaFree(dx); cudaFree(dy); cudaFree(dz); cudaFree(dg);
}
// kernel-only timing hooks: device pointers in, no marshalling
extern "C" void ref_gpu_forward_dev(const ref_grid* g, float* dx, float* dy, float* dz, int n, float* dres) {
    if (g->num == 0) density_forward(dres, dx, dy, dz, (NanoFloatGridT*)g->dev[0], n);
    else color_forward(dres, dx, dy, dz, g->dev_arr, n, g->num);
}
extern "C" void ref_gpu_adam(ref_grid* p, ref_grid* gr, ref_grid* m, ref_grid* v, int mode, float stepsz, float eps, float b0, float b1,
                             ref_grid* perlr) {   // plenvdb.h:751-767, 774-789
    if (p->num == 0) {
        auto *P = (NanoFloatGridT*)p->dev[0], *G = (NanoFloatGridT*)gr->dev[0], *M = (NanoFloatGridT*)m->dev[0], *V = (NanoFloatGridT*)v->dev[0];
        if (mode == 2) density_updateDataWithPerlr(P, G, M, V, stepsz, eps, b0, b1, p->nleafCount, (NanoFloatGridT*)perlr->dev[0]);
        else if (mode == 1) density_updateDataSkipGrad(P, G, M, V, stepsz, eps, b0, b1, p->nleafCount);
        else density_updateData(P, G, M, V, stepsz, eps, b0, b1, p->nleafCount);
    } else {
        if (mode == 1) color_updateDataSkipGrad(p->dev_arr, gr->dev_arr, m->dev_arr, v->dev_arr, stepsz, eps, b0, b1, p->nleafCount, p->num);
        else color_updateData(p->dev_arr, gr->dev_arr, m->dev_arr, v->dev_arr, stepsz, eps, b0, b1, p->nleafCount, p->num);
    }
}
extern "C" void ref_gpu_zero_grad(ref_grid* g) {
    if (g->num == 0) density_zero_grad((NanoFloatGridT*)g->dev[0], g->nleafCount);
    else color_zero_grad(g->dev_arr, g->nleafCount, g->num);
}

// MGRenderer (plenvdb.h:933-1068) reduced to one call: set up the PODs exactly as RenderKwargs::create /
// SceneInfo::create+load / MLP::create+load do, run render_an_image_cuda, return the image and per-pixel counts.
extern "C" void ref_gpu_render(const ref_grid* idx, const float* den, const float* col, int N1, int dcol,
                               const float* w0, const float* b0, const float* w1, const float* b1, const float* w2, const float* b2,
                               const int* reso, const float* K, const float* xyz_min, const float* xyz_max,
                               float near, float stepdist, float act_shift, float interval, float thres, float bg, int inverse_y,
                               int H, int Wd, const float* c2w, float* out_rgb, int32_t* n_samples_out, float* seconds) {
    MLP mlp; mlp.Dcol = dcol; mlp.Dpe = 27; mlp.Din = dcol + 27; mlp.Dhid = 128; mlp.Dout = 3;
    mlp.w0 = to_dev(w0, (size_t)mlp.Din * 128); mlp.b0 = to_dev(b0, 128); mlp.w1 = to_dev(w1, 128 * 128); mlp.b1 = to_dev(b1, 128);
    mlp.w2 = to_dev(w2, 128 * 3); mlp.b2 = to_dev(b2, 3);
    cublasCreate(&mlp.cuHandle);
    SceneInfo scene; scene.reso = to_dev(reso, 3); scene.K = to_dev(K, 9); scene.xyz_min = to_dev(xyz_min, 3); scene.xyz_max = to_dev(xyz_max, 3);
    RenderKwargs a; a.near = near; a.far = 1e9; a.stepdist = stepdist; a.act_shift = act_shift; a.interval = interval;
    a.fast_color_thres = thres; a.bg = bg; a.inverse_y = inverse_y != 0; a.H = H; a.W = Wd; a.HW = H * Wd;
    const int HW = a.HW;
    cudaMalloc(&a.i_starts, HW * sizeof(int)); cudaMalloc(&a.i_ends, HW * sizeof(int)); cudaMalloc(&a.tmins, HW * sizeof(float));
    cudaMalloc(&a.tmaxs, HW * sizeof(float)); cudaMalloc(&a.steplens, HW * sizeof(float)); cudaMalloc(&a.pefeat, (size_t)HW * 27 * sizeof(float));
    cudaMalloc(&a.n_samples, HW * sizeof(int)); cudaMalloc(&a.rays_o, 3 * sizeof(float)); cudaMalloc(&a.rays_d, (size_t)HW * 3 * sizeof(float));
    cudaMalloc(&a.data, (size_t)HW * 3 * sizeof(float));
    float* dden = to_dev(den, N1); float* dcolp = to_dev(col, (size_t)N1 * dcol); float* dc2w = to_dev(c2w, 16);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    render_an_image_cuda(a, mlp, scene, dc2w, (NanoFloatGridT*)idx->dev[0], dden, dcolp);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    if (seconds) *seconds = ms * 1e-3f;
    cudaMemcpy(out_rgb, a.data, (size_t)HW * 3 * sizeof(float), cudaMemcpyDeviceToHost);
    if (n_samples_out) cudaMemcpy(n_samples_out, a.n_samples, HW * sizeof(int), cudaMemcpyDeviceToHost);
    cudaFree(a.i_starts); cudaFree(a.i_ends); cudaFree(a.tmins); cudaFree(a.tmaxs); cudaFree(a.steplens); cudaFree(a.pefeat);
    cudaFree(a.n_samples); cudaFree(a.rays_o); cudaFree(a.rays_d); cudaFree(a.data); cudaFree(dden); cudaFree(dcolp); cudaFree(dc2w);
    cudaFree(mlp.w0); cudaFree(mlp.b0); cudaFree(mlp.w1); cudaFree(mlp.b1); cudaFree(mlp.w2); cudaFree(mlp.b2); cublasDestroy(mlp.cuHandle);
    cudaFree(scene.reso); cudaFree(scene.K); cudaFree(scene.xyz_min); cudaFree(scene.xyz_max);
}
#endif
