// topology.cpp — host-side builder of the shared sparse-tree topology.
//
// Produces the same node set and the same node ORDER that the reference obtains from OpenVDB
// denseFill / copyFromDense followed by openToNanoVDB (plenvdb/lib/vdb/plenvdb.h:117-125, 197-210;
// openvdb/nanovdb/nanovdb/util/OpenToNanoVDB.h:521-533): upper nodes in root order, lower nodes
// depth-first by increasing child offset inside each upper node, leaves depth-first by increasing child
// offset inside each lower node.  Only non-negative index space is built (the reference's grids live in
// [0, reso)).  The result is exported as flat int32 tables (see pvdb_tree in include/plenvdb_b200.h).
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>
#include <vector>

#include "../../include/plenvdb_b200.h"

static thread_local char g_err[512] = "";
static thread_local int g_launches = 0;

void pvdb_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void pvdb_count_launch(int n) { g_launches += n; }
void pvdb_reset_launch_count() { g_launches = 0; }

extern "C" const char* pvdb_last_error(void) { return g_err; }
extern "C" int pvdb_abi_version(void) { return 1; }
extern "C" int pvdb_last_launch_count(void) { return g_launches; }

struct pvdb_topo {
    std::vector<uint64_t> root_keys;
    std::vector<int32_t> upper_child;   // [n_upper][32768]
    std::vector<int32_t> lower_child;   // [n_lower][4096]
    std::vector<int32_t> leaf_origin;   // [n_leaf][3]
    std::vector<uint64_t> leaf_mask;    // [n_leaf][8]
};

namespace {

inline uint64_t root_key(int x, int y, int z) {   // NanoVDB.h:2702-2709
    return (uint64_t)((uint32_t)z >> 12) | ((uint64_t)((uint32_t)y >> 12) << 21) | ((uint64_t)((uint32_t)x >> 12) << 42);
}

// `block_words(bx,by,bz,w)` fills the 8 value-mask words of the 8^3 block and returns true if the block
// becomes a leaf.
template <class BlockFn>
pvdb_topo* build(int rx, int ry, int rz, BlockFn block_words) {
    if (rx <= 0 || ry <= 0 || rz <= 0) {
        pvdb_set_error("pvdb_topo: resolution must be positive (got %d,%d,%d)", rx, ry, rz);
        return nullptr;
    }
    pvdb_topo* t = new (std::nothrow) pvdb_topo();
    if (!t) { pvdb_set_error("pvdb_topo: out of memory"); return nullptr; }
    const int nbx = (rx + 7) / 8, nby = (ry + 7) / 8, nbz = (rz + 7) / 8;        // leaf blocks
    const int nlx = (rx + 127) / 128, nly = (ry + 127) / 128, nlz = (rz + 127) / 128;  // lower nodes
    const int nux = (rx + 4095) / 4096, nuy = (ry + 4095) / 4096, nuz = (rz + 4095) / 4096;
    uint64_t w[8];
    for (int ux = 0; ux < nux; ++ux)
        for (int uy = 0; uy < nuy; ++uy)
            for (int uz = 0; uz < nuz; ++uz) {
                int upper = -1;
                // children of the upper node in increasing offset (x<<10 | y<<5 | z)
                for (int lx = ux * 32; lx < std::min(nlx, ux * 32 + 32); ++lx)
                    for (int ly = uy * 32; ly < std::min(nly, uy * 32 + 32); ++ly)
                        for (int lz = uz * 32; lz < std::min(nlz, uz * 32 + 32); ++lz) {
                            int lower = -1;
                            for (int bx = lx * 16; bx < std::min(nbx, lx * 16 + 16); ++bx)
                                for (int by = ly * 16; by < std::min(nby, ly * 16 + 16); ++by)
                                    for (int bz = lz * 16; bz < std::min(nbz, lz * 16 + 16); ++bz) {
                                        if (!block_words(bx, by, bz, w)) continue;
                                        if (upper < 0) {
                                            upper = (int)t->root_keys.size();
                                            t->root_keys.push_back(root_key(ux * 4096, uy * 4096, uz * 4096));
                                            t->upper_child.resize((size_t)(upper + 1) * 32768, -1);
                                        }
                                        if (lower < 0) {
                                            lower = (int)(t->lower_child.size() / 4096);
                                            t->lower_child.resize((size_t)(lower + 1) * 4096, -1);
                                            const int uoff = ((lx & 31) << 10) | ((ly & 31) << 5) | (lz & 31);
                                            t->upper_child[(size_t)upper * 32768 + uoff] = lower;
                                        }
                                        const int leaf = (int)(t->leaf_origin.size() / 3);
                                        const int loff = ((bx & 15) << 8) | ((by & 15) << 4) | (bz & 15);
                                        t->lower_child[(size_t)lower * 4096 + loff] = leaf;
                                        t->leaf_origin.push_back(bx * 8);
                                        t->leaf_origin.push_back(by * 8);
                                        t->leaf_origin.push_back(bz * 8);
                                        t->leaf_mask.insert(t->leaf_mask.end(), w, w + 8);
                                    }
                        }
            }
    return t;
}

}  // namespace

extern "C" pvdb_topo* pvdb_topo_create_dense(int rx, int ry, int rz) {
    return build(rx, ry, rz, [=](int bx, int by, int bz, uint64_t* w) {
        for (int i = 0; i < 8; ++i) w[i] = 0;
        for (int dx = 0; dx < 8; ++dx) {
            if (bx * 8 + dx >= rx) break;
            for (int dy = 0; dy < 8; ++dy) {
                if (by * 8 + dy >= ry) break;
                for (int dz = 0; dz < 8; ++dz) {
                    if (bz * 8 + dz >= rz) break;
                    const int n = (dx << 6) | (dy << 3) | dz;   // NanoVDB.h:3893-3900
                    w[n >> 6] |= 1ull << (n & 63);
                }
            }
        }
        return true;
    });
}

extern "C" pvdb_topo* pvdb_topo_create_from_mask(const uint8_t* active, int rx, int ry, int rz) {
    if (!active) { pvdb_set_error("pvdb_topo_create_from_mask: null mask"); return nullptr; }
    return build(rx, ry, rz, [=](int bx, int by, int bz, uint64_t* w) {
        bool any = false;
        for (int i = 0; i < 8; ++i) w[i] = 0;
        for (int dx = 0; dx < 8; ++dx) {
            const int x = bx * 8 + dx;
            if (x >= rx) break;
            for (int dy = 0; dy < 8; ++dy) {
                const int y = by * 8 + dy;
                if (y >= ry) break;
                const uint8_t* row = active + ((size_t)x * ry + y) * rz;
                for (int dz = 0; dz < 8; ++dz) {
                    const int z = bz * 8 + dz;
                    if (z >= rz) break;
                    if (row[z]) {
                        const int n = (dx << 6) | (dy << 3) | dz;
                        w[n >> 6] |= 1ull << (n & 63);
                        any = true;
                    }
                }
            }
        }
        return any;
    });
}

extern "C" void pvdb_topo_destroy(pvdb_topo* t) { delete t; }

extern "C" int pvdb_topo_counts(const pvdb_topo* t, int32_t* n_upper, int32_t* n_lower, int32_t* n_leaf) {
    if (!t) { pvdb_set_error("pvdb_topo_counts: null topology"); return PVDB_ERR_ARG; }
    if (n_upper) *n_upper = (int32_t)t->root_keys.size();
    if (n_lower) *n_lower = (int32_t)(t->lower_child.size() / 4096);
    if (n_leaf) *n_leaf = (int32_t)(t->leaf_origin.size() / 3);
    return PVDB_OK;
}

extern "C" int pvdb_topo_export(const pvdb_topo* t, uint64_t* root_keys, int32_t* upper_child, int32_t* lower_child,
                                int32_t* leaf_origin, uint64_t* leaf_mask) {
    if (!t) { pvdb_set_error("pvdb_topo_export: null topology"); return PVDB_ERR_ARG; }
    if (root_keys) std::memcpy(root_keys, t->root_keys.data(), t->root_keys.size() * sizeof(uint64_t));
    if (upper_child) std::memcpy(upper_child, t->upper_child.data(), t->upper_child.size() * sizeof(int32_t));
    if (lower_child) std::memcpy(lower_child, t->lower_child.data(), t->lower_child.size() * sizeof(int32_t));
    if (leaf_origin) std::memcpy(leaf_origin, t->leaf_origin.data(), t->leaf_origin.size() * sizeof(int32_t));
    if (leaf_mask) std::memcpy(leaf_mask, t->leaf_mask.data(), t->leaf_mask.size() * sizeof(uint64_t));
    return PVDB_OK;
}
