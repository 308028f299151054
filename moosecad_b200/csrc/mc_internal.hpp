// Internal host-side structures behind the opaque C handles.
#pragma once
#include <cuda.h>  // CUtensorMap (types only)
#include <cuda_runtime.h>

#include <cstdint>
#include <string>
#include <vector>

#include "../../include/moosecad_b200.h"
#include "mc_error.hpp"
#include "mc_lattice.cuh"
#include "mc_tape.hpp"

namespace mc {

struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
};

// one side-set: (elem, side) pairs on the device, host copy on demand
struct SideSetDev {
    DevBuf d_pairs;
    uint64_t n = 0;
    std::vector<uint64_t> h;
    bool have_h = false;
};

}  // namespace mc

struct mc_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    uint64_t launches = 0;
    int sm_count = 148;
    size_t smem_optin = 0;
    mc::DevBuf d_warp_table;            // visit-order table of warp_vertices (k_warp_codes)
    std::vector<mc::DevBuf> free_list;  // cached device allocations
    uint64_t held_bytes = 0;

    uint64_t tiled_min_flops = 16;   // programs at least this expensive per point take the tile-specialised, sign-proving SDF path (env MC_TILED_MIN_FLOPS)
    uint32_t tile_program_cap_words = MC_TILE_PROGRAM_CAP_WORDS;  // env MC_TILE_PROGRAM_CAP_WORDS overrides (tuning)
    bool tet_emit_attr_set = false;     // cudaFuncSetAttribute for k_tet_emit done on this device
    bool boundary_attr_set = false;     // ... and for k_boundary_sides
    bool classify_attr_set = false;     // ... and for k_classify_cells
    mc::DevBuf d_shapes;                // shape tables (mc_shapes.hpp), uploaded at context creation
    bool debug = false, debug_sync = false;  // MC_DEBUG=1: poison allocations + check every stage (2: poison, 3: checks)

    int alloc(size_t bytes, mc::DevBuf& out);          // pool allocation (+ poison fill in debug mode)
    int alloc_raw(size_t bytes, mc::DevBuf& out);
    int stage_check(const char* stage);                // debug mode: synchronise + report faults by stage
    void release(mc::DevBuf& b);
    void trim();
};

struct mc_mesh {
    mc_ctx* ctx = nullptr;
    int phase = 0;  // 1 after begin, 2 after emit_vertices, 3 after emit_tets
    mc::Lattice L{};
    uint32_t c0 = 0, c1 = 0;    // owned cell layers [c0, c1)
    uint32_t cl0 = 0, cl1 = 0;  // classified cell layers
    uint32_t oz0 = 0, oz1 = 0;  // owned point planes [oz0, oz1]
    uint32_t wz0 = 0, wz1 = 0;  // planes with warp codes
    double resolution = 0, error_threshold = 0;
    uint32_t flags = 0, max_bisect = 200;
    mc::Program prog;

    alignas(64) CUtensorMap tmap_phi;  // {NX, NY, planes} f64 tensor over d_phi, box 18 x 18 x 9 (mc_tma.cuh)
    mc::TileGrid g{};
    mc::DevBuf d_prog, d_phi, d_wcode, d_vmask, d_vloc, d_cinfo, d_cshape;
    mc::DevBuf d_tflag, d_vuni, d_cuni, d_tsign, d_sub;
    mc::DevBuf d_ttile_tot, d_ttile_base, d_vtile_tot, d_vtile_vbase, d_vtile_cbase;
    mc::DevBuf own_lower_table, own_lower_coords;  // copies pulled from the slab below (mc_tessellate_multi)
    const void* lower_table = nullptr;  // id table of plane c0 from the rank below (caller-owned)
    const double* lower_coords = nullptr;  // coordinates of the rank below's top tile layer (caller-owned)
    uint64_t lower_first = 0, lower_count = 0;  // global ids [first, first + count) of that slice
    uint64_t top_layer_first = 0;  // local index of the first vertex of the tile layer holding plane c1
    bool tiled_sdf = false;
    bool vertices_local_done = false;  // mc_tessellate_emit_vertices_local has run
    mc::DevBuf d_counters, d_cutlist, d_coords, d_tets, d_plane_table;
    mc::DevBuf d_inst_start, d_word_inst, d_dec;  // tile specialisation (mc_prune.cuh)
    mc::DevBuf d_lists;                           // work lists of the tile kernels: 4 x n_tiles uint32 (TileLists)
    mc::DevBuf d_scan_sums;                       // block sums of the tile scans (k_scan_tiles_local / _add)
    mc::DevBuf d_tile_list;                       // tiles the SDF kernel evaluates (the others are proven)
    uint64_t n_tiles = 0;
    double avg_pruned_words = 0;

    uint64_t counters[40] = {0};
    uint64_t n_verts = 0, n_cuts = 0, n_tets = 0, n_kept = 0;
    uint64_t points_evaluated = 0, points_owned = 0;
    uint64_t vertex_base = 0, tet_base = 0;
    int index_bytes = 4;
    uint64_t dim[3] = {0, 0, 0};

    cudaEvent_t ev[12] = {nullptr};
    double stage_ms[9] = {0};

    // host copies made on demand
    std::vector<double> h_coords;
    std::vector<uint64_t> h_tets;
    bool have_h_coords = false, have_h_tets = false;

    // boundary / side-sets (mc_boundary.cu)
    bool have_boundary = false, have_h_sides = false;
    mc::DevBuf d_sides;  // (elem, side) pairs, uint64 x 2, sorted
    uint64_t n_sides = 0;
    std::vector<uint64_t> h_sides;
    mc::DevBuf d_side_flag, d_side_nodes;  // per-vertex flags / list of the surface vertices (side-sets)
    std::vector<mc::SideSetDev> sidesets;
};

// Mesh::to_boundary result: a triangle mesh (mesh/src/lib.rs:359-386)
struct mc_surface {
    mc_ctx* ctx = nullptr;
    uint64_t n_verts = 0, n_tris = 0;
    mc::DevBuf d_coords, d_tris;       // xyz f64; 3 x uint64 node ids per triangle
    std::vector<double> h_coords;
    std::vector<uint64_t> h_tris;
};

namespace mc {
// pooled scratch buffers that go back to the context's cache on every path out of a scope
struct PoolScope {
    mc_ctx* c;
    std::vector<DevBuf*> bufs;
    explicit PoolScope(mc_ctx* ctx) : c(ctx) {}
    PoolScope(const PoolScope&) = delete;
    PoolScope& operator=(const PoolScope&) = delete;
    int alloc(size_t bytes, DevBuf& b) { bufs.push_back(&b); return c->alloc(bytes, b); }
    void adopt(DevBuf& b) { bufs.push_back(&b); }
    ~PoolScope() { for (DevBuf* b : bufs) c->release(*b); }
};
int cuda_fail(cudaError_t e, const char* what);
#define MC_CUDA(call)                                              \
    do {                                                           \
        cudaError_t e_ = (call);                                   \
        if (e_ != cudaSuccess) return mc::cuda_fail(e_, #call);    \
    } while (0)

// upload a compiled program into a pooled device buffer
int upload_program(mc_ctx* ctx, const Program& p, DevBuf& out);
// exclusive scan of packed per-tile totals (mc_mesher.cu)
int scan_tiles(mc_ctx* c, const void* tot, uint64_t n, void* base_lo, void* base_hi, void* sums, void* totals);
// boundary + side-sets implementation (mc_boundary.cu)
int compute_boundary(mc_mesh* m);
int boundary_host_view(mc_mesh* m);
int compute_sideset(mc_mesh* m, const mc_tape* tape, int32_t root, SideSetDev& out);
int compute_surface(mc_mesh* m, mc_surface* out);
}  // namespace mc
