// C ABI — device context, point evaluation and the mesher (include/moosecad_b200.h).
// Host orchestration of the kernels in mc_kernels.cuh.  No CPU fallback anywhere: every
// entry point here launches CUDA kernels or fails with MC_ERR_CUDA.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>

#include "mc_internal.hpp"
#include "mc_kernels.cuh"
#include "mc_prune.cuh"

using namespace mc;

static_assert(CN_COUNT <= 40, "mc_mesh::counters holds 40 entries, the device slots above are work cursors");
// d_counters has 64 slots: the upper ones are the work cursors of the list kernels, one per launch site
enum Cursor : int { CUR_WARP = 40, CUR_CLASSIFY, CUR_VCOUNT, CUR_VEMIT_INT, CUR_VEMIT, CUR_BISECT, CUR_TET_FULL, CUR_TET_SURF };

namespace mc {

int cuda_fail(cudaError_t e, const char* what) {
    return fail(MC_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}

int upload_program(mc_ctx* ctx, const Program& p, DevBuf& out) {
    int rc = ctx->alloc(p.words.size() * 8, out);
    if (rc != MC_OK) return rc;
    MC_CUDA(cudaMemcpyAsync(out.p, p.words.data(), p.words.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
    // the host vector must outlive the copy: pageable memcpy is staged before returning
    return MC_OK;
}

// Visit order of warp_vertices around one vertex (tessellate3d/src/lib.rs:146-173): tets in
// cut_lattice order (8 cells x-major, Tile::TETS order), edges in Tet4::edge order on the flipped
// node order.  It only depends on the parity of the vertex' coordinates.  Entry layout: see
// k_warp_codes.
static void build_warp_table(std::vector<unsigned long long>& tab) {
    static const int LAT[8][3] = {{0, 0, 0}, {1, 0, 0}, {1, 1, 0}, {0, 1, 0}, {0, 0, 1}, {1, 0, 1}, {1, 1, 1}, {0, 1, 1}};
    static const int TETS[5][4] = {{0, 1, 5, 2}, {0, 2, 5, 7}, {0, 7, 5, 4}, {2, 6, 5, 7}, {0, 2, 7, 3}};
    static const int EDGE[6][2] = {{0, 1}, {0, 2}, {0, 3}, {1, 2}, {1, 3}, {2, 3}};
    static const int DIR[18][3] = {{1, 0, 0},  {0, 1, 0},  {0, 0, 1},  {1, 1, 0},   {-1, 1, 0}, {1, 0, 1},  {-1, 0, 1}, {0, 1, 1},  {0, -1, 1},
                                   {-1, 0, 0}, {0, -1, 0}, {0, 0, -1}, {-1, -1, 0}, {1, -1, 0}, {-1, 0, -1}, {1, 0, -1}, {0, -1, -1}, {0, 1, -1}};
    tab.assign(8 * (WT_STRIDE + WP_STRIDE), 0ull);
    for (int pc = 0; pc < 8; ++pc) {
        const int par[3] = {pc & 1, (pc >> 1) & 1, (pc >> 2) & 1};
        int n = 0;
        for (int ax = -1; ax <= 0; ++ax)
            for (int ay = -1; ay <= 0; ++ay)
                for (int az = -1; az <= 0; ++az) {
                    const int off[3] = {ax, ay, az};
                    bool mir[3];
                    for (int k = 0; k < 3; ++k) mir[k] = ((par[k] + (off[k] ? 1 : 0)) & 1) != 0;  // cell coordinate odd
                    const bool flip = mir[0] ^ mir[1] ^ mir[2];
                    // offset of tile corner k relative to the vertex
                    auto corner_off = [&](int k, int o[3]) {
                        for (int q = 0; q < 3; ++q) o[q] = off[q] + (mir[q] ? 1 - LAT[k][q] : LAT[k][q]);
                    };
                    for (int t = 0; t < 5; ++t) {
                        int node[4];
                        for (int m = 0; m < 4; ++m) node[m] = TETS[t][(flip && m < 2) ? 1 - m : m];
                        int me = -1;
                        for (int m = 0; m < 4; ++m) {
                            int o[3];
                            corner_off(node[m], o);
                            if (o[0] == 0 && o[1] == 0 && o[2] == 0) me = m;
                        }
                        if (me < 0) continue;
                        unsigned long long corners = 0ull;
                        for (int m = 0; m < 4; ++m) {
                            int o[3];
                            corner_off(node[m], o);
                            corners |= (unsigned long long)((o[0] + 1) | ((o[1] + 1) << 2) | ((o[2] + 1) << 4)) << (6 * m);
                        }
                        for (int e = 0; e < 6; ++e) {
                            const int i0 = EDGE[e][0], i1 = EDGE[e][1];
                            if (i0 != me && i1 != me) continue;
                            int o[3];
                            corner_off(node[i0 == me ? i1 : i0], o);
                            int d = -1;
                            for (int q = 0; q < 18; ++q)
                                if (DIR[q][0] == o[0] && DIR[q][1] == o[1] && DIR[q][2] == o[2]) d = q;
                            const unsigned long long orient = i0 == me ? 0ull : 1ull;
                            tab[pc * WT_STRIDE + 1 + n++] = (unsigned long long)((ax ? 1 : 0) | (ay ? 2 : 0) | (az ? 4 : 0)) |
                                                            ((unsigned long long)d << 3) | (orient << 8) | (corners << 9);
                        }
                    }
                }
        tab[pc * WT_STRIDE] = (unsigned long long)n;
        // distinct (direction, orientation) pairs in the order of their first visit: the whole visit
        // order of a vertex whose 8 cells all exist and whose candidate tets are all kept
        unsigned long long* pairs = tab.data() + 8 * WT_STRIDE + pc * WP_STRIDE;
        unsigned long long seen = 0ull;
        int np = 0;
        for (int i = 1; i <= n; ++i) {
            const unsigned long long e = tab[pc * WT_STRIDE + i];
            const unsigned bit = (unsigned)((e >> 3) & 31ull) * 2u + (unsigned)((e >> 8) & 1ull);
            if ((seen >> bit) & 1ull) continue;
            seen |= 1ull << bit;
            pairs[1 + np++] = (unsigned long long)bit;
        }
        pairs[0] = (unsigned long long)np;
    }
}

static inline unsigned blocks_for(uint64_t n, unsigned per_block) { return (unsigned)((n + per_block - 1) / per_block); }

// Tensor map of the lattice field for the TMA staging of the tile kernels (mc_tma.cuh): a 3-D f64 tensor
// {NX, NY, stored planes} with the padded row pitch, box BOX_W x BOX_H x BOX_HZ.  The driver's encoder is
// looked up at run time (no link dependency on libcuda).
int make_phi_tensor_map(mc_mesh* m) {
    typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeTiled encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
        if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn)
            return fail(MC_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
        encode = (EncodeTiled)fn;
    }
    const Lattice& L = m->L;
    const cuuint64_t dims[3] = {L.NX, L.NY, (cuuint64_t)(L.pz1 - L.pz0 + 1)};
    const cuuint64_t strides[2] = {(cuuint64_t)L.PX * 8, (cuuint64_t)L.PX * L.NY * 8};
    const cuuint32_t box[3] = {(cuuint32_t)BOX_W, (cuuint32_t)BOX_H, (cuuint32_t)BOX_HZ};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = encode(&m->tmap_phi, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, m->d_phi.p, dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(MC_ERR_CUDA, "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
    return MC_OK;
}

// exclusive scan of n packed per-tile totals (k_scan_tiles_local / _add); `sums` holds 2 words per SCAN_BLOCK entries
int scan_tiles(mc_ctx* c, const void* tot, uint64_t n, void* base_lo, void* base_hi, void* sums, void* totals) {
    if (n == 0) {
        MC_CUDA(cudaMemsetAsync(totals, 0, 16, c->stream));
        return MC_OK;
    }
    const unsigned nb = blocks_for(n, SCAN_BLOCK);
    k_scan_tiles_local<<<nb, SCAN_THREADS, 0, c->stream>>>((const unsigned long long*)tot, n, (unsigned long long*)base_lo,
                                                           (unsigned long long*)base_hi, (unsigned long long*)sums);
    k_scan_tiles_add<<<nb, SCAN_THREADS, 0, c->stream>>>(n, (unsigned long long*)base_lo, (unsigned long long*)base_hi,
                                                         (const unsigned long long*)sums, (unsigned long long*)totals);
    c->launches += 2;
    MC_CUDA(cudaGetLastError());
    return MC_OK;
}

}  // namespace mc

// ---- pooled device memory ---------------------------------------------------------------
int mc_ctx::alloc_raw(size_t bytes, DevBuf& out) {
    if (bytes == 0) bytes = 8;
    bytes = (bytes + 255) & ~(size_t)255;
    // Size classes (eight per power of two, <= 12.5 % slack) above 1 MB: a job that comes back with slightly
    // different array sizes (another slab of the same lattice, a refined partition) reuses the cached
    // buffers instead of going to cudaMalloc / growing the cache.
    if (bytes > (1u << 20)) {
        size_t step = (size_t)1 << 17;
        while ((step << 4) < bytes) step <<= 1;   // step = 2^(floor(log2(bytes - 1)) - 3)
        bytes = (bytes + step - 1) / step * step;
    }
    int best = -1;
    for (int i = 0; i < (int)free_list.size(); ++i)
        if (free_list[i].bytes >= bytes && (best < 0 || free_list[i].bytes < free_list[best].bytes)) best = i;
    if (best >= 0 && free_list[best].bytes <= bytes + bytes / 4 + (1u << 20)) {
        out = free_list[best];
        free_list.erase(free_list.begin() + best);
        return MC_OK;
    }
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) {
        // release the cache and retry once
        trim();
        cudaGetLastError();
        e = cudaMalloc(&p, bytes);
        if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc");
    }
    held_bytes += bytes;
    out.p = p;
    out.bytes = bytes;
    return MC_OK;
}
// MC_DEBUG=1 in the environment: every buffer handed out is filled with 0xFF first (reads of
// memory no kernel wrote show up as NaN / huge indices instead of stale-but-plausible data) and
// every stage is synchronised and checked, so that a fault is reported with the stage's name.
int mc_ctx::alloc(size_t bytes, mc::DevBuf& out) {
    int rc = alloc_raw(bytes, out);
    if (rc == MC_OK && debug) {
        cudaError_t e = cudaMemsetAsync(out.p, 0xFF, out.bytes, stream);
        if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(poison)");
    }
    return rc;
}
int mc_ctx::stage_check(const char* stage) {
    if (!debug_sync) return MC_OK;
    cudaError_t e = cudaStreamSynchronize(stream);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, stage);
    return MC_OK;
}
void mc_ctx::release(DevBuf& b) {
    if (b.p) free_list.push_back(b);
    b = DevBuf();
}
void mc_ctx::trim() {
    for (auto& b : free_list) {
        cudaFree(b.p);
        held_bytes -= b.bytes;
    }
    free_list.clear();
}

static void mesh_release_all(mc_mesh* m) {
    mc_ctx* c = m->ctx;
    DevBuf* bufs[] = {&m->d_prog, &m->d_phi, &m->d_wcode, &m->d_vmask, &m->d_vloc, &m->d_cinfo, &m->d_cshape, &m->d_ttile_tot,
                      &m->d_ttile_base, &m->d_vtile_tot, &m->d_vtile_vbase, &m->d_vtile_cbase, &m->d_counters,
                      &m->d_tflag, &m->d_vuni, &m->d_cuni, &m->d_tsign,
                      &m->d_cutlist, &m->d_coords, &m->d_tets, &m->d_plane_table, &m->d_sides,
                      &m->d_inst_start, &m->d_word_inst, &m->d_dec, &m->d_sub, &m->d_tile_list, &m->d_scan_sums, &m->d_lists};
    for (DevBuf* b : bufs) c->release(*b);
    c->release(m->d_side_flag);
    c->release(m->d_side_nodes);
    c->release(m->own_lower_table);
    c->release(m->own_lower_coords);
    for (auto& ss : m->sidesets) c->release(ss.d_pairs);
    for (auto& e : m->ev)
        if (e) { cudaEventDestroy(e); e = nullptr; }
}

extern "C" {

// ---- context ----------------------------------------------------------------------------
mc_ctx* mc_ctx_new(int device, void* stream) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        fail(MC_ERR_CUDA, std::string("mc_ctx_new: no CUDA device (") + cudaGetErrorString(e) + "); this library has no CPU fallback");
        return nullptr;
    }
    if (device < 0 || device >= n) { fail(MC_ERR_ARG, "mc_ctx_new: bad device ordinal"); return nullptr; }
    if (cudaSetDevice(device) != cudaSuccess) { fail(MC_ERR_CUDA, "mc_ctx_new: cudaSetDevice failed"); return nullptr; }
    mc_ctx* c = new mc_ctx();
    { const char* dbg = std::getenv("MC_DEBUG"); const int lvl = dbg ? std::atoi(dbg) : 0; c->debug = lvl == 1 || lvl == 2; c->debug_sync = lvl == 1 || lvl == 3; }
    { const char* cap = std::getenv("MC_TILE_PROGRAM_CAP_WORDS"); if (cap && std::atoi(cap) > 0) c->tile_program_cap_words = (uint32_t)std::atoi(cap); }
    { const char* mf = std::getenv("MC_TILED_MIN_FLOPS"); if (mf && std::atoi(mf) >= 0) c->tiled_min_flops = (uint64_t)std::atoi(mf); }
    c->device = device;
    c->stream = (cudaStream_t)stream;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) {
        c->sm_count = prop.multiProcessorCount;
        c->smem_optin = prop.sharedMemPerBlockOptin;
    }
    {
        std::vector<unsigned long long> tab;
        build_warp_table(tab);
        void* p = nullptr;
        if (cudaMalloc(&p, tab.size() * 8) != cudaSuccess ||
            cudaMemcpy(p, tab.data(), tab.size() * 8, cudaMemcpyHostToDevice) != cudaSuccess) {
            fail(MC_ERR_CUDA, "mc_ctx_new: cannot upload the warp visit table");
            delete c;
            return nullptr;
        }
        c->d_warp_table.p = p;
        c->d_warp_table.bytes = tab.size() * 8;
    }
    {
        static ShapeTables tables;   // 28 KB, the same for every context
        static const bool ok = build_shape_tables(tables);
        void* p = nullptr;
        if (!ok || cudaMalloc(&p, sizeof(ShapeTables)) != cudaSuccess ||
            cudaMemcpy(p, &tables, sizeof(ShapeTables), cudaMemcpyHostToDevice) != cudaSuccess) {
            fail(ok ? MC_ERR_CUDA : MC_ERR_STATE, "mc_ctx_new: cannot build / upload the shape tables");
            if (p) cudaFree(p);
            cudaFree(c->d_warp_table.p);
            delete c;
            return nullptr;
        }
        c->d_shapes.p = p;
        c->d_shapes.bytes = sizeof(ShapeTables);
    }
    return c;
}
void mc_ctx_free(mc_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    c->trim();
    if (c->d_warp_table.p) cudaFree(c->d_warp_table.p);
    if (c->d_shapes.p) cudaFree(c->d_shapes.p);
    delete c;
}
int mc_ctx_device(const mc_ctx* c) { return c ? c->device : MC_ERR_ARG; }
uint64_t mc_ctx_launch_count(const mc_ctx* c) { return c ? c->launches : 0; }
uint64_t mc_ctx_workspace_bytes(const mc_ctx* c) { return c ? c->held_bytes : 0; }
int mc_ctx_trim(mc_ctx* c) {
    if (!c) return fail(MC_ERR_ARG, "mc_ctx_trim: null context");
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    c->trim();
    return MC_OK;
}

// ---- ObjectAdaptor::value ---------------------------------------------------------------
int mc_eval_points_device(mc_ctx* c, const mc_tape* t, int32_t root, double slack, const void* d_points, uint64_t n,
                          void* d_out) {
    if (!c || !t) return fail(MC_ERR_ARG, "mc_eval_points_device: null handle");
    if (n == 0) return MC_OK;
    MC_CUDA(cudaSetDevice(c->device));
    Program p;
    std::string err;
    int rc = t->compile(root, slack, p, err);
    if (rc != MC_OK) return fail(rc, err);
    PoolScope scratch(c);
    DevBuf dp;
    scratch.adopt(dp);
    rc = upload_program(c, p, dp);
    if (rc != MC_OK) return rc;
    k_eval_points<<<blocks_for(n, 256), 256, 0, c->stream>>>((const uint64_t*)dp.p, (const double*)d_points, n,
                                                              (double*)d_out);
    c->launches++;
    MC_CUDA(cudaGetLastError());
    MC_CUDA(cudaStreamSynchronize(c->stream));  // dp is recycled when `scratch` goes out of scope
    return MC_OK;
}

int mc_eval_points(mc_ctx* c, const mc_tape* t, int32_t root, double slack, const double* points, uint64_t n,
                   double* out) {
    if (!c || !t || (n && (!points || !out))) return fail(MC_ERR_ARG, "mc_eval_points: bad argument");
    if (n == 0) return MC_OK;
    MC_CUDA(cudaSetDevice(c->device));
    PoolScope scratch(c);
    DevBuf dpts, dout;
    int rc = scratch.alloc(n * 24, dpts);
    if (rc == MC_OK) rc = scratch.alloc(n * 8, dout);
    if (rc != MC_OK) return rc;
    MC_CUDA(cudaMemcpyAsync(dpts.p, points, n * 24, cudaMemcpyHostToDevice, c->stream));
    rc = mc_eval_points_device(c, t, root, slack, dpts.p, n, dout.p);
    if (rc == MC_OK) {
        cudaError_t e = cudaMemcpyAsync(out, dout.p, n * 8, cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) rc = cuda_fail(e, "copy back");
    }
    return rc;
}

// ---- the mesher -------------------------------------------------------------------------
extern "C++" {
namespace mc {
// kernel-side view of a mesh's slab
Stuff make_stuff(const mc_mesh* m) {
    Stuff S;
    S.L = m->L;
    S.g = m->g;
    S.c0 = m->c0; S.c1 = m->c1; S.cl0 = m->cl0; S.cl1 = m->cl1;
    S.oz0 = m->oz0; S.oz1 = m->oz1; S.wz0 = m->wz0; S.wz1 = m->wz1;
    S.tflag = (uint8_t*)m->d_tflag.p; S.vuni = (uint8_t*)m->d_vuni.p; S.cuni = (uint8_t*)m->d_cuni.p;
    S.vmask = (uint16_t*)m->d_vmask.p; S.vloc = (uint16_t*)m->d_vloc.p; S.cinfo = (uint16_t*)m->d_cinfo.p;
    S.tvbase = (const unsigned long long*)m->d_vtile_vbase.p;
    S.ttbase = (const unsigned long long*)m->d_ttile_base.p;
    S.lower_table = (const unsigned long long*)m->lower_table;
    S.cshape = (unsigned long long*)m->d_cshape.p;
    S.shapes = (const ShapeTables*)m->ctx->d_shapes.p;
    S.emit_reclassify = (m->flags & MC_OPT_RECLASSIFY_EMIT) != 0;
    S.warp_table = (const unsigned long long*)m->ctx->d_warp_table.p;
    S.vertex_base = (long long)m->vertex_base;
    return S;
}
}  // namespace mc
}  // extern "C++"

static TileLists make_lists(const mc_mesh* m) {
    TileLists T;
    const uint32_t* base = (const uint32_t*)m->d_lists.p;
    const unsigned long long* cnt = (const unsigned long long*)m->d_counters.p;
    T.v_active = base; T.v_neg = base + m->n_tiles; T.c_active = base + 2 * m->n_tiles; T.c_full = base + 3 * m->n_tiles;
    T.n_v_active = cnt + CN_ACTIVE_VTILES; T.n_v_neg = cnt + CN_LIST_VNEG;
    T.n_c_active = cnt + CN_ACTIVE_CTILES; T.n_c_full = cnt + CN_LIST_CFULL;
    return T;
}
static inline unsigned list_grid(const mc_mesh* m) {
    const uint64_t g = (uint64_t)m->ctx->sm_count * TILE_GRID_PER_SM;
    return (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(g, m->n_tiles));
}

// tile-specialised evaluation (mc_prune.cuh): worth it once the program is more than a few nodes
static int launch_sdf_field_tiled(mc_mesh* m, bool* done) {
    mc_ctx* c = m->ctx;
    *done = false;
    const Program& p = m->prog;
    const uint32_t n_inst = (uint32_t)p.inst_start.size(), n_words = (uint32_t)p.words.size();
    if (n_inst >= 0xffffu || p.n_decisions == 0) return MC_OK;
    // shared-memory budget of a tile-specialised program: long programs (hundreds of primitives)
    // specialise to a small fraction of their length; the rare tile that does not fit evaluates the
    // full program from global memory
    const uint32_t cap_words = std::min<uint32_t>(n_words, c->tile_program_cap_words);
    const size_t smem = (size_t)cap_words * 8 + ((p.n_decisions + 15u) & ~15u) + (size_t)(n_inst + 1) * 8;
    const size_t smem_cap = c->smem_optin ? c->smem_optin : 48 * 1024;
    if (smem > smem_cap) return MC_OK;
    const Lattice& L = m->L;
    const uint64_t ntiles = m->n_tiles;
    int rc = c->alloc((size_t)n_inst * 4, m->d_inst_start);
    if (rc == MC_OK) rc = c->alloc((size_t)n_words * 2, m->d_word_inst);
    if (rc == MC_OK) rc = c->alloc((size_t)p.n_decisions * ntiles, m->d_dec);
    if (rc == MC_OK) rc = c->alloc(ntiles + 4, m->d_tsign);
    if (rc != MC_OK) return rc;
    MC_CUDA(cudaMemcpyAsync(m->d_inst_start.p, p.inst_start.data(), (size_t)n_inst * 4, cudaMemcpyHostToDevice, c->stream));
    MC_CUDA(cudaMemcpyAsync(m->d_word_inst.p, p.word_inst.data(), (size_t)n_words * 2, cudaMemcpyHostToDevice, c->stream));
    k_tile_decide<<<blocks_for(ntiles, 128), 128, 0, c->stream>>>(L, m->g, (const uint64_t*)m->d_prog.p, ntiles,
                                                                  (uint8_t*)m->d_dec.p, (int8_t*)m->d_tsign.p);
    c->launches++;
    MC_CUDA(cudaGetLastError());
    const bool skip_uniform = !(m->flags & MC_OPT_KEEP_PHI);
    unsigned long long* cnt = (unsigned long long*)m->d_counters.p;
    if (skip_uniform) {
        rc = c->alloc(ntiles * 8, m->d_sub);
        if (rc == MC_OK) rc = c->alloc(ntiles * 4, m->d_tile_list);
        if (rc != MC_OK) return rc;
        k_subbox_decide<<<blocks_for(ntiles, 4), 128, 4 * (size_t)((p.n_decisions + 15u) & ~15u), c->stream>>>(
            L, m->g, (const uint64_t*)m->d_prog.p, ntiles, p.n_decisions, (const uint8_t*)m->d_dec.p,
            (int8_t*)m->d_tsign.p, (uint32_t*)m->d_sub.p, cnt + CN_SUB_BOXES_PROVEN, 1u, (uint32_t*)m->d_tile_list.p,
            cnt + CN_SDF_LIST, (uint8_t*)m->d_tflag.p, cnt + CN_SKIPPED_TILES);
        c->launches++;
        MC_CUDA(cudaGetLastError());
        // from here on readers take the sign of a proven tile from the table: nothing is stored there
        m->L.tsign = (const int8_t*)m->d_tsign.p;
        m->L.tntx = m->g.ntx;
        m->L.tnty = m->g.nty;
    }
    const bool count = (m->flags & MC_OPT_COUNT_FLOPS) != 0;
    const bool small = p.small();
    const bool capped = cap_words < n_words;
    auto kern = capped ? (count ? (small ? k_sdf_field_tiled<8, 4, true, true, true> : k_sdf_field_tiled<8, 4, true, false, true>)
                                : (small ? k_sdf_field_tiled<8, 4, false, true, true> : k_sdf_field_tiled<8, 4, false, false, true>))
                       : (count ? (small ? k_sdf_field_tiled<8, 4, true, true, false> : k_sdf_field_tiled<8, 4, true, false, false>)
                                : (small ? k_sdf_field_tiled<8, 4, false, true, false> : k_sdf_field_tiled<8, 4, false, false, false>));
    if (smem > 48 * 1024) MC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    MC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, smem));
    if (per_sm < 1) per_sm = 1;
    uint64_t grid = (uint64_t)c->sm_count * per_sm;
    if (grid > ntiles) grid = ntiles;
    unsigned long long* pruned_total = (unsigned long long*)m->d_counters.p + CN_PRUNED_WORDS;
    kern<<<(unsigned)grid, 256, smem, c->stream>>>(L, m->g, (double*)m->d_phi.p, (const uint64_t*)m->d_prog.p,
                                                   (const uint32_t*)m->d_inst_start.p, (const uint16_t*)m->d_word_inst.p,
                                                   n_inst, cap_words, p.n_decisions, (const uint8_t*)m->d_dec.p, ntiles,
                                                   (const int8_t*)m->d_tsign.p, skip_uniform ? 1 : 0,
                                                   pruned_total, cnt + CN_SDF_FLOPS,
                                                   skip_uniform ? (const uint32_t*)m->d_sub.p : nullptr, (uint8_t*)m->d_tflag.p,
                                                   skip_uniform ? (const uint32_t*)m->d_tile_list.p : nullptr,
                                                   cnt + CN_SDF_LIST, cnt + CN_SDF_NEXT);
    c->launches++;
    MC_CUDA(cudaGetLastError());
    m->tiled_sdf = true;
    *done = true;
    return MC_OK;
}

static int launch_sdf_field(mc_mesh* m) {
    mc_ctx* c = m->ctx;
    if ((m->prog.flops >= c->tiled_min_flops || (m->flags & MC_OPT_FORCE_TILE_PRUNING)) && !(m->flags & MC_OPT_NO_TILE_PRUNING)) {
        bool done = false;
        int rc = launch_sdf_field_tiled(m, &done);
        if (rc != MC_OK || done) return rc;
    }
    const int nwords = (int)m->prog.words.size();
    const size_t prog_bytes = (size_t)nwords * 8;
    // stage the program in shared memory when it fits
    const size_t smem_cap = c->smem_optin ? c->smem_optin : 48 * 1024;
    const int in_smem = prog_bytes <= smem_cap ? 1 : 0;
    const size_t smem = in_smem ? prog_bytes : 0;
    const uint32_t nplanes = m->L.pz1 - m->L.pz0 + 1;
    // cheap programs are HBM-write bound: 32 consecutive x per warp; heavy ones want compact
    // warp footprints so that bbox guards are warp-coherent
    const bool linear = m->prog.flops < 64;
    const bool count = (m->flags & MC_OPT_COUNT_FLOPS) != 0;
    const bool small = m->prog.small();
    auto kern = linear ? (count ? (small ? k_sdf_field<32, 1, true, true> : k_sdf_field<32, 1, true, false>)
                                : (small ? k_sdf_field<32, 1, false, true> : k_sdf_field<32, 1, false, false>))
                       : (count ? (small ? k_sdf_field<8, 4, true, true> : k_sdf_field<8, 4, true, false>)
                                : (small ? k_sdf_field<8, 4, false, true> : k_sdf_field<8, 4, false, false>));
    if (smem > 48 * 1024) MC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    MC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 512, smem));
    if (per_sm < 1) per_sm = 1;
    const uint64_t nsteps = (uint64_t)((m->L.NX + (linear ? 31 : 7)) / (linear ? 32 : 8)) *
                            ((m->L.NY + (linear ? 0 : 3)) / (linear ? 1 : 4)) * nplanes;
    uint64_t grid = (uint64_t)c->sm_count * per_sm;
    const uint64_t need = (nsteps + 15) / 16;
    if (grid > need) grid = need ? need : 1;
    kern<<<(unsigned)grid, 512, smem, c->stream>>>(m->L, (double*)m->d_phi.p, (const uint64_t*)m->d_prog.p, nwords,
                                                   in_smem, m->L.pz0, nplanes, (unsigned long long*)m->d_counters.p + CN_SDF_FLOPS);
    c->launches++;
    MC_CUDA(cudaGetLastError());
    return MC_OK;
}

// Interval pass over the whole lattice -> per tile layer, how many tiles fall in each cost class:
// 0 sign not proven (full SDF evaluation, possibly surface work), 1 proven inside with a proven-inside
// 27-neighbourhood (no SDF evaluation, arithmetic output), 2 proven inside otherwise (shell evaluation
// + output), 3 proven outside with a proven-outside neighbourhood (nothing), 4 proven outside
// otherwise (shell evaluation).
static int tile_layer_classes(mc_ctx* c, const mc_tape* volume, int32_t root, double resolution, const char* who,
                              uint64_t dim[3], uint64_t layer_begin, uint64_t layer_end,
                              std::vector<unsigned long long>& counts, uint32_t& ntz) {
    if (!volume->valid(root) || !(resolution > 0.0)) return fail(MC_ERR_ARG, std::string(who) + ": bad argument");
    MC_CUDA(cudaSetDevice(c->device));
    const Box& bb = volume->nodes[(size_t)root].bbox;
    for (int k = 0; k < 3; ++k) {
        const double d = std::ceil((bb.hi[k] - bb.lo[k]) / resolution);
        if (!(d >= 0.0) || !(d < 4294967295.0) || !std::isfinite(bb.lo[k]))
            return fail(MC_ERR_BBOX, std::string(who) + ": the volume's bounding box is not finite");
        dim[k] = (uint64_t)d;
    }
    Program prog;
    std::string err;
    int rc = volume->compile(root, resolution, prog, err);
    if (rc != MC_OK) return fail(rc, err);
    Lattice L{};
    L.nx = (uint32_t)dim[0]; L.ny = (uint32_t)dim[1]; L.nz = (uint32_t)dim[2];
    L.NX = L.nx + 1; L.NY = L.ny + 1; L.NZ = L.nz + 1;
    L.PX = (L.NX + 15u) & ~15u;
    L.ox = bb.lo[0]; L.oy = bb.lo[1]; L.oz = bb.lo[2];
    L.res = resolution;
    L.pz0 = 0; L.pz1 = L.nz;
    TileGrid g;
    g.ntx = (L.NX + TS - 1) / TS; g.nty = (L.NY + TS - 1) / TS; g.ntz = (L.NZ + TS - 1) / TS; g.z0 = 0;
    ntz = g.ntz;
    if (layer_end > ntz) layer_end = ntz;
    if (layer_begin > layer_end) layer_begin = layer_end;
    counts.assign((size_t)ntz * MC_TILE_CLASS_STRIDE, 0ull);
    if (layer_begin == layer_end) return MC_OK;
    // the tiles of the requested layers plus one layer either side (the neighbourhood test)
    const uint32_t k_lo = layer_begin > 0 ? (uint32_t)layer_begin - 1 : 0u, k_hi = (uint32_t)std::min<uint64_t>(ntz, layer_end + 1);
    g.z0 = k_lo * TS;
    g.ntz = k_hi - k_lo;
    const uint32_t first = (uint32_t)layer_begin - k_lo, last = (uint32_t)layer_end - k_lo;  // local range that is counted
    const uint64_t ntiles = (uint64_t)g.ntx * g.nty * g.ntz;
    if (prog.n_decisions == 0) {
        // nothing to prove with: every tile is evaluated in full
        for (uint64_t k = layer_begin; k < layer_end; ++k) {
            counts[(size_t)k * MC_TILE_CLASS_STRIDE] = (unsigned long long)g.ntx * g.nty;
            counts[(size_t)k * MC_TILE_CLASS_STRIDE + 5] = (unsigned long long)g.ntx * g.nty * (unsigned long long)prog.flops;
        }
        return MC_OK;
    }
    if (ntiles >= (1ull << 31)) return fail(MC_ERR_LIMIT, std::string(who) + ": lattice too large");
    DevBuf dprog, ddec, dsign, dcnt, dwork, dsub;
    const uint32_t sample = 4;  // sub-box proofs on every 4th tile are enough for a load estimate
    rc = upload_program(c, prog, dprog);
    if (rc == MC_OK) rc = c->alloc((size_t)prog.n_decisions * ntiles, ddec);
    if (rc == MC_OK) rc = c->alloc(ntiles + 4, dsign);
    if (rc == MC_OK) rc = c->alloc((counts.size() + 1) * 8, dcnt);  // + a scratch counter
    if (rc == MC_OK) rc = c->alloc((ntiles + 1) * 4, dwork);
    if (rc == MC_OK) rc = c->alloc(ntiles * 8, dsub);
    cudaError_t e = cudaSuccess;
    if (rc == MC_OK) {
        e = cudaMemsetAsync(dcnt.p, 0, (counts.size() + 1) * 8, c->stream);
        k_tile_decide<<<blocks_for(ntiles, 128), 128, 0, c->stream>>>(L, g, (const uint64_t*)dprog.p, ntiles,
                                                                      (uint8_t*)ddec.p, (int8_t*)dsign.p, (float*)dwork.p);
        k_subbox_decide<<<blocks_for(ntiles, 4), 128, 4 * (size_t)((prog.n_decisions + 15u) & ~15u), c->stream>>>(
            L, g, (const uint64_t*)dprog.p, ntiles, prog.n_decisions, (const uint8_t*)ddec.p, (int8_t*)dsign.p,
            (uint32_t*)dsub.p, (unsigned long long*)dcnt.p + counts.size(), sample);
        k_tile_layer_classes<<<blocks_for(ntiles, 256), 256, 0, c->stream>>>(
            g, (const int8_t*)dsign.p, (const float*)dwork.p, (const uint32_t*)dsub.p, sample, ntiles, first, last,
            (unsigned long long*)dcnt.p + (size_t)k_lo * MC_TILE_CLASS_STRIDE);
        c->launches++;
        c->launches += 2;
        if (e == cudaSuccess) e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaMemcpyAsync(counts.data(), dcnt.p, counts.size() * 8, cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    }
    c->release(dprog); c->release(ddec); c->release(dsign); c->release(dcnt); c->release(dwork); c->release(dsub);
    if (rc != MC_OK) return rc;
    if (e != cudaSuccess) return cuda_fail(e, who);
    return MC_OK;
}

int mc_tile_layer_classes(mc_ctx* c, const mc_tape* volume, int32_t root, double resolution, uint64_t layer_begin,
                          uint64_t layer_end, uint64_t* counts, uint64_t n) {
    if (!c || !volume || !counts) return fail(MC_ERR_ARG, "mc_tile_layer_classes: bad argument");
    uint64_t dim[3];
    std::vector<unsigned long long> cnt;
    uint32_t ntz = 0;
    int rc = tile_layer_classes(c, volume, root, resolution, "mc_tile_layer_classes", dim, layer_begin, layer_end, cnt, ntz);
    if (rc != MC_OK) return rc;
    if (n != ntz) return fail(MC_ERR_ARG, "mc_tile_layer_classes: expected one entry per tile layer (" + std::to_string(ntz) + ")");
    for (size_t i = 0; i < cnt.size(); ++i) counts[i] = cnt[i];
    return MC_OK;
}

// relative cost of one tile of each class (fitted to per-slab device times of BASELINE config 4,
// profiles/r01_notes.md "slab balance")
static const double SLAB_CLASS_COST[MC_TILE_CLASS_STRIDE] = {
    MC_SLAB_COST_UNPROVEN, MC_SLAB_COST_NEG_SKIPPED, MC_SLAB_COST_NEG_SHELL, MC_SLAB_COST_POS_SKIPPED,
    MC_SLAB_COST_POS_SHELL, MC_SLAB_COST_UNPROVEN_WORK, MC_SLAB_COST_SHELL_WORK};

void mc_slab_class_weights(double* out) {
    if (out) for (int q = 0; q < MC_TILE_CLASS_STRIDE; ++q) out[q] = SLAB_CLASS_COST[q];
}

int mc_estimate_layer_costs(mc_ctx* c, const mc_tape* volume, int32_t root, double resolution, double* cost, uint64_t n) {
    if (!c || !volume || !cost) return fail(MC_ERR_ARG, "mc_estimate_layer_costs: bad argument");
    uint64_t dim[3];
    std::vector<unsigned long long> cnt;
    uint32_t ntz = 0;
    int rc = tile_layer_classes(c, volume, root, resolution, "mc_estimate_layer_costs", dim, 0, ~0ull, cnt, ntz);
    if (rc != MC_OK) return rc;
    if (n != dim[2]) return fail(MC_ERR_ARG, "mc_estimate_layer_costs: expected one entry per cell layer (" + std::to_string(dim[2]) + ")");
    for (uint64_t z = 0; z < n; ++z) {
        const size_t k = (size_t)(z / TS);
        double s = 0.0;
        for (int q = 0; q < MC_TILE_CLASS_STRIDE; ++q) s += SLAB_CLASS_COST[q] * (double)cnt[k * MC_TILE_CLASS_STRIDE + q];
        cost[z] = s / (double)TS;
    }
    return MC_OK;
}

int mc_tessellate_begin(mc_ctx* c, const mc_tape* volume, int32_t root, double resolution, double relative_error,
                        const mc_slab* slab, const mc_options* opts, mc_mesh** out) {
    if (!c || !volume || !out) return fail(MC_ERR_ARG, "mc_tessellate_begin: null handle");
    *out = nullptr;
    if (!volume->valid(root)) return fail(MC_ERR_ARG, "mc_tessellate_begin: invalid root node");
    if (!(resolution > 0.0)) return fail(MC_ERR_ARG, "mc_tessellate_begin: resolution must be positive");
    MC_CUDA(cudaSetDevice(c->device));

    // IsosurfaceStuffing::new (tessellate3d/src/lib.rs:36-52): dim = ceil(bbox.dim / resolution)
    const Box& bb = volume->nodes[(size_t)root].bbox;
    uint64_t dim[3];
    for (int k = 0; k < 3; ++k) {
        const double d = std::ceil((bb.hi[k] - bb.lo[k]) / resolution);
        if (!(d >= 0.0) || !(d < 4294967295.0) || !std::isfinite(bb.lo[k]))
            return fail(MC_ERR_BBOX, "mc_tessellate_begin: the volume's bounding box is not finite (or the lattice exceeds 2^32 cells per axis)");
        dim[k] = (uint64_t)d;
    }
    mc_mesh* m = new mc_mesh();
    m->ctx = c;
    for (int k = 0; k < 3; ++k) m->dim[k] = dim[k];
    m->resolution = resolution;
    m->error_threshold = resolution * relative_error;
    if (opts && opts->struct_size >= sizeof(mc_options)) {
        m->flags = opts->flags;
        if (opts->max_bisect_iters) m->max_bisect = opts->max_bisect_iters;
    }
    std::string err;
    // ObjectAdaptor::value = approx_value(p, tessellation_resolution), src/main.rs:33-35
    int rc = volume->compile(root, resolution, m->prog, err);
    if (rc != MC_OK) { delete m; return fail(rc, err); }

    Lattice& L = m->L;
    L.nx = (uint32_t)dim[0]; L.ny = (uint32_t)dim[1]; L.nz = (uint32_t)dim[2];
    L.NX = L.nx + 1; L.NY = L.ny + 1; L.NZ = L.nz + 1;
    L.PX = (L.NX + 15u) & ~15u;
    L.ox = bb.lo[0]; L.oy = bb.lo[1]; L.oz = bb.lo[2];
    L.res = resolution;
    uint64_t zb = 0, ze = dim[2];
    if (slab) { zb = slab->z_begin; ze = slab->z_end; }  // NULL = the whole lattice; an empty range owns nothing
    if (zb > ze || ze > dim[2]) { delete m; return fail(MC_ERR_ARG, "mc_tessellate_begin: bad slab range"); }
    m->c0 = (uint32_t)zb; m->c1 = (uint32_t)ze;
    const uint32_t nz = L.nz;
    L.pz0 = m->c0 >= 2 ? m->c0 - 2 : 0;
    L.pz1 = (uint32_t)std::min<uint64_t>(nz, (uint64_t)m->c1 + 2);
    m->wz0 = m->c0 >= 1 ? m->c0 - 1 : 0;
    m->wz1 = (uint32_t)std::min<uint64_t>(nz, (uint64_t)m->c1 + 1);
    m->cl0 = m->c0 >= 1 ? m->c0 - 1 : 0;
    m->cl1 = (uint32_t)std::min<uint64_t>(nz, (uint64_t)m->c1 + 1);
    m->oz0 = m->c0 == 0 ? 0 : m->c0 + 1;
    m->oz1 = m->c1;
    if (m->c0 == m->c1 && nz > 0) { m->oz0 = 1; m->oz1 = 0; }  // empty slab owns nothing

    const uint64_t per_plane = (uint64_t)L.NX * L.NY;
    const uint32_t nplanes = L.pz1 - L.pz0 + 1;
    const uint64_t npts = (uint64_t)L.PX * L.NY * nplanes;   // entries of the per-point arrays (padded rows)
    const uint64_t ncls = (uint64_t)L.nx * L.ny * (m->cl1 - m->cl0);
    const uint64_t nown = (m->oz1 >= m->oz0) ? per_plane * (m->oz1 - m->oz0 + 1) : 0;
    m->points_evaluated = per_plane * nplanes;
    m->points_owned = nown;
    m->g.ntx = (L.NX + TS - 1) / TS; m->g.nty = (L.NY + TS - 1) / TS; m->g.ntz = (nplanes + TS - 1) / TS;
    m->g.z0 = L.pz0;
    const uint64_t ntiles = (uint64_t)m->g.ntx * m->g.nty * m->g.ntz;
    m->n_tiles = ntiles;
    if (ntiles >= (1ull << 31)) { delete m; return fail(MC_ERR_LIMIT, "lattice too large for one slab"); }

#define MC_TRY(x) do { rc = (x); if (rc != MC_OK) { mesh_release_all(m); delete m; return rc; } } while (0)
#define MC_TRY_CUDA(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { mesh_release_all(m); delete m; return cuda_fail(e_, #x); } } while (0)
    for (auto& e : m->ev) MC_TRY_CUDA(cudaEventCreate(&e));
    MC_TRY(upload_program(c, m->prog, m->d_prog));
    MC_TRY(c->alloc(npts * 8, m->d_phi));
    MC_TRY(c->alloc(npts, m->d_wcode));
    MC_TRY(c->alloc(npts * 2, m->d_vmask));
    MC_TRY(c->alloc(npts * 2, m->d_vloc));
    MC_TRY(c->alloc(std::max<uint64_t>(ncls, 1) * 2, m->d_cinfo));
    MC_TRY(c->alloc(std::max<uint64_t>(ncls, 1) * 8, m->d_cshape));
    MC_TRY(c->alloc(ntiles + 4, m->d_tflag));
    MC_TRY(c->alloc(ntiles + 4, m->d_vuni));
    MC_TRY(c->alloc(ntiles + 4, m->d_cuni));
    MC_TRY(c->alloc((ntiles + 1) * 8, m->d_ttile_tot));
    MC_TRY(c->alloc((ntiles + 1) * 8, m->d_ttile_base));
    MC_TRY(c->alloc((ntiles + 1) * 8, m->d_vtile_tot));
    MC_TRY(c->alloc((ntiles + 1) * 8, m->d_vtile_vbase));
    MC_TRY(c->alloc((ntiles + 1) * 8, m->d_vtile_cbase));
    MC_TRY(c->alloc(64 * 8, m->d_counters));
    MC_TRY(c->alloc((ntiles / SCAN_BLOCK + 2) * 16, m->d_scan_sums));
    MC_TRY(c->alloc(std::max<uint64_t>(ntiles, 1) * 4 * 4, m->d_lists));
    L.phi = (const double*)m->d_phi.p;
    L.wcode = (uint8_t*)m->d_wcode.p;
    MC_TRY(make_phi_tensor_map(m));

    cudaStream_t s = c->stream;
    {
        uint64_t init[64] = {0};
        const double one = 1.0;
        std::memcpy(&init[CN_AR_MIN_BITS], &one, 8);  // aspect_ratio = (one, zero), tessellate3d/src/lib.rs:404
        MC_TRY_CUDA(cudaMemcpyAsync(m->d_counters.p, init, sizeof(init), cudaMemcpyHostToDevice, s));
    }
    MC_TRY_CUDA(cudaMemsetAsync(m->d_wcode.p, 0, npts, s));
    unsigned long long* counters = (unsigned long long*)m->d_counters.p;
    const unsigned tgrid = (unsigned)ntiles;

    // a2: SDF field on all stored planes
    MC_TRY_CUDA(cudaEventRecord(m->ev[0], s));
    MC_TRY(launch_sdf_field(m));
    const Stuff S = make_stuff(m);  // after the SDF stage: it decides whether the tile-sign table is in use
    MC_TRY(c->stage_check("stage sdf_field"));
    MC_TRY_CUDA(cudaEventRecord(m->ev[1], s));
    // tile sign summary (the tiled SDF kernel writes it itself) + a6: warp codes
    const TileLists T = make_lists(m);
    const unsigned lgrid = list_grid(m);
    if (!m->tiled_sdf) k_tile_minmax<<<tgrid, TILE_THREADS, 0, s>>>(S);
    k_tile_neighbourhood<<<blocks_for(ntiles, 256), 256, 0, s>>>(S, ntiles, counters, (uint32_t*)T.v_active, (uint32_t*)T.v_neg,
                                                                 (unsigned long long*)m->d_vtile_tot.p);
    MC_TRY(c->stage_check("stage tile summary"));
    k_warp_codes<<<lgrid, TILE_THREADS, 0, s>>>(S, m->tmap_phi, counters, T.v_active, T.n_v_active, counters + CUR_WARP);
    c->launches += m->tiled_sdf ? 2 : 3;
    MC_TRY_CUDA(cudaGetLastError());
    MC_TRY(c->stage_check("stage warp codes"));
    MC_TRY_CUDA(cudaEventRecord(m->ev[2], s));
    // a5/a7/a9: classification + counts
    k_cell_tile_class<<<blocks_for(ntiles, 256), 256, 0, s>>>(S, ntiles, counters, (uint32_t*)T.c_active, (uint32_t*)T.c_full,
                                                              (unsigned long long*)m->d_ttile_tot.p);
    if (!c->classify_attr_set) {
        MC_TRY_CUDA(cudaFuncSetAttribute(k_classify_cells, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CLASSIFY_SMEM));
        c->classify_attr_set = true;
    }
    k_classify_cells<<<lgrid, TILE_THREADS, CLASSIFY_SMEM, s>>>(S, (const uint64_t*)m->d_prog.p, (unsigned long long*)m->d_ttile_tot.p,
                                                    counters, T.c_active, T.n_c_active, counters + CUR_CLASSIFY);
    c->launches += 2;
    MC_TRY_CUDA(cudaGetLastError());
    MC_TRY(scan_tiles(c, m->d_ttile_tot.p, ntiles, m->d_ttile_base.p, nullptr, m->d_scan_sums.p, counters + CN_TETS));
    // CN_TETS / CN_KEPT are adjacent: totals[0] = output tets, totals[1] = kept lattice tets
    static_assert(CN_KEPT == CN_TETS + 1, "scan totals layout");
    MC_TRY(c->stage_check("stage classify"));
    MC_TRY_CUDA(cudaEventRecord(m->ev[3], s));
    // a10 pass 1: vertex masks + counts
    k_vertex_count<<<lgrid, TILE_THREADS, 0, s>>>(S, (unsigned long long*)m->d_vtile_tot.p, T.v_active, T.n_v_active,
                                                  counters + CUR_VCOUNT);
    c->launches += 1;
    MC_TRY_CUDA(cudaGetLastError());
    MC_TRY(scan_tiles(c, m->d_vtile_tot.p, ntiles, m->d_vtile_vbase.p, m->d_vtile_cbase.p, m->d_scan_sums.p, counters + CN_VERTS));
    static_assert(CN_CUTS == CN_VERTS + 1, "scan totals layout");
    MC_TRY(c->stage_check("stage vertex count"));
    MC_TRY_CUDA(cudaEventRecord(m->ev[4], s));
    MC_TRY_CUDA(cudaMemcpyAsync(m->counters, m->d_counters.p, CN_COUNT * 8, cudaMemcpyDeviceToHost, s));
    // first vertex of the tile layer that holds the top owned plane: from there to the end, the
    // coordinates are what the rank above needs for the tets that touch the shared plane
    unsigned long long top_first = 0ull;
    const bool has_upper = m->c1 < L.nz && nown > 0;
    if (has_upper) {
        const uint64_t kz = (m->oz1 - L.pz0) / TS;
        MC_TRY_CUDA(cudaMemcpyAsync(&top_first, (const unsigned long long*)m->d_vtile_vbase.p + kz * m->g.ntx * m->g.nty, 8,
                                    cudaMemcpyDeviceToHost, s));
    }
    MC_TRY_CUDA(cudaStreamSynchronize(s));
    m->top_layer_first = top_first;
    m->n_verts = m->counters[CN_VERTS];
    m->n_cuts = m->counters[CN_CUTS];
    m->n_tets = m->counters[CN_TETS];
    m->n_kept = m->counters[CN_KEPT];
    if (m->n_verts >= (1ull << 40)) {
        mesh_release_all(m);
        delete m;
        return fail(MC_ERR_LIMIT, "too many vertices in one slab");
    }
    m->phase = 1;
    *out = m;
    return MC_OK;
}

// The part of the vertex emission that does not depend on the slab's global vertex base:
// coordinates of the lattice vertices, the cut-edge work list and the bisection.  Only launches
// work; a multi-rank caller issues it before the (blocking) exchange of the counts so that the
// device is busy while the ranks wait for each other.
int mc_tessellate_emit_vertices_local(mc_mesh* m) {
    if (!m || m->phase != 1 || m->vertices_local_done)
        return fail(MC_ERR_STATE, "mc_tessellate_emit_vertices_local: call it once, after mc_tessellate_begin");
    mc_ctx* c = m->ctx;
    MC_CUDA(cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    int rc;
    rc = c->alloc(std::max<uint64_t>(m->n_verts, 1) * 24, m->d_coords);
    if (rc != MC_OK) return rc;
    rc = c->alloc(std::max<uint64_t>(m->n_cuts, 1) * 8, m->d_cutlist);
    if (rc != MC_OK) return rc;
    const Lattice& L = m->L;
    const Stuff S = make_stuff(m);
    unsigned long long* counters = (unsigned long long*)m->d_counters.p;
    MC_CUDA(cudaEventRecord(m->ev[5], s));
    const TileLists T = make_lists(m);
    const unsigned lgrid = list_grid(m);
    k_vertex_emit_interior<<<lgrid, TILE_THREADS, 0, s>>>(S, (double*)m->d_coords.p, T.v_neg, T.n_v_neg, counters + CUR_VEMIT_INT);
    k_vertex_emit<<<lgrid, TILE_THREADS, 0, s>>>(S, (const unsigned long long*)m->d_vtile_cbase.p, (double*)m->d_coords.p,
                                                 (unsigned long long*)m->d_cutlist.p, T.v_active, T.n_v_active,
                                                 counters + CUR_VEMIT);
    c->launches += 2;
    MC_CUDA(cudaGetLastError());
    if ((rc = c->stage_check("stage vertex emit")) != MC_OK) return rc;
    MC_CUDA(cudaEventRecord(m->ev[6], s));
    if (m->n_cuts > 0) {
        if (m->tiled_sdf && m->d_dec.p) {
            // the tile-specialised programs are valid on every lattice edge that starts in the tile
            const Program& p = m->prog;
            const uint32_t n_inst = (uint32_t)p.inst_start.size();
            const uint32_t n_words = std::min<uint32_t>((uint32_t)p.words.size(), c->tile_program_cap_words);
            const size_t smem = (size_t)n_words * 8 + ((p.n_decisions + 15u) & ~15u) + (size_t)(n_inst + 1) * 8;
            const bool capped = n_words < (uint32_t)p.words.size();
            auto kern = capped ? (p.small() ? k_bisect_edges_tiled<true, true> : k_bisect_edges_tiled<false, true>)
                               : (p.small() ? k_bisect_edges_tiled<true, false> : k_bisect_edges_tiled<false, false>);
            if (smem > 48 * 1024) MC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            kern<<<lgrid, TILE_THREADS, smem, s>>>(
                S, (const uint64_t*)m->d_prog.p, (const uint32_t*)m->d_inst_start.p, (const uint16_t*)m->d_word_inst.p,
                n_inst, n_words, p.n_decisions, (const uint8_t*)m->d_dec.p, m->n_tiles,
                (const unsigned long long*)m->d_vtile_cbase.p, m->n_cuts, (const unsigned long long*)m->d_cutlist.p,
                m->error_threshold, m->max_bisect, (double*)m->d_coords.p, counters, T.v_active, T.n_v_active,
                counters + CUR_BISECT);
        } else {
            k_bisect_edges<<<blocks_for(m->n_cuts, 256), 256, 0, s>>>(
                S, (const uint64_t*)m->d_prog.p, (const unsigned long long*)m->d_cutlist.p, m->n_cuts, m->error_threshold,
                m->max_bisect, (double*)m->d_coords.p, counters);
        }
        c->launches++;
        MC_CUDA(cudaGetLastError());
    }
    if ((rc = c->stage_check("stage bisect")) != MC_OK) return rc;
    MC_CUDA(cudaEventRecord(m->ev[7], s));
    m->vertices_local_done = true;
    return MC_OK;
}

int mc_tessellate_emit_vertices(mc_mesh* m, uint64_t vertex_base) {
    if (!m || m->phase != 1) return fail(MC_ERR_STATE, "mc_tessellate_emit_vertices: call mc_tessellate_begin first");
    int rc;
    if (!m->vertices_local_done && (rc = mc_tessellate_emit_vertices_local(m)) != MC_OK) return rc;
    mc_ctx* c = m->ctx;
    MC_CUDA(cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    m->vertex_base = vertex_base;
    const Lattice& L = m->L;
    const Stuff S = make_stuff(m);
    // id table of the top owned plane for the rank above
    if (m->c1 < L.nz && m->points_owned > 0) {
        const uint64_t n = (uint64_t)L.NX * L.NY;
        rc = c->alloc(n * 8, m->d_plane_table);
        if (rc != MC_OK) return rc;
        k_export_plane<<<blocks_for(n, 256), 256, 0, s>>>(S, m->oz1, (unsigned long long*)m->d_plane_table.p);
        c->launches++;
        MC_CUDA(cudaGetLastError());
        if ((rc = c->stage_check("stage export plane")) != MC_OK) return rc;
    }
    m->phase = 2;
    return MC_OK;
}

// Per-axis step classes for the global plain-lattice quality pass (k_quality_classes): returns false
// when an axis has more than QC_K / 2 distinct steps (the per-tile kernels are used instead).
// Layout of the device buffer: [3][QC_K / 2] doubles, [QC_K * QC_K] uint32 presence words, then the
// class bytes of the x, y and z cells.
static bool build_quality_classes(const Lattice& L, std::vector<double>& steps, std::vector<uint8_t>& ids) {
    steps.assign(3 * (QC_K / 2), 0.0);
    ids.clear();
    const uint32_t n[3] = {L.nx, L.ny, L.nz};
    const double o[3] = {L.ox, L.oy, L.oz};
    ids.reserve((size_t)n[0] + n[1] + n[2]);
    for (int a = 0; a < 3; ++a) {
        int nd = 0;
        double* st = steps.data() + a * (QC_K / 2);
        for (uint32_t c = 0; c < n[a]; ++c) {
            // lat_x / lat_y / lat_z of mc_lattice.cuh, same arithmetic (no contraction on either side)
            const double p0 = (double)c * L.res + o[a], p1 = (double)(c + 1) * L.res + o[a];
            const double d = p1 - p0;
            int k = 0;
            while (k < nd && std::memcmp(&st[k], &d, 8) != 0) ++k;
            if (k == nd) {
                if (nd == QC_K / 2) return false;
                st[nd++] = d;
            }
            ids.push_back((uint8_t)(2 * k + (int)(c & 1u)));
        }
    }
    return true;
}

int mc_tessellate_set_lower_coords(mc_mesh* m, const void* d_coords, uint64_t first_global_id, uint64_t count) {
    if (!m || m->phase >= 3) return fail(MC_ERR_STATE, "mc_tessellate_set_lower_coords: call it before mc_tessellate_emit_tets");
    m->lower_coords = count ? (const double*)d_coords : nullptr;
    m->lower_first = first_global_id;
    m->lower_count = m->lower_coords ? count : 0;
    return MC_OK;
}

int mc_mesh_upper_layer_vertices(const mc_mesh* m, uint64_t* first_local, uint64_t* count) {
    if (!m || m->phase < 1) return fail(MC_ERR_STATE, "mc_mesh_upper_layer_vertices: call mc_tessellate_begin first");
    const bool has_upper = m->c1 < m->L.nz && m->points_owned > 0;
    if (first_local) *first_local = has_upper ? m->top_layer_first : m->n_verts;
    if (count) *count = has_upper ? m->n_verts - m->top_layer_first : 0;
    return MC_OK;
}

int mc_mesh_upper_layer_coords(mc_mesh* m, void** d_coords) {
    if (!m || !d_coords || !(m->vertices_local_done || m->phase >= 2))
        return fail(MC_ERR_STATE, "mc_mesh_upper_layer_coords: call mc_tessellate_emit_vertices(_local) first");
    uint64_t first = 0, count = 0;
    mc_mesh_upper_layer_vertices(m, &first, &count);
    *d_coords = count ? (void*)((double*)m->d_coords.p + 3 * first) : nullptr;
    return MC_OK;
}

int mc_mesh_copy_upper_layer_coords(mc_mesh* m, void* d_dst) {
    void* src = nullptr;
    int rc = mc_mesh_upper_layer_coords(m, &src);
    if (rc != MC_OK) return rc;
    uint64_t first = 0, count = 0;
    mc_mesh_upper_layer_vertices(m, &first, &count);
    if (!count) return MC_OK;
    if (!d_dst) return fail(MC_ERR_ARG, "mc_mesh_copy_upper_layer_coords: null destination");
    MC_CUDA(cudaSetDevice(m->ctx->device));
    MC_CUDA(cudaMemcpyAsync(d_dst, src, count * 24, cudaMemcpyDeviceToDevice, m->ctx->stream));
    return MC_OK;
}

int mc_tessellate_emit_tets(mc_mesh* m, uint64_t tet_base, const void* d_lower_plane) {
    if (!m || m->phase != 2) return fail(MC_ERR_STATE, "mc_tessellate_emit_tets: call mc_tessellate_emit_vertices first");
    mc_ctx* c = m->ctx;
    MC_CUDA(cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    m->tet_base = tet_base;
    unsigned long long* counters = (unsigned long long*)m->d_counters.p;
    if (m->c0 > 0 && !d_lower_plane)
        return fail(MC_ERR_ARG, "mc_tessellate_emit_tets: slab does not start at z = 0, the lower plane table is required");
    m->lower_table = m->c0 > 0 ? d_lower_plane : nullptr;
    m->index_bytes = (m->vertex_base + m->n_verts < (1ull << 32)) ? 4 : 8;
    int rc = c->alloc(std::max<uint64_t>(m->n_tets, 1) * 4 * (size_t)m->index_bytes, m->d_tets);
    if (rc != MC_OK) return rc;
    const Stuff S = make_stuff(m);
    const bool quality = !(m->flags & MC_OPT_SKIP_QUALITY);
    if (!c->tet_emit_attr_set) {
        // per context (= per device): opt in to the dynamic shared memory of the surface-tile emission
        MC_CUDA(cudaFuncSetAttribute(k_tet_emit<uint32_t, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tet_emit_smem_bytes<uint32_t>()));
        MC_CUDA(cudaFuncSetAttribute(k_tet_emit<uint32_t, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tet_emit_smem_bytes<uint32_t>()));
        MC_CUDA(cudaFuncSetAttribute(k_tet_emit<unsigned long long, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tet_emit_smem_bytes<unsigned long long>()));
        MC_CUDA(cudaFuncSetAttribute(k_tet_emit<unsigned long long, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tet_emit_smem_bytes<unsigned long long>()));
        c->tet_emit_attr_set = true;
    }
    MC_CUDA(cudaEventRecord(m->ev[8], s));
    if (m->n_tets > 0) {
        const unsigned grid = (unsigned)m->n_tiles;
        const TileLists T = make_lists(m);
        const unsigned lgrid = list_grid(m);
        const double* coords = (const double*)m->d_coords.p;
        const uint64_t nlist = quality ? m->counters[CN_LISTED_TETS] : 0;
        PoolScope scratch(c);   // (kernels queued on the stream keep using the buffers: the cache hands them out in stream order)
        DevBuf qlist;
        rc = scratch.alloc(std::max<uint64_t>(nlist, 1) * 8, qlist);
        if (rc != MC_OK) return rc;
        unsigned long long* ql = (unsigned long long*)qlist.p;
        // plain lattice tets: one evaluation per global (step, parity) class when the steps take few
        // distinct values, else per tile
        DevBuf qcls;
        QualityClasses Q{};
        bool class_quality = false;
        if (quality) {
            std::vector<double> steps;
            std::vector<uint8_t> ids;
            if (!(m->flags & MC_OPT_PER_TILE_QUALITY) && build_quality_classes(m->L, steps, ids)) {
                const size_t off_present = steps.size() * 8, off_ids = off_present + (size_t)QC_K * QC_K * 4;
                rc = scratch.alloc(off_ids + ids.size() + 16, qcls);
                if (rc != MC_OK) return rc;
                char* base = (char*)qcls.p;
                MC_CUDA(cudaMemcpyAsync(base, steps.data(), steps.size() * 8, cudaMemcpyHostToDevice, s));
                MC_CUDA(cudaMemsetAsync(base + off_present, 0, (size_t)QC_K * QC_K * 4, s));
                MC_CUDA(cudaMemcpyAsync(base + off_ids, ids.data(), ids.size(), cudaMemcpyHostToDevice, s));
                Q.steps = (const double*)base;
                Q.present = (uint32_t*)(base + off_present);
                Q.kx = (const uint8_t*)(base + off_ids);
                Q.ky = Q.kx + m->L.nx;
                Q.kz = Q.ky + m->L.ny;
                class_quality = true;
            }
        }
        if (m->index_bytes == 4) {
            uint32_t* out = (uint32_t*)m->d_tets.p;
            k_tet_emit_full<uint32_t><<<lgrid, TILE_THREADS, 0, s>>>(S, out, T.c_full, T.n_c_full, counters + CUR_TET_FULL);
            if ((rc = c->stage_check("stage tet emit (interior tiles)")) != MC_OK) return rc;
            if (quality) {
                if (class_quality) {
                    k_quality_mark_full<<<blocks_for(m->n_tiles, 8), 256, 0, s>>>(S, Q, T.c_full, T.n_c_full);
                } else {
                    k_full_tile_quality<<<grid, 64, 0, s>>>(S, counters);
                    k_surface_tile_quality<<<grid, TILE_THREADS, 0, s>>>(S, counters);
                }
                if ((rc = c->stage_check("stage tile quality")) != MC_OK) return rc;
                // (the surface tiles' plain cells are marked by the emission itself)
                k_tet_emit<uint32_t, true><<<lgrid, TILE_THREADS, tet_emit_smem_bytes<uint32_t>(), s>>>(S, out, ql, Q, counters, T.c_active, T.n_c_active, counters + CUR_TET_SURF);
                if (class_quality) k_quality_classes<<<QC_K * QC_K * QC_K / 256, 256, 0, s>>>(Q, counters);
                if ((rc = c->stage_check("stage tet emit (surface tiles)")) != MC_OK) return rc;
                if (nlist) k_tet_quality_list<uint32_t><<<blocks_for(nlist, 256), 256, 0, s>>>(coords, out, ql, nlist, (long long)m->vertex_base, m->lower_coords, (long long)m->lower_first, (long long)m->lower_count, counters);
            } else {
                k_tet_emit<uint32_t, false><<<lgrid, TILE_THREADS, tet_emit_smem_bytes<uint32_t>(), s>>>(S, out, ql, Q, counters, T.c_active, T.n_c_active, counters + CUR_TET_SURF);
            }
        } else {
            unsigned long long* out = (unsigned long long*)m->d_tets.p;
            k_tet_emit_full<unsigned long long><<<lgrid, TILE_THREADS, 0, s>>>(S, out, T.c_full, T.n_c_full, counters + CUR_TET_FULL);
            if ((rc = c->stage_check("stage tet emit (interior tiles)")) != MC_OK) return rc;
            if (quality) {
                if (class_quality) {
                    k_quality_mark_full<<<blocks_for(m->n_tiles, 8), 256, 0, s>>>(S, Q, T.c_full, T.n_c_full);
                } else {
                    k_full_tile_quality<<<grid, 64, 0, s>>>(S, counters);
                    k_surface_tile_quality<<<grid, TILE_THREADS, 0, s>>>(S, counters);
                }
                if ((rc = c->stage_check("stage tile quality")) != MC_OK) return rc;
                // (the surface tiles' plain cells are marked by the emission itself)
                k_tet_emit<unsigned long long, true><<<lgrid, TILE_THREADS, tet_emit_smem_bytes<unsigned long long>(), s>>>(S, out, ql, Q, counters, T.c_active, T.n_c_active, counters + CUR_TET_SURF);
                if (class_quality) k_quality_classes<<<QC_K * QC_K * QC_K / 256, 256, 0, s>>>(Q, counters);
                if ((rc = c->stage_check("stage tet emit (surface tiles)")) != MC_OK) return rc;
                if (nlist) k_tet_quality_list<unsigned long long><<<blocks_for(nlist, 256), 256, 0, s>>>(coords, out, ql, nlist, (long long)m->vertex_base, m->lower_coords, (long long)m->lower_first, (long long)m->lower_count, counters);
            } else {
                k_tet_emit<unsigned long long, false><<<lgrid, TILE_THREADS, tet_emit_smem_bytes<unsigned long long>(), s>>>(S, out, ql, Q, counters, T.c_active, T.n_c_active, counters + CUR_TET_SURF);
            }
        }
        c->launches += quality ? 5 : 2;
        MC_CUDA(cudaGetLastError());
        if ((rc = c->stage_check("stage tet emit / quality list")) != MC_OK) return rc;
    }
    MC_CUDA(cudaEventRecord(m->ev[9], s));
    MC_CUDA(cudaMemcpyAsync(m->counters, m->d_counters.p, CN_COUNT * 8, cudaMemcpyDeviceToHost, s));
    MC_CUDA(cudaStreamSynchronize(s));
    if (m->counters[CN_CLASS_INVERTED]) {
        // some plain lattice tet is inverted (degenerate lattice): the count needs multiplicities
        const Stuff S2 = make_stuff(m);
        k_full_tile_quality<<<(unsigned)m->n_tiles, 64, 0, s>>>(S2, counters);
        k_surface_tile_quality<<<(unsigned)m->n_tiles, TILE_THREADS, 0, s>>>(S2, counters);
        c->launches += 2;
        MC_CUDA(cudaGetLastError());
        MC_CUDA(cudaMemcpyAsync(m->counters, m->d_counters.p, CN_COUNT * 8, cudaMemcpyDeviceToHost, s));
        MC_CUDA(cudaStreamSynchronize(s));
    }
    float ms;
    const int pairs[6][2] = {{0, 1}, {1, 2}, {2, 3}, {5, 6}, {6, 7}, {8, 9}};
    for (int i = 0; i < 6; ++i) {
        MC_CUDA(cudaEventElapsedTime(&ms, m->ev[pairs[i][0]], m->ev[pairs[i][1]]));
        m->stage_ms[i] = ms;
    }
    m->stage_ms[6] = 0.0;  // quality is fused into the tet emission
    MC_CUDA(cudaEventElapsedTime(&ms, m->ev[3], m->ev[4]));
    m->stage_ms[3] += ms;  // vertex count + scan belongs to "vertex scan+emit"
    // scratch that is no longer needed goes back to the pool now
    c->release(m->d_cutlist);
    c->release(m->d_dec);
    m->phase = 3;
    if (m->counters[CN_BISECT_FAIL])
        return fail(MC_ERR_DIVERGED, "edge bisection did not reach the error threshold within the iteration cap on " +
                                         std::to_string(m->counters[CN_BISECT_FAIL]) + " edges");
    return MC_OK;
}

int mc_tessellate(mc_ctx* c, const mc_tape* volume, int32_t root, double resolution, double relative_error,
                  const mc_options* opts, mc_mesh** out) {
    mc_mesh* m = nullptr;
    int rc = mc_tessellate_begin(c, volume, root, resolution, relative_error, nullptr, opts, &m);
    if (rc != MC_OK) return rc;
    rc = mc_tessellate_emit_vertices(m, 0);
    if (rc == MC_OK) rc = mc_tessellate_emit_tets(m, 0, nullptr);
    if (rc != MC_OK && rc != MC_ERR_DIVERGED) { mc_mesh_free(m); return rc; }
    *out = m;
    return rc;
}

void mc_mesh_free(mc_mesh* m) {
    if (!m) return;
    cudaSetDevice(m->ctx->device);
    cudaStreamSynchronize(m->ctx->stream);
    mesh_release_all(m);
    delete m;
}

int mc_mesh_counts(const mc_mesh* m, uint64_t counts[6]) {
    if (!m || !counts) return fail(MC_ERR_ARG, "mc_mesh_counts: bad argument");
    counts[0] = m->n_verts; counts[1] = m->n_tets; counts[2] = m->points_evaluated;
    counts[3] = m->points_owned; counts[4] = m->n_cuts; counts[5] = m->counters[CN_WARPED];
    return MC_OK;
}

int mc_mesh_prune_stats(const mc_mesh* m, uint64_t* n_tiles, uint64_t* pruned_words_total, uint64_t* full_words) {
    if (!m) return fail(MC_ERR_ARG, "mc_mesh_prune_stats: null mesh");
    if (n_tiles) *n_tiles = m->tiled_sdf ? m->n_tiles : 0;
    if (pruned_words_total) *pruned_words_total = m->counters[CN_PRUNED_WORDS];
    if (full_words) *full_words = m->prog.words.size();
    return MC_OK;
}

int mc_mesh_tile_stats(const mc_mesh* m, uint64_t out[6]) {
    if (!m || !out) return fail(MC_ERR_ARG, "mc_mesh_tile_stats: bad argument");
    out[0] = m->n_tiles;
    out[1] = m->counters[CN_SKIPPED_TILES];
    out[2] = m->counters[CN_ACTIVE_VTILES];
    out[3] = m->counters[CN_ACTIVE_CTILES];
    out[4] = m->counters[CN_SDF_FLOPS];
    out[5] = m->counters[CN_SUB_BOXES_PROVEN];
    return MC_OK;
}

int mc_mesh_upper_plane_table(mc_mesh* m, void** d_table, uint64_t* n_entries) {
    if (!m || !d_table || !n_entries) return fail(MC_ERR_ARG, "mc_mesh_upper_plane_table: bad argument");
    if (m->phase < 2) return fail(MC_ERR_STATE, "mc_mesh_upper_plane_table: emit the vertices first");
    *d_table = m->d_plane_table.p;
    *n_entries = m->d_plane_table.p ? (uint64_t)m->L.NX * m->L.NY : 0;
    return MC_OK;
}

int mc_mesh_copy_upper_plane_table(mc_mesh* m, void* d_dst) {
    if (!m || !d_dst) return fail(MC_ERR_ARG, "mc_mesh_copy_upper_plane_table: bad argument");
    if (m->phase < 2 || !m->d_plane_table.p) return fail(MC_ERR_STATE, "mc_mesh_copy_upper_plane_table: no table (last slab or vertices not emitted)");
    MC_CUDA(cudaSetDevice(m->ctx->device));
    MC_CUDA(cudaMemcpyAsync(d_dst, m->d_plane_table.p, (size_t)m->L.NX * m->L.NY * 8, cudaMemcpyDeviceToDevice, m->ctx->stream));
    return MC_OK;
}

int mc_mesh_stats(const mc_mesh* m, uint64_t dim[3], uint64_t ntets_stage[3], uint64_t stats[7], double ar_minmax[2],
                  uint64_t* inverted) {
    if (!m) return fail(MC_ERR_ARG, "mc_mesh_stats: null mesh");
    if (dim) for (int k = 0; k < 3; ++k) dim[k] = m->dim[k];
    if (ntets_stage) {
        ntets_stage[0] = m->n_kept;                                   // after cut_lattice (:424)
        ntets_stage[1] = m->n_kept;                                   // trim_spikes entry (:247)
        ntets_stage[2] = m->n_tets + m->counters[CN_EXT_REMOVED];     // trim_spikes exit (:366)
    }
    if (stats) for (int k = 0; k < 7; ++k) stats[k] = m->counters[CN_STAT0 + k];
    if (ar_minmax) {
        std::memcpy(&ar_minmax[0], &m->counters[CN_AR_MIN_BITS], 8);
        std::memcpy(&ar_minmax[1], &m->counters[CN_AR_MAX_BITS], 8);
    }
    if (inverted) *inverted = m->counters[CN_INVERTED];
    return MC_OK;
}

int mc_mesh_stage_ms(const mc_mesh* m, double ms[9]) {
    if (!m || !ms) return fail(MC_ERR_ARG, "mc_mesh_stage_ms: bad argument");
    for (int i = 0; i < 9; ++i) ms[i] = m->stage_ms[i];
    return MC_OK;
}

uint64_t mc_mesh_num_vertices(const mc_mesh* m) { return m ? m->n_verts : 0; }
uint64_t mc_mesh_num_tets(const mc_mesh* m) { return m ? m->n_tets : 0; }

int mc_mesh_device_views(mc_mesh* m, void** d_coords, void** d_tets, int* index_bytes) {
    if (!m || m->phase < 3) return fail(MC_ERR_STATE, "mc_mesh_device_views: mesh not emitted");
    if (d_coords) *d_coords = m->d_coords.p;
    if (d_tets) *d_tets = m->d_tets.p;
    if (index_bytes) *index_bytes = m->index_bytes;
    return MC_OK;
}

int mc_mesh_copy_to_host(mc_mesh* m, double* coords, void* tets, int index_bytes) {
    if (!m || m->phase < 3) return fail(MC_ERR_STATE, "mc_mesh_copy_to_host: mesh not emitted");
    if (index_bytes != 4 && index_bytes != 8) return fail(MC_ERR_ARG, "mc_mesh_copy_to_host: index_bytes must be 4 or 8");
    mc_ctx* c = m->ctx;
    MC_CUDA(cudaSetDevice(c->device));
    if (coords && m->n_verts)
        MC_CUDA(cudaMemcpyAsync(coords, m->d_coords.p, m->n_verts * 24, cudaMemcpyDeviceToHost, c->stream));
    if (tets && m->n_tets) {
        if (index_bytes == m->index_bytes) {
            MC_CUDA(cudaMemcpyAsync(tets, m->d_tets.p, m->n_tets * 4 * (size_t)index_bytes, cudaMemcpyDeviceToHost, c->stream));
            MC_CUDA(cudaStreamSynchronize(c->stream));
        } else if (index_bytes == 8) {  // widen on the host
            std::vector<uint32_t> tmp(m->n_tets * 4);
            MC_CUDA(cudaMemcpyAsync(tmp.data(), m->d_tets.p, tmp.size() * 4, cudaMemcpyDeviceToHost, c->stream));
            MC_CUDA(cudaStreamSynchronize(c->stream));
            uint64_t* o = (uint64_t*)tets;
            for (size_t i = 0; i < tmp.size(); ++i) o[i] = tmp[i];
        } else {
            return fail(MC_ERR_ARG, "mc_mesh_copy_to_host: ids do not fit 32 bits");
        }
    }
    MC_CUDA(cudaStreamSynchronize(c->stream));
    return MC_OK;
}

const double* mc_mesh_coords(mc_mesh* m) {
    if (!m || m->phase < 3) { fail(MC_ERR_STATE, "mc_mesh_coords: mesh not emitted"); return nullptr; }
    if (!m->have_h_coords) {
        m->h_coords.resize(m->n_verts * 3);
        if (mc_mesh_copy_to_host(m, m->h_coords.data(), nullptr, 8) != MC_OK) return nullptr;
        m->have_h_coords = true;
    }
    return m->h_coords.data();
}
const uint64_t* mc_mesh_tets(mc_mesh* m) {
    if (!m || m->phase < 3) { fail(MC_ERR_STATE, "mc_mesh_tets: mesh not emitted"); return nullptr; }
    if (!m->have_h_tets) {
        m->h_tets.resize(m->n_tets * 4);
        if (mc_mesh_copy_to_host(m, nullptr, m->h_tets.data(), 8) != MC_OK) return nullptr;
        m->have_h_tets = true;
    }
    return m->h_tets.data();
}

int mc_mesh_copy_phi(mc_mesh* m, double* out, uint64_t n) {
    if (!m || !out) return fail(MC_ERR_ARG, "mc_mesh_copy_phi: bad argument");
    const uint64_t have = (uint64_t)m->L.NX * m->L.NY * (m->L.pz1 - m->L.pz0 + 1);
    if (n != have) return fail(MC_ERR_ARG, "mc_mesh_copy_phi: expected " + std::to_string(have) + " values");
    if (m->tiled_sdf && !(m->flags & MC_OPT_KEEP_PHI))
        return fail(MC_ERR_STATE, "mc_mesh_copy_phi: the field holds sign placeholders where the sign was proven; "
                                  "tessellate with MC_OPT_KEEP_PHI to read it");
    MC_CUDA(cudaSetDevice(m->ctx->device));
    // the rows of the field are padded to L.PX points on the device
    MC_CUDA(cudaMemcpy2DAsync(out, (size_t)m->L.NX * 8, m->d_phi.p, (size_t)m->L.PX * 8, (size_t)m->L.NX * 8,
                              (size_t)m->L.NY * (m->L.pz1 - m->L.pz0 + 1), cudaMemcpyDeviceToHost, m->ctx->stream));
    MC_CUDA(cudaStreamSynchronize(m->ctx->stream));
    return MC_OK;
}

// ---- boundary / side-sets (implementation in mc_boundary.cu) -----------------------------
int mc_mesh_boundary(mc_mesh* m, uint64_t* n_sides) {
    if (!m || m->phase < 3) return fail(MC_ERR_STATE, "mc_mesh_boundary: mesh not emitted");
    if (!m->have_boundary) {
        int rc = compute_boundary(m);
        if (rc != MC_OK) return rc;
    }
    if (n_sides) *n_sides = m->n_sides;
    return MC_OK;
}
const uint64_t* mc_mesh_boundary_sides(mc_mesh* m) {
    if (mc_mesh_boundary(m, nullptr) != MC_OK) return nullptr;
    if (boundary_host_view(m) != MC_OK) return nullptr;
    return m->h_sides.data();
}
int mc_mesh_boundary_device(mc_mesh* m, void** d_sides, uint64_t* n_sides) {
    int rc = mc_mesh_boundary(m, n_sides);
    if (rc != MC_OK) return rc;
    if (d_sides) *d_sides = m->d_sides.p;
    return MC_OK;
}
int mc_mesh_add_sideset(mc_mesh* m, const mc_tape* boundary, int32_t root) {
    if (!m || !boundary) return fail(MC_ERR_ARG, "mc_mesh_add_sideset: null handle");
    int rc = mc_mesh_boundary(m, nullptr);
    if (rc != MC_OK) return rc;
    SideSetDev ss;
    rc = compute_sideset(m, boundary, root, ss);
    if (rc != MC_OK) { m->ctx->release(ss.d_pairs); return rc; }
    m->sidesets.push_back(std::move(ss));
    return (int)m->sidesets.size() - 1;
}
uint64_t mc_mesh_sideset_len(const mc_mesh* m, int i) {
    if (!m || i < 0 || (size_t)i >= m->sidesets.size()) return 0;
    return m->sidesets[(size_t)i].n;
}
const uint64_t* mc_mesh_sideset(mc_mesh* m, int i) {
    if (!m || i < 0 || (size_t)i >= m->sidesets.size()) { fail(MC_ERR_ARG, "mc_mesh_sideset: bad index"); return nullptr; }
    SideSetDev& ss = m->sidesets[(size_t)i];
    if (!ss.have_h) {
        ss.h.resize(ss.n * 2 + 1);
        if (ss.n) {
            cudaSetDevice(m->ctx->device);
            cudaError_t e = cudaMemcpyAsync(ss.h.data(), ss.d_pairs.p, ss.n * 16, cudaMemcpyDeviceToHost, m->ctx->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(m->ctx->stream);
            if (e != cudaSuccess) { cuda_fail(e, "mc_mesh_sideset"); return nullptr; }
        }
        ss.have_h = true;
    }
    return ss.h.data();
}
int mc_mesh_sideset_device(mc_mesh* m, int i, void** d_pairs, uint64_t* n) {
    if (!m || i < 0 || (size_t)i >= m->sidesets.size()) return fail(MC_ERR_ARG, "mc_mesh_sideset_device: bad index");
    if (d_pairs) *d_pairs = m->sidesets[(size_t)i].d_pairs.p;
    if (n) *n = m->sidesets[(size_t)i].n;
    return MC_OK;
}

// ---- Mesh::to_boundary -------------------------------------------------------------------
int mc_mesh_to_boundary(mc_mesh* m, mc_surface** out) {
    if (!m || !out) return fail(MC_ERR_ARG, "mc_mesh_to_boundary: bad argument");
    *out = nullptr;
    int rc = mc_mesh_boundary(m, nullptr);
    if (rc != MC_OK) return rc;
    mc_surface* sf = new mc_surface();
    sf->ctx = m->ctx;
    rc = compute_surface(m, sf);
    if (rc != MC_OK) { mc_surface_free(sf); return rc; }
    *out = sf;
    return MC_OK;
}
void mc_surface_free(mc_surface* sf) {
    if (!sf) return;
    cudaSetDevice(sf->ctx->device);
    cudaStreamSynchronize(sf->ctx->stream);
    sf->ctx->release(sf->d_coords);
    sf->ctx->release(sf->d_tris);
    delete sf;
}
uint64_t mc_surface_num_vertices(const mc_surface* sf) { return sf ? sf->n_verts : 0; }
uint64_t mc_surface_num_tris(const mc_surface* sf) { return sf ? sf->n_tris : 0; }
const double* mc_surface_coords(const mc_surface* sf) { return sf ? sf->h_coords.data() : nullptr; }
const uint64_t* mc_surface_tris(const mc_surface* sf) { return sf ? sf->h_tris.data() : nullptr; }

// ---- mesh-crate predicates ---------------------------------------------------------------
int mc_tet_quality(mc_ctx* c, const double* coords, uint64_t nv, const uint64_t* tets, uint64_t nt, double* volume,
                   double* aspect_ratio) {
    if (!c || !coords || !tets || !volume || !aspect_ratio) return fail(MC_ERR_ARG, "mc_tet_quality: bad argument");
    if (nt == 0) return MC_OK;
    MC_CUDA(cudaSetDevice(c->device));
    DevBuf dc, dt, dv, da;
    int rc = c->alloc(nv * 24, dc);
    if (rc == MC_OK) rc = c->alloc(nt * 32, dt);
    if (rc == MC_OK) rc = c->alloc(nt * 8, dv);
    if (rc == MC_OK) rc = c->alloc(nt * 8, da);
    if (rc == MC_OK) {
        cudaError_t e = cudaMemcpyAsync(dc.p, coords, nv * 24, cudaMemcpyHostToDevice, c->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(dt.p, tets, nt * 32, cudaMemcpyHostToDevice, c->stream);
        if (e == cudaSuccess) {
            k_tet_quality_arrays<<<blocks_for(nt, 256), 256, 0, c->stream>>>(
                (const double*)dc.p, (const unsigned long long*)dt.p, nt, (double*)dv.p, (double*)da.p);
            c->launches++;
            e = cudaGetLastError();
        }
        if (e == cudaSuccess) e = cudaMemcpyAsync(volume, dv.p, nt * 8, cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(aspect_ratio, da.p, nt * 8, cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) rc = cuda_fail(e, "mc_tet_quality");
    }
    c->release(dc); c->release(dt); c->release(dv); c->release(da);
    return rc;
}

// ---- host-only self description -------------------------------------------------------------
int64_t mc_shape_tables(void* out, uint64_t bytes) {
    if (!out) return (int64_t)sizeof(ShapeTables);
    if (bytes != sizeof(ShapeTables)) return fail(MC_ERR_ARG, "mc_shape_tables: bytes != sizeof(ShapeTables)");
    static ShapeTables tables;
    if (!build_shape_tables(tables)) return fail(MC_ERR_STATE, "mc_shape_tables: the tables failed their consistency checks");
    memcpy(out, &tables, sizeof(ShapeTables));
    return MC_OK;
}

// ---- probes -----------------------------------------------------------------------------
int mc_probe_fp64_nofma(mc_ctx* c, double* flops_per_s) {
    if (!c || !flops_per_s) return fail(MC_ERR_ARG, "mc_probe_fp64_nofma: bad argument");
    MC_CUDA(cudaSetDevice(c->device));
    DevBuf d;
    int rc = c->alloc(8, d);
    if (rc != MC_OK) return rc;
    cudaEvent_t a, b;
    MC_CUDA(cudaEventCreate(&a));
    MC_CUDA(cudaEventCreate(&b));
    const int iters = 4096, threads = 256, blocks = c->sm_count * 8;
    k_probe_fp64<<<blocks, threads, 0, c->stream>>>((double*)d.p, 64);  // warm-up
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        MC_CUDA(cudaEventRecord(a, c->stream));
        k_probe_fp64<<<blocks, threads, 0, c->stream>>>((double*)d.p, iters);
        MC_CUDA(cudaEventRecord(b, c->stream));
        MC_CUDA(cudaEventSynchronize(b));
        float ms = 0;
        MC_CUDA(cudaEventElapsedTime(&ms, a, b));
        const double flops = (double)blocks * threads * (double)iters * 16.0;
        best = std::max(best, flops / (ms * 1e-3));
        c->launches++;
    }
    c->launches++;
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    c->release(d);
    *flops_per_s = best;
    return MC_OK;
}

}  // extern "C"
