// CUDA kernels of the meshing hot path (sm_100a, f64, --fmad=false).
// Stage map (SURVEY.md §8a -> kernel):
//   a2  ObjectAdaptor::value over the lattice      k_sdf_field / k_sdf_field_tiled (mc_prune.cuh)
//   --  per-tile sign summary                      k_tile_minmax, k_tile_neighbourhood, k_cell_tile_class
//   a6  warp_vertices                              k_warp_codes
//   a5/a7/a9 cut_lattice, trim_spikes, exterior    k_classify_cells
//   a10 compact (vertex numbering)                 k_vertex_count / k_vertex_emit (+ k_scan_tiles)
//   a8  cut_edge bisection                         k_bisect_edges
//   a7/a10/a11 tet emission + quality              k_tet_emit
//
// All stuffing kernels run one block per TS^3 tile.  A tile whose neighbourhood has one strict
// sign takes a fast path that never reads the SDF field: exterior tiles produce nothing,
// interior tiles produce lattice vertices / the five lattice tets of every cell with ids that
// are pure arithmetic (vertex and tet numbering is tile-major), so the bulk of the lattice costs
// only the bytes of its output.
#pragma once
#include "mc_shapes.hpp"
#include "mc_stuffing.cuh"
#include "mc_tma.cuh"

namespace mc {

// device-side counters, one uint64 each
enum Counter : int {
    CN_VERTS = 0, CN_CUTS, CN_TETS, CN_KEPT, CN_STAT0, /* ..CN_STAT0+6 */
    CN_WARPED = CN_STAT0 + 7, CN_EXT_REMOVED, CN_BISECT_FAIL, CN_BISECT_ITERS, CN_INVERTED,
    CN_AR_MIN_BITS, CN_AR_MAX_BITS, CN_PRUNED_WORDS, CN_ACTIVE_VTILES, CN_ACTIVE_CTILES, CN_SKIPPED_TILES, CN_LISTED_TETS, CN_QLIST_FILL, CN_SDF_FLOPS, CN_SUB_BOXES_PROVEN, CN_CLASS_INVERTED, CN_SDF_LIST, CN_SDF_NEXT, CN_LIST_VNEG, CN_LIST_CFULL, CN_SURF_CELLS, CN_SURF_FILL, CN_COUNT
};

constexpr int TILE_THREADS = 256;
constexpr uint32_t TILE_POINTS = TS * TS * TS;
constexpr int TILE_ROUNDS = (int)(TILE_POINTS / TILE_THREADS);

// per-tile flags
constexpr uint8_t TF_NEG = 1, TF_POS = 2, TF_HASWARP = 4;
constexpr uint8_t VU_ACTIVE = 0, VU_NEG = 1, VU_POS = 2;    // point tile: 27-neighbourhood all NEG / all POS
constexpr uint8_t CU_ACTIVE = 0, CU_FULL = 1, CU_EMPTY = 2; // cell tile: every cell emits its 5 lattice tets / nothing
// per-cell info written by k_classify_cells (ACTIVE cell tiles only):
// bits 2t..2t+1 = number of output tets of lattice tet t (0..3), bits 10..14 = lattice tet t is emitted
// untouched (kept, not trimmed, not dropped by the exterior test: its one output tet is the lattice
// tet itself), bit 15 = all five lattice tets emitted untouched
constexpr uint16_t CI_UNTOUCHED_SHIFT = 10;
constexpr uint16_t CI_FULL = 1u << 15;
constexpr uint16_t CI_ALL_FULL = (uint16_t)(0x155u | CI_FULL);  // one output tet per lattice tet
// on a FULL cell the removed bits are free: bit 10 marks "some corner is zero (possibly warped)",
// i.e. the cell's tets are not congruent to plain lattice tets and get their quality evaluated
// individually (k_tet_quality_list)
constexpr uint16_t CI_FULL_ZERO_CORNER = 1u << 10;
__host__ __device__ __forceinline__ uint32_t ci_tet_n(uint16_t ci, int t) { return (ci >> (2 * t)) & 3u; }
__host__ __device__ __forceinline__ uint32_t ci_count(uint16_t ci) {
    return (ci & 3u) + ((ci >> 2) & 3u) + ((ci >> 4) & 3u) + ((ci >> 6) & 3u) + ((ci >> 8) & 3u);
}
// per-vertex mask written by k_vertex_count (ACTIVE point tiles only)
constexpr uint16_t VM_CUT_MASK = 0x01ff;   // bit s: the edge along canonical direction s is cut
constexpr uint16_t VM_USED = 1u << 15;     // the lattice vertex itself is in the output

// Work lists of the tile kernels: only the tiles that have something to do are visited, blocks take
// them one at a time from a cursor (tiles differ in cost by orders of magnitude).
struct TileLists {
    const uint32_t* v_active;  // point tiles on the general path (VU_ACTIVE)
    const uint32_t* v_neg;     // interior point tiles (VU_NEG): arithmetic vertices
    const uint32_t* c_active;  // cell tiles on the general path (CU_ACTIVE)
    const uint32_t* c_full;    // interior cell tiles (CU_FULL): five lattice tets per cell
    const unsigned long long* n_v_active;  // list lengths (device counters)
    const unsigned long long* n_v_neg;
    const unsigned long long* n_c_active;
    const unsigned long long* n_c_full;
};
constexpr int TILE_GRID_PER_SM = 8;  // blocks launched per SM by the list kernels (they loop)

// next work item of this block: false when the list is exhausted.  All threads must call it.
__device__ __forceinline__ bool next_tile(const uint32_t* __restrict__ list, const unsigned long long* __restrict__ n,
                                          unsigned long long* __restrict__ cursor, uint32_t& tile) {
    __shared__ unsigned long long s_item;
    __syncthreads();  // the previous tile is done with shared memory (and s_item has been read)
    if (threadIdx.x == 0) s_item = atomicAdd(cursor, 1ull);
    __syncthreads();
    const unsigned long long i = s_item;
    if (i >= *n) return false;
    tile = list[i];
    return true;
}

// everything the stuffing kernels need about one slab
struct Stuff {
    Lattice L;
    TileGrid g;
    uint32_t c0, c1;      // owned cell layers [c0, c1)
    uint32_t cl0, cl1;    // classified cell layers (halo layers only contribute "used" flags)
    uint32_t oz0, oz1;    // owned point planes [oz0, oz1]
    uint32_t wz0, wz1;    // planes with warp codes
    uint8_t* tflag;       // [tiles] TF_*
    uint8_t* vuni;        // [tiles] VU_*
    uint8_t* cuni;        // [tiles] CU_*
    uint16_t* vmask;      // [points] VM_*
    uint16_t* vloc;       // [points] offset of the vertex's first output inside its tile
    uint16_t* cinfo;      // [cells of classified layers] CI_*
    // Shape bytes (mc_shapes.hpp) of the five lattice tets of every SURFACE cell (classified, neither full nor
    // empty), byte t of cshape[cidx]; written by k_classify_cells, read by the tet emission and by
    // Mesh::boundary (mc_boundary.cu).  Full / empty cells are told apart by cinfo alone.
    unsigned long long* cshape;
    const ShapeTables* shapes;         // device copy of the shape tables (per context)
    bool emit_reclassify;              // MC_OPT_RECLASSIFY_EMIT: the tet emission sorts again instead of using cshape (A/B check)
    const unsigned long long* warp_table;  // visit-order table of warp_vertices, see build_warp_table
    const unsigned long long* tvbase;  // [tiles] first output vertex of the tile (slab-local)
    const unsigned long long* ttbase;  // [tiles] first output tet of the tile (slab-local)
    const unsigned long long* lower_table;  // id table of plane c0 from the rank below, or null
    long long vertex_base;             // global id of this slab's first vertex
};

constexpr int QC_K = 32;  // classes per axis (16 distinct steps x parity); more -> per-tile kernels
// classes of the plain lattice cells for check_for_inverted_tets (see k_quality_classes)
struct QualityClasses {
    const uint8_t* kx;      // class of every cell column / row / layer (global cell index)
    const uint8_t* ky;
    const uint8_t* kz;
    const double* steps;    // [3][QC_K / 2] distinct step values per axis
    uint32_t* present;      // [kz * QC_K + ky] bit kx
};

__device__ __forceinline__ size_t cidx(const Stuff& S, uint32_t cx, uint32_t cy, uint32_t cz) {
    return ((size_t)(cz - S.cl0) * S.L.ny + cy) * S.L.nx + cx;
}

// ---------------------------------------------------------------------------------------
// block-wide exclusive scan of one uint32 per thread (TILE_THREADS threads); returns the
// exclusive prefix, *total receives the block sum
__device__ __forceinline__ uint32_t block_exscan(uint32_t v, uint32_t* total, uint32_t* smem /*[9]*/) {
    const unsigned lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t n = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= (unsigned)o) inc += n;
    }
    __syncthreads();  // protect smem reuse between calls
    if (lane == 31u) smem[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        uint32_t w = (lane < TILE_THREADS / 32) ? smem[lane] : 0u;
        uint32_t winc = w;
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) {
            const uint32_t n = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= (unsigned)o) winc += n;
        }
        if (lane < TILE_THREADS / 32) smem[lane] = winc - w;
        if (lane == TILE_THREADS / 32 - 1) smem[8] = winc;
    }
    __syncthreads();
    *total = smem[8];
    return smem[wid] + inc - v;
}

// block-wide sum of one uint64 per thread (every thread gets it)
__device__ __forceinline__ unsigned long long block_sum(unsigned long long v, unsigned long long* s_acc) {
    if (threadIdx.x == 0) *s_acc = 0ull;
    __syncthreads();
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31u) == 0u && v) atomicAdd(s_acc, v);
    __syncthreads();
    const unsigned long long r = *s_acc;
    __syncthreads();  // s_acc may be reset by the next call
    return r;
}

// ---------------------------------------------------------------------------------------
// a2: the lattice SDF field, plain evaluator.  Persistent blocks; the program is staged once per
// block into shared memory; each warp evaluates WX x WY points of one z-plane per step with a
// warp-uniform program counter and writes them with coalesced stores.
template <int WX, int WY, bool COUNT, bool SMALL>
__global__ void __launch_bounds__(512)
k_sdf_field(Lattice L, double* __restrict__ phi_out, const uint64_t* __restrict__ prog_g, int nwords,
            int prog_in_smem, uint32_t zplane0, uint32_t nplanes, unsigned long long* __restrict__ flops_total) {
    unsigned long long my_flops = 0ull;
    extern __shared__ uint64_t s_prog[];
    const uint64_t* prog = prog_g;
    if (prog_in_smem) {
        for (int i = threadIdx.x; i < nwords; i += blockDim.x) s_prog[i] = prog_g[i];
        __syncthreads();
        prog = s_prog;
    }
    const uint32_t ntx = (L.NX + WX - 1) / WX, nty = (L.NY + WY - 1) / WY;
    const uint64_t ntiles = (uint64_t)ntx * nty * nplanes;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t lx = lane % WX, ly = lane / WX;
    const uint64_t warps_total = (uint64_t)gridDim.x * (blockDim.x >> 5);
    for (uint64_t tile = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); tile < ntiles;
         tile += warps_total) {
        const uint32_t tx = (uint32_t)(tile % ntx);
        const uint64_t r = tile / ntx;
        const uint32_t ty = (uint32_t)(r % nty);
        const uint32_t Z = zplane0 + (uint32_t)(r / nty);
        const uint32_t X = tx * WX + lx, Y = ty * WY + ly;
        const bool inside = X < L.NX && Y < L.NY;
        // out-of-range lanes evaluate a clamped point so that the warp stays converged
        const uint32_t Xc = X < L.NX ? X : L.NX - 1, Yc = Y < L.NY ? Y : L.NY - 1;
        unsigned long long fl = 0ull;
        const double v = vm_eval_impl<true, COUNT, SMALL>(prog, lat_x(L, Xc), lat_y(L, Yc), lat_z(L, Z), fl);
        if (inside) { phi_out[pidx(L, X, Y, Z)] = v; my_flops += fl; }
    }
    if (COUNT) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) my_flops += __shfl_xor_sync(0xffffffffu, my_flops, o);
        if ((threadIdx.x & 31u) == 0u && my_flops) atomicAdd(flops_total, my_flops);
    }
}

// ObjectAdaptor::value for arbitrary points (src/main.rs:33-35,95): one point per thread
static __global__ void k_eval_points(const uint64_t* __restrict__ prog, const double* __restrict__ pts, uint64_t n,
                                     double* __restrict__ out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t ic = i < n ? i : (n - 1);
    const double v = vm_eval<true>(prog, pts[3 * ic], pts[3 * ic + 1], pts[3 * ic + 2]);
    if (i < n) out[i] = v;
}

// Stage a cube of H^3 lattice points (origin = tile origin + OFF per axis) through a functor:
// f(idx, inside, phi, wcode) for every idx in [0, H^3).  Four points per thread are in flight at
// once: the loads go to clamped (always valid) addresses, so that they are unconditional and can
// be issued back to back — the staging loops are latency bound otherwise.
template <int H, int OFF, bool WCODE, typename F>
__device__ __forceinline__ void stage_points(const Lattice& L, uint32_t X0, uint32_t Y0, uint32_t Z0, F f) {
    constexpr int N = H * H * H, B = 4;
    for (int base = (int)threadIdx.x; base < N; base += B * TILE_THREADS) {
        double v[B];
        uint8_t wc[B];
        bool in[B];
#pragma unroll
        for (int u = 0; u < B; ++u) {
            const int idx = min(base + u * TILE_THREADS, N - 1);
            const int64_t X = (int64_t)X0 + idx % H + OFF, Y = (int64_t)Y0 + (idx / H) % H + OFF, Z = (int64_t)Z0 + idx / (H * H) + OFF;
            in[u] = X >= 0 && Y >= 0 && X < (int64_t)L.NX && Y < (int64_t)L.NY && Z >= (int64_t)L.pz0 && Z <= (int64_t)L.pz1;
            const uint32_t Xc = (uint32_t)min(max(X, (int64_t)0), (int64_t)L.NX - 1), Yc = (uint32_t)min(max(Y, (int64_t)0), (int64_t)L.NY - 1);
            const uint32_t Zc = (uint32_t)min(max(Z, (int64_t)L.pz0), (int64_t)L.pz1);
            const size_t pi = pidx(L, Xc, Yc, Zc);
            const int8_t ts = tile_sign_at(L, Xc, Yc, Zc);  // proven tile: nothing stored, the sign is all there is
            v[u] = ts ? (double)ts : L.phi[pi];
            wc[u] = (WCODE && !ts) ? L.wcode[pi] : (uint8_t)0;
        }
#pragma unroll
        for (int u = 0; u < B; ++u) {
            const int idx = base + u * TILE_THREADS;
            if (idx < N) f(idx, in[u], v[u], wc[u]);
        }
    }
}

// The same staging through the TMA unit: the (TS + 2)^3 box of phi around a tile (origin = tile origin + OFF)
// is brought into shared memory by two bulk tensor copies of BOX_HZ planes each (one thread issues them, the
// block waits on an mbarrier), then f(idx, inside, phi) is called for every idx in [0, 18^3).  Nothing is stored
// for points of proven tiles: their sign comes from the tile-sign table (27 neighbours, staged first).
// The innermost start coordinate of a TMA box must be 16-byte aligned (measured: scripts/probe/tma_probe.cu),
// i.e. even for doubles: the box starts at x = tile origin - 2 and is BOX_W = 20 wide, columns 1..18 (OFF = -1)
// or 2..19 (OFF = 0) are used.
constexpr int BOX_H = (int)TS + 2, BOX_HZ = BOX_H / 2, BOX_W = BOX_H + 2;
constexpr uint32_t BOX_BYTES = (uint32_t)(BOX_HZ * BOX_H * BOX_W * 8);
template <int OFF, typename F>
__device__ __forceinline__ void tma_stage_phi(const Lattice& L, const TileGrid& g, const CUtensorMap* tmap, unsigned long long* bar,
                                              uint32_t& parity, double* box, int8_t* s_ts, uint32_t tile, uint32_t X0, uint32_t Y0,
                                              uint32_t Z0, F f) {
    if (threadIdx.x < 27) {
        const int dx = (int)(threadIdx.x % 3) - 1, dy = (int)((threadIdx.x / 3) % 3) - 1, dz = (int)(threadIdx.x / 9) - 1;
        const int tx = (int)(tile % g.ntx) + dx, ty = (int)((tile / g.ntx) % g.nty) + dy, tz = (int)(tile / (g.ntx * g.nty)) + dz;
        int8_t sgn = 0;
        if (L.tsign != nullptr && tx >= 0 && ty >= 0 && tz >= 0 && tx < (int)g.ntx && ty < (int)g.nty && tz < (int)g.ntz)
            sgn = L.tsign[((size_t)tz * g.nty + ty) * g.ntx + tx];
        s_ts[threadIdx.x] = sgn;
    }
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {
        __syncthreads();  // the box is free again (and the tile signs are visible)
        if (threadIdx.x == 0) {
            fence_proxy_async();
            mbar_expect_tx(bar, BOX_BYTES);
            tma_load_3d(box, tmap, (int)X0 - 2, (int)Y0 + OFF, (int)(Z0 - L.pz0) + OFF + BOX_HZ * half, bar);
        }
        mbar_wait(bar, parity);
        parity ^= 1u;
        for (int i = (int)threadIdx.x; i < BOX_HZ * BOX_H * BOX_H; i += TILE_THREADS) {
            const int bx = i % BOX_H, by = (i / BOX_H) % BOX_H, bz = i / (BOX_H * BOX_H) + BOX_HZ * half;
            const int lx = bx + OFF, ly = by + OFF, lz = bz + OFF;   // relative to the tile origin
            const int64_t X = (int64_t)X0 + lx, Y = (int64_t)Y0 + ly, Z = (int64_t)Z0 + lz;
            const bool in = X >= 0 && Y >= 0 && X < (int64_t)L.NX && Y < (int64_t)L.NY && Z >= (int64_t)L.pz0 && Z <= (int64_t)L.pz1;
            const int tsx = lx < 0 ? 0 : (lx >= (int)TS ? 2 : 1), tsy = ly < 0 ? 0 : (ly >= (int)TS ? 2 : 1), tsz = lz < 0 ? 0 : (lz >= (int)TS ? 2 : 1);
            const int8_t ts = s_ts[(tsz * 3 + tsy) * 3 + tsx];
            f((bz * BOX_H + by) * BOX_H + bx, in, ts ? (double)ts : box[((bz - BOX_HZ * half) * BOX_H + by) * BOX_W + (lx + 2)]);
        }
    }
}

// ---------------------------------------------------------------------------------------
// per-tile sign summary of the raw field: TF_NEG (max < 0), TF_POS (min > 0)
static __global__ void __launch_bounds__(TILE_THREADS)
k_tile_minmax(Stuff S) {
    __shared__ double s_lo[TILE_THREADS / 32], s_hi[TILE_THREADS / 32];
    const Lattice& L = S.L;
    uint32_t X0, Y0, Z0;
    tile_origin(S.g, blockIdx.x, X0, Y0, Z0);
    const double INF = __longlong_as_double(0x7ff0000000000000LL);
    double lo = INF, hi = -INF;
    for (int r = 0; r < TILE_ROUNDS; ++r) {
        const uint32_t j = (uint32_t)r * TILE_THREADS + threadIdx.x;
        const uint32_t X = X0 + (j & 15u), Y = Y0 + ((j >> 4) & 15u), Z = Z0 + (j >> 8);
        if (X < L.NX && Y < L.NY && Z <= L.pz1) {
            const double v = L.phi[pidx(L, X, Y, Z)];
            if (v != v) { lo = -INF; hi = INF; }  // NaN: neither sign
            lo = fmin(lo, v); hi = fmax(hi, v);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if ((threadIdx.x & 31u) == 0u) { s_lo[threadIdx.x >> 5] = lo; s_hi[threadIdx.x >> 5] = hi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < TILE_THREADS / 32; ++w) { lo = fmin(lo, s_lo[w]); hi = fmax(hi, s_hi[w]); }
        S.tflag[blockIdx.x] = (uint8_t)((hi < 0.0 ? TF_NEG : 0) | (lo > 0.0 ? TF_POS : 0));
    }
}

// owned points of a tile: x, y extents and the owned z range [zlo, zhi) (may be empty)
struct TileBox { uint32_t X0, Y0, Z0, nxt, nyt, zlo, zhi; };
__device__ __forceinline__ TileBox tile_box(const Stuff& S, uint64_t t) {
    TileBox b;
    tile_origin(S.g, t, b.X0, b.Y0, b.Z0);
    b.nxt = min(TS, S.L.NX - b.X0);
    b.nyt = min(TS, S.L.NY - b.Y0);
    b.zlo = max(b.Z0, S.oz0);
    b.zhi = min(b.Z0 + TS, S.oz1 + 1);
    if (b.zhi < b.zlo) b.zhi = b.zlo;
    return b;
}

// point tiles: uniformly negative / positive together with everything one lattice step around them?
// Either the interval pass proved it (tile sign: the proof covers the tile's box grown by one step),
// or the stored values say so for the whole 27-neighbourhood of tiles.  Uniform tiles get their
// vertex totals here (interior: every owned point is an output vertex, no cut edges); active and
// interior tiles are appended to the work lists of the vertex kernels.
static __global__ void k_tile_neighbourhood(Stuff S, uint64_t ntiles, unsigned long long* __restrict__ counters,
                                            uint32_t* __restrict__ list_active, uint32_t* __restrict__ list_neg,
                                            unsigned long long* __restrict__ vtile_tot) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ntiles) return;
    uint8_t u;
    if (S.L.tsign != nullptr && S.L.tsign[t] != 0) {
        u = S.L.tsign[t] < 0 ? VU_NEG : VU_POS;
    } else {
        const int tix = (int)(t % S.g.ntx), tiy = (int)((t / S.g.ntx) % S.g.nty), tiz = (int)(t / ((uint64_t)S.g.ntx * S.g.nty));
        uint8_t all = TF_NEG | TF_POS;
        for (int dz = -1; dz <= 1; ++dz)
            for (int dy = -1; dy <= 1; ++dy)
                for (int dx = -1; dx <= 1; ++dx) {
                    const int x = tix + dx, y = tiy + dy, z = tiz + dz;
                    if (x < 0 || y < 0 || z < 0 || x >= (int)S.g.ntx || y >= (int)S.g.nty || z >= (int)S.g.ntz) continue;
                    all &= S.tflag[((size_t)z * S.g.nty + y) * S.g.ntx + x];
                }
        u = (all & TF_NEG) ? VU_NEG : ((all & TF_POS) ? VU_POS : VU_ACTIVE);
    }
    S.vuni[t] = u;
    if (u == VU_ACTIVE) {
        list_active[atomicAdd(&counters[CN_ACTIVE_VTILES], 1ull)] = (uint32_t)t;
    } else {
        const TileBox b = tile_box(S, t);
        const unsigned long long n = u == VU_NEG ? (unsigned long long)b.nxt * b.nyt * (b.zhi - b.zlo) : 0ull;
        vtile_tot[t] = n;
        if (n) list_neg[atomicAdd(&counters[CN_LIST_VNEG], 1ull)] = (uint32_t)t;
    }
}

// cell tiles: the 8 tiles holding the corners are all negative and un-warped (every cell emits
// its five lattice tets) / all positive (nothing is kept)?  Uniform tiles get their tet totals here
// (low 32 bits: output tets, high 32 bits: kept lattice tets, owned cells only).
static __global__ void k_cell_tile_class(Stuff S, uint64_t ntiles, unsigned long long* __restrict__ counters,
                                         uint32_t* __restrict__ list_active, uint32_t* __restrict__ list_full,
                                         unsigned long long* __restrict__ ttile_tot) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ntiles) return;
    const int tix = (int)(t % S.g.ntx), tiy = (int)((t / S.g.ntx) % S.g.nty), tiz = (int)(t / ((uint64_t)S.g.ntx * S.g.nty));
    uint8_t all = TF_NEG | TF_POS, any = 0;
    for (int dz = 0; dz <= 1; ++dz)
        for (int dy = 0; dy <= 1; ++dy)
            for (int dx = 0; dx <= 1; ++dx) {
                const int x = tix + dx, y = tiy + dy, z = tiz + dz;
                if (x >= (int)S.g.ntx || y >= (int)S.g.nty || z >= (int)S.g.ntz) continue;
                const uint8_t f = S.tflag[((size_t)z * S.g.nty + y) * S.g.ntx + x];
                all &= f; any |= f;
            }
    const uint8_t u = ((all & TF_NEG) && !(any & TF_HASWARP)) ? CU_FULL : ((all & TF_POS) ? CU_EMPTY : CU_ACTIVE);
    S.cuni[t] = u;
    const Lattice& L = S.L;
    uint32_t X0, Y0, Z0;
    tile_origin(S.g, t, X0, Y0, Z0);
    const uint32_t zlo = max(Z0, S.c0), zhi = min(Z0 + TS, S.c1);
    const bool has_cells = X0 < L.nx && Y0 < L.ny;
    if (u == CU_ACTIVE) {
        // (classification also covers the halo layers, so every active tile with cells is listed)
        if (has_cells) list_active[atomicAdd(&counters[CN_ACTIVE_CTILES], 1ull)] = (uint32_t)t;
        else ttile_tot[t] = 0ull;
    } else {
        unsigned long long tot = 0ull;
        if (u == CU_FULL && has_cells && zhi > zlo) {
            const unsigned long long n = (unsigned long long)min(TS, L.nx - X0) * min(TS, L.ny - Y0) * (zhi - zlo) * 5ull;
            tot = n | (n << 32);
            list_full[atomicAdd(&counters[CN_LIST_CFULL], 1ull)] = (uint32_t)t;
        }
        ttile_tot[t] = tot;
    }
}

// ---------------------------------------------------------------------------------------
// a6: warp codes (only tiles whose neighbourhood is not of one strict sign can have any).
//
// warp_vertices' visit order around a vertex (8 cells x-major, 5 tets, 6 edges on the flipped
// node order) only depends on the parity of the vertex' coordinates, so it is tabulated once on
// the host (mc_mesher.cu build_warp_table): entry = cell offset | neighbour direction << 3 |
// orientation << 8 | offsets of the tet's 4 corners << 9.  The tile's sign classes (with a
// one-point halo) are staged in shared memory; vertices with an edge that does not join two
// vertices of the same strict sign are collected in a list and processed one per thread.
constexpr int WT_STRIDE = 97;  // [count, 96 entries] per parity class
constexpr int WP_STRIDE = 37;  // [count, <= 36 distinct (direction, orientation) pairs in first-visit order]
constexpr int8_t SG_NEG = -1, SG_ZERO = 0, SG_POS = 1, SG_NONE = 2, SG_NAN = 3;
// C_DIR as compile-time constants: unrolled neighbour scans get immediate shared-memory offsets
struct KDir { int x, y, z; };
__host__ __device__ constexpr KDir kdir(int d) {
    constexpr int T[18][3] = {{1, 0, 0},  {0, 1, 0},  {0, 0, 1},  {1, 1, 0},   {-1, 1, 0}, {1, 0, 1},  {-1, 0, 1}, {0, 1, 1},  {0, -1, 1},
                              {-1, 0, 0}, {0, -1, 0}, {0, 0, -1}, {-1, -1, 0}, {1, -1, 0}, {-1, 0, -1}, {1, 0, -1}, {0, -1, -1}, {0, 1, -1}};
    return KDir{T[d][0], T[d][1], T[d][2]};
}

static __global__ void __launch_bounds__(TILE_THREADS, 4)
k_warp_codes(Stuff S, const __grid_constant__ CUtensorMap tmap_phi, unsigned long long* __restrict__ counters,
             const uint32_t* __restrict__ tile_list, const unsigned long long* __restrict__ n_list,
             unsigned long long* __restrict__ cursor) {
    constexpr int H = (int)TS + 2;
    static_assert(H == BOX_H, "the TMA box is the tile with a one-point halo");
    __shared__ __align__(128) double s_box[BOX_HZ * BOX_H * BOX_W];
    __shared__ __align__(8) unsigned long long s_bar;
    __shared__ int8_t s_ts[27];
    uint32_t parity = 0u;
    if (threadIdx.x == 0) mbar_init(&s_bar, 1u);
    __shared__ int8_t s_sg[H * H * H];
    __shared__ unsigned long long s_tab[8 * WT_STRIDE];
    __shared__ uint8_t s_pairs[8 * WP_STRIDE];
    __shared__ uint32_t s_nan;
    uint16_t* s_list = reinterpret_cast<uint16_t*>(s_box);   // [TILE_POINTS]; the box is dead once the sign classes are staged
    static_assert(sizeof(s_box) >= TILE_POINTS * sizeof(uint16_t), "the vertex list reuses the TMA box");
    __shared__ uint32_t s_n;
    __shared__ uint32_t s_rowm[3][H * H];   // bit masks of the box's point rows
    const Lattice& L = S.L;
    for (int i = (int)threadIdx.x; i < 8 * WT_STRIDE; i += TILE_THREADS) s_tab[i] = S.warp_table[i];
    for (int i = (int)threadIdx.x; i < 8 * WP_STRIDE; i += TILE_THREADS) s_pairs[i] = (uint8_t)S.warp_table[8 * WT_STRIDE + i];
    uint32_t tile;
    while (next_tile(tile_list, n_list, cursor, tile)) {
    uint32_t X0, Y0, Z0;
    tile_origin(S.g, tile, X0, Y0, Z0);
    if (threadIdx.x == 0) { s_n = 0u; s_nan = 0u; }
    tma_stage_phi<-1>(L, S.g, &tmap_phi, &s_bar, parity, s_box, s_ts, tile, X0, Y0, Z0, [&](int idx, bool in, double v) {
        const int8_t sg = !in ? SG_NONE : (v > 0.0 ? SG_POS : (v < 0.0 ? SG_NEG : (v == 0.0 ? SG_ZERO : SG_NAN)));
        s_sg[idx] = sg;
        if (sg == SG_NAN) s_nan = 1u;
    });
    __syncthreads();
    // Which vertices have an edge that does not join two vertices of one strict sign?  Per point row of the box
    // three bit masks (bit bx): the point exists and is not negative / not positive / exists at all ...
    for (int r = (int)threadIdx.x; r < H * H; r += TILE_THREADS) {
        uint32_t ex = 0u, ng = 0u, ps = 0u;
#pragma unroll
        for (int i = 0; i < H; ++i) {
            const int sg = s_sg[r * H + i];
            ex |= (sg != SG_NONE ? 1u : 0u) << i;
            ng |= (sg == SG_NEG ? 1u : 0u) << i;
            ps |= (sg == SG_POS ? 1u : 0u) << i;
        }
        s_rowm[0][r] = ex & ~ng; s_rowm[1][r] = ex & ~ps; s_rowm[2][r] = ex;
    }
    __syncthreads();
    // ... and every thread settles one x-row of 16 vertices: a negative vertex looks for a neighbour that is not
    // negative, a positive one for one that is not positive, any other for any neighbour at all; vertices with
    // even X + Y + Z have 18 neighbours (axes + face diagonals), the others the 6 along the axes.
    {
        const uint32_t ly = threadIdx.x & 15u, lz = threadIdx.x >> 4;
        const uint32_t Y = Y0 + ly, Z = Z0 + lz;
        if (X0 < L.NX && Y < L.NY && Z >= S.wz0 && Z <= S.wz1) {
            const int r = (int)(lz + 1u) * H + (int)(ly + 1u);
            uint32_t any[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const uint32_t* M = s_rowm[k];
                const uint32_t c = M[r], yp = M[r + 1], ym = M[r - 1], zp = M[r + H], zm = M[r - H];
                const uint32_t axis = (c >> 1) | (c << 1) | yp | ym | zp | zm;
                const uint32_t diag = (yp >> 1) | (yp << 1) | (ym >> 1) | (ym << 1) | (zp >> 1) | (zp << 1) | (zm >> 1) | (zm << 1) |
                                      M[r + H + 1] | M[r + H - 1] | M[r - H + 1] | M[r - H - 1];
                any[k] = axis | (diag & (((X0 + Y + Z) & 1u) ? 0x55555555u : 0xaaaaaaaau));   // bit bx <-> X = X0 + bx - 1
            }
            const uint32_t ng = s_rowm[2][r] & ~s_rowm[0][r], ps = s_rowm[2][r] & ~s_rowm[1][r];
            const uint32_t nxt = min(TS, L.NX - X0);
            const uint32_t valid = (nxt >= 16u ? 0xffffu : ((1u << nxt) - 1u)) << 1;
            uint32_t sel = ((ng & any[0]) | (ps & any[1]) | (~ng & ~ps & any[2])) & valid;
            if (sel) {
                uint32_t at = atomicAdd(&s_n, (uint32_t)__popc(sel));
                for (; sel; sel &= sel - 1u) s_list[at++] = (uint16_t)(threadIdx.x * TS + (uint32_t)(__ffs((int)sel) - 2));
            }
        }
    }
    __syncthreads();
    const uint32_t n = s_n;
    unsigned warped_owned = 0;
    bool warped_any = false;
    for (uint32_t q = threadIdx.x; q < n; q += TILE_THREADS) {
        const uint32_t j = s_list[q];
        const uint32_t lx = j & 15u, ly = (j >> 4) & 15u, lz = j >> 8;
        const uint32_t X = X0 + lx, Y = Y0 + ly, Z = Z0 + lz;
        const int me = (int)((lz + 1) * H + (ly + 1)) * H + (int)(lx + 1);
        const int sv = s_sg[me];
        const double pv = L.phi[pidx(L, X, Y, Z)];
        const double xv = lat_x(L, X), yv = lat_y(L, Y), zv = lat_z(L, Z);
        const uint32_t pclass = (X & 1u) | ((Y & 1u) << 1) | ((Z & 1u) << 2);
        bool has = false;
        double best = 0.0;
        uint32_t code = 0u;
        // one candidate: the edge to the neighbour in direction d, seen from this vertex as the
        // first (o = 0) or second (o = 1) node of the tet edge
        auto candidate = [&](uint32_t d, uint32_t o) {
            const uint32_t WX = X + C_DIR[d][0], WY = Y + C_DIR[d][1], WZ = Z + C_DIR[d][2];
            const double pw = L.phi[pidx(L, WX, WY, WZ)];
            // Cheap exact pre-test: the ends differ in sign class, so alpha = num / t lies in [0, 1]
            // (or is NaN).  |num| > 0.3000001 |t| means the rounded quotient is certainly >= 0.3
            // (o = 0; mirror for o = 1: certainly <= 0.7), i.e. no candidate — most edges stop here,
            // before the division and the square root.  NaN / inf make the comparison false.
            const double t = o == 0u ? pv - pw : pw - pv;
            if (o == 0u ? fabs(pv) > 0.3000001 * fabs(t) : fabs(pw) < 0.6999999 * fabs(t)) return;
            const double alpha = o == 0u ? pv / t : pw / t;
            const double dx = lat_x(L, WX) - xv, dy = lat_y(L, WY) - yv, dz = lat_z(L, WZ) - zv;
            double dd;
            if (o == 0u) {
                if (!(alpha < 0.3)) return;
                dd = alpha * sqrt((dx * dx + dy * dy) + dz * dz);
            } else {
                if (!(alpha > 1.0 - 0.3)) return;
                dd = (1.0 - alpha) * sqrt((dx * dx + dy * dy) + dz * dz);
            }
            if (!has || dd < best) { has = true; best = dd; code = 1u + 2u * d + o; }
        };
        if (!s_nan && X >= 1u && X < L.nx && Y >= 1u && Y < L.ny && Z >= 1u && Z < L.nz) {
            // All 8 cells exist and (no NaN) an edge that does not join two vertices of one strict
            // sign has an end with value <= 0, so every tet containing it was kept by cut_lattice:
            // the visit order of the distinct (direction, orientation) pairs is a constant of the
            // parity class.
            const uint8_t* pr = s_pairs + pclass * WP_STRIDE;
            const uint32_t np = pr[0];
            for (uint32_t i = 1; i <= np; ++i) {
                const uint32_t d = pr[i] >> 1, o = pr[i] & 1u;
                const int sw = s_sg[me + (C_DIR[d][2] * H + C_DIR[d][1]) * H + C_DIR[d][0]];
                if ((sv == SG_POS || sv == SG_NEG) && sw == sv) continue;  // both ends of one strict sign
                candidate(d, o);
            }
        } else {
            const unsigned long long* tab = s_tab + pclass * WT_STRIDE;
            const uint32_t nvis = (uint32_t)tab[0];
            unsigned long long seen = 0ull;
            for (uint32_t i = 1; i <= nvis; ++i) {
                const unsigned long long e = tab[i];
                // the cell must exist
                if (((e & 1ull) && X == 0u) || (!(e & 1ull) && X >= L.nx) || ((e & 2ull) && Y == 0u) || (!(e & 2ull) && Y >= L.ny) ||
                    ((e & 4ull) && Z == 0u) || (!(e & 4ull) && Z >= L.nz)) continue;
                const uint32_t d = (uint32_t)(e >> 3) & 31u, o = (uint32_t)(e >> 8) & 1u;
                const uint32_t bit = d * 2u + o;
                if ((seen >> bit) & 1ull) continue;
                const int sw = s_sg[me + (C_DIR[d][2] * H + C_DIR[d][1]) * H + C_DIR[d][0]];
                if ((sv == SG_POS || sv == SG_NEG) && sw == sv) continue;  // both ends of one strict sign
                bool kept = false;  // cut_lattice kept the tet iff a corner has value <= 0
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const uint32_t off = (uint32_t)(e >> (9 + 6 * c)) & 63u;
                    const int cs = s_sg[me + (((int)((off >> 4) & 3u) - 1) * H + ((int)((off >> 2) & 3u) - 1)) * H + ((int)(off & 3u) - 1)];
                    kept = kept || cs == SG_NEG || cs == SG_ZERO;
                }
                if (!kept) continue;
                seen |= 1ull << bit;
                candidate(d, o);
            }
        }
        if (has) {
            L.wcode[pidx(L, X, Y, Z)] = (uint8_t)code;
            warped_any = true;
            warped_owned += (Z >= S.oz0 && Z <= S.oz1) ? 1u : 0u;
        }
    }
    // tflag bytes are updated through their aligned 32-bit word
    if (__any_sync(0xffffffffu, warped_any) && (threadIdx.x & 31u) == 0u)
        atomicOr((unsigned int*)(S.tflag + (tile & ~3u)), (unsigned int)TF_HASWARP << (8u * (tile & 3u)));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) warped_owned += __shfl_xor_sync(0xffffffffu, warped_owned, o);
    if (warped_owned && (threadIdx.x & 31u) == 0u) atomicAdd(&counters[CN_WARPED], (unsigned long long)warped_owned);
    }  // tiles
}

// ---------------------------------------------------------------------------------------
// a5/a7/a9: classify the five tets of every cell: output tets per lattice tet, exterior-test
// results, "full cell" flag; mark zero-valued vertices that appear in the output.  Per-tile
// totals (low 32 bits: output tets, high 32 bits: kept lattice tets) count owned cells only.
// Cells whose corners are all negative / all positive are settled from the staged classes;
// the tets of the remaining (surface) cells are collected and classified one per thread.
// tie-break of the 4-sort between two equal non-zero values: first-touch order of nodes i and k of tet (c, t)
// (rare: symmetric scenes; kept out of line)
static __device__ __noinline__ bool first_touch_less(const Lattice& L, const Cell& c, int t, int i, int k) {
    uint32_t Xi, Yi, Zi, Xk, Yk, Zk;
    corner_xyz(c, tet_corner(c, t, i), Xi, Yi, Zi);
    corner_xyz(c, tet_corner(c, t, k), Xk, Yk, Zk);
    return first_touch_key(L, Xi, Yi, Zi) < first_touch_key(L, Xk, Yk, Zk);
}

constexpr int CC_H = (int)TS + 1, CC_NP = CC_H * CC_H * CC_H;
constexpr uint32_t CC_PASS = TILE_POINTS / 4;    // surface cells assembled per pass
constexpr size_t CLASSIFY_SMEM = (size_t)CC_NP * 8;   // post-warp phi of the tile's cell corners (dynamic shared memory)
static __global__ void __launch_bounds__(TILE_THREADS, 3)
k_classify_cells(Stuff S, const uint64_t* __restrict__ prog, unsigned long long* __restrict__ tile_tot,
                 unsigned long long* __restrict__ counters, const uint32_t* __restrict__ tile_list,
                 const unsigned long long* __restrict__ n_list, unsigned long long* __restrict__ cursor) {
    extern __shared__ __align__(16) double s_phi[];   // [CC_NP] post-warp values (0 for a warped vertex)
    __shared__ unsigned long long s_acc;
    const Lattice& L = S.L;
    uint32_t tile;
    while (next_tile(tile_list, n_list, cursor, tile)) {  // active cell tiles (the others were settled by k_cell_tile_class)
    uint32_t X0, Y0, Z0;
    tile_origin(S.g, tile, X0, Y0, Z0);
    // corner classes of the tile's cells: bit 0 = post-warp phi < 0, bit 1 = raw phi > 0, bit 2 = raw phi <= 0
    constexpr int H = CC_H;
    __shared__ uint8_t s_cls[CC_NP];
    __shared__ uint8_t s_wc[CC_NP];            // warp codes of the corners (final since k_warp_codes)
    __shared__ uint32_t s_rowneg[H * H], s_rowpos[H * H];  // bit i of a point row: class bit 0 / bit 1 of its point i
    __shared__ uint16_t s_list[TILE_POINTS];   // surface cells (local index)
    __shared__ uint32_t s_ci[CC_PASS];         // cinfo of the surface cells being assembled, by list slot
    __shared__ __align__(8) uint8_t s_shape[CC_PASS][8];  // their shape bytes (mc_shapes.hpp), one per lattice tet
    __shared__ uint32_t s_n;
    __shared__ uint32_t s_stat[8];             // trim cases 0..6 + exterior removals of the owned cells
    if (threadIdx.x == 0) s_n = 0u;
    if (threadIdx.x < 8) s_stat[threadIdx.x] = 0u;
    __syncthreads();
    stage_points<H, 0, true>(L, X0, Y0, Z0, [&](int idx, bool in, double raw, uint8_t wc) {
        const double w = (wc & 0x7f) ? 0.0 : raw;
        s_phi[idx] = w;
        s_wc[idx] = (uint8_t)(wc & 0x7f);
        s_cls[idx] = in ? (uint8_t)((w < 0.0 ? 1 : 0) | (raw > 0.0 ? 2 : 0) | (raw <= 0.0 ? 4 : 0)) : (uint8_t)0;
    });
    __syncthreads();
    for (int r = (int)threadIdx.x; r < H * H; r += TILE_THREADS) {
        uint32_t neg = 0u, pos = 0u;
#pragma unroll
        for (int i = 0; i < H; ++i) {
            const uint32_t c = s_cls[r * H + i];
            neg |= (c & 1u) << i;
            pos |= ((c >> 1) & 1u) << i;
        }
        s_rowneg[r] = neg; s_rowpos[r] = pos;
    }
    __syncthreads();
    uint32_t my_out = 0, my_kept = 0, my_listed = 0, my_surf = 0;
    {
        // every thread settles one x-row of 16 cells from the four point rows around it: a cell whose eight
        // corners are all negative emits five untouched lattice tets, all positive: nothing; the rest is listed
        const uint32_t ly = threadIdx.x & 15u, lz = threadIdx.x >> 4;
        const uint32_t cy = Y0 + ly, cz = Z0 + lz;
        if (X0 < L.nx && cy < L.ny && cz >= S.cl0 && cz < S.cl1) {
            const uint32_t nxt = min(TS, L.nx - X0), valid = nxt >= 16u ? 0xffffu : ((1u << nxt) - 1u);
            const uint32_t r00 = lz * H + ly, r01 = r00 + 1u, r10 = r00 + H, r11 = r10 + 1u;
            const uint32_t n4 = s_rowneg[r00] & s_rowneg[r01] & s_rowneg[r10] & s_rowneg[r11];
            const uint32_t p4 = s_rowpos[r00] & s_rowpos[r01] & s_rowpos[r10] & s_rowpos[r11];
            const uint32_t full = n4 & (n4 >> 1) & valid, empty = p4 & (p4 >> 1) & valid;
            const uint32_t surf = valid & ~full & ~empty;
            uint16_t* row = S.cinfo + cidx(S, X0, cy, cz);
            if (nxt == 16u && (reinterpret_cast<uintptr_t>(row) & 15u) == 0u) {
                uint32_t wds[8];
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    wds[i] = (((full >> (2 * i)) & 1u) ? (uint32_t)CI_ALL_FULL : 0u) | ((((full >> (2 * i + 1)) & 1u) ? (uint32_t)CI_ALL_FULL : 0u) << 16);
                reinterpret_cast<uint4*>(row)[0] = make_uint4(wds[0], wds[1], wds[2], wds[3]);
                reinterpret_cast<uint4*>(row)[1] = make_uint4(wds[4], wds[5], wds[6], wds[7]);
            } else {
                for (uint32_t i = 0; i < nxt; ++i) row[i] = ((full >> i) & 1u) ? CI_ALL_FULL : (uint16_t)0;
            }
            if (cz >= S.c0 && cz < S.c1) { my_out += 5u * (uint32_t)__popc(full); my_kept += 5u * (uint32_t)__popc(full); }
            if (surf) {
                uint32_t at = atomicAdd(&s_n, (uint32_t)__popc(surf));
                for (uint32_t m = surf; m; m &= m - 1u) s_list[at++] = (uint16_t)(threadIdx.x * TS + (uint32_t)(__ffs((int)m) - 1));
            }
        }
    }
    __syncthreads();
    const uint32_t n = s_n;
    // a tile has at most 4096 surface cells; the assembly buffer holds CC_PASS list slots per pass
    for (uint32_t pass = 0; pass < n; pass += CC_PASS) {
        const uint32_t m = min(n - pass, CC_PASS);
        for (uint32_t q = threadIdx.x; q < m; q += TILE_THREADS) {
            s_ci[q] = 0u;
            *reinterpret_cast<unsigned long long*>(s_shape[q]) = 0ull;
        }
        __syncthreads();
        for (uint32_t w = threadIdx.x; w < 5u * m; w += TILE_THREADS) {
            const uint32_t q = w / 5u, t = w % 5u;
            const uint32_t j = s_list[pass + q];
            const uint32_t cx = X0 + (j & 15u), cy = Y0 + ((j >> 4) & 15u), cz = Z0 + (j >> 8);
            const Cell c = make_cell(cx, cy, cz);
            const bool own = cz >= S.c0 && cz < S.c1;  // halo layers only contribute flags
            // the four nodes (flipped order): local index in the staged box, sign classes, post-warp values
            int li[4];
            uint8_t andc = 3, orc = 0;
#pragma unroll
            for (int nn = 0; nn < 4; ++nn) {
                uint32_t X, Y, Z;
                corner_xyz(c, tet_corner(c, (int)t, nn), X, Y, Z);
                li[nn] = (int)((Z - Z0) * H + (Y - Y0)) * H + (int)(X - X0);
                const uint8_t cl = s_cls[li[nn]];
                andc &= cl; orc |= cl;
            }
            if (!(orc & 4)) continue;   // no corner with a raw value <= 0: not kept by cut_lattice
            if (andc & 1) {             // all strictly negative: passed through untouched
                atomicOr(&s_ci[q], (1u << (2u * t)) | (1u << (16u + t)));
                s_shape[q][t] = (uint8_t)SHAPE_WHOLE;
                continue;
            }
            const double p0 = s_phi[li[0]], p1 = s_phi[li[1]], p2 = s_phi[li[2]], p3 = s_phi[li[3]];
            uint32_t nout, shape;
            uint32_t bits = 1u << (16u + t);  // bit 16+t: kept
            if (p0 <= 0.0 && p1 <= 0.0 && p2 <= 0.0 && p3 <= 0.0) {
                // no positive vertex: passed through, unless all four are zero and the exterior test drops it
                nout = 1u; shape = SHAPE_WHOLE;
                if (p0 == 0.0 && p1 == 0.0 && p2 == 0.0 && p3 == 0.0) {
                    TetNodes tn;
                    load_tet(L, c, (int)t, tn);
                    TetOut to;
                    classify_tet<true, true>(L, c, (int)t, tn, prog, -1, to);
                    if (to.ext_removed) {
                        nout = 0u; shape = SHAPE_NONE;
                        bits |= 1u << (10u + t);   // bit 10+t: dropped by the exterior test
                        if (own) atomicAdd(&s_stat[7], 1u);
                    }
                }
            } else {
                int kase, perm;
                sort_case(p0, p1, p2, p3, [&](int i, int k) { return first_touch_less(L, c, (int)t, i, k); }, kase, perm);
                nout = C_CASE_N[kase];
                shape = nout ? (uint32_t)(SHAPE_TRIM0 + 24 * (kase - 1) + perm) : (uint32_t)SHAPE_NONE;
                bits |= 1u << (24u + t);           // bit 24+t: trimmed
                if (own) atomicAdd(&s_stat[kase], 1u);
            }
            bits |= nout << (2u * t);
            atomicOr(&s_ci[q], bits);
            s_shape[q][t] = (uint8_t)shape;
            // zero-valued original vertices that made it into the output are "used"
            const uint32_t used = S.shapes->used[shape];
#pragma unroll
            for (int nn = 0; nn < 4; ++nn) {
                const double pn = nn == 0 ? p0 : (nn == 1 ? p1 : (nn == 2 ? p2 : p3));
                if (((used >> nn) & 1u) && pn == 0.0) {
                    uint32_t X, Y, Z;
                    corner_xyz(c, tet_corner(c, (int)t, nn), X, Y, Z);
                    // (a plain store: the warp code in the low bits is final, everybody who marks the vertex writes the same byte)
                    L.wcode[pidx(L, X, Y, Z)] = (uint8_t)(s_wc[li[nn]] | 0x80u);
                }
            }
        }
        __syncthreads();
        for (uint32_t q = threadIdx.x; q < m; q += TILE_THREADS) {
            const uint32_t j = s_list[pass + q];
            const uint32_t cx = X0 + (j & 15u), cy = Y0 + ((j >> 4) & 15u), cz = Z0 + (j >> 8);
            const uint32_t b = s_ci[q];
            const uint32_t kept_bits = (b >> 16) & 31u, removed_bits = (b >> 10) & 31u, trimmed_bits = (b >> 24) & 31u;
            const uint32_t untouched_bits = kept_bits & ~removed_bits & ~trimmed_bits;
            uint16_t ci = (uint16_t)((b & 0x3ffu) | (untouched_bits << CI_UNTOUCHED_SHIFT));
            const uint32_t kept = (uint32_t)__popc(kept_bits);
            // (on a FULL cell bit 10 doubles as CI_FULL_ZERO_CORNER: all five untouched bits are set anyway)
            if (untouched_bits == 31u && ci_count(ci) == 5u) ci |= (uint16_t)(CI_FULL | CI_FULL_ZERO_CORNER);
            S.cinfo[cidx(S, cx, cy, cz)] = ci;
            S.cshape[cidx(S, cx, cy, cz)] = *reinterpret_cast<const unsigned long long*>(s_shape[q]);
            if (cz >= S.c0 && cz < S.c1) {
                my_out += ci_count(ci); my_kept += kept; my_listed += ci_count(ci);
                if (ci_count(ci) != 0u && !(ci & CI_FULL)) ++my_surf;
            }
        }
        __syncthreads();
    }
    const unsigned long long tot = block_sum((unsigned long long)my_out | ((unsigned long long)my_kept << 32), &s_acc);
    if (threadIdx.x == 0) tile_tot[tile] = tot;
    const unsigned long long listed = block_sum((unsigned long long)my_listed | ((unsigned long long)my_surf << 32), &s_acc);
    if (threadIdx.x == 0 && listed) {
        if (listed & 0xffffffffull) atomicAdd(&counters[CN_LISTED_TETS], listed & 0xffffffffull);
        if (listed >> 32) atomicAdd(&counters[CN_SURF_CELLS], listed >> 32);
    }
    // the statistics leave the block as at most 8 global atomics
    if (threadIdx.x < 8 && s_stat[threadIdx.x])
        atomicAdd(&counters[threadIdx.x < 7 ? CN_STAT0 + (int)threadIdx.x : (int)CN_EXT_REMOVED], (unsigned long long)s_stat[threadIdx.x]);
    }  // tiles
}

// ---------------------------------------------------------------------------------------
// exclusive scan of per-tile totals in two launches.  Entries pack two 32-bit counters; the low
// halves are scanned into base_lo, the high halves into base_hi (either may be null); totals[0] /
// totals[1] receive the sums.
//   k_scan_tiles_local: every block scans SCAN_BLOCK entries in place (block-local prefixes) and
//                       writes its pair of sums to block_sums
//   k_scan_tiles_add:   every block adds the sums of the blocks before it (<= a few hundred values,
//                       summed by the block itself: no third launch, no look-back spinning)
constexpr int SCAN_THREADS = 256, SCAN_ITEMS = 8, SCAN_BLOCK = SCAN_THREADS * SCAN_ITEMS;
static __global__ void __launch_bounds__(SCAN_THREADS)
k_scan_tiles_local(const unsigned long long* __restrict__ tot, uint64_t n, unsigned long long* __restrict__ base_lo,
                   unsigned long long* __restrict__ base_hi, unsigned long long* __restrict__ block_sums) {
    __shared__ unsigned long long s_w[2][SCAN_THREADS / 32];
    const unsigned lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
    const uint64_t first = (uint64_t)blockIdx.x * SCAN_BLOCK + (uint64_t)threadIdx.x * SCAN_ITEMS;
    unsigned long long v0[SCAN_ITEMS], v1[SCAN_ITEMS], a0 = 0ull, a1 = 0ull;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        const unsigned long long e = first + i < n ? tot[first + i] : 0ull;
        v0[i] = a0; v1[i] = a1;           // exclusive prefix inside the thread
        a0 += e & 0xffffffffull; a1 += e >> 32;
    }
    unsigned long long i0 = a0, i1 = a1;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long n0 = __shfl_up_sync(0xffffffffu, i0, o), n1 = __shfl_up_sync(0xffffffffu, i1, o);
        if (lane >= (unsigned)o) { i0 += n0; i1 += n1; }
    }
    if (lane == 31u) { s_w[0][wid] = i0; s_w[1][wid] = i1; }
    __syncthreads();
    unsigned long long w0 = 0ull, w1 = 0ull, t0 = 0ull, t1 = 0ull;
#pragma unroll
    for (int w = 0; w < SCAN_THREADS / 32; ++w) {
        if (w < (int)wid) { w0 += s_w[0][w]; w1 += s_w[1][w]; }
        t0 += s_w[0][w]; t1 += s_w[1][w];
    }
    const unsigned long long ex0 = w0 + i0 - a0, ex1 = w1 + i1 - a1;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i)
        if (first + i < n) {
            if (base_lo) base_lo[first + i] = ex0 + v0[i];
            if (base_hi) base_hi[first + i] = ex1 + v1[i];
        }
    if (threadIdx.x == 0) { block_sums[2 * blockIdx.x] = t0; block_sums[2 * blockIdx.x + 1] = t1; }
}
static __global__ void __launch_bounds__(SCAN_THREADS)
k_scan_tiles_add(uint64_t n, unsigned long long* __restrict__ base_lo, unsigned long long* __restrict__ base_hi,
                 const unsigned long long* __restrict__ block_sums, unsigned long long* __restrict__ totals) {
    __shared__ unsigned long long s_acc;
    // sums of the blocks before this one (block 0 also produces the grand totals)
    unsigned long long p0 = 0ull, p1 = 0ull, g0 = 0ull, g1 = 0ull;
    for (unsigned q = threadIdx.x; q < gridDim.x; q += SCAN_THREADS) {
        const unsigned long long b0 = block_sums[2 * q], b1 = block_sums[2 * q + 1];
        if (q < blockIdx.x) { p0 += b0; p1 += b1; }
        g0 += b0; g1 += b1;
    }
    p0 = block_sum(p0, &s_acc);
    p1 = block_sum(p1, &s_acc);
    if (blockIdx.x == 0) {
        g0 = block_sum(g0, &s_acc);
        g1 = block_sum(g1, &s_acc);
        if (threadIdx.x == 0) { totals[0] = g0; totals[1] = g1; }
    }
    if (blockIdx.x == 0) return;  // nothing to add
    const uint64_t first = (uint64_t)blockIdx.x * SCAN_BLOCK;
    for (uint32_t i = threadIdx.x; i < (uint32_t)SCAN_BLOCK; i += SCAN_THREADS)
        if (first + i < n) {
            if (base_lo) base_lo[first + i] += p0;
            if (base_hi) base_hi[first + i] += p1;
        }
}

// ---------------------------------------------------------------------------------------
// a10 (vertices), pass 1: for every owned lattice vertex decide whether it is in the output
// and which of its nine owned edges carry a cut vertex.  Tile totals: low = output vertices,
// high = cut edges.
static __global__ void __launch_bounds__(TILE_THREADS)
k_vertex_count(Stuff S, unsigned long long* __restrict__ tile_tot, const uint32_t* __restrict__ tile_list,
               const unsigned long long* __restrict__ n_list, unsigned long long* __restrict__ cursor) {
    __shared__ unsigned long long s_acc;
    const Lattice& L = S.L;
    uint32_t tile;
    while (next_tile(tile_list, n_list, cursor, tile)) {  // active point tiles (the others were settled by k_tile_neighbourhood)
    const TileBox b = tile_box(S, tile);
    // post-warp classes of the tile's points with a one-point halo:
    // 0 none, 1 negative, 2 positive, 3 zero (not in the output), 4 zero used by some output tet
    constexpr int H = (int)TS + 2;
    __shared__ uint8_t s_cl[H * H * H];
    stage_points<H, -1, true>(L, b.X0, b.Y0, b.Z0, [&](int idx, bool in, double raw, uint8_t wc) {
        const double pv = (wc & 0x7f) ? 0.0 : raw;
        s_cl[idx] = !in ? (uint8_t)0 : (pv < 0.0 ? 1 : (pv > 0.0 ? 2 : (pv == 0.0 ? ((wc & 0x80) ? 4 : 3) : 0)));
    });
    __syncthreads();
    // bit masks of the box's point rows (bit bx): negative / positive / zero-but-used
    __shared__ uint32_t s_rn[H * H], s_rp[H * H], s_ru[H * H];
    for (int r = (int)threadIdx.x; r < H * H; r += TILE_THREADS) {
        uint32_t ng = 0u, ps = 0u, us = 0u;
#pragma unroll
        for (int i = 0; i < H; ++i) {
            const uint32_t c = s_cl[r * H + i];
            ng |= (c == 1u ? 1u : 0u) << i;
            ps |= (c == 2u ? 1u : 0u) << i;
            us |= (c == 4u ? 1u : 0u) << i;
        }
        s_rn[r] = ng; s_rp[r] = ps; s_ru[r] = us;
    }
    __syncthreads();
    // every thread settles one x-row of 16 vertices: the edge along canonical direction s is cut iff its ends
    // have strictly opposite signs — nine row masks, then transposed into the per-vertex words
    unsigned long long mine = 0ull;
    {
        const uint32_t ly = threadIdx.x & 15u, lz = threadIdx.x >> 4;
        const uint32_t Y = b.Y0 + ly, Z = b.Z0 + lz;
        if (b.X0 < L.NX && Y < L.NY && Z >= b.zlo && Z < b.zhi) {
            const int r = (int)(lz + 1u) * H + (int)(ly + 1u);
            const uint32_t nxt = min(TS, L.NX - b.X0);
            const uint32_t valid = (nxt >= 16u ? 0xffffu : ((1u << nxt) - 1u)) << 1;
            const uint32_t ng = s_rn[r], ps = s_rp[r];
            // vertices with even X + Y + Z own nine edges, the others the three along the axes (bit bx <-> X = X0 + bx - 1)
            const uint32_t even = ((b.X0 + Y + Z) & 1u) ? 0x55555555u : 0xaaaaaaaau;
            uint32_t cut[9];
#pragma unroll
            for (int sdir = 0; sdir < 9; ++sdir) {
                const KDir kd = kdir(sdir);
                const int rn = r + kd.z * H + kd.y;
                uint32_t nn = s_rn[rn], pp = s_rp[rn];
                if (kd.x > 0) { nn >>= 1; pp >>= 1; }
                if (kd.x < 0) { nn <<= 1; pp <<= 1; }
                cut[sdir] = ((ng & pp) | (ps & nn)) & valid & (sdir < 3 ? 0xffffffffu : even);
            }
            const uint32_t used = (ng | s_ru[r]) & valid;
            uint32_t ncut = 0u;
#pragma unroll
            for (int sdir = 0; sdir < 9; ++sdir) ncut += (uint32_t)__popc(cut[sdir]);
            mine = (unsigned long long)(ncut + (uint32_t)__popc(used)) | ((unsigned long long)ncut << 32);
            uint32_t wds[8];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                uint32_t m = ((used >> (i + 1)) & 1u) ? (uint32_t)VM_USED : 0u;
#pragma unroll
                for (int sdir = 0; sdir < 9; ++sdir) m |= ((cut[sdir] >> (i + 1)) & 1u) << sdir;
                if (i & 1) wds[i >> 1] |= m << 16; else wds[i >> 1] = m;
            }
            // the row of masks is 32 aligned bytes (row pitch and tile origin are multiples of 16 points)
            uint4* out = reinterpret_cast<uint4*>(S.vmask + pidx(L, b.X0, Y, Z));
            out[0] = make_uint4(wds[0], wds[1], wds[2], wds[3]);
            out[1] = make_uint4(wds[4], wds[5], wds[6], wds[7]);
        }
    }
    const unsigned long long tot = block_sum(mine, &s_acc);
    if (threadIdx.x == 0) tile_tot[tile] = tot;
    }  // tiles
}

// global id of a lattice vertex's first output + its mask.  Interior tiles are pure arithmetic.
__device__ __forceinline__ long long vertex_gid(const Stuff& S, uint32_t X, uint32_t Y, uint32_t Z, uint16_t& mask) {
    const Lattice& L = S.L;
    if (S.lower_table != nullptr && Z == S.c0) {
        const unsigned long long e = S.lower_table[(size_t)Y * L.NX + X];
        mask = (uint16_t)(e & 0xffffull);
        return (long long)(e >> 16);
    }
    const uint32_t t = tile_of(S.g, X, Y, Z);
    if (S.vuni[t] == VU_NEG) {
        const TileBox b = tile_box(S, t);
        mask = VM_USED;
        return S.vertex_base + (long long)S.tvbase[t] +
               (long long)(((Z - b.zlo) * b.nyt + (Y - b.Y0)) * b.nxt + (X - b.X0));
    }
    const size_t pi = pidx(L, X, Y, Z);
    mask = S.vmask[pi];
    return S.vertex_base + (long long)S.tvbase[t] + (long long)S.vloc[pi];
}
// id of the cut vertex on the edge leaving the owner along canonical slot s
__device__ __forceinline__ long long cut_vertex_id(long long owner_base, uint16_t owner_mask, int s) {
    return owner_base + ((owner_mask & VM_USED) ? 1 : 0) + __popc(owner_mask & ((1u << s) - 1u));
}

// a10 (vertices), pass 2: number the output vertices tile by tile (lattice vertex first, then
// its cut edges in slot order), write the coordinates of the lattice vertices (warped where
// applicable) and append the cut edges to the bisection work list.
// interior point tiles: every owned point is an un-warped output vertex, id = tile base + linear
// index.  256 vertices at a time go through shared memory so that the xyz triples leave as three
// fully coalesced stores per thread instead of 24-byte strided ones.
static __global__ void __launch_bounds__(TILE_THREADS)
k_vertex_emit_interior(Stuff S, double* __restrict__ coords, const uint32_t* __restrict__ tile_list,
                       const unsigned long long* __restrict__ n_list, unsigned long long* __restrict__ cursor) {
    __shared__ double s_xyz[3 * TILE_THREADS];
    const Lattice& L = S.L;
    uint32_t tile;
    while (next_tile(tile_list, n_list, cursor, tile)) {
        const TileBox b = tile_box(S, tile);
        const uint64_t vb = S.tvbase[tile];
        const uint32_t n = b.nxt * b.nyt * (b.zhi - b.zlo);
        for (uint32_t base = 0; base < n; base += TILE_THREADS) {
            const uint32_t j = base + threadIdx.x;
            if (j < n) {
                const uint32_t x = j % b.nxt, y = (j / b.nxt) % b.nyt, z = j / (b.nxt * b.nyt);
                s_xyz[3 * threadIdx.x] = lat_x(L, b.X0 + x);
                s_xyz[3 * threadIdx.x + 1] = lat_y(L, b.Y0 + y);
                s_xyz[3 * threadIdx.x + 2] = lat_z(L, b.zlo + z);
            }
            __syncthreads();
            const uint32_t m3 = 3u * min((uint32_t)TILE_THREADS, n - base);
            double* o = coords + 3 * (vb + base);
            for (uint32_t e = threadIdx.x; e < m3; e += TILE_THREADS) o[e] = s_xyz[e];
            __syncthreads();
        }
    }
}

static __global__ void __launch_bounds__(TILE_THREADS)
k_vertex_emit(Stuff S, const unsigned long long* __restrict__ tile_cbase, double* __restrict__ coords,
              unsigned long long* __restrict__ cutlist, const uint32_t* __restrict__ tile_list,
              const unsigned long long* __restrict__ n_list, unsigned long long* __restrict__ cursor) {
    __shared__ uint32_t s_scan[9];
    const Lattice& L = S.L;
    uint32_t tile;
    while (next_tile(tile_list, n_list, cursor, tile)) {  // active point tiles
    const TileBox b = tile_box(S, tile);
    if (b.zhi == b.zlo) continue;
    const uint64_t vb = S.tvbase[tile];
    // Numbering: every thread takes one x-row of 16 points (consecutive in the numbering) and one
    // block scan of the row totals (low half: output vertices, high half: cut edges; <= 40960 /
    // 36864 per tile) gives each point's first local id and cut-list slot ...
    __shared__ uint16_t s_m[TILE_POINTS], s_loc[TILE_POINTS], s_cut[TILE_POINTS];
    {
        const uint32_t ly = threadIdx.x & 15u, lz = threadIdx.x >> 4;
        const uint32_t Y = b.Y0 + ly, Z = b.Z0 + lz;
        const bool row_in = Y < L.NY && Z >= b.zlo && Z < b.zhi;
        const uint32_t nrow = row_in ? min(TS, L.NX - b.X0) : 0u;
        const uint16_t* __restrict__ row = S.vmask + (row_in ? pidx(L, b.X0, Y, Z) : 0);
        uint32_t acc = 0;
        uint32_t pre[TS];
#pragma unroll
        for (uint32_t i = 0; i < TS; ++i) {
            const uint16_t m = i < nrow ? row[i] : (uint16_t)0;
            s_m[threadIdx.x * TS + i] = m;
            const uint32_t ncut = (uint32_t)__popc(m & VM_CUT_MASK);
            pre[i] = acc;
            acc += (ncut + ((m & VM_USED) ? 1u : 0u)) | (ncut << 16);
        }
        uint32_t tot;
        const uint32_t ex = block_exscan(acc, &tot, s_scan);
#pragma unroll
        for (uint32_t i = 0; i < TS; ++i) {
            const uint32_t v = ex + pre[i];
            s_loc[threadIdx.x * TS + i] = (uint16_t)(v & 0xffffu);
            s_cut[threadIdx.x * TS + i] = (uint16_t)(v >> 16);
        }
    }
    __syncthreads();
    // ... and the outputs are written with consecutive lanes on consecutive points
    const uint64_t cb = tile_cbase[tile];
    for (int r = 0; r < TILE_ROUNDS; ++r) {
        const uint32_t j = (uint32_t)r * TILE_THREADS + threadIdx.x;
        const uint32_t X = b.X0 + (j & 15u), Y = b.Y0 + ((j >> 4) & 15u), Z = b.Z0 + (j >> 8);
        if (!(X < L.NX && Y < L.NY && Z >= b.zlo && Z < b.zhi)) continue;
        const size_t pi = pidx(L, X, Y, Z);
        const uint16_t m = s_m[j];
        const uint32_t loc = s_loc[j];
        S.vloc[pi] = (uint16_t)loc;
        if (m & VM_USED) {
            double p[3];
            warped_pos(L, X, Y, Z, p);
            double* o = coords + 3 * (vb + loc);
            o[0] = p[0]; o[1] = p[1]; o[2] = p[2];
        }
        uint64_t ci = cb + s_cut[j];
        for (uint32_t cm = m & VM_CUT_MASK; cm; cm &= cm - 1u)
            cutlist[ci++] = ((unsigned long long)pi << 4) | (unsigned long long)(__ffs((int)cm) - 1);
    }
    }  // tiles
}

// a8: cut_edge — bisection of the real SDF between the positive and the negative end of
// every cut edge until |phi| <= resolution * relative_error (tessellate3d/src/lib.rs:198-210).
// One edge per thread; lanes that have converged idle until the warp is done so that the
// evaluator stays warp-uniform.
static __global__ void __launch_bounds__(256)
k_bisect_edges(Stuff S, const uint64_t* __restrict__ prog, const unsigned long long* __restrict__ cutlist,
               uint64_t ncut, double error_threshold, uint32_t max_iters, double* __restrict__ coords,
               unsigned long long* __restrict__ counters) {
    const Lattice& L = S.L;
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = i < ncut;
    const unsigned long long e = cutlist[active ? i : (ncut - 1)];
    const size_t pi = (size_t)(e >> 4);
    const int s = (int)(e & 15ull);
    const uint64_t per_plane = (uint64_t)L.PX * L.NY;
    const uint32_t Z = L.pz0 + (uint32_t)(pi / per_plane);
    const uint64_t r = pi % per_plane;
    const uint32_t Y = (uint32_t)(r / L.PX), X = (uint32_t)(r % L.PX);
    const uint32_t WX = X + C_DIR[s][0], WY = Y + C_DIR[s][1], WZ = Z + C_DIR[s][2];
    const double pv = L.phi[pi];
    // `a` starts at the positive end (every call site passes the phi > 0 vertex first)
    double ax, ay, az, bx, by, bz;
    if (pv > 0.0) {
        ax = lat_x(L, X); ay = lat_y(L, Y); az = lat_z(L, Z);
        bx = lat_x(L, WX); by = lat_y(L, WY); bz = lat_z(L, WZ);
    } else {
        bx = lat_x(L, X); by = lat_y(L, Y); bz = lat_z(L, Z);
        ax = lat_x(L, WX); ay = lat_y(L, WY); az = lat_z(L, WZ);
    }
    double cx = 0.0, cy = 0.0, cz = 0.0;
    bool done = !active;
    uint32_t iters = 0;
    while (__any_sync(0xffffffffu, !done)) {
        const double mx = (ax + bx) * 0.5, my = (ay + by) * 0.5, mz = (az + bz) * 0.5;
        const double phi = vm_eval<true>(prog, mx, my, mz);
        if (!done) {
            cx = mx; cy = my; cz = mz;
            ++iters;
            if (fabs(phi) <= error_threshold) done = true;
            else if (phi > 0.0) { ax = mx; ay = my; az = mz; }   // phia stays positive
            else { bx = mx; by = my; bz = mz; }
            if (!done && iters >= max_iters) {
                done = true;
                atomicAdd(&counters[CN_BISECT_FAIL], 1ull);
            }
        }
    }
    if (active) {
        uint16_t mask;
        const long long base = vertex_gid(S, X, Y, Z, mask);
        const long long id = cut_vertex_id(base, mask, s) - S.vertex_base;
        coords[3 * id] = cx; coords[3 * id + 1] = cy; coords[3 * id + 2] = cz;
    }
    unsigned long long it64 = active ? iters : 0u;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) it64 += __shfl_xor_sync(0xffffffffu, it64, o);
    if ((threadIdx.x & 31u) == 0u && it64) atomicAdd(&counters[CN_BISECT_ITERS], it64);
}

// ---------------------------------------------------------------------------------------
// a11: Mesh::volume and Mesh::aspect_ratio (mesh/src/lib.rs:292-316)
__device__ __forceinline__ void tet_quality(const double a[3], const double b[3], const double c[3],
                                            const double d[3], double& vol, double& ar) {
    // determinant of the matrix with columns a-d, b-d, c-d (nalgebra 3x3 closed form)
    const double m11 = a[0] - d[0], m21 = a[1] - d[1], m31 = a[2] - d[2];
    const double m12 = b[0] - d[0], m22 = b[1] - d[1], m32 = b[2] - d[2];
    const double m13 = c[0] - d[0], m23 = c[1] - d[1], m33 = c[2] - d[2];
    const double minor_m12_m23 = m22 * m33 - m32 * m23;
    const double minor_m11_m23 = m21 * m33 - m31 * m23;
    const double minor_m11_m22 = m21 * m32 - m31 * m22;
    vol = (m11 * minor_m12_m23 - m12 * minor_m11_m23 + m13 * minor_m11_m22) / 6.0;
    // Tet4::face / Tet4::edge tables as compile-time constants so that everything stays in registers
    constexpr int F[4][3] = {{0, 1, 2}, {0, 2, 3}, {0, 3, 1}, {1, 3, 2}};
    constexpr int E[6][2] = {{0, 1}, {0, 2}, {0, 3}, {1, 2}, {1, 3}, {2, 3}};
    const double P[4][3] = {{a[0], a[1], a[2]}, {b[0], b[1], b[2]}, {c[0], c[1], c[2]}, {d[0], d[1], d[2]}};
    double ab[3], ac[3], ad[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) { ab[k] = b[k] - a[k]; ac[k] = c[k] - a[k]; ad[k] = d[k] - a[k]; }
    const double cr0 = ac[1] * ad[2] - ac[2] * ad[1], cr1 = ac[2] * ad[0] - ac[0] * ad[2], cr2 = ac[0] * ad[1] - ac[1] * ad[0];
    const double alpha = fabs((ab[0] * cr0 + ab[1] * cr1) + ab[2] * cr2);
    double sum_normals = 0.0;
#pragma unroll
    for (int f = 0; f < 4; ++f) {
        const double u0 = P[F[f][1]][0] - P[F[f][0]][0], u1 = P[F[f][1]][1] - P[F[f][0]][1], u2 = P[F[f][1]][2] - P[F[f][0]][2];
        const double v0 = P[F[f][2]][0] - P[F[f][0]][0], v1 = P[F[f][2]][1] - P[F[f][0]][1], v2 = P[F[f][2]][2] - P[F[f][0]][2];
        const double n0 = u1 * v2 - u2 * v1, n1 = u2 * v0 - u0 * v2, n2 = u0 * v1 - u1 * v0;
        sum_normals += sqrt((n0 * n0 + n1 * n1) + n2 * n2);
    }
    double max_length2 = 0.0;
#pragma unroll
    for (int e = 0; e < 6; ++e) {
        const double u0 = P[E[e][1]][0] - P[E[e][0]][0], u1 = P[E[e][1]][1] - P[E[e][0]][1], u2 = P[E[e][1]][2] - P[E[e][0]][2];
        max_length2 = fmax(max_length2, (u0 * u0 + u1 * u1) + u2 * u2);
    }
    const double max_length = sqrt(max_length2);
    const double radius = alpha / sum_normals;
    ar = 1.0 / (max_length / radius / (2.0 * sqrt(6.0)));
}

// per-tet quality for explicit arrays (mc_tet_quality)
static __global__ void k_tet_quality_arrays(const double* __restrict__ coords, const unsigned long long* __restrict__ tets,
                                            uint64_t nt, double* __restrict__ vol, double* __restrict__ ar) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nt) return;
    double p[4][3];
#pragma unroll
    for (int n = 0; n < 4; ++n)
#pragma unroll
        for (int k = 0; k < 3; ++k) p[n][k] = coords[3 * tets[4 * i + n] + k];
    tet_quality(p[0], p[1], p[2], p[3], vol[i], ar[i]);
}

struct QualityAcc {
    double armin, armax;
    unsigned inverted;
    // `mult`: number of congruent tets this one stands for
    __device__ __forceinline__ void add(const double a[3], const double b[3], const double c[3], const double d[3],
                                        unsigned mult = 1u) {
        double vol, ar;
        tet_quality(a, b, c, d, vol, ar);
        if (vol <= 0.0) inverted += mult;
        armin = fmin(armin, ar);  // Float::min / max ignore NaN like fmin / fmax
        armax = fmax(armax, ar);
    }
    // warp reduction + global min / max / count (all 32 lanes must call)
    __device__ __forceinline__ void flush(unsigned long long* __restrict__ counters) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            armin = fmin(armin, __shfl_xor_sync(0xffffffffu, armin, o));
            armax = fmax(armax, __shfl_xor_sync(0xffffffffu, armax, o));
            inverted += __shfl_xor_sync(0xffffffffu, inverted, o);
        }
        if ((threadIdx.x & 31u) == 0u) {
            // aspect ratios are >= 0, so their bit patterns order like unsigned integers
            if (inverted) atomicAdd(&counters[CN_INVERTED], (unsigned long long)inverted);
            const unsigned long long bmin = (unsigned long long)__double_as_longlong(armin);
            const unsigned long long bmax = (unsigned long long)__double_as_longlong(armax);
            if (bmin < counters[CN_AR_MIN_BITS]) atomicMin(&counters[CN_AR_MIN_BITS], bmin);
            if (bmax > counters[CN_AR_MAX_BITS]) atomicMax(&counters[CN_AR_MAX_BITS], bmax);
        }
    }
};

// ---------------------------------------------------------------------------------------
// a7/a10/a11: emit the output tets of the owned cells, tile by tile, with global vertex ids, and
// accumulate check_for_inverted_tets' statistics (min / max aspect ratio, inverted count).
template <typename IndexT>
__device__ __forceinline__ IndexT node_global_id(const Stuff& S, const TetNodes& tn, int code) {
    uint16_t mask;
    if (code < 4) return (IndexT)vertex_gid(S, tet_X(tn, code), tet_Y(tn, code), tet_Z(tn, code), mask);
    const int i = (code - 4) >> 2, j = (code - 4) & 3;
    const uint32_t Xi = tet_X(tn, i), Yi = tet_Y(tn, i), Zi = tet_Z(tn, i);
    const uint32_t Xj = tet_X(tn, j), Yj = tet_Y(tn, j), Zj = tet_Z(tn, j);
    int d = dir_index((int)Xj - (int)Xi, (int)Yj - (int)Yi, (int)Zj - (int)Zi);
    const bool own_i = d < 9;
    if (!own_i) d -= 9;
    const long long base = own_i ? vertex_gid(S, Xi, Yi, Zi, mask) : vertex_gid(S, Xj, Yj, Zj, mask);
    return (IndexT)cut_vertex_id(base, mask, d);
}

// quality of the five un-warped lattice tets of one cell (kept out of line: it is evaluated for a
// handful of representative cells per tile and would otherwise dictate the register budget)
static __device__ __forceinline__ void lattice_cell_quality(const Lattice& L, uint32_t cx, uint32_t cy, uint32_t cz,
                                                         unsigned mult, QualityAcc& qa) {
    constexpr int T[5][4] = {{0, 1, 5, 2}, {0, 2, 5, 7}, {0, 7, 5, 4}, {2, 6, 5, 7}, {0, 2, 7, 3}};  // Tile::TETS
    const Cell c = make_cell(cx, cy, cz);
    const bool flip = c.flip();
    double pos[8][3];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        uint32_t X, Y, Z;
        corner_xyz(c, k, X, Y, Z);
        pos[k][0] = lat_x(L, X); pos[k][1] = lat_y(L, Y); pos[k][2] = lat_z(L, Z);
    }
#pragma unroll
    for (int t = 0; t < 5; ++t) {
        if (flip) qa.add(pos[T[t][1]], pos[T[t][0]], pos[T[t][2]], pos[T[t][3]], mult);
        else qa.add(pos[T[t][0]], pos[T[t][1]], pos[T[t][2]], pos[T[t][3]], mult);
    }
}

// Interior tiles (every cell emits its five lattice tets, no vertex is warped): ids are pure
// arithmetic, the tets of 256 cells at a time are staged in shared memory and written with
// fully coalesced 16-byte stores.  Quality: an un-warped lattice tet only depends on the
// coordinate steps of its cell along the three axes and on the cell's mirroring, so it is
// evaluated once per distinct (step, parity) combination of the tile, weighted by multiplicity.
template <typename IndexT>
__global__ void __launch_bounds__(TILE_THREADS)
k_tet_emit_full(Stuff S, IndexT* __restrict__ tets, const uint32_t* __restrict__ tile_list,
                const unsigned long long* __restrict__ n_list, unsigned long long* __restrict__ cursor) {
    const Lattice& L = S.L;
    uint32_t tile;
    while (next_tile(tile_list, n_list, cursor, tile)) {  // interior cell tiles with owned cells
    uint32_t X0, Y0, Z0;
    tile_origin(S.g, tile, X0, Y0, Z0);
    const uint32_t zlo = max(Z0, S.c0), zhi = min(Z0 + TS, S.c1);
    const uint64_t tb = S.ttbase[tile];
    constexpr int T[5][4] = {{0, 1, 5, 2}, {0, 2, 5, 7}, {0, 7, 5, 4}, {2, 6, 5, 7}, {0, 2, 7, 3}};  // Tile::TETS
    const uint32_t nxt = min(TS, L.nx - X0), nyt = min(TS, L.ny - Y0);
    // the (up to) 8 point tiles holding the corners of this cell tile
    __shared__ uint8_t s_nvu[8];
    __shared__ unsigned long long s_nbase[8];
    __shared__ TileBox s_nbox[8];
    // (row stride TILE_THREADS + 5 tets: the copy-out below reads (tet g % 5, cell g / 5) and 5 t + c takes every
    // residue mod 8 over eight consecutive g, so its 16-byte shared-memory loads are free of bank conflicts)
    constexpr uint32_t OUT_STRIDE = TILE_THREADS + 5;
    __shared__ __align__(16) IndexT s_out[5][OUT_STRIDE][4];
    if (threadIdx.x < 8) {
        const uint32_t tix = (uint32_t)(tile % S.g.ntx) + (threadIdx.x & 1u);
        const uint32_t tiy = (uint32_t)((tile / S.g.ntx) % S.g.nty) + ((threadIdx.x >> 1) & 1u);
        const uint32_t tiz = (uint32_t)(tile / (S.g.ntx * S.g.nty)) + (threadIdx.x >> 2);
        uint8_t vu = VU_ACTIVE;
        if (tix < S.g.ntx && tiy < S.g.nty && tiz < S.g.ntz) {
            const uint32_t t = (tiz * S.g.nty + tiy) * S.g.ntx + tix;
            vu = S.vuni[t];
            s_nbase[threadIdx.x] = S.tvbase[t];
            s_nbox[threadIdx.x] = tile_box(S, t);
        }
        s_nvu[threadIdx.x] = vu;
    }
    __syncthreads();
    const uint32_t n = nxt * nyt * (zhi - zlo);
    if (sizeof(IndexT) == 4) {
        // 32-bit ids: the ids of the tile's (TS + 1)^3 corner points are tabulated once (every point is
        // shared by up to 8 cells), a cell then costs eight shared-memory loads and a few adds
        constexpr int H = (int)TS + 1;
        __shared__ uint32_t s_id[sizeof(IndexT) == 4 ? H * H * H : 1];  // (unused by the 64-bit instantiation)
        for (int idx = (int)threadIdx.x; idx < H * H * H; idx += TILE_THREADS) {
            const uint32_t lx = idx % H, ly = (idx / H) % H, lz = idx / (H * H);
            const uint32_t X = X0 + lx, Y = Y0 + ly, Z = Z0 + lz;
            uint32_t id = 0u;
            if (lx <= nxt && ly <= nyt && Z >= zlo && Z <= zhi) {
                const uint32_t which = (lx == TS ? 1u : 0u) | (ly == TS ? 2u : 0u) | (lz == TS ? 4u : 0u);
                if (s_nvu[which] == VU_NEG && !(S.lower_table != nullptr && Z == S.c0)) {
                    const TileBox& nb = s_nbox[which];
                    id = (uint32_t)(S.vertex_base + (long long)s_nbase[which] +
                                    (long long)(((Z - nb.zlo) * nb.nyt + (Y - nb.Y0)) * nb.nxt + (X - nb.X0)));
                } else {
                    uint16_t mask;
                    id = (uint32_t)vertex_gid(S, X, Y, Z, mask);
                }
            }
            s_id[idx] = id;
        }
        __syncthreads();
        uint4* s_out4 = reinterpret_cast<uint4*>(&s_out[0][0][0]);
        for (uint32_t base = 0; base < n; base += TILE_THREADS) {
            const uint32_t j = base + threadIdx.x;
            if (j < n) {
                const uint32_t lx = j % nxt, ly = (j / nxt) % nyt, lz = zlo - Z0 + j / (nxt * nyt);
                // corner k of the (possibly mirrored) cell: offset 0 or 1 per axis, mirrored axes count down
                const uint32_t fx = (X0 + lx) & 1u, fy = (Y0 + ly) & 1u, fz = (Z0 + lz) & 1u;
                const int at = (int)((lz + fz) * H + (ly + fy)) * H + (int)(lx + fx);
                const int ax = fx ? -1 : 1, ay = fy ? -H : H, az = fz ? -H * H : H * H;
                // Tile::LATTICE: 0 (0,0,0) 1 (1,0,0) 2 (1,1,0) 3 (0,1,0) 4 (0,0,1) 5 (1,0,1) 6 (1,1,1) 7 (0,1,1)
                const uint32_t i0 = s_id[at], i1 = s_id[at + ax], i2 = s_id[at + ax + ay], i3 = s_id[at + ay];
                const uint32_t i4 = s_id[at + az], i5 = s_id[at + ax + az], i6 = s_id[at + ax + ay + az], i7 = s_id[at + ay + az];
                const bool flip = (fx ^ fy ^ fz) != 0u;
                // Tile::TETS, nodes 0 and 1 swapped when the cell is mirrored an odd number of times
                s_out4[0 * OUT_STRIDE + threadIdx.x] = flip ? make_uint4(i1, i0, i5, i2) : make_uint4(i0, i1, i5, i2);
                s_out4[1 * OUT_STRIDE + threadIdx.x] = flip ? make_uint4(i2, i0, i5, i7) : make_uint4(i0, i2, i5, i7);
                s_out4[2 * OUT_STRIDE + threadIdx.x] = flip ? make_uint4(i7, i0, i5, i4) : make_uint4(i0, i7, i5, i4);
                s_out4[3 * OUT_STRIDE + threadIdx.x] = flip ? make_uint4(i6, i2, i5, i7) : make_uint4(i2, i6, i5, i7);
                s_out4[4 * OUT_STRIDE + threadIdx.x] = flip ? make_uint4(i2, i0, i7, i3) : make_uint4(i0, i2, i7, i3);
            }
            __syncthreads();
            // coalesced copy: output tet g of this round = cell g / 5, lattice tet g % 5
            const uint32_t ncell = min((uint32_t)TILE_THREADS, n - base);
            uint4* dst4 = reinterpret_cast<uint4*>(tets + 4 * (tb + 5ull * base));
            for (uint32_t g = threadIdx.x; g < 5u * ncell; g += TILE_THREADS) dst4[g] = s_out4[(g % 5u) * OUT_STRIDE + g / 5u];
            __syncthreads();
        }
        continue;
    }
    for (uint32_t base = 0; base < n; base += TILE_THREADS) {
        const uint32_t j = base + threadIdx.x;
        if (j < n) {
            const uint32_t cx = X0 + j % nxt, cy = Y0 + (j / nxt) % nyt, cz = zlo + j / (nxt * nyt);
            const Cell c = make_cell(cx, cy, cz);
            const bool flip = c.flip();
            IndexT id[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                uint32_t X, Y, Z;
                corner_xyz(c, k, X, Y, Z);
                const uint32_t which = (X >= X0 + TS ? 1u : 0u) | (Y >= Y0 + TS ? 2u : 0u) | (Z >= Z0 + TS ? 4u : 0u);
                if (s_nvu[which] == VU_NEG && !(S.lower_table != nullptr && Z == S.c0)) {
                    const TileBox& nb = s_nbox[which];
                    id[k] = (IndexT)(S.vertex_base + (long long)s_nbase[which] +
                                     (long long)(((Z - nb.zlo) * nb.nyt + (Y - nb.Y0)) * nb.nxt + (X - nb.X0)));
                } else {
                    uint16_t mask;
                    id[k] = (IndexT)vertex_gid(S, X, Y, Z, mask);
                }
            }
#pragma unroll
            for (int t = 0; t < 5; ++t) {
                // nodes 0 and 1 swapped when the cell is mirrored an odd number of times
                s_out[t][threadIdx.x][0] = flip ? id[T[t][1]] : id[T[t][0]];
                s_out[t][threadIdx.x][1] = flip ? id[T[t][0]] : id[T[t][1]];
                s_out[t][threadIdx.x][2] = id[T[t][2]];
                s_out[t][threadIdx.x][3] = id[T[t][3]];
            }
        }
        __syncthreads();
        const uint32_t ncell = min((uint32_t)TILE_THREADS, n - base);
        IndexT* dst = tets + 4 * (tb + 5ull * base);
        for (uint32_t g = threadIdx.x; g < 20u * ncell; g += TILE_THREADS)
            dst[g] = s_out[(g >> 2) % 5u][(g >> 2) / 5u][g & 3u];
        __syncthreads();
    }
    }  // tiles
}

// check_for_inverted_tets over the interior tiles: an un-warped lattice tet only depends on the
// coordinate steps of its cell along the three axes and on the cell's mirroring, so it is
// evaluated once per distinct (step, parity) combination of the tile, weighted by multiplicity.
static __global__ void __launch_bounds__(64)
k_full_tile_quality(Stuff S, unsigned long long* __restrict__ counters) {
    const Lattice& L = S.L;
    uint32_t X0, Y0, Z0;
    tile_origin(S.g, blockIdx.x, X0, Y0, Z0);
    if (S.cuni[blockIdx.x] != CU_FULL || X0 >= L.nx || Y0 >= L.ny) return;
    const uint32_t zlo = max(Z0, S.c0), zhi = min(Z0 + TS, S.c1);
    if (zhi <= zlo) return;
    __shared__ double s_step[3][TS];
    __shared__ uint8_t s_rep[3][TS];
    __shared__ uint16_t s_mult[3][TS];
    if (threadIdx.x < 3 * TS) {
        const uint32_t a = threadIdx.x / TS, i = threadIdx.x % TS;
        const uint32_t cc = (a == 0 ? X0 : (a == 1 ? Y0 : Z0)) + i;
        const bool valid = a == 0 ? cc < L.nx : (a == 1 ? cc < L.ny : (cc >= zlo && cc < zhi));
        double d = 0.0;
        if (valid) d = a == 0 ? lat_x(L, cc + 1) - lat_x(L, cc) : (a == 1 ? lat_y(L, cc + 1) - lat_y(L, cc) : lat_z(L, cc + 1) - lat_z(L, cc));
        s_step[a][i] = d;
        s_rep[a][i] = valid ? 1 : 0;
    }
    __syncthreads();
    if (threadIdx.x < 3 * TS) {
        const uint32_t a = threadIdx.x / TS, i = threadIdx.x % TS;
        const uint32_t base = a == 0 ? X0 : (a == 1 ? Y0 : Z0);
        bool rep = s_rep[a][i] != 0;
        uint32_t mult = 0;
        if (rep) {
            for (uint32_t q = 0; q < TS; ++q) {
                const bool same = s_rep[a][q] != 0 && s_step[a][q] == s_step[a][i] && (((base + q) ^ (base + i)) & 1u) == 0u;
                if (same) { ++mult; if (q < i) rep = false; }
            }
        }
        s_mult[a][i] = (uint16_t)(rep ? mult : 0u);  // 0 = not a representative
    }
    __syncthreads();
    QualityAcc qa = {1.0, 0.0, 0u};
    for (uint32_t j = threadIdx.x; j < TS * TS * TS; j += 64) {
        const uint32_t ix = j & 15u, iy = (j >> 4) & 15u, iz = j >> 8;
        const uint32_t mult = (uint32_t)s_mult[0][ix] * s_mult[1][iy] * s_mult[2][iz];
        if (mult == 0u) continue;
        lattice_cell_quality(L, X0 + ix, Y0 + iy, Z0 + iz, mult, qa);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        qa.armin = fmin(qa.armin, __shfl_xor_sync(0xffffffffu, qa.armin, o));
        qa.armax = fmax(qa.armax, __shfl_xor_sync(0xffffffffu, qa.armax, o));
        qa.inverted += __shfl_xor_sync(0xffffffffu, qa.inverted, o);
    }
    if ((threadIdx.x & 31u) == 0u) {
        if (qa.inverted) atomicAdd(&counters[CN_INVERTED], (unsigned long long)qa.inverted);
        const unsigned long long bmin = (unsigned long long)__double_as_longlong(qa.armin);
        const unsigned long long bmax = (unsigned long long)__double_as_longlong(qa.armax);
        if (bmin < counters[CN_AR_MIN_BITS]) atomicMin(&counters[CN_AR_MIN_BITS], bmin);
        if (bmax > counters[CN_AR_MAX_BITS]) atomicMax(&counters[CN_AR_MAX_BITS], bmax);
    }
}

// Surface tiles.  The ids / masks / warp flags of the tile's (TS+1)^3 lattice vertices are staged in
// shared memory first (independent loads), so that the per-cell and per-tet work below never
// chases global pointers.  Dynamic shared memory layout: see tet_emit_smem_bytes().
constexpr int TE_H = (int)TS + 1;
constexpr int TE_NP = TE_H * TE_H * TE_H;
// direction index (C_DIR) of the lattice edge between two corners of the staged box, by the difference of their
// box indices (dz * TE_H^2 + dy * TE_H + dx + TE_DIRMID), 255 = not an edge
constexpr int TE_DIRMID = TE_H * TE_H + TE_H + 1, TE_DIRTAB = 2 * TE_DIRMID + 2;
template <typename IndexT>
constexpr size_t tet_emit_smem_bytes() {
    return (size_t)TE_NP * sizeof(IndexT) + (size_t)TE_NP * 2 /*mask*/ + 15 +
           (size_t)TILE_POINTS * 2 * 4 /*ci, off, list, qoff*/ + (size_t)TILE_THREADS * 4 * 4 /*node indices*/ + TE_DIRTAB /*directions*/;
}

// one output tet = one (u32 ids) or two (u64 ids) 16-byte stores
__device__ __forceinline__ void store_tet(uint32_t* tets, uint64_t o, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    reinterpret_cast<uint4*>(tets)[o] = make_uint4(a, b, c, d);
}
__device__ __forceinline__ void store_tet(unsigned long long* tets, uint64_t o, unsigned long long a, unsigned long long b,
                                          unsigned long long c, unsigned long long d) {
    reinterpret_cast<ulonglong2*>(tets)[2 * o] = make_ulonglong2(a, b);
    reinterpret_cast<ulonglong2*>(tets)[2 * o + 1] = make_ulonglong2(c, d);
}

template <typename IndexT, bool QUALITY>
__global__ void __launch_bounds__(TILE_THREADS)
k_tet_emit(Stuff S, IndexT* __restrict__ tets, unsigned long long* __restrict__ qlist, QualityClasses Q,
           unsigned long long* __restrict__ counters, const uint32_t* __restrict__ tile_list,
           const unsigned long long* __restrict__ n_list, unsigned long long* __restrict__ cursor) {
    extern __shared__ __align__(16) unsigned char te_smem[];
    __shared__ uint32_t s_scan[9];
    __shared__ uint32_t s_n;
    const Lattice& L = S.L;
    {
        // (same carve-up as below; filled once per block)
        IndexT* g0 = reinterpret_cast<IndexT*>(te_smem);
        uint16_t* m0 = reinterpret_cast<uint16_t*>(g0 + TE_NP);
        uint16_t* c0 = reinterpret_cast<uint16_t*>((reinterpret_cast<uintptr_t>(m0 + TE_NP) + 15) & ~(uintptr_t)15);
        uint8_t* tab = reinterpret_cast<uint8_t*>(reinterpret_cast<int*>(c0 + 4 * TILE_POINTS) + 4 * TILE_THREADS);
        for (int i = (int)threadIdx.x; i < TE_DIRTAB; i += TILE_THREADS) tab[i] = 255;
        __syncthreads();
        if (threadIdx.x < 18) tab[(C_DIR[threadIdx.x][2] * TE_H + C_DIR[threadIdx.x][1]) * TE_H + C_DIR[threadIdx.x][0] + TE_DIRMID] = (uint8_t)threadIdx.x;
    }
    uint32_t tile;
    while (next_tile(tile_list, n_list, cursor, tile)) {  // active cell tiles
    uint32_t X0, Y0, Z0;
    tile_origin(S.g, tile, X0, Y0, Z0);
    const uint32_t zlo = max(Z0, S.c0), zhi = min(Z0 + TS, S.c1);
    if (zhi <= zlo) continue;  // halo layers only
    IndexT* s_gid = reinterpret_cast<IndexT*>(te_smem);
    uint16_t* s_mask = reinterpret_cast<uint16_t*>(s_gid + TE_NP);
    uint16_t* s_ci = reinterpret_cast<uint16_t*>((reinterpret_cast<uintptr_t>(s_mask + TE_NP) + 15) & ~(uintptr_t)15);
    uint16_t* s_off = s_ci + TILE_POINTS;   // first output tet of the cell, relative to the tile
    uint16_t* s_list = s_off + TILE_POINTS; // cells with trimmed / dropped tets
    uint16_t* s_qoff = s_list + TILE_POINTS; // first quality-list slot of the cell, relative to the tile
    int* s_li = reinterpret_cast<int*>(s_qoff + TILE_POINTS) + 4 * threadIdx.x;   // this thread's four node indices (run-time indexed)
    uint8_t* s_dirtab = reinterpret_cast<uint8_t*>(reinterpret_cast<int*>(s_qoff + TILE_POINTS) + 4 * TILE_THREADS);
    __shared__ unsigned long long s_qbase;
    constexpr int T[5][4] = {{0, 1, 5, 2}, {0, 2, 5, 7}, {0, 7, 5, 4}, {2, 6, 5, 7}, {0, 2, 7, 3}};  // Tile::TETS
    const uint64_t tb = S.ttbase[tile];
    if (threadIdx.x == 0) s_n = 0u;
    // (issued first, used in step 2) this thread's x-row of cell info words: 32 contiguous bytes
    uint32_t crow[TS / 2];
    {
        const uint32_t ly = threadIdx.x & 15u, lz = threadIdx.x >> 4;
        const uint32_t cy = Y0 + ly, cz = Z0 + lz;
        const bool row_owned = cy < L.ny && cz >= zlo && cz < zhi;
        const uint16_t* __restrict__ row = S.cinfo + (row_owned ? cidx(S, X0, cy, cz) : 0);
        const uint32_t nx_row = row_owned ? min(TS, L.nx - X0) : 0u;
        if (nx_row == TS && (reinterpret_cast<uintptr_t>(row) & 15u) == 0u) {
            const uint4 a = reinterpret_cast<const uint4*>(row)[0], b = reinterpret_cast<const uint4*>(row)[1];
            crow[0] = a.x; crow[1] = a.y; crow[2] = a.z; crow[3] = a.w; crow[4] = b.x; crow[5] = b.y; crow[6] = b.z; crow[7] = b.w;
        } else {
#pragma unroll
            for (uint32_t i = 0; i < TS / 2; ++i)
                crow[i] = (2u * i < nx_row ? (uint32_t)row[2u * i] : 0u) | ((2u * i + 1u < nx_row ? (uint32_t)row[2u * i + 1u] : 0u) << 16);
        }
    }
    // 1. vertex ids / masks of the tile's corners.  The (TS+1)^3 points lie in at most 8 tiles:
    //    their class, id base and owned box are staged once, so that a point costs no dependent
    //    global load (interior tiles: arithmetic; surface tiles: its mask and local index).
    __shared__ TileBox s_tb[8];
    __shared__ long long s_tbase[8];
    __shared__ uint8_t s_tvu[8];
    if (threadIdx.x < 8) {
        const uint32_t q = threadIdx.x;
        const uint32_t X = X0 + ((q & 1u) ? TS : 0u), Y = Y0 + ((q & 2u) ? TS : 0u), Z = Z0 + ((q & 4u) ? TS : 0u);
        uint8_t vu = VU_POS;
        if (X < L.NX && Y < L.NY && Z <= L.pz1) {
            const uint32_t t = tile_of(S.g, X, Y, Z);
            vu = S.vuni[t];
            s_tb[q] = tile_box(S, t);
            s_tbase[q] = S.vertex_base + (long long)S.tvbase[t];
        }
        s_tvu[q] = vu;
    }
    __syncthreads();
    // One x-row of 17 points per thread and round: the masks / local indices of its first 16 points are two
    // aligned 32-byte runs (the row pitch is a multiple of 16 points), fetched with 16-byte loads.
    for (int r = (int)threadIdx.x; r < TE_H * TE_H; r += TILE_THREADS) {
        const uint32_t ly = (uint32_t)r % TE_H, lz = (uint32_t)r / TE_H;
        const uint32_t Y = Y0 + ly, Z = Z0 + lz;
        const uint32_t q0 = (ly == TS ? 2u : 0u) | (lz == TS ? 4u : 0u);
        const bool row_in = Y < L.NY && Z <= L.pz1;
        const bool from_table = row_in && S.lower_table != nullptr && Z == S.c0;
        const uint8_t vu0 = row_in ? s_tvu[q0] : VU_POS, vu1 = (row_in && X0 + TS < L.NX) ? s_tvu[q0 | 1u] : VU_POS;
        uint4 m0 = make_uint4(0u, 0u, 0u, 0u), m1 = m0, l0 = m0, l1 = m0;
        uint32_t mlast = 0u, llast = 0u;
        if (!from_table) {
            const size_t pi = pidx(L, X0, Y, Z);
            if (vu0 == VU_ACTIVE) {
                m0 = *reinterpret_cast<const uint4*>(S.vmask + pi); m1 = *reinterpret_cast<const uint4*>(S.vmask + pi + 8);
                l0 = *reinterpret_cast<const uint4*>(S.vloc + pi);  l1 = *reinterpret_cast<const uint4*>(S.vloc + pi + 8);
            }
            if (vu1 == VU_ACTIVE) { mlast = S.vmask[pi + TS]; llast = S.vloc[pi + TS]; }
        }
        const uint32_t mw[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
        const uint32_t lw[8] = {l0.x, l0.y, l0.z, l0.w, l1.x, l1.y, l1.z, l1.w};
        const int idx0 = r * TE_H;
#pragma unroll
        for (uint32_t i = 0; i <= TS; ++i) {
            const uint32_t X = X0 + i;
            const uint32_t q = i == TS ? (q0 | 1u) : q0;
            const uint8_t vu = i == TS ? vu1 : vu0;
            IndexT gid = 0;
            uint16_t mask = 0;
            if (X < L.NX) {
                if (from_table) {
                    const unsigned long long e = vu != VU_POS ? S.lower_table[(size_t)Y * L.NX + X] : 0ull;
                    mask = (uint16_t)(e & 0xffffull);
                    gid = (IndexT)(e >> 16);
                } else if (vu == VU_NEG) {
                    const TileBox& tbx = s_tb[q];
                    mask = VM_USED;
                    gid = (IndexT)(s_tbase[q] + (long long)(((Z - tbx.zlo) * tbx.nyt + (Y - tbx.Y0)) * tbx.nxt + (X - tbx.X0)));
                } else if (vu == VU_ACTIVE) {
                    const uint32_t mv = i == TS ? mlast : ((mw[i >> 1] >> (16u * (i & 1u))) & 0xffffu);
                    const uint32_t lv = i == TS ? llast : ((lw[i >> 1] >> (16u * (i & 1u))) & 0xffffu);
                    mask = (uint16_t)mv;
                    gid = (IndexT)(s_tbase[q] + (long long)lv);
                }
            }
            s_gid[idx0 + (int)i] = gid; s_mask[idx0 + (int)i] = mask;
        }
    }
    // 2. per-cell info + offsets: every thread takes one x-row of 16 cells (consecutive in the
    //    output order), one block scan of the row totals gives the offsets of the output tets
    //    (low half) and of the tets listed for the per-tet quality pass (high half; <= 61440 each)
    {
        uint32_t acc = 0, qbits = 0;
        uint32_t pre[TS];
#pragma unroll
        for (uint32_t i = 0; i < TS; ++i) {
            const uint16_t ci = (uint16_t)(crow[i >> 1] >> (16u * (i & 1u)));
            const uint32_t count = ci_count(ci);
            const bool listed = QUALITY && count && (!(ci & CI_FULL) || (ci & CI_FULL_ZERO_CORNER));
            // plain cells (untouched, all corners strictly negative): their (step, parity) class occurs
            if (QUALITY && Q.present != nullptr && (ci & CI_FULL) && !(ci & CI_FULL_ZERO_CORNER)) qbits |= 1u << Q.kx[X0 + i];
            s_ci[threadIdx.x * TS + i] = ci;
            pre[i] = acc;
            acc += count | (listed ? count << 16 : 0u);
            if (count && !(ci & CI_FULL)) s_list[atomicAdd(&s_n, 1u)] = (uint16_t)(threadIdx.x * TS + i);
        }
        uint32_t tot;
        const uint32_t ex = block_exscan(acc, &tot, s_scan);
#pragma unroll
        for (uint32_t i = 0; i < TS; ++i) {
            const uint32_t v = ex + pre[i];
            s_off[threadIdx.x * TS + i] = (uint16_t)(v & 0xffffu);
            if (QUALITY) s_qoff[threadIdx.x * TS + i] = (uint16_t)(v >> 16);
        }
        if (QUALITY && threadIdx.x == 0) s_qbase = (tot >> 16) ? atomicAdd(&counters[CN_QLIST_FILL], (unsigned long long)(tot >> 16)) : 0ull;
        if (QUALITY && qbits) {
            uint32_t* w = Q.present + (uint32_t)Q.kz[Z0 + (threadIdx.x >> 4)] * QC_K + Q.ky[Y0 + (threadIdx.x & 15u)];
            if ((*w & qbits) != qbits) atomicOr(w, qbits);
        }
    }
    __syncthreads();
    // 3a. cells whose five lattice tets are emitted untouched
    for (uint32_t j = threadIdx.x; j < TILE_POINTS; j += TILE_THREADS) {
        const uint16_t ci = s_ci[j];
        if (!(ci & CI_FULL)) continue;
        const uint32_t lx = j & 15u, ly = (j >> 4) & 15u, lz = j >> 8;
        // corner k of Tile::LATTICE sits at (lx ^ fx, ly ^ fy, lz ^ fz) of the cell: six offsets per cell
        const uint32_t fx = (X0 + lx) & 1u, fy = (Y0 + ly) & 1u, fz = (Z0 + lz) & 1u;
        const bool flip = ((fx ^ fy ^ fz) & 1u) != 0u;
        const int base = (int)((lz * TE_H + ly) * TE_H + lx);
        const int ox[2] = {(int)fx, 1 - (int)fx}, oy[2] = {(int)fy * TE_H, (1 - (int)fy) * TE_H},
                  oz[2] = {(int)fz * TE_H * TE_H, (1 - (int)fz) * TE_H * TE_H};
        constexpr int LB[8][3] = {{0, 0, 0}, {1, 0, 0}, {1, 1, 0}, {0, 1, 0}, {0, 0, 1}, {1, 0, 1}, {1, 1, 1}, {0, 1, 1}};  // Tile::LATTICE
        IndexT id[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) id[k] = s_gid[base + ox[LB[k][0]] + oy[LB[k][1]] + oz[LB[k][2]]];
        const uint64_t o = tb + s_off[j];
#pragma unroll
        for (int t = 0; t < 5; ++t)
            store_tet(tets, o + (uint64_t)t, flip ? id[T[t][1]] : id[T[t][0]], flip ? id[T[t][0]] : id[T[t][1]],
                      id[T[t][2]], id[T[t][3]]);
        // zero-cornered untouched cells are not congruent to plain lattice tets: listed for the
        // per-tet quality pass (the plain ones are handled per class by k_surface_tile_quality)
        if (QUALITY && (ci & CI_FULL_ZERO_CORNER)) {
            const unsigned long long at = s_qbase + s_qoff[j];
#pragma unroll
            for (int t = 0; t < 5; ++t) qlist[at + t] = o + (uint64_t)t;
        }
    }
    // 3b. the lattice tets of the surface cells, one per thread
    const uint32_t n = s_n;
    for (uint32_t w = threadIdx.x; w < 5u * n; w += TILE_THREADS) {
        const uint32_t j = s_list[w / 5u];
        const int t = (int)(w % 5u);
        const uint16_t ci = s_ci[j];
        const uint32_t nt = ci_tet_n(ci, t);
        if (nt == 0u) continue;
        uint32_t before = 0;
        for (int u = 0; u < t; ++u) before += ci_tet_n(ci, u);
        uint64_t o = tb + s_off[j] + before;
        const Cell c = make_cell(X0 + (j & 15u), Y0 + ((j >> 4) & 15u), Z0 + (j >> 8));
        TetNodes tn;
        TetOut to;
        if ((ci >> (CI_UNTOUCHED_SHIFT + t)) & 1u) {
            // untouched: the lattice tet itself, no need to look at the field
#pragma unroll
            for (int m = 0; m < 4; ++m) corner_xyz(c, tet_corner(c, t, m), tn.X[m], tn.Y[m], tn.Z[m]);
            to.n = 1; to.kase = -1; to.ext_removed = false;
            to.nodes = 0ull | (1ull << 5) | (2ull << 10) | (3ull << 15);
        } else if (!S.emit_reclassify) {
            // trimmed: the pieces follow from the shape byte k_classify_cells recorded (no look at the field)
#pragma unroll
            for (int m = 0; m < 4; ++m) corner_xyz(c, tet_corner(c, t, m), tn.X[m], tn.Y[m], tn.Z[m]);
            const uint32_t shape = (uint32_t)(S.cshape[cidx(S, c.cx, c.cy, c.cz)] >> (8 * t)) & 0xffu;
            const unsigned long long e = S.shapes->nodes[shape];
            to.n = (int)(e >> 60); to.kase = 0; to.ext_removed = false;
            to.nodes = e & 0x0fffffffffffffffull;
        } else {
            // the same from the field (nt != 0, so kept and not dropped by the exterior test): vertices on both sides of zero
            load_tet_mixed(L, c, t, tn);
            classify_tet<false, true>(L, c, t, tn, nullptr, 0, to);
        }
        // the four nodes' indices in the staged box live in shared memory: the node codes index them at run time
#pragma unroll
        for (int m = 0; m < 4; ++m) s_li[m] = (int)((tn.Z[m] - Z0) * TE_H + (tn.Y[m] - Y0)) * TE_H + (int)(tn.X[m] - X0);
        const unsigned long long at = QUALITY ? s_qbase + s_qoff[j] + before : 0ull;
        for (int q = 0; q < to.n; ++q) {
            IndexT id[4];
#pragma unroll
            for (int m = 0; m < 4; ++m) {
                const int code = to.node(q, m);
                if (code < 4) {
                    id[m] = s_gid[s_li[code]];
                } else {
                    // cut vertex on the edge i -> jn: owned by the end from which the direction is canonical (d < 9)
                    const int li_i = s_li[(code - 4) >> 2], li_j = s_li[(code - 4) & 3];
                    int d = (int)s_dirtab[li_j - li_i + TE_DIRMID];
                    const bool own_i = d < 9;
                    if (!own_i) d -= 9;
                    const int lo = own_i ? li_i : li_j;
                    id[m] = (IndexT)cut_vertex_id((long long)s_gid[lo], s_mask[lo], d);
                }
            }
            store_tet(tets, o, id[0], id[1], id[2], id[3]);
            if (QUALITY) qlist[at + (unsigned)q] = o;
            ++o;
        }
    }
    }  // tiles
}

// check_for_inverted_tets for the plain lattice tets of the SURFACE tiles (untouched cells with
// all corners strictly negative): counted per (step, parity) class of the tile, one evaluation
// per class — see k_full_tile_quality.
static __global__ void __launch_bounds__(TILE_THREADS)
k_surface_tile_quality(Stuff S, unsigned long long* __restrict__ counters) {
    const Lattice& L = S.L;
    uint32_t X0, Y0, Z0;
    tile_origin(S.g, blockIdx.x, X0, Y0, Z0);
    if (S.cuni[blockIdx.x] != CU_ACTIVE || X0 >= L.nx || Y0 >= L.ny) return;
    const uint32_t zlo = max(Z0, S.c0), zhi = min(Z0 + TS, S.c1);
    if (zhi <= zlo) return;
    __shared__ double s_step[3][TS];
    __shared__ uint8_t s_repidx[3][TS];
    __shared__ __align__(4) uint16_t s_cnt[TILE_POINTS];
    __shared__ uint32_t s_need_counts;
    for (uint32_t j = threadIdx.x; j < TILE_POINTS; j += TILE_THREADS) s_cnt[j] = 0;
    if (threadIdx.x == 0) s_need_counts = 0u;
    if (threadIdx.x < 3 * TS) {
        const uint32_t a = threadIdx.x / TS, i = threadIdx.x % TS;
        const uint32_t cc = (a == 0 ? X0 : (a == 1 ? Y0 : Z0)) + i;
        s_step[a][i] = a == 0 ? lat_x(L, cc + 1) - lat_x(L, cc) : (a == 1 ? lat_y(L, cc + 1) - lat_y(L, cc) : lat_z(L, cc + 1) - lat_z(L, cc));
    }
    __syncthreads();
    if (threadIdx.x < 3 * TS) {
        const uint32_t a = threadIdx.x / TS, i = threadIdx.x % TS;
        uint32_t rep = i;
        for (uint32_t q = 0; q < i; ++q)
            if (s_step[a][q] == s_step[a][i] && ((q ^ i) & 1u) == 0u) { rep = q; break; }
        s_repidx[a][i] = (uint8_t)rep;
    }
    __syncthreads();
    // The aspect-ratio range only needs to know WHICH classes occur (plain stores, no atomics);
    // multiplicities matter for the inverted count alone, and un-warped lattice tets are not
    // inverted: the counting pass below runs only if some class says otherwise.
    auto class_of = [&](uint32_t j, uint32_t& cls) -> bool {
        const uint32_t lx = j & 15u, ly = (j >> 4) & 15u, lz = j >> 8;
        const uint32_t cx = X0 + lx, cy = Y0 + ly, cz = Z0 + lz;
        if (cx >= L.nx || cy >= L.ny || cz < zlo || cz >= zhi) return false;
        const uint16_t ci = S.cinfo[cidx(S, cx, cy, cz)];
        if (!((ci & CI_FULL) && !(ci & CI_FULL_ZERO_CORNER))) return false;
        cls = ((uint32_t)s_repidx[2][lz] * TS + s_repidx[1][ly]) * TS + s_repidx[0][lx];
        return true;
    };
    for (uint32_t j = threadIdx.x; j < TILE_POINTS; j += TILE_THREADS) {
        uint32_t cls;
        if (class_of(j, cls)) s_cnt[cls] = 1;
    }
    __syncthreads();
    QualityAcc qa = {1.0, 0.0, 0u};
    for (uint32_t j = threadIdx.x; j < TILE_POINTS; j += TILE_THREADS)
        if (s_cnt[j]) lattice_cell_quality(L, X0 + (j & 15u), Y0 + ((j >> 4) & 15u), Z0 + (j >> 8), 1u, qa);
    if (qa.inverted) s_need_counts = 1u;
    qa.inverted = 0u;
    __syncthreads();
    if (s_need_counts) {
        // exact multiplicities (16-bit counters updated through their aligned 32-bit word)
        for (uint32_t j = threadIdx.x; j < TILE_POINTS; j += TILE_THREADS) s_cnt[j] = 0;
        __syncthreads();
        for (uint32_t j = threadIdx.x; j < TILE_POINTS; j += TILE_THREADS) {
            uint32_t cls;
            if (class_of(j, cls)) atomicAdd(reinterpret_cast<unsigned int*>(s_cnt + (cls & ~1u)), 1u << (16u * (cls & 1u)));
        }
        __syncthreads();
        for (uint32_t j = threadIdx.x; j < TILE_POINTS; j += TILE_THREADS) {
            const uint32_t cnt = s_cnt[j];
            if (cnt) {
                QualityAcc one = {1.0, 0.0, 0u};
                lattice_cell_quality(L, X0 + (j & 15u), Y0 + ((j >> 4) & 15u), Z0 + (j >> 8), cnt, one);
                qa.inverted += one.inverted;
            }
        }
    }
    qa.flush(counters);
}

// ---------------------------------------------------------------------------------------
// check_for_inverted_tets for ALL plain lattice tets at once.  An un-warped lattice tet is built from
// coordinate differences that are 0 or +-(the step of its cell along an axis), so its volume and
// aspect ratio are functions of (step_x, step_y, step_z, mirroring) only.  Over the whole lattice
// the steps take a handful of distinct values (they differ in the last bits), so every axis gets a
// small class id per cell (host: step value id * 2 + parity, at most QC_K per axis), the kernels
// below record which (kx, ky, kz) combinations occur among the plain cells, and each combination is
// evaluated exactly once.  The aspect-ratio range needs nothing else; the inverted COUNT would
// need multiplicities, so a combination that turns out inverted (it never does for a sane
// lattice) only raises a flag and the host re-runs the per-tile counting kernels above.

// interior cell tiles: every cell is plain, all combinations of the tile's classes occur
static __global__ void __launch_bounds__(256)
k_quality_mark_full(Stuff S, QualityClasses Q, const uint32_t* __restrict__ tile_list,
                    const unsigned long long* __restrict__ n_list) {
    const Lattice& L = S.L;
    // one warp per listed tile (interior cell tiles with owned cells)
    const unsigned long long item = (unsigned long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (item >= *n_list) return;
    const uint32_t tile = tile_list[item];
    uint32_t X0, Y0, Z0;
    tile_origin(S.g, tile, X0, Y0, Z0);
    const uint32_t zlo = max(Z0, S.c0), zhi = min(Z0 + TS, S.c1);
    const uint32_t lane = threadIdx.x & 31u, i = lane & 15u;
    uint32_t bx = 0u, by = 0u, bz = 0u;
    if (lane < 16u) {
        if (X0 + i < L.nx) bx = 1u << Q.kx[X0 + i];
        if (Y0 + i < L.ny) by = 1u << Q.ky[Y0 + i];
    } else if (Z0 + i >= zlo && Z0 + i < zhi) {
        bz = 1u << Q.kz[Z0 + i];
    }
    const uint32_t mx = __reduce_or_sync(0xffffffffu, bx), my = __reduce_or_sync(0xffffffffu, by), mz = __reduce_or_sync(0xffffffffu, bz);
    const uint32_t ny_ = (uint32_t)__popc(my), npairs = ny_ * (uint32_t)__popc(mz);
    for (uint32_t p = lane; p < npairs; p += 32u) {
        const uint32_t ky = __fns(my, 0u, (int)(p % ny_) + 1), kz = __fns(mz, 0u, (int)(p / ny_) + 1);
        uint32_t* w = Q.present + kz * QC_K + ky;
        if ((*w & mx) != mx) atomicOr(w, mx);
    }
}

// one thread per (kx, ky, kz): the five lattice tets of a cell with those steps and mirroring
static __global__ void __launch_bounds__(256)
k_quality_classes(QualityClasses Q, unsigned long long* __restrict__ counters) {
    constexpr int T[5][4] = {{0, 1, 5, 2}, {0, 2, 5, 7}, {0, 7, 5, 4}, {2, 6, 5, 7}, {0, 2, 7, 3}};  // Tile::TETS
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t kx = t % QC_K, ky = (t / QC_K) % QC_K, kz = t / (QC_K * QC_K);
    QualityAcc qa = {1.0, 0.0, 0u};
    if (kz < QC_K && ((Q.present[kz * QC_K + ky] >> kx) & 1u)) {
        const double d[3] = {Q.steps[kx >> 1], Q.steps[QC_K / 2 + (ky >> 1)], Q.steps[QC_K + (kz >> 1)]};
        const bool f[3] = {(kx & 1u) != 0u, (ky & 1u) != 0u, (kz & 1u) != 0u};
        double pos[8][3];
#pragma unroll
        for (int k = 0; k < 8; ++k)
#pragma unroll
            for (int a = 0; a < 3; ++a) pos[k][a] = ((f[a] ? 1 - C_LAT[k][a] : C_LAT[k][a]) != 0) ? d[a] : 0.0;
        const bool flip = f[0] ^ f[1] ^ f[2];
#pragma unroll
        for (int q = 0; q < 5; ++q) {
            if (flip) qa.add(pos[T[q][1]], pos[T[q][0]], pos[T[q][2]], pos[T[q][3]]);
            else qa.add(pos[T[q][0]], pos[T[q][1]], pos[T[q][2]], pos[T[q][3]]);
        }
    }
    // an inverted class needs multiplicities: flag it, count nothing here
    if (qa.inverted) atomicAdd(&counters[CN_CLASS_INVERTED], 1ull);
    qa.inverted = 0u;
    qa.flush(counters);
}

// check_for_inverted_tets for the listed (trimmed or zero-cornered) tets: one tet per thread,
// positions gathered from the emitted coordinates.  Vertices owned by the rank below (ids under
// vertex_base) are looked up in the slice of ITS coordinates that came with the plane table
// (lower_coords: global ids [lower_first, lower_first + lower_count)); without that slice such
// tets are left out of the statistics.
template <typename IndexT>
__global__ void __launch_bounds__(256)
k_tet_quality_list(const double* __restrict__ coords, const IndexT* __restrict__ tets,
                   const unsigned long long* __restrict__ qlist, uint64_t n, long long vertex_base,
                   const double* __restrict__ lower_coords, long long lower_first, long long lower_count,
                   unsigned long long* __restrict__ counters) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    QualityAcc qa = {1.0, 0.0, 0u};
    if (i < n) {
        const unsigned long long e = qlist[i];
        double p[4][3];
        bool have = true;
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            const long long gid = (long long)tets[4 * e + m];
            const double* src = coords + 3 * (gid - vertex_base);
            if (gid < vertex_base) {
                const long long lid = gid - lower_first;
                if (lower_coords == nullptr || lid < 0 || lid >= lower_count) { have = false; break; }
                src = lower_coords + 3 * lid;
            }
#pragma unroll
            for (int k = 0; k < 3; ++k) p[m][k] = src[k];
        }
        if (have) qa.add(p[0], p[1], p[2], p[3]);
    }
    qa.flush(counters);
}

// id table of the plane z = Z for the rank above: (global id of the vertex's first output << 16) | mask
static __global__ void k_export_plane(Stuff S, uint32_t Z, unsigned long long* __restrict__ table) {
    const uint64_t n = (uint64_t)S.L.NX * S.L.NY;
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t Y = (uint32_t)(i / S.L.NX), X = (uint32_t)(i % S.L.NX);
    uint16_t mask = 0;
    long long base = 0;
    if (S.vuni[tile_of(S.g, X, Y, Z)] != VU_POS) base = vertex_gid(S, X, Y, Z, mask);
    table[i] = ((unsigned long long)base << 16) | (unsigned long long)mask;
}

// DADD + DMUL throughput probe (no FMA): 8 independent chains per thread
static __global__ void k_probe_fp64(double* out, int iters) {
    double a0 = 1.0 + threadIdx.x * 1e-9, a1 = a0 + 1e-3, a2 = a0 + 2e-3, a3 = a0 + 3e-3;
    double a4 = a0 + 4e-3, a5 = a0 + 5e-3, a6 = a0 + 6e-3, a7 = a0 + 7e-3;
    const double m = 0.999999999, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
        a0 = a0 * m; a1 = a1 * m; a2 = a2 * m; a3 = a3 * m; a4 = a4 * m; a5 = a5 * m; a6 = a6 * m; a7 = a7 * m;
        a0 = a0 + c; a1 = a1 + c; a2 = a2 + c; a3 = a3 + c; a4 = a4 + c; a5 = a5 + c; a6 = a6 + c; a7 = a7 + c;
    }
    const double s = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
    if (s == 12345.678) out[0] = s;
}

}  // namespace mc
