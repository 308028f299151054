// Mesh::boundary() (mesh/src/lib.rs:341-357) and the side-set loop (src/main.rs:87-106).
//
// The reference toggles every face of every tet in and out of a HashMap keyed by the sorted node
// triple: a face is a boundary side iff it occurs an odd number of times, and the side that is
// reported is its last occurrence.  No map and no sort is needed to evaluate that rule here,
// because the mesh is structured: every output tet is a piece of ONE lattice tet (whole, or one
// of the <= 3 tets trim_spikes cuts it into), so a face of an output tet either
//   * lies inside its lattice tet: it can only occur again among the other pieces of the same
//     lattice tet (the node ids of a face name the lattice vertices / lattice edges it hangs on), or
//   * lies in one of the four faces of its lattice tet: it can only occur again among the pieces of
//     the same lattice tet and of the ONE lattice tet on the other side of that lattice face.
// What trim_spikes made of a lattice tet is one of 146 SHAPES (mc_shapes.hpp; k_classify_cells records
// one byte per lattice tet), and which lattice tet lies across a lattice face, how its nodes line up
// with ours and which of the two comes later in the fixed order only depends on the mirroring class of
// the cell, the tet number and the face.  So the hash-map rule (including the faces that occur three
// times in the prism split of trim case 3) is evaluated per face with a handful of table look-ups:
//   own[shape]                 occurrences of every piece-face among the other pieces of its lattice tet
//   pfk[shape][k]              the piece-faces that lie in lattice face k
//   geo[class][tet][k]         the lattice tet across face k, rank permutations of the shared nodes
//   pfsig / fsig[shape][..]    face signatures (which of the three shared lattice points / edges a
//                              sub-face hangs on) in the rank order both sides agree on
// Only cells next to the surface do any work: a cell whose five lattice tets are emitted untouched
// ("full") next to full cells is skipped outright.
//
// Everything is slab-aware: the cell layers next to the slab are classified too (halo), so the tet
// across a slab boundary is known locally and no exchange is needed.  Sides come out sorted by
// (elem, side) because tiles, cells, lattice tets and pieces are visited in emission order.
#include <algorithm>
#include <cstdlib>
#include <limits>

#include "mc_internal.hpp"
#include "mc_kernels.cuh"

namespace mc {

Stuff make_stuff(const mc_mesh* m);  // mc_mesher.cu

// A cell as the boundary kernels see it: bits 0..15 its info word (CI_*), bits 16 + 8 t .. its five shape bytes.
constexpr unsigned long long CELL_ALL_FULL = (unsigned long long)CI_ALL_FULL | (0x0101010101ull << 16);
__device__ __forceinline__ uint32_t cell_shape(unsigned long long w, int t) { return (uint32_t)(w >> (16 + 8 * t)) & 0xffu; }

// The words of a tile's cells with a one-cell halo are staged in shared memory (BH^3 entries, 0 = nothing
// there / outside the lattice); `at` = index of the cell in that box.
constexpr int BH = (int)TS + 2;
constexpr int BH3 = BH * BH * BH;
// Stage the cell words of tile (X0, Y0, Z0) + halo.  Two rounds of independent loads (every thread issues a
// batch before it uses any): the info words of all cells, then the shape bytes of the surface cells among them.
// s_cu[27]: class (CU_*) of the 27 cell tiles around, CU_EMPTY where there is none.
__device__ __forceinline__ void stage_cell_words(const Stuff& S, uint32_t tile, uint32_t X0, uint32_t Y0, uint32_t Z0,
                                                 unsigned long long* __restrict__ s_cell, uint8_t* __restrict__ s_cu) {
    const Lattice& L = S.L;
    if (threadIdx.x < 27) {
        const int dx = (int)(threadIdx.x % 3u) - 1, dy = (int)((threadIdx.x / 3u) % 3u) - 1, dz = (int)(threadIdx.x / 9u) - 1;
        const int tix = (int)(tile % S.g.ntx) + dx, tiy = (int)((tile / S.g.ntx) % S.g.nty) + dy,
                  tiz = (int)(tile / ((uint64_t)S.g.ntx * S.g.nty)) + dz;
        uint8_t cu = CU_EMPTY;
        if (tix >= 0 && tiy >= 0 && tiz >= 0 && tix < (int)S.g.ntx && tiy < (int)S.g.nty && tiz < (int)S.g.ntz)
            cu = S.cuni[((size_t)tiz * S.g.nty + tiy) * S.g.ntx + tix];
        s_cu[threadIdx.x] = cu;
    }
    __syncthreads();
    constexpr int BATCH = 4;
    // round 1: info words
    for (int base = (int)threadIdx.x; base < BH3; base += BATCH * TILE_THREADS) {
        uint16_t ci[BATCH];
        uint8_t cu[BATCH];
#pragma unroll
        for (int b = 0; b < BATCH; ++b) {
            const int idx = base + b * TILE_THREADS;
            ci[b] = 0;
            cu[b] = CU_EMPTY;
            if (idx < BH3) {
                const int lx = idx % BH - 1, ly = (idx / BH) % BH - 1, lz = idx / (BH * BH) - 1;
                const int64_t cx = (int64_t)X0 + lx, cy = (int64_t)Y0 + ly, cz = (int64_t)Z0 + lz;
                // (cells outside the classified layers are never reached for neighbours of owned cells)
                if (cx >= 0 && cy >= 0 && cx < (int64_t)L.nx && cy < (int64_t)L.ny && cz >= (int64_t)S.cl0 && cz < (int64_t)S.cl1) {
                    cu[b] = s_cu[((lz < 0 ? 0 : (lz >= (int)TS ? 2 : 1)) * 3 + (ly < 0 ? 0 : (ly >= (int)TS ? 2 : 1))) * 3 +
                                 (lx < 0 ? 0 : (lx >= (int)TS ? 2 : 1))];
                    if (cu[b] == CU_ACTIVE) ci[b] = S.cinfo[cidx(S, (uint32_t)cx, (uint32_t)cy, (uint32_t)cz)];
                }
            }
        }
#pragma unroll
        for (int b = 0; b < BATCH; ++b) {
            const int idx = base + b * TILE_THREADS;
            if (idx >= BH3) continue;
            unsigned long long w = 0ull;
            if (cu[b] == CU_FULL || (ci[b] & CI_FULL)) w = CELL_ALL_FULL;
            else if (ci_count(ci[b]) != 0u) w = (unsigned long long)ci[b];
            s_cell[idx] = w;
        }
    }
    __syncthreads();
    // round 2: shape bytes of the surface cells (a word without shape bytes so far)
    for (int base = (int)threadIdx.x; base < BH3; base += BATCH * TILE_THREADS) {
        unsigned long long sh[BATCH];
        bool need[BATCH];
#pragma unroll
        for (int b = 0; b < BATCH; ++b) {
            const int idx = base + b * TILE_THREADS;
            sh[b] = 0ull;
            need[b] = false;
            if (idx < BH3) {
                const unsigned long long w = s_cell[idx];
                need[b] = w != 0ull && (w >> 16) == 0ull;
                if (need[b]) {
                    const int lx = idx % BH - 1, ly = (idx / BH) % BH - 1, lz = idx / (BH * BH) - 1;
                    sh[b] = S.cshape[cidx(S, X0 + (uint32_t)lx, Y0 + (uint32_t)ly, Z0 + (uint32_t)lz)];
                }
            }
        }
#pragma unroll
        for (int b = 0; b < BATCH; ++b)
            if (need[b]) s_cell[base + b * TILE_THREADS] |= sh[b] << 16;
    }
}

// Boundary sides of the pieces of lattice tet t (shape A != 0) of the cell at `at` (mirroring class p): bit 4 q + s.
__device__ __forceinline__ uint32_t tet_boundary_mask(const ShapeTables* __restrict__ T, const unsigned long long* __restrict__ s_cell, int at,
                                                      uint32_t p, int t, uint32_t A) {
    const uint32_t own = __ldg(&T->own[A]);
    uint32_t odd = own & 0xfffu, notlast = (own >> 12) & 0xfffu;
    const uint32_t valid = (1u << (own >> 24)) - 1u;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        uint32_t pf = __ldg(&T->pfk[A][k]);
        if (pf == 0u) continue;
        const uint32_t g = __ldg(&T->geo[p][t][k]);
        const int dx = (int)((g >> GEO_DX_SHIFT) & 3u) - 1, dy = (int)((g >> GEO_DY_SHIFT) & 3u) - 1, dz = (int)((g >> GEO_DZ_SHIFT) & 3u) - 1;
        const uint32_t B = cell_shape(s_cell[at + (dz * BH + dy) * BH + dx], (int)((g >> GEO_T2_SHIFT) & 7u));
        if (B == SHAPE_NONE) continue;
        const uint32_t sigs = __ldg(&T->fsig[B][(g >> GEO_K2_SHIFT) & 3u][(g >> GEO_PNB_SHIFT) & 7u]);
        const uint8_t* __restrict__ mysig = T->pfsig[A][(g >> GEO_PME_SHIFT) & 7u];
        while (pf) {
            const int i = __ffs((int)pf) - 1;
            pf &= pf - 1u;
            const uint32_t mine = __ldg(&mysig[i]);
            // (a signature is never 0, unused bytes of `sigs` are)
            const uint32_t cnt = (((sigs) & 0xffu) == mine ? 1u : 0u) + (((sigs >> 8) & 0xffu) == mine ? 1u : 0u) +
                                 (((sigs >> 16) & 0xffu) == mine ? 1u : 0u) + ((sigs >> 24) == mine ? 1u : 0u);
            if (cnt & 1u) odd ^= 1u << i;
            if (cnt != 0u && (g & GEO_AFTER)) notlast |= 1u << i;
        }
    }
    return ~odd & ~notlast & valid;   // the other occurrences are even in number and none comes later
}

// A cell tile takes part when it can hold a boundary side: any tile that is not uniformly full / empty, and
// full tiles on the border of the lattice.  A full tile elsewhere has none: all 17^3 corner points of its cells
// are strictly negative and un-warped, so the lattice tet across any of its lattice faces shares three strictly
// negative nodes with it — it was kept, it is not all-zero (no exterior test), and whatever trim_spikes did to
// it (nothing, or case 2 when the fourth node is positive) the shared face is a whole face of one of its pieces.
__device__ __forceinline__ bool boundary_tile_candidate(const Stuff& S, uint64_t tile, uint8_t cu) {
    if (cu == CU_EMPTY) return false;
    if (cu == CU_ACTIVE) return true;
    uint32_t X0, Y0, Z0;
    tile_origin(S.g, tile, X0, Y0, Z0);
    return X0 == 0u || Y0 == 0u || Z0 == 0u || X0 + TS >= S.L.nx || Y0 + TS >= S.L.ny || Z0 + TS >= S.L.nz;
}

// The cell tiles that can hold a boundary side, as an ORDERED work list (flag, scan, scatter): the sides are
// written tile by tile in list order, so the list decides their order.
static __global__ void k_boundary_tile_flags(Stuff S, uint64_t ntiles, unsigned long long* __restrict__ flag) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ntiles) return;
    uint32_t X0, Y0, Z0;
    tile_origin(S.g, t, X0, Y0, Z0);
    const uint32_t zlo = max(Z0, S.c0), zhi = min(Z0 + TS, S.c1);
    flag[t] = (X0 >= S.L.nx || Y0 >= S.L.ny || zhi <= zlo || !boundary_tile_candidate(S, t, S.cuni[t])) ? 0ull : 1ull;
}
static __global__ void k_boundary_tile_list(uint64_t ntiles, const unsigned long long* __restrict__ flag,
                                            const unsigned long long* __restrict__ pos, uint32_t* __restrict__ list) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < ntiles && flag[t]) list[pos[t]] = (uint32_t)t;
}

constexpr size_t BND_SMEM = (size_t)BH3 * 8 /*cell words*/ + (size_t)TILE_POINTS * 2 /*list*/ + (size_t)5 * TILE_POINTS * 2 /*masks*/;

// ONE pass over the listed tiles, a tile at a time per block:
//   1. the boundary sides of every lattice tet of the tile's owned cells as a 12-bit mask (bit 4 q + s) in
//      shared memory.  Only cells next to the surface do any work: the cells that are not full, and full cells
//      with a neighbour that is not; their lattice tets are dealt to the threads one by one;
//   2. room for the tile's sides is taken from a running counter (nobody waits for anybody);
//   3. the masks are expanded into (elem, side) pairs in emission order (x-rows of cells, lattice tets, pieces,
//      sides) at that place of a STAGING buffer; slot_first / slot_count record where.
// k_boundary_order then copies the tiles' runs into list order (a scan of the counts gives the places), so that
// the result is sorted by (elem, side) whatever order the blocks ran in.
// `capacity` sides fit the staging buffer; a tile that would run over it is counted but not written (the host
// sees total > capacity and repeats the pass with the exact size).
static __global__ void __launch_bounds__(TILE_THREADS)
k_boundary_sides(Stuff S, const uint32_t* __restrict__ tile_list, const unsigned long long* __restrict__ n_list,
                 unsigned long long* __restrict__ cursor, unsigned long long* __restrict__ slot_first,
                 unsigned long long* __restrict__ slot_count, unsigned long long capacity,
                 unsigned long long* __restrict__ ctl_total, unsigned long long tet_base, unsigned long long* __restrict__ sides) {
    extern __shared__ __align__(16) unsigned char bnd_smem[];
    unsigned long long* s_cell = reinterpret_cast<unsigned long long*>(bnd_smem);
    uint16_t* s_list = reinterpret_cast<uint16_t*>(s_cell + BH3);
    uint16_t* s_mask = s_list + TILE_POINTS;
    __shared__ uint8_t s_cu[27];
    __shared__ unsigned long long s_item, s_base;
    __shared__ uint32_t s_n;
    __shared__ uint32_t s_scan[9];
    __shared__ uint32_t s_rowsides[TILE_THREADS];   // sides of every x-row of cells
    const Lattice& L = S.L;
    const ShapeTables* __restrict__ T = S.shapes;
    for (;;) {
        __syncthreads();  // the previous tile is done with shared memory
        if (threadIdx.x == 0) s_item = atomicAdd(cursor, 1ull);
        __syncthreads();
        const unsigned long long slot = s_item, nslots = *n_list;
        if (slot >= nslots) break;
        const uint32_t tile = tile_list[slot];
        uint32_t X0, Y0, Z0;
        tile_origin(S.g, tile, X0, Y0, Z0);
        const uint32_t zlo = max(Z0, S.c0), zhi = min(Z0 + TS, S.c1);
        if (threadIdx.x == 0) s_n = 0u;
        s_rowsides[threadIdx.x] = 0u;
        // the tile's cells and the cells around it
        stage_cell_words(S, tile, X0, Y0, Z0, s_cell, s_cu);
        {
            uint4* mz = reinterpret_cast<uint4*>(s_mask);
            for (uint32_t i = threadIdx.x; i < 5u * TILE_POINTS * 2u / 16u; i += TILE_THREADS) mz[i] = make_uint4(0u, 0u, 0u, 0u);
        }
        __syncthreads();
        // candidate cells: owned, emitting something, and not a full cell surrounded by full cells
        for (int rnd = 0; rnd < TILE_ROUNDS; ++rnd) {
            const uint32_t j = (uint32_t)rnd * TILE_THREADS + threadIdx.x;
            const uint32_t lx = j & 15u, ly = (j >> 4) & 15u, lz = j >> 8;
            if (X0 + lx >= L.nx || Y0 + ly >= L.ny || Z0 + lz < zlo || Z0 + lz >= zhi) continue;
            const int at = (int)((lz + 1u) * BH + (ly + 1u)) * BH + (int)(lx + 1u);
            const unsigned long long w = s_cell[at];
            if (w == 0ull) continue;
            const bool cand = !(w & CI_FULL) ||
                              !(s_cell[at - 1] & s_cell[at + 1] & s_cell[at - BH] & s_cell[at + BH] & s_cell[at - BH * BH] & s_cell[at + BH * BH] & CI_FULL);
            if (cand) s_list[atomicAdd(&s_n, 1u)] = (uint16_t)j;
        }
        __syncthreads();
        const uint32_t ncand = s_n;
        for (uint32_t w = threadIdx.x; w < 5u * ncand; w += TILE_THREADS) {
            const uint32_t j = s_list[w / 5u], t = w % 5u;
            const uint32_t lx = j & 15u, ly = (j >> 4) & 15u, lz = j >> 8;
            const int at = (int)((lz + 1u) * BH + (ly + 1u)) * BH + (int)(lx + 1u);
            const uint32_t A = cell_shape(s_cell[at], (int)t);
            if (A == SHAPE_NONE) continue;
            const uint32_t p = ((X0 + lx) & 1u) | (((Y0 + ly) & 1u) << 1) | (((Z0 + lz) & 1u) << 2);
            const uint32_t mask = tet_boundary_mask(T, s_cell, at, p, (int)t, A);
            if (mask) {
                s_mask[t * TILE_POINTS + j] = (uint16_t)mask;
                atomicAdd(&s_rowsides[j >> 4], (uint32_t)__popc(mask));
            }
        }
        __syncthreads();
        // every thread owns one x-row of cells: its tets and sides, then the offsets from one block scan
        const uint32_t ly = threadIdx.x & 15u, lz = threadIdx.x >> 4;
        const bool row_owned = Y0 + ly < L.ny && Z0 + lz >= zlo && Z0 + lz < zhi;
        const uint32_t nx_row = (row_owned && ncand != 0u) ? min(TS, L.nx - X0) : 0u;
        const int at0 = (int)((lz + 1u) * BH + (ly + 1u)) * BH + 1;
        uint32_t ntet = 0;
        const uint32_t nside = s_rowsides[threadIdx.x];
        for (uint32_t i = 0; i < nx_row; ++i) ntet += ci_count((uint16_t)s_cell[at0 + (int)i]);
        uint32_t tot;
        const uint32_t ex = block_exscan(ntet | (nside << 16), &tot, s_scan);  // <= 61440 tets, sides are far fewer
        const unsigned long long tile_sides = tot >> 16;
        if (threadIdx.x == 0) {
            const unsigned long long base = tile_sides ? atomicAdd(ctl_total, tile_sides) : 0ull;
            s_base = base;
            slot_first[slot] = base;
            slot_count[slot] = base + tile_sides <= capacity ? tile_sides : 0ull;   // (low half: scanned by scan_tiles)
        }
        __syncthreads();
        if (tile_sides == 0ull || s_base + tile_sides > capacity) continue;
        uint64_t elem = tet_base + S.ttbase[tile] + (ex & 0xffffu);
        uint64_t o = s_base + (ex >> 16);
        for (uint32_t i = 0; i < nx_row && nside != 0u; ++i) {
            const uint32_t j = threadIdx.x * TS + i;
            const uint16_t ci = (uint16_t)s_cell[at0 + (int)i];
            for (uint32_t t = 0; t < 5u; ++t) {
                for (uint32_t mask = s_mask[t * TILE_POINTS + j]; mask; mask &= mask - 1u) {
                    const uint32_t bb = (uint32_t)__ffs((int)mask) - 1u;
                    reinterpret_cast<ulonglong2*>(sides)[o++] = make_ulonglong2(elem + (bb >> 2), bb & 3u);
                }
                elem += ci_tet_n(ci, (int)t);
            }
        }
    }
}

// the tiles' runs of sides, from where they were written to their place in list order
static __global__ void __launch_bounds__(256)
k_boundary_order(uint64_t nslots_max, const unsigned long long* __restrict__ n_list, const unsigned long long* __restrict__ slot_first,
                 const unsigned long long* __restrict__ slot_count, const unsigned long long* __restrict__ slot_pos,
                 const ulonglong2* __restrict__ staged, ulonglong2* __restrict__ sides) {
    const uint64_t nslots = min((uint64_t)*n_list, nslots_max);
    for (uint64_t slot = blockIdx.x; slot < nslots; slot += gridDim.x) {
        const unsigned long long n = slot_count[slot], src = slot_first[slot], dst = slot_pos[slot];
        for (unsigned long long i = threadIdx.x; i < n; i += blockDim.x) sides[dst + i] = staged[src + i];
    }
}

// ---- side-sets ----------------------------------------------------------------------------------
// src/main.rs:87-106: a surface side belongs to a boundary object's side-set iff all three nodes satisfy
// |approx_value(p, EPSILON)| <= EPSILON.  The object is evaluated once per surface VERTEX:
//   1 + 2. k_mark_side_nodes: flag the nodes of the surface sides (own vertices, and the slice of the rank
//      below's vertices a slab may reference); whoever flags a vertex first appends it to a dense list
//   3. k_eval_side_nodes: value at every listed vertex -> flag 2 (on the object) or 1
//   4. k_sideset_select: sides whose three nodes carry flag 2, counted per block, scanned, written in order
template <typename IndexT>
__device__ __forceinline__ long long side_node(const IndexT* __restrict__ tets, unsigned long long tet_base,
                                               const unsigned long long* __restrict__ sides, uint64_t i, int k) {
    const unsigned long long e = sides[2 * i] - tet_base, s = sides[2 * i + 1];
    return (long long)tets[4 * e + C_FACE[s][k]];
}
// flag index of a global vertex id: own vertices first, then the slice of the rank below; -1: not here
__device__ __forceinline__ long long flag_index(long long gid, long long vertex_base, long long n_verts, long long lower_first,
                                                long long lower_count) {
    if (gid >= vertex_base) return gid - vertex_base;
    const long long lid = gid - lower_first;
    return (lid >= 0 && lid < lower_count) ? n_verts + lid : -1;
}

// list != null: the thread that flags a vertex first (claimed through the aligned 32-bit word of its flag byte)
// also appends it to the list of surface vertices (any order), *count = its length
template <typename IndexT>
__global__ void k_mark_side_nodes(const IndexT* __restrict__ tets, unsigned long long tet_base,
                                  const unsigned long long* __restrict__ sides, uint64_t n, long long vertex_base,
                                  long long n_verts, long long lower_first, long long lower_count, uint8_t* __restrict__ flag,
                                  unsigned long long* __restrict__ list, unsigned long long* __restrict__ count) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const long long f = flag_index(side_node(tets, tet_base, sides, i, k), vertex_base, n_verts, lower_first, lower_count);
        if (f < 0 || flag[f] != 0) continue;
        if (list == nullptr) { flag[f] = 1; continue; }
        const unsigned bit = 1u << (8u * (unsigned)(f & 3));
        const unsigned old = atomicOr(reinterpret_cast<unsigned*>(flag + (f & ~3ll)), bit);
        if (!(old & bit)) list[atomicAdd(count, 1ull)] = (unsigned long long)f;
    }
}

static __global__ void __launch_bounds__(256)
k_eval_side_nodes(const uint64_t* __restrict__ prog, const double* __restrict__ coords, long long n_verts,
                  const double* __restrict__ lower_coords, const unsigned long long* __restrict__ list,
                  const unsigned long long* __restrict__ count, uint8_t* __restrict__ flag) {
    const uint64_t n = *count;
    const double EPS = 2.220446049250313e-16;  // f64::EPSILON
    for (uint64_t base = (uint64_t)blockIdx.x * blockDim.x; base < n; base += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t i = base + threadIdx.x;
        const unsigned long long f = list[i < n ? i : n - 1];
        const double* p = (long long)f < n_verts ? coords + 3 * f : lower_coords + 3 * (f - (unsigned long long)n_verts);
        const double phi = vm_eval<true>(prog, p[0], p[1], p[2]);
        if (i < n) flag[f] = (phi > EPS || phi < -EPS) ? 1 : 2;
    }
}

// EMIT = false: per block of 256 sides the number of members -> block_io; EMIT = true: write them
template <typename IndexT, bool EMIT>
__global__ void __launch_bounds__(256)
k_sideset_select(const IndexT* __restrict__ tets, unsigned long long tet_base, const unsigned long long* __restrict__ sides,
                 uint64_t n, long long vertex_base, long long n_verts, long long lower_first, long long lower_count,
                 const uint8_t* __restrict__ flag, unsigned long long* __restrict__ block_io,
                 unsigned long long* __restrict__ out) {
    __shared__ uint32_t s_w[8];
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool member = i < n;
    if (member) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const long long f = flag_index(side_node(tets, tet_base, sides, i, k), vertex_base, n_verts, lower_first, lower_count);
            member = member && f >= 0 && flag[f] == 2;
        }
    }
    const unsigned b = __ballot_sync(0xffffffffu, member);
    const unsigned lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
    if (lane == 0u) s_w[wid] = (uint32_t)__popc(b);
    __syncthreads();
    uint32_t before = 0, total = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) { if (w < (int)wid) before += s_w[w]; total += s_w[w]; }
    if (!EMIT) {
        if (threadIdx.x == 0) block_io[blockIdx.x] = total;
    } else if (member) {
        const uint64_t o = block_io[blockIdx.x] + before + __popc(b & ((1u << lane) - 1u));
        out[2 * o] = sides[2 * i];
        out[2 * o + 1] = sides[2 * i + 1];
    }
}

// ---- host orchestration ---------------------------------------------------------------------------
namespace {
struct Scratch {  // pooled buffers + events, released on every path
    mc_ctx* c;
    std::vector<DevBuf*> bufs;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    explicit Scratch(mc_ctx* ctx) : c(ctx) {}
    int alloc(size_t bytes, DevBuf& b) { bufs.push_back(&b); return c->alloc(bytes, b); }
    ~Scratch() {
        for (DevBuf* b : bufs) c->release(*b);
        if (e0) cudaEventDestroy(e0);
        if (e1) cudaEventDestroy(e1);
    }
};
}  // namespace

int compute_boundary(mc_mesh* m) {
    mc_ctx* c = m->ctx;
    MC_CUDA(cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    const Lattice& L = m->L;
    if (L.nx >= (1u << 21) || L.ny >= (1u << 21) || L.nz >= (1u << 21))
        return fail(MC_ERR_LIMIT, "Mesh::boundary: more than 2^21 cells along an axis");
    if (!c->boundary_attr_set) {
        MC_CUDA(cudaFuncSetAttribute(k_boundary_sides, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BND_SMEM));
        c->boundary_attr_set = true;
    }
    Scratch sc(c);
    MC_CUDA(cudaEventCreate(&sc.e0));
    MC_CUDA(cudaEventCreate(&sc.e1));
    MC_CUDA(cudaEventRecord(sc.e0, s));
    DevBuf fflag, fpos, dctl, dlist, sfirst, scount, spos, staged;
    int rc = sc.alloc((m->n_tiles + 1) * 8, fflag);
    if (rc == MC_OK) rc = sc.alloc((m->n_tiles + 1) * 8, fpos);
    if (rc == MC_OK) rc = sc.alloc(8 * 8, dctl);   // [0] number of sides, [2..3] list length (scan totals), [4] cursor, [6..7] scan totals
    if (rc == MC_OK) rc = sc.alloc(std::max<uint64_t>(m->n_tiles, 1) * 4, dlist);
    if (rc == MC_OK) rc = sc.alloc((m->n_tiles + 1) * 8, sfirst);
    if (rc == MC_OK) rc = sc.alloc((m->n_tiles + 1) * 8, scount);
    if (rc == MC_OK) rc = sc.alloc((m->n_tiles + 1) * 8, spos);
    if (rc != MC_OK) return rc;
    unsigned long long* ctl = (unsigned long long*)dctl.p;
    const Stuff S = make_stuff(m);
    m->n_sides = 0;
    m->h_sides.clear();
    m->have_h_sides = false;
    c->release(m->d_sides);
    if (m->n_tiles > 0 && m->n_tets > 0) {
        const unsigned tb = (unsigned)((m->n_tiles + 255) / 256);
        k_boundary_tile_flags<<<tb, 256, 0, s>>>(S, m->n_tiles, (unsigned long long*)fflag.p);
        MC_CUDA(cudaMemsetAsync(dctl.p, 0, 64, s));
        rc = scan_tiles(c, fflag.p, m->n_tiles, fpos.p, nullptr, m->d_scan_sums.p, ctl + 2);
        if (rc != MC_OK) return rc;
        k_boundary_tile_list<<<tb, 256, 0, s>>>(m->n_tiles, (const unsigned long long*)fflag.p, (const unsigned long long*)fpos.p, (uint32_t*)dlist.p);
        c->launches += 2;
        // Room for the sides: a generous guess from what the host knows (surface cells, border of the slab);
        // the pass counts exactly, and is repeated with the exact size in the rare case the guess was short.
        const uint64_t nl = m->c1 - m->c0;
        uint64_t capacity = std::max<uint64_t>(1ull << 20, 8ull * m->counters[CN_SURF_CELLS] +
                                                           12ull * ((uint64_t)L.nx * L.ny + ((uint64_t)L.nx + L.ny) * nl));
        capacity = std::min<uint64_t>(capacity, 4ull * m->n_tets);
        if (const char* cap = std::getenv("MC_BOUNDARY_CAPACITY")) capacity = std::max<long long>(1, std::atoll(cap));  // tests: force the second pass
        const unsigned grid = (unsigned)std::min<uint64_t>(m->n_tiles, (uint64_t)c->sm_count * 2);  // 2 blocks of BND_SMEM fit an SM
        for (int attempt = 0; attempt < 2; ++attempt) {
            rc = c->alloc(capacity * 16, staged);
            if (rc != MC_OK) return rc;
            MC_CUDA(cudaMemsetAsync(scount.p, 0, (m->n_tiles + 1) * 8, s));
            MC_CUDA(cudaMemsetAsync(ctl, 0, 8, s));
            MC_CUDA(cudaMemsetAsync(ctl + 4, 0, 8, s));
            k_boundary_sides<<<grid, TILE_THREADS, BND_SMEM, s>>>(S, (const uint32_t*)dlist.p, ctl + 2, ctl + 4, (unsigned long long*)sfirst.p,
                                                                 (unsigned long long*)scount.p, capacity, ctl, m->tet_base,
                                                                 (unsigned long long*)staged.p);
            c->launches++;
            MC_CUDA(cudaGetLastError());
            unsigned long long total = 0;
            MC_CUDA(cudaMemcpyAsync(&total, ctl, 8, cudaMemcpyDeviceToHost, s));
            MC_CUDA(cudaStreamSynchronize(s));
            m->n_sides = total;
            if (total <= capacity) break;
            if (attempt == 1) { c->release(staged); return fail(MC_ERR_STATE, "Mesh::boundary: the side count changed between two passes"); }
            c->release(staged);
            capacity = total;
        }
        if (m->n_sides) {
            rc = c->alloc(m->n_sides * 16, m->d_sides);
            if (rc == MC_OK) rc = scan_tiles(c, scount.p, m->n_tiles, spos.p, nullptr, m->d_scan_sums.p, ctl + 6);
            if (rc != MC_OK) { c->release(staged); return rc; }
            k_boundary_order<<<(unsigned)std::min<uint64_t>(m->n_tiles, (uint64_t)c->sm_count * 16), 256, 0, s>>>(
                m->n_tiles, ctl + 2, (const unsigned long long*)sfirst.p, (const unsigned long long*)scount.p,
                (const unsigned long long*)spos.p, (const ulonglong2*)staged.p, (ulonglong2*)m->d_sides.p);
            c->launches++;
            MC_CUDA(cudaGetLastError());
        }
        sc.bufs.push_back(&staged);   // released with the other scratch buffers (after the event below)
    }
    MC_CUDA(cudaEventRecord(sc.e1, s));
    MC_CUDA(cudaEventSynchronize(sc.e1));
    float ms = 0;
    MC_CUDA(cudaEventElapsedTime(&ms, sc.e0, sc.e1));
    m->stage_ms[7] = ms;
    m->have_boundary = true;
    return MC_OK;
}

int boundary_host_view(mc_mesh* m) {
    if (m->have_h_sides) return MC_OK;
    mc_ctx* c = m->ctx;
    MC_CUDA(cudaSetDevice(c->device));
    m->h_sides.resize(m->n_sides * 2);
    if (m->n_sides) {
        MC_CUDA(cudaMemcpyAsync(m->h_sides.data(), m->d_sides.p, m->n_sides * 16, cudaMemcpyDeviceToHost, c->stream));
        MC_CUDA(cudaStreamSynchronize(c->stream));
    }
    m->have_h_sides = true;
    return MC_OK;
}

template <typename IndexT>
static int sideset_impl(mc_mesh* m, const Program& p, SideSetDev& out) {
    mc_ctx* c = m->ctx;
    cudaStream_t s = c->stream;
    Scratch sc(c);
    MC_CUDA(cudaEventCreate(&sc.e0));
    MC_CUDA(cudaEventCreate(&sc.e1));
    MC_CUDA(cudaEventRecord(sc.e0, s));
    const uint64_t ns = m->n_sides;
    const long long nv = (long long)m->n_verts, lf = (long long)m->lower_first, lc = (long long)m->lower_count;
    const uint64_t nflag = (uint64_t)nv + (uint64_t)lc;
    const IndexT* tets = (const IndexT*)m->d_tets.p;
    const unsigned long long* sides = (const unsigned long long*)m->d_sides.p;
    const unsigned nb = (unsigned)((ns + 255) / 256);
    int rc = MC_OK;
    // 1 + 2: the surface vertices, once per mesh
    if (!m->d_side_nodes.p) {
        rc = c->alloc(nflag + 4, m->d_side_flag);
        if (rc == MC_OK) rc = c->alloc(std::max<uint64_t>(std::min<uint64_t>(3 * ns, nflag), 1) * 8 + 8, m->d_side_nodes);
        if (rc != MC_OK) return rc;
        MC_CUDA(cudaMemsetAsync(m->d_side_flag.p, 0, nflag + 4, s));
        MC_CUDA(cudaMemsetAsync(m->d_side_nodes.p, 0, 8, s));
        k_mark_side_nodes<IndexT><<<nb, 256, 0, s>>>(tets, m->tet_base, sides, ns, (long long)m->vertex_base, nv, lf, lc,
                                                     (uint8_t*)m->d_side_flag.p, (unsigned long long*)m->d_side_nodes.p + 1,
                                                     (unsigned long long*)m->d_side_nodes.p);
        c->launches += 1;
        MC_CUDA(cudaGetLastError());
    }
    // 3: the boundary object at every surface vertex
    DevBuf dp, bcnt, bbase, dtot, sums;
    sc.bufs.push_back(&dp);
    rc = upload_program(c, p, dp);
    if (rc == MC_OK) rc = sc.alloc(((size_t)nb + 1) * 8, bcnt);
    if (rc == MC_OK) rc = sc.alloc(((size_t)nb + 1) * 8, bbase);
    if (rc == MC_OK) rc = sc.alloc(32, dtot);
    if (rc == MC_OK) rc = sc.alloc(((size_t)nb / SCAN_BLOCK + 2) * 16, sums);
    if (rc != MC_OK) return rc;
    k_eval_side_nodes<<<(unsigned)c->sm_count * 8u, 256, 0, s>>>((const uint64_t*)dp.p, (const double*)m->d_coords.p, nv, m->lower_coords,
                                                                (const unsigned long long*)m->d_side_nodes.p + 1,
                                                                (const unsigned long long*)m->d_side_nodes.p, (uint8_t*)m->d_side_flag.p);
    // 4: select
    k_sideset_select<IndexT, false><<<nb, 256, 0, s>>>(tets, m->tet_base, sides, ns, (long long)m->vertex_base, nv, lf, lc,
                                                       (const uint8_t*)m->d_side_flag.p, (unsigned long long*)bcnt.p, nullptr);
    c->launches += 2;
    MC_CUDA(cudaGetLastError());
    rc = scan_tiles(c, bcnt.p, nb, bbase.p, nullptr, sums.p, dtot.p);
    if (rc != MC_OK) return rc;
    unsigned long long h[2];
    MC_CUDA(cudaMemcpyAsync(h, dtot.p, 16, cudaMemcpyDeviceToHost, s));
    MC_CUDA(cudaStreamSynchronize(s));
    out.n = h[0];
    if (out.n) {
        rc = c->alloc(out.n * 16, out.d_pairs);
        if (rc != MC_OK) return rc;
        k_sideset_select<IndexT, true><<<nb, 256, 0, s>>>(tets, m->tet_base, sides, ns, (long long)m->vertex_base, nv, lf, lc,
                                                          (const uint8_t*)m->d_side_flag.p, (unsigned long long*)bbase.p,
                                                          (unsigned long long*)out.d_pairs.p);
        c->launches++;
        MC_CUDA(cudaGetLastError());
    }
    MC_CUDA(cudaEventRecord(sc.e1, s));
    MC_CUDA(cudaEventSynchronize(sc.e1));  // dp is recycled when `sc` goes out of scope
    float ms = 0;
    MC_CUDA(cudaEventElapsedTime(&ms, sc.e0, sc.e1));
    m->stage_ms[8] += ms;
    return MC_OK;
}

int compute_sideset(mc_mesh* m, const mc_tape* tape, int32_t root, SideSetDev& out) {
    mc_ctx* c = m->ctx;
    MC_CUDA(cudaSetDevice(c->device));
    out.n = 0;
    if (m->n_sides == 0) return MC_OK;
    Program p;
    std::string err;
    // boundary.approx_value(point, f64::EPSILON), src/main.rs:95
    int rc = tape->compile(root, std::numeric_limits<double>::epsilon(), p, err);
    if (rc != MC_OK) return fail(rc, err);
    return m->index_bytes == 4 ? sideset_impl<uint32_t>(m, p, out) : sideset_impl<unsigned long long>(m, p, out);
}

}  // namespace mc

// ---- Mesh::to_boundary (mesh/src/lib.rs:359-386) -------------------------------------------------
// The surface as a triangle mesh: one Tri3 per boundary side (the side's nodes in Tet4::face order), the
// vertices compacted to the referenced ones (Mesh::compact, mesh/src/lib.rs:318-339; numbering by vertex
// id instead of first appearance — compared after canonical sorting like every other mesh here).
namespace mc {

static __global__ void k_flag_prefix(const uint8_t* __restrict__ flag, uint64_t n, unsigned long long* __restrict__ block_cnt) {
    __shared__ uint32_t s_w[8];
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned b = __ballot_sync(0xffffffffu, i < n && flag[i] != 0);
    if ((threadIdx.x & 31u) == 0u) s_w[threadIdx.x >> 5] = (uint32_t)__popc(b);
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int w = 0; w < 8; ++w) t += s_w[w];
        block_cnt[blockIdx.x] = t;
    }
}
// new id of every flagged vertex (ids in vertex order) + its coordinates
static __global__ void k_surface_vertices(const uint8_t* __restrict__ flag, uint64_t n, const unsigned long long* __restrict__ block_base,
                                          const double* __restrict__ coords, unsigned long long* __restrict__ new_id,
                                          double* __restrict__ out_xyz) {
    __shared__ uint32_t s_w[8];
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool on = i < n && flag[i] != 0;
    const unsigned b = __ballot_sync(0xffffffffu, on);
    const unsigned lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
    if (lane == 0u) s_w[wid] = (uint32_t)__popc(b);
    __syncthreads();
    uint32_t before = 0;
    for (unsigned w = 0; w < wid; ++w) before += s_w[w];
    if (on) {
        const unsigned long long id = block_base[blockIdx.x] + before + __popc(b & ((1u << lane) - 1u));
        new_id[i] = id;
        out_xyz[3 * id] = coords[3 * i]; out_xyz[3 * id + 1] = coords[3 * i + 1]; out_xyz[3 * id + 2] = coords[3 * i + 2];
    }
}
template <typename IndexT>
__global__ void k_surface_tris(const IndexT* __restrict__ tets, unsigned long long tet_base, const unsigned long long* __restrict__ sides,
                               uint64_t n, long long vertex_base, const unsigned long long* __restrict__ new_id,
                               unsigned long long* __restrict__ tris) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
#pragma unroll
    for (int k = 0; k < 3; ++k) tris[3 * i + k] = new_id[side_node(tets, tet_base, sides, i, k) - vertex_base];
}

int compute_surface(mc_mesh* m, mc_surface* out) {
    mc_ctx* c = m->ctx;
    MC_CUDA(cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    if (m->lower_count != 0 || m->vertex_base != 0)
        return fail(MC_ERR_UNSUPPORTED, "Mesh::to_boundary on a z-slab: the surface vertices of the shared plane live on the slab below");
    const uint64_t ns = m->n_sides, nv = m->n_verts;
    out->n_tris = ns;
    out->n_verts = 0;
    if (ns == 0) return MC_OK;
    Scratch sc(c);
    DevBuf flag, newid, bcnt, bbase, dtot, sums;
    const unsigned nbv = (unsigned)((nv + 255) / 256), nbs = (unsigned)((ns + 255) / 256);
    int rc = sc.alloc(std::max<uint64_t>(nv, 1), flag);
    if (rc == MC_OK) rc = sc.alloc(std::max<uint64_t>(nv, 1) * 8, newid);
    if (rc == MC_OK) rc = sc.alloc(((size_t)nbv + 1) * 8, bcnt);
    if (rc == MC_OK) rc = sc.alloc(((size_t)nbv + 1) * 8, bbase);
    if (rc == MC_OK) rc = sc.alloc(32, dtot);
    if (rc == MC_OK) rc = sc.alloc(((size_t)nbv / SCAN_BLOCK + 2) * 16, sums);
    if (rc != MC_OK) return rc;
    MC_CUDA(cudaMemsetAsync(flag.p, 0, std::max<uint64_t>(nv, 1), s));
    const unsigned long long* sides = (const unsigned long long*)m->d_sides.p;
    if (m->index_bytes == 4)
        k_mark_side_nodes<uint32_t><<<nbs, 256, 0, s>>>((const uint32_t*)m->d_tets.p, m->tet_base, sides, ns, 0, (long long)nv, 0, 0, (uint8_t*)flag.p, nullptr, nullptr);
    else
        k_mark_side_nodes<unsigned long long><<<nbs, 256, 0, s>>>((const unsigned long long*)m->d_tets.p, m->tet_base, sides, ns, 0, (long long)nv, 0, 0, (uint8_t*)flag.p, nullptr, nullptr);
    k_flag_prefix<<<nbv, 256, 0, s>>>((const uint8_t*)flag.p, nv, (unsigned long long*)bcnt.p);
    c->launches += 2;
    MC_CUDA(cudaGetLastError());
    rc = scan_tiles(c, bcnt.p, nbv, bbase.p, nullptr, sums.p, dtot.p);
    if (rc != MC_OK) return rc;
    unsigned long long h[2];
    MC_CUDA(cudaMemcpyAsync(h, dtot.p, 16, cudaMemcpyDeviceToHost, s));
    MC_CUDA(cudaStreamSynchronize(s));
    out->n_verts = h[0];
    rc = c->alloc(std::max<uint64_t>(out->n_verts, 1) * 24, out->d_coords);
    if (rc == MC_OK) rc = c->alloc(ns * 24, out->d_tris);
    if (rc != MC_OK) return rc;
    k_surface_vertices<<<nbv, 256, 0, s>>>((const uint8_t*)flag.p, nv, (const unsigned long long*)bbase.p, (const double*)m->d_coords.p,
                                           (unsigned long long*)newid.p, (double*)out->d_coords.p);
    if (m->index_bytes == 4)
        k_surface_tris<uint32_t><<<nbs, 256, 0, s>>>((const uint32_t*)m->d_tets.p, m->tet_base, sides, ns, 0, (const unsigned long long*)newid.p,
                                                     (unsigned long long*)out->d_tris.p);
    else
        k_surface_tris<unsigned long long><<<nbs, 256, 0, s>>>((const unsigned long long*)m->d_tets.p, m->tet_base, sides, ns, 0,
                                                               (const unsigned long long*)newid.p, (unsigned long long*)out->d_tris.p);
    c->launches += 2;
    MC_CUDA(cudaGetLastError());
    out->h_coords.resize(out->n_verts * 3);
    out->h_tris.resize(ns * 3);
    if (out->n_verts) MC_CUDA(cudaMemcpyAsync(out->h_coords.data(), out->d_coords.p, out->n_verts * 24, cudaMemcpyDeviceToHost, s));
    MC_CUDA(cudaMemcpyAsync(out->h_tris.data(), out->d_tris.p, ns * 24, cudaMemcpyDeviceToHost, s));
    MC_CUDA(cudaStreamSynchronize(s));
    return MC_OK;
}

}  // namespace mc
