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
// So every (lattice tet, piece, side) counts the other occurrences of its face locally —
// re-classifying the neighbouring lattice tet from the SDF field, like the tet emission does —
// and reports the face when the total is odd and it is the last of them in a fixed order.  This is
// the hash-map rule verbatim (it even reproduces the faces that occur three times in the prism
// split of trim case 3), evaluated per face.  Only cells next to the surface do any work: a cell
// whose five lattice tets are emitted untouched ("full") next to full cells is skipped outright.
//
// Everything is slab-aware: the cell layers next to the slab are classified too (halo), so the tet
// across a slab boundary is re-classified locally and no exchange is needed.  Sides come out sorted by
// (elem, side) because tiles, cells, lattice tets and pieces are visited in emission order.
#include <algorithm>
#include <cstdlib>
#include <limits>

#include "mc_internal.hpp"
#include "mc_kernels.cuh"

namespace mc {

Stuff make_stuff(const mc_mesh* m);  // mc_mesher.cu

// ---- lattice topology helpers -----------------------------------------------------------------
// The lattice tet across lattice face k (= opposite node k of the flipped node order) of tet t of a
// cell with mirroring class p = fx | fy << 1 | fz << 2: (dx + 1) | (dy + 1) << 2 | (dz + 1) << 4 |
// t2 << 6.  Built on the host from Tile::LATTICE / Tile::TETS by brute force (build_nb_table).
static __constant__ uint16_t C_NB_TABLE[8][5][4];

static void build_nb_table(uint16_t tab[8][5][4]) {
    static const int LAT[8][3] = {{0, 0, 0}, {1, 0, 0}, {1, 1, 0}, {0, 1, 0}, {0, 0, 1}, {1, 0, 1}, {1, 1, 1}, {0, 1, 1}};
    static const int TETS[5][4] = {{0, 1, 5, 2}, {0, 2, 5, 7}, {0, 7, 5, 4}, {2, 6, 5, 7}, {0, 2, 7, 3}};
    // node n (flipped order) of tet t of the cell at offset o (cells), mirroring from the parity of p0 + o
    auto node = [&](const int par[3], const int off[3], int t, int n, int out[3]) {
        bool mir[3];
        for (int a = 0; a < 3; ++a) mir[a] = ((par[a] + off[a]) & 1) != 0;
        const bool flip = mir[0] ^ mir[1] ^ mir[2];
        const int k = TETS[t][(flip && n < 2) ? 1 - n : n];
        for (int a = 0; a < 3; ++a) out[a] = off[a] + (mir[a] ? 1 - LAT[k][a] : LAT[k][a]);
    };
    for (int p = 0; p < 8; ++p) {
        const int par[3] = {p & 1, (p >> 1) & 1, (p >> 2) & 1};
        const int zero[3] = {0, 0, 0};
        for (int t = 0; t < 5; ++t)
            for (int k = 0; k < 4; ++k) {
                int face[3][3], w = 0;
                for (int n = 0; n < 4; ++n)
                    if (n != k) node(par, zero, t, n, face[w++]);
                uint16_t found = 0xffffu;
                for (int dx = -1; dx <= 1; ++dx)
                    for (int dy = -1; dy <= 1; ++dy)
                        for (int dz = -1; dz <= 1; ++dz) {
                            if (std::abs(dx) + std::abs(dy) + std::abs(dz) > 1) continue;
                            const int off[3] = {dx, dy, dz};
                            for (int t2 = 0; t2 < 5; ++t2) {
                                if (dx == 0 && dy == 0 && dz == 0 && t2 == t) continue;
                                int hit = 0;
                                for (int n = 0; n < 4; ++n) {
                                    int q[3];
                                    node(par, off, t2, n, q);
                                    for (int f = 0; f < 3; ++f) hit += (q[0] == face[f][0] && q[1] == face[f][1] && q[2] == face[f][2]) ? 1 : 0;
                                }
                                if (hit == 3) found = (uint16_t)((dx + 1) | ((dy + 1) << 2) | ((dz + 1) << 4) | (t2 << 6));
                            }
                        }
                tab[p][t][k] = found;
            }
    }
}

__device__ __forceinline__ unsigned long long pt_id(uint32_t X, uint32_t Y, uint32_t Z) {
    return ((unsigned long long)Z << 42) | ((unsigned long long)Y << 21) | (unsigned long long)X;
}

// what a cell emitted: its info word, whether all five lattice tets came out untouched, whether anything did
struct CellInfo { uint16_t ci; bool full; bool any; };
__device__ __forceinline__ CellInfo cell_info(const Stuff& S, int64_t cx, int64_t cy, int64_t cz) {
    CellInfo r = {0, false, false};
    const Lattice& L = S.L;
    if (cx < 0 || cy < 0 || cz < 0 || cx >= (int64_t)L.nx || cy >= (int64_t)L.ny || cz >= (int64_t)L.nz) return r;
    if (cz < (int64_t)S.cl0 || cz >= (int64_t)S.cl1) return r;  // never reached for neighbours of owned cells
    const uint8_t cu = S.cuni[tile_of(S.g, (uint32_t)cx, (uint32_t)cy, (uint32_t)cz)];
    if (cu == CU_EMPTY) return r;
    if (cu == CU_FULL) { r.ci = CI_ALL_FULL; r.full = true; r.any = true; return r; }
    r.ci = S.cinfo[cidx(S, (uint32_t)cx, (uint32_t)cy, (uint32_t)cz)];
    r.full = (r.ci & CI_FULL) != 0;
    r.any = ci_count(r.ci) != 0;
    return r;
}

// state of a lattice tet, from its cell's info word: nothing emitted / emitted untouched (one piece,
// the lattice tet itself) / trimmed (needs classify_tet)
enum : int { TET_NONE = 0, TET_WHOLE = 1, TET_TRIM = 2 };
__device__ __forceinline__ int tet_state(uint16_t ci, int t) {
    if (ci & CI_FULL) return TET_WHOLE;
    if (ci_tet_n(ci, t) == 0u) return TET_NONE;
    return ((ci >> (CI_UNTOUCHED_SHIFT + t)) & 1u) ? TET_WHOLE : TET_TRIM;
}
// signature (see face_sig) of a whole lattice face: its three original nodes
constexpr uint32_t SIG_WHOLE_FACE = (1u << 1) | (1u << 2) | (1u << 4);
// side (Tet4::face row) opposite node k: FACE = {0,1,2} {0,2,3} {0,3,1} {1,3,2}
__device__ __forceinline__ int side_opposite(int k) { return k == 0 ? 3 : (k == 1 ? 1 : (k == 2 ? 2 : 0)); }

// the pieces of a TRIMMED lattice tet (node codes: 0..3 original node, 4 + 4 i + j cut on edge i-j); one
// copy of the classification code per kernel
static __device__ __noinline__ void trimmed_pieces(const Lattice& L, const Cell& c, int t, TetNodes& tn, TetOut& to) {
    // trimmed: kept by cut_lattice with vertices on both sides of zero, and not dropped by the exterior test
    load_tet_mixed(L, c, t, tn);
    classify_tet<false, true>(L, c, t, tn, nullptr, 0, to);
}

// lattice coordinates of the (flipped) nodes of lattice tet (c, t), without touching the field
__device__ __forceinline__ void tet_coords(const Cell& c, int t, TetNodes& tn) {
#pragma unroll
    for (int n = 0; n < 4; ++n) corner_xyz(c, tet_corner(c, t, n), tn.X[n], tn.Y[n], tn.Z[n]);
}

// signature of the face (q, s) of `to` relative to lattice face k of its lattice tet, or 0 when the
// face does not lie in it.  r0..r3 = rank (0..2) of node n among the three lattice points of face k
// by point id (the rank of node k is unused): both sides of the lattice face compute the same ranks.
// A node becomes a 3-bit set of ranks (original node: one bit, cut vertex: the two ends of its edge),
// the face the set of its three node sets (8 bits).
static __device__ __noinline__ uint32_t face_sig(unsigned long long nodes, int q, int s, int k, int r0, int r1, int r2, int r3) {
    uint32_t sig = 0u;
    for (int m = 0; m < 3; ++m) {
        const int code = (int)((nodes >> (5 * (4 * q + C_FACE[s][m]))) & 31ull);
        uint32_t nm;
        if (code < 4) {
            if (code == k) return 0u;
            nm = 1u << pick4(r0, r1, r2, r3, code);
        } else {
            const int i = (code - 4) >> 2, j = (code - 4) & 3;
            if (i == k || j == k) return 0u;
            nm = (1u << pick4(r0, r1, r2, r3, i)) | (1u << pick4(r0, r1, r2, r3, j));
        }
        sig |= 1u << nm;
    }
    return sig;
}

// order of two occurrences of a face in different lattice tets: a fixed total order of the lattice tets
__device__ __forceinline__ bool tet_after(const Cell& a, int ta, const Cell& b, int tb) {
    if (a.cz != b.cz) return a.cz > b.cz;
    if (a.cy != b.cy) return a.cy > b.cy;
    if (a.cx != b.cx) return a.cx > b.cx;
    return ta > tb;
}

// The info words of a tile's cells with a one-cell halo are staged in shared memory (BH^3 entries,
// 0 = nothing there / outside the lattice); `at` = index of the cell in that box.
constexpr int BH = (int)TS + 2;
// The lattice tet on the other side of lattice face k of tet t of the cell at `at` (mirroring class p):
// returns its state (TET_NONE also when there is none: lattice border); *at2 / *t2 / *ci2 describe it.
__device__ __forceinline__ int neighbour_tet(const uint16_t* __restrict__ s_cih, int at, uint32_t p, int t, int k, int* at2, int* t2,
                                             uint16_t* ci2) {
    const uint32_t e = C_NB_TABLE[p][t][k];
    *t2 = (int)(e >> 6);
    const int dx = (int)(e & 3u) - 1, dy = (int)((e >> 2) & 3u) - 1, dz = (int)((e >> 4) & 3u) - 1;
    *at2 = at + (dz * BH + dy) * BH + dx;
    *ci2 = s_cih[*at2];
    return tet_state(*ci2, *t2);
}

// node-code set (bit per code 0..19) of face (q, s) of a piece list, and the lattice faces (bit k =
// face opposite node k) every node of it lies in
static __device__ __noinline__ void face_codes(unsigned long long nodes, int q, int s, uint32_t& codes, uint32_t& fmask) {
    codes = 0u;
    fmask = 0xfu;
    for (int m = 0; m < 3; ++m) {
        const int code = (int)((nodes >> (5 * (4 * q + C_FACE[s][m]))) & 31ull);
        codes |= 1u << code;
        fmask &= code < 4 ? ~(1u << code) : ~((1u << ((code - 4) >> 2)) | (1u << ((code - 4) & 3)));
    }
    fmask &= 0xfu;
}

// what a block knows about the tile it works on
struct TileView {
    const uint16_t* cih;    // info words of the tile's cells with a one-cell halo (BH^3)
    const uint16_t* soff;   // surface-cell number of the tile's owned cells (piece cache), TS^3
    unsigned long long sbase;
    uint32_t X0, Y0, Z0, zlo, zhi;
    bool cached;            // the piece cache holds this tile's trimmed tets
};

// the pieces of TRIMMED tet t of cell c: from the cache k_tet_emit filled (owned cells of this tile) or by
// classifying it again; tn receives the node coordinates (and, on the second path, the values)
__device__ __forceinline__ void pieces_of_trimmed(const Stuff& S, const TileView& V, const Cell& c, int t, TetNodes& tn, TetOut& to) {
    const uint32_t lx = c.cx - V.X0, ly = c.cy - V.Y0, lz = c.cz - V.Z0;   // (wraps for cells before the tile)
    if (V.cached && lx < TS && ly < TS && lz < TS && c.cz >= V.zlo && c.cz < V.zhi) {
        const unsigned long long e = S.pieces[5ull * (V.sbase + V.soff[(lz * TS + ly) * TS + lx]) + (unsigned)t];
        tet_coords(c, t, tn);
        to.nodes = e & 0x0fffffffffffffffull;
        to.n = (int)(e >> 60);
        return;
    }
    trimmed_pieces(S.L, c, t, tn, to);
}

// signatures (up to four, 8 bits each) of the faces that the pieces of TRIMMED tet (c2, t2) have in the
// lattice face whose three points are the nodes != k of (gid[0..3])
static __device__ __noinline__ uint32_t trimmed_sigs_on_face(const Stuff& S, const TileView& V, const Cell& c2, int t2, unsigned long long g0,
                                                             unsigned long long g1, unsigned long long g2_, unsigned long long g3, int k) {
    TetNodes tn2;
    TetOut to2;
    pieces_of_trimmed(S, V, c2, t2, tn2, to2);
    const unsigned long long gid[4] = {g0, g1, g2_, g3};
    unsigned long long h[4];
    for (int n = 0; n < 4; ++n) h[n] = pt_id(tn2.X[n], tn2.Y[n], tn2.Z[n]);
    int k2 = 0, r2[4];
    for (int n = 0; n < 4; ++n) {
        int rk = 0;
        bool on = false;
        for (int o = 0; o < 4; ++o)
            if (o != k) { on = on || h[n] == gid[o]; rk += gid[o] < h[n] ? 1 : 0; }
        if (!on) k2 = n;
        r2[n] = rk;
    }
    uint32_t sigs = 0u;
    int nsig = 0;
    for (int q2 = 0; q2 < to2.n; ++q2)
        for (int s2 = 0; s2 < 4; ++s2) {
            uint32_t codes, fmask;
            face_codes(to2.nodes, q2, s2, codes, fmask);
            if (!((fmask >> k2) & 1u)) continue;  // not in the shared lattice face
            const uint32_t sg = face_sig(to2.nodes, q2, s2, k2, r2[0], r2[1], r2[2], r2[3]);
            if (sg && nsig < 4) { sigs |= sg << (8 * nsig); ++nsig; }
        }
    return sigs;
}

// Boundary sides of an UNTOUCHED lattice tet (one piece) from the info words alone: bit s.  Its four faces
// are whole lattice faces: matched by an untouched neighbour, unmatched next to nothing.  *heavy is set (and
// 0 returned) when a neighbour is trimmed: the caller queues the tet for tet_boundary_mask.
__device__ __forceinline__ uint32_t whole_tet_mask(const uint16_t* __restrict__ s_cih, int at, uint32_t p, int t, bool* heavy) {
    uint32_t out = 0u;
    *heavy = false;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        int at2, t2;
        uint16_t ci2;
        const int st2 = neighbour_tet(s_cih, at, p, t, k, &at2, &t2, &ci2);
        if (st2 == TET_TRIM) *heavy = true;
        else if (st2 == TET_NONE) out |= 1u << side_opposite(k);
    }
    return *heavy ? 0u : out;
}

// Boundary sides of the pieces of lattice tet (c, t), general procedure: bit 4 q + s.
static __device__ __noinline__ uint32_t tet_boundary_mask(const Stuff& S, const TileView& V, int at, const Cell& c, int t) {
    const uint16_t* __restrict__ s_cih = V.cih;
    const uint16_t ci = s_cih[at];
    const uint32_t p = (c.fx ? 1u : 0u) | (c.fy ? 2u : 0u) | (c.fz ? 4u : 0u);
    const int st = tet_state(ci, t);
    if (st == TET_NONE) return 0u;
    TetNodes tn;
    TetOut to;
    if (st == TET_WHOLE) {
        tet_coords(c, t, tn);
        to.n = 1;
        to.nodes = 0ull | (1ull << 5) | (2ull << 10) | (3ull << 15);
    } else {
        pieces_of_trimmed(S, V, c, t, tn, to);
    }
    if (to.n == 0) return 0u;
    unsigned long long gid[4];
    for (int n = 0; n < 4; ++n) gid[n] = pt_id(tn.X[n], tn.Y[n], tn.Z[n]);
    // the faces of the pieces: node-code sets and the lattice face each lies in
    uint32_t fc[12], fm[12];
    for (int i = 0; i < 12; ++i) {
        fc[i] = 0u; fm[i] = 0u;
        if ((i >> 2) < to.n) face_codes(to.nodes, i >> 2, i & 3, fc[i], fm[i]);
    }
    // the other side of the four lattice faces, looked at only when a face of ours lies in it
    uint32_t nb_sigs[4] = {0u, 0u, 0u, 0u};
    bool nb_after[4] = {false, false, false, false};
    uint32_t used = 0u;
    for (int i = 0; i < 12; ++i) used |= fm[i];
    for (int k = 0; k < 4; ++k) {
        if (!((used >> k) & 1u)) continue;
        int at2, t2;
        uint16_t ci2;
        const int st2 = neighbour_tet(s_cih, at, p, t, k, &at2, &t2, &ci2);
        if (st2 == TET_NONE) continue;
        const int d = at2 - at;  // the neighbour cell relative to this one
        const Cell c2 = make_cell(c.cx + (uint32_t)(d == 1 ? 1 : (d == -1 ? -1 : 0)), c.cy + (uint32_t)(d == BH ? 1 : (d == -BH ? -1 : 0)),
                                  c.cz + (uint32_t)(d == BH * BH ? 1 : (d == -BH * BH ? -1 : 0)));
        nb_after[k] = tet_after(c2, t2, c, t);
        nb_sigs[k] = st2 == TET_WHOLE ? SIG_WHOLE_FACE : trimmed_sigs_on_face(S, V, c2, t2, gid[0], gid[1], gid[2], gid[3], k);
    }
    uint32_t out = 0u;
    for (int i = 0; i < 12; ++i) {
        if ((i >> 2) >= to.n) continue;
        uint32_t cnt = 0u;
        bool last = true;
        // ... occurrences among the other pieces of this lattice tet
        for (int j = 0; j < 12; ++j)
            if (j != i && fc[j] == fc[i]) { ++cnt; if (j > i) last = false; }
        // ... and among the pieces of the lattice tet across the lattice face it lies in
        if (fm[i]) {
            const int k = __ffs((int)fm[i]) - 1;
            int r[4];
            for (int n = 0; n < 4; ++n) {
                int rk = 0;
                for (int o = 0; o < 4; ++o) rk += (o != k && o != n && gid[o] < gid[n]) ? 1 : 0;
                r[n] = rk;
            }
            const uint32_t mine = face_sig(to.nodes, i >> 2, i & 3, k, r[0], r[1], r[2], r[3]);
            const uint32_t sigs = pick4(nb_sigs[0], nb_sigs[1], nb_sigs[2], nb_sigs[3], k);
            const bool after = pick4(nb_after[0], nb_after[1], nb_after[2], nb_after[3], k);
            for (int f = 0; f < 4; ++f)
                if (mine != 0u && ((sigs >> (8 * f)) & 0xffu) == mine) { ++cnt; if (after) last = false; }
        }
        if ((cnt & 1u) == 0u && last) out |= 1u << i;
    }
    return out;
}

// A cell tile takes part when it can hold a boundary side: any tile that is not uniformly full /
// empty, and full tiles that touch a tile that is not full (or the lattice border)
__device__ __forceinline__ bool boundary_tile_candidate(const Stuff& S, uint64_t tile, uint8_t cu) {
    if (cu == CU_EMPTY) return false;
    if (cu == CU_ACTIVE) return true;
    const int tix = (int)(tile % S.g.ntx), tiy = (int)((tile / S.g.ntx) % S.g.nty), tiz = (int)(tile / ((uint64_t)S.g.ntx * S.g.nty));
    uint32_t X0, Y0, Z0;
    tile_origin(S.g, tile, X0, Y0, Z0);
    if (X0 == 0u || Y0 == 0u || Z0 == 0u || X0 + TS >= S.L.nx || Y0 + TS >= S.L.ny || Z0 + TS >= S.L.nz) return true;
    const int d[6][3] = {{-1, 0, 0}, {1, 0, 0}, {0, -1, 0}, {0, 1, 0}, {0, 0, -1}, {0, 0, 1}};
    for (int f = 0; f < 6; ++f) {
        const int x = tix + d[f][0], y = tiy + d[f][1], z = tiz + d[f][2];
        if (x < 0 || y < 0 || z < 0 || x >= (int)S.g.ntx || y >= (int)S.g.nty || z >= (int)S.g.ntz) return true;
        if (S.cuni[((size_t)z * S.g.nty + y) * S.g.ntx + x] != CU_FULL) return true;
    }
    return false;
}

// the cell tiles that can hold a boundary side -> work list (its index is the tile's slot in the mask scratch)
static __global__ void k_boundary_tiles(Stuff S, uint64_t ntiles, uint32_t* __restrict__ list, unsigned long long* __restrict__ n_list) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ntiles) return;
    uint32_t X0, Y0, Z0;
    tile_origin(S.g, t, X0, Y0, Z0);
    const uint32_t zlo = max(Z0, S.c0), zhi = min(Z0 + TS, S.c1);
    if (X0 >= S.L.nx || Y0 >= S.L.ny || zhi <= zlo || !boundary_tile_candidate(S, t, S.cuni[t])) return;
    list[atomicAdd(n_list, 1ull)] = (uint32_t)t;
}

// per-cell info of a tile's owned cells -> s_ci (0 for cells that are not owned / outside the lattice)
__device__ __forceinline__ void stage_cell_info(const Stuff& S, uint32_t tile, uint16_t* s_ci, uint32_t& X0, uint32_t& Y0, uint32_t& Z0,
                                                uint32_t& zlo, uint32_t& zhi) {
    const Lattice& L = S.L;
    tile_origin(S.g, tile, X0, Y0, Z0);
    const uint8_t cu = S.cuni[tile];
    zlo = max(Z0, S.c0); zhi = min(Z0 + TS, S.c1);
    const uint32_t nxt = min(TS, L.nx - X0);
    const uint32_t ly = threadIdx.x & 15u, lz = threadIdx.x >> 4;
    const uint32_t cy = Y0 + ly, cz = Z0 + lz;
    const bool row_owned = cy < L.ny && cz >= zlo && cz < zhi;
    const uint32_t nx_row = row_owned ? nxt : 0u;
#pragma unroll
    for (uint32_t i = 0; i < TS; ++i) {
        uint16_t ci = 0;
        if (i < nx_row) ci = cu == CU_FULL ? CI_ALL_FULL : S.cinfo[cidx(S, X0 + i, cy, cz)];
        s_ci[threadIdx.x * TS + i] = ci;
    }
}

// Pass 1, one listed tile at a time: the boundary sides of every lattice tet of the tile's owned cells as a
// 12-bit mask (bit 4 q + s), written to masks[slot][t][cell] (5 x 4096 uint16 per tile), and the tile's
// number of sides -> tile_tot[tile].  Only cells next to the surface do any work: the cells that are not
// full, and full cells with a neighbour that is not; their lattice tets are dealt to the threads one by one.
static __global__ void __launch_bounds__(TILE_THREADS)
k_boundary_masks(Stuff S, const uint32_t* __restrict__ tile_list, const unsigned long long* __restrict__ n_list,
                 unsigned long long* __restrict__ cursor, uint16_t* __restrict__ masks, unsigned long long* __restrict__ tile_tot) {
    __shared__ unsigned long long s_acc;
    __shared__ unsigned long long s_item;
    __shared__ uint16_t s_cih[BH * BH * BH], s_list[TILE_POINTS], s_heavy[8 * TILE_THREADS], s_soff[TILE_POINTS];
    __shared__ uint32_t s_n, s_nheavy;
    __shared__ uint32_t s_scan[9];
    const Lattice& L = S.L;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_item = atomicAdd(cursor, 1ull);
        __syncthreads();
        const unsigned long long slot = s_item;
        if (slot >= *n_list) break;
        const uint32_t tile = tile_list[slot];
        uint32_t X0, Y0, Z0;
        tile_origin(S.g, tile, X0, Y0, Z0);
        const uint32_t zlo = max(Z0, S.c0), zhi = min(Z0 + TS, S.c1);
        if (threadIdx.x == 0) s_n = 0u;
        // info words of the tile's cells and of the cells around it
        for (int idx = (int)threadIdx.x; idx < BH * BH * BH; idx += TILE_THREADS) {
            const int lx = idx % BH - 1, ly = (idx / BH) % BH - 1, lz = idx / (BH * BH) - 1;
            s_cih[idx] = cell_info(S, (int64_t)X0 + lx, (int64_t)Y0 + ly, (int64_t)Z0 + lz).ci;
        }
        // this tile's masks start out empty
        uint4* mz = reinterpret_cast<uint4*>(masks + slot * (5ull * TILE_POINTS));
        for (uint32_t i = threadIdx.x; i < 5u * TILE_POINTS * 2u / 16u; i += TILE_THREADS) mz[i] = make_uint4(0u, 0u, 0u, 0u);
        __syncthreads();
        // surface-cell numbers of the piece cache: the owned cells that are not full, in emission order
        // (x-rows of cells, like k_tet_emit)
        {
            const uint32_t ly = threadIdx.x & 15u, lz = threadIdx.x >> 4;
            const bool row_owned = Y0 + ly < L.ny && Z0 + lz >= zlo && Z0 + lz < zhi;
            uint32_t surf = 0;
#pragma unroll
            for (uint32_t i = 0; i < TS; ++i) {
                const uint16_t ci = s_cih[(int)((lz + 1u) * BH + (ly + 1u)) * BH + (int)(i + 1u)];
                s_soff[threadIdx.x * TS + i] = (uint16_t)surf;
                if (row_owned && X0 + i < L.nx && ci_count(ci) != 0u && !(ci & CI_FULL)) ++surf;
            }
            uint32_t stot;
            const uint32_t sex = block_exscan(surf, &stot, s_scan);
#pragma unroll
            for (uint32_t i = 0; i < TS; ++i) s_soff[threadIdx.x * TS + i] = (uint16_t)(s_soff[threadIdx.x * TS + i] + sex);
        }
        TileView V;
        V.cih = s_cih; V.soff = s_soff; V.X0 = X0; V.Y0 = Y0; V.Z0 = Z0; V.zlo = zlo; V.zhi = zhi;
        V.cached = S.pieces != nullptr && S.cuni[tile] == CU_ACTIVE;
        V.sbase = V.cached ? S.tile_sbase[tile] : 0ull;
        __syncthreads();
        // candidate cells: owned, emitting something, and not a full cell surrounded by full cells
        for (int rnd = 0; rnd < TILE_ROUNDS; ++rnd) {
            const uint32_t j = (uint32_t)rnd * TILE_THREADS + threadIdx.x;
            const uint32_t lx = j & 15u, ly = (j >> 4) & 15u, lz = j >> 8;
            if (Z0 + lz < zlo || Z0 + lz >= zhi) continue;
            const int at = (int)((lz + 1u) * BH + (ly + 1u)) * BH + (int)(lx + 1u);
            const uint16_t ci = s_cih[at];
            if (ci_count(ci) == 0u) continue;
            const bool cand = !(ci & CI_FULL) ||
                              !(s_cih[at - 1] & s_cih[at + 1] & s_cih[at - BH] & s_cih[at + BH] & s_cih[at - BH * BH] & s_cih[at + BH * BH] & CI_FULL);
            if (cand) s_list[atomicAdd(&s_n, 1u)] = (uint16_t)j;
        }
        __syncthreads();
        const uint32_t ncand = s_n;
        unsigned long long mine = 0ull;
        uint16_t* mt = masks + slot * (5ull * TILE_POINTS);
        // Two speeds: an untouched lattice tet whose neighbours are untouched / absent is settled from the
        // info words alone; trimmed tets (and untouched ones next to a trimmed one) are queued and then
        // processed densely, so that the expensive re-classification does not run in half-empty warps.
        constexpr uint32_t CHUNK = 8u * TILE_THREADS;
        for (uint32_t base = 0; base < 5u * ncand; base += CHUNK) {
            if (threadIdx.x == 0) s_nheavy = 0u;
            __syncthreads();
            for (uint32_t w = base + threadIdx.x; w < min(base + CHUNK, 5u * ncand); w += TILE_THREADS) {
                const uint32_t j = s_list[w / 5u], t = w % 5u;
                const uint32_t lx = j & 15u, ly = (j >> 4) & 15u, lz = j >> 8;
                const int at = (int)((lz + 1u) * BH + (ly + 1u)) * BH + (int)(lx + 1u);
                const int st = tet_state(s_cih[at], (int)t);
                if (st == TET_NONE) continue;
                bool heavy = st == TET_TRIM;
                if (!heavy) {
                    const uint32_t p = ((X0 + lx) & 1u) | (((Y0 + ly) & 1u) << 1) | (((Z0 + lz) & 1u) << 2);
                    const uint32_t mask = whole_tet_mask(s_cih, at, p, (int)t, &heavy);
                    if (mask) { mt[t * TILE_POINTS + j] = (uint16_t)mask; mine += (unsigned long long)__popc(mask); }
                }
                if (heavy) s_heavy[atomicAdd(&s_nheavy, 1u)] = (uint16_t)(j | (t << 12));
            }
            __syncthreads();
            const uint32_t nh = s_nheavy;
            for (uint32_t w = threadIdx.x; w < nh; w += TILE_THREADS) {
                const uint32_t j = s_heavy[w] & 0xfffu, t = s_heavy[w] >> 12;
                const uint32_t lx = j & 15u, ly = (j >> 4) & 15u, lz = j >> 8;
                const int at = (int)((lz + 1u) * BH + (ly + 1u)) * BH + (int)(lx + 1u);
                const Cell c = make_cell(X0 + lx, Y0 + ly, Z0 + lz);
                const uint32_t mask = tet_boundary_mask(S, V, at, c, (int)t);
                if (mask) { mt[t * TILE_POINTS + j] = (uint16_t)mask; mine += (unsigned long long)__popc(mask); }
            }
            __syncthreads();
        }
        const unsigned long long tot = block_sum(mine, &s_acc);
        if (threadIdx.x == 0) tile_tot[tile] = tot;
    }
}

// Pass 2: expand the masks into (elem, side) pairs in emission order (tiles, x-rows of cells, lattice tets,
// pieces, sides): the output is sorted by (elem, side).  tile_base[tile] = index of the tile's first side.
static __global__ void __launch_bounds__(TILE_THREADS)
k_boundary_emit(Stuff S, const uint32_t* __restrict__ tile_list, const unsigned long long* __restrict__ n_list,
                unsigned long long* __restrict__ cursor, const uint16_t* __restrict__ masks,
                const unsigned long long* __restrict__ tile_base, unsigned long long tet_base, unsigned long long* __restrict__ sides) {
    __shared__ unsigned long long s_item;
    __shared__ uint32_t s_scan[9];
    __shared__ uint16_t s_ci[TILE_POINTS];
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_item = atomicAdd(cursor, 1ull);
        __syncthreads();
        const unsigned long long slot = s_item;
        if (slot >= *n_list) break;
        const uint32_t tile = tile_list[slot];
        uint32_t X0, Y0, Z0, zlo, zhi;
        stage_cell_info(S, tile, s_ci, X0, Y0, Z0, zlo, zhi);   // every thread stages (and then owns) one x-row
        const uint16_t* mt = masks + slot * (5ull * TILE_POINTS);
        // tets and sides of this thread's row
        uint32_t ntet = 0, nside = 0;
#pragma unroll
        for (uint32_t i = 0; i < TS; ++i) {
            const uint32_t j = threadIdx.x * TS + i;
            ntet += ci_count(s_ci[j]);
#pragma unroll
            for (uint32_t t = 0; t < 5u; ++t) nside += (uint32_t)__popc((uint32_t)mt[t * TILE_POINTS + j]);
        }
        uint32_t tot;
        const uint32_t ex = block_exscan(ntet | (nside << 16), &tot, s_scan);  // <= 61440 tets, sides are far fewer
        if ((tot >> 16) == 0u) continue;
        uint64_t elem = tet_base + S.ttbase[tile] + (ex & 0xffffu);
        uint64_t o = tile_base[tile] + (ex >> 16);
        for (uint32_t i = 0; i < TS && nside != 0u; ++i) {
            const uint32_t j = threadIdx.x * TS + i;
            const uint16_t ci = s_ci[j];
            for (uint32_t t = 0; t < 5u; ++t) {
                const uint32_t mask = mt[t * TILE_POINTS + j];
                for (uint32_t b = 0; b < 12u && mask; ++b)
                    if ((mask >> b) & 1u) { sides[2 * o] = elem + (b >> 2); sides[2 * o + 1] = b & 3u; ++o; }
                elem += ci_tet_n(ci, (int)t);
            }
        }
    }
}

// ---- side-sets ----------------------------------------------------------------------------------
// src/main.rs:87-106: a surface side belongs to a boundary object's side-set iff all three nodes satisfy
// |approx_value(p, EPSILON)| <= EPSILON.  The object is evaluated once per surface VERTEX:
//   1. k_mark_side_nodes: flag the nodes of the surface sides (own vertices, and the slice of the rank
//      below's vertices a slab may reference)          2. k_list_flagged: gather them into a dense list
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

template <typename IndexT>
__global__ void k_mark_side_nodes(const IndexT* __restrict__ tets, unsigned long long tet_base,
                                  const unsigned long long* __restrict__ sides, uint64_t n, long long vertex_base,
                                  long long n_verts, long long lower_first, long long lower_count, uint8_t* __restrict__ flag) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const long long f = flag_index(side_node(tets, tet_base, sides, i, k), vertex_base, n_verts, lower_first, lower_count);
        if (f >= 0 && flag[f] == 0) flag[f] = 1;
    }
}

static __global__ void k_list_flagged(const uint8_t* __restrict__ flag, uint64_t n, unsigned long long* __restrict__ list,
                                      unsigned long long* __restrict__ count) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool on = i < n && flag[i] != 0;
    const unsigned b = __ballot_sync(0xffffffffu, on);
    if (b == 0u) return;
    const unsigned lane = threadIdx.x & 31u;
    const int leader = __ffs((int)b) - 1;
    unsigned long long base = 0ull;
    if ((int)lane == leader) base = atomicAdd(count, (unsigned long long)__popc(b));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (on) list[base + __popc(b & ((1u << lane) - 1u))] = i;
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
    {
        // constant memory is per device; the upload is 320 bytes
        uint16_t tab[8][5][4];
        build_nb_table(tab);
        MC_CUDA(cudaMemcpyToSymbolAsync(C_NB_TABLE, tab, sizeof(tab), 0, cudaMemcpyHostToDevice, s));
    }
    Scratch sc(c);
    MC_CUDA(cudaEventCreate(&sc.e0));
    MC_CUDA(cudaEventCreate(&sc.e1));
    MC_CUDA(cudaEventRecord(sc.e0, s));
    DevBuf ftot, fbase, dctl, dlist, dmasks;
    int rc = sc.alloc((m->n_tiles + 1) * 8, ftot);
    if (rc == MC_OK) rc = sc.alloc((m->n_tiles + 1) * 8, fbase);
    if (rc == MC_OK) rc = sc.alloc(8 * 8, dctl);   // [0..1] totals, [2] list length, [3] / [4] cursors
    if (rc == MC_OK) rc = sc.alloc(std::max<uint64_t>(m->n_tiles, 1) * 4, dlist);
    if (rc != MC_OK) return rc;
    unsigned long long* ctl = (unsigned long long*)dctl.p;
    const Stuff S = make_stuff(m);
    m->n_sides = 0;
    m->h_sides.clear();
    m->have_h_sides = false;
    c->release(m->d_sides);
    if (m->n_tiles > 0 && m->n_tets > 0) {
        MC_CUDA(cudaMemsetAsync(dctl.p, 0, 64, s));
        MC_CUDA(cudaMemsetAsync(ftot.p, 0, (m->n_tiles + 1) * 8, s));
        k_boundary_tiles<<<(unsigned)((m->n_tiles + 255) / 256), 256, 0, s>>>(S, m->n_tiles, (uint32_t*)dlist.p, ctl + 2);
        c->launches++;
        MC_CUDA(cudaGetLastError());
        unsigned long long nlist = 0;
        MC_CUDA(cudaMemcpyAsync(&nlist, ctl + 2, 8, cudaMemcpyDeviceToHost, s));
        MC_CUDA(cudaStreamSynchronize(s));
        if (nlist) {
            // 5 x 4096 uint16 masks per listed tile
            rc = sc.alloc(nlist * 5ull * TILE_POINTS * 2ull, dmasks);
            if (rc != MC_OK) return rc;
            const unsigned grid = (unsigned)std::min<uint64_t>(nlist, (uint64_t)c->sm_count * TILE_GRID_PER_SM);
            k_boundary_masks<<<grid, TILE_THREADS, 0, s>>>(S, (const uint32_t*)dlist.p, ctl + 2, ctl + 3, (uint16_t*)dmasks.p,
                                                          (unsigned long long*)ftot.p);
            c->launches++;
            MC_CUDA(cudaGetLastError());
            rc = scan_tiles(c, ftot.p, m->n_tiles, fbase.p, nullptr, m->d_scan_sums.p, ctl);
            if (rc != MC_OK) return rc;
            unsigned long long h[2];
            MC_CUDA(cudaMemcpyAsync(h, ctl, 16, cudaMemcpyDeviceToHost, s));
            MC_CUDA(cudaStreamSynchronize(s));
            m->n_sides = h[0];
            if (m->n_sides) {
                rc = c->alloc(m->n_sides * 16, m->d_sides);
                if (rc != MC_OK) return rc;
                k_boundary_emit<<<grid, TILE_THREADS, 0, s>>>(S, (const uint32_t*)dlist.p, ctl + 2, ctl + 4, (const uint16_t*)dmasks.p,
                                                             (const unsigned long long*)fbase.p, m->tet_base,
                                                             (unsigned long long*)m->d_sides.p);
                c->launches++;
                MC_CUDA(cudaGetLastError());
            }
        }
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
        rc = c->alloc(std::max<uint64_t>(nflag, 1), m->d_side_flag);
        if (rc == MC_OK) rc = c->alloc(std::max<uint64_t>(std::min<uint64_t>(3 * ns, nflag), 1) * 8 + 8, m->d_side_nodes);
        if (rc != MC_OK) return rc;
        MC_CUDA(cudaMemsetAsync(m->d_side_flag.p, 0, std::max<uint64_t>(nflag, 1), s));
        MC_CUDA(cudaMemsetAsync(m->d_side_nodes.p, 0, 8, s));
        k_mark_side_nodes<IndexT><<<nb, 256, 0, s>>>(tets, m->tet_base, sides, ns, (long long)m->vertex_base, nv, lf, lc,
                                                     (uint8_t*)m->d_side_flag.p);
        k_list_flagged<<<(unsigned)((nflag + 255) / 256), 256, 0, s>>>((const uint8_t*)m->d_side_flag.p, nflag,
                                                                      (unsigned long long*)m->d_side_nodes.p + 1,
                                                                      (unsigned long long*)m->d_side_nodes.p);
        c->launches += 2;
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
        k_mark_side_nodes<uint32_t><<<nbs, 256, 0, s>>>((const uint32_t*)m->d_tets.p, m->tet_base, sides, ns, 0, (long long)nv, 0, 0, (uint8_t*)flag.p);
    else
        k_mark_side_nodes<unsigned long long><<<nbs, 256, 0, s>>>((const unsigned long long*)m->d_tets.p, m->tet_base, sides, ns, 0, (long long)nv, 0, 0, (uint8_t*)flag.p);
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
