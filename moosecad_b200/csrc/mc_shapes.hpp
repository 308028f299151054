// Shape bytes: what trim_spikes (tessellate3d/src/lib.rs:242-367) made of one lattice tet, as ONE byte.
//
// The output of trim_spikes for a kept lattice tet only depends on the trim case (1..6) and on the
// order the 4-sort puts the four nodes in (24 permutations; `flipped` is the permutation's parity):
//     0                              nothing came out (not kept, case 0, dropped by the exterior test)
//     1                              the lattice tet itself (passed through untouched)
//     2 + 24 (case - 1) + perm       trimmed; perm = dense index of (ia, ib, ic, id), see shape_perm_index
// k_classify_cells records the byte per lattice tet (five per cell, Stuff::cshape); the tet emission
// turns it back into node codes with a table instead of reading the field and sorting again, and
// Mesh::boundary (mesh/src/lib.rs:341-357) becomes table look-ups: everything the face-parity rule needs
// to know about a lattice tet and the lattice tet across one of its faces is a function of
// (shape, mirroring class of the cell, tet number, face) — see mc_boundary.cu.
//
// The tables are built on the host by brute force from the same symbolic case table the classification
// uses (C_CASE_TET) and from Tile::LATTICE / Tile::TETS, once per context.
#pragma once
#include <stdint.h>
#include <string.h>

namespace mc {

constexpr int SHAPE_NONE = 0, SHAPE_WHOLE = 1, SHAPE_TRIM0 = 2, N_SHAPES = 2 + 6 * 24;

// dense index of the permutation (ia, ib, ic, id) of 0..3
#if defined(__CUDACC__)
__host__ __device__
#endif
inline int shape_perm_index(int ia, int ib, int ic) {
    const int rb = ib - (ib > ia ? 1 : 0);
    const int rc = ic - (ic > ia ? 1 : 0) - (ic > ib ? 1 : 0);
    return ia * 6 + rb * 2 + rc;
}

// geo[p][t][k]: the lattice tet across lattice face k (opposite node k of the flipped node order) of tet t
// of a cell with mirroring class p = fx | fy << 1 | fz << 2
constexpr uint32_t GEO_DX_SHIFT = 0, GEO_DY_SHIFT = 2, GEO_DZ_SHIFT = 4;  // offset + 1 of the neighbour's cell
constexpr uint32_t GEO_T2_SHIFT = 6;      // 3 bits: its tet number
constexpr uint32_t GEO_K2_SHIFT = 9;      // 2 bits: the shared face seen from it
constexpr uint32_t GEO_PME_SHIFT = 11;    // 3 bits: rank permutation of my three face nodes (index into S3)
constexpr uint32_t GEO_PNB_SHIFT = 14;    // 3 bits: rank permutation of its three face nodes
constexpr uint32_t GEO_AFTER = 1u << 17;  // the neighbour comes after me in the fixed order of the lattice tets
constexpr uint32_t GEO_VALID = 1u << 18;

struct ShapeTables {
    unsigned long long nodes[N_SHAPES];  // TetOut::nodes | number of pieces << 60
    // piece-face i = 4 q + s (piece q, side s).  own: bits 0..11 = the face occurs an odd number of times among
    // the OTHER faces of the same lattice tet's pieces, bits 12..23 = it is not the last of its occurrences there,
    // bits 24..27 = 4 x number of pieces
    uint32_t own[N_SHAPES];
    uint16_t pfk[N_SHAPES][4];        // piece-faces lying in lattice face k
    uint8_t pfsig[N_SHAPES][6][12];   // signature of piece-face i for rank permutation pi of its lattice face's nodes
    uint32_t fsig[N_SHAPES][4][6];    // the (<= 4) signatures of the piece-faces in lattice face k, 8 bits each
    uint32_t geo[8][5][4];
    uint8_t used[N_SHAPES];           // bit n: original node n (flipped node order) occurs in some piece
};

namespace shapes_detail {

constexpr int FACE[4][3] = {{0, 1, 2}, {0, 2, 3}, {0, 3, 1}, {1, 3, 2}};  // Tet4::face, mesh/src/lib.rs:661-668
constexpr int S3[6][3] = {{0, 1, 2}, {0, 2, 1}, {1, 0, 2}, {1, 2, 0}, {2, 0, 1}, {2, 1, 0}};
inline int cut(int x, int y) { return 4 + 4 * x + y; }

inline int code_of(unsigned long long nodes, int q, int m) { return (int)((nodes >> (5 * (4 * q + m))) & 31ull); }

// node-code set of face (q, s) and the lattice faces (bit k) all of its nodes lie in
inline void face_codes(unsigned long long nodes, int q, int s, uint32_t& codes, uint32_t& fmask) {
    codes = 0u;
    fmask = 0xfu;
    for (int m = 0; m < 3; ++m) {
        const int code = code_of(nodes, q, FACE[s][m]);
        codes |= 1u << code;
        fmask &= code < 4 ? ~(1u << code) : ~((1u << ((code - 4) >> 2)) | (1u << ((code - 4) & 3)));
    }
    fmask &= 0xfu;
}

// signature of face (q, s) inside lattice face k: every node becomes the set of ranks (among the three
// lattice points of face k) it hangs on — one for an original node, two for a cut vertex — and the face the
// set of its three node sets.  0 when the face does not lie in lattice face k.
inline uint32_t face_sig(unsigned long long nodes, int q, int s, int k, const int r[4]) {
    uint32_t sig = 0u;
    for (int m = 0; m < 3; ++m) {
        const int code = code_of(nodes, q, FACE[s][m]);
        uint32_t nm;
        if (code < 4) {
            if (code == k) return 0u;
            nm = 1u << r[code];
        } else {
            const int i = (code - 4) >> 2, j = (code - 4) & 3;
            if (i == k || j == k) return 0u;
            nm = (1u << r[i]) | (1u << r[j]);
        }
        sig |= 1u << nm;
    }
    return sig;
}

}  // namespace shapes_detail

// returns false when an internal consistency check fails (more than four sub-faces in a lattice face, ...)
inline bool build_shape_tables(ShapeTables& T) {
    using namespace shapes_detail;
    memset(&T, 0, sizeof(T));
    // symbolic a, b, c, d = 0..3 of the sorted order; {n0, n1, n2, n3, toggles_flip}: the same table as C_CASE_TET
    // (mc_stuffing.cuh), cases 1..6
    static const int CASE_N[7] = {0, 1, 3, 3, 1, 2, 1};
    const int CASE_TET[7][3][5] = {
        {{0, 0, 0, 0, 0}, {0, 0, 0, 0, 0}, {0, 0, 0, 0, 0}},
        {{cut(0, 3), cut(1, 3), cut(2, 3), 3, 0}, {0, 0, 0, 0, 0}, {0, 0, 0, 0, 0}},
        {{cut(0, 1), 1, 2, 3, 0}, {2, cut(0, 1), cut(0, 2), 3, 1}, {cut(0, 1), 3, cut(0, 2), cut(0, 3), 0}},
        {{cut(0, 2), cut(1, 3), 2, 3, 0}, {cut(0, 2), cut(1, 3), cut(1, 2), 3, 0}, {cut(0, 2), cut(0, 3), cut(1, 3), 3, 0}},
        {{cut(0, 3), 1, 2, 3, 0}, {0, 0, 0, 0, 0}, {0, 0, 0, 0, 0}},
        {{1, cut(0, 2), 2, 3, 1}, {cut(0, 3), 1, cut(0, 2), 3, 0}, {0, 0, 0, 0, 0}},
        {{cut(0, 3), cut(1, 3), 2, 3, 0}, {0, 0, 0, 0, 0}, {0, 0, 0, 0, 0}},
    };
    T.nodes[SHAPE_NONE] = 0ull;
    T.nodes[SHAPE_WHOLE] = (0ull | (1ull << 5) | (2ull << 10) | (3ull << 15)) | (1ull << 60);
    for (int kase = 1; kase <= 6; ++kase)
        for (int ia = 0; ia < 4; ++ia)
            for (int ib = 0; ib < 4; ++ib)
                for (int ic = 0; ic < 4; ++ic)
                    for (int id = 0; id < 4; ++id) {
                        if (ia == ib || ia == ic || ia == id || ib == ic || ib == id || ic == id) continue;
                        const int pm[4] = {ia, ib, ic, id};
                        int inv = 0;
                        for (int x = 0; x < 4; ++x)
                            for (int y = x + 1; y < 4; ++y) inv += pm[x] > pm[y] ? 1 : 0;
                        const bool flipped = (inv & 1) != 0;  // every swap of the sorting network toggles it
                        unsigned long long nodes = 0ull;
                        for (int q = 0; q < CASE_N[kase]; ++q) {
                            int nd[4];
                            for (int m = 0; m < 4; ++m) {
                                const int sy = CASE_TET[kase][q][m];
                                nd[m] = sy < 4 ? pm[sy] : 4 + 4 * pm[(sy - 4) >> 2] + pm[(sy - 4) & 3];
                            }
                            if (flipped != (CASE_TET[kase][q][4] != 0)) { const int tmp = nd[0]; nd[0] = nd[1]; nd[1] = tmp; }
                            for (int m = 0; m < 4; ++m) nodes |= (unsigned long long)nd[m] << (5 * (4 * q + m));
                        }
                        const int shape = SHAPE_TRIM0 + 24 * (kase - 1) + shape_perm_index(ia, ib, ic);
                        if (shape >= N_SHAPES || T.nodes[shape] != 0ull) return false;
                        T.nodes[shape] = nodes | ((unsigned long long)CASE_N[kase] << 60);
                    }
    for (int A = 1; A < N_SHAPES; ++A) {
        const unsigned long long nodes = T.nodes[A] & 0x0fffffffffffffffull;
        const int n = (int)(T.nodes[A] >> 60);
        if (n < 1 || n > 3) return false;
        uint32_t fc[12], fm[12];
        for (int i = 0; i < 4 * n; ++i) face_codes(nodes, i >> 2, i & 3, fc[i], fm[i]);
        uint32_t odd = 0u, notlast = 0u;
        for (int i = 0; i < 4 * n; ++i) {
            int cnt = 0;
            for (int j = 0; j < 4 * n; ++j)
                if (j != i && fc[j] == fc[i]) { ++cnt; if (j > i) notlast |= 1u << i; }
            if (cnt & 1) odd |= 1u << i;
            if (fm[i]) {
                if (fm[i] & (fm[i] - 1u)) return false;  // a face lies in at most one lattice face
                int k = 0;
                while (!((fm[i] >> k) & 1u)) ++k;
                T.pfk[A][k] = (uint16_t)(T.pfk[A][k] | (1u << i));
            }
        }
        T.own[A] = odd | (notlast << 12) | ((uint32_t)(4 * n) << 24);
        for (int q = 0; q < n; ++q)
            for (int m = 0; m < 4; ++m)
                if (code_of(nodes, q, m) < 4) T.used[A] = (uint8_t)(T.used[A] | (1u << code_of(nodes, q, m)));
        for (int k = 0; k < 4; ++k)
            for (int pi = 0; pi < 6; ++pi) {
                int r[4] = {0, 0, 0, 0}, pos = 0;
                for (int node = 0; node < 4; ++node)
                    if (node != k) r[node] = S3[pi][pos++];
                uint32_t sigs = 0u;
                int nsig = 0;
                for (int i = 0; i < 4 * n; ++i) {
                    if (!((T.pfk[A][k] >> i) & 1u)) continue;
                    const uint32_t sg = face_sig(nodes, i >> 2, i & 3, k, r);
                    if (sg == 0u || sg > 0xffu || nsig >= 4) return false;
                    T.pfsig[A][pi][i] = (uint8_t)sg;
                    sigs |= sg << (8 * nsig);
                    ++nsig;
                }
                T.fsig[A][k][pi] = sigs;
            }
    }
    // geometry: Tile::LATTICE / Tile::TETS (tessellate3d/src/lib.rs:434-454) with the per-axis mirroring of odd
    // cells and the node 0 <-> 1 flip (:59-74,126-128)
    static const int LAT[8][3] = {{0, 0, 0}, {1, 0, 0}, {1, 1, 0}, {0, 1, 0}, {0, 0, 1}, {1, 0, 1}, {1, 1, 1}, {0, 1, 1}};
    static const int TETS[5][4] = {{0, 1, 5, 2}, {0, 2, 5, 7}, {0, 7, 5, 4}, {2, 6, 5, 7}, {0, 2, 7, 3}};
    auto node = [&](const int par[3], const int off[3], int t, int n, int out[3]) {
        bool mir[3];
        for (int a = 0; a < 3; ++a) mir[a] = ((par[a] + off[a]) & 1) != 0;
        const bool flip = mir[0] ^ mir[1] ^ mir[2];
        const int c = TETS[t][(flip && n < 2) ? 1 - n : n];
        for (int a = 0; a < 3; ++a) out[a] = off[a] + (mir[a] ? 1 - LAT[c][a] : LAT[c][a]);
    };
    auto before = [](const int a[3], const int b[3]) {  // order of the point ids: z, then y, then x
        if (a[2] != b[2]) return a[2] < b[2];
        if (a[1] != b[1]) return a[1] < b[1];
        return a[0] < b[0];
    };
    auto perm_index = [&](const int pts[3][3]) {  // rank permutation of three points (position -> rank)
        int rk[3];
        for (int x = 0; x < 3; ++x) {
            rk[x] = 0;
            for (int y = 0; y < 3; ++y) rk[x] += (y != x && before(pts[y], pts[x])) ? 1 : 0;
        }
        for (int pi = 0; pi < 6; ++pi)
            if (S3[pi][0] == rk[0] && S3[pi][1] == rk[1] && S3[pi][2] == rk[2]) return pi;
        return -1;
    };
    for (int p = 0; p < 8; ++p) {
        const int par[3] = {p & 1, (p >> 1) & 1, (p >> 2) & 1};
        const int zero[3] = {0, 0, 0};
        for (int t = 0; t < 5; ++t)
            for (int k = 0; k < 4; ++k) {
                int face[3][3], w = 0;
                for (int n = 0; n < 4; ++n)
                    if (n != k) node(par, zero, t, n, face[w++]);
                const int pme = perm_index(face);
                int found = 0;
                for (int dx = -1; dx <= 1; ++dx)
                    for (int dy = -1; dy <= 1; ++dy)
                        for (int dz = -1; dz <= 1; ++dz) {
                            if ((dx != 0) + (dy != 0) + (dz != 0) > 1) continue;
                            const int off[3] = {dx, dy, dz};
                            for (int t2 = 0; t2 < 5; ++t2) {
                                if (dx == 0 && dy == 0 && dz == 0 && t2 == t) continue;
                                int hit = 0, k2 = -1, f2[3][3], w2 = 0;
                                for (int n = 0; n < 4; ++n) {
                                    int q[3];
                                    node(par, off, t2, n, q);
                                    bool on = false;
                                    for (int f = 0; f < 3; ++f) on = on || (q[0] == face[f][0] && q[1] == face[f][1] && q[2] == face[f][2]);
                                    if (on) { ++hit; if (w2 < 3) { f2[w2][0] = q[0]; f2[w2][1] = q[1]; f2[w2][2] = q[2]; ++w2; } }
                                    else k2 = n;
                                }
                                if (hit != 3 || k2 < 0) continue;
                                const int pnb = perm_index(f2);
                                if (pme < 0 || pnb < 0) return false;
                                // tet_after: cells by (z, y, x), then the tet number
                                const bool after = dz != 0 ? dz > 0 : (dy != 0 ? dy > 0 : (dx != 0 ? dx > 0 : t2 > t));
                                T.geo[p][t][k] = ((uint32_t)(dx + 1) << GEO_DX_SHIFT) | ((uint32_t)(dy + 1) << GEO_DY_SHIFT) |
                                                 ((uint32_t)(dz + 1) << GEO_DZ_SHIFT) | ((uint32_t)t2 << GEO_T2_SHIFT) |
                                                 ((uint32_t)k2 << GEO_K2_SHIFT) | ((uint32_t)pme << GEO_PME_SHIFT) |
                                                 ((uint32_t)pnb << GEO_PNB_SHIFT) | (after ? GEO_AFTER : 0u) | GEO_VALID;
                                ++found;
                            }
                        }
                if (found != 1) return false;  // every lattice face has exactly one tet on the other side
            }
    }
    return true;
}

}  // namespace mc
