// Order-free device restatement of tessellate3d's isosurface stuffing:
//   warp_vertices         tessellate3d/src/lib.rs:138-181   -> k_warp_codes (mc_kernels.cuh) / warped_pos
//   first-touch vertex id tessellate3d/src/lib.rs:114-124   -> first_touch_key
//   trim_spikes           tessellate3d/src/lib.rs:242-367   -> classify_tet
//   remove_exterior_tets  tessellate3d/src/lib.rs:370-398   -> exterior test inside classify_tet
// Everything is a pure function of the lattice SDF field, so any thread can recompute any
// piece; nothing depends on a traversal order.
#pragma once
#include "mc_lattice.cuh"
#include "mc_shapes.hpp"
#include "mc_vm.cuh"

namespace mc {

__device__ __forceinline__ double raw_phi(const Lattice& L, uint32_t X, uint32_t Y, uint32_t Z) {
    return phi_at(L, X, Y, Z);
}
// phi after warp_vertices: warped vertices are snapped to 0 (tessellate3d/src/lib.rs:178).
// (A vertex of a proven tile is never warped: all its neighbours have its strict sign.)
__device__ __forceinline__ double warped_phi(const Lattice& L, uint32_t X, uint32_t Y, uint32_t Z) {
    const int8_t s = tile_sign_at(L, X, Y, Z);
    if (s) return (double)s;
    const size_t i = pidx(L, X, Y, Z);
    return (L.wcode[i] & 0x7f) ? 0.0 : L.phi[i];
}

// cut_lattice keeps a tet iff any corner has value <= 0 (tessellate3d/src/lib.rs:105-111)
__device__ __forceinline__ bool tet_kept(const Lattice& L, const Cell& c, int t) {
#pragma unroll
    for (int n = 0; n < 4; ++n) {
        uint32_t X, Y, Z;
        corner_xyz(c, C_TETS[t][n], X, Y, Z);
        if (raw_phi(L, X, Y, Z) <= 0.0) return true;
    }
    return false;
}

// Order-isomorphic stand-in for the reference's sequential vertex id: position of the vertex's
// first appearance in cut_lattice's loop (cells x-major, then tet, then node, kept tets only).
static __device__ uint64_t first_touch_key(const Lattice& L, uint32_t X, uint32_t Y, uint32_t Z) {
    for (int ax = -1; ax <= 0; ++ax) {
        const int64_t cx = (int64_t)X + ax;
        if (cx < 0 || cx >= (int64_t)L.nx) continue;
        for (int ay = -1; ay <= 0; ++ay) {
            const int64_t cy = (int64_t)Y + ay;
            if (cy < 0 || cy >= (int64_t)L.ny) continue;
            for (int az = -1; az <= 0; ++az) {
                const int64_t cz = (int64_t)Z + az;
                if (cz < 0 || cz >= (int64_t)L.nz) continue;
                const Cell c = make_cell((uint32_t)cx, (uint32_t)cy, (uint32_t)cz);
                const int k = corner_of(c, X, Y, Z);
                for (int t = 0; t < 5; ++t) {
                    int pos = -1;
#pragma unroll
                    for (int n = 0; n < 4; ++n)
                        if (C_TETS[t][n] == k) pos = n;
                    if (pos < 0) continue;
                    if (!tet_kept(L, c, t)) continue;
                    const uint64_t cell_linear = ((uint64_t)cx * L.ny + (uint64_t)cy) * L.nz + (uint64_t)cz;
                    return (cell_linear * 5u + (uint64_t)t) * 4u + (uint64_t)pos;
                }
            }
        }
    }
    return ~0ull;
}

// warp codes (k_warp_codes): 0 (not warped) or 1 + 2*dir + orient where dir indexes C_DIR (the
// neighbour the winning edge leads to) and orient says whether the vertex was node 0 (orient 0,
// `alpha < 0.3` branch) or node 1 (orient 1, `alpha > 0.7` branch) of the winning edge visit.
// position after warp_vertices (tessellate3d/src/lib.rs:163,170,177): x += delta
static __device__ void warped_pos(const Lattice& L, uint32_t X, uint32_t Y, uint32_t Z, double out[3]) {
    const double xv = lat_x(L, X), yv = lat_y(L, Y), zv = lat_z(L, Z);
    const uint32_t code = L.wcode[pidx(L, X, Y, Z)] & 0x7fu;
    if (code == 0u) { out[0] = xv; out[1] = yv; out[2] = zv; return; }
    const uint32_t di = (code - 1u) >> 1, orient = (code - 1u) & 1u;
    const uint32_t WX = X + C_DIR[di][0], WY = Y + C_DIR[di][1], WZ = Z + C_DIR[di][2];
    // a warped vertex and the far end of its winning edge differ in sign class: neither lies in a
    // proven tile, the stored values are read directly
    const double pv = L.phi[pidx(L, X, Y, Z)], pw = L.phi[pidx(L, WX, WY, WZ)];
    const double dx = lat_x(L, WX) - xv, dy = lat_y(L, WY) - yv, dz = lat_z(L, WZ) - zv;
    double f;
    if (orient == 0u) {
        f = pv / (pv - pw);             // delta = (x1 - x0) * alpha, me = node 0
    } else {
        f = 1.0 - pw / (pw - pv);       // delta = (x0 - x1) * (1 - alpha), me = node 1
    }
    out[0] = xv + dx * f; out[1] = yv + dy * f; out[2] = zv + dz * f;
}

// ---------------------------------------------------------------------------------------
// trim_spikes + remove_exterior_tets for one lattice tet.
// Output tets reference either an original node of the tet (code 0..3, index into the flipped
// node order) or the cut vertex on the edge between a positive node i and a negative node j
// (code 4 + 4*i + j).
// `pick` helpers keep the small per-tet arrays in registers (dynamic indexing would put them in
// local memory)
template <typename T>
__device__ __forceinline__ T pick4(T a, T b, T c, T d, int i) { return i == 0 ? a : (i == 1 ? b : (i == 2 ? c : d)); }

struct TetOut {
    int n;               // number of output tets (0..3)
    int kase;            // -2 not kept, -1 passed through, 0..6 trim case
    int perm;            // dense index of the sorted node order (shape_perm_index), trim cases only
    bool ext_removed;    // all-zero tet dropped by the centroid test
    unsigned long long nodes;  // 5 bits per node code, node m of output tet q at bit 5 * (4 q + m)
    __device__ __forceinline__ int node(int q, int m) const { return (int)((nodes >> (5 * (4 * q + m))) & 31ull); }
};

// symbolic a,b,c,d = 0..3; cut(x,y) = 4 + 4x + y
#define MC_CUT(x, y) (4 + 4 * (x) + (y))
static __constant__ uint8_t C_CASE_N[7] = {0, 1, 3, 3, 1, 2, 1};
static __constant__ uint8_t C_CASE_TET[7][3][5] = {
    // {n0, n1, n2, n3, toggles_flip}
    {{0, 0, 0, 0, 0}, {0, 0, 0, 0, 0}, {0, 0, 0, 0, 0}},
    {{MC_CUT(0, 3), MC_CUT(1, 3), MC_CUT(2, 3), 3, 0}, {0, 0, 0, 0, 0}, {0, 0, 0, 0, 0}},
    {{MC_CUT(0, 1), 1, 2, 3, 0}, {2, MC_CUT(0, 1), MC_CUT(0, 2), 3, 1}, {MC_CUT(0, 1), 3, MC_CUT(0, 2), MC_CUT(0, 3), 0}},
    {{MC_CUT(0, 2), MC_CUT(1, 3), 2, 3, 0}, {MC_CUT(0, 2), MC_CUT(1, 3), MC_CUT(1, 2), 3, 0}, {MC_CUT(0, 2), MC_CUT(0, 3), MC_CUT(1, 3), 3, 0}},
    {{MC_CUT(0, 3), 1, 2, 3, 0}, {0, 0, 0, 0, 0}, {0, 0, 0, 0, 0}},
    {{1, MC_CUT(0, 2), 2, 3, 1}, {MC_CUT(0, 3), 1, MC_CUT(0, 2), 3, 0}, {0, 0, 0, 0, 0}},
    {{MC_CUT(0, 3), MC_CUT(1, 3), 2, 3, 0}, {0, 0, 0, 0, 0}, {0, 0, 0, 0, 0}},
};

struct TetNodes {
    uint32_t X[4], Y[4], Z[4];  // lattice coordinates of the flipped node order
    double p[4];                // post-warp phi
};
__device__ __forceinline__ uint32_t tet_X(const TetNodes& t, int i) { return pick4(t.X[0], t.X[1], t.X[2], t.X[3], i); }
__device__ __forceinline__ uint32_t tet_Y(const TetNodes& t, int i) { return pick4(t.Y[0], t.Y[1], t.Y[2], t.Y[3], i); }
__device__ __forceinline__ uint32_t tet_Z(const TetNodes& t, int i) { return pick4(t.Z[0], t.Z[1], t.Z[2], t.Z[3], i); }
__device__ __forceinline__ double tet_p(const TetNodes& t, int i) { return pick4(t.p[0], t.p[1], t.p[2], t.p[3], i); }

__device__ __forceinline__ void load_tet(const Lattice& L, const Cell& c, int t, TetNodes& tn) {
#pragma unroll
    for (int n = 0; n < 4; ++n) {
        corner_xyz(c, tet_corner(c, t, n), tn.X[n], tn.Y[n], tn.Z[n]);
        tn.p[n] = warped_phi(L, tn.X[n], tn.Y[n], tn.Z[n]);
    }
}

// the same for a tet that is known to have a vertex of each side of zero (a trimmed tet): none of its
// vertices lies in a proven tile, the stored values are read directly
__device__ __forceinline__ void load_tet_mixed(const Lattice& L, const Cell& c, int t, TetNodes& tn) {
#pragma unroll
    for (int n = 0; n < 4; ++n) {
        corner_xyz(c, tet_corner(c, t, n), tn.X[n], tn.Y[n], tn.Z[n]);
        const size_t i = pidx(L, tn.X[n], tn.Y[n], tn.Z[n]);
        tn.p[n] = (L.wcode[i] & 0x7f) ? 0.0 : L.phi[i];
    }
}

// The 4-sort of trim_spikes on its own (tessellate3d/src/lib.rs:262-282) for a tet with a positive vertex:
// trim case and the dense index of the sorted node order (mc_shapes.hpp).  The values travel with the
// labels through the comparator network, so nothing is indexed dynamically; `tie(i, j)` is only asked for
// two equal non-zero values (first-touch order of nodes i and j, like the reference's vertex ids).
template <typename Tie>
__device__ __forceinline__ void sort_case(double p0, double p1, double p2, double p3, Tie tie, int& kase, int& perm) {
    int ia = 0, ib = 1, ic = 2, id = 3;
    double pa = p0, pb = p1, pc = p2, pd = p3;
#define MC_CSWAP2(iu, iv, pu, pv)                                             \
    if (pu < pv || (pu == pv && pu != 0.0 && tie(iu, iv))) {                  \
        const int ti_ = iu; iu = iv; iv = ti_;                                \
        const double tp_ = pu; pu = pv; pv = tp_;                             \
    }
    MC_CSWAP2(ia, ib, pa, pb)
    MC_CSWAP2(ic, id, pc, pd)
    MC_CSWAP2(ia, ic, pa, pc)
    MC_CSWAP2(ib, id, pb, pd)
    MC_CSWAP2(ib, ic, pb, pc)
#undef MC_CSWAP2
    if (pd == 0.0) kase = 0;
    else if (pc > 0.0) kase = 1;
    else if (pb < 0.0) kase = 2;
    else if (pb > 0.0 && pc < 0.0) kase = 3;
    else if (pb == 0.0 && pc == 0.0) kase = 4;
    else if (pb == 0.0) kase = 5;
    else kase = 6;
    perm = shape_perm_index(ia, ib, ic);
}

// `removed_hint`: < 0 -> evaluate the exterior test with the SDF program; 0/1 -> cached result
// KNOWN_KEPT: the caller knows that cut_lattice kept the tet (it produced output)
template <bool CAN_EVAL, bool KNOWN_KEPT = false>
__device__ void classify_tet(const Lattice& L, const Cell& c, int t, const TetNodes& tn,
                             const uint64_t* __restrict__ prog, int removed_hint, TetOut& out) {
    out.n = 0;
    out.kase = -2;
    out.ext_removed = false;
    out.perm = 0;
    out.nodes = 0ull;
    if (!KNOWN_KEPT && !tet_kept(L, c, t)) return;
    const double* p = tn.p;
    if (p[0] <= 0.0 && p[1] <= 0.0 && p[2] <= 0.0 && p[3] <= 0.0) {
        out.kase = -1;
        if (p[0] == 0.0 && p[1] == 0.0 && p[2] == 0.0 && p[3] == 0.0) {
            bool removed;
            if (!CAN_EVAL || removed_hint >= 0) {
                removed = removed_hint > 0;
            } else if (CAN_EVAL) {
                // centroid = (((a + b) + c) + d) / 4 of the warped positions
                double a[3], b[3], cc[3], d[3];
                warped_pos(L, tn.X[0], tn.Y[0], tn.Z[0], a);
                warped_pos(L, tn.X[1], tn.Y[1], tn.Z[1], b);
                warped_pos(L, tn.X[2], tn.Y[2], tn.Z[2], cc);
                warped_pos(L, tn.X[3], tn.Y[3], tn.Z[3], d);
                const double qx = (((a[0] + b[0]) + cc[0]) + d[0]) / 4.0;
                const double qy = (((a[1] + b[1]) + cc[1]) + d[1]) / 4.0;
                const double qz = (((a[2] + b[2]) + cc[2]) + d[2]) / 4.0;
                removed = vm_eval<false>(prog, qx, qy, qz) > 0.0;
            }
            if (removed) { out.ext_removed = true; return; }
        }
        out.n = 1;
        out.nodes = 0ull | (1ull << 5) | (2ull << 10) | (3ull << 15);
        return;
    }
    // 5-comparator network, descending by (phi, first-touch order); tessellate3d/src/lib.rs:262-282
    int ia = 0, ib = 1, ic = 2, id = 3;
    bool flipped = false;
    // Ties between two exact zeros (warped vertices) are left in place: they only occur as the
    // (b, c) pair of case 4 or inside case 0, where swapping the labels together with `flipped`
    // yields the same oriented tet, so the reference's vertex-id tie-break cannot matter there.
    // Ties between equal non-zero values (symmetric scenes) do change the split and use the
    // first-touch order like the reference.
#define MC_LESS(i, j)                                                                               \
    (tet_p(tn, i) < tet_p(tn, j) ||                                                                 \
     (tet_p(tn, i) == tet_p(tn, j) && tet_p(tn, i) != 0.0 &&                                        \
      first_touch_key(L, tet_X(tn, i), tet_Y(tn, i), tet_Z(tn, i)) <                                \
          first_touch_key(L, tet_X(tn, j), tet_Y(tn, j), tet_Z(tn, j))))
#define MC_CSWAP(u, v) if (MC_LESS(u, v)) { const int tmp_ = u; u = v; v = tmp_; flipped = !flipped; }
    MC_CSWAP(ia, ib)
    MC_CSWAP(ic, id)
    MC_CSWAP(ia, ic)
    MC_CSWAP(ib, id)
    MC_CSWAP(ib, ic)
#undef MC_CSWAP
#undef MC_LESS
    const double pb = tet_p(tn, ib), pc = tet_p(tn, ic), pd = tet_p(tn, id);
    int kase;
    if (pd == 0.0) kase = 0;
    else if (pc > 0.0) kase = 1;
    else if (pb < 0.0) kase = 2;
    else if (pb > 0.0 && pc < 0.0) kase = 3;
    else if (pb == 0.0 && pc == 0.0) kase = 4;
    else if (pb == 0.0) kase = 5;
    else kase = 6;
    out.kase = kase;
    out.perm = shape_perm_index(ia, ib, ic);
    out.n = C_CASE_N[kase];
    unsigned long long nodes = 0ull;
    for (int q = 0; q < out.n; ++q) {
        int nd[4];
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            const int sy = C_CASE_TET[kase][q][m];
            nd[m] = (sy < 4) ? pick4(ia, ib, ic, id, sy)
                             : (4 + 4 * pick4(ia, ib, ic, id, (sy - 4) >> 2) + pick4(ia, ib, ic, id, (sy - 4) & 3));
        }
        const bool f = flipped != (C_CASE_TET[kase][q][4] != 0);
        if (f) { const int tmp = nd[0]; nd[0] = nd[1]; nd[1] = tmp; }
#pragma unroll
        for (int m = 0; m < 4; ++m) nodes |= (unsigned long long)nd[m] << (5 * (4 * q + m));
    }
    out.nodes = nodes;
}

}  // namespace mc
