// Tile specialisation of the SDF program.
//
// The reference evaluates the whole CSG tree at every lattice point; on a 1024^3 lattice with
// a 64-primitive scene that is ~10^13 operations although, around any given point, only a
// handful of primitives can influence the value.  Here the lattice is cut into tiles; for
// every tile an INTERVAL evaluation of the program (lanes = tiles, the program counter is
// warp-uniform) decides, rigorously,
//   * for every bbox guard: passes at every point of the tile / fails at every point / mixed,
//   * for every child of a smooth boolean: whether it can influence the result anywhere in
//     the tile.  A child x_i of rvmin can be dropped when x_i >= m + max(r, exact_range) at
//     every point (m = the minimum): it then never changes the running minimum, never sets
//     the `close` flag and is excluded from the exponential sum, so rvmin returns the same
//     bits with or without it (mirror image for rvmax).
// The evaluation kernel then compacts, per tile, the program into shared memory (dead
// subtrees removed, always-pass guards dropped, always-fail guards replaced by a plain bbox
// distance, single-child booleans elided, jump targets re-based) and runs the ordinary point
// evaluator on it.  Values are bit-identical to evaluating the full program; only work that
// provably cannot matter is skipped.
//
// Interval bounds are computed in floating point and then widened by PAD (1e-9 relative +
// absolute), many orders of magnitude above the rounding error of the few dozen operations
// involved, and every decision keeps that margin; NaN bounds fall through to "keep".
#pragma once
#include "mc_kernels.cuh"

namespace mc {

struct Ival { double lo, hi; };

__device__ __forceinline__ double pad_of(double v) { return 1e-9 * (1.0 + fabs(v)); }
__device__ __forceinline__ Ival widen(Ival a) { return {a.lo - pad_of(a.lo), a.hi + pad_of(a.hi)}; }
__device__ __forceinline__ Ival iv_scale(double m, double q0, double q1) {
    const double a = m * q0, b = m * q1;
    return {fmin(a, b), fmax(a, b)};
}
// interval of |q| over [q0, q1]
__device__ __forceinline__ Ival iv_abs(double q0, double q1) {
    const double a = fabs(q0), b = fabs(q1);
    if (q0 <= 0.0 && q1 >= 0.0) return {0.0, fmax(a, b)};
    return {fmin(a, b), fmax(a, b)};
}
// interval of bbox.distance over a box
__device__ __forceinline__ Ival iv_box_distance(const uint64_t* __restrict__ w, int at, const double* b) {
    double lo = -__longlong_as_double(0x7ff0000000000000LL), hi = lo;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const double blo = ld_f64(w, at + k), bhi = ld_f64(w, at + 3 + k);
        const double q0 = b[2 * k], q1 = b[2 * k + 1];
        hi = fmax(hi, fmax(q1 - bhi, blo - q0));
        lo = fmax(lo, fmax(q0 - bhi, blo - q1));
    }
    return widen({lo, hi});
}

// Interval evaluation of the program for one tile (this lane); decisions are written to
// dec[id * dec_stride].  All 32 lanes of the warp run in lock step.
// *work (optional) receives a rough count of the FP64 operations one point evaluation of the
// tile-specialised program costs (the load estimate of mc_tile_layer_classes).
// pdec (optional, then dec must be null): decisions already taken for a box that CONTAINS box0 (all
// lanes of the warp share them): failed guards and dropped children are skipped outright, passing
// guards are not re-tested — the evaluation costs what the tile-specialised program costs.
__device__ __noinline__ Ival vm_decide(const uint64_t* __restrict__ w, const double box0[6], uint8_t* __restrict__ dec,
                          size_t dec_stride, float* work = nullptr, const uint8_t* __restrict__ pdec = nullptr,
                          size_t pstride = 0) {
#define MC_PD(id) (pdec ? pdec[(size_t)(id) * pstride] : (uint8_t)MC_DEC_KEEP)
    float wk = 0.0f;
    int dead_end = -1;  // this lane's tile is inside a failed guard up to this word
    Ival vs[MC_MAX_VALUE_DEPTH + 1];
    double bs[6 * MC_MAX_POINT_DEPTH];
    double b[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) b[k] = box0[k];
    int sp = 0, bsp = 0, pc = 0;
    Ival top = {0.0, 0.0};
    const double INF = __longlong_as_double(0x7ff0000000000000LL);
#define MC_IPUSH(v) do { vs[sp] = top; ++sp; top = (v); } while (0)
    for (;;) {
        const uint64_t h = w[pc];
        const uint32_t op = MC_HDR_OP(h);
        if (op == MC_OP_END) { if (work) *work = wk; return top; }
        const bool alive = pc >= dead_end;
        switch (op) {
        case MC_OP_PLANE: {
            if (alive) wk += 3.0f;
            const unsigned sm = MC_HDR_SMALL(h);
            const unsigned axis = sm & 3u;
            const double d = ld_f64(w, pc + 1);
            const double q0 = b[2 * axis], q1 = b[2 * axis + 1];
            Ival v;
            if (sm & 4u) v = {-q1 - d, -q0 - d}; else v = {q0 - d, q1 - d};
            MC_IPUSH(widen(v));
            break;
        }
        case MC_OP_NPLANE: {
            if (alive) wk += 6.0f;
            Ival s = {-ld_f64(w, pc + 4), -ld_f64(w, pc + 4)};
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const Ival t = iv_scale(ld_f64(w, pc + 1 + k), b[2 * k], b[2 * k + 1]);
                s.lo += t.lo; s.hi += t.hi;
            }
            MC_IPUSH(widen(s));
            break;
        }
        case MC_OP_USQ: {
            if (alive) wk += 6.0f;
            Ival s = {-1.0, -1.0};
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const Ival a = iv_abs(b[2 * k], b[2 * k + 1]);
                s.lo += a.lo * a.lo; s.hi += a.hi * a.hi;
            }
            MC_IPUSH(widen(s));
            break;
        }
        case MC_OP_CONE: {
            if (alive) wk += 12.0f;
            const Ival ax = iv_abs(b[0], b[1]), ay = iv_abs(b[2], b[3]);
            const Ival rho = {sqrt(ax.lo * ax.lo + ay.lo * ay.lo), sqrt(ax.hi * ax.hi + ay.hi * ay.hi)};
            const double off = ld_f64(w, pc + 2);
            const Ival u = iv_scale(ld_f64(w, pc + 1), b[4] + off, b[5] + off);
            const Ival rad = iv_abs(u.lo, u.hi);
            const Ival v = iv_scale(ld_f64(w, pc + 3), rho.lo - rad.hi, rho.hi - rad.lo);
            MC_IPUSH(widen(widen(v)));
            break;
        }
        case MC_OP_GSPHERE:
        case MC_OP_GCYL:
        case MC_OP_SPHERE:
        case MC_OP_CYL: {
            // SPHERE / CYL: the guard-free forms the tile compaction leaves behind (radius at pc + 1)
            const bool guarded = op == MC_OP_GSPHERE || op == MC_OP_GCYL;
            const double r = ld_f64(w, pc + (guarded ? 8 : 1));
            const Ival ax = iv_abs(b[0], b[1]), ay = iv_abs(b[2], b[3]), az = iv_abs(b[4], b[5]);
            Ival e;
            if (op == MC_OP_GSPHERE || op == MC_OP_SPHERE)
                e = {sqrt((ax.lo * ax.lo + ay.lo * ay.lo) + az.lo * az.lo) - r,
                     sqrt((ax.hi * ax.hi + ay.hi * ay.hi) + az.hi * az.hi) - r};
            else
                e = {sqrt(ax.lo * ax.lo + ay.lo * ay.lo) - r, sqrt(ax.hi * ax.hi + ay.hi * ay.hi) - r};
            e = widen(e);
            if (!guarded) { if (alive) wk += 8.0f; MC_IPUSH(e); break; }
            const uint8_t pd = MC_PD(MC_HDR_DEC(h));
            if (pd == MC_DEC_PASS) { MC_IPUSH(e); break; }
            const Ival a = iv_box_distance(w, pc + 1, b);
            if (pd == MC_DEC_FAIL) { MC_IPUSH(a); break; }
            const double slack = ld_f64(w, pc + 7);
            uint8_t d = MC_DEC_KEEP;
            Ival v = {fmin(a.lo, e.lo), fmax(a.hi, e.hi)};
            if (a.hi <= slack) { d = MC_DEC_PASS; v = e; }
            else if (a.lo > slack) { d = MC_DEC_FAIL; v = a; }
            if (dec) dec[(size_t)MC_HDR_DEC(h) * dec_stride] = d;
            if (alive) wk += d == MC_DEC_FAIL ? 12.0f : 20.0f;
            MC_IPUSH(v);
            break;
        }
        case MC_OP_BOXDIST:
            if (alive) wk += 12.0f;
            MC_IPUSH(iv_box_distance(w, pc + 1, b));
            break;
        case MC_OP_GBOX:
        case MC_OP_BOX: {
            // BOX: the guard-free form of the tile compaction (planes at pc + 1.., r, er at pc + 7, 8).
            // `small` = planes that take part (all six in an uncompacted program).
            const bool guarded = op == MC_OP_GBOX;
            const int pl = pc + (guarded ? 8 : 1);
            unsigned mask = MC_HDR_SMALL(h);
            uint8_t pd = MC_DEC_KEEP;
            if (guarded && pdec) {
                pd = MC_PD(MC_HDR_DEC(h));
                if (pd == MC_DEC_FAIL) { MC_IPUSH(iv_box_distance(w, pc + 1, b)); break; }
#pragma unroll
                for (int i = 0; i < 6; ++i)
                    if (MC_PD(MC_HDR_DEC(h) + 1u + (uint32_t)i) == MC_DEC_FAIL) mask &= ~(1u << i);
            }
            // the six plane intervals; a plane is droppable like any rvmax child
            const double r = ld_f64(w, pl + 6), er = ld_f64(w, pl + 7);
            const double margin = fmax(r, er);
            Ival pv[6];
            double ext = -INF;
#pragma unroll
            for (int i = 0; i < 6; ++i) {
                const double dd = ld_f64(w, pl + i);
                const double q0 = b[2 * (i % 3)], q1 = b[2 * (i % 3) + 1];
                pv[i] = widen(i >= 3 ? Ival{-q1 - dd, -q0 - dd} : Ival{q0 - dd, q1 - dd});
                if ((mask >> i) & 1u) ext = fmax(ext, pv[i].lo);
            }
            Ival e = {-INF, -INF};
            int kept = 0;
            const uint32_t dbase = MC_HDR_DEC(h) + 1u;
#pragma unroll
            for (int i = 0; i < 6; ++i) {
                const bool drop = !((mask >> i) & 1u) || pv[i].hi <= ext - margin - pad_of(ext - margin);
                if (dec && guarded) dec[(size_t)(dbase + (uint32_t)i) * dec_stride] = drop ? MC_DEC_FAIL : MC_DEC_KEEP;
                if (!drop) { ++kept; e.lo = fmax(e.lo, pv[i].lo); e.hi = fmax(e.hi, pv[i].hi); }
            }
            if (kept > 1) {
                // smooth branch: rvmax <= r/4 ln sum exp(hi_i / (r/4)) over the kept planes (<= max + r/4 ln 6);
                // the exact branch returns the max, which is below that too
                double up = 0.25 * r * 1.791759469228055;
                if (r > 0.0) {
                    double ssum = 0.0;
#pragma unroll
                    for (int i = 0; i < 6; ++i) {
                        const bool drop = !((mask >> i) & 1u) || pv[i].hi <= ext - margin - pad_of(ext - margin);
                        if (!drop) ssum += exp((pv[i].hi - e.hi) / (0.25 * r));
                    }
                    up = fmin(up, 0.25 * r * log(ssum));
                }
                e.hi += up;
                e = widen(e);
            }
            if (!guarded || pd == MC_DEC_PASS) {
                if (alive) wk += 6.0f * (float)kept + (kept > 1 ? 30.0f * (float)(kept + 1) : 0.0f);
                MC_IPUSH(e);
                break;
            }
            const Ival a = iv_box_distance(w, pc + 1, b);
            const double slack = ld_f64(w, pc + 7);
            uint8_t d = MC_DEC_KEEP;
            Ival v = {fmin(a.lo, e.lo), fmax(a.hi, e.hi)};
            if (a.hi <= slack) { d = MC_DEC_PASS; v = e; }
            else if (a.lo > slack) { d = MC_DEC_FAIL; v = a; }
            if (dec) dec[(size_t)MC_HDR_DEC(h) * dec_stride] = d;
            if (alive) wk += d == MC_DEC_FAIL ? 12.0f : 12.0f + 6.0f * (float)kept + (kept > 1 ? 30.0f * (float)(kept + 1) : 0.0f);
            MC_IPUSH(v);
            break;
        }
        case MC_OP_GAPUSH:
        case MC_OP_GUARD: {
            const uint8_t pd = MC_PD(MC_HDR_DEC(h));
            const Ival a = pd == MC_DEC_PASS ? Ival{0.0, 0.0} : iv_box_distance(w, pc + 1, b);
            if (pd == MC_DEC_FAIL) {
                MC_IPUSH(a);
                pc = (int)MC_HDR_IDX(h);
                continue;
            }
            const double slack = ld_f64(w, pc + 7);
            uint8_t d = MC_DEC_KEEP;
            if (pd == MC_DEC_PASS || a.hi <= slack) d = MC_DEC_PASS;
            else if (a.lo > slack) d = MC_DEC_FAIL;
            if (dec) dec[(size_t)MC_HDR_DEC(h) * dec_stride] = d;
            if (alive) {
                wk += op == MC_OP_GAPUSH ? 30.0f : 12.0f;
                if (d == MC_DEC_FAIL) dead_end = (int)MC_HDR_IDX(h);
            }
            // the subtree is skipped only when it is dead for all 32 tiles of the warp
            if (__all_sync(0xffffffffu, d == MC_DEC_FAIL)) {
                MC_IPUSH(a);
                pc = (int)MC_HDR_IDX(h);
                continue;
            }
            if (op == MC_OP_GAPUSH) {
#pragma unroll
                for (int k = 0; k < 6; ++k) bs[bsp + k] = b[k];
                bsp += 6;
                double nb[6];
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    Ival s = {ld_f64(w, pc + 11 + 4 * i), ld_f64(w, pc + 11 + 4 * i)};
#pragma unroll
                    for (int j = 0; j < 3; ++j) {
                        const Ival t = iv_scale(ld_f64(w, pc + 8 + 4 * i + j), b[2 * j], b[2 * j + 1]);
                        s.lo += t.lo; s.hi += t.hi;
                    }
                    s = widen(s);
                    nb[2 * i] = s.lo; nb[2 * i + 1] = s.hi;
                }
#pragma unroll
                for (int k = 0; k < 6; ++k) b[k] = nb[k];
            }
            break;
        }
        case MC_OP_ENDG: {
            const int g = (int)MC_HDR_IDX(h);
            if (MC_PD(MC_HDR_DEC(w[g])) == MC_DEC_PASS) break;  // passes everywhere: subtree value
            const Ival a = iv_box_distance(w, g + 1, b);
            const double slack = ld_f64(w, g + 7);
            if (a.hi <= slack) { /* pass everywhere: subtree value */ }
            else if (a.lo > slack) top = a;
            else top = {fmin(top.lo, a.lo), fmax(top.hi, a.hi)};
            break;
        }
        case MC_OP_NEG:
            top = {-top.hi, -top.lo};
            break;
        case MC_OP_CHILD:
            if (pdec && MC_PD(MC_HDR_DEC(h)) == MC_DEC_FAIL) { pc = (int)MC_HDR_IDX(h); continue; }  // dropped child
            break;
        case MC_OP_BOOL: {
            const unsigned sm = MC_HDR_SMALL(h);
            int k = (int)(sm & 0x7fu);
            if (pdec) {
                // only the children the containing box kept are on the stack
                int kept_children = 0;
                for (int i = 0; i < k; ++i) kept_children += MC_PD(MC_HDR_DEC(h) + (uint32_t)i) != MC_DEC_FAIL;
                k = kept_children;
                if (k <= 1) break;
            }
            const bool is_union = (sm & 0x80u) != 0;
            const double r = ld_f64(w, pc + 1), er = ld_f64(w, pc + 2);
            const double margin = fmax(r, er);
            const int base = sp - k + 1;
            // bound of the extremum: min of the upper bounds (union) / max of the lower bounds
            double ext = is_union ? INF : -INF;
            for (int i = 0; i < k; ++i) {
                const Ival x = (i == k - 1) ? top : vs[base + i];
                ext = is_union ? fmin(ext, x.hi) : fmax(ext, x.lo);
            }
            int kept = 0;
            Ival res = is_union ? Ival{INF, INF} : Ival{-INF, -INF};
            const uint32_t dbase = MC_HDR_DEC(h);
            for (int i = 0; i < k; ++i) {
                const Ival x = (i == k - 1) ? top : vs[base + i];
                bool drop;
                if (is_union) drop = x.lo >= ext + margin + pad_of(ext + margin);
                else drop = x.hi <= ext - margin - pad_of(ext - margin);
                if (dec) dec[(size_t)(dbase + (uint32_t)i) * dec_stride] = drop ? MC_DEC_FAIL : MC_DEC_KEEP;
                if (!drop) {
                    ++kept;
                    if (is_union) { res.lo = fmin(res.lo, x.lo); res.hi = fmin(res.hi, x.hi); }
                    else { res.lo = fmax(res.lo, x.lo); res.hi = fmax(res.hi, x.hi); }
                }
            }
            if (alive && kept > 1) wk += 30.0f * (float)(kept + 1);
            if (kept > 1) {
                // smooth branch: rvmin = -r/4 ln sum exp(-x_i / (r/4)) over a subset of the kept children
                // >= -r/4 ln sum exp(-lo_i / (r/4)) over all of them (>= min lo - r/4 ln k); the exact
                // branch returns the min, which is above that too.  Mirror image for rvmax.  Written
                // relative to the extremum of the bounds so that no term exceeds 1.
                double wdn = 0.25 * r * 1.3862943611198906 * (double)(kept <= 4 ? 1 : 2) * (kept <= 16 ? 1.0 : 2.5);
                if (r > 0.0) {
                    double ssum = 0.0;
                    for (int i = 0; i < k; ++i) {
                        const Ival x = (i == k - 1) ? top : vs[base + i];
                        bool drop;
                        if (is_union) drop = x.lo >= ext + margin + pad_of(ext + margin);
                        else drop = x.hi <= ext - margin - pad_of(ext - margin);
                        if (!drop) ssum += is_union ? exp(-(x.lo - res.lo) / (0.25 * r)) : exp((x.hi - res.hi) / (0.25 * r));
                    }
                    wdn = fmin(wdn, 0.25 * r * log(ssum));
                }
                if (is_union) res.lo -= wdn; else res.hi += wdn;
                res = widen(res);
            }
            sp -= k - 1;
            top = res;
            break;
        }
        case MC_OP_APUSH: {
            if (alive) wk += 18.0f;
#pragma unroll
            for (int k = 0; k < 6; ++k) bs[bsp + k] = b[k];
            bsp += 6;
            double nb[6];
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                Ival s = {ld_f64(w, pc + 4 + 4 * i), ld_f64(w, pc + 4 + 4 * i)};
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    const Ival t = iv_scale(ld_f64(w, pc + 1 + 4 * i + j), b[2 * j], b[2 * j + 1]);
                    s.lo += t.lo; s.hi += t.hi;
                }
                s = widen(s);
                nb[2 * i] = s.lo; nb[2 * i + 1] = s.hi;
            }
#pragma unroll
            for (int k = 0; k < 6; ++k) b[k] = nb[k];
            break;
        }
        case MC_OP_APOP: {
            bsp -= 6;
#pragma unroll
            for (int k = 0; k < 6; ++k) b[k] = bs[bsp + k];
            top = widen(iv_scale(ld_f64(w, pc + 1), top.lo, top.hi));
            break;
        }
        default:
            if (work) *work = wk;
            return Ival{-INF, INF};
        }
        pc += mc_op_words(op);
    }
#undef MC_IPUSH
#undef MC_PD
}

// decisions for every tile: one tile per lane
static __global__ void __launch_bounds__(128)
k_tile_decide(Lattice L, TileGrid g, const uint64_t* __restrict__ prog, uint64_t ntiles, uint8_t* __restrict__ dec,
              int8_t* __restrict__ tsign, float* __restrict__ twork = nullptr) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t tc = t < ntiles ? t : ntiles - 1;
    uint32_t X0, Y0, Z0;
    tile_origin(g, tc, X0, Y0, Z0);
    // The box covers the tile's points plus one lattice step on every side: the reach of the lattice
    // edges OWNED by them (canonical directions: -1..+1 in x and y, 0..+1 in z), so that the
    // decisions also hold along every cut edge the tile bisects, and the neighbourhood the sub-box
    // sign proofs (k_subbox_decide) look at.  Lattice coordinates are monotone in the index, so
    // the end points bound the box exactly.
    const uint32_t Xa = X0 > 0 ? X0 - 1 : 0, Ya = Y0 > 0 ? Y0 - 1 : 0, Za = Z0 > 0 ? Z0 - 1 : 0;
    const uint32_t X1 = min(X0 + TS, L.nx), Y1 = min(Y0 + TS, L.ny), Z1 = min(Z0 + TS, L.nz);
    const double box[6] = {lat_x(L, Xa), lat_x(L, X1), lat_y(L, Ya), lat_y(L, Y1), lat_z(L, Za), lat_z(L, Z1)};
    // lanes past the end redo the last tile (same decisions, benign duplicate stores)
    float wk = 0.0f;
    const Ival root = vm_decide(prog, box, dec + tc, (size_t)ntiles, twork ? &wk : nullptr);
    if (twork) twork[tc] = wk;
    // proven sign of the whole tile (the bounds are padded; NaN proves nothing)
    tsign[tc] = root.hi < 0.0 ? (int8_t)-1 : (root.lo > 0.0 ? (int8_t)1 : (int8_t)0);
}

// 1 / sample of the tiles, spread evenly (the load estimate looks at a sample of the sub-box proofs)
__device__ __forceinline__ bool tile_sampled(const TileGrid& g, uint64_t tile, uint32_t sample) {
    if (sample <= 1u) return true;
    const uint32_t tix = (uint32_t)(tile % g.ntx), tiy = (uint32_t)((tile / g.ntx) % g.nty), tiz = (uint32_t)(tile / ((uint64_t)g.ntx * g.nty)) + g.z0 / TS;
    return (tix + 2u * tiy + 3u * tiz) % sample == 0u;  // global tile coordinates: independent of the shard
}

// Sign proofs below the tile level: the 32 sub-boxes (8 x 4 x 4 points) of every tile whose sign is
// not proven for its whole 27-neighbourhood.  One warp per tile, one sub-box per lane: an interval
// evaluation of the program over the sub-box plus one lattice step on every side.  A definite
// sign there means that no tet or edge touching the sub-box's points has a vertex of another sign
// class, i.e. their values are never read (same argument as for whole tiles) and k_sdf_field_tiled
// writes the placeholder instead of evaluating them.  The lanes of a warp share the tile, so
// subtrees whose bounding box is far from the tile are skipped warp-uniformly.
// sub_masks[2 * tile] = sub-boxes proven negative, [2 * tile + 1] = proven positive.
// With `tile_list` (the mesher's call; not the load estimate's): a tile all of whose sub-boxes are
// proven with ONE sign is promoted to a proven tile (tsign), proven tiles get their sign summary
// (tflag) here and are not visited by the SDF kernel at all — nothing is stored for them —, the other
// tiles are appended to the SDF kernel's work list (tile_list[0 .. *n_list)).
static __global__ void __launch_bounds__(128)
k_subbox_decide(Lattice L, TileGrid g, const uint64_t* __restrict__ prog, uint64_t ntiles, uint32_t n_dec,
                const uint8_t* __restrict__ dec, int8_t* __restrict__ tsign, uint32_t* __restrict__ sub_masks,
                unsigned long long* __restrict__ sub_boxes_proven, uint32_t sample = 1u,
                uint32_t* __restrict__ tile_list = nullptr, unsigned long long* __restrict__ n_list = nullptr,
                uint8_t* __restrict__ tflag = nullptr, unsigned long long* __restrict__ skipped_tiles = nullptr) {
    extern __shared__ uint8_t s_pd_all[];  // the tiles' decisions, one slice per warp
    const uint64_t tile = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (tile >= ntiles) return;  // warp-uniform
    if (!tile_sampled(g, tile, sample)) return;  // load estimate: every sample-th tile only
    const uint32_t lane = threadIdx.x & 31u;
    uint8_t* s_pd = s_pd_all + (size_t)(threadIdx.x >> 5) * ((n_dec + 15u) & ~15u);
    const int sg = tsign[tile];
    if (sg != 0) {
        // the tile's sign is proven on its box plus one lattice step on every side (k_tile_decide),
        // which contains every sub-box plus its step: all 32 are proven
        if (lane == 0) {
            sub_masks[2 * tile] = sg < 0 ? 0xffffffffu : 0u; sub_masks[2 * tile + 1] = sg < 0 ? 0u : 0xffffffffu;
            if (tile_list) { tflag[tile] = sg < 0 ? TF_NEG : TF_POS; atomicAdd(skipped_tiles, 1ull); }
        }
        return;
    }
    uint32_t X0, Y0, Z0;
    tile_origin(g, tile, X0, Y0, Z0);
    const uint32_t bx = X0 + (lane & 1u) * 8u, by = Y0 + ((lane >> 1) & 3u) * 4u, bz = Z0 + (lane >> 3) * 4u;
    const uint32_t xa = bx > 0 ? bx - 1 : 0, ya = by > 0 ? by - 1 : 0, za = bz > 0 ? bz - 1 : 0;
    const double box[6] = {lat_x(L, min(xa, L.nx)), lat_x(L, min(bx + 8u, L.nx)), lat_y(L, min(ya, L.ny)),
                           lat_y(L, min(by + 4u, L.ny)), lat_z(L, min(za, L.nz)), lat_z(L, min(bz + 4u, L.nz))};
    // the tile's own decisions hold on its box plus one step on every side, which contains `box`;
    // they are gathered into shared memory once (the table is decision-major)
    for (uint32_t i = lane; i < n_dec; i += 32u) s_pd[i] = dec[(size_t)i * ntiles + tile];
    __syncwarp();
    const Ival r = vm_decide(prog, box, nullptr, 0, nullptr, s_pd, 1);
    const unsigned neg = __ballot_sync(0xffffffffu, r.hi < 0.0), pos = __ballot_sync(0xffffffffu, r.lo > 0.0);
    if (lane == 0) {
        sub_masks[2 * tile] = neg;
        sub_masks[2 * tile + 1] = pos;
        if (tile_list && (neg == 0xffffffffu || pos == 0xffffffffu)) {
            // the 32 grown sub-boxes cover the tile's grown box: the tile is proven after all
            tsign[tile] = neg ? (int8_t)-1 : (int8_t)1;
            tflag[tile] = neg ? TF_NEG : TF_POS;
            atomicAdd(skipped_tiles, 1ull);
        } else {
            if (neg | pos) atomicAdd(sub_boxes_proven, (unsigned long long)__popc(neg | pos));
            if (tile_list) tile_list[atomicAdd(n_list, 1ull)] = (uint32_t)tile;
        }
    }
}

// cost classes of the tiles per tile layer (see tile_layer_classes in mc_mesher.cu):
// counts[layer * MC_TILE_CLASS_STRIDE + class], + 5: work of the class-0 tiles, + 6: of the shell tiles
static __global__ void k_tile_layer_classes(TileGrid g, const int8_t* __restrict__ tsign, const float* __restrict__ twork,
                                            const uint32_t* __restrict__ sub_masks, uint32_t sample,
                                            uint64_t ntiles, uint32_t first, uint32_t last,
                                            unsigned long long* __restrict__ counts) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ntiles) return;
    const int tix = (int)(t % g.ntx), tiy = (int)((t / g.ntx) % g.nty), tiz = (int)(t / ((uint64_t)g.ntx * g.nty));
    if ((uint32_t)tiz < first || (uint32_t)tiz >= last) return;  // halo layer: only looked at
    const int sg = tsign[t];
    int cls = 0;
    if (sg != 0) {
        bool uniform = true;
        for (int q = 0; q < 27 && uniform; ++q) {
            const int x = tix + q % 3 - 1, y = tiy + (q / 3) % 3 - 1, z = tiz + q / 9 - 1;
            if (x < 0 || y < 0 || z < 0 || x >= (int)g.ntx || y >= (int)g.nty || z >= (int)g.ntz) continue;
            if (tsign[((size_t)z * g.nty + y) * g.ntx + x] != sg) uniform = false;
        }
        cls = sg < 0 ? (uniform ? 1 : 2) : (uniform ? 3 : 4);
    }
    atomicAdd(&counts[(size_t)tiz * MC_TILE_CLASS_STRIDE + cls], 1ull);
    // per-point work of the tile programs that are evaluated in full / on the shell only
    const unsigned long long wk = (unsigned long long)twork[t];
    if (cls == 0) {
        // work of the sub-boxes that stay unproven (from the sampled tiles, scaled up)
        if (tile_sampled(g, t, sample)) {
            const uint32_t un = (uint32_t)__popc(~(sub_masks[2 * t] | sub_masks[2 * t + 1]));
            atomicAdd(&counts[(size_t)tiz * MC_TILE_CLASS_STRIDE + 5], wk * un * sample / 32ull);
        }
    } else if (cls == 2 || cls == 4) {
        atomicAdd(&counts[(size_t)tiz * MC_TILE_CLASS_STRIDE + 6], wk);
    }
}

// Evaluation of the un-specialised program from global memory, out of line: the fallback for the
// rare tile whose specialised program exceeds the shared-memory budget (same values by
// construction).  All 32 lanes must call it.
static __device__ __noinline__ double vm_eval_full_program(const uint64_t* __restrict__ prog, double x, double y, double z) {
    return vm_eval<true>(prog, x, y, z);
}

// Block-cooperative compaction of the program for one tile into shared memory (s_prog), using the
// tile's decisions.  All threads of the block must call it; it ends with the result visible to
// the whole block once the caller has synchronised.  Returns the number of words via s_newpos[n_inst].
__device__ __forceinline__ void compact_tile_program(const uint64_t* __restrict__ prog,
                                                     const uint32_t* __restrict__ inst_start,
                                                     const uint16_t* __restrict__ word_inst, uint32_t n_inst,
                                                     uint32_t n_dec, const uint8_t* __restrict__ dec_g, uint64_t ntiles,
                                                     uint64_t tile, uint64_t* s_prog, uint8_t* s_dec, int* s_delta,
                                                     uint32_t* s_newpos, uint32_t* s_scan, uint32_t cap_words) {
    const uint32_t tid = threadIdx.x;
    for (uint32_t i = tid; i < n_dec; i += blockDim.x) s_dec[i] = dec_g[(size_t)i * ntiles + tile];
    for (uint32_t i = tid; i <= n_inst; i += blockDim.x) s_delta[i] = 0;
    __syncthreads();
    // 1. dead ranges: dropped boolean children and always-failing guards
    for (uint32_t i = tid; i < n_inst; i += blockDim.x) {
        const uint64_t h = prog[inst_start[i]];
        const uint32_t op = MC_HDR_OP(h);
        if (op == MC_OP_CHILD && s_dec[MC_HDR_DEC(h)] == MC_DEC_FAIL) {
            atomicAdd(&s_delta[i + 1], 1);
            atomicAdd(&s_delta[word_inst[MC_HDR_IDX(h)]], -1);
        } else if ((op == MC_OP_GUARD || op == MC_OP_GAPUSH) && s_dec[MC_HDR_DEC(h)] == MC_DEC_FAIL) {
            atomicAdd(&s_delta[i + 1], 1);
            atomicAdd(&s_delta[word_inst[MC_HDR_IDX(h)]], -1);
        }
    }
    __syncthreads();
    // 2. inclusive scan of the deltas -> nesting depth of dead ranges; new length of every
    //    instruction; exclusive scan of the lengths -> new word positions
    uint32_t run_depth = 0, run_pos = 0;
    for (uint32_t base = 0; base <= n_inst; base += blockDim.x) {
        const uint32_t i = base + tid;
        const int dlt = (i <= n_inst) ? s_delta[i] : 0;
        uint32_t tot;
        // depths are non-negative at every prefix, so unsigned wrap-around arithmetic is safe
        const uint32_t ex = block_exscan((uint32_t)dlt, &tot, s_scan);
        const uint32_t depth = run_depth + ex + (uint32_t)dlt;
        run_depth += tot;
        uint32_t len = 0;
        if (i < n_inst && depth == 0) {
            const uint64_t h = prog[inst_start[i]];
            const uint32_t op = MC_HDR_OP(h);
            const uint8_t d = (op == MC_OP_GUARD || op == MC_OP_ENDG || op == MC_OP_GSPHERE || op == MC_OP_GCYL ||
                               op == MC_OP_GBOX || op == MC_OP_GAPUSH)
                                  ? s_dec[MC_HDR_DEC(h)] : (uint8_t)MC_DEC_KEEP;
            switch (op) {
            case MC_OP_GUARD: len = d == MC_DEC_PASS ? 0u : (d == MC_DEC_FAIL ? 7u : 8u); break;
            case MC_OP_GAPUSH: len = d == MC_DEC_PASS ? 13u : (d == MC_DEC_FAIL ? 7u : 20u); break;
            case MC_OP_GBOX: len = d == MC_DEC_PASS ? 9u : (d == MC_DEC_FAIL ? 7u : 16u); break;
            case MC_OP_ENDG: len = d == MC_DEC_KEEP ? 1u : 0u; break;
            case MC_OP_GSPHERE: case MC_OP_GCYL: len = d == MC_DEC_PASS ? 2u : (d == MC_DEC_FAIL ? 7u : 9u); break;
            case MC_OP_CHILD: len = 0u; break;
            case MC_OP_BOOL: {
                const uint32_t k = MC_HDR_SMALL(h) & 0x7fu, db = MC_HDR_DEC(h);
                uint32_t kept = 0;
                for (uint32_t c = 0; c < k; ++c) kept += s_dec[db + c] != MC_DEC_FAIL;
                len = kept > 1 ? 3u : 0u;
                break;
            }
            default: len = (uint32_t)mc_op_words(op); break;
            }
        }
        const uint32_t pex = block_exscan(len, &tot, s_scan);
        if (i <= n_inst) { s_newpos[i] = run_pos + pex; s_delta[i] = (int)len; }
        run_pos += tot;
    }
    __syncthreads();
    // 3. write the compacted program (the caller evaluates the full program instead when it does
    //    not fit the shared-memory budget: s_newpos[n_inst] > cap_words)
    if (s_newpos[n_inst] > cap_words) return;
    for (uint32_t i = tid; i < n_inst; i += blockDim.x) {
        const uint32_t len = (uint32_t)s_delta[i];
        if (len == 0) continue;
        const uint32_t src = inst_start[i], dst = s_newpos[i];
        const uint64_t h = prog[src];
        const uint32_t op = MC_HDR_OP(h);
        switch (op) {
        case MC_OP_GUARD:
            if (len == 7u) {
                s_prog[dst] = MC_HDR(MC_OP_BOXDIST, 0, 0, 0, 0);
                for (uint32_t q = 1; q < 7; ++q) s_prog[dst + q] = prog[src + q];
            } else {
                s_prog[dst] = MC_HDR(MC_OP_GUARD, 0, MC_HDR_LEVEL(h), 0, s_newpos[word_inst[MC_HDR_IDX(h)]]);
                for (uint32_t q = 1; q < 8; ++q) s_prog[dst + q] = prog[src + q];
            }
            break;
        case MC_OP_GAPUSH:
            if (len == 7u) {
                s_prog[dst] = MC_HDR(MC_OP_BOXDIST, 0, 0, 0, 0);
                for (uint32_t q = 1; q < 7; ++q) s_prog[dst + q] = prog[src + q];
            } else if (len == 13u) {
                s_prog[dst] = MC_HDR(MC_OP_APUSH, 0, 0, 0, 0);
                for (uint32_t q = 1; q < 13; ++q) s_prog[dst + q] = prog[src + 7 + q];
            } else {
                s_prog[dst] = MC_HDR(MC_OP_GAPUSH, 0, MC_HDR_LEVEL(h), 0, s_newpos[word_inst[MC_HDR_IDX(h)]]);
                for (uint32_t q = 1; q < 20; ++q) s_prog[dst + q] = prog[src + q];
            }
            break;
        case MC_OP_GBOX:
            if (len == 7u) {
                s_prog[dst] = MC_HDR(MC_OP_BOXDIST, 0, 0, 0, 0);
                for (uint32_t q = 1; q < 7; ++q) s_prog[dst + q] = prog[src + q];
            } else {
                uint32_t mask = 0;
                for (uint32_t q = 0; q < 6; ++q) mask |= (s_dec[MC_HDR_DEC(h) + 1u + q] != MC_DEC_FAIL ? 1u : 0u) << q;
                if (len == 9u) {
                    s_prog[dst] = MC_HDR(MC_OP_BOX, mask, 0, 0, 0);
                    for (uint32_t q = 1; q < 9; ++q) s_prog[dst + q] = prog[src + 7 + q];
                } else {
                    s_prog[dst] = MC_HDR(MC_OP_GBOX, mask, 0, 0, 0);
                    for (uint32_t q = 1; q < 16; ++q) s_prog[dst + q] = prog[src + q];
                }
            }
            break;
        case MC_OP_ENDG:
            s_prog[dst] = MC_HDR(MC_OP_ENDG, 0, MC_HDR_LEVEL(h), 0, s_newpos[word_inst[MC_HDR_IDX(h)]]);
            break;
        case MC_OP_GSPHERE: case MC_OP_GCYL:
            if (len == 2u) {
                s_prog[dst] = MC_HDR(op == MC_OP_GSPHERE ? MC_OP_SPHERE : MC_OP_CYL, 0, 0, 0, 0);
                s_prog[dst + 1] = prog[src + 8];
            } else if (len == 7u) {
                s_prog[dst] = MC_HDR(MC_OP_BOXDIST, 0, 0, 0, 0);
                for (uint32_t q = 1; q < 7; ++q) s_prog[dst + q] = prog[src + q];
            } else {
                for (uint32_t q = 0; q < 9; ++q) s_prog[dst + q] = prog[src + q];
            }
            break;
        case MC_OP_BOOL: {
            const uint32_t k = MC_HDR_SMALL(h) & 0x7fu, db = MC_HDR_DEC(h);
            uint32_t kept = 0;
            for (uint32_t c = 0; c < k; ++c) kept += s_dec[db + c] != MC_DEC_FAIL;
            s_prog[dst] = MC_HDR(MC_OP_BOOL, kept | (MC_HDR_SMALL(h) & 0x80u), 0, 0, 0);
            s_prog[dst + 1] = prog[src + 1];
            s_prog[dst + 2] = prog[src + 2];
            break;
        }
        default:
            for (uint32_t q = 0; q < len; ++q) s_prog[dst + q] = prog[src + q];
            break;
        }
    }
}

// per-tile compaction of the program into shared memory + evaluation of the tile's points
// smem layout: [out program (n_words = its capacity: the program's length, capped by the host)] [dec (n_dec
// bytes, padded)] [delta int32 (n_inst+1)] [newpos u32 (n_inst+1)]
// CAPPED: the program is longer than the shared-memory budget, a tile's specialised program may
// not fit (the fallback code is compiled out otherwise)
template <int WX, int WY, bool COUNT, bool SMALL, bool CAPPED>
__global__ void __launch_bounds__(256, 4)
k_sdf_field_tiled(Lattice L, TileGrid g, double* __restrict__ phi_out, const uint64_t* __restrict__ prog,
                  const uint32_t* __restrict__ inst_start, const uint16_t* __restrict__ word_inst, uint32_t n_inst,
                  uint32_t n_words, uint32_t n_dec, const uint8_t* __restrict__ dec_g, uint64_t ntiles,
                  const int8_t* __restrict__ tsign, int skip_uniform, unsigned long long* __restrict__ pruned_words_total,
                  unsigned long long* __restrict__ flops_total, const uint32_t* __restrict__ sub_masks,
                  uint8_t* __restrict__ tflag, const uint32_t* __restrict__ tile_list,
                  const unsigned long long* __restrict__ n_list, unsigned long long* __restrict__ next_item) {
    unsigned long long my_flops = 0ull;
    extern __shared__ uint64_t smem[];
    uint64_t* s_prog = smem;
    uint8_t* s_dec = (uint8_t*)(s_prog + n_words);
    int* s_delta = (int*)(s_dec + ((n_dec + 15u) & ~15u));
    uint32_t* s_newpos = (uint32_t*)(s_delta + (n_inst + 1));
    __shared__ uint32_t s_scan[9];
    const uint32_t tid = threadIdx.x;

    // Work items: the tiles k_subbox_decide listed (not proven), or every tile when the caller wants
    // the whole field.  Blocks take them one at a time from a global counter: the cost of a tile
    // (unproven sub-boxes x length of its program) varies by two orders of magnitude.
    __shared__ unsigned long long s_item;
    const unsigned long long n_items = tile_list ? *n_list : (unsigned long long)ntiles;
    for (;;) {
        __syncthreads();  // previous tile's evaluation is done with s_prog (and s_item has been read)
        if (tid == 0) s_item = atomicAdd(next_item, 1ull);
        __syncthreads();
        if (s_item >= n_items) break;
        const uint64_t tile = tile_list ? (uint64_t)tile_list[s_item] : (uint64_t)s_item;
        // A point whose whole neighbourhood (one lattice step on every side) is PROVEN to have one
        // strict sign is never read for its value: every consumer (warp, the 4-sort, bisection end
        // points) only looks at values of tets / edges that contain a vertex of another sign class,
        // and those lie within one lattice step.  Proven sub-boxes of a listed tile get a placeholder
        // of the right sign instead of being evaluated; proven TILES are not even listed.
        uint32_t X0, Y0, Z0;
        tile_origin(g, tile, X0, Y0, Z0);
        uint32_t sub_neg = 0u, sub_pos = 0u;
        if (skip_uniform) { sub_neg = sub_masks[2 * tile]; sub_pos = sub_masks[2 * tile + 1]; }
        compact_tile_program(prog, inst_start, word_inst, n_inst, n_dec, dec_g, ntiles, tile, s_prog, s_dec, s_delta,
                             s_newpos, s_scan, n_words);
        if (tid == 0 && pruned_words_total) atomicAdd(pruned_words_total, (unsigned long long)s_newpos[n_inst]);
        const bool fits = !CAPPED || s_newpos[n_inst] <= n_words;
        __syncthreads();
        // evaluate the tile: WX x WY points of one z-plane per warp step; the footprint of a step
        // lies in one sub-box
        static_assert(WX == 8 && WY == 4, "warp footprint = sub-box cross-section");
        const uint32_t lane = tid & 31u;
        bool not_neg = false, not_pos = false;  // a stored value that is not < 0 / not > 0 (NaN: both)
        // placeholders of the proven sub-boxes
        if (sub_neg | sub_pos) {
            // (two points per thread, one 16-byte store: rows are padded to an even pitch, X is even)
            for (uint32_t j = tid; j < TS * TS * TS / 2u; j += blockDim.x) {
                const uint32_t lx = (j & 7u) * 2u, ly = (j >> 3) & 15u, lz = j >> 7;
                const uint32_t sub = (lx >> 3) | ((ly >> 2) << 1) | ((lz >> 2) << 3);
                const bool dn = (sub_neg >> sub) & 1u, dp = (sub_pos >> sub) & 1u;
                const uint32_t X = X0 + lx, Y = Y0 + ly, Z = Z0 + lz;
                if ((dn || dp) && X < L.NX && Y < L.NY && Z <= L.pz1) {
                    const double ph = dn ? -1.0 : 1.0;
                    *reinterpret_cast<double2*>(phi_out + pidx(L, X, Y, Z)) = make_double2(ph, ph);
                    not_pos = not_pos || dn; not_neg = not_neg || dp;
                }
            }
        }
        // the other sub-boxes, four 8 x 4 warp steps each, dealt round-robin to the warps.  The loop is
        // instantiated twice so that the hot one holds nothing but the inlined evaluator.
        const uint32_t unproven = ~(sub_neg | sub_pos);
        const uint32_t nsteps = 4u * (uint32_t)__popc(unproven);
        auto run = [&](auto eval) {
            for (uint32_t t = tid >> 5; t < nsteps; t += (blockDim.x >> 5)) {
                const uint32_t sub = __fns(unproven, 0u, (int)(t >> 2) + 1);
                const uint32_t lx = (sub & 1u) * 8u + lane % WX, ly = ((sub >> 1) & 3u) * 4u + lane / WX, lz = (sub >> 3) * 4u + (t & 3u);
                const uint32_t X = X0 + lx, Y = Y0 + ly, Z = Z0 + lz;
                if (Z > L.pz1) continue;  // warp-uniform
                const bool inside = X < L.NX && Y < L.NY;
                const uint32_t Xc = X < L.NX ? X : L.NX - 1, Yc = Y < L.NY ? Y : L.NY - 1;
                unsigned long long fl = 0ull;
                const double v = eval(lat_x(L, Xc), lat_y(L, Yc), lat_z(L, Z), fl);
                // even lanes store their own value and their right neighbour's with one 16-byte store (the row
                // pitch is even, so the slot after the last point of a row exists)
                const double vr = __shfl_down_sync(0xffffffffu, v, 1);
                if (inside && (lane & 1u) == 0u) *reinterpret_cast<double2*>(phi_out + pidx(L, X, Y, Z)) = make_double2(v, vr);
                if (inside) {
                    my_flops += fl;
                    not_neg = not_neg || !(v < 0.0);
                    not_pos = not_pos || !(v > 0.0);
                }
            }
        };
        if (CAPPED && !fits) run([&](double x, double y, double z, unsigned long long&) { return vm_eval_full_program(prog, x, y, z); });
        else run([&](double x, double y, double z, unsigned long long& fl) { return vm_eval_impl<true, COUNT, SMALL>(s_prog, x, y, z, fl); });
        // sign summary of the tile (TF_NEG: every value < 0, TF_POS: every value > 0), the job of
        // k_tile_minmax on the untiled path
        const int any_not_neg = __syncthreads_or(not_neg ? 1 : 0), any_not_pos = __syncthreads_or(not_pos ? 1 : 0);
        if (tid == 0) tflag[tile] = (uint8_t)((any_not_neg ? 0 : TF_NEG) | (any_not_pos ? 0 : TF_POS));
    }
    if (COUNT) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) my_flops += __shfl_xor_sync(0xffffffffu, my_flops, o);
        if ((tid & 31u) == 0u && my_flops) atomicAdd(flops_total, my_flops);
    }
}


// a8 with the tile-specialised program: one block per tile that owns cut edges.  The tile's cut
// edges are contiguous in the work list (vertex numbering is tile-major).
template <bool SMALL, bool CAPPED>
__global__ void __launch_bounds__(TILE_THREADS)
k_bisect_edges_tiled(Stuff S, const uint64_t* __restrict__ prog, const uint32_t* __restrict__ inst_start,
                     const uint16_t* __restrict__ word_inst, uint32_t n_inst, uint32_t n_words, uint32_t n_dec,
                     const uint8_t* __restrict__ dec_g, uint64_t ntiles, const unsigned long long* __restrict__ tile_cbase,
                     uint64_t ncut_total, const unsigned long long* __restrict__ cutlist, double error_threshold,
                     uint32_t max_iters, double* __restrict__ coords, unsigned long long* __restrict__ counters,
                     const uint32_t* __restrict__ tile_list, const unsigned long long* __restrict__ n_list,
                     unsigned long long* __restrict__ cursor) {
    extern __shared__ uint64_t smem[];
    uint64_t* s_prog = smem;
    uint8_t* s_dec = (uint8_t*)(s_prog + n_words);
    int* s_delta = (int*)(s_dec + ((n_dec + 15u) & ~15u));
    uint32_t* s_newpos = (uint32_t*)(s_delta + (n_inst + 1));
    __shared__ uint32_t s_scan[9];
    unsigned long long it64 = 0ull;
    uint32_t tile32;
    while (next_tile(tile_list, n_list, cursor, tile32)) {  // active point tiles
    const uint64_t tile = tile32;
    const uint64_t c0 = tile_cbase[tile], c1 = tile + 1 < ntiles ? tile_cbase[tile + 1] : ncut_total;
    if (c1 <= c0) continue;
    compact_tile_program(prog, inst_start, word_inst, n_inst, n_dec, dec_g, ntiles, tile, s_prog, s_dec, s_delta, s_newpos,
                         s_scan, n_words);
    const bool fits = !CAPPED || s_newpos[n_inst] <= n_words;
    __syncthreads();
    const Lattice& L = S.L;
    const uint64_t per_plane = (uint64_t)L.PX * L.NY;
    const uint64_t n = c1 - c0;
    for (uint64_t base = 0; base < n; base += TILE_THREADS) {
        const uint64_t i = base + threadIdx.x;
        const bool active = i < n;
        const unsigned long long e = cutlist[c0 + (active ? i : n - 1)];
        const size_t pi = (size_t)(e >> 4);
        const int s = (int)(e & 15ull);
        const uint32_t Z = L.pz0 + (uint32_t)(pi / per_plane);
        const uint64_t r = pi % per_plane;
        const uint32_t Y = (uint32_t)(r / L.PX), X = (uint32_t)(r % L.PX);
        const uint32_t WX = X + C_DIR[s][0], WY = Y + C_DIR[s][1], WZ = Z + C_DIR[s][2];
        const double pv = L.phi[pi];
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
            unsigned long long fl = 0ull;
            const double phi = (CAPPED && !fits) ? vm_eval_full_program(prog, mx, my, mz)
                                                 : vm_eval_impl<true, false, SMALL>(s_prog, mx, my, mz, fl);
            if (!done) {
                cx = mx; cy = my; cz = mz;
                ++iters;
                if (fabs(phi) <= error_threshold) done = true;
                else if (phi > 0.0) { ax = mx; ay = my; az = mz; }
                else { bx = mx; by = my; bz = mz; }
                if (!done && iters >= max_iters) {
                    done = true;
                    atomicAdd(&counters[CN_BISECT_FAIL], 1ull);
                }
            }
        }
        if (active) {
            uint16_t mask;
            const long long vb = vertex_gid(S, X, Y, Z, mask);
            const long long id = cut_vertex_id(vb, mask, s) - S.vertex_base;
            coords[3 * id] = cx; coords[3 * id + 1] = cy; coords[3 * id + 2] = cz;
            it64 += iters;
        }
    }
    }  // tiles
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) it64 += __shfl_xor_sync(0xffffffffu, it64, o);
    if ((threadIdx.x & 31u) == 0u && it64) atomicAdd(&counters[CN_BISECT_ITERS], it64);
}

}  // namespace mc
