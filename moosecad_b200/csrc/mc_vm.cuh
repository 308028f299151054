// Device evaluator of the flat SDF program (mc_program.h): one lattice point per lane,
// the program counter is warp-uniform when UNIFORM is set (all 32 lanes must be converged):
// a bbox guard is skipped only when NO lane passes it, otherwise every lane evaluates the
// subtree and the guard's select is applied per lane — value-identical to the reference's
// `if approx <= slack { exact } else { approx }` because the tree is side-effect free.
//
// Arithmetic follows implicit3d 0.14.2 / bbox 0.11.2 / nalgebra 0.22.1 operation order
// (DESIGN.md §SDF); compiled with --fmad=false.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "mc_math.cuh"
#include "mc_program.h"

namespace mc {

__device__ __forceinline__ double ld_f64(const uint64_t* __restrict__ w, int i) {
    return __longlong_as_double((long long)w[i]);
}

// BoundingBox::distance: max(xval, max(yval, zval)), *val = max(p - hi, lo - p)
__device__ __forceinline__ double box_distance(const uint64_t* __restrict__ w, int at, double x, double y,
                                               double z) {
    const double xv = fmax(x - ld_f64(w, at + 3), ld_f64(w, at + 0) - x);
    const double yv = fmax(y - ld_f64(w, at + 4), ld_f64(w, at + 1) - y);
    const double zv = fmax(z - ld_f64(w, at + 5), ld_f64(w, at + 2) - z);
    return fmax(xv, fmax(yv, zv));
}

// rvmin / rvmax over k stacked child values v[0..k) (v[k-1] is `top`)
// FP64 operation counts used by the counting build of the evaluator (bench.py's roofline):
// one per add / sub / mul / max / abs / compare, division and sqrt count 1, mc_exp = 23, mc_log = 27
// (their main paths)
#define MC_FLOPS_EXP 23
#define MC_FLOPS_LOG 27

template <bool IS_UNION, bool COUNT>
__device__ __forceinline__ double smooth_fold(const double* __restrict__ vs, int base, int k, double top,
                                              double r, double exact_range, unsigned long long& flops) {
    // pass 1: running extremum + "two children are within exact_range" flag
    bool close = false;
    double m = IS_UNION ? __longlong_as_double(0x7ff0000000000000LL) : __longlong_as_double(0xfff0000000000000LL);
    for (int i = 0; i < k; ++i) {
        const double x = (i == k - 1) ? top : vs[base + i];
        if (IS_UNION) {
            if (x < m) { close = (m - x) < exact_range; m = x; }
            else if ((x - m) < exact_range) close = true;
        } else {
            if (x > m) { close = (x - m) < exact_range; m = x; }
            else if ((m - x) < exact_range) close = true;
        }
    }
    if (COUNT) flops += 3ull * (unsigned)k;
    if (!close) return m;
    const double lim = IS_UNION ? (m + r) : (m - r);
    const double r4 = r / 4.0;
    double s = 0.0;
    for (int i = 0; i < k; ++i) {
        const double x = (i == k - 1) ? top : vs[base + i];
        if (IS_UNION) { if (x < lim) { s = s + mc_exp(-x / r4); if (COUNT) flops += 4 + MC_FLOPS_EXP; } }
        else          { if (x > lim) { s = s + mc_exp(x / r4); if (COUNT) flops += 3 + MC_FLOPS_EXP; } }
    }
    if (COUNT) flops += 3 + MC_FLOPS_LOG;
    return IS_UNION ? (mc_log(s) * -r4) : (mc_log(s) * r4);
}

// rvmin / rvmax over k <= 4 operands held in registers (v0..v3 in push order, only the first k used)
template <bool IS_UNION, bool COUNT>
__device__ __forceinline__ double smooth_fold4(double v0, double v1, double v2, double v3, int k, double r,
                                               double exact_range, unsigned long long& flops) {
    const double v[4] = {v0, v1, v2, v3};
    bool close = false;
    double m = IS_UNION ? __longlong_as_double(0x7ff0000000000000LL) : __longlong_as_double(0xfff0000000000000LL);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (i >= k) break;
        const double x = v[i];
        if (IS_UNION) {
            if (x < m) { close = (m - x) < exact_range; m = x; }
            else if ((x - m) < exact_range) close = true;
        } else {
            if (x > m) { close = (x - m) < exact_range; m = x; }
            else if ((m - x) < exact_range) close = true;
        }
    }
    if (COUNT) flops += 3ull * (unsigned)k;
    if (!close) return m;
    const double lim = IS_UNION ? (m + r) : (m - r);
    const double r4 = r / 4.0;
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (i >= k) break;
        const double x = v[i];
        if (IS_UNION) { if (x < lim) { s = s + mc_exp(-x / r4); if (COUNT) flops += 4 + MC_FLOPS_EXP; } }
        else          { if (x > lim) { s = s + mc_exp(x / r4); if (COUNT) flops += 3 + MC_FLOPS_EXP; } }
    }
    if (COUNT) flops += 3 + MC_FLOPS_LOG;
    return IS_UNION ? (mc_log(s) * -r4) : (mc_log(s) * r4);
}

// geom.cube: rvmax over the six axis planes, everything in registers
// (`mask`: planes that can influence the result on this tile; dropped ones are skipped, which is
// value-neutral, see mc_prune.cuh)
template <bool COUNT>
__device__ __forceinline__ double box_fold(const uint64_t* __restrict__ w, int at, unsigned mask, double x, double y,
                                           double z, unsigned long long& flops) {
    double v[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        if (!((mask >> i) & 1u)) { v[i] = 0.0; continue; }
        const double q = (i % 3) == 0 ? x : ((i % 3) == 1 ? y : z);
        const double d = ld_f64(w, at + i);
        const double ap = fabs(q);
        const bool sign_positive = __double_as_longlong(q) >= 0;
        const bool inverted = i >= 3;
        v[i] = (sign_positive != inverted) ? (ap - d) : (-ap - d);
    }
    const double r = ld_f64(w, at + 6), exact_range = ld_f64(w, at + 7);
    bool close = false;
    double m = __longlong_as_double(0xfff0000000000000LL);
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        if (!((mask >> i) & 1u)) continue;
        const double xx = v[i];
        if (xx > m) { close = (xx - m) < exact_range; m = xx; }
        else if ((m - xx) < exact_range) close = true;
    }
    if (COUNT) flops += 5ull * (unsigned)__popc(mask);
    if (!close) return m;
    const double lim = m - r, r4 = r / 4.0;
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 6; ++i)
        if (((mask >> i) & 1u) && v[i] > lim) { s = s + mc_exp(v[i] / r4); if (COUNT) flops += 3 + MC_FLOPS_EXP; }
    if (COUNT) flops += 3 + MC_FLOPS_LOG;
    return mc_log(s) * r4;
}

// Evaluate the program at (x, y, z).  `w` may point to shared or global memory.
// SMALL: programs whose value stack never holds more than 4 entries, whose booleans have at most
// 4 children and whose affine nodes do not nest keep everything in registers (s1..s3 below `top`,
// one saved point) instead of the local-memory stacks; the host picks the variant per program.
template <bool UNIFORM, bool COUNT, bool SMALL>
__device__ __forceinline__ double vm_eval_impl(const uint64_t* __restrict__ w, double x, double y, double z, unsigned long long& flops) {
    double vs[SMALL ? 1 : MC_MAX_VALUE_DEPTH + 1];   // element i lives in vs[i + 1]; the newest is `top`
    double ps[SMALL ? 1 : 3 * MC_MAX_POINT_DEPTH];
    double s1 = 0.0, s2 = 0.0, s3 = 0.0, sx = 0.0, sy = 0.0, sz = 0.0;
    int sp = 0, psp = 0, pc = 0;
    double top = 0.0;
    unsigned long long allpass = 0ull;  // bit L: every lane passed the guard at nesting level L
    const unsigned FULL = 0xffffffffu;
#define MC_PUSH(v) do { if (SMALL) { s3 = s2; s2 = s1; s1 = top; } else { vs[sp] = top; ++sp; } top = (v); } while (0)
    for (;;) {
        const uint64_t h = w[pc];
        switch (MC_HDR_OP(h)) {
        case MC_OP_END:
            return top;
        case MC_OP_PLANE: {
            const unsigned sm = MC_HDR_SMALL(h);
            const unsigned axis = sm & 3u;
            const double q = axis == 0 ? x : (axis == 1 ? y : z);
            const double d = ld_f64(w, pc + 1);
            const double ap = fabs(q);
            const bool sign_positive = __double_as_longlong(q) >= 0;
            const bool inverted = (sm & 4u) != 0;
            MC_PUSH((sign_positive != inverted) ? (ap - d) : (-ap - d));
            if (COUNT) flops += 2;
            pc += 2;
            break;
        }
        case MC_OP_NPLANE: {
            const double v = ((x * ld_f64(w, pc + 1) + y * ld_f64(w, pc + 2)) + z * ld_f64(w, pc + 3)) -
                             ld_f64(w, pc + 4);
            MC_PUSH(v);
            if (COUNT) flops += 6;
            pc += 5;
            break;
        }
        case MC_OP_USQ: {
            MC_PUSH(((x * x + y * y) + z * z) - 1.0);
            if (COUNT) flops += 6;
            pc += 1;
            break;
        }
        case MC_OP_CONE: {
            const double radius = fabs(ld_f64(w, pc + 1) * (z + ld_f64(w, pc + 2)));
            MC_PUSH((sqrt(x * x + y * y) - radius) * ld_f64(w, pc + 3));
            if (COUNT) flops += 9;
            pc += 4;
            break;
        }
        case MC_OP_GSPHERE:
        case MC_OP_GCYL: {
            const double a = box_distance(w, pc + 1, x, y, z);
            const bool pass = a <= ld_f64(w, pc + 7);
            double v = a;
            if (UNIFORM ? __any_sync(FULL, pass) : pass) {
                const double r = ld_f64(w, pc + 8);
                const double e = (MC_HDR_OP(h) == MC_OP_GSPHERE) ? (sqrt((x * x + y * y) + z * z) - r)
                                                                 : (sqrt(x * x + y * y) - r);
                if (pass) v = e;
                if (COUNT) flops += (MC_HDR_OP(h) == MC_OP_GSPHERE) ? 7 : 5;
            }
            if (COUNT) flops += 12;
            MC_PUSH(v);
            pc += 9;
            break;
        }
        case MC_OP_GBOX: {
            const double a = box_distance(w, pc + 1, x, y, z);
            const bool pass = a <= ld_f64(w, pc + 7);
            double v = a;
            if (COUNT) flops += 12;
            if (UNIFORM ? __any_sync(FULL, pass) : pass) {
                const double e = box_fold<COUNT>(w, pc + 8, MC_HDR_SMALL(h), x, y, z, flops);
                if (pass) v = e;
            }
            MC_PUSH(v);
            pc += 16;
            break;
        }
        case MC_OP_BOX: {
            const double e = box_fold<COUNT>(w, pc + 1, MC_HDR_SMALL(h), x, y, z, flops);
            MC_PUSH(e);
            pc += 9;
            break;
        }
        case MC_OP_GAPUSH:
        case MC_OP_GUARD: {
            const double a = box_distance(w, pc + 1, x, y, z);
            const bool pass = a <= ld_f64(w, pc + 7);
            const unsigned level = MC_HDR_LEVEL(h);
            if (COUNT) flops += 12;
            if (UNIFORM) {
                const unsigned b = __ballot_sync(FULL, pass);
                if (b == 0u) { MC_PUSH(a); pc = (int)MC_HDR_IDX(h); break; }
                if (level < 64u) {
                    if (b == FULL) allpass |= (1ull << level);
                    else allpass &= ~(1ull << level);
                }
            } else {
                if (!pass) { MC_PUSH(a); pc = (int)MC_HDR_IDX(h); break; }
            }
            pc += 8;
            if (MC_HDR_OP(h) == MC_OP_GUARD) break;
            pc -= 1;  // GAPUSH: the matrix follows the guard words; fall into APUSH with pc + 1 = first entry
        }
        // fallthrough
        case MC_OP_APUSH: {
            if (SMALL) { sx = x; sy = y; sz = z; }
            else { ps[psp] = x; ps[psp + 1] = y; ps[psp + 2] = z; psp += 3; }
            // nalgebra transform_point: ((m0*x + m1*y) + m2*z) + t   (n == 1 for affine matrices)
            const double nx = ((ld_f64(w, pc + 1) * x + ld_f64(w, pc + 2) * y) + ld_f64(w, pc + 3) * z) + ld_f64(w, pc + 4);
            const double ny = ((ld_f64(w, pc + 5) * x + ld_f64(w, pc + 6) * y) + ld_f64(w, pc + 7) * z) + ld_f64(w, pc + 8);
            const double nz = ((ld_f64(w, pc + 9) * x + ld_f64(w, pc + 10) * y) + ld_f64(w, pc + 11) * z) + ld_f64(w, pc + 12);
            x = nx; y = ny; z = nz;
            if (COUNT) flops += 18;
            pc += 13;
            break;
        }
        case MC_OP_ENDG: {
            // lanes that failed the guard take bbox.distance instead of the subtree value
            if (UNIFORM) {
                const unsigned level = MC_HDR_LEVEL(h);
                if (level >= 64u || !((allpass >> level) & 1ull)) {
                    const int g = (int)MC_HDR_IDX(h);
                    const double a = box_distance(w, g + 1, x, y, z);
                    if (!(a <= ld_f64(w, g + 7))) top = a;
                    if (COUNT) flops += 12;
                }
            }
            // non-uniform mode: only lanes that passed reach ENDG, nothing to select
            pc += 1;
            break;
        }
        case MC_OP_NEG:
            top = -top;
            if (COUNT) flops += 1;
            pc += 1;
            break;
        case MC_OP_CHILD:
            pc += 1;
            break;
        case MC_OP_BOXDIST:
            MC_PUSH(box_distance(w, pc + 1, x, y, z));
            if (COUNT) flops += 11;
            pc += 7;
            break;
        case MC_OP_SPHERE:
            MC_PUSH(sqrt((x * x + y * y) + z * z) - ld_f64(w, pc + 1));
            if (COUNT) flops += 7;
            pc += 2;
            break;
        case MC_OP_CYL:
            MC_PUSH(sqrt(x * x + y * y) - ld_f64(w, pc + 1));
            if (COUNT) flops += 5;
            pc += 2;
            break;
        case MC_OP_BOOL: {
            const unsigned sm = MC_HDR_SMALL(h);
            const int k = (int)(sm & 0x7fu);
            const double r = ld_f64(w, pc + 1);
            const double er = ld_f64(w, pc + 2);
            double res;
            if (SMALL) {
                // operands in push order: (.., s3, s2, s1, top)
                const double o0 = k == 2 ? s1 : (k == 3 ? s2 : s3);
                const double o1 = k == 2 ? top : (k == 3 ? s1 : s2);
                const double o2 = k == 3 ? top : s1;
                res = (sm & 0x80u) ? smooth_fold4<true, COUNT>(o0, o1, o2, top, k, r, er, flops)
                                   : smooth_fold4<false, COUNT>(o0, o1, o2, top, k, r, er, flops);
                if (k == 2) { s1 = s2; s2 = s3; }
                else if (k == 3) { s1 = s3; }
            } else {
                const int base = sp - k + 1;  // v[i] (i < k-1) is vs[base + i]
                res = (sm & 0x80u) ? smooth_fold<true, COUNT>(vs, base, k, top, r, er, flops)
                                   : smooth_fold<false, COUNT>(vs, base, k, top, r, er, flops);
                sp -= k - 1;
            }
            top = res;
            pc += 3;
            break;
        }
        case MC_OP_APOP: {
            if (SMALL) { x = sx; y = sy; z = sz; }
            else { psp -= 3; x = ps[psp]; y = ps[psp + 1]; z = ps[psp + 2]; }
            top = top * ld_f64(w, pc + 1);
            if (COUNT) flops += 1;
            pc += 2;
            break;
        }
        default:
            return __longlong_as_double(0x7ff8000000000000LL);
        }
    }
#undef MC_PUSH
}

template <bool UNIFORM>
__device__ __forceinline__ double vm_eval(const uint64_t* __restrict__ w, double x, double y, double z) {
    unsigned long long unused = 0ull;
    return vm_eval_impl<UNIFORM, false, false>(w, x, y, z, unused);
}

}  // namespace mc
