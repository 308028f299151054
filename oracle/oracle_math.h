// ORACLE — TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may build, load or call it, and only as the checker / the CPU baseline.
//
// exp / ln for the smooth booleans (implicit3d 0.14.2 `rvmin`/`rvmax`, un-vendored crate,
// reference call sites luascad/src/lib.rs:252-273).  The reference calls Rust's
// f64::exp / f64::ln, i.e. the platform libm.  Two modes:
//   libm_mode 0 : platform libm (what the reference binary would execute on this host)
//   libm_mode 1 : a portable restatement of the classic Sun-style argument-reduction +
//                 rational/polynomial kernels, written with + - * / only, no FMA, so that
//                 IEEE-754 guarantees the same bits on any host and on the device.
// Mode 1 is what the device implements; tests compare device == oracle(mode 1) bitwise and
// oracle(mode 0) vs oracle(mode 1) within the north-star tolerance (1e-12 relative).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

namespace orc {

static inline uint64_t d2u(double d) { uint64_t u; std::memcpy(&u, &d, 8); return u; }
static inline double u2d(uint64_t u) { double d; std::memcpy(&d, &u, 8); return d; }

// e^x, |error| < 1 ulp.  x = k*ln2 + r, |r| <= 0.5*ln2;  e^r = 1 + 2r/(2 - c(r)) with
// c(r) = r - r^2*(P1 + r^2*(P2 + ...)) (Remez fit of r*(e^r+1)/(e^r-1)).
static inline double portable_exp(double x) {
    const double LN2_HI = 6.93147180369123816490e-01;   // high 32 bits of ln2
    const double LN2_LO = 1.90821492927058770002e-10;   // ln2 - LN2_HI
    const double INV_LN2 = 1.44269504088896338700e+00;
    const double P1 = 1.66666666666666019037e-01;
    const double P2 = -2.77777777770155933842e-03;
    const double P3 = 6.61375632143793436117e-05;
    const double P4 = -1.65339022054652515390e-06;
    const double P5 = 4.13813679705723846039e-08;
    if (x != x) return x;
    if (x > 7.09782712893383973096e+02) return HUGE_VAL;
    if (x < -7.45133219101941108420e+02) return 0.0;
    double ax = std::fabs(x);
    double hi = x, lo = 0.0;
    int k = 0;
    if (ax > 0.34657359027997264 /* 0.5*ln2 */) {
        k = (int)(INV_LN2 * x + (x < 0.0 ? -0.5 : 0.5));
        double t = (double)k;
        hi = x - t * LN2_HI;
        lo = t * LN2_LO;
        x = hi - lo;
    } else if (ax < 3.725290298461914e-09 /* 2^-28 */) {
        return 1.0 + x;
    }
    double t = x * x;
    double c = x - t * (P1 + t * (P2 + t * (P3 + t * (P4 + t * P5))));
    if (k == 0) return 1.0 - ((x * c) / (c - 2.0) - x);
    double y = 1.0 - ((lo - (x * c) / (2.0 - c)) - hi);
    if (k >= -1021) {
        return u2d(d2u(y) + ((uint64_t)(int64_t)k << 52));
    }
    y = u2d(d2u(y) + ((uint64_t)(int64_t)(k + 1000) << 52));
    return y * 9.33263618503218878990e-302;  // 2^-1000
}

// ln(x), |error| < 1 ulp.  x = 2^k * (1+f), sqrt(2)/2 < 1+f < sqrt(2); s = f/(2+f);
// ln(1+f) = f - f^2/2 + s*(f^2/2 + R(s^2)).
static inline double portable_log(double x) {
    const double LN2_HI = 6.93147180369123816490e-01;
    const double LN2_LO = 1.90821492927058770002e-10;
    const double L1 = 6.666666666666735130e-01;
    const double L2 = 3.999999999940941908e-01;
    const double L3 = 2.857142874366239149e-01;
    const double L4 = 2.222219843214978396e-01;
    const double L5 = 1.818357216161805012e-01;
    const double L6 = 1.531383769920937332e-01;
    const double L7 = 1.479819860511658591e-01;
    if (x != x) return x;
    if (x == 0.0) return -HUGE_VAL;
    if (x < 0.0) return NAN;
    if (x == HUGE_VAL) return x;
    int k = 0;
    uint64_t ux = d2u(x);
    if ((ux >> 52) == 0) {  // subnormal: scale up by 2^54
        x = x * 18014398509481984.0;
        ux = d2u(x);
        k -= 54;
    }
    uint32_t hx = (uint32_t)(ux >> 32);
    k += (int)(hx >> 20) - 1023;
    hx &= 0x000fffffu;
    uint32_t i = (hx + 0x95f64u) & 0x100000u;  // 1+f >= sqrt(2) -> halve
    ux = ((uint64_t)(hx | (i ^ 0x3ff00000u)) << 32) | (ux & 0xffffffffull);
    x = u2d(ux);
    k += (int)(i >> 20);
    double f = x - 1.0;
    double dk = (double)k;
    if ((0x000fffffu & (2u + hx)) < 3u) {  // |f| < 2^-20
        if (f == 0.0) {
            if (k == 0) return 0.0;
            return dk * LN2_HI + dk * LN2_LO;
        }
        double R = f * f * (0.5 - 0.33333333333333333 * f);
        if (k == 0) return f - R;
        return dk * LN2_HI - ((R - dk * LN2_LO) - f);
    }
    double s = f / (2.0 + f);
    double z = s * s;
    double w = z * z;
    double t1 = w * (L2 + w * (L4 + w * L6));
    double t2 = z * (L1 + w * (L3 + w * (L5 + w * L7)));
    double R = t2 + t1;
    int32_t ii = (int32_t)hx - 0x6147a;
    int32_t jj = 0x6b851 - (int32_t)hx;
    if ((ii | jj) > 0) {
        double hfsq = 0.5 * f * f;
        if (k == 0) return f - (hfsq - s * (hfsq + R));
        return dk * LN2_HI - ((hfsq - (s * (hfsq + R) + dk * LN2_LO)) - f);
    }
    if (k == 0) return f - s * (f - R);
    return dk * LN2_HI - ((s * (f - R) - dk * LN2_LO) - f);
}

}  // namespace orc
