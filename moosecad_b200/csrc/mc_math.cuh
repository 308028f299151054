// exp / ln for the smooth booleans (implicit3d `rvmin` / `rvmax`), device side.
//
// The reference calls the host libm; CUDA's exp()/log() are <= 1 ulp but not the same
// bits.  To make the SDF field reproducible bit for bit on any host and on the device, both
// are evaluated with the classic argument-reduction + rational kernel scheme using only
// IEEE-754 + - * / (this file is compiled with --fmad=false, so nothing is contracted).
// Accuracy: < 1 ulp (measured against mpmath in tests/test_math_accuracy.py).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mc {

static __device__ __noinline__ double mc_exp(double x) {
    const double LN2_HI = 6.93147180369123816490e-01;
    const double LN2_LO = 1.90821492927058770002e-10;
    const double INV_LN2 = 1.44269504088896338700e+00;
    const double P1 = 1.66666666666666019037e-01;
    const double P2 = -2.77777777770155933842e-03;
    const double P3 = 6.61375632143793436117e-05;
    const double P4 = -1.65339022054652515390e-06;
    const double P5 = 4.13813679705723846039e-08;
    if (x != x) return x;
    if (x > 7.09782712893383973096e+02) return __longlong_as_double(0x7ff0000000000000LL);
    if (x < -7.45133219101941108420e+02) return 0.0;
    const double ax = fabs(x);
    double hi = x, lo = 0.0;
    int k = 0;
    if (ax > 0.34657359027997264) {
        k = (int)(INV_LN2 * x + (x < 0.0 ? -0.5 : 0.5));
        const double t = (double)k;
        hi = x - t * LN2_HI;
        lo = t * LN2_LO;
        x = hi - lo;
    } else if (ax < 3.725290298461914e-09) {
        return 1.0 + x;
    }
    const double t = x * x;
    const double c = x - t * (P1 + t * (P2 + t * (P3 + t * (P4 + t * P5))));
    if (k == 0) return 1.0 - ((x * c) / (c - 2.0) - x);
    double y = 1.0 - ((lo - (x * c) / (2.0 - c)) - hi);
    if (k >= -1021) {
        return __longlong_as_double(__double_as_longlong(y) + ((long long)k << 52));
    }
    y = __longlong_as_double(__double_as_longlong(y) + ((long long)(k + 1000) << 52));
    return y * 9.33263618503218878990e-302;
}

static __device__ __noinline__ double mc_log(double x) {
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
    if (x == 0.0) return __longlong_as_double(0xfff0000000000000LL);
    if (x < 0.0) return __longlong_as_double(0x7ff8000000000000LL);
    if (x == __longlong_as_double(0x7ff0000000000000LL)) return x;
    int k = 0;
    unsigned long long ux = (unsigned long long)__double_as_longlong(x);
    if ((ux >> 52) == 0) {
        x = x * 18014398509481984.0;
        ux = (unsigned long long)__double_as_longlong(x);
        k -= 54;
    }
    unsigned int hx = (unsigned int)(ux >> 32);
    k += (int)(hx >> 20) - 1023;
    hx &= 0x000fffffu;
    const unsigned int i = (hx + 0x95f64u) & 0x100000u;
    ux = ((unsigned long long)(hx | (i ^ 0x3ff00000u)) << 32) | (ux & 0xffffffffull);
    x = __longlong_as_double((long long)ux);
    k += (int)(i >> 20);
    const double f = x - 1.0;
    const double dk = (double)k;
    if ((0x000fffffu & (2u + hx)) < 3u) {
        if (f == 0.0) {
            if (k == 0) return 0.0;
            return dk * LN2_HI + dk * LN2_LO;
        }
        const double R = f * f * (0.5 - 0.33333333333333333 * f);
        if (k == 0) return f - R;
        return dk * LN2_HI - ((R - dk * LN2_LO) - f);
    }
    const double s = f / (2.0 + f);
    const double z = s * s;
    const double w = z * z;
    const double t1 = w * (L2 + w * (L4 + w * L6));
    const double t2 = z * (L1 + w * (L3 + w * (L5 + w * L7)));
    const double R = t2 + t1;
    const int ii = (int)hx - 0x6147a;
    const int jj = 0x6b851 - (int)hx;
    if ((ii | jj) > 0) {
        const double hfsq = 0.5 * f * f;
        if (k == 0) return f - (hfsq - s * (hfsq + R));
        return dk * LN2_HI - ((hfsq - (s * (hfsq + R) + dk * LN2_LO)) - f);
    }
    if (k == 0) return f - s * (f - R);
    return dk * LN2_HI - ((s * (f - R) - dk * LN2_LO) - f);
}

}  // namespace mc
