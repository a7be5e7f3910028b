// fastmath.cuh -- fast-mode elementary functions on the argument ranges the transport path uses.
//
// sin/cos on [0, pi], acos on (-1, 1), log on (0, 1], exp(-x) on [0, 700]: the ranges of the scattering angles, the
// peel-off angle, tau = -log(1 - xi) and the attenuation e^-tau.  Without the special-case and huge-argument paths
// of the CUDA math library they are 40-50 instructions each instead of 80-256, which matters twice in the
// interaction event: fewer instructions, and a third less code for the instruction cache to stream (ncu:
// no_instruction was the top stall).  Kernels and coefficients are fdlibm's (k_sin.c, k_cos.c, e_acos.c, e_log.c,
// e_exp.c; Sun Microsystems, freely distributable); validated on the host against glibc over 4e6 random arguments
// per function: <= 1 ulp (tools/fastmath_check.cc).
#pragma once
#ifdef __CUDACC__
#define FM_HD __device__ __forceinline__
#define FM_DIV(a, b) fdiv(a, b)          // transport.cuh: MUFU seed + Newton steps
#define FM_SQRT(a) fsqrt(a)
FM_HD long long fm_d2ll(double x) { return __double_as_longlong(x); }
FM_HD double fm_ll2d(long long v) { return __longlong_as_double(v); }
#else
#include <cmath>
#include <cstring>
#define FM_HD inline
#define FM_DIV(a, b) ((a) / (b))
#define FM_SQRT(a) std::sqrt(a)
FM_HD long long fm_d2ll(double x) { long long v; std::memcpy(&v, &x, 8); return v; }
FM_HD double fm_ll2d(long long v) { double x; std::memcpy(&x, &v, 8); return x; }
#endif

// sin and cos of x in [0, pi]
FM_HD void fm_sincos_0pi(double x, double* s, double* c) {
    const int n = (int)(x * 6.36619772367581382433e-01 + 0.5);          // 0, 1 or 2 quarter turns
    double r = fma(-(double)n, 1.57079632679489655800e+00, x);
    r = fma(-(double)n, 6.12323399573676603587e-17, r);                 // |r| <= pi/4
    const double z = r * r;
    const double ps = -1.66666666666666324348e-01 + z * (8.33333333332248946124e-03 + z * (-1.98412698298579493134e-04 +
                      z * (2.75573137070700676789e-06 + z * (-2.50507602534068634195e-08 + z * 1.58969099521155010221e-10))));
    const double sr = fma(z * r, ps, r);
    const double pc = 4.16666666666666019037e-02 + z * (-1.38888888888741095749e-03 + z * (2.48015872894767294178e-05 +
                      z * (-2.75573143513906633035e-07 + z * (2.08757232129817482790e-09 + z * -1.13596475577881948265e-11))));
    const double hz = 0.5 * z, w = 1.0 - hz;
    const double cr = w + (((1.0 - w) - hz) + z * z * pc);
    *s = (n == 1) ? cr : ((n == 2) ? -sr : sr);
    *c = (n == 1) ? -sr : ((n == 2) ? -cr : cr);
}

// acos(x), |x| < 1
FM_HD double fm_acos(double x) {
    const double a = fabs(x);
    const bool small = a < 0.5;
    const double z = small ? x * x : 0.5 * (1.0 - a);
    const double p = z * (1.66666666666666657415e-01 + z * (-3.25565818622400915405e-01 + z * (2.01212532134862925881e-01 +
                     z * (-4.00555345006794114027e-02 + z * (7.91534994289814532176e-04 + z * 3.47933107596021167570e-05)))));
    const double q = 1.0 + z * (-2.40339491173441421878e+00 + z * (2.02094576023350569471e+00 + z * (-6.88283971605453293030e-01 +
                     z * 7.70381505559019352791e-02)));
    const double r = FM_DIV(p, q);
    if (small) return 1.57079632679489655800e+00 - (x - (6.12323399573676603587e-17 - x * r));
    const double s = FM_SQRT(z);
    const double t = 2.0 * (s + s * r);
    return (x > 0.0) ? t : 3.14159265358979311600e+00 - t;
}

// log(x), 1e-300 < x <= 1 (normal numbers)
FM_HD double fm_log(double x) {
    long long b = fm_d2ll(x);
    int hx = (int)(b >> 32);
    int k = (hx >> 20) - 1023;
    hx &= 0x000fffff;
    const int i = (hx + 0x95f64) & 0x100000;
    b = (b & 0xffffffffll) | ((long long)(hx | (i ^ 0x3ff00000)) << 32);
    k += (i >> 20);
    const double f = fm_ll2d(b) - 1.0;
    const double s = FM_DIV(f, 2.0 + f);
    const double z = s * s, w = z * z;
    const double t1 = w * (3.999999999940941908e-01 + w * (2.222219843214978396e-01 + w * 1.531383769920937332e-01));
    const double t2 = z * (6.666666666666735130e-01 + w * (2.857142874366239149e-01 + w * (1.818357216161805012e-01 + w * 1.479819860511658591e-01)));
    const double R = t2 + t1, hfsq = 0.5 * f * f, dk = (double)k;
    return dk * 6.93147180369123816490e-01 - ((hfsq - (s * (hfsq + R) + dk * 1.90821492927058770002e-10)) - f);
}

// exp(-x), 0 <= x <= 700
FM_HD double fm_exp_neg(double x) {
    const double y = -x;
    const int k = (int)(1.44269504088896338700e+00 * y - 0.5);
    const double hi = fma(-(double)k, 6.93147180369123816490e-01, y), lo = (double)k * 1.90821492927058770002e-10;
    const double r = hi - lo;
    const double t = r * r;
    const double c = r - t * (1.66666666666666019037e-01 + t * (-2.77777777770155933842e-03 + t * (6.61375632143793436117e-05 +
                     t * (-1.65339022054652515390e-06 + t * 4.13813679705723846039e-08))));
    const double e = 1.0 - ((lo - FM_DIV(r * c, 2.0 - c)) - hi);
    return fm_ll2d(fm_d2ll(e) + ((long long)k << 52));
}
