// Host check of artes_b200/csrc/fastmath.cuh against glibc: g++ -O2 -ffp-contract=off tools/fastmath_check.cc -o /tmp/fmc && /tmp/fmc
#include <cmath>
#include <cstring>
#include <cstdio>
#include <random>
#include "../artes_b200/csrc/fastmath.cuh"
static double ulps(double a, double b) { if (a == b) return 0; double u = std::fabs(b) > 0 ? std::ldexp(1.0, std::ilogb(b) - 52) : 5e-324; return std::fabs(a - b) / u; }
int main() {
    std::mt19937_64 g(1); std::uniform_real_distribution<double> U(0.0, 1.0);
    double ms = 0, mc = 0, ma = 0, ml = 0, me = 0, mabs = 0;
    for (int i = 0; i < 4000000; ++i) {
        double x = U(g) * M_PI; double s, c; fm_sincos_0pi(x, &s, &c);
        ms = std::fmax(ms, ulps(s, std::sin(x))); mc = std::fmax(mc, ulps(c, std::cos(x)));
        mabs = std::fmax(mabs, std::fmax(std::fabs(s - std::sin(x)), std::fabs(c - std::cos(x))));
        double v = 2.0 * U(g) - 1.0; if (i % 7 == 0) v = (v > 0 ? 1 : -1) * (1.0 - 1e-10 * U(g));
        ma = std::fmax(ma, ulps(fm_acos(v), std::acos(v)));
        double l = U(g); if (i % 5 == 0) l = std::pow(10.0, -10 * U(g)); if (l <= 0) l = 1e-10;
        ml = std::fmax(ml, (std::log(l) == 0.0) ? std::fabs(fm_log(l)) * 1e16 : ulps(fm_log(l), std::log(l)));
        double e = 50.0 * U(g); if (i % 3 == 0) e = U(g) * 1e-3;
        me = std::fmax(me, ulps(fm_exp_neg(e), std::exp(-e)));
    }
    double s, c; fm_sincos_0pi(0.0, &s, &c); printf("sincos(0) %g %g; ", s, c); fm_sincos_0pi(M_PI, &s, &c); printf("sincos(pi) %g %g\n", s, c);
    printf("max ulp: sin %.2f cos %.2f (abs %.2e) acos %.2f log %.2f exp %.2f; log(1)=%g exp(0)=%g acos(0)=%.17g\n", ms, mc, mabs, ma, ml, me, fm_log(1.0), fm_exp_neg(0.0), fm_acos(0.0));
}
