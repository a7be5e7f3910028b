// transport.cuh -- sm_100a photon-packet transport: device functions + kernels.
//
// Included by two translation units:
//   transport_faithful.cu  (-fmad=false, ARTES_FAITHFUL=1): reference operation order, no FMA
//                          contraction, sequential 180-bin CDFs -> bit-comparable with the oracle
//   transport_fast.cu      (-fmad=true,  ARTES_FAITHFUL=0): same physics; prefix-table CDF inversion
//
// Design: one persistent warp-synchronous state machine per lane.  Every lane always
// carries one photon; the tau pre-pass (src/ARTES.f90:633-656), both walks (:691-778, :850-941) and the
// three peel-off walks (:4542-4569, :4651-4673, :4739-4761) all advance through ONE shared inlined
// `cell_face` step per loop iteration, so a warp stays converged on the dominant work no matter which
// of the six walks each lane is in.  The rare, heavy events (emission, CDF sampling + Mueller update,
// detector deposit) are executed for the lanes that need them, optionally deferred until enough lanes
// of the warp wait for the same event (ballot regrouping).  Grid tables live in shared memory, the
// opacity / matrix tables in L2; the random stream is counter-based Philox4x32-10 keyed by photon id.
#pragma once

#include <cuda_runtime.h>
#include <cstdio>
#include <math_constants.h>
#include "kernel_args.h"

#ifndef ARTES_FAITHFUL
#error "define ARTES_FAITHFUL to 0 or 1"
#endif


namespace artes {
namespace ARTES_NS {

constexpr unsigned FULL = 0xffffffffu;
// pi = 4*atan(1) (src/ARTES.f90:9) is the correctly rounded double below.
constexpr double PI = 3.14159265358979323846;

enum Phase : int { PH_NEW = 0, PH_PRE, PH_WALK, PH_PEEL, PH_PEELDONE, PH_SCAT, PH_SCAT2, PH_IDLE, PH_LAMBERT,
                   PH_PREDONE, PH_SURFHIT, PH_RETIRE };
enum PeelKind : int { PK_SCATTER = 0, PK_SURFACE = 1, PK_THERMAL = 2 };

// ---------------------------------------------------------------------------------------------
// shared-memory table layout (doubles): rfront | thetafront | ttan | tcos | psin | pcos | phifront |
// sinbeta | cos2beta | sin2beta | tplane(int)
// ---------------------------------------------------------------------------------------------
struct SmLayout {
    int o_tf, o_tt, o_tc, o_ps, o_pc, o_pf, o_sb, o_c2, o_s2, o_tp, total;
    __host__ __device__ SmLayout(int nr, int nt, int np) {
        o_tf = nr + 1; o_tt = o_tf + nt + 1; o_tc = o_tt + nt + 1; o_ps = o_tc + nt + 1;
        o_pc = o_ps + np; o_pf = o_pc + np; o_sb = o_pf + np; o_c2 = o_sb + 180; o_s2 = o_c2 + 180;
        o_tp = o_s2 + 180; total = o_tp + (nt + 2) / 2 + 1;
    }
};

// ---------------------------------------------------------------------------------------------
// Philox4x32-10, counter = (photon id lo, hi, draw block, 0), key = seed.  xi = (u + 0.5) / 2^32.
// Replaces the per-thread Marsaglia-Zaman state of src/ARTES.f90:4197-4230.
// ---------------------------------------------------------------------------------------------
struct Rng {
    unsigned long long id;
    unsigned nd;
    unsigned b0, b1, b2, b3;
    bool exhausted;
};

// (kept out of line: ~100 integer instructions needed once per four draws at some twenty call sites)
__device__ __noinline__ uint4 philox4(unsigned long long id, unsigned blk, unsigned long long seed) {
    unsigned c0 = (unsigned)id, c1 = (unsigned)(id >> 32), c2 = blk, c3 = 0u;
    unsigned k0 = (unsigned)seed, k1 = (unsigned)(seed >> 32);
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        unsigned hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        unsigned hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        unsigned n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}
__device__ __forceinline__ void philox_block(unsigned long long id, unsigned blk, unsigned long long seed, Rng& r) {
    const uint4 v = philox4(id, blk, seed);
    r.b0 = v.x; r.b1 = v.y; r.b2 = v.z; r.b3 = v.w;
}

template <bool TRACE>
__device__ __forceinline__ double rng_next(Rng& r, const KernelArgs& A) {
    if (TRACE) {
        if ((int)r.nd >= A.R.max_draws) { r.exhausted = true; ++r.nd; return 0.5; }
        double v = A.R.xi[(size_t)(r.id - A.L.id_base) * A.R.max_draws + r.nd];
        ++r.nd;
        return v;
    }
    unsigned w = r.nd & 3u;
    if (w == 0u) philox_block(r.id, r.nd >> 2, A.L.seed, r);
    unsigned u = (w == 0u) ? r.b0 : (w == 1u) ? r.b1 : (w == 2u) ? r.b2 : r.b3;
    ++r.nd;
    return ((double)u + 0.5) * (1.0 / 4294967296.0);
}

// ---------------------------------------------------------------------------------------------
// geometry
// ---------------------------------------------------------------------------------------------
struct CellFace {
    int nf0, nf1, co0, co1, co2;
    double dist;
    bool exit;
    int err;
};

// quadratic_equation src/ARTES.f90:4154-4173 followed by the root choice of :2897-2907 (thr = 1e-15)
// or :2944-2954 (thr = 1e-3); `mirror` = +1/-1 applies the cone nappe rejection of :3036-3052.
__device__ __forceinline__ double solve_pick(double qa, double qb, double qc, double thr, int mirror, double z, double n2) {
    double s1 = 0.0, s2 = 0.0;
    double disc = qb * qb - 4.0 * qa * qc;
    if (disc >= 0.0) {
        double q = -0.5 * (qb + copysign(1.0, qb) * sqrt(disc));
        if (fabs(qa) > 1.e-100) s1 = q / qa;
        if (fabs(q) > 1.e-100) s2 = qc / q;
    }
    if (mirror != 0) {
        // mirror > 0: thetafront > pi/2 (reject z_test > 0); mirror < 0: thetafront < pi/2 (reject z_test < 0)
        if (s1 > 1.e-15) { double zt = z + s1 * n2; if ((zt > 0.0 && mirror > 0) || (zt < 0.0 && mirror < 0)) s1 = 0.0; }
        if (s2 > 1.e-15) { double zt = z + s2 * n2; if ((zt > 0.0 && mirror > 0) || (zt < 0.0 && mirror < 0)) s2 = 0.0; }
    }
    double d = 0.0;
    if (s1 > thr && s2 <= thr && s1 < 1.e100) d = s1;
    else if (s2 > thr && s1 <= thr && s2 < 1.e100) d = s2;
    else if (s1 > thr && s2 > thr) {
        if (s1 < 1.e100 && s1 < s2) d = s1;
        else if (s2 < 1.e100 && s2 < s1) d = s2;
    }
    return d;
}

// cell_face + next_cell, src/ARTES.f90:2800-3470 and :2671-2798.  The sub-expressions shared between
// the quadrics keep the reference's association, so the faithful build is bit-identical to evaluating
// every qa/qb/qc from scratch as the reference does.
__device__ __forceinline__ void cell_face(const double* __restrict__ sm, const SmLayout& lay, const DevTables& T,
                                          double x, double y, double z, double n0, double n1, double n2,
                                          int cf0, int cf1, int c0, int c1, int c2, CellFace& o) {
    const int* tplane = reinterpret_cast<const int*>(sm + lay.o_tp);
    const double a = 1.0 / T.ox, b = 1.0 / T.oy, c = 1.0 / T.oz;

    int f00 = c0, f01 = c0 + 1, f02 = -999;
    int f10 = c1, f11 = c1 + 1, f12 = -999;
    int f20 = c2, f21 = c2 + 1;
    if (f21 == T.np) f21 = 0;
    if (cf0 == 1) { f00 = cf1 - 1; f01 = cf1 + 1; f02 = cf1; }
    else if (cf0 == 2) { f10 = cf1 - 1; f11 = cf1 + 1; f12 = cf1; }
    else if (cf0 == 3) { f20 = (cf1 == 0) ? T.np - 1 : cf1 - 1; f21 = (cf1 == T.np - 1) ? 0 : cf1 + 1; }

    const double A1 = a * a * n0 * n0 + b * b * n1 * n1, A2 = c * c * n2 * n2;
    const double B1 = a * a * x * n0 + b * b * y * n1,   B2 = c * c * z * n2;
    const double C1 = a * a * x * x + b * b * y * y,     C2 = c * c * z * z;
    const double qa_s = A1 + A2, qb_s = 2.0 * (B1 + B2), qc_s = C1 + C2;

    // Which of the (up to) six quadric candidates exist for this lane: bit k of `act`,
    // k = 0..2 spheres (inner, outer, same), k = 3..5 cones / equatorial plane (inner, outer, same).
    unsigned act = 0u;
    if (cf0 == 1) {  // radial :2885-2964
        if (c0 == cf1 - 1) act |= 1u | 4u;
        else if (c0 == cf1) act |= 2u;
    } else act |= 1u | 2u;
    if (cf0 == 2) {  // polar :3020-3171
        if (c1 == cf1 - 1 && f10 != 0) act |= 8u;
        else if (c1 == cf1 && f11 != T.nt) act |= 16u;
        const double tf = sm[lay.o_tf + f12];
        if (((tf < PI / 2.0 && c1 == cf1 - 1) || (tf > PI / 2.0 && c1 == cf1)) && tplane[f12] == 1) act |= 32u;
    } else {
        if (f10 < 0 || f10 > T.nt) f10 = 0;  // error 029 (log only)
        if (f10 != 0) act |= 8u;
        if (f11 != T.nt) act |= 16u;
    }
    const unsigned long long packS = (unsigned long long)(unsigned)(f00 & 0xffff) | ((unsigned long long)(unsigned)(f01 & 0xffff) << 16) |
                                     ((unsigned long long)(unsigned)(f02 & 0xffff) << 32);
    const unsigned long long packC = (unsigned long long)(unsigned)(f10 & 0xffff) | ((unsigned long long)(unsigned)(f11 & 0xffff) << 16) |
                                     ((unsigned long long)(unsigned)(f12 & 0xffff) << 32);

    // Nearest face, src/ARTES.f90:3358-3418: the reference scans distance(i,j) with j (inner, outer, same)
    // outer and i (r, theta, phi) inner and a strict <, first over candidates > 1e-9 and, if there is none,
    // over candidates > 1e-12.  Equivalent: minimum by (distance, scan position ord = 3 j + i).
    double fd = 1.e100, fd12 = 1.e100;
    int sel = -1, sel12 = -1;       // ord | face << 4
    auto consider = [&](double d, int ord, int face) {
        if (d > 1.e-9 && (d < fd || (d == fd && ord < (sel & 15)))) { fd = d; sel = ord | (face << 4); }
        if (d > 1.e-12 && (d < fd12 || (d == fd12 && ord < (sel12 & 15)))) { fd12 = d; sel12 = ord | (face << 4); }
    };

    // One solver body for all quadrics.  Every lane walks through ITS OWN list of existing candidates
    // (bit scan of `act`), so the i-th trip of the loop solves the i-th candidate of every lane in one
    // converged instruction stream: ~4 trips with nearly all lanes busy instead of 6-8 inlined solver
    // copies that each run for a subset of the lanes.  The per-candidate arithmetic is unchanged.
    while (act) {
        const int k = __ffs(act) - 1;
        act &= act - 1u;
        const bool cone = k >= 3;
        const int j = cone ? k - 3 : k;
        const int f = (int)(((cone ? packC : packS) >> (16 * j)) & 0xffffull);
        const double v = sm[cone ? lay.o_tt + f : f];                // tan(theta_f) or r_f
        const double tf = sm[lay.o_tf + (cone ? f : 0)];
        const int tp = cone ? tplane[f] : 1;
        const double vv = v * v;
        int mir = 0;
        if (cone) mir = (tf > PI / 2.0) ? 1 : ((tf < PI / 2.0) ? -1 : 0);
        const double qa = cone ? A1 - A2 * v * v : qa_s;
        const double qb = cone ? 2.0 * (B1 - B2 * v * v) : qb_s;
        const double qc = cone ? C1 - C2 * v * v : qc_s - vv;
        double d;
        if (tp != 1) {  // equatorial plane :3068, :3118 (never a "same" candidate)
            d = 0.0;
            if (tp == 2) {
                if (j == 0) { if (-z / n2 > 0.0 && n2 > 1.e-15) d = -z / n2; }
                else if (j == 1) { if (-z / n2 > 0.0 && n2 < -1.e-15) d = -z / n2; }
            }
        } else d = solve_pick(qa, qb, qc, j == 2 ? 1.e-3 : 1.e-15, mir, z, n2);
        consider(d, 3 * j + (cone ? 1 : 0), f);
    }

    // ---- azimuthal :3292-3350 (full planes; guards as in the reference, including solutions_p(1) in the
    // test of solutions_p(2) and the missing a, b in the outward denominator check)
    {
        double sp1 = 0.0;
        bool do_in, do_out, plain_guard = false;
        if (cf0 == 3) {
            do_in = (c2 == cf1 - 1) || (c2 == T.np - 1 && cf1 == 0);
            do_out = (!do_in) && (c2 == cf1);
        } else { do_in = do_out = (T.np > 1); plain_guard = true; }
        if (do_in) {
            double ps = sm[lay.o_ps + f20], pc = sm[lay.o_pc + f20];
            double den = b * n1 * pc - a * n0 * ps;
            if (fabs(den) > 0.0) {
                sp1 = (a * x * ps - b * y * pc) / den;
                if (sp1 > 1.e-15 && sp1 < 1.e100) consider(sp1, 2, f20);
            }
        }
        if (do_out) {
            double ps = sm[lay.o_ps + f21], pc = sm[lay.o_pc + f21];
            double den = b * n1 * pc - a * n0 * ps;
            double guard = plain_guard ? (n1 * pc - n0 * ps) : den;
            if (fabs(guard) > 0.0) {
                double sp2 = (a * x * ps - b * y * pc) / den;
                if (sp2 > 1.e-15 && sp1 < 1.e100) consider(sp2, 5, f21);
            }
        }
    }
    if (sel < 0) { sel = sel12; fd = fd12; }
    const int li = (sel < 0) ? -1 : ((sel & 15) % 3);
    const int lf = (sel < 0) ? 0 : (sel >> 4);
    o.dist = fd; o.exit = false; o.err = 0;
    o.nf0 = 0; o.nf1 = 0; o.co0 = 0; o.co1 = 0; o.co2 = 0;
    if (li < 0) { o.err = 31; return; }
    o.nf0 = li + 1; o.nf1 = lf;

    // next_cell :2671-2798
    if (li == 0) {
        o.co1 = c1; o.co2 = c2;
        if (cf0 == 1 && lf == cf1) o.co0 = c0 + 1;
        else if (lf == c0) o.co0 = c0 - 1;
        else if (lf == c0 + 1) o.co0 = c0 + 1;
        else { o.co0 = 0; o.co1 = 0; o.co2 = 0; }  // error 022
    } else if (li == 1) {
        o.co0 = c0; o.co2 = c2;
        double tf = sm[lay.o_tf + lf];
        if (cf0 == 2 && lf == cf1 && tf < PI / 2.0) o.co1 = c1 + 1;
        else if (cf0 == 2 && lf == cf1 && tf > PI / 2.0) o.co1 = c1 - 1;
        else if (lf == c1) o.co1 = c1 - 1;
        else if (lf == c1 + 1) o.co1 = c1 + 1;
        else { o.co0 = 0; o.co1 = 0; o.co2 = 0; }  // error 023
    } else {
        o.co0 = c0; o.co1 = c1;
        if (c2 == T.np - 1 && lf == 0) o.co2 = 0;
        else if (c2 == 0 && lf == 0) o.co2 = T.np - 1;
        else if (lf == c2 + 1) o.co2 = c2 + 1;
        else if (lf == c2) o.co2 = c2 - 1;
        else { o.co0 = 0; o.co1 = 0; o.co2 = 0; }  // error 024
    }
    if (o.nf0 == 1 && o.nf1 == T.nr) o.exit = true;  // :3436
    if (cf0 == 1 && cf1 == T.cell_depth && o.nf0 == 1 && o.nf1 == T.cell_depth) o.err = 34;
    else if (o.co0 == T.nr && !o.exit) o.err = 35;
    else if (o.co1 == T.nt) o.err = 36;
    else if (c0 == o.co0 && c1 == o.co1 && c2 == o.co2 && !o.exit) o.err = 37;
}

// mueller_matrix_filler :1934-1960: returns c2p, s2p ((1,1)=(2,2)=c2p, (1,2)=s2p, (2,1)=-s2p)
__device__ __forceinline__ void mueller(double psi, double& c2p, double& s2p) {
    c2p = cos(2.0 * psi);
    s2p = sqrt(1.0 - c2p * c2p);
    if (psi > PI / 2.0 && psi < PI) s2p = -s2p;
    else if (psi > 3.0 * PI / 2.0 && psi < 2.0 * PI) s2p = -s2p;
    else if (psi > -PI / 2.0 && psi < 0.0) s2p = -s2p;
    else if (psi > -2.0 * PI && psi < -3.0 * PI / 2.0) s2p = -s2p;
}

// direction_cosine :1962-2052; returns 0 or the error code of an undefined result
__device__ __forceinline__ int direction_cosine(double alpha, double beta, double d0, double d1, double d2,
                                                double& e0, double& e1, double& e2) {
    double cto = d2 / sqrt(d0 * d0 + d1 * d1 + d2 * d2);
    double sto = sqrt(1.0 - cto * cto);
    double phi_old = atan2(d1, d0);
    if (phi_old < 0.0) phi_old = phi_old + 2.0 * PI;
    double ctn;
    if (beta >= PI && beta < 2.0 * PI) ctn = cto * alpha + sto * sqrt(1.0 - alpha * alpha) * cos(2.0 * PI - beta);
    else if (beta >= 0.0 && beta < PI) ctn = cto * alpha + sto * sqrt(1.0 - alpha * alpha) * cos(beta);
    else return 18;
    double stn = sqrt(1.0 - ctn * ctn);
    double nc = (alpha - ctn * cto) / (stn * sto);
    if (nc >= 1.0) nc = 1.0 - 1.e-10;
    else if (nc <= -1.0) nc = -1.0 + 1.e-10;
    double phi_new;
    if (fabs(nc) <= 1.0) {
        if (beta >= PI && beta < 2.0 * PI) phi_new = phi_old - acos(nc);
        else phi_new = phi_old + acos(nc);
    } else return 20;
    if (phi_new < 0.0) phi_new = phi_new + 2.0 * PI;
    if (phi_new > 2.0 * PI) phi_new = phi_new - 2.0 * PI;
    double cpn = cos(phi_new), spn;
    if (phi_new >= 0.0 && phi_new < PI) spn = sqrt(1.0 - cpn * cpn);
    else if (phi_new >= PI && phi_new <= 2.0 * PI) spn = -sqrt(1.0 - cpn * cpn);
    else return 21;
    e0 = stn * cpn; e1 = stn * spn; e2 = ctn;
    return 0;
}

// scatter matrix at angle acos_a: bracket + interpolation of :1448-1530 / :4780-4862
__device__ __forceinline__ void matrix_at(const DevTables& T, int cellidx, double acos_a, double F[16]) {
    const double deg = acos_a * 180.0 / PI;
    int lo, up;
    if (fmod(deg, 1.0) > 0.5) { up = (int)deg + 2; lo = (int)deg + 1; }
    else { up = (int)deg + 1; lo = (int)deg; }
    const double* base = T.M + (size_t)__ldg(T.c2u + cellidx) * (180 * 16);
    if (up == 1 || lo == 180) {
        const double2* m = reinterpret_cast<const double2*>(base + (up == 1 ? 0 : 179) * 16);
#pragma unroll
        for (int i = 0; i < 8; ++i) { double2 v = __ldg(m + i); F[2 * i] = v.x; F[2 * i + 1] = v.y; }
    } else {
        const double2* m0 = reinterpret_cast<const double2*>(base + (lo - 1) * 16);
        const double2* m1 = reinterpret_cast<const double2*>(base + (up - 1) * 16);
        const double y0 = (double)lo - 0.5, y1 = (double)up - 0.5;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            double2 v0 = __ldg(m0 + i), v1 = __ldg(m1 + i);
            F[2 * i] = (v1.x - v0.x) * (deg - y0) / (y1 - y0) + v0.x;
            F[2 * i + 1] = (v1.y - v0.y) * (deg - y0) / (y1 - y0) + v0.y;
        }
    }
}

// polarization_rotation :1663-1932.  Returns 0, or the code of an undefined output (11, 16);
// `soft` collects the log-only codes 12..15.
__device__ __forceinline__ int polarization_rotation(double alpha, double beta, const double Sin[4], const double F[16],
                                                     double d2, double e2, double Sout[4], bool peeling, int& soft) {
    double norm;
    if (fabs(alpha) < 1.0 && fabs(e2) < 1.0) {
        double nc2 = (d2 - e2 * alpha) / (sqrt(1.0 - alpha * alpha) * sqrt(1.0 - e2 * e2));
        double beta2;
        if (fabs(nc2) <= 1.0) beta2 = acos(nc2);
        else if (nc2 > 1.0 && nc2 < 1.00001) beta2 = 0.0;
        else if (nc2 < -1.0 && nc2 > -1.00001) beta2 = PI;
        else return 11;
        double c2, s2;
        mueller(beta, c2, s2);
        double r0 = Sin[0], r1 = c2 * Sin[1] + s2 * Sin[2], r2 = -s2 * Sin[1] + c2 * Sin[2], r3 = Sin[3];
        double den = sqrt(r1 * r1 + r2 * r2 + r3 * r3);
        if (den > 0.0) norm = sqrt(Sin[1] * Sin[1] + Sin[2] * Sin[2] + Sin[3] * Sin[3]) / den;
        else norm = 1.0;
        if (norm < 1.0 || norm > 1.0) { r1 = r1 * norm; r2 = r2 * norm; r3 = r3 * norm; }
        double s[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) s[r] = F[4 * r] * r0 + F[4 * r + 1] * r1 + F[4 * r + 2] * r2 + F[4 * r + 3] * r3;
        if (!peeling) {
            if (s[0] > 0.0) { norm = r0 / s[0]; s[0] = norm * s[0]; s[1] = norm * s[1]; s[2] = norm * s[2]; s[3] = norm * s[3]; }
            else soft = 12;
        }
        if (beta >= 0.0 && beta < PI) mueller(beta2, c2, s2);
        else if (beta >= PI && beta < 2.0 * PI) mueller(-beta2, c2, s2);
        Sout[0] = s[0];
        Sout[1] = c2 * s[1] + s2 * s[2];
        Sout[2] = -s2 * s[1] + c2 * s[2];
        Sout[3] = s[3];
        den = sqrt(Sout[1] * Sout[1] + Sout[2] * Sout[2] + Sout[3] * Sout[3]);
        if (den > 0.0) norm = sqrt(s[1] * s[1] + s[2] * s[2] + s[3] * s[3]) / den;
        else norm = 1.0;
        if (norm < 1.0 || norm > 1.0) { Sout[1] = Sout[1] * norm; Sout[2] = Sout[2] * norm; Sout[3] = Sout[3] * norm; }
        return 0;
    } else if (alpha >= 1.0 && alpha < 1.0001) {
        Sout[0] = Sin[0]; Sout[1] = Sin[1]; Sout[2] = Sin[2]; Sout[3] = Sin[3];
        soft = 13;
        return 0;
    } else if (alpha <= -1.0 && alpha > -1.0001) {
        double s[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) s[r] = F[4 * r] * Sin[0] + F[4 * r + 1] * Sin[1] + F[4 * r + 2] * Sin[2] + F[4 * r + 3] * Sin[3];
        soft = 15;
        if (peeling) { Sout[0] = s[0]; Sout[1] = s[1]; Sout[2] = s[2]; Sout[3] = s[3]; }
        else if (s[0] > 0.0) { norm = Sin[0] / s[0]; Sout[0] = norm * s[0]; Sout[1] = norm * s[1]; Sout[2] = norm * s[2]; Sout[3] = norm * s[3]; }
        else { Sout[0] = Sout[1] = Sout[2] = Sout[3] = 0.0; soft = 14; }
        return 0;
    }
    return 16;
}

// scattering_angle_sampling :1534-1661.
// Faithful: the two 180-bin running sums in the reference's order; the search pass recomputes the same
// partial sums instead of storing intensity_cumulative(0:180), which is bit-identical and keeps the
// per-thread array out of local memory.
// Fast: cum(i) is a 3-/4-term combination of host-built prefix tables -> binary search over 180 bins.
template <bool TRACE>
__device__ __forceinline__ int sample_angles(const double* __restrict__ sm, const SmLayout& lay, const KernelArgs& A,
                                             Rng& rng, const double S[4], int cellidx, double& alpha, double& beta) {
    const DevTables& T = A.T;
    const int u = __ldg(T.c2u + cellidx);
    const double p11 = __ldg(T.p1k + 4 * u), p12 = __ldg(T.p1k + 4 * u + 1), p13 = __ldg(T.p1k + 4 * u + 2), p14 = __ldg(T.p1k + 4 * u + 3);
#if ARTES_FAITHFUL
    const double* c2t = sm + lay.o_c2;
    const double* s2t = sm + lay.o_s2;
    double cum = 0.0;
    for (int i = 0; i < 180; ++i) {
        double v = p11 * S[0] + p12 * S[1] * c2t[i] + p12 * S[2] * s2t[i] - p13 * S[1] * s2t[i] + p13 * S[2] * c2t[i] + p14 * S[3];
        cum = cum + v;
    }
    double xi = rng_next<TRACE>(rng, A);
    double samp = xi * cum;
    bool found = false;
    double prev = 0.0;
    beta = 0.0;
    for (int i = 0; i < 180; ++i) {
        double v = p11 * S[0] + p12 * S[1] * c2t[i] + p12 * S[2] * s2t[i] - p13 * S[1] * s2t[i] + p13 * S[2] * c2t[i] + p14 * S[3];
        double cur = prev + v;
        if (samp >= prev && samp <= cur) {
            double x0 = (double)i, x1 = (double)(i + 1);
            beta = (x1 - x0) * (samp - prev) / (cur - prev) + x0;
            beta = beta * PI / 180.0;
            found = true;
            break;
        }
        prev = cur;
    }
#else
    const double Ac = p11 * S[0] + p14 * S[3], Bc = p12 * S[1] + p13 * S[2], Cc = p12 * S[2] - p13 * S[1];
    const double* pc2 = T.cdfA;
    const double* ps2 = T.cdfA + 181;
    auto cumA = [&](int i) { return Ac * (double)i + Bc * __ldg(pc2 + i) + Cc * __ldg(ps2 + i); };
    double xi = rng_next<TRACE>(rng, A);
    double samp = xi * cumA(180);
    int lo = 0, hi = 180;  // smallest i in 1..180 with cum(i) >= samp
    while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (cumA(mid) >= samp) hi = mid; else lo = mid; }
    bool found = true;
    {
        double y0 = cumA(hi - 1), y1 = cumA(hi);
        double fr = (samp - y0) / (y1 - y0);
        if (!(fr == fr)) found = false;  // NaN: degenerate CDF
        fr = fmin(fmax(fr, 0.0), 1.0);
        beta = (fr + (double)(hi - 1)) * (PI / 180.0);
    }
#endif
    xi = rng_next<TRACE>(rng, A);
    if (!found) return 6;
    if (xi > 0.5) beta = beta + PI;
    if (beta >= 2.0 * PI) beta = 2.0 * PI - 1.e-10;
    if (beta <= 0.0) beta = -2.0 * PI + 1.e-10;

    double c2b = cos(2.0 * beta);
    double s2b = sqrt(1.0 - c2b * c2b);
    if (beta > PI / 2.0 && beta < PI) s2b = -s2b;
    else if (beta > 3.0 * PI / 2.0 && beta < 2.0 * PI) s2b = -s2b;
    else if (beta > -PI / 2.0 && beta < 0.0) s2b = -s2b;
    else if (beta > -2.0 * PI && beta < -3.0 * PI / 2.0) s2b = -s2b;

#if ARTES_FAITHFUL
    const double* sbt = sm + lay.o_sb;
    const double2* row = reinterpret_cast<const double2*>(T.Mrow + (size_t)u * 720);
    cum = 0.0;
    for (int i = 0; i < 180; ++i) {
        double2 m01 = __ldg(row + 2 * i), m23 = __ldg(row + 2 * i + 1);
        double v = m01.x * S[0] + m01.y * c2b * S[1] + m01.y * s2b * S[2] - m23.x * s2b * S[1] + m23.x * c2b * S[2] + m23.y * S[3];
        v = v * sbt[i] * PI / 180.0;
        cum = cum + v;
    }
    xi = rng_next<TRACE>(rng, A);
    samp = xi * cum;
    found = false;
    prev = 0.0;
    alpha = 0.0;
    for (int i = 0; i < 180; ++i) {
        double2 m01 = __ldg(row + 2 * i), m23 = __ldg(row + 2 * i + 1);
        double v = m01.x * S[0] + m01.y * c2b * S[1] + m01.y * s2b * S[2] - m23.x * s2b * S[1] + m23.x * c2b * S[2] + m23.y * S[3];
        v = v * sbt[i] * PI / 180.0;
        double cur = prev + v;
        if (samp >= prev && samp <= cur) {
            double x0 = (double)i, x1 = (double)(i + 1);
            alpha = (x1 - x0) * (samp - prev) / (cur - prev) + x0;
            alpha = cos(alpha * PI / 180.0);
            found = true;
            break;
        }
        prev = cur;
    }
#else
    const double w1 = S[0], w2 = c2b * S[1] + s2b * S[2], w3 = c2b * S[2] - s2b * S[1], w4 = S[3];
    const double2* tab = reinterpret_cast<const double2*>(T.cdfP + (size_t)u * (181 * 4));
    auto cumP = [&](int i) {
        double2 q01 = __ldg(tab + 2 * i), q23 = __ldg(tab + 2 * i + 1);
        return w1 * q01.x + w2 * q01.y + w3 * q23.x + w4 * q23.y;
    };
    xi = rng_next<TRACE>(rng, A);
    samp = xi * cumP(180);
    lo = 0; hi = 180;
    while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (cumP(mid) >= samp) hi = mid; else lo = mid; }
    {
        double y0 = cumP(hi - 1), y1 = cumP(hi);
        double fr = (samp - y0) / (y1 - y0);
        found = (fr == fr);
        fr = fmin(fmax(fr, 0.0), 1.0);
        alpha = cos((fr + (double)(hi - 1)) * (PI / 180.0));
    }
#endif
    if (!found) return 7;
    if (alpha >= 1.0) alpha = 1.0 - 1.e-10;
    if (alpha <= -1.0) alpha = -1.0 + 1.e-10;
    return 0;
}


#if !ARTES_FAITHFUL
// ---------------------------------------------------------------------------------------------
// Fast-mode reciprocal, division and square root: MUFU seed (20 bits) + two Newton / Goldschmidt steps,
// without the IEEE special-case paths of div.rn.f64 / sqrt.rn.f64 (~25 / ~21 instructions each).  Results are
// within an ulp or two for normal operands; a zero or infinite operand gives NaN where IEEE gives inf/0, so
// callers guard those cases exactly where the reference's logic depends on them.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double frcp(double b) {
    double x;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(b));
    double e = fma(-b, x, 1.0);
    x = fma(x, e, x);
    e = fma(-b, x, 1.0);
    return fma(x, e, x);
}
// The quotient and the root end with a residual correction (q += (a - b q) x, g += (a - g^2) h) that squares the
// error once more, so ONE Newton step on the seed is enough in front of it: seed 2^-20 -> 2^-39 -> below the rounding.
__device__ __forceinline__ double fdiv(double a, double b) {
    double x;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(b));
    x = fma(x, fma(-b, x, 1.0), x);
    const double q = a * x;
    return fma(fma(-b, q, a), x, q);
}
__device__ __forceinline__ double fsqrt(double a) {   // a >= 0
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
    double g = a * y, h = 0.5 * y;
    const double r = fma(-g, h, 0.5);
    g = fma(g, r, g); h = fma(h, r, h);
    g = fma(fma(-g, g, a), h, g);
    return (a > 0.0) ? g : 0.0;
}
#endif

#if !ARTES_FAITHFUL
// ---------------------------------------------------------------------------------------------
// Fast-mode event math.  Same formulas as the reference's spherical trigonometry, with every
// acos/cos/atan2 round trip replaced by its algebraic identity:
//   mueller(psi) == (cos 2psi, sin 2psi)                       (src/ARTES.f90:1942-1953 is the sign of sin 2psi)
//   beta2 = acos(nc2)  ->  cos 2beta2 = 2 nc2^2 - 1, sin 2beta2 = 2 nc2 sqrt(1 - nc2^2)      (:1730-1735)
//   phi_new = phi_old +- acos(nc) -> rotation of (cos phi_old, sin phi_old) by (nc, +-sqrt(1-nc^2)) (:1999-2050)
//   the quadrant test of peel_photon (:4904-4914) == sign of (d x det)_z
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void matrix_at_deg(const DevTables& T, int cellidx, double deg, double F[16]) {
    int lo, up;
    const double fl = floor(deg);
    if (deg - fl > 0.5) { up = (int)fl + 2; lo = (int)fl + 1; }
    else { up = (int)fl + 1; lo = (int)fl; }
    const double* base = T.M + (size_t)__ldg(T.c2u + cellidx) * (180 * 16);
    if (up <= 1 || lo >= 180) {
        const double2* m = reinterpret_cast<const double2*>(base + (up <= 1 ? 0 : 179) * 16);
#pragma unroll
        for (int i = 0; i < 8; ++i) { double2 v = __ldg(m + i); F[2 * i] = v.x; F[2 * i + 1] = v.y; }
    } else {
        const double2* m0 = reinterpret_cast<const double2*>(base + (lo - 1) * 16);
        const double2* m1 = reinterpret_cast<const double2*>(base + (up - 1) * 16);
        const double w = deg - ((double)lo - 0.5);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            double2 v0 = __ldg(m0 + i), v1 = __ldg(m1 + i);
            F[2 * i] = (v1.x - v0.x) * w + v0.x;
            F[2 * i + 1] = (v1.y - v0.y) * w + v0.y;
        }
    }
}

// polarization_rotation with the rotation angles given by their double-angle cosines / sines:
// (c2a, s2a) = mueller(beta); nc2 = cos(beta2); flip = (beta >= pi) selects mueller(-beta2).
__device__ __forceinline__ int polrot_fast(double c2a, double s2a, bool flip, double nc2, const double Sin[4],
                                           const double F[16], double Sout[4], bool peeling, int& soft) {
    if (!(fabs(nc2) < 1.00001)) return 11;
    nc2 = fmin(fmax(nc2, -1.0), 1.0);
    double r0 = Sin[0], r1 = c2a * Sin[1] + s2a * Sin[2], r2 = c2a * Sin[2] - s2a * Sin[1], r3 = Sin[3];
    double s[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) s[r] = F[4 * r] * r0 + F[4 * r + 1] * r1 + F[4 * r + 2] * r2 + F[4 * r + 3] * r3;
    if (!peeling) {
        if (s[0] > 0.0) { double nrm = r0 / s[0]; s[0] = r0; s[1] *= nrm; s[2] *= nrm; s[3] *= nrm; }
        else soft = 12;
    }
    const double c2 = 2.0 * nc2 * nc2 - 1.0;
    double s2 = 2.0 * nc2 * sqrt(fmax(1.0 - nc2 * nc2, 0.0));
    if (flip) s2 = -s2;
    Sout[0] = s[0];
    Sout[1] = c2 * s[1] + s2 * s[2];
    Sout[2] = c2 * s[2] - s2 * s[1];
    Sout[3] = s[3];
    return 0;
}
#endif


#if !ARTES_FAITHFUL
struct FastAngles { double alpha, sT, deg, cb, sb; bool flip; };

// scattering_angle_sampling :1534-1661 in fast mode: cum(i) of both CDFs is a linear combination of
// host-built prefix tables, inverted by binary search; returns cos/sin of the sampled angles directly.
template <bool TRACE>
__device__ __forceinline__ int sample_angles_fast(const KernelArgs& A, Rng& rng, const double S[4], int cellidx, FastAngles& g) {
    const DevTables& T = A.T;
    const int u = __ldg(T.c2u + cellidx);
    const double p11 = __ldg(T.p1k + 4 * u), p12 = __ldg(T.p1k + 4 * u + 1), p13 = __ldg(T.p1k + 4 * u + 2), p14 = __ldg(T.p1k + 4 * u + 3);
    const double Ac = p11 * S[0] + p14 * S[3], Bc = p12 * S[1] + p13 * S[2], Cc = p12 * S[2] - p13 * S[1];
    const double* pc2 = T.cdfA;
    const double* ps2 = T.cdfA + 181;
    auto cumA = [&](int i) { return Ac * (double)i + Bc * __ldg(pc2 + i) + Cc * __ldg(ps2 + i); };
    double xi = rng_next<TRACE>(rng, A);
    double samp = xi * cumA(180);
    int lo = 0, hi = 180;  // smallest i in 1..180 with cum(i) >= samp
    double ylo = 0.0, yhi = cumA(180);
#pragma unroll 1
    while (hi - lo > 1) { int mid = (lo + hi) >> 1; double y = cumA(mid); if (y >= samp) { hi = mid; yhi = y; } else { lo = mid; ylo = y; } }
    double fr = (samp - ylo) / (yhi - ylo);
    if (!(fr == fr)) { rng_next<TRACE>(rng, A); return 6; }
    fr = fmin(fmax(fr, 0.0), 1.0);
    double beta = (fr + (double)lo) * (PI / 180.0);
    xi = rng_next<TRACE>(rng, A);
    sincos(beta, &g.sb, &g.cb);
    g.flip = xi > 0.5;                       // beta + pi  (:1589-1590)
    if (g.flip) { g.sb = -g.sb; g.cb = -g.cb; }
    const double c2b = g.cb * g.cb - g.sb * g.sb, s2b = 2.0 * g.sb * g.cb;
    const double w1 = S[0], w2 = c2b * S[1] + s2b * S[2], w3 = c2b * S[2] - s2b * S[1], w4 = S[3];
    const double2* tab = reinterpret_cast<const double2*>(T.cdfP + (size_t)u * (181 * 4));
    auto cumP = [&](int i) {
        double2 q01 = __ldg(tab + 2 * i), q23 = __ldg(tab + 2 * i + 1);
        return w1 * q01.x + w2 * q01.y + w3 * q23.x + w4 * q23.y;
    };
    xi = rng_next<TRACE>(rng, A);
    yhi = cumP(180); ylo = 0.0;
    samp = xi * yhi;
    lo = 0; hi = 180;
#pragma unroll 1
    while (hi - lo > 1) { int mid = (lo + hi) >> 1; double y = cumP(mid); if (y >= samp) { hi = mid; yhi = y; } else { lo = mid; ylo = y; } }
    fr = (samp - ylo) / (yhi - ylo);
    if (!(fr == fr)) return 7;
    fr = fmin(fmax(fr, 0.0), 1.0);
    g.deg = fr + (double)lo;
    sincos(g.deg * (PI / 180.0), &g.sT, &g.alpha);
    if (g.alpha >= 1.0) { g.alpha = 1.0 - 1.e-10; g.sT = sqrt(1.0 - g.alpha * g.alpha); }
    if (g.alpha <= -1.0) { g.alpha = -1.0 + 1.e-10; g.sT = sqrt(1.0 - g.alpha * g.alpha); }
    return 0;
}
#endif

// initial_cell :2605-2669
__device__ __forceinline__ void initial_cell(const double* __restrict__ sm, const SmLayout& lay, const DevTables& T,
                                             double x, double y, double z, int& c0, int& c1, int& c2) {
    double r = sqrt(x * x + y * y + z * z);
    double theta = acos(z / r);
    double phi = atan2(y, x);
    if (phi < 0.0) phi = phi + 2.0 * PI;
    c0 = T.nr - 1; c1 = 0; c2 = 0;
    for (int j = 0; j < T.nt; ++j)
        if (theta > sm[lay.o_tf + j] && theta < sm[lay.o_tf + j + 1]) { c1 = j; break; }
    for (int j = 0; j < T.np; ++j) {
        double hi = (j < T.np - 1) ? sm[lay.o_pf + j + 1] : 2.0 * PI;
        if (phi > sm[lay.o_pf + j] && phi < hi) { c2 = j; break; }
    }
}

__device__ __forceinline__ void tuple_hash(unsigned long long& h, int v) { h ^= (unsigned)v; h *= 1099511628211ull; }

// ---------------------------------------------------------------------------------------------
// photon state, event bodies, engines
// ---------------------------------------------------------------------------------------------
#include "engine.cuh"
#include "engine3.cuh"
#if !ARTES_FAITHFUL
#include "engine2.cuh"
#endif

// ---------------------------------------------------------------------------------------------
// isolated cell_face evaluations (unit-test hook)
// ---------------------------------------------------------------------------------------------
__global__ void cell_face_kernel(DevTables T, unsigned long long n, const double* __restrict__ pos, const double* __restrict__ dir,
                                 const int* __restrict__ face, const int* __restrict__ cell, int* __restrict__ out_i,
                                 double* __restrict__ out_d) {
    extern __shared__ double sm[];
    const SmLayout lay(T.nr, T.nt, T.np);
    for (int i = threadIdx.x; i <= T.nr; i += blockDim.x) sm[i] = T.rfront[i];
    for (int i = threadIdx.x; i <= T.nt; i += blockDim.x) {
        sm[lay.o_tf + i] = T.thetafront[i]; sm[lay.o_tt + i] = T.ttan[i]; sm[lay.o_tc + i] = T.tcos[i];
        reinterpret_cast<int*>(sm + lay.o_tp)[i] = T.tplane[i];
    }
    for (int i = threadIdx.x; i < T.np; i += blockDim.x) { sm[lay.o_ps + i] = T.psin[i]; sm[lay.o_pc + i] = T.pcos[i]; sm[lay.o_pf + i] = T.phifront[i]; }
    __syncthreads();
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * blockDim.x) {
        CellFace o;
        cell_face(sm, lay, T, pos[3 * i], pos[3 * i + 1], pos[3 * i + 2], dir[3 * i], dir[3 * i + 1], dir[3 * i + 2],
                  face[2 * i], face[2 * i + 1], cell[3 * i], cell[3 * i + 1], cell[3 * i + 2], o);
        int* p = out_i + 7 * i;
        p[0] = o.nf0; p[1] = o.nf1; p[2] = o.co0; p[3] = o.co1; p[4] = o.co2; p[5] = o.exit ? 1 : 0; p[6] = o.err;
        out_d[i] = o.dist;
    }
}

}  // namespace ARTES_NS
}  // namespace artes
