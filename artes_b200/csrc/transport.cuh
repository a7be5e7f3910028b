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

enum Phase : int { PH_NEW = 0, PH_PRE, PH_WALK, PH_PEEL, PH_PEELDONE, PH_SCAT, PH_SCAT2, PH_IDLE };
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

__device__ __forceinline__ void philox_block(unsigned long long id, unsigned blk, unsigned long long seed, Rng& r) {
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
    r.b0 = c0; r.b1 = c1; r.b2 = c2; r.b3 = c3;
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

    // One solver body for all quadrics: the loop keeps the warp converged (every lane solves its k-th
    // candidate in the same instruction stream) and keeps the kernel small enough for the instruction cache.
#pragma unroll 1
    for (int k = 0; k < 6; ++k) {
        if (!((act >> k) & 1u)) continue;
        const bool cone = k >= 3;
        const int j = cone ? k - 3 : k;
        const int f = (int)(((cone ? packC : packS) >> (16 * j)) & 0xffffull);
        double qa = qa_s, qb = qb_s, qc;
        int mir = 0;
        double d;
        if (!cone) { const double r = sm[f]; qc = qc_s - r * r; }
        else {
            const double t = sm[lay.o_tt + f], tf = sm[lay.o_tf + f];
            mir = (tf > PI / 2.0) ? 1 : ((tf < PI / 2.0) ? -1 : 0);
            qa = A1 - A2 * t * t; qb = 2.0 * (B1 - B2 * t * t); qc = C1 - C2 * t * t;
        }
        if (cone && tplane[f] != 1) {  // equatorial plane :3068, :3118 (never a "same" candidate)
            d = 0.0;
            if (tplane[f] == 2) {
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
// the transport kernel
// ---------------------------------------------------------------------------------------------
// Regrouping thresholds are launch parameters (LaunchArgs::defer_events / defer_refill): the number of
// lanes of a warp that must wait for the heavy events / for new photons before those code paths run.

template <bool TRACE>
__global__ void __launch_bounds__(128, 4) transport_kernel(const __grid_constant__ KernelArgs A) {
    extern __shared__ double sm[];
    const DevTables& T = A.T;
    const LaunchArgs& L = A.L;
    const SmLayout lay(T.nr, T.nt, T.np);
    {   // stage the grid tables (coalesced)
        for (int i = threadIdx.x; i <= T.nr; i += blockDim.x) sm[i] = T.rfront[i];
        for (int i = threadIdx.x; i <= T.nt; i += blockDim.x) {
            sm[lay.o_tf + i] = T.thetafront[i]; sm[lay.o_tt + i] = T.ttan[i]; sm[lay.o_tc + i] = T.tcos[i];
            reinterpret_cast<int*>(sm + lay.o_tp)[i] = T.tplane[i];
        }
        for (int i = threadIdx.x; i < T.np; i += blockDim.x) {
            sm[lay.o_ps + i] = T.psin[i]; sm[lay.o_pc + i] = T.pcos[i]; sm[lay.o_pf + i] = T.phifront[i];
        }
        for (int i = threadIdx.x; i < 540; i += blockDim.x) sm[lay.o_sb + i] = T.trig[i];
        __syncthreads();
    }
    const int lane = threadIdx.x & 31;

    // ---- per-lane photon state
    Rng rng; rng.id = 0; rng.nd = 0; rng.b0 = rng.b1 = rng.b2 = rng.b3 = 0; rng.exhausted = false;
    int ph = PH_NEW, pk = PK_SCATTER;
    double px = 0, py = 0, pz = 0;          // photon position ("home" while a probe walk runs)
    double dx = 0, dy = 0, dz = 0;          // photon direction
    double S[4] = {0, 0, 0, 0};             // Stokes vector
    int c0 = 0, c1 = 0, c2 = 0, f0 = 0, f1 = 0;
    double tau = 0, tau_run = 0;
    double wx = 0, wy = 0, wz = 0;          // walker (the point cell_face is evaluated at)
    int wc0 = 0, wc1 = 0, wc2 = 0, wf0 = 0, wf1 = 0;
    double tacc = 0;                        // optical depth of the running probe walk (pre-pass / peel)
    bool peel_exit = false;
    // counters
    unsigned long long n_cf = 0;
    unsigned n_emit = 0, n_sc = 0, n_peel = 0, n_surf = 0, n_err = 0, n_draw = 0;
    // trace
    int t_len = 0, t_nsc = 0;
    unsigned long long t_hash = 1469598103934665603ull;

    auto err_count = [&](int code) { atomicAdd(A.O.err + code, 1ull); };
    auto record = [&](int a, int b, int c, int d, int e) {
        if (TRACE) {
            if (A.R.seq_head && t_len < A.R.max_rec) {
                int* p = A.R.seq_head + ((size_t)(rng.id - L.id_base) * A.R.max_rec + t_len) * 5;
                p[0] = a; p[1] = b; p[2] = c; p[3] = d; p[4] = e;
            }
            tuple_hash(t_hash, a); tuple_hash(t_hash, b); tuple_hash(t_hash, c); tuple_hash(t_hash, d); tuple_hash(t_hash, e);
            ++t_len;
        }
    };
    auto retire = [&]() {  // photon finished: publish trace record, ask for a new one
        if (TRACE) {
            size_t k = (size_t)(rng.id - L.id_base);
            A.R.seq_len[k] = t_len; A.R.seq_hash[k] = t_hash;
            if (A.R.fstate) {
                double* f = A.R.fstate + k * 8;
                const bool live = (ph == PH_WALK);
                f[0] = live ? wx : px; f[1] = live ? wy : py; f[2] = live ? wz : pz;
                f[3] = S[0]; f[4] = S[1]; f[5] = S[2]; f[6] = S[3]; f[7] = (double)t_nsc;
            }
        }
        n_draw += rng.nd;
        ph = PH_NEW;
    };
    // detector deposit :4947-4972 / :4575-4585 / :4683-4693
    auto deposit = [&](double W0, double W1, double W2, double W3, bool all4) {
        double x_im = py * L.cos_dp - px * L.sin_dp;
        double y_im = pz * L.sin_dt - py * L.cos_dt * L.sin_dp - px * L.cos_dt * L.cos_dp;
        int ix = (int)(L.nx * (x_im + L.x_max) / (2.0 * L.x_max)) + 1;
        int iy = (int)(L.ny * (y_im + L.y_max) / (2.0 * L.y_max)) + 1;
        if (ix < 1 || ix > L.nx || iy < 1 || iy > L.ny) { err_count(60); return; }
        record(100, ix, iy, 0, 0);
        const size_t npx = (size_t)L.nx * L.ny;
        double* d = A.O.det + (size_t)(ix - 1) + (size_t)L.nx * (iy - 1);
        atomicAdd(d, W0); atomicAdd(d + 4 * npx, W0 * W0); atomicAdd(d + 8 * npx, 1.0);
        if (all4) {
            atomicAdd(d + npx, W1); atomicAdd(d + 2 * npx, W2); atomicAdd(d + 3 * npx, W3);
            atomicAdd(d + 5 * npx, W1 * W1); atomicAdd(d + 6 * npx, W2 * W2); atomicAdd(d + 7 * npx, W3 * W3);
            atomicAdd(d + 9 * npx, 1.0);
        }
    };
    auto start_probe = [&](int r_shift) { wx = px; wy = py; wz = pz; wc0 = c0 + r_shift; wc1 = c1; wc2 = c2; wf0 = f0; wf1 = f1; tacc = 0.0; };

    for (;;) {
        // ================= A. refill + emission (emit_photon :1008-1268) =================
        const unsigned need0 = __ballot_sync(FULL, ph == PH_NEW);
        const unsigned walk0 = __ballot_sync(FULL, ph == PH_PRE || ph == PH_WALK || ph == PH_PEEL);
        const unsigned need = (__popc(need0) >= L.defer_refill || walk0 == 0u) ? need0 : 0u;
        if (need) {
            const int leader = __ffs(need) - 1;
            unsigned long long base = 0;
            if (lane == leader) base = atomicAdd(A.O.counter, (unsigned long long)__popc(need));
            base = __shfl_sync(FULL, base, leader);
            if (ph == PH_NEW) {
                const unsigned long long k = base + (unsigned long long)__popc(need & ((1u << lane) - 1u));
                if (k >= L.n_photons) ph = PH_IDLE;
                else {
                    rng.id = L.id_base + k; rng.nd = 0; rng.exhausted = false;
                    t_len = 0; t_nsc = 0; t_hash = 1469598103934665603ull;
                    ++n_emit;
                    S[0] = 1.0; S[1] = 0.0; S[2] = 0.0; S[3] = 0.0;
                    int e = 0;
                    double bias_weight = 1.0;
                    if (L.photon_source == 1) {
                        f0 = 1; f1 = T.nr;
                        double xi, r_disk;
                        if (L.limb_emission) {
                            for (;;) { xi = rng_next<TRACE>(rng, A); r_disk = sqrt(xi); if (r_disk > 0.9 || rng.exhausted) break; }
                        } else { xi = rng_next<TRACE>(rng, A); r_disk = sqrt(xi); }
                        xi = rng_next<TRACE>(rng, A);
                        const double phi_disk = 2.0 * PI * xi;
                        const double R = sm[T.nr];
                        const double d1 = R * r_disk * sin(phi_disk);
                        const double d2 = R * r_disk * cos(phi_disk);
                        dx = -1.0; dy = 0.0; dz = 0.0;
                        px = sqrt(R * R - d1 * d1 - d2 * d2); py = d1; pz = d2;
                        if (L.stellar_direction) {  // :1080-1111
                            double tx = px * L.rot_y_cos + py * 0.0 + pz * L.rot_y_sin;
                            double ty = px * 0.0 + py * 1.0 + pz * 0.0;
                            double tz = px * (-L.rot_y_sin) + py * 0.0 + pz * L.rot_y_cos;
                            px = tx * L.rot_z_cos + ty * (-L.rot_z_sin) + tz * 0.0;
                            py = tx * L.rot_z_sin + ty * L.rot_z_cos + tz * 0.0;
                            pz = tx * 0.0 + ty * 0.0 + tz * 1.0;
                            dx = L.star_dir[0]; dy = L.star_dir[1]; dz = L.star_dir[2];
                        }
                        initial_cell(sm, lay, T, px, py, pz, c0, c1, c2);
                    } else {  // thermal :1117-1266
                        f0 = 0; f1 = 0;
                        double xi = rng_next<TRACE>(rng, A);
                        const int ncdf = (T.nr - T.cell_depth) * T.nt * T.np;
                        const double samp = xi * __ldg(T.emis_cdf + ncdf - 1);
                        int lo = -1, hi = ncdf - 1;  // first p with cdf[p] >= samp (== the linear scan of :1132-1155)
                        while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (__ldg(T.emis_cdf + mid) >= samp) hi = mid; else lo = mid; }
                        c2 = hi % T.np; c1 = (hi / T.np) % T.nt; c0 = T.cell_depth + hi / (T.np * T.nt);
                        xi = rng_next<TRACE>(rng, A);
                        double rs = xi * (sm[c0 + 1] - sm[c0]); rs = sm[c0] + rs;
                        xi = rng_next<TRACE>(rng, A);
                        double ct = xi * (sm[lay.o_tc + c1 + 1] - sm[lay.o_tc + c1]); ct = sm[lay.o_tc + c1] + ct;
                        double st = sqrt(1.0 - ct * ct);
                        xi = rng_next<TRACE>(rng, A);
                        double phs;
                        if (T.np == 1) phs = 2.0 * PI * xi;
                        else if (c2 < T.np - 1) { phs = xi * (sm[lay.o_pf + c2 + 1] - sm[lay.o_pf + c2]); phs = sm[lay.o_pf + c2] + phs; }
                        else { phs = xi * (2.0 * PI - sm[lay.o_pf + c2]); phs = sm[lay.o_pf + c2] + phs; }
                        double cp = cos(phs), sp = sqrt(1.0 - cp * cp);
                        if (phs > PI) sp = -sp;
                        px = rs * st * cp; py = rs * st * sp; pz = rs * ct;
                        px = T.ox * px; py = T.oy * py; pz = T.oz * pz;
                        if (L.photon_emission == 1) {
                            xi = rng_next<TRACE>(rng, A);
                            double al = 2.0 * xi - 1.0;
                            xi = rng_next<TRACE>(rng, A);
                            double be = 2.0 * PI * xi;
                            double cb = cos(be), sb = sqrt(1.0 - cb * cb);
                            if (be > PI) sb = -sb;
                            dx = sqrt(1.0 - al * al) * cb; dy = sqrt(1.0 - al * al) * sb; dz = al;
                        } else {
                            xi = rng_next<TRACE>(rng, A);
                            double yb = (1.0 + L.photon_bias) * tan(PI * xi / 2.0) / sqrt(1.0 - L.photon_bias * L.photon_bias);
                            double ths = acos((1.0 - yb * yb) / (1.0 + yb * yb));
                            xi = rng_next<TRACE>(rng, A);
                            double be = 2.0 * PI * xi;
                            double r0 = px / (T.ox * T.ox), r1 = py / (T.oy * T.oy), r2 = pz / (T.oz * T.oz);
                            double nrm = sqrt(r0 * r0 + r1 * r1 + r2 * r2);
                            r0 = r0 / nrm; r1 = r1 / nrm; r2 = r2 / nrm;
                            e = direction_cosine(cos(PI - ths), be, r0, r1, r2, dx, dy, dz);
                            bias_weight = (PI * sin(ths) * (1.0 + L.photon_bias * cos(ths))) / (2.0 * sqrt(1.0 - L.photon_bias * L.photon_bias));
                        }
                        if (e == 0 && fabs(dz) >= 1.0) err_count(54);
                    }
                    if (e) { err_count(e); ++n_err; retire(); }
                    else if (L.photon_source == 2) {  // :599-621
                        S[0] = S[0] * bias_weight / __ldg(T.cell_weight + c0 + T.nr * (c1 + T.nt * c2));
                        atomicAdd(A.O.flux, S[0]);
                        ++n_peel; pk = PK_THERMAL; start_probe(0); ph = PH_PEEL;
                    } else { start_probe(0); ph = PH_PRE; }
                }
            }
        }

        // ================= B. one cell crossing for every walking lane =================
        if (ph == PH_PRE || ph == PH_WALK || ph == PH_PEEL) {
            const bool peel = (ph == PH_PEEL);
            const double n0 = peel ? L.det[0] : dx, n1 = peel ? L.det[1] : dy, n2 = peel ? L.det[2] : dz;
            CellFace o;
            cell_face(sm, lay, T, wx, wy, wz, n0, n1, n2, wf0, wf1, wc0, wc1, wc2, o);
            ++n_cf;
            record(o.nf0, o.nf1, o.co0, o.co1, o.co2);
            const int wci = wc0 + T.nr * (wc1 + T.nt * wc2);
            if (o.err) {
                err_count(o.err);
                if (ph == PH_PRE) { err_count(2); ++n_err; retire(); }
                else if (ph == PH_WALK) { err_count(3); ++n_err; retire(); }
                else if (pk == PK_SCATTER) { err_count(43); ++n_err; retire(); }
                else if (pk == PK_THERMAL) { err_count(46); err_count(47); ++n_err; retire(); }
                else { err_count(42); peel_exit = false; ph = PH_PEELDONE; }
            } else if (ph == PH_WALK) {
                const double kap = __ldg(T.kext + wci);
                const double tau_cell = o.dist * kap;
                if (tau_run + tau_cell > tau) {  // :705-720 / :862-879 interaction inside this cell
                    const double s = (tau - tau_run) / kap;
                    px = wx + s * dx; py = wy + s * dy; pz = wz + s * dz;
                    c0 = wc0; c1 = wc1; c2 = wc2; f0 = 0; f1 = 0;
                    if (L.flow_global) {  // add_flow_global :4992-5014
                        double th = acos(pz / sqrt(px * px + py * py + pz * pz)), phh = atan2(py, px);
                        double* f = A.O.flow3 + (size_t)3 * wci;
                        atomicAdd(f, (sin(th) * cos(phh) * dx + sin(th) * sin(phh) * dy + cos(th) * dz) * s * S[0]);
                        atomicAdd(f + 1, (cos(th) * cos(phh) * dx + cos(th) * sin(phh) * dy - sin(th) * dz) * s * S[0]);
                        atomicAdd(f + 2, (-sin(phh) * dx + cos(phh) * dy) * s * S[0]);
                    }
                    ph = PH_SCAT;
                } else {
                    wx = wx + o.dist * dx; wy = wy + o.dist * dy; wz = wz + o.dist * dz;
                    if (L.flow_global) {
                        double th = acos(wz / sqrt(wx * wx + wy * wy + wz * wz)), phh = atan2(wy, wx);
                        double* f = A.O.flow3 + (size_t)3 * wci;
                        atomicAdd(f, (sin(th) * cos(phh) * dx + sin(th) * sin(phh) * dy + cos(th) * dz) * o.dist * S[0]);
                        atomicAdd(f + 1, (cos(th) * cos(phh) * dx + cos(th) * sin(phh) * dy - sin(th) * dz) * o.dist * S[0]);
                        atomicAdd(f + 2, (-sin(phh) * dx + cos(phh) * dy) * o.dist * S[0]);
                    }
                    if (L.flow_theta) {  // :730-744
                        double* f = A.O.flow4 + (size_t)4 * wci;
                        if (o.nf0 == 1) { if (o.co0 > wc0) atomicAdd(f, S[0]); else if (o.co0 < wc0) atomicAdd(f + 1, S[0]); }
                        else if (o.nf0 == 2) { if (o.co1 > wc1) atomicAdd(f + 2, S[0]); else if (o.co1 < wc1) atomicAdd(f + 3, S[0]); }
                    }
                    wf0 = o.nf0; wf1 = o.nf1; wc0 = o.co0; wc1 = o.co1; wc2 = o.co2;
                    if (o.exit) {
                        if (L.photon_source == 2) atomicAdd(A.O.flux + 1, S[0]);  // :780 / :953
                        retire();
                    } else {
                        if (o.nf0 == 1 && o.nf1 == T.cell_depth) {  // surface :755-774
                            ++n_surf;
                            double xi = rng_next<TRACE>(rng, A);
                            if (xi > L.surface_albedo) retire();
                            else {  // lambertian :1369-1402, then peel_surface :4600-4708
                                double s0 = wx / (T.ox * T.ox), s1 = wy / (T.oy * T.oy), s2 = wz / (T.oz * T.oz);
                                double nrm = sqrt(s0 * s0 + s1 * s1 + s2 * s2);
                                s0 = s0 / nrm; s1 = s1 / nrm; s2 = s2 / nrm;
                                xi = rng_next<TRACE>(rng, A);
                                double al = sqrt(xi);
                                xi = rng_next<TRACE>(rng, A);
                                double be = 2.0 * PI * xi;
                                double e0, e1, e2;
                                int e = direction_cosine(al, be, s0, s1, s2, e0, e1, e2);
                                if (e) { err_count(e); ++n_err; retire(); }
                                else {
                                    dx = e0; dy = e1; dz = e2;
                                    px = wx; py = wy; pz = wz; c0 = wc0; c1 = wc1; c2 = wc2; f0 = wf0; f1 = wf1;
                                    // cos of the angle between the surface normal and the detector :4628-4634
                                    double nth = acos(s2 / sqrt(s0 * s0 + s1 * s1 + s2 * s2));
                                    double nph = atan2(s1, s0);
                                    if (nph < 0.0) nph = nph + 2.0 * PI;
                                    double cos_angle = sin(L.det_sph_theta) * cos(L.det_sph_phi) * sin(nth) * cos(nph) +
                                                       sin(L.det_sph_theta) * sin(L.det_sph_phi) * sin(nth) * sin(nph) +
                                                       cos(L.det_sph_theta) * cos(nth);
                                    tau_run = tau_run + tau_cell;
                                    S[1] = 0.0; S[2] = 0.0; S[3] = 0.0;
                                    if (cos_angle > 0.0) { ++n_peel; pk = PK_SURFACE; start_probe(1); ph = PH_PEEL; }
                                    else { c0 = c0 + 1; wc0 = c0; }
                                }
                            }
                        } else tau_run = tau_run + tau_cell;
                    }
                }
            } else {
                // probe walks: tau pre-pass :633-656 and the three peel walks
                tacc = tacc + o.dist * __ldg(T.kext + wci);
                wx = wx + o.dist * n0; wy = wy + o.dist * n1; wz = wz + o.dist * n2;
                const bool hit_surface = (o.nf0 == 1 && o.nf1 == T.cell_depth);
                if (o.exit || hit_surface) {
                    if (ph == PH_PRE) {
                        // first optical depth :660-685
                        bool go = true;
                        if (tacc < 1.e-6 && !hit_surface) { go = false; retire(); }
                        else if (tacc < 1.e-6 && hit_surface) { double xi = rng_next<TRACE>(rng, A); tau = -log(1.0 - xi); }
                        else {
                            double xi = rng_next<TRACE>(rng, A);
                            if (tacc < 50.0) {
                                tau = -log(1.0 - xi * (1.0 - exp(-tacc)));
                                double f = 1.0 - exp(-tacc);
                                S[0] = S[0] * f; S[1] = S[1] * f; S[2] = S[2] * f; S[3] = S[3] * f;
                            } else tau = -log(1.0 - xi);
                        }
                        if (go) { tau_run = 0.0; start_probe(0); ph = PH_WALK; }
                    } else { peel_exit = o.exit; ph = PH_PEELDONE; }
                } else { wf0 = o.nf0; wf1 = o.nf1; wc0 = o.co0; wc1 = o.co1; wc2 = o.co2; }
            }
        }

        // ================= D. interaction point reached: survival + start of the peel-off =================
        if (ph == PH_SCAT) {  // :788-815
            bool alive = L.photon_scattering != 0;
            if (TRACE && rng.exhausted) alive = false;
            if (alive) {
                double xi = rng_next<TRACE>(rng, A);
                if (xi < L.fstop) alive = false;
            }
            if (alive) {
                const double alb = __ldg(T.albedo + c0 + T.nr * (c1 + T.nt * c2));
                if (alb < 1.0 && alb > 0.0) {
                    double gamma = alb / (1.0 - L.fstop);
                    S[0] = gamma * S[0]; S[1] = gamma * S[1]; S[2] = gamma * S[2]; S[3] = gamma * S[3];
                }
                if (S[0] <= L.photon_minimum) alive = false;
            }
            if (!alive) retire();
            else { ++n_peel; pk = PK_SCATTER; start_probe(0); ph = PH_PEEL; }
        }

        // ---- ballot regrouping: the heavy events below run only when enough lanes of the warp wait for
        // them (or nobody is left walking), so that their instructions are issued for many lanes at once.
        const unsigned m_evt = __ballot_sync(FULL, ph == PH_PEELDONE || ph == PH_SCAT2);
        const unsigned m_walk = __ballot_sync(FULL, ph == PH_PRE || ph == PH_WALK || ph == PH_PEEL);
        const bool run_events = m_evt && (__popc(m_evt) >= L.defer_events || m_walk == 0u);

        // ================= C. a peel walk ended: weight + deposit =================
        if (run_events && ph == PH_PEELDONE) {
            const bool ok = peel_exit && tacc < 50.0;
            if (pk == PK_THERMAL) {  // :4571-4596
                if (ok) {
                    double w = exp(-tacc) / (4.0 * PI);
                    double W0 = w * S[0];
                    if (W0 > 0.0 && W0 < 1.e100) deposit(W0, 0, 0, 0, false); else err_count(51);
                }
                start_probe(0); ph = PH_PRE;
            } else if (pk == PK_SURFACE) {  // :4675-4704
                if (ok) {
                    double s0 = px / (T.ox * T.ox), s1 = py / (T.oy * T.oy), s2 = pz / (T.oz * T.oz);
                    double nrm = sqrt(s0 * s0 + s1 * s1 + s2 * s2);
                    s0 = s0 / nrm; s1 = s1 / nrm; s2 = s2 / nrm;
                    double nth = acos(s2 / sqrt(s0 * s0 + s1 * s1 + s2 * s2));
                    double nph = atan2(s1, s0);
                    if (nph < 0.0) nph = nph + 2.0 * PI;
                    double cos_angle = sin(L.det_sph_theta) * cos(L.det_sph_phi) * sin(nth) * cos(nph) +
                                       sin(L.det_sph_theta) * sin(L.det_sph_phi) * sin(nth) * sin(nph) +
                                       cos(L.det_sph_theta) * cos(nth);
                    double w = exp(-tacc) * cos_angle / PI;
                    double W0 = w * S[0];
                    if (W0 > 0.0 && W0 < 1.e100) deposit(W0, 0, 0, 0, false); else err_count(52);
                }
                c0 = c0 + 1;  // :770
                start_probe(0); ph = PH_WALK;
            } else {  // peel_photon :4763-4986
                if (ok) {
                    const double w = exp(-tacc);
                    double mu = dx * L.det[0] + dy * L.det[1] + dz * L.det[2];
                    if (mu >= 1.0) mu = 1.0 - 1.e-10;
                    else if (mu <= -1.0) mu = -1.0 + 1.e-10;
                    double F[16];
#if ARTES_FAITHFUL
                    matrix_at(T, c0 + T.nr * (c1 + T.nt * c2), acos(mu), F);
                    double phi_old = atan2(dy, dx);
                    if (phi_old < 0.0) phi_old = phi_old + 2.0 * PI;
                    if (phi_old > 2.0 * PI) phi_old = phi_old - 2.0 * PI;
                    const double phi_new = L.det_atan2;
                    if (!(fabs(dz) < 1.0)) err_count(45);
                    else {
                        double nc = (L.det[2] - dz * mu) / (sqrt(1.0 - mu * mu) * sqrt(1.0 - dz * dz));
                        double phs = 0.0;
                        bool good = true;
                        if (fabs(nc) < 1.0) phs = acos(nc);
                        else if (nc >= 1.0) phs = 0.0 + 1.e-10;
                        else if (nc <= -1.0) phs = PI - 1.e-10;
                        else { good = false; err_count(44); }
                        if (good) {
                            if (phi_old - phi_new >= 0.0 && phi_old - phi_new < PI) phs = 2.0 * PI - phs;
                            if (2.0 * PI + phi_old - phi_new >= 0.0 && 2.0 * PI + phi_old - phi_new < PI) phs = 2.0 * PI - phs;
                            if (phs < 0.0) phs = phs + 2.0 * PI;
                            double so[4];
                            int soft = 0;
                            int e = polarization_rotation(mu, phs, S, F, dz, L.det[2], so, true, soft);
                            if (soft) err_count(soft);
                            if (e) err_count(e);
                            else if (w * so[0] > 0.0 && w * so[0] < 1.e100) deposit(w * so[0], -(w * so[1]), w * so[2], w * so[3], true);
                            else err_count(53);
                        }
                    }
#else
                    matrix_at_deg(T, c0 + T.nr * (c1 + T.nt * c2), acos(mu) * (180.0 / PI), F);
                    if (!(fabs(dz) < 1.0)) err_count(45);
                    else {
                        const double smu = sqrt(1.0 - mu * mu);
                        double nc = (L.det[2] - dz * mu) / (smu * sqrt(1.0 - dz * dz));
                        if (!(nc == nc)) err_count(44);
                        else {
                            nc = fmin(fmax(nc, -1.0), 1.0);
                            const double cr = dy * L.det[0] - dx * L.det[1];                 // sin(phi_old - phi_new) > 0 ?
                            const bool flip = (cr > 0.0) || (cr == 0.0 && dx * L.det[0] + dy * L.det[1] > 0.0);
                            const double c2a = 2.0 * nc * nc - 1.0;
                            double s2a = 2.0 * nc * sqrt(fmax(1.0 - nc * nc, 0.0));
                            if (flip) s2a = -s2a;
                            const double nc2 = (dz - L.det[2] * mu) / (smu * sqrt(1.0 - L.det[2] * L.det[2]));
                            double so[4];
                            int soft = 0;
                            int e = (fabs(L.det[2]) < 1.0) ? polrot_fast(c2a, s2a, flip, nc2, S, F, so, true, soft) : 16;
                            if (e) err_count(e);
                            else if (w * so[0] > 0.0 && w * so[0] < 1.e100) deposit(w * so[0], -(w * so[1]), w * so[2], w * so[3], true);
                            else err_count(53);
                        }
                    }
#endif
                }
                ph = PH_SCAT2;
            }
        }

        // ================= E. scattering event (scatter_photon :1434-1532 + polarization_rotation) ==========
        if (run_events && ph == PH_SCAT2) {
            ++n_sc; ++t_nsc;
            const int ci = c0 + T.nr * (c1 + T.nt * c2);
            double e0 = 0, e1 = 0, e2 = 0;
            int e;
#if ARTES_FAITHFUL
            double alpha, beta;
            e = sample_angles<TRACE>(sm, lay, A, rng, S, ci, alpha, beta);
            if (!e) e = direction_cosine(alpha, beta, dx, dy, dz, e0, e1, e2);
            if (!e && !(fabs(alpha) < 1.0)) e = 50;
            if (!e) {
                double F[16], Sn[4];
                matrix_at(T, ci, acos(alpha), F);
                int soft = 0;
                e = polarization_rotation(alpha, beta, S, F, dz, e2, Sn, false, soft);
                if (soft) err_count(soft);
                if (!e) { S[0] = Sn[0]; S[1] = Sn[1]; S[2] = Sn[2]; S[3] = Sn[3]; dx = e0; dy = e1; dz = e2; }
            }
#else
            FastAngles g;
            e = sample_angles_fast<TRACE>(A, rng, S, ci, g);
            if (!e) {
                // direction_cosine :1962-2052 without the acos / cos round trip
                const double cto = dz / sqrt(dx * dx + dy * dy + dz * dz);
                const double sto = sqrt(1.0 - cto * cto);
                const double ctn = cto * g.alpha + sto * g.sT * g.cb;
                const double stn = sqrt(1.0 - ctn * ctn);
                double nc = (g.alpha - ctn * cto) / (stn * sto);
                if (!(nc == nc)) e = 20;
                else {
                    if (nc >= 1.0) nc = 1.0 - 1.e-10; else if (nc <= -1.0) nc = -1.0 + 1.e-10;
                    const double sD = sqrt(1.0 - nc * nc) * (g.flip ? -1.0 : 1.0);
                    const double rho = sqrt(dx * dx + dy * dy);
                    const double cph = rho > 0.0 ? dx / rho : 1.0, sph = rho > 0.0 ? dy / rho : 0.0;
                    e0 = stn * (cph * nc - sph * sD); e1 = stn * (sph * nc + cph * sD); e2 = ctn;
                    if (!(fabs(e2) < 1.0)) e = 16;
                }
            }
            if (!e) {
                double F[16], Sn[4];
                matrix_at_deg(T, ci, g.deg, F);
                const double nc2 = (dz - e2 * g.alpha) / (g.sT * sqrt(1.0 - e2 * e2));
                int soft = 0;
                e = polrot_fast(g.cb * g.cb - g.sb * g.sb, 2.0 * g.sb * g.cb, g.flip, nc2, S, F, Sn, false, soft);
                if (soft) err_count(soft);
                if (!e) { S[0] = Sn[0]; S[1] = Sn[1]; S[2] = Sn[2]; S[3] = Sn[3]; dx = e0; dy = e1; dz = e2; }
            }
#endif
            if (e) { err_count(e); ++n_err; retire(); }
            else {
                double xi = rng_next<TRACE>(rng, A);  // :845
                tau = -log(1.0 - xi);
                tau_run = 0.0;
                start_probe(0); ph = PH_WALK;
            }
        }

        if (__all_sync(FULL, ph == PH_IDLE)) break;
    }

    // ---- flush event counters (warp reduce, one atomic per warp and counter)
    unsigned long long v[7] = {n_emit, n_cf, n_sc, n_peel, n_surf, n_draw, n_err};
#pragma unroll
    for (int k = 0; k < 7; ++k) {
        unsigned long long x = v[k];
        for (int off = 16; off > 0; off >>= 1) x += __shfl_xor_sync(FULL, x, off);
        if (lane == 0 && x) atomicAdd(A.O.stats + k, x);
    }
}

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
