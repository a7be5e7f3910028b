// artes_oracle.cc -- CPU restatement of the ARTES photon-packet transport loop.
//
// TEST INFRASTRUCTURE ONLY (see artes_oracle.h).  PARITY UNPINNED: no golden vectors exist in
// the reference and it cannot be compiled here (no Fortran compiler).  What stands in for the pin (tests/test_oracle.py, DESIGN.md
// section 2): Chandrasekhar's exact solution of the semi-infinite isotropic atmosphere, and numerical solutions of the invariance
// equation for anisotropic (Henyey-Greenstein), polarised (Rayleigh) and polarised anisotropic scattering -- geometric albedo, limb
// darkening, radial limb polarisation, the polarised phase curve -- which this restatement reproduces to its Monte Carlo noise.
//
// Every function cites the src/ARTES.f90 range it follows.  Arithmetic is IEEE double in the
// reference's operation order; build with -ffp-contract=off (gfortran -O3 on generic x86-64
// emits no FMA).  Deliberate deviations (all on error paths the reference leaves undefined):
//   * a failed cell_face (error 031/033-037) drops the photon at once instead of walking on
//     with face_distance = 1e100 (:638-649, :695-701, :854-860);
//   * uninitialised results (beta2 :1745, cos_theta_new :1987, phi_new :2013, phi_scatter :4895,
//     stokes_out :4936) drop the photon / the contribution and count the error code;
//   * pixel indices outside the image are dropped (error slot 60) instead of written out of bounds.

#include "artes_oracle.h"

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <chrono>
#include <vector>
#include <sys/stat.h>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

const double PI = 4.0 * std::atan(1.0);  // :9

// ------------------------------------------------------------------------------------------
// Tables (program-scope arrays of the reference, :58-91)
// ------------------------------------------------------------------------------------------
struct Tables {
    int nr = 0, nt = 0, np = 0, cells = 0;
    std::vector<double> rfront, thetafront, phifront, tcos, ttan, psin, pcos;
    std::vector<int> thetaplane;
    double ox = 1, oy = 1, oz = 1;
    double sinbeta[360], cosbeta[360], sin2beta[360], cos2beta[360];
    // per wavelength
    std::vector<double> ksca, kabs, kext, albedo;
    int n_uniq = 0;
    std::vector<double> M;    // [u][180][16]
    std::vector<double> p1k;  // [u][4]  cell_p11..p14_int
    std::vector<int> c2u;
    int cell_depth = 0;
    bool thermal = false;
    std::vector<double> cell_weight, emis_cdf;

    int idx(const int c[3]) const { return c[0] + nr * (c[1] + nt * c[2]); }
    const double* mat(int cellidx, int angle /*0..179*/) const {
        return &M[((size_t)c2u[cellidx] * 180 + angle) * 16];
    }
};

// ------------------------------------------------------------------------------------------
// Random numbers
// ------------------------------------------------------------------------------------------
inline uint32_t mulhi32(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }

void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = mulhi32(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        uint32_t hi1 = mulhi32(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

struct Rng {
    int kind = 1;  // 0 MZ, 1 Philox, 2 injected
    // MZ (:4197-4230): state(thread,1:4)
    int32_t s1 = 0, s2 = 362436069, s3 = 16163801, s4 = 1131199299;  // default_seed(2:4) :114
    // Philox
    uint64_t seed = 0, id = 0;
    uint32_t buf[4] = {0, 0, 0, 0};
    // injected
    const double* xi = nullptr;
    int max_draws = 0;
    bool exhausted = false;
    uint32_t ndraw = 0;  // draws of the current photon
    uint64_t total = 0;
    int err55 = 0;

    void start_photon(uint64_t photon_id, const double* stream) {
        id = photon_id; ndraw = 0; xi = stream; exhausted = false;
    }
    double next() {
        ++total;
        if (kind == 0) {
            int32_t imz = s1 - s3;                                 // :4205
            if (imz < 0) imz += 2147483579;                        // :4207
            s1 = s2; s2 = s3; s3 = imz;                            // :4209-4211
            s4 = (int32_t)(69069u * (uint32_t)s4 + 1013904243u);   // :4212 (32-bit wrap)
            imz = (int32_t)((uint32_t)imz + (uint32_t)s4);         // :4214
            double v = 0.5 + 0.23283064e-9 * (double)imz;          // :4216
            if (!(v > 0.0 && v < 1.0)) ++err55;
            ++ndraw;
            return v;
        } else if (kind == 1) {
            uint32_t w = ndraw & 3u;
            if (w == 0) {
                uint32_t ctr[4] = {(uint32_t)id, (uint32_t)(id >> 32), ndraw >> 2, 0u};
                uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
                philox4x32_10(ctr, key, buf);
            }
            ++ndraw;
            return ((double)buf[w] + 0.5) * (1.0 / 4294967296.0);
        } else {
            if ((int)ndraw >= max_draws) { exhausted = true; ++ndraw; return 0.5; }
            return xi[ndraw++];
        }
    }
};

// ------------------------------------------------------------------------------------------
// Recorder for the injected-stream walk parity test
// ------------------------------------------------------------------------------------------
struct Recorder {
    int32_t len = 0;
    uint64_t hash = 1469598103934665603ull;
    int32_t* head = nullptr;
    int max_rec = 0;
    void word(int32_t v) { hash ^= (uint32_t)v; hash *= 1099511628211ull; }
    void tuple(int a, int b, int c, int d, int e) {
        if (head && len < max_rec) { int32_t* p = head + (size_t)len * 5; p[0] = a; p[1] = b; p[2] = c; p[3] = d; p[4] = e; }
        word(a); word(b); word(c); word(d); word(e);
        ++len;
    }
};

// ------------------------------------------------------------------------------------------
// Per-thread accumulators
// ------------------------------------------------------------------------------------------
struct Accum {
    std::vector<double> det;          // detector_thread(nx,ny,4,3)
    double flux_emitted = 0, flux_exit = 0;
    std::vector<double> flow4, flow3;
    uint64_t err[ARTES_ERR_SLOTS];
    uint64_t n_emit = 0, n_cf = 0, n_sc = 0, n_peel = 0, n_surf = 0, n_error = 0;
    Accum() { std::memset(err, 0, sizeof(err)); }
    void error(int code) { if (code >= 0 && code < ARTES_ERR_SLOTS) ++err[code]; }
};

struct Run {
    const Tables& T;
    artes_launch_t L;
    double det_dir[3];
    double sin_dt, cos_dt, sin_dp, cos_dp;
    bool stat_emul = false;
    Accum* A = nullptr;
    Rng* R = nullptr;
    Recorder* rec = nullptr;
    explicit Run(const Tables& t) : T(t) {}
};

inline void emulate_stat_call() {
    struct stat sb;
    (void)::stat("/tmp/artes_oracle_error.log", &sb);  // :549, :2820
}

// quadratic_equation :4154-4173
inline void quadratic(double a, double b, double c, double s[2]) {
    s[0] = 0.0; s[1] = 0.0;
    double disc = b * b - 4.0 * a * c;
    if (disc >= 0.0) {
        double q = -0.5 * (b + std::copysign(1.0, b) * std::sqrt(disc));
        if (std::fabs(a) > 1.e-100) s[0] = q / a;
        if (std::fabs(q) > 1.e-100) s[1] = c / q;
    }
}

// root choice repeated at :2897-2907, :2944-2954, :3054-3064, ...
inline double pick_root(double s1, double s2, double thr) {
    double d = 0.0;
    if (s1 > thr && s2 <= thr && s1 < 1.e100) d = s1;
    else if (s2 > thr && s1 <= thr && s2 < 1.e100) d = s2;
    else if (s1 > thr && s2 > thr) {
        if (s1 < 1.e100 && s1 < s2) d = s1;
        else if (s2 < 1.e100 && s2 < s1) d = s2;
    }
    return d;
}

struct CellFace {
    int nf[2];
    int co[3];
    double dist;
    bool exit, err;
    int code;
};

// next_cell :2671-2798
inline void next_cell(const Tables& T, const int cf[2], const int nf[2], const int ci[3], int co[3], Accum* A) {
    co[0] = co[1] = co[2] = 0;
    if (nf[0] == 1) {
        if (cf[0] == 1 && nf[1] == cf[1]) { co[0] = ci[0] + 1; co[1] = ci[1]; co[2] = ci[2]; }
        else if (nf[1] == ci[0])          { co[0] = ci[0] - 1; co[1] = ci[1]; co[2] = ci[2]; }
        else if (nf[1] == ci[0] + 1)      { co[0] = ci[0] + 1; co[1] = ci[1]; co[2] = ci[2]; }
        else if (A) A->error(22);
    }
    if (nf[0] == 2) {
        if (cf[0] == 2 && nf[1] == cf[1] && T.thetafront[nf[1]] < PI / 2.0)      { co[0] = ci[0]; co[1] = ci[1] + 1; co[2] = ci[2]; }
        else if (cf[0] == 2 && nf[1] == cf[1] && T.thetafront[nf[1]] > PI / 2.0) { co[0] = ci[0]; co[1] = ci[1] - 1; co[2] = ci[2]; }
        else if (nf[1] == ci[1])     { co[0] = ci[0]; co[1] = ci[1] - 1; co[2] = ci[2]; }
        else if (nf[1] == ci[1] + 1) { co[0] = ci[0]; co[1] = ci[1] + 1; co[2] = ci[2]; }
        else if (A) A->error(23);
    }
    if (nf[0] == 3) {
        if (ci[2] == T.np - 1 && nf[1] == 0) { co[0] = ci[0]; co[1] = ci[1]; co[2] = 0; }
        else if (ci[2] == 0 && nf[1] == 0)   { co[0] = ci[0]; co[1] = ci[1]; co[2] = T.np - 1; }
        else if (nf[1] == ci[2] + 1)         { co[0] = ci[0]; co[1] = ci[1]; co[2] = ci[2] + 1; }
        else if (nf[1] == ci[2])             { co[0] = ci[0]; co[1] = ci[1]; co[2] = ci[2] - 1; }
        else if (A) A->error(24);
    }
}

// cell_face :2800-3470
void cell_face(const Run& run, double x, double y, double z, const double n[3], const int cf[2],
               const int cell[3], CellFace& o) {
    const Tables& T = run.T;
    Accum* A = run.A;
    if (A) ++A->n_cf;
    if (run.stat_emul) emulate_stat_call();
    if (cell[0] < T.cell_depth && A) A->error(25);  // :2811 (log only)

    o.exit = false; o.err = false; o.code = 0; o.dist = 0.0;
    o.nf[0] = o.nf[1] = 0; o.co[0] = o.co[1] = o.co[2] = 0;

    double d[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};  // distance(i,j) :2835
    const double a = 1.0 / T.ox, b = 1.0 / T.oy, c = 1.0 / T.oz;  // :2838-2840

    int face[3][3];  // face(i,j) :2842-2853
    face[0][0] = cell[0]; face[0][1] = cell[0] + 1; face[0][2] = -999;
    face[1][0] = cell[1]; face[1][1] = cell[1] + 1; face[1][2] = -999;
    face[2][0] = cell[2]; face[2][1] = cell[2] + 1; face[2][2] = -999;
    if (face[2][1] == T.np) face[2][1] = 0;

    if (cf[0] == 1) { face[0][0] = cf[1] - 1; face[0][1] = cf[1] + 1; face[0][2] = cf[1]; }        // :2855-2859
    else if (cf[0] == 2) { face[1][0] = cf[1] - 1; face[1][1] = cf[1] + 1; face[1][2] = cf[1]; }   // :2861-2865
    else if (cf[0] == 3) {                                                                          // :2867-2879
        face[2][0] = (cf[1] == 0) ? T.np - 1 : cf[1] - 1;
        face[2][1] = (cf[1] == T.np - 1) ? 0 : cf[1] + 1;
    }

    auto sphere = [&](int k, double thr) -> double {  // :2891-2907 and copies
        double qa = a * a * n[0] * n[0] + b * b * n[1] * n[1] + c * c * n[2] * n[2];
        double qb = 2.0 * (a * a * x * n[0] + b * b * y * n[1] + c * c * z * n[2]);
        double qc = a * a * x * x + b * b * y * y + c * c * z * z - T.rfront[k] * T.rfront[k];
        double s[2];
        quadratic(qa, qb, qc, s);
        return pick_root(s[0], s[1], thr);
    };
    auto cone = [&](int k, double thr) -> double {  // :3028-3064 and copies
        double t = T.ttan[k];
        double qa = a * a * n[0] * n[0] + b * b * n[1] * n[1] - c * c * n[2] * n[2] * t * t;
        double qb = 2.0 * (a * a * x * n[0] + b * b * y * n[1] - c * c * z * n[2] * t * t);
        double qc = a * a * x * x + b * b * y * y - c * c * z * z * t * t;
        double s[2];
        quadratic(qa, qb, qc, s);
        for (int m = 0; m < 2; ++m) {
            if (s[m] > 1.e-15) {
                double zt = z + s[m] * n[2];
                if ((zt > 0.0 && T.thetafront[k] > PI / 2.0) || (zt < 0.0 && T.thetafront[k] < PI / 2.0)) s[m] = 0.0;
            }
        }
        return pick_root(s[0], s[1], thr);
    };
    auto theta_inner = [&](int k) {  // face(2,1)
        if (T.thetaplane[k] == 1) d[1][0] = cone(k, 1.e-15);
        else if (T.thetaplane[k] == 2) { if (-z / n[2] > 0.0 && n[2] > 1.e-15) d[1][0] = -z / n[2]; }   // :3068
    };
    auto theta_outer = [&](int k) {  // face(2,2)
        if (T.thetaplane[k] == 1) d[1][1] = cone(k, 1.e-15);
        else if (T.thetaplane[k] == 2) { if (-z / n[2] > 0.0 && n[2] < -1.e-15) d[1][1] = -z / n[2]; }  // :3118
    };

    // ---- radial :2885-3010
    if (cf[0] == 1) {
        if (cell[0] == cf[1] - 1) d[0][0] = sphere(face[0][0], 1.e-15);
        else if (cell[0] == cf[1]) d[0][1] = sphere(face[0][1], 1.e-15);
        if (cell[0] == cf[1] - 1) d[0][2] = sphere(face[0][2], 1.e-3);  // same face :2933-2954
        else if (cell[0] == cf[1]) {}
        else if (A) A->error(27);
    } else {
        d[0][0] = sphere(face[0][0], 1.e-15);
        d[0][1] = sphere(face[0][1], 1.e-15);
    }

    // ---- polar :3014-3290
    if (cf[0] == 2) {
        if (cell[1] == cf[1] - 1 && face[1][0] != 0) theta_inner(face[1][0]);
        else if (cell[1] == cf[1] && face[1][1] != T.nt) theta_outer(face[1][1]);
        int k = face[1][2];
        if ((T.thetafront[k] < PI / 2.0 && cell[1] == cf[1] - 1) || (T.thetafront[k] > PI / 2.0 && cell[1] == cf[1])) {
            if (T.thetaplane[k] == 1) d[1][2] = cone(k, 1.e-3);  // :3129-3169
        }
    } else {
        if (face[1][0] < 0 || face[1][0] > T.nt) { if (A) A->error(29); face[1][0] = 0; }
        if (face[1][0] != 0) theta_inner(face[1][0]);
        if (face[1][1] != T.nt) theta_outer(face[1][1]);
    }

    // ---- azimuthal :3292-3350 (full planes; guards reproduced verbatim incl. their typos)
    double sp1 = 0.0, sp2 = 0.0;
    if (cf[0] == 3) {
        if (cell[2] == cf[1] - 1 || (cell[2] == T.np - 1 && cf[1] == 0)) {
            int k = face[2][0];
            if (std::fabs(b * n[1] * T.pcos[k] - a * n[0] * T.psin[k]) > 0.0) {
                sp1 = (a * x * T.psin[k] - b * y * T.pcos[k]) / (b * n[1] * T.pcos[k] - a * n[0] * T.psin[k]);
                if (sp1 > 1.e-15 && sp1 < 1.e100) d[2][0] = sp1;
            }
        } else if (cell[2] == cf[1]) {
            int k = face[2][1];
            if (std::fabs(b * n[1] * T.pcos[k] - a * n[0] * T.psin[k]) > 0.0) {
                sp2 = (a * x * T.psin[k] - b * y * T.pcos[k]) / (b * n[1] * T.pcos[k] - a * n[0] * T.psin[k]);
                if (sp2 > 1.e-15 && sp1 < 1.e100) d[2][1] = sp2;  // :3318 tests solutions_p(1)
            }
        }
    } else if (T.np > 1) {
        int k = face[2][0];
        if (std::fabs(b * n[1] * T.pcos[k] - a * n[0] * T.psin[k]) > 0.0) {
            sp1 = (a * x * T.psin[k] - b * y * T.pcos[k]) / (b * n[1] * T.pcos[k] - a * n[0] * T.psin[k]);
            if (sp1 > 1.e-15 && sp1 < 1.e100) d[2][0] = sp1;
        }
        k = face[2][1];
        if (std::fabs(n[1] * T.pcos[k] - n[0] * T.psin[k]) > 0.0) {  // :3341 omits a, b
            sp2 = (a * x * T.psin[k] - b * y * T.pcos[k]) / (b * n[1] * T.pcos[k] - a * n[0] * T.psin[k]);
            if (sp2 > 1.e-15 && sp1 < 1.e100) d[2][1] = sp2;  // :3346 tests solutions_p(1)
        }
    }

    // ---- nearest face :3358-3418 (j outer, i inner, strict <)
    double fd = 1.e100;
    int li = -1, lj = -1;
    for (int j = 0; j < 3; ++j)
        for (int i = 0; i < 3; ++i)
            if (d[i][j] > 1.e-9 && d[i][j] < fd) { fd = d[i][j]; li = i; lj = j; }
    if (li < 0) {
        fd = 1.e100;
        for (int j = 0; j < 3; ++j)
            for (int i = 0; i < 3; ++i)
                if (d[i][j] > 1.e-12 && d[i][j] < fd) { fd = d[i][j]; li = i; lj = j; }
        if (li < 0) { o.err = true; o.code = 31; o.dist = fd; if (A) A->error(31); return; }  // deviation: return at once
    }
    o.dist = fd;
    o.nf[0] = li + 1;
    o.nf[1] = face[li][lj];
    if (o.nf[1] == -999) { o.err = true; o.code = 33; if (A) A->error(33); return; }  // :3423-3428
    next_cell(T, cf, o.nf, cell, o.co, A);  // :3432

    if (o.nf[0] == 1 && o.nf[1] == T.nr) o.exit = true;  // :3436

    if (cf[0] == 1 && cf[1] == T.cell_depth && o.nf[0] == 1 && o.nf[1] == T.cell_depth) { o.err = true; o.code = 34; }
    else if (o.co[0] == T.nr && !o.exit) { o.err = true; o.code = 35; }
    else if (o.co[1] == T.nt) { o.err = true; o.code = 36; }
    else if (cell[0] == o.co[0] && cell[1] == o.co[1] && cell[2] == o.co[2] && !o.exit) { o.err = true; o.code = 37; }
    if (o.err && A) A->error(o.code);
}

// mueller_matrix_filler :1934-1960  -> m[0]=(1,1) m[1]=(1,2) m[2]=(2,1) m[3]=(2,2)
inline void mueller(double psi, double m[4]) {
    double c2p = std::cos(2.0 * psi);
    double s2p = std::sqrt(1.0 - c2p * c2p);
    if (psi > PI / 2.0 && psi < PI) s2p = -s2p;
    else if (psi > 3.0 * PI / 2.0 && psi < 2.0 * PI) s2p = -s2p;
    else if (psi > -PI / 2.0 && psi < 0.0) s2p = -s2p;
    else if (psi > -2.0 * PI && psi < -3.0 * PI / 2.0) s2p = -s2p;
    m[0] = c2p; m[2] = -s2p; m[1] = s2p; m[3] = c2p;
}

// direction_cosine :1962-2052.  Returns an error code (18/19/20/21) when the reference would
// read an uninitialised variable, 0 otherwise.
int direction_cosine(double alpha, double beta, const double dir[3], double dn[3]) {
    double cto = dir[2] / std::sqrt(dir[0] * dir[0] + dir[1] * dir[1] + dir[2] * dir[2]);
    double sto = std::sqrt(1.0 - cto * cto);
    double phi_old = std::atan2(dir[1], dir[0]);
    if (phi_old < 0.0) phi_old = phi_old + 2.0 * PI;
    double ctn;
    if (beta >= PI && beta < 2.0 * PI) ctn = cto * alpha + sto * std::sqrt(1.0 - alpha * alpha) * std::cos(2.0 * PI - beta);
    else if (beta >= 0.0 && beta < PI) ctn = cto * alpha + sto * std::sqrt(1.0 - alpha * alpha) * std::cos(beta);
    else return 18;
    double stn = std::sqrt(1.0 - ctn * ctn);
    double nc = (alpha - ctn * cto) / (stn * sto);
    if (nc >= 1.0) nc = 1.0 - 1.e-10;
    else if (nc <= -1.0) nc = -1.0 + 1.e-10;
    double phi_new;
    if (std::fabs(nc) <= 1.0) {
        if (beta >= PI && beta < 2.0 * PI) phi_new = phi_old - std::acos(nc);
        else phi_new = phi_old + std::acos(nc);
    } else return 20;  // NaN
    if (phi_new < 0.0) phi_new = phi_new + 2.0 * PI;
    if (phi_new > 2.0 * PI) phi_new = phi_new - 2.0 * PI;
    double cpn = std::cos(phi_new), spn;
    if (phi_new >= 0.0 && phi_new < PI) spn = std::sqrt(1.0 - cpn * cpn);
    else if (phi_new >= PI && phi_new <= 2.0 * PI) spn = -std::sqrt(1.0 - cpn * cpn);
    else return 21;
    dn[0] = stn * cpn; dn[1] = stn * spn; dn[2] = ctn;
    return 0;
}

// Matrix at a scattering angle: bracket choice and interpolation of :1448-1530 / :4780-4862.
inline void matrix_at(const Tables& T, int cellidx, double acos_a, double F[16]) {
    double deg = acos_a * 180.0 / PI;
    int lo, up;
    if (std::fmod(deg, 1.0) > 0.5) { up = (int)deg + 2; lo = (int)deg + 1; }
    else { up = (int)deg + 1; lo = (int)deg; }
    if (up == 1) { const double* m = T.mat(cellidx, 0); for (int i = 0; i < 16; ++i) F[i] = m[i]; }
    else if (lo == 180) { const double* m = T.mat(cellidx, 179); for (int i = 0; i < 16; ++i) F[i] = m[i]; }
    else {
        const double* m0 = T.mat(cellidx, lo - 1);
        const double* m1 = T.mat(cellidx, up - 1);
        double y0 = (double)lo - 0.5, y1 = (double)up - 0.5;
        for (int i = 0; i < 16; ++i) F[i] = (m1[i] - m0[i]) * (deg - y0) / (y1 - y0) + m0[i];
    }
}

// polarization_rotation :1663-1932.  F is row-major (F[4*r+c] = scatter(r+1,c+1)).
// Returns an error code when the output would be undefined (11, 16), else 0; *soft receives the
// log-only codes (12, 13, 14, 15).
int polarization_rotation(double alpha, double beta, const double Sin[4], const double F[16],
                          const double dir[3], const double dn[3], double Sout[4], bool peeling, Accum* A) {
    double norm;
    if (std::fabs(alpha) < 1.0 && std::fabs(dn[2]) < 1.0) {
        double nc2 = (dir[2] - dn[2] * alpha) / (std::sqrt(1.0 - alpha * alpha) * std::sqrt(1.0 - dn[2] * dn[2]));
        double beta2;
        if (std::fabs(nc2) <= 1.0) beta2 = std::acos(nc2);
        else if (nc2 > 1.0 && nc2 < 1.00001) beta2 = 0.0;
        else if (nc2 < -1.0 && nc2 > -1.00001) beta2 = PI;
        else return 11;
        double mm[4];
        mueller(beta, mm);
        double sr[4];
        sr[0] = Sin[0];
        sr[1] = mm[0] * Sin[1] + mm[1] * Sin[2];
        sr[2] = mm[2] * Sin[1] + mm[3] * Sin[2];
        sr[3] = Sin[3];
        if (std::sqrt(sr[1] * sr[1] + sr[2] * sr[2] + sr[3] * sr[3]) > 0.0)
            norm = std::sqrt(Sin[1] * Sin[1] + Sin[2] * Sin[2] + Sin[3] * Sin[3]) /
                   std::sqrt(sr[1] * sr[1] + sr[2] * sr[2] + sr[3] * sr[3]);
        else norm = 1.0;
        if (norm < 1.0 || norm > 1.0) { sr[1] = sr[1] * norm; sr[2] = sr[2] * norm; sr[3] = sr[3] * norm; }
        double ss[4];
        for (int r = 0; r < 4; ++r)
            ss[r] = F[4 * r] * sr[0] + F[4 * r + 1] * sr[1] + F[4 * r + 2] * sr[2] + F[4 * r + 3] * sr[3];
        if (!peeling) {
            if (ss[0] > 0.0) { norm = sr[0] / ss[0]; for (int r = 0; r < 4; ++r) ss[r] = norm * ss[r]; }
            else if (A) A->error(12);
        }
        // :1818-1826 (beta outside [0,2pi) keeps the previous matrix = mueller(beta))
        if (beta >= 0.0 && beta < PI) mueller(beta2, mm);
        else if (beta >= PI && beta < 2.0 * PI) mueller(-beta2, mm);
        Sout[0] = ss[0];
        Sout[1] = mm[0] * ss[1] + mm[1] * ss[2];
        Sout[2] = mm[2] * ss[1] + mm[3] * ss[2];
        Sout[3] = ss[3];
        if (std::sqrt(Sout[1] * Sout[1] + Sout[2] * Sout[2] + Sout[3] * Sout[3]) > 0.0)
            norm = std::sqrt(ss[1] * ss[1] + ss[2] * ss[2] + ss[3] * ss[3]) /
                   std::sqrt(Sout[1] * Sout[1] + Sout[2] * Sout[2] + Sout[3] * Sout[3]);
        else norm = 1.0;
        if (norm < 1.0 || norm > 1.0) { Sout[1] = Sout[1] * norm; Sout[2] = Sout[2] * norm; Sout[3] = Sout[3] * norm; }
        return 0;
    } else if (alpha >= 1.0 && alpha < 1.0001) {
        for (int r = 0; r < 4; ++r) Sout[r] = Sin[r];
        if (A) A->error(13);
        return 0;
    } else if (alpha <= -1.0 && alpha > -1.0001) {
        double ss[4];
        for (int r = 0; r < 4; ++r)
            ss[r] = F[4 * r] * Sin[0] + F[4 * r + 1] * Sin[1] + F[4 * r + 2] * Sin[2] + F[4 * r + 3] * Sin[3];
        if (peeling) { for (int r = 0; r < 4; ++r) Sout[r] = ss[r]; }
        else if (ss[0] > 0.0) { norm = Sin[0] / ss[0]; for (int r = 0; r < 4; ++r) Sout[r] = norm * ss[r]; }
        else { for (int r = 0; r < 4; ++r) Sout[r] = 0.0; if (A) A->error(14); }
        if (A) A->error(15);
        return 0;
    }
    return 16;  // |alpha|<1 but |dn_z|>=1: output undefined in the reference
}

// scattering_angle_sampling :1534-1661.  Returns 0 or the error code (6/7) of a failed search.
int scattering_angle_sampling(const Run& run, const double S[4], int cellidx, double& alpha, double& beta) {
    const Tables& T = run.T;
    Rng& R = *run.R;
    const double* p = &T.p1k[(size_t)T.c2u[cellidx] * 4];
    double cum[181];
    cum[0] = 0.0;
    for (int i = 1; i <= 180; ++i) {  // :1547-1560
        double v = p[0] * S[0] + p[1] * S[1] * T.cos2beta[i - 1] + p[1] * S[2] * T.sin2beta[i - 1]
                 - p[2] * S[1] * T.sin2beta[i - 1] + p[2] * S[2] * T.cos2beta[i - 1] + p[3] * S[3];
        cum[i] = cum[i - 1] + v;
    }
    double xi = R.next();
    double samp = xi * cum[180];
    bool found = false;
    beta = 0.0;
    for (int i = 1; i <= 180; ++i) {  // :1565-1587
        if (samp >= cum[i - 1] && samp <= cum[i]) {
            double x0 = (double)(i - 1), x1 = (double)i, y0 = cum[i - 1], y1 = cum[i];
            beta = (x1 - x0) * (samp - y0) / (y1 - y0) + x0;
            beta = beta * PI / 180.0;
            found = true;
            break;
        }
    }
    xi = R.next();  // :1589
    if (!found) return 6;
    if (xi > 0.5) beta = beta + PI;
    if (beta >= 2.0 * PI) beta = 2.0 * PI - 1.e-10;
    if (beta <= 0.0) beta = -2.0 * PI + 1.e-10;

    double c2b = std::cos(2.0 * beta);
    double s2b = std::sqrt(1.0 - c2b * c2b);
    if (beta > PI / 2.0 && beta < PI) s2b = -s2b;
    else if (beta > 3.0 * PI / 2.0 && beta < 2.0 * PI) s2b = -s2b;
    else if (beta > -PI / 2.0 && beta < 0.0) s2b = -s2b;
    else if (beta > -2.0 * PI && beta < -3.0 * PI / 2.0) s2b = -s2b;

    for (int i = 1; i <= 180; ++i) {  // :1610-1623
        const double* m = T.mat(cellidx, i - 1);
        double v = m[0] * S[0] + m[1] * c2b * S[1] + m[1] * s2b * S[2] - m[2] * s2b * S[1] + m[2] * c2b * S[2] + m[3] * S[3];
        v = v * T.sinbeta[i - 1] * PI / 180.0;
        cum[i] = cum[i - 1] + v;
    }
    xi = R.next();  // :1625
    samp = xi * cum[180];
    found = false;
    alpha = 0.0;
    for (int i = 1; i <= 180; ++i) {  // :1628-1656
        if (samp >= cum[i - 1] && samp <= cum[i]) {
            double x0 = (double)(i - 1), x1 = (double)i, y0 = cum[i - 1], y1 = cum[i];
            alpha = (x1 - x0) * (samp - y0) / (y1 - y0) + x0;
            alpha = std::cos(alpha * PI / 180.0);
            if (std::fabs(alpha) >= 1.0 && run.A) run.A->error(56);
            found = true;
            break;
        }
    }
    if (!found) return 7;
    if (alpha >= 1.0) alpha = 1.0 - 1.e-10;
    if (alpha <= -1.0) alpha = -1.0 + 1.e-10;
    return 0;
}

// Detector deposit :4947-4972 (k4 = true) and :4575-4585 / :4683-4693 (I only).
inline void deposit(const Run& run, double xp, double yp, double zp, const double W[4], bool all4) {
    const artes_launch_t& L = run.L;
    double x_im = yp * run.cos_dp - xp * run.sin_dp;
    double y_im = zp * run.sin_dt - yp * run.cos_dt * run.sin_dp - xp * run.cos_dt * run.cos_dp;
    int ix = (int)(L.nx * (x_im + L.x_max) / (2.0 * L.x_max)) + 1;
    int iy = (int)(L.ny * (y_im + L.y_max) / (2.0 * L.y_max)) + 1;
    if (ix < 1 || ix > L.nx || iy < 1 || iy > L.ny) { run.A->error(60); return; }
    if (run.rec) run.rec->tuple(100, ix, iy, 0, 0);
    const size_t npx = (size_t)L.nx * L.ny;
    double* det = run.A->det.data();
    size_t p = (size_t)(ix - 1) + (size_t)L.nx * (iy - 1);
    int nk = all4 ? 4 : 1;
    for (int k = 0; k < nk; ++k) {
        det[p + npx * (k + 4 * 0)] += W[k];
        det[p + npx * (k + 4 * 1)] += W[k] * W[k];
        det[p + npx * (k + 4 * 2)] += 1.0;
    }
}

// Straight walk to the grid exit or the surface, summing optical depth (:4739-4761 and copies).
// Returns 0 ok-exit, 1 hit surface, 2 cell error.
int peel_walk(const Run& run, double x, double y, double z, const int face_in[2], const int cell_in[3], double& tau_total) {
    const Tables& T = run.T;
    int cf[2] = {face_in[0], face_in[1]};
    int cell[3] = {cell_in[0], cell_in[1], cell_in[2]};
    tau_total = 0.0;
    CellFace o;
    for (;;) {
        cell_face(run, x, y, z, run.det_dir, cf, cell, o);
        if (run.rec) run.rec->tuple(o.nf[0], o.nf[1], o.co[0], o.co[1], o.co[2]);
        if (o.err) return 2;
        double tau_cell = o.dist * T.kext[T.idx(cell)];
        tau_total = tau_total + tau_cell;
        x = x + o.dist * run.det_dir[0];
        y = y + o.dist * run.det_dir[1];
        z = z + o.dist * run.det_dir[2];
        if (o.exit) return 0;
        if (o.nf[0] == 1 && o.nf[1] == T.cell_depth) return 1;
        cf[0] = o.nf[0]; cf[1] = o.nf[1];
        cell[0] = o.co[0]; cell[1] = o.co[1]; cell[2] = o.co[2];
    }
}

// peel_thermal :4519-4598.  Returns true on cell error (photon dropped with error 047).
bool peel_thermal(const Run& run, double x, double y, double z, const double S[4], const int cell[3], const int face[2]) {
    ++run.A->n_peel;
    double tau;
    int r = peel_walk(run, x, y, z, face, cell, tau);
    if (r == 2) { run.A->error(46); return true; }
    if (r == 0 && tau < 50.0) {
        double w = std::exp(-tau) / (4.0 * PI);
        double W[4] = {w * S[0], 0, 0, 0};
        if (W[0] > 0.0 && W[0] < 1.e100) deposit(run, x, y, z, W, false);
        else run.A->error(51);
    }
    return false;
}

// peel_surface :4600-4708
void peel_surface(const Run& run, double x, double y, double z, const double S[4], const int cell_in[3], const int face[2]) {
    const Tables& T = run.T;
    double sn[3] = {x / (T.ox * T.ox), y / (T.oy * T.oy), z / (T.oz * T.oz)};
    double norm = std::sqrt(sn[0] * sn[0] + sn[1] * sn[1] + sn[2] * sn[2]);
    sn[0] = sn[0] / norm; sn[1] = sn[1] / norm; sn[2] = sn[2] / norm;
    // cartesian_spherical of both vectors :4628-4630, :1404-1419
    auto sph = [](const double v[3], double& th, double& ph) {
        double r = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
        th = std::acos(v[2] / r);
        ph = std::atan2(v[1], v[0]);
        if (ph < 0.0) ph = ph + 2.0 * PI;
    };
    double dth, dph, nth, nph;
    sph(run.det_dir, dth, dph);
    sph(sn, nth, nph);
    double cos_angle = std::sin(dth) * std::cos(dph) * std::sin(nth) * std::cos(nph) +
                       std::sin(dth) * std::sin(dph) * std::sin(nth) * std::sin(nph) + std::cos(dth) * std::cos(nth);
    if (cos_angle > 0.0) {
        ++run.A->n_peel;
        int cell[3] = {cell_in[0] + 1, cell_in[1], cell_in[2]};
        double tau;
        int r = peel_walk(run, x, y, z, face, cell, tau);
        if (r == 2) { run.A->error(42); return; }  // deviation: stop the peel instead of walking on
        if (r == 0 && tau < 50.0) {
            double w = std::exp(-tau) * cos_angle / PI;
            double W[4] = {w * S[0], 0, 0, 0};
            if (W[0] > 0.0 && W[0] < 1.e100) deposit(run, x, y, z, W, false);
            else run.A->error(52);
        }
    }
}

// peel_photon :4710-4990.  Returns true when the photon must be dropped (cell_error).
bool peel_photon(const Run& run, double x, double y, double z, const double S[4], const double dir[3],
                 const int cell[3], const int face[2]) {
    const Tables& T = run.T;
    ++run.A->n_peel;
    double tau;
    int r = peel_walk(run, x, y, z, face, cell, tau);
    if (r == 2) { run.A->error(43); return true; }
    if (!(r == 0 && tau < 50.0)) return false;
    const double* det = run.det_dir;
    double w = std::exp(-tau);
    double mu = dir[0] * det[0] + dir[1] * det[1] + dir[2] * det[2];
    if (mu >= 1.0) mu = 1.0 - 1.e-10;
    else if (mu <= -1.0) mu = -1.0 + 1.e-10;
    double F[16];
    matrix_at(T, T.idx(cell), std::acos(mu), F);
    double phi_old = std::atan2(dir[1], dir[0]);
    if (phi_old < 0.0) phi_old = phi_old + 2.0 * PI;
    if (phi_old > 2.0 * PI) phi_old = phi_old - 2.0 * PI;
    double phi_new = std::atan2(det[1], det[0]);
    if (phi_new < 0.0) phi_new = phi_new + 2.0 * PI;
    if (phi_new > 2.0 * PI) phi_new = phi_new - 2.0 * PI;
    if (!(std::fabs(dir[2]) < 1.0)) { run.A->error(45); return false; }  // stokes_out undefined: drop contribution
    double nc = (det[2] - dir[2] * mu) / (std::sqrt(1.0 - mu * mu) * std::sqrt(1.0 - dir[2] * dir[2]));
    double phs;
    if (std::fabs(nc) < 1.0) phs = std::acos(nc);
    else if (nc >= 1.0) phs = 0.0 + 1.e-10;
    else if (nc <= -1.0) phs = PI - 1.e-10;
    else { run.A->error(44); return false; }  // NaN
    if (phi_old - phi_new >= 0.0 && phi_old - phi_new < PI) phs = 2.0 * PI - phs;
    if (2.0 * PI + phi_old - phi_new >= 0.0 && 2.0 * PI + phi_old - phi_new < PI) phs = 2.0 * PI - phs;
    if (phs < 0.0) phs = phs + 2.0 * PI;
    double so[4];
    int e = polarization_rotation(mu, phs, S, F, dir, det, so, true, run.A);
    if (e) { run.A->error(e); return false; }
    if (w * so[0] > 0.0 && w * so[0] < 1.e100) {
        double W[4] = {w * so[0], -(w * so[1]), w * so[2], w * so[3]};  // Q sign flipped :4956
        deposit(run, x, y, z, W, true);
    } else run.A->error(53);
    return false;
}

// add_flow_global :4992-5014
inline void add_flow_global(const Run& run, double x, double y, double z, const double dir[3], double e, double dist, const int cell[3]) {
    double th = std::acos(z / std::sqrt(x * x + y * y + z * z));
    double ph = std::atan2(y, x);
    double rd = std::sin(th) * std::cos(ph) * dir[0] + std::sin(th) * std::sin(ph) * dir[1] + std::cos(th) * dir[2];
    double td = std::cos(th) * std::cos(ph) * dir[0] + std::cos(th) * std::sin(ph) * dir[1] - std::sin(th) * dir[2];
    double pd = -std::sin(ph) * dir[0] + std::cos(ph) * dir[1];
    double* f = &run.A->flow3[(size_t)3 * run.T.idx(cell)];
    f[0] += rd * dist * e; f[1] += td * dist * e; f[2] += pd * dist * e;
}

// lambertian :1369-1402.  Returns direction_cosine's error code.
int lambertian(const Run& run, double x, double y, double z, double dir[3]) {
    const Tables& T = run.T;
    double sn[3] = {x / (T.ox * T.ox), y / (T.oy * T.oy), z / (T.oz * T.oz)};
    double norm = std::sqrt(sn[0] * sn[0] + sn[1] * sn[1] + sn[2] * sn[2]);
    sn[0] = sn[0] / norm; sn[1] = sn[1] / norm; sn[2] = sn[2] / norm;
    double xi = run.R->next();
    double alpha = std::sqrt(xi);
    xi = run.R->next();
    double beta = 2.0 * PI * xi;
    double dn[3];
    int e = direction_cosine(alpha, beta, sn, dn);
    if (e) return e;
    dir[0] = dn[0]; dir[1] = dn[1]; dir[2] = dn[2];
    return 0;
}

// initial_cell :2605-2669 (star photons only)
void initial_cell(const Tables& T, double x, double y, double z, int cell[3]) {
    cell[0] = cell[1] = cell[2] = 0;
    double r = std::sqrt(x * x + y * y + z * z);
    double theta = std::acos(z / r);
    double phi = std::atan2(y, x);
    if (phi < 0.0) phi = phi + 2.0 * PI;
    cell[0] = T.nr - 1;
    for (int j = 0; j < T.nt; ++j)
        if (theta > T.thetafront[j] && theta < T.thetafront[j + 1]) { cell[1] = j; break; }
    for (int j = 0; j < T.np; ++j) {
        if (j < T.np - 1) { if (phi > T.phifront[j] && phi < T.phifront[j + 1]) { cell[2] = j; break; } }
        else if (phi > T.phifront[j] && phi < 2.0 * PI) { cell[2] = j; break; }
    }
}

// rotation_matrix :1270-1326 applied to a vector (axis 2 or 3)
inline void rotate(int axis, double ang, const double v[3], double o[3]) {
    double c = std::cos(ang), s = std::sin(ang);
    if (axis == 2) {
        o[0] = v[0] * c + v[1] * 0.0 + v[2] * s;
        o[1] = v[0] * 0.0 + v[1] * 1.0 + v[2] * 0.0;
        o[2] = v[0] * (-s) + v[1] * 0.0 + v[2] * c;
    } else {
        o[0] = v[0] * c + v[1] * (-s) + v[2] * 0.0;
        o[1] = v[0] * s + v[1] * c + v[2] * 0.0;
        o[2] = v[0] * 0.0 + v[1] * 0.0 + v[2] * 1.0;
    }
}

// emit_photon :1008-1268.  Returns an error code (18..21 from direction_cosine) or 0.
int emit_photon(const Run& run, double pos[3], double dir[3], int face[2], int cell[3], double& bias_weight) {
    const Tables& T = run.T;
    const artes_launch_t& L = run.L;
    Rng& R = *run.R;
    bias_weight = 1.0;
    if (L.photon_source == 1) {
        face[0] = 1; face[1] = T.nr;
        double r_disk, xi;
        if (L.limb_emission) {
            for (;;) { xi = R.next(); r_disk = std::sqrt(xi); if (r_disk > 0.9) break; if (R.exhausted) break; }
        } else { xi = R.next(); r_disk = std::sqrt(xi); }
        xi = R.next();
        double phi_disk = 2.0 * PI * xi;
        double d1 = T.rfront[T.nr] * r_disk * std::sin(phi_disk);
        double d2 = T.rfront[T.nr] * r_disk * std::cos(phi_disk);
        dir[0] = -1.0; dir[1] = 0.0; dir[2] = 0.0;
        pos[0] = std::sqrt(T.rfront[T.nr] * T.rfront[T.nr] - d1 * d1 - d2 * d2);
        pos[1] = d1; pos[2] = d2;
        if (L.stellar_direction) {  // :1080-1111
            double t[3], u[3];
            rotate(2, -(PI / 2.0 - L.theta_star), pos, t);
            rotate(3, L.phi_star, t, u);
            pos[0] = u[0]; pos[1] = u[1]; pos[2] = u[2];
            double td = PI - L.theta_star, pd = PI + L.phi_star;
            if (td < 0.0) td = td + 2.0 * PI;
            if (td > 2.0 * PI) td = td - 2.0 * PI;
            if (pd < 0.0) pd = pd + 2.0 * PI;
            if (pd > 2.0 * PI) pd = pd - 2.0 * PI;
            dir[0] = 1.0 * std::sin(td) * std::cos(pd);  // spherical_cartesian :1421-1432
            dir[1] = 1.0 * std::sin(td) * std::sin(pd);
            dir[2] = 1.0 * std::cos(td);
        }
        initial_cell(T, pos[0], pos[1], pos[2], cell);
        return 0;
    }
    // thermal emission :1117-1266
    face[0] = 0; face[1] = 0;
    double xi = R.next();
    double samp = xi * T.emis_cdf[T.cells - 1];
    double prev = 0.0;
    bool ok = false;
    cell[0] = cell[1] = cell[2] = 0;
    for (int i = T.cell_depth; i < T.nr && !ok; ++i)
        for (int j = 0; j < T.nt && !ok; ++j)
            for (int k = 0; k < T.np; ++k) {
                double e = T.emis_cdf[i + T.nr * (j + T.nt * k)];
                if (samp >= prev && samp <= e) { cell[0] = i; cell[1] = j; cell[2] = k; ok = true; break; }
                prev = e;
            }
    xi = R.next();
    double rs = xi * (T.rfront[cell[0] + 1] - T.rfront[cell[0]]);
    rs = T.rfront[cell[0]] + rs;
    xi = R.next();
    double ct = xi * (T.tcos[cell[1] + 1] - T.tcos[cell[1]]);
    ct = T.tcos[cell[1]] + ct;
    double st = std::sqrt(1.0 - ct * ct);
    double ph = 0.0;
    xi = R.next();
    if (T.np == 1) ph = 2.0 * PI * xi;
    else if (cell[2] < T.np - 1) { ph = xi * (T.phifront[cell[2] + 1] - T.phifront[cell[2]]); ph = T.phifront[cell[2]] + ph; }
    else { ph = xi * (2.0 * PI - T.phifront[cell[2]]); ph = T.phifront[cell[2]] + ph; }
    double cp = std::cos(ph);
    double sp = std::sqrt(1.0 - cp * cp);
    if (ph > PI) sp = -sp;
    pos[0] = rs * st * cp; pos[1] = rs * st * sp; pos[2] = rs * ct;
    pos[0] = T.ox * pos[0]; pos[1] = T.oy * pos[1]; pos[2] = T.oz * pos[2];
    if (L.photon_emission == 1) {
        xi = R.next();
        double alpha = 2.0 * xi - 1.0;
        xi = R.next();
        double beta = 2.0 * PI * xi;
        double cb = std::cos(beta);
        double sb = std::sqrt(1.0 - cb * cb);
        if (beta > PI) sb = -sb;
        dir[0] = std::sqrt(1.0 - alpha * alpha) * cb;
        dir[1] = std::sqrt(1.0 - alpha * alpha) * sb;
        dir[2] = alpha;
    } else {
        xi = R.next();
        double yb = (1.0 + L.photon_bias) * std::tan(PI * xi / 2.0) / std::sqrt(1.0 - L.photon_bias * L.photon_bias);
        double ths = std::acos((1.0 - yb * yb) / (1.0 + yb * yb));
        xi = R.next();
        double beta = 2.0 * PI * xi;
        double ru[3] = {pos[0] / (T.ox * T.ox), pos[1] / (T.oy * T.oy), pos[2] / (T.oz * T.oz)};
        double norm = std::sqrt(ru[0] * ru[0] + ru[1] * ru[1] + ru[2] * ru[2]);
        ru[0] = ru[0] / norm; ru[1] = ru[1] / norm; ru[2] = ru[2] / norm;
        int e = direction_cosine(std::cos(PI - ths), beta, ru, dir);
        if (e) return e;
        bias_weight = (PI * std::sin(ths) * (1.0 + L.photon_bias * std::cos(ths))) / (2.0 * std::sqrt(1.0 - L.photon_bias * L.photon_bias));
    }
    if (std::fabs(dir[2]) >= 1.0) run.A->error(54);
    return 0;
}

struct PhotonResult {
    double x, y, z, S[4];
    int n_scatter;
};

// One iteration of the photon loop, radiative_transfer :546-955.
void photon(const Run& run, PhotonResult* res) {
    const Tables& T = run.T;
    const artes_launch_t& L = run.L;
    Accum& A = *run.A;
    Rng& R = *run.R;
    if (run.stat_emul) emulate_stat_call();  // :549

    double pos[3] = {0, 0, 0}, dir[3], bias_weight = 0.0;
    int cf[2] = {0, 0}, cell[3] = {0, 0, 0};
    double S[4] = {1.0, 0.0, 0.0, 0.0};
    int nsc = 0;
    auto finish = [&]() {
        if (res) { res->x = pos[0]; res->y = pos[1]; res->z = pos[2]; for (int k = 0; k < 4; ++k) res->S[k] = S[k]; res->n_scatter = nsc; }
    };

    ++A.n_emit;
    int e = emit_photon(run, pos, dir, cf, cell, bias_weight);
    if (e) { A.error(e); ++A.n_error; finish(); return; }

    if (L.photon_source == 2) {  // :599-621
        S[0] = S[0] * bias_weight / T.cell_weight[T.idx(cell)];
        A.flux_emitted = A.flux_emitted + S[0];
        if (peel_thermal(run, pos[0], pos[1], pos[2], S, cell, cf)) { A.error(47); ++A.n_error; finish(); return; }
    }

    // ---- optical depth to the grid boundary or the surface :625-656
    CellFace o;
    double tau_first = 0.0;
    {
        double xc = pos[0], yc = pos[1], zc = pos[2];
        int cfc[2] = {cf[0], cf[1]}, cc[3] = {cell[0], cell[1], cell[2]};
        for (;;) {
            cell_face(run, xc, yc, zc, dir, cfc, cc, o);
            if (run.rec) run.rec->tuple(o.nf[0], o.nf[1], o.co[0], o.co[1], o.co[2]);
            if (o.err) { A.error(2); ++A.n_error; finish(); return; }
            double tau_cell = o.dist * T.kext[T.idx(cc)];
            tau_first = tau_first + tau_cell;
            xc = xc + o.dist * dir[0]; yc = yc + o.dist * dir[1]; zc = zc + o.dist * dir[2];
            if (o.exit || (o.nf[0] == 1 && o.nf[1] == T.cell_depth)) break;
            cfc[0] = o.nf[0]; cfc[1] = o.nf[1];
            cc[0] = o.co[0]; cc[1] = o.co[1]; cc[2] = o.co[2];
        }
    }
    const bool to_surface = (o.nf[0] == 1 && o.nf[1] == T.cell_depth);
    double tau;
    if (tau_first < 1.e-6 && !to_surface) { finish(); return; }  // :660-664
    else if (tau_first < 1.e-6 && to_surface) { double xi = R.next(); tau = -std::log(1.0 - xi); }
    else {
        double xi = R.next();
        if (tau_first < 50.0) {
            tau = -std::log(1.0 - xi * (1.0 - std::exp(-tau_first)));
            double f = 1.0 - std::exp(-tau_first);
            for (int k = 0; k < 4; ++k) S[k] = S[k] * f;
        } else tau = -std::log(1.0 - xi);
    }

    // One walk to the next interaction point (:689-778 and :848-941).
    // Returns 0 interaction, 1 grid exit, 2 absorbed by the surface, 3 error.
    auto walk = [&]() -> int {
        double tau_run = 0.0;
        for (;;) {
            cell_face(run, pos[0], pos[1], pos[2], dir, cf, cell, o);
            if (run.rec) run.rec->tuple(o.nf[0], o.nf[1], o.co[0], o.co[1], o.co[2]);
            if (o.err) { A.error(3); return 3; }
            const double kap = T.kext[T.idx(cell)];
            double tau_cell = o.dist * kap;
            if (tau_run + tau_cell > tau) {
                double s = (tau - tau_run) / kap;
                pos[0] = pos[0] + s * dir[0]; pos[1] = pos[1] + s * dir[1]; pos[2] = pos[2] + s * dir[2];
                if (L.flow_global) add_flow_global(run, pos[0], pos[1], pos[2], dir, S[0], s, cell);
                cf[0] = 0; cf[1] = 0;
                return 0;
            }
            pos[0] = pos[0] + o.dist * dir[0]; pos[1] = pos[1] + o.dist * dir[1]; pos[2] = pos[2] + o.dist * dir[2];
            if (L.flow_global) add_flow_global(run, pos[0], pos[1], pos[2], dir, S[0], o.dist, cell);
            if (L.flow_theta) {  // :730-744
                double* f = &A.flow4[(size_t)4 * T.idx(cell)];
                if (o.nf[0] == 1) { if (o.co[0] > cell[0]) f[0] += S[0]; else if (o.co[0] < cell[0]) f[1] += S[0]; }
                else if (o.nf[0] == 2) { if (o.co[1] > cell[1]) f[2] += S[0]; else if (o.co[1] < cell[1]) f[3] += S[0]; }
            }
            cf[0] = o.nf[0]; cf[1] = o.nf[1];
            cell[0] = o.co[0]; cell[1] = o.co[1]; cell[2] = o.co[2];
            if (o.exit) return 1;
            if (o.nf[0] == 1 && o.nf[1] == T.cell_depth) {  // :755-774
                ++A.n_surf;
                double xi = R.next();
                if (xi > L.surface_albedo) return 2;
                double Sold[4] = {S[0], S[1], S[2], S[3]};
                int le = lambertian(run, pos[0], pos[1], pos[2], dir);
                if (le) { A.error(le); return 3; }
                peel_surface(run, pos[0], pos[1], pos[2], Sold, cell, cf);
                S[1] = 0.0; S[2] = 0.0; S[3] = 0.0;
                cell[0] = cell[0] + 1;
            }
            tau_run = tau_run + tau_cell;
        }
    };

    int w = walk();
    if (w == 1 && L.photon_source == 2) A.flux_exit = A.flux_exit + S[0];  // :780
    if (w != 0) { if (w == 3) ++A.n_error; finish(); return; }

    // ---- scattering loop :788-951
    for (;;) {
        if (!L.photon_scattering) break;
        if (R.exhausted) break;  // injected stream used up (test hook only)
        double xi = R.next();
        if (xi < L.fstop) break;
        const int ci = T.idx(cell);
        if (T.albedo[ci] < 1.0 && T.albedo[ci] > 0.0) {
            double gamma = T.albedo[ci] / (1.0 - L.fstop);
            for (int k = 0; k < 4; ++k) S[k] = gamma * S[k];
        }
        if (S[0] <= L.photon_minimum) break;
        if (peel_photon(run, pos[0], pos[1], pos[2], S, dir, cell, cf)) { ++A.n_error; break; }
        // scatter_photon :1434-1532
        ++A.n_sc; ++nsc;
        double alpha, beta, dn[3], F[16], Sn[4];
        e = scattering_angle_sampling(run, S, ci, alpha, beta);
        if (e) { A.error(e); ++A.n_error; break; }
        e = direction_cosine(alpha, beta, dir, dn);
        if (e) { A.error(e); ++A.n_error; break; }
        matrix_at(T, ci, std::acos(alpha), F);
        if (std::fabs(alpha) < 1.0) {
            e = polarization_rotation(alpha, beta, S, F, dir, dn, Sn, false, &A);
            if (e) { A.error(e); ++A.n_error; break; }
            for (int k = 0; k < 4; ++k) S[k] = Sn[k];
            dir[0] = dn[0]; dir[1] = dn[1]; dir[2] = dn[2];
        } else { A.error(50); ++A.n_error; break; }
        xi = R.next();  // :845
        tau = -std::log(1.0 - xi);
        w = walk();
        if (w == 3) { A.error(5); ++A.n_error; }
        if (w != 0) break;
    }
    if (w == 1 && L.photon_source == 2) A.flux_exit = A.flux_exit + S[0];  // :953
    finish();
}

void setup_run(Run& run, const artes_launch_t* L) {
    run.L = *L;
    // spherical_cartesian(1, det_theta, det_phi) :495, :1421-1432 ; :499-502
    run.det_dir[0] = 1.0 * std::sin(L->det_theta) * std::cos(L->det_phi);
    run.det_dir[1] = 1.0 * std::sin(L->det_theta) * std::sin(L->det_phi);
    run.det_dir[2] = 1.0 * std::cos(L->det_theta);
    run.sin_dt = std::sin(L->det_theta); run.cos_dt = std::cos(L->det_theta);
    run.sin_dp = std::sin(L->det_phi); run.cos_dp = std::cos(L->det_phi);
}

}  // namespace

struct artes_ref_ctx {
    Tables T;
};

extern "C" {

artes_ref_ctx* artes_ref_create(void) { return new artes_ref_ctx(); }
void artes_ref_destroy(artes_ref_ctx* c) { delete c; }

int artes_ref_set_grid(artes_ref_ctx* c, int nr, int ntheta, int nphi, const double* rfront,
                       const double* thetafront, const int32_t* thetaplane, const double* phifront,
                       double ox, double oy, double oz) {
    if (!c || nr < 1 || ntheta < 1 || nphi < 1) return -1;
    Tables& T = c->T;
    T.nr = nr; T.nt = ntheta; T.np = nphi; T.cells = nr * ntheta * nphi;
    T.rfront.assign(rfront, rfront + nr + 1);
    T.thetafront.assign(thetafront, thetafront + ntheta + 1);
    T.thetaplane.assign(thetaplane, thetaplane + ntheta + 1);
    T.phifront.assign(phifront, phifront + nphi);
    T.ox = ox; T.oy = oy; T.oz = oz;
    T.tcos.resize(ntheta + 1); T.ttan.resize(ntheta + 1);
    for (int i = 0; i <= ntheta; ++i) { T.tcos[i] = std::cos(T.thetafront[i]); T.ttan[i] = std::tan(T.thetafront[i]); }  // :2261-2264
    T.pcos.resize(nphi); T.psin.resize(nphi);
    for (int i = 0; i < nphi; ++i) { T.pcos[i] = std::cos(T.phifront[i]); T.psin[i] = std::sin(T.phifront[i]); }        // :2267-2270
    for (int i = 1; i <= 180; ++i) {  // :409-420
        T.cosbeta[i - 1] = (std::cos((double)i * PI / 180.0) + std::cos((double)(i - 1) * PI / 180.0)) / 2.0;
        T.cosbeta[i + 179] = -T.cosbeta[i - 1];
        T.sinbeta[i - 1] = (std::sin((double)i * PI / 180.0) + std::sin((double)(i - 1) * PI / 180.0)) / 2.0;
        T.sinbeta[i + 179] = -T.sinbeta[i - 1];
        T.cos2beta[i - 1] = (std::cos(2.0 * (double)i * PI / 180.0) + std::cos(2.0 * (double)(i - 1) * PI / 180.0)) / 2.0;
        T.cos2beta[i + 179] = T.cos2beta[i - 1];
        T.sin2beta[i - 1] = (std::sin(2.0 * (double)i * PI / 180.0) + std::sin(2.0 * (double)(i - 1) * PI / 180.0)) / 2.0;
        T.sin2beta[i + 179] = T.sin2beta[i - 1];
    }
    return 0;
}

int artes_ref_set_wavelength(artes_ref_ctx* c, const double* k_sca, const double* k_abs, int n_uniq,
                             const double* uniq_matrix, const int32_t* cell_to_uniq, int cell_depth,
                             const double* cell_weight, const double* emis_cdf) {
    if (!c || c->T.cells == 0 || n_uniq < 1) return -1;
    Tables& T = c->T;
    const int n = T.cells;
    T.ksca.assign(k_sca, k_sca + n);
    T.kabs.assign(k_abs, k_abs + n);
    T.kext.assign(n, 0.0);
    T.albedo.assign(n, 0.0);
    for (int i = 0; i < n; ++i) {  // :2178-2188
        T.kext[i] = T.ksca[i] + T.kabs[i];
        if (T.kext[i] > 0.0) T.albedo[i] = T.ksca[i] / T.kext[i];
        if (T.albedo[i] < 1.e-20) T.albedo[i] = 1.e-20;
    }
    T.n_uniq = n_uniq;
    T.M.assign(uniq_matrix, uniq_matrix + (size_t)n_uniq * 180 * 16);
    T.c2u.assign(cell_to_uniq, cell_to_uniq + n);
    for (int i = 0; i < n; ++i) if (T.c2u[i] < 0 || T.c2u[i] >= n_uniq) return -2;
    T.p1k.assign((size_t)n_uniq * 4, 0.0);
    for (int u = 0; u < n_uniq; ++u)  // :2215-2230
        for (int a = 0; a < 180; ++a)
            for (int k = 0; k < 4; ++k)
                T.p1k[(size_t)u * 4 + k] = T.p1k[(size_t)u * 4 + k] + T.M[((size_t)u * 180 + a) * 16 + k] * T.sinbeta[a] * PI / 180.0;
    T.cell_depth = cell_depth;
    T.thermal = (cell_weight && emis_cdf);
    if (T.thermal) { T.cell_weight.assign(cell_weight, cell_weight + n); T.emis_cdf.assign(emis_cdf, emis_cdf + n); }
    return 0;
}

int artes_ref_cell_depth(const artes_ref_ctx* c, int photon_source, int ring) {
    // grid_initialize(2) :2329-2393
    const Tables& T = c->T;
    int cell_max = 1000000, cell_depth = 0;
    const int grid_out = (photon_source == 2 && ring) ? 2 : 0;
    const double limit = (photon_source == 1) ? 30.0 : 5.0;
    const std::vector<double>& kap = (photon_source == 1) ? T.kext : T.kabs;
    for (int j = 0; j < T.nt; ++j)
        for (int k = 0; k < T.np; ++k) {
            double tot = 0.0;
            for (int i = grid_out; i < T.nr; ++i) {
                tot = tot + kap[(T.nr - i - 1) + T.nr * (j + T.nt * k)] * (T.rfront[T.nr - i] - T.rfront[T.nr - i - 1]);
                cell_depth = T.nr - i - 1;
                if (tot > limit) break;
            }
            if (cell_depth < cell_max) cell_max = cell_depth;
        }
    return cell_max;
}

// ---- host side of radiative_transfer / write_output (the product's driver and Python mirror are checked against these) ----

// planck_function :1350-1367 (wavelength in metres)
double artes_ref_planck(double temperature, double wavelength, int photon_source) {
    const double pi = 4.0 * std::atan(1.0);
    const double k_b = 1.3806488e-23, hh = 6.62606957e-34, cc = 2.99792458e8;      // :9-16
    if (photon_source == 1) return (2.0 * pi * hh * cc * cc / std::pow(wavelength, 5.0)) / (std::exp(hh * cc / (wavelength * k_b * temperature)) - 1.0);
    return (2.0 * hh * cc * cc / std::pow(wavelength, 5.0)) / (std::exp(hh * cc / (wavelength * k_b * temperature)) - 1.0);
}

// photon_package :2509-2539
double artes_ref_package_energy(int photon_source, double t_star, double r_star, double orbit, double distance_planet, double rfront_nr,
                                double wavelength, double packages, int phase_curve, double det_phi, double emissivity_total) {
    const double pi = 4.0 * std::atan(1.0);
    double package_energy = 0.0;
    if (photon_source == 1) {
        const double planck_flux = artes_ref_planck(t_star, wavelength, 1);
        package_energy = pi * planck_flux * rfront_nr * rfront_nr * r_star * r_star / (orbit * orbit * distance_planet * distance_planet * packages);
        if (phase_curve && det_phi * 180.0 / pi >= 170.0)
            package_energy = package_energy * (pi * r_star * r_star - 0.9 * 0.9 * pi * r_star * r_star) / (pi * r_star * r_star);
    } else if (photon_source == 2) {
        package_energy = emissivity_total / (distance_planet * distance_planet * packages);
    }
    return package_energy;
}

// The tail of radiative_transfer :957-1004: detector(nx,ny,4,3) from the thread sum and the package energy, photometry(11).
// det_sum / detector are in the reference's array order (ix fastest, then iy, Stokes index, l); sums run in that order like
// the Fortran intrinsic.  photometry(10) = PI / I is 0 for I = 0 (the reference divides by zero there).
void artes_ref_finish_detector(int nx, int ny, const double* det_sum, double package_energy, double* detector, double* photometry) {
    const size_t npx = (size_t)nx * ny;
    for (int l = 1; l <= 3; ++l)
        for (int k = 0; k < 4; ++k)
            for (size_t i = 0; i < npx; ++i) {
                const size_t x = ((size_t)(l - 1) * 4 + k) * npx + i;
                if (l == 1) detector[x] = det_sum[x] * package_energy;
                else if (l == 2) detector[x] = det_sum[x] * package_energy * package_energy;
                else detector[x] = det_sum[x];
            }
    auto plane_sum = [&](int k, int l) { double t = 0.0; const double* p = detector + ((size_t)(l - 1) * 4 + k) * npx; for (size_t i = 0; i < npx; ++i) t = t + p[i]; return t; };
    for (int i = 0; i < 11; ++i) photometry[i] = 0.0;
    photometry[0] = plane_sum(0, 1); photometry[2] = plane_sum(1, 1); photometry[4] = plane_sum(2, 1); photometry[6] = plane_sum(3, 1);
    photometry[8] = std::sqrt(plane_sum(1, 1) * plane_sum(1, 1) + plane_sum(2, 1) * plane_sum(2, 1));
    photometry[9] = photometry[0] != 0.0 ? photometry[8] / photometry[0] : 0.0;
    for (int i = 1; i <= 4; ++i) {
        if (plane_sum(i - 1, 3) > 0.0) {
            const double dummy = (plane_sum(i - 1, 2) / plane_sum(i - 1, 3)) - std::pow(plane_sum(i - 1, 1) / plane_sum(i - 1, 3), 2);
            if (dummy > 0.0) photometry[i * 2 - 1] = std::sqrt(dummy) * std::sqrt(plane_sum(i - 1, 3));
        }
    }
    if (photometry[2] * photometry[2] + photometry[4] * photometry[4] > 0.0) {
        const double dpi = std::sqrt((std::pow(photometry[2] * photometry[3], 2) + std::pow(photometry[4] * photometry[5], 2)) /
                                     (2.0 * (photometry[2] * photometry[2] + photometry[4] * photometry[4])));
        photometry[10] = photometry[9] * std::sqrt(std::pow(dpi / photometry[8], 2) + std::pow(photometry[1] / photometry[0], 2));
    }
}

// write_output :3481-3519: error(nx,ny,5) = sigma of I, Q, U, V and of the degree of polarisation.  A pixel with Q = U = 0
// reads `pol` and `dpol` uninitialised / stale in the reference (:3506-3516); here its sigma_P is 0.
void artes_ref_stokes_error(int nx, int ny, const double* detector, double* error) {
    const size_t npx = (size_t)nx * ny;
    auto det = [&](size_t i, int k, int l) { return detector[((size_t)(l - 1) * 4 + (k - 1)) * npx + i]; };
    for (size_t i = 0; i < 5 * npx; ++i) error[i] = 0.0;
    for (int k = 1; k <= 4; ++k)
        for (size_t i = 0; i < npx; ++i)
            if (det(i, k, 3) > 0.0) {
                const double dummy = (det(i, k, 2) / det(i, k, 3)) - std::pow(det(i, k, 1) / det(i, k, 3), 2);
                if (dummy > 0.0) error[(size_t)(k - 1) * npx + i] = std::sqrt(dummy) * std::sqrt(det(i, k, 3));
            }
    for (size_t i = 0; i < npx; ++i) {
        const double q = det(i, 2, 1), u = det(i, 3, 1);
        if (q * q + u * u > 0.0) {
            const double pol = std::sqrt(q * q + u * u);
            const double dpol = std::sqrt((std::pow(q * error[npx + i], 2) + std::pow(u * error[2 * npx + i], 2)) / (2.0 * (q * q + u * u)));
            if (det(i, 1, 1) > 0.0)
                error[4 * npx + i] = (pol / det(i, 1, 1)) * std::sqrt(std::pow(dpol / pol, 2) + std::pow(error[i] / det(i, 1, 1), 2));
        }
    }
}

int artes_ref_run(artes_ref_ctx* c, const artes_launch_t* L, int rng_kind, int nthreads, int emulate_stat,
                  double* det_sum, double* flux, double* flow4, double* flow3, uint64_t* err_hist,
                  artes_stats_t* stats) {
    if (!c || !L || L->struct_size != sizeof(artes_launch_t)) return -1;
    const Tables& T = c->T;
    if (L->photon_source == 2 && !T.thermal) return -3;
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
    nthreads = 1;
#endif
    const size_t ndet = (size_t)L->nx * L->ny * 12;
    std::vector<Accum> acc(nthreads);
    for (auto& a : acc) {
        a.det.assign(ndet, 0.0);
        if (L->flow_theta) a.flow4.assign((size_t)4 * T.cells, 0.0);
        if (L->flow_global) a.flow3.assign((size_t)3 * T.cells, 0.0);
    }
    std::vector<uint64_t> draws(nthreads, 0);
    const int64_t N = (int64_t)L->n_photons;
    auto t0 = std::chrono::steady_clock::now();
#pragma omp parallel num_threads(nthreads)
    {
#ifdef _OPENMP
        const int tid = omp_get_thread_num();
#else
        const int tid = 0;
#endif
        Run run(T);
        setup_run(run, L);
        run.stat_emul = emulate_stat != 0;
        run.A = &acc[tid];
        Rng rng;
        rng.kind = rng_kind;
        rng.seed = L->seed;
        rng.s1 = (int32_t)((L->seed * 7919ull + 104729ull * (uint64_t)tid) % 1000000ull);  // stands in for :443-444
        // Streams of this generator that differ only in s1 (all the reference ever varies: s1 = int(xi * 1e6), s2..s4 are
        // constants, :114, :443-446) are far from independent: (i) the difference of two such subtract-with-borrow
        // sequences obeys d(n) = d(n-3) - d(n-1) and grows by only 1.15x per draw from |ds1| * 2^-31, so their first ~60
        // draws coincide -- 32 runs x 8 threads put all 20 329 deposits of their first packets into ONE pixel; (ii) the
        // congruential half s4 is the SAME number in every stream at the same draw index, which pins draw n of all streams
        // into one common half of (0,1): averaging over streams does not cancel that, only averaging along a stream does,
        // so runs made of many SHORT streams are biased (C2 phase point at 30 deg: +1.3 % = 5 sigma with 94 packets per
        // stream, gone at 12 500 packets per stream, where the generator agrees with Philox to 0.02 +- 0.02 %).  The
        // reference drowns both in 1e6+ packets per thread; the statistical tests, made of many small runs, cannot.  For
        // test batches the oracle therefore starts every thread's generator at its own point: the congruential half at a
        // seed- and thread-dependent phase (every 32-bit value is a phase of that full-period sequence) and the whole
        // generator 512..8703 draws ahead.  Same generator, decorrelated streams.
        if (rng_kind == 0) {
            const uint64_t hsh = (L->seed + 1ull) * 0x9e3779b97f4a7c15ull + (uint64_t)(tid + 1) * 0xbf58476d1ce4e5b9ull;
            uint64_t z = hsh ^ (hsh >> 31); z *= 0x94d049bb133111ebull; z ^= z >> 29;
            rng.s4 = (int32_t)(uint32_t)(z >> 16);
            const int burn = 512 + (int)((hsh >> 40) % 8192ull);
            for (int w = 0; w < burn; ++w) (void)rng.next();
            rng.total = 0; rng.err55 = 0;
        }
        run.R = &rng;
#pragma omp for schedule(static)
        for (int64_t i = 0; i < N; ++i) {
            rng.start_photon(L->photon_id_base + (uint64_t)i, nullptr);
            photon(run, nullptr);
        }
        draws[tid] = rng.total;
        acc[tid].err[55] += rng.err55;
    }
    auto t1 = std::chrono::steady_clock::now();
    // thread sum :959-975 (without package_energy)
    if (det_sum) { for (size_t p = 0; p < ndet; ++p) { double s = 0.0; for (int t = 0; t < nthreads; ++t) s += acc[t].det[p]; det_sum[p] = s; } }
    if (flux) { flux[0] = flux[1] = 0.0; for (int t = 0; t < nthreads; ++t) { flux[0] += acc[t].flux_emitted; flux[1] += acc[t].flux_exit; } }
    if (flow4) { for (size_t p = 0; p < (size_t)4 * T.cells; ++p) { double s = 0.0; if (L->flow_theta) for (int t = 0; t < nthreads; ++t) s += acc[t].flow4[p]; flow4[p] = s; } }
    if (flow3) { for (size_t p = 0; p < (size_t)3 * T.cells; ++p) { double s = 0.0; if (L->flow_global) for (int t = 0; t < nthreads; ++t) s += acc[t].flow3[p]; flow3[p] = s; } }
    if (err_hist) { for (int k = 0; k < ARTES_ERR_SLOTS; ++k) { err_hist[k] = 0; for (int t = 0; t < nthreads; ++t) err_hist[k] += acc[t].err[k]; } }
    if (stats) {
        std::memset(stats, 0, sizeof(*stats));
        for (int t = 0; t < nthreads; ++t) {
            stats->n_emit += acc[t].n_emit; stats->n_cell_face += acc[t].n_cf; stats->n_scatter += acc[t].n_sc;
            stats->n_peel += acc[t].n_peel; stats->n_surface += acc[t].n_surf; stats->n_error += acc[t].n_error;
            stats->n_draws += draws[t];
        }
        stats->kernel_ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
        stats->reserved = (uint64_t)nthreads;
    }
    return 0;
}

int artes_ref_trace(artes_ref_ctx* c, const artes_launch_t* L, const double* xi, uint64_t n, int max_draws,
                    int32_t* seq_len, uint64_t* seq_hash, int32_t* seq_head, int max_rec, double* fstate) {
    if (!c || !L || L->struct_size != sizeof(artes_launch_t)) return -1;
    const Tables& T = c->T;
    if (L->photon_source == 2 && !T.thermal) return -3;
    Run run(T);
    setup_run(run, L);
    Accum acc;
    acc.det.assign((size_t)L->nx * L->ny * 12, 0.0);
    if (L->flow_theta) acc.flow4.assign((size_t)4 * T.cells, 0.0);
    if (L->flow_global) acc.flow3.assign((size_t)3 * T.cells, 0.0);
    run.A = &acc;
    Rng rng;
    rng.kind = 2;
    rng.max_draws = max_draws;
    run.R = &rng;
    for (uint64_t i = 0; i < n; ++i) {
        Recorder rec;
        rec.head = (seq_head && max_rec > 0) ? seq_head + (size_t)i * max_rec * 5 : nullptr;
        rec.max_rec = max_rec;
        run.rec = &rec;
        rng.start_photon(L->photon_id_base + i, xi + (size_t)i * max_draws);
        PhotonResult res;
        photon(run, &res);
        if (seq_len) seq_len[i] = rec.len;
        if (seq_hash) seq_hash[i] = rec.hash;
        if (fstate) {
            double* f = fstate + (size_t)i * 8;
            f[0] = res.x; f[1] = res.y; f[2] = res.z; f[3] = res.S[0]; f[4] = res.S[1]; f[5] = res.S[2]; f[6] = res.S[3]; f[7] = (double)res.n_scatter;
        }
    }
    return 0;
}

int artes_ref_cell_face(artes_ref_ctx* c, uint64_t n, const double* pos, const double* dir, const int32_t* face,
                        const int32_t* cell, int32_t* out_i, double* out_d) {
    if (!c) return -1;
    Run run(c->T);
    run.A = nullptr;
    for (uint64_t i = 0; i < n; ++i) {
        int cf[2] = {face[2 * i], face[2 * i + 1]};
        int ce[3] = {cell[3 * i], cell[3 * i + 1], cell[3 * i + 2]};
        CellFace o;
        cell_face(run, pos[3 * i], pos[3 * i + 1], pos[3 * i + 2], dir + 3 * i, cf, ce, o);
        int32_t* p = out_i + 7 * i;
        p[0] = o.nf[0]; p[1] = o.nf[1]; p[2] = o.co[0]; p[3] = o.co[1]; p[4] = o.co[2]; p[5] = o.exit ? 1 : 0; p[6] = o.err ? o.code : 0;
        out_d[i] = o.dist;
    }
    return 0;
}

int artes_ref_scatter(artes_ref_ctx* c, uint64_t n, const double* stokes, const double* dir, const int32_t* cell_idx,
                      const double* xi, double* out) {
    if (!c) return -1;
    Run run(c->T);
    Accum acc;
    run.A = &acc;
    Rng rng;
    rng.kind = 2; rng.max_draws = 3;
    run.R = &rng;
    for (uint64_t i = 0; i < n; ++i) {
        rng.start_photon(i, xi + 3 * i);
        double alpha = 0, beta = 0, dn[3] = {0, 0, 0}, F[16], Sn[4] = {0, 0, 0, 0};
        double* o = out + 9 * i;
        for (int k = 0; k < 9; ++k) o[k] = NAN;
        int e = scattering_angle_sampling(run, stokes + 4 * i, cell_idx[i], alpha, beta);
        if (e) continue;
        o[0] = alpha; o[1] = beta;
        e = direction_cosine(alpha, beta, dir + 3 * i, dn);
        if (e) continue;
        o[2] = dn[0]; o[3] = dn[1]; o[4] = dn[2];
        matrix_at(c->T, cell_idx[i], std::acos(alpha), F);
        e = polarization_rotation(alpha, beta, stokes + 4 * i, F, dir + 3 * i, dn, Sn, false, &acc);
        if (e) continue;
        o[5] = Sn[0]; o[6] = Sn[1]; o[7] = Sn[2]; o[8] = Sn[3];
    }
    return 0;
}

void artes_ref_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) { philox4x32_10(ctr, key, out); }

void artes_ref_philox_uniforms(uint64_t seed, uint64_t id, int n, double* out) {
    Rng r; r.kind = 1; r.seed = seed; r.start_photon(id, nullptr);
    for (int i = 0; i < n; ++i) out[i] = r.next();
}

void artes_ref_mz_uniforms(int32_t s1, int n, double* out) {
    Rng r; r.kind = 0; r.s1 = s1;
    for (int i = 0; i < n; ++i) out[i] = r.next();
}

}  // extern "C"
