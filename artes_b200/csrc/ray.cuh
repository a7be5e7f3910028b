// ray.cuh -- fast-mode geometry: incremental ray marching in an absolute ray parameter.
// Included from transport.cuh in the fast translation unit only.
//
// The reference re-solves every bounding surface of the current cell from the current position at every
// crossing (cell_face, src/ARTES.f90:2800-3470: 4-8 quadratics per step).  Along one straight ray
// X(t) = X0 + t n all those surfaces are fixed quadrics in t, so the march only has to keep, per axis
// (r, theta, phi), the parameter of the next crossing and re-solve the ONE axis whose cell index just
// changed: spheres need one square root and no division (1/qa is a ray constant), cones one square root
// and two divisions, half-planes one division.  Roots are compared with the current parameter by
// exact inequality: the surface just crossed reproduces its root bit for bit, so no distance
// thresholds (1e-9 / 1e-12 / 1e-15 of :3358-3392) are needed; only a ray that STARTS on a face excludes
// that face's roots below 1e-3 m, the reference's own same-face threshold (:2944, :3157).
// The crossing sequence is the reference's except when two faces are crossed within ~1e-8 m.

struct Ray {
    double x0, y0, z0, n0, n1, n2;     // origin and direction
    double t;                          // current parameter
    double A1, A2, B1, B2, C1, C2;     // a^2 n.n, a^2 x.n, a^2 x.x split in (xy | z) parts, oblate-scaled
    double inv_qs;                     // 1 / (A1 + A2)
    double tr, tt, tp;                 // parameter of the next radial / polar / azimuthal crossing
    int fr, ft, fp;                    // face index of that crossing
    int dr, dt, dp;                    // -1: towards the lower cell index, +1: towards the higher one
    int sf0, sf1;                      // face the ray started on (0,0 = interior point)
};

constexpr double RAY_NONE = 1.0e300;

__device__ __forceinline__ double start_floor(const Ray& R, int axis, int face) {
    // roots of the face the ray started on must lie beyond the reference's same-face threshold
    return (R.sf0 == axis && R.sf1 == face) ? 1.e-3 : -RAY_NONE;
}

// smallest root of sphere k beyond the current parameter
__device__ __forceinline__ double sphere_next(const Ray& R, double r, double tfloor) {
    const double hb = R.B1 + R.B2, qa = R.A1 + R.A2, qc = (R.C1 + R.C2) - r * r;
    const double disc = hb * hb - qa * qc;
    if (!(disc >= 0.0)) return RAY_NONE;
    const double sq = sqrt(disc);
    const double t1 = (-hb - sq) * R.inv_qs, t2 = (-hb + sq) * R.inv_qs;
    const double lim = fmax(R.t, tfloor);
    if (t1 > lim) return t1;
    if (t2 > lim) return t2;
    return RAY_NONE;
}

// smallest valid root of cone (tan^2 = T2, hemisphere sign hs = +1 north / -1 south / 0 exactly 90 deg)
__device__ __forceinline__ double cone_next(const Ray& R, double T2, int hs, double tfloor) {
    const double qa = R.A1 - R.A2 * T2, hb = R.B1 - R.B2 * T2, qc = R.C1 - R.C2 * T2;
    const double disc = hb * hb - qa * qc;
    if (!(disc >= 0.0)) return RAY_NONE;
    const double q = -(hb + copysign(sqrt(disc), hb));
    double t1 = (fabs(qa) > 1.e-100) ? q / qa : RAY_NONE;
    double t2 = (fabs(q) > 1.e-100) ? qc / q : RAY_NONE;
    const double lim = fmax(R.t, tfloor);
    // mirror nappe: the quadric also contains the cone of the other hemisphere (:3036-3052)
    if (!(t1 > lim) || (R.z0 + t1 * R.n2) * (double)hs < 0.0) t1 = RAY_NONE;
    if (!(t2 > lim) || (R.z0 + t2 * R.n2) * (double)hs < 0.0) t2 = RAY_NONE;
    return fmin(t1, t2);
}

__device__ __forceinline__ void radial_update(const double* __restrict__ sm, Ray& R, int c0) {
    const double ti = (c0 >= 0) ? sphere_next(R, sm[c0], start_floor(R, 1, c0)) : RAY_NONE;
    const double to = sphere_next(R, sm[c0 + 1], start_floor(R, 1, c0 + 1));
    if (ti <= to) { R.tr = ti; R.fr = c0; R.dr = -1; } else { R.tr = to; R.fr = c0 + 1; R.dr = 1; }
}

__device__ __forceinline__ double polar_face(const double* __restrict__ sm, const SmLayout& lay, const Ray& R, int k) {
    const int* tplane = reinterpret_cast<const int*>(sm + lay.o_tp);
    const double fl = start_floor(R, 2, k);
    if (tplane[k] == 1) {
        const double tn = sm[lay.o_tt + k], tf = sm[lay.o_tf + k];
        const int hs = (tf < PI / 2.0) ? 1 : ((tf > PI / 2.0) ? -1 : 0);
        return cone_next(R, tn * tn, hs, fl);
    }
    // equatorial plane z = 0 (:3068, :3118)
    const double t = -R.z0 / R.n2;
    return (t > fmax(R.t, fl)) ? t : RAY_NONE;
}

__device__ __forceinline__ void polar_update(const double* __restrict__ sm, const SmLayout& lay, const DevTables& T, Ray& R, int c1) {
    const double ti = (c1 != 0) ? polar_face(sm, lay, R, c1) : RAY_NONE;
    const double to = (c1 + 1 != T.nt) ? polar_face(sm, lay, R, c1 + 1) : RAY_NONE;
    if (ti <= to) { R.tt = ti; R.ft = c1; R.dt = -1; } else { R.tt = to; R.ft = c1 + 1; R.dt = 1; }
}

__device__ __forceinline__ double plane_next(const double* __restrict__ sm, const SmLayout& lay, const DevTables& T, const Ray& R, int k) {
    const double a = 1.0 / T.ox, b = 1.0 / T.oy;
    const double ps = sm[lay.o_ps + k], pc = sm[lay.o_pc + k];
    const double den = b * R.n1 * pc - a * R.n0 * ps;
    const double t = (a * R.x0 * ps - b * R.y0 * pc) / den;      // full plane through the z axis (:3300-3348)
    return (t > fmax(R.t, start_floor(R, 3, k))) ? t : RAY_NONE;  // den == 0 -> inf/nan -> no crossing
}

__device__ __forceinline__ void azimuthal_update(const double* __restrict__ sm, const SmLayout& lay, const DevTables& T, Ray& R, int c2) {
    if (T.np <= 1) { R.tp = RAY_NONE; R.fp = 0; R.dp = 1; return; }
    const int ku = (c2 + 1 == T.np) ? 0 : c2 + 1;
    const double ti = plane_next(sm, lay, T, R, c2);
    const double to = plane_next(sm, lay, T, R, ku);
    if (ti <= to) { R.tp = ti; R.fp = c2; R.dp = -1; } else { R.tp = to; R.fp = ku; R.dp = 1; }
}

// start a ray at the walker position of P along (n0,n1,n2)
__device__ __forceinline__ void ray_setup(const Ctx& X, const Photon& P, Ray& R, double n0, double n1, double n2) {
    const DevTables& T = X.A.T;
    const double a = 1.0 / T.ox, b = 1.0 / T.oy, c = 1.0 / T.oz;
    R.x0 = P.wx; R.y0 = P.wy; R.z0 = P.wz; R.n0 = n0; R.n1 = n1; R.n2 = n2;
    R.t = 0.0;
    R.A1 = a * a * n0 * n0 + b * b * n1 * n1; R.A2 = c * c * n2 * n2;
    R.B1 = a * a * P.wx * n0 + b * b * P.wy * n1; R.B2 = c * c * P.wz * n2;
    R.C1 = a * a * P.wx * P.wx + b * b * P.wy * P.wy; R.C2 = c * c * P.wz * P.wz;
    R.inv_qs = 1.0 / (R.A1 + R.A2);
    R.sf0 = P.wf0; R.sf1 = P.wf1;
    radial_update(X.sm, R, P.wc0);
    polar_update(X.sm, X.lay, T, R, P.wc1);
    azimuthal_update(X.sm, X.lay, T, R, P.wc2);
}

// geometry of the next crossing: fills `o` like cell_face does; `axis` = 0 r, 1 theta, 2 phi
__device__ __forceinline__ void ray_next(const DevTables& T, const Photon& P, const Ray& R, CellFace& o, int& axis) {
    double tn = R.tr; axis = 0;
    if (R.tt < tn) { tn = R.tt; axis = 1; }
    if (R.tp < tn) { tn = R.tp; axis = 2; }
    o.err = 0; o.exit = false;
    o.co0 = P.wc0; o.co1 = P.wc1; o.co2 = P.wc2;
    if (!(tn < RAY_NONE)) { o.err = 31; o.nf0 = 0; o.nf1 = 0; o.co0 = 0; o.co1 = 0; o.co2 = 0; o.dist = 1.e100; return; }
    o.dist = tn - R.t;
    o.nf0 = axis + 1;
    if (axis == 0) { o.nf1 = R.fr; o.co0 = P.wc0 + R.dr; o.exit = (R.fr == T.nr); }
    else if (axis == 1) { o.nf1 = R.ft; o.co1 = P.wc1 + R.dt; }
    else { o.nf1 = R.fp; o.co2 = (R.dp > 0) ? ((P.wc2 + 1 == T.np) ? 0 : P.wc2 + 1) : ((P.wc2 == 0) ? T.np - 1 : P.wc2 - 1); }
}

// after the photon moved across `axis` (P.wc* already updated): advance the parameter, re-solve that axis
__device__ __forceinline__ void ray_advance(const Ctx& X, const Photon& P, Ray& R, int axis) {
    R.t = (axis == 0) ? R.tr : ((axis == 1) ? R.tt : R.tp);
    if (axis == 0) radial_update(X.sm, R, P.wc0);
    else if (axis == 1) polar_update(X.sm, X.lay, X.A.T, R, P.wc1);
    else azimuthal_update(X.sm, X.lay, X.A.T, R, P.wc2);
}
