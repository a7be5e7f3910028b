// ray.cuh -- fast-mode geometry: incremental ray marching in an absolute ray parameter.
// Included from engine.cuh in the fast translation unit only.
//
// The reference re-solves every bounding surface of the current cell from the current position at every
// crossing (cell_face, src/ARTES.f90:2800-3470: 4-8 quadratics per step).  Along one straight ray
// X(t) = X0 + t n all those surfaces are fixed quadrics  qa t^2 + 2 hb t + qc = 0  in t, so the march only
// keeps, per axis (r, theta, phi), the parameter of the next crossing and re-solves the ONE axis whose
// cell index just changed: two faces, i.e. two quadric solves per step instead of 4-8.
//   sphere r_k      qa = A1 + A2        hb = B1 + B2         qc = C1 + C2 - r_k^2          (:2891-2893)
//   cone   theta_k  qa = A1 - A2 T      hb = B1 - B2 T       qc = C1 - C2 T,  T = tan^2    (:3030-3032)
//   plane  z = 0    qa = 0              hb = n_z / 2         qc = z0                       (:3068)
//   plane  phi_k    qa = 0              hb = den / 2         qc = -num                     (:3302-3303)
// with A1 = a^2 n_x^2 + b^2 n_y^2, A2 = c^2 n_z^2, B1 = a^2 x n_x + b^2 y n_y, ... (oblate scaling a, b, c).
// One generic solver serves all of them, so every lane of a warp runs the same instruction stream whatever
// axis it just crossed.  Roots are compared with the current parameter by exact inequality: the surface just
// crossed reproduces its root bit for bit, so none of the distance thresholds of :3358-3392 is needed; only
// a ray that STARTS on a face excludes that face's roots below 1e-3 m, the reference's own same-face
// threshold (:2944, :3157).  The crossing sequence equals the reference's unless two faces are crossed
// within ~1e-8 m of each other.

struct Ray {
    double x0, y0, z0;                 // origin
    double t;                          // current parameter
    double A1, A2, B1, B2, C1, C2;     // quadric constants of this ray
    double tr, tt, tp;                 // parameter of the next radial / polar / azimuthal crossing
    int fr, ft, fp;                    // face index of that crossing
    int dirs;                          // bit a set: the crossing on axis a leads to the higher cell index
    int sf0, sf1;                      // face the ray started on (0,0 = interior point)
};

constexpr double RAY_NONE = 1.0e300;

// smallest root > lim of qa t^2 + 2 hb t + qc = 0 whose point lies on the right nappe (hs = +1: z >= 0,
// -1: z <= 0, 0: no test); stable form of quadratic_equation :4154-4173
__device__ __forceinline__ double quadric_next(double qa, double hb, double qc, int hs, double lim, double z0, double n2) {
    const double disc = hb * hb - qa * qc;
    if (!(disc >= 0.0)) return RAY_NONE;
    const double q = -(hb + copysign(sqrt(disc), hb));
    double t1 = (fabs(qa) > 1.e-100) ? q / qa : RAY_NONE;
    double t2 = (fabs(q) > 1.e-100) ? qc / q : RAY_NONE;
    if (!(t1 > lim) || (z0 + t1 * n2) * (double)hs < 0.0) t1 = RAY_NONE;
    if (!(t2 > lim) || (z0 + t2 * n2) * (double)hs < 0.0) t2 = RAY_NONE;
    return fmin(t1, t2);
}

// Re-solve ONE pending axis of `mask` (bit 0 r, bit 1 theta, bit 2 phi; the lowest set bit is taken and
// cleared) for the cell (c0,c1,c2).  The coefficient sets of the two faces of that axis are built with
// selects, so lanes that crossed different kinds of faces stay converged in the solver.  A lane that just
// started a ray has all three bits set and needs three calls (three trips of the march loop) before it
// can step; a lane that merely crossed a face needs one.
__device__ __forceinline__ void ray_update(const Ctx& X, Ray& R, int& mask, int c0, int c1, int c2,
                                           double n0, double n1, double n2) {
    const DevTables& T = X.A.T;
    const double* sm = X.sm;
    const SmLayout& lay = X.lay;
    const int* tplane = reinterpret_cast<const int*>(sm + lay.o_tp);
    const double a = 1.0 / T.ox, b = 1.0 / T.oy;
    {
        const int ax = __ffs(mask) - 1;
        mask &= mask - 1;
        // the two faces of this axis: lower index kl, upper index ku (-1: no such face)
        int kl, ku;
        if (ax == 0) { kl = c0; ku = c0 + 1; }
        else if (ax == 1) { kl = (c1 != 0) ? c1 : -1; ku = (c1 + 1 != T.nt) ? c1 + 1 : -1; }
        else { kl = c2; ku = (c2 + 1 == T.np) ? 0 : c2 + 1; }
        double tl = RAY_NONE, tu = RAY_NONE;
#pragma unroll 1
        for (int side = 0; side < 2; ++side) {
            const int k = side ? ku : kl;
            double qa, hb, qc;
            int hs = 0;
            if (ax == 0) {
                const double r = sm[k < 0 ? 0 : k];
                qa = R.A1 + R.A2; hb = R.B1 + R.B2; qc = (R.C1 + R.C2) - r * r;
            } else if (ax == 1) {
                const int kk = k < 0 ? 0 : k;
                const double tn = sm[lay.o_tt + kk], tf = sm[lay.o_tf + kk];
                const double T2 = tn * tn;
                if (tplane[kk] == 1) {
                    qa = R.A1 - R.A2 * T2; hb = R.B1 - R.B2 * T2; qc = R.C1 - R.C2 * T2;
                    hs = (tf < PI / 2.0) ? 1 : ((tf > PI / 2.0) ? -1 : 0);
                } else { qa = 0.0; hb = 0.5 * n2; qc = R.z0; }
            } else {
                const double ps = sm[lay.o_ps + k], pc = sm[lay.o_pc + k];
                qa = 0.0; hb = 0.5 * (b * n1 * pc - a * n0 * ps); qc = -(a * R.x0 * ps - b * R.y0 * pc);
            }
            // a ray that started on this very face ignores its roots below the same-face threshold
            const double lim = (R.sf0 == ax + 1 && R.sf1 == k) ? fmax(R.t, 1.e-3) : R.t;
            const double tk = (k >= 0) ? quadric_next(qa, hb, qc, hs, lim, R.z0, n2) : RAY_NONE;
            if (side) tu = tk; else tl = tk;
        }
        const bool up = !(tl <= tu);
        const double tn_ = up ? tu : tl;
        const int fk = up ? ku : kl;
        if (ax == 0) { R.tr = tn_; R.fr = fk; }
        else if (ax == 1) { R.tt = tn_; R.ft = fk; }
        else { R.tp = tn_; R.fp = fk; }
        R.dirs = up ? (R.dirs | (1 << ax)) : (R.dirs & ~(1 << ax));
    }
}

// start a ray at the walker position of P along (n0,n1,n2)
__device__ __forceinline__ void ray_setup(const Ctx& X, const Photon& P, Ray& R, double n0, double n1, double n2) {
    const DevTables& T = X.A.T;
    const double a = 1.0 / T.ox, b = 1.0 / T.oy, c = 1.0 / T.oz;
    R.x0 = P.wx; R.y0 = P.wy; R.z0 = P.wz;
    R.t = 0.0;
    R.A1 = a * a * n0 * n0 + b * b * n1 * n1; R.A2 = c * c * n2 * n2;
    R.B1 = a * a * P.wx * n0 + b * b * P.wy * n1; R.B2 = c * c * P.wz * n2;
    R.C1 = a * a * P.wx * P.wx + b * b * P.wy * P.wy; R.C2 = c * c * P.wz * P.wz;
    R.sf0 = P.wf0; R.sf1 = P.wf1;
    R.dirs = 0;
    R.tr = R.tt = R.tp = RAY_NONE; R.fr = R.ft = R.fp = 0;
}

// axes a fresh ray has to solve: grids without theta / phi faces skip those axes
__device__ __forceinline__ int ray_axes(const DevTables& T) { return 1 | (T.nt > 1 ? 2 : 0) | (T.np > 1 ? 4 : 0); }

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
    const bool up = (R.dirs >> axis) & 1;
    if (axis == 0) { o.nf1 = R.fr; o.co0 = P.wc0 + (up ? 1 : -1); o.exit = (R.fr == T.nr); }
    else if (axis == 1) { o.nf1 = R.ft; o.co1 = P.wc1 + (up ? 1 : -1); }
    else { o.nf1 = R.fp; o.co2 = up ? ((P.wc2 + 1 == T.np) ? 0 : P.wc2 + 1) : ((P.wc2 == 0) ? T.np - 1 : P.wc2 - 1); }
}
