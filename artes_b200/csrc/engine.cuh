// engine.cuh -- photon state, the five event bodies, and the two engines that schedule them.
// Included from transport.cuh (inside namespace artes::ARTES_NS).
//
// Event bodies (each a restatement of one part of radiative_transfer, src/ARTES.f90:546-955):
//   ev_emit      A  emit_photon :1008-1268 (+ thermal weight and peel_thermal start :599-621)
//   ev_cross     B  one cell_face step of whichever walk the photon is in: tau pre-pass :633-656, first
//                   optical depth :660-685, the two transport walks :691-778 / :850-941 (incl. surface
//                   hit and Lambert reflection), the peel walks :4542-4569 / :4651-4673 / :4739-4761
//   ev_survive   D  survival test and albedo weight before a scattering :791-815
//   ev_peel_done C  detector deposit of peel_photon :4763-4986, peel_surface :4675-4704, peel_thermal :4571-4596
//   ev_scatter   E  scatter_photon :1434-1532 + polarization_rotation :1663-1932 + next tau :845-846
//
// Engines:
//   transport_kernel   "persistent" engine: one lane owns one photon until it dies; B runs converged for
//                      all walking lanes, C/E/A are ballot-deferred until enough lanes wait.
//   wf_* kernels       "wavefront" engine: photons live in an HBM pool (SoA); a persistent march kernel
//                      pulls photons from a queue and runs B/D until the next heavy event, then hands the
//                      photon to the queue of that event; converged event / emit kernels process those
//                      queues and feed the march queue of the next pass.

struct Photon {
    Rng rng;
    int ph, pk;
    bool peel_exit;
    bool at_walker;             // a retiring photon reports the walker position (it was in a transport walk)
    double px, py, pz;          // photon position ("home" while a probe walk runs)
    double dx, dy, dz;          // photon direction
    double S[4];                // Stokes vector
    int c0, c1, c2, f0, f1;     // cell and current face of the photon
    double tau, tau_run;
    double wx, wy, wz;          // walker: the point cell_face is evaluated at
    int wc0, wc1, wc2, wf0, wf1;
    double tacc;                // optical depth of the running probe walk (pre-pass / peel)
    int t_len, t_nsc;           // trace
    unsigned long long t_hash;
};

struct Counters {
    unsigned long long n_cf;
    unsigned n_emit, n_sc, n_peel, n_surf, n_err, n_draw;
    __device__ __forceinline__ void zero() { n_cf = 0; n_emit = n_sc = n_peel = n_surf = n_err = n_draw = 0; }
    __device__ __forceinline__ void flush(unsigned long long* stats) {
        const int lane = threadIdx.x & 31;
        unsigned long long v[7] = {n_emit, n_cf, n_sc, n_peel, n_surf, n_draw, n_err};
#pragma unroll
        for (int k = 0; k < 7; ++k) {
            unsigned long long x = v[k];
            for (int off = 16; off > 0; off >>= 1) x += __shfl_xor_sync(FULL, x, off);
            if (lane == 0 && x) atomicAdd(stats + k, x);
        }
    }
};

struct Ctx {
    const double* sm;
    SmLayout lay;
    const KernelArgs& A;
    __device__ __forceinline__ Ctx(const double* s, const KernelArgs& a) : sm(s), lay(a.T.nr, a.T.nt, a.T.np), A(a) {}
};

#if !ARTES_FAITHFUL
#include "ray.cuh"
#endif

__device__ __forceinline__ void stage_tables(double* sm, const DevTables& T) {
    const SmLayout lay(T.nr, T.nt, T.np);
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

__device__ __forceinline__ void err_count(const KernelArgs& A, int code) { atomicAdd(A.O.err + code, 1ull); }

template <bool TRACE>
__device__ __forceinline__ void record(const KernelArgs& A, Photon& P, int a, int b, int c, int d, int e) {
    if (TRACE) {
        if (A.R.seq_head && P.t_len < A.R.max_rec) {
            int* p = A.R.seq_head + ((size_t)(P.rng.id - A.L.id_base) * A.R.max_rec + P.t_len) * 5;
            p[0] = a; p[1] = b; p[2] = c; p[3] = d; p[4] = e;
        }
        tuple_hash(P.t_hash, a); tuple_hash(P.t_hash, b); tuple_hash(P.t_hash, c); tuple_hash(P.t_hash, d); tuple_hash(P.t_hash, e);
        ++P.t_len;
    }
}

// photon finished: publish the trace record, mark the lane / slot free
template <bool TRACE>
__device__ __forceinline__ void retire(const KernelArgs& A, Photon& P, Counters& C) {
    if (TRACE) {
        size_t k = (size_t)(P.rng.id - A.L.id_base);
        A.R.seq_len[k] = P.t_len; A.R.seq_hash[k] = P.t_hash;
        if (A.R.fstate) {
            double* f = A.R.fstate + k * 8;
            const bool live = (P.ph == PH_WALK) || P.at_walker;
            f[0] = live ? P.wx : P.px; f[1] = live ? P.wy : P.py; f[2] = live ? P.wz : P.pz;
            f[3] = P.S[0]; f[4] = P.S[1]; f[5] = P.S[2]; f[6] = P.S[3]; f[7] = (double)P.t_nsc;
        }
    }
    C.n_draw += P.rng.nd;
    P.at_walker = false;
    P.ph = PH_NEW;
}

// detector deposit :4947-4972 / :4575-4585 / :4683-4693
template <bool TRACE>
__device__ __forceinline__ void deposit(const KernelArgs& A, Photon& P, double W0, double W1, double W2, double W3, bool all4) {
    const LaunchArgs& L = A.L;
    double x_im = P.py * L.cos_dp - P.px * L.sin_dp;
    double y_im = P.pz * L.sin_dt - P.py * L.cos_dt * L.sin_dp - P.px * L.cos_dt * L.cos_dp;
    int ix = (int)(L.nx * (x_im + L.x_max) / (2.0 * L.x_max)) + 1;
    int iy = (int)(L.ny * (y_im + L.y_max) / (2.0 * L.y_max)) + 1;
    if (ix < 1 || ix > L.nx || iy < 1 || iy > L.ny) { err_count(A, 60); return; }
    record<TRACE>(A, P, 100, ix, iy, 0, 0);
    const size_t npx = (size_t)L.nx * L.ny;
    double* d = A.O.det + (size_t)(ix - 1) + (size_t)L.nx * (iy - 1);
    atomicAdd(d, W0); atomicAdd(d + 4 * npx, W0 * W0); atomicAdd(d + 8 * npx, 1.0);
    if (all4) {
        atomicAdd(d + npx, W1); atomicAdd(d + 2 * npx, W2); atomicAdd(d + 3 * npx, W3);
        atomicAdd(d + 5 * npx, W1 * W1); atomicAdd(d + 6 * npx, W2 * W2); atomicAdd(d + 7 * npx, W3 * W3);
        atomicAdd(d + 9 * npx, 1.0);
    }
}

__device__ __forceinline__ void start_probe(Photon& P, int r_shift) {
    P.wx = P.px; P.wy = P.py; P.wz = P.pz;
    P.wc0 = P.c0 + r_shift; P.wc1 = P.c1; P.wc2 = P.c2; P.wf0 = P.f0; P.wf1 = P.f1;
    P.tacc = 0.0;
}

// cos of the angle between the (oblate) surface normal at (x,y,z) and the detector :4609-4634
__device__ __noinline__ double surface_cos_angle(const KernelArgs& A, double x, double y, double z) {
    const DevTables& T = A.T;
    const LaunchArgs& L = A.L;
    double s0 = x / (T.ox * T.ox), s1 = y / (T.oy * T.oy), s2 = z / (T.oz * T.oz);
    double nrm = sqrt(s0 * s0 + s1 * s1 + s2 * s2);
    s0 = s0 / nrm; s1 = s1 / nrm; s2 = s2 / nrm;
    double nth = acos(s2 / sqrt(s0 * s0 + s1 * s1 + s2 * s2));
    double nph = atan2(s1, s0);
    if (nph < 0.0) nph = nph + 2.0 * PI;
    return sin(L.det_sph_theta) * cos(L.det_sph_phi) * sin(nth) * cos(nph) +
           sin(L.det_sph_theta) * sin(L.det_sph_phi) * sin(nth) * sin(nph) + cos(L.det_sph_theta) * cos(nth);
}

// add_flow_global :4992-5014
__device__ __noinline__ void add_flow_global(const KernelArgs& A, double x, double y, double z, double dx, double dy, double dz,
                                                double e, double dist, int cellidx) {
    double th = acos(z / sqrt(x * x + y * y + z * z)), phh = atan2(y, x);
    double* f = A.O.flow3 + (size_t)3 * cellidx;
    atomicAdd(f, (sin(th) * cos(phh) * dx + sin(th) * sin(phh) * dy + cos(th) * dz) * dist * e);
    atomicAdd(f + 1, (cos(th) * cos(phh) * dx + cos(th) * sin(phh) * dy - sin(th) * dz) * dist * e);
    atomicAdd(f + 2, (-sin(phh) * dx + cos(phh) * dy) * dist * e);
}

// ================= A. emission: photon k of this launch (emit_photon :1008-1268) =================
template <bool TRACE, bool GEN>
__device__ __forceinline__ void ev_emit(const Ctx& X, Photon& P, Counters& C, unsigned long long k) {
    const KernelArgs& A = X.A;
    const DevTables& T = A.T;
    const LaunchArgs& L = A.L;
    const double* sm = X.sm;
    const SmLayout& lay = X.lay;
    P.rng.id = L.id_base + k; P.rng.nd = 0; P.rng.exhausted = false;
    P.t_len = 0; P.t_nsc = 0; P.t_hash = 1469598103934665603ull;
    ++C.n_emit;
    P.S[0] = 1.0; P.S[1] = 0.0; P.S[2] = 0.0; P.S[3] = 0.0;
    P.tau = 0.0; P.tau_run = 0.0; P.peel_exit = false; P.at_walker = false; P.pk = PK_SCATTER;
    int e = 0;
    double bias_weight = 1.0;
    if (!GEN || L.photon_source == 1) {
        P.f0 = 1; P.f1 = T.nr;
        double xi, r_disk;
        if (L.limb_emission) {
            for (;;) { xi = rng_next<TRACE>(P.rng, A); r_disk = sqrt(xi); if (r_disk > 0.9 || P.rng.exhausted) break; }
        } else { xi = rng_next<TRACE>(P.rng, A); r_disk = sqrt(xi); }
        xi = rng_next<TRACE>(P.rng, A);
        const double phi_disk = 2.0 * PI * xi;
        const double R = sm[T.nr];
        double sphi, cphi;
#if ARTES_FAITHFUL
        sphi = sin(phi_disk); cphi = cos(phi_disk);
#else
        sincos(phi_disk, &sphi, &cphi);
#endif
        const double d1 = R * r_disk * sphi;
        const double d2 = R * r_disk * cphi;
        P.dx = -1.0; P.dy = 0.0; P.dz = 0.0;
        P.px = sqrt(R * R - d1 * d1 - d2 * d2); P.py = d1; P.pz = d2;
        if (L.stellar_direction) {  // :1080-1111
            double tx = P.px * L.rot_y_cos + P.py * 0.0 + P.pz * L.rot_y_sin;
            double ty = P.px * 0.0 + P.py * 1.0 + P.pz * 0.0;
            double tz = P.px * (-L.rot_y_sin) + P.py * 0.0 + P.pz * L.rot_y_cos;
            P.px = tx * L.rot_z_cos + ty * (-L.rot_z_sin) + tz * 0.0;
            P.py = tx * L.rot_z_sin + ty * L.rot_z_cos + tz * 0.0;
            P.pz = tx * 0.0 + ty * 0.0 + tz * 1.0;
            P.dx = L.star_dir[0]; P.dy = L.star_dir[1]; P.dz = L.star_dir[2];
        }
        initial_cell(sm, lay, T, P.px, P.py, P.pz, P.c0, P.c1, P.c2);
    } else {  // thermal :1117-1266
        P.f0 = 0; P.f1 = 0;
        double xi = rng_next<TRACE>(P.rng, A);
        const int ncdf = (T.nr - T.cell_depth) * T.nt * T.np;
        const double samp = xi * __ldg(T.emis_cdf + ncdf - 1);
        int lo = -1, hi = ncdf - 1;  // first p with cdf[p] >= samp (== the linear scan of :1132-1155)
        while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (__ldg(T.emis_cdf + mid) >= samp) hi = mid; else lo = mid; }
        P.c2 = hi % T.np; P.c1 = (hi / T.np) % T.nt; P.c0 = T.cell_depth + hi / (T.np * T.nt);
        xi = rng_next<TRACE>(P.rng, A);
        double rs = xi * (sm[P.c0 + 1] - sm[P.c0]); rs = sm[P.c0] + rs;
        xi = rng_next<TRACE>(P.rng, A);
        double ct = xi * (sm[lay.o_tc + P.c1 + 1] - sm[lay.o_tc + P.c1]); ct = sm[lay.o_tc + P.c1] + ct;
        double st = sqrt(1.0 - ct * ct);
        xi = rng_next<TRACE>(P.rng, A);
        double phs;
        if (T.np == 1) phs = 2.0 * PI * xi;
        else if (P.c2 < T.np - 1) { phs = xi * (sm[lay.o_pf + P.c2 + 1] - sm[lay.o_pf + P.c2]); phs = sm[lay.o_pf + P.c2] + phs; }
        else { phs = xi * (2.0 * PI - sm[lay.o_pf + P.c2]); phs = sm[lay.o_pf + P.c2] + phs; }
        double cp = cos(phs), sp = sqrt(1.0 - cp * cp);
        if (phs > PI) sp = -sp;
        P.px = rs * st * cp; P.py = rs * st * sp; P.pz = rs * ct;
        P.px = T.ox * P.px; P.py = T.oy * P.py; P.pz = T.oz * P.pz;
        if (L.photon_emission == 1) {
            xi = rng_next<TRACE>(P.rng, A);
            double al = 2.0 * xi - 1.0;
            xi = rng_next<TRACE>(P.rng, A);
            double be = 2.0 * PI * xi;
            double cb = cos(be), sb = sqrt(1.0 - cb * cb);
            if (be > PI) sb = -sb;
            P.dx = sqrt(1.0 - al * al) * cb; P.dy = sqrt(1.0 - al * al) * sb; P.dz = al;
        } else {
            xi = rng_next<TRACE>(P.rng, A);
            double yb = (1.0 + L.photon_bias) * tan(PI * xi / 2.0) / sqrt(1.0 - L.photon_bias * L.photon_bias);
            double ths = acos((1.0 - yb * yb) / (1.0 + yb * yb));
            xi = rng_next<TRACE>(P.rng, A);
            double be = 2.0 * PI * xi;
            double r0 = P.px / (T.ox * T.ox), r1 = P.py / (T.oy * T.oy), r2 = P.pz / (T.oz * T.oz);
            double nrm = sqrt(r0 * r0 + r1 * r1 + r2 * r2);
            r0 = r0 / nrm; r1 = r1 / nrm; r2 = r2 / nrm;
            e = direction_cosine(cos(PI - ths), be, r0, r1, r2, P.dx, P.dy, P.dz);
            bias_weight = (PI * sin(ths) * (1.0 + L.photon_bias * cos(ths))) / (2.0 * sqrt(1.0 - L.photon_bias * L.photon_bias));
        }
        if (e == 0 && fabs(P.dz) >= 1.0) err_count(A, 54);
    }
    if (e) { err_count(A, e); ++C.n_err; retire<TRACE>(A, P, C); }
    else if (GEN && L.photon_source == 2) {  // :599-621
        P.S[0] = P.S[0] * bias_weight / __ldg(T.cell_weight + P.c0 + T.nr * (P.c1 + T.nt * P.c2));
        atomicAdd(A.O.flux, P.S[0]);
        ++C.n_peel; P.pk = PK_THERMAL; start_probe(P, 0); P.ph = PH_PEEL;
    } else { start_probe(P, 0); P.ph = PH_PRE; }
}

// ================= D. interaction point reached: survival + start of the peel-off (:788-815) =================
template <bool TRACE>
__device__ __forceinline__ void ev_survive(const Ctx& X, Photon& P, Counters& C) {
    const KernelArgs& A = X.A;
    const DevTables& T = A.T;
    const LaunchArgs& L = A.L;
    bool alive = L.photon_scattering != 0;
    if (TRACE && P.rng.exhausted) alive = false;
    if (alive) {
        double xi = rng_next<TRACE>(P.rng, A);
        if (xi < L.fstop) alive = false;
    }
    if (alive) {
        const double alb = __ldg(T.albedo + P.c0 + T.nr * (P.c1 + T.nt * P.c2));
        if (alb < 1.0 && alb > 0.0) {
            double gamma = alb / (1.0 - L.fstop);
            P.S[0] = gamma * P.S[0]; P.S[1] = gamma * P.S[1]; P.S[2] = gamma * P.S[2]; P.S[3] = gamma * P.S[3];
        }
        if (P.S[0] <= L.photon_minimum) alive = false;
    }
    if (!alive) retire<TRACE>(A, P, C);
    else { ++C.n_peel; P.pk = PK_SCATTER; start_probe(P, 0); P.ph = PH_PEEL; }
}

// ================= B. one cell crossing of the walk the photon is in =================
// On return P.ph tells what comes next: PH_PRE / PH_WALK / PH_PEEL (keep walking), PH_SCAT (interaction
// reached -> ev_survive), PH_PEELDONE (-> ev_peel_done), PH_LAMBERT (surface reflection event), PH_NEW (retired).
template <bool TRACE, bool GEN>
__device__ __forceinline__ void apply_crossing(const Ctx& X, Photon& P, Counters& C, const CellFace& o, double n0, double n1, double n2);

template <bool TRACE, bool GEN>
__device__ __forceinline__ void ev_cross(const Ctx& X, Photon& P, Counters& C) {
    const LaunchArgs& L = X.A.L;
    const bool peel = (P.ph == PH_PEEL);
    const double n0 = peel ? L.det[0] : P.dx, n1 = peel ? L.det[1] : P.dy, n2 = peel ? L.det[2] : P.dz;
    CellFace o;
    cell_face(X.sm, X.lay, X.A.T, P.wx, P.wy, P.wz, n0, n1, n2, P.wf0, P.wf1, P.wc0, P.wc1, P.wc2, o);
    apply_crossing<TRACE, GEN>(X, P, C, o, n0, n1, n2);
}

// What one crossing does to the photon, given the geometry result `o` along direction (n0,n1,n2).
// Touches only the "hot" walker state (w, wc, wf, tau, tau_run, tacc, S[0] for the flow counters); whatever
// needs the random stream or the full Stokes vector is left to a follow-up handler chosen through P.ph:
//   PH_PREDONE -> ev_pre_done, PH_SCAT -> ev_survive, PH_SURFHIT -> ev_surface_hit, PH_RETIRE -> ev_retire.
template <bool TRACE, bool GEN>
__device__ __forceinline__ void apply_crossing(const Ctx& X, Photon& P, Counters& C, const CellFace& o, double n0, double n1, double n2) {
    const KernelArgs& A = X.A;
    const DevTables& T = A.T;
    const LaunchArgs& L = A.L;
    ++C.n_cf;
    record<TRACE>(A, P, o.nf0, o.nf1, o.co0, o.co1, o.co2);
    const int wci = P.wc0 + T.nr * (P.wc1 + T.nt * P.wc2);
    if (o.err) {
        err_count(A, o.err);
        P.peel_exit = false;                      // not a grid exit
        P.at_walker = (P.ph == PH_WALK);
        if (P.ph == PH_PRE) { err_count(A, 2); ++C.n_err; P.ph = PH_RETIRE; }
        else if (P.ph == PH_WALK) { err_count(A, 3); ++C.n_err; P.ph = PH_RETIRE; }
        else if (P.pk == PK_SCATTER) { err_count(A, 43); ++C.n_err; P.ph = PH_RETIRE; }
        else if (P.pk == PK_THERMAL) { err_count(A, 46); err_count(A, 47); ++C.n_err; P.ph = PH_RETIRE; }
        else { err_count(A, 42); P.ph = PH_PEELDONE; }
    } else if (P.ph == PH_WALK) {
        const double kap = __ldg(T.kext + wci);
        const double tau_cell = o.dist * kap;
        if (P.tau_run + tau_cell > P.tau) {  // :705-720 / :862-879 interaction inside this cell
            const double s = (P.tau - P.tau_run) / kap;
            P.px = P.wx + s * n0; P.py = P.wy + s * n1; P.pz = P.wz + s * n2;
            P.c0 = P.wc0; P.c1 = P.wc1; P.c2 = P.wc2; P.f0 = 0; P.f1 = 0;
            if (GEN && L.flow_global) add_flow_global(A, P.px, P.py, P.pz, n0, n1, n2, P.S[0], s, wci);
            P.ph = PH_SCAT;
        } else {
            P.wx = P.wx + o.dist * n0; P.wy = P.wy + o.dist * n1; P.wz = P.wz + o.dist * n2;
            if (GEN && L.flow_global) add_flow_global(A, P.wx, P.wy, P.wz, n0, n1, n2, P.S[0], o.dist, wci);
            if (GEN && L.flow_theta) {  // :730-744
                double* f = A.O.flow4 + (size_t)4 * wci;
                if (o.nf0 == 1) { if (o.co0 > P.wc0) atomicAdd(f, P.S[0]); else if (o.co0 < P.wc0) atomicAdd(f + 1, P.S[0]); }
                else if (o.nf0 == 2) { if (o.co1 > P.wc1) atomicAdd(f + 2, P.S[0]); else if (o.co1 < P.wc1) atomicAdd(f + 3, P.S[0]); }
            }
            P.wf0 = o.nf0; P.wf1 = o.nf1; P.wc0 = o.co0; P.wc1 = o.co1; P.wc2 = o.co2;
            P.tau_run = P.tau_run + tau_cell;      // :776 (after a grid exit / absorption it is never read again)
            if (o.exit) { P.peel_exit = true; P.at_walker = true; P.ph = PH_RETIRE; }
            else if (o.nf0 == 1 && o.nf1 == T.cell_depth) { ++C.n_surf; P.ph = PH_SURFHIT; }   // surface :755-774
        }
    } else {
        // probe walks: tau pre-pass :633-656 and the three peel walks
        P.tacc = P.tacc + o.dist * __ldg(T.kext + wci);
        P.wx = P.wx + o.dist * n0; P.wy = P.wy + o.dist * n1; P.wz = P.wz + o.dist * n2;
        const bool hit_surface = (o.nf0 == 1 && o.nf1 == T.cell_depth);
        if (o.exit || hit_surface) {
            P.peel_exit = o.exit;
            P.ph = (P.ph == PH_PRE) ? PH_PREDONE : PH_PEELDONE;
        } else { P.wf0 = o.nf0; P.wf1 = o.nf1; P.wc0 = o.co0; P.wc1 = o.co1; P.wc2 = o.co2; }
    }
}

// first optical depth :660-685 (the pre-pass ended on the grid boundary, peel_exit, or on the surface)
template <bool TRACE>
__device__ __forceinline__ void ev_pre_done(const Ctx& X, Photon& P, Counters& C) {
    const KernelArgs& A = X.A;
    const bool hit_surface = !P.peel_exit;
    if (P.tacc < 1.e-6 && !hit_surface) { retire<TRACE>(A, P, C); return; }
    // one draw in every branch of :666-685; the forced first interaction rescales the Stokes vector
    const double xi = rng_next<TRACE>(P.rng, A);
    double arg = 1.0 - xi;
    if (!(P.tacc < 1.e-6) && P.tacc < 50.0) {
        const double f = 1.0 - exp(-P.tacc);
        arg = 1.0 - xi * f;
        P.S[0] = P.S[0] * f; P.S[1] = P.S[1] * f; P.S[2] = P.S[2] * f; P.S[3] = P.S[3] * f;
    }
    P.tau = -log(arg);
    P.tau_run = 0.0; start_probe(P, 0); P.ph = PH_WALK;
}

// the walk reached the surface :755-764: absorbed, or on to the Lambert reflection event
template <bool TRACE, bool GEN>
__device__ __forceinline__ void ev_surface_hit(const Ctx& X, Photon& P, Counters& C) {
    const KernelArgs& A = X.A;
    double xi = rng_next<TRACE>(P.rng, A);
    if (!GEN || xi > A.L.surface_albedo) { P.at_walker = true; retire<TRACE>(A, P, C); }
    else P.ph = PH_LAMBERT;
}

// a photon left the grid (peel_exit) or was dropped by an error path while walking
template <bool TRACE, bool GEN>
__device__ __forceinline__ void ev_retire(const Ctx& X, Photon& P, Counters& C) {
    const KernelArgs& A = X.A;
    if (GEN && P.peel_exit && A.L.photon_source == 2) atomicAdd(A.O.flux + 1, P.S[0]);  // :780 / :953
    retire<TRACE>(A, P, C);
}

// every cheap follow-up of a crossing, for the engines that keep the whole photon in registers
template <bool TRACE, bool GEN>
__device__ __forceinline__ void cheap_handlers(const Ctx& X, Photon& P, Counters& C) {
    if (P.ph == PH_PREDONE) ev_pre_done<TRACE>(X, P, C);
    else if (P.ph == PH_SURFHIT) ev_surface_hit<TRACE, GEN>(X, P, C);
    else if (P.ph == PH_RETIRE) ev_retire<TRACE, GEN>(X, P, C);
    if (P.ph == PH_SCAT) ev_survive<TRACE>(X, P, C);
}

// ================= surface reflection: lambertian :1369-1402, then the start of peel_surface :4600-4650 ==========
template <bool TRACE>
__device__ __forceinline__ void ev_lambert(const Ctx& X, Photon& P, Counters& C) {
    const KernelArgs& A = X.A;
    const DevTables& T = A.T;
    double s0 = P.wx / (T.ox * T.ox), s1 = P.wy / (T.oy * T.oy), s2 = P.wz / (T.oz * T.oz);
    double nrm = sqrt(s0 * s0 + s1 * s1 + s2 * s2);
    s0 = s0 / nrm; s1 = s1 / nrm; s2 = s2 / nrm;
    double xi = rng_next<TRACE>(P.rng, A);
    double al = sqrt(xi);
    xi = rng_next<TRACE>(P.rng, A);
    double be = 2.0 * PI * xi;
    double e0, e1, e2;
    int e = direction_cosine(al, be, s0, s1, s2, e0, e1, e2);
    if (e) { err_count(A, e); ++C.n_err; retire<TRACE>(A, P, C); return; }
    P.dx = e0; P.dy = e1; P.dz = e2;
    P.px = P.wx; P.py = P.wy; P.pz = P.wz; P.c0 = P.wc0; P.c1 = P.wc1; P.c2 = P.wc2; P.f0 = P.wf0; P.f1 = P.wf1;
    const double cos_angle = surface_cos_angle(A, P.px, P.py, P.pz);
    // stokes_new :1397-1400 -- the surface peel only uses I, so Q,U,V can be cleared before it runs
    P.S[1] = 0.0; P.S[2] = 0.0; P.S[3] = 0.0;
    if (cos_angle > 0.0) { ++C.n_peel; P.pk = PK_SURFACE; start_probe(P, 1); P.ph = PH_PEEL; }
    else { P.c0 = P.c0 + 1; P.wc0 = P.c0; P.ph = PH_WALK; }
}

// ================= C. a peel walk ended: weight + deposit =================
template <bool TRACE, bool GEN>
__device__ __forceinline__ void ev_peel_done(const Ctx& X, Photon& P, Counters& C) {
    const KernelArgs& A = X.A;
    const DevTables& T = A.T;
    const LaunchArgs& L = A.L;
    const bool ok = P.peel_exit && P.tacc < 50.0;
    if (GEN && P.pk == PK_THERMAL) {  // :4571-4596
        if (ok) {
            double w = exp(-P.tacc) / (4.0 * PI);
            double W0 = w * P.S[0];
            if (W0 > 0.0 && W0 < 1.e100) deposit<TRACE>(A, P, W0, 0, 0, 0, false); else err_count(A, 51);
        }
        start_probe(P, 0); P.ph = PH_PRE;
    } else if (GEN && P.pk == PK_SURFACE) {  // :4675-4704
        if (ok) {
            const double cos_angle = surface_cos_angle(A, P.px, P.py, P.pz);
            double w = exp(-P.tacc) * cos_angle / PI;
            double W0 = w * P.S[0];
            if (W0 > 0.0 && W0 < 1.e100) deposit<TRACE>(A, P, W0, 0, 0, 0, false); else err_count(A, 52);
        }
        P.c0 = P.c0 + 1;  // :770
        start_probe(P, 0); P.ph = PH_WALK;
    } else {  // peel_photon :4763-4986
        if (ok) {
            const double w = exp(-P.tacc);
            double mu = P.dx * L.det[0] + P.dy * L.det[1] + P.dz * L.det[2];
            if (mu >= 1.0) mu = 1.0 - 1.e-10;
            else if (mu <= -1.0) mu = -1.0 + 1.e-10;
            double F[16];
            const int ci = P.c0 + T.nr * (P.c1 + T.nt * P.c2);
#if ARTES_FAITHFUL
            matrix_at(T, ci, acos(mu), F);
            double phi_old = atan2(P.dy, P.dx);
            if (phi_old < 0.0) phi_old = phi_old + 2.0 * PI;
            if (phi_old > 2.0 * PI) phi_old = phi_old - 2.0 * PI;
            const double phi_new = L.det_atan2;
            if (!(fabs(P.dz) < 1.0)) err_count(A, 45);
            else {
                double nc = (L.det[2] - P.dz * mu) / (sqrt(1.0 - mu * mu) * sqrt(1.0 - P.dz * P.dz));
                double phs = 0.0;
                bool good = true;
                if (fabs(nc) < 1.0) phs = acos(nc);
                else if (nc >= 1.0) phs = 0.0 + 1.e-10;
                else if (nc <= -1.0) phs = PI - 1.e-10;
                else { good = false; err_count(A, 44); }
                if (good) {
                    if (phi_old - phi_new >= 0.0 && phi_old - phi_new < PI) phs = 2.0 * PI - phs;
                    if (2.0 * PI + phi_old - phi_new >= 0.0 && 2.0 * PI + phi_old - phi_new < PI) phs = 2.0 * PI - phs;
                    if (phs < 0.0) phs = phs + 2.0 * PI;
                    double so[4];
                    int soft = 0;
                    int e = polarization_rotation(mu, phs, P.S, F, P.dz, L.det[2], so, true, soft);
                    if (soft) err_count(A, soft);
                    if (e) err_count(A, e);
                    else if (w * so[0] > 0.0 && w * so[0] < 1.e100) deposit<TRACE>(A, P, w * so[0], -(w * so[1]), w * so[2], w * so[3], true);
                    else err_count(A, 53);
                }
            }
#else
            matrix_at_deg(T, ci, acos(mu) * (180.0 / PI), F);
            if (!(fabs(P.dz) < 1.0)) err_count(A, 45);
            else {
                const double smu = sqrt(1.0 - mu * mu);
                double nc = (L.det[2] - P.dz * mu) / (smu * sqrt(1.0 - P.dz * P.dz));
                if (!(nc == nc)) err_count(A, 44);
                else {
                    nc = fmin(fmax(nc, -1.0), 1.0);
                    const double cr = P.dy * L.det[0] - P.dx * L.det[1];                 // sin(phi_old - phi_new) > 0 ?
                    const bool flip = (cr > 0.0) || (cr == 0.0 && P.dx * L.det[0] + P.dy * L.det[1] > 0.0);
                    const double c2a = 2.0 * nc * nc - 1.0;
                    double s2a = 2.0 * nc * sqrt(fmax(1.0 - nc * nc, 0.0));
                    if (flip) s2a = -s2a;
                    const double nc2 = (P.dz - L.det[2] * mu) / (smu * sqrt(1.0 - L.det[2] * L.det[2]));
                    double so[4];
                    int soft = 0;
                    int e = (fabs(L.det[2]) < 1.0) ? polrot_fast(c2a, s2a, flip, nc2, P.S, F, so, true, soft) : 16;
                    if (e) err_count(A, e);
                    else if (w * so[0] > 0.0 && w * so[0] < 1.e100) deposit<TRACE>(A, P, w * so[0], -(w * so[1]), w * so[2], w * so[3], true);
                    else err_count(A, 53);
                }
            }
#endif
        }
        P.ph = PH_SCAT2;
    }
}

// ================= E. scattering event (scatter_photon :1434-1532 + polarization_rotation + :845) ==========
template <bool TRACE>
__device__ __forceinline__ void ev_scatter(const Ctx& X, Photon& P, Counters& C) {
    const KernelArgs& A = X.A;
    const DevTables& T = A.T;
    ++C.n_sc; ++P.t_nsc;
    const int ci = P.c0 + T.nr * (P.c1 + T.nt * P.c2);
    double e0 = 0, e1 = 0, e2 = 0;
    int e;
#if ARTES_FAITHFUL
    double alpha, beta;
    e = sample_angles<TRACE>(X.sm, X.lay, A, P.rng, P.S, ci, alpha, beta);
    if (!e) e = direction_cosine(alpha, beta, P.dx, P.dy, P.dz, e0, e1, e2);
    if (!e && !(fabs(alpha) < 1.0)) e = 50;
    if (!e) {
        double F[16], Sn[4];
        matrix_at(T, ci, acos(alpha), F);
        int soft = 0;
        e = polarization_rotation(alpha, beta, P.S, F, P.dz, e2, Sn, false, soft);
        if (soft) err_count(A, soft);
        if (!e) { P.S[0] = Sn[0]; P.S[1] = Sn[1]; P.S[2] = Sn[2]; P.S[3] = Sn[3]; P.dx = e0; P.dy = e1; P.dz = e2; }
    }
#else
    FastAngles g;
    e = sample_angles_fast<TRACE>(A, P.rng, P.S, ci, g);
    if (!e) {
        // direction_cosine :1962-2052 without the acos / cos round trip
        const double cto = P.dz / sqrt(P.dx * P.dx + P.dy * P.dy + P.dz * P.dz);
        const double sto = sqrt(1.0 - cto * cto);
        const double ctn = cto * g.alpha + sto * g.sT * g.cb;
        const double stn = sqrt(1.0 - ctn * ctn);
        double nc = (g.alpha - ctn * cto) / (stn * sto);
        if (!(nc == nc)) e = 20;
        else {
            if (nc >= 1.0) nc = 1.0 - 1.e-10; else if (nc <= -1.0) nc = -1.0 + 1.e-10;
            const double sD = sqrt(1.0 - nc * nc) * (g.flip ? -1.0 : 1.0);
            const double rho = sqrt(P.dx * P.dx + P.dy * P.dy);
            const double cph = rho > 0.0 ? P.dx / rho : 1.0, sph = rho > 0.0 ? P.dy / rho : 0.0;
            e0 = stn * (cph * nc - sph * sD); e1 = stn * (sph * nc + cph * sD); e2 = ctn;
            if (!(fabs(e2) < 1.0)) e = 16;
        }
    }
    if (!e) {
        double F[16], Sn[4];
        matrix_at_deg(T, ci, g.deg, F);
        const double nc2 = (P.dz - e2 * g.alpha) / (g.sT * sqrt(1.0 - e2 * e2));
        int soft = 0;
        e = polrot_fast(g.cb * g.cb - g.sb * g.sb, 2.0 * g.sb * g.cb, g.flip, nc2, P.S, F, Sn, false, soft);
        if (soft) err_count(A, soft);
        if (!e) { P.S[0] = Sn[0]; P.S[1] = Sn[1]; P.S[2] = Sn[2]; P.S[3] = Sn[3]; P.dx = e0; P.dy = e1; P.dz = e2; }
    }
#endif
    if (e) { err_count(A, e); ++C.n_err; retire<TRACE>(A, P, C); }
    else {
        double xi = rng_next<TRACE>(P.rng, A);  // :845
        P.tau = -log(1.0 - xi);
        P.tau_run = 0.0;
        start_probe(P, 0); P.ph = PH_WALK;
    }
}

// =====================================================================================================
// Engine 1: persistent lanes (one lane keeps one photon), events ballot-deferred
// =====================================================================================================
template <bool TRACE, bool GEN>
__global__ void __launch_bounds__(128, 4) transport_kernel(const __grid_constant__ KernelArgs A) {
    extern __shared__ double sm[];
    stage_tables(sm, A.T);
    const Ctx X(sm, A);
    const LaunchArgs& L = A.L;
    const int lane = threadIdx.x & 31;
    Photon P;
    P.ph = PH_NEW; P.rng.nd = 0; P.rng.id = 0; P.rng.exhausted = false; P.at_walker = false;
    Counters C; C.zero();
#if !ARTES_FAITHFUL && ARTES_PERSISTENT_RAY
    Ray R;
    int upd = 0, ray_ph = -1;      // ray_ph: the walk the current ray belongs to (-1: none)
    double rn0 = 0, rn1 = 0, rn2 = 0;
#endif

    for (;;) {
        // A. refill: new photons are emitted when enough lanes are free (or nobody walks any more)
        const unsigned need0 = __ballot_sync(FULL, P.ph == PH_NEW);
        const unsigned walk0 = __ballot_sync(FULL, P.ph == PH_PRE || P.ph == PH_WALK || P.ph == PH_PEEL);
        const unsigned need = (__popc(need0) >= L.defer_refill || walk0 == 0u) ? need0 : 0u;
        if (need) {
            const int leader = __ffs(need) - 1;
            unsigned long long base = 0;
            if (lane == leader) base = atomicAdd(A.O.counter, (unsigned long long)__popc(need));
            base = __shfl_sync(FULL, base, leader);
            if (P.ph == PH_NEW) {
                const unsigned long long k = base + (unsigned long long)__popc(need & ((1u << lane) - 1u));
                if (k >= L.n_photons) P.ph = PH_IDLE;
                else ev_emit<TRACE, GEN>(X, P, C, k);
            }
        }
        // B. one cell crossing for every walking lane
#if ARTES_FAITHFUL || !ARTES_PERSISTENT_RAY
        if (P.ph == PH_PRE || P.ph == PH_WALK || P.ph == PH_PEEL) { ev_cross<TRACE, GEN>(X, P, C); cheap_handlers<TRACE, GEN>(X, P, C); }
#else
        {   // fast mode: incremental ray marching (ray.cuh); a walk that just started first solves its three axes
            const bool walking = (P.ph == PH_PRE || P.ph == PH_WALK || P.ph == PH_PEEL);
            if (walking && ray_ph != P.ph) {
                const bool peel = (P.ph == PH_PEEL);
                rn0 = peel ? L.det[0] : P.dx; rn1 = peel ? L.det[1] : P.dy; rn2 = peel ? L.det[2] : P.dz;
                ray_setup(X, P, R, rn0, rn1, rn2); upd = ray_axes(A.T); ray_ph = P.ph;
            }
            if (walking && upd) ray_update(X, R, upd, P.wc0, P.wc1, P.wc2, rn0, rn1, rn2);
            if (walking && upd == 0) {
                CellFace o;
                int axis;
                ray_next(A.T, P, R, o, axis);
                const int ph0 = P.ph;
                apply_crossing<TRACE, GEN>(X, P, C, o, rn0, rn1, rn2);
                if (P.ph == ph0) { R.t = (axis == 0) ? R.tr : ((axis == 1) ? R.tt : R.tp); upd = 1 << axis; }
                else ray_ph = -1;
                cheap_handlers<TRACE, GEN>(X, P, C);
            }
        }
#endif
        if (GEN && P.ph == PH_LAMBERT) { ev_lambert<TRACE>(X, P, C); }
        // C, E. heavy events run once enough lanes of the warp wait for them
        const unsigned m_evt = __ballot_sync(FULL, P.ph == PH_PEELDONE || P.ph == PH_SCAT2);
        const unsigned m_walk = __ballot_sync(FULL, P.ph == PH_PRE || P.ph == PH_WALK || P.ph == PH_PEEL);
        const bool run_events = m_evt && (__popc(m_evt) >= L.defer_events || m_walk == 0u);
        if (run_events && P.ph == PH_PEELDONE) ev_peel_done<TRACE, GEN>(X, P, C);
        if (run_events && P.ph == PH_SCAT2) ev_scatter<TRACE>(X, P, C);
        if (__all_sync(FULL, P.ph == PH_IDLE)) break;
    }
    C.flush(A.O.stats);
}


__device__ __forceinline__ unsigned long long pack_cf(int c0, int c1, int c2, int f0, int f1) {
    return (unsigned long long)(unsigned)((c0 + 1) & 0xffff) | ((unsigned long long)(unsigned)(c1 & 0xffff) << 16) |
           ((unsigned long long)(unsigned)(c2 & 0xffff) << 32) | ((unsigned long long)(unsigned)(f0 & 3) << 48) |
           ((unsigned long long)(unsigned)(f1 & 0x3fff) << 50);
}
__device__ __forceinline__ void unpack_cf(unsigned long long v, int& c0, int& c1, int& c2, int& f0, int& f1) {
    c0 = (int)(v & 0xffff) - 1; c1 = (int)((v >> 16) & 0xffff); c2 = (int)((v >> 32) & 0xffff);
    f0 = (int)((v >> 48) & 3); f1 = (int)((v >> 50) & 0x3fff);
}

// =====================================================================================================
// Engine 3: regroup.  Every warp owns 64 photons whose state lives in shared memory; the 32 lanes march
// the photons that are ready to walk, and as soon as a photon needs a heavy event its lane parks it in its
// slot and picks up another ready photon.  When 32 photons wait (peel deposit + scattering, Lambert
// reflection, or an empty slot to emit into) the whole warp runs that event code once, fully converged,
// and the photons become ready again.  No lane idles while others march, no event runs for a fraction of
// a warp, and nothing leaves the SM: this is the "regroup by next event with ballot / compaction" design.
// =====================================================================================================
constexpr int RG_SLOTS = 64;     // photons per warp
constexpr int RG_STRIDE = 23;    // doubles per slot, odd -> conflict-free when lanes touch distinct slots
// slot layout (doubles): 0-2 p | 3-5 d | 6-9 S | 10 tau | 11 tau_run | 12 tacc | 13-15 w | 16 hcf | 17 wcf | 18 id |
//                        19 nd, misc | 20 t_len, t_nsc | 21 t_hash

__device__ __forceinline__ unsigned long long d2u(double v) { return (unsigned long long)__double_as_longlong(v); }
__device__ __forceinline__ double u2d(unsigned long long v) { return __longlong_as_double((long long)v); }
__device__ __forceinline__ unsigned pack_misc(const Photon& P) {
    return (unsigned)P.ph | ((unsigned)P.pk << 4) | ((P.peel_exit ? 1u : 0u) << 6) | ((P.rng.exhausted ? 1u : 0u) << 7) |
           ((P.at_walker ? 1u : 0u) << 8);
}
__device__ __forceinline__ void unpack_misc(unsigned m, Photon& P) {
    P.ph = (int)(m & 15u); P.pk = (int)((m >> 4) & 3u); P.peel_exit = ((m >> 6) & 1u) != 0u;
    P.rng.exhausted = ((m >> 7) & 1u) != 0u; P.at_walker = ((m >> 8) & 1u) != 0u;
}

// hot = what a marching lane keeps in registers
template <bool TRACE>
__device__ __forceinline__ void slot_store_hot(double* sl, const Photon& P) {
    sl[10] = P.tau; sl[11] = P.tau_run; sl[12] = P.tacc; sl[13] = P.wx; sl[14] = P.wy; sl[15] = P.wz;
    sl[17] = u2d(pack_cf(P.wc0, P.wc1, P.wc2, P.wf0, P.wf1));
    sl[19] = u2d((unsigned long long)P.rng.nd | ((unsigned long long)pack_misc(P) << 32));
    if (TRACE) { sl[20] = u2d((unsigned long long)(unsigned)P.t_len | ((unsigned long long)(unsigned)P.t_nsc << 32)); sl[21] = u2d(P.t_hash); }
}
template <bool TRACE>
__device__ __forceinline__ void slot_load_hot(const double* sl, Photon& P) {
    P.S[0] = sl[6];
    P.tau = sl[10]; P.tau_run = sl[11]; P.tacc = sl[12]; P.wx = sl[13]; P.wy = sl[14]; P.wz = sl[15];
    unpack_cf(d2u(sl[17]), P.wc0, P.wc1, P.wc2, P.wf0, P.wf1);
    const unsigned long long nm = d2u(sl[19]);
    P.rng.nd = (unsigned)nm; unpack_misc((unsigned)(nm >> 32), P);
    if (TRACE) { const unsigned long long t = d2u(sl[20]); P.t_len = (int)(unsigned)t; P.t_nsc = (int)(unsigned)(t >> 32); P.t_hash = d2u(sl[21]); }
}
// cold = the rest (home position, direction, Stokes vector, random stream)
template <bool TRACE>
__device__ __forceinline__ void slot_load_cold(const double* sl, Photon& P, unsigned long long seed, bool with_home) {
    if (with_home) { P.px = sl[0]; P.py = sl[1]; P.pz = sl[2]; unpack_cf(d2u(sl[16]), P.c0, P.c1, P.c2, P.f0, P.f1); }
    P.dx = sl[3]; P.dy = sl[4]; P.dz = sl[5];
    P.S[0] = sl[6]; P.S[1] = sl[7]; P.S[2] = sl[8]; P.S[3] = sl[9];
    P.rng.id = d2u(sl[18]);
    if (!TRACE && (P.rng.nd & 3u)) philox_block(P.rng.id, P.rng.nd >> 2, seed, P.rng);
}
__device__ __forceinline__ void slot_store_cold(double* sl, const Photon& P) {
    sl[0] = P.px; sl[1] = P.py; sl[2] = P.pz; sl[3] = P.dx; sl[4] = P.dy; sl[5] = P.dz;
    sl[6] = P.S[0]; sl[7] = P.S[1]; sl[8] = P.S[2]; sl[9] = P.S[3];
    sl[16] = u2d(pack_cf(P.c0, P.c1, P.c2, P.f0, P.f1));
    sl[18] = u2d(P.rng.id);
}

__device__ __forceinline__ int nth_set_bit64(unsigned long long m, int n) {   // position of the n-th (0-based) set bit
    const unsigned lo = (unsigned)m, hi = (unsigned)(m >> 32);
    const int cl = __popc(lo);
    return (n < cl) ? (int)__fns(lo, 0, n + 1) : 32 + (int)__fns(hi, 0, n - cl + 1);
}
__device__ __forceinline__ unsigned long long warp_or64(bool pred, int bit) {
    const unsigned long long v = pred ? (1ull << bit) : 0ull;
    const unsigned lo = __reduce_or_sync(FULL, (unsigned)v), hi = __reduce_or_sync(FULL, (unsigned)(v >> 32));
    return (unsigned long long)lo | ((unsigned long long)hi << 32);
}


// One heavy event on the photon in slot `sl` (or an emission into it), kept out of line: the march loop then
// holds no event code and keeps its registers, and the event code sees a clean register file.
// Returns x: 1 = slot now ready to march, 2 = slot now free, 0 = untouched; y,z,w: packed counter increments.
template <bool TRACE>
__device__ __noinline__ uint4 rg_event_fn(const KernelArgs* Ap, const double* sm, double* sl, int is_emit, unsigned long long k) {
    constexpr bool GEN = true;
    const KernelArgs& A = *Ap;
    const Ctx X(sm, A);
    Counters C; C.zero();
    Photon E;
    if (is_emit) ev_emit<TRACE, GEN>(X, E, C, k);
    else {
        slot_load_hot<TRACE>(sl, E);
        slot_load_cold<TRACE>(sl, E, A.L.seed, true);
        if (E.ph == PH_LAMBERT) ev_lambert<TRACE>(X, E, C);
        else {
            ev_peel_done<TRACE, GEN>(X, E, C);
            if (E.ph == PH_SCAT2) ev_scatter<TRACE>(X, E, C);
        }
    }
    unsigned res;
    if (E.ph == PH_NEW) res = is_emit ? 0u : 2u;     // a failed emission leaves the slot free as it was
    else { slot_store_cold(sl, E); slot_store_hot<TRACE>(sl, E); res = 1u; }
    return make_uint4(res | (C.n_emit << 8) | (C.n_err << 16), C.n_sc | (C.n_peel << 16), C.n_draw, C.n_surf);
}

template <bool TRACE>
__global__ void __launch_bounds__(128, 4) regroup_kernel(const __grid_constant__ KernelArgs A) {
    constexpr bool GEN = true;
    extern __shared__ double sm[];
    stage_tables(sm, A.T);
    const Ctx X(sm, A);
    const LaunchArgs& L = A.L;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* slots = sm + X.lay.total + (size_t)warp * RG_SLOTS * RG_STRIDE;
    unsigned long long ready = 0ull, pend = 0ull, freem = ~0ull;   // warp-uniform slot sets
    bool supply = true;                                            // photons left to emit (warp-uniform)
    Counters C; C.zero();
    Photon P;                        // the photon this lane is marching (hot fields; cold ones only transiently)
    P.ph = PH_NEW; P.at_walker = false; P.rng.exhausted = false; P.rng.nd = 0; P.rng.id = 0;
    int cur = -1;
    double n0 = 0, n1 = 0, n2 = 0;   // walker direction
#if !ARTES_FAITHFUL
    Ray R;
    bool need_ray = false;
    int upd = 0;
#endif

    for (;;) {
        const unsigned marching = __ballot_sync(FULL, cur >= 0);
        const int n_pend = __popcll(pend), n_free = supply ? __popcll(freem) : 0;
        const int items = n_pend + n_free;
        if (marching == 0u && ready == 0ull && items == 0) break;

        // ------------------------------------------------------------------ event round (converged)
        if (items >= 32 || (marching == 0u && ready == 0ull)) {
            __syncwarp();
            int my = -1;
            bool is_emit = false;
            if (lane < n_pend) my = nth_set_bit64(pend, lane);
            else if (lane - n_pend < n_free) { my = nth_set_bit64(freem, lane - n_pend); is_emit = true; }
            unsigned long long k = 0;
            {   // photon ids for the emitting lanes, warp-aggregated
                const unsigned em = __ballot_sync(FULL, is_emit);
                unsigned long long k0 = 0;
                if (em) {
                    const int leader = __ffs(em) - 1;
                    if (lane == leader) k0 = atomicAdd(A.O.counter, (unsigned long long)__popc(em));
                    k0 = __shfl_sync(FULL, k0, leader);
                }
                k = k0 + (unsigned long long)__popc(em & ((1u << lane) - 1u));
                const bool dry = is_emit && k >= L.n_photons;
                if (dry) { is_emit = false; my = -1; }
                if (__any_sync(FULL, dry)) supply = false;
            }
            bool now_ready = false, now_free = false;
            const bool was_pend = (my >= 0) && !is_emit;
            if (my >= 0) {
                const uint4 r = rg_event_fn<TRACE>(&A, sm, slots + my * RG_STRIDE, is_emit ? 1 : 0, k);
                now_ready = (r.x & 255u) == 1u; now_free = (r.x & 255u) == 2u;
                C.n_emit += (r.x >> 8) & 255u; C.n_err += r.x >> 16; C.n_sc += r.y & 0xffffu; C.n_peel += r.y >> 16;
                C.n_draw += r.z; C.n_surf += r.w;
            }
            const int bit = my < 0 ? 0 : my;
            pend &= ~warp_or64(was_pend, bit);
            freem = (freem & ~warp_or64(is_emit && now_ready, bit)) | warp_or64(now_free, bit);
            ready |= warp_or64(now_ready, bit);
            __syncwarp();
            continue;
        }

        // ------------------------------------------------------------------ empty lanes pick up ready photons
        {
            const unsigned empty = __ballot_sync(FULL, cur < 0);
            if (empty && ready) {
                const int rank = __popc(empty & ((1u << lane) - 1u));
                bool took = false;
                if (cur < 0 && rank < __popcll(ready)) {
                    cur = nth_set_bit64(ready, rank);
                    const double* sl = slots + cur * RG_STRIDE;
                    slot_load_hot<TRACE>(sl, P);
                    if (P.ph == PH_PEEL) { n0 = L.det[0]; n1 = L.det[1]; n2 = L.det[2]; }
                    else { n0 = sl[3]; n1 = sl[4]; n2 = sl[5]; }
                    if (TRACE) P.rng.id = d2u(sl[18]);
#if !ARTES_FAITHFUL
                    need_ray = true;
#endif
                    took = true;
                }
                ready &= ~warp_or64(took, took ? cur : 0);
            }
        }

        // ------------------------------------------------------------------ one crossing for every marching lane
#if !ARTES_FAITHFUL
        // fast mode: a fresh ray first gets its constants, then every lane with a pending axis solves ONE
        // axis per trip (same converged solver for all); a lane steps once nothing is pending.
        if (cur >= 0 && need_ray) { ray_setup(X, P, R, n0, n1, n2); upd = ray_axes(A.T); need_ray = false; }
        if (cur >= 0 && upd) ray_update(X, R, upd, P.wc0, P.wc1, P.wc2, n0, n1, n2);
        const bool step = cur >= 0 && upd == 0;
#else
        const bool step = cur >= 0;
#endif
        bool to_pend = false, to_free = false;
        if (step) {
            double* sl = slots + cur * RG_STRIDE;
            CellFace o;
#if ARTES_FAITHFUL
            cell_face(X.sm, X.lay, A.T, P.wx, P.wy, P.wz, n0, n1, n2, P.wf0, P.wf1, P.wc0, P.wc1, P.wc2, o);
            apply_crossing<TRACE, GEN>(X, P, C, o, n0, n1, n2);
#else
            int axis;
            ray_next(A.T, P, R, o, axis);
            const int ph0 = P.ph;
            apply_crossing<TRACE, GEN>(X, P, C, o, n0, n1, n2);
            if (P.ph == ph0) { R.t = (axis == 0) ? R.tr : ((axis == 1) ? R.tt : R.tp); upd = 1 << axis; }
#endif
            if (P.ph == PH_PREDONE || P.ph == PH_SCAT || P.ph == PH_SURFHIT || P.ph == PH_RETIRE) {
                // cheap follow-ups that need the random stream / Stokes vector: fetch them from the slot
                slot_load_cold<TRACE>(sl, P, L.seed, P.ph != PH_SCAT);
                cheap_handlers<TRACE, GEN>(X, P, C);
                if (P.ph != PH_NEW) {
                    slot_store_cold(sl, P);
                    if (P.ph == PH_PEEL) { n0 = L.det[0]; n1 = L.det[1]; n2 = L.det[2]; }
                    else { n0 = P.dx; n1 = P.dy; n2 = P.dz; }
#if !ARTES_FAITHFUL
                    need_ray = true;
#endif
                }
            }
            if (P.ph == PH_PEELDONE || P.ph == PH_LAMBERT) { slot_store_hot<TRACE>(sl, P); to_pend = true; }
            else if (P.ph == PH_NEW) to_free = true;
        }
        if (__any_sync(FULL, to_pend || to_free)) {
            const int bit = cur < 0 ? 0 : cur;
            pend |= warp_or64(to_pend, bit);
            freem |= warp_or64(to_free, bit);
            if (to_pend || to_free) cur = -1;
        }
    }
    C.flush(A.O.stats);
}

// =====================================================================================================
// Engine 2: wavefront.  Photon pool in HBM (SoA), queues of slot indices, three kernels per pass.
// =====================================================================================================
// what the march kernel needs of a photon
template <bool TRACE>
__device__ __forceinline__ void pool_load_all(const PoolArgs& Q, int s, Photon& P) {
    const size_t M = Q.capacity;
    const double* d = Q.d + s;
    P.px = d[0 * M]; P.py = d[1 * M]; P.pz = d[2 * M]; P.dx = d[3 * M]; P.dy = d[4 * M]; P.dz = d[5 * M];
    P.S[0] = d[6 * M]; P.S[1] = d[7 * M]; P.S[2] = d[8 * M]; P.S[3] = d[9 * M];
    P.tau = d[10 * M]; P.tau_run = d[11 * M]; P.tacc = d[12 * M]; P.wx = d[13 * M]; P.wy = d[14 * M]; P.wz = d[15 * M];
    unpack_cf(Q.hcf[s], P.c0, P.c1, P.c2, P.f0, P.f1);
    unpack_cf(Q.wcf[s], P.wc0, P.wc1, P.wc2, P.wf0, P.wf1);
    P.rng.id = Q.id[s];
    const unsigned m = Q.misc[s];
    P.ph = (int)(m & 15u); P.pk = (int)((m >> 4) & 3u); P.peel_exit = ((m >> 6) & 1u) != 0u; P.rng.exhausted = ((m >> 7) & 1u) != 0u;
    P.at_walker = false;
    P.rng.nd = Q.nd[s];
    if (!TRACE && (P.rng.nd & 3u)) philox_block(P.rng.id, P.rng.nd >> 2, Q.seed, P.rng);   // rebuild the buffered block
    if (TRACE) { P.t_len = Q.t_len[s]; P.t_nsc = Q.t_nsc[s]; P.t_hash = Q.t_hash[s]; }
}

template <bool TRACE>
__device__ __forceinline__ void pool_store_all(const PoolArgs& Q, int s, const Photon& P) {
    const size_t M = Q.capacity;
    double* d = Q.d + s;
    d[0 * M] = P.px; d[1 * M] = P.py; d[2 * M] = P.pz; d[3 * M] = P.dx; d[4 * M] = P.dy; d[5 * M] = P.dz;
    d[6 * M] = P.S[0]; d[7 * M] = P.S[1]; d[8 * M] = P.S[2]; d[9 * M] = P.S[3];
    d[10 * M] = P.tau; d[11 * M] = P.tau_run; d[12 * M] = P.tacc; d[13 * M] = P.wx; d[14 * M] = P.wy; d[15 * M] = P.wz;
    Q.hcf[s] = pack_cf(P.c0, P.c1, P.c2, P.f0, P.f1);
    Q.wcf[s] = pack_cf(P.wc0, P.wc1, P.wc2, P.wf0, P.wf1);
    Q.id[s] = P.rng.id;
    Q.nd[s] = P.rng.nd;
    Q.misc[s] = (unsigned)P.ph | ((unsigned)P.pk << 4) | ((P.peel_exit ? 1u : 0u) << 6) | ((P.rng.exhausted ? 1u : 0u) << 7);
    if (TRACE) { Q.t_len[s] = P.t_len; Q.t_nsc[s] = P.t_nsc; Q.t_hash[s] = P.t_hash; }
}

// warp-aggregated append of `slot` to queue q (counter *n) for the lanes with `pred`
// (must be reached by all 32 lanes of the warp)
__device__ __forceinline__ void queue_push(int* q, unsigned* n, bool pred, int slot) {
    const unsigned m = __ballot_sync(FULL, pred);
    if (!m) return;
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(m) - 1;
    unsigned base = 0;
    if (lane == leader) base = atomicAdd(n, (unsigned)__popc(m));
    base = __shfl_sync(FULL, base, leader);
    if (pred) q[base + __popc(m & ((1u << lane) - 1u))] = slot;
}

// queue control block in device memory
//   ctl[0] n_march[0]   ctl[1] n_march[1]   ctl[2] n_event   ctl[3] n_free   ctl[4] march cursor
//   ctl[5] cur (which march queue is being consumed)         ctl[6] photons in flight after the pass
enum { Q_NMARCH0 = 0, Q_NMARCH1 = 1, Q_NEVENT = 2, Q_NFREE = 3, Q_CURSOR = 4, Q_CUR = 5, Q_INFLIGHT = 6 };

// ---- emit: every free slot takes the next photon id, if any is left -----------------------------------
template <bool TRACE>
__global__ void __launch_bounds__(128) wf_emit_kernel(const __grid_constant__ KernelArgs A, const __grid_constant__ PoolArgs Q) {
    constexpr bool GEN = true;
    extern __shared__ double sm[];
    stage_tables(sm, A.T);
    const Ctx X(sm, A);
    Counters C; C.zero();
    const unsigned n_free = Q.ctl[Q_NFREE];
    const unsigned nxt = Q.ctl[Q_CUR] ^ 1u;
    int* q_out = Q.q_march + (size_t)nxt * Q.capacity;
    const int lane = threadIdx.x & 31;
    for (unsigned base_i = (blockIdx.x * blockDim.x + threadIdx.x) & ~31u; base_i < n_free; base_i += gridDim.x * blockDim.x) {
        const unsigned i = base_i + lane;
        const bool have = i < n_free;
        // photon ids are handed out warp-aggregated
        const unsigned m = __ballot_sync(FULL, have);
        unsigned long long k0 = 0;
        if (lane == 0) k0 = atomicAdd(A.O.counter, (unsigned long long)__popc(m));
        k0 = __shfl_sync(FULL, k0, 0);
        const unsigned long long k = k0 + __popc(m & ((1u << lane) - 1u));
        bool go = false;
        int slot = 0;
        if (have && k < A.L.n_photons) {
            slot = Q.q_free[i];
            Photon P;
            ev_emit<TRACE, GEN>(X, P, C, k);
            if (P.ph != PH_NEW) { pool_store_all<TRACE>(Q, slot, P); go = true; }
            // an emission error retires the photon at once: the slot is simply not re-queued this pass
        }
        queue_push(q_out, Q.ctl + nxt, go, slot);
    }
    C.flush(A.O.stats);
}

// ---- march: persistent lanes pull photons and walk them to their next heavy event -----------------------
template <bool TRACE>
__global__ void __launch_bounds__(128, 4) wf_march_kernel(const __grid_constant__ KernelArgs A, const __grid_constant__ PoolArgs Q) {
    constexpr bool GEN = true;
    extern __shared__ double sm[];
    stage_tables(sm, A.T);
    const Ctx X(sm, A);
    Counters C; C.zero();
    const unsigned cur = Q.ctl[Q_CUR];
    const unsigned n_in = Q.ctl[cur];
    const int* q_in = Q.q_march + (size_t)cur * Q.capacity;
    const int lane = threadIdx.x & 31;
    Photon P;
    P.ph = PH_NEW;
    int slot = -1;
    bool drained = false;
#if !ARTES_FAITHFUL
    Ray R;
    bool need_ray = true;
    int upd = 0;
    double rn0 = 0, rn1 = 0, rn2 = 0;
#endif
    for (;;) {
        // refill free lanes from the march queue
        const unsigned need = __ballot_sync(FULL, slot < 0 && !drained);
        if (need) {
            const int leader = __ffs(need) - 1;
            unsigned base = 0;
            if (lane == leader) base = atomicAdd(Q.ctl + Q_CURSOR, (unsigned)__popc(need));
            base = __shfl_sync(FULL, base, leader);
            if (slot < 0 && !drained) {
                const unsigned i = base + __popc(need & ((1u << lane) - 1u));
                if (i < n_in) {
                    slot = q_in[i]; pool_load_all<TRACE>(Q, slot, P);
#if !ARTES_FAITHFUL
                    need_ray = true;
#endif
                }
                else drained = true;
            }
        }
        if (__all_sync(FULL, slot < 0)) break;
        bool to_event = false, to_free = false;
#if ARTES_FAITHFUL
        if (slot >= 0) {
            ev_cross<TRACE, GEN>(X, P, C);
            cheap_handlers<TRACE, GEN>(X, P, C);
            to_event = (P.ph == PH_PEELDONE || P.ph == PH_LAMBERT);
            to_free = (P.ph == PH_NEW);
            if (to_event) pool_store_all<TRACE>(Q, slot, P);
        }
#else
        // fast mode: incremental ray marching (ray.cuh)
        if (slot >= 0 && need_ray) {
            const bool peel = (P.ph == PH_PEEL);
            rn0 = peel ? A.L.det[0] : P.dx; rn1 = peel ? A.L.det[1] : P.dy; rn2 = peel ? A.L.det[2] : P.dz;
            ray_setup(X, P, R, rn0, rn1, rn2); upd = ray_axes(A.T); need_ray = false;
        }
        if (slot >= 0 && upd) ray_update(X, R, upd, P.wc0, P.wc1, P.wc2, rn0, rn1, rn2);
        if (slot >= 0 && upd == 0) {
            CellFace o;
            int axis;
            ray_next(A.T, P, R, o, axis);
            const int ph0 = P.ph;
            apply_crossing<TRACE, GEN>(X, P, C, o, rn0, rn1, rn2);
            cheap_handlers<TRACE, GEN>(X, P, C);
            if (P.ph == ph0) { R.t = (axis == 0) ? R.tr : ((axis == 1) ? R.tt : R.tp); upd = 1 << axis; }
            else need_ray = true;
            to_event = (P.ph == PH_PEELDONE || P.ph == PH_LAMBERT);
            to_free = (P.ph == PH_NEW);
            if (to_event) pool_store_all<TRACE>(Q, slot, P);
        }
#endif
        if (__any_sync(FULL, to_event || to_free)) {
            queue_push(Q.q_event, Q.ctl + Q_NEVENT, to_event, slot);
            queue_push(Q.q_free, Q.ctl + Q_NFREE, to_free, slot);
            if (to_event || to_free) slot = -1;
        }
    }
    C.flush(A.O.stats);
}

// ---- event: deposit + scattering (or the continuation of a surface / thermal peel), fully converged ------
template <bool TRACE>
__global__ void __launch_bounds__(128) wf_event_kernel(const __grid_constant__ KernelArgs A, const __grid_constant__ PoolArgs Q) {
    constexpr bool GEN = true;
    extern __shared__ double sm[];
    stage_tables(sm, A.T);
    const Ctx X(sm, A);
    Counters C; C.zero();
    const unsigned n_ev = Q.ctl[Q_NEVENT];
    const unsigned nxt = Q.ctl[Q_CUR] ^ 1u;
    int* q_out = Q.q_march + (size_t)nxt * Q.capacity;
    const int lane = threadIdx.x & 31;
    for (unsigned base_i = (blockIdx.x * blockDim.x + threadIdx.x) & ~31u; base_i < n_ev; base_i += gridDim.x * blockDim.x) {
        const unsigned i = base_i + lane;
        bool go = false, freed = false;
        int slot = 0;
        if (i < n_ev) {
            slot = Q.q_event[i];
            Photon P;
            pool_load_all<TRACE>(Q, slot, P);
            if (P.ph == PH_LAMBERT) ev_lambert<TRACE>(X, P, C);
            else {
                ev_peel_done<TRACE, GEN>(X, P, C);
                if (P.ph == PH_SCAT2) ev_scatter<TRACE>(X, P, C);
            }
            if (P.ph == PH_NEW) freed = true;
            else { pool_store_all<TRACE>(Q, slot, P); go = true; }
        }
        queue_push(q_out, Q.ctl + nxt, go, slot);
        queue_push(Q.q_free, Q.ctl + Q_NFREE, freed, slot);
    }
    C.flush(A.O.stats);
}

// ---- queue bookkeeping between the stages (one thread) ---------------------------------------------------
// stage 0: before event+emit of a pass   stage 1: after them (flip queues, publish the in-flight count)
__global__ void wf_ctl_kernel(PoolArgs Q, int stage) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    if (stage == 0) {
        Q.ctl[Q.ctl[Q_CUR] ^ 1u] = 0u;   // the queue the event / emit kernels append to
        Q.ctl[Q_CURSOR] = 0u;
    } else {
        const unsigned nxt = Q.ctl[Q_CUR] ^ 1u;
        Q.ctl[Q_CUR] = nxt;
        Q.ctl[Q_NEVENT] = 0u;
        Q.ctl[Q_NFREE] = 0u;
        Q.ctl[Q_INFLIGHT] = Q.ctl[nxt];
        *Q.host_inflight = Q.ctl[nxt];
    }
}

__global__ void wf_init_kernel(PoolArgs Q) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < Q.capacity) Q.q_free[i] = (int)i;
    if (i == 0) {
        Q.ctl[Q_NMARCH0] = 0u; Q.ctl[Q_NMARCH1] = 0u; Q.ctl[Q_NEVENT] = 0u; Q.ctl[Q_NFREE] = (unsigned)Q.capacity;
        Q.ctl[Q_CURSOR] = 0u; Q.ctl[Q_CUR] = 0u; Q.ctl[Q_INFLIGHT] = 0u;
    }
}
