// engine.cuh -- photon state, the event bodies, and the persistent kernel that schedules them.
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
// Engine: transport_kernel -- persistent lanes.  One lane owns one photon until it dies; the crossing step B
// runs converged for all walking lanes, the heavy events C/E/A are ballot-deferred until enough lanes of the
// warp wait for them.  Two other schedulers were built and measured in round 1 (git history: "wavefront" =
// HBM photon pool + march/event/emit kernels with device queues; "regroup" = 64 photons per warp in shared
// memory with fully converged event rounds); both lost to this one on every configuration.

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

#ifdef ARTES_TUNING   // comparison kernel, tuning builds only: the product schedules these event bodies from engine3.cuh
// =====================================================================================================
// Engine 1: persistent lanes (one lane keeps one photon), events ballot-deferred
// =====================================================================================================
template <bool TRACE, bool GEN, bool RAY>
__global__ void __launch_bounds__(128, 4) transport_kernel(const __grid_constant__ KernelArgs A) {
    extern __shared__ double sm[];
    stage_tables(sm, A.T);
    const Ctx X(sm, A);
    const LaunchArgs& L = A.L;
    const int lane = threadIdx.x & 31;
    Photon P;
    P.ph = PH_NEW; P.rng.nd = 0; P.rng.id = 0; P.rng.exhausted = false; P.at_walker = false;
    Counters C; C.zero();
#if !ARTES_FAITHFUL
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
#if ARTES_FAITHFUL
        if (P.ph == PH_PRE || P.ph == PH_WALK || P.ph == PH_PEEL) { ev_cross<TRACE, GEN>(X, P, C); cheap_handlers<TRACE, GEN>(X, P, C); }
#else
        if (!RAY) {
            if (P.ph == PH_PRE || P.ph == PH_WALK || P.ph == PH_PEEL) { ev_cross<TRACE, GEN>(X, P, C); cheap_handlers<TRACE, GEN>(X, P, C); }
        } else {   // incremental ray marching (ray.cuh); a walk that just started first solves its axes, one per trip
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
#endif  // ARTES_TUNING
