// engine2.cuh -- the fast-mode production engine: ray marchers + block-level event batches.
// Included from transport.cuh (inside namespace artes::ARTES_NS), fast translation unit only.
//
// Why a second engine.  ncu on the persistent-lane engine (engine.cuh) at C4: 13 of 32 lanes active on
// average, 35 % issue utilisation, `no_instruction` (i-cache) the top stall, 12 900 thread instructions
// per scattering cycle (profiles/r01_g_*).  The lanes diverge because every lane runs its own photon's
// whole life -- emission, first optical depth, survival, peel-off weight, CDF sampling, deposit -- inline
// between cell crossings.  Here the photon's life is cut into RAYS and EVENTS:
//
//   ray     a straight segment through the r-theta-phi grid (tau pre-pass :633-656, transport walk
//           :691-778 / :850-941, peel-off walk :4739-4761).  A lane ("marcher") owns one ray at a time and
//           does nothing but step it radially in a tight loop; its whole state is ~30 registers.
//   event   everything else.  A photon lives in a SLOT: the hot ray state in shared memory, the cold state
//           (position, direction, Stokes vector, optical depth; the photon id and its stream position are hot ints) in a 160-byte
//           L2-resident record.  When its ray ends the marcher writes (t, tau) back to the slot and pushes
//           the slot on the list of the event it needs; whenever a list holds 32 entries some warp takes
//           them and runs the event fully converged.  Each event ends by setting up the slot's next ray
//           and pushing the slot on the ready list, from which free marcher lanes claim.
//
//   EMIT  emit_photon :1008-1115 (star)                     -> pre-pass ray
//   PRE   first optical depth :660-685                      -> transport ray
//   H     interaction: survival :791-813, peel-off weight + pixel (peel_photon :4763-4951 without the
//         e^-tau factor), scatter_photon + polarization_rotation :819-823, next tau :845   -> peel ray
//   DEP   e^-tau, detector deposit :4955-4972               -> transport ray
//   RES   a theta or phi face was crossed: move the cell index, re-solve that axis        -> same ray
//
// Geometry.  Along X(t) = X0 + t n every bounding surface is a fixed quadric in t (ray.cuh).  The three
// axes are independent crossing sequences that only need merging by t:
//   r      |X|^2 has one minimum: inward the next sphere is the inner one (if the ray reaches it), after
//          the turning point the outer one: t = -hb/qa -+ sqrt(D0 + r_k^2/qa): one sqrt, no division
//   theta  cos(theta) has one extremum (g(t) = g0 + g1 t is the sign of its derivative): try the face in
//          the direction of motion, then the other one
//   phi    the azimuth is monotonic (sign of the angular momentum L_z): one linear solve
// Radial crossings dominate by far (10 km layers against 10^4 km wide theta/phi cells), so only the
// radial re-solve is in the marcher; theta/phi crossings are RES events.
//
// The draw order of the random stream and every physical formula are those of the oracle, so for the
// same Philox stream the trajectories agree with the oracle's to rounding.
//
// Two kernels schedule the same marcher and events: transport3_kernel (asynchronous, no block barrier; the
// default) and transport2_kernel (bulk-synchronous rounds; kept for comparison, a few per cent slower).
//
// Scope: everything except flow_global and oblate planets (those run on the persistent-lane engine).  The GEN
// instantiation compiles in the thermal source, the reflecting surface (SURF event) and the latitudinal flow
// counters; the TRACE instantiation the injected-stream walk recorder; the BATCH instantiation a launch index per photon
// (artes_gpu_run_batch: the 73 launches of a phase curve or the wavelengths of a spectrum as ONE kernel -- the work
// counter covers all launches, detector geometry / cell_depth / wavelength of each launch sit in a shared-memory table,
// images and fluxes of the launches follow each other in the output buffer).
//
// Warp-uniform decisions.  Every condition the whole warp acts on (claim, take a batch, leave the kernel) is read by ONE
// lane and broadcast, or is the result of a vote: a volatile shared-memory read per lane is not warp-uniform, and a warp
// that splits on it dead-locks in the next *_sync.

namespace e2 {

#include "fastmath.cuh"

// measurement switches (tools/build_variants.sh builds the library with single optimisations turned off)
#ifndef E2_OPT_R2S
#define E2_OPT_R2S 1       // squared radii through ld.shared with a 32-bit address
#endif
#ifndef E2_EARLY_ROOT
#define E2_EARLY_ROOT 0    // the radial root of the NEXT cell is computed at the top of a step, next to the min / optical-depth chain it does
#endif                     // not depend on (two independent dependency chains per step instead of one long one); same operations, same results
#ifndef E2_RELEASE
#define E2_RELEASE 0       // 1: a warp keeps no ray across passes -- unfinished rays go back to the ready list, and a pass starts only
#endif                     //    with at least E2_MARCH_MIN ready rays (or when there is no full event batch to run instead)
#ifndef E2_MARCH_MIN
#define E2_MARCH_MIN 24
#endif
#ifndef E2_EVENT_POLICY
#define E2_EVENT_POLICY 1  // which event list a warp serves: 1 the one with the largest backlog, 0 a fixed priority order (RES, DEP, H, SURF, PRE, EMIT)
#endif
#ifndef E2_WATCHDOG
#define E2_WATCHDOG 1      // 1: bounded waits on list cells (see ring_take); 2: also the idle-turn watchdog of the scheduling loop
#endif                     //    (measured 4-5 % on C4, profiles/r02_g_*: debug / stress builds only, tools/gpu_stress.py)

// A block has NT marcher lanes and NP >= NT photon slots: with more photons than lanes a lane whose ray
// ended finds another ready ray at once, and a round collects enough events to keep every warp busy in
// the event phase.  RC = ring capacity of the lists (power of two >= NP).

enum : int { K_PRE = 0, K_WALK = 1, K_PEEL = 2, K_DEAD = 3 };
enum : int { L_EMIT = 0, L_PRE, L_H, L_DEP, L_RES, L_SURF, L_FAN, L_SC, L_RDY, N_LISTS };   // event lists (L_FAN, L_SC: multi-detector walks only), then the ready list
constexpr int N_EVENT_LISTS = L_RDY;
enum : int { O_NONE = 0, O_LIMIT, O_EXIT, O_SURF, O_REST, O_RESP, O_ERR, O_DEAD };
// info word: bits 0-1 kind, 2 radial inward, 3 next theta face is the upper one, 4 phi increasing, 8-11 outcome
enum : int { B_INWARD = 4, B_TUPPER = 8, B_PUP = 16, PK_SHIFT = 5 };   // bits 5-6: peel kind of a K_PEEL ray (PK_SCATTER/SURFACE/THERMAL)

// Slot fields.  COLD fields (touched by events only) live in a 160-byte record per slot in global memory
// (L2-resident scratch, one region per block); HOT fields (the ray a marcher loads and stores) live in
// shared memory as a structure of arrays [field][NP].  Keeping the cold part out of shared memory is what
// lets a block hold several times more photons than lanes.
enum : int { F_PX = 0, F_PY, F_PZ, F_DX, F_DY, F_DZ, F_S0, F_S1, F_S2, F_S3, F_TAU, F_W0, F_W1, F_W2, F_W3, NF_COLD,
             F_T = NF_COLD, F_ACC, F_TR, F_TT, F_TP, F_HBN, F_D0, F_IQ, F_LIM, NF_D };
enum : int { I_CELL = 0, I_INFO, I_ND, I_IDLO, I_IDHI, I_BATCH, NI_HOT, I_HCELL = NI_HOT, I_PIX,
             I_TLEN, I_TNSC, I_THLO, I_THHI, I_FLAG, NF_I };
// I_ND draw counter, I_IDLO/HI photon id, I_BATCH launch index of the photon in a batched launch: hot (shared memory), every interaction
// and every ray load of a wavelength batch reads them; I_T*: walk recorder of the trace hook; I_FLAG bit 0: injected stream used up
constexpr int NF_HOT = NF_D - NF_COLD;
constexpr int REC = 20;       // doubles per cold record: 15 doubles + 10 ints = 160 bytes

__host__ __device__ constexpr int ring_cap(int np_slots) { int c = 32; while (c < np_slots) c <<= 1; return c; }

// Shared-memory layout.  The slot arrays, lists and list counters come FIRST, at offsets that depend only on the
// template parameter NP: their addresses are immediates in the instruction stream instead of values the compiler has
// to keep in (or spill from) registers around the whole scheduling loop.  The grid tables, whose sizes are run-time
// values, follow.
__host__ __device__ constexpr int fixed_doubles(int NP) {
    return (NF_HOT * NP * 8 + NI_HOT * NP * 4 + N_LISTS * ring_cap(NP) * 2 + 64 * 4 + 15) / 16 * 2;      // 64 ints: head[16] | tail[16] | misc[32]
}
constexpr int GEO = 12;     // doubles per launch in the launch table: det(3), sin_dt, cos_dt, sin_dp, cos_dp, limb_emission, det_sph_theta, det_sph_phi,
                            // cell_depth and wavelength index of the launch (batches over wavelengths: LaunchArgs::wl_batch)
#ifndef E2_COMPACT_M
#define E2_COMPACT_M 1     // block-diagonal scattering matrices are read from the eight-element copy (DevTables::Mc)
#endif
struct Lay {
    int o_r, o_r2, o_tf, o_tt, o_ps, o_pc, o_pf, o_tp, o_ca, o_geo, o_det, n_end;   // offsets in doubles
    size_t bytes;
    // nb: launches in the launch table; ndet: doubles of the block-private detector image (0: the image stays in global memory)
    __host__ __device__ Lay(int nr, int nt, int np, int NP, int nb = 1, int ndet = 0) {
        o_r = fixed_doubles(NP);
        o_r2 = o_r + nr + 1; o_tf = o_r2 + nr + 1; o_tt = o_tf + nt + 1; o_ps = o_tt + nt + 1; o_pc = o_ps + np; o_pf = o_pc + np;
        o_tp = o_pf + np; o_ca = ((o_tp + (nt + 2) / 2 + 1) + 1) & ~1; o_geo = o_ca + 2 * 181; o_det = o_geo + GEO * (nb < 1 ? 1 : nb);   // o_ca, o_geo: 16-byte aligned
        n_end = o_det + ndet;
        bytes = (size_t)n_end * 8;
    }
};

template <int NP, bool TR = false, bool GN = false, bool BT = false, bool MD = false>
struct ShT {                     // pointers into the block's shared memory
    static constexpr bool TRACE = TR;   // the injected-stream walk recorder (test hook) is compiled in
    static constexpr bool BATCH = BT;   // batched launches: per-photon launch index, detector geometry from the shared-memory table
    static constexpr bool GEN = GN;     // thermal source, reflecting surface and latitudinal flow counters are compiled in
    static constexpr bool MULTI = MD;   // ONE walk, peel-off towards every detector of the launch table (artes_gpu_run_multi)
    double* sdet;                // block-private detector image in shared memory (null: global atomics)
    const double* r; const double* r2; const double* tf; const double* ttan; const double* ps; const double* pc; const double* pf;
    const int* tplane;
    const double* geo;           // [n_batch][GEO] detector geometry of the launch(es)
    const double2* cdfa;         // azimuth prefix table [181] (cos2beta, sin2beta): 17 probes per scattering, kept out of the L1 global path
    double* sd; int* si; short* q; int* head; int* tail;
    int* misc;                   // [0] slots retired for good, [1] event batch counter, [2] ready-list tail at the end of the last event phase
    double* cold;                // this block's records in global memory
    static constexpr int RC = ring_cap(NP);
    __device__ __forceinline__ double& D(int f, int s) const { return (f >= NF_COLD) ? sd[(f - NF_COLD) * NP + s] : cold[(size_t)s * REC + f]; }
    __device__ __forceinline__ int& I(int f, int s) const {
        return (f < NI_HOT) ? si[f * NP + s] : reinterpret_cast<int*>(cold + (size_t)s * REC + NF_COLD)[f - NI_HOT];
    }
    __device__ __forceinline__ short& Q(int l, int pos) const { return q[l * RC + (pos & (RC - 1))]; }
};

// 256-bit global accesses (LDG.E.256 / STG.E.256 on sm_100a): a warp instruction whose lanes read scattered
// 32-byte pieces costs one L1 wavefront per lane whatever the width, so the wide form halves the load of the L1
// data pipe (65 % busy in the profile of the 128-bit version) for the matrix rows, CDF entries and photon records.
__device__ __forceinline__ void ldg256_nc(const double* p, double& a, double& b, double& c, double& d) {   // read-only tables
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
}
__device__ __forceinline__ void ldg256(const double* p, double& a, double& b, double& c, double& d) {      // data written by this kernel
    asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p) : "memory");
}
__device__ __forceinline__ void stg256(double* p, double a, double b, double c, double d) {
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}

__device__ __forceinline__ int pack_cell(int c0, int c1, int c2) { return c0 | (c1 << 10) | (c2 << 20); }

// Philox draws nd .. nd+4 of photon `id` (at most two counter blocks)
struct Draws {
    uint4 a, b;
    unsigned off;
    __device__ __forceinline__ Draws(unsigned long long id, unsigned nd, unsigned long long seed, int count) {
        off = nd & 3u;
        a = philox4(id, nd >> 2, seed);
        b = a;
        if (off + (unsigned)count > 4u) b = philox4(id, (nd >> 2) + 1u, seed);
    }
    __device__ __forceinline__ double get(int i) const {
        const unsigned j = off + (unsigned)i, w = j & 3u;
        const unsigned x = (j < 4u) ? ((w == 0u) ? a.x : (w == 1u) ? a.y : (w == 2u) ? a.z : a.w)
                                    : ((w == 0u) ? b.x : (w == 1u) ? b.y : (w == 2u) ? b.z : b.w);
        return ((double)x + 0.5) * (1.0 / 4294967296.0);
    }
};

// `count` (<= 5) consecutive draws of slot s starting at draw index nd: the Philox stream, or -- trace hook -- the
// injected stream xi[photon][max_draws] (a draw past its end returns 0.5 and marks the stream as used up,
// like rng_next<true> of transport.cuh)
template <class Sh>
__device__ __forceinline__ void draws(const Sh& X, const KernelArgs& A, int s, unsigned long long id, unsigned nd, int count, double* xi) {
    if (Sh::TRACE) {
        for (int i = 0; i < count; ++i) {
            if ((int)(nd + i) >= A.R.max_draws) { xi[i] = 0.5; X.I(I_FLAG, s) |= 1; }
            else xi[i] = A.R.xi[(size_t)(id - A.L.id_base) * A.R.max_draws + nd + i];
        }
    } else {
        const Draws d(id, nd, A.L.seed, count);
#pragma unroll
        for (int i = 0; i < 5; ++i) if (i < count) xi[i] = d.get(i);
    }
}

// walk recorder of the trace hook: one (a,b,c,d,e) tuple per cell_face outcome / detector deposit
__device__ __forceinline__ void trace_tuple(const KernelArgs& A, unsigned long long id, int& tlen, unsigned long long& thash,
                                            int a, int b, int c, int d, int e) {
    if (A.R.seq_head && tlen < A.R.max_rec) {
        int* p = A.R.seq_head + ((size_t)(id - A.L.id_base) * A.R.max_rec + tlen) * 5;
        p[0] = a; p[1] = b; p[2] = c; p[3] = d; p[4] = e;
    }
    tuple_hash(thash, a); tuple_hash(thash, b); tuple_hash(thash, c); tuple_hash(thash, d); tuple_hash(thash, e);
    ++tlen;
}

// quadric constants of a ray, event side
struct RayK { double A1, A2, B1, B2, C1, C2, g0, g1, Xx, Xy, Nx, Ny, z0, n2; };

__device__ __forceinline__ void ray_consts(const DevTables& T, double x, double y, double z, double n0, double n1, double n2,
                                           RayK& K, double& hbn, double& D0, double& iq) {
    const double a = T.inv_ox, b = T.inv_oy, a2 = a * a, b2 = b * b, c2 = T.inv_oz * T.inv_oz;
    K.A1 = a2 * n0 * n0 + b2 * n1 * n1; K.A2 = c2 * n2 * n2;
    K.B1 = a2 * x * n0 + b2 * y * n1;   K.B2 = c2 * z * n2;
    K.C1 = a2 * x * x + b2 * y * y;     K.C2 = c2 * z * z;
    const double qa = K.A1 + K.A2, hb = K.B1 + K.B2, Cs = K.C1 + K.C2;
    iq = frcp(qa); hbn = -hb * iq; D0 = fma(hbn, hbn, -(Cs * iq));
    K.g0 = n2 * Cs - z * hb; K.g1 = n2 * hb - z * qa;      // sign of d(cos theta)/dt = sign(g0 + g1 t)
    K.Xx = a * x; K.Xy = b * y; K.Nx = a * n0; K.Ny = b * n1; K.z0 = z; K.n2 = n2;
}

// first radial crossing of a fresh ray in cell c0 (sface = radial face the ray starts on, or -1)
template <class Sh>
__device__ __forceinline__ double radial_first(const Sh& X, int c0, int sface, double hbn, double D0, double iq, int& inward) {
    inward = hbn > 0.0;
    if (inward) {
        const double r = X.r[c0], disc = fma(r * r, iq, D0);
        if (disc >= 0.0) { const double t = hbn - fsqrt(disc); if (t > 1.e-15) return t; }
        inward = 0;
    }
    const double r = X.r[c0 + 1], disc = fma(r * r, iq, D0);
    if (disc >= 0.0) {
        const double t = hbn + fsqrt(disc);
        if (t > ((sface == c0 + 1) ? 1.e-3 : 1.e-15)) return t;     // same-face threshold :2944
    }
    return RAY_NONE;
}

// smallest root > lim of qa t^2 + 2 hb t + qc = 0 on the right nappe (ray.cuh quadric_next with the fast operations)
__device__ __forceinline__ double quadric_next_f(double qa, double hb, double qc, int hs, double lim, double z0, double n2) {
    const double disc = hb * hb - qa * qc;
    if (!(disc >= 0.0)) return RAY_NONE;
    const double q = -(hb + copysign(fsqrt(disc), hb));
    double t1 = (fabs(qa) > 1.e-100) ? fdiv(q, qa) : RAY_NONE;
    double t2 = (fabs(q) > 1.e-100) ? fdiv(qc, q) : RAY_NONE;
    if (!(t1 > lim) || (z0 + t1 * n2) * (double)hs < 0.0) t1 = RAY_NONE;
    if (!(t2 > lim) || (z0 + t2 * n2) * (double)hs < 0.0) t2 = RAY_NONE;
    return fmin(t1, t2);
}

template <class Sh>
__device__ __forceinline__ double cone_root(const Sh& X, int k, double t, const RayK& K) {
    const int tp = X.tplane[k];
    if (tp != 1) {   // equatorial plane :3068 / :3118
        if (tp != 2 || K.n2 == 0.0) return RAY_NONE;
        const double r = fdiv(-K.z0, K.n2);
        return (r > t) ? r : RAY_NONE;
    }
    const double tn = X.ttan[k], T2 = tn * tn, tf = X.tf[k];
    const int hs = (tf < PI / 2.0) ? 1 : ((tf > PI / 2.0) ? -1 : 0);
    return quadric_next_f(K.A1 - K.A2 * T2, K.B1 - K.B2 * T2, K.C1 - K.C2 * T2, hs, t, K.z0, K.n2);
}

// next polar crossing after parameter t from cell c1: face in the direction of motion first, then the other
template <class Sh>
__device__ __forceinline__ double theta_next(const Sh& X, int nt, int c1, double t, const RayK& K, int& upper) {
    bool down = (K.g0 + K.g1 * t) > 0.0;     // cos(theta) increasing: towards the lower face index
    upper = 0;
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
        const int k = down ? c1 : c1 + 1;
        if (k != 0 && k != nt) {
            const double tk = cone_root(X, k, t, K);
            if (tk < RAY_NONE) { upper = down ? 0 : 1; return tk; }
        }
        down = !down;
    }
    return RAY_NONE;
}

template <class Sh>
__device__ __forceinline__ double phi_next(const Sh& X, int np, int c2, double t, const RayK& K, int& up) {
    up = (K.Xx * K.Ny - K.Xy * K.Nx) > 0.0;
    if (np <= 1) return RAY_NONE;
    const int k = up ? ((c2 + 1 == np) ? 0 : c2 + 1) : c2;
    const double ps = X.ps[k], pc = X.pc[k];
    const double den = K.Ny * pc - K.Nx * ps;
    if (den == 0.0) return RAY_NONE;
    const double r = fdiv(K.Xx * ps - K.Xy * pc, den);
    return (r > t) ? r : RAY_NONE;
}

// What an event asks for: the next ray of its slot.  The set-up itself (three axis solves) is done once, after the
// event switch (run_event), so that the kernel holds one copy of that code instead of one per event.
struct RaySpec {
    bool make;
    double x, y, z, n0, n1, n2, lim, acc0;
    int c0, c1, c2, sface, kind, pk;
    __device__ __forceinline__ void set(double x_, double y_, double z_, double n0_, double n1_, double n2_, int c0_, int c1_, int c2_,
                                        int sface_, int kind_, double lim_, double acc0_ = 0.0, int pk_ = 0) {
        make = true; x = x_; y = y_; z = z_; n0 = n0_; n1 = n1_; n2 = n2_; c0 = c0_; c1 = c1_; c2 = c2_;
        sface = sface_; kind = kind_; lim = lim_; acc0 = acc0_; pk = pk_;
    }
};

// set up the next ray of slot s: origin (x,y,z) in cell (c0,c1,c2), direction n
template <class Sh>
__device__ __forceinline__ void ray_setup(const Sh& X, const DevTables& T, int s, double x, double y, double z,
                                          double n0, double n1, double n2, int c0, int c1, int c2, int sface, int kind, double lim,
                                          double acc0 = 0.0, int pk = 0) {
    RayK K;
    double hbn, D0, iq;
    ray_consts(T, x, y, z, n0, n1, n2, K, hbn, D0, iq);
    int inward, upper = 0, up = 0;
    const double tr = radial_first(X, c0, sface, hbn, D0, iq, inward);
    const double tt = (T.nt > 1) ? theta_next(X, T.nt, c1, 0.0, K, upper) : RAY_NONE;
    const double tp = phi_next(X, T.np, c2, 0.0, K, up);
    X.D(F_T, s) = 0.0; X.D(F_ACC, s) = acc0; X.D(F_TR, s) = tr; X.D(F_TT, s) = tt; X.D(F_TP, s) = tp;
    X.D(F_HBN, s) = hbn; X.D(F_D0, s) = D0; X.D(F_IQ, s) = iq; X.D(F_LIM, s) = lim;
    X.I(I_CELL, s) = pack_cell(c0, c1, c2);
    X.I(I_INFO, s) = kind | (inward ? B_INWARD : 0) | (upper ? B_TUPPER : 0) | (up ? B_PUP : 0) | (pk << PK_SHIFT);
}

// matrix_at_deg (transport.cuh) with 256-bit row reads
__device__ __forceinline__ void matrix_at_deg_f(const DevTables& T, int u, double deg, double F[16]) {
    int lo, up;
    const double fl = floor(deg);
    if (deg - fl > 0.5) { up = (int)fl + 2; lo = (int)fl + 1; }
    else { up = (int)fl + 1; lo = (int)fl; }
    const double* base = T.M + (size_t)u * (180 * 16);
    if (up <= 1 || lo >= 180) {
        const double* m = base + (up <= 1 ? 0 : 179) * 16;
#pragma unroll
        for (int i = 0; i < 4; ++i) ldg256_nc(m + 4 * i, F[4 * i], F[4 * i + 1], F[4 * i + 2], F[4 * i + 3]);
    } else {
        const double* m0 = base + (lo - 1) * 16;
        const double* m1 = base + (up - 1) * 16;
        const double w = deg - ((double)lo - 0.5);
        double v0[16], v1[16];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            ldg256_nc(m0 + 4 * i, v0[4 * i], v0[4 * i + 1], v0[4 * i + 2], v0[4 * i + 3]);
            ldg256_nc(m1 + 4 * i, v1[4 * i], v1[4 * i + 1], v1[4 * i + 2], v1[4 * i + 3]);
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) F[i] = (v1[i] - v0[i]) * w + v0[i];
    }
}

// polrot_fast (transport.cuh) with the fast division / square root
__device__ __forceinline__ int polrot_f(double c2a, double s2a, bool flip, double nc2, const double Sin[4],
                                        const double F[16], double Sout[4], bool peeling, int& soft) {
    if (!(fabs(nc2) < 1.00001)) return 11;
    nc2 = fmin(fmax(nc2, -1.0), 1.0);
    const double r0 = Sin[0], r1 = c2a * Sin[1] + s2a * Sin[2], r2 = c2a * Sin[2] - s2a * Sin[1], r3 = Sin[3];
    double s[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) s[r] = F[4 * r] * r0 + F[4 * r + 1] * r1 + F[4 * r + 2] * r2 + F[4 * r + 3] * r3;
    if (!peeling) {
        if (s[0] > 0.0) { const double nrm = fdiv(r0, s[0]); s[0] = r0; s[1] *= nrm; s[2] *= nrm; s[3] *= nrm; }
        else soft = 12;
    }
    const double c2 = 2.0 * nc2 * nc2 - 1.0;
    double s2 = 2.0 * nc2 * fsqrt(fmax(1.0 - nc2 * nc2, 0.0));
    if (flip) s2 = -s2;
    Sout[0] = s[0];
    Sout[1] = c2 * s[1] + s2 * s[2];
    Sout[2] = c2 * s[2] - s2 * s[1];
    Sout[3] = s[3];
    return 0;
}

// polrot_f with the scattering matrix taken straight from the table: F(deg) is interpolated between the two bracketing matrix
// rows (matrix_at_deg_f) and multiplied with the rotated Stokes vector ROW BY ROW, so that only eight table values are live at a
// time instead of F[16] and the two 16-value table rows (96 registers): the interaction event is what sets the kernel's
// register budget.  Same arithmetic as matrix_at_deg_f + polrot_f.
__device__ __forceinline__ int polrot_deg_f(const DevTables& T, int u, double deg, double c2a, double s2a, bool flip, double nc2,
                                            const double Sin[4], double Sout[4], bool peeling, int& soft) {
    if (!(fabs(nc2) < 1.00001)) return 11;
    nc2 = fmin(fmax(nc2, -1.0), 1.0);
    const double r0 = Sin[0], r1 = c2a * Sin[1] + s2a * Sin[2], r2 = c2a * Sin[2] - s2a * Sin[1], r3 = Sin[3];
    int lo, up;
    const double fl = floor(deg);
    if (deg - fl > 0.5) { up = (int)fl + 2; lo = (int)fl + 1; }
    else { up = (int)fl + 1; lo = (int)fl; }
    const bool edge = (up <= 1 || lo >= 180);
    const int row0 = edge ? (up <= 1 ? 0 : 179) : (lo - 1), row1 = edge ? row0 : (up - 1);
    const double w = edge ? 0.0 : deg - ((double)lo - 0.5);
    double s[4];
    if (E2_COMPACT_M && T.Mc) {
        // block-diagonal matrices (DevTables::Mc): rows 0, 1 act on (r0, r1), rows 2, 3 on (r2, r3); the terms left out are products
        // with exact zeros.  Half the table bytes and half the 32-byte reads (one L1 wavefront per lane and read) of the general path.
        const double* base = T.Mc + (size_t)u * (180 * 8);
        const double* m0 = base + row0 * 8;
        const double* m1 = base + row1 * 8;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            double a0, a1, a2, a3, b0, b1, b2, b3;
            ldg256_nc(m0 + 4 * h, a0, a1, a2, a3);
            ldg256_nc(m1 + 4 * h, b0, b1, b2, b3);
            const double x = h ? r2 : r0, y = h ? r3 : r1;
            // the operation order of the general path's contracted sum: fma(F_r3, r3, fma(F_r2, r2, fma(F_r0, r0, F_r1 * r1)))
            const double f0 = (b0 - a0) * w + a0, f1 = (b1 - a1) * w + a1, f2 = (b2 - a2) * w + a2, f3 = (b3 - a3) * w + a3;
            s[2 * h] = h ? fma(f1, y, f0 * x) : fma(f0, x, f1 * y);
            s[2 * h + 1] = h ? fma(f3, y, f2 * x) : fma(f2, x, f3 * y);
        }
    } else {
        const double* base = T.M + (size_t)u * (180 * 16);
        const double* m0 = base + row0 * 16;
        const double* m1 = base + row1 * 16;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            double a0, a1, a2, a3, b0, b1, b2, b3;
            ldg256_nc(m0 + 4 * r, a0, a1, a2, a3);
            ldg256_nc(m1 + 4 * r, b0, b1, b2, b3);
            s[r] = ((b0 - a0) * w + a0) * r0 + ((b1 - a1) * w + a1) * r1 + ((b2 - a2) * w + a2) * r2 + ((b3 - a3) * w + a3) * r3;
        }
    }
    if (!peeling) {
        if (s[0] > 0.0) { const double nrm = fdiv(r0, s[0]); s[0] = r0; s[1] *= nrm; s[2] *= nrm; s[3] *= nrm; }
        else soft = 12;
    }
    const double c2 = 2.0 * nc2 * nc2 - 1.0;
    double s2 = 2.0 * nc2 * fsqrt(fmax(1.0 - nc2 * nc2, 0.0));
    if (flip) s2 = -s2;
    Sout[0] = s[0];
    Sout[1] = c2 * s[1] + s2 * s[2];
    Sout[2] = c2 * s[2] - s2 * s[1];
    Sout[3] = s[3];
    return 0;
}

// Inversion of a 180-bin cumulative table: returns the bin lo with cum(lo) < samp <= cum(lo + 1) and the two
// table values.  Three 6-ary rounds (steps 30, 5, 1; independent probes each) instead of eight dependent
// binary-search steps: the same number of table reads, a third of the load round trips.  The last round reads
// the six entries lo .. lo+5, so the bracketing pair comes out of registers instead of a fourth round trip.
#ifndef E2_BSEARCH
#define E2_BSEARCH 3       // bit 0: polar CDF, bit 1: azimuth CDF inverted by binary search (8 dependent probes instead of 16 in three rounds)
#endif
// The same bin by bisection: 8 probes instead of 16.  y180 = cum(180) (the caller has it), cum(0) = 0 (prefix tables).
template <class F>
__device__ __forceinline__ int search2(F cum, double samp, double y180, double& ylo, double& yhi) {
    int lo = 0, hi = 180;          // cum(lo) < samp (or lo = 0), cum(hi) >= samp (or hi = 180)
    ylo = 0.0; yhi = y180;
#pragma unroll
    for (int it = 0; it < 8; ++it) {
        const int mid = (lo + hi) >> 1;
        const double v = cum(mid);
        const bool below = (v < samp) && (hi - lo > 1);
        const bool above = !(v < samp) && (hi - lo > 1);
        if (below) { lo = mid; ylo = v; }
        if (above) { hi = mid; yhi = v; }
    }
    return lo;
}

// cum30(k) = cum(30 k) for k = 1..5: the first round's probes, which may come from a copy that is cheaper to read
template <class F, class F30>
__device__ __forceinline__ int search6(F cum, F30 cum30, double samp, double& ylo, double& yhi) {
    int lo = 0;
#pragma unroll
    for (int round = 0; round < 2; ++round) {
        const int step = (round == 0) ? 30 : 5;
        double y[5];
#pragma unroll
        for (int k = 0; k < 5; ++k) y[k] = (round == 0) ? cum30(k + 1) : cum(lo + (k + 1) * step);
        int c = 0;
#pragma unroll
        for (int k = 0; k < 5; ++k) c += (y[k] < samp) ? 1 : 0;
        lo += c * step;
    }
    double y[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) y[k] = cum(lo + k);
    int c = 0;
#pragma unroll
    for (int k = 1; k < 5; ++k) c += (y[k] < samp) ? 1 : 0;
    ylo = y[0]; yhi = y[1];
#pragma unroll
    for (int k = 1; k < 5; ++k) if (c == k) { ylo = y[k]; yhi = y[k + 1]; }
    return lo + c;
}

// sample_angles_fast with the three random numbers supplied by the caller (engine2: stateless Philox draws)
template <class Sh>
__device__ __forceinline__ int sample_angles_f(const Sh& X, const KernelArgs& A, double xi1, double xi2, double xi3, const double S[4],
                                                     int cellidx, FastAngles& g, int u) {
    const DevTables& T = A.T;
    (void)cellidx;
    double p11, p12, p13, p14;
    ldg256_nc(T.p1k + 4 * u, p11, p12, p13, p14);
    const double Ac = p11 * S[0] + p14 * S[3], Bc = p12 * S[1] + p13 * S[2], Cc = p12 * S[2] - p13 * S[1];
    auto cumA = [&](int i) { const double2 q = X.cdfa[i]; return Ac * (double)i + Bc * q.x + Cc * q.y; };
    const double a180 = cumA(180);
    double samp = xi1 * a180;
    double ylo, yhi;
    int lo = (E2_BSEARCH & 2) ? search2(cumA, samp, a180, ylo, yhi)
                              : search6(cumA, [&](int k) { return cumA(30 * k); }, samp, ylo, yhi);   // bin lo: cum(lo) < samp <= cum(lo + 1)
    double fr = fdiv(samp - ylo, yhi - ylo);
    if (!(fr == fr)) return 6;
    fr = fmin(fmax(fr, 0.0), 1.0);
    const double beta = (fr + (double)lo) * (PI / 180.0);
    fm_sincos_0pi(beta, &g.sb, &g.cb);
    g.flip = xi2 > 0.5;                       // beta + pi  (:1589-1590)
    if (g.flip) { g.sb = -g.sb; g.cb = -g.cb; }
    const double c2b = g.cb * g.cb - g.sb * g.sb, s2b = 2.0 * g.sb * g.cb;
    const double w1 = S[0], w2 = c2b * S[1] + s2b * S[2], w3 = c2b * S[2] - s2b * S[1], w4 = S[3];
    const double* tab = T.cdfP + (size_t)u * (181 * 4);
    auto cumP = [&](int i) {
        double q0, q1, q2, q3;
        ldg256_nc(tab + 4 * i, q0, q1, q2, q3);
        return w1 * q0 + w2 * q1 + w3 * q2 + w4 * q3;
    };
    // cumP(180) without reading the table: row 180 of cdfP is p1k, built by the same additions in the same order (artes_gpu.cu)
    const double p180 = w1 * p11 + w2 * p12 + w3 * p13 + w4 * p14;
    samp = xi3 * p180;
    lo = (E2_BSEARCH & 1) ? search2(cumP, samp, p180, ylo, yhi) : search6(cumP, [&](int k) { return cumP(30 * k); }, samp, ylo, yhi);
    fr = fdiv(samp - ylo, yhi - ylo);
    if (!(fr == fr)) return 7;
    fr = fmin(fmax(fr, 0.0), 1.0);
    g.deg = fr + (double)lo;
    fm_sincos_0pi(g.deg * (PI / 180.0), &g.sT, &g.alpha);
    if (g.alpha >= 1.0) { g.alpha = 1.0 - 1.e-10; g.sT = sqrt(1.0 - g.alpha * g.alpha); }
    if (g.alpha <= -1.0) { g.alpha = -1.0 + 1.e-10; g.sT = sqrt(1.0 - g.alpha * g.alpha); }
    return 0;
}

// detector geometry of launch kb (shared memory; one entry for a plain launch)
// cos of the angle between the surface normal at (x,y,z) and the detector (:4609-4634), detector angles of the photon's launch
__device__ __noinline__ double surface_cos_angle2(const DevTables& T, double x, double y, double z, double dth, double dph) {
    double s0 = x / (T.ox * T.ox), s1 = y / (T.oy * T.oy), s2 = z / (T.oz * T.oz);
    double nrm = sqrt(s0 * s0 + s1 * s1 + s2 * s2);
    s0 = s0 / nrm; s1 = s1 / nrm; s2 = s2 / nrm;
    double nth = acos(s2 / sqrt(s0 * s0 + s1 * s1 + s2 * s2));
    double nph = atan2(s1, s0);
    if (nph < 0.0) nph = nph + 2.0 * PI;
    return sin(dth) * cos(dph) * sin(nth) * cos(nph) + sin(dth) * sin(dph) * sin(nth) * sin(nph) + cos(dth) * cos(nth);
}
struct Geo { double d0, d1, d2, sdt, cdt, sdp, cdp, limb; };
template <class Sh>
__device__ __forceinline__ Geo geo_of(const Sh& X, const LaunchArgs& L, int kb) {
    if (!Sh::BATCH) { Geo G; G.d0 = L.det[0]; G.d1 = L.det[1]; G.d2 = L.det[2]; G.sdt = L.sin_dt; G.cdt = L.cos_dt; G.sdp = L.sin_dp; G.cdp = L.cos_dp; G.limb = 0.0; return G; }
    const double4 a = *reinterpret_cast<const double4*>(X.geo + GEO * kb), b = *reinterpret_cast<const double4*>(X.geo + GEO * kb + 4);
    Geo G; G.d0 = a.x; G.d1 = a.y; G.d2 = a.z; G.sdt = a.w; G.cdt = b.x; G.sdp = b.y; G.cdp = b.z; G.limb = b.w;
    return G;
}

// cell_depth / wavelength index of the photon's launch (they differ between the launches of a wavelength batch only)
template <class Sh>
__device__ __forceinline__ int depth_of(const Sh& X, const KernelArgs& A, int kb) {
    return (Sh::BATCH && A.L.wl_batch) ? (int)X.geo[GEO * kb + 10] : A.T.cell_depth;
}
template <class Sh>
__device__ __forceinline__ int wl_of(const Sh& X, const KernelArgs& A, int kb) {
    return (Sh::BATCH && A.L.wl_batch) ? (int)X.geo[GEO * kb + 11] : 0;
}

struct Cnt {
    unsigned long long n_cf;
    unsigned n_emit, n_sc, n_peel, n_surf, n_err, n_draw;
};

// ---------------------------------------------------------------------------------------------------
// events (called warp-converged: every lane of the warp handles one slot of the same list; !valid lanes idle)
// ---------------------------------------------------------------------------------------------------

// EMIT, planet source: emit_photon :1117-1266 + the thermal weight and the start of peel_thermal :599-621.
// The cell comes from a binary search on the emissivity CDF (the reference scans it linearly, :1132-1155).
template <class Sh>
__device__ __forceinline__ bool ev_emit_thermal(const Sh& X, const KernelArgs& A, int s, unsigned long long id, int kb, Cnt& C, RaySpec& rs) {
    const DevTables& T = A.T;
    const LaunchArgs& L = A.L;
    const Geo G = geo_of(X, L, kb);
    double xr[5], xq;
    unsigned nd = 0;
    draws(X, A, s, id, nd, 4, xr); nd += 4;
    const int depth = depth_of(X, A, kb);
    const size_t wl_off = (size_t)wl_of(X, A, kb) * T.cells;      // tables of the photon's wavelength (wavelength batches)
    const double* ecdf = T.emis_cdf + wl_off;
    const int ncdf = (T.nr - depth) * T.nt * T.np;
    const double samp = xr[0] * __ldg(ecdf + ncdf - 1);
    int lo = -1, hi = ncdf - 1;      // first p with cdf[p] >= samp
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (__ldg(ecdf + mid) >= samp) hi = mid; else lo = mid; }
    const int c2 = hi % T.np, c1 = (hi / T.np) % T.nt, c0 = depth + hi / (T.np * T.nt);
    double rsam = xr[1] * (X.r[c0 + 1] - X.r[c0]); rsam = X.r[c0] + rsam;
    const double tc0 = __ldg(T.tcos + c1), tc1 = __ldg(T.tcos + c1 + 1);
    double ct = xr[2] * (tc1 - tc0); ct = tc0 + ct;
    const double st = sqrt(1.0 - ct * ct);
    double phs;
    if (T.np == 1) phs = 2.0 * PI * xr[3];
    else if (c2 < T.np - 1) { phs = xr[3] * (X.pf[c2 + 1] - X.pf[c2]); phs = X.pf[c2] + phs; }
    else { phs = xr[3] * (2.0 * PI - X.pf[c2]); phs = X.pf[c2] + phs; }
    const double cp = cos(phs);
    double sp = sqrt(1.0 - cp * cp);
    if (phs > PI) sp = -sp;
    double px = rsam * st * cp, py = rsam * st * sp, pz = rsam * ct;
    px = T.ox * px; py = T.oy * py; pz = T.oz * pz;
    double dx, dy, dz, bias_weight = 1.0;
    int e = 0;
    if (L.photon_emission == 1) {
        draws(X, A, s, id, nd, 2, xr); nd += 2;
        const double al = 2.0 * xr[0] - 1.0, be = 2.0 * PI * xr[1];
        const double cb = cos(be);
        double sb = sqrt(1.0 - cb * cb);
        if (be > PI) sb = -sb;
        dx = sqrt(1.0 - al * al) * cb; dy = sqrt(1.0 - al * al) * sb; dz = al;
    } else {
        draws(X, A, s, id, nd, 2, xr); nd += 2;
        const double yb = (1.0 + L.photon_bias) * tan(PI * xr[0] / 2.0) / sqrt(1.0 - L.photon_bias * L.photon_bias);
        const double ths = acos((1.0 - yb * yb) / (1.0 + yb * yb));
        const double be = 2.0 * PI * xr[1];
        double r0 = px / (T.ox * T.ox), r1 = py / (T.oy * T.oy), r2 = pz / (T.oz * T.oz);
        const double nrm = sqrt(r0 * r0 + r1 * r1 + r2 * r2);
        r0 = r0 / nrm; r1 = r1 / nrm; r2 = r2 / nrm;
        e = direction_cosine(cos(PI - ths), be, r0, r1, r2, dx, dy, dz);
        bias_weight = (PI * sin(ths) * (1.0 + L.photon_bias * cos(ths))) / (2.0 * sqrt(1.0 - L.photon_bias * L.photon_bias));
    }
    (void)xq;
    X.I(I_ND, s) = (int)nd;
    X.I(I_HCELL, s) = pack_cell(c0, c1, c2);
    X.D(F_PX, s) = px; X.D(F_PY, s) = py; X.D(F_PZ, s) = pz;
    X.D(F_S1, s) = 0.0; X.D(F_S2, s) = 0.0; X.D(F_S3, s) = 0.0; X.D(F_TAU, s) = 0.0;
    if (e) { err_count(A, e); ++C.n_err; X.D(F_S0, s) = 1.0; X.I(I_INFO, s) = K_DEAD; return true; }
    if (fabs(dz) >= 1.0) err_count(A, 54);
    X.D(F_DX, s) = dx; X.D(F_DY, s) = dy; X.D(F_DZ, s) = dz;
    const double S0 = 1.0 * bias_weight / __ldg(T.cell_weight + wl_off + c0 + T.nr * (c1 + T.nt * c2));
    X.D(F_S0, s) = S0;
    atomicAdd(A.O.flux + 2 * kb, S0);
    // peel_thermal :4519-4598: walk to the detector, deposit e^-tau / 4 pi x I (weight applied by DEP)
    ++C.n_peel;
    X.D(F_W0, s) = S0;
    {
        const double x_im = py * G.cdp - px * G.sdp;
        const double y_im = pz * G.sdt - py * G.cdt * G.sdp - px * G.cdt * G.cdp;
        const int ix = (int)(L.nx * (x_im + L.x_max) / (2.0 * L.x_max)) + 1;
        const int iy = (int)(L.ny * (y_im + L.y_max) / (2.0 * L.y_max)) + 1;
        X.I(I_PIX, s) = (ix < 1 || ix > L.nx || iy < 1 || iy > L.ny) ? -2 : (ix - 1) + L.nx * (iy - 1) + kb * 10 * L.nx * L.ny;
    }
    rs.set(px, py, pz, G.d0, G.d1, G.d2, c0, c1, c2, -1, K_PEEL, CUDART_INF, 0.0, PK_THERMAL);
    return true;
}

// SURF (reflecting surface only): the transport walk reached the surface :755-774 -> absorbed, or Lambert reflection
// (lambertian :1369-1402) followed by the start of peel_surface :4600-4650.  The optical depth of the walk is NOT
// resampled: the reflected photon goes on with the same tau and the running sum (:766-776).
template <class Sh>
__device__ __forceinline__ bool ev_surface(const Sh& X, const KernelArgs& A, bool valid, int s, Cnt& C, RaySpec& rs) {
    if (!valid) return false;
    const DevTables& T = A.T;
    const LaunchArgs& L = A.L;
    const unsigned long long id = (unsigned long long)(unsigned)X.I(I_IDLO, s) | ((unsigned long long)(unsigned)X.I(I_IDHI, s) << 32);
    unsigned nd = (unsigned)X.I(I_ND, s);
    double xr[5];
    draws(X, A, s, id, nd, 1, xr); ++nd;
    const double tw = X.D(F_T, s);
    const double wx = X.D(F_PX, s) + tw * X.D(F_DX, s), wy = X.D(F_PY, s) + tw * X.D(F_DY, s), wz = X.D(F_PZ, s) + tw * X.D(F_DZ, s);
    X.D(F_PX, s) = wx; X.D(F_PY, s) = wy; X.D(F_PZ, s) = wz;      // the photon now sits on the surface
    if (xr[0] > L.surface_albedo) { X.I(I_ND, s) = (int)nd; X.I(I_INFO, s) = K_DEAD; return true; }
    double s0 = wx / (T.ox * T.ox), s1 = wy / (T.oy * T.oy), s2 = wz / (T.oz * T.oz);
    const double nrm = sqrt(s0 * s0 + s1 * s1 + s2 * s2);
    s0 = s0 / nrm; s1 = s1 / nrm; s2 = s2 / nrm;
    draws(X, A, s, id, nd, 2, xr); nd += 2;
    X.I(I_ND, s) = (int)nd;
    const double al = sqrt(xr[0]), be = 2.0 * PI * xr[1];
    double e0, e1, e2;
    const int e = direction_cosine(al, be, s0, s1, s2, e0, e1, e2);
    if (e) { err_count(A, e); ++C.n_err; X.I(I_INFO, s) = K_DEAD; return true; }
    X.D(F_DX, s) = e0; X.D(F_DY, s) = e1; X.D(F_DZ, s) = e2;
    X.D(F_S1, s) = 0.0; X.D(F_S2, s) = 0.0; X.D(F_S3, s) = 0.0;      // fully depolarised :1397-1400
    const int cell = X.I(I_CELL, s);                                 // the cell above the surface face
    const int c0 = cell & 1023, c1 = (cell >> 10) & 1023, c2 = (cell >> 20) & 1023;
    X.I(I_HCELL, s) = cell;
    const double tau = X.D(F_TAU, s), acc = X.D(F_ACC, s);
    const int kb = Sh::BATCH ? X.I(I_BATCH, s) : 0;
    const double cos_angle = Sh::BATCH ? surface_cos_angle2(T, wx, wy, wz, X.geo[GEO * kb + 8], X.geo[GEO * kb + 9])
                                       : surface_cos_angle2(T, wx, wy, wz, L.det_sph_theta, L.det_sph_phi);
    if (cos_angle > 0.0) {
        ++C.n_peel;
        X.D(F_W0, s) = X.D(F_S0, s); X.D(F_W1, s) = acc; X.D(F_W2, s) = cos_angle;
        const Geo G = geo_of(X, L, kb);
        const double x_im = wy * G.cdp - wx * G.sdp;
        const double y_im = wz * G.sdt - wy * G.cdt * G.sdp - wx * G.cdt * G.cdp;
        const int ix = (int)(L.nx * (x_im + L.x_max) / (2.0 * L.x_max)) + 1;
        const int iy = (int)(L.ny * (y_im + L.y_max) / (2.0 * L.y_max)) + 1;
        X.I(I_PIX, s) = (ix < 1 || ix > L.nx || iy < 1 || iy > L.ny) ? -2 : (ix - 1) + L.nx * (iy - 1) + kb * 10 * L.nx * L.ny;
        rs.set(wx, wy, wz, G.d0, G.d1, G.d2, c0, c1, c2, depth_of(X, A, kb), K_PEEL, CUDART_INF, 0.0, PK_SURFACE);
    } else
        rs.set(wx, wy, wz, e0, e1, e2, c0, c1, c2, depth_of(X, A, kb), K_WALK, tau, acc);
    return true;
}

// EMIT: star emission (emit_photon :1008-1115 + initial_cell).  Returns true if a ray was set up.
template <class Sh>
__device__ __forceinline__ bool ev_emit(const Sh& X, const KernelArgs& A, bool valid, int s, Cnt& C, RaySpec& rs) {
    const DevTables& T = A.T;
    const LaunchArgs& L = A.L;
    const int lane = threadIdx.x & 31;
    const unsigned vm = __ballot_sync(FULL, valid);
    unsigned long long base = 0;
    const int leader = __ffs(vm) - 1;
    if (lane == leader) base = atomicAdd(A.O.counter, (unsigned long long)__popc(vm));
    base = __shfl_sync(FULL, base, leader < 0 ? 0 : leader);
    if (!valid) return false;
    C.n_draw += (unsigned)X.I(I_ND, s);          // draws of the photon that lived in this slot before
    if (Sh::TRACE && X.I(I_ND, s) > 0) {         // ... and its walk record
        const unsigned long long pid = (unsigned long long)(unsigned)X.I(I_IDLO, s) | ((unsigned long long)(unsigned)X.I(I_IDHI, s) << 32);
        const size_t kk = (size_t)(pid - L.id_base);
        A.R.seq_len[kk] = X.I(I_TLEN, s);
        A.R.seq_hash[kk] = (unsigned long long)(unsigned)X.I(I_THLO, s) | ((unsigned long long)(unsigned)X.I(I_THHI, s) << 32);
        if (A.R.fstate) {
            const int pinfo = X.I(I_INFO, s);
            const int pout = (pinfo >> 8) & 15;
            const double tw = ((pinfo & 3) == K_WALK && (pout == O_EXIT || pout == O_SURF || pout == O_ERR)) ? X.D(F_T, s) : 0.0;
            double* f = A.R.fstate + kk * 8;
            f[0] = X.D(F_PX, s) + tw * X.D(F_DX, s); f[1] = X.D(F_PY, s) + tw * X.D(F_DY, s); f[2] = X.D(F_PZ, s) + tw * X.D(F_DZ, s);
            f[3] = X.D(F_S0, s); f[4] = X.D(F_S1, s); f[5] = X.D(F_S2, s); f[6] = X.D(F_S3, s); f[7] = (double)X.I(I_TNSC, s);
        }
    }
    if (Sh::GEN && L.photon_source == 2 && X.I(I_ND, s) > 0) {   // emergent flux of a thermal photon that left the grid (:780, :953)
        const int pinfo = X.I(I_INFO, s);
        if ((pinfo & 3) == K_WALK && ((pinfo >> 8) & 15) == O_EXIT) atomicAdd(A.O.flux + 1 + (Sh::BATCH ? 2 * X.I(I_BATCH, s) : 0), X.D(F_S0, s));
    }
    X.I(I_ND, s) = 0;
    const unsigned long long k = base + (unsigned long long)__popc(vm & ((1u << lane) - 1u));
    if (k >= L.n_photons) { atomicAdd(X.misc, 1); return false; }
    ++C.n_emit;
    const unsigned long long id = L.id_base + k;
    X.I(I_IDLO, s) = (int)(unsigned)id; X.I(I_IDHI, s) = (int)(unsigned)(id >> 32);
    int kb = 0;                              // launch of a batch this work item belongs to
    if (Sh::BATCH) { kb = (int)((id - L.batch_base) / L.per_launch); X.I(I_BATCH, s) = kb; }
    if (Sh::TRACE) {
        X.I(I_TLEN, s) = 0; X.I(I_TNSC, s) = 0; X.I(I_FLAG, s) = 0;
        X.I(I_THLO, s) = (int)(unsigned)1469598103934665603ull; X.I(I_THHI, s) = (int)(unsigned)(1469598103934665603ull >> 32);
    }
    unsigned nd = 0;
    if (Sh::GEN && L.photon_source == 2) return ev_emit_thermal(X, A, s, id, kb, C, rs);
    double xi, r_disk;
    if (Sh::BATCH ? (X.geo[GEO * kb + 7] != 0.0) : (L.limb_emission != 0)) {
        for (;;) { draws(X, A, s, id, nd, 1, &xi); ++nd; r_disk = sqrt(xi); if (r_disk > 0.9 || (Sh::TRACE && (X.I(I_FLAG, s) & 1))) break; }
    } else { draws(X, A, s, id, nd, 1, &xi); ++nd; r_disk = sqrt(xi); }
    draws(X, A, s, id, nd, 1, &xi); ++nd;
    const double phi_disk = 2.0 * PI * xi;
    const double R = X.r[T.nr];
    double sphi, cphi;
    sincos(phi_disk, &sphi, &cphi);
    const double d1 = R * r_disk * sphi, d2 = R * r_disk * cphi;
    double dx = -1.0, dy = 0.0, dz = 0.0;
    double px = sqrt(R * R - d1 * d1 - d2 * d2), py = d1, pz = d2;
    if (L.stellar_direction) {  // :1080-1111
        const double tx = px * L.rot_y_cos + pz * L.rot_y_sin, ty = py, tz = px * (-L.rot_y_sin) + pz * L.rot_y_cos;
        px = tx * L.rot_z_cos + ty * (-L.rot_z_sin); py = tx * L.rot_z_sin + ty * L.rot_z_cos; pz = tz;
        dx = L.star_dir[0]; dy = L.star_dir[1]; dz = L.star_dir[2];
    }
    // initial_cell :2605-2669
    int c0 = T.nr - 1, c1 = 0, c2 = 0;
    {
        const double r = sqrt(px * px + py * py + pz * pz);
        const double theta = acos(pz / r);
        double phi = atan2(py, px);
        if (phi < 0.0) phi = phi + 2.0 * PI;
        for (int j = 0; j < T.nt; ++j) if (theta > X.tf[j] && theta < X.tf[j + 1]) { c1 = j; break; }
        for (int j = 0; j < T.np; ++j) {
            const double hi = (j < T.np - 1) ? X.pf[j + 1] : 2.0 * PI;
            if (phi > X.pf[j] && phi < hi) { c2 = j; break; }
        }
    }
    X.D(F_PX, s) = px; X.D(F_PY, s) = py; X.D(F_PZ, s) = pz; X.D(F_DX, s) = dx; X.D(F_DY, s) = dy; X.D(F_DZ, s) = dz;
    X.D(F_S0, s) = 1.0; X.D(F_S1, s) = 0.0; X.D(F_S2, s) = 0.0; X.D(F_S3, s) = 0.0; X.D(F_TAU, s) = 0.0;
    X.I(I_ND, s) = (int)nd;
    X.I(I_HCELL, s) = pack_cell(c0, c1, c2) | (1 << 30);      // bit 30: the photon sits on the outer radial face
    rs.set(px, py, pz, dx, dy, dz, c0, c1, c2, T.nr, K_PRE, CUDART_INF);
    return true;
}

// PRE: the tau pre-pass ended -> first optical depth :660-685
template <class Sh>
__device__ __forceinline__ bool ev_pre(const Sh& X, const KernelArgs& A, bool valid, int s, Cnt& C, RaySpec& rs) {
    if (!valid) return false;
    const LaunchArgs& L = A.L;
    const int out = (X.I(I_INFO, s) >> 8) & 15;
    const double tacc = X.D(F_ACC, s);
    const bool hit_surface = (out == O_SURF);
    if (tacc < 1.e-6 && !hit_surface) { X.I(I_INFO, s) = K_DEAD; return true; }
    const unsigned long long id = (unsigned long long)(unsigned)X.I(I_IDLO, s) | ((unsigned long long)(unsigned)X.I(I_IDHI, s) << 32);
    const unsigned nd = (unsigned)X.I(I_ND, s);
    double xi;
    draws(X, A, s, id, nd, 1, &xi);
    X.I(I_ND, s) = (int)(nd + 1u);
    double arg = 1.0 - xi;
    if (!(tacc < 1.e-6) && tacc < 50.0) {
        const double f = 1.0 - fm_exp_neg(tacc);
        arg = 1.0 - xi * f;
        X.D(F_S0, s) *= f; X.D(F_S1, s) *= f; X.D(F_S2, s) *= f; X.D(F_S3, s) *= f;
    }
    const double tau = -fm_log(arg);
    X.D(F_TAU, s) = tau;
    const int hc = X.I(I_HCELL, s);
    rs.set(X.D(F_PX, s), X.D(F_PY, s), X.D(F_PZ, s), X.D(F_DX, s), X.D(F_DY, s), X.D(F_DZ, s),
              hc & 1023, (hc >> 10) & 1023, (hc >> 20) & 1023, (hc >> 30) ? A.T.nr : -1, K_WALK, tau);
    return true;
}

// H: the transport walk reached its optical depth.  Survival :791-813, peel-off weight and pixel (peel_photon
// :4763-4951; the e^-tau factor is applied by DEP once the peel ray has been walked), scattering :819-845.
// (defined after the marcher, below)
template <class Sh>
__device__ __forceinline__ int march_inline(const Sh& X, const KernelArgs& A, double px, double py, double pz, double n0, double n1, double n2,
                                            int cell, int slot, double lim, double& acc_out, unsigned& n_step);
__device__ __forceinline__ void deposit_scatter_warp(const KernelArgs& A, bool dep, int pix, const double v[8]);

#ifndef E2_INLINE_PEEL
#define E2_INLINE_PEEL 1   // 1: the peel-off walk of a scattering is marched inside the interaction event (no peel ray, no DEP event)
#endif
#ifndef E2_KEEP_W
#define E2_KEEP_W 0
#endif
struct HOut { double px, py, pz, dx, dy, dz, tau, W[4]; int pix, cell, kb; };
// returns 0: idle lane, 1: the photon died (survival test), 2: scattered -- `o` holds the peel-off weights and the new state
template <class Sh>
__device__ __forceinline__ int interact_core(const Sh& X, const KernelArgs& A, bool valid, int s, Cnt& C, HOut& o) {
    if (!valid) return 0;
    const DevTables& T = A.T;
    const LaunchArgs& L = A.L;
    const int cell = X.I(I_CELL, s);
    const int c0 = cell & 1023, c1 = (cell >> 10) & 1023, c2 = (cell >> 20) & 1023;
    const int ci = c0 + T.nr * (c1 + T.nt * c2);
    // the cold record in 32-byte pieces: [px py pz dx] [dy dz S0 S1] [S2 S3 tau W0] [W1 W2 W3 hcell|pix] [tlen|tnsc ...]; photon id and draw counter are hot ints
    double* rec = X.cold + (size_t)s * REC;
    double hx, hy, hz, dx, dy, dz, S[4], tau0, w0_;
    ldg256(rec, hx, hy, hz, dx);
    ldg256(rec + 4, dy, dz, S[0], S[1]);
    ldg256(rec + 8, S[2], S[3], tau0, w0_);
    (void)w0_;
    // per-cell record {cell_opacity, cell_albedo, unique-matrix index}: one 256-bit read instead of three scattered ones
    double kap_c, alb_c, u_bits, pad_c;
    const int kb = Sh::BATCH ? X.I(I_BATCH, s) : 0;
    ldg256_nc(T.cellrec + (size_t)4 * ((size_t)ci + (size_t)wl_of(X, A, kb) * T.cells), kap_c, alb_c, u_bits, pad_c);
    (void)pad_c;
    const int u = (int)__double_as_longlong(u_bits);
    // the marcher stopped after adding the crossing that overshoots tau: step back by the overshoot (:705-720)
    const double tpos = X.D(F_T, s) - fdiv(X.D(F_ACC, s) - tau0, kap_c);
    const double px = hx + tpos * dx, py = hy + tpos * dy, pz = hz + tpos * dz;
    const Geo G = geo_of(X, L, kb);
    double mu = dx * G.d0 + dy * G.d1 + dz * G.d2;
    if (mu >= 1.0) mu = 1.0 - 1.e-10; else if (mu <= -1.0) mu = -1.0 + 1.e-10;
    const double peel_deg = fm_acos(mu) * (180.0 / PI);
    // photon id and position in its random stream: hot fields (shared memory)
    const unsigned long long id = (unsigned long long)(unsigned)X.I(I_IDLO, s) | ((unsigned long long)(unsigned)X.I(I_IDHI, s) << 32);
    unsigned nd = (unsigned)X.I(I_ND, s);
    bool alive = L.photon_scattering != 0;
    if (Sh::TRACE && (X.I(I_FLAG, s) & 1)) alive = false;       // injected stream used up (test hook only)
    double xr[5];
    if (alive) draws(X, A, s, id, nd, 5, xr);
    if (Sh::TRACE && alive) X.I(I_FLAG, s) &= ~1;                // (only the draws actually consumed below may exhaust it)
    if (alive) { const double xi = xr[0]; ++nd; if (Sh::TRACE && (int)nd > A.R.max_draws) X.I(I_FLAG, s) |= 1; if (xi < L.fstop) alive = false; }
    if (alive) {
        const double alb = alb_c;
        if (alb < 1.0 && alb > 0.0) { const double g = fdiv(alb, 1.0 - L.fstop); S[0] *= g; S[1] *= g; S[2] *= g; S[3] *= g; }
        if (S[0] <= L.photon_minimum) alive = false;
    }
    if (!alive) {
        X.I(I_ND, s) = (int)nd; X.I(I_INFO, s) = K_DEAD;
        if (Sh::TRACE) {   // final state of the photon for the trace hook
            X.D(F_PX, s) = px; X.D(F_PY, s) = py; X.D(F_PZ, s) = pz;
            X.D(F_S0, s) = S[0]; X.D(F_S1, s) = S[1]; X.D(F_S2, s) = S[2]; X.D(F_S3, s) = S[3];
        }
        return 1;
    }
    // ---- peel-off towards the detector: Stokes vector scattered into det, pixel
    ++C.n_peel;
    int pix = -1;
    double W[4] = {0.0, 0.0, 0.0, 0.0};
    {
        if (!(fabs(dz) < 1.0)) err_count(A, 45);
        else {
            const double smu = fsqrt(1.0 - mu * mu);
            double nc = fdiv(G.d2 - dz * mu, smu * fsqrt(1.0 - dz * dz));
            if (!(nc == nc)) err_count(A, 44);
            else {
                nc = fmin(fmax(nc, -1.0), 1.0);
                const double cr = dy * G.d0 - dx * G.d1;
                const bool flip = (cr > 0.0) || (cr == 0.0 && dx * G.d0 + dy * G.d1 > 0.0);
                const double c2a = 2.0 * nc * nc - 1.0;
                double s2a = 2.0 * nc * fsqrt(fmax(1.0 - nc * nc, 0.0));
                if (flip) s2a = -s2a;
                const double nc2 = fdiv(dz - G.d2 * mu, smu * fsqrt(1.0 - G.d2 * G.d2));
                int soft = 0;
                const int e = (fabs(G.d2) < 1.0) ? polrot_deg_f(T, u, peel_deg, c2a, s2a, flip, nc2, S, W, true, soft) : 16;
                if (e) err_count(A, e);
                else if (!(W[0] > 0.0 && W[0] < 1.e100)) err_count(A, 53);
                else {
                    const double x_im = py * G.cdp - px * G.sdp;
                    const double y_im = pz * G.sdt - py * G.cdt * G.sdp - px * G.cdt * G.cdp;
                    const int ix = (int)fdiv(L.nx * (x_im + L.x_max), 2.0 * L.x_max) + 1;
                    const int iy = (int)fdiv(L.ny * (y_im + L.y_max), 2.0 * L.y_max) + 1;
                    if (ix < 1 || ix > L.nx || iy < 1 || iy > L.ny) err_count(A, 60);
                    else pix = (ix - 1) + L.nx * (iy - 1) + kb * 10 * L.nx * L.ny;
                }
            }
        }
    }
    // ---- scattering: new direction and Stokes vector (ev_scatter, fast branch)
    ++C.n_sc;
    double tau = -1.0;                      // < 0: the photon dies after its peel-off has been deposited
    {
        FastAngles g;
        int e = sample_angles_f(X, A, xr[1], xr[2], xr[3], S, ci, g, u);
        nd += (e == 6) ? 2u : 3u;
        if (Sh::TRACE) { X.I(I_TNSC, s) += 1; if ((int)nd > A.R.max_draws) X.I(I_FLAG, s) |= 1; }
        double e0 = 0, e1 = 0, e2 = 0;
        if (!e) {
            const double cto = fdiv(dz, fsqrt(dx * dx + dy * dy + dz * dz));
            const double sto = fsqrt(1.0 - cto * cto);
            const double ctn = cto * g.alpha + sto * g.sT * g.cb;
            const double stn = fsqrt(1.0 - ctn * ctn);
            double nc = fdiv(g.alpha - ctn * cto, stn * sto);
            if (!(nc == nc)) e = 20;
            else {
                if (nc >= 1.0) nc = 1.0 - 1.e-10; else if (nc <= -1.0) nc = -1.0 + 1.e-10;
                const double sD = fsqrt(1.0 - nc * nc) * (g.flip ? -1.0 : 1.0);
                const double rho = fsqrt(dx * dx + dy * dy);
                const double irho = frcp(rho);
                const double cph = rho > 0.0 ? dx * irho : 1.0, sph = rho > 0.0 ? dy * irho : 0.0;
                e0 = stn * (cph * nc - sph * sD); e1 = stn * (sph * nc + cph * sD); e2 = ctn;
                if (!(fabs(e2) < 1.0)) e = 16;
            }
        }
        if (!e) {
            double Sn[4];
            const double nc2 = fdiv(dz - e2 * g.alpha, g.sT * fsqrt(1.0 - e2 * e2));
            int soft = 0;
            e = polrot_deg_f(T, u, g.deg, g.cb * g.cb - g.sb * g.sb, 2.0 * g.sb * g.cb, g.flip, nc2, S, Sn, false, soft);
            if (soft) err_count(A, soft);
            if (!e) { S[0] = Sn[0]; S[1] = Sn[1]; S[2] = Sn[2]; S[3] = Sn[3]; dx = e0; dy = e1; dz = e2; }
        }
        if (e) { err_count(A, e); ++C.n_err; }
        else { const double xi = xr[4]; ++nd; if (Sh::TRACE && (int)nd > A.R.max_draws) X.I(I_FLAG, s) |= 1; tau = -fm_log(1.0 - xi); }
    }
    stg256(rec, px, py, pz, dx);
    stg256(rec + 4, dy, dz, S[0], S[1]);
    stg256(rec + 8, S[2], S[3], tau, W[0]);
    // (the peel-off weights and the pixel are for a DEP event: with the walk inside this event nobody reads them)
    if (!E2_INLINE_PEEL || E2_KEEP_W) stg256(rec + 12, W[1], W[2], W[3], __longlong_as_double((long long)((unsigned long long)(unsigned)cell | ((unsigned long long)(unsigned)pix << 32))));
#ifdef E2_DEBUG
    if (!(W[0] < 1.e10) || !(S[0] < 1.e10)) printf("E2 interact: slot %d cell %d %d %d W %g %g %g %g S %g tpos %g tau %g acc %g t %g\n", s, c0, c1, c2, W[0], W[1], W[2], W[3], S[0], tpos, tau0, X.D(F_ACC, s), X.D(F_T, s));
#endif
    X.I(I_ND, s) = (int)nd;
    o.px = px; o.py = py; o.pz = pz; o.dx = dx; o.dy = dy; o.dz = dz; o.tau = tau;
    o.W[0] = W[0]; o.W[1] = W[1]; o.W[2] = W[2]; o.W[3] = W[3]; o.pix = pix; o.cell = cell; o.kb = kb;
    return 2;
}

template <class Sh>
__device__ __forceinline__ bool ev_interact(const Sh& X, const KernelArgs& A, bool valid, int s, Cnt& C, RaySpec& rs) {
    HOut o;
    const int st = interact_core(X, A, valid, s, C, o);
    const LaunchArgs& L = A.L;
    if (!E2_INLINE_PEEL) {
        if (st == 2) {
            const Geo G = geo_of(X, L, o.kb);
            rs.set(o.px, o.py, o.pz, G.d0, G.d1, G.d2, o.cell & 1023, (o.cell >> 10) & 1023, (o.cell >> 20) & 1023, -1, K_PEEL, CUDART_INF);
        }
        return st != 0;
    }
    // ---- the walk to the detector in the lane (:4739-4761), e^-tau and the deposit (:4955-4972), then the transport ray of the
    // scattered photon: what the peel ray and the DEP event did, without their trips through the lists and the photon record
    bool dep = false, dead = false;
    double v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (st == 2) {
        const Geo G = geo_of(X, L, o.kb);
        unsigned n_step = 0;
        double acc = 0.0;
        const int out = march_inline(X, A, o.px, o.py, o.pz, G.d0, G.d1, G.d2, o.cell, s, CUDART_INF, acc, n_step);
        C.n_cf += n_step;
        if (out == O_ERR) { err_count(A, 31); ++C.n_err; err_count(A, 43); dead = true; }      // (as Marcher::finish: the photon is dropped)
        dep = (out == O_EXIT && acc < 50.0 && o.pix >= 0);
        if (dep) {
            const double w = fm_exp_neg(acc);
            v[0] = w * o.W[0]; v[1] = -(w * o.W[1]); v[2] = w * o.W[2]; v[3] = w * o.W[3];
            v[4] = v[0] * v[0]; v[5] = v[1] * v[1]; v[6] = v[2] * v[2]; v[7] = v[3] * v[3];
            if (Sh::TRACE) {   // the reference's deposit point in the walk record: (100, ix, iy, 0, 0)
                const unsigned long long pid = (unsigned long long)(unsigned)X.I(I_IDLO, s) | ((unsigned long long)(unsigned)X.I(I_IDHI, s) << 32);
                int tl = X.I(I_TLEN, s);
                unsigned long long th = (unsigned long long)(unsigned)X.I(I_THLO, s) | ((unsigned long long)(unsigned)X.I(I_THHI, s) << 32);
                trace_tuple(A, pid, tl, th, 100, o.pix % L.nx + 1, o.pix / L.nx + 1, 0, 0);
                X.I(I_TLEN, s) = tl; X.I(I_THLO, s) = (int)(unsigned)th; X.I(I_THHI, s) = (int)(unsigned)(th >> 32);
            }
        }
    }
    deposit_scatter_warp(A, dep, o.pix, v);
    if (st == 2) {
        if (dead || o.tau < 0.0) X.I(I_INFO, s) = K_DEAD;
        else rs.set(o.px, o.py, o.pz, o.dx, o.dy, o.dz, o.cell & 1023, (o.cell >> 10) & 1023, (o.cell >> 20) & 1023, -1, K_WALK, o.tau);
    }
    return st != 0;
}

// DEP: the peel ray ended -> e^-tau, detector deposit :4955-4972; then the transport ray of the scattered photon
template <class Sh>
__device__ __forceinline__ bool ev_deposit(const Sh& X, const KernelArgs& A, bool valid, int s, Cnt& C, RaySpec& rs) {
    const LaunchArgs& L = A.L;
    const int out = valid ? ((X.I(I_INFO, s) >> 8) & 15) : O_NONE;
    const double tacc = valid ? X.D(F_ACC, s) : 0.0;
    // the cold record in 32-byte pieces (see ev_interact); an idle lane reads slot 0's record and ignores it
    const double* rec = X.cold + (size_t)s * REC;
    double hx, hy, hz, dx, dy, dz, s0_, s1_, s2_, s3_, tau, W0, W1, W2, W3, ipk;
    ldg256(rec, hx, hy, hz, dx);
    ldg256(rec + 4, dy, dz, s0_, s1_);
    ldg256(rec + 8, s2_, s3_, tau, W0);
    ldg256(rec + 12, W1, W2, W3, ipk);
    (void)s0_; (void)s1_; (void)s2_; (void)s3_;
    const unsigned long long w15 = (unsigned long long)__double_as_longlong(ipk);
    const int hc = (int)(unsigned)w15;
    const int pix = valid ? (int)(unsigned)(w15 >> 32) : -1;
    // Deposit.  When every depositing lane of the batch hits the same pixel (1x1 "photometry" detectors: phase
    // curves, spectra) the ten sums are reduced across the warp first, so the L2 sees one atomic per plane and
    // batch instead of 32 serialised ones on the same address.
    const int pk = (Sh::GEN && valid) ? ((X.I(I_INFO, s) >> PK_SHIFT) & 3) : PK_SCATTER;
    bool dep = valid && out == O_EXIT && tacc < 50.0 && pix >= 0;
    double w_i = 0.0;            // surface / thermal peel: the only deposited quantity
    if (Sh::GEN && pk != PK_SCATTER) {
        dep = false;
        if (valid && out == O_EXIT && tacc < 50.0) {
            const double w = (pk == PK_THERMAL) ? fm_exp_neg(tacc) / (4.0 * PI) : fm_exp_neg(tacc) * W2 / PI;
            w_i = w * W0;
            if (!(w_i > 0.0 && w_i < 1.e100)) err_count(A, pk == PK_THERMAL ? 51 : 52);
            else if (pix == -2) err_count(A, 60);
            else dep = true;
        }
    }
    if (Sh::TRACE && dep) {   // the reference's deposit point in the walk record: (100, ix, iy, 0, 0)
        const unsigned long long pid = (unsigned long long)(unsigned)X.I(I_IDLO, s) | ((unsigned long long)(unsigned)X.I(I_IDHI, s) << 32);
        int tl = X.I(I_TLEN, s);
        unsigned long long th = (unsigned long long)(unsigned)X.I(I_THLO, s) | ((unsigned long long)(unsigned)X.I(I_THHI, s) << 32);
        trace_tuple(A, pid, tl, th, 100, pix % L.nx + 1, pix / L.nx + 1, 0, 0);
        X.I(I_TLEN, s) = tl; X.I(I_THLO, s) = (int)(unsigned)th; X.I(I_THHI, s) = (int)(unsigned)(th >> 32);
    }
    const unsigned dm = __ballot_sync(FULL, dep);
    if (dm) {
        double v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        if (dep) {
            const double w = fm_exp_neg(tacc);
            v[0] = w * W0; v[1] = -(w * W1); v[2] = w * W2; v[3] = w * W3;
            v[4] = v[0] * v[0]; v[5] = v[1] * v[1]; v[6] = v[2] * v[2]; v[7] = v[3] * v[3];
        }
        const size_t npx = (size_t)L.nx * L.ny;
        // always the global image here: a pointer that may be shared OR global would turn the reductions below into generic-address
        // atomics.  (Block-private images are used by the multi-detector walks only, ev_fan; measured: profiles/r02_ab_variants.txt.)
        double* const detbase = A.O.det;
        if (Sh::GEN && dep && pk != PK_SCATTER) {   // :4583-4585 / :4691-4693: Stokes I only
            double* d = detbase + pix;
            atomicAdd(d, w_i); atomicAdd(d + 4 * npx, w_i * w_i); atomicAdd(d + 8 * npx, 1.0);
        }
        // Lanes that hit the same pixel are summed in the warp first (up to four pixel groups per batch: a batched
        // launch has the photons of two launches in flight around each launch boundary); what is left goes lane by lane.
        unsigned rem = __ballot_sync(FULL, dep && pk == PK_SCATTER), left = 0u;
        const int lane = threadIdx.x & 31;
#pragma unroll 1
        for (int it = 0; it < 4 && rem; ++it) {
            const int pixg = __shfl_sync(FULL, pix, __ffs(rem) - 1);
            const unsigned grp = __ballot_sync(FULL, ((rem >> lane) & 1u) && pix == pixg);
            rem &= ~grp;
            if (__popc(grp) <= 2) { left |= grp; continue; }
            const bool in = (grp >> lane) & 1u;
            double x = 0.0;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                double r = in ? v[k] : 0.0;
                for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(FULL, r, o);
                if (lane == k) x = r;
            }
            double* d = detbase + pixg;
            if (lane < 8) atomicAdd(d + (size_t)lane * npx, x);
            else if (lane < 10) atomicAdd(d + (size_t)lane * npx, (double)__popc(grp));
        }
        left |= rem;
        if ((left >> lane) & 1u) {
            double* d = detbase + pix;
#pragma unroll
            for (int k = 0; k < 8; ++k) atomicAdd(d + (size_t)k * npx, v[k]);
            atomicAdd(d + 8 * npx, 1.0); atomicAdd(d + 9 * npx, 1.0);
        }
    }
    if (!valid) return false;
    if (Sh::GEN && pk == PK_THERMAL) {          // after peel_thermal the photon starts its tau pre-pass (:621-656)
        rs.set(hx, hy, hz, dx, dy, dz,
                  hc & 1023, (hc >> 10) & 1023, (hc >> 20) & 1023, -1, K_PRE, CUDART_INF);
        return true;
    }
    if (Sh::GEN && pk == PK_SURFACE) {          // the reflected photon goes on with its old tau and running sum (:766-776)
        rs.set(hx, hy, hz, dx, dy, dz,
                  hc & 1023, (hc >> 10) & 1023, (hc >> 20) & 1023, depth_of(X, A, Sh::BATCH ? X.I(I_BATCH, s) : 0), K_WALK, tau, W1);
        return true;
    }
    if (tau < 0.0) { X.I(I_INFO, s) = K_DEAD; return true; }
    rs.set(hx, hy, hz, dx, dy, dz,
              hc & 1023, (hc >> 10) & 1023, (hc >> 20) & 1023, -1, K_WALK, tau);
    return true;
}

// RES: a polar or azimuthal face was crossed at parameter t: step the cell index and re-solve that axis
template <class Sh>
__device__ __forceinline__ bool ev_resolve(const Sh& X, const KernelArgs& A, bool valid, int s) {
    if (!valid) return false;
    const DevTables& T = A.T;
    const LaunchArgs& L = A.L;
    int info = X.I(I_INFO, s);
    const int out = (info >> 8) & 15, kind = info & 3;
    const int cell = X.I(I_CELL, s);
    int c0 = cell & 1023, c1 = (cell >> 10) & 1023, c2 = (cell >> 20) & 1023;
    const bool peel = (kind == K_PEEL);
    const double* rec = X.cold + (size_t)s * REC;
    double hx, hy, hz, rdx, rdy, rdz, s0_, s1_;
    ldg256(rec, hx, hy, hz, rdx);
    ldg256(rec + 4, rdy, rdz, s0_, s1_);
    (void)s1_;
    double n0 = rdx, n1 = rdy, n2 = rdz;
    if (peel) {
        if (Sh::BATCH) { const double* gd = X.geo + GEO * X.I(I_BATCH, s); n0 = gd[0]; n1 = gd[1]; n2 = gd[2]; }
        else { n0 = L.det[0]; n1 = L.det[1]; n2 = L.det[2]; }
    }
    RayK K;
    double hbn, D0, iq;
    ray_consts(T, hx, hy, hz, n0, n1, n2, K, hbn, D0, iq);
    const double t = X.D(F_T, s);
    info &= 0xff;
    if (out == O_REST) {
        if (Sh::GEN && L.flow_theta && kind == K_WALK)      // add_flow :5016-5047, polar crossings (:736-742)
            atomicAdd(A.O.flow4 + (size_t)4 * (c0 + T.nr * (c1 + T.nt * c2)) + ((info & B_TUPPER) ? 2 : 3), s0_);
        c1 += (info & B_TUPPER) ? 1 : -1;
        int upper;
        X.D(F_TT, s) = theta_next(X, T.nt, c1, t, K, upper);
        info = (info & ~B_TUPPER) | (upper ? B_TUPPER : 0);
    } else {
        if (info & B_PUP) c2 = (c2 + 1 == T.np) ? 0 : c2 + 1; else c2 = (c2 == 0) ? T.np - 1 : c2 - 1;
        int up;
        X.D(F_TP, s) = phi_next(X, T.np, c2, t, K, up);
    }
#ifdef E2_DEBUG
    if (c1 < 0 || c1 >= T.nt || c2 < 0 || c2 >= T.np) printf("E2 resolve: slot %d out %d cell %d %d %d info %x t %.17g\n", s, out, c0, c1, c2, info, t);
#endif
    X.I(I_CELL, s) = pack_cell(c0, c1, c2);
    X.I(I_INFO, s) = info;
    return true;
}

// ---------------------------------------------------------------------------------------------------
// block set-up and the marcher
// ---------------------------------------------------------------------------------------------------
template <int NT, int NP, bool TR, bool GN, bool BT, bool MD = false>
__device__ __forceinline__ void block_setup(const KernelArgs& A, double* sm, ShT<NP, TR, GN, BT, MD>& X, bool mark_empty) {
    const DevTables& T = A.T;
    const int nb = A.L.n_batch > 1 ? A.L.n_batch : 1;
    const Lay lay(T.nr, T.nt, T.np, NP, nb, A.L.sdet_doubles);
    constexpr int RC = ShT<NP, TR, GN, BT, MD>::RC;
    X.sd = sm;
    X.si = reinterpret_cast<int*>(X.sd + NF_HOT * NP);
    X.q = reinterpret_cast<short*>(X.si + NI_HOT * NP);
    X.head = reinterpret_cast<int*>(X.q + N_LISTS * RC);
    X.tail = X.head + 16; X.misc = X.head + 32;
    X.sdet = A.L.sdet_doubles > 0 ? sm + lay.o_det : nullptr;
    for (int i = threadIdx.x; i < A.L.sdet_doubles; i += NT) sm[lay.o_det + i] = 0.0;
    X.r = sm + lay.o_r; X.r2 = sm + lay.o_r2; X.tf = sm + lay.o_tf; X.ttan = sm + lay.o_tt; X.ps = sm + lay.o_ps; X.pc = sm + lay.o_pc; X.pf = sm + lay.o_pf;
    X.tplane = reinterpret_cast<const int*>(sm + lay.o_tp);
    X.cdfa = reinterpret_cast<const double2*>(sm + lay.o_ca);
    X.geo = sm + lay.o_geo;
    X.cold = A.O.scratch + (size_t)blockIdx.x * NP * REC;
    const int tid = threadIdx.x;
    for (int i = tid; i <= T.nr; i += NT) { const double r = T.rfront[i]; sm[lay.o_r + i] = r; sm[lay.o_r2 + i] = r * r; }
    for (int i = tid; i <= T.nt; i += NT) {
        sm[lay.o_tf + i] = T.thetafront[i]; sm[lay.o_tt + i] = T.ttan[i];
        reinterpret_cast<int*>(sm + lay.o_tp)[i] = T.tplane[i];
    }
    for (int i = tid; i < T.np; i += NT) { sm[lay.o_ps + i] = T.psin[i]; sm[lay.o_pc + i] = T.pcos[i]; sm[lay.o_pf + i] = T.phifront[i]; }
    for (int i = tid; i < 2 * 181; i += NT) sm[lay.o_ca + i] = T.cdfA2[i];
    if (BT && A.L.n_batch > 1) { for (int i = tid; i < GEO * nb; i += NT) sm[lay.o_geo + i] = A.L.geo[i]; }
    else if (BT && tid < GEO) {
        const LaunchArgs& L = A.L;
        sm[lay.o_geo + tid] = (tid < 3) ? L.det[tid] : (tid == 3) ? L.sin_dt : (tid == 4) ? L.cos_dt : (tid == 5) ? L.sin_dp
                            : (tid == 6) ? L.cos_dp : (tid == 7) ? (double)L.limb_emission : (tid == 8) ? L.det_sph_theta
                            : (tid == 9) ? L.det_sph_phi : (tid == 10) ? (double)A.T.cell_depth : 0.0;
    }
    if (mark_empty) for (int i = tid; i < N_LISTS * RC; i += NT) X.q[i] = (short)-1;
    __syncthreads();
    for (int i = tid; i < NP; i += NT) { X.Q(L_EMIT, i) = (short)i; X.I(I_ND, i) = 0; X.I(I_INFO, i) = 0; }   // every slot starts by asking for a photon
    if (tid < 16) { X.head[tid] = 0; X.tail[tid] = (tid == L_EMIT) ? NP : 0; }
    if (tid < 32) X.misc[tid] = 0;      // [0] retired slots, [1], [2] bulk-synchronous kernel, [8..31] watchdog words of the (<= 8) warps
    __syncthreads();
}

// list a ray goes on when it ends, by [kind][outcome]  (-1: it does not end)
__constant__ signed char c_list[4][8] = {
    //  NONE  LIMIT   EXIT    SURF    REST   RESP   ERR     DEAD
    {-1, L_PRE, L_PRE, L_PRE, L_RES, L_RES, L_EMIT, L_EMIT},     // K_PRE
    {-1, L_H, L_EMIT, L_EMIT, L_RES, L_RES, L_EMIT, L_EMIT},     // K_WALK
    {-1, L_DEP, L_DEP, L_DEP, L_RES, L_RES, L_EMIT, L_EMIT},     // K_PEEL
    {-1, L_EMIT, L_EMIT, L_EMIT, L_EMIT, L_EMIT, L_EMIT, L_EMIT} // K_DEAD
};

// A marcher lane: the ray it is stepping and one step of it.  The step is written for a short instruction stream:
// every update is committed unconditionally (an interaction inside the cell is recognised by acc > lim AFTER the
// crossing was added, and the event subtracts the overshoot), the radial direction is a +-1 register, the opacity
// row pointer is kept, and the sphere radii come squared from shared memory.
struct Marcher {
    int slot, c0, dr, cell12, info;
    double t, acc, tr, tt, tp, hbn, D0, iq, lim, kap, ds;
    const double* kb;            // kext + nr * col: the opacity row of the ray's (theta, phi) column
    int col;                     // column index c1 + nt*c2 (+ nt*np * wavelength in a wavelength batch)
    int nr, nt, depth;           // launch invariants kept in registers (the kernel parameters live in constant memory)
    const double* r2g;           // (the generic pointer: measurement variant without E2_OPT_R2S)
    unsigned r2s;                // shared-space byte address of the squared radii: ld.shared with a 32-bit address instead of a generic
                                 // pointer (the compiler rebuilt the generic shared base with S2R + LEA in every step)
    int tl; unsigned long long th, pid;   // walk recorder (trace hook only; dead code otherwise)
    double s0w;                           // Stokes I of the photon (latitudinal flow counters only)

    __device__ __forceinline__ void init(const DevTables& T) {
        slot = -1; c0 = cell12 = info = 0; dr = 1;
        t = acc = tr = tt = tp = hbn = D0 = iq = lim = kap = 0.0; ds = 1.0;
        col = 0;
        nr = T.nr; nt = T.nt; depth = T.cell_depth; kb = T.kext; r2s = 0u;
    }
    // the ray is in column `column` from now on
    __device__ __forceinline__ void set_column(const DevTables& T, int column) {
        col = column; kb = T.kext + (size_t)T.nr * (size_t)column;
    }
    template <class Sh>
    __device__ __forceinline__ void bind(const Sh& X) {
        r2g = X.r2;
        unsigned long long sh;      // volatile: converted ONCE and kept, not rematerialised (S2UR + ULEA) at every use
        asm volatile("cvta.to.shared.u64 %0, %1;" : "=l"(sh) : "l"((unsigned long long)X.r2));
        r2s = (unsigned)sh;
    }
    // opacity of layer c0 (measured and dropped, twice: four layers per 256-bit load and a select -- the branch and the selects in
    // the stepping loop cost 20-28 %, profiles/r02_ab_variants.txt; the next layer loaded one step ahead -- 2-6 %)
    __device__ __forceinline__ void load_kap() { kap = __ldg(kb + c0); }
    __device__ __forceinline__ double r2_at(int i) const {
        if (!E2_OPT_R2S) return r2g[i];
        double v;
        asm("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(r2s + 8u * (unsigned)i));
        return v;
    }

    template <class Sh>
    __device__ __forceinline__ void load(const Sh& X, const KernelArgs& A, int s) {
        slot = s;
        if (Sh::GEN && A.L.flow_theta) s0w = X.D(F_S0, s);
        t = X.D(F_T, s); acc = X.D(F_ACC, s); tr = X.D(F_TR, s); tt = X.D(F_TT, s); tp = X.D(F_TP, s);
        hbn = X.D(F_HBN, s); D0 = X.D(F_D0, s); iq = X.D(F_IQ, s); lim = X.D(F_LIM, s);
        const int cell = X.I(I_CELL, s);
        info = X.I(I_INFO, s);
        c0 = cell & 1023; cell12 = cell & ~1023;
        dr = (info & B_INWARD) ? -1 : 1; ds = (info & B_INWARD) ? -1.0 : 1.0;
        int column = ((cell >> 10) & 1023) + nt * ((cell >> 20) & 1023);
        if (Sh::BATCH && A.L.wl_batch) {       // wavelength batch: the opacity table and the surface layer of the photon's launch
            const int kbi = X.I(I_BATCH, s);
            column += wl_of(X, A, kbi) * (A.T.nt * A.T.np);
            depth = depth_of(X, A, kbi);
        }
        set_column(A.T, column);
        load_kap();
        if (Sh::TRACE) {
            tl = X.I(I_TLEN, s);
            th = (unsigned long long)(unsigned)X.I(I_THLO, s) | ((unsigned long long)(unsigned)X.I(I_THHI, s) << 32);
            pid = (unsigned long long)(unsigned)X.I(I_IDLO, s) | ((unsigned long long)(unsigned)X.I(I_IDHI, s) << 32);
        }
    }

    // One step: advance to the next crossing.  Returns the outcome (O_NONE: the ray goes on).
    template <class Sh>
    __device__ __forceinline__ int step(const Sh& X, const KernelArgs& A, unsigned& n_cf) {
        // (E2_EARLY_ROOT) the crossing after this one, should this one be radial: cell cn = c0 + dr, inner sphere while the ray still
        // reaches it (inward), else the outer one.  Nothing here depends on which face is crossed now, so the square root's chain
        // (MUFU seed + Newton step + residual, ~10 dependent FP64 operations) runs beside the min / accumulate / compare chain below
        // instead of after it.  A step that ends on a polar / azimuthal face or on the optical-depth limit discards it.
        double trn = 0.0;
        bool inner = false;
        if (E2_EARLY_ROOT) {
            const int cn = min(max(c0 + dr, 0), nr - 1);
            const double da = fma(r2_at(cn), iq, D0), db = fma(r2_at(cn + 1), iq, D0);
            inner = (dr < 0) && !(da < 0.0);
            const double disc = inner ? da : db;
            trn = fma(inner ? -1.0 : 1.0, fsqrt(disc), hbn);
            if (disc < 0.0) trn = RAY_NONE;
            asm volatile("" : "+d"(trn));      // computed HERE, not sunk below the branches that follow
        }
        double tn = tr;
        if (tt < tn) tn = tt;
        if (tp < tn) tn = tp;
        ++n_cf;
        if (Sh::TRACE) {   // the cell_face outcome the reference would record here: (next_face(2), cell_out(3))
            const int c1 = (cell12 >> 10) & 1023, c2 = (cell12 >> 20) & 1023, np = A.T.np;
            if (!(tn < RAY_NONE)) trace_tuple(A, pid, tl, th, 0, 0, 0, 0, 0);
            else if (tn == tr) trace_tuple(A, pid, tl, th, 1, c0 + (dr > 0 ? 1 : 0), c0 + dr, c1, c2);
            else if (tn == tt) trace_tuple(A, pid, tl, th, 2, (info & B_TUPPER) ? c1 + 1 : c1, c0, (info & B_TUPPER) ? c1 + 1 : c1 - 1, c2);
            else {
                const int up2 = (c2 + 1 == np) ? 0 : c2 + 1, dn2 = (c2 == 0) ? np - 1 : c2 - 1;
                trace_tuple(A, pid, tl, th, 3, (info & B_PUP) ? up2 : c2, c0, c1, (info & B_PUP) ? up2 : dn2);
            }
        }
        if (!(tn < RAY_NONE)) return O_ERR;
        acc = fma(tn - t, kap, acc);
        const bool radial = (tn == tr);
        t = tn;
        if (acc > lim) return O_LIMIT;          // lim = +inf for the probe walks
        if (!radial) return (tn == tt) ? O_REST : O_RESP;
        const int up = (dr > 0) ? 1 : 0;
        const int f = c0 + up;
        if (Sh::GEN && A.L.flow_theta && (info & 3) == K_WALK)      // add_flow :5016-5047, radial crossings (:730-735)
            atomicAdd(A.O.flow4 + (size_t)4 * ((size_t)nr * col + c0) + (up ? 0 : 1), s0w);
        // outward the ray can only leave through the top face, inward only reach the surface face: one comparison.  The surface
        // face of a plain launch is read from the constant bank (as a register it was spilled and reloaded in every step).
        const int dep = (Sh::BATCH && A.L.wl_batch) ? depth : A.T.cell_depth;
        if (f == (up ? nr : dep)) return up ? O_EXIT : O_SURF;
        c0 += dr;
        load_kap();
        if (E2_EARLY_ROOT) {
            if (dr < 0 && !inner) { dr = 1; ds = 1.0; }      // turning point passed: outward from here on
            tr = trn;
            return O_NONE;
        }
        // inward: the inner sphere if the ray reaches it, else (turning point passed) the outer one
        double disc = fma(r2_at(c0 + up), iq, D0);
        if (disc < 0.0) {
            if (dr < 0) { dr = 1; ds = 1.0; disc = fma(r2_at(c0 + 1), iq, D0); }
            if (disc < 0.0) { tr = RAY_NONE; return O_NONE; }
        }
        tr = fma(ds, fsqrt(disc), hbn);
        return O_NONE;
    }

    // The ray ended with outcome `out`: write its state back to the slot; returns the event list the slot goes on.
    template <class Sh>
    __device__ __forceinline__ int finish(const Sh& X, const KernelArgs& A, Cnt& C, int out) {
        const int kind = info & 3;
        X.D(F_T, slot) = t; X.D(F_ACC, slot) = acc; X.D(F_TR, slot) = tr;   // (tr: a re-solved ray goes on)
        X.I(I_CELL, slot) = cell12 | c0;
        X.I(I_INFO, slot) = (info & 0xff & ~B_INWARD) | (dr < 0 ? B_INWARD : 0) | (out << 8);
        if (Sh::TRACE) { X.I(I_TLEN, slot) = tl; X.I(I_THLO, slot) = (int)(unsigned)th; X.I(I_THHI, slot) = (int)(unsigned)(th >> 32); }
        if (out == O_ERR) {
            err_count(A, 31); ++C.n_err;
            err_count(A, kind == K_PRE ? 2 : (kind == K_WALK ? 3 : 43));
        } else if (out == O_SURF && kind == K_WALK) {
            ++C.n_surf;
            if (Sh::GEN && A.L.surface_albedo > 0.0) return L_SURF;      // absorbed or reflected: decided by the SURF event
            X.I(I_ND, slot) += 1;                                        // black surface: absorbed (:755-764: one draw)
        }
        return c_list[kind][out];
    }

    // the pass ended before the ray did: write back what a step changes, so that any lane can go on with it (E2_RELEASE)
    template <class Sh>
    __device__ __forceinline__ void release(const Sh& X) {
        X.D(F_T, slot) = t; X.D(F_ACC, slot) = acc; X.D(F_TR, slot) = tr;
        X.I(I_CELL, slot) = cell12 | c0;
        X.I(I_INFO, slot) = (info & 0xff & ~B_INWARD) | (dr < 0 ? B_INWARD : 0);
        if (Sh::TRACE) { X.I(I_TLEN, slot) = tl; X.I(I_THLO, slot) = (int)(unsigned)th; X.I(I_THHI, slot) = (int)(unsigned)(th >> 32); }
    }

    template <class Sh>
    __device__ __forceinline__ int trip(const Sh& X, const KernelArgs& A, Cnt& C) {
        if ((info & 3) == K_DEAD) return finish(X, A, C, O_DEAD);
        unsigned n = 0;
        const int out = step(X, A, n);
        C.n_cf += n;
        return (out == O_NONE) ? -1 : finish(X, A, C, out);
    }
};


// ---------------------------------------------------------------------------------------------------
// A ray marched INSIDE an event, by the lane that owns the photon (multi-detector fan-out; peel-off walk of the interaction
// event with E2_INLINE_PEEL): no slot, no list, no record traffic -- the ray state never leaves registers, polar / azimuthal
// faces are re-solved inline.  Returns the outcome (O_EXIT, O_SURF, O_LIMIT, O_ERR) and the optical depth walked.
// ---------------------------------------------------------------------------------------------------
template <class Sh>
__device__ __forceinline__ int march_inline(const Sh& X, const KernelArgs& A, double px, double py, double pz, double n0, double n1, double n2,
                                            int cell, int slot, double lim, double& acc_out, unsigned& n_step) {
    const DevTables& T = A.T;
    const int c0 = cell & 1023, c1 = (cell >> 10) & 1023, c2 = (cell >> 20) & 1023;
    Marcher M;
    M.init(T); M.bind(X);
    int inward, upper = 0, up = 0;
    {
        RayK Kc;       // the quadric constants are needed for the first solves only; a polar / azimuthal crossing (rare) rebuilds them,
        double hbn, D0, iq;      // so that they do not occupy 28 registers while the ray is marched
        ray_consts(T, px, py, pz, n0, n1, n2, Kc, hbn, D0, iq);
        M.tr = radial_first(X, c0, -1, hbn, D0, iq, inward);
        M.tt = (T.nt > 1) ? theta_next(X, T.nt, c1, 0.0, Kc, upper) : RAY_NONE;
        M.tp = phi_next(X, T.np, c2, 0.0, Kc, up);
        M.hbn = hbn; M.D0 = D0; M.iq = iq;
    }
    M.t = 0.0; M.acc = 0.0; M.lim = lim;
    M.c0 = c0; M.cell12 = cell & ~1023;
    M.info = K_PEEL | (inward ? B_INWARD : 0) | (upper ? B_TUPPER : 0) | (up ? B_PUP : 0);
    M.dr = inward ? -1 : 1; M.ds = inward ? -1.0 : 1.0;
    int wlcol = 0;                         // first column of the photon's wavelength (0 unless a wavelength batch)
    if (Sh::BATCH && A.L.wl_batch) {       // wavelength batch: the opacity table and the surface layer of the photon's launch
        const int kbi = X.I(I_BATCH, slot);
        wlcol = wl_of(X, A, kbi) * (T.nt * T.np);
        M.depth = depth_of(X, A, kbi);
    }
    M.set_column(T, wlcol + c1 + T.nt * c2);
    M.load_kap();
    M.slot = slot;
    if (Sh::TRACE) {
        M.tl = X.I(I_TLEN, slot);
        M.th = (unsigned long long)(unsigned)X.I(I_THLO, slot) | ((unsigned long long)(unsigned)X.I(I_THHI, slot) << 32);
        M.pid = (unsigned long long)(unsigned)X.I(I_IDLO, slot) | ((unsigned long long)(unsigned)X.I(I_IDHI, slot) << 32);
    }
    int out;
#pragma unroll 1
    for (;;) {
        out = M.step(X, A, n_step);
        if (out == O_NONE) continue;
        if (out != O_REST && out != O_RESP) break;
        // a polar or azimuthal face: move the cell index and re-solve that axis (the RES event, inline)
        int cc1 = (M.cell12 >> 10) & 1023, cc2 = (M.cell12 >> 20) & 1023;
        RayK Kc;
        double h_, d_, i_;
        ray_consts(T, px, py, pz, n0, n1, n2, Kc, h_, d_, i_);
        if (out == O_REST) {
            cc1 += (M.info & B_TUPPER) ? 1 : -1;
            int up2;
            M.tt = theta_next(X, T.nt, cc1, M.t, Kc, up2);
            M.info = (M.info & ~B_TUPPER) | (up2 ? B_TUPPER : 0);
        } else {
            if (M.info & B_PUP) cc2 = (cc2 + 1 == T.np) ? 0 : cc2 + 1; else cc2 = (cc2 == 0) ? T.np - 1 : cc2 - 1;
            int up2;
            M.tp = phi_next(X, T.np, cc2, M.t, Kc, up2);
        }
        M.cell12 = (cc1 << 10) | (cc2 << 20);
        M.set_column(T, wlcol + cc1 + T.nt * cc2);
        M.load_kap();
    }
    if (Sh::TRACE) { X.I(I_TLEN, slot) = M.tl; X.I(I_THLO, slot) = (int)(unsigned)M.th; X.I(I_THHI, slot) = (int)(unsigned)(M.th >> 32); }
    acc_out = M.acc;
    return out;
}

// Deposit of a scattering peel-off (:4955-4972), called by the WHOLE warp: lanes that hit the same pixel are summed in the warp first
// (up to four pixel groups: 1x1 detectors, launch boundaries of a batched launch); what is left goes lane by lane.  Always to the
// global image: a pointer that may be shared OR global would turn these reductions into generic-address atomics.
__device__ __forceinline__ void deposit_scatter_warp(const KernelArgs& A, bool dep, int pix, const double v[8]) {
    const unsigned dm = __ballot_sync(FULL, dep);
    if (!dm) return;
    const LaunchArgs& L = A.L;
    const size_t npx = (size_t)L.nx * L.ny;
    unsigned rem = dm, left = 0u;
    const int lane = threadIdx.x & 31;
#pragma unroll 1
    for (int it = 0; it < 4 && rem; ++it) {
        const int pixg = __shfl_sync(FULL, pix, __ffs(rem) - 1);
        const unsigned grp = __ballot_sync(FULL, ((rem >> lane) & 1u) && pix == pixg);
        rem &= ~grp;
        if (__popc(grp) <= 2) { left |= grp; continue; }
        const bool in = (grp >> lane) & 1u;
        double x = 0.0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            double r = in ? v[k] : 0.0;
            for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(FULL, r, o);
            if (lane == k) x = r;
        }
        double* d = A.O.det + pixg;
        if (lane < 8) atomicAdd(d + (size_t)lane * npx, x);
        else if (lane < 10) atomicAdd(d + (size_t)lane * npx, (double)__popc(grp));
    }
    left |= rem;
    if ((left >> lane) & 1u) {
        double* d = A.O.det + pix;
#pragma unroll
        for (int k = 0; k < 8; ++k) atomicAdd(d + (size_t)k * npx, v[k]);
        atomicAdd(d + 8 * npx, 1.0); atomicAdd(d + 9 * npx, 1.0);
    }
}

// ---------------------------------------------------------------------------------------------------
// Multi-detector walks (ShT::MULTI, artes_gpu_run_multi).  The reference runs the whole random walk once per detector
// azimuth of a phase curve (src/ARTES.f90:215-245: 73 calls of radiative_transfer that differ in det_phi only).  Peel-off
// is a next-event estimate: at every scattering the walk may be observed from ANY direction without disturbing it, so
// ONE walk can feed all K detectors of the launch table -- every detector receives exactly the deposits the reference's
// run with that det_phi would make along this walk (same weights, same pixel rule); only the walk is shared, i.e. the
// K images are statistically correlated instead of independent.  The interaction event is cut in three:
//   H    survival :791-813                                              -> FAN list
//   FAN  one warp per photon, lanes = detectors: peel-off weight and pixel (:4763-4951), the walk to the grid exit
//        (:4739-4761) marched INLINE by the lane (no slot, no list: the ray state never leaves registers), e^-tau and the
//        deposit (:4955-4972), 32 detectors per round                      -> SC list
//   SC   scattering :819-845 (new direction, Stokes vector, optical depth)  -> transport ray
// Star source, black surface, no flow counters (the general paths keep their per-detector launches).
// ---------------------------------------------------------------------------------------------------

// H (multi): the transport walk reached its optical depth: step back to the interaction point, survival :791-813
template <class Sh>
__device__ __forceinline__ int ev_survive(const Sh& X, const KernelArgs& A, bool valid, int s, Cnt& C) {
    if (!valid) return -1;
    const DevTables& T = A.T;
    const LaunchArgs& L = A.L;
    const int cell = X.I(I_CELL, s);
    const int c0 = cell & 1023, c1 = (cell >> 10) & 1023, c2 = (cell >> 20) & 1023;
    const int ci = c0 + T.nr * (c1 + T.nt * c2);
    double* rec = X.cold + (size_t)s * REC;
    double hx, hy, hz, dx, dy, dz, S[4], tau0, w0_;
    ldg256(rec, hx, hy, hz, dx);
    ldg256(rec + 4, dy, dz, S[0], S[1]);
    ldg256(rec + 8, S[2], S[3], tau0, w0_);
    (void)w0_;
    double kap_c, alb_c, u_bits, pad_c;
    ldg256_nc(T.cellrec + (size_t)4 * ci, kap_c, alb_c, u_bits, pad_c);
    (void)pad_c; (void)u_bits;
    const double tpos = X.D(F_T, s) - fdiv(X.D(F_ACC, s) - tau0, kap_c);      // step back by the overshoot (:705-720)
    const double px = hx + tpos * dx, py = hy + tpos * dy, pz = hz + tpos * dz;
    const unsigned long long id = (unsigned long long)(unsigned)X.I(I_IDLO, s) | ((unsigned long long)(unsigned)X.I(I_IDHI, s) << 32);
    unsigned nd = (unsigned)X.I(I_ND, s);
    bool alive = L.photon_scattering != 0;
    if (alive) { double xi; draws(X, A, s, id, nd, 1, &xi); ++nd; if (xi < L.fstop) alive = false; }
    if (alive) {
        if (alb_c < 1.0 && alb_c > 0.0) { const double g = fdiv(alb_c, 1.0 - L.fstop); S[0] *= g; S[1] *= g; S[2] *= g; S[3] *= g; }
        if (S[0] <= L.photon_minimum) alive = false;
    }
    X.I(I_ND, s) = (int)nd;
    if (!alive) { X.I(I_INFO, s) = K_DEAD; return L_EMIT; }
    stg256(rec, px, py, pz, dx);
    stg256(rec + 4, dy, dz, S[0], S[1]);
    stg256(rec + 8, S[2], S[3], tau0, 0.0);
    X.I(I_HCELL, s) = cell;
    return L_FAN;
}

// SC (multi): scattering :819-845 -- new direction and Stokes vector, next optical depth, transport ray
template <class Sh>
__device__ __forceinline__ int ev_scatter_md(const Sh& X, const KernelArgs& A, bool valid, int s, Cnt& C, RaySpec& rs) {
    if (!valid) return -1;
    const DevTables& T = A.T;
    double* rec = X.cold + (size_t)s * REC;
    double px, py, pz, dx, dy, dz, S[4], t0_, t1_;
    ldg256(rec, px, py, pz, dx);
    ldg256(rec + 4, dy, dz, S[0], S[1]);
    ldg256(rec + 8, S[2], S[3], t0_, t1_);
    (void)t0_; (void)t1_;
    const int cell = X.I(I_HCELL, s);
    const int c0 = cell & 1023, c1 = (cell >> 10) & 1023, c2 = (cell >> 20) & 1023;
    const int ci = c0 + T.nr * (c1 + T.nt * c2);
    double kap_c, alb_c, u_bits, pad_c;
    ldg256_nc(T.cellrec + (size_t)4 * ci, kap_c, alb_c, u_bits, pad_c);
    (void)kap_c; (void)alb_c; (void)pad_c;
    const int u = (int)__double_as_longlong(u_bits);
    const unsigned long long id = (unsigned long long)(unsigned)X.I(I_IDLO, s) | ((unsigned long long)(unsigned)X.I(I_IDHI, s) << 32);
    unsigned nd = (unsigned)X.I(I_ND, s);
    double xr[5];
    draws(X, A, s, id, nd, 4, xr);
    ++C.n_sc;
    FastAngles g;
    int e = sample_angles_f(X, A, xr[0], xr[1], xr[2], S, ci, g, u);
    nd += (e == 6) ? 2u : 3u;
    double e0 = 0, e1 = 0, e2 = 0, tau = -1.0;
    if (!e) {
        const double cto = fdiv(dz, fsqrt(dx * dx + dy * dy + dz * dz));
        const double sto = fsqrt(1.0 - cto * cto);
        const double ctn = cto * g.alpha + sto * g.sT * g.cb;
        const double stn = fsqrt(1.0 - ctn * ctn);
        double nc = fdiv(g.alpha - ctn * cto, stn * sto);
        if (!(nc == nc)) e = 20;
        else {
            if (nc >= 1.0) nc = 1.0 - 1.e-10; else if (nc <= -1.0) nc = -1.0 + 1.e-10;
            const double sD = fsqrt(1.0 - nc * nc) * (g.flip ? -1.0 : 1.0);
            const double rho = fsqrt(dx * dx + dy * dy);
            const double irho = frcp(rho);
            const double cph = rho > 0.0 ? dx * irho : 1.0, sph = rho > 0.0 ? dy * irho : 0.0;
            e0 = stn * (cph * nc - sph * sD); e1 = stn * (sph * nc + cph * sD); e2 = ctn;
            if (!(fabs(e2) < 1.0)) e = 16;
        }
    }
    if (!e) {
        double Sn[4];
        const double nc2 = fdiv(dz - e2 * g.alpha, g.sT * fsqrt(1.0 - e2 * e2));
        int soft = 0;
        e = polrot_deg_f(T, u, g.deg, g.cb * g.cb - g.sb * g.sb, 2.0 * g.sb * g.cb, g.flip, nc2, S, Sn, false, soft);
        if (soft) err_count(A, soft);
        if (!e) { S[0] = Sn[0]; S[1] = Sn[1]; S[2] = Sn[2]; S[3] = Sn[3]; dx = e0; dy = e1; dz = e2; }
    }
    if (e) { err_count(A, e); ++C.n_err; X.I(I_ND, s) = (int)nd; X.I(I_INFO, s) = K_DEAD; return L_EMIT; }
    ++nd;
    tau = -fm_log(1.0 - xr[3]);
    stg256(rec, px, py, pz, dx);
    stg256(rec + 4, dy, dz, S[0], S[1]);
    stg256(rec + 8, S[2], S[3], tau, 0.0);
    X.I(I_ND, s) = (int)nd;
    rs.set(px, py, pz, dx, dy, dz, c0, c1, c2, -1, K_WALK, tau);
    return L_RDY;
}

// FAN (multi): peel-off of the n photons of the batch (lane j holds slot s of photon j) towards every detector.
// The n x K (photon, detector) pairs are dealt to the lanes 32 at a time, so that a round is full whatever K is (68 detectors
// as 32 + 32 + 4 lanes per photon wasted a third of the lanes).  A round touches at most two or three photons: their records
// are read with one or two distinct addresses per instruction, the matrix rows of the detectors come from the same 23 KB
// blocks, and the walk to the detector runs in the lane.
template <class Sh>
__device__ __forceinline__ void ev_fan(const Sh& X, const KernelArgs& A, int n, int s_lane, Cnt& C) {
    const DevTables& T = A.T;
    const LaunchArgs& L = A.L;
    const int lane = threadIdx.x & 31;
    const int K = L.n_batch;
    const int total = n * K;
    const size_t npx = (size_t)L.nx * L.ny;
#pragma unroll 1
    for (int it = 0; it < total; it += 32) {
        {
            const int item = it + lane;
            const bool act = item < total;
            const int j = act ? item / K : 0;
            const int kd = act ? item - j * K : 0;
            const int sp = __shfl_sync(FULL, s_lane, j);
            const double* rec = X.cold + (size_t)sp * REC;
            double px, py, pz, dx, dy, dz, S[4], t0_, t1_;
            ldg256(rec, px, py, pz, dx);
            ldg256(rec + 4, dy, dz, S[0], S[1]);
            ldg256(rec + 8, S[2], S[3], t0_, t1_);
            (void)t0_; (void)t1_;
            const int cell = X.I(I_HCELL, sp);
            const int c0 = cell & 1023, c1 = (cell >> 10) & 1023, c2 = (cell >> 20) & 1023;
            const int ci = c0 + T.nr * (c1 + T.nt * c2);
            double kap_c, alb_c, u_bits, pad_c;
            ldg256_nc(T.cellrec + (size_t)4 * ci, kap_c, alb_c, u_bits, pad_c);
            (void)alb_c; (void)pad_c; (void)kap_c;
            const int u = (int)__double_as_longlong(u_bits);
            const bool dz_ok = fabs(dz) < 1.0;
            const double idz = dz_ok ? frcp(fsqrt(1.0 - dz * dz)) : 0.0;
            double W[4] = {0.0, 0.0, 0.0, 0.0};
            int pix = -1;
            Geo G = geo_of(X, L, kd);
            if (act) {
                ++C.n_peel;
                double mu = dx * G.d0 + dy * G.d1 + dz * G.d2;
                if (mu >= 1.0) mu = 1.0 - 1.e-10; else if (mu <= -1.0) mu = -1.0 + 1.e-10;
                const double peel_deg = fm_acos(mu) * (180.0 / PI);
                if (!dz_ok) err_count(A, 45);
                else {
                    const double smu = fsqrt(1.0 - mu * mu);
                    double nc = fdiv(G.d2 - dz * mu, smu) * idz;
                    if (!(nc == nc)) err_count(A, 44);
                    else {
                        nc = fmin(fmax(nc, -1.0), 1.0);
                        const double cr = dy * G.d0 - dx * G.d1;
                        const bool flip = (cr > 0.0) || (cr == 0.0 && dx * G.d0 + dy * G.d1 > 0.0);
                        const double c2a = 2.0 * nc * nc - 1.0;
                        double s2a = 2.0 * nc * fsqrt(fmax(1.0 - nc * nc, 0.0));
                        if (flip) s2a = -s2a;
                        const double nc2 = fdiv(dz - G.d2 * mu, smu * fsqrt(1.0 - G.d2 * G.d2));
                        int soft = 0;
                        const int e = (fabs(G.d2) < 1.0) ? polrot_deg_f(T, u, peel_deg, c2a, s2a, flip, nc2, S, W, true, soft) : 16;
                        if (e) err_count(A, e);
                        else if (!(W[0] > 0.0 && W[0] < 1.e100)) err_count(A, 53);
                        else {
                            const double x_im = py * G.cdp - px * G.sdp;
                            const double y_im = pz * G.sdt - py * G.cdt * G.sdp - px * G.cdt * G.cdp;
                            const int ix = (int)fdiv(L.nx * (x_im + L.x_max), 2.0 * L.x_max) + 1;
                            const int iy = (int)fdiv(L.ny * (y_im + L.y_max), 2.0 * L.y_max) + 1;
                            if (ix < 1 || ix > L.nx || iy < 1 || iy > L.ny) err_count(A, 60);
                            else pix = (ix - 1) + L.nx * (iy - 1) + kd * 10 * L.nx * L.ny;
                        }
                    }
                }
            }
            // ---- the walk to the detector (:4739-4761), in the lane; stops once tau >= 50 (the reference drops those, :4765)
            if (pix >= 0) {
                unsigned n_step = 0;
                double acc_w = 0.0;
                const int out = march_inline(X, A, px, py, pz, G.d0, G.d1, G.d2, cell, sp, 50.0, acc_w, n_step);
                C.n_cf += n_step;
                if (out == O_ERR) { err_count(A, 31); err_count(A, 43); ++C.n_err; }
                if (out == O_EXIT && acc_w < 50.0) {        // reached the detector: e^-tau, deposit :4955-4972
                    const double w = fm_exp_neg(acc_w);
                    const double v0 = w * W[0], v1 = -(w * W[1]), v2 = w * W[2], v3 = w * W[3];
                    if (X.sdet) {
                        double* d = X.sdet + pix;
                        atomicAdd(d, v0); atomicAdd(d + npx, v1); atomicAdd(d + 2 * npx, v2); atomicAdd(d + 3 * npx, v3);
                        atomicAdd(d + 4 * npx, v0 * v0); atomicAdd(d + 5 * npx, v1 * v1); atomicAdd(d + 6 * npx, v2 * v2); atomicAdd(d + 7 * npx, v3 * v3);
                        atomicAdd(d + 8 * npx, 1.0); atomicAdd(d + 9 * npx, 1.0);
                    } else {
                        double* d = A.O.det + pix;
                        atomicAdd(d, v0); atomicAdd(d + npx, v1); atomicAdd(d + 2 * npx, v2); atomicAdd(d + 3 * npx, v3);
                        atomicAdd(d + 4 * npx, v0 * v0); atomicAdd(d + 5 * npx, v1 * v1); atomicAdd(d + 6 * npx, v2 * v2); atomicAdd(d + 7 * npx, v3 * v3);
                        atomicAdd(d + 8 * npx, 1.0); atomicAdd(d + 9 * npx, 1.0);
                    }
                }
            }
            __syncwarp();
        }
    }
}

#ifndef E2_HEAD_STEPS
#define E2_HEAD_STEPS 0    // > 0: the event that sets up a transport ray also marches its first E2_HEAD_STEPS crossings; a ray that ends
#endif                     //      within them goes straight to its event list, without the trip through the ready list and a marcher

// the first crossings of the transport ray an event has just set up; returns the list the slot goes on (L_RDY: the ray goes on)
template <class Sh>
__device__ __forceinline__ int head_march(const Sh& X, const KernelArgs& A, int s, Cnt& C) {
    Marcher Mw;
    Mw.init(A.T); Mw.bind(X);
    Mw.load(X, A, s);
    unsigned n_step = 0;
    int out = O_NONE;
#pragma unroll 1
    for (int k = 0; k < E2_HEAD_STEPS && out == O_NONE; ++k) out = Mw.step(X, A, n_step);
    C.n_cf += n_step;
    if (out != O_NONE) return Mw.finish(X, A, C, out);
    Mw.release(X);
    return L_RDY;
}

// the event of list l; returns the list the lane's slot goes on next (-1: none)
template <class Sh>
__device__ __forceinline__ int run_event(const Sh& X, const KernelArgs& A, int l, bool valid, int s, Cnt& C) {
    RaySpec rs;
    rs.make = false;
    bool push;
    if (l == L_H) push = ev_interact(X, A, valid, s, C, rs);
    else if (l == L_DEP) push = ev_deposit(X, A, valid, s, C, rs);
    else if (l == L_RES) push = ev_resolve(X, A, valid, s);
    else if (l == L_PRE) push = ev_pre(X, A, valid, s, C, rs);
    else if (Sh::GEN && l == L_SURF) push = ev_surface(X, A, valid, s, C, rs);
    else push = ev_emit(X, A, valid, s, C, rs);
    if (rs.make) ray_setup(X, A.T, s, rs.x, rs.y, rs.z, rs.n0, rs.n1, rs.n2, rs.c0, rs.c1, rs.c2, rs.sface, rs.kind, rs.lim, rs.acc0, rs.pk);
    if (E2_HEAD_STEPS > 0 && rs.make && rs.kind == K_WALK) return head_march(X, A, s, C);
    return push ? L_RDY : -1;
}

// multi-detector walks: the event of list l for a batch of n slots; returns the list the lane's slot goes on next (-1: none)
template <class Sh>
__device__ __forceinline__ int run_event_md(const Sh& X, const KernelArgs& A, int l, int n, bool valid, int s, Cnt& C) {
    RaySpec rs;
    rs.make = false;
    int tgt;
    if (l == L_FAN) { ev_fan(X, A, n, s, C); tgt = valid ? L_SC : -1; }
    else if (l == L_SC) tgt = ev_scatter_md(X, A, valid, s, C, rs);
    else if (l == L_H) tgt = ev_survive(X, A, valid, s, C);
    else if (l == L_RES) tgt = ev_resolve(X, A, valid, s) ? L_RDY : -1;
    else if (l == L_PRE) tgt = ev_pre(X, A, valid, s, C, rs) ? L_RDY : -1;
    else tgt = ev_emit(X, A, valid, s, C, rs) ? L_RDY : -1;
    if (rs.make) ray_setup(X, A.T, s, rs.x, rs.y, rs.z, rs.n0, rs.n1, rs.n2, rs.c0, rs.c1, rs.c2, rs.sface, rs.kind, rs.lim, rs.acc0, rs.pk);
    return tgt;
}

__device__ __forceinline__ void flush_counters(const KernelArgs& A, const Cnt& C) {
    const int lane = threadIdx.x & 31;
    unsigned long long v[7] = {C.n_emit, C.n_cf, C.n_sc, C.n_peel, C.n_surf, C.n_draw, C.n_err};
#pragma unroll
    for (int k = 0; k < 7; ++k) {
        unsigned long long x = v[k];
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(FULL, x, o);
        if (lane == 0 && x) atomicAdd(A.O.stats + k, x);
    }
}

#ifdef ARTES_TUNING   // comparison kernel, tuning builds only (make TUNING=1)
// ---------------------------------------------------------------------------------------------------
// kernel A: bulk-synchronous rounds (marcher phase of `trips` trips | barrier | event phase | barrier)
// ---------------------------------------------------------------------------------------------------
template <int NT, int NP, int MINB, int NRAY>
__global__ void __launch_bounds__(NT, MINB) transport2_kernel(const __grid_constant__ KernelArgs A) {
    extern __shared__ double smraw[];
    const DevTables& T = A.T;
    using Sh = ShT<NP, false, false>;
    Sh X;
    block_setup<NT, NP, false, false, false>(A, smraw, X, false);
    const int tid = threadIdx.x, lane = tid & 31;
    const unsigned lt = (1u << lane) - 1u;
    const int trips = A.L.e2_trips > 0 ? A.L.e2_trips : 32;      // marcher steps per round
    const int inner = A.L.e2_inner > 0 ? A.L.e2_inner : 4;       // steps per bookkeeping pass
    Cnt C; C.n_cf = 0; C.n_emit = C.n_sc = C.n_peel = C.n_surf = C.n_err = C.n_draw = 0;
    Marcher M[NRAY];              // NRAY independent rays per lane: their dependency chains interleave (ILP)
#pragma unroll
    for (int j = 0; j < NRAY; ++j) { M[j].init(T); M[j].bind(X); }
    volatile int* vhead = X.head;
    volatile int* vtail = X.tail;

    for (;;) {
        // ================= marcher phase =================
        // `trips` steps per round in passes of `inner`: claim rays for the free lanes, step every lane `inner` times
        // in a tight loop (a lane whose ray ends inside the pass waits for its end), write back and push what ended.
        for (int trip = 0; trip < trips; trip += inner) {
            // ---- free lanes claim ready rays
#pragma unroll
            for (int j = 0; j < NRAY; ++j) {
                const unsigned fm = __ballot_sync(FULL, M[j].slot < 0);
                if (fm && (vtail[L_RDY] - vhead[L_RDY]) > 0) {
                    int base = 0, n = 0;
                    if (lane == 0) {
                        const int want = __popc(fm);
                        int h = vhead[L_RDY];
                        for (;;) {
                            n = min(want, vtail[L_RDY] - h);
                            if (n <= 0) { n = 0; break; }
                            const int old = atomicCAS(X.head + L_RDY, h, h + n);
                            if (old == h) { base = h; break; }
                            h = old;
                        }
                    }
                    base = __shfl_sync(FULL, base, 0); n = __shfl_sync(FULL, n, 0);
                    const int rank = __popc(fm & lt);
                    if (M[j].slot < 0 && rank < n) M[j].load(X, A, X.Q(L_RDY, base + rank));
                }
            }
            // ---- `inner` steps of every ray
            int out[NRAY];
            unsigned n_step = 0;
#pragma unroll
            for (int j = 0; j < NRAY; ++j) out[j] = (M[j].slot >= 0 && (M[j].info & 3) == K_DEAD) ? O_DEAD : O_NONE;
#pragma unroll 1
            for (int k = 0; k < inner; ++k) {
#pragma unroll
                for (int j = 0; j < NRAY; ++j)
                    if (M[j].slot >= 0 && out[j] == O_NONE) out[j] = M[j].step(X, A, n_step);
            }
            C.n_cf += n_step;
            // ---- write back and push ended rays on their event lists (one shared-memory atomic per list present in the warp)
#pragma unroll
            for (int j = 0; j < NRAY; ++j) {
                int lst = -1;
                if (M[j].slot >= 0 && out[j] != O_NONE) lst = M[j].finish(X, A, C, out[j]);
                if (__any_sync(FULL, lst >= 0)) {
                    const unsigned g = __match_any_sync(FULL, lst);
                    const int leader = __ffs(g) - 1;
                    int base = 0;
                    if (lane == leader && lst >= 0) base = atomicAdd(X.tail + lst, __popc(g));
                    base = __shfl_sync(FULL, base, leader);
                    if (lst >= 0) { X.Q(lst, base + __popc(g & lt)) = (short)M[j].slot; M[j].slot = -1; }
                }
            }
        }
        __syncthreads();
        // ================= event phase =================
        // Every warp claims 32-event batches of one type until none is left, heaviest type first.  Only full
        // batches are served unless the marchers would run short of ready rays.
        {
            const int order[5] = {L_H, L_DEP, L_RES, L_PRE, L_EMIT};
            int hd[5], m[5], boff[6], full = 0;
#pragma unroll
            for (int k = 0; k < 5; ++k) { hd[k] = X.head[order[k]]; m[k] = X.tail[order[k]] - hd[k]; full += m[k] >> 5; }
            const bool partial = (X.misc[2] - X.head[L_RDY]) + 32 * full < NRAY * (NT + NT / 2);
            boff[0] = 0;
#pragma unroll
            for (int k = 0; k < 5; ++k) { if (!partial) m[k] &= ~31; boff[k + 1] = boff[k] + ((m[k] + 31) >> 5); }
            for (;;) {
                int b = 0;
                if (lane == 0) b = atomicAdd(X.misc + 1, 1);
                b = __shfl_sync(FULL, b, 0);
                if (b >= boff[5]) break;
                int k = 0;
#pragma unroll
                for (int j = 1; j < 5; ++j) if (b >= boff[j]) k = j;
                const int idx = ((b - boff[k]) << 5) + lane;
                const bool valid = idx < m[k];
                const int l = order[k];
                const int s = valid ? (int)X.Q(l, hd[k] + idx) : 0;
                const bool push = run_event(X, A, l, valid, s, C) >= 0;
                const unsigned pm = __ballot_sync(FULL, push);
                if (pm) {
                    const int leader = __ffs(pm) - 1;
                    int base = 0;
                    if (lane == leader) base = atomicAdd(X.tail + L_RDY, __popc(pm));
                    base = __shfl_sync(FULL, base, leader);
                    if (push) X.Q(L_RDY, base + __popc(pm & lt)) = (short)s;
                }
            }
            __syncthreads();
            if (tid == 0) {
#pragma unroll
                for (int k = 0; k < 5; ++k) X.head[order[k]] = hd[k] + m[k];
                X.misc[1] = 0; X.misc[2] = X.tail[L_RDY];
            }
        }
        if (X.misc[0] >= NP) break;
    }
    flush_counters(A, C);
}

#endif  // ARTES_TUNING

// ---------------------------------------------------------------------------------------------------
// kernel B: asynchronous.  No block barrier after set-up: every warp loops  claim rays -> one trip -> push
// ended rays -> if some event list holds a full batch, take it and run it.  List entries are published
// through the ring cell itself (-1 = empty): a producer reserves a position with an atomic on `tail`, fences
// its slot writes and stores the slot id; a consumer reserves positions with a CAS on `head`, waits for the
// id to appear, clears the cell and fences before touching the slot.  A warp with (almost) nothing to march
// and no ready ray to claim also takes partial batches, which is what drains the lists at the end.
// ---------------------------------------------------------------------------------------------------
// Watchdog.  A list cell is published by exactly one producer and a consumer may wait for it; if the scheduling logic were ever
// wrong such a wait (or a block whose warps all find nothing to do) would hang the GPU.  Waits therefore count their polls:
// after ~10 s without progress a warp raises the launch's abort word (error slot 63), every loop that sees it leaves, and the host
// returns an error instead of a hung device.  The polls of a healthy launch end within microseconds, the counter costs nothing.
constexpr int ERR_WATCHDOG = 63;
// the waits themselves are out of line: the common case (the cell is already published / free) costs one load and no register
__device__ __noinline__ int ring_wait_take(volatile short* e, unsigned long long* abort_word) {
    unsigned spin = 0;
    for (;;) {
        const short v = *e;
        if (v >= 0) return (int)v;
        if (!E2_WATCHDOG || (++spin & 0xfffffu) != 0u) continue;
        if (*(volatile unsigned long long*)abort_word) return 0;
        if (spin >= (400u << 20)) { atomicExch(abort_word, 1ull); return 0; }
    }
}
__device__ __noinline__ void ring_wait_put(volatile short* e, unsigned long long* abort_word) {
    unsigned spin = 0;
    while (*e >= 0) {
        if (!E2_WATCHDOG || (++spin & 0xfffffu) != 0u) continue;
        if (*(volatile unsigned long long*)abort_word) return;
        if (spin >= (400u << 20)) { atomicExch(abort_word, 1ull); return; }
    }
}
__device__ __forceinline__ int ring_take(volatile short* e, unsigned long long* abort_word) {
    int v = *e;
    if (v < 0) v = ring_wait_take(e, abort_word);
    *e = (short)-1;
    return v;
}
__device__ __forceinline__ void ring_put(volatile short* e, int s, unsigned long long* abort_word) {
    if (*e >= 0) ring_wait_put(e, abort_word);
    *e = (short)s;
}
// Watchdog for the turns in which a warp found neither a ray nor an event.  It must not slow the polling of an idle warp (which is
// what picks up the next event): lane 0 alone counts the idle turns in shared memory (misc[8 + warp]: one load and one store);
// every 65 536 of them it looks whether ANY list of the block moved (sum of the list tails) and, after ~1e8 idle turns (seconds)
// without a single push anywhere in the block, raises the abort word and retires the block (misc[0] = NP: every warp leaves at
// the top of its next turn).  The drain of a launch is not idle in this sense: the warps finishing the last photons keep pushing.
__device__ __noinline__ void watchdog_check(volatile int* vmisc, volatile int* vtail, int c, int np_slots, unsigned long long* abort_word) {
    const int w = threadIdx.x >> 5;
    int sig = 0;
    for (int l = 0; l < N_LISTS; ++l) sig += vtail[l];
    if (sig != vmisc[16 + w]) { vmisc[16 + w] = sig; vmisc[24 + w] = c; }
    else if (c - vmisc[24 + w] >= (1 << 27)) atomicExch(abort_word, 1ull);
    if (*(volatile unsigned long long*)abort_word != 0ull) atomicMax((int*)vmisc, np_slots);
}

template <int NT, int NP, int MINB, bool TR, bool GN, bool BT = false, bool MD = false>
__global__ void __launch_bounds__(NT, MINB) transport3_kernel(const __grid_constant__ KernelArgs A) {
    extern __shared__ double smraw[];
    const DevTables& T = A.T;
    using Sh = ShT<NP, TR, GN, BT, MD>;
    Sh X;
    block_setup<NT, NP, TR, GN, BT, MD>(A, smraw, X, true);
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    Cnt C; C.n_cf = 0; C.n_emit = C.n_sc = C.n_peel = C.n_surf = C.n_err = C.n_draw = 0;
    Marcher M; M.init(T); M.bind(X);
    volatile int* vhead = X.head;
    volatile int* vtail = X.tail;
    volatile int* vmisc = X.misc;
#ifndef E2_STARVE
#define E2_STARVE 8
#endif
#ifndef E2_INNER_SCALE
#define E2_INNER_SCALE 2       // steps per pass = nr / E2_INNER_SCALE + 2, within [6, 16]
#endif
    const int starve = E2_STARVE;                                // take partial batches when fewer lanes than this march
    // (Measured and dropped: soft warp specialisation -- the last warps of a block only run events, the others only march --
    // to shrink the code each warp loops over: 5-10 % slower on every workload, the event warps idle too often.)
#ifdef E2_STATS
    unsigned long long st_pass = 0, st_act0 = 0, st_rdy0 = 0, st_it = 0, st_lane = 0, st_ev = 0, st_rdy = 0, st_evb = 0, st_evl = 0;
    unsigned long long st_av = 0, st_lb = 0, st_ll = 0;      // per lane = per list: backlog summed over the passes, batches run, lanes in them
    bool rdy_empty = false;
#endif
    // steps per bookkeeping pass: rays are about as long as the grid has radial layers (measured best: 4-8 at nr = 2, 12 at nr = 20, 16 at nr = 100).
    // Since the peel-off walks moved into the interaction event the marchers carry transport rays only, and on 3-D grids, where a ray also ends
    // at the polar and azimuthal faces, shorter passes pay (nr = 20: 8 steps +2 % on C4, -1.7 % on the 2-D grid of C2; profiles/r02_ab_variants.txt)
    const int inner = A.L.e2_inner > 0 ? A.L.e2_inner : min(16, max(6, T.nr / (T.np > 1 ? E2_INNER_SCALE + 1 : E2_INNER_SCALE) + 2));

    for (;;) {
        // every slot retired: the block is done.  One lane reads, so that the whole warp leaves together (a volatile read per lane
        // is not warp-uniform by construction, and a warp that splits here would wait for its exited lanes in the ballots below)
        if (__shfl_sync(FULL, (int)vmisc[0], 0) >= NP) break;
        // ---- free lanes claim ready rays
        const unsigned fm = __ballot_sync(FULL, M.slot < 0);
#ifdef E2_STATS
        rdy_empty = false;
#else
        bool rdy_empty = false;
#endif
        if (fm) {
            int base = 0, n = 0;
            if (lane == 0) {
                const int want = __popc(fm);
                int h = vhead[L_RDY];
                for (;;) {
                    n = min(want, vtail[L_RDY] - h);
                    if (n <= 0) { n = 0; break; }
                    if (E2_RELEASE && n < E2_MARCH_MIN) {      // a thin pass only if no event list holds a full batch to run instead
                        bool full_event = false;
                        for (int q = 0; q < N_EVENT_LISTS; ++q) full_event = full_event || (vtail[q] - vhead[q] >= 32);
                        if (full_event) { n = 0; break; }
                    }
                    const int old = atomicCAS(X.head + L_RDY, h, h + n);
                    if (old == h) { base = h; break; }
                    h = old;
                }
            }
            base = __shfl_sync(FULL, base, 0); n = __shfl_sync(FULL, n, 0);
            rdy_empty = n < __popc(fm);
            if (E2_RELEASE) rdy_empty = (n == 0);      // (no ray is kept across passes: "starving" = this turn marched nothing)
            const int rank = __popc(fm & lt);
            if (M.slot < 0 && rank < n) {
                const int s = ring_take(&X.Q(L_RDY, base + rank), A.O.err + ERR_WATCHDOG);
                __threadfence_block();
                M.load(X, A, s);
            }
        }
        // ---- `inner` steps in a tight loop, then the write-back of what ended
        int lst = -1;
        {
            int out = (M.slot >= 0 && (M.info & 3) == K_DEAD) ? O_DEAD : O_NONE;
            unsigned n_step = 0;
#ifdef E2_STATS   // -DE2_STATS: occupancy counters of the pass structure in err slots 50-58 (tools/gpu_tune.py prints them)
            { const int na = __popc(__ballot_sync(FULL, M.slot >= 0 && out == O_NONE));
              st_pass++; st_act0 += na; st_rdy0 += rdy_empty ? 1 : 0; st_ev += (vtail[L_H]-vhead[L_H]) + (vtail[L_DEP]-vhead[L_DEP]) + (vtail[L_RES]-vhead[L_RES]); st_rdy += vtail[L_RDY]-vhead[L_RDY]; }
#endif
#pragma unroll 1
            for (int k = 0; k < inner; ++k) {
                const bool go = M.slot >= 0 && out == O_NONE;
                if (!__any_sync(FULL, go)) break;          // every ray of the warp has ended: no empty turns
#ifdef E2_STATS
                { const int na = __popc(__ballot_sync(FULL, go)); st_it += na > 0; st_lane += na; }
#endif
                if (go) out = M.step(X, A, n_step);
            }
            C.n_cf += n_step;
            if (M.slot >= 0 && out != O_NONE) lst = M.finish(X, A, C, out);
            else if (E2_RELEASE && M.slot >= 0) { M.release(X); lst = L_RDY; }
        }
        // ---- push ended rays on their event lists
        if (__any_sync(FULL, lst >= 0)) {
            __threadfence_block();
            const unsigned g = __match_any_sync(FULL, lst);
            const int leader = __ffs(g) - 1;
            int base = 0;
            if (lane == leader && lst >= 0) base = atomicAdd(X.tail + lst, __popc(g));
            base = __shfl_sync(FULL, base, leader);
            if (lst >= 0) { ring_put(&X.Q(lst, base + __popc(g & lt)), M.slot, A.O.err + ERR_WATCHDOG); M.slot = -1; }
        }
        // ---- events: a full batch if there is one; a partial one if this warp has little else to do
        int av = 0;
        if (lane < N_EVENT_LISTS) av = vtail[lane] - vhead[lane];
#ifdef E2_STATS
        st_av += (lane < N_EVENT_LISTS) ? av : (lane == L_RDY ? max(vtail[L_RDY] - vhead[L_RDY], 0) : 0);
#endif
        const unsigned fullm = __ballot_sync(FULL, av >= 32);
        const unsigned anym = __ballot_sync(FULL, av > 0);
        const int nactive = __popc(__ballot_sync(FULL, M.slot >= 0));
        int l = -1;
        if (E2_WATCHDOG >= 2 && nactive == 0 && !anym && lane == 0) {
            const int c = vmisc[8 + (threadIdx.x >> 5)] + 1;
            vmisc[8 + (threadIdx.x >> 5)] = c;
            if ((c & 0xffff) == 0) watchdog_check(vmisc, vtail, c, NP, A.O.err + ERR_WATCHDOG);
        }
        // priority: re-solves and deposits first (cheap, they hand rays straight back), then interactions
        if (Sh::MULTI) {
            // multi-detector walks: a FAN event is a full warp's work for ONE photon, so any waiting photon is taken (a few at a
            // time: the event is long); the others as usual
            if (anym & (1u << L_FAN)) l = L_FAN;
            else if (fullm)
                l = (fullm & (1u << L_RES)) ? L_RES : (fullm & (1u << L_H)) ? L_H : (fullm & (1u << L_SC)) ? L_SC
                    : (fullm & (1u << L_PRE)) ? L_PRE : L_EMIT;
            else if (anym && rdy_empty && nactive < starve)
                l = (anym & (1u << L_RES)) ? L_RES : (anym & (1u << L_H)) ? L_H : (anym & (1u << L_SC)) ? L_SC
                    : (anym & (1u << L_PRE)) ? L_PRE : L_EMIT;
        } else if (E2_EVENT_POLICY == 1) {
            // the list with the largest backlog (full batches first; partial ones only when this warp is starving).  With a fixed
            // priority order the lists at its end -- EMIT above all -- were served only when nothing else had a full batch: dead
            // slots piled up there, the block walked a fraction of its 512 photons and the lanes ran short of rays
            // (-DE2_STATS: 12.7 of 32 lanes held a ray at the start of a pass on C4, 61 % of the claims found the ready list empty).
            const int mine = (lane < N_EVENT_LISTS) ? av : 0;
            const int mx = __reduce_max_sync(FULL, mine);
            if (mx >= 32 || (mx > 0 && rdy_empty && nactive < starve)) l = __ffs(__ballot_sync(FULL, lane < N_EVENT_LISTS && av == mx)) - 1;
        } else if (fullm)
            l = (fullm & (1u << L_RES)) ? L_RES : (fullm & (1u << L_DEP)) ? L_DEP : (fullm & (1u << L_H)) ? L_H
                : (fullm & (1u << L_SURF)) ? L_SURF : (fullm & (1u << L_PRE)) ? L_PRE : L_EMIT;
        else if (anym && rdy_empty && nactive < starve)
            l = (anym & (1u << L_RES)) ? L_RES : (anym & (1u << L_DEP)) ? L_DEP : (anym & (1u << L_H)) ? L_H
                : (anym & (1u << L_SURF)) ? L_SURF : (anym & (1u << L_PRE)) ? L_PRE : L_EMIT;
        if (l >= 0) {
            int base = 0, n = 0;
            if (lane == 0) {
                const int h = vhead[l];
                n = min((Sh::MULTI && l == L_FAN) ? 8 : 32, vtail[l] - h);
                if (n > 0 && atomicCAS(X.head + l, h, h + n) == h) base = h; else n = 0;
            }
            base = __shfl_sync(FULL, base, 0); n = __shfl_sync(FULL, n, 0);
            if (n > 0) {
                const bool valid = lane < n;
                int s = 0;
                if (valid) s = ring_take(&X.Q(l, base + lane), A.O.err + ERR_WATCHDOG);
                __threadfence_block();
#ifdef E2_STATS
                st_evb++; st_evl += n;
                if (lane == l) { st_lb++; st_ll += n; }
#endif
                if (Sh::MULTI) {
                    const int tgt = run_event_md(X, A, l, n, valid, s, C);
                    __threadfence_block();
                    if (__any_sync(FULL, tgt >= 0)) {
                        const unsigned g = __match_any_sync(FULL, tgt);
                        const int leader = __ffs(g) - 1;
                        int pb = 0;
                        if (lane == leader && tgt >= 0) pb = atomicAdd(X.tail + tgt, __popc(g));
                        pb = __shfl_sync(FULL, pb, leader);
                        if (tgt >= 0) ring_put(&X.Q(tgt, pb + __popc(g & lt)), s, A.O.err + ERR_WATCHDOG);
                    }
                } else {
                    const int tgt = run_event(X, A, l, valid, s, C);
                    __threadfence_block();
                    if (E2_HEAD_STEPS > 0) {       // a head-marched ray may have ended: any list
                        if (__any_sync(FULL, tgt >= 0)) {
                            const unsigned g = __match_any_sync(FULL, tgt);
                            const int leader = __ffs(g) - 1;
                            int pb = 0;
                            if (lane == leader && tgt >= 0) pb = atomicAdd(X.tail + tgt, __popc(g));
                            pb = __shfl_sync(FULL, pb, leader);
                            if (tgt >= 0) ring_put(&X.Q(tgt, pb + __popc(g & lt)), s, A.O.err + ERR_WATCHDOG);
                        }
                    } else {
                        const bool push = tgt >= 0;
                        const unsigned pm = __ballot_sync(FULL, push);
                        if (pm) {
                            const int leader = __ffs(pm) - 1;
                            int pb = 0;
                            if (lane == leader) pb = atomicAdd(X.tail + L_RDY, __popc(pm));
                            pb = __shfl_sync(FULL, pb, leader);
                            if (push) ring_put(&X.Q(L_RDY, pb + __popc(pm & lt)), s, A.O.err + ERR_WATCHDOG);
                        }
                    }
                }
            }
        }
    }
    // block-private detector image -> the global one (every warp leaves the loop only when all slots are retired, i.e. after
    // the last deposit of the block)
    if (X.sdet) {
        __syncthreads();
        for (int i = threadIdx.x; i < A.L.sdet_doubles; i += NT) { const double v = X.sdet[i]; if (v != 0.0) atomicAdd(A.O.det + i, v); }
    }
#ifdef E2_STATS
    if (lane == 0) {
        atomicAdd(A.O.err + 50, st_pass); atomicAdd(A.O.err + 51, st_act0); atomicAdd(A.O.err + 52, st_rdy0);
        atomicAdd(A.O.err + 53, st_it); atomicAdd(A.O.err + 54, st_lane); atomicAdd(A.O.err + 55, st_ev);
        atomicAdd(A.O.err + 56, st_rdy); atomicAdd(A.O.err + 57, st_evb); atomicAdd(A.O.err + 58, st_evl);
    }
    if (lane < N_LISTS) { atomicAdd(A.O.err + 4 + lane, st_av); atomicAdd(A.O.err + 13 + lane, st_lb); atomicAdd(A.O.err + 22 + lane, st_ll); }
#endif
    flush_counters(A, C);
}

}  // namespace e2
