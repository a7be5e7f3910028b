// engine3.cuh -- the reference-order ("faithful") arithmetic on an event-list scheduler.
// Included from transport.cuh (inside namespace artes::ARTES_NS) after engine.cuh, in both translation units.
//
// The persistent-lane engine (engine.cuh, transport_kernel) keeps one photon per lane for its whole life; ncu shows 13 of 32
// lanes active and the instruction cache as the top stall, because every lane runs its own mix of cell crossings, peel-off
// weights, CDF scans and emissions.  Here the SAME event bodies -- ev_emit, ev_cross (one cell_face evaluation in the
// reference's operation order, src/ARTES.f90:2800-3470, plus the bookkeeping of the walk it belongs to), ev_peel_done,
// ev_scatter, ev_lambert and the cheap follow-ups -- are scheduled the way engine2.cuh schedules the fast mode:
//
//   * a photon lives in a SLOT: its whole state (struct Photon, 384 bytes) in an L2-resident record of the block's pool;
//   * a warp claims up to 32 slots from the READY list, loads their photons and walks them `inner` cell crossings in lock
//     step (every lane executes ev_cross: converged); the cheap follow-ups of a crossing (first optical depth :660-685,
//     survival :791-813, surface hit, retirement) run inline;
//   * a photon that reached a heavy event is stored back and its slot pushed on that event's list: EMIT (a new photon is
//     needed), DEP (a peel walk ended: weight and deposit, ev_peel_done), SC (scattering, ev_scatter), SURF (Lambert
//     reflection); whenever a list holds 32 slots a warp runs the event for all of them, converged, and pushes the slots
//     on the list their photons need next;
//   * all photons go back to the pool at the end of a pass, so a lane holds ONE photon at a time (the register budget of the
//     persistent-lane engine), and both the walk and the events always start with full warps.
//
// The arithmetic is untouched -- the event bodies are the ones the bit-exact gates certify on the persistent-lane engine
// (crossing sequences, cell_face distances) -- only the order in which photons are served changes, and with it the order of
// the floating-point sums in the image.  Covers everything the persistent-lane engine covers (thermal source, reflecting
// surface, both flow counters, oblate planets, the trace hook).

namespace e3 {

enum : int { L_EMIT = 0, L_DEP, L_SC, L_SURF, L_RDY, N_LISTS };
constexpr int ERR_WATCHDOG = 63;

__host__ __device__ constexpr int ring_cap(int np_slots) { int c = 32; while (c < np_slots) c <<= 1; return c; }

struct alignas(32) PhotonSlot { Photon p; };
constexpr int SLOT_DOUBLES = (int)(sizeof(PhotonSlot) / sizeof(double));

// shared memory of a block: the grid tables of engine.cuh (SmLayout) | lists | head[8] tail[8] misc[16]
__host__ __device__ inline size_t smem_bytes(int nr, int nt, int np, int NP) {
    const size_t tab = ((size_t)SmLayout(nr, nt, np).total * sizeof(double) + 15) / 16 * 16;
    return tab + (size_t)N_LISTS * ring_cap(NP) * sizeof(short) + 32 * sizeof(int);
}

// waits on list cells are bounded (see engine2.cuh: the same watchdog word, error slot 63)
__device__ __noinline__ int wait_take(volatile short* e, unsigned long long* abort_word) {
    unsigned spin = 0;
    for (;;) {
        const short v = *e;
        if (v >= 0) return (int)v;
        if ((++spin & 0xfffffu) != 0u) continue;
        if (*(volatile unsigned long long*)abort_word) return 0;
        if (spin >= (400u << 20)) { atomicExch(abort_word, 1ull); return 0; }
    }
}
__device__ __noinline__ void wait_put(volatile short* e, unsigned long long* abort_word) {
    unsigned spin = 0;
    while (*e >= 0) {
        if ((++spin & 0xfffffu) != 0u) continue;
        if (*(volatile unsigned long long*)abort_word) return;
        if (spin >= (400u << 20)) { atomicExch(abort_word, 1ull); return; }
    }
}
__device__ __forceinline__ int ring_take(volatile short* e, unsigned long long* abort_word) {
    int v = *e;
    if (v < 0) v = wait_take(e, abort_word);
    *e = (short)-1;
    return v;
}
__device__ __forceinline__ void ring_put(volatile short* e, int s, unsigned long long* abort_word) {
    if (*e >= 0) wait_put(e, abort_word);
    *e = (short)s;
}

__device__ __forceinline__ bool walking(int ph) { return ph == PH_PRE || ph == PH_WALK || ph == PH_PEEL; }

// the list a photon's slot goes on, by its phase (-1: the slot is retired for good)
__device__ __forceinline__ int list_of(int ph) {
    if (walking(ph)) return L_RDY;
    if (ph == PH_NEW) return L_EMIT;
    if (ph == PH_PEELDONE) return L_DEP;
    if (ph == PH_SCAT2) return L_SC;
    if (ph == PH_LAMBERT) return L_SURF;
    return -1;      // PH_IDLE
}

template <int NT, int NP, int MINB, bool TRACE, bool GEN>
__global__ void __launch_bounds__(NT, MINB) transport4_kernel(const __grid_constant__ KernelArgs A) {
    extern __shared__ double sm[];
    constexpr int RC = ring_cap(NP);
    const DevTables& T = A.T;
    const LaunchArgs& L = A.L;
    const size_t tab = ((size_t)SmLayout(T.nr, T.nt, T.np).total * sizeof(double) + 15) / 16 * 16;
    short* q = reinterpret_cast<short*>(reinterpret_cast<char*>(sm) + tab);
    int* head = reinterpret_cast<int*>(q + N_LISTS * RC);
    int* tail = head + 8;
    int* misc = head + 16;                   // [0] slots retired for good
    PhotonSlot* pool = reinterpret_cast<PhotonSlot*>(A.O.scratch) + (size_t)blockIdx.x * NP;
    const int tid = threadIdx.x, lane = tid & 31;
    const unsigned lt = (1u << lane) - 1u;
    unsigned long long* const abort_word = A.O.err + ERR_WATCHDOG;

    // ---- block set-up: every slot starts by asking for a photon
    for (int i = tid; i < N_LISTS * RC; i += NT) q[i] = (short)-1;
    for (int i = tid; i < NP; i += NT) {
        Photon P0;
        memset(&P0, 0, sizeof(P0));
        P0.ph = PH_NEW;
        pool[i].p = P0;
    }
    stage_tables(sm, T);                     // ends with a barrier
    for (int i = tid; i < NP; i += NT) q[L_EMIT * RC + i] = (short)i;
    if (tid < 8) { head[tid] = 0; tail[tid] = (tid == L_EMIT) ? NP : 0; }
    if (tid < 16) misc[tid] = 0;
    __threadfence_block();
    __syncthreads();

    const Ctx X(sm, A);
    Counters C; C.zero();
    volatile int* vhead = head;
    volatile int* vtail = tail;
    volatile int* vmisc = misc;
    auto Q = [&](int l, int pos) -> volatile short* { return q + l * RC + (pos & (RC - 1)); };
    const int inner = min(16, max(6, T.nr / 2 + 2));      // cell crossings per pass

    // push the slots of the warp's lanes on the lists their photons need next (lst < 0: nothing to push)
    auto push = [&](int lst, int s) {
        if (__any_sync(FULL, lst >= 0)) {
            __threadfence_block();                        // the photon records (global memory, same block) before the list cells
            const unsigned g = __match_any_sync(FULL, lst);
            const int leader = __ffs(g) - 1;
            int base = 0;
            if (lane == leader && lst >= 0) base = atomicAdd(tail + lst, __popc(g));
            base = __shfl_sync(FULL, base, leader);
            if (lst >= 0) ring_put(Q(lst, base + __popc(g & lt)), s, abort_word);
        }
    };
    // take up to `want` slots from list l (one CAS per warp); returns the number taken, the lane's slot in s
    auto take = [&](int l, int want, int& s) -> int {
        int base = 0, n = 0;
        if (lane == 0) {
            int h = vhead[l];
            for (;;) {
                n = min(want, vtail[l] - h);
                if (n <= 0) { n = 0; break; }
                const int old = atomicCAS(head + l, h, h + n);
                if (old == h) { base = h; break; }
                h = old;
            }
        }
        base = __shfl_sync(FULL, base, 0); n = __shfl_sync(FULL, n, 0);
        if (lane < n) s = ring_take(Q(l, base + lane), abort_word);
        if (n > 0) __threadfence_block();
        return n;
    };

    Photon P;
    for (;;) {
        if (__shfl_sync(FULL, (int)vmisc[0], 0) >= NP) break;
        // ---- a pass of the walk: claim ready photons, `inner` crossings in lock step, everything back to the pool
        int s = -1;
        const int nw = take(L_RDY, 32, s);
        if (nw > 0) {
            const bool mine = lane < nw;
            if (mine) P = pool[s].p; else P.ph = PH_IDLE;
#pragma unroll 1
            for (int k = 0; k < inner; ++k) {
                const bool go = mine && walking(P.ph);
                if (!__any_sync(FULL, go)) break;
                if (go) { ev_cross<TRACE, GEN>(X, P, C); cheap_handlers<TRACE, GEN>(X, P, C); }
            }
            int lst = -1;
            if (mine) {
                lst = list_of(P.ph);
                if (lst >= 0) pool[s].p = P; else atomicAdd(misc, 1);
            }
            push(lst, s);
        }
        // ---- events: a full batch if some list holds one; a partial one if there was nothing to walk
        int av = 0;
        if (lane < L_RDY) av = vtail[lane] - vhead[lane];
        const unsigned fullm = __ballot_sync(FULL, av >= 32);
        const unsigned anym = __ballot_sync(FULL, av > 0);
        int l = -1;
        if (fullm) l = (fullm & (1u << L_DEP)) ? L_DEP : (fullm & (1u << L_SC)) ? L_SC : (fullm & (1u << L_SURF)) ? L_SURF : L_EMIT;
        else if (anym && nw < 8) l = (anym & (1u << L_DEP)) ? L_DEP : (anym & (1u << L_SC)) ? L_SC : (anym & (1u << L_SURF)) ? L_SURF : L_EMIT;
        if (l >= 0) {
            int se = -1;
            const int n = take(l, 32, se);
            if (n > 0) {
                const bool valid = lane < n;
                if (l == L_EMIT) {
                    const unsigned vm = __ballot_sync(FULL, valid);
                    unsigned long long base = 0;
                    if (lane == 0) base = atomicAdd(A.O.counter, (unsigned long long)__popc(vm));
                    base = __shfl_sync(FULL, base, 0);
                    if (valid) {
                        P = pool[se].p;
                        const unsigned long long k = base + (unsigned long long)lane;
                        if (k >= L.n_photons) P.ph = PH_IDLE;
                        else ev_emit<TRACE, GEN>(X, P, C, k);
                    }
                } else if (valid) {
                    P = pool[se].p;
                    if (l == L_DEP) ev_peel_done<TRACE, GEN>(X, P, C);
                    else if (l == L_SC) ev_scatter<TRACE>(X, P, C);
                    else if (GEN) ev_lambert<TRACE>(X, P, C);
                }
                int lst = -1;
                if (valid) {
                    lst = list_of(P.ph);
                    if (lst >= 0) pool[se].p = P; else atomicAdd(misc, 1);
                }
                push(lst, se);
            }
        } else if (nw == 0) {
            // nothing to walk and no event: an idle turn; leave if a bounded wait somewhere raised the abort word
            if (__shfl_sync(FULL, (int)(*(volatile unsigned long long*)abort_word != 0ull), 0)) break;
        }
    }
    C.flush(A.O.stats);
}

}  // namespace e3
