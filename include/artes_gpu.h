/*
 * artes_gpu.h -- C-ABI of libartes_gpu: the B200 (sm_100a) replacement for the
 * photon-packet transport loop of ARTES.
 *
 * The reference has no FFI; its seam is the argument-less internal subroutine
 * `radiative_transfer` (src/ARTES.f90:518-1006) which `run` calls once per
 * wavelength / detector azimuth (src/ARTES.f90:146,185,241,255) and which talks
 * to the rest of the program through program-scope variables
 * (src/ARTES.f90:19-115; the OpenMP clause :534-544 lists exactly what crosses).
 * Every entry point below names the reference state it replaces.
 *
 * Conventions
 *  - plain C, fixed-width integers, IEEE doubles, caller owns all host buffers;
 *  - all 3-D cell arrays are in the reference's Fortran order, radial index
 *    fastest: idx = ir + nr*(itheta + ntheta*iphi)   (src/ARTES.f90:64-70);
 *  - every function returns 0 on success, <0 on error
 *    (artes_gpu_last_error() gives the text); nothing exits or throws across
 *    the ABI (the reference calls exit(0), src/ARTES.f90:554,2825);
 *  - there is NO CPU fallback: without a CUDA device artes_gpu_create fails.
 */
#ifndef ARTES_GPU_H
#define ARTES_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ARTES_GPU_ABI_VERSION 3 /* 2: artes_gpu_run_batch, artes_gpu_set_wavelengths, artes_launch_t::wl_index (was reserved0); 3: artes_gpu_set_wavelength_dense_wl, artes_gpu_run_multi */

/* arithmetic modes */
#define ARTES_MODE_FAITHFUL 0 /* reference operation order, no FMA contraction, sequential 180-bin CDFs */
#define ARTES_MODE_FAST     1 /* same physics; prefix-table CDF inversion, FMA allowed                   */

#define ARTES_ERR_SLOTS 64 /* err_hist[code]: code = the reference's `error NNN` number (src/ARTES.f90, ~50 sites) */

typedef struct artes_gpu_ctx artes_gpu_ctx;

/* Per-launch parameters = the scalar program-scope inputs of radiative_transfer. */
typedef struct artes_launch_t {
    uint32_t struct_size;       /* sizeof(artes_launch_t), ABI check                                   */
    int32_t  mode;              /* ARTES_MODE_*                                                        */
    uint64_t n_photons;         /* `packages` handled by THIS call (src/ARTES.f90:26,546)              */
    uint64_t photon_id_base;    /* global id of the first photon (Philox counter; multi-GPU sharding)  */
    uint64_t seed;              /* Philox key; replaces the clock-seeded state(:,:) (:433-449)         */
    int32_t  photon_source;     /* 1 = star, 2 = planet (:20)                                          */
    int32_t  photon_scattering; /* (:32)                                                               */
    int32_t  photon_emission;   /* 1 isotropic, 2 biased (:33)                                         */
    int32_t  stellar_direction; /* (:37)                                                               */
    int32_t  limb_emission;     /* phase_curve .and. det_phi*180/pi >= 170 (:1041)                     */
    int32_t  flow_global;       /* (:53)                                                               */
    int32_t  flow_theta;        /* (:54)                                                               */
    int32_t  nx, ny;            /* detector pixels (:48-49)                                            */
    int32_t  wl_index;          /* wavelength of the launch: index into the tables of artes_gpu_set_wavelengths (0 otherwise) */
    double   fstop;             /* (:29)                                                               */
    double   photon_minimum;    /* (:30)                                                               */
    double   photon_bias;       /* (:34)                                                               */
    double   surface_albedo;    /* (:40)                                                               */
    double   theta_star;        /* [rad] (:38)                                                         */
    double   phi_star;          /* [rad] (:39)                                                         */
    double   det_theta;         /* [rad] det_dir(4) (:91,496)                                          */
    double   det_phi;           /* [rad] det_dir(5) (:91,497)                                          */
    double   x_max, y_max;      /* image half-size [m] (:50-51,475-479)                                */
} artes_launch_t;

/* Exact event counters of one launch (the flop accounting uses them). */
typedef struct artes_stats_t {
    uint64_t n_emit;       /* photons emitted                       */
    uint64_t n_cell_face;  /* cell_face evaluations (all walks)     */
    uint64_t n_scatter;    /* scatter_photon calls                  */
    uint64_t n_peel;       /* peel_photon + peel_surface + peel_thermal walks started */
    uint64_t n_surface;    /* surface hits                          */
    uint64_t n_draws;      /* random numbers consumed               */
    uint64_t n_error;      /* photons dropped through an error path */
    uint64_t reserved;     /* kernels launched by the call (all devices of the context) */
    double   kernel_ms;    /* CUDA-event time of the transport kernel(s), max over this ctx's devices */
    double   reduce_ms;    /* CUDA-event time of the NCCL reduce (0 if single device)                 */
    double   h2d_ms;       /* last table upload                                                        */
    double   d2h_ms;       /* result download                                                          */
} artes_stats_t;

/* ---- lifetime ------------------------------------------------------------------------- */

/* One context owns `ndev` devices (dev_ids may be NULL = 0..ndev-1) of this process; the
 * OpenMP team of the reference (src/ARTES.f90:360-363) becomes these devices. */
int  artes_gpu_create(artes_gpu_ctx** ctx, int ndev, const int* dev_ids);
int  artes_gpu_destroy(artes_gpu_ctx* ctx);
const char* artes_gpu_last_error(const artes_gpu_ctx* ctx); /* ctx may be NULL: last create error */
int  artes_gpu_abi_version(void);

/* ---- static inputs (per run) ------------------------------------------------------------ */

/* Replaces rfront/thetafront/thetaplane/phifront (src/ARTES.f90:58-61, filled at :2071-2122)
 * and oblate_x/y/z (:42,469-471). The derived tables theta_grid_cos/tan, phi_grid_sin/cos
 * (:2261-2270) and sinbeta/cos2beta/sin2beta (:409-420) are rebuilt inside, on the host,
 * in double, with the same expressions.
 *   rfront[nr+1] [m], thetafront[ntheta+1] [rad], thetaplane[ntheta+1] (1 cone, 2 plane),
 *   phifront[nphi] [rad] */
int  artes_gpu_set_grid(artes_gpu_ctx* ctx, int nr, int ntheta, int nphi,
                        const double* rfront, const double* thetafront, const int32_t* thetaplane,
                        const double* phifront, double oblate_x, double oblate_y, double oblate_z);

/* ---- per-wavelength inputs ---------------------------------------------------------------- */

/* Replaces the wl_count slices of cell_scattering_opacity / cell_absorption_opacity
 * (src/ARTES.f90:64-65) and cell_scatter_matrix (:69) plus the host-derived cell_opacity,
 * cell_albedo (:2172-2188), cell_p11..p14_int (:2203-2230), cell_depth (:2329-2393),
 * cell_weight and emissivity_cumulative (:2395-2453).
 *   k_sca, k_abs       [cells] [m^-1]
 *   uniq_matrix        [n_uniq][180][16]  (angle bin, then element 4*row+col; = one cell's
 *                      block of HDU 8 of atmosphere.fits)
 *   cell_to_uniq       [cells] index into uniq_matrix
 *   cell_depth         surface face index
 *   cell_weight, emis_cdf [cells] or NULL (only photon_source = 2) */
int  artes_gpu_set_wavelength(artes_gpu_ctx* ctx, const double* k_sca, const double* k_abs,
                              int n_uniq, const double* uniq_matrix, const int32_t* cell_to_uniq,
                              int cell_depth, const double* cell_weight, const double* emis_cdf);

/* The tables of n_wl wavelengths at once (the wl_count loop of the spectrum / broadband modes, src/ARTES.f90:132-204):
 * k_sca, k_abs, cell_to_uniq are [n_wl][cells], cell_to_uniq indexes ONE common list of n_uniq matrix blocks,
 * cell_depths[n_wl].  A launch picks its wavelength with artes_launch_t::wl_index; launches of different wavelengths
 * can then share one batched kernel launch (artes_gpu_run_batch).  cell_weight, emis_cdf: [n_wl][cells] for the
 * thermal source (as in artes_gpu_set_wavelength, per wavelength) or NULL. */
int  artes_gpu_set_wavelengths(artes_gpu_ctx* ctx, int n_wl, const double* k_sca, const double* k_abs,
                               int n_uniq, const double* uniq_matrix, const int32_t* cell_to_uniq,
                               const int32_t* cell_depths, const double* cell_weight, const double* emis_cdf);

/* Same, taking the reference's dense array for ONE wavelength exactly as it sits in memory
 * after ftgpvd (src/ARTES.f90:2196-2198): element (cell, e, a) at cell + cells*(e + 16*a),
 * e = 0..15, a = 0..179.  De-duplicated internally (hash of each cell's 2880 doubles). */
int  artes_gpu_set_wavelength_dense(artes_gpu_ctx* ctx, const double* k_sca, const double* k_abs,
                                    const double* matrix_dense, int cell_depth,
                                    const double* cell_weight, const double* emis_cdf);

/* Same for wavelength `wl_index` (0-based) out of the reference's WHOLE program-scope arrays, without a slice copy:
 * cell_scattering_opacity / cell_absorption_opacity (cells, n_wl) (src/ARTES.f90:64-65) and
 * cell_scatter_matrix(cells, n_wl, 16, 180) (:69): element (cell, wl, e, a) at cell + cells*(wl + n_wl*(e + 16*a)).
 * (A Fortran slice cell_scatter_matrix(:,:,:,wl,:,:) is not contiguous for n_wl > 1; passing it to the entry above makes
 * the compiler copy 16.6 GB at the scale configuration.)  The 2880 (e, a) planes of the wavelength are streamed to the
 * device, hashed per cell there and de-duplicated (get_atmosphere :2054-2235 reads them cell by cell on one core);
 * equal hashes are verified element by element on the device before two cells share a block. */
int  artes_gpu_set_wavelength_dense_wl(artes_gpu_ctx* ctx, int n_wl, int wl_index, const double* k_sca_all,
                                       const double* k_abs_all, const double* matrix_all, int cell_depth,
                                       const double* cell_weight, const double* emis_cdf);

/* ---- the hot path --------------------------------------------------------------------------- */

/* Replaces `call radiative_transfer` up to and excluding the package_energy scaling
 * (src/ARTES.f90:546-955 and the thread sum of :959-975).
 *   det_sum  [nx*ny*4*3], Fortran order detector(ix,iy,stokes,l): l=1 sum W, l=2 sum W^2,
 *            l=3 counts (:84-85); UNscaled (the driver applies package_energy, :957-975)
 *   flux     [2] = sum(flux_emitted), sum(flux_exit) (:86-87,607,780,953)
 *   flow4    [4*cells] cell_flow summed over threads, (dir, cell) dir fastest (:82) or NULL
 *   flow3    [3*cells] cell_flow_global likewise (:81) or NULL
 *   err_hist [ARTES_ERR_SLOTS] per-code error counts (replaces error.log appends) or NULL
 * With several devices (or after artes_gpu_nccl_init_rank) the outputs are the NCCL sum
 * over all devices/ranks; every rank receives the sum. */
int  artes_gpu_run(artes_gpu_ctx* ctx, const artes_launch_t* launch,
                   double* det_sum, double* flux, double* flow4, double* flow3,
                   uint64_t* err_hist, artes_stats_t* stats);

/* Asynchronous form: enqueue on the context's stream(s) and return; artes_gpu_wait copies
 * the results out.  Lets the driver overlap launch k's reduce/download with launch k+1's
 * host work (phase curve: 73 launches, src/ARTES.f90:215-245). */
int  artes_gpu_run_async(artes_gpu_ctx* ctx, const artes_launch_t* launch);
int  artes_gpu_wait(artes_gpu_ctx* ctx, double* det_sum, double* flux, double* flow4, double* flow3,
                    uint64_t* err_hist, artes_stats_t* stats);

/* Several launches that share every parameter except the detector direction (det_theta, det_phi,
 * limb_emission) and the wavelength (wl_index, after artes_gpu_set_wavelengths: the spectrum loop :132-165)
 * as ONE kernel launch: the phase-curve loop of the reference (src/ARTES.f90:215-245,
 * 73 calls of radiative_transfer, one per det_phi) pays the drain of a launch once instead of 73 times.
 * Launch k uses the photon ids launches[0].photon_id_base + k*n_photons + [0, n_photons), i.e. it equals
 * artes_gpu_run with that photon_id_base (the launches are statistically independent).
 *   det_sum [n][nx*ny*4*3], flux [n][2] (either may be NULL); err_hist / stats: sums over the batch.
 * Flow counters are not available in a batch (error). Where the batched kernel does not apply
 * (faithful mode, oblate planets) the launches run one after the other with the same results. */
#define ARTES_MAX_BATCH 256
int  artes_gpu_run_batch(artes_gpu_ctx* ctx, const artes_launch_t* launches, int n,
                         double* det_sum, double* flux, uint64_t* err_hist, artes_stats_t* stats);

/* ONE random walk observed by n detectors.  The reference repeats the whole walk for every detector azimuth of a
 * phase curve (src/ARTES.f90:215-245); peel-off (peel_photon :4710-4990) is a next-event estimate that does not disturb
 * the walk, so a single walk of launches[0].n_photons packets can peel off towards ALL n detector directions at every
 * scattering.  Every detector receives exactly the deposits the reference's run with that det_theta / det_phi would make
 * along this walk: image k has the expectation value and the per-image noise of artes_gpu_run with launches[k] and the
 * same n_photons, but the n images share the walks, i.e. their noise is CORRELATED between detectors instead of
 * independent (a phase curve comes out smooth; error bars per point stay valid, chi^2 over points does not).
 * The launches may differ in det_theta / det_phi only; limb_emission (a different emission law, :1041-1055) and wl_index
 * must be equal, so a phase curve takes two calls: the angles below 170 deg and the limb-biased ones.
 * Same outputs as artes_gpu_run_batch.  Star source over a black surface in fast mode; anything else (thermal source,
 * reflecting surface, faithful mode, oblate planet) runs as artes_gpu_run_batch, i.e. with independent walks. */
int  artes_gpu_run_multi(artes_gpu_ctx* ctx, const artes_launch_t* launches, int n,
                         double* det_sum, double* flux, uint64_t* err_hist, artes_stats_t* stats);

/* ---- multi-process NCCL (one process per GPU, e.g. under torchrun / MPI) ---------------------- */

#define ARTES_NCCL_ID_BYTES 128
int  artes_gpu_nccl_unique_id(void* id_out /* ARTES_NCCL_ID_BYTES */);
int  artes_gpu_nccl_init_rank(artes_gpu_ctx* ctx, int nranks, int rank, const void* id);

/* ---- test hooks --------------------------------------------------------------------------------- */

/* Walk `n` photons with an INJECTED random stream xi[n][max_draws] (draw order of SURVEY
 * App. C) and report every cell_face outcome in program order (tau pre-pass, walks, peel
 * walks): tuple (next_face(1), next_face(2), cell_out(1..3)); a detector deposit is the
 * tuple (100, ix, iy, 0, 0).
 *   seq_len  [n]   number of tuples of photon i
 *   seq_hash [n]   FNV-1a-64 over all tuple words
 *   seq_head [n][max_rec][5] first max_rec tuples (may be NULL / max_rec = 0)
 *   fstate   [n][8] final x,y,z,I,Q,U,V,n_scatter (may be NULL)
 * Deterministic: single device, no atomics on the outputs. */
int  artes_gpu_trace(artes_gpu_ctx* ctx, const artes_launch_t* launch,
                     const double* xi, uint64_t n, int max_draws,
                     int32_t* seq_len, uint64_t* seq_hash, int32_t* seq_head, int max_rec,
                     double* fstate);

/* Batch of isolated cell_face evaluations (src/ARTES.f90:2800-3470):
 *   pos[n][3], dir[n][3], face[n][2], cell[n][3]  ->
 *   out_i[n][7] = next_face(2), cell_out(3), grid_exit, cell_error ; out_d[n] = face_distance */
int  artes_gpu_cell_face(artes_gpu_ctx* ctx, int mode, uint64_t n, const double* pos, const double* dir,
                         const int32_t* face, const int32_t* cell, int32_t* out_i, double* out_d);

/* Device properties of the context's first device (bench/roofline bookkeeping). */
int  artes_gpu_device_info(const artes_gpu_ctx* ctx, int* sm_count, int* cc_major, int* cc_minor,
                           char* name, int name_len);

/* Which transport engine the last run / trace of this context used: 2 = ray/event engine (fast mode), 3 = event-list engine
 * around the reference-order event bodies (faithful mode; fast mode with flow_global or an oblate planet), 1 = persistent-lane
 * engine (tuning builds only).  Lets tests assert that the production path is the one that ran. */
int  artes_gpu_last_engine(const artes_gpu_ctx* ctx);

/* Test hook for the dense-matrix ingest: planes > 0 makes artes_gpu_set_wavelength_dense[_wl] stream the 2880 (element, angle)
 * planes in chunks of that many (two passes: hash, verify), the path taken when a wavelength's planes do not fit into free HBM;
 * 0 restores the automatic choice (resident planes whenever they fit). */
int  artes_gpu_test_ingest_chunk(int planes);

/* Test hook for the scattering-matrix tables: block-diagonal matrices (exact zeros in the two off-diagonal 2x2 quarters of every
 * block -- what python/opacity*.py write) are read by the interaction event from an eight-element copy; on != 0 makes the
 * following artes_gpu_set_wavelength* calls keep the 16-element path for them as well, so that a test can compare the two. */
int  artes_gpu_test_full_matrix(int on);

/* FP64 / FP32 FMA peak microbenchmark (roofline denominator): returns TFLOP/s. */
int  artes_gpu_fma_peak(artes_gpu_ctx* ctx, double* fp64_tflops, double* fp32_tflops);

#ifdef __cplusplus
}
#endif
#endif /* ARTES_GPU_H */
