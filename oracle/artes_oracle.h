/*
 * artes_oracle.h -- C API of the CPU oracle (TEST INFRASTRUCTURE, not product code).
 *
 * PARITY UNPINNED: the reference (bgin/ARTES, src/ARTES.f90) ships no tests, no golden
 * vectors and seeds its generator from the clock (src/ARTES.f90:4187-4191), and no Fortran
 * compiler exists in this image, so this restatement could not be checked against the
 * reference binary.  It is pinned only by line-by-line review against the cited ranges and
 * by analytic anchors (tests/test_oracle_*.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library.
 */
#ifndef ARTES_ORACLE_H
#define ARTES_ORACLE_H

#include "../include/artes_gpu.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct artes_ref_ctx artes_ref_ctx;

#define ARTES_REF_RNG_MZ       0 /* Marsaglia-Zaman generator of the reference (src/ARTES.f90:4197-4230) */
#define ARTES_REF_RNG_PHILOX   1 /* the product's Philox4x32-10 stream, one per photon id                */

artes_ref_ctx* artes_ref_create(void);
void artes_ref_destroy(artes_ref_ctx* c);

int artes_ref_set_grid(artes_ref_ctx* c, int nr, int ntheta, int nphi,
                       const double* rfront, const double* thetafront, const int32_t* thetaplane,
                       const double* phifront, double oblate_x, double oblate_y, double oblate_z);

int artes_ref_set_wavelength(artes_ref_ctx* c, const double* k_sca, const double* k_abs,
                             int n_uniq, const double* uniq_matrix, const int32_t* cell_to_uniq,
                             int cell_depth, const double* cell_weight, const double* emis_cdf);

/* radiative_transfer (src/ARTES.f90:518-975, before the package_energy scaling).
 * nthreads<=0: all cores.  emulate_stat!=0 keeps the stat() syscall per photon and per
 * cell_face call of the reference (:549,:2820).  stats->kernel_ms = wall time of the
 * photon loop. */
int artes_ref_run(artes_ref_ctx* c, const artes_launch_t* launch, int rng_kind, int nthreads,
                  int emulate_stat, double* det_sum, double* flux, double* flow4, double* flow3,
                  uint64_t* err_hist, artes_stats_t* stats);

/* Same contract as artes_gpu_trace. */
int artes_ref_trace(artes_ref_ctx* c, const artes_launch_t* launch, const double* xi, uint64_t n,
                    int max_draws, int32_t* seq_len, uint64_t* seq_hash, int32_t* seq_head,
                    int max_rec, double* fstate);

/* Same contract as artes_gpu_cell_face. */
int artes_ref_cell_face(artes_ref_ctx* c, uint64_t n, const double* pos, const double* dir,
                        const int32_t* face, const int32_t* cell, int32_t* out_i, double* out_d);

/* One scattering event in isolation (src/ARTES.f90:1434-1661 + :1663-1932):
 *   stokes[n][4], dir[n][3], cell_idx[n], xi[n][3] -> out[n][9] = alpha, beta, dir_new(3), stokes_new(4) */
int artes_ref_scatter(artes_ref_ctx* c, uint64_t n, const double* stokes, const double* dir,
                      const int32_t* cell_idx, const double* xi, double* out);

/* Philox4x32-10 block (known-answer tests). */
void artes_ref_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
/* The first `n` uniforms photon `id` sees under `seed` (must equal the device stream). */
void artes_ref_philox_uniforms(uint64_t seed, uint64_t id, int n, double* out);
/* The first `n` uniforms of the reference generator for seed word s1. */
void artes_ref_mz_uniforms(int32_t s1, int n, double* out);

/* Host-side derived arrays the reference computes in grid_initialize(2) (src/ARTES.f90:2329-2453);
 * restated here so that tests can check the product's host helpers. */
int artes_ref_cell_depth(const artes_ref_ctx* c, int photon_source, int ring);

/* Host side of the path, restated for checks of the driver / Python mirror:
 * planck_function (:1350-1367), photon_package (:2509-2539), the tail of radiative_transfer (:957-1004: detector from the
 * thread sums, photometry(11)) and the error planes of write_output (:3481-3519).  Arrays in the reference's order
 * detector(nx,ny,4,3), error(nx,ny,5). */
double artes_ref_planck(double temperature, double wavelength, int photon_source);
double artes_ref_package_energy(int photon_source, double t_star, double r_star, double orbit, double distance_planet, double rfront_nr,
                                double wavelength, double packages, int phase_curve, double det_phi, double emissivity_total);
void artes_ref_finish_detector(int nx, int ny, const double* det_sum, double package_energy, double* detector, double* photometry);
void artes_ref_stokes_error(int nx, int ny, const double* detector, double* error);

#ifdef __cplusplus
}
#endif
#endif
