// kernel_args.h -- POD blocks passed from the host C-ABI (artes_gpu.cu) to the sm_100a kernels.
#pragma once
#include <stdint.h>

namespace artes {

// Per-cell / per-grid tables resident in HBM (uploaded by artes_gpu_set_grid / set_wavelength).
// They replace the program-scope arrays of src/ARTES.f90:58-91.
struct DevTables {
    int nr, nt, np, cells, cell_depth, n_uniq;
    double ox, oy, oz;               // oblate_x/y/z (:42)
    double inv_ox, inv_oy, inv_oz;   // their reciprocals (a, b, c of :2838-2840)
    const double* rfront;            // [nr+1]           (:58)
    const double* thetafront;        // [nt+1] rad       (:59)
    const double* ttan;              // [nt+1] theta_grid_tan (:77)
    const double* tcos;              // [nt+1] theta_grid_cos (:76)
    const int*    tplane;            // [nt+1] thetaplane (:60)
    const double* phifront;          // [np]             (:61)
    const double* psin;              // [np] phi_grid_sin (:78)
    const double* pcos;              // [np] phi_grid_cos (:79)
    const double* trig;              // sinbeta[180] | cos2beta[180] | sin2beta[180]  (:88-89, first half)
    const double* kext;              // [cells] cell_opacity (:66)
    const double* albedo;            // [cells] cell_albedo  (:70)
    const int*    c2u;               // [cells] cell -> unique matrix block
    const double* M;                 // [n_uniq][180][16] cell_scatter_matrix, de-duplicated (:69)
    const double* Mrow;              // [n_uniq][180][4]  first matrix row, compact (polar CDF loop)
    const double* Mc;                // [n_uniq][180][8]  {F11 F12 F21 F22 F33 F34 F43 F44} when EVERY block has exact zeros in the two off-diagonal
                                     // 2x2 quarters (spheres, Rayleigh, Henyey-Greenstein: all the reference's opacity tools), else null
    const double* p1k;               // [n_uniq][4] cell_p11..p14_int (:72-75)
    const double* cdfA;              // fast mode: prefix sums of cos2beta | sin2beta, [2][181]
    const double* cdfA2;             // the same, interleaved [181][2] (staged in shared memory by the ray/event engine)
    const double* cellrec;           // [cells][4]: cell_opacity, cell_albedo, unique-matrix index (as int64 bits), 0 -- one 256-bit read per interaction
    const double* cdfP;              // fast mode: [n_uniq][181][4] prefix sums of P1k(i)*sinbeta(i)*pi/180
    const double* cell_weight;       // [cells] or null (:68)
    const double* emis_cdf;          // [(nr-cell_depth)*nt*np] emissivity_cumulative in the (i,j,k) loop order of :2425-2427
};

// Scalars of one launch (artes_launch_t + host-derived detector geometry, :495-502).
struct LaunchArgs {
    unsigned long long n_photons, id_base, seed;
    int photon_source, photon_scattering, photon_emission, stellar_direction, limb_emission;
    int flow_global, flow_theta, nx, ny;
    int defer_events, defer_refill;   // ballot-regrouping thresholds (lanes), persistent-lane engine
    int e2_trips, e2_pad;             // ray/event engine: marcher steps per round, block shape
    int e2_inner, e2_pad2;            // steps per bookkeeping pass
    double fstop, photon_minimum, photon_bias, surface_albedo, theta_star, phi_star;
    double x_max, y_max;
    double det[3];                   // det_dir(1:3)
    double sin_dt, cos_dt, sin_dp, cos_dp;
    double det_atan2;                // atan2(det(2),det(1)) folded into [0,2pi] (:4871-4874)
    double det_sph_theta, det_sph_phi;  // cartesian_spherical(det) used by peel_surface (:4628)
    // star rotation (:1080-1109), host-evaluated trigonometry
    double rot_y_cos, rot_y_sin, rot_z_cos, rot_z_sin, star_dir[3];
    // batched launches (artes_gpu_run_batch; ray/event engine only): n_photons = n_batch * per_launch work items, item w
    // belongs to launch (id_base + w - batch_base) / per_launch; geo = [n_batch][12] {det(3), sin_dt, cos_dt, sin_dp, cos_dp,
    // limb_emission, det_sph_theta, det_sph_phi, cell_depth, wavelength index} in device memory (null for a single launch: the fields above are used).  Images and fluxes of the
    // launches follow each other in DevOutputs::det ([n_batch][10][ny][nx]) and ::flux ([n_batch][2]).
    unsigned long long per_launch, batch_base;
    const double* geo;
    int n_batch;
    int wl_batch;                    // 1: the launches of the batch use different wavelengths: kext / cellrec are [n_wl][cells], geo[k][10..11] = cell_depth, wavelength index
    int multi;                       // 1: ONE walk of n_photons packets peels off towards all n_batch detectors of geo (artes_gpu_run_multi)
    int sdet_doubles;                // > 0: the detector images ([n_batch][10][ny][nx] doubles) are accumulated per block in shared memory and added to det at the end
};

// Device accumulators of one launch.
struct DevOutputs {
    double* det;                     // [10][ny][nx]: sumI,Q,U,V | sqI,Q,U,V | nI | nQUV
    double* flux;                    // [2] flux_emitted, flux_exit (:86-87)
    double* flow4;                   // [cells][4] or null (:82)
    double* flow3;                   // [cells][3] or null (:81)
    unsigned long long* err;         // [64]
    unsigned long long* stats;       // [8]: emit, cell_face, scatter, peel, surface, draws, error, -
    unsigned long long* counter;     // work counter (photon ids handed out)
    double* scratch;                 // ray/event engine: cold photon records, blocks x slots x 160 B
};

// Injected-stream walk trace (test hook).
struct TraceArgs {
    const double* xi;                // [n][max_draws]
    int max_draws, max_rec;
    int* seq_len;                    // [n]
    unsigned long long* seq_hash;    // [n]
    int* seq_head;                   // [n][max_rec][5] or null
    double* fstate;                  // [n][8] or null
};

struct KernelArgs {
    DevTables T;
    LaunchArgs L;
    DevOutputs O;
    TraceArgs R;
};

// Host-side launchers, one pair per arithmetic mode (separate translation units: the faithful one is
// compiled with -fmad=false).

}  // namespace artes
