! artes_gpu_mod.f90 -- iso_c_binding interface of libartes_gpu (include/artes_gpu.h).
!
! This is the binding a maintainer of the reference adds to src/ARTES.f90 so that `call radiative_transfer`
! (src/ARTES.f90:146,185,241,255) runs on the GPU: every procedure below is the C function of the same name.
! No Fortran compiler exists in the build image of this repository, so this file is shipped as source and is
! compiled only when $(FC) is available (fortran/Makefile); the driver that is built and tested is the C++
! one (src/host/artes_main.cc), which makes the same calls in the same order.
module artes_gpu_mod

  use, intrinsic :: iso_c_binding
  implicit none

  integer(c_int), parameter :: ARTES_MODE_FAITHFUL = 0, ARTES_MODE_FAST = 1, ARTES_ERR_SLOTS = 64

  ! artes_launch_t: the scalar program-scope inputs of radiative_transfer (src/ARTES.f90:19-55, 91-96)
  type, bind(c) :: artes_launch_t
     integer(c_int32_t) :: struct_size
     integer(c_int32_t) :: mode
     integer(c_int64_t) :: n_photons          ! packages
     integer(c_int64_t) :: photon_id_base
     integer(c_int64_t) :: seed
     integer(c_int32_t) :: photon_source, photon_scattering, photon_emission, stellar_direction
     integer(c_int32_t) :: limb_emission, flow_global, flow_theta, nx, ny, wl_index
     real(c_double)     :: fstop, photon_minimum, photon_bias, surface_albedo
     real(c_double)     :: theta_star, phi_star, det_theta, det_phi, x_max, y_max
  end type artes_launch_t

  type, bind(c) :: artes_stats_t
     integer(c_int64_t) :: n_emit, n_cell_face, n_scatter, n_peel, n_surface, n_draws, n_error, reserved
     real(c_double)     :: kernel_ms, reduce_ms, h2d_ms, d2h_ms
  end type artes_stats_t

  interface

     integer(c_int) function artes_gpu_create(ctx, ndev, dev_ids) bind(c, name="artes_gpu_create")
       import :: c_ptr, c_int
       type(c_ptr), intent(out)   :: ctx
       integer(c_int), value      :: ndev
       type(c_ptr), value         :: dev_ids          ! c_null_ptr: devices 0..ndev-1
     end function artes_gpu_create

     integer(c_int) function artes_gpu_destroy(ctx) bind(c, name="artes_gpu_destroy")
       import :: c_ptr, c_int
       type(c_ptr), value :: ctx
     end function artes_gpu_destroy

     type(c_ptr) function artes_gpu_last_error(ctx) bind(c, name="artes_gpu_last_error")
       import :: c_ptr
       type(c_ptr), value :: ctx
     end function artes_gpu_last_error

     ! rfront(0:nr) [m], thetafront(0:ntheta) [rad], thetaplane(0:ntheta), phifront(0:nphi-1) [rad]  (:58-61)
     integer(c_int) function artes_gpu_set_grid(ctx, nr, ntheta, nphi, rfront, thetafront, thetaplane, phifront, &
          oblate_x, oblate_y, oblate_z) bind(c, name="artes_gpu_set_grid")
       import :: c_ptr, c_int, c_double, c_int32_t
       type(c_ptr), value             :: ctx
       integer(c_int), value          :: nr, ntheta, nphi
       real(c_double), intent(in)     :: rfront(*), thetafront(*), phifront(*)
       integer(c_int32_t), intent(in) :: thetaplane(*)
       real(c_double), value          :: oblate_x, oblate_y, oblate_z
     end function artes_gpu_set_grid

     ! cell_scattering_opacity(:,:,:,wl), cell_absorption_opacity(:,:,:,wl) and cell_scatter_matrix(:,:,:,wl,:,:) exactly
     ! as they sit in memory (cell index fastest, :64-69); de-duplicated inside the library
     integer(c_int) function artes_gpu_set_wavelength_dense(ctx, k_sca, k_abs, matrix_dense, cell_depth, cell_weight, emis_cdf) &
          bind(c, name="artes_gpu_set_wavelength_dense")
       import :: c_ptr, c_int, c_double
       type(c_ptr), value         :: ctx
       real(c_double), intent(in) :: k_sca(*), k_abs(*), matrix_dense(*)
       integer(c_int), value      :: cell_depth
       type(c_ptr), value         :: cell_weight, emis_cdf      ! c_null_ptr unless photon_source = 2
     end function artes_gpu_set_wavelength_dense

     ! the same for wavelength wl_index (0-based) out of the WHOLE program-scope arrays cell_scattering_opacity(:,:,:,:),
     ! cell_absorption_opacity(:,:,:,:), cell_scatter_matrix(:,:,:,:,:,:) -- no slice, no copy-in temporary; the
     ! matrix blocks are de-duplicated on the GPU
     integer(c_int) function artes_gpu_set_wavelength_dense_wl(ctx, n_wl, wl_index, k_sca_all, k_abs_all, matrix_all, cell_depth, &
          cell_weight, emis_cdf) bind(c, name="artes_gpu_set_wavelength_dense_wl")
       import :: c_ptr, c_int, c_double
       type(c_ptr), value         :: ctx
       integer(c_int), value      :: n_wl, wl_index, cell_depth
       real(c_double), intent(in) :: k_sca_all(*), k_abs_all(*), matrix_all(*)
       type(c_ptr), value         :: cell_weight, emis_cdf      ! c_null_ptr unless photon_source = 2
     end function artes_gpu_set_wavelength_dense_wl

     ! det_sum(nx,ny,4,3) = the thread sum of detector_thread (:959-975, before the package_energy scaling)
     integer(c_int) function artes_gpu_run(ctx, launch, det_sum, flux, flow4, flow3, err_hist, stats) bind(c, name="artes_gpu_run")
       import :: c_ptr, c_int, c_double, c_int64_t, artes_launch_t, artes_stats_t
       type(c_ptr), value               :: ctx
       type(artes_launch_t), intent(in) :: launch
       real(c_double), intent(out)      :: det_sum(*), flux(2)
       type(c_ptr), value               :: flow4, flow3           ! cell_flow / cell_flow_global sums or c_null_ptr
       integer(c_int64_t), intent(out)  :: err_hist(*)
       type(artes_stats_t), intent(out) :: stats
     end function artes_gpu_run

     ! all wl_count slices at once (k_sca, k_abs, cell_to_uniq: (cells, n_wl); one common list of matrix blocks); a launch
     ! then picks its wavelength with launch%wl_index and the wavelength loop (:132-204) can be one artes_gpu_run_batch
     integer(c_int) function artes_gpu_set_wavelengths(ctx, n_wl, k_sca, k_abs, n_uniq, uniq_matrix, cell_to_uniq, cell_depths, &
          cell_weight, emis_cdf) bind(c, name="artes_gpu_set_wavelengths")
       import :: c_ptr, c_int, c_double, c_int32_t
       type(c_ptr), value             :: ctx
       integer(c_int), value          :: n_wl, n_uniq
       real(c_double), intent(in)     :: k_sca(*), k_abs(*), uniq_matrix(*)
       integer(c_int32_t), intent(in) :: cell_to_uniq(*), cell_depths(*)
       type(c_ptr), value             :: cell_weight, emis_cdf   ! (cells, n_wl) for the thermal source, else c_null_ptr
     end function artes_gpu_set_wavelengths

     ! the phase-curve loop (:215-245) as one launch: launches(n) differ in det_theta / det_phi / limb_emission only;
     ! det_sum(nx,ny,4,3,n), flux(2,n)
     integer(c_int) function artes_gpu_run_batch(ctx, launches, n, det_sum, flux, err_hist, stats) bind(c, name="artes_gpu_run_batch")
       import :: c_ptr, c_int, c_double, c_int64_t, artes_launch_t, artes_stats_t
       type(c_ptr), value               :: ctx
       type(artes_launch_t), intent(in) :: launches(*)
       integer(c_int), value            :: n
       real(c_double), intent(out)      :: det_sum(*), flux(*)
       integer(c_int64_t), intent(out)  :: err_hist(*)
       type(artes_stats_t), intent(out) :: stats
     end function artes_gpu_run_batch

  end interface

end module artes_gpu_mod
