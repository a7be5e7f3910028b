! radiative_transfer_gpu.f90 -- drop-in body for `subroutine radiative_transfer` of src/ARTES.f90 (:518-1006).
!
! How to wire it into the reference (see INTEGRATION.md):
!   1. add `use artes_gpu_mod` after `program artes`, and the two program-scope variables
!        type(c_ptr) :: gpu_ctx = c_null_ptr
!        integer     :: gpu_wl_loaded = 0
!   2. replace the body of `radiative_transfer` (everything between `:546` and `:955`, the OpenMP photon loop) by
!      the include below; the reduction / photometry tail `:957-1004` stays as it is, reading `detector_thread`
!      with threads = 1;
!   3. link with -lartes_gpu.
! The fragment uses only variables the reference already has at program scope (src/ARTES.f90:19-115).

    type(artes_launch_t) :: launch
    type(artes_stats_t)  :: stats
    integer(c_int64_t)   :: err_hist(ARTES_ERR_SLOTS)
    real(c_double)       :: flux(2)
    integer(c_int)       :: rc
    integer              :: code

    if (.not.c_associated(gpu_ctx)) then
       rc = artes_gpu_create(gpu_ctx, 1_c_int, c_null_ptr)
       if (rc.ne.0) stop "artes_gpu_create failed: no CUDA device (there is no CPU fallback)"
       rc = artes_gpu_set_grid(gpu_ctx, nr, ntheta, nphi, rfront, thetafront, thetaplane, phifront, oblate_x, oblate_y, oblate_z)
    end if

    if (gpu_wl_loaded.ne.wl_count) then
       ! the wl_count slices are contiguous because the cell indices come first (:64-69)
       if (photon_source.eq.2) then
          rc = artes_gpu_set_wavelength_dense(gpu_ctx, cell_scattering_opacity(:,:,:,wl_count), cell_absorption_opacity(:,:,:,wl_count), &
               cell_scatter_matrix(:,:,:,wl_count,:,:), cell_depth, c_loc(cell_weight), c_loc(emissivity_cumulative))
       else
          rc = artes_gpu_set_wavelength_dense(gpu_ctx, cell_scattering_opacity(:,:,:,wl_count), cell_absorption_opacity(:,:,:,wl_count), &
               cell_scatter_matrix(:,:,:,wl_count,:,:), cell_depth, c_null_ptr, c_null_ptr)
       end if
       gpu_wl_loaded = wl_count
    end if

    launch%struct_size = int(c_sizeof(launch), c_int32_t)
    launch%mode = ARTES_MODE_FAST
    launch%n_photons = int(packages, c_int64_t)
    launch%photon_id_base = 0_c_int64_t
    launch%seed = int(state(1,1), c_int64_t)              ! the clock-derived seed word of :433-449
    launch%photon_source = photon_source
    launch%photon_scattering = merge(1, 0, photon_scattering)
    launch%photon_emission = photon_emission
    launch%stellar_direction = merge(1, 0, stellar_direction)
    launch%limb_emission = merge(1, 0, phase_curve.and.det_phi*180._dp/pi.ge.170._dp)      ! :1041
    launch%flow_global = merge(1, 0, flow_global)
    launch%flow_theta = merge(1, 0, flow_theta)
    launch%nx = nx
    launch%ny = ny
    launch%wl_index = 0
    launch%fstop = fstop
    launch%photon_minimum = photon_minimum
    launch%photon_bias = photon_bias
    launch%surface_albedo = surface_albedo
    launch%theta_star = theta_star
    launch%phi_star = phi_star
    launch%det_theta = det_dir(4)
    launch%det_phi = det_dir(5)
    launch%x_max = x_max
    launch%y_max = y_max

    ! detector_thread(nx,ny,4,3,1) receives the sum over all GPU threads; imaging_broad keeps accumulating (:175-180)
    if (flow_global.and.flow_theta) then
       rc = artes_gpu_run(gpu_ctx, launch, detector_gpu, flux, c_loc(cell_flow_gpu), c_loc(cell_flow_global_gpu), err_hist, stats)
    else
       rc = artes_gpu_run(gpu_ctx, launch, detector_gpu, flux, c_null_ptr, c_null_ptr, err_hist, stats)
    end if
    if (rc.ne.0) stop "artes_gpu_run failed"
    detector_thread(:,:,:,:,1) = detector_thread(:,:,:,:,1) + reshape(detector_gpu, (/ nx, ny, 4, 3 /))
    if (photon_source.eq.2) then
       flux_emitted(1) = flux(1)
       flux_exit(1) = flux(2)
    end if

    ! the per-photon anomalies the reference appends to error.log one by one
    open (11, file=trim(error_log), position="append")
    do code = 1, ARTES_ERR_SLOTS
       if (err_hist(code).gt.0) write (11,'(a,i3.3,a,i0)') "error ", code-1, " x ", err_hist(code)
    end do
    close (11)
