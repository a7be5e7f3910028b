! radiative_transfer_gpu.f90 -- drop-in body for `subroutine radiative_transfer` of src/ARTES.f90 (:518-1006).
!
! How to wire it into the reference (see INTEGRATION.md):
!   1. add `use artes_gpu_mod` after `program artes`, and the program-scope variables
!        type(c_ptr)        :: gpu_ctx = c_null_ptr
!        integer            :: gpu_wl_loaded = 0
!        integer(c_int64_t) :: gpu_calls = 0                        ! calls of radiative_transfer so far
!        real(c_double), allocatable, target :: detector_gpu(:), cell_flow_gpu(:,:), cell_flow_global_gpu(:,:)
!      allocated next to detector_thread (:2543-2547): detector_gpu(nx*ny*12), cell_flow_gpu(4,cells), cell_flow_global_gpu(3,cells)
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
    type(c_ptr)          :: flow4_ptr, flow3_ptr

    if (.not.c_associated(gpu_ctx)) then
       rc = artes_gpu_create(gpu_ctx, 1_c_int, c_null_ptr)
       if (rc.ne.0) stop "artes_gpu_create failed: no CUDA device (there is no CPU fallback)"
       rc = artes_gpu_set_grid(gpu_ctx, nr, ntheta, nphi, rfront, thetafront, thetaplane, phifront, oblate_x, oblate_y, oblate_z)
    end if

    if (gpu_wl_loaded.ne.wl_count) then
       ! The WHOLE arrays are passed, not wl_count slices: cell_scatter_matrix(:,:,:,wl_count,:,:) is not contiguous for
       ! n_wavelength > 1 (the wavelength index sits between the cell indices and the matrix indices, :69), so a slice
       ! argument would make the compiler copy 16.6 GB at the scale configuration.  The library picks the wavelength
       ! with strides (element (cell, wl, e, a) at cell + cells*(wl + n_wl*(e + 16*a))) and de-duplicates on the GPU.
       if (photon_source.eq.2) then
          rc = artes_gpu_set_wavelength_dense_wl(gpu_ctx, n_wavelength, wl_count - 1, cell_scattering_opacity, cell_absorption_opacity, &
               cell_scatter_matrix, cell_depth, c_loc(cell_weight), c_loc(emissivity_cumulative))
       else
          rc = artes_gpu_set_wavelength_dense_wl(gpu_ctx, n_wavelength, wl_count - 1, cell_scattering_opacity, cell_absorption_opacity, &
               cell_scatter_matrix, cell_depth, c_null_ptr, c_null_ptr)
       end if
       if (rc.ne.0) stop "artes_gpu_set_wavelength_dense_wl failed"
       gpu_wl_loaded = wl_count
    end if

    launch%struct_size = int(c_sizeof(launch), c_int32_t)
    launch%mode = ARTES_MODE_FAST
    launch%n_photons = int(packages, c_int64_t)
    ! every call walks its own photon-id range: the reference carries its generator state from call to call (:4197-4230),
    ! so the launches of a spectrum / phase curve are statistically independent
    launch%photon_id_base = gpu_calls * int(packages, c_int64_t)
    gpu_calls = gpu_calls + 1
    launch%seed = int(iand(state(1,1), huge(state(1,1))), c_int64_t)   ! the clock-derived seed word of :433-449, sign bit cleared
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
    ! the two flow counters are independent keywords (:53-54): each buffer is passed iff its own flag is on
    flow4_ptr = c_null_ptr
    flow3_ptr = c_null_ptr
    if (flow_theta) flow4_ptr = c_loc(cell_flow_gpu)
    if (flow_global) flow3_ptr = c_loc(cell_flow_global_gpu)
    rc = artes_gpu_run(gpu_ctx, launch, detector_gpu, flux, flow4_ptr, flow3_ptr, err_hist, stats)
    if (rc.ne.0) stop "artes_gpu_run failed"
    detector_thread(:,:,:,:,1) = detector_thread(:,:,:,:,1) + reshape(detector_gpu, (/ nx, ny, 4, 3 /))
    ! the library returns the flows summed over its threads as (dir, cell): back into thread 1 of cell_flow(thread,4,i,j,k)
    ! (:82) and cell_flow_global(thread,3,i,j,k) (:81), which write_output reads (:3715-3766)
    if (flow_theta) cell_flow(1,:,:,:,:) = reshape(cell_flow_gpu, (/ 4, nr, ntheta, nphi /))
    if (flow_global) cell_flow_global(1,:,:,:,:) = reshape(cell_flow_global_gpu, (/ 3, nr, ntheta, nphi /))
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
