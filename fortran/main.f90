program raytrace
! The reference's src/main.f90 with its two `!$OMP do` ray loops replaced by ort_trace().
! Everything else -- setup_sim, the loaders and dispersion laws of src/lens.f90, the scalar
! prologue, trans-stats.dat, the transmission prints, writeImage -- is the reference's own code,
! linked unmodified (constants, utils, vector_class, stackMod, random_mod, stokes, imageMod,
! surfaces, lens, sourceMod, setupMod).  All five source_types take this path (image: the budget
! init_emit_image built in setup_sim is handed to the library), and so does the ray tracker.
!
! UNCOMPILED: this repository's image has no Fortran compiler; tests/test_fortran_lint.py is the
! stand-in (declarations, argument counts against the interfaces); see INTEGRATION.md.

    use lensMod,      only : plano_convex, achromatic_doublet, glass_bottle
    use utils,        only : str
    use constants,    only : pi
    use setup
    use imageMod
    use ort_interface
    use iso_c_binding
    use iso_fortran_env, only: int64

    implicit none

    type(plano_convex)       :: L2
    type(achromatic_doublet) :: L3
    type(glass_bottle)       :: bottle

    integer, allocatable :: image(:, :, :), imgin(:,:)
    integer(c_int64_t), allocatable :: counts(:)
    integer(c_int64_t) :: lost(1), hist(ORT_NSTATUS)
    integer(int64) :: rcount, pcount
    real           :: angle, cosThetaMax, r1, r2
    real           :: besselDiameter, distance, img_plane_1
    logical        :: file_exists
    integer        :: nphotonsLocal, upoint, rc, ngpus
    integer(c_int32_t) :: sizes(8)
    character(len=:), allocatable :: filename
    character(len=32) :: env
    type(ort_scene)  :: scene
    type(ort_job)    :: job
    type(ort_timing) :: timing

    allocate(image(-200:200, -200:200, 2), imgin(512, 512), counts(ORT_IMG_BINS))
    image = 0

    call setup_sim(L2, L3, bottle, imgin, nphotonsLocal)

    filename = trim(adjustl(source_type))//"_bottle_"//str(use_bottle)//"_Ra_"// &
            str(bottle%radiusa,7)//"_Rb_"//str(bottle%radiusb,7)//"_offset_"//&
            str(bottle%centre%z,7)//"_"//str(iris)//"_"//str(iris_radius, 7)//"_L2f_"//str(L2%f,6)//"_L3f_"//str(L3%f,6)//&
            "_fo_"//str(fibre_offset, 7)//"_alp_"//str(alpha*180/pi,7)//"_bwidth_"//str(ringWidth,7)//"_sep_"//str(isors_offset,7)

    ! scalar prologue, unchanged from the reference (src/main.f90:51-70,81)
    angle = atan(L2%radius / L2%fb)
    cosThetaMax = cos(angle)
    if(l2%fb <= bottle%radiusa + bottle%centre%z)then
        print*,"Bottle offset too large! Adjusting so that there is a minimum of 2mm offset from lens."
        bottle%centre%z = L2%fb - bottle%radiusa - 2d-3
        print*,"Now bottle set at z position:",bottle%centre%z
    end if
    if(isors_source)then
        distance = bottle%radiusa + isors_offset
    else
        distance = (bottle%radiusa + bottle%centre%z)
    end if
    besselDiameter = distance*97.3d-3*tan(alpha* (n - 1)) /(l2%fb)
    r1 = besselDiameter - ringWidth
    r2 = (besselDiameter / 2.d0)**2
    r1 = r1**2
    img_plane_1 = 2.*(L2%fb + L3%fb) + L2%thickness + L3%thickness

    ! ---- device set-up: install.sh -n N exports ORT_NUM_GPUS -------------------------------
    rc = ort_struct_sizes(sizes)
    if(sizes(4) /= c_sizeof(scene) .or. sizes(5) /= c_sizeof(job))error stop "ort_interface.f90 out of date"
    call get_environment_variable("ORT_NUM_GPUS", env)
    ngpus = 0
    if(len_trim(env) > 0)read(env,*)ngpus
    if(ort_init(int(ngpus, c_int)) < 0)error stop "ort_init failed (no CUDA device?)"

    job%use_bottle = merge(1, 0, use_bottle)
    job%iris_before = merge(1, 0, iris(1));  job%iris_after = merge(1, 0, iris(2))
    job%precision = 64;  job%flags = 0;  job%stop_after = 0
    job%source_kind = ORT_SRC_POINT
    if(crs_source)job%source_kind = ORT_SRC_CRS
    if(isors_source)job%source_kind = ORT_SRC_ISORS
    if(spot_source)job%source_kind = ORT_SRC_SPOT
    if(image_source)then
        ! emit_image (src/sourceMod.f90:303-361) scans imgin; the library searches its prefix sums
        job%source_kind = ORT_SRC_IMAGE
        if(ort_set_image_source(imgin) /= 0)error stop "ort_set_image_source failed"
    end if
    job%iris_radius = iris_radius;  job%fibre_offset = fibre_offset;  job%image_diameter = image_diameter
    job%uniform_override = -1.d0
    job%seed = 123456789_c_int64_t          ! init_rng(123456789), src/main.f90:79
    job%first_ray = 0;  job%nrays = nphotons;  job%total_rays = nphotons

    ! ---- ring loop, was src/main.f90:90-109 ------------------------------------------------------
    call ort_pack_scene(bottle, L2, L3, cosThetaMax, r1, r2, img_plane_1, 0.d0, scene, spot_size, isors_offset, ringWidth)
    job%phase = ORT_PHASE_RING
    rc = ort_trace(job, [scene], 1_c_int, counts, lost, hist, timing)
    if(rc /= 0 .and. rc /= ORT_ETRACE)error stop "ort_trace (ring) failed"
    rcount = lost(1)
    image(:, :, 1) = reshape(int(counts), [401, 401])
    ! tracker, was src/main.f90:72-74,103,107 (+ src/optics_system.f90:28-50)
    if(use_tracker)then
        rc = ort_write_tracks(job, scene, folder//filename//"-ringtrace.dat"//c_null_char)
        if(rc /= 0 .and. rc /= ORT_ETRACE)error stop "ort_write_tracks (ring) failed"
    end if

    ! lenses at 843 nm for the point loop, src/main.f90:113-117
    wavelength = 843d-9
    L2 = plano_convex("../res/"//trim(L2file), wavelength)
    L3 = achromatic_doublet("../res/"//trim(L3file), wavelength, 2.*L2%fb+ L2%thickness)

    ! ---- point loop, was src/main.f90:127-162 --------------------------------------------------
    call ort_pack_scene(bottle, L2, L3, cosThetaMax, r1, r2, img_plane_1, merge(bottle%centre%z, 0.d0, isors_source), scene, &
                        spot_size, isors_offset, ringWidth)
    job%phase = ORT_PHASE_POINT
    rc = ort_trace(job, [scene], 1_c_int, counts, lost, hist, timing)
    if(rc /= 0 .and. rc /= ORT_ETRACE)error stop "ort_trace (point) failed"
    pcount = lost(1)
    image(:, :, 2) = reshape(int(counts), [401, 401])
    ! tracker, was src/main.f90:121-124,144-160
    if(use_tracker)then
        rc = ort_write_tracks(job, scene, folder//filename//"-pointtrace.dat"//c_null_char)
        if(rc /= 0 .and. rc /= ORT_ETRACE)error stop "ort_write_tracks (point) failed"
    end if
    rc = ort_finalize()

    ! ---- epilogue, unchanged (src/main.f90:168-185) -------------------------------------------
    inquire(file=folder//"trans-stats.dat", exist=file_exists)
    if(.not. file_exists)then
        open(newunit=upoint, file=folder//"trans-stats.dat")
        write(upoint, *)"r/%, p/%, l2%f, l3%f, bottle?, radiusA, radiusB, iris_pos, iris_radius, offset, source_type, seperation"
    else
        open(newunit=upoint, file=folder//"trans-stats.dat", position="append")
    end if
    write(upoint,*)100.*(1.-(rcount/(real(nphotons)))),",", 100.*(1.-(pcount/(real(nphotons)))),",", L2%f,",",&
              L3%f,",", use_bottle,",", bottle%radiusa,",", bottle%radiusb,",", iris,",", str(iris_radius,7), ",", bottle%centre%z, &
              ",", trim(adjustl(source_type))//",", isors_offset
    close(upoint)

    print"(A,1X,f8.2,A)","Ring  transmitted: ",100.*(1.-(rcount/(real(nphotons)))),"%"
    print"(A,1X,f8.2,A)","Point transmitted: ",100.*(1.-(pcount/(real(nphotons)))),"%"

    if(makeImages)then
        call writeImage(image, folder//filename//"_image")
    end if

end program raytrace
