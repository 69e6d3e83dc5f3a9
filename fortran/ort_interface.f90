module ort_interface
! ISO_C_BINDING view of include/ort.h -- the seam the Fortran host calls instead of its two
! `!$OMP do` ray loops (reference src/main.f90:90-109 and :127-162).
!
! Every type below mirrors the C struct of the same name field for field (all members are
! c_double / c_int32_t / c_int64_t, so there is no padding surprise); `ort_struct_sizes` lets a
! program check that at start-up (see fortran/main.f90).
!
! NOTE: the build image has no Fortran compiler, so this file has never been through one
! ("uncompiled").  What stands in for the compiler is tests/test_fortran_lint.py: it tokenises this
! file and fails on any dummy argument without a declaration, on any `type, bind(C)` whose field
! list (names, order, kinds, extents) differs from the C struct of the same name in include/ort.h,
! on any parameter whose value differs from the #define of the same name, and on any bind(C)
! name or argument count that differs from the C prototype; tests/test_abi.py checks that
! libort.so exports every bind(C) name.  Build line once a compiler is available:
!   gfortran -O2 -cpp -freal-4-real-8 fortran/ort_interface.f90 fortran/main.f90 \
!            -Lopticalraytrace_b200 -lort -Wl,-rpath,$PWD/opticalraytrace_b200 -o bin/raytrace

    use iso_c_binding

    implicit none

    integer(c_int), parameter :: ORT_PHASE_RING = 1, ORT_PHASE_POINT = 2
    integer(c_int), parameter :: ORT_IMG_N = 401, ORT_IMG_BINS = 401*401, ORT_NSTATUS = 32
    integer(c_int), parameter :: ORT_ETRACE = -7
    integer(c_int), parameter :: ORT_FLAG_NO_FILTER = 8, ORT_FLAG_VERIFY_FILTER = 16   ! ort_job%flags, ring loop
    integer(c_int), parameter :: ORT_SRC_POINT = 0, ORT_SRC_CRS = 1, ORT_SRC_ISORS = 2, ORT_SRC_SPOT = 3, ORT_SRC_IMAGE = 4
    integer(c_int), parameter :: ORT_SRCIMG_N = 512

    type, bind(C) :: ort_plano                 ! reference src/lens.f90:8-20
        real(c_double) :: thickness, diameter, radius, fb, f, n1, n2, curve_radius
        real(c_double) :: centre(3), flat_normal(3)
    end type ort_plano

    type, bind(C) :: ort_doublet               ! reference src/lens.f90:27-33
        real(c_double) :: thickness, diameter, radius, fb, f, n1, n2, n3
        real(c_double) :: thickness1, thickness2, R1, R2, R3
        real(c_double) :: centre1(3), centre2(3), centre3(3)
    end type ort_doublet

    type, bind(C) :: ort_bottle                ! reference src/lens.f90:40-48
        real(c_double) :: nbottle, ncontents, thickness, radiusa, radiusb
        real(c_double) :: mua_b, mus_b, mua_c, mus_c
        real(c_double) :: centre(3)
        integer(c_int32_t) :: ellipse, scatter_b, scatter_c, pad_
    end type ort_bottle

    type, bind(C) :: ort_scene
        type(ort_bottle)  :: bottle
        type(ort_plano)   :: L2
        type(ort_doublet) :: L3
        real(c_double) :: cos_theta_max, r1, r2, img_plane, point_offset
        real(c_double) :: spot_size, isors_offset, ring_width
    end type ort_scene

    type, bind(C) :: ort_job
        integer(c_int32_t) :: phase, use_bottle, iris_before, iris_after, precision, flags, stop_after, source_kind
        real(c_double)     :: iris_radius, fibre_offset, image_diameter, uniform_override
        integer(c_int64_t) :: seed, first_ray, nrays, total_rays
    end type ort_job

    type, bind(C) :: ort_timing
        real(c_double)     :: trace_seconds, reduce_seconds, wall_seconds, d2h_seconds
        integer(c_int64_t) :: kernel_launches, h2d_bytes, d2h_bytes
    end type ort_timing

    interface
        integer(c_int) function ort_init(ngpus) bind(C, name="ort_init")
            import :: c_int
            integer(c_int), value :: ngpus
        end function ort_init

        integer(c_int) function ort_init_rank(device, rank, nranks, nccl_id) bind(C, name="ort_init_rank")
            import :: c_int, c_ptr
            integer(c_int), value :: device, rank, nranks
            type(c_ptr),    value :: nccl_id
        end function ort_init_rank

        integer(c_int) function ort_finalize() bind(C, name="ort_finalize")
            import :: c_int
        end function ort_finalize

        type(c_ptr) function ort_last_error() bind(C, name="ort_last_error")
            import :: c_ptr
        end function ort_last_error

        integer(c_int) function ort_struct_sizes(out) bind(C, name="ort_struct_sizes")
            import :: c_int, c_int32_t
            integer(c_int32_t), intent(out) :: out(8)
        end function ort_struct_sizes

        ! image(401*401, nscenes) is integer(c_int64_t): same bits as the library's uint64 counts
        integer(c_int) function ort_trace(job, scenes, nscenes, image, lost, status_hist, timing) &
                bind(C, name="ort_trace")
            import :: c_int, c_int64_t, ort_job, ort_scene, ort_timing
            type(ort_job),      intent(in)  :: job
            type(ort_scene),    intent(in)  :: scenes(*)
            integer(c_int),     value       :: nscenes
            integer(c_int64_t), intent(out) :: image(*)
            integer(c_int64_t), intent(out) :: lost(*)
            integer(c_int64_t), intent(out) :: status_hist(*)
            type(ort_timing),   intent(out) :: timing
        end function ort_trace

        integer(c_int) function ort_trace_rays(job, scene, n, pos_in, dir_in, pos_out, dir_out, status, bin_xy) &
                bind(C, name="ort_trace_rays")
            import :: c_int, c_int32_t, c_int64_t, c_double, ort_job, ort_scene
            type(ort_job),      intent(in)  :: job
            type(ort_scene),    intent(in)  :: scene
            integer(c_int64_t), value       :: n
            real(c_double),     intent(in)  :: pos_in(*), dir_in(*)
            real(c_double),     intent(out) :: pos_out(*), dir_out(*)
            integer(c_int32_t), intent(out) :: status(*), bin_xy(*)
        end function ort_trace_rays

        ! makeImage3D (src/imageMod.f90:61-90): volume(401*401*200) is integer(c_int32_t), the memory
        ! order of image(-200:200, -200:200, 200, layer)
        integer(c_int) function ort_trace_volume(job, scene, volume, lost, status_hist) &
                bind(C, name="ort_trace_volume")
            import :: c_int, c_int32_t, c_int64_t, ort_job, ort_scene
            type(ort_job),      intent(in)  :: job
            type(ort_scene),    intent(in)  :: scene
            integer(c_int32_t), intent(out) :: volume(*)
            integer(c_int64_t), intent(out) :: lost
            integer(c_int64_t), intent(out) :: status_hist(*)
        end function ort_trace_volume

        ! source_type image: `budget` is imgin(512, 512) exactly as the reference's init_emit_image
        ! (src/sourceMod.f90:363-408) leaves it -- default integer, Fortran memory order
        integer(c_int) function ort_set_image_source(budget) bind(C, name="ort_set_image_source")
            import :: c_int, c_int32_t
            integer(c_int32_t), intent(in) :: budget(*)
        end function ort_set_image_source

        ! the library's own init_emit_image (counter-based draws instead of ran2): path is a
        ! C string (trim(name)//c_null_char)
        integer(c_int) function ort_load_image_source(path, nphotons, seed, budget) &
                bind(C, name="ort_load_image_source")
            import :: c_int, c_int32_t, c_int64_t, c_char
            character(kind=c_char), intent(in)  :: path(*)
            integer(c_int64_t),     value       :: nphotons, seed
            integer(c_int32_t),     intent(out) :: budget(*)
        end function ort_load_image_source

        ! the ray tracker (src/stackMod.f90, src/main.f90:103-107,144-160): writes the trace file
        ! of job%phase for rays first_ray .. first_ray+nrays-1 (nrays <= 10000)
        integer(c_int) function ort_write_tracks(job, scene, path) bind(C, name="ort_write_tracks")
            import :: c_int, c_char, ort_job, ort_scene
            type(ort_job),          intent(in) :: job
            type(ort_scene),        intent(in) :: scene
            character(kind=c_char), intent(in) :: path(*)
        end function ort_write_tracks
    end interface

contains

    ! Copy the reference's derived types into the POD scene.  `use lensMod` types come from the
    ! UNMODIFIED reference src/lens.f90, whose loaders keep doing the file parsing + dispersion.
    subroutine ort_pack_scene(bottle, L2, L3, cosThetaMax, r1, r2, img_plane, point_offset, s, &
                              spot_size, isors_offset, ring_width)
        use lensMod, only : plano_convex, achromatic_doublet, glass_bottle
        type(glass_bottle),       intent(in)  :: bottle
        type(plano_convex),       intent(in)  :: L2
        type(achromatic_doublet), intent(in)  :: L3
        real,                     intent(in)  :: cosThetaMax, r1, r2, img_plane, point_offset
        type(ort_scene),          intent(out) :: s
        real(c_double), optional, intent(in)  :: spot_size, isors_offset, ring_width

        s%bottle%nbottle = bottle%nbottle;   s%bottle%ncontents = bottle%ncontents
        s%bottle%thickness = bottle%thickness
        s%bottle%radiusa = bottle%radiusa;   s%bottle%radiusb = bottle%radiusb
        s%bottle%mua_b = bottle%mua_b;       s%bottle%mus_b = bottle%mus_b
        s%bottle%mua_c = bottle%mua_c;       s%bottle%mus_c = bottle%mus_c
        s%bottle%centre = [bottle%centre%x, bottle%centre%y, bottle%centre%z]
        s%bottle%ellipse = merge(1, 0, bottle%ellipse)
        s%bottle%scatter_b = merge(1, 0, bottle%scatter_b)
        s%bottle%scatter_c = merge(1, 0, bottle%scatter_c)
        s%bottle%pad_ = 0

        s%L2%thickness = L2%thickness; s%L2%diameter = L2%diameter; s%L2%radius = L2%radius
        s%L2%fb = L2%fb; s%L2%f = L2%f; s%L2%n1 = L2%n1; s%L2%n2 = L2%n2
        s%L2%curve_radius = L2%curve_radius
        s%L2%centre = [L2%centre%x, L2%centre%y, L2%centre%z]
        s%L2%flat_normal = [L2%flatNormal%x, L2%flatNormal%y, L2%flatNormal%z]

        s%L3%thickness = L3%thickness; s%L3%diameter = L3%diameter; s%L3%radius = L3%radius
        s%L3%fb = L3%fb; s%L3%f = L3%f; s%L3%n1 = L3%n1; s%L3%n2 = L3%n2; s%L3%n3 = L3%n3
        s%L3%thickness1 = L3%thickness1; s%L3%thickness2 = L3%thickness2
        s%L3%R1 = L3%R1; s%L3%R2 = L3%R2; s%L3%R3 = L3%R3
        s%L3%centre1 = [L3%centre1%x, L3%centre1%y, L3%centre1%z]
        s%L3%centre2 = [L3%centre2%x, L3%centre2%y, L3%centre2%z]
        s%L3%centre3 = [L3%centre3%x, L3%centre3%y, L3%centre3%z]

        s%cos_theta_max = cosThetaMax
        s%r1 = r1;  s%r2 = r2
        s%img_plane = img_plane
        s%point_offset = point_offset
        s%spot_size = 0.d0;  s%isors_offset = 0.d0;  s%ring_width = 0.d0
        if(present(spot_size))s%spot_size = spot_size
        if(present(isors_offset))s%isors_offset = isors_offset
        if(present(ring_width))s%ring_width = ring_width
    end subroutine ort_pack_scene

end module ort_interface
