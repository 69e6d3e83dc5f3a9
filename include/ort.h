/*
 * ort.h -- C-ABI of the B200-native per-ray trace loop of OpticalRayTrace.
 *
 * This is the seam the reference's two OpenMP ray loops are replaced by
 * (reference src/main.f90:90-109 "ring" loop, src/main.f90:127-162 "point" loop).
 * Everything is `extern "C"`, every struct is plain-old-data made of double /
 * int32_t / int64_t / uint64_t so it can be mirrored 1:1 by a Fortran
 * `type, bind(C)` (fortran/ort_interface.f90), by ctypes, or by cgo.
 *
 * Conventions
 *   - every entry point returns 0 on success and a negative ORT_E* code otherwise;
 *     nothing in the library calls exit()/abort().  ort_last_error() gives the text.
 *   - the caller owns all host buffers; the library owns device memory between
 *     ort_init*() and ort_finalize().
 *   - one host thread drives the library (the OpenMP region of the reference is gone).
 *   - there is NO CPU fallback: without a CUDA device every compute entry fails
 *     with ORT_ENODEVICE.
 *   - all lengths are metres, all reals fp64 (the reference is built with
 *     -freal-4-real-8, reference src/Makefile:2).
 */
#ifndef ORT_H
#define ORT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORT_VERSION 100

/* detector: image(-200:200,-200:200,layer), reference src/imageMod.f90:23 */
#define ORT_IMG_N 401
#define ORT_IMG_HALF 200
#define ORT_IMG_BINS (ORT_IMG_N * ORT_IMG_N)

/* error codes */
#define ORT_OK 0
#define ORT_EINVAL (-1)    /* bad argument */
#define ORT_ENODEVICE (-2) /* no CUDA device / library not initialised */
#define ORT_ECUDA (-3)     /* CUDA runtime error (text in ort_last_error) */
#define ORT_ENCCL (-4)     /* NCCL error or NCCL not loadable */
#define ORT_EIO (-5)       /* file could not be read / written */
#define ORT_EPARSE (-6)    /* malformed params file */
#define ORT_ETRACE (-7)    /* the kernels hit one of the reference's `error stop` invariants;
                              results are still returned, see status codes 18 and 24 */

/* phases */
#define ORT_PHASE_RING 1  /* reference src/main.f90:90-109 */
#define ORT_PHASE_POINT 2 /* reference src/main.f90:127-162 */

/* ort_job.flags */
#define ORT_FLAG_FIX_OUTER_ELLIPSE 1 /* opt-in: outer ellipse wall uses (Ra,Rb), not the
                                        reference's (Ra/2,Rb/2) (src/lens.f90:301) */
#define ORT_FLAG_NO_REDUCE 2         /* rank mode: leave per-rank images un-reduced */
#define ORT_FLAG_NO_COMPACTION 4     /* diagnostic: one-thread-per-ray kernel without the
                                        warp-level live-ray compaction */
#define ORT_FLAG_NO_FILTER 8         /* ring loop: do not use the single-precision culling filter
                                        (every ray past L2's aperture is traced in fp64; results are
                                        the same either way) */
#define ORT_FLAG_VERIFY_FILTER 16    /* ring loop, diagnostic: run the filter AND fp64 on every ray;
                                        status_hist[ORT_FILTER_SLOT_CALLED] = rays the filter called,
                                        status_hist[ORT_FILTER_SLOT_WRONG] = calls that disagree with
                                        fp64 (0 expected); image and the other counters as usual */

#define ORT_FLAG_ONE_LANE 32          /* diagnostic: a batched call with few rays per scene runs its scenes back to
                                        back on one stream instead of overlapping them on several */

/* ort_job.source_kind = settings.params source_type.  Which routine emits, per loop
 * (reference src/main.f90:95-101 and :132-142):
 *                 ring loop (phase 1)                      point loop (phase 2)
 *   POINT         ring()             sourceMod.f90:250     point()            :12
 *   CRS           point_on_bottle()  :50                   point()
 *   ISORS         iSORS(ring=.true.) :162                  point(offset = bottle centre z)
 *   SPOT          ring()                                   create_spot()      :122
 *   IMAGE         ring()                                   emit_image()       :303 (needs
 *                                                          ort_set_image_source first) */
#define ORT_SRC_POINT 0
#define ORT_SRC_CRS 1
#define ORT_SRC_ISORS 2
#define ORT_SRC_SPOT 3
#define ORT_SRC_IMAGE 4
#define ORT_SRCIMG_N 512 /* the image source is 512 x 512, src/sourceMod.f90:372 */

/* ort_job.stop_after (explicit-ray entry point only): where pos_out/dir_out are sampled */
#define ORT_STOP_NONE 0   /* full path, sampled at the image plane */
#define ORT_STOP_SOURCE 1 /* after ring()/point() */
#define ORT_STOP_BOTTLE 2 /* after glass_bottle%forward (point phase) */
#define ORT_STOP_L2 3     /* after plano_convex%forward */
#define ORT_STOP_L3 4     /* after achromatic_doublet%forward */

/* Final state of a ray ("which stage killed it").  0 = counted in the image. */
enum ort_status {
    ORT_ST_BINNED = 0,
    ORT_ST_BOTTLE_INNER_MISS = 1,    /* src/lens.f90:257-260 */
    ORT_ST_CONTENTS_ABSORBED = 2,    /* src/lens.f90:270-274 */
    ORT_ST_CONTENTS_BACKWARD = 3,    /* src/lens.f90:278-281 */
    ORT_ST_BOTTLE_INNER_REFLECT = 4, /* src/lens.f90:293-297 */
    ORT_ST_BOTTLE_OUTER_MISS = 5,    /* src/lens.f90:305-308 */
    ORT_ST_WALL_ABSORBED = 6,        /* src/lens.f90:321-325 */
    ORT_ST_WALL_BACKWARD = 7,        /* src/lens.f90:329-332 */
    ORT_ST_BOTTLE_OUTER_REFLECT = 8, /* src/lens.f90:344-348 */
    ORT_ST_L2_APERTURE = 9,          /* src/lens.f90:450-454 */
    ORT_ST_L2_SPHERE_MISS = 10,      /* src/lens.f90:462-467 */
    ORT_ST_L2_CURVED_REFLECT = 11,   /* src/lens.f90:475-479 */
    ORT_ST_L3_IRIS_BEFORE = 12,      /* src/lens.f90:551-565 */
    ORT_ST_L3_S1_MISS = 13,          /* src/lens.f90:568-572 */
    ORT_ST_L3_APERTURE = 14,         /* src/lens.f90:576-580 */
    ORT_ST_L3_S1_REFLECT = 15,       /* src/lens.f90:586-590 */
    ORT_ST_L3_S2_MISS = 16,          /* src/lens.f90:595-599 */
    ORT_ST_L3_S2_REFLECT = 17,       /* src/lens.f90:606-610 */
    ORT_ST_L3_S3_MISS = 18,          /* src/lens.f90:617  `error stop "Help3"` */
    ORT_ST_L3_S3_REFLECT = 19,       /* src/lens.f90:624-628 */
    ORT_ST_L3_IRIS_AFTER = 20,       /* src/lens.f90:632-644 */
    ORT_ST_NA_REJECT = 21,           /* src/imageMod.f90:42-44 */
    ORT_ST_FAR = 22,                 /* src/imageMod.f90:48, plus non-finite positions */
    ORT_ST_OFF_DETECTOR = 23,        /* src/imageMod.f90:52-54 */
    ORT_ST_TAUINT_MISS = 24,         /* src/surfaces.f90:33-39 `error stop "no intersection"` */
    ORT_ST_STOPPED = 25,             /* explicit-ray mode: reached ort_job.stop_after alive */
    ORT_ST_SOURCE_MISS = 26          /* iSORS: src/sourceMod.f90:216-218 `error stop "no intersection
                                        with bottle!"`; crs: spot point beside the bottle */
};
#define ORT_NSTATUS 32
#define ORT_SCATTER_EVENTS_SLOT 27   /* not a status: the number of scatter events (stokes calls, src/lens.f90:268,
                                       319) of the launch, for the roofline's flop count */
#define ORT_FILTER_SLOT_OVERFLOW 29 /* internal: non-zero makes ort_trace fail with ORT_ECUDA */
#define ORT_FILTER_SLOT_CALLED 30   /* only with ORT_FLAG_VERIFY_FILTER */
#define ORT_FILTER_SLOT_WRONG 31
/* statuses 1..20 and 24 are what the reference adds to rcount / pcount
 * (src/optics_system.f90:32,42 and src/main.f90:150-151); 26 is an abort there */
#define ORT_STATUS_IS_LOST(s) (((s) >= 1 && (s) <= 20) || (s) == 24 || (s) == 26)

/* reference src/lens.f90:8-20 (type lens + plano_convex) */
typedef struct {
    double thickness, diameter, radius, fb, f, n1, n2, curve_radius;
    double centre[3];
    double flat_normal[3];
} ort_plano;

/* reference src/lens.f90:8-12,27-33 (type lens + achromatic_doublet) */
typedef struct {
    double thickness, diameter, radius, fb, f, n1, n2, n3;
    double thickness1, thickness2, R1, R2, R3;
    double centre1[3], centre2[3], centre3[3];
} ort_doublet;

/* reference src/lens.f90:40-48 (type glass_bottle) */
typedef struct {
    double nbottle, ncontents, thickness, radiusa, radiusb;
    double mua_b, mus_b, mua_c, mus_c;
    double centre[3];
    int32_t ellipse, scatter_b, scatter_c, _pad;
} ort_bottle;

/* One optical configuration: what reference src/main.f90 holds in (bottle, L2, L3) plus the
 * scalars its prologue derives from them (src/main.f90:51-70,81).  A batched launch traces
 * many scenes at once (the runner.py sweeps become one call). */
typedef struct {
    ort_bottle bottle;
    ort_plano L2;
    ort_doublet L3;
    double cos_theta_max; /* src/main.f90:51-52 */
    double r1, r2;        /* squared annulus radii, src/main.f90:66-70 */
    double img_plane;     /* src/main.f90:81 */
    double point_offset;  /* z of the point source (0; bottle centre z for isors, main.f90:140) */
    double spot_size;     /* crs: sigma of the Gaussian spot on the bottle, already rescaled as in
                             src/setupMod.f90:135-136 */
    double isors_offset;  /* isors: ring / spot separation, src/setupMod.f90:132 */
    double ring_width;    /* isors: beam width on the axicon (ringWidth, src/setupMod.f90:57) */
} ort_scene;

/* One launch of a ray loop. */
typedef struct {
    int32_t phase;       /* ORT_PHASE_RING | ORT_PHASE_POINT */
    int32_t use_bottle;  /* src/setupMod.f90:63 (point phase only, src/main.f90:145) */
    int32_t iris_before; /* iris(1), src/setupMod.f90:103-108 */
    int32_t iris_after;  /* iris(2) */
    int32_t precision;   /* 64: fp64, the reference's arithmetic (1e-9 parity);
                            32: the fp32 variant (1e-5 parity, same draws, same decisions except
                            within 2^-24 of a threshold) */
    int32_t flags;       /* ORT_FLAG_* */
    int32_t stop_after;  /* ORT_STOP_* (ort_trace_rays only) */
    int32_t source_kind; /* ORT_SRC_*: which of the reference's sources feeds the loop */
    double iris_radius;      /* fraction of the lens radius, src/setupMod.f90:101 */
    double fibre_offset;     /* src/setupMod.f90:84 */
    double image_diameter;   /* src/setupMod.f90:83 */
    double uniform_override; /* < 0: draws come from the counter-based generator;
                                in [0,1): every draw returns this value (known-answer tests) */
    uint64_t seed;     /* key of the counter-based generator (reference seeds 123456789,
                          src/main.f90:79) */
    int64_t first_ray; /* index of the first ray; uniforms are a pure function of
                          (seed, phase, ray index, draw slot) */
    int64_t nrays;     /* rays per scene */
    int64_t total_rays; /* nphotons of the whole job (only create_spot depends on it,
                           src/sourceMod.f90:134-143); 0 means nrays */
} ort_job;

typedef struct {
    double trace_seconds;  /* CUDA-event time: image clear + trace kernels, max over devices */
    double reduce_seconds; /* CUDA-event time of the NCCL image reduce (0 on one GPU) */
    double wall_seconds;   /* host wall clock of the whole call incl. H2D/D2H */
    double d2h_seconds;    /* CUDA-event time of the image + histogram read-back on device 0 */
    int64_t kernel_launches;
    int64_t h2d_bytes, d2h_bytes;
} ort_timing;

/* parsed settings.params, reference src/setupMod.f90:57-133 */
typedef struct {
    double ring_width, wavelength, alpha_deg, n_axicon, image_diameter, fibre_offset;
    double iris_radius, isors_offset, spot_size;
    int64_t nphotons;
    int32_t use_bottle, use_tracker, make_images, iris_before, iris_after, _pad;
    char source_type[64];
    char iris_name[64];
    char bottle_file[256], l2_file[256], l3_file[256], image_file[256], folder[256];
} ort_settings;

/* ---- lifetime ------------------------------------------------------------------------ */
/* Single-process mode: drive devices 0..ngpus-1 from this process (ngpus<=0: all visible).
 * `./install.sh -n N` maps to this.  Returns the number of devices in use (>0) or ORT_E*.
 * With more than one device it also runs one small reduce, so that NCCL's channels exist before
 * the first job is timed. */
int ort_init(int ngpus);
/* One-process-per-GPU mode (torchrun / MPI style): this process owns `device`; when
 * nranks > 1, `nccl_id` is the 128-byte ncclUniqueId rank 0 got from ort_nccl_unique_id()
 * and distributed by any side channel.  Collective when nranks > 1: every rank calls it (it
 * creates the communicator and runs one small reduce on it). */
int ort_init_rank(int device, int rank, int nranks, const void* nccl_id);
int ort_nccl_unique_id(void* out128);
int ort_finalize(void);
/* cudaDeviceSynchronize on every device the library drives (ort_trace already returns only when its
 * own work is done; this is for callers that bracket a timed region) */
int ort_synchronize(void);
const char* ort_last_error(void);
int ort_device_count(void);
/* sizeof() of {ort_plano, ort_doublet, ort_bottle, ort_scene, ort_job, ort_timing, ort_settings}
 * followed by ORT_VERSION -- lets a foreign-language binding verify its struct mirrors. */
int ort_struct_sizes(int32_t out[8]);

/* ---- the hot path -------------------------------------------------------------------- */
/* Replaces one ray loop of reference src/main.f90 for `nscenes` optical configurations.
 *   image        [nscenes][ORT_IMG_BINS] uint64, host, OVERWRITTEN with the counts of this
 *                phase; element ((yp+200)*401 + (xp+200)) as in src/imageMod.f90:102-112.
 *                May be NULL (timing runs).
 *   lost         [nscenes] rays lost in bottle/lenses = rcount | pcount of src/main.f90:28
 *   status_hist  [nscenes][ORT_NSTATUS] histogram of enum ort_status (may be NULL)
 * In single-process multi-GPU mode the ray range is split over the devices and the
 * per-device images are summed with one ncclReduce before the call returns; in rank mode
 * each rank traces the [first_ray, first_ray+nrays) it is given and rank 0 receives the sum
 * (the other ranks' outputs are their private partial results). */
int ort_trace(const ort_job* job, const ort_scene* scenes, int nscenes, uint64_t* image,
              int64_t* lost, int64_t* status_hist, ort_timing* timing);

/* Parity entry point: trace an explicit list of n rays through ONE scene.
 *   pos_in/dir_in   structure-of-arrays [3][n] (x block, y block, z block); NULL => rays are
 *                   emitted by the device-side source of job->phase (ray index first_ray+i)
 *   pos_out/dir_out [3][n] state where the ray stopped (see ort_job.stop_after)
 *   status          [n] enum ort_status
 *   bin_xy          [2][n] (xp, yp) in -200..200, or INT32_MIN when not binned
 * Draws use ray index job->first_ray + i, so the same rays give the same decisions as
 * ort_trace. */
int ort_trace_rays(const ort_job* job, const ort_scene* scene, int64_t n, const double* pos_in,
                   const double* dir_in, double* pos_out, double* dir_out, int32_t* status,
                   int32_t* bin_xy);

/* The reference's ray tracker (src/stackMod.f90 + src/main.f90:103,107,144-160, read by
 * debug-plot.py): for rays [first_ray, first_ray+nrays) of job->phase, append to `path` the
 * positions each surviving ray visited (image plane first -- the stack is popped -- one
 * "3(F10.7,1x)" line per position) followed by three blank lines; rays lost in a lens leave
 * only the blank lines, rays lost in the bottle their source and bottle positions plus six.
 * nrays <= 10000 like the reference (src/setupMod.f90:75). */
int ort_write_tracks(const ort_job* job, const ort_scene* scene, const char* path);

/* Image source (source_type "image", reference src/sourceMod.f90:303-408).
 * ort_load_image_source = init_emit_image (:363-408): reads the 512x512 fp64 intensity file
 * bpm.py writes, and turns it into a per-pixel ray budget for `nphotons` rays (fractional parts
 * rounded up with probability = fraction; the draws come from the counter-based generator,
 * stream 3).  budget[(j-1)*512 + (i-1)] = imgin(i,j), the Fortran memory order, which is also
 * the order emit_image scans (:313-321).
 * ort_set_image_source hands a budget to the library (copied to every device); the point loop of
 * a job with source_kind = ORT_SRC_IMAGE then emits ray k from the pixel the reference's scan
 * would have reached after k rays.  Rays beyond the total budget end with ORT_ST_SOURCE_MISS
 * (the reference re-uses stale state for them). */
int ort_load_image_source(const char* path, int64_t nphotons, uint64_t seed, int32_t* budget);
int ort_set_image_source(const int32_t* budget);

/* The uniforms a ray sees: out[i] = uniform of slot `first_slot+i` (testing the generator). */
int ort_uniforms(uint64_t seed, int32_t phase, int64_t ray, int32_t first_slot, int32_t n,
                 double* out);

/* FP64 FMA peak of device 0 measured with a DFMA micro-kernel (roofline denominator).
 * Returns TFLOP/s in *tflops. */
int ort_measure_fp64_peak(double* tflops, double* sm_clock_mhz);

/* Accuracy of the library's fast fp64 reciprocal / division / sqrt / rsqrt on device 0: worst
 * error in units of the last place over n pseudo-random operands (order: rcp, div, sqrt, rsqrt). */
int ort_math_selftest(int64_t n, uint64_t max_ulp[4]);

/* The ring loop's single-precision filter rests its error bound on the
 * accuracy of five hardware approximations (rcp, rsqrt, sqrt, sin, cos .approx.ftz.f32).  This runs
 * all 2^32 fp32 bit patterns through them on device 0 against fp64 references:
 * worst[0..2] = largest relative error of rcp / rsqrt / sqrt (|x| in [2^-64, 2^64]),
 * worst[3..4] = largest absolute error of sin / cos (|x| <= 3.1416);  assumed[] (may be NULL) receives
 * what the filter's bounds assume -- at least twice the measured value, or the proof does not hold
 * on this device. */
int ort_mufu_selftest(double worst[5], double assumed[5]);

/* ---- host side of the drop-in surface (no device needed) -------------------------------- */
int ort_load_plano(const char* path, double wavelength, double offset, ort_plano* out);
int ort_load_doublet(const char* path, double wavelength, double offset, ort_doublet* out);
int ort_load_bottle(const char* path, double wavelength, ort_bottle* out);
int ort_read_settings(const char* path, ort_settings* out);
/* Build the scene for one phase as src/main.f90 does: loads the three files named in the
 * settings from `resdir`, the bottle at settings->wavelength and the lenses at
 * `lens_wavelength` (785 nm settings value for the ring phase, 843e-9 for the point phase,
 * src/main.f90:113-117), applies the offset guard (src/main.f90:54-58) and derives
 * cos_theta_max, r1, r2, img_plane.  `pre_guard_offset` (may be NULL) receives
 * bottle%centre%z before the guard (used by the output file name, src/main.f90:45-48). */
int ort_build_scene(const ort_settings* settings, const char* resdir, double lens_wavelength,
                    ort_scene* out, double* pre_guard_offset);
int ort_job_from_settings(const ort_settings* settings, int32_t phase, ort_job* out);
/* Output base name, src/main.f90:45-48 (without folder and without "_image-*.dat"). */
int ort_output_basename(const ort_settings* settings, const ort_scene* scene,
                        double pre_guard_offset, char* buf, size_t buflen);
/* Three raw fp64 401x401 files <base>_image-{ring,point,total}.dat, src/imageMod.f90:93-114 */
int ort_write_images(const char* base_with_folder, const uint64_t* ring, const uint64_t* point);
/* Append one line to <folder>/trans-stats.dat, src/main.f90:168-178 */
int ort_append_trans_stats(const char* folder, const ort_settings* settings,
                           const ort_scene* scene_after, int64_t rcount, int64_t pcount);

/* ------------------------------------------------------------------------------------------
 * Volume image: makeImage3D / writeImage3D, src/imageMod.f90:61-90,117-133.
 * The reference's generic `makeImage` picks this routine when `image` has rank 4; its main program
 * only ever passes rank 3, so this is the path a maintainer gets by changing that declaration.
 * From the image plane every surviving ray is sampled at ORT_VOL_DEPTH depths diameter/200 apart
 * (no NA test); each sample inside the 401 x 401 window increments its voxel, the first one
 * outside ends the ray.  volume[(depth * 401 + (yp + 200)) * 401 + (xp + 200)] -- the memory order
 * of image(-200:200, -200:200, 200, layer); uint32 counts (one layer: 128.6 MB).  Status BINNED =
 * at least one voxel hit, OFF_DETECTOR = none.  Runs on the library's first device.
 * ---------------------------------------------------------------------------------------- */
#define ORT_VOL_DEPTH 200
int ort_trace_volume(const ort_job* job, const ort_scene* scene, uint32_t* volume, int64_t* lost,
                     int64_t* status_hist);
/* <base>-vol-ring.dat and <base>-vol-point.dat: raw fp64, src/imageMod.f90:117-133.  Either
 * volume may be NULL (that file is then not written). */
int ort_write_volume(const char* base_with_folder, const uint32_t* vol_ring, const uint32_t* vol_point);

/* ------------------------------------------------------------------------------------------
 * Beam-propagation pre-processor: replaces the reference's bpm.py, the
 * script that writes `bessel-normal.dat`, the nxy x nxy fp64 intensity map of the Bessel beam that
 * the `image` source samples (ort_load_image_source reads 512 x 512).  Lengths in micrometres as in
 * the script.  The library must be initialised; the work runs on its first device (cuFFT).
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    double w0;          /* beam waist, bpm.py:84 (582*4) */
    double wavelength;  /* :85 (0.785) */
    double axicon_deg;  /* :87 (5) */
    double n_axicon;    /* :88 (1.45) */
    double xymax;       /* lateral extent of the grid, :93 (5000) */
    double ring_radius; /* centre of the ring-shaped start field, :120 (1612) */
    double ring_width;  /* its 1/e half width, :120 (300) */
    int32_t nxy;        /* grid points per side, :94 (512) */
    int32_t nz;         /* axial voxels: dz = 3 w0 (k / k_r) / nz, :95-100 (1000) */
    int32_t steps;      /* free-space steps before the lens; < 0: nz / 10 as in :126 */
    int32_t reserved;
} ort_bpm;
int ort_bpm_defaults(ort_bpm* p);
int ort_bpm_struct_size(void);
/* intensity[nxy * nxy] in the element order of the file bpm.py writes (|e^T|^2, bpm.py:203-204) */
int ort_bpm_bessel(const ort_bpm* p, double* intensity);
/* the same, written as raw fp64 to `path` (bpm.py:205 writes ./bessel-normal.dat) */
int ort_bpm_write_file(const ort_bpm* p, const char* path);

#ifdef __cplusplus
}
#endif
#endif /* ORT_H */
