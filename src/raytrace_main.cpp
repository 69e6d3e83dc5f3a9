/*
 * raytrace -- host program around the B200 trace loop.
 *
 * Does what the reference's `program raytrace` (src/main.f90) does, with its two `!$OMP do`
 * loops (src/main.f90:90-109 and :127-162) replaced by two ort_trace() calls:
 *   settings.params from ../res/<argv[1]>  ->  optics at the settings wavelength (ring phase)
 *   -> ring loop -> lenses re-built at 843 nm (src/main.f90:113-117) -> point loop
 *   -> ../data/<folder>/trans-stats.dat, the two "transmitted" lines, three raw images.
 * Run from bin/ exactly like the reference binary: `cd bin && ./raytrace settings.params`.
 * The Fortran equivalent (fortran/main.f90) is the same program over the same C-ABI; this C++
 * twin exists because the build image has no Fortran compiler.
 *
 * ORT_NUM_GPUS (exported by `install.sh -n N`) selects how many devices share the ray range.
 */
#include <sys/stat.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../include/ort.h"

static int fail(const char* what) {
    std::fprintf(stderr, "raytrace: %s: %s\n", what, ort_last_error());
    return 1;
}

int main(int argc, char** argv) {
    const char* arg = argc > 1 ? argv[1] : "settings.params";
    const char* resdir = std::getenv("ORT_RES_DIR") ? std::getenv("ORT_RES_DIR") : "../res";
    const char* datadir = std::getenv("ORT_DATA_DIR") ? std::getenv("ORT_DATA_DIR") : "../data";
    std::printf(" Using %s settings.\n", arg);

    ort_settings st;
    if (ort_read_settings((std::string(resdir) + "/" + arg).c_str(), &st)) return fail("settings");
    if (st.use_tracker) { /* src/setupMod.f90:75-82 */
        if (st.nphotons > 10000) {
            std::fprintf(stderr, "ERROR STOP Too many photons for tracker use!\n");
            return 1;
        }
        if (st.make_images) {
            std::printf(" ***************\n Cannot track packets and make images!\n"
                        " Deselecting makeImages\n ***************\n");
            st.make_images = 0;
        }
    }
    if (std::strcmp(st.source_type, "image") == 0) { /* init_emit_image, src/setupMod.f90:120-121 */
        std::vector<int32_t> budget((size_t)ORT_SRCIMG_N * ORT_SRCIMG_N);
        if (ort_load_image_source((std::string(resdir) + "/" + st.image_file).c_str(), st.nphotons, 123456789ull,
                                  budget.data()))
            return fail("image source");
        ort_set_image_source(budget.data());
    }
    int want = std::getenv("ORT_NUM_GPUS") ? std::atoi(std::getenv("ORT_NUM_GPUS")) : 0;
    int ngpu = ort_init(want);
    if (ngpu < 0) return fail("ort_init");
    if (want > ngpu) /* install.sh -n N with N above the visible device count (its default is 32 = "all") */
        std::printf(" ORT_NUM_GPUS=%d but %d CUDA device%s visible: using %d\n", want, ngpu, ngpu == 1 ? "" : "s", ngpu);

    ort_scene ring_scene, point_scene;
    double pre_guard = 0.0;
    if (ort_build_scene(&st, resdir, st.wavelength, &ring_scene, &pre_guard)) return fail("scene");
    if (ort_build_scene(&st, resdir, 843e-9, &point_scene, nullptr)) return fail("scene");
    if (ring_scene.bottle.centre[2] != pre_guard) { /* src/main.f90:54-58 */
        std::printf(" Bottle offset too large! Adjusting so that there is a minimum of 2mm offset from lens.\n");
        std::printf(" Now bottle set at z position:   %.16E\n", ring_scene.bottle.centre[2]);
    }
    char name[1024];
    if (ort_output_basename(&st, &ring_scene, pre_guard, name, sizeof name)) return fail("name");

    std::vector<uint64_t> ring(ORT_IMG_BINS), point(ORT_IMG_BINS);
    int64_t rcount = 0, pcount = 0;
    ort_job job;
    ort_timing tr, tp;
    ort_job_from_settings(&st, ORT_PHASE_RING, &job);
    int rc = ort_trace(&job, &ring_scene, 1, ring.data(), &rcount, nullptr, &tr);
    if (rc && rc != ORT_ETRACE) return fail("ring loop");
    ort_job_from_settings(&st, ORT_PHASE_POINT, &job);
    rc = ort_trace(&job, &point_scene, 1, point.data(), &pcount, nullptr, &tp);
    if (rc && rc != ORT_ETRACE) return fail("point loop");
    if (rc == ORT_ETRACE) std::fprintf(stderr, "raytrace: warning: %s\n", ort_last_error());

    std::string folder = std::string(datadir) + "/" + st.folder + "/";
    ::mkdir(datadir, 0777);
    ::mkdir(folder.c_str(), 0777);
    if (st.use_tracker) { /* src/main.f90:72-74,121-124: <name>-ringtrace.dat / -pointtrace.dat */
        ort_job_from_settings(&st, ORT_PHASE_RING, &job);
        if (ort_write_tracks(&job, &ring_scene, (folder + name + "-ringtrace.dat").c_str())) return fail("ring tracks");
        ort_job_from_settings(&st, ORT_PHASE_POINT, &job);
        if (ort_write_tracks(&job, &point_scene, (folder + name + "-pointtrace.dat").c_str())) return fail("point tracks");
    }
    if (ort_append_trans_stats(folder.c_str(), &st, &point_scene, rcount, pcount)) return fail("trans-stats");
    double n = (double)st.nphotons;
    std::printf("Ring  transmitted:  %8.2f%%\n", 100. * (1. - (rcount / n)));
    std::printf("Point transmitted:  %8.2f%%\n", 100. * (1. - (pcount / n)));
    if (st.make_images) {
        if (ort_write_images((folder + name + "_image").c_str(), ring.data(), point.data())) return fail("images");
    }
    std::fprintf(stderr,
                 "[ort] %d GPU(s): ring %.3e rays/s (%.3f ms), point %.3e rays/s (%.3f ms) device-timed\n", ngpu,
                 n / (tr.trace_seconds + tr.reduce_seconds), (tr.trace_seconds + tr.reduce_seconds) * 1e3,
                 n / (tp.trace_seconds + tp.reduce_seconds), (tp.trace_seconds + tp.reduce_seconds) * 1e3);
    ort_finalize();
    return 0;
}
