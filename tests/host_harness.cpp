/*
 * host_harness.cpp -- TEST INFRASTRUCTURE ONLY (never part of libort.so, never a fallback).
 *
 * Compiles the product's per-ray optics header (opticalraytrace_b200/csrc/ort_optics.cuh) and
 * scene flattening for the HOST, so the reformulated arithmetic the kernels run (one-division
 * quadratics, fused Fresnel+Snell, hoisted normals, squared-radius tests ...) can be compared
 * against the oracle on a machine without a GPU.  The GPU tests (-m gpu) then check the same
 * header compiled for sm_100a through the real C-ABI.
 */
#include <cstdint>
#include <cstring>
#include <vector>

#include "../opticalraytrace_b200/csrc/ort_flatten.h"
#include "../opticalraytrace_b200/csrc/ort_optics.cuh"

template <typename R>
static void run_rays(const DevSceneT<R>& S, const DevJob& J, int64_t n, const double* pin, const double* din,
                     double* pout, double* dout, int32_t* status, int32_t* bin) {
    const bool have = pin != nullptr;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        OrtRng g;
        uint64_t ray = (uint64_t)J.first_ray + (uint64_t)i;
        g.k0 = (uint32_t)J.seed; g.k1 = (uint32_t)(J.seed >> 32);
        g.rk = J.round_keys;
        g.r0 = (uint32_t)ray; g.r1 = (uint32_t)(ray >> 32);
        g.phase = (uint32_t)J.phase;
        g.override_u = J.uniform_override;
        OrtRayT<R> r = {0, 0, 0, 0, 0, 1};
        if (have) {
            r.px = (R)pin[i]; r.py = (R)pin[n + i]; r.pz = (R)pin[2 * n + i];
            r.dx = (R)din[i]; r.dy = (R)din[n + i]; r.dz = (R)din[2 * n + i];
        }
        int x = 0, y = 0;
        int st = ort_full_path(S, J, g, have, r, &x, &y);
        pout[i] = r.px; pout[n + i] = r.py; pout[2 * n + i] = r.pz;
        dout[i] = r.dx; dout[n + i] = r.dy; dout[2 * n + i] = r.dz;
        status[i] = st;
        bin[i] = st == ORT_ST_BINNED ? x : INT32_MIN;
        bin[n + i] = st == ORT_ST_BINNED ? y : INT32_MIN;
    }
}

static std::vector<long long> g_cdf;
extern "C" int hh_set_image_source(const int32_t* budget) {
    g_cdf.clear();
    if (!budget) return 0;
    long long acc = 0;
    for (size_t k = 0; k < (size_t)ORT_SRCIMG_N * ORT_SRCIMG_N; ++k) {
        acc += budget[k] > 0 ? budget[k] : 0;
        g_cdf.push_back(acc);
    }
    return 0;
}

extern "C" int hh_trace_rays(const ort_job* job, const ort_scene* scene, int64_t n, const double* pin,
                             const double* din, double* pout, double* dout, int32_t* status, int32_t* bin) {
    DevScene S;
    DevJob J;
    ort_flatten_scene(*scene, *job, S);
    ort_make_dev_job(*job, 1, job->first_ray, n, J);
    J.image_cdf = g_cdf.empty() ? nullptr : g_cdf.data();
    if (job->precision == 32) {
        DevSceneT<float> Sf;
        ort_scene_to_float(S, Sf);
        run_rays<float>(Sf, J, n, pin, din, pout, dout, status, bin);
    } else {
        run_rays<double>(S, J, n, pin, din, pout, dout, status, bin);
    }
    return 0;
}

/* the ring loop's fp32 culling filter on rays [first, first+n): verdict[i] = -1 when stage A
 * already ends the ray (aim outside L2's aperture), else ort_ring_filter's answer (0 = hand to
 * fp64, s > 0 = certain status).  Returns the scene's ring_shortcut flag (the filter is only used
 * when it is set). */
extern "C" int hh_ring_filter(const ort_job* job, const ort_scene* scene, int64_t n, int32_t* verdict) {
    DevScene S;
    DevJob J;
    ort_flatten_scene(*scene, *job, S);
    ort_make_dev_job(*job, 1, job->first_ray, n, J);
    DevSceneT<float> F;
    ort_scene_to_float(S, F);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        OrtRng g;
        uint64_t ray = (uint64_t)J.first_ray + (uint64_t)i;
        g.k0 = (uint32_t)J.seed; g.k1 = (uint32_t)(J.seed >> 32);
        g.rk = J.round_keys;
        g.r0 = (uint32_t)ray; g.r1 = (uint32_t)(ray >> 32);
        g.phase = (uint32_t)J.phase;
        g.override_u = -1.0;
        uint32_t w[4];
        ort_block(g, 1u, w);
        const double u2 = ort_bits_to_uniform<double>(w[0], w[1]);
        verdict[i] = ort_ring_aims_outside_aperture(S, u2) ? -1 : ort_ring_filter(F, J, g, w[1], w[2], w[3]);
    }
    return S.ring_shortcut;
}

/* the integer form of stage A's aperture test: returns the cut (0 when none exists) */
extern "C" unsigned long long hh_ring_aim_cut(const ort_job* job, const ort_scene* scene, int* have) {
    DevScene S;
    ort_flatten_scene(*scene, *job, S);
    unsigned long long cut = 0;
    *have = ort_ring_aim_cut(S, &cut) ? 1 : 0;
    return cut;
}

/* the launcher's range guard for the ring filter */
extern "C" int hh_ring_filter_in_range(const ort_job* job, const ort_scene* scene) {
    DevScene S;
    ort_flatten_scene(*scene, *job, S);
    return ort_ring_filter_in_range(S, job->iris_before != 0) ? 1 : 0;
}
