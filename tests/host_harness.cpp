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

#include <cmath>
#include <map>

#include "../opticalraytrace_b200/csrc/ort_flatten.h"
#include "../opticalraytrace_b200/csrc/ort_optics.cuh"
#include "../opticalraytrace_b200/csrc/ort_filter.cuh"

#ifdef ORTF_FUZZ
/* ORTF_FUZZ build: every MUFU stand-in is pushed to +- the error the bounds assume (half of
 * ORTF_E_*, i.e. the largest error measured on the device), sign from a per-thread LCG */
static thread_local uint32_t g_fuzz_state = 12345u;
extern "C" float ortf_fuzz_sign(void) {
    g_fuzz_state = g_fuzz_state * 1664525u + 1013904223u;
    return (g_fuzz_state & 0x80000000u) ? 1.0f : -1.0f;
}
#endif

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

static OrtRng harness_rng(const DevJob& J, int64_t i) {
    OrtRng g;
    uint64_t ray = (uint64_t)J.first_ray + (uint64_t)i;
    g.k0 = (uint32_t)J.seed; g.k1 = (uint32_t)(J.seed >> 32);
    g.rk = J.round_keys;
    g.r0 = (uint32_t)ray; g.r1 = (uint32_t)(ray >> 32);
    g.phase = (uint32_t)J.phase;
    g.override_u = -1.0;
    return g;
}

/* the ring loop's fp32 culling filter on rays [first, first+n): verdict[i] = -1 when stage A
 * already ends the ray (aim outside L2's aperture), else ort_ring_filter's answer (0 = hand to
 * fp64, s > 0 = proven status).  Returns 1 when the launcher would use the filter on this scene
 * (ring_shortcut and usable bound constants). */
extern "C" int hh_ring_filter(const ort_job* job, const ort_scene* scene, int64_t n, int32_t* verdict) {
    DevScene S;
    DevJob J;
    ort_flatten_scene(*scene, *job, S);
    ort_make_dev_job(*job, 1, job->first_ray, n, J);
    DevSceneT<float> F;
    ort_scene_to_float(S, F);
    DevFilter K;
    ort_make_filter(S, job->iris_before != 0, K);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        OrtRng g = harness_rng(J, i);
        uint32_t w[4];
        ort_block(g, 1u, w);
        const uint32_t hi = ort_aim_hi(g);
        const double u2 = ort_bits_to_uniform<double>(w[0], hi);
        verdict[i] = ort_ring_aims_outside_aperture(S, u2) ? -1 : ort_ring_filter(F, K, J, g, hi);
    }
    return (S.ring_shortcut && K.usable == 2) ? 1 : 0;
}

/* The filter's two-rays-per-lane instantiation (what the culling kernel runs; here with the packed
 * operations done per half): ray i in the low half, ray (i + shift) % n in the high half of the same lane.
 * lo[i] = verdict of ray i, hi[(i + shift) % n] = verdict of its partner; both must equal the one-ray
 * instantiation's verdict whatever the partner does (ends earlier, ends later, trips a guard, ...). */
extern "C" int hh_ring_filter_pairs(const ort_job* job, const ort_scene* scene, int64_t n, int64_t shift, int32_t* lo,
                                    int32_t* hi) {
    DevScene S;
    DevJob J;
    ort_flatten_scene(*scene, *job, S);
    ort_make_dev_job(*job, 1, job->first_ray, n, J);
    DevSceneT<float> F;
    ort_scene_to_float(S, F);
    DevFilter K;
    ort_make_filter(S, job->iris_before != 0, K);
    OrtfParamsT<OrtfV2> P;
    ortf_make_params<OrtfV2>(F, K, J.iris_before, P);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        const int64_t j = (i + shift) % n;
        OrtRng g0 = harness_rng(J, i), g1 = harness_rng(J, j);
        uint32_t a0[4], b0[4], a1[4], b1[4];
        ort_block(g0, 0u, a0);
        ort_block(g0, 1u, b0);
        ort_block(g1, 0u, a1);
        ort_block(g1, 1u, b1);
        const OrtfS2 st = ortf_filter<OrtfTwo>(P, OrtfW2{a0[1], a1[1]}, OrtfW2{a0[2], a1[2]}, OrtfW2{a0[3], a1[3]},
                                               OrtfW2{b0[2], b1[2]}, OrtfW2{b0[3], b1[3]}, OrtfW2{ort_aim_hi(g0), ort_aim_hi(g1)});
        lo[i] = st.a;
        hi[j] = st.b;
    }
    return (S.ring_shortcut && K.usable == 2) ? 1 : 0;
}

/* ---- double-precision twin of ort_ring_filter: the same quantities, from the exact draws and the
 * fp64 scene, recorded under the same tags ----------------------------------------------------- */
struct TwinRec { double v[3]; };
typedef std::map<int, TwinRec> TwinTrace;
static void trec(TwinTrace& T, int tag, double a, double b = 0, double c = 0) { T[tag] = TwinRec{{a, b, c}}; }
struct DRay { double px, py, pz, dx, dy, dz; };

static bool twin_sphere(const DRay& r, double cx, double cy, double cz, double R2, double* t, TwinTrace& T, int surf) {
    double lx = r.px - cx, ly = r.py - cy, lz = r.pz - cz;
    double h = r.dx * lx + r.dy * ly + r.dz * lz;
    double c = lx * lx + ly * ly + lz * lz - R2;
    double disc = h * h - c;
    trec(T, surf + ORTF_T_H, h);
    trec(T, surf + ORTF_T_C, c);
    trec(T, surf + ORTF_T_DISC, disc);
    bool hpos = h > 0.0;
    if (disc < 0.0 || (hpos && c > 0.0)) return false;
    double sq = std::sqrt(disc);
    double q = hpos ? -(h + sq) : (sq - h);
    *t = (!hpos && c < 0.0) ? q : c / q;
    trec(T, surf + ORTF_T_T, *t);
    return true;
}
static bool twin_interface(DRay& r, double nx, double ny, double nz, const DevIface& f, double u, TwinTrace& T, int surf) {
    double c = nx * r.dx + ny * r.dy + nz * r.dz;
    double costt = std::fabs(c);
    double s2 = 1.0 - costt * costt;
    double ct2 = 1.0 - f.eta2 * s2;
    trec(T, surf + ORTF_T_NI, c);
    trec(T, surf + ORTF_T_S2, s2);
    trec(T, surf + ORTF_T_CT2, ct2);
    bool tir = !(ct2 > 0.0);
    bool reflect = tir;
    double A = 0;
    if (!tir) {
        double cost2 = std::sqrt(ct2);
        trec(T, surf + ORTF_T_COST, cost2);
        double ec = f.eta * costt, e2 = f.eta * cost2;
        A = ec - cost2;
        double B = ec + cost2, C = e2 - costt, D = e2 + costt;
        double F = (u + u) * (B * B) * (D * D) - (A * A) * (D * D) - (C * C) * (B * B);
        trec(T, surf + ORTF_T_F, F);
        reflect = !(F > 0.0);
    }
    double a = reflect ? 1.0 : f.eta;
    double kk = reflect ? -2.0 * c : ((c < 0.0) ? A : -A);
    r.dx = a * r.dx + kk * nx;
    r.dy = a * r.dy + kk * ny;
    r.dz = a * r.dz + kk * nz;
    trec(T, surf + ORTF_T_DIR, r.dx, r.dy, r.dz);
    return reflect;
}
static int twin_filter(const DevScene& S, const DevJob& J, const OrtRng& g, uint32_t hi, TwinTrace& T) {
    uint32_t a[4], b[4];
    ort_block(g, 0u, a);
    ort_block(g, 1u, b);
    const double PI2 = 6.283185307179586476925286766559;
    const double u0 = ort_bits_to_uniform<double>(a[0], a[1]), u1 = ort_word_to_uniform<double>(a[2]);
    const double u2 = ort_bits_to_uniform<double>(b[0], hi), u3 = ort_word_to_uniform<double>(b[2]);
    double rr = std::sqrt(S.r1 + u0 * S.r2_m_r1);
    DRay r;
    double sx = rr * std::cos(PI2 * u1), sy = rr * std::sin(PI2 * u1);
    double q = S.ellipse ? sy * S.ra_over_rb : sy;
    double sz = S.bcz + std::sqrt(S.ra2 - q * q);
    double rl = std::sqrt(u2 * S.lens_r2);
    double ax = rl * std::cos(PI2 * u3), ay = rl * std::sin(PI2 * u3);
    double ex = ax - sx, ey = ay - sy, ez = S.l2_fb - sz;
    double inv = 1.0 / std::sqrt(ex * ex + ey * ey + ez * ez);
    r.dx = ex * inv; r.dy = ey * inv; r.dz = ez * inv;
    double t = (S.l2_flat_z - sz) / r.dz; /* the fp64 path's ort_l2_enter */
    r.px = sx + r.dx * t; r.py = sy + r.dy * t; r.pz = sz + r.dz * t;
    trec(T, 0 + ORTF_T_POS, r.px, r.py, r.pz);
    trec(T, 0 + ORTF_T_DIR, r.dx, r.dy, r.dz);
    (void)twin_interface(r, S.l2_fnx, S.l2_fny, S.l2_fnz, S.l2_in, ort_word_to_uniform<double>(a[3]), T, 100);
    if (!twin_sphere(r, S.l2_cx, S.l2_cy, S.l2_cz, S.l2_R2, &t, T, 200)) return ORT_ST_L2_SPHERE_MISS;
    r.px += r.dx * t; r.py += r.dy * t; r.pz += r.dz * t;
    double nx = (S.l2_cx - r.px) * S.l2_invR, ny = (S.l2_cy - r.py) * S.l2_invR, nz = (S.l2_cz - r.pz) * S.l2_invR;
    trec(T, 200 + ORTF_T_POS, r.px, r.py, r.pz);
    trec(T, 200 + ORTF_T_NORMAL, nx, ny, nz);
    if (twin_interface(r, nx, ny, nz, S.l2_out, ort_word_to_uniform<double>(b[3]), T, 300)) return ORT_ST_L2_CURVED_REFLECT;
    if (J.iris_before) {
        double ti = (S.l3_iris1_z - r.pz) / r.dz;
        double x = r.px + r.dx * ti, y = r.py + r.dy * ti;
        trec(T, 400 + ORTF_T_RHO2, x * x + y * y);
        if (x * x + y * y > S.l3_iris_r2) return ORT_ST_L3_IRIS_BEFORE;
    }
    if (!twin_sphere(r, S.l3_c1x, S.l3_c1y, S.l3_c1z, S.l3_R1_2, &t, T, 500)) return ORT_ST_L3_S1_MISS;
    r.px += r.dx * t; r.py += r.dy * t; r.pz += r.dz * t;
    trec(T, 500 + ORTF_T_POS, r.px, r.py, r.pz);
    trec(T, 600 + ORTF_T_RHO2, r.px * r.px + r.py * r.py);
    return (r.px * r.px + r.py * r.py > S.l3_radius2) ? ORT_ST_L3_APERTURE : 0;
}

/* Runs ort_ring_filter with its trace and the double twin on rays [first, first+n) that pass stage A
 * and compares every quantity the filter holds a bound for:
 *   max_ratio[tag % 100]  largest |fp32 - exact| / bound over the records made while no guard had tripped
 *   counts[0] records compared, [1] bound violations, [2] rays the filter called, [3] calls whose status
 *   differs from the twin's, [4] rays that passed stage A
 * Returns bit 0: the filter's premises hold (ring_shortcut -- L2's flat face is the aim plane -- and the ones
 * ort_make_filter checks; without them the launcher never runs the filter and the comparison means nothing),
 * bit 1: the bounds are also small enough for the launcher to use the filter. */
extern "C" int hh_filter_bounds(const ort_job* job, const ort_scene* scene, int64_t n, double* max_ratio /*[16]*/,
                                int64_t* counts /*[8]*/, int64_t* first_doubt /*[700] or NULL: by tag, which record
                                was the first made after something could not be proved*/) {
    DevScene S;
    DevJob J;
    ort_flatten_scene(*scene, *job, S);
    ort_make_dev_job(*job, 1, job->first_ray, n, J);
    DevSceneT<float> F;
    ort_scene_to_float(S, F);
    DevFilter K;
    ort_make_filter(S, job->iris_before != 0, K);
    for (int k = 0; k < 16; ++k) max_ratio[k] = 0.0;
    int64_t nrec = 0, nviol = 0, ncalled = 0, nwrong = 0, npass = 0;
#pragma omp parallel
    {
        double mr[16] = {0};
        int64_t lrec = 0, lviol = 0, lcalled = 0, lwrong = 0, lpass = 0;
#pragma omp for schedule(static)
        for (int64_t i = 0; i < n; ++i) {
            OrtRng g = harness_rng(J, i);
#ifdef ORTF_FUZZ
            g_fuzz_state = (uint32_t)(i * 2654435761u) ^ 0x9e3779b9u;
#endif
            uint32_t w[4];
            ort_block(g, 1u, w);
            const uint32_t hi = ort_aim_hi(g);
            if (ort_ring_aims_outside_aperture(S, ort_bits_to_uniform<double>(w[0], hi))) continue;
            ++lpass;
            OrtFilterTrace tr;
            tr.n = 0;
            int verdict = ort_ring_filter(F, K, J, g, hi, &tr);
            TwinTrace T;
            int exact = twin_filter(S, J, g, hi, T);
            if (verdict > 0) {
                ++lcalled;
                if (verdict != exact) ++lwrong;
            }
            if (first_doubt) {
                int tag = 699; /* no doubt at any record */
                for (int k = 0; k < tr.n; ++k)
                    if (!tr.rec[k].valid) { tag = tr.rec[k].tag; break; }
                if (verdict == 0) {
#pragma omp atomic
                    first_doubt[tag] += 1;
                }
            }
            for (int k = 0; k < tr.n; ++k) {
                const OrtFilterTrace::Rec& rc = tr.rec[k];
                if (!rc.valid || rc.tag % 100 >= 20) continue; /* 20..29: diagnostic checkpoints, not quantities */
                auto it = T.find(rc.tag);
                ++lrec;
                if (it == T.end()) { /* the twin went another way although nothing was in doubt */
                    ++lviol;
                    continue;
                }
                double dx = (double)rc.v[0] - it->second.v[0], dy = (double)rc.v[1] - it->second.v[1],
                       dz = (double)rc.v[2] - it->second.v[2];
                double err = std::sqrt(dx * dx + dy * dy + dz * dz);
                double ratio = err / (double)rc.bound;
                if (!(ratio <= 1.0)) ++lviol;
                int kind = rc.tag % 100;
                if (ratio > mr[kind] || !(ratio == ratio)) mr[kind] = ratio;
            }
        }
#pragma omp critical
        {
            for (int k = 0; k < 16; ++k)
                if (mr[k] > max_ratio[k] || !(mr[k] == mr[k])) max_ratio[k] = mr[k];
            nrec += lrec; nviol += lviol; ncalled += lcalled; nwrong += lwrong; npass += lpass;
        }
    }
    counts[0] = nrec; counts[1] = nviol; counts[2] = ncalled; counts[3] = nwrong; counts[4] = npass;
    return (S.ring_shortcut && K.usable >= 1 ? 1 : 0) | (K.usable == 2 ? 2 : 0);
}

/* the integer form of stage A's aperture test: returns the cut (0 when none exists) */
extern "C" unsigned long long hh_ring_aim_cut(const ort_job* job, const ort_scene* scene, int* have) {
    DevScene S;
    ort_flatten_scene(*scene, *job, S);
    unsigned long long cut = 0;
    *have = ort_ring_aim_cut(S, &cut) ? 1 : 0;
    return cut;
}

/* debugging aid: the filter's trace of one ray (tag, valid, value, bound) */
extern "C" int hh_filter_trace(const ort_job* job, const ort_scene* scene, int64_t i, int32_t* tags, float* vals /*[n][4]*/) {
    DevScene S;
    DevJob J;
    ort_flatten_scene(*scene, *job, S);
    ort_make_dev_job(*job, 1, job->first_ray, 1, J);
    DevSceneT<float> F;
    ort_scene_to_float(S, F);
    DevFilter K;
    ort_make_filter(S, job->iris_before != 0, K);
    OrtRng g = harness_rng(J, i);
    OrtFilterTrace tr;
    tr.n = 0;
    int verdict = ort_ring_filter(F, K, J, g, ort_aim_hi(g), &tr);
    for (int k = 0; k < tr.n; ++k) {
        tags[2 * k] = tr.rec[k].tag; tags[2 * k + 1] = tr.rec[k].valid;
        for (int c = 0; c < 3; ++c) vals[4 * k + c] = tr.rec[k].v[c];
        vals[4 * k + 3] = tr.rec[k].bound;
    }
    tags[2 * tr.n] = -1;
    return verdict;
}

/* whether the launcher uses the ring filter on this scene (usable bound constants) */
extern "C" int hh_ring_filter_in_range(const ort_job* job, const ort_scene* scene) {
    DevScene S;
    ort_flatten_scene(*scene, *job, S);
    DevFilter K;
    ort_make_filter(S, job->iris_before != 0, K);
    return K.usable == 2;
}
