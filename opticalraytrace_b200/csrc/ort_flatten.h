/* ort_flatten.h -- host-only: public scene + job -> the flattened constant-memory form the
 * kernels read (DevScene / DevJob, ort_dev_types.h).  Every launch-invariant scalar is hoisted
 * here, once per launch. */
#ifndef ORT_FLATTEN_H
#define ORT_FLATTEN_H

#include <cmath>
#include <cstring>

#include "ort_dev_types.h"

/* ------------------------------------------------------------------------------------------
 * scene flattening (host, once per launch)
 * ---------------------------------------------------------------------------------------- */
inline DevIface ort_mk_iface(double na, double nb) {
    DevIface f;
    f.na = na;
    f.nb = nb;
    f.eta = na / nb;
    f.eta2 = f.eta * f.eta;
    return f;
}
inline double ort_sq(double x) { return x * x; }

inline void ort_flatten_scene(const ort_scene& s, const ort_job& j, DevScene& d) {
    memset(&d, 0, sizeof d);
    const ort_bottle& b = s.bottle;
    d.bcx = b.centre[0]; d.bcy = b.centre[1]; d.bcz = b.centre[2];
    d.b_in_r = b.radiusa - b.thickness;
    d.b_in_r2 = ort_sq(d.b_in_r);
    d.b_out_r = b.radiusa;
    d.b_out_r2 = ort_sq(b.radiusa);
    d.b_in_invr = 1.0 / d.b_in_r;
    d.b_out_invr = 1.0 / d.b_out_r;
    d.b_in_ia2 = 1.0 / ort_sq(b.radiusa - b.thickness);
    d.b_in_ib2 = 1.0 / ort_sq(b.radiusb - b.thickness);
    if (j.flags & ORT_FLAG_FIX_OUTER_ELLIPSE) {
        d.b_out_ia2 = 1.0 / ort_sq(b.radiusa);
        d.b_out_ib2 = 1.0 / ort_sq(b.radiusb);
    } else { /* reference src/lens.f90:301 halves the outer radii */
        d.b_out_ia2 = 1.0 / ort_sq(b.radiusa / 2.0);
        d.b_out_ib2 = 1.0 / ort_sq(b.radiusb / 2.0);
    }
    d.b_in = ort_mk_iface(b.ncontents, b.nbottle);
    d.b_out = ort_mk_iface(b.nbottle, 1.0);
    d.mutot_c = b.mua_c + b.mus_c;
    d.inv_mutot_c = d.mutot_c != 0.0 ? 1.0 / d.mutot_c : 0.0;
    d.albedo_c = d.mutot_c != 0.0 ? b.mus_c / (b.mus_c + b.mua_c) : 0.0;
    d.mutot_b = b.mua_b + b.mus_b;
    d.inv_mutot_b = d.mutot_b != 0.0 ? 1.0 / d.mutot_b : 0.0;
    d.albedo_b = d.mutot_b != 0.0 ? b.mus_b / (b.mus_b + b.mua_b) : 0.0;
    d.ellipse = b.ellipse;
    d.scatter_b = b.scatter_b;
    d.scatter_c = b.scatter_c;

    d.cos_theta_max = s.cos_theta_max;
    d.one_m_ctm = 1.0 - s.cos_theta_max;
    d.point_offset = s.point_offset;
    d.r1 = s.r1;
    d.r2_m_r1 = s.r2 - s.r1;
    d.ra2 = ort_sq(b.radiusa);
    d.ra_over_rb = b.radiusa / b.radiusb;
    d.lens_r2 = ort_sq(s.L2.radius + 10e-3);
    d.l2_fb = s.L2.fb;

    { /* crs: src/sourceMod.f90:82; isors: src/sourceMod.f90:180-188,206 */
        const double axicon_n = 1.4, radius = 12.7e-3, height = 1.1e-3;
        d.spot_size = s.spot_size;
        d.crs_r2 = ort_sq(b.radiusa + b.thickness);
        d.isors_beam = s.ring_width;
        double alpha = std::atan(height / radius);
        d.isors_base = (s.isors_offset + s.ring_width) / std::tan(alpha * (axicon_n - 1.));
        d.isors_k = ort_sq(radius / height);
        d.isors_h = height;
        d.isors_z = b.radiusa + b.centre[2] + 2.220446049250313e-16;
        d.isors_lens_r2 = ort_sq(s.L2.radius);
        d.isors_axicon = ort_mk_iface(axicon_n, 1.0);
    }
    const ort_plano& p = s.L2;
    d.l2_cx = p.centre[0]; d.l2_cy = p.centre[1]; d.l2_cz = p.centre[2];
    d.l2_flat_z = p.centre[2] + p.curve_radius - p.thickness;
    d.l2_radius2 = ort_sq(p.radius);
    d.l2_R2 = ort_sq(p.curve_radius);
    d.l2_invR = 1.0 / p.curve_radius;
    d.l2_fnx = p.flat_normal[0]; d.l2_fny = p.flat_normal[1]; d.l2_fnz = p.flat_normal[2];
    d.l2_in = ort_mk_iface(p.n1, p.n2);
    d.l2_out = ort_mk_iface(p.n2, p.n1);

    const ort_doublet& q = s.L3;
    d.l3_c1x = q.centre1[0]; d.l3_c1y = q.centre1[1]; d.l3_c1z = q.centre1[2];
    d.l3_c2x = q.centre2[0]; d.l3_c2y = q.centre2[1]; d.l3_c2z = q.centre2[2];
    d.l3_c3x = q.centre3[0]; d.l3_c3y = q.centre3[1]; d.l3_c3z = q.centre3[2];
    d.l3_R1_2 = ort_sq(q.R1); d.l3_R2_2 = ort_sq(q.R2); d.l3_R3_2 = ort_sq(q.R3);
    d.l3_invR1 = 1.0 / q.R1; d.l3_invR2 = 1.0 / q.R2; d.l3_invR3 = 1.0 / q.R3;
    d.l3_radius2 = ort_sq(q.radius * 1.0);
    d.l3_iris_r2 = ort_sq(q.radius * j.iris_radius);
    d.l3_iris1_z = q.centre1[2] - q.R1;
    d.l3_iris2_z = q.centre3[2] + q.R3;
    d.l3_s1 = ort_mk_iface(q.n1, q.n2);
    d.l3_s2 = ort_mk_iface(q.n2, q.n3);
    d.l3_s3 = ort_mk_iface(q.n3, q.n1);

    d.img_z = s.img_plane + j.fibre_offset;
    d.binwid = j.image_diameter / 401.0;
    d.inv_binwid = 401.0 / j.image_diameter;
    d.cos_na2 = 1.0 - 0.22 * 0.22; /* cos^2(asin(0.22)), reference src/imageMod.f90:40 */
    /* flat_z = (fb + th - R) + R - th: equal to fb up to rounding for loader-built lenses */
    d.ring_shortcut = (p.centre[0] == 0.0 && p.centre[1] == 0.0 &&
                       std::fabs(d.l2_flat_z - d.l2_fb) <= 4e-16 * std::fabs(d.l2_fb)) ? 1 : 0;
}

/* Stage A of the ring loop on the raw draw: the smallest 64-bit word w for which
 * u2 = (w >> 11) 2^-53 fails L2's aperture test exactly as the kernels evaluate it,
 * u2 * lens_r2 > l2_radius2 (ort_ring_aims_outside_aperture).  The expression is monotone in w,
 * so a bisection finds it.  Returns false when no draw fails (then the caller does not use the
 * integer test). */
inline bool ort_ring_aim_cut(const DevScene& d, unsigned long long* cut) {
    auto outside = [&](unsigned long long bits) {
        double u2 = (double)bits * (1.0 / 9007199254740992.0);
        return u2 * d.lens_r2 > d.l2_radius2;
    };
    const unsigned long long top = (1ull << 53) - 1;
    if (!outside(top)) return false;
    unsigned long long lo = 0, hi = top; /* outside(hi) holds */
    while (lo < hi) {
        unsigned long long mid = lo + ((hi - lo) >> 1);
        if (outside(mid)) hi = mid;
        else lo = mid + 1;
    }
    *cut = lo << 11;
    return true;
}

/* The ring loop's fp32 culling filter (ort_ring_filter) was validated on geometries whose largest
 * coordinate is ~12x their smallest aperture (profiles/r01_filter_margin.txt).  Its margins are
 * relative, but fp32 coordinates are absolute: with everything shifted far from the origin, or a
 * pin-hole aperture, rounding of the coordinates themselves reaches the margins.  The launcher
 * therefore uses the filter only while (largest |coordinate| + radius of curvature) <= 200 x
 * (smallest aperture radius) -- fp32 coordinate rounding <= 1.2e-5 of an aperture, forty times
 * below the 5e-4 margin -- and falls back to the all-fp64 kernel otherwise. */
inline bool ort_ring_filter_in_range(const DevScene& d, bool iris_before) {
    double big = std::fabs(d.bcz) + std::sqrt(d.ra2);
    big = std::fmax(big, std::fabs(d.l2_fb));
    big = std::fmax(big, std::fabs(d.l2_cz) + std::sqrt(d.l2_R2));
    big = std::fmax(big, std::fabs(d.l3_c1z) + std::sqrt(d.l3_R1_2));
    big = std::fmax(big, std::sqrt(d.lens_r2));
    double small = std::fmin(std::sqrt(d.l2_radius2), std::sqrt(d.l3_radius2));
    if (iris_before) {
        big = std::fmax(big, std::fabs(d.l3_iris1_z));
        small = std::fmin(small, std::sqrt(d.l3_iris_r2));
    }
    return std::isfinite(big) && small > 0.0 && big <= 200.0 * small;
}

/* fp32 variant: every hoisted scalar is computed in double above and rounded once here.
 * DevSceneT<R> is N reals followed by 4 int32, in the same order for every R. */
inline void ort_scene_to_float(const DevScene& s, DevSceneT<float>& d) {
    constexpr size_t N = (sizeof(DevScene) - 4 * sizeof(int32_t)) / sizeof(double);
    static_assert(sizeof(DevScene) == N * sizeof(double) + 4 * sizeof(int32_t), "DevScene layout");
    static_assert(sizeof(DevSceneT<float>) == N * sizeof(float) + 4 * sizeof(int32_t), "DevSceneT<float> layout");
    const double* sp = reinterpret_cast<const double*>(&s);
    float* dp = reinterpret_cast<float*>(&d);
    for (size_t i = 0; i < N; ++i) dp[i] = (float)sp[i];
    std::memcpy(dp + N, sp + N, 4 * sizeof(int32_t));
}

inline void ort_make_dev_job(const ort_job& j, int nscenes, int64_t first, int64_t n, DevJob& d) {
    memset(&d, 0, sizeof d);
    d.seed = j.seed;
    d.first_ray = first;
    d.nrays = n;
    d.uniform_override = j.uniform_override;
    d.phase = j.phase;
    d.use_bottle = j.use_bottle;
    d.iris_before = j.iris_before;
    d.iris_after = j.iris_after;
    d.nscenes = nscenes;
    d.stop_after = j.stop_after;
    d.flags = j.flags;
    d.source_kind = j.source_kind;
    d.total_rays = j.total_rays > 0 ? j.total_rays : j.nrays;
    d.image_cdf = nullptr; /* the launcher fills in its device copy */
    for (int r = 0; r < 10; ++r) {
        d.round_keys[2 * r] = (uint32_t)j.seed + (uint32_t)r * 0x9E3779B9u;
        d.round_keys[2 * r + 1] = (uint32_t)(j.seed >> 32) + (uint32_t)r * 0xBB67AE85u;
    }
}


#endif
