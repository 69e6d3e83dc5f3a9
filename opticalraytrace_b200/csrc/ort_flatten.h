/* ort_flatten.h -- host-only: public scene + job -> the flattened constant-memory form the
 * kernels read (DevScene / DevJob, ort_dev_types.h).  Every launch-invariant scalar is hoisted
 * here, once per launch. */
#ifndef ORT_FLATTEN_H
#define ORT_FLATTEN_H

#include <cmath>
#include <cstddef>
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
inline bool ort_ring_aim_cut(const DevScene& d, unsigned long long* cut, bool fp32 = false) {
    auto outside = [&](unsigned long long bits) {
        if (fp32) { /* the fp32 variant's own expression (ort_bits_to_uniform<float>, float scene): also monotone */
            float u2 = fminf((float)bits * (1.0f / 9007199254740992.0f), 0.99999994f);
            return u2 * (float)d.lens_r2 > (float)d.l2_radius2;
        }
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

/* ------------------------------------------------------------------------------------------
 * Error-bound constants of the ring loop's single-precision filter (ort_filter.cuh; derivation in
 * DESIGN.md section 3.1c).  Everything here is a bound on |fp32 value - exact value| that holds for
 * every ray of the launch, computed in double and rounded UP to float.  u = 2^-24; where u
 * multiplies a scene length it is counted twice (U = 2u: the rounding of the fp32 scene constant
 * and the rounding of the operation), and every constant carries a factor >= 1.01.
 * K.usable = 0: a premise of the filter fails or a constant is not finite (the constants mean nothing);
 * 1: the bounds hold but are too large to decide much; 2: use the filter.  The launcher runs the
 * all-fp64 kernel unless it is 2.
 * ---------------------------------------------------------------------------------------- */
inline float ort_up(double x) { /* >= x as a float; NaN / negative / overflow -> +inf */
    if (!(x >= 0.0) || x > 1e30) return INFINITY;
    return (float)(x * 1.0000003); /* more than the 2^-24 a round-to-nearest conversion can lose */
}
/* L2's flat face: eta < 1, normal exactly (0,0,-1) (checked by ort_make_filter).  All bounds are linear in
 * the bound ed of the incoming direction, ed < 2^-7 (guard G1). */
inline bool ort_filter_flat(const DevIface& f, DevFilterFlat& o) {
    const double u = 5.9604644775390625e-8, E_RSQ = 2.6e-7;
    const double eta = f.eta, eta2 = f.eta2, g = 1.0 - eta2;
    /* sin^2 = 1 - dz^2:  es2 = ed (2 dz + ed) + u <= 2.02 ed + u */
    o.s2_a = ort_up(2.02);
    o.s2_b = ort_up(1.01 * u);
    /* cos^2 theta_t = 1 - eta^2 sin^2 >= g exactly; in fp32 / on the segment >= gmin (es2 <= 0.0159 under G1) */
    const double gmin = g - eta2 * (1.01 * 0.0159 + 2.0 * u) - u;
    if (!(eta < 1.0) || !(gmin > 0.0)) return false;
    const double ICT = 1.0 / std::sqrt(gmin);
    /* R(c1) = (rs^2 + rp^2) / 2, rs = A / B, rp = C / D:  rs' = 2 eta (1 - eta^2) / (c2 B^2), rp' = -2 eta (1 - eta^2) / (c2 D^2),
     * |rs|, |rp| <= 1, B >= c2 >= sqrt(gmin), D >= eta c2:  |dR / dc1| <= DR = 2 eta g ICT (1 + 1 / eta^2) / gmin.
     * lhs - num = 2 den (u - R): certain when |lhs - num| > den (2 DR ed + 2 |du| + 26u rounding) */
    const double DR = 2.0 * eta * g * ICT * (1.0 + 1.0 / eta2) / gmin;
    o.f_a = ort_up(2.0 * 1.05 * DR);
    o.f_b = ort_up(30.0 * u);
    /* T = (eta dx, eta dy, cos theta_t):  dT/dI = diag(eta, eta, eta^2 dz / c2), norm eta (eta dz <= c2);
     * evaluation: cos^2 within (2 eta^2 + 1) u, its root within that / (2 c2) + (E_RSQ + u); eta dx within 2u eta */
    o.d_a = ort_up(1.1 * eta);
    o.d_b = ort_up(1.01 * ((2.0 * eta2 + 1.0) * u * ICT / 2.0 + E_RSQ + u) + 2.02 * u * eta);
    return true;
}
/* L2's curved face: eta > 1 (checked by ort_make_filter) */
inline bool ort_filter_exit(const DevIface& f, DevFilterIface& o) {
    const double u = 5.9604644775390625e-8, E_RSQ = 2.6e-7;
    const double eta = f.eta, eta2 = f.eta2;
    if (!(eta > 1.0)) return false;
    o.ni_0 = ort_up(4.0 * u);                                   /* N.I: three products, |N|, |I| <= 1.01 */
    /* cos^2 theta_t = 1 - eta^2 s2, s2 known to es2 = 2.02 eni + u:  ect2 = eta^2 (1.01 es2 + u) + 1.01 u eta^2, held x 1.01 */
    o.ct2_a = ort_up(1.01 * eta2 * 1.01 * 2.02);
    o.ct2_b = ort_up(1.01 * (eta2 * 2.01 * u + 1.01 * u * eta2));
    o.cs_0 = ort_up(1.01 * (E_RSQ + u));
    /* without total reflection eta c1 > sqrt(eta^2 - 1): B >= eta c1, D >= c1, so 1 / B^2 + 1 / D^2 <= (1 + eta^2) / (eta^2 - 1) and
     * |dR / dc1| <= 2 eta (eta^2 - 1) / c2 * that = 2 eta (1 + eta^2) / c2; 1.07: c2 over the segment (G7) */
    o.f_a = ort_up(2.0 * 1.05 * 1.07 * 2.0 * eta * (1.0 + eta2));
    o.f_b = ort_up(30.0 * u);
    o.k_a = ort_up(1.01 * eta);                                 /* k = eta c1 - c2 */
    o.k_0 = ort_up(4.04 * u * (1.0 + eta));
    /* T = eta I + k N' (ort_filter.cuh, ortf_exit_face):
     *   propagated: 1.1 eta (eta c1 / c2) ed + 1.1 sqrt2 max(1, kmax) (|k| / c2) en,  kmax = sqrt(eta^2 - 1) >= |k|
     *   evaluation: N.I within 3.03u; s2 within 7.1u; ct2 within k2 u, k2 = 9.2 eta^2; c2 within 1.01 k2 u / (2 c2)
     *               + (E_RSQ + u) c2;  k and the three fma: 10u eta */
    const double kmax = std::sqrt(eta2 - 1.0), k2 = 9.2 * eta2;
    o.d_d = ort_up(1.1 * eta);
    o.d_n = ort_up(1.1 * 1.4143 * std::fmax(1.0, kmax));
    o.d_a = ort_up(1.01 * k2 * u / 2.0);
    o.d_0 = ort_up(1.01 * (E_RSQ + u) + 10.1 * u * eta);
    return true;
}
/* centre (cx,cy,cz), radius R; LS >= distance from any admissible start point to the centre; PM >= distance of
 * any point of the sphere from the origin; ep0 >= 0: the start point's own bound when it is a scene constant
 * (L2's flat face), folded into h_0, c_0, d_0, p_0 */
inline void ort_filter_sphere(double cx, double cy, double cz, double R, double LS, double PM, double ep0,
                              DevFilterSphere& o) {
    const double u = 5.9604644775390625e-8, U = 2.0 * u;
    const double Cn = std::sqrt(cx * cx + cy * cy + cz * cz);
    const double L = 1.02 * LS;         /* |l~|: the start point is within ep <= ep_max of where it should be */
    const double el0 = ep0 + U * (Cn + L); /* l~ = pos~ - centre~: el = ep + el0 */
    const double c0 = 2.02 * L * el0 + 6.0 * u * (L * L + R * R);
    const double d0 = 1.01 * 1.01 * u * (2.0 * L * L + R * R);
    o.h_d = ort_up(1.01 * L);                                     /* h = dir . l */
    o.h_p = ort_up(1.01);
    o.h_0 = ort_up(1.01 * el0 + 3.2 * u * L);
    o.c_p = ort_up(2.02 * L);                                     /* c = l . l - R^2 */
    o.c_0 = ort_up(c0);
    o.d_h = ort_up(1.01 * 2.04 * L);                              /* disc = h^2 - c, held x 1.01 */
    o.d_0 = ort_up(ep0 > 0.0 ? d0 + 1.01 * c0 : d0);              /* from the flat face bound(c) is the constant c_0 */
    o.p_0 = ort_up(ep0 + 1.01 * U * PM);                          /* pos' = pos + t dir */
    o.n_p = ort_up(1.01 / R);                                     /* normal = (centre - pos') / R */
    o.n_0 = ort_up(1.01 * U * (Cn + R) / R + 2.02 * u);
}
inline void ort_make_filter(const DevScene& d, bool iris_before, DevFilter& K) {
    const double u = 5.9604644775390625e-8, U = 2.0 * u;
    const double E_RSQ = 2.6e-7, E_SQRT = 2.384185791015625e-7, E_SIN = 1.1e-6;
    const double PI2 = 6.283185307179586;
    memset(&K, 0, sizeof K);
    /* ---- ring source position: (rr cos, rr sin, bcz + sqrt(Ra^2 - q^2)), rr^2 = r1 + u0 (r2 - r1) */
    const double r1 = d.r1, D = d.r2_m_r1, r2 = r1 + D;
    const double rmax2 = std::fmax(r1, r2), rmin2 = std::fmin(r1, r2), rmax = std::sqrt(rmax2);
    const double E_rr2 = 2.0 * U * (std::fabs(D) + rmax2);            /* draw within 2^-23, constants, fma */
    const double E_rr = E_rr2 / std::sqrt(rmin2) + 1.01 * E_SQRT * rmax;
    const double E_trig = PI2 * U + E_SIN;                            /* angle within 4 pi u, MUFU sin / cos */
    const double E_pxy = 1.01 * rmax * (E_trig + u) + E_rr;
    const double kappa = d.ellipse ? d.ra_over_rb : 1.0;
    const double E_q = kappa * E_pxy + (d.ellipse ? 2.0 * u * kappa * rmax : 0.0);
    const double qmax = 1.01 * kappa * rmax, Ra = std::sqrt(d.ra2);
    const double wmin = d.ra2 - qmax * qmax;
    const double E_w = 2.02 * qmax * E_q + 2.0 * u * d.ra2;
    const double E_sw = E_w / std::sqrt(0.99 * wmin) + 1.01 * E_SQRT * Ra;
    const double E_pz = E_sw + U * (std::fabs(d.bcz) + Ra);
    const double EP0 = 1.4143 * E_pxy + E_pz;
    /* ---- aim point rl (cos, sin), rl = aim2 * rsqrt(aim2) <= L2.radius, draw >= 2^-16 (guard G5) */
    const double lens_r = std::sqrt(d.lens_r2), l2r = std::sqrt(d.l2_radius2);
    const double E_rl = 1.02 * l2r * (2.5 * u + E_RSQ) + 1.02 * std::ldexp(1.0, -25) * lens_r;
    const double EA = 1.01 * l2r * 1.4143 * (E_trig + u) + E_rl;
    /* the flat face is the aim plane only to 4e-16 (ring_shortcut): the exact path meets it this far
     * from the aim point at most */
    const double pzmax = d.bcz + Ra, pzmin = d.bcz + std::sqrt(std::fmax(wmin, 0.0));
    const double Dmin = d.l2_fb - pzmax;
    const double mismatch = 4.1e-16 * std::fabs(d.l2_fb) * (l2r + rmax) / std::fmax(Dmin, 1e-300);
    const double ep_flat = EA + U * std::fabs(d.l2_flat_z) + mismatch;
    K.ep_flat = ort_up(ep_flat);
    /* ---- direction (aim - source) / |aim - source| */
    const double Emax = std::hypot(l2r + rmax, d.l2_fb - pzmin);
    const double EE = EA + EP0 + U * std::fabs(d.l2_fb) + 2.0 * u * Emax;
    K.ed_a = ort_up(2.02 * EE);
    K.ed_b = ort_up(1.01 * (5.0 * u + E_RSQ));
    /* scene constants that multiply a raw word: the fp32 scene's value (rounded once from double, as
     * ort_scene_to_float does) times 2^-32, exact */
    K.r2m_s = (float)d.r2_m_r1 * 2.3283064365386963e-10f;
    K.lens_r2_s = (float)d.lens_r2 * 2.3283064365386963e-10f;
    /* ---- L2's two faces: premises of the specialised code */
    bool ok = ort_filter_flat(d.l2_in, K.flat) && ort_filter_exit(d.l2_out, K.curved);
    ok = ok && d.l2_fnx == 0.0 && d.l2_fny == 0.0 && d.l2_fnz == -1.0 && d.l2_cx == 0.0 && d.l2_cy == 0.0;
    ok = ok && d.l3_c1x == 0.0 && d.l3_c1y == 0.0; /* both sphere centres on the axis */
    /* ---- spheres.  L2: from the flat face inside the aperture.  L3 surface 1: from L2's sphere. */
    const double R2 = std::sqrt(d.l2_R2), R1 = std::sqrt(d.l3_R1_2);
    const double c2n = std::fabs(d.l2_cz);
    const double c3n = std::sqrt(d.l3_c1x * d.l3_c1x + d.l3_c1y * d.l3_c1y + d.l3_c1z * d.l3_c1z);
    const double LS2 = std::hypot(l2r, d.l2_flat_z - d.l2_cz);
    const double dc = std::sqrt(d.l3_c1x * d.l3_c1x + d.l3_c1y * d.l3_c1y + (d.l3_c1z - d.l2_cz) * (d.l3_c1z - d.l2_cz));
    ort_filter_sphere(d.l2_cx, d.l2_cy, d.l2_cz, R2, LS2, c2n + R2, ep_flat, K.s2);
    ort_filter_sphere(d.l3_c1x, d.l3_c1y, d.l3_c1z, R1, dc + R2, c3n + R1, 0.0, K.s3);
    /* ---- aperture / iris: |rho~^2 - rho*^2| <= ep (2 rho + ep) <= 1.02 ep (rho^2 / r + r) for ep <= r / 64;
     * + 4u rho^2 of rounding; threshold within u r^2 */
    const double l3r = std::sqrt(d.l3_radius2), ir = std::sqrt(d.l3_iris_r2);
    K.ap_inv = ort_up(1.02 / l3r);
    K.ap_r = ort_up(1.02 * l3r);
    K.ap_0 = ort_up(U * d.l3_radius2);
    double small = std::fmin(std::fmin(l2r, l3r), std::fmin(R2, R1));
    if (iris_before) {
        K.iris_inv = ort_up(1.02 / ir);
        K.iris_r = ort_up(1.02 * ir);
        K.iris_0 = ort_up(U * d.l3_iris_r2);
        K.iris_z0 = ort_up(U * (2.0 * std::fabs(d.l3_iris1_z) + c2n + R2));
        small = std::fmin(small, ir);
    }
    K.ep_max = (float)(small / 64.0);
    /* usable: the premises hold, every constant is finite, the geometry is non-degenerate, and the bounds are
     * small enough for the filter to decide anything (a system far from the origin fails here; so does an
     * index ratio within ~1 % of 1, which leaves the Fresnel decision no provable margin) */
    ok = ok && wmin > 0.0 && rmin2 > 0.0 && Dmin > 0.0 && small > 0.0;
    const float* kf = reinterpret_cast<const float*>(&K);
    for (size_t i = 0; i < offsetof(DevFilter, usable) / sizeof(float); ++i) ok = ok && std::isfinite(kf[i]);
    /* valid (1): the bounds hold.  useful (2): ... and are small enough to be worth a culling pass */
    bool useful = ok && 16.0 * (ep_flat + K.s2.p_0 + K.s3.p_0) < K.ep_max &&
                  2.02 * EE / std::fmax(Dmin, 1e-300) < 1.0 / 512.0 && K.flat.f_a < 1e3f && K.curved.f_a < 1e3f;
    K.usable = useful ? 2 : ok ? 1 : 0;
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
