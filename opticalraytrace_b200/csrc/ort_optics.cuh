/*
 * ort_optics.cuh -- per-ray optics of the trace loop, written for one-thread-per-ray fp64 on
 * sm_100a.  Each function names the reference routine whose result it reproduces (to 1e-9
 * relative; see DESIGN.md "numerics" for why the two need not be bit-identical).
 *
 * This is not a transliteration of the Fortran.  What changed and why (fp64 div / sqrt / sincos
 * are 15-70 instruction sequences on the FP64 pipe, FMAs are 1):
 *   - quadratics are solved in half-b form with ONE division (c/q) for unit directions and one
 *     reciprocal for the cylinder/ellipse (q/a and c/q share 1/(a*q));
 *   - Fresnel + Snell are fused: the reference's fresnel() and refract() each take
 *     sqrt(1 - eta^2 sin^2), here it is taken once, and the two squared amplitude ratios share
 *     one division; sin(theta_i) is never formed (the TIR test is done on eta^2 sin^2);
 *   - surface normals on spheres are (centre - pos) * (1/R) with 1/R hoisted to the host
 *     (the point is on the sphere by construction), no sqrt + 3 divisions;
 *   - aperture / iris tests compare squared radii; the fibre-NA test compares
 *     dz^2 >= cos^2(asin 0.22) |d|^2 instead of acos() > asin();
 *   - sin/cos of 2*pi*u use sincospi (no range-reduction slow path);
 *   - every launch-invariant scalar comes pre-computed in DevScene.
 *
 * The header is also compiled for the host by tests/host_harness (test-only) so the
 * reformulated arithmetic can be checked against the oracle without a GPU; the product never
 * runs it on the CPU.
 */
#ifndef ORT_OPTICS_CUH
#define ORT_OPTICS_CUH

#include <math.h>
#include <stdint.h>

#include "ort_dev_types.h"

#if defined(__CUDACC__)
#define ORT_HD __host__ __device__ __forceinline__
#else
#define ORT_HD inline
#endif

#define ORT_PI 3.14159265358979323846
#define ORT_TWOPI 6.28318530717958647692

/* -------------------------------------------------------------------------------------------
 * small wrappers over device intrinsics (host versions only serve the test harness)
 * ----------------------------------------------------------------------------------------- */
ORT_HD void ort_sincospi(double x, double* s, double* c) {
#ifdef __CUDA_ARCH__
    sincospi(x, s, c);
#else
    *s = sin(ORT_PI * x);
    *c = cos(ORT_PI * x);
#endif
}
ORT_HD void ort_sincos(double x, double* s, double* c) {
#ifdef __CUDA_ARCH__
    sincos(x, s, c);
#else
    *s = sin(x);
    *c = cos(x);
#endif
}
ORT_HD uint32_t ort_mulhi(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
#endif
}

/* -------------------------------------------------------------------------------------------
 * fp64 reciprocal / division / square root without the compiler's special-case slow paths.
 * nvcc's `a / b` and `sqrt(x)` are correctly rounded and carry a range check plus a call into a
 * denormal/overflow subroutine (a BSSY/CALL/BSYNC cluster per use -- the 41 CALLs of the first
 * build).  Every operand on this path is a normal number between ~1e-12 and ~1e3, so the
 * MUFU seed + Newton/Goldschmidt refinement below is enough: <= 1 ulp (measured on the device by
 * ort_math_selftest), far inside the 1e-9 parity tolerance.  Zero and NaN keep the meaning the
 * callers rely on: ort_sqrt(0) = 0, ort_sqrt(<0) = NaN, ort_div(x, 0) = +-inf.
 * ----------------------------------------------------------------------------------------- */
#ifdef __CUDA_ARCH__
__device__ __forceinline__ double ort_mufu_rcp(double x) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    return y;
}
__device__ __forceinline__ double ort_mufu_rsq(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    return y;
}
#endif
#ifndef ORT_NR
#define ORT_NR 2 /* Newton / Goldschmidt refinement steps after the MUFU seed */
#endif
ORT_HD double ort_rcp(double x) {
#ifdef __CUDA_ARCH__
    double y = ort_mufu_rcp(x);
#pragma unroll
    for (int i = 0; i < ORT_NR; ++i) {
        double e = fma(-x, y, 1.0);
        y = fma(y, e, y);
    }
    return y;
#else
    return 1.0 / x;
#endif
}
ORT_HD double ort_div(double a, double b) {
#ifdef __CUDA_ARCH__
    /* one Newton step on the reciprocal, then the residual correction of the quotient (itself a
     * quadratic step): the seed's ~2^-23 becomes < 2^-60 */
    double y = ort_mufu_rcp(b);
    double e = fma(-b, y, 1.0);
    y = fma(y, e, y);
#if ORT_NR >= 3
    e = fma(-b, y, 1.0);
    y = fma(y, e, y);
#endif
    double q = a * y;
    double r = fma(-b, q, a);
    return fma(r, y, q);
#else
    return a / b;
#endif
}
/* division whose divisor may be exactly zero (ray parallel to a plane): +-inf like IEEE */
ORT_HD double ort_div_z(double a, double b) {
#ifdef __CUDA_ARCH__
    double q = ort_div(a, b);
    if ((__double_as_longlong(b) << 1) == 0ll) q = (a > 0.0) ? INFINITY : ((a < 0.0) ? -INFINITY : NAN);
    return q;
#else
    return a / b;
#endif
}
ORT_HD double ort_rsqrt(double x) {
#ifdef __CUDA_ARCH__
    double y = ort_mufu_rsq(x);
    double g = x * y, h = 0.5 * y;
#pragma unroll
    for (int i = 0; i < ORT_NR; ++i) {
        double r = fma(-g, h, 0.5);
        if (i + 1 < ORT_NR) g = fma(g, r, g);
        h = fma(h, r, h);
    }
    return h + h;
#else
    return 1.0 / sqrt(x);
#endif
}
/* Goldschmidt: g -> sqrt(x), h -> 1/(2 sqrt(x)) */
ORT_HD double ort_sqrt(double x) {
#ifdef __CUDA_ARCH__
    /* one Goldschmidt step, then the Newton correction of the root (also quadratic) */
    double y = ort_mufu_rsq(x);
    double g = x * y, h = 0.5 * y;
#pragma unroll
    for (int i = 0; i < ORT_NR - 1; ++i) {
        double r = fma(-g, h, 0.5);
        g = fma(g, r, g);
        h = fma(h, r, h);
    }
    double d = fma(-g, g, x);
    g = fma(d, h, g);
    return (x == 0.0) ? 0.0 : g;
#else
    return sqrt(x);
#endif
}

/* -------------------------------------------------------------------------------------------
 * Counter-based uniforms -- replaces the reference's ran2() (src/random_mod.f90:39-46).
 * Philox4x32-10, counter = (ray_lo, ray_hi, phase, block), key = (seed_lo, seed_hi).
 * Block b yields draw slots 2b and 2b+1; u = (64 bits >> 11) * 2^-53 in [0,1).
 * Slot map (fixed, independent of control flow):
 *   0,1  source (ring: r, theta | point: phi, cos theta)
 *   2,3  ring: aim-disc r, theta | point: bottle inner, outer reflect_refract
 *   4,5  L2 flat, L2 curved      6,7,8  L3 surfaces 1,2,3
 *   16.. scatter loops (tauint / albedo / stokes), consumed sequentially
 * ----------------------------------------------------------------------------------------- */
struct OrtRng {
    uint32_t k0, k1;   /* seed */
    uint32_t r0, r1;   /* ray index */
    uint32_t phase;
    double override_u; /* >= 0: every draw returns this */
};

ORT_HD void ort_philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                              uint32_t k1, uint32_t* o) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = ort_mulhi(M0, c0), lo0 = M0 * c0;
        uint32_t hi1 = ort_mulhi(M1, c2), lo1 = M1 * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += W0; k1 += W1;
    }
    o[0] = c0; o[1] = c1; o[2] = c2; o[3] = c3;
}

ORT_HD double ort_bits_to_uniform(uint32_t lo, uint32_t hi) {
    uint64_t bits = (((uint64_t)hi << 32) | lo) >> 11;
    return (double)bits * (1.0 / 9007199254740992.0);
}

/* the two uniforms of Philox block `block`: slots 2*block and 2*block+1 */
ORT_HD void ort_draw2(const OrtRng& g, uint32_t block, double* ua, double* ub) {
    if (g.override_u >= 0.0) {
        *ua = g.override_u;
        *ub = g.override_u;
        return;
    }
    uint32_t w[4];
    ort_philox4x32_10(g.r0, g.r1, g.phase, block, g.k0, g.k1, w);
    *ua = ort_bits_to_uniform(w[0], w[1]);
    *ub = ort_bits_to_uniform(w[2], w[3]);
}

/* sequential draws for the scatter loops (slots 16, 17, ...) */
struct OrtScatterRng {
    uint32_t next; /* next slot */
    double spare;  /* odd-slot value of the last generated block */
};
ORT_HD double ort_scatter_draw(const OrtRng& g, OrtScatterRng& s) {
    if (g.override_u >= 0.0) return g.override_u;
    uint32_t slot = s.next++;
    if (slot & 1u) return s.spare;
    double a, b;
    ort_draw2(g, slot >> 1, &a, &b);
    s.spare = b;
    return a;
}

struct OrtRay {
    double px, py, pz, dx, dy, dz;
};

/* -------------------------------------------------------------------------------------------
 * Quadratic root selection -- reproduces solveQuadratic + the root picking shared by the
 * reference's intersect_* (src/surfaces.f90:227-260 and :74-87).  Half-b form:
 * a t^2 + 2 h t + c = 0.  `inv_aq_needed`: a != 1 (cylinder / ellipse).
 * ----------------------------------------------------------------------------------------- */
/* sign-bit tests on the high word: integer ALU work instead of a DSETP on the FP64 pipe.  Only
 * used where a signed zero cannot occur (differences of distinct quantities). */
ORT_HD bool ort_signbit(double x) {
#ifdef __CUDA_ARCH__
    return __double2hiint(x) < 0;
#else
    return signbit(x);
#endif
}
ORT_HD bool ort_either_negative(double x, double y) {
#ifdef __CUDA_ARCH__
    return (__double2hiint(x) | __double2hiint(y)) < 0;
#else
    return signbit(x) || signbit(y);
#endif
}
ORT_HD bool ort_is_zero(double x) { /* +0 only */
#ifdef __CUDA_ARCH__
    return __double_as_longlong(x) == 0ll;
#else
    return x == 0.0 && !signbit(x);
#endif
}

ORT_HD bool ort_both_zero(double x, double y) { /* +-0 */
#ifdef __CUDA_ARCH__
    return ((__double_as_longlong(x) | __double_as_longlong(y)) << 1) == 0ll;
#else
    return x == 0.0 && y == 0.0;
#endif
}

/* Unit normal of a sphere at a point on it: (centre - pos) / R with 1/R hoisted.  A ray that runs
 * exactly along the axis must see EXACTLY (0,0,+-1): the reference normalises by the computed
 * length, gets |N.I| == 1 bit for bit, and its fresnel() then returns 0 -- the
 * on-axis rays of create_spot and of the known-answer tests depend on it. */
ORT_HD void ort_sphere_normal(const OrtRay& r, double cx, double cy, double cz, double invR, double* nx,
                              double* ny, double* nz) {
    double vx = cx - r.px, vy = cy - r.py, vz = cz - r.pz;
    *nx = vx * invR;
    *ny = vy * invR;
    *nz = ort_both_zero(vx, vy) ? copysign(1.0, vz) : vz * invR;
}

/* The reference sorts the two roots, takes the smaller unless it is negative, and misses when
 * both are (src/surfaces.f90:75-84).  With the stable pair q = -(h + sgn(h) s), {q/a, c/q} the
 * outcome is decided by the signs of h and c alone, and only ONE quotient is ever needed:
 *   h > 0 : q < 0, so q/a < 0 and c/q has the sign of -c:  c > 0 -> miss, else t = c/q
 *   h <= 0: q >= 0:  c < 0 -> c/q < 0, t = q/a ;  c >= 0 -> both >= 0 and c/q is the smaller */
ORT_HD bool ort_pick_root_unit(double h, double c, double* t) {
    /* a == 1 (unit direction): q/a = q */
    double disc = fma(h, h, -c);
    if (ort_signbit(disc)) return false;
    double s = ort_sqrt(disc);
    bool hpos = h > 0.0;
    if (hpos && c > 0.0) return false;
    double q = hpos ? -(h + s) : (s - h);
    double x1 = ort_div(c, q);
    double tt = (!hpos && c < 0.0) ? q : x1;
    if (q == 0.0) tt = 0.0; /* h = c = 0: the ray starts on the surface, tangent */
    *t = tt;
    return true;
}
ORT_HD bool ort_pick_root(double a, double h, double c, double* t) {
    double disc = fma(h, h, -a * c);
    if (ort_signbit(disc)) return false;
    double s = ort_sqrt(disc);
    bool hpos = h > 0.0;
    if (hpos && c > 0.0) return false;
    double q = hpos ? -(h + s) : (s - h);
    bool use_q = !hpos && c < 0.0;
    double tt = ort_div(use_q ? q : c, use_q ? a : q);
    if (!(tt >= 0.0)) return false; /* a = 0 (ray along the axis) or q = 0 */
    *t = tt;
    return true;
}

/* intersect_sphere (src/surfaces.f90:52-89) for a unit direction */
ORT_HD bool ort_hit_sphere(const OrtRay& r, double cx, double cy, double cz, double R2, double* t) {
    double lx = r.px - cx, ly = r.py - cy, lz = r.pz - cz;
    double h = fma(r.dx, lx, fma(r.dy, ly, r.dz * lz));
    double c = fma(lx, lx, fma(ly, ly, fma(lz, lz, -R2)));
    return ort_pick_root_unit(h, c, t);
}
/* intersect_cylinder (src/surfaces.f90:91-130): axis along x, only (y,z) enter */
ORT_HD bool ort_hit_cylinder(const OrtRay& r, double cy, double cz, double R2, double* t) {
    double ly = r.py - cy, lz = r.pz - cz;
    double a = fma(r.dz, r.dz, r.dy * r.dy);
    double h = fma(r.dz, lz, r.dy * ly);
    double c = fma(lz, lz, fma(ly, ly, -R2));
    return ort_pick_root(a, h, c, t);
}
/* intersect_ellipse (src/surfaces.f90:133-176): ia2 = 1/semia^2 (z), ib2 = 1/semib^2 (y) */
ORT_HD bool ort_hit_ellipse(const OrtRay& r, double cy, double cz, double ia2, double ib2, double* t) {
    double ly = r.py - cy, lz = r.pz - cz;
    double a = fma(ia2 * r.dz, r.dz, ib2 * r.dy * r.dy);
    double h = fma(ia2 * r.dz, lz, ib2 * r.dy * ly);
    double c = fma(ia2 * lz, lz, fma(ib2 * ly, ly, -1.0));
    return ort_pick_root(a, h, c, t);
}

ORT_HD void ort_advance(OrtRay& r, double t) {
    r.px = fma(r.dx, t, r.px);
    r.py = fma(r.dy, t, r.py);
    r.pz = fma(r.dz, t, r.pz);
}

/* -------------------------------------------------------------------------------------------
 * One dielectric interface: reflect_refract = fresnel + reflect | refract
 * (src/surfaces.f90:262-372) fused.  (nx,ny,nz) is a unit normal of either orientation.
 * Returns true when the ray was reflected.  Branch semantics kept from the reference:
 *   R = 1 when eta*sin > 1 (TIR) or when 1 - cos^2 < 0 (the reference's sqrt gives NaN there and
 *   its NaN guard turns that into 1);  R = 0 at exactly normal incidence (cos == 1);
 *   reflect when u <= R.
 * ----------------------------------------------------------------------------------------- */
ORT_HD bool ort_interface(OrtRay& r, double nx, double ny, double nz, const DevIface& f, double u) {
    double c = fma(nx, r.dx, fma(ny, r.dy, nz * r.dz)); /* N . I */
    double costt = fabs(c);
    double s2 = fma(-costt, costt, 1.0); /* sin^2(theta_i) */
    double ct2 = fma(-f.eta2, s2, 1.0);  /* cos^2(theta_t) = 1 - eta^2 sin^2 */
    double cost2 = ort_sqrt(fmax(ct2, 0.0));
    /* Fresnel amplitudes in units of nb (the ratios do not change): A/B = r_s, C/D = r_p.
     * R = (A^2 D^2 + C^2 B^2) / (2 B^2 D^2); the draw is compared without forming the quotient:
     *   u > R  <=>  2u B^2 D^2 > A^2 D^2 + C^2 B^2.
     * 0 <= R <= 1 by construction (|A| <= B, |C| <= D); a NaN fails the comparison and reflects,
     * which is what the reference's NaN guard (R = 1) does. */
    double ec = f.eta * costt, e2 = f.eta * cost2;
    double A = ec - cost2, B = ec + cost2, C = e2 - costt, D = e2 + costt;
    double B2 = B * B, D2 = D * D;
    double num = fma(A * A, D2, (C * C) * B2);
    double lhs = (u + u) * (B2 * D2);
    bool transmit = lhs > num;
    if (ort_either_negative(ct2, s2)) transmit = false; /* TIR, or |N.I| > 1 by rounding (reference: NaN -> R = 1) */
    else if (ort_is_zero(s2)) transmit = u > 0.0;      /* exactly normal incidence: the reference returns R = 0 */
    if (!transmit) { /* reflect, src/surfaces.f90:285-300 */
        double k = -2.0 * c;
        r.dx = fma(k, nx, r.dx);
        r.dy = fma(k, ny, r.dy);
        r.dz = fma(k, nz, r.dz);
        return true;
    }
    /* refract, src/surfaces.f90:303-333: T = eta I + (eta c1 - c2) N', N' opposing I */
    double k = (c < 0.0) ? A : -A; /* eta c1 - c2 */
    r.dx = fma(f.eta, r.dx, k * nx);
    r.dy = fma(f.eta, r.dy, k * ny);
    r.dz = fma(f.eta, r.dz, k * nz);
    return false;
}

/* -------------------------------------------------------------------------------------------
 * Sources (src/sourceMod.f90)
 * ----------------------------------------------------------------------------------------- */
/* point, src/sourceMod.f90:12-47 */
ORT_HD void ort_source_point(const DevScene& S, const OrtRng& g, OrtRay& r) {
    double u0, u1, sp, cp;
    ort_draw2(g, 0, &u0, &u1);
    ort_sincospi(2.0 * u0, &sp, &cp);
    double cost = fma(u1, S.cos_theta_max, 1.0 - u1);
    double sint = ort_sqrt(fma(-cost, cost, 1.0));
    r.dx = sint * cp;
    r.dy = sint * sp;
    r.dz = cost;
    r.px = 0.0;
    r.py = 0.0;
    r.pz = S.point_offset;
}

/* ring, src/sourceMod.f90:250-300; (u0,u1) place the ray on the annulus, (u2,u3) pick the aim
 * point on the disc of radius L2.radius + 10 mm in the plane z = L2.fb */
ORT_HD void ort_source_ring_u(const DevScene& S, double u0, double u1, double u2, double u3, OrtRay& r) {
    double s, c;
    double rr = ort_sqrt(fma(u0, S.r2_m_r1, S.r1));
    ort_sincospi(2.0 * u1, &s, &c);
    double px = rr * c, py = rr * s;
    double q = S.ellipse ? py * S.ra_over_rb : py;
    double pz = S.bcz + ort_sqrt(fma(-q, q, S.ra2));
    double rl = ort_sqrt(u2 * S.lens_r2);
    ort_sincospi(2.0 * u3, &s, &c);
    double ex = fma(rl, c, -px), ey = fma(rl, s, -py), ez = S.l2_fb - pz;
    double inv = ort_rsqrt(fma(ex, ex, fma(ey, ey, ez * ez)));
    r.px = px;
    r.py = py;
    r.pz = pz;
    r.dx = ex * inv;
    r.dy = ey * inv;
    r.dz = ez * inv;
}
ORT_HD void ort_source_ring(const DevScene& S, const OrtRng& g, OrtRay& r) {
    double u0, u1, u2, u3;
    ort_draw2(g, 0, &u0, &u1);
    ort_draw2(g, 1, &u2, &u3);
    ort_source_ring_u(S, u0, u1, u2, u3, r);
}
/* When L2's flat face lies in the aim plane (DevScene.ring_shortcut) the ray meets that face AT
 * its aim point, so the aperture test of src/lens.f90:450-454 is a test on u2 alone: 69 % of the
 * ring rays of the shipped geometries end here, before any position, direction, sqrt or sincos
 * has been computed. */
ORT_HD bool ort_ring_aims_outside_aperture(const DevScene& S, double u2) {
    return u2 * S.lens_r2 > S.l2_radius2;
}

/* ---- the other sources of settings.params ---------------------------- */
/* rang, src/random_mod.f90:59-85: polar Box-Muller; its rejection loop takes the sequential
 * draws (slots 16, 17, ...) */
ORT_HD void ort_rang(const OrtRng& g, OrtScatterRng& sr, double sigma, double* x, double* y) {
    double a, b, s;
    do {
        a = fma(ort_scatter_draw(g, sr), 2.0, -1.0);
        b = fma(ort_scatter_draw(g, sr), 2.0, -1.0);
        s = fma(b, b, a * a);
    } while (s >= 1.0 && g.override_u < 0.0);
    double cst = sqrt(-2.0 * log(s) / s);
    *x = sigma * (a * cst);
    *y = sigma * (b * cst);
}

/* point_on_bottle, src/sourceMod.f90:50-89 (crs, ring loop): a Gaussian spot projected along -z
 * onto the cylinder of radius Ra + thickness, emitting into the cone of point() */
ORT_HD bool ort_source_crs(const DevScene& S, const OrtRng& g, OrtRay& r) {
    OrtScatterRng sr;
    sr.next = 16;
    sr.spare = 0.0;
    ort_source_point(S, g, r); /* same two draws, same direction formulas (:65-77) */
    double dx = r.dx, dy = r.dy, dz = r.dz, x, y, t;
    ort_rang(g, sr, S.spot_size, &x, &y);
    OrtRay probe = {x, y, 1.0, 0.0, 0.0, -1.0};
    if (!ort_hit_cylinder(probe, S.bcy, S.bcz, S.crs_r2, &t)) return false;
    r.px = x;
    r.py = y;
    r.pz = 1.0 - t;
    r.dx = dx;
    r.dy = dy;
    r.dz = dz;
    return true;
}

/* create_spot, src/sourceMod.f90:122-159 (spot, point loop): deterministic angular grid;
 * n = 1-based loop index, nrays = nphotons */
ORT_HD void ort_source_spot(const DevScene& S, long long nrays, long long n, OrtRay& r) {
    double nrays_sqrt = sqrt((double)nrays);
    double dphi = ORT_TWOPI / nrays_sqrt;
    double dtheta = acos(S.cos_theta_max) / nrays_sqrt;
    double phi = dphi * (double)(n % 10), theta = dtheta * (double)(n / 10);
    double sp, cp, st_, ct;
    ort_sincos(phi, &sp, &cp);
    ort_sincos(theta, &st_, &ct);
    double sint = sqrt(fma(-ct, ct, 1.0));
    r.dx = sint * cp;
    r.dy = sint * sp;
    r.dz = ct;
    r.px = r.py = r.pz = 0.0;
}

/* intersect_cone, src/surfaces.f90:179-224, for the axicon of iSORS */
ORT_HD bool ort_hit_cone(const OrtRay& r, double k, double height, double* t) {
    double lz = r.pz - height;
    double a = fma(r.dx, r.dx, fma(r.dy, r.dy, -k * r.dz * r.dz));
    double h = fma(r.dx, r.px, fma(r.dy, r.py, -k * r.dz * lz));
    double c = fma(r.px, r.px, fma(r.py, r.py, -k * lz * lz));
    /* a < 0 here (steep ray), so the sign logic of ort_pick_root does not apply: both roots */
    double disc = fma(h, h, -a * c);
    if (disc < 0.0) return false;
    double s = sqrt(disc);
    double q = (h > 0.0) ? -(h + s) : (s - h);
    double x0 = (disc == 0.0) ? -h / a : q / a, x1 = (disc == 0.0) ? x0 : c / q;
    double t0 = fmin(x0, x1), t1 = fmax(x0, x1);
    double tt = (t0 < 0.0) ? t1 : t0;
    if (tt < 0.0) return false;
    *t = tt;
    return true;
}

/* iSORS(ring = .true.), src/sourceMod.f90:162-247 (isors, ring loop).  false = the reference's
 * `error stop "no intersection with bottle!"` (every ray the axicon face reflects, ~2.8 %). */
ORT_HD bool ort_source_isors(const DevScene& S, const OrtRng& g, OrtRay& r) {
    OrtScatterRng sr;
    sr.next = 16;
    sr.spare = 0.0;
    double x, y, t, u_r, u_th, u_ax, unused;
    ort_rang(g, sr, S.isors_beam, &x, &y);
    r.px = x; r.py = y; r.pz = 2.0 * S.isors_h;
    r.dx = 0.0; r.dy = 0.0; r.dz = -1.0;
    ort_draw2(g, 0, &u_r, &u_th);
    ort_draw2(g, 1, &u_ax, &unused);
    if (ort_hit_cone(r, S.isors_k, S.isors_h, &t)) {
        ort_advance(r, t);
        /* gradient of the cone, inverted (upper nappe), normalised */
        double nx = -(2.0 * r.px / S.isors_k), ny = -(2.0 * r.py / S.isors_k), nz = -(-2.0 * r.pz + 2.0 * S.isors_h);
        double inv = 1.0 / sqrt(fma(nx, nx, fma(ny, ny, nz * nz)));
        (void)ort_interface(r, nx * inv, ny * inv, nz * inv, S.isors_axicon, u_ax); /* flag ignored */
        ort_advance(r, S.isors_base / r.dz);
        r.pz = S.isors_z;
        bool hit = S.ellipse ? ort_hit_ellipse(r, S.bcy, S.bcz, S.b_in_ia2, S.b_in_ib2, &t)
                             : ort_hit_cylinder(r, S.bcy, S.bcz, S.b_in_r2, &t);
        if (!hit) return false;
        ort_advance(r, t);
    }
    double rl = sqrt(u_r * S.isors_lens_r2), s, c;
    ort_sincospi(2.0 * u_th, &s, &c);
    double ex = fma(rl, c, -r.px), ey = fma(rl, s, -r.py), ez = S.l2_fb - r.pz;
    double inv = 1.0 / sqrt(fma(ex, ex, fma(ey, ey, ez * ez)));
    r.dx = ex * inv;
    r.dy = ey * inv;
    r.dz = ez * inv;
    return true;
}

/* source dispatch of src/main.f90:95-101 (ring loop) and :132-142 (point loop); SRC is
 * ort_job.source_kind.  Returns 0 or ORT_ST_SOURCE_MISS. */
template <int PHASE, int SRC>
ORT_HD int ort_emit(const DevScene& S, const DevJob& J, const OrtRng& g, long long ray, OrtRay& r) {
    if (PHASE == ORT_PHASE_RING) {
        if (SRC == ORT_SRC_CRS) return ort_source_crs(S, g, r) ? 0 : ORT_ST_SOURCE_MISS;
        if (SRC == ORT_SRC_ISORS) return ort_source_isors(S, g, r) ? 0 : ORT_ST_SOURCE_MISS;
        ort_source_ring(S, g, r);
    } else {
        if (SRC == ORT_SRC_SPOT) ort_source_spot(S, J.total_rays, ray + 1, r);
        else ort_source_point(S, g, r);
    }
    return 0;
}

/* -------------------------------------------------------------------------------------------
 * Scatter: tauint (src/surfaces.f90:13-50) and stokes (src/stokes.f90:7-166)
 * ----------------------------------------------------------------------------------------- */
/* returns false where the reference would `error stop "no intersection"` */
ORT_HD bool ort_tauint(const OrtRay& r, double mutot, double inv_mutot, double cy, double cz,
                       double R2, double u, double* dist, bool* tflag) {
    double tau = -log(u);
    double d;
    if (!ort_hit_cylinder(r, cy, cz, R2, &d)) return false;
    if (tau < d * mutot) {
        *dist = tau * inv_mutot;
        *tflag = false;
    } else {
        *dist = d;
        *tflag = true;
    }
    return true;
}

/* The spherical-triangle update below is ill-conditioned for small deflections (cosi2 is a
 * difference of nearly equal quotients), so here -- and only here -- the arithmetic keeps the
 * reference's operation order and is protected from FMA contraction (ORT_MUL / ORT_ADD): any
 * other rounding is amplified by up to ~1/sin^2 of the deflection angle. */
#ifdef __CUDA_ARCH__
#define ORT_MUL(a, b) __dmul_rn((a), (b))
#define ORT_ADD(a, b) __dadd_rn((a), (b))
#define ORT_SUB(a, b) __dsub_rn((a), (b))
#else
#define ORT_MUL(a, b) ((a) * (b))
#define ORT_ADD(a, b) ((a) + (b))
#define ORT_SUB(a, b) ((a) - (b))
#endif
ORT_HD void ort_stokes(OrtRay& r, double hgg, const OrtRng& g, OrtScatterRng& sr) {
    double cost = r.dz;
    double sint = sqrt(ORT_SUB(1.0, ORT_MUL(cost, cost)));
    double phi = atan2(r.dy, r.dx);
    double sinp, cosp;
    if (hgg == 0.0) { /* isotropic, src/stokes.f90:33-48 */
        cost = ORT_SUB(ORT_MUL(2.0, ort_scatter_draw(g, sr)), 1.0);
        sint = ORT_SUB(1.0, ORT_MUL(cost, cost));
        sint = (sint <= 0.0) ? 0.0 : sqrt(sint);
        ort_sincos(ORT_MUL(ORT_TWOPI, ort_scatter_draw(g, sr)), &sinp, &cosp);
    } else { /* Henyey-Greenstein, src/stokes.f90:54-158 */
        double g2 = ORT_MUL(hgg, hgg);
        double costp = cost, sintp = sint;
        double den = ORT_ADD(ORT_SUB(1.0, hgg), ORT_MUL(ORT_MUL(2.0, hgg), ort_scatter_draw(g, sr)));
        double tq = ORT_SUB(1.0, g2) / den;
        double bmu = ORT_SUB(ORT_ADD(1.0, g2), ORT_MUL(tq, tq)) / ORT_MUL(2.0, hgg);
        double cosb2 = ORT_MUL(bmu, bmu);
        if (fabs(bmu) > 1.0) {
            bmu = (bmu > 1.0) ? 1.0 : -1.0;
            cosb2 = 1.0;
        }
        double sinbt = sqrt(ORT_SUB(1.0, cosb2));
        double ri1 = ORT_MUL(ORT_TWOPI, ort_scatter_draw(g, sr));
        /* the reference's two branches (ri1 > pi uses ri3 = 2pi - ri1 and adds acos; otherwise
         * subtracts) differ only in the sign applied to acos(cosdph) */
        bool upper = ri1 > ORT_PI;
        double ang = upper ? ORT_SUB(ORT_TWOPI, ri1) : ri1;
        double sini, cosi;
        ort_sincos(ang, &sini, &cosi);
        if (bmu == 1.0 || bmu == -1.0) return; /* goto 100: direction unchanged */
        cost = ORT_ADD(ORT_MUL(costp, bmu), ORT_MUL(ORT_MUL(sintp, sinbt), cosi));
        double sini2, cosi2 = 0.0;
        if (fabs(cost) < 1.0) {
            sint = fabs(sqrt(ORT_SUB(1.0, ORT_MUL(cost, cost))));
            sini2 = ORT_MUL(sini, sintp) / sint;
            double bott = ORT_MUL(sint, sinbt);
            cosi2 = ORT_SUB(costp / bott, ORT_MUL(cost, bmu) / bott);
        } else {
            sint = 0.0;
            sini2 = 0.0;
            if (cost >= 1.0) cosi2 = -1.0;
            if (cost <= -1.0) cosi2 = 1.0;
        }
        double cosdph = ORT_ADD(-ORT_MUL(cosi2, cosi), ORT_MUL(ORT_MUL(sini2, sini), bmu));
        if (fabs(cosdph) > 1.0) cosdph = (cosdph > 1.0) ? 1.0 : -1.0;
        double dph = acos(cosdph);
        phi = upper ? ORT_ADD(phi, dph) : ORT_SUB(phi, dph);
        if (phi > ORT_TWOPI) phi = ORT_SUB(phi, ORT_TWOPI);
        if (phi < 0.0) phi = ORT_ADD(phi, ORT_TWOPI);
        ort_sincos(phi, &sinp, &cosp);
    }
    r.dx = ORT_MUL(sint, cosp);
    r.dy = ORT_MUL(sint, sinp);
    r.dz = cost;
}

/* -------------------------------------------------------------------------------------------
 * glass_bottle%forward, src/lens.f90:230-350.  Returns 0 or the ort_status that ended the ray.
 * ----------------------------------------------------------------------------------------- */
/* one scatter loop (contents :262-282, wall :312-333); *t is the step still to be taken */
ORT_HD int ort_scatter_loop(const DevScene& S, const OrtRng& g, OrtScatterRng& sr, OrtRay& r,
                            double mutot, double inv_mutot, double albedo, double hgg, double Rlim,
                            double Rlim2, int st_absorbed, int st_backward, double* t) {
    bool flag;
    if (!ort_tauint(r, mutot, inv_mutot, S.bcy, S.bcz, Rlim2, ort_scatter_draw(g, sr), t, &flag))
        return ORT_ST_TAUINT_MISS;
    while (!flag) {
        ort_advance(r, *t);
        if (ort_scatter_draw(g, sr) < albedo) {
            ort_stokes(r, hgg, g, sr);
        } else {
            return st_absorbed;
        }
        if (!ort_tauint(r, mutot, inv_mutot, S.bcy, S.bcz, Rlim2, ort_scatter_draw(g, sr), t, &flag))
            return ORT_ST_TAUINT_MISS;
        /* the reference's exit test uses (x,z) although the axis is x */
        if (sqrt(fma(r.px, r.px, r.pz * r.pz)) >= Rlim) break;
    }
    if (r.dz < 0.0) return st_backward;
    return 0;
}

template <bool SCATTER>
ORT_HD int ort_bottle_forward(const DevScene& S, const OrtRng& g, OrtRay& r) {
    double t, u_in, u_out;
    OrtScatterRng sr;
    sr.next = 16;
    sr.spare = 0.0;
    bool hit = S.ellipse ? ort_hit_ellipse(r, S.bcy, S.bcz, S.b_in_ia2, S.b_in_ib2, &t)
                         : ort_hit_cylinder(r, S.bcy, S.bcz, S.b_in_r2, &t);
    if (!hit) return ORT_ST_BOTTLE_INNER_MISS;
    if (SCATTER && S.scatter_c) {
        int st = ort_scatter_loop(S, g, sr, r, S.mutot_c, S.inv_mutot_c, S.albedo_c, 0.65, S.b_in_r,
                                  S.b_in_r2, ORT_ST_CONTENTS_ABSORBED, ORT_ST_CONTENTS_BACKWARD, &t);
        if (st) return st;
    }
    ort_advance(r, t);
    ort_draw2(g, 1, &u_in, &u_out);
    {   /* radial normal in the (y,z) plane, also for the ellipse (src/lens.f90:288-290) */
        double ny = S.bcy - r.py, nz = S.bcz - r.pz;
        /* on a clear cylindrical wall the hit point is on the cylinder: |(ny,nz)| = radius.  After
         * a scatter loop (quirk 4) or on an ellipse it is not, and the length is computed. */
        double inv = (SCATTER || S.ellipse) ? ort_rsqrt(fma(ny, ny, nz * nz)) : S.b_in_invr;
        double nzu = ort_both_zero(ny, 0.0) ? copysign(1.0, nz) : nz * inv; /* on-axis ray: exactly +-1 */
        if (ort_interface(r, 0.0, ny * inv, nzu, S.b_in, u_in)) return ORT_ST_BOTTLE_INNER_REFLECT;
    }
    hit = S.ellipse ? ort_hit_ellipse(r, S.bcy, S.bcz, S.b_out_ia2, S.b_out_ib2, &t)
                    : ort_hit_cylinder(r, S.bcy, S.bcz, S.b_out_r2, &t);
    if (!hit) return ORT_ST_BOTTLE_OUTER_MISS;
    if (SCATTER && S.scatter_b) {
        int st = ort_scatter_loop(S, g, sr, r, S.mutot_b, S.inv_mutot_b, S.albedo_b, 0.9, S.b_out_r,
                                  S.b_out_r2, ORT_ST_WALL_ABSORBED, ORT_ST_WALL_BACKWARD, &t);
        if (st) return st;
    }
    ort_advance(r, t);
    {
        double ny = S.bcy - r.py, nz = S.bcz - r.pz;
        double inv = (SCATTER || S.ellipse) ? ort_rsqrt(fma(ny, ny, nz * nz)) : S.b_out_invr;
        double nzu = ort_both_zero(ny, 0.0) ? copysign(1.0, nz) : nz * inv;
        if (ort_interface(r, 0.0, ny * inv, nzu, S.b_out, u_out)) return ORT_ST_BOTTLE_OUTER_REFLECT;
    }
    return 0;
}

/* -------------------------------------------------------------------------------------------
 * plano_convex%forward, src/lens.f90:425-481, split at the aperture test
 * ----------------------------------------------------------------------------------------- */
ORT_HD int ort_l2_enter(const DevScene& S, OrtRay& r) { /* :447-454 */
    double d = ort_div_z(S.l2_flat_z - r.pz, r.dz);
    ort_advance(r, d);
    if (fma(r.px, r.px, r.py * r.py) > S.l2_radius2) return ORT_ST_L2_APERTURE;
    return 0;
}
ORT_HD int ort_l2_body(const DevScene& S, const OrtRng& g, OrtRay& r) { /* :458-479 */
    double u_flat, u_curved, t;
    ort_draw2(g, 2, &u_flat, &u_curved);
    /* a reflection at the flat face is computed but never tested */
    (void)ort_interface(r, S.l2_fnx, S.l2_fny, S.l2_fnz, S.l2_in, u_flat);
    if (!ort_hit_sphere(r, S.l2_cx, S.l2_cy, S.l2_cz, S.l2_R2, &t)) return ORT_ST_L2_SPHERE_MISS;
    ort_advance(r, t);
    double nx, ny, nz;
    ort_sphere_normal(r, S.l2_cx, S.l2_cy, S.l2_cz, S.l2_invR, &nx, &ny, &nz);
    if (ort_interface(r, nx, ny, nz, S.l2_out, u_curved)) return ORT_ST_L2_CURVED_REFLECT;
    return 0;
}

/* -------------------------------------------------------------------------------------------
 * achromatic_doublet%forward, src/lens.f90:531-645, split after the first-surface aperture test
 * ----------------------------------------------------------------------------------------- */
ORT_HD int ort_l3_enter(const DevScene& S, bool iris_before, OrtRay& r) { /* :551-580 */
    double t;
    if (iris_before) {
        t = ort_div_z(S.l3_iris1_z - r.pz, r.dz);
        double x = fma(r.dx, t, r.px), y = fma(r.dy, t, r.py);
        if (fma(x, x, y * y) > S.l3_iris_r2) { /* the reference leaves pos on the iris plane */
            r.px = x; r.py = y; r.pz = fma(r.dz, t, r.pz);
            return ORT_ST_L3_IRIS_BEFORE;
        }
    }
    if (!ort_hit_sphere(r, S.l3_c1x, S.l3_c1y, S.l3_c1z, S.l3_R1_2, &t)) return ORT_ST_L3_S1_MISS;
    ort_advance(r, t);
    if (fma(r.px, r.px, r.py * r.py) > S.l3_radius2) return ORT_ST_L3_APERTURE;
    return 0;
}
ORT_HD int ort_l3_body(const DevScene& S, const OrtRng& g, bool iris_after, OrtRay& r) { /* :582-644 */
    double u1, u2, u3, unused, t, nx, ny, nz;
    ort_draw2(g, 3, &u1, &u2);
    ort_sphere_normal(r, S.l3_c1x, S.l3_c1y, S.l3_c1z, S.l3_invR1, &nx, &ny, &nz);
    if (ort_interface(r, nx, ny, nz, S.l3_s1, u1)) return ORT_ST_L3_S1_REFLECT;
    if (!ort_hit_sphere(r, S.l3_c2x, S.l3_c2y, S.l3_c2z, S.l3_R2_2, &t)) return ORT_ST_L3_S2_MISS;
    ort_advance(r, t);
    ort_sphere_normal(r, S.l3_c2x, S.l3_c2y, S.l3_c2z, S.l3_invR2, &nx, &ny, &nz);
    if (ort_interface(r, nx, ny, nz, S.l3_s2, u2)) return ORT_ST_L3_S2_REFLECT;
    /* the reference aborts here on a miss (error stop "Help3", :617); we count it */
    if (!ort_hit_sphere(r, S.l3_c3x, S.l3_c3y, S.l3_c3z, S.l3_R3_2, &t)) return ORT_ST_L3_S3_MISS;
    ort_advance(r, t);
    ort_draw2(g, 4, &u3, &unused);
    ort_sphere_normal(r, S.l3_c3x, S.l3_c3y, S.l3_c3z, S.l3_invR3, &nx, &ny, &nz);
    if (ort_interface(r, nx, ny, nz, S.l3_s3, u3)) return ORT_ST_L3_S3_REFLECT;
    if (iris_after) {
        t = ort_div_z(S.l3_iris2_z - r.pz, r.dz);
        double x = fma(r.dx, t, r.px), y = fma(r.dy, t, r.py);
        if (fma(x, x, y * y) > S.l3_iris_r2) {
            r.px = x; r.py = y; r.pz = fma(r.dz, t, r.pz);
            return ORT_ST_L3_IRIS_AFTER;
        }
    }
    return 0;
}

/* -------------------------------------------------------------------------------------------
 * transfer to the image plane (src/optics_system.f90:48-49) + makeImage2D
 * (src/imageMod.f90:19-58).  Returns the status; *bin = (yp+200)*401 + (xp+200) when binned.
 * ----------------------------------------------------------------------------------------- */
ORT_HD int ort_image(const DevScene& S, OrtRay& r, int* xp, int* yp) {
    double d = ort_div_z(S.img_z - r.pz, r.dz);
    ort_advance(r, d);
    /* angle = acos(dz/|d|) > asin(0.22)  <=>  dz < cos_na |d|;  a NaN angle passes (reference) */
    double dd = fma(r.dx, r.dx, fma(r.dy, r.dy, r.dz * r.dz));
    if (r.dz <= 0.0 || r.dz * r.dz < S.cos_na2 * dd) return ORT_ST_NA_REJECT;
    if (r.px > 1000.0 || r.py > 1000.0) return ORT_ST_FAR;
    double fx = floor(r.px * S.inv_binwid), fy = floor(r.py * S.inv_binwid);
    if (!(fabs(fx) < 2.0e9) || !(fabs(fy) < 2.0e9)) return ORT_ST_FAR;
    if (fabs(fx) > 200.0 || fabs(fy) > 200.0) return ORT_ST_OFF_DETECTOR;
    *xp = (int)fx;
    *yp = (int)fy;
    return ORT_ST_BINNED;
}

/* -------------------------------------------------------------------------------------------
 * One whole iteration of the reference's ray loops for a single ray (src/main.f90:90-109 /
 * :127-162 incl. telescope, src/optics_system.f90:6-52), with the explicit-ray conveniences of
 * ort_trace_rays: optional caller-supplied start state and ort_job.stop_after.
 * ----------------------------------------------------------------------------------------- */
ORT_HD int ort_full_path(const DevScene& S, const DevJob& J, const OrtRng& g, bool have_input, OrtRay& r,
                         int* xp, int* yp) {
    const int stop = J.stop_after;
    int st;
    if (!have_input) {
        long long ray = ((long long)g.r1 << 32) | g.r0;
        int es;
        if (J.phase == ORT_PHASE_RING) {
            es = J.source_kind == ORT_SRC_CRS ? ort_emit<ORT_PHASE_RING, ORT_SRC_CRS>(S, J, g, ray, r)
               : J.source_kind == ORT_SRC_ISORS ? ort_emit<ORT_PHASE_RING, ORT_SRC_ISORS>(S, J, g, ray, r)
                                                : ort_emit<ORT_PHASE_RING, ORT_SRC_POINT>(S, J, g, ray, r);
        } else {
            es = J.source_kind == ORT_SRC_SPOT ? ort_emit<ORT_PHASE_POINT, ORT_SRC_SPOT>(S, J, g, ray, r)
                                               : ort_emit<ORT_PHASE_POINT, ORT_SRC_POINT>(S, J, g, ray, r);
        }
        if (es) return es;
    }
    if (stop == ORT_STOP_SOURCE) return ORT_ST_STOPPED;
    if (J.phase == ORT_PHASE_POINT && J.use_bottle) {
        st = (S.scatter_b | S.scatter_c) ? ort_bottle_forward<true>(S, g, r) : ort_bottle_forward<false>(S, g, r);
        if (st) return st;
    }
    if (stop == ORT_STOP_BOTTLE) return ORT_ST_STOPPED;
    st = ort_l2_enter(S, r);
    if (st) return st;
    st = ort_l2_body(S, g, r);
    if (st) return st;
    if (stop == ORT_STOP_L2) return ORT_ST_STOPPED;
    st = ort_l3_enter(S, J.iris_before != 0, r);
    if (st) return st;
    st = ort_l3_body(S, g, J.iris_after != 0, r);
    if (st) return st;
    if (stop == ORT_STOP_L3) return ORT_ST_STOPPED;
    return ort_image(S, r, xp, yp);
}

#endif /* ORT_OPTICS_CUH */
