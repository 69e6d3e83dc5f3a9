/*
 * ort_optics.cuh -- per-ray optics of the trace loop, one thread per ray, templated on the real
 * type R (double = the reference's arithmetic, 1e-9 parity; float = the fp32 variant, 1e-5).
 * Each function names the reference routine whose result it reproduces; DESIGN.md section 3.3
 * explains why the two need not be bit-identical.
 *
 * This is not a transliteration of the Fortran.  Measured on this chip (tools/microbench): a
 * DFMA with three distinct register operands issues every 3 cycles per scheduler, fp64 div and
 * sqrt are 6-9 dependent FP64 instructions, and two thirds of what the loops execute is NOT
 * floating point -- so the code below minimises instructions of every kind:
 *   - quadratics in half-b form; the root the reference picks by sorting follows from the signs
 *     of h and c, so only ONE quotient is formed (c/q or q/a);
 *   - Fresnel + Snell fused: one sqrt, amplitudes in units of n_b, the draw compared as
 *     2u B^2 D^2 > A^2 D^2 + C^2 B^2 (no division), eta c1 - c2 shared with the refraction;
 *   - sphere normals are (centre - pos) * (1/R), clear-cylinder normals * (1/R), 1/R hoisted;
 *   - aperture / iris tests on squared radii; the fibre-NA test as dz^2 >= cos^2(asin .22) |d|^2;
 *   - rcp / div / sqrt from the MUFU seed + one refinement + residual correction, without the
 *     compiler's range checks and slow-path calls (0 ulp off the correctly rounded operators over
 *     1.3e8 device samples, ort_math_selftest); sincospi restricted to [0, 2);
 *   - sign / zero tests on the integer pipe where a signed zero cannot occur;
 *   - every launch-invariant scalar comes pre-computed in DevScene (kernel-parameter constant
 *     bank), the Philox key schedule in DevJob.
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
/* sin(pi x), cos(pi x) for x in [0, 2) -- all the sources ever ask for (x = 2u).  The library
 * sincospi spends ~75 instructions on arbitrary arguments; here: q = nearest half-integer count,
 * r = x - q/2 in [-1/4, 1/4] (exact), Taylor polynomials of sin(pi r), cos(pi r) in r^2
 * (truncation < 5e-17), quadrant fix-up.  <= 2 ulp, like the library.  The coefficients sit in
 * constant memory so that each DFMA reads its own as a c[][] operand (as literals the compiler
 * builds every one of them in a pair of uniform registers first: 38 UMOV per call). */
#ifdef __CUDACC__
static __constant__ double ort_sincos_poly[19] = {
    7.95205400147551261e-07, -2.19153534478302173e-05, 4.66302805767612554e-04, -7.37043094571435044e-03,
    8.21458866111282326e-02, -5.99264529320792105e-01, 2.55016403987734552e+00, -5.16771278004997026e+00,
    3.14159265358979312e+00,
    -1.38789524622137714e-07, 4.30306958703294729e-06, -1.04638104924845705e-04, 1.92957430940392314e-03,
    -2.58068913900140612e-02, 2.35330630358893206e-01, -1.33526276885458950e+00, 4.05871212641676848e+00,
    -4.93480220054467900e+00, 1.0};
#endif
ORT_HD void ort_sincospi(double x, double* s, double* c) {
#ifdef __CUDA_ARCH__
    const double q = rint(x + x);
    const double r = fma(q, -0.5, x);
    const double t = r * r;
    const double* P = ort_sincos_poly;
    double ps = P[0];
#pragma unroll
    for (int i = 1; i < 9; ++i) ps = fma(ps, t, P[i]);
    ps *= r;
    double pc = P[9];
#pragma unroll
    for (int i = 10; i < 19; ++i) pc = fma(pc, t, P[i]);
    const int k = (int)q; /* 0..4 */
    const double a = (k & 1) ? pc : ps, b = (k & 1) ? ps : pc; /* odd quadrant: swap */
    *s = (k & 2) ? -a : a;
    *c = ((k + 1) & 2) ? -b : b;
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
ORT_HD double ort_sqrt_nz(double x) { /* x > 0: no zero guard (0 would give NaN) */
#ifdef __CUDA_ARCH__
    double y = ort_mufu_rsq(x);
    double g = x * y, h = 0.5 * y;
#pragma unroll
    for (int i = 0; i < ORT_NR - 1; ++i) {
        double r = fma(-g, h, 0.5);
        g = fma(g, r, g);
        h = fma(h, r, h);
    }
    double d = fma(-g, g, x);
    return fma(d, h, g);
#else
    return sqrt(x);
#endif
}
ORT_HD double ort_sqrt(double x) {
#ifdef __CUDA_ARCH__
    /* one Goldschmidt step, then the Newton correction of the root (also quadratic).  The seed is
     * taken of x + 1e-300 (== x for every x this path can produce except 0), so that x = 0 gives
     * g = 0 * 1e150 = 0 straight through instead of 0 * inf and needs no select; x < 0 stays NaN. */
    double y = ort_mufu_rsq(x + 1e-300);
    double g = x * y, h = 0.5 * y;
#pragma unroll
    for (int i = 0; i < ORT_NR - 1; ++i) {
        double r = fma(-g, h, 0.5);
        g = fma(g, r, g);
        h = fma(h, r, h);
    }
    double d = fma(-g, g, x);
    return fma(d, h, g);
#else
    return sqrt(x);
#endif
}

/* -------------------------------------------------------------------------------------------
 * fp32 variant (ort_job.precision = 32, tolerance 1e-5): the same functions on float.  MUFU.RCP
 * and MUFU.RSQ are ~1 ulp in fp32, so no refinement is needed.
 * ----------------------------------------------------------------------------------------- */
ORT_HD void ort_sincospi(float x, float* s, float* c) {
#ifdef __CUDA_ARCH__
    sincospif(x, s, c);
#else
    *s = sinf((float)ORT_PI * x);
    *c = cosf((float)ORT_PI * x);
#endif
}
ORT_HD void ort_sincos(float x, float* s, float* c) {
#ifdef __CUDA_ARCH__
    sincosf(x, s, c);
#else
    *s = sinf(x);
    *c = cosf(x);
#endif
}
/* MUFU seed (~1 ulp) + one Newton step: FP32 FMAs are nearly free next to everything else and
 * keep the variant inside its 1e-5 budget through seven surfaces */
ORT_HD float ort_rcp(float x) {
#ifdef __CUDA_ARCH__
    float y = __frcp_rn(x);
    return y;
#else
    return 1.0f / x;
#endif
}
ORT_HD float ort_div(float a, float b) {
#ifdef __CUDA_ARCH__
    float y = __fdividef(1.0f, b);
    float q = a * y;
    return fmaf(fmaf(-b, q, a), y, q);
#else
    return a / b;
#endif
}
ORT_HD float ort_div_z(float a, float b) { return a / b; } /* IEEE: +-inf for b == 0 */
ORT_HD float ort_rsqrt(float x) {
#ifdef __CUDA_ARCH__
    float y = rsqrtf(x);
    return fmaf(y * fmaf(-x * y, y, 1.0f), 0.5f, y);
#else
    return 1.0f / sqrtf(x);
#endif
}
ORT_HD float ort_sqrt_nz(float x) { /* x > 0 */
#ifdef __CUDA_ARCH__
    float y = rsqrtf(x), g = x * y;
    return fmaf(fmaf(-g, g, x), 0.5f * y, g);
#else
    return sqrtf(x);
#endif
}
ORT_HD float ort_sqrt(float x) {
#ifdef __CUDA_ARCH__
    float y = rsqrtf(x), g = x * y;
    g = fmaf(fmaf(-g, g, x), 0.5f * y, g);
    return (x == 0.0f) ? 0.0f : g;
#else
    return sqrtf(x);
#endif
}

/* sign-bit tests on the high word: integer ALU work instead of a DSETP on the FP64 pipe.  Only
 * used where a signed zero cannot occur (differences of distinct quantities). */
ORT_HD bool ort_signbit(double x) {
#ifdef __CUDA_ARCH__
    return __double2hiint(x) < 0;
#else
    return signbit(x);
#endif
}
ORT_HD bool ort_signbit(float x) { return signbit(x); }
ORT_HD bool ort_either_negative(double x, double y) {
#ifdef __CUDA_ARCH__
    return (__double2hiint(x) | __double2hiint(y)) < 0;
#else
    return signbit(x) || signbit(y);
#endif
}
ORT_HD bool ort_either_negative(float x, float y) { return signbit(x) || signbit(y); }
ORT_HD bool ort_is_zero(double x) { /* +0 only */
#ifdef __CUDA_ARCH__
    return __double_as_longlong(x) == 0ll;
#else
    return x == 0.0 && !signbit(x);
#endif
}
ORT_HD bool ort_is_zero(float x) { return x == 0.0f && !signbit(x); }
ORT_HD bool ort_both_zero(double x, double y) { /* +-0 */
#ifdef __CUDA_ARCH__
    return ((__double_as_longlong(x) | __double_as_longlong(y)) << 1) == 0ll;
#else
    return x == 0.0 && y == 0.0;
#endif
}
ORT_HD bool ort_both_zero(float x, float y) { return x == 0.0f && y == 0.0f; }

/* x with its sign flipped when `flip`: one integer op on the high word instead of a negation and
 * a two-register select */
ORT_HD double ort_flip_if(double x, bool flip) {
#ifdef __CUDA_ARCH__
    return __hiloint2double(__double2hiint(x) ^ (flip ? (int)0x80000000 : 0), __double2loint(x));
#else
    return flip ? -x : x;
#endif
}
ORT_HD float ort_flip_if(float x, bool flip) { return flip ? -x : x; }
/* x if c carries a sign bit, else -x -- (c < 0) ? x : -x wherever c cannot be a negative zero */
ORT_HD double ort_neg_unless_negative(double x, double c) {
#ifdef __CUDA_ARCH__
    return __hiloint2double(__double2hiint(x) ^ (~__double2hiint(c) & (int)0x80000000), __double2loint(x));
#else
    return (c < 0.0) ? x : -x;
#endif
}
ORT_HD float ort_neg_unless_negative(float x, float c) { return (c < 0.0f) ? x : -x; }

/* uncontracted arithmetic for stokes (see there) */
ORT_HD double ort_mul_rn(double a, double b) {
#ifdef __CUDA_ARCH__
    return __dmul_rn(a, b);
#else
    return a * b;
#endif
}
ORT_HD double ort_add_rn(double a, double b) {
#ifdef __CUDA_ARCH__
    return __dadd_rn(a, b);
#else
    return a + b;
#endif
}
ORT_HD double ort_sub_rn(double a, double b) {
#ifdef __CUDA_ARCH__
    return __dsub_rn(a, b);
#else
    return a - b;
#endif
}
ORT_HD float ort_mul_rn(float a, float b) {
#ifdef __CUDA_ARCH__
    return __fmul_rn(a, b);
#else
    return a * b;
#endif
}
ORT_HD float ort_add_rn(float a, float b) {
#ifdef __CUDA_ARCH__
    return __fadd_rn(a, b);
#else
    return a + b;
#endif
}
ORT_HD float ort_sub_rn(float a, float b) {
#ifdef __CUDA_ARCH__
    return __fsub_rn(a, b);
#else
    return a - b;
#endif
}

/* -------------------------------------------------------------------------------------------
 * Counter-based uniforms -- replaces the reference's ran2() (src/random_mod.f90:39-46).
 * Philox4x32-7 (Salmon et al., SC'11: the 7-round variant is the fewest rounds that pass BigCrush,
 * "Crush-resistant"; 10 is Random123's default with its safety margin).  counter = (ray_lo, ray_hi,
 * phase, block), key = (seed_lo, seed_hi).  A block is four 32-bit words w0..w3 and serves
 *     one WIDE draw   u = ((w1:w0) >> 11) * 2^-53   53 bits like gfortran's random_number: radial
 *                     draws (annulus / aim-disc radius^2, cos theta), whose small values matter;
 *     NARROW draws    u = w * 2^-32                 angles and reflect-or-refract decisions.
 * Slot map, rev 2 (fixed, independent of control flow; `ort_uniforms` exposes it):
 *     slot : block.words      ring loop                  point loop
 *        0 : 0.w0w1 wide      annulus r^2                cos theta
 *        1 : 0.w2             annulus angle              phi
 *        2 : 1.w0w1 wide      aim-disc r^2 (*)           bottle inner wall decision
 *        3 : 1.w2             aim-disc angle             bottle outer wall decision
 *        4 : 0.w3             L2 flat decision           L2 flat decision
 *        5 : 1.w3             L2 curved decision         L2 curved decision
 *    6,7,8 : 2.w0 w1 w2       L3 surfaces 1, 2, 3        L3 surfaces 1, 2, 3       (9: 2.w3 spare)
 *       10 : 3.w0w1 wide, 11 : 3.w2, 12 : 3.w3          image source: aim r^2, aim angle
 *       13 : 4.w0w1 wide, 14 : 4.w2, 15 : 4.w3          spare
 *     16.. : block slot/2, both halves wide              scatter loops / rang, consumed in order
 * (*) in the ring loop the HIGH word of slot 2 is not 1.w1 but word (ray & 3) of a block that four
 *     consecutive rays share -- counter (ray >> 2, 0, 16 + phase, 0): that word alone decides L2's
 *     aperture for 69 % of the ring rays, and this way one Philox block decides it for four of them
 *     (words of one block are as independent as words of different blocks); the low word stays 1.w0.
 * A whole ray needs three blocks (rev 1: five blocks of ten rounds); the ring loop's culling kernel
 * needs block 1 for every ray and block 0 for the rays that pass L2's aperture.
 * ----------------------------------------------------------------------------------------- */
#define ORT_PHILOX_ROUNDS 7
struct OrtRng {
    uint32_t k0, k1;   /* seed */
    const uint32_t* rk; /* optional precomputed key schedule (DevJob.round_keys), or NULL */
    uint32_t r0, r1;   /* ray index */
    uint32_t phase;
    double override_u; /* >= 0: every draw returns this */
};

/* rounds [first, first + n) of Philox4x32; the key of round r is key + r * (W0, W1) */
ORT_HD void ort_philox4x32_rounds(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                  uint32_t k1, uint32_t* o, const uint32_t* rk, int n) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < n; ++r) {
        uint32_t hi0 = ort_mulhi(M0, c0), lo0 = M0 * c0;
        uint32_t hi1 = ort_mulhi(M1, c2), lo1 = M1 * c2;
        uint32_t ka = rk ? rk[2 * r] : k0, kb = rk ? rk[2 * r + 1] : k1;
        uint32_t n0 = hi1 ^ c1 ^ ka, n2 = hi0 ^ c3 ^ kb;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += W0; k1 += W1;
    }
    o[0] = c0; o[1] = c1; o[2] = c2; o[3] = c3;
}
ORT_HD void ort_philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                           uint32_t k1, uint32_t* o, const uint32_t* rk = nullptr) {
    ort_philox4x32_rounds(c0, c1, c2, c3, k0, k1, o, rk, ORT_PHILOX_ROUNDS);
}
/* the four words of block `block` of this ray */
ORT_HD void ort_block(const OrtRng& g, uint32_t block, uint32_t* w) {
    ort_philox4x32(g.r0, g.r1, g.phase, block, g.k0, g.k1, w, g.rk);
}
/* the block the four rays 4q .. 4q+3 share (ring loop: the high words of their slot-2 draws) */
#define ORT_SHARED_PHASE 16u
ORT_HD void ort_shared_block(const OrtRng& g, uint32_t* w) {
    const uint32_t q0 = (g.r0 >> 2) | (g.r1 << 30), q1 = g.r1 >> 2;
    ort_philox4x32(q0, q1, ORT_SHARED_PHASE + g.phase, 0u, g.k0, g.k1, w, g.rk);
}
ORT_HD uint32_t ort_pick_word(const uint32_t* w, uint32_t k) {
    return k == 0u ? w[0] : k == 1u ? w[1] : k == 2u ? w[2] : w[3];
}
/* the high word of this ray's slot-2 draw in the ring loop */
ORT_HD uint32_t ort_aim_hi(const OrtRng& g) {
    uint32_t w[4];
    ort_shared_block(g, w);
    return ort_pick_word(w, g.r0 & 3u);
}

/* 64 random bits -> uniform in [0,1): 53 bits for double (like gfortran's random_number); the
 * fp32 variant rounds the SAME number to float, so both variants make the same decisions except
 * within 2^-24 of a threshold */
template <typename R> ORT_HD R ort_bits_to_uniform(uint32_t lo, uint32_t hi);
template <> ORT_HD double ort_bits_to_uniform<double>(uint32_t lo, uint32_t hi) {
    uint64_t bits = (((uint64_t)hi << 32) | lo) >> 11;
    return (double)bits * (1.0 / 9007199254740992.0);
}
template <> ORT_HD float ort_bits_to_uniform<float>(uint32_t lo, uint32_t hi) {
    /* full relative precision for small draws (an annulus radius is sqrt(u)); never 1.0f */
    uint64_t bits = (((uint64_t)hi << 32) | lo) >> 11;
    float u = (float)bits * (1.0f / 9007199254740992.0f);
    return fminf(u, 0.99999994f);
}
/* one word -> uniform in [0,1): exact in double */
template <typename R> ORT_HD R ort_word_to_uniform(uint32_t w);
template <> ORT_HD double ort_word_to_uniform<double>(uint32_t w) { return (double)w * (1.0 / 4294967296.0); }
template <> ORT_HD float ort_word_to_uniform<float>(uint32_t w) {
    return fminf((float)w * (1.0f / 4294967296.0f), 0.99999994f);
}
/* the same with the known-answer-test override (ort_trace_rays only; folds away in the loops) */
template <typename R>
ORT_HD R ort_wide(const OrtRng& g, uint32_t lo, uint32_t hi) {
    return g.override_u >= 0.0 ? (R)g.override_u : ort_bits_to_uniform<R>(lo, hi);
}
template <typename R>
ORT_HD R ort_narrow(const OrtRng& g, uint32_t w) {
    return g.override_u >= 0.0 ? (R)g.override_u : ort_word_to_uniform<R>(w);
}

/* twice the draw, for ort_interface: a scaling by 2^-31 (2^-52) instead of 2^-32 (2^-53), exact */
template <typename R>
ORT_HD R ort_narrow2(const OrtRng& g, uint32_t w) {
    if (g.override_u >= 0.0) return (R)(g.override_u + g.override_u);
    return sizeof(R) == 8 ? (R)((double)w * (1.0 / 2147483648.0)) : ort_word_to_uniform<R>(w) * R(2.0);
}
template <typename R>
ORT_HD R ort_wide2(const OrtRng& g, uint32_t lo, uint32_t hi) {
    const R u = ort_wide<R>(g, lo, hi);
    return u + u;
}

/* draw slot -> (block, which words): the table above */
template <typename R>
ORT_HD R ort_slot(const OrtRng& g, uint32_t slot) {
    if (g.override_u >= 0.0) return (R)g.override_u;
    uint32_t w[4];
    if (slot >= 16u) {
        ort_block(g, slot >> 1, w);
        return (slot & 1u) ? ort_bits_to_uniform<R>(w[2], w[3]) : ort_bits_to_uniform<R>(w[0], w[1]);
    }
    /* kind: 0 wide, 1 = w2, 2 = w3, 3..6 = w0..w3 of block 2 */
    const uint32_t blk = slot < 6u ? ((slot == 4u) ? 0u : (slot == 5u) ? 1u : (slot >> 1)) : slot < 10u ? 2u : slot < 13u ? 3u : 4u;
    ort_block(g, blk, w);
    if (slot == 2u && g.phase == (uint32_t)ORT_PHASE_RING) return ort_bits_to_uniform<R>(w[0], ort_aim_hi(g));
    if (slot < 6u) return slot < 4u ? ((slot & 1u) ? ort_word_to_uniform<R>(w[2]) : ort_bits_to_uniform<R>(w[0], w[1]))
                                    : ort_word_to_uniform<R>(w[3]);
    if (slot < 10u) return ort_word_to_uniform<R>(w[slot - 6u]);
    const uint32_t k = (slot - 10u) % 3u;
    return k == 0u ? ort_bits_to_uniform<R>(w[0], w[1]) : ort_word_to_uniform<R>(w[1u + k]);
}

/* the two wide uniforms of Philox block `block` (scatter draws: slots 2*block and 2*block+1) */
template <typename R>
ORT_HD void ort_draw2(const OrtRng& g, uint32_t block, R* ua, R* ub) {
    if (g.override_u >= 0.0) {
        *ua = (R)g.override_u;
        *ub = (R)g.override_u;
        return;
    }
    uint32_t w[4];
    ort_block(g, block, w);
    *ua = ort_bits_to_uniform<R>(w[0], w[1]);
    *ub = ort_bits_to_uniform<R>(w[2], w[3]);
}

/* sequential draws for the scatter loops (slots 16, 17, ...) */
template <typename R>
struct OrtScatterRngT {
    uint32_t next; /* next slot */
    R spare;       /* odd-slot value of the last generated block */
};
template <typename R>
ORT_HD R ort_scatter_draw(const OrtRng& g, OrtScatterRngT<R>& s) {
    if (g.override_u >= 0.0) return (R)g.override_u;
    uint32_t slot = s.next++;
    if (slot & 1u) return s.spare;
    R a, b;
    ort_draw2(g, slot >> 1, &a, &b);
    s.spare = b;
    return a;
}

template <typename R>
struct OrtRayT {
    R px, py, pz, dx, dy, dz;
};
typedef OrtRayT<double> OrtRay;

/* -------------------------------------------------------------------------------------------
 * Quadratic root selection -- reproduces solveQuadratic + the root picking shared by the
 * reference's intersect_* (src/surfaces.f90:227-260 and :74-87).  Half-b form:
 * a t^2 + 2 h t + c = 0 (a = 1 for the spheres: unit directions).
 * ----------------------------------------------------------------------------------------- */
/* Unit normal of a sphere at a point on it: (centre - pos) / R with 1/R hoisted.  A ray that runs
 * exactly along the axis must see EXACTLY (0,0,+-1): the reference normalises by the computed
 * length, gets |N.I| == 1 bit for bit, and its fresnel() then returns 0 -- the
 * on-axis rays of create_spot and of the known-answer tests depend on it. */
template <typename R>
ORT_HD void ort_sphere_normal(const OrtRayT<R>& r, R cx, R cy, R cz, R invR, R* nx,
                              R* ny, R* nz) {
    R vx = cx - r.px, vy = cy - r.py, vz = cz - r.pz;
    *nx = vx * invR;
    *ny = vy * invR;
    *nz = ort_both_zero(vx, vy) ? copysign(R(1.0), vz) : vz * invR;
}

/* The reference sorts the two roots, takes the smaller unless it is negative, and misses when
 * both are (src/surfaces.f90:75-84).  With the stable pair q = -(h + sgn(h) s), {q/a, c/q} the
 * outcome is decided by the signs of h and c alone, and only ONE quotient is ever needed:
 *   h > 0 : q < 0, so q/a < 0 and c/q has the sign of -c:  c > 0 -> miss, else t = c/q
 *   h <= 0: q >= 0:  c < 0 -> c/q < 0, t = q/a ;  c >= 0 -> both >= 0 and c/q is the smaller */
template <typename R>
ORT_HD bool ort_pick_root_unit(R h, R c, R* t) {
    /* a == 1 (unit direction): q/a = q */
    R disc = fma(h, h, -c);
    if (ort_signbit(disc)) return false;
    R s = ort_sqrt(disc);
    bool hpos = h > R(0.0);
    if (hpos && c > R(0.0)) return false;
    R q = -(h + ort_flip_if(s, !hpos)); /* hpos ? -(h + s) : (s - h) */
    R x1 = ort_div(c, q);
    R tt = (!hpos && c < R(0.0)) ? q : x1;
    if (q == R(0.0)) tt = R(0.0); /* h = c = 0: the ray starts on the surface, tangent */
    *t = tt;
    return true;
}
template <typename R>
ORT_HD bool ort_pick_root(R a, R h, R c, R* t) {
    R disc = fma(h, h, -a * c);
    if (ort_signbit(disc)) return false;
    R s = ort_sqrt(disc);
    bool hpos = h > R(0.0);
    if (hpos && c > R(0.0)) return false;
    R q = hpos ? -(h + s) : (s - h);
    bool use_q = !hpos && c < R(0.0);
    R tt = ort_div(use_q ? q : c, use_q ? a : q);
    if (!(tt >= R(0.0))) return false; /* a = 0 (ray along the axis) or q = 0 */
    *t = tt;
    return true;
}

/* intersect_sphere (src/surfaces.f90:52-89) for a unit direction */
template <typename R>
ORT_HD bool ort_hit_sphere(const OrtRayT<R>& r, R cx, R cy, R cz, R R2, R* t) {
    R lx = r.px - cx, ly = r.py - cy, lz = r.pz - cz;
    R h = fma(r.dx, lx, fma(r.dy, ly, r.dz * lz));
    R c = fma(lx, lx, fma(ly, ly, fma(lz, lz, -R2)));
    return ort_pick_root_unit(h, c, t);
}
/* intersect_cylinder (src/surfaces.f90:91-130): axis along x, only (y,z) enter */
template <typename R>
ORT_HD bool ort_hit_cylinder(const OrtRayT<R>& r, R cy, R cz, R R2, R* t) {
    R ly = r.py - cy, lz = r.pz - cz;
    R a = fma(r.dz, r.dz, r.dy * r.dy);
    R h = fma(r.dz, lz, r.dy * ly);
    R c = fma(lz, lz, fma(ly, ly, -R2));
    return ort_pick_root(a, h, c, t);
}
/* intersect_ellipse (src/surfaces.f90:133-176): ia2 = 1/semia^2 (z), ib2 = 1/semib^2 (y) */
template <typename R>
ORT_HD bool ort_hit_ellipse(const OrtRayT<R>& r, R cy, R cz, R ia2, R ib2, R* t) {
    R ly = r.py - cy, lz = r.pz - cz;
    R a = fma(ia2 * r.dz, r.dz, ib2 * r.dy * r.dy);
    R h = fma(ia2 * r.dz, lz, ib2 * r.dy * ly);
    R c = fma(ia2 * lz, lz, fma(ib2 * ly, ly, -R(1.0)));
    return ort_pick_root(a, h, c, t);
}

template <typename R>
ORT_HD void ort_advance(OrtRayT<R>& r, R t) {
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
/* u2 = TWICE the draw (ort_narrow2 / ort_wide2 form it at no extra cost) */
template <typename R>
ORT_HD bool ort_interface(OrtRayT<R>& r, R nx, R ny, R nz, const DevIfaceT<R>& f, R u2) {
    R c = fma(nx, r.dx, fma(ny, r.dy, nz * r.dz)); /* N . I */
    R costt = fabs(c);
    R s2 = fma(-costt, costt, R(1.0)); /* sin^2(theta_i) */
    R ct2 = fma(-f.eta2, s2, R(1.0));  /* cos^2(theta_t) = 1 - eta^2 sin^2 */
    R cost2 = ort_sqrt_nz(ct2); /* NaN under total internal reflection, where it is not used */
    /* Fresnel amplitudes in units of nb (the ratios do not change): A/B = r_s, C/D = r_p.
     * R = (A^2 D^2 + C^2 B^2) / (2 B^2 D^2); the draw is compared without forming the quotient:
     *   u > R  <=>  2u B^2 D^2 > A^2 D^2 + C^2 B^2   (2u is what the caller hands in).
     * 0 <= R <= 1 by construction (|A| <= B, |C| <= D); a NaN fails the comparison and reflects,
     * which is what the reference's NaN guard (R = 1) does. */
    R ec = f.eta * costt, e2 = f.eta * cost2;
    R A = ec - cost2, B = ec + cost2, C = e2 - costt, D = e2 + costt;
    R B2 = B * B, D2 = D * D;
    R num = fma(A * A, D2, (C * C) * B2);
    R lhs = u2 * (B2 * D2);
    bool transmit = lhs > num;
    if (ort_either_negative(ct2, s2)) transmit = false; /* TIR, or |N.I| > 1 by rounding (reference: NaN -> R = 1) */
    else if (ort_is_zero(s2)) transmit = u2 > R(0.0);     /* exactly normal incidence: the reference returns R = 0 */
    if (!transmit) { /* reflect, src/surfaces.f90:285-300 */
        R k = -R(2.0) * c;
        r.dx = fma(k, nx, r.dx);
        r.dy = fma(k, ny, r.dy);
        r.dz = fma(k, nz, r.dz);
        return true;
    }
    /* refract, src/surfaces.f90:303-333: T = eta I + (eta c1 - c2) N', N' opposing I */
    /* k = (c < 0) ? A : -A, on the sign bit of c: N.I is a sum of three products and cannot be a negative
     * zero unless every product is one (a ray in the tangent plane with two vanishing components) */
    R k = ort_neg_unless_negative(A, c); /* eta c1 - c2, opposing the normal */
    r.dx = fma(f.eta, r.dx, k * nx);
    r.dy = fma(f.eta, r.dy, k * ny);
    r.dz = fma(f.eta, r.dz, k * nz);
    return false;
}

/* -------------------------------------------------------------------------------------------
 * Sources (src/sourceMod.f90)
 * ----------------------------------------------------------------------------------------- */
/* point, src/sourceMod.f90:12-47 */
/* blocks 0 and 1 of a ray: every emitter and the two surfaces after it draw from these, and the
 * two L2 decisions (a[3], b[3]) travel with the ray to L2 */
struct OrtDraws01 {
    uint32_t a[4], b[4];
    uint32_t hi2; /* high word of slot 2: b[1], in the ring loop the word of the shared block */
};
/* PHASE: ORT_PHASE_RING / ORT_PHASE_POINT where the caller knows it at compile time (the kernels), 0 =
 * look at g.phase */
template <int PHASE = 0>
ORT_HD void ort_draws01(const OrtRng& g, OrtDraws01& D) {
    ort_block(g, 0u, D.a);
    ort_block(g, 1u, D.b);
    const bool ring = PHASE ? PHASE == ORT_PHASE_RING : g.phase == (uint32_t)ORT_PHASE_RING;
    D.hi2 = ring ? ort_aim_hi(g) : D.b[1];
}

template <typename R>
ORT_HD void ort_source_point(const DevSceneT<R>& S, const OrtRng& g, const OrtDraws01& D, OrtRayT<R>& r) {
    R sp, cp;
    const R u1 = ort_wide<R>(g, D.a[0], D.a[1]);  /* slot 0: cos theta */
    const R u0 = ort_narrow<R>(g, D.a[2]);        /* slot 1: phi */
    ort_sincospi(R(2.0) * u0, &sp, &cp);
    R cost = fma(u1, S.cos_theta_max, R(1.0) - u1);
    R sint;
    if (sizeof(R) == 4) { /* fp32: 1 - cost^2 cancels for small cones; (1-cost)(1+cost) does not */
        sint = ort_sqrt((u1 * S.one_m_ctm) * (R(1.0) + cost));
    } else {
        sint = ort_sqrt(fma(-cost, cost, R(1.0)));
    }
    r.dx = sint * cp;
    r.dy = sint * sp;
    r.dz = cost;
    r.px = R(0.0);
    r.py = R(0.0);
    r.pz = S.point_offset;
}

/* ring, src/sourceMod.f90:250-300; (u0,u1) place the ray on the annulus, (u2,u3) pick the aim
 * point on the disc of radius L2.radius + 10 mm in the plane z = L2.fb */
template <typename R>
ORT_HD void ort_source_ring_u(const DevSceneT<R>& S, R u0, R u1, R u2, R u3, OrtRayT<R>& r) {
    R s, c;
    R rr = ort_sqrt(fma(u0, S.r2_m_r1, S.r1));
    ort_sincospi(R(2.0) * u1, &s, &c);
    R px = rr * c, py = rr * s;
    R q = S.ellipse ? py * S.ra_over_rb : py;
    R pz = S.bcz + ort_sqrt(fma(-q, q, S.ra2));
    R rl = ort_sqrt(u2 * S.lens_r2);
    ort_sincospi(R(2.0) * u3, &s, &c);
    R ex = fma(rl, c, -px), ey = fma(rl, s, -py), ez = S.l2_fb - pz;
    R inv = ort_rsqrt(fma(ex, ex, fma(ey, ey, ez * ez)));
    r.px = px;
    r.py = py;
    r.pz = pz;
    r.dx = ex * inv;
    r.dy = ey * inv;
    r.dz = ez * inv;
}
template <typename R>
ORT_HD void ort_source_ring(const DevSceneT<R>& S, const OrtRng& g, const OrtDraws01& D, OrtRayT<R>& r) {
    ort_source_ring_u(S, ort_wide<R>(g, D.a[0], D.a[1]), ort_narrow<R>(g, D.a[2]), ort_wide<R>(g, D.b[0], D.hi2),
                      ort_narrow<R>(g, D.b[2]), r);
}
/* When L2's flat face lies in the aim plane (DevSceneT<R>.ring_shortcut) the ray meets that face AT
 * its aim point, so the aperture test of src/lens.f90:450-454 is a test on u2 alone: 69 % of the
 * ring rays of the shipped geometries end here, before any position, direction, sqrt or sincos
 * has been computed. */
template <typename R>
ORT_HD bool ort_ring_aims_outside_aperture(const DevSceneT<R>& S, R u2) {
    return u2 * S.lens_r2 > S.l2_radius2;
}

/* ---- the other sources of settings.params ---------------------------- */
/* rang, src/random_mod.f90:59-85: polar Box-Muller; its rejection loop takes the sequential
 * draws (slots 16, 17, ...) */
template <typename R>
ORT_HD void ort_rang(const OrtRng& g, OrtScatterRngT<R>& sr, R sigma, R* x, R* y) {
    R a, b, s;
    do {
        a = fma(ort_scatter_draw(g, sr), R(2.0), -R(1.0));
        b = fma(ort_scatter_draw(g, sr), R(2.0), -R(1.0));
        s = fma(b, b, a * a);
    } while (s >= R(1.0) && g.override_u < R(0.0));
    R cst = sqrt(-R(2.0) * log(s) / s);
    *x = sigma * (a * cst);
    *y = sigma * (b * cst);
}

/* point_on_bottle, src/sourceMod.f90:50-89 (crs, ring loop): a Gaussian spot projected along -z
 * onto the cylinder of radius Ra + thickness, emitting into the cone of point() */
template <typename R>
ORT_HD bool ort_source_crs(const DevSceneT<R>& S, const OrtRng& g, const OrtDraws01& D, OrtRayT<R>& r) {
    OrtScatterRngT<R> sr;
    sr.next = 16;
    sr.spare = R(0.0);
    ort_source_point(S, g, D, r); /* same two draws, same direction formulas (:65-77) */
    R dx = r.dx, dy = r.dy, dz = r.dz, x, y;
    ort_rang(g, sr, S.spot_size, &x, &y);
    /* the reference drops the point from z = 1 along -z onto the cylinder of radius Ra + thickness
     * (intersect_cylinder, first root): z = cz + sqrt(R^2 - (y - cy)^2), formed directly so that
     * the fp32 variant does not subtract two numbers close to 1 */
    R qy = y - S.bcy;
    R disc = fma(-qy, qy, S.crs_r2);
    if (disc < R(0.0)) return false;
    r.px = x;
    r.py = y;
    r.pz = S.bcz + ort_sqrt(disc);
    r.dx = dx;
    r.dy = dy;
    r.dz = dz;
    return true;
}

/* create_spot, src/sourceMod.f90:122-159 (spot, point loop): deterministic angular grid;
 * n = 1-based loop index, nrays = nphotons */
template <typename R>
ORT_HD void ort_source_spot(const DevSceneT<R>& S, long long nrays, long long n, OrtRayT<R>& r) {
    R nrays_sqrt = sqrt((R)nrays);
    R dphi = R(ORT_TWOPI) / nrays_sqrt;
    R dtheta = acos(S.cos_theta_max) / nrays_sqrt;
    R phi = dphi * (R)(n % 10), theta = dtheta * (R)(n / 10);
    R sp, cp, st_, ct;
    ort_sincos(phi, &sp, &cp);
    ort_sincos(theta, &st_, &ct);
    R sint = sqrt(fma(-ct, ct, R(1.0)));
    r.dx = sint * cp;
    r.dy = sint * sp;
    r.dz = ct;
    r.px = r.py = r.pz = R(0.0);
}

/* intersect_cone, src/surfaces.f90:179-224, for the axicon of iSORS.  The Gaussian beam is centred
 * on the apex, where the discriminant b^2 - 4ac = 4 k rho^2 is the difference of two numbers ~1e9
 * times larger: the reference's own answer is only good to ~1e-7 there, and any other operation
 * order gives a different one.  So here, as in stokes, the arithmetic keeps the reference's order
 * and is protected from FMA contraction (a once-per-ray source routine, not a hot spot). */
template <typename R>
ORT_HD bool ort_hit_cone(const OrtRayT<R>& r, R k, R height, R* t) {
    const R lz = ort_sub_rn(r.pz, height);
    const R a = ort_sub_rn(ort_add_rn(ort_mul_rn(r.dx, r.dx), ort_mul_rn(r.dy, r.dy)), ort_mul_rn(k, ort_mul_rn(r.dz, r.dz)));
    const R b = ort_mul_rn(R(2.0), ort_sub_rn(ort_add_rn(ort_mul_rn(r.dx, r.px), ort_mul_rn(r.dy, r.py)),
                                              ort_mul_rn(ort_mul_rn(k, r.dz), lz)));
    const R c = ort_sub_rn(ort_add_rn(ort_mul_rn(r.px, r.px), ort_mul_rn(r.py, r.py)), ort_mul_rn(k, ort_mul_rn(lz, lz)));
    /* solveQuadratic, src/surfaces.f90:227-260, and the root picking of :212-221 */
    const R disc = ort_sub_rn(ort_mul_rn(b, b), ort_mul_rn(ort_mul_rn(R(4.0), a), c));
    if (disc < R(0.0)) return false;
    R x0, x1;
    if (disc == R(0.0)) {
        x0 = x1 = ort_mul_rn(-R(0.5), b) / a;
    } else {
        const R s = sqrt(disc);
        const R q = ort_mul_rn(-R(0.5), (b > R(0.0)) ? ort_add_rn(b, s) : ort_sub_rn(b, s));
        x0 = q / a;
        x1 = c / q;
    }
    R t0 = fmin(x0, x1), t1 = fmax(x0, x1);
    R tt = (t0 < R(0.0)) ? t1 : t0;
    if (tt < R(0.0)) return false;
    *t = tt;
    return true;
}

/* iSORS(ring = .true.), src/sourceMod.f90:162-247 (isors, ring loop).  false = the reference's
 * `error stop "no intersection with bottle!"` (every ray the axicon face reflects, ~2.8 %). */
template <typename R>
ORT_HD bool ort_source_isors(const DevSceneT<R>& S, const OrtRng& g, const OrtDraws01& D, OrtRayT<R>& r) {
    OrtScatterRngT<R> sr;
    sr.next = 16;
    sr.spare = R(0.0);
    R x, y, t;
    ort_rang(g, sr, S.isors_beam, &x, &y);
    r.px = x; r.py = y; r.pz = R(2.0) * S.isors_h;
    r.dx = R(0.0); r.dy = R(0.0); r.dz = -R(1.0);
    const R u_r = ort_wide<R>(g, D.a[0], D.a[1]), u_th = ort_narrow<R>(g, D.a[2]); /* slots 0, 1 */
    const R u_ax = ort_wide2<R>(g, D.b[0], D.hi2);                                 /* slot 2, doubled */
    if (ort_hit_cone(r, S.isors_k, S.isors_h, &t)) {
        ort_advance(r, t);
        /* gradient of the cone, inverted (upper nappe), normalised */
        R nx = -(R(2.0) * r.px / S.isors_k), ny = -(R(2.0) * r.py / S.isors_k), nz = -(-R(2.0) * r.pz + R(2.0) * S.isors_h);
        R inv = R(1.0) / sqrt(fma(nx, nx, fma(ny, ny, nz * nz)));
        (void)ort_interface(r, nx * inv, ny * inv, nz * inv, S.isors_axicon, u_ax); /* flag ignored */
        ort_advance(r, S.isors_base / r.dz);
        r.pz = S.isors_z;
        bool hit = S.ellipse ? ort_hit_ellipse(r, S.bcy, S.bcz, S.b_in_ia2, S.b_in_ib2, &t)
                             : ort_hit_cylinder(r, S.bcy, S.bcz, S.b_in_r2, &t);
        if (!hit) return false;
        ort_advance(r, t);
    }
    R rl = sqrt(u_r * S.isors_lens_r2), s, c;
    ort_sincospi(R(2.0) * u_th, &s, &c);
    R ex = fma(rl, c, -r.px), ey = fma(rl, s, -r.py), ez = S.l2_fb - r.pz;
    R inv = R(1.0) / sqrt(fma(ex, ex, fma(ey, ey, ez * ez)));
    r.dx = ex * inv;
    r.dy = ey * inv;
    r.dz = ez * inv;
    return true;
}

/* emit_image + emit, src/sourceMod.f90:303-361 (image source, point loop): ray k leaves the pixel
 * the reference's budget scan reaches after k rays (binary search in the prefix sums), from a
 * uniform point inside it, aimed at a uniform point of L2's aperture disc.
 * Slots: x 0, y 1, aim radius^2 10, aim angle 11. */
template <typename R>
ORT_HD bool ort_source_image(const DevSceneT<R>& S, const DevJob& J, const OrtRng& g, const OrtDraws01& D, long long k,
                             OrtRayT<R>& r) {
    const long long npix = (long long)ORT_SRCIMG_N * ORT_SRCIMG_N;
    const long long* cdf = J.image_cdf;
    if (cdf == nullptr || k >= cdf[npix - 1]) return false;
    long long lo = 0, hi = npix - 1;
    while (lo < hi) {
        long long mid = (lo + hi) >> 1;
        if (cdf[mid] > k) hi = mid;
        else lo = mid + 1;
    }
    const R dx = R(5000e-6 / 512.);
    const R fj = (R)(lo % ORT_SRCIMG_N), fi = (R)(lo / ORT_SRCIMG_N); /* zero-based pixel */
    R s, c;
    uint32_t w3[4];
    ort_block(g, 3u, w3);
    const R u0 = ort_wide<R>(g, D.a[0], D.a[1]), u1 = ort_narrow<R>(g, D.a[2]);   /* slots 0, 1 */
    const R u2 = ort_wide<R>(g, w3[0], w3[1]), u3 = ort_narrow<R>(g, w3[2]);      /* slots 10, 11 */
    R ax = fj * dx, bx = (fj + R(1.0)) * dx, ay = fi * dx, by = (fi + R(1.0)) * dx;
    r.px = fma(u0, bx - ax, ax) - R(2500e-6);
    r.py = fma(u1, by - ay, ay) - R(2500e-6);
    r.pz = R(0.0);
    R rl = ort_sqrt(u2 * S.isors_lens_r2); /* L2.radius^2 */
    ort_sincospi(R(2.0) * u3, &s, &c);
    R ex = fma(rl, c, -r.px), ey = fma(rl, s, -r.py), ez = S.l2_fb - r.pz;
    R inv = ort_rsqrt(fma(ex, ex, fma(ey, ey, ez * ez)));
    r.dx = ex * inv;
    r.dy = ey * inv;
    r.dz = ez * inv;
    return true;
}

/* source dispatch of src/main.f90:95-101 (ring loop) and :132-142 (point loop); SRC is
 * ort_job.source_kind.  Returns 0 or ORT_ST_SOURCE_MISS. */
template <int PHASE, int SRC, typename R>
ORT_HD int ort_emit(const DevSceneT<R>& S, const DevJob& J, const OrtRng& g, const OrtDraws01& D, long long ray,
                    OrtRayT<R>& r) {
    if (PHASE == ORT_PHASE_RING) {
        if (SRC == ORT_SRC_CRS) return ort_source_crs(S, g, D, r) ? 0 : ORT_ST_SOURCE_MISS;
        if (SRC == ORT_SRC_ISORS) return ort_source_isors(S, g, D, r) ? 0 : ORT_ST_SOURCE_MISS;
        ort_source_ring(S, g, D, r);
    } else {
        if (SRC == ORT_SRC_IMAGE) return ort_source_image(S, J, g, D, ray, r) ? 0 : ORT_ST_SOURCE_MISS;
        if (SRC == ORT_SRC_SPOT) ort_source_spot(S, J.total_rays, ray + 1, r);
        else ort_source_point(S, g, D, r);
    }
    return 0;
}

/* -------------------------------------------------------------------------------------------
 * Scatter: tauint (src/surfaces.f90:13-50) and stokes (src/stokes.f90:7-166)
 * ----------------------------------------------------------------------------------------- */
/* returns false where the reference would `error stop "no intersection"` */
template <typename R>
ORT_HD bool ort_tauint(const OrtRayT<R>& r, R mutot, R inv_mutot, R cy, R cz,
                       R R2, R u, R* dist, bool* tflag) {
    R tau = -log(u);
    R d;
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
template <typename R>
ORT_HD void ort_stokes(OrtRayT<R>& r, R hgg, const OrtRng& g, OrtScatterRngT<R>& sr) {
    R cost = r.dz;
    R sint = sqrt(ort_sub_rn(R(1.0), ort_mul_rn(cost, cost)));
    R phi = atan2(r.dy, r.dx);
    R sinp, cosp;
    if (hgg == R(0.0)) { /* isotropic, src/stokes.f90:33-48 */
        cost = ort_sub_rn(ort_mul_rn(R(2.0), ort_scatter_draw(g, sr)), R(1.0));
        sint = ort_sub_rn(R(1.0), ort_mul_rn(cost, cost));
        sint = (sint <= R(0.0)) ? R(0.0) : sqrt(sint);
        ort_sincos(ort_mul_rn(R(ORT_TWOPI), ort_scatter_draw(g, sr)), &sinp, &cosp);
    } else { /* Henyey-Greenstein, src/stokes.f90:54-158 */
        R g2 = ort_mul_rn(hgg, hgg);
        R costp = cost, sintp = sint;
        R den = ort_add_rn(ort_sub_rn(R(1.0), hgg), ort_mul_rn(ort_mul_rn(R(2.0), hgg), ort_scatter_draw(g, sr)));
        R tq = ort_sub_rn(R(1.0), g2) / den;
        R bmu = ort_sub_rn(ort_add_rn(R(1.0), g2), ort_mul_rn(tq, tq)) / ort_mul_rn(R(2.0), hgg);
        R cosb2 = ort_mul_rn(bmu, bmu);
        if (fabs(bmu) > R(1.0)) {
            bmu = (bmu > R(1.0)) ? R(1.0) : -R(1.0);
            cosb2 = R(1.0);
        }
        R sinbt = sqrt(ort_sub_rn(R(1.0), cosb2));
        R ri1 = ort_mul_rn(R(ORT_TWOPI), ort_scatter_draw(g, sr));
        /* the reference's two branches (ri1 > pi uses ri3 = 2pi - ri1 and adds acos; otherwise
         * subtracts) differ only in the sign applied to acos(cosdph) */
        bool upper = ri1 > R(ORT_PI);
        R ang = upper ? ort_sub_rn(R(ORT_TWOPI), ri1) : ri1;
        R sini, cosi;
        ort_sincos(ang, &sini, &cosi);
        if (bmu == R(1.0) || bmu == -R(1.0)) return; /* goto 100: direction unchanged */
        cost = ort_add_rn(ort_mul_rn(costp, bmu), ort_mul_rn(ort_mul_rn(sintp, sinbt), cosi));
        R sini2, cosi2 = R(0.0);
        if (fabs(cost) < R(1.0)) {
            sint = fabs(sqrt(ort_sub_rn(R(1.0), ort_mul_rn(cost, cost))));
            sini2 = ort_mul_rn(sini, sintp) / sint;
            R bott = ort_mul_rn(sint, sinbt);
            cosi2 = ort_sub_rn(costp / bott, ort_mul_rn(cost, bmu) / bott);
        } else {
            sint = R(0.0);
            sini2 = R(0.0);
            if (cost >= R(1.0)) cosi2 = -R(1.0);
            if (cost <= -R(1.0)) cosi2 = R(1.0);
        }
        R cosdph = ort_add_rn(-ort_mul_rn(cosi2, cosi), ort_mul_rn(ort_mul_rn(sini2, sini), bmu));
        if (fabs(cosdph) > R(1.0)) cosdph = (cosdph > R(1.0)) ? R(1.0) : -R(1.0);
        R dph = acos(cosdph);
        phi = upper ? ort_add_rn(phi, dph) : ort_sub_rn(phi, dph);
        if (phi > R(ORT_TWOPI)) phi = ort_sub_rn(phi, R(ORT_TWOPI));
        if (phi < R(0.0)) phi = ort_add_rn(phi, R(ORT_TWOPI));
        ort_sincos(phi, &sinp, &cosp);
    }
    r.dx = ort_mul_rn(sint, cosp);
    r.dy = ort_mul_rn(sint, sinp);
    r.dz = cost;
}

/* -------------------------------------------------------------------------------------------
 * glass_bottle%forward, src/lens.f90:230-350.  Returns 0 or the ort_status that ended the ray.
 * ----------------------------------------------------------------------------------------- */
/* The clear bottle: straight through. */
template <typename R>
ORT_HD int ort_bottle_clear(const DevSceneT<R>& S, const OrtRng& g, const OrtDraws01& D, OrtRayT<R>& r) {
    R t;
    const R u_in = ort_wide2<R>(g, D.b[0], D.hi2), u_out = ort_narrow2<R>(g, D.b[2]); /* slots 2, 3, doubled */
    bool hit = S.ellipse ? ort_hit_ellipse(r, S.bcy, S.bcz, S.b_in_ia2, S.b_in_ib2, &t)
                         : ort_hit_cylinder(r, S.bcy, S.bcz, S.b_in_r2, &t);
    if (!hit) return ORT_ST_BOTTLE_INNER_MISS;
    ort_advance(r, t);
    {   /* radial normal in the (y,z) plane, also for the ellipse (src/lens.f90:288-290) */
        R ny = S.bcy - r.py, nz = S.bcz - r.pz;
        /* on a clear cylindrical wall the hit point is on the cylinder: |(ny,nz)| = radius; on an
         * ellipse it is not, and the length is computed */
        R inv = S.ellipse ? ort_rsqrt(fma(ny, ny, nz * nz)) : S.b_in_invr;
        R nzu = ort_both_zero(ny, R(0.0)) ? copysign(R(1.0), nz) : nz * inv; /* on-axis ray: exactly +-1 */
        if (ort_interface(r, R(0.0), ny * inv, nzu, S.b_in, u_in)) return ORT_ST_BOTTLE_INNER_REFLECT;
    }
    hit = S.ellipse ? ort_hit_ellipse(r, S.bcy, S.bcz, S.b_out_ia2, S.b_out_ib2, &t)
                    : ort_hit_cylinder(r, S.bcy, S.bcz, S.b_out_r2, &t);
    if (!hit) return ORT_ST_BOTTLE_OUTER_MISS;
    ort_advance(r, t);
    {
        R ny = S.bcy - r.py, nz = S.bcz - r.pz;
        R inv = S.ellipse ? ort_rsqrt(fma(ny, ny, nz * nz)) : S.b_out_invr;
        R nzu = ort_both_zero(ny, R(0.0)) ? copysign(R(1.0), nz) : nz * inv;
        if (ort_interface(r, R(0.0), ny * inv, nzu, S.b_out, u_out)) return ORT_ST_BOTTLE_OUTER_REFLECT;
    }
    return 0;
}

/* The scattering bottle as a RESUMABLE walk.  The reference runs two `do while` loops per ray (contents
 * src/lens.f90:262-282, wall :312-333) whose trip counts are geometric random numbers: run per lane, a
 * warp idles on its longest chain (9.7 of 32 lanes busy in BASELINE config 4).  Here one pass of a loop
 * body is a unit of work of its own (ort_scatter_event), the walk between the loops is ort_bottle_resume,
 * and the state in between -- the pending step, the position in the sequential draw stream, which loop
 * -- is explicit, so that the kernel can queue rays between events and always run 32 of them.
 *   ort_bottle_resume(from): 0 = the ray has just been emitted, 1 = its contents loop has ended,
 *   2 = its wall loop has ended.  Returns a final status (> 0), 0 (through the bottle), or
 *   ORT_BOTTLE_EVENT: the ray is inside loop ss.loop and owes one pass of its body. */
#define ORT_BOTTLE_EVENT (-2)
template <typename R>
struct OrtScatterStateT {
    OrtScatterRngT<R> sr; /* next sequential slot (16, 17, ...) and the spare of the last block */
    R t;                  /* the step still to be taken */
    int loop;             /* 0 contents, 1 wall */
};
template <typename R>
ORT_HD int ort_bottle_resume(const DevSceneT<R>& S, const OrtRng& g, const OrtDraws01& D, OrtRayT<R>& r,
                             OrtScatterStateT<R>& ss, int from) {
    bool flag;
    if (from == 0) {
        ss.sr.next = 16;
        ss.sr.spare = R(0.0);
        bool hit = S.ellipse ? ort_hit_ellipse(r, S.bcy, S.bcz, S.b_in_ia2, S.b_in_ib2, &ss.t)
                             : ort_hit_cylinder(r, S.bcy, S.bcz, S.b_in_r2, &ss.t);
        if (!hit) return ORT_ST_BOTTLE_INNER_MISS;
        if (S.scatter_c) { /* :263-265; tauint always uses the cylinder */
            if (!ort_tauint(r, S.mutot_c, S.inv_mutot_c, S.bcy, S.bcz, S.b_in_r2, ort_scatter_draw(g, ss.sr), &ss.t, &flag))
                return ORT_ST_TAUINT_MISS;
            if (!flag) {
                ss.loop = 0;
                return ORT_BOTTLE_EVENT;
            }
            if (r.dz < R(0.0)) return ORT_ST_CONTENTS_BACKWARD;
        }
    } else if (from == 1) {
        if (r.dz < R(0.0)) return ORT_ST_CONTENTS_BACKWARD; /* :278-281 */
    }
    if (from <= 1) {
        ort_advance(r, ss.t);
        {   /* radial normal in the (y,z) plane, also for the ellipse (src/lens.f90:288-290); after a scatter
             * loop (quirk 4) the point need not lie on the cylinder: the length is computed */
            R ny = S.bcy - r.py, nz = S.bcz - r.pz;
            R inv = ort_rsqrt(fma(ny, ny, nz * nz));
            R nzu = ort_both_zero(ny, R(0.0)) ? copysign(R(1.0), nz) : nz * inv; /* on-axis ray: exactly +-1 */
            if (ort_interface(r, R(0.0), ny * inv, nzu, S.b_in, ort_wide2<R>(g, D.b[0], D.hi2))) /* slot 2 */
                return ORT_ST_BOTTLE_INNER_REFLECT;
        }
        bool hit = S.ellipse ? ort_hit_ellipse(r, S.bcy, S.bcz, S.b_out_ia2, S.b_out_ib2, &ss.t)
                             : ort_hit_cylinder(r, S.bcy, S.bcz, S.b_out_r2, &ss.t);
        if (!hit) return ORT_ST_BOTTLE_OUTER_MISS;
        if (S.scatter_b) { /* :313-315 */
            if (!ort_tauint(r, S.mutot_b, S.inv_mutot_b, S.bcy, S.bcz, S.b_out_r2, ort_scatter_draw(g, ss.sr), &ss.t, &flag))
                return ORT_ST_TAUINT_MISS;
            if (!flag) {
                ss.loop = 1;
                return ORT_BOTTLE_EVENT;
            }
            if (r.dz < R(0.0)) return ORT_ST_WALL_BACKWARD;
        }
    } else {
        if (r.dz < R(0.0)) return ORT_ST_WALL_BACKWARD; /* :329-332 */
    }
    ort_advance(r, ss.t);
    {
        R ny = S.bcy - r.py, nz = S.bcz - r.pz;
        R inv = ort_rsqrt(fma(ny, ny, nz * nz));
        R nzu = ort_both_zero(ny, R(0.0)) ? copysign(R(1.0), nz) : nz * inv;
        if (ort_interface(r, R(0.0), ny * inv, nzu, S.b_out, ort_narrow2<R>(g, D.b[2]))) /* slot 3 */
            return ORT_ST_BOTTLE_OUTER_REFLECT;
    }
    return 0;
}
/* One pass of the body of loop ss.loop (:266-277 / :316-328): step, albedo draw, stokes, next tauint, the
 * reference's exit test.  Returns a final status (> 0), ORT_BOTTLE_EVENT (another pass is owed) or 0 (the
 * loop has ended: ort_bottle_resume(ss.loop + 1) goes on).  *scattered = stokes was called. */
template <typename R>
ORT_HD int ort_scatter_event(const DevSceneT<R>& S, const OrtRng& g, OrtRayT<R>& r, OrtScatterStateT<R>& ss,
                             bool* scattered) {
    const bool wall = ss.loop != 0;
    const R mutot = wall ? S.mutot_b : S.mutot_c, inv_mutot = wall ? S.inv_mutot_b : S.inv_mutot_c;
    const R albedo = wall ? S.albedo_b : S.albedo_c, hgg = wall ? R(0.9) : R(0.65);
    const R Rlim = wall ? S.b_out_r : S.b_in_r, Rlim2 = wall ? S.b_out_r2 : S.b_in_r2;
    *scattered = false;
    ort_advance(r, ss.t);
    if (!(ort_scatter_draw(g, ss.sr) < albedo)) return wall ? ORT_ST_WALL_ABSORBED : ORT_ST_CONTENTS_ABSORBED;
    ort_stokes(r, hgg, g, ss.sr);
    *scattered = true;
    bool flag;
    if (!ort_tauint(r, mutot, inv_mutot, S.bcy, S.bcz, Rlim2, ort_scatter_draw(g, ss.sr), &ss.t, &flag))
        return ORT_ST_TAUINT_MISS;
    /* the reference's exit test uses (x,z) although the axis is x */
    if (sqrt(fma(r.px, r.px, r.pz * r.pz)) >= Rlim) return 0;
    return flag ? 0 : ORT_BOTTLE_EVENT;
}

/* glass_bottle%forward for one ray from start to end.  SCATTER: the two functions above driven per ray
 * (explicit-ray entry point, diagnostic flat kernel, host harness; the production kernel for scattering
 * bottles, ort_trace_scatter_kernel, queues the rays between events instead).  *nevents counts the
 * scatter events (stokes calls) of this ray. */
template <bool SCATTER, typename R>
ORT_HD int ort_bottle_forward(const DevSceneT<R>& S, const OrtRng& g, const OrtDraws01& D, OrtRayT<R>& r, int* nevents = nullptr) {
    if (!SCATTER) return ort_bottle_clear(S, g, D, r);
    OrtScatterStateT<R> ss;
    int st = ort_bottle_resume(S, g, D, r, ss, 0);
    while (st == ORT_BOTTLE_EVENT) {
        bool scattered;
        st = ort_scatter_event(S, g, r, ss, &scattered);
        if (scattered && nevents) ++*nevents;
        if (st == 0) st = ort_bottle_resume(S, g, D, r, ss, ss.loop + 1);
    }
    return st;
}

/* -------------------------------------------------------------------------------------------
 * plano_convex%forward, src/lens.f90:425-481, split at the aperture test
 * ----------------------------------------------------------------------------------------- */
template <typename R>
ORT_HD int ort_l2_enter(const DevSceneT<R>& S, OrtRayT<R>& r) { /* :447-454 */
    R d = ort_div_z(S.l2_flat_z - r.pz, r.dz);
    ort_advance(r, d);
    if (fma(r.px, r.px, r.py * r.py) > S.l2_radius2) return ORT_ST_L2_APERTURE;
    return 0;
}
/* w_flat, w_curved: the decision words of slots 4 and 5 (words 3 of blocks 0 and 1) */
template <typename R>
ORT_HD int ort_l2_body(const DevSceneT<R>& S, const OrtRng& g, uint32_t w_flat, uint32_t w_curved, OrtRayT<R>& r) { /* :458-479 */
    R t;
    const R u_flat = ort_narrow2<R>(g, w_flat), u_curved = ort_narrow2<R>(g, w_curved); /* doubled */
    /* a reflection at the flat face is computed but never tested */
    (void)ort_interface(r, S.l2_fnx, S.l2_fny, S.l2_fnz, S.l2_in, u_flat);
    if (!ort_hit_sphere(r, S.l2_cx, S.l2_cy, S.l2_cz, S.l2_R2, &t)) return ORT_ST_L2_SPHERE_MISS;
    ort_advance(r, t);
    R nx, ny, nz;
    ort_sphere_normal(r, S.l2_cx, S.l2_cy, S.l2_cz, S.l2_invR, &nx, &ny, &nz);
    if (ort_interface(r, nx, ny, nz, S.l2_out, u_curved)) return ORT_ST_L2_CURVED_REFLECT;
    return 0;
}

/* -------------------------------------------------------------------------------------------
 * achromatic_doublet%forward, src/lens.f90:531-645, split after the first-surface aperture test
 * ----------------------------------------------------------------------------------------- */
template <typename R>
ORT_HD int ort_l3_enter(const DevSceneT<R>& S, bool iris_before, OrtRayT<R>& r) { /* :551-580 */
    R t;
    if (iris_before) {
        t = ort_div_z(S.l3_iris1_z - r.pz, r.dz);
        R x = fma(r.dx, t, r.px), y = fma(r.dy, t, r.py);
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
template <typename R>
ORT_HD int ort_l3_body(const DevSceneT<R>& S, const OrtRng& g, bool iris_after, OrtRayT<R>& r) { /* :582-644 */
    R t, nx, ny, nz;
    uint32_t w[4];
    ort_block(g, 2u, w); /* slots 6, 7, 8 */
    const R u1 = ort_narrow2<R>(g, w[0]), u2 = ort_narrow2<R>(g, w[1]), u3 = ort_narrow2<R>(g, w[2]); /* doubled */
    ort_sphere_normal(r, S.l3_c1x, S.l3_c1y, S.l3_c1z, S.l3_invR1, &nx, &ny, &nz);
    if (ort_interface(r, nx, ny, nz, S.l3_s1, u1)) return ORT_ST_L3_S1_REFLECT;
    if (!ort_hit_sphere(r, S.l3_c2x, S.l3_c2y, S.l3_c2z, S.l3_R2_2, &t)) return ORT_ST_L3_S2_MISS;
    ort_advance(r, t);
    ort_sphere_normal(r, S.l3_c2x, S.l3_c2y, S.l3_c2z, S.l3_invR2, &nx, &ny, &nz);
    if (ort_interface(r, nx, ny, nz, S.l3_s2, u2)) return ORT_ST_L3_S2_REFLECT;
    /* the reference aborts here on a miss (error stop "Help3", :617); we count it */
    if (!ort_hit_sphere(r, S.l3_c3x, S.l3_c3y, S.l3_c3z, S.l3_R3_2, &t)) return ORT_ST_L3_S3_MISS;
    ort_advance(r, t);
    ort_sphere_normal(r, S.l3_c3x, S.l3_c3y, S.l3_c3z, S.l3_invR3, &nx, &ny, &nz);
    if (ort_interface(r, nx, ny, nz, S.l3_s3, u3)) return ORT_ST_L3_S3_REFLECT;
    if (iris_after) {
        t = ort_div_z(S.l3_iris2_z - r.pz, r.dz);
        R x = fma(r.dx, t, r.px), y = fma(r.dy, t, r.py);
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
template <typename R>
ORT_HD int ort_image(const DevSceneT<R>& S, OrtRayT<R>& r, int* xp, int* yp) {
    R d = ort_div_z(S.img_z - r.pz, r.dz);
    ort_advance(r, d);
    /* angle = acos(dz/|d|) > asin(0.22)  <=>  dz < cos_na |d|;  a NaN angle passes (reference) */
    R dd = fma(r.dx, r.dx, fma(r.dy, r.dy, r.dz * r.dz));
    if (r.dz <= R(0.0) || r.dz * r.dz < S.cos_na2 * dd) return ORT_ST_NA_REJECT;
    R fx = floor(r.px * S.inv_binwid), fy = floor(r.py * S.inv_binwid);
    if (!(fmax(fabs(fx), fabs(fy)) <= R(200.0))) { /* off the detector, far away, or not finite */
        if (r.px > R(1000.0) || r.py > R(1000.0)) return ORT_ST_FAR;
        if (!(fabs(fx) < R(2.0e9)) || !(fabs(fy) < R(2.0e9))) return ORT_ST_FAR;
        return ORT_ST_OFF_DETECTOR;
    }
    *xp = (int)fx;
    *yp = (int)fy;
    return ORT_ST_BINNED;
}

/* -------------------------------------------------------------------------------------------
 * One whole iteration of the reference's ray loops for a single ray (src/main.f90:90-109 /
 * :127-162 incl. telescope, src/optics_system.f90:6-52), with the explicit-ray conveniences of
 * ort_trace_rays: optional caller-supplied start state and ort_job.stop_after.
 * ----------------------------------------------------------------------------------------- */
template <typename R>
ORT_HD int ort_full_path(const DevSceneT<R>& S, const DevJob& J, const OrtRng& g, bool have_input, OrtRayT<R>& r,
                         int* xp, int* yp) {
    const int stop = J.stop_after;
    int st;
    OrtDraws01 D;
    ort_draws01(g, D);
    if (!have_input) {
        long long ray = ((long long)g.r1 << 32) | g.r0;
        int es;
        if (J.phase == ORT_PHASE_RING) {
            es = J.source_kind == ORT_SRC_CRS ? ort_emit<ORT_PHASE_RING, ORT_SRC_CRS>(S, J, g, D, ray, r)
               : J.source_kind == ORT_SRC_ISORS ? ort_emit<ORT_PHASE_RING, ORT_SRC_ISORS>(S, J, g, D, ray, r)
                                                : ort_emit<ORT_PHASE_RING, ORT_SRC_POINT>(S, J, g, D, ray, r);
        } else {
            es = J.source_kind == ORT_SRC_SPOT ? ort_emit<ORT_PHASE_POINT, ORT_SRC_SPOT>(S, J, g, D, ray, r)
               : J.source_kind == ORT_SRC_IMAGE ? ort_emit<ORT_PHASE_POINT, ORT_SRC_IMAGE>(S, J, g, D, ray, r)
                                                : ort_emit<ORT_PHASE_POINT, ORT_SRC_POINT>(S, J, g, D, ray, r);
        }
        if (es) return es;
    }
    if (stop == ORT_STOP_SOURCE) return ORT_ST_STOPPED;
    if (J.phase == ORT_PHASE_POINT && J.use_bottle) {
        st = (S.scatter_b | S.scatter_c) ? ort_bottle_forward<true>(S, g, D, r) : ort_bottle_forward<false>(S, g, D, r);
        if (st) return st;
    }
    if (stop == ORT_STOP_BOTTLE) return ORT_ST_STOPPED;
    st = ort_l2_enter(S, r);
    if (st) return st;
    st = ort_l2_body(S, g, D.a[3], D.b[3], r);
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
