/*
 * ort_filter.cuh -- the ring loop's single-precision culling filter, with a running error bound
 * (no counterpart in the reference).
 *
 * 99.5 % of the ring rays that pass L2's aperture still end between L2's curved face and L3's
 * aperture test (total reflection in L2, no intersection with L3's first sphere, outside L3's
 * aperture): they add one to a status counter and nothing else.  WHICH counter is a chain of sign
 * decisions -- discriminants, aperture radii, the reflect-or-refract draw against the Fresnel
 * reflectance.  ortf_filter walks source -> L2 -> L3's first surface in fp32 (one issue slot per
 * FFMA instead of a multi-cycle DFMA, one MUFU per rcp / rsqrt / sin / cos) and returns
 *     s > 0 : the ray ends with status s -- PROVABLY what the fp64 path decides;
 *     0     : the ray survives to L3, or some decision could not be proved -> the caller traces it
 *             in fp64 (ort_stage_b), which alone moves rays forward.
 *
 * The proof is a running forward error analysis (DESIGN.md section 3.1c has the derivation; the
 * comments below name the rule each line applies).  Next to the ray the filter carries
 *     ep >= | pos~ - pos* |_2      ed >= | dir~ - dir* |_2      en >= | normal~ - normal* |_2
 * where x~ is the fp32 value and x* what exact arithmetic gives on the exact inputs (the 53-bit /
 * 32-bit draws and the fp64 scene).  Every decision value v~ gets a bound ev >= |v~ - v*| from
 *   (R1) fl(a op b) = (a op b)(1 + d), |d| <= u = 2^-24, for +, -, *, fma (round to nearest);
 *   (R2) |a~ b~ - a* b*| <= |a~| eb + |b*| ea, and its dot-product form with Euclidean norms;
 *   (R3) |sqrt a~ - sqrt a*| <= ea / sqrt a~;  |1/a~ - 1/a*| <= ea / (|a~| (|a~| - ea));
 *   (R4) the MUFU approximations: relative error <= ORTF_E_RCP / ORTF_E_RSQ / ORTF_E_SQRT, absolute
 *        error of sin / cos on [-pi, pi] <= ORTF_E_SIN -- twice the largest error
 *        ort_mufu_selftest finds when it tries EVERY fp32 argument on the device;
 * and the filter calls the decision only when |v~| > ev.  The fp64 path computes the same
 * functions with unit roundoff 2^-53, so it sits within 2^-29 of these same bounds from x*; the
 * constants carry a factor >= 1.01 (and u is counted as 2^-23 where it multiplies a scene length),
 * which covers that, the rounding of the bound arithmetic itself (sums and products of
 * non-negative numbers: relative error <= 60 u), and the second-order terms under the guards
 * G1..G5 below.  A ray that trips a guard, or any test involving a NaN, goes to fp64.
 *
 * tests/test_filter_bound.py checks every one of these bounds against a double-precision twin of
 * the filter on millions of rays and on randomised / extreme scenes, with the MUFU results
 * perturbed by the full assumed error; ORT_FLAG_VERIFY_FILTER runs filter and fp64 on every ray on
 * the device and counts disagreements.
 */
#ifndef ORT_FILTER_CUH
#define ORT_FILTER_CUH

#include "ort_optics.cuh"

#define ORTF_U 5.9604644775390625e-8f    /* 2^-24 */
#define ORTF_E_RCP 2.384185791015625e-7f  /* 2^-22, relative */
#define ORTF_E_RSQ 2.6e-7f                /* relative */
#define ORTF_E_SQRT 2.384185791015625e-7f /* 2^-22, relative */
#define ORTF_E_SIN 1.1e-6f                /* absolute, arguments in [-3.1416, 3.1416] */
/* measured on a B200 (profiles/r02_filter_verify.txt): rcp 2^-23.28, rsqrt 2^-22.94, sqrt 2^-23.25,
 * sin 2^-20.94, cos 2^-21.24 */

/* Test hook (host harness only): ORTF_FUZZ multiplies every MUFU result by 1 +- the assumed error,
 * sign chosen by the caller's generator, so that the bounds are exercised at their limit. */
#if defined(ORTF_FUZZ) && !defined(__CUDA_ARCH__)
extern "C" float ortf_fuzz_sign(void);
#define ORTF_PERTURB_REL(y, e) ((y) * (1.0f + ortf_fuzz_sign() * (e)))
#define ORTF_PERTURB_ABS(y, e) ((y) + ortf_fuzz_sign() * (e))
#else
#define ORTF_PERTURB_REL(y, e) (y)
#define ORTF_PERTURB_ABS(y, e) (y)
#endif

ORT_HD float ortf_rcp(float x) {
#ifdef __CUDA_ARCH__
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
#else
    return ORTF_PERTURB_REL(1.0f / x, 0.5f * ORTF_E_RCP);
#endif
}
ORT_HD float ortf_rsqrt(float x) {
#ifdef __CUDA_ARCH__
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
#else
    return ORTF_PERTURB_REL(1.0f / sqrtf(x), 0.5f * ORTF_E_RSQ);
#endif
}
ORT_HD float ortf_sqrt(float x) {
#ifdef __CUDA_ARCH__
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
#else
    return ORTF_PERTURB_REL(sqrtf(x), 0.5f * ORTF_E_SQRT);
#endif
}
/* sin, cos of 2 pi u for the narrow draw u = w 2^-32.  Read as a SIGNED integer the word is 2^32 (u - [u >= 1/2]):
 * the same angle, folded to [-pi, pi), from one conversion and one product.  The conversion is off by
 * <= 2^-25 |w_s| <= 2^-26 of a turn (less than the 2^-25 the bound allows for an unsigned word). */
ORT_HD void ortf_sincos_word(uint32_t w, float* s, float* c) {
    float a = (float)(int32_t)w * 1.4629180792671596e-9f; /* 2 pi 2^-32 */
#ifdef __CUDA_ARCH__
    asm("sin.approx.ftz.f32 %0, %1;" : "=f"(*s) : "f"(a));
    asm("cos.approx.ftz.f32 %0, %1;" : "=f"(*c) : "f"(a));
#else
    *s = ORTF_PERTURB_ABS(sinf(a), 0.5f * ORTF_E_SIN);
    *c = ORTF_PERTURB_ABS(cosf(a), 0.5f * ORTF_E_SIN);
#endif
}
/* A 32-bit word (a narrow draw, or the high word of a wide one) as a float: 2^32 times a uniform that is
 * off the exact draw by <= 2^-25 (+ 2^-32 for the unseen low word of a wide draw).  The factor 2^-32 is
 * folded into whatever the draw multiplies (powers of two: the products are the same floats). */
ORT_HD float ortf_word(uint32_t w) {
    return (float)w;
}

/* ---- lanes --------------------------------------------------------------------------------------
 * The filter is written ONCE, over a policy P that says what a "value" is:
 *   OrtfOne  a float: one ray per lane -- the host harness (bound tests, double twin), documentation;
 *   OrtfTwo  two floats in one 64-bit register: TWO rays per lane, arithmetic (on the device) as fma / mul / add / sub
 *            .rn.f32x2 (FFMA2 & co. on sm_100a: one issue slot does the operation for both rays -- the
 *            culling kernel is bound by instruction issue, and a third of its instructions are these);
 *            comparisons, selects and the MUFU approximations are done per half.
 * Both perform the same operations in the same order on each ray, so the error analysis is one (where the
 * compiler fuses a product into the following sum there is one rounding instead of two: (R1) charges each
 * operation its own, which covers both).
 * Negations are written so that they cost nothing in either form (a constant stored negated, a
 * select that negates, or an explicit multiplication by -1, which is exact). */
struct OrtfOne {
    typedef float V;      /* value */
    typedef bool M;       /* outcome of a comparison */
    typedef int S;        /* ray status */
    typedef uint32_t W;   /* a 32-bit word of the generator */
    static constexpr bool kTrace = true;
    static ORT_HD V lit(float k) { return k; }
    static ORT_HD V fma(V a, V b, V c) { return fmaf(a, b, c); }
    static ORT_HD V mul(V a, V b) { return a * b; }
    static ORT_HD V add(V a, V b) { return a + b; }
    static ORT_HD V sub(V a, V b) { return a - b; }
    static ORT_HD V neg(V a) { return -a; }
    static ORT_HD V abs(V a) { return fabsf(a); }
    static ORT_HD M gt(V a, V b) { return a > b; }
    static ORT_HD M lt(V a, V b) { return a < b; }
    static ORT_HD M ngt(V a, V b) { return !(a > b); }        /* also true for a NaN */
    static ORT_HD M nlt(V a, V b) { return !(a < b); }
    static ORT_HD M abs_ngt(V a, V b) { return !(fabsf(a) > b); }
    static ORT_HD V sel(M m, V a, V b) { return m ? a : b; }
    static ORT_HD V sel_na(M m, V a, V b) { return m ? -a : b; } /* m ? -a : b */
    static ORT_HD M mor(M a, M b) { return a || b; }
    static ORT_HD M mand(M a, M b) { return a && b; }
    static ORT_HD M mnot(M a) { return !a; }
    static ORT_HD M mfalse() { return false; }
    static ORT_HD bool all(M a) { return a; }
    static ORT_HD void sset(S& st, M m, int code) { st = m ? code : st; }
    static ORT_HD S snone() { return 0; }
    static ORT_HD V rcp(V a) { return ortf_rcp(a); }
    static ORT_HD V rsqrt(V a) { return ortf_rsqrt(a); }
    static ORT_HD V sqrt(V a) { return ortf_sqrt(a); }
    static ORT_HD V word(W w) { return ortf_word(w); }
    static ORT_HD void sincos_word(W w, V* s, V* c) { ortf_sincos_word(w, s, c); }
    static ORT_HD M wlt(W w, uint32_t k) { return w < k; }
    /* what the trace records (host only) */
    static ORT_HD float first(V a) { return a; }
    static ORT_HD bool first(M a) { return a; }
};

/* two floats in one 64-bit register; on the host (harness: the pair logic under test) the same bits, the
 * operations done per half */
struct OrtfV2 { unsigned long long v; };
struct OrtfM2 { bool a, b; };
struct OrtfS2 { int a, b; };
struct OrtfW2 { uint32_t x, y; };
ORT_HD OrtfV2 ortf_pk(float a, float b) {
    OrtfV2 r;
#ifdef __CUDA_ARCH__
    asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(a), "f"(b));
#else
    uint32_t x, y;
    memcpy(&x, &a, 4);
    memcpy(&y, &b, 4);
    r.v = ((unsigned long long)y << 32) | x;
#endif
    return r;
}
ORT_HD float ortf_lo(OrtfV2 a) {
    float x;
#ifdef __CUDA_ARCH__
    asm("{\n\t.reg .f32 t;\n\tmov.b64 {%0, t}, %1;\n\t}" : "=f"(x) : "l"(a.v));
#else
    const uint32_t b = (uint32_t)a.v;
    memcpy(&x, &b, 4);
#endif
    return x;
}
ORT_HD float ortf_hi(OrtfV2 a) {
    float y;
#ifdef __CUDA_ARCH__
    asm("{\n\t.reg .f32 t;\n\tmov.b64 {t, %0}, %1;\n\t}" : "=f"(y) : "l"(a.v));
#else
    const uint32_t b = (uint32_t)(a.v >> 32);
    memcpy(&y, &b, 4);
#endif
    return y;
}
struct OrtfTwo {
    typedef OrtfV2 V;
    typedef OrtfM2 M;
    typedef OrtfS2 S;
    typedef OrtfW2 W;
    static constexpr bool kTrace = false;
    static ORT_HD V lit(float k) { return ortf_pk(k, k); }
    static ORT_HD V fma(V a, V b, V c) {
#ifdef __CUDA_ARCH__
        V r;
        asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v));
        return r;
#else
        return ortf_pk(fmaf(ortf_lo(a), ortf_lo(b), ortf_lo(c)), fmaf(ortf_hi(a), ortf_hi(b), ortf_hi(c)));
#endif
    }
    static ORT_HD V mul(V a, V b) {
#ifdef __CUDA_ARCH__
        V r;
        asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
        return r;
#else
        return ortf_pk(ortf_lo(a) * ortf_lo(b), ortf_hi(a) * ortf_hi(b));
#endif
    }
    static ORT_HD V add(V a, V b) {
#ifdef __CUDA_ARCH__
        V r;
        asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
        return r;
#else
        return ortf_pk(ortf_lo(a) + ortf_lo(b), ortf_hi(a) + ortf_hi(b));
#endif
    }
    static ORT_HD V sub(V a, V b) {
#ifdef __CUDA_ARCH__
        V r;
        asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
        return r;
#else
        return ortf_pk(ortf_lo(a) - ortf_lo(b), ortf_hi(a) - ortf_hi(b));
#endif
    }
    static ORT_HD V neg(V a) { return mul(a, lit(-1.0f)); } /* exact */
    static ORT_HD V abs(V a) { return ortf_pk(fabsf(ortf_lo(a)), fabsf(ortf_hi(a))); }
    static ORT_HD M gt(V a, V b) { return M{ortf_lo(a) > ortf_lo(b), ortf_hi(a) > ortf_hi(b)}; }
    static ORT_HD M lt(V a, V b) { return M{ortf_lo(a) < ortf_lo(b), ortf_hi(a) < ortf_hi(b)}; }
    static ORT_HD M ngt(V a, V b) { return M{!(ortf_lo(a) > ortf_lo(b)), !(ortf_hi(a) > ortf_hi(b))}; }
    static ORT_HD M nlt(V a, V b) { return M{!(ortf_lo(a) < ortf_lo(b)), !(ortf_hi(a) < ortf_hi(b))}; }
    static ORT_HD M abs_ngt(V a, V b) {
        return M{!(fabsf(ortf_lo(a)) > ortf_lo(b)), !(fabsf(ortf_hi(a)) > ortf_hi(b))};
    }
    static ORT_HD V sel(M m, V a, V b) { return ortf_pk(m.a ? ortf_lo(a) : ortf_lo(b), m.b ? ortf_hi(a) : ortf_hi(b)); }
    static ORT_HD V sel_na(M m, V a, V b) { return ortf_pk(m.a ? -ortf_lo(a) : ortf_lo(b), m.b ? -ortf_hi(a) : ortf_hi(b)); }
    static ORT_HD M mor(M x, M y) { return M{x.a || y.a, x.b || y.b}; }
    static ORT_HD M mand(M x, M y) { return M{x.a && y.a, x.b && y.b}; }
    static ORT_HD M mnot(M x) { return M{!x.a, !x.b}; }
    static ORT_HD M mfalse() { return M{false, false}; }
    static ORT_HD bool all(M x) { return x.a && x.b; }
    static ORT_HD void sset(S& st, M m, int code) {
        st.a = m.a ? code : st.a;
        st.b = m.b ? code : st.b;
    }
    static ORT_HD S snone() { return S{0, 0}; }
    static ORT_HD V rcp(V a) { return ortf_pk(ortf_rcp(ortf_lo(a)), ortf_rcp(ortf_hi(a))); }
    static ORT_HD V rsqrt(V a) { return ortf_pk(ortf_rsqrt(ortf_lo(a)), ortf_rsqrt(ortf_hi(a))); }
    static ORT_HD V sqrt(V a) { return ortf_pk(ortf_sqrt(ortf_lo(a)), ortf_sqrt(ortf_hi(a))); }
    static ORT_HD V word(W w) { return ortf_pk(ortf_word(w.x), ortf_word(w.y)); }
    static ORT_HD void sincos_word(W w, V* s, V* c) {
        float s0, c0, s1, c1;
        ortf_sincos_word(w.x, &s0, &c0);
        ortf_sincos_word(w.y, &s1, &c1);
        *s = ortf_pk(s0, s1);
        *c = ortf_pk(c0, c1);
    }
    static ORT_HD M wlt(W w, uint32_t k) { return M{w.x < k, w.y < k}; }
    static ORT_HD float first(V a) { return ortf_lo(a); }
    static ORT_HD bool first(M a) { return a.a; }
};

/* Every scene or bound constant the filter touches, in the policy's value type (for OrtfTwo: the float
 * twice, so that it is one 64-bit constant-bank operand).  ortf_make_params fills it from the fp32 scene
 * and ort_make_filter's constants; the names are theirs. */
template <typename C>
struct OrtfParamsT {
    C r2m_s, r1, ra_over_rb, bcz, ra2, lens_r2_s, l2_radius2_lim, l2_fb, l2_flat_z, ed_a, ed_b, ep_flat, ep_max;
    /* L2 flat face */
    C in_neg_eta2, in_eta, fl_s2_a, fl_s2_b, fl_f_a, fl_f_b, fl_d_a, fl_d_b;
    /* L2 sphere, from the flat face */
    C l2_cz, l2_R2, l2_invR, l2_neg_invR, s2_h_d, s2_h_0, s2_c_0, s2_d_h, s2_d_0, s2_p_0, s2_n_p, s2_n_0;
    /* L2 curved face */
    C out_neg_eta2, out_eta, cv_ni_0, cv_ct2_a, cv_ct2_b, cv_cs_0, cv_f_a, cv_f_b, cv_k_a, cv_k_0, cv_d_d, cv_d_n, cv_d_a, cv_d_0;
    /* L3: iris, first sphere, aperture */
    C iris1_z, iris_z0, iris_r2, iris_inv, iris_r, iris_0;
    C l3_c1z, l3_R1_2, s3_h_d, s3_h_p, s3_h_0, s3_c_p, s3_c_0, s3_d_h, s3_d_0, s3_p_0;
    C l3_radius2, ap_inv, ap_r, ap_0;
    int32_t ellipse, iris_before;
};
template <typename C> ORT_HD C ortf_dup(float k);
template <> ORT_HD float ortf_dup<float>(float k) { return k; }
template <> ORT_HD unsigned long long ortf_dup<unsigned long long>(float k) {
    uint32_t b;
    memcpy(&b, &k, 4);
    return ((unsigned long long)b << 32) | b;
}
template <> ORT_HD OrtfV2 ortf_dup<OrtfV2>(float k) {
    OrtfV2 r;
    r.v = ortf_dup<unsigned long long>(k);
    return r;
}
template <typename C>
ORT_HD void ortf_make_params(const DevSceneT<float>& F, const DevFilter& K, int iris_before, OrtfParamsT<C>& q) {
#define ORTF_SET(field, value) q.field = ortf_dup<C>(value)
    ORTF_SET(r2m_s, K.r2m_s); ORTF_SET(r1, F.r1); ORTF_SET(ra_over_rb, F.ra_over_rb); ORTF_SET(bcz, F.bcz); ORTF_SET(ra2, F.ra2);
    ORTF_SET(lens_r2_s, K.lens_r2_s); ORTF_SET(l2_radius2_lim, F.l2_radius2 * 0.999996f); ORTF_SET(l2_fb, F.l2_fb);
    ORTF_SET(l2_flat_z, F.l2_flat_z); ORTF_SET(ed_a, K.ed_a); ORTF_SET(ed_b, K.ed_b); ORTF_SET(ep_flat, K.ep_flat); ORTF_SET(ep_max, K.ep_max);
    ORTF_SET(in_neg_eta2, -F.l2_in.eta2); ORTF_SET(in_eta, F.l2_in.eta);
    ORTF_SET(fl_s2_a, K.flat.s2_a); ORTF_SET(fl_s2_b, K.flat.s2_b); ORTF_SET(fl_f_a, K.flat.f_a); ORTF_SET(fl_f_b, K.flat.f_b);
    ORTF_SET(fl_d_a, K.flat.d_a); ORTF_SET(fl_d_b, K.flat.d_b);
    ORTF_SET(l2_cz, F.l2_cz); ORTF_SET(l2_R2, F.l2_R2); ORTF_SET(l2_invR, F.l2_invR); ORTF_SET(l2_neg_invR, -F.l2_invR);
    ORTF_SET(s2_h_d, K.s2.h_d); ORTF_SET(s2_h_0, K.s2.h_0); ORTF_SET(s2_c_0, K.s2.c_0); ORTF_SET(s2_d_h, K.s2.d_h); ORTF_SET(s2_d_0, K.s2.d_0);
    ORTF_SET(s2_p_0, K.s2.p_0); ORTF_SET(s2_n_p, K.s2.n_p); ORTF_SET(s2_n_0, K.s2.n_0);
    ORTF_SET(out_neg_eta2, -F.l2_out.eta2); ORTF_SET(out_eta, F.l2_out.eta);
    ORTF_SET(cv_ni_0, K.curved.ni_0); ORTF_SET(cv_ct2_a, K.curved.ct2_a); ORTF_SET(cv_ct2_b, K.curved.ct2_b); ORTF_SET(cv_cs_0, K.curved.cs_0);
    ORTF_SET(cv_f_a, K.curved.f_a); ORTF_SET(cv_f_b, K.curved.f_b); ORTF_SET(cv_k_a, K.curved.k_a); ORTF_SET(cv_k_0, K.curved.k_0);
    ORTF_SET(cv_d_d, K.curved.d_d); ORTF_SET(cv_d_n, K.curved.d_n); ORTF_SET(cv_d_a, K.curved.d_a); ORTF_SET(cv_d_0, K.curved.d_0);
    ORTF_SET(iris1_z, F.l3_iris1_z); ORTF_SET(iris_z0, K.iris_z0); ORTF_SET(iris_r2, F.l3_iris_r2); ORTF_SET(iris_inv, K.iris_inv);
    ORTF_SET(iris_r, K.iris_r); ORTF_SET(iris_0, K.iris_0);
    ORTF_SET(l3_c1z, F.l3_c1z); ORTF_SET(l3_R1_2, F.l3_R1_2);
    ORTF_SET(s3_h_d, K.s3.h_d); ORTF_SET(s3_h_p, K.s3.h_p); ORTF_SET(s3_h_0, K.s3.h_0); ORTF_SET(s3_c_p, K.s3.c_p); ORTF_SET(s3_c_0, K.s3.c_0);
    ORTF_SET(s3_d_h, K.s3.d_h); ORTF_SET(s3_d_0, K.s3.d_0); ORTF_SET(s3_p_0, K.s3.p_0);
    ORTF_SET(l3_radius2, F.l3_radius2); ORTF_SET(ap_inv, K.ap_inv); ORTF_SET(ap_r, K.ap_r); ORTF_SET(ap_0, K.ap_0);
#undef ORTF_SET
    q.ellipse = F.ellipse;
    q.iris_before = iris_before;
}

/* what the bound test in tests/test_filter_bound.py reads back (host harness only) */
struct OrtFilterTrace {
    int n;
    struct Rec {
        int tag;      /* ORTF_T_* + 100 * (which surface) */
        int valid;    /* no guard had tripped when the record was made */
        float v[3];   /* the fp32 value (scalars in v[0]) */
        float bound;  /* the bound the filter holds for it */
    } rec[96];
};
enum { ORTF_T_POS = 1, ORTF_T_DIR, ORTF_T_NORMAL, ORTF_T_NI, ORTF_T_S2, ORTF_T_CT2, ORTF_T_COST, ORTF_T_F,
       ORTF_T_H, ORTF_T_C, ORTF_T_DISC, ORTF_T_T, ORTF_T_RHO2 };
template <typename P>
ORT_HD void ortf_trace(OrtFilterTrace* tr, int tag, typename P::M unc, typename P::V a, typename P::V b, typename P::V c,
                       typename P::V bound) {
#ifndef __CUDA_ARCH__
    if (P::kTrace && tr && tr->n < 96) {
        OrtFilterTrace::Rec& r = tr->rec[tr->n++];
        r.tag = tag; r.valid = P::first(unc) ? 0 : 1;
        r.v[0] = P::first(a); r.v[1] = P::first(b); r.v[2] = P::first(c); r.bound = P::first(bound);
    }
#else
    (void)tr; (void)tag; (void)unc; (void)a; (void)b; (void)c; (void)bound;
#endif
}
/* a guard's outcome, recorded under its own tag */
template <typename P>
ORT_HD void ortf_trace_guard(OrtFilterTrace* tr, int tag, typename P::M tripped) {
    ortf_trace<P>(tr, tag, tripped, P::lit(0.f), P::lit(0.f), P::lit(0.f), P::lit(0.f));
}

template <typename P>
struct OrtfRay {
    typename P::V px, py, pz, dx, dy, dz;
};

/* The helpers do not branch on a decision they cannot prove: they OR it into `unc` and carry on
 * with what fp32 says; the caller closes a ray at its first definite end and answers 0 (ask fp64)
 * when anything before it was unproven.  One exit per possible end (taken when every ray of the lane has
 * ended), everything else straight-line.  A ray that has ended (`done`) stays in the arithmetic of its
 * lane: what is computed for it afterwards, `unc` included, is never looked at again.
 * Guards (a ray that trips one goes to fp64):
 *   G1  ed + en < 2^-7 at an interface              (second-order terms stay inside the 1.01 / 1.1 factors)
 *   G2  ep < ep_max = (smallest radius) / 64 at every surface
 *   G4  bound(q) < |q| / 256 before a quotient c / q  (keeps |t*| <= 1.01 |t~|)
 *   G5  the aim-disc draw is >= 2^-16 (its low word, unseen here, is then < 2^-16 of it)
 *   G6  cos theta_i > 32 bound(N.I),  G7  bound(cos theta_t) < cos theta_t / 16   (refraction Jacobian) */

/* Ray-sphere intersection, ort_hit_sphere / ort_pick_root_unit; the centre is on the axis (ort_make_filter
 * checks).  Returns the mask of rays that HIT; the outcome hangs on the signs of disc, h and c.
 * In: ep, ed.  Out: *t and *et >= |t~ - t*|.  Bound constants: (h_d, h_p, h_0, c_p, c_0, d_h, d_0); from L2's
 * flat face (FROM_FLAT) the start point's bound ep_flat is inside h_0, c_0, d_0 and h_p, c_p are not used. */
template <typename P, bool FROM_FLAT>
ORT_HD typename P::M ortf_hit_sphere(const OrtfRay<P>& r, typename P::V cz, typename P::V R2, typename P::V h_d,
                                     typename P::V h_p, typename P::V h_0, typename P::V c_p, typename P::V c_0,
                                     typename P::V d_h, typename P::V d_0, typename P::V ep, typename P::V ed,
                                     typename P::V* t, typename P::V* et, typename P::M& unc, OrtFilterTrace* tr, int surf) {
    typedef typename P::V V;
    typedef typename P::M M;
    V lx = r.px, ly = r.py, lz = P::sub(r.pz, cz);
    V h = P::fma(r.dx, lx, P::fma(r.dy, ly, P::mul(r.dz, lz)));
    V l2 = P::fma(lx, lx, P::fma(ly, ly, P::mul(lz, lz)));
    V c = P::sub(l2, R2);
    V disc = P::fma(h, h, P::neg(c));
    /* (R1, R2) with |l| <= L, the scene's bound on the distance to this centre:
     *   eh    = 1.01 L ed + 1.01 ep + (rounding)          h = dir . l
     *   ec    = 2.02 L ep + (rounding)                    c = l . l - R^2
     *   edisc = 2.04 L eh + ec + (rounding)               disc = h^2 - c        (held times 1.01) */
    V eh = FROM_FLAT ? P::fma(h_d, ed, h_0) : P::fma(h_d, ed, P::fma(h_p, ep, h_0));
    V ec = FROM_FLAT ? c_0 : P::fma(c_p, ep, c_0);
    V edisc = FROM_FLAT ? P::fma(d_h, eh, d_0) : P::fma(d_h, eh, P::fma(P::lit(1.01f), ec, d_0));
    ortf_trace<P>(tr, surf + ORTF_T_H, unc, h, P::lit(0.f), P::lit(0.f), eh);
    ortf_trace<P>(tr, surf + ORTF_T_C, unc, c, P::lit(0.f), P::lit(0.f), ec);
    ortf_trace<P>(tr, surf + ORTF_T_DISC, unc, disc, P::lit(0.f), P::lit(0.f), edisc);
    /* each test also catches a NaN */
    unc = P::mor(unc, P::mor(P::abs_ngt(disc, edisc), P::mor(P::abs_ngt(h, eh), P::abs_ngt(c, ec))));
    ortf_trace_guard<P>(tr, surf + 20, P::abs_ngt(disc, edisc));
    ortf_trace_guard<P>(tr, surf + 21, P::abs_ngt(h, eh));
    ortf_trace_guard<P>(tr, surf + 22, P::abs_ngt(c, ec));
    const V zero = P::lit(0.0f);
    M hpos = P::gt(h, zero);
    M miss = P::mor(P::lt(disc, zero), P::mand(hpos, P::gt(c, zero)));
    /* (R3, R4) sq = disc * rsqrt(disc):  esq = 1.01 edisc / sq + 1.01 (E_RSQ + u) sq */
    V isq = P::rsqrt(disc);
    V sq = P::mul(disc, isq);
    V esq = P::fma(edisc, isq, P::mul(P::lit(1.01f * (ORTF_E_RSQ + ORTF_U)), sq));
    /* q = -(h + sgn(h) sq): |q| = |h| + sq, no cancellation.  eq = eh + esq + u |q| */
    V q = P::sel_na(hpos, P::add(h, sq), P::sub(sq, h));
    V aq = P::abs(q);
    V eq = P::fma(P::lit(ORTF_U), aq, P::add(eh, esq));
    /* G4 -- for a ray that hits: a miss has ended, and what is computed past this point means nothing for it */
    M g4 = P::mand(P::mnot(miss), P::nlt(eq, P::mul(P::lit(0.00390625f), aq)));
    unc = P::mor(unc, g4);
    ortf_trace_guard<P>(tr, surf + 23, g4);
    /* the reference's root: q itself when the ray starts inside with the centre ahead, else c / q.
     * (R3, R4) quotient: et = 1.02 (ec + 1.02 |t| eq) / |q| + 1.01 (E_RCP + u) |t| */
    V rq = P::rcp(q);
    V tq = P::mul(c, rq);
    M use_q = P::mand(P::mnot(hpos), P::lt(c, zero));
    V atq = P::abs(tq);
    V etq = P::fma(P::mul(P::lit(1.02f), P::abs(rq)), P::fma(P::mul(P::lit(1.02f), atq), eq, ec),
                   P::mul(P::lit(1.01f * (ORTF_E_RCP + ORTF_U)), atq));
    *t = P::sel(use_q, q, tq);
    *et = P::sel(use_q, eq, etq);
    ortf_trace<P>(tr, surf + ORTF_T_T, P::mor(unc, miss), *t, P::lit(0.f), P::lit(0.f), *et);
    return P::mnot(miss);
}

/* L2's flat face (ort_interface with the normal (0,0,-1) and eta < 1; ort_make_filter checks both):
 * N.I = -dz exactly, no total reflection, T = (eta dx, eta dy, cos theta_t).  A reflected ray goes on
 * either way (the reference does not test that flag, quirk 1).  uw: the decision word as a float. */
template <typename P>
ORT_HD void ortf_flat_face(OrtfRay<P>& r, const OrtfParamsT<typename P::V>& k, typename P::V uw, typename P::V& ed,
                           typename P::M& unc, OrtFilterTrace* tr, int surf) {
    typedef typename P::V V;
    typedef typename P::M M;
    const V one = P::lit(1.0f);
    V costt = r.dz; /* > 0: the aim plane lies beyond the bottle (Dmin > 0) */
    V s2 = P::fma(P::neg(costt), costt, one);
    V ct2 = P::fma(k.in_neg_eta2, s2, one);
    /* es2 = ed (2 |dz| + ed) + u <= 2.02 ed + u;  s2 > es2: not the reference's special case at EXACTLY
     * normal incidence.  ct2 >= 1 - eta^2 > 0: no decision hangs on it. */
    V es2 = P::fma(k.fl_s2_a, ed, k.fl_s2_b);
    ortf_trace<P>(tr, surf + ORTF_T_NI, unc, P::neg(costt), P::lit(0.f), P::lit(0.f), ed);
    ortf_trace<P>(tr, surf + ORTF_T_S2, unc, s2, P::lit(0.f), P::lit(0.f), es2);
    M g1 = P::nlt(ed, P::lit(0.0078125f)), gs = P::ngt(s2, es2);
    unc = P::mor(unc, P::mor(g1, P::mor(gs, P::ngt(costt, P::lit(0.0f))))); /* G1 */
    ortf_trace_guard<P>(tr, surf + 24, g1);
    ortf_trace_guard<P>(tr, surf + 25, gs);
    V cost2 = P::sqrt(ct2); /* E_SQRT <= E_RSQ + u, what d_b and f_b allow for it */
    V ec = P::mul(k.in_eta, costt), e2 = P::mul(k.in_eta, cost2);
    V A = P::sub(ec, cost2), B = P::add(ec, cost2), C = P::sub(e2, costt), D = P::add(e2, costt);
    V B2 = P::mul(B, B), D2 = P::mul(D, D), den = P::mul(B2, D2);
    V num = P::fma(P::mul(A, A), D2, P::mul(P::mul(C, C), B2));
    V lhs = P::mul(P::mul(uw, den), P::lit(4.656612873077393e-10f)); /* 2 u den, u = uw 2^-32 */
    /* lhs - num = 2 den (u - R(cos_i)):  |dR / dcos_i| = 2 eta (1 - eta^2) / cos_t |A / B^3 - C / D^3|, bounded
     * over the whole face by a scene constant, so  ef = den (f_a ed + f_b)  (f_b: the draw and the rounding) */
    V ef = P::mul(den, P::fma(k.fl_f_a, ed, k.fl_f_b));
    V F = P::sub(lhs, num);
    ortf_trace<P>(tr, surf + ORTF_T_F, unc, F, P::lit(0.f), P::lit(0.f), ef);
    M gf = P::abs_ngt(F, ef);
    unc = P::mor(unc, gf);
    ortf_trace_guard<P>(tr, surf + 28, gf);
    M reflect = P::ngt(lhs, num);
    /* refract: |dT| <= eta max(1, eta cos_i / cos_t) |dI| = eta |dI|  ->  ed' = 1.1 eta ed + (evaluation);
     * reflect: (dx, dy, -dz), ed' = ed */
    V ed_refr = P::fma(k.fl_d_a, ed, k.fl_d_b);
    r.dx = P::sel(reflect, r.dx, P::mul(k.in_eta, r.dx));
    r.dy = P::sel(reflect, r.dy, P::mul(k.in_eta, r.dy));
    r.dz = P::sel_na(reflect, costt, cost2);
    ed = P::sel(reflect, ed, ed_refr);
    ortf_trace<P>(tr, surf + ORTF_T_DIR, unc, r.dx, r.dy, r.dz, ed);
}

/* L2's curved face, out of the denser medium (ort_interface with eta > 1): returns the mask of rays that
 * are reflected (they end; their direction is then meaningless).  In: en (normal), ed (direction); out: ed
 * of the refracted direction. */
template <typename P>
ORT_HD typename P::M ortf_exit_face(OrtfRay<P>& r, typename P::V nx, typename P::V ny, typename P::V nz,
                                    const OrtfParamsT<typename P::V>& k, typename P::V uw, typename P::V en, typename P::V& ed,
                                    typename P::M& unc, OrtFilterTrace* tr, int surf) {
    typedef typename P::V V;
    typedef typename P::M M;
    const V one = P::lit(1.0f), zero = P::lit(0.0f);
    V c = P::fma(nx, r.dx, P::fma(ny, r.dy, P::mul(nz, r.dz)));
    V costt = P::abs(c);
    V s2 = P::fma(P::neg(costt), costt, one);
    V ct2 = P::fma(k.out_neg_eta2, s2, one);
    /* (R1, R2) eni = 1.02 (ed + en) + 4u;  es2 = 2.02 eni + u;  ect2 = eta^2 (1.01 es2 + u) + 1.01 u eta^2 */
    V edn = P::add(ed, en);
    V eni = P::fma(P::lit(1.02f), edn, k.cv_ni_0);
    V es2 = P::fma(P::lit(2.02f), eni, P::lit(ORTF_U));
    V ect2 = P::fma(k.cv_ct2_a, eni, k.cv_ct2_b);
    ortf_trace<P>(tr, surf + ORTF_T_NI, unc, c, P::lit(0.f), P::lit(0.f), eni);
    ortf_trace<P>(tr, surf + ORTF_T_S2, unc, s2, P::lit(0.f), P::lit(0.f), es2);
    ortf_trace<P>(tr, surf + ORTF_T_CT2, unc, ct2, P::lit(0.f), P::lit(0.f), ect2);
    /* G1;  s2 > 0 for certain;  the sign of ct2 decides total reflection */
    M g1 = P::nlt(edn, P::lit(0.0078125f)), gs = P::ngt(s2, es2), gc = P::abs_ngt(ct2, ect2);
    unc = P::mor(unc, P::mor(g1, P::mor(gs, gc)));
    ortf_trace_guard<P>(tr, surf + 24, g1);
    ortf_trace_guard<P>(tr, surf + 25, gs);
    ortf_trace_guard<P>(tr, surf + 26, gc);
    M tir = P::ngt(ct2, zero); /* total reflection */
    if (P::all(tir)) return tir;
    /* (R3, R4) cos theta_t = ct2 * rsqrt(ct2):  ecs = ect2 / cos + 1.01 (E_RSQ + u) cos.  Close to the
     * critical angle 1 / cos is large and everything downstream becomes unprovable by itself. */
    V ict = P::rsqrt(ct2);
    V cost2 = P::mul(ct2, ict);
    V ecs = P::fma(ect2, ict, P::mul(k.cv_cs_0, cost2));
    ortf_trace<P>(tr, surf + ORTF_T_COST, P::mor(unc, tir), cost2, P::lit(0.f), P::lit(0.f), ecs);
    V ec = P::mul(k.out_eta, costt), e2 = P::mul(k.out_eta, cost2);
    V A = P::sub(ec, cost2), B = P::add(ec, cost2), C = P::sub(e2, costt), D = P::add(e2, costt);
    V B2 = P::mul(B, B), D2 = P::mul(D, D), den = P::mul(B2, D2);
    V num = P::fma(P::mul(A, A), D2, P::mul(P::mul(C, C), B2));
    V lhs = P::mul(P::mul(uw, den), P::lit(4.656612873077393e-10f)); /* 2 u den, u = uw 2^-32 */
    /* lhs - num = 2 den (u - R(cos_i)),  |dR / dcos_i| <= 2 eta |1 - eta^2| / cos_t (1 / B^2 + 1 / D^2), and B, D
     * are bounded below without total reflection:  ef = den (f_a eni / cos_t + f_b);  G6, G7 keep cos_i, cos_t
     * within the 1.07 inside f_a (and d_d, d_n below) over the segment between fp32 and exact arguments */
    V ef = P::mul(den, P::fma(P::mul(k.cv_f_a, ict), eni, k.cv_f_b));
    V F = P::sub(lhs, num);
    ortf_trace<P>(tr, surf + ORTF_T_F, P::mor(unc, tir), F, P::lit(0.f), P::lit(0.f), ef);
    M gf = P::abs_ngt(F, ef), g6 = P::ngt(costt, P::mul(P::lit(32.0f), eni)), g7 = P::nlt(ecs, P::mul(P::lit(0.0625f), cost2));
    /* a ray in total reflection has ended: what follows is not computed on anything meaningful for it */
    unc = P::mor(unc, P::mand(P::mnot(tir), P::mor(gf, P::mor(g6, g7))));
    ortf_trace_guard<P>(tr, surf + 28, P::mand(P::mnot(tir), gf));
    ortf_trace_guard<P>(tr, surf + 30, P::mand(P::mnot(tir), g6));
    ortf_trace_guard<P>(tr, surf + 31, P::mand(P::mnot(tir), g7));
    M reflected = P::mor(tir, P::ngt(lhs, num));
    if (P::all(reflected)) return reflected;
    /* refract, T = eta I + k N', k = eta c1 - c2 (N' opposing I):
     *     dT = eta [dI_perp + (eta c1 / c2)(N'.dI) N'] + k dN + eta (k / c2)(dN.I) N'
     * so |dT| <= eta (eta c1 / c2) |dI| + sqrt2 max(1, |k|) (|k| / c2) |dN|  (eta c1 >= c2 when eta > 1), with
     * |k| known to ek = k_a eni + ecs + k_0; the rounding of the evaluation adds d_a / c2 + d_0 */
    V m = P::mul(ec, ict);
    V gk = P::mul(P::add(P::abs(A), P::fma(k.cv_k_a, eni, P::add(ecs, k.cv_k_0))), ict);
    V ed_new = P::fma(m, P::mul(k.cv_d_d, ed), P::fma(gk, P::mul(k.cv_d_n, en), P::fma(k.cv_d_a, ict, k.cv_d_0)));
    V kk = P::sel_na(P::nlt(c, zero), A, A); /* (c < 0) ? A : -A */
    r.dx = P::fma(k.out_eta, r.dx, P::mul(kk, nx));
    r.dy = P::fma(k.out_eta, r.dy, P::mul(kk, ny));
    r.dz = P::fma(k.out_eta, r.dz, P::mul(kk, nz));
    ed = ed_new;
    ortf_trace<P>(tr, surf + ORTF_T_DIR, P::mor(unc, reflected), r.dx, r.dy, r.dz, ed);
    return reflected;
}

/* aperture / iris test rho^2 > r^2 on a point known to ep:
 * |rho~^2 - rho*^2| <= ep (2 rho + ep) <= 1.02 ep (rho^2 / r + r)  (2 rho <= rho^2 / r + r, ep <= ep_max <= r / 64) */
template <typename P>
ORT_HD typename P::M ortf_outside(typename P::V x, typename P::V y, typename P::V r2, typename P::V inv_r, typename P::V rr,
                                  typename P::V e0, typename P::V ep, typename P::M& unc, OrtFilterTrace* tr, int surf) {
    typedef typename P::V V;
    V rho2 = P::fma(x, x, P::mul(y, y));
    V e = P::fma(ep, P::fma(rho2, inv_r, rr), P::fma(P::lit(4.0f * ORTF_U), rho2, e0));
    ortf_trace<P>(tr, surf + ORTF_T_RHO2, unc, rho2, P::lit(0.f), P::lit(0.f), e);
    unc = P::mor(unc, P::abs_ngt(P::sub(rho2, r2), e));
    return P::gt(rho2, r2);
}

/* The filter proper.  Words of the ray's own blocks: w1 (high word of the annulus r^2 draw), w2 (annulus
 * angle), w3 (L2's flat-face decision) of block 0; wa (aim angle), wc (L2's curved-face decision) of block 1;
 * h2: the high word of the aim-disc r^2 draw, which stage A has tested.  Returns the status per ray:
 * s > 0 proven, 0 = ask fp64. */
template <typename P>
ORT_HD typename P::S ortf_filter(const OrtfParamsT<typename P::V>& k, typename P::W w1, typename P::W w2, typename P::W w3,
                                 typename P::W wa, typename P::W wc, typename P::W h2, OrtFilterTrace* tr = nullptr) {
    typedef typename P::V V;
    typedef typename P::M M;
    typename P::S st = P::snone();
    /* ring source, ort_source_ring_u.  Its position error is a scene constant (inside k.ed_a) */
    OrtfRay<P> r;
    V s, c;
    V rr = P::sqrt(P::fma(P::word(w1), k.r2m_s, k.r1)); /* r1 + u0 (r2 - r1) */
    P::sincos_word(w2, &s, &c);
    V sx = P::mul(rr, c), sy = P::mul(rr, s);
    V q = k.ellipse ? P::mul(sy, k.ra_over_rb) : sy;
    V sz = P::add(k.bcz, P::sqrt(P::fma(P::neg(q), q, k.ra2)));
    V aim2 = P::mul(P::word(h2), k.lens_r2_s); /* u2 lens_r2 */
    /* G5;  and L2's aperture: stage A decided it on the exact draw, but the fp64 path re-tests the
     * computed aim point (ort_l2_enter), which can differ within its rounding of the edge -- those
     * rays are fp64's */
    M unc = P::mor(P::wlt(h2, 65536u), P::nlt(aim2, k.l2_radius2_lim));
    V rl = P::sqrt(aim2); /* E_SQRT <= E_RSQ + 2.5u, what E_rl allows */
    P::sincos_word(wa, &s, &c);
    V ax = P::mul(rl, c), ay = P::mul(rl, s);
    V ex = P::sub(ax, sx), ey = P::sub(ay, sy), ez = P::sub(k.l2_fb, sz);
    V inv = P::rsqrt(P::fma(ex, ex, P::fma(ey, ey, P::mul(ez, ez))));
    r.dx = P::mul(ex, inv);
    r.dy = P::mul(ey, inv);
    r.dz = P::mul(ez, inv);
    /* |e~ - e*| <= EE (scene constant);  normalising: ed = 2.02 EE / |e~| + 1.01 (5u + E_RSQ) */
    V ed = P::fma(k.ed_a, inv, k.ed_b);
    /* the flat face lies in the aim plane (ring_shortcut): the ray meets it at the aim point */
    r.px = ax;
    r.py = ay;
    r.pz = k.l2_flat_z;
    ortf_trace<P>(tr, 0 + ORTF_T_POS, unc, r.px, r.py, r.pz, k.ep_flat);
    ortf_trace<P>(tr, 0 + ORTF_T_DIR, unc, r.dx, r.dy, r.dz, ed);
    /* L2, ort_l2_body; a reflection at the flat face is not tested by the reference: the ray goes on */
    ortf_flat_face<P>(r, k, P::word(w3), ed, unc, tr, 100);
    V t, et;
    const V none = P::lit(0.0f);
    M hit = ortf_hit_sphere<P, true>(r, k.l2_cz, k.l2_R2, k.s2_h_d, none, k.s2_h_0, none, k.s2_c_0, k.s2_d_h, k.s2_d_0, none,
                                     ed, &t, &et, unc, tr, 200);
    M done = P::mnot(hit);
    P::sset(st, P::mand(done, P::mnot(unc)), ORT_ST_L2_SPHERE_MISS);
    if (P::all(done)) return st;
    r.px = P::fma(r.dx, t, r.px);
    r.py = P::fma(r.dy, t, r.py);
    r.pz = P::fma(r.dz, t, r.pz);
    /* (R2) the hit point: ep' = ep + |t| ed + 1.02 et + rounding (|t*| <= |t~| + et, |dir~| <= 1.01);  G2 */
    V ep = P::fma(P::abs(t), ed, P::fma(P::lit(1.02f), et, k.s2_p_0));
    V en = P::fma(k.s2_n_p, ep, k.s2_n_0);
    V nx = P::mul(r.px, k.l2_neg_invR), ny = P::mul(r.py, k.l2_neg_invR), nz = P::mul(P::sub(k.l2_cz, r.pz), k.l2_invR);
    ortf_trace<P>(tr, 200 + ORTF_T_POS, P::mor(unc, done), r.px, r.py, r.pz, ep);
    ortf_trace<P>(tr, 200 + ORTF_T_NORMAL, P::mor(unc, done), nx, ny, nz, en);
    unc = P::mor(unc, P::nlt(ep, k.ep_max));
    {
        M refl = ortf_exit_face<P>(r, nx, ny, nz, k, P::word(wc), en, ed, unc, tr, 300);
        M now = P::mand(refl, P::mnot(done));
        P::sset(st, P::mand(now, P::mnot(unc)), ORT_ST_L2_CURVED_REFLECT);
        done = P::mor(done, refl);
        if (P::all(done)) return st;
    }
    /* L3 up to its aperture, ort_l3_enter */
    if (k.iris_before) {
        /* ti = (z - pz) / dz:  (R3, R4) eti = 1.02 (ep + u |z| + 1.02 |ti| ed) / |dz| + 1.01 (E_RCP + 2u) |ti|;
         * G4 as for the sphere; the point in the iris plane is known to ep + |ti| ed + 1.01 eti (its own
         * rounding is inside the 4u rho^2 term) */
        V adz = P::abs(r.dz);
        unc = P::mor(unc, P::nlt(ed, P::mul(P::lit(0.00390625f), adz)));
        V rdz = P::rcp(r.dz);
        V ti = P::mul(P::sub(k.iris1_z, r.pz), rdz);
        V ati = P::abs(ti);
        V eti = P::fma(P::mul(P::lit(1.02f), P::abs(rdz)), P::fma(P::mul(P::lit(1.02f), ati), ed, P::add(ep, k.iris_z0)),
                       P::mul(P::lit(1.01f * (ORTF_E_RCP + 2.0f * ORTF_U)), ati));
        V x = P::fma(r.dx, ti, r.px), y = P::fma(r.dy, ti, r.py);
        V epi = P::fma(ati, ed, P::fma(P::lit(1.01f), eti, ep));
        unc = P::mor(unc, P::nlt(epi, k.ep_max));
        M out = ortf_outside<P>(x, y, k.iris_r2, k.iris_inv, k.iris_r, k.iris_0, epi, unc, tr, 400);
        M now = P::mand(out, P::mnot(done));
        P::sset(st, P::mand(now, P::mnot(unc)), ORT_ST_L3_IRIS_BEFORE);
        done = P::mor(done, out);
        if (P::all(done)) return st;
    }
    {
        M hit3 = ortf_hit_sphere<P, false>(r, k.l3_c1z, k.l3_R1_2, k.s3_h_d, k.s3_h_p, k.s3_h_0, k.s3_c_p, k.s3_c_0, k.s3_d_h,
                                           k.s3_d_0, ep, ed, &t, &et, unc, tr, 500);
        M now = P::mand(P::mnot(hit3), P::mnot(done));
        P::sset(st, P::mand(now, P::mnot(unc)), ORT_ST_L3_S1_MISS);
        done = P::mor(done, P::mnot(hit3));
        if (P::all(done)) return st;
    }
    r.px = P::fma(r.dx, t, r.px);
    r.py = P::fma(r.dy, t, r.py);
    r.pz = P::fma(r.dz, t, r.pz);
    ep = P::fma(P::abs(t), ed, P::fma(P::lit(1.02f), et, P::add(ep, k.s3_p_0)));
    unc = P::mor(unc, P::nlt(ep, k.ep_max));
    ortf_trace<P>(tr, 500 + ORTF_T_POS, P::mor(unc, done), r.px, r.py, r.pz, ep);
    M out = ortf_outside<P>(r.px, r.py, k.l3_radius2, k.ap_inv, k.ap_r, k.ap_0, ep, unc, tr, 600);
    P::sset(st, P::mand(P::mand(out, P::mnot(done)), P::mnot(unc)), ORT_ST_L3_APERTURE);
    return st;
}

/* One ray, from its generator state (host harness; h2 as ort_aim_hi gives it) */
ORT_HD int ort_ring_filter(const DevSceneT<float>& F, const DevFilter& K, const DevJob& J, const OrtRng& g, uint32_t h2,
                           OrtFilterTrace* tr = nullptr) {
    uint32_t w[4], v[4];
    ort_block(g, 0u, w);
    ort_block(g, 1u, v);
    OrtfParamsT<float> k;
    ortf_make_params<float>(F, K, J.iris_before, k);
    return ortf_filter<OrtfOne>(k, w[1], w[2], w[3], v[2], v[3], h2, tr);
}

#endif
