/*
 * ort_filter.cuh -- the ring loop's single-precision culling filter, with a running error bound
 * (no counterpart in the reference).
 *
 * 99.5 % of the ring rays that pass L2's aperture still end between L2's curved face and L3's
 * aperture test (total reflection in L2, no intersection with L3's first sphere, outside L3's
 * aperture): they add one to a status counter and nothing else.  WHICH counter is a chain of sign
 * decisions -- discriminants, aperture radii, the reflect-or-refract draw against the Fresnel
 * reflectance.  ort_ring_filter walks source -> L2 -> L3's first surface in fp32 (one issue slot per
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
ORT_HD void ortf_trace(OrtFilterTrace* tr, int tag, bool unc, float a, float b, float c, float bound) {
#ifndef __CUDA_ARCH__
    if (tr && tr->n < 96) {
        OrtFilterTrace::Rec& r = tr->rec[tr->n++];
        r.tag = tag; r.valid = unc ? 0 : 1; r.v[0] = a; r.v[1] = b; r.v[2] = c; r.bound = bound;
    }
#else
    (void)tr; (void)tag; (void)unc; (void)a; (void)b; (void)c; (void)bound;
#endif
}

/* The helpers do not branch on a decision they cannot prove: they OR it into `unc` and carry on
 * with what fp32 says; the caller leaves at the ray's first definite end and answers 0 (ask fp64)
 * when anything before it was unproven.  One exit per possible end, everything else straight-line.
 * Guards (a ray that trips one goes to fp64):
 *   G1  ed + en < 2^-7 at an interface              (second-order terms stay inside the 1.01 / 1.1 factors)
 *   G2  ep < ep_max = (smallest radius) / 64 at every surface
 *   G4  bound(q) < |q| / 256 before a quotient c / q  (keeps |t*| <= 1.01 |t~|)
 *   G5  the aim-disc draw is >= 2^-16 (its low word, unseen here, is then < 2^-16 of it)
 *   G6  cos theta_i > 32 bound(N.I),  G7  bound(cos theta_t) < cos theta_t / 16   (refraction Jacobian) */

/* Ray-sphere intersection, ort_hit_sphere / ort_pick_root_unit: false = miss.  The outcome hangs on
 * the signs of disc, h and c.  In: ep, ed.  Out: *t and *et >= |t~ - t*|. */
template <bool FROM_FLAT> /* the start point is on L2's flat face: its bound ep_flat is inside h_0, c_0, d_0 */
ORT_HD bool ortf_hit_sphere(const OrtRayT<float>& r, float cz, float R2, const DevFilterSphere& k,
                            float ep, float ed, float* t, float* et, bool& unc, OrtFilterTrace* tr, int surf) {
    float lx = r.px, ly = r.py, lz = r.pz - cz; /* the centre is on the axis (ort_make_filter checks) */
    float h = fmaf(r.dx, lx, fmaf(r.dy, ly, r.dz * lz));
    float l2 = fmaf(lx, lx, fmaf(ly, ly, lz * lz));
    float c = l2 - R2;
    float disc = fmaf(h, h, -c);
    /* (R1, R2) with |l| <= L, the scene's bound on the distance to this centre:
     *   eh    = 1.01 L ed + 1.01 ep + (rounding)          h = dir . l
     *   ec    = 2.02 L ep + (rounding)                    c = l . l - R^2
     *   edisc = 2.04 L eh + ec + (rounding)               disc = h^2 - c        (held times 1.01) */
    float eh = FROM_FLAT ? fmaf(k.h_d, ed, k.h_0) : fmaf(k.h_d, ed, fmaf(k.h_p, ep, k.h_0));
    float ec = FROM_FLAT ? k.c_0 : fmaf(k.c_p, ep, k.c_0);
    float edisc = FROM_FLAT ? fmaf(k.d_h, eh, k.d_0) : fmaf(k.d_h, eh, fmaf(1.01f, ec, k.d_0));
    ortf_trace(tr, surf + ORTF_T_H, unc, h, 0.f, 0.f, eh);
    ortf_trace(tr, surf + ORTF_T_C, unc, c, 0.f, 0.f, ec);
    ortf_trace(tr, surf + ORTF_T_DISC, unc, disc, 0.f, 0.f, edisc);
    unc |= !(fabsf(disc) > edisc) || !(fabsf(h) > eh) || !(fabsf(c) > ec); /* each test also catches a NaN */
    ortf_trace(tr, surf + 20, !(fabsf(disc) > edisc), 0.f, 0.f, 0.f, 0.f);
    ortf_trace(tr, surf + 21, !(fabsf(h) > eh), 0.f, 0.f, 0.f, 0.f);
    ortf_trace(tr, surf + 22, !(fabsf(c) > ec), 0.f, 0.f, 0.f, 0.f);
    bool hpos = h > 0.0f;
    if (disc < 0.0f || (hpos && c > 0.0f)) return false;
    /* (R3, R4) sq = disc * rsqrt(disc):  esq = 1.01 edisc / sq + 1.01 (E_RSQ + u) sq */
    float isq = ortf_rsqrt(disc);
    float sq = disc * isq;
    float esq = fmaf(edisc, isq, (1.01f * (ORTF_E_RSQ + ORTF_U)) * sq);
    /* q = -(h + sgn(h) sq): |q| = |h| + sq, no cancellation.  eq = eh + esq + u |q| */
    float q = hpos ? -(h + sq) : (sq - h);
    float aq = fabsf(q);
    float eq = fmaf(ORTF_U, aq, eh + esq);
    unc |= !(eq < 0.00390625f * aq); /* G4 */
    ortf_trace(tr, surf + 23, !(eq < 0.00390625f * aq), 0.f, 0.f, 0.f, 0.f);
    /* the reference's root: q itself when the ray starts inside with the centre ahead, else c / q.
     * (R3, R4) quotient: et = 1.02 (ec + 1.02 |t| eq) / |q| + 1.01 (E_RCP + u) |t| */
    float rq = ortf_rcp(q);
    float tq = c * rq;
    bool use_q = !hpos && c < 0.0f;
    float atq = fabsf(tq);
    float etq = fmaf(1.02f * fabsf(rq), fmaf(1.02f * atq, eq, ec), (1.01f * (ORTF_E_RCP + ORTF_U)) * atq);
    *t = use_q ? q : tq;
    *et = use_q ? eq : etq;
    ortf_trace(tr, surf + ORTF_T_T, unc, *t, 0.f, 0.f, *et);
    return true;
}

/* L2's flat face (ort_interface with the normal (0,0,-1) and eta < 1; ort_make_filter checks both):
 * N.I = -dz exactly, no total reflection, T = (eta dx, eta dy, cos theta_t).  true = reflected
 * (the reference does not test that flag: the ray goes on either way, quirk 1). */
ORT_HD bool ortf_flat_face(OrtRayT<float>& r, const DevIfaceT<float>& f, const DevFilterFlat& k, float uw, float& ed,
                           bool& unc, OrtFilterTrace* tr, int surf) {
    float costt = r.dz; /* > 0: the aim plane lies beyond the bottle (Dmin > 0) */
    float s2 = fmaf(-costt, costt, 1.0f);
    float ct2 = fmaf(-f.eta2, s2, 1.0f);
    /* es2 = ed (2 |dz| + ed) + u <= 2.02 ed + u;  s2 > es2: not the reference's special case at EXACTLY
     * normal incidence.  ct2 >= 1 - eta^2 > 0: no decision hangs on it. */
    float es2 = fmaf(k.s2_a, ed, k.s2_b);
    ortf_trace(tr, surf + ORTF_T_NI, unc, -costt, 0.f, 0.f, ed);
    ortf_trace(tr, surf + ORTF_T_S2, unc, s2, 0.f, 0.f, es2);
    unc |= !(ed < 0.0078125f) || !(s2 > es2) || !(costt > 0.0f); /* G1 */
    ortf_trace(tr, surf + 24, !(ed < 0.0078125f), 0.f, 0.f, 0.f, 0.f);
    ortf_trace(tr, surf + 25, !(s2 > es2), 0.f, 0.f, 0.f, 0.f);
    float cost2 = ortf_sqrt(ct2); /* E_SQRT <= E_RSQ + u, what d_b and f_b allow for it */
    float ec = f.eta * costt, e2 = f.eta * cost2;
    float A = ec - cost2, B = ec + cost2, C = e2 - costt, D = e2 + costt;
    float B2 = B * B, D2 = D * D, den = B2 * D2;
    float num = fmaf(A * A, D2, (C * C) * B2);
    float lhs = (uw * den) * 4.656612873077393e-10f; /* 2 u den, u = uw 2^-32 */
    /* lhs - num = 2 den (u - R(cos_i)):  |dR / dcos_i| = 2 eta (1 - eta^2) / cos_t |A / B^3 - C / D^3|, bounded
     * over the whole face by a scene constant, so  ef = den (f_a ed + f_b)  (f_b: the draw and the rounding) */
    float ef = den * fmaf(k.f_a, ed, k.f_b);
    ortf_trace(tr, surf + ORTF_T_F, unc, lhs - num, 0.f, 0.f, ef);
    unc |= !(fabsf(lhs - num) > ef);
    ortf_trace(tr, surf + 28, !(fabsf(lhs - num) > ef), 0.f, 0.f, 0.f, 0.f);
    bool reflect = !(lhs > num);
    /* refract: |dT| <= eta max(1, eta cos_i / cos_t) |dI| = eta |dI|  ->  ed' = 1.1 eta ed + (evaluation);
     * reflect: (dx, dy, -dz), ed' = ed */
    float ed_refr = fmaf(k.d_a, ed, k.d_b);
    r.dx = reflect ? r.dx : f.eta * r.dx;
    r.dy = reflect ? r.dy : f.eta * r.dy;
    r.dz = reflect ? -costt : cost2;
    ed = reflect ? ed : ed_refr;
    ortf_trace(tr, surf + ORTF_T_DIR, unc, r.dx, r.dy, r.dz, ed);
    return reflect;
}

/* A curved interface out of the denser medium, ort_interface with eta > 1: true = reflected, and
 * then the ray ends (its direction is not updated).  In: en (normal), ed (direction); out: ed of the
 * refracted direction. */
ORT_HD bool ortf_exit_face(OrtRayT<float>& r, float nx, float ny, float nz, const DevIfaceT<float>& f,
                           const DevFilterIface& k, float uw, float en, float& ed, bool& unc, OrtFilterTrace* tr,
                           int surf) {
    float c = fmaf(nx, r.dx, fmaf(ny, r.dy, nz * r.dz));
    float costt = fabsf(c);
    float s2 = fmaf(-costt, costt, 1.0f);
    float ct2 = fmaf(-f.eta2, s2, 1.0f);
    /* (R1, R2) eni = 1.02 (ed + en) + 4u;  es2 = 2.02 eni + u;  ect2 = eta^2 (1.01 es2 + u) + 1.01 u eta^2 */
    float edn = ed + en;
    float eni = fmaf(1.02f, edn, k.ni_0);
    float es2 = fmaf(2.02f, eni, ORTF_U);
    float ect2 = fmaf(k.ct2_a, eni, k.ct2_b);
    ortf_trace(tr, surf + ORTF_T_NI, unc, c, 0.f, 0.f, eni);
    ortf_trace(tr, surf + ORTF_T_S2, unc, s2, 0.f, 0.f, es2);
    ortf_trace(tr, surf + ORTF_T_CT2, unc, ct2, 0.f, 0.f, ect2);
    /* G1;  s2 > 0 for certain;  the sign of ct2 decides total reflection */
    unc |= !(edn < 0.0078125f) || !(s2 > es2) || !(fabsf(ct2) > ect2);
    ortf_trace(tr, surf + 24, !(edn < 0.0078125f), 0.f, 0.f, 0.f, 0.f);
    ortf_trace(tr, surf + 25, !(s2 > es2), 0.f, 0.f, 0.f, 0.f);
    ortf_trace(tr, surf + 26, !(fabsf(ct2) > ect2), 0.f, 0.f, 0.f, 0.f);
    if (!(ct2 > 0.0f)) return true; /* total reflection */
    /* (R3, R4) cos theta_t = ct2 * rsqrt(ct2):  ecs = ect2 / cos + 1.01 (E_RSQ + u) cos.  Close to the
     * critical angle 1 / cos is large and everything downstream becomes unprovable by itself. */
    float ict = ortf_rsqrt(ct2);
    float cost2 = ct2 * ict;
    float ecs = fmaf(ect2, ict, k.cs_0 * cost2);
    ortf_trace(tr, surf + ORTF_T_COST, unc, cost2, 0.f, 0.f, ecs);
    float ec = f.eta * costt, e2 = f.eta * cost2;
    float A = ec - cost2, B = ec + cost2, C = e2 - costt, D = e2 + costt;
    float B2 = B * B, D2 = D * D, den = B2 * D2;
    float num = fmaf(A * A, D2, (C * C) * B2);
    float lhs = (uw * den) * 4.656612873077393e-10f; /* 2 u den, u = uw 2^-32 */
    /* lhs - num = 2 den (u - R(cos_i)),  |dR / dcos_i| <= 2 eta |1 - eta^2| / cos_t (1 / B^2 + 1 / D^2), and B, D
     * are bounded below without total reflection:  ef = den (f_a eni / cos_t + f_b);  G6, G7 keep cos_i, cos_t
     * within the 1.07 inside f_a (and d_d, d_n below) over the segment between fp32 and exact arguments */
    float ef = den * fmaf(k.f_a * ict, eni, k.f_b);
    ortf_trace(tr, surf + ORTF_T_F, unc, lhs - num, 0.f, 0.f, ef);
    unc |= !(fabsf(lhs - num) > ef) || !(costt > 32.0f * eni) || !(ecs < 0.0625f * cost2);
    ortf_trace(tr, surf + 28, !(fabsf(lhs - num) > ef), 0.f, 0.f, 0.f, 0.f);
    ortf_trace(tr, surf + 30, !(costt > 32.0f * eni), 0.f, 0.f, 0.f, 0.f);
    ortf_trace(tr, surf + 31, !(ecs < 0.0625f * cost2), 0.f, 0.f, 0.f, 0.f);
    if (!(lhs > num)) return true; /* reflected */
    /* refract, T = eta I + k N', k = eta c1 - c2 (N' opposing I):
     *     dT = eta [dI_perp + (eta c1 / c2)(N'.dI) N'] + k dN + eta (k / c2)(dN.I) N'
     * so |dT| <= eta (eta c1 / c2) |dI| + sqrt2 max(1, |k|) (|k| / c2) |dN|  (eta c1 >= c2 when eta > 1), with
     * |k| known to ek = k_a eni + ecs + k_0; the rounding of the evaluation adds d_a / c2 + d_0 */
    float m = ec * ict;
    float gk = (fabsf(A) + fmaf(k.k_a, eni, ecs + k.k_0)) * ict;
    float ed_new = fmaf(m, k.d_d * ed, fmaf(gk, k.d_n * en, fmaf(k.d_a, ict, k.d_0)));
    float kk = (c < 0.0f) ? A : -A;
    r.dx = fmaf(f.eta, r.dx, kk * nx);
    r.dy = fmaf(f.eta, r.dy, kk * ny);
    r.dz = fmaf(f.eta, r.dz, kk * nz);
    ed = ed_new;
    ortf_trace(tr, surf + ORTF_T_DIR, unc, r.dx, r.dy, r.dz, ed);
    return false;
}

/* aperture / iris test rho^2 > r^2 on a point known to ep:
 * |rho~^2 - rho*^2| <= ep (2 rho + ep) <= 1.02 ep (rho^2 / r + r)  (2 rho <= rho^2 / r + r, ep <= ep_max <= r / 64) */
ORT_HD bool ortf_outside(float x, float y, float r2, float inv_r, float rr, float e0, float ep, bool& unc,
                         OrtFilterTrace* tr, int surf) {
    float rho2 = fmaf(x, x, y * y);
    float e = fmaf(ep, fmaf(rho2, inv_r, rr), fmaf(4.0f * ORTF_U, rho2, e0));
    ortf_trace(tr, surf + ORTF_T_RHO2, unc, rho2, 0.f, 0.f, e);
    unc |= !(fabsf(rho2 - r2) > e);
    return rho2 > r2;
}

/* h2: the high word of the aim-disc r^2 draw, which stage A has tested; the ray's blocks 0 (annulus r^2,
 * annulus angle, L2's flat-face decision) and 1 (aim angle, L2's curved-face decision) are generated here */
ORT_HD int ort_ring_filter(const DevSceneT<float>& F, const DevFilter& K, const DevJob& J, const OrtRng& g, uint32_t h2,
                           OrtFilterTrace* tr = nullptr) {
    uint32_t w[4], v[4];
    ort_block(g, 0u, w);
    ort_block(g, 1u, v);
    const uint32_t w_aim = v[2], w_curved = v[3];
    /* ring source, ort_source_ring_u.  Its position error is a scene constant (inside K.ed_a) */
    OrtRayT<float> r;
    float s, c;
    float rr = ortf_sqrt(fmaf(ortf_word(w[1]), K.r2m_s, F.r1)); /* r1 + u0 (r2 - r1) */
    ortf_sincos_word(w[2], &s, &c);
    r.px = rr * c;
    r.py = rr * s;
    float q = F.ellipse ? r.py * F.ra_over_rb : r.py;
    r.pz = F.bcz + ortf_sqrt(fmaf(-q, q, F.ra2));
    float aim2 = ortf_word(h2) * K.lens_r2_s; /* u2 lens_r2 */
    /* G5;  and L2's aperture: stage A decided it on the exact draw, but the fp64 path re-tests the
     * computed aim point (ort_l2_enter), which can differ within its rounding of the edge -- those
     * rays are fp64's */
    bool unc = h2 < 65536u || !(aim2 < F.l2_radius2 * 0.999996f);
    float rl = ortf_sqrt(aim2); /* E_SQRT <= E_RSQ + 2.5u, what E_rl allows */
    ortf_sincos_word(w_aim, &s, &c);
    float ax = rl * c, ay = rl * s;
    float ex = ax - r.px, ey = ay - r.py, ez = F.l2_fb - r.pz;
    float inv = ortf_rsqrt(fmaf(ex, ex, fmaf(ey, ey, ez * ez)));
    r.dx = ex * inv;
    r.dy = ey * inv;
    r.dz = ez * inv;
    /* |e~ - e*| <= EE (scene constant);  normalising: ed = 2.02 EE / |e~| + 1.01 (5u + E_RSQ) */
    float ed = fmaf(K.ed_a, inv, K.ed_b);
    /* the flat face lies in the aim plane (ring_shortcut): the ray meets it at the aim point */
    r.px = ax;
    r.py = ay;
    r.pz = F.l2_flat_z;
    ortf_trace(tr, 0 + ORTF_T_POS, unc, r.px, r.py, r.pz, K.ep_flat);
    ortf_trace(tr, 0 + ORTF_T_DIR, unc, r.dx, r.dy, r.dz, ed);
    /* L2, ort_l2_body; a reflection at the flat face is not tested by the reference: the ray goes on */
    (void)ortf_flat_face(r, F.l2_in, K.flat, ortf_word(w[3]), ed, unc, tr, 100);
    float t, et;
    if (!ortf_hit_sphere<true>(r, F.l2_cz, F.l2_R2, K.s2, 0.0f, ed, &t, &et, unc, tr, 200))
        return unc ? 0 : ORT_ST_L2_SPHERE_MISS;
    ort_advance(r, t);
    /* (R2) the hit point: ep' = ep + |t| ed + 1.02 et + rounding (|t*| <= |t~| + et, |dir~| <= 1.01);  G2 */
    float ep = fmaf(fabsf(t), ed, fmaf(1.02f, et, K.s2.p_0));
    float en = fmaf(K.s2.n_p, ep, K.s2.n_0);
    float nx = r.px * -F.l2_invR, ny = r.py * -F.l2_invR, nz = (F.l2_cz - r.pz) * F.l2_invR; /* centre on the axis */
    ortf_trace(tr, 200 + ORTF_T_POS, unc, r.px, r.py, r.pz, ep);
    ortf_trace(tr, 200 + ORTF_T_NORMAL, unc, nx, ny, nz, en);
    unc |= !(ep < K.ep_max);
    if (ortf_exit_face(r, nx, ny, nz, F.l2_out, K.curved, ortf_word(w_curved), en, ed, unc, tr, 300))
        return unc ? 0 : ORT_ST_L2_CURVED_REFLECT;
    /* L3 up to its aperture, ort_l3_enter */
    if (J.iris_before) {
        /* ti = (z - pz) / dz:  (R3, R4) eti = 1.02 (ep + u |z| + 1.02 |ti| ed) / |dz| + 1.01 (E_RCP + 2u) |ti|;
         * G4 as for the sphere; the point in the iris plane is known to ep + |ti| ed + 1.01 eti (its own
         * rounding is inside the 4u rho^2 term) */
        float adz = fabsf(r.dz);
        unc |= !(ed < 0.00390625f * adz);
        float rdz = ortf_rcp(r.dz);
        float ti = (F.l3_iris1_z - r.pz) * rdz;
        float ati = fabsf(ti);
        float eti = fmaf(1.02f * fabsf(rdz), fmaf(1.02f * ati, ed, ep + K.iris_z0), (1.01f * (ORTF_E_RCP + 2.0f * ORTF_U)) * ati);
        float x = fmaf(r.dx, ti, r.px), y = fmaf(r.dy, ti, r.py);
        float epi = fmaf(ati, ed, fmaf(1.01f, eti, ep));
        unc |= !(epi < K.ep_max);
        if (ortf_outside(x, y, F.l3_iris_r2, K.iris_inv, K.iris_r, K.iris_0, epi, unc, tr, 400))
            return unc ? 0 : ORT_ST_L3_IRIS_BEFORE;
    }
    if (!ortf_hit_sphere<false>(r, F.l3_c1z, F.l3_R1_2, K.s3, ep, ed, &t, &et, unc, tr, 500))
        return unc ? 0 : ORT_ST_L3_S1_MISS;
    ort_advance(r, t);
    ep = fmaf(fabsf(t), ed, fmaf(1.02f, et, ep + K.s3.p_0));
    unc |= !(ep < K.ep_max);
    ortf_trace(tr, 500 + ORTF_T_POS, unc, r.px, r.py, r.pz, ep);
    bool out = ortf_outside(r.px, r.py, F.l3_radius2, K.ap_inv, K.ap_r, K.ap_0, ep, unc, tr, 600);
    return (!unc && out) ? ORT_ST_L3_APERTURE : 0;
}

#endif /* ORT_FILTER_CUH */
