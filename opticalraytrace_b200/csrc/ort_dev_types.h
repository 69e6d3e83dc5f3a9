/*
 * ort_dev_types.h -- device-side (constant-memory) view of a scene and a job.
 *
 * The public structs of include/ort.h mirror the reference's derived types field by field; the
 * kernels read this flattened form instead, in which everything that is invariant over a launch
 * has been hoisted on the host once (squared radii, refractive-index ratios, plane positions,
 * the cosine of the acceptance angle, 401/diameter ...), so that the per-ray code is only the
 * work that depends on the ray.
 */
#ifndef ORT_DEV_TYPES_H
#define ORT_DEV_TYPES_H

#include <stdint.h>

#include "../../include/ort.h"

#define ORT_MAX_SCENES 4096 /* scenes per ort_trace call; each scene is its own kernel launch and
                               travels as a __grid_constant__ kernel parameter (712 B) */

/* One refracting interface n_a -> n_b */
template <typename R>
struct DevIfaceT {
    R na, nb;  /* refractive indices on the incoming / outgoing side */
    R eta;     /* na / nb  (reference computes n1/n2 per call, src/surfaces.f90:279,352) */
    R eta2;    /* eta * eta */
};

template <typename R>
struct DevSceneT {
    /* --- bottle (reference src/lens.f90:230-350) --- */
    R bcx, bcy, bcz;         /* centre */
    R b_in_r, b_in_r2;       /* inner cylinder radius Ra - th and its square */
    R b_out_r, b_out_r2;     /* outer cylinder radius Ra */
    R b_in_invr, b_out_invr; /* 1/radius: the radial normal of a clear cylindrical wall */
    R b_in_ia2, b_in_ib2;    /* inner ellipse 1/semia^2 (z), 1/semib^2 (y) */
    R b_out_ia2, b_out_ib2;  /* outer ellipse (reference: Ra/2, Rb/2; fixed: Ra, Rb) */
    DevIfaceT<R> b_in;                /* contents -> glass */
    DevIfaceT<R> b_out;               /* glass -> air (1.0) */
    R mutot_c, inv_mutot_c, albedo_c; /* contents: mua+mus, 1/(mua+mus), mus/(mus+mua) */
    R mutot_b, inv_mutot_b, albedo_b; /* wall */
    /* --- sources (reference src/sourceMod.f90:12-47,250-300) --- */
    R cos_theta_max, one_m_ctm, point_offset;
    R r1, r2_m_r1;           /* annulus: r = r1 + u (r2 - r1) */
    R ra2, ra_over_rb;       /* Ra^2, Ra/Rb */
    R lens_r2;               /* (L2.radius + 10e-3)^2 */
    R l2_fb;                 /* z of the aim disc */
    /* --- crs / isors sources (reference src/sourceMod.f90:50-89,162-247) --- */
    R spot_size;             /* crs: sigma of the spot (already rescaled) */
    R crs_r2;                /* crs: (Ra + thickness)^2, the cylinder the spot is projected on */
    R isors_beam;            /* isors: beam width (sigma of the Gaussian on the axicon) */
    R isors_base;            /* isors: (separation + beam) / tan(alpha (n_axicon - 1)) */
    R isors_k, isors_h;      /* isors: axicon (radius/height)^2 and height */
    R isors_z;               /* isors: Ra + bottle z + epsilon(1.) */
    R isors_lens_r2;         /* isors: L2.radius^2 (aim disc) */
    DevIfaceT<R> isors_axicon;        /* 1.4 -> 1.0 */
    /* --- L2 plano-convex (reference src/lens.f90:425-481) --- */
    R l2_cx, l2_cy, l2_cz;   /* sphere centre */
    R l2_flat_z;             /* centre.z + R - thickness */
    R l2_radius2;            /* aperture radius squared */
    R l2_R2, l2_invR;        /* curve radius squared, 1/R */
    R l2_fnx, l2_fny, l2_fnz;/* flat-face normal */
    DevIfaceT<R> l2_in, l2_out;       /* n1 -> n2, n2 -> n1 */
    /* --- L3 achromatic doublet (reference src/lens.f90:531-645) --- */
    R l3_c1x, l3_c1y, l3_c1z, l3_c2x, l3_c2y, l3_c2z, l3_c3x, l3_c3y, l3_c3z;
    R l3_R1_2, l3_R2_2, l3_R3_2, l3_invR1, l3_invR2, l3_invR3;
    R l3_radius2;            /* aperture */
    R l3_iris_r2;            /* (radius * iris_radius)^2 */
    R l3_iris1_z, l3_iris2_z;/* centre1.z - R1, centre3.z + R3 */
    DevIfaceT<R> l3_s1, l3_s2, l3_s3; /* n1->n2, n2->n3, n3->n1 */
    /* --- image plane (reference src/optics_system.f90:48-49, src/imageMod.f90:19-58) --- */
    R img_z;                 /* img_plane + fibre_offset */
    R inv_binwid;            /* 401 / diameter */
    R binwid;                /* diameter / 401 */
    R cos_na2;               /* cos(asin(0.22))^2 */
    int32_t ellipse, scatter_b, scatter_c;
    int32_t ring_shortcut;        /* L2's flat face lies in the ring source's aim plane z = fb (true for
                                     every lens the loaders build): the aperture test reduces to
                                     u * lens_r2 > radius^2 and is taken before anything else */
};
typedef DevIfaceT<double> DevIface;
typedef DevSceneT<double> DevScene;

/* ---- the ring loop's single-precision culling filter: error-bound constants ------------------
 * ort_ring_filter (ort_filter.cuh) carries a bound on the distance between each single-precision
 * quantity and the value exact arithmetic would give, and calls a ray only when every decision
 * value is further from zero than its bound.  Everything in those bounds that does not depend on
 * the ray is computed here, once per launch, in double and rounded up (ort_make_filter,
 * ort_flatten.h).  The derivation is DESIGN.md section 3.1c; units: lengths in metres, directions and
 * normals dimensionless. */
struct DevFilterFlat {    /* L2's flat face: entering the denser medium through a plane whose normal is EXACTLY
                             (0,0,-1): no total reflection, cos theta_t >= sqrt(1 - eta^2), every bound is
                             linear in the bound of the incoming direction */
    float s2_a, s2_b;     /* bound(sin^2 theta_i)  = s2_a * ed + s2_b */
    float f_a, f_b;       /* bound(lhs - num)      = den * (f_a * ed + f_b) */
    float d_a, d_b;       /* refraction: bound(dir') = d_a * ed + d_b */
};
struct DevFilterIface {   /* a curved interface leaving the denser medium (eta > 1): conditioning ~ 1 / cos theta_t */
    float ni_0;           /* bound(N.I)            = 1.02 (ed + en) + ni_0 */
    float ct2_a, ct2_b;   /* bound(cos^2 theta_t)  = ct2_a * bound(N.I) + ct2_b   (already times 1.01) */
    float cs_0;           /* bound(cos theta_t)    = bound(cos^2) / cos + cs_0 * cos */
    float f_a, f_b;       /* bound(lhs - num)      = den * (f_a / cos * bound(N.I) + f_b) */
    float k_a, k_0;       /* bound(k), k = eta cos_i - cos_t:  k_a * bound(N.I) + bound(cos theta_t) + k_0 */
    float d_d, d_n, d_a, d_0; /* refraction: bound(dir') = d_d (eta cos_i / cos_t) ed + d_n (|k| / cos_t) en + d_a / cos_t + d_0 */
};
struct DevFilterSphere {  /* one ray-sphere intersection + the normal at the hit point */
    float h_d, h_p, h_0;  /* bound(h)    = h_d * ed + h_p * ep + h_0 */
    float c_p, c_0;       /* bound(c)    = c_p * ep + c_0 */
    float d_h, d_0;       /* bound(disc) = d_h * bound(h) + 1.01 bound(c) + d_0   (already times 1.01) */
    float p_0;            /* rounding of the advance to the hit point */
    float n_p, n_0;       /* bound(normal) = n_p * ep + n_0 */
};
struct DevFilter {
    float ed_a, ed_b;     /* bound(dir) of the emitted ray = ed_a / |aim - source| + ed_b */
    float ep_flat;        /* bound(pos) on L2's flat face */
    float ep_max;         /* a bound(pos) above this hands the ray to fp64 */
    float ap_inv, ap_r, ap_0;       /* L3 aperture: bound(rho^2) = ep (rho^2 * ap_inv + ap_r) + 4u rho^2 + ap_0 */
    float iris_inv, iris_r, iris_0; /* L3 iris, same form */
    float iris_z0;                  /* rounding of the iris plane distance */
    float r2m_s, lens_r2_s;         /* the fp32 scene's r2_m_r1 and lens_r2 times 2^-32 (they multiply raw 32-bit words) */
    DevFilterFlat flat;
    DevFilterIface curved;
    DevFilterSphere s2, s3; /* s2: from the flat face, where ep = ep_flat is inside h_0, c_0 */
    int32_t usable;       /* 0: a premise fails (see ort_make_filter) or a constant is not finite; 1: valid bounds,
                             too large to be useful; 2: use the filter (the launcher runs all-fp64 otherwise) */
};

struct DevJob {
    uint64_t seed;
    int64_t first_ray;   /* ray index of local ray 0 of this launch */
    int64_t nrays;       /* rays per scene in this launch (< 2^32) */
    double uniform_override;
    int32_t phase, use_bottle, iris_before, iris_after;
    int32_t nscenes, stop_after, flags, source_kind;
    int64_t total_rays;  /* nphotons of the whole job (create_spot) */
    uint32_t round_keys[20];    /* Philox key schedule seed + r * (W0, W1), r = 0..9: launch constants,
                                   so the rounds read them straight from the constant bank */
    uint64_t aim_cut;           /* ring loop: the smallest 64-bit slot-2 draw that fails L2's aperture test
                                   (ort_ring_aim_cut), or 0 when the test is not a test on that draw */
    const long long* image_cdf; /* image source: inclusive prefix sums of the 512x512 ray budget in
                                   emit_image's scan order (device memory), or NULL */
};

#endif
