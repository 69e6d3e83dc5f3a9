/*
 * ort_oracle.cpp -- CPU ORACLE.  TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * A C++17 restatement, statement by statement and in the reference's operation order, of the
 * per-ray path of lewisfish/OpticalRayTrace (Fortran, every `real` is fp64 because the
 * reference is built with -freal-4-real-8, src/Makefile:2).  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this file; the product
 * (opticalraytrace_b200/csrc) never does and has no CPU fallback.
 *
 * PARITY UNPINNED: the reference has no tests or golden vectors and no
 * Fortran compiler exists in this image or on the GPU boxes, so the reference binary could not
 * be run to pin this restatement.  What pins it instead: (1) the physics / closed-form
 * known-answer checks of SURVEY.md section 8(c) (tests/test_oracle_kat.py), (2) an independent
 * second restatement in pure Python written from the Fortran separately
 * (tests/pyref.py, compared ray by ray), (3) Random123's published Philox4x32-10 vectors for
 * the generator.
 *
 * The one deliberate difference from the reference: libgfortran's random_number (xoshiro256**,
 * src/random_mod.f90:39-46) is replaced -- as BASELINE.json's north_star prescribes -- by a
 * counter-based generator (Philox4x32-7) whose every uniform is a pure function of
 * (seed, phase, ray index, draw slot).  Build: see oracle/Makefile (-O2 -ffp-contract=off).
 */
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/ort.h"

namespace {

/* ---------------------------------------------------------------------------------------
 * type(vector), reference src/vector_class.f90:3-31 and its operators :48-186
 * ------------------------------------------------------------------------------------- */
struct vec {
    double x, y, z;
};
inline vec operator-(vec a, vec b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; } /* :48-57 */
inline vec operator+(vec a, vec b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; } /* :84-93 */
inline vec operator*(vec a, double b) { return {a.x * b, a.y * b, a.z * b}; }    /* :139-148 */
inline vec operator*(double a, vec b) { return {a * b.x, a * b.y, a * b.z}; }    /* :151-160 */
inline vec operator/(vec a, double b) { return {a.x / b, a.y / b, a.z / b}; }    /* :163-172 */
inline double dot(vec a, vec b) { return (a.x * b.x) + (a.y * b.y) + (a.z * b.z); } /* :96-106 */
inline vec magnitude(vec a) { /* :175-186 -- despite the name, this normalises */
    double tmp = std::sqrt(a.x * a.x + a.y * a.y + a.z * a.z);
    return a / tmp;
}

inline vec v3(const double* p) { return {p[0], p[1], p[2]}; }

/* reference src/constants.f90:5 -- 4.*atan(1.) folds to the correctly rounded pi */
const double PI = 3.14159265358979323846;
const double TWOPI = 2.0 * 3.14159265358979323846;

/* ---------------------------------------------------------------------------------------
 * Counter-based uniforms: Philox4x32-7 (Salmon et al., SC'11; Random123 v1.14 constants; seven
 * rounds is the variant that paper reports as the fewest that pass BigCrush).
 * Replaces src/random_mod.f90:39-46 (ran2).  counter = (ray_lo, ray_hi, phase, block),
 * key = (seed_lo, seed_hi).  A block of four words w0..w3 serves one WIDE draw
 * u = ((w1:w0) >> 11) * 2^-53 in [0,1) (53 bits like gfortran's random_number: the radial draws)
 * and NARROW draws u = w * 2^-32 (angles, reflect-or-refract decisions).  Slot map, rev 2:
 *    0: 0.w0w1 wide   1: 0.w2   2: 1.w0w1 wide   3: 1.w2   4: 0.w3   5: 1.w3
 *    6..9: 2.w0..w3   10: 3.w0w1 wide  11: 3.w2  12: 3.w3  13: 4.w0w1 wide  14: 4.w2  15: 4.w3
 *    16..: block slot/2, (w1:w0) for even and (w3:w2) for odd slots, both wide
 * In the ring loop (phase 1) the HIGH word of slot 2 is word (ray & 3) of a block four consecutive rays
 * share, counter (ray >> 2, 16 + phase, 0): that word decides L2's aperture for 69 % of the ring rays.
 * ------------------------------------------------------------------------------------- */
const int PHILOX_ROUNDS = 7;
/* rounds [first, first + n) of Philox4x32 (the key of round r is key + r * (W0, W1)) */
inline void philox4x32_rounds(const uint32_t ctr_in[4], const uint32_t key_in[2], int first, int n, uint32_t out[4]) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
    uint32_t c0 = ctr_in[0], c1 = ctr_in[1], c2 = ctr_in[2], c3 = ctr_in[3];
    uint32_t k0 = key_in[0] + (uint32_t)first * W0, k1 = key_in[1] + (uint32_t)first * W1;
    for (int r = 0; r < n; ++r) {
        uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
        uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
        uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
        uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* draw slots (fixed, independent of control-flow history) */
enum {
    SLOT_SRC0 = 0, /* wide.    ring: r^2       | point: cos theta */
    SLOT_SRC1 = 1, /* narrow.  ring: theta     | point: phi */
    SLOT_SRC2 = 2, /* wide.    ring: r^2 lens  | point: bottle inner reflect_refract */
    SLOT_SRC3 = 3, /* narrow.  ring: th lens   | point: bottle outer reflect_refract */
    SLOT_L2_FLAT = 4,
    SLOT_L2_CURVED = 5,
    SLOT_L3_S1 = 6,
    SLOT_L3_S2 = 7,
    SLOT_L3_S3 = 8,
    SLOT_SCATTER0 = 16 /* scatter loops: slots 16,17,... consumed sequentially */
};

struct Draws {
    uint64_t seed, ray;
    uint32_t phase;
    double override_u;
    uint32_t fixed[5][4]; /* blocks 0..4, generated on first use */
    unsigned have_fixed;
    uint32_t cached_block; /* the scatter block in use */
    bool have;
    uint32_t w[4];
    uint32_t scatter_next;
    Draws(uint64_t seed_, uint32_t phase_, uint64_t ray_, double ov)
        : seed(seed_), ray(ray_), phase(phase_), override_u(ov), fixed{}, have_fixed(0), cached_block(0),
          have(false), w{}, scatter_next(SLOT_SCATTER0) {}
    void generate(uint32_t b, uint32_t* out) const {
        uint32_t ctr[4] = {(uint32_t)ray, (uint32_t)(ray >> 32), phase, b};
        uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
        philox4x32_rounds(ctr, key, 0, PHILOX_ROUNDS, out);
    }
    static double wide(uint32_t lo, uint32_t hi) {
        uint64_t bits = ((uint64_t)hi << 32) | lo;
        return (double)(bits >> 11) * (1.0 / 9007199254740992.0);
    }
    static double narrow(uint32_t x) { return (double)x * (1.0 / 4294967296.0); }
    double slot(uint32_t k) {
        if (override_u >= 0.0) return override_u;
        if (k >= 16) {
            uint32_t b = k >> 1;
            if (!have || b != cached_block) {
                generate(b, w);
                cached_block = b;
                have = true;
            }
            return (k & 1) ? wide(w[2], w[3]) : wide(w[0], w[1]);
        }
        static const unsigned char blk[16] = {0, 0, 1, 1, 0, 1, 2, 2, 2, 2, 3, 3, 3, 4, 4, 4};
        static const signed char word[16] = {-1, 2, -1, 2, 3, 3, 0, 1, 2, 3, -1, 2, 3, -1, 2, 3}; /* -1: wide */
        const unsigned b = blk[k];
        if (!(have_fixed & (1u << b))) {
            generate(b, fixed[b]);
            have_fixed |= 1u << b;
        }
        const uint32_t* f = fixed[b];
        if (k == 2 && phase == ORT_PHASE_RING) {
            /* ring loop: the high word of slot 2 is word (ray & 3) of the block rays 4q .. 4q+3 share,
             * counter (ray >> 2, 16 + phase, 0) */
            uint64_t q = ray >> 2;
            uint32_t ctr[4] = {(uint32_t)q, (uint32_t)(q >> 32), 16u + phase, 0u};
            uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
            uint32_t sh[4];
            philox4x32_rounds(ctr, key, 0, PHILOX_ROUNDS, sh);
            return wide(f[0], sh[ray & 3]);
        }
        return word[k] < 0 ? wide(f[0], f[1]) : narrow(f[word[k]]);
    }
    double scatter() { return slot(scatter_next++); }
};

/* ---------------------------------------------------------------------------------------
 * reference src/surfaces.f90
 * ------------------------------------------------------------------------------------- */
/* solveQuadratic, src/surfaces.f90:227-260 */
inline bool solveQuadratic(double a, double b, double c, double& x0, double& x1) {
    double discrim = b * b - 4.0 * a * c;
    if (discrim < 0.0) {
        return false;
    } else if (discrim == 0.0) {
        x0 = -0.5 * b / a;
        x1 = x0;
    } else {
        double q;
        if (b > 0.0) {
            q = -0.5 * (b + std::sqrt(discrim));
        } else {
            q = -0.5 * (b - std::sqrt(discrim));
        }
        x0 = q / a;
        x1 = c / q;
    }
    return true;
}

/* shared tail of the intersect_* functions, e.g. src/surfaces.f90:74-87 */
inline bool pick_root(double a, double b, double c, double& t) {
    double t0, t1;
    if (!solveQuadratic(a, b, c, t0, t1)) return false;
    if (t0 > t1) {
        double tmp = t1;
        t1 = t0;
        t0 = tmp;
    }
    if (t0 < 0.0) {
        t0 = t1;
        if (t0 < 0.0) return false;
    }
    t = t0;
    return true;
}

/* intersect_sphere, src/surfaces.f90:52-89 */
inline bool intersect_sphere(vec orig, vec dir, double& t, vec centre, double radius) {
    vec L = orig - centre;
    double a = dot(dir, dir);
    double b = 2.0 * dot(dir, L);
    double c = dot(L, L) - radius * radius;
    return pick_root(a, b, c, t);
}

/* intersect_cylinder, src/surfaces.f90:91-130 -- axis along x: only y,z enter */
inline bool intersect_cylinder(vec orig, vec dir, double& t, vec centre, double radius) {
    vec L = orig - centre;
    double a = dir.z * dir.z + dir.y * dir.y;
    double b = 2 * (dir.z * L.z + dir.y * L.y);
    double c = L.z * L.z + L.y * L.y - radius * radius;
    return pick_root(a, b, c, t);
}

/* intersect_ellipse, src/surfaces.f90:133-176 -- semia <-> z, semib <-> y */
inline bool intersect_ellipse(vec orig, vec dir, double& t, vec centre, double semia, double semib) {
    double semia2div = 1. / (semia * semia);
    double semib2div = 1. / (semib * semib);
    vec L = orig - centre;
    double a = semia2div * (dir.z * dir.z) + semib2div * (dir.y * dir.y);
    double b = 2 * (semia2div * dir.z * L.z + semib2div * dir.y * L.y);
    double c = semia2div * (L.z * L.z) + semib2div * (L.y * L.y) - 1;
    return pick_root(a, b, c, t);
}

/* intersect_cone, src/surfaces.f90:179-224 */
inline bool intersect_cone(vec orig, vec dir, double& t, vec centre, double radius, double height) {
    double k = radius / height;
    k = k * k;
    vec L = orig - centre;
    double a = dir.x * dir.x + dir.y * dir.y - (k * (dir.z * dir.z));
    double b = 2. * ((dir.x * L.x) + (dir.y * L.y) - (k * dir.z * (L.z - height)));
    double c = L.x * L.x + L.y * L.y - (k * ((L.z - height) * (L.z - height)));
    return pick_root(a, b, c, t);
}

/* fresnel, src/surfaces.f90:336-372 */
inline double fresnel(vec I, vec N, double n1, double n2) {
    double costt = std::fabs(dot(I, N));
    double sintt = std::sqrt(1. - costt * costt);
    double sint2 = n1 / n2 * sintt;
    if (sint2 > 1.) {
        return 1.0;
    } else if (costt == 1.) {
        return 0.;
    } else {
        sint2 = (n1 / n2) * sintt;
        double cost2 = std::sqrt(1. - sint2 * sint2);
        double r1 = std::fabs((n1 * costt - n2 * cost2) / (n1 * costt + n2 * cost2));
        double f1 = r1 * r1;
        double r2 = std::fabs((n1 * cost2 - n2 * costt) / (n1 * cost2 + n2 * costt));
        double f2 = r2 * r2;
        double tir = 0.5 * (f1 + f2);
        if (std::isnan(tir) || tir > 1. || tir < 0.) tir = 1.;
        return tir;
    }
}

/* reflect, src/surfaces.f90:285-300 */
inline void reflect(vec& I, vec N) {
    vec R = I - (2. * dot(N, I)) * N;
    I = R;
}

/* refract, src/surfaces.f90:303-333 */
inline void refract(vec& I, vec N, double eta) {
    vec Ntmp = N;
    double c1 = dot(Ntmp, I);
    if (c1 < 0.) {
        c1 = -c1;
    } else {
        Ntmp = (-1.) * N;
    }
    double c2 = std::sqrt(1.0 - (eta * eta) * (1.0 - c1 * c1));
    vec T = eta * I + (eta * c1 - c2) * Ntmp;
    I = T;
}

/* reflect_refract, src/surfaces.f90:262-282; `u` is the ran2() of :275 */
inline void reflect_refract(vec& I, vec N, double n1, double n2, bool& rflag, double u) {
    rflag = false;
    if (u <= fresnel(I, N, n1, n2)) {
        reflect(I, N);
        rflag = true;
    } else {
        refract(I, N, n1 / n2);
    }
}

/* ---------------------------------------------------------------------------------------
 * Conditioning probe for the scatter path (tests only; off unless orc_set_jitter(seed != 0)).
 * The GPU's libm differs from glibc's by up to 2 ulp per call (log, atan2, sin, cos, acos) and
 * its fused multiply-adds move intersection distances by an ulp; stokes' spherical-triangle
 * update amplifies such differences by up to ~1/sin^2 of the deflection.  With the probe on,
 * every one of those results is moved by a pseudo-random -2..+2 ulp, so that running the oracle
 * against itself measures, ray by ray, how far two correct implementations may be apart.
 * ------------------------------------------------------------------------------------- */
static uint64_t g_jitter_seed = 0;
static thread_local uint64_t t_jitter = 0;
inline void jitter_begin(int64_t ray) { t_jitter = g_jitter_seed ? (g_jitter_seed ^ ((uint64_t)ray * 0x9E3779B97F4A7C15ull)) | 1ull : 0; }
inline double jit(double x) {
    if (!t_jitter || x == 0.0 || !std::isfinite(x)) return x;
    t_jitter = t_jitter * 6364136223846793005ull + 1442695040888963407ull;
    int k = (int)((t_jitter >> 33) % 5u) - 2; /* -2 .. +2 ulp */
    int64_t bits;
    std::memcpy(&bits, &x, sizeof bits);
    bits += (x > 0.0) ? k : -k;
    std::memcpy(&x, &bits, sizeof bits);
    return x;
}

/* tauint, src/surfaces.f90:13-50; returns false where the reference does `error stop` */
inline bool tauint(vec pos, vec dir, double mua, double mus, vec centre, double radius, double& dist,
                   bool& tflag, double u) {
    double mu_tot = mua + mus;
    double tau = -jit(std::log(u));
    tflag = false;
    bool flag = intersect_cylinder(pos, dir, dist, centre, radius);
    if (!flag) return false;
    dist = jit(dist);
    double tauradius = dist * mu_tot;
    if (tau < tauradius) {
        dist = tau / mu_tot;
    } else {
        tflag = true;
    }
    return true;
}

/* ---------------------------------------------------------------------------------------
 * stokes, reference src/stokes.f90:7-166 (Henyey-Greenstein direction update)
 * ------------------------------------------------------------------------------------- */
static thread_local int64_t t_scatter_events = 0; /* stokes() calls of this thread (status_hist slot 27) */
inline void stokes(vec& dir, double hgg, Draws& rng) {
    ++t_scatter_events;
    double nxp = dir.x, nyp = dir.y, nzp = dir.z;
    double cost = dir.z;
    double sint = std::sqrt(1. - cost * cost);
    double g2 = hgg * hgg;
    double phi = jit(std::atan2(dir.y, dir.x));
    double cosp, sinp;

    if (hgg == 0.0) { /* :33-48 isotropic */
        cost = 2. * rng.scatter() - 1.;
        sint = (1. - cost * cost);
        if (sint <= 0.) {
            sint = 0.;
        } else {
            sint = std::sqrt(sint);
        }
        phi = TWOPI * rng.scatter();
        sinp = jit(std::sin(phi));
        cosp = jit(std::cos(phi));
        nxp = sint * cosp;
        nyp = sint * sinp;
        nzp = cost;
    } else { /* :54-158 */
        double costp = cost, sintp = sint, phip = phi;
        double t = (1. - g2) / (1. - hgg + 2. * hgg * rng.scatter());
        double bmu = ((1. + g2) - t * t) / (2. * hgg);
        double cosb2 = bmu * bmu;
        double b = cosb2 - 1.;
        (void)b;
        if (std::fabs(bmu) > 1.) {
            if (bmu > 1.) {
                bmu = 1.;
                cosb2 = 1.;
            } else {
                bmu = -1.;
                cosb2 = 1.;
            }
        }
        double sinbt = std::sqrt(1. - cosb2);
        double ri1 = TWOPI * rng.scatter();
        double sini2, cosi2 = 0., bott, cosdph;

        if (ri1 > PI) { /* :76-113 */
            double ri3 = TWOPI - ri1;
            double cosi3 = jit(std::cos(ri3));
            double sini3 = jit(std::sin(ri3));
            if (bmu == 1. || bmu == -1.) goto L100; /* :81-87 */
            cost = costp * bmu + sintp * sinbt * cosi3;
            if (std::fabs(cost) < 1.) {
                sint = std::fabs(std::sqrt(1. - cost * cost));
                sini2 = sini3 * sintp / sint;
                bott = sint * sinbt;
                cosi2 = costp / bott - cost * bmu / bott;
            } else {
                sint = 0.;
                sini2 = 0.;
                if (cost >= 1.) cosi2 = -1.;
                if (cost <= -1.) cosi2 = 1.;
            }
            cosdph = -cosi2 * cosi3 + sini2 * sini3 * bmu;
            if (std::fabs(cosdph) > 1.) {
                if (cosdph > 1.) {
                    cosdph = 1.;
                } else {
                    cosdph = -1.;
                }
            }
            phi = phip + jit(std::acos(cosdph));
            if (phi > TWOPI) phi = phi - TWOPI;
            if (phi < 0.) phi = phi + TWOPI;
        } else { /* :116-151 */
            double cosi1 = jit(std::cos(ri1));
            double sini1 = jit(std::sin(ri1));
            if (bmu == 1. || bmu == -1.) goto L100; /* :119-125 */
            cost = costp * bmu + sintp * sinbt * cosi1;
            if (std::fabs(cost) < 1.) {
                sint = std::fabs(std::sqrt(1. - cost * cost));
                sini2 = sini1 * sintp / sint;
                bott = sint * sinbt;
                cosi2 = costp / bott - cost * bmu / bott;
            } else {
                sint = 0.;
                sini2 = 0.;
                if (cost >= 1.) cosi2 = -1.;
                if (cost <= -1.) cosi2 = 1.;
            }
            cosdph = -cosi1 * cosi2 + sini1 * sini2 * bmu;
            if (std::fabs(cosdph) > 1.) {
                if (cosdph > 1.) {
                    cosdph = 1.;
                } else {
                    cosdph = -1.;
                }
            }
            phi = phip - jit(std::acos(cosdph));
            if (phi > TWOPI) phi = phi - TWOPI;
            if (phi < 0.) phi = phi + TWOPI;
        }
        cosp = jit(std::cos(phi));
        sinp = jit(std::sin(phi));
        nxp = sint * cosp;
        nyp = sint * sinp;
        nzp = cost;
    }
L100:
    dir = {nxp, nyp, nzp};
}

/* ---------------------------------------------------------------------------------------
 * sources, reference src/sourceMod.f90
 * ------------------------------------------------------------------------------------- */
/* point, src/sourceMod.f90:12-47 */
inline void source_point(vec& pos, vec& dir, double cosThetaMax, double offset, Draws& rng) {
    double phi = TWOPI * rng.slot(SLOT_SRC1);
    double cosp = std::cos(phi);
    double sinp = std::sin(phi);
    double ran = rng.slot(SLOT_SRC0);
    double cost = (1.0 - ran) + ran * cosThetaMax;
    double sint = std::sqrt(1.0 - cost * cost);
    dir = {sint * cosp, sint * sinp, cost};
    pos = {0.0, 0.0, 0.0 + offset};
}

/* ring, src/sourceMod.f90:250-300 */
inline void source_ring(vec& pos, vec& dir, const ort_plano& lens, double r1, double r2, double Ra,
                        double Rb, bool ellipse, double bottleOffset, Draws& rng) {
    double r = r1 + rng.slot(SLOT_SRC0) * (r2 - r1); /* ranu, src/random_mod.f90:48-57 */
    double theta = rng.slot(SLOT_SRC1) * TWOPI;
    double posx = std::sqrt(r) * std::cos(theta);
    double posy = std::sqrt(r) * std::sin(theta);
    double posz;
    if (ellipse) {
        double q = posy * Ra / Rb;
        posz = bottleOffset + std::sqrt(Ra * Ra - q * q);
    } else {
        posz = bottleOffset + std::sqrt(Ra * Ra - posy * posy);
    }
    pos = {posx, posy, posz};

    double rl = lens.radius + 10e-3;
    r = 0. + rng.slot(SLOT_SRC2) * (rl * rl - 0.);
    theta = rng.slot(SLOT_SRC3) * TWOPI;
    posx = std::sqrt(r) * std::cos(theta);
    posy = std::sqrt(r) * std::sin(theta);
    vec lenspoint = {posx, posy, lens.fb};

    double dx = lenspoint.x - pos.x, dy = lenspoint.y - pos.y, dz = lenspoint.z - pos.z;
    double dist = std::sqrt(dx * dx + dy * dy + dz * dz);
    dir = {(lenspoint.x - pos.x) / dist, (lenspoint.y - pos.y) / dist, (lenspoint.z - pos.z) / dist};
    dir = magnitude(dir);
}

/* rang, src/random_mod.f90:59-85: polar Box-Muller; the rejection loop consumes the sequential
 * draws (slots 16, 17, ...) */
inline void rang(double& x, double& y, double avg, double sigma, Draws& rng) {
    double s = 1.;
    while (s >= 1.) {
        x = -1. + rng.scatter() * (1. - -1.);
        y = -1. + rng.scatter() * (1. - -1.);
        s = y * y + x * x;
        if (rng.override_u >= 0.) break; /* constant test draws would never leave the loop */
    }
    double cst = std::sqrt(-2. * std::log(s) / s);
    double tmp = x * cst;
    x = avg + sigma * tmp;
    tmp = y * cst;
    y = avg + sigma * tmp;
}

/* point_on_bottle, src/sourceMod.f90:50-89 (crs source of the ring loop).  Returns false when
 * the spot point misses the cylinder (the reference then uses an undefined t). */
inline bool source_crs(vec& pos, vec& dir, double cosThetaMax, const ort_bottle& B, double spot_radius,
                       Draws& rng) {
    double phi = TWOPI * rng.slot(SLOT_SRC1);
    double cosp = std::cos(phi);
    double sinp = std::sin(phi);
    double ran = rng.slot(SLOT_SRC0);
    double cost = (1.0 - ran) + ran * cosThetaMax;
    double sint = std::sqrt(1.0 - cost * cost);
    double nxp = sint * cosp, nyp = sint * sinp, nzp = cost;
    double tmp1, tmp2;
    rang(tmp1, tmp2, 0., spot_radius, rng);
    pos = {tmp1, tmp2, 1.0};
    dir = {0., 0., -1.};
    double t = 0.;
    bool flag = intersect_cylinder(pos, dir, t, v3(B.centre), B.radiusa + B.thickness);
    if (!flag) return false;
    pos = pos + dir * t;
    dir = {nxp, nyp, nzp};
    return true;
}

/* create_spot, src/sourceMod.f90:122-159; n is the 1-based loop index, nrays = nphotons */
inline void source_spot(vec& pos, vec& dir, double cosThetaMax, int64_t nrays, int64_t n) {
    double nrays_sqrt = std::sqrt((double)nrays);
    double phimax = TWOPI;
    double thetaMax = std::acos(cosThetaMax);
    double deltaPhi = phimax / nrays_sqrt;
    double deltaTheta = thetaMax / nrays_sqrt;
    double phi = deltaPhi * (double)(n % 10);
    double theta = deltaTheta * (double)(n / 10);
    double sinp = std::sin(phi), cosp = std::cos(phi);
    double cost = std::cos(theta);
    double sint = std::sqrt(1. - cost * cost);
    dir = {sint * cosp, sint * sinp, cost};
    pos = {0., 0., 0.};
}

/* emit_image + emit, src/sourceMod.f90:303-361.  `cdf` = inclusive prefix sums of the budget in
 * the order emit_image scans it (second index outer, first index inner = memory order);
 * ray k goes to the first pixel whose prefix sum exceeds k.  Slots: x 0, y 1, aim r 10, aim theta 11. */
inline bool source_image(vec& pos, vec& dir, const ort_plano& lens, const int64_t* cdf, int64_t k, Draws& rng) {
    const int64_t npix = (int64_t)ORT_SRCIMG_N * ORT_SRCIMG_N;
    if (!cdf || k >= cdf[npix - 1]) return false;
    int64_t lo = 0, hi = npix - 1;
    while (lo < hi) {
        int64_t mid = (lo + hi) / 2;
        if (cdf[mid] > k) hi = mid; else lo = mid + 1;
    }
    int64_t j = lo % ORT_SRCIMG_N + 1; /* first index of img(j,i): passed to emit as its `i` */
    int64_t i = lo / ORT_SRCIMG_N + 1; /* second index: emit's `j` */
    double dx = 5000e-6 / 512.;
    double ax = (j - 1.) * dx, bx = j * dx, ay = (i - 1.) * dx, by = i * dx;
    double x = (ax + rng.slot(0) * (bx - ax)) - 2500e-6;
    double y = (ay + rng.slot(1) * (by - ay)) - 2500e-6;
    pos = {x, y, 0.0};
    double r = 0. + rng.slot(10) * ((lens.radius * lens.radius) - 0.);
    double theta = rng.slot(11) * TWOPI;
    double posx = std::sqrt(r) * std::cos(theta);
    double posy = std::sqrt(r) * std::sin(theta);
    vec lenspoint = {posx, posy, lens.fb};
    double ex = lenspoint.x - pos.x, ey = lenspoint.y - pos.y, ez = lenspoint.z - pos.z;
    double dist = std::sqrt(ex * ex + ey * ey + ez * ez);
    dir = {(lenspoint.x - pos.x) / dist, (lenspoint.y - pos.y) / dist, (lenspoint.z - pos.z) / dist};
    dir = magnitude(dir);
    return true;
}
std::vector<int64_t> g_image_cdf; /* set by orc_set_image_source */

/* iSORS with ring = .true., src/sourceMod.f90:162-247.  Returns false at the reference's
 * `error stop "no intersection with bottle!"`. */
inline bool source_isors(vec& pos, vec& dir, const ort_bottle& B, const ort_plano& L1, double seperation,
                         double beam_width, Draws& rng) {
    double axicon_n = 1.4;
    double radius = 12.7e-3;
    double height = 1.1e-3;
    double alpha = std::atan(height / radius);
    double k = (radius / height) * (radius / height);
    double base_pos = (seperation + beam_width) / std::tan(alpha * (axicon_n - 1.));
    vec centre = {0., 0., 0.};
    double posx = 0., posy = 0., t;
    rang(posx, posy, 0., beam_width, rng);
    pos = centre + vec{posx, posy, 2 * height};
    dir = {0., 0., -1.};
    bool flag = intersect_cone(pos, dir, t, centre, radius, height);
    if (flag) {
        pos = pos + t * dir;
        vec normal = {2 * (pos.x - centre.x) / k, 2 * (pos.y - centre.y) / k, -2 * (pos.z - centre.z) + 2 * height};
        normal = normal * (-1.);
        normal = magnitude(normal);
        reflect_refract(dir, normal, axicon_n, 1., flag, rng.slot(SLOT_SRC2));
        t = (base_pos) / dir.z;
        pos = pos + t * dir;
        pos.z = B.radiusa + B.centre[2] + 2.220446049250313e-16; /* epsilon(1.) with -freal-4-real-8 */
        if (B.ellipse) {
            double rad1 = B.radiusa - B.thickness;
            double rad2 = B.radiusb - B.thickness;
            flag = intersect_ellipse(pos, dir, t, v3(B.centre), rad1, rad2);
        } else {
            flag = intersect_cylinder(pos, dir, t, v3(B.centre), B.radiusa - B.thickness);
        }
        if (!flag) return false;
        pos = pos + t * dir;
    }
    double r = 0. + rng.slot(SLOT_SRC0) * ((L1.radius * L1.radius) - 0.);
    double theta = rng.slot(SLOT_SRC1) * TWOPI;
    posx = std::sqrt(r) * std::cos(theta);
    posy = std::sqrt(r) * std::sin(theta);
    vec lenspoint = {posx, posy, L1.fb};
    double dx = lenspoint.x - pos.x, dy = lenspoint.y - pos.y, dz = lenspoint.z - pos.z;
    double dist = std::sqrt(dx * dx + dy * dy + dz * dz);
    dir = {(lenspoint.x - pos.x) / dist, (lenspoint.y - pos.y) / dist, (lenspoint.z - pos.z) / dist};
    dir = magnitude(dir);
    return true;
}

/* ---------------------------------------------------------------------------------------
 * optical elements, reference src/lens.f90
 * ------------------------------------------------------------------------------------- */

/* bottle_forward_sub, src/lens.f90:230-350.  Returns 0 when the ray leaves the bottle. */
inline int bottle_forward(const ort_bottle& B, vec centre, vec& pos, vec& dir, Draws& rng, int flags) {
    double t = 0.;
    bool flag;
    /* inner surface :249-260 */
    if (B.ellipse) {
        double rad1 = B.radiusa - B.thickness;
        double rad2 = B.radiusb - B.thickness;
        flag = intersect_ellipse(pos, dir, t, centre, rad1, rad2);
    } else {
        flag = intersect_cylinder(pos, dir, t, centre, B.radiusa - B.thickness);
    }
    if (!flag) return ORT_ST_BOTTLE_INNER_MISS;

    if (B.scatter_c) { /* :262-282 */
        flag = false;
        if (!tauint(pos, dir, B.mua_c, B.mus_c, centre, B.radiusa - B.thickness, t, flag, rng.scatter()))
            return ORT_ST_TAUINT_MISS;
        while (!flag) {
            pos = pos + t * dir;
            if (rng.scatter() < B.mus_c / (B.mus_c + B.mua_c)) {
                stokes(dir, .65, rng);
            } else {
                return ORT_ST_CONTENTS_ABSORBED;
            }
            if (!tauint(pos, dir, B.mua_c, B.mus_c, centre, B.radiusa - B.thickness, t, flag, rng.scatter()))
                return ORT_ST_TAUINT_MISS;
            if (std::sqrt(pos.x * pos.x + pos.z * pos.z) >= B.radiusa - B.thickness) break;
        }
        if (dir.z < 0.) return ORT_ST_CONTENTS_BACKWARD;
    }

    pos = pos + t * dir; /* :284 */
    vec orig = pos;
    orig.x = centre.x;
    vec normal = centre - orig;
    normal = magnitude(normal);

    reflect_refract(dir, normal, B.ncontents, B.nbottle, flag, rng.slot(SLOT_SRC2)); /* :293 */
    if (flag) return ORT_ST_BOTTLE_INNER_REFLECT;

    /* outer surface :299-308 */
    if (B.ellipse) {
        if (flags & ORT_FLAG_FIX_OUTER_ELLIPSE)
            flag = intersect_ellipse(pos, dir, t, centre, B.radiusa, B.radiusb);
        else
            flag = intersect_ellipse(pos, dir, t, centre, B.radiusa / 2., B.radiusb / 2.);
    } else {
        flag = intersect_cylinder(pos, dir, t, centre, B.radiusa);
    }
    if (!flag) return ORT_ST_BOTTLE_OUTER_MISS;

    if (B.scatter_b) { /* :312-333 */
        flag = false;
        if (!tauint(pos, dir, B.mua_b, B.mus_b, centre, B.radiusa, t, flag, rng.scatter()))
            return ORT_ST_TAUINT_MISS;
        while (!flag) {
            pos = pos + t * dir;
            if (rng.scatter() < B.mus_b / (B.mus_b + B.mua_b)) {
                stokes(dir, 0.9, rng);
            } else {
                return ORT_ST_WALL_ABSORBED;
            }
            if (!tauint(pos, dir, B.mua_b, B.mus_b, centre, B.radiusa, t, flag, rng.scatter()))
                return ORT_ST_TAUINT_MISS;
            if (std::sqrt(pos.x * pos.x + pos.z * pos.z) >= B.radiusa) break;
        }
        if (dir.z < 0.) return ORT_ST_WALL_BACKWARD;
    }

    pos = pos + t * dir; /* :335 */
    orig = pos;
    orig.x = centre.x;
    normal = centre - orig;
    normal = magnitude(normal);

    reflect_refract(dir, normal, B.nbottle, 1.0, flag, rng.slot(SLOT_SRC3)); /* :344 */
    if (flag) return ORT_ST_BOTTLE_OUTER_REFLECT;
    return 0;
}

/* plano_forward_sub, src/lens.f90:425-481 */
inline int plano_forward(const ort_plano& L, vec& pos, vec& dir, Draws& rng) {
    vec centre = v3(L.centre), flatNormal = v3(L.flat_normal);
    double a = centre.z + L.curve_radius - L.thickness;
    double d = (a - pos.z) / dir.z;
    pos = pos + dir * d;
    double r = std::sqrt(pos.x * pos.x + pos.y * pos.y);
    if (r > L.radius) return ORT_ST_L2_APERTURE;

    bool flag;
    reflect_refract(dir, flatNormal, L.n1, L.n2, flag, rng.slot(SLOT_L2_FLAT)); /* flag ignored :458-459 */

    double t = 0.;
    flag = intersect_sphere(pos, dir, t, centre, L.curve_radius);
    if (!flag) return ORT_ST_L2_SPHERE_MISS;
    pos = pos + t * dir;

    vec curvedNormal = centre - pos;
    curvedNormal = magnitude(curvedNormal);
    reflect_refract(dir, curvedNormal, L.n2, L.n1, flag, rng.slot(SLOT_L2_CURVED));
    if (flag) return ORT_ST_L2_CURVED_REFLECT;
    return 0;
}

/* doublet_forward_sub, src/lens.f90:531-645 */
inline int doublet_forward(const ort_doublet& L, vec& pos, vec& dir, bool iris1, bool iris2,
                           double iris_radius, Draws& rng) {
    vec c1 = v3(L.centre1), c2 = v3(L.centre2), c3 = v3(L.centre3);
    double t, r;
    bool flag;
    if (iris1) { /* :551-565 */
        vec origpos = pos;
        t = ((c1.z - L.R1) - pos.z) / dir.z;
        pos = pos + dir * t;
        r = std::sqrt(pos.x * pos.x + pos.y * pos.y);
        if (r > L.radius * iris_radius) return ORT_ST_L3_IRIS_BEFORE;
        pos = origpos;
    }
    flag = intersect_sphere(pos, dir, t, c1, L.R1); /* :568 */
    if (!flag) return ORT_ST_L3_S1_MISS;
    pos = pos + t * dir;
    r = std::sqrt(pos.x * pos.x + pos.y * pos.y);
    if (r > (L.radius * 1.0)) return ORT_ST_L3_APERTURE;

    vec normal = pos - c1;
    normal = magnitude(normal);
    reflect_refract(dir, normal, L.n1, L.n2, flag, rng.slot(SLOT_L3_S1));
    if (flag) return ORT_ST_L3_S1_REFLECT;

    flag = intersect_sphere(pos, dir, t, c2, L.R2); /* :595 */
    if (!flag) return ORT_ST_L3_S2_MISS;
    pos = pos + t * dir;
    normal = c2 - pos;
    normal = magnitude(normal);
    reflect_refract(dir, normal, L.n2, L.n3, flag, rng.slot(SLOT_L3_S2));
    if (flag) return ORT_ST_L3_S2_REFLECT;

    flag = intersect_sphere(pos, dir, t, c3, L.R3); /* :616 */
    if (!flag) return ORT_ST_L3_S3_MISS;          /* reference: error stop "Help3" */
    pos = pos + t * dir;
    normal = c3 - pos;
    normal = magnitude(normal);
    reflect_refract(dir, normal, L.n3, L.n1, flag, rng.slot(SLOT_L3_S3));
    if (flag) return ORT_ST_L3_S3_REFLECT;

    if (iris2) { /* :632-644 */
        vec origpos = pos;
        t = ((c3.z + L.R3) - pos.z) / dir.z;
        pos = pos + dir * t;
        r = std::sqrt(pos.x * pos.x + pos.y * pos.y);
        if (r > L.radius * iris_radius) return ORT_ST_L3_IRIS_AFTER;
        pos = origpos;
    }
    return 0;
}

/* makeImage2D, src/imageMod.f90:19-58.  Returns status; bin in (xp, yp). */
inline int make_image(vec dir, vec pos, double diameter, int& xp, int& yp) {
    vec n = {0., 0., -1.};
    n = magnitude(n);
    vec d = magnitude(dir);
    d = (-1.) * d;
    double top = dot(n, d);
    double bottom = std::sqrt(dot(d, d)) * std::sqrt(dot(n, n));
    double angle = std::acos(top / bottom);
    double na = std::asin(0.22);
    if (angle > na) return ORT_ST_NA_REJECT; /* NaN angle passes, as in the reference */
    double binwid = diameter / 401.;
    if (pos.x > 1000 || pos.y > 1000) return ORT_ST_FAR;
    double fx = std::floor(pos.x / binwid), fy = std::floor(pos.y / binwid);
    /* the reference converts to int32 here (UB when huge): non-finite or
     * unrepresentable values are a miss */
    if (!(std::fabs(fx) < 2.0e9) || !(std::fabs(fy) < 2.0e9)) return ORT_ST_FAR;
    xp = (int)fx;
    yp = (int)fy;
    if (std::abs(xp) > 200 || std::abs(yp) > 200) return ORT_ST_OFF_DETECTOR;
    return ORT_ST_BINNED;
}

struct RayOut {
    vec pos, dir;
    int status, xp, yp;
};

/* One iteration of src/main.f90:90-109 (ring) or :127-162 (point), incl. telescope
 * (src/optics_system.f90:6-52).  `have_input`: the ray comes from the caller instead of the
 * source. */
inline RayOut trace_one(const ort_job& J, const ort_scene& S, int64_t ray, bool have_input, vec pos,
                        vec dir) {
    Draws rng(J.seed, (uint32_t)J.phase, (uint64_t)ray, J.uniform_override);
    jitter_begin(ray);
    RayOut o;
    o.xp = o.yp = INT32_MIN;
    int st = 0;
    vec bcentre = v3(S.bottle.centre);
    auto done = [&](int status) {
        o.pos = pos;
        o.dir = dir;
        o.status = status;
        return o;
    };
    if (!have_input) { /* source dispatch of src/main.f90:95-101 and :132-142 */
        if (J.phase == ORT_PHASE_RING) {
            if (J.source_kind == ORT_SRC_ISORS) {
                if (!source_isors(pos, dir, S.bottle, S.L2, S.isors_offset, S.ring_width, rng))
                    return done(ORT_ST_SOURCE_MISS);
            } else if (J.source_kind == ORT_SRC_CRS) {
                if (!source_crs(pos, dir, S.cos_theta_max, S.bottle, S.spot_size, rng))
                    return done(ORT_ST_SOURCE_MISS);
            } else {
                source_ring(pos, dir, S.L2, S.r1, S.r2, S.bottle.radiusa, S.bottle.radiusb,
                            S.bottle.ellipse != 0, bcentre.z, rng);
            }
        } else {
            if (J.source_kind == ORT_SRC_IMAGE) {
                if (!source_image(pos, dir, S.L2, g_image_cdf.empty() ? nullptr : g_image_cdf.data(), ray, rng))
                    return done(ORT_ST_SOURCE_MISS);
            } else if (J.source_kind == ORT_SRC_SPOT) {
                source_spot(pos, dir, S.cos_theta_max, J.total_rays > 0 ? J.total_rays : J.nrays, ray + 1);
            } else {
                source_point(pos, dir, S.cos_theta_max, S.point_offset, rng);
            }
        }
    }
    if (J.stop_after == ORT_STOP_SOURCE) return done(ORT_ST_STOPPED);
    if (J.phase == ORT_PHASE_POINT && J.use_bottle) { /* src/main.f90:145-155 */
        st = bottle_forward(S.bottle, bcentre, pos, dir, rng, J.flags);
        if (st) return done(st);
    }
    if (J.stop_after == ORT_STOP_BOTTLE) return done(ORT_ST_STOPPED);
    st = plano_forward(S.L2, pos, dir, rng); /* src/optics_system.f90:28 */
    if (st) return done(st);
    if (J.stop_after == ORT_STOP_L2) return done(ORT_ST_STOPPED);
    st = doublet_forward(S.L3, pos, dir, J.iris_before != 0, J.iris_after != 0, J.iris_radius, rng);
    if (st) return done(st);
    if (J.stop_after == ORT_STOP_L3) return done(ORT_ST_STOPPED);
    /* src/optics_system.f90:48-49 */
    double d = ((S.img_plane + J.fibre_offset) - pos.z) / dir.z;
    pos = pos + dir * d;
    st = make_image(dir, pos, J.image_diameter, o.xp, o.yp);
    if (st != ORT_ST_BINNED) o.xp = o.yp = INT32_MIN;
    return done(st);
}

/* ---------------------------------------------------------------------------------------
 * loaders + dispersion laws, reference src/lens.f90:73-227,647-695
 * ------------------------------------------------------------------------------------- */
/* first token of every line, Fortran list-directed style ("2.10d-3" -> 2.10e-3) */
bool first_tokens(const char* path, std::vector<std::string>& toks) {
    FILE* fh = std::fopen(path, "r");
    if (!fh) return false;
    char line[4096];
    while (std::fgets(line, sizeof line, fh)) {
        char* p = line;
        while (*p == ' ' || *p == '\t') ++p;
        char* q = p;
        while (*q && *q != ' ' && *q != '\t' && *q != '\n' && *q != '\r' && *q != ',') ++q;
        if (q == p) continue; /* blank line: list-directed read skips it */
        toks.emplace_back(p, q - p);
    }
    std::fclose(fh);
    return true;
}
double freal(const std::string& s) {
    std::string t = s;
    for (auto& ch : t)
        if (ch == 'd' || ch == 'D') ch = 'e';
    return std::strtod(t.c_str(), nullptr);
}

/* Sellmeier, src/lens.f90:647-665 */
double Sellmeier(double wave, double b1, double b2, double b3, double c1, double c2, double c3) {
    double w = wave * 1e6;
    double wave2 = w * w;
    double a = (b1 * wave2) / (wave2 - c1);
    double b = (b2 * wave2) / (wave2 - c2);
    double c = (b3 * wave2) / (wave2 - c3);
    return std::sqrt(1.0 + (a + b + c));
}
/* cauchy, src/lens.f90:667-680 (x**(-2), x**(-4) as gfortran expands integer powers) */
double cauchy(double wave, double a, double b, double c) {
    double w = wave * 1e6;
    double w2 = w * w;
    return a + b * (1.0 / w2) + c * (1.0 / (w2 * w2));
}
/* dispersion, src/lens.f90:682-695 */
double dispersion(double wave, double a, double b, double c) {
    double w = wave * 1e6;
    double wave2 = w * w;
    return a - b * wave2 + (c / wave2);
}

}  // namespace

extern "C" {

/* init_plano_convex, src/lens.f90:129-167 */
int orc_load_plano(const char* path, double wavelength, double offset, ort_plano* o) {
    std::vector<std::string> t;
    if (!first_tokens(path, t)) return ORT_EIO;
    if (t.size() < 12) return ORT_EPARSE;
    o->thickness = freal(t[0]);
    o->curve_radius = freal(t[1]);
    o->diameter = freal(t[2]);
    o->f = freal(t[3]);
    o->fb = freal(t[4]);
    o->n1 = freal(t[5]);
    o->n2 = Sellmeier(wavelength, freal(t[6]), freal(t[7]), freal(t[8]), freal(t[9]), freal(t[10]),
                      freal(t[11]));
    o->radius = o->diameter / 2.0;
    o->centre[0] = 0.;
    o->centre[1] = 0.;
    o->centre[2] = offset + (o->fb + o->thickness) - o->curve_radius;
    o->flat_normal[0] = 0.;
    o->flat_normal[1] = 0.;
    o->flat_normal[2] = -1.;
    return 0;
}

/* init_achromatic_doublet, src/lens.f90:73-126 */
int orc_load_doublet(const char* path, double wavelength, double offset, ort_doublet* o) {
    std::vector<std::string> t;
    if (!first_tokens(path, t)) return ORT_EIO;
    if (t.size() < 21) return ORT_EPARSE;
    o->thickness1 = freal(t[0]);
    o->thickness2 = freal(t[1]);
    o->R1 = freal(t[2]);
    o->R2 = freal(t[3]);
    o->R3 = freal(t[4]);
    o->diameter = freal(t[5]);
    o->f = freal(t[6]);
    o->fb = freal(t[7]);
    o->n1 = freal(t[8]);
    o->n2 = Sellmeier(wavelength, freal(t[9]), freal(t[10]), freal(t[11]), freal(t[12]), freal(t[13]),
                      freal(t[14]));
    o->n3 = Sellmeier(wavelength, freal(t[15]), freal(t[16]), freal(t[17]), freal(t[18]),
                      freal(t[19]), freal(t[20]));
    o->radius = o->diameter / 2.0;
    o->thickness = o->thickness1 + o->thickness2;
    for (int i = 0; i < 2; ++i) o->centre1[i] = o->centre2[i] = o->centre3[i] = 0.;
    o->centre1[2] = offset + o->fb + o->R1;
    o->centre2[2] = offset + o->fb + o->thickness1 - o->R2;
    o->centre3[2] = offset + o->fb + o->thickness - o->R3;
    return 0;
}

/* init_bottle, src/lens.f90:170-227 (tolerant of the 14-line file) */
int orc_load_bottle(const char* path, double wavelength, ort_bottle* o) {
    std::vector<std::string> t;
    if (!first_tokens(path, t)) return ORT_EIO;
    if (t.size() < 12) return ORT_EPARSE;
    o->thickness = freal(t[0]);
    o->radiusa = freal(t[1]);
    o->radiusb = freal(t[2]);
    o->centre[0] = freal(t[3]);
    o->centre[1] = freal(t[4]);
    o->centre[2] = freal(t[5]);
    double mu[4] = {0., 0., 0., 0.};
    for (size_t i = 12; i < t.size() && i < 16; ++i) mu[i - 12] = freal(t[i]);
    o->mua_b = mu[0];
    o->mus_b = mu[1];
    o->mua_c = mu[2];
    o->mus_c = mu[3];
    o->nbottle = dispersion(wavelength, freal(t[6]), freal(t[7]), freal(t[8]));
    o->ncontents = cauchy(wavelength, freal(t[9]), freal(t[10]), freal(t[11]));
    o->scatter_b = (o->mua_b + o->mus_b != 0.0) ? 1 : 0;
    o->scatter_c = (o->mua_c + o->mus_c != 0.0) ? 1 : 0;
    o->ellipse = (o->radiusa != o->radiusb) ? 1 : 0;
    o->_pad = 0;
    return 0;
}

/* The prologue of src/main.f90:51-70,81: fills the derived scalars of a scene whose bottle, L2
 * and L3 are already loaded; applies the offset guard (:54-58).  alpha in degrees as read from
 * settings.params (converted at src/setupMod.f90:61). */
int orc_derive_scene(ort_scene* S, double alpha_deg, double n_axicon, double ring_width,
                     int isors_source, double isors_offset, double spot_size) {
    { /* src/setupMod.f90:135-136, before the offset guard of main.f90 */
        double offset = S->bottle.radiusa + S->bottle.centre[2];
        S->spot_size = (spot_size * (S->L2.fb - offset)) / S->L2.fb;
        S->isors_offset = isors_offset;
        S->ring_width = ring_width;
    }
    double alpha = alpha_deg * PI / 180.;
    double angle = std::atan(S->L2.radius / S->L2.fb);
    S->cos_theta_max = std::cos(angle);
    if (S->L2.fb <= S->bottle.radiusa + S->bottle.centre[2]) {
        S->bottle.centre[2] = S->L2.fb - S->bottle.radiusa - 2e-3;
    }
    double distance;
    if (isors_source) {
        distance = S->bottle.radiusa + isors_offset;
    } else {
        distance = (S->bottle.radiusa + S->bottle.centre[2]);
    }
    double besselDiameter = distance * 97.3e-3 * std::tan(alpha * (n_axicon - 1)) / (S->L2.fb);
    double r1 = besselDiameter - ring_width;
    S->r2 = (besselDiameter / 2.0) * (besselDiameter / 2.0);
    S->r1 = r1 * r1;
    S->img_plane = 2. * (S->L2.fb + S->L3.fb) + S->L2.thickness + S->L3.thickness;
    S->point_offset = isors_source ? S->bottle.centre[2] : 0.0;
    return 0;
}

/* init_emit_image, src/sourceMod.f90:363-408 (serial build: nphotonsLocal = nphotons).
 * budget[(j-1)*512 + (i-1)] = imgin(i,j).  The rounding draw of pixel (i,j) is slot 0 of ray
 * (i-1)*512 + (j-1) of stream 3 (the reference draws them in that loop order, before init_rng). */
int orc_load_image_source(const char* path, int64_t nphotons, uint64_t seed, int32_t* budget) {
    const int N = ORT_SRCIMG_N;
    std::vector<double> f((size_t)N * N);
    FILE* fh = std::fopen(path, "rb");
    if (!fh) return ORT_EIO;
    size_t got = std::fread(f.data(), sizeof(double), f.size(), fh);
    std::fclose(fh);
    if (got != f.size()) return ORT_EPARSE;
    /* imgout(a,b) = f[(b-1)*N + (a-1)]; after the transpose imgout(i,j) = f[(i-1)*N + (j-1)] */
    double tot = 0.;
    for (int b = 1; b <= N; ++b)         /* sum() runs over the array in memory order */
        for (int a = 1; a <= N; ++a) tot += f[(size_t)(a - 1) * N + (b - 1)];
    for (int i = 1; i <= N; ++i) {
        for (int j = 1; j <= N; ++j) {
            double v = f[(size_t)(i - 1) * N + (j - 1)];
            double tmp = ((double)nphotons * v) / tot;
            double diff = tmp - (double)(int64_t)tmp;
            Draws rng(seed, 3, (uint64_t)((i - 1) * N + (j - 1)), -1.0);
            int32_t n = (int32_t)(int64_t)tmp;
            if (rng.slot(0) < diff && diff > 0) n += 1;
            budget[(size_t)(j - 1) * N + (i - 1)] = n;
        }
    }
    return 0;
}
int orc_set_image_source(const int32_t* budget) {
    g_image_cdf.clear();
    if (!budget) return 0;
    const size_t n = (size_t)ORT_SRCIMG_N * ORT_SRCIMG_N;
    g_image_cdf.resize(n);
    int64_t acc = 0;
    for (size_t k = 0; k < n; ++k) {
        acc += budget[k] > 0 ? budget[k] : 0;
        g_image_cdf[k] = acc;
    }
    return 0;
}

/* conditioning probe of the scatter path (see jit()): 0 = off */
int orc_set_jitter(uint64_t seed) {
    g_jitter_seed = seed;
    return 0;
}

int orc_uniforms(uint64_t seed, int32_t phase, int64_t ray, int32_t first_slot, int32_t n, double* out) {
    Draws rng(seed, (uint32_t)phase, (uint64_t)ray, -1.0);
    for (int i = 0; i < n; ++i) out[i] = rng.slot((uint32_t)(first_slot + i));
    return 0;
}

/* rounds [first, first + n) of Philox4x32: (0, 10) is Random123's philox4x32-10, (0, 7) the
 * generator used here, and (7, 3) applied to the latter's output must give the former */
int orc_philox(const uint32_t* ctr, const uint32_t* key, int32_t first, int32_t n, uint32_t* out) {
    philox4x32_rounds(ctr, key, first, n, out);
    return 0;
}
int orc_philox_rounds(void) { return PHILOX_ROUNDS; }

/* Ray indices in [first, first + n) whose ring-loop aim word (the high word of slot 2, see Draws::slot) equals
 * `word`: the rays that sit on the integer aperture cut of the CUDA kernels' stage A (2^-32 of all rays),
 * found by brute force for tests/golden/make_edge_rays.py.  Returns how many were found (at most `cap` stored). */
int64_t orc_find_aim_word(uint64_t seed, int32_t phase, uint32_t word, int64_t first, int64_t n, int64_t* out, int64_t cap) {
    const uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    const uint64_t q0 = (uint64_t)first >> 2, q1 = ((uint64_t)(first + n) + 3) >> 2;
    int64_t found = 0;
#pragma omp parallel for schedule(static)
    for (uint64_t q = q0; q < q1; ++q) {
        uint32_t ctr[4] = {(uint32_t)q, (uint32_t)(q >> 32), 16u + (uint32_t)phase, 0u}, sh[4];
        philox4x32_rounds(ctr, key, 0, PHILOX_ROUNDS, sh);
        for (int k = 0; k < 4; ++k) {
            const int64_t ray = (int64_t)(4 * q + k);
            if (sh[k] == word && ray >= first && ray < first + n) {
                int64_t slot;
#pragma omp atomic capture
                slot = found++;
                if (slot < cap) out[slot] = ray;
            }
        }
    }
    return found;
}

/* Same contract as ort_trace_rays (include/ort.h). */
int orc_trace_rays(const ort_job* job, const ort_scene* scene, int64_t n, const double* pos_in,
                   const double* dir_in, double* pos_out, double* dir_out, int32_t* status,
                   int32_t* bin_xy) {
    bool have = pos_in != nullptr && dir_in != nullptr;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        vec p = {0, 0, 0}, d = {0, 0, 1};
        if (have) {
            p = {pos_in[i], pos_in[n + i], pos_in[2 * n + i]};
            d = {dir_in[i], dir_in[n + i], dir_in[2 * n + i]};
        }
        RayOut o = trace_one(*job, *scene, job->first_ray + i, have, p, d);
        if (pos_out) {
            pos_out[i] = o.pos.x;
            pos_out[n + i] = o.pos.y;
            pos_out[2 * n + i] = o.pos.z;
        }
        if (dir_out) {
            dir_out[i] = o.dir.x;
            dir_out[n + i] = o.dir.y;
            dir_out[2 * n + i] = o.dir.z;
        }
        if (status) status[i] = o.status;
        if (bin_xy) {
            bin_xy[i] = o.xp;
            bin_xy[n + i] = o.yp;
        }
    }
    return 0;
}

/* Same contract as ort_trace (include/ort.h); the OpenMP loop mirrors src/main.f90:83-110 with
 * atomic image increments (src/imageMod.f90:55) and a reduction on the loss counter. */
int orc_trace(const ort_job* job, const ort_scene* scenes, int nscenes, uint64_t* image, int64_t* lost,
              int64_t* status_hist, int nthreads) {
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
    for (int s = 0; s < nscenes; ++s) {
        const ort_scene& S = scenes[s];
        uint64_t* img = image ? image + (size_t)s * ORT_IMG_BINS : nullptr;
        if (img) std::memset(img, 0, sizeof(uint64_t) * ORT_IMG_BINS);
        int64_t hist[ORT_NSTATUS] = {0};
#pragma omp parallel
        {
            int64_t h[ORT_NSTATUS] = {0};
            t_scatter_events = 0;
#pragma omp for schedule(static)
            for (int64_t i = 0; i < job->nrays; ++i) {
                RayOut o = trace_one(*job, S, job->first_ray + i, false, {0, 0, 0}, {0, 0, 1});
                h[o.status]++;
                if (o.status == ORT_ST_BINNED && img) {
                    size_t idx = (size_t)(o.yp + ORT_IMG_HALF) * ORT_IMG_N + (size_t)(o.xp + ORT_IMG_HALF);
#pragma omp atomic
                    img[idx] += 1;
                }
            }
            h[ORT_SCATTER_EVENTS_SLOT] = t_scatter_events;
#pragma omp critical
            for (int k = 0; k < ORT_NSTATUS; ++k) hist[k] += h[k];
        }
        int64_t l = 0;
        for (int k = 0; k < ORT_NSTATUS; ++k)
            if (ORT_STATUS_IS_LOST(k)) l += hist[k];
        if (lost) lost[s] = l;
        if (status_hist)
            for (int k = 0; k < ORT_NSTATUS; ++k) status_hist[(size_t)s * ORT_NSTATUS + k] = hist[k];
    }
    return 0;
}

/* Volume image: makeImage3D, src/imageMod.f90:61-90, in the place of makeImage2D (same call,
 * `image` of rank 4).  The reference never reaches it (main's image has rank 3); restated so that
 * the CUDA version has something to be compared with.  No NA test, no `> 1000` test: from the
 * image plane the ray is sampled at 200 depths dz = diameter / 200 apart and every sample inside
 * the 401 x 401 window increments its voxel; the first sample outside ends the ray.
 * volume[(i * 401 + (yp + 200)) * 401 + (xp + 200)], i = depth: the memory order of
 * image(-200:200, -200:200, 200, layer).  Status: BINNED when at least one voxel was hit, else
 * OFF_DETECTOR; non-finite samples count as outside. */
int orc_trace_volume(const ort_job* job, const ort_scene* scene, uint32_t* volume, int64_t* lost,
                     int64_t* status_hist, int nthreads) {
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
    const ort_scene& S = *scene;
    ort_job J = *job;
    J.stop_after = ORT_STOP_L3;
    std::memset(volume, 0, sizeof(uint32_t) * (size_t)ORT_VOL_DEPTH * ORT_IMG_BINS);
    int64_t hist[ORT_NSTATUS] = {0};
    const double binwid = J.image_diameter / 401.;
    const double dz = J.image_diameter / 200.;
#pragma omp parallel
    {
        int64_t h[ORT_NSTATUS] = {0};
#pragma omp for schedule(static)
        for (int64_t i = 0; i < J.nrays; ++i) {
            RayOut o = trace_one(J, S, J.first_ray + i, false, {0, 0, 0}, {0, 0, 1});
            if (o.status != ORT_ST_STOPPED) {
                h[o.status]++;
                continue;
            }
            vec pos = o.pos, dir = o.dir;
            double d = ((S.img_plane + J.fibre_offset) - pos.z) / dir.z; /* src/optics_system.f90:48-49 */
            pos = pos + dir * d;
            int hits = 0;
            for (int k = 0; k < ORT_VOL_DEPTH; ++k) {
                vec np = pos + dir * (k * dz);
                double fx = std::floor(np.x / binwid), fy = std::floor(np.y / binwid);
                if (!(std::fabs(fx) <= 200.0) || !(std::fabs(fy) <= 200.0)) break;
                size_t idx = ((size_t)k * ORT_IMG_N + (size_t)((int)fy + ORT_IMG_HALF)) * ORT_IMG_N +
                             (size_t)((int)fx + ORT_IMG_HALF);
#pragma omp atomic
                volume[idx] += 1;
                ++hits;
            }
            h[hits ? ORT_ST_BINNED : ORT_ST_OFF_DETECTOR]++;
        }
#pragma omp critical
        for (int k = 0; k < ORT_NSTATUS; ++k) hist[k] += h[k];
    }
    int64_t l = 0;
    for (int k = 0; k < ORT_NSTATUS; ++k)
        if (ORT_STATUS_IS_LOST(k)) l += hist[k];
    if (lost) *lost = l;
    if (status_hist)
        for (int k = 0; k < ORT_NSTATUS; ++k) status_hist[k] = hist[k];
    return 0;
}

int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* single-function probes used by the known-answer tests */
double orc_fresnel(const double* I, const double* N, double n1, double n2) {
    return fresnel(v3(I), v3(N), n1, n2);
}
int orc_intersect_sphere(const double* o, const double* d, const double* c, double R, double* t) {
    return intersect_sphere(v3(o), v3(d), *t, v3(c), R) ? 1 : 0;
}
int orc_intersect_cylinder(const double* o, const double* d, const double* c, double R, double* t) {
    return intersect_cylinder(v3(o), v3(d), *t, v3(c), R) ? 1 : 0;
}
int orc_intersect_ellipse(const double* o, const double* d, const double* c, double a, double b, double* t) {
    return intersect_ellipse(v3(o), v3(d), *t, v3(c), a, b) ? 1 : 0;
}
void orc_refract(double* I, const double* N, double eta) {
    vec i = v3(I);
    refract(i, v3(N), eta);
    I[0] = i.x; I[1] = i.y; I[2] = i.z;
}
void orc_reflect(double* I, const double* N) {
    vec i = v3(I);
    reflect(i, v3(N));
    I[0] = i.x; I[1] = i.y; I[2] = i.z;
}
void orc_stokes(double* dir, double hgg, uint64_t seed, int64_t ray) {
    Draws rng(seed, 2, (uint64_t)ray, -1.0);
    vec d = v3(dir);
    stokes(d, hgg, rng);
    dir[0] = d.x; dir[1] = d.y; dir[2] = d.z;
}
double orc_sellmeier(double w, double b1, double b2, double b3, double c1, double c2, double c3) {
    return Sellmeier(w, b1, b2, b3, c1, c2, c3);
}

}  // extern "C"
