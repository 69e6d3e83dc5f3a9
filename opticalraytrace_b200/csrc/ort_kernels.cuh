/*
 * ort_kernels.cuh -- the CUDA kernels of the trace loop (sm_100a).
 *
 *   ort_trace_kernel<PHASE,BOTTLE>   the production megakernel: persistent warps, one thread
 *                                    per ray, warp-private shared-memory queues that compact
 *                                    the live rays between surfaces, warp-aggregated 64-bit
 *                                    reductions into the detector image.
 *   ort_ring_cull_kernel             the ring loop's first kernel: integer aim-point test + the
 *                                    single-precision culling filter on two rays per lane (packed
 *                                    f32x2 arithmetic); no fp64.
 *   ort_ring_survivors_kernel        fp64 stages over the ray indices the cull kernel lists.
 *   ort_trace_flat_kernel<...>       the same path without compaction (diagnostic / evidence).
 *   ort_volume_kernel                makeImage3D (opt-in volume image).
 *   ort_rays_kernel                  explicit ray list, per-ray outputs (parity entry point).
 *   ort_uniforms_kernel              exposes the counter-based generator.
 *   ort_dfma_peak_kernel             FP64 FMA peak micro-benchmark (roofline denominator).
 *   ort_math_selftest_kernel         ulp error of the fast reciprocal / division / sqrt.
 *
 * Replaces the two `!$OMP do` loops of reference src/main.f90:90-109 and :127-162.
 *
 * Scene and job arrive as __grid_constant__ kernel parameters: they sit in the constant bank the
 * launch already uses, so every scene scalar is an immediate c[0][off] operand of the DFMA that
 * needs it (the first build indexed a __constant__ array with the scene number, which cost an
 * LDC through the address-divergence unit per scalar: 16-19 % ADU utilisation in ncu).
 */
#ifndef ORT_KERNELS_CUH
#define ORT_KERNELS_CUH

#include <cuda_runtime.h>

#include "ort_optics.cuh"
#include "ort_filter.cuh"

/* `make DEBUG=1` (install.sh -d): own bounds checks on every queue slot and image bin -- the
 * substitute for compute-sanitizer, which is closed on this GPU pool */
#ifdef ORT_DEBUG
#include <cassert>
#define ORT_ASSERT(c) assert(c)
#else
#define ORT_ASSERT(c) ((void)0)
#endif

#ifndef ORT_TPB
#define ORT_TPB 256
#endif
#define ORT_WPB (ORT_TPB / 32)
#define ORT_QCAP 64 /* a queue holds < 32 leftovers + <= 32 new survivors */
#define ORT_FULL 0xffffffffu
#ifndef ORT_MIN_BLOCKS
#define ORT_MIN_BLOCKS 3 /* 24 warps per SM: measured best (profiles/): +4-5 % over 2, 4 is no better */
#endif

/* structure-of-arrays ray queue private to one warp */
template <typename R>
struct WarpQueue {
    R px[ORT_QCAP], py[ORT_QCAP], pz[ORT_QCAP], dx[ORT_QCAP], dy[ORT_QCAP], dz[ORT_QCAP];
    uint32_t id[ORT_QCAP]; /* ray index relative to DevJob.first_ray */
};
/* The queue in front of L2.  Its rays stand on L2's flat face, the plane z = l2_flat_z, so z is not
 * stored; it carries L2's two decision words instead (slots 4 and 5: words 3 of the Philox blocks
 * 0 and 1 the emitting stage generated), so that L2 costs no generator call.  (With the aim-plane
 * shortcut the ring loop keeps its own 8-byte entries -- tested word, ray index -- in this memory:
 * ort_ring_quads_pass.) */
template <typename R>
struct WarpQueueL2 {
    R px[ORT_QCAP], py[ORT_QCAP], dx[ORT_QCAP], dy[ORT_QCAP], dz[ORT_QCAP];
    uint32_t id[ORT_QCAP], wf[ORT_QCAP], wc[ORT_QCAP];
};
template <typename R>
struct WarpShared {
    WarpQueueL2<R> q0;
    WarpQueue<R> q1;
    unsigned hist[ORT_NSTATUS]; /* this warp's histogram of final ray statuses */
};
#ifndef ORT_COUNT_SMEM
#define ORT_COUNT_SMEM 1 /* 1: statuses are tallied in the warp's shared-memory histogram (one ATOMS per
                            stage); 0: in 27 warp-uniform registers (vote + popc + add per status) */
#endif
/* one shared-memory increment per ray that ended in this stage */
__device__ __forceinline__ void ort_tally_smem(unsigned* hist, int st, bool ended) {
    if (ended) atomicAdd(hist + st, 1u);
}
__device__ __forceinline__ void ort_hist_clear(unsigned* hist, unsigned lane) {
    hist[lane] = 0u;
    __syncwarp();
}
__device__ __forceinline__ void ort_hist_flush(const unsigned* hist, unsigned lane, unsigned long long* __restrict__ counters) {
    __syncwarp();
    const unsigned mine = hist[lane];
    if (mine) atomicAdd(counters + lane, (unsigned long long)mine);
}

__device__ __forceinline__ OrtRng ort_make_rng(const DevJob& J, uint32_t local_id) {
    OrtRng g;
    unsigned long long ray = (unsigned long long)J.first_ray + local_id;
    g.k0 = (uint32_t)J.seed;
    g.k1 = (uint32_t)(J.seed >> 32);
    g.rk = J.round_keys;
    g.r0 = (uint32_t)ray;
    g.r1 = (uint32_t)(ray >> 32);
    g.phase = (uint32_t)J.phase;
    g.override_u = J.uniform_override;
    return g;
}
/* the trace kernels never use the test-only constant draw (ort_trace rejects it), so its
 * branches fold away */
__device__ __forceinline__ OrtRng ort_make_rng_prod(const DevJob& J, uint32_t local_id) {
    OrtRng g = ort_make_rng(J, local_id);
    g.override_u = -1.0;
    return g;
}

/* Returns the slot the entry went to (or -1). */
template <typename R>
__device__ __forceinline__ int ort_q_push(WarpQueue<R>& q, int& n, bool alive, const OrtRayT<R>& r,
                                          uint32_t id, unsigned lane) {
    unsigned m = __ballot_sync(ORT_FULL, alive);
    int p = -1;
    if (alive) {
        p = n + __popc(m & ((1u << lane) - 1u));
        ORT_ASSERT(p >= 0 && p < ORT_QCAP);
        q.px[p] = r.px; q.py[p] = r.py; q.pz[p] = r.pz;
        q.dx[p] = r.dx; q.dy[p] = r.dy; q.dz[p] = r.dz;
        q.id[p] = id;
    }
    n += __popc(m);
    return p;
}
/* ... and the slot the entry came from (or -1) */
template <typename R>
__device__ __forceinline__ int ort_q_pop(WarpQueue<R>& q, int& n, OrtRayT<R>& r, uint32_t& id, unsigned lane) {
    int cnt = n < 32 ? n : 32;
    int base = n - cnt;
    int p = -1;
    if ((int)lane < cnt) {
        p = base + lane;
        ORT_ASSERT(p >= 0 && p < ORT_QCAP);
        r.px = q.px[p]; r.py = q.py[p]; r.pz = q.pz[p];
        r.dx = q.dx[p]; r.dy = q.dy[p]; r.dz = q.dz[p];
        id = q.id[p];
    }
    n = base;
    return p;
}
/* the queue in front of L2 (the start point is on its flat face: no pz) */
template <typename R>
__device__ __forceinline__ void ort_q0_push(WarpQueueL2<R>& q, int& n, bool alive, const OrtRayT<R>& r, uint32_t id,
                                            uint32_t wf, uint32_t wc, unsigned lane) {
    unsigned m = __ballot_sync(ORT_FULL, alive);
    if (alive) {
        const int p = n + __popc(m & ((1u << lane) - 1u));
        ORT_ASSERT(p >= 0 && p < ORT_QCAP);
        q.px[p] = r.px; q.py[p] = r.py;
        q.dx[p] = r.dx; q.dy[p] = r.dy; q.dz[p] = r.dz;
        q.wf[p] = wf;
        q.wc[p] = wc;
        q.id[p] = id;
    }
    n += __popc(m);
}
template <typename R>
__device__ __forceinline__ bool ort_q0_pop(WarpQueueL2<R>& q, int& n, R flat_z, OrtRayT<R>& r, uint32_t& id, uint32_t& wf,
                                           uint32_t& wc, unsigned lane) {
    int cnt = n < 32 ? n : 32;
    int base = n - cnt;
    const bool act = (int)lane < cnt;
    if (act) {
        const int p = base + lane;
        ORT_ASSERT(p >= 0 && p < ORT_QCAP);
        r.px = q.px[p]; r.py = q.py[p];
        r.pz = flat_z;
        r.dx = q.dx[p]; r.dy = q.dy[p]; r.dz = q.dz[p];
        wf = q.wf[p];
        wc = q.wc[p];
        id = q.id[p];
    }
    n = base;
    return act;
}

/* Histogram of final ray statuses.  Every warp keeps one counter per status it can produce, as
 * plain (warp-uniform) registers: a stage adds popc(ballot(status == k)) -- vote, popc, add -- for
 * the few statuses that end most of its rays; the rest share one cold path behind an any().  No
 * atomics, no match.any, no shared memory; the counters reach memory once, at the end of the
 * launch (lane k adds counter k). */
struct OrtCounts {
    unsigned c[ORT_NSTATUS];
};
template <int K>
__device__ __forceinline__ void ort_count_one(OrtCounts& cnt, int st) {
    cnt.c[K] += __popc(__ballot_sync(ORT_FULL, st == K));
}
template <int... KS>
__device__ __forceinline__ void ort_count(OrtCounts& cnt, int st) {
    (ort_count_one<KS>(cnt, st), ...);
}
/* statuses outside the common list of a stage: cold */
template <int... KS>
__device__ __forceinline__ void ort_count_rare(OrtCounts& cnt, int st, bool is_rare) {
    if (__any_sync(ORT_FULL, is_rare)) {
        (ort_count_one<KS>(cnt, st), ...);
    }
}
__device__ __forceinline__ void ort_counts_flush(const OrtCounts& cnt, unsigned lane,
                                                 unsigned long long* __restrict__ counters) {
    unsigned mine = 0;
#pragma unroll
    for (int k = 0; k < ORT_NSTATUS; ++k)
        if (lane == (unsigned)k) mine = cnt.c[k];
    if (mine) atomicAdd(counters + lane, (unsigned long long)mine);
}

/* detector increment, reference src/imageMod.f90:55 (`!$omp atomic`): lanes that hit the same
 * bin are merged with match.any and one 64-bit RED is issued per distinct bin */
__device__ __forceinline__ void ort_bin(unsigned long long* img, bool binned, int xp, int yp, unsigned lane) {
    if (!__any_sync(ORT_FULL, binned)) return;
    unsigned key = binned ? (unsigned)((yp + ORT_IMG_HALF) * ORT_IMG_N + (xp + ORT_IMG_HALF)) : 0xffffffffu;
    ORT_ASSERT(!binned || key < (unsigned)ORT_IMG_BINS);
    unsigned m = __match_any_sync(ORT_FULL, key);
    if (binned && (int)lane == __ffs(m) - 1) atomicAdd(img + key, (unsigned long long)__popc(m));
}

/* ---- ring loop, stage A on QUADS -------------------------------------------------------------
 * The four rays 4q .. 4q+3 (global ray index) share the block that holds the high words of their aim-disc
 * draws (ort_shared_block), so one lane decides L2's aperture for four rays with one Philox block: a pass
 * takes the 32 quads [32 b, 32 b + 32) -- 128 rays -- and appends the survivors as (high word, ray id) to
 * the warp's queue q (room for 32 leftovers + 128).  Ray ids are offsets from J.first_ray; `mis` is how
 * far that is from a multiple of four.  on_edge(id, word) makes the call for a draw whose high word
 * EQUALS the cut's (2^-32 of the rays).  Returns the number of rays that ended here (warp-uniform). */
#define ORT_QUAD 4
#define ORT_QUAD_QCAP (32 + 32 * ORT_QUAD)
struct OrtQuadRange {
    uint32_t mis, nquads, npasses, nrays, cut_hi;
};
__device__ __forceinline__ OrtQuadRange ort_quad_range(const DevJob& J) {
    OrtQuadRange q;
    q.nrays = (uint32_t)J.nrays;
    q.mis = (uint32_t)(J.first_ray & 3);
    q.nquads = (q.nrays + q.mis + 3u) >> 2;
    q.npasses = (q.nquads + 31u) >> 5;
    q.cut_hi = (uint32_t)(J.aim_cut >> 32);
    return q;
}
template <int CAP = ORT_QUAD_QCAP, typename EdgeFn>
__device__ __forceinline__ unsigned ort_ring_quads_pass(const DevJob& J, const OrtQuadRange& Q, uint32_t b, uint2* q, int& n0,
                                                        unsigned lane, EdgeFn on_edge) {
    unsigned below;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(below));
    const uint32_t quad = b * 32u + lane;
    /* a pass at either end of the slice can hold ids outside it: counted exactly there, 128 elsewhere */
    const bool edge = (b == 0u && Q.mis != 0u) || (b * 32u + 32u >= Q.nquads);
    OrtRng g = ort_make_rng_prod(J, 0u); /* any ray of the quad names the shared block */
    {
        const unsigned long long ray = ((unsigned long long)J.first_ray & ~3ull) + 4ull * quad;
        g.r0 = (uint32_t)ray;
        g.r1 = (uint32_t)(ray >> 32);
    }
    uint32_t w[4];
    ort_shared_block(g, w);
    const int before = n0;
    unsigned nvalid = 128u;
    if (!edge) { /* every id of the pass is inside the slice */
#pragma unroll
        for (int k = 0; k < ORT_QUAD; ++k) {
            bool pass = w[k] <= Q.cut_hi;
            if (pass && w[k] == Q.cut_hi) pass = on_edge(quad * 4u + (uint32_t)k - Q.mis, w[k]);
            const unsigned m = __ballot_sync(ORT_FULL, pass);
            if (pass) {
                int p = n0 + __popc(m & below);
                ORT_ASSERT(p >= 0 && p < CAP);
                q[p] = make_uint2(w[k], quad * 4u + (uint32_t)k - Q.mis);
            }
            n0 += __popc(m);
        }
    } else {
        nvalid = 0u;
#pragma unroll
        for (int k = 0; k < ORT_QUAD; ++k) {
            const uint32_t id = quad * 4u + (uint32_t)k - Q.mis; /* wraps above nrays when it is before the slice */
            const bool valid = id < Q.nrays;
            bool pass = valid && w[k] <= Q.cut_hi;
            if (pass && w[k] == Q.cut_hi) pass = on_edge(id, w[k]);
            const unsigned m = __ballot_sync(ORT_FULL, pass);
            nvalid += __popc(__ballot_sync(ORT_FULL, valid));
            if (pass) {
                int p = n0 + __popc(m & below);
                ORT_ASSERT(p >= 0 && p < CAP);
                q[p] = make_uint2(w[k], id);
            }
            n0 += __popc(m);
        }
    }
    __syncwarp();
    return nvalid - (unsigned)(n0 - before);
}
/* up to 32 entries off the end of the queue, one per lane */
__device__ __forceinline__ bool ort_quads_pop(const uint2* q, int& n, uint2& e, unsigned lane) {
    int cnt = n < 32 ? n : 32;
    int base = n - cnt;
    bool act = (int)lane < cnt;
    if (act) {
        int p = base + lane;
        ORT_ASSERT(p >= 0 && p < ORT_QUAD_QCAP);
        e = q[p];
    }
    n = base;
    __syncwarp();
    return act;
}

/* ---- stages ------------------------------------------------------------------------------
 * A: emit.  Ring phase with the aim-plane shortcut: ort_ring_quads_pass above (69 % of the rays are aimed
 *    outside L2's aperture; the survivors carry the tested word in the queue).  Otherwise:
 *    full source, (point phase) both bottle walls, carry to L2's flat face, aperture test.
 * B: (ring shortcut: the rest of the source, then) through L2, up to and including the aperture
 *    test on L3's first surface.
 * C: the three refractions of L3, transfer to the image plane, acceptance + binning. */
template <int PHASE, int BOTTLE, int SRC, typename R>
__device__ __forceinline__ int ort_stage_a(const DevSceneT<R>& S, const DevJob& J, const OrtRng& g, uint32_t id,
                                           OrtRayT<R>& r, uint32_t& wf, uint32_t& wc, int* nevents = nullptr) {
    OrtDraws01 D;
    ort_draws01<PHASE>(g, D);
    wf = D.a[3];
    wc = D.b[3];
    int es = ort_emit<PHASE, SRC>(S, J, g, D, J.first_ray + (long long)id, r);
    if (es) return es;
    if (PHASE == ORT_PHASE_POINT) {
        if (BOTTLE == 1) {
            int st = ort_bottle_forward<false>(S, g, D, r);
            if (st) return st;
        } else if (BOTTLE == 2) {
            int st = ort_bottle_forward<true>(S, g, D, r, nevents);
            if (st) return st;
        }
    }
    return ort_l2_enter(S, r);
}
/* FROM_WORD: in the ring loop with the aim-plane shortcut the ray arrives as the word stage A tested (in wc) */
template <int PHASE, int SRC, bool FROM_WORD = true, typename R>
__device__ __forceinline__ int ort_stage_b(const DevSceneT<R>& S, const DevJob& J, const OrtRng& g, OrtRayT<R>& r,
                                           uint32_t wf, uint32_t wc) {
    if (FROM_WORD && PHASE == ORT_PHASE_RING && SRC == ORT_SRC_POINT && S.ring_shortcut) {
        uint32_t a[4], b[4];
        ort_block(g, 0u, a);
        ort_block(g, 1u, b);
        const R u2 = ort_bits_to_uniform<R>(b[0], wc), u3 = ort_word_to_uniform<R>(b[2]); /* wc: the high word stage A tested */
        wf = a[3];
        wc = b[3];
        ort_source_ring_u(S, ort_bits_to_uniform<R>(a[0], a[1]), ort_word_to_uniform<R>(a[2]), u2, u3, r);
        int st0 = ort_l2_enter(S, r); /* same arithmetic as the general path; cannot fail except
                                         within rounding of the aperture edge */
        if (st0) return st0;
    }
    int st = ort_l2_body(S, g, wf, wc, r);
    if (st) return st;
    return ort_l3_enter(S, J.iris_before != 0, r);
}
template <typename R>
__device__ __forceinline__ int ort_stage_c(const DevSceneT<R>& S, const DevJob& J, const OrtRng& g, OrtRayT<R>& r,
                                           int* xp, int* yp) {
    int st = ort_l3_body(S, g, J.iris_after != 0, r);
    if (st) return st;
    return ort_image(S, r, xp, yp);
}

/* statuses a ray can end with, per stage: the common ones are counted unconditionally, the rest
 * behind one any() */
template <int PHASE, int BOTTLE, int SRC>
__device__ __forceinline__ void ort_count_a(OrtCounts& cnt, int st) {
    if (PHASE == ORT_PHASE_RING || BOTTLE == 0) {
        ort_count<ORT_ST_L2_APERTURE>(cnt, st);
        if (SRC != ORT_SRC_POINT) ort_count_rare<ORT_ST_SOURCE_MISS>(cnt, st, st == ORT_ST_SOURCE_MISS);
    } else {
        ort_count<ORT_ST_BOTTLE_INNER_REFLECT, ORT_ST_BOTTLE_OUTER_REFLECT, ORT_ST_L2_APERTURE>(cnt, st);
        ort_count_rare<1, 2, 3, 5, 6, 7, 24, 26>(cnt, st, st > 0 && st != 4 && st != 8 && st != 9);
    }
}
__device__ __forceinline__ void ort_count_b(OrtCounts& cnt, int st) {
    ort_count<ORT_ST_L2_CURVED_REFLECT, ORT_ST_L3_S1_MISS, ORT_ST_L3_APERTURE>(cnt, st);
    ort_count_rare<ORT_ST_L2_APERTURE, ORT_ST_L2_SPHERE_MISS, ORT_ST_L3_IRIS_BEFORE>(cnt, st, st == 9 || st == 10 || st == 12);
}
__device__ __forceinline__ void ort_count_c(OrtCounts& cnt, int st) {
    ort_count<ORT_ST_BINNED, ORT_ST_L3_S1_REFLECT, ORT_ST_L3_S3_REFLECT, ORT_ST_NA_REJECT>(cnt, st);
    ort_count_rare<16, 17, 18, 20, 22, 23>(cnt, st, st == 16 || st == 17 || st == 18 || st == 20 || st >= 22);
}

template <int PHASE, int BOTTLE, int SRC, typename R>
__global__ void __launch_bounds__(ORT_TPB, ORT_MIN_BLOCKS)
ort_trace_kernel(const __grid_constant__ DevSceneT<R> S, const __grid_constant__ DevJob J,
                 unsigned long long* __restrict__ img, unsigned long long* __restrict__ counters) {
    extern __shared__ __align__(16) unsigned char ort_smem[];
    WarpShared<R>& ws = reinterpret_cast<WarpShared<R>*>(ort_smem)[threadIdx.x >> 5];
    const unsigned lane = threadIdx.x & 31u;
    const uint32_t nwarps = gridDim.x * ORT_WPB;
    const uint32_t gwarp = blockIdx.x * ORT_WPB + (threadIdx.x >> 5);
    const uint32_t nrays = (uint32_t)J.nrays;
    const uint32_t nbatches = (nrays + 31u) >> 5;

#if ORT_COUNT_SMEM
    ort_hist_clear(ws.hist, lane);
#else
    OrtCounts cnt;
#pragma unroll
    for (int k = 0; k < ORT_NSTATUS; ++k) cnt.c[k] = 0;
#endif
    int n1 = 0, n2 = 0;
    uint32_t b = gwarp;
    /* ring loop with the aim-plane shortcut: stage 0 works on quads and hands on (tested word, ray id);
     * these entries live in q0's memory */
    const bool slim = PHASE == ORT_PHASE_RING && SRC == ORT_SRC_POINT && S.ring_shortcut;
    static_assert(sizeof(WarpQueueL2<R>) >= ORT_QUAD_QCAP * sizeof(uint2), "q0 too small for the quad queue");
    uint2* const qq = reinterpret_cast<uint2*>(&ws.q0);
    const OrtQuadRange Q = ort_quad_range(J);
    const uint32_t nemit = slim ? Q.npasses : nbatches;
    for (;;) {
        int stage;
        if (n2 >= 32) stage = 2;
        else if (n1 >= 32) stage = 1;
        else if (b < nemit) stage = 0;
        else if (n2 > 0) stage = 2;
        else if (n1 > 0) stage = 1;
        else break;

        OrtRayT<R> r;
        uint32_t id = 0;
        if (PHASE == ORT_PHASE_RING && SRC == ORT_SRC_POINT && stage == 0 && slim) {
            const unsigned ended = ort_ring_quads_pass(J, Q, b, qq, n1, lane, [&](uint32_t eid, uint32_t hi) {
                /* the expression itself (ort_ring_aims_outside_aperture on the whole draw) */
                OrtRng g = ort_make_rng_prod(J, eid);
                uint32_t v[4];
                ort_block(g, 1u, v);
                return !ort_ring_aims_outside_aperture(S, ort_bits_to_uniform<R>(v[0], hi));
            });
            b += nwarps;
#if ORT_COUNT_SMEM
            if (lane == 0 && ended) atomicAdd(ws.hist + ORT_ST_L2_APERTURE, ended);
#else
            cnt.c[ORT_ST_L2_APERTURE] += ended;
#endif
        } else if (stage == 0) {
            id = b * 32u + lane;
            b += nwarps;
            int st = -1;
            uint32_t wf = 0u, wc = 0u;
            if (id < nrays) {
                OrtRng g = ort_make_rng_prod(J, id);
                if (PHASE == ORT_PHASE_POINT && BOTTLE == 2) { /* scatter events of this ray -> histogram slot 27 */
                    int nev = 0;
                    st = ort_stage_a<PHASE, BOTTLE, SRC>(S, J, g, id, r, wf, wc, &nev);
#if ORT_COUNT_SMEM
                    if (nev) atomicAdd(ws.hist + ORT_SCATTER_EVENTS_SLOT, (unsigned)nev);
#else
                    if (nev) atomicAdd(counters + ORT_SCATTER_EVENTS_SLOT, (unsigned long long)nev);
#endif
                } else {
                    st = ort_stage_a<PHASE, BOTTLE, SRC>(S, J, g, id, r, wf, wc);
                }
            }
            ort_q0_push(ws.q0, n1, st == 0, r, id, wf, wc, lane);
            __syncwarp();
#if ORT_COUNT_SMEM
            ort_tally_smem(ws.hist, st, st > 0);
#else
            ort_count_a<PHASE, BOTTLE, SRC>(cnt, st);
#endif
        } else if (stage == 1) {
            uint32_t wf = 0u, wc = 0u;
            bool act;
            if (slim) {
                uint2 e = make_uint2(0u, 0u);
                act = ort_quads_pop(qq, n1, e, lane);
                wc = e.x; /* the word stage A tested */
                id = e.y;
            } else {
                act = ort_q0_pop(ws.q0, n1, S.l2_flat_z, r, id, wf, wc, lane);
            }
            int st = -1;
            if (act) {
                OrtRng g = ort_make_rng_prod(J, id);
                st = ort_stage_b<PHASE, SRC>(S, J, g, r, wf, wc);
            }
            __syncwarp();
            ort_q_push(ws.q1, n2, st == 0, r, id, lane);
            __syncwarp();
#if ORT_COUNT_SMEM
            ort_tally_smem(ws.hist, st, st > 0);
#else
            ort_count_b(cnt, st);
#endif
        } else {
            const bool act = ort_q_pop(ws.q1, n2, r, id, lane) >= 0;
            __syncwarp();
            int st = -1, xp = 0, yp = 0;
            if (act) {
                OrtRng g = ort_make_rng_prod(J, id);
                st = ort_stage_c(S, J, g, r, &xp, &yp);
            }
            ort_bin(img, st == ORT_ST_BINNED, xp, yp, lane);
#if ORT_COUNT_SMEM
            ort_tally_smem(ws.hist, st, st >= 0);
#else
            ort_count_c(cnt, st);
#endif
        }
    }
#if ORT_COUNT_SMEM
    ort_hist_flush(ws.hist, lane, counters);
#else
    ort_counts_flush(cnt, lane, counters);
#endif
}

/* ---- point loop through a SCATTERING bottle -----------------------------------------------------
 * The reference's scatter loops (src/lens.f90:262-282, :312-333) have geometric trip counts; a warp that
 * runs them per lane waits for its longest chain (measured: 9.7 of 32 lanes busy in BASELINE config 4).
 * Here a scatter EVENT is a stage of its own with its own queue: rays that owe an event wait in qs with
 * their explicit loop state (ort_bottle_resume / ort_scatter_event), and an event pass always has 32 of
 * them.  A ray whose loop has ended walks on (the rest of the bottle, possibly into the second loop and
 * back into qs) inside the same pass; that walk is short.  Stages B and C are those of ort_trace_kernel. */
template <typename R>
struct ScatterQueue {
    R px[ORT_QCAP], py[ORT_QCAP], pz[ORT_QCAP], dx[ORT_QCAP], dy[ORT_QCAP], dz[ORT_QCAP];
    R t[ORT_QCAP], spare[ORT_QCAP];   /* pending step; odd-slot draw of the last generated block */
    uint32_t id[ORT_QCAP], next[ORT_QCAP], loop[ORT_QCAP];
};
template <typename R>
struct ScatterShared {
    ScatterQueue<R> qs;
    WarpQueueL2<R> q0;
    WarpQueue<R> q1;
    unsigned hist[ORT_NSTATUS];
};
template <typename R>
__device__ __forceinline__ void ort_qs_push(ScatterQueue<R>& q, int& n, bool alive, const OrtRayT<R>& r,
                                            const OrtScatterStateT<R>& ss, uint32_t id, unsigned lane) {
    unsigned m = __ballot_sync(ORT_FULL, alive);
    if (alive) {
        const int p = n + __popc(m & ((1u << lane) - 1u));
        ORT_ASSERT(p >= 0 && p < ORT_QCAP);
        q.px[p] = r.px; q.py[p] = r.py; q.pz[p] = r.pz;
        q.dx[p] = r.dx; q.dy[p] = r.dy; q.dz[p] = r.dz;
        q.t[p] = ss.t; q.spare[p] = ss.sr.spare;
        q.id[p] = id; q.next[p] = ss.sr.next; q.loop[p] = (uint32_t)ss.loop;
    }
    n += __popc(m);
}
template <typename R>
__device__ __forceinline__ bool ort_qs_pop(ScatterQueue<R>& q, int& n, OrtRayT<R>& r, OrtScatterStateT<R>& ss,
                                           uint32_t& id, unsigned lane) {
    int cnt = n < 32 ? n : 32;
    int base = n - cnt;
    const bool act = (int)lane < cnt;
    if (act) {
        const int p = base + lane;
        ORT_ASSERT(p >= 0 && p < ORT_QCAP);
        r.px = q.px[p]; r.py = q.py[p]; r.pz = q.pz[p];
        r.dx = q.dx[p]; r.dy = q.dy[p]; r.dz = q.dz[p];
        ss.t = q.t[p]; ss.sr.spare = q.spare[p];
        id = q.id[p]; ss.sr.next = q.next[p]; ss.loop = (int)q.loop[p];
    }
    n = base;
    return act;
}

template <int SRC, typename R>
__global__ void __launch_bounds__(ORT_TPB, 2)
ort_trace_scatter_kernel(const __grid_constant__ DevSceneT<R> S, const __grid_constant__ DevJob J,
                         unsigned long long* __restrict__ img, unsigned long long* __restrict__ counters) {
    extern __shared__ __align__(16) unsigned char ort_smem[];
    ScatterShared<R>& ws = reinterpret_cast<ScatterShared<R>*>(ort_smem)[threadIdx.x >> 5];
    const unsigned lane = threadIdx.x & 31u;
    const uint32_t nwarps = gridDim.x * ORT_WPB;
    const uint32_t gwarp = blockIdx.x * ORT_WPB + (threadIdx.x >> 5);
    const uint32_t nrays = (uint32_t)J.nrays;
    const uint32_t nbatches = (nrays + 31u) >> 5;
    ort_hist_clear(ws.hist, lane);
    int ns = 0, n1 = 0, n2 = 0;
    uint32_t b = gwarp;
    for (;;) {
        /* deepest stage with a full batch first; new rays only when nothing is full; then drain */
        int stage;
        if (n2 >= 32) stage = 3;
        else if (n1 >= 32) stage = 2;
        else if (ns >= 32) stage = 1;
        else if (b < nbatches) stage = 0;
        else if (ns > 0) stage = 1; /* events feed the later stages: drain them first */
        else if (n2 > 0) stage = 3;
        else if (n1 > 0) stage = 2;
        else break;

        OrtRayT<R> r;
        uint32_t id = 0;
        if (stage <= 1) {
            /* 0: emit and walk into the bottle; 1: one scatter event, then walk on if the loop has ended */
            OrtScatterStateT<R> ss;
            ss.t = R(0.0); ss.sr.next = 16u; ss.sr.spare = R(0.0); ss.loop = 0;
            bool act;
            if (stage == 0) {
                id = b * 32u + lane;
                b += nwarps;
                act = id < nrays;
            } else {
                act = ort_qs_pop(ws.qs, ns, r, ss, id, lane);
            }
            __syncwarp();
            int st = -1;
            uint32_t wf = 0u, wc = 0u;
            if (act) {
                OrtRng g = ort_make_rng_prod(J, id);
                int from = 0;
                if (stage == 1) {
                    bool scattered;
                    st = ort_scatter_event(S, g, r, ss, &scattered);
                    if (scattered) atomicAdd(ws.hist + ORT_SCATTER_EVENTS_SLOT, 1u);
                    from = ss.loop + 1;
                } else {
                    st = 0;
                }
                if (st == 0) {
                    OrtDraws01 D;
                    ort_draws01<ORT_PHASE_POINT>(g, D);
                    wf = D.a[3];
                    wc = D.b[3];
                    if (stage == 0) st = ort_emit<ORT_PHASE_POINT, SRC>(S, J, g, D, J.first_ray + (long long)id, r);
                    if (st == 0) st = ort_bottle_resume(S, g, D, r, ss, from);
                    if (st == 0) st = ort_l2_enter(S, r);
                }
            }
            ort_qs_push(ws.qs, ns, st == ORT_BOTTLE_EVENT, r, ss, id, lane);
            ort_q0_push(ws.q0, n1, st == 0, r, id, wf, wc, lane);
            __syncwarp();
            ort_tally_smem(ws.hist, st, st > 0);
        } else if (stage == 2) {
            uint32_t wf = 0u, wc = 0u;
            const bool act = ort_q0_pop(ws.q0, n1, S.l2_flat_z, r, id, wf, wc, lane);
            int st = -1;
            if (act) {
                OrtRng g = ort_make_rng_prod(J, id);
                st = ort_stage_b<ORT_PHASE_POINT, SRC>(S, J, g, r, wf, wc);
            }
            __syncwarp();
            ort_q_push(ws.q1, n2, st == 0, r, id, lane);
            __syncwarp();
            ort_tally_smem(ws.hist, st, st > 0);
        } else {
            const bool act = ort_q_pop(ws.q1, n2, r, id, lane) >= 0;
            __syncwarp();
            int st = -1, xp = 0, yp = 0;
            if (act) {
                OrtRng g = ort_make_rng_prod(J, id);
                st = ort_stage_c(S, J, g, r, &xp, &yp);
            }
            ort_bin(img, st == ORT_ST_BINNED, xp, yp, lane);
            ort_tally_smem(ws.hist, st, st >= 0);
        }
    }
    ort_hist_flush(ws.hist, lane, counters);
}

/* ---- ring loop with the fp32 culling filter (ortf_filter) ---------------------------------
 * Two kernels per slice of the ray range, because 99 % of the ring rays never need fp64 and a
 * kernel that also contained the fp64 stages would carry their registers and code next to the
 * filter's (which fills its own 79 at two rays per lane):
 *
 *   ort_ring_cull_kernel       integer + fp32 only, 79 registers, 3 blocks (24 warps) per SM.
 *       A: one Philox block per FOUR rays (ort_ring_quads_pass): the aim-point aperture test on the high
 *          word of the raw draw (aim_cut, see ort_ring_aim_cut) ends 69 % of the rays;
 *       F: the single-precision filter on the compacted survivors, 64 per pass -- two per lane, in
 *          packed f32x2 arithmetic (ortf_filter<OrtfTwo>); the rays it can call are counted,
 *          the others (they reach L3, or a decision was too close) are appended to a list of ray
 *          indices in global memory -- 4 bytes per listed ray, ~0.2 % .. 4 % of the rays.
 *   ort_ring_survivors_kernel  the ordinary fp64 stages B and C over that list (the draws are
 *       regenerated from the ray index).
 *
 * Used when the scene has ring_shortcut, precision is 64 and ORT_FLAG_NO_FILTER is not set; the
 * results are identical to ort_trace_kernel's.  VERIFY (ORT_FLAG_VERIFY_FILTER): F lists every ray
 * (with the verdict of its filter above the ray index) and the survivors kernel compares that verdict with
 * what fp64 finds:
 * counters[ORT_FILTER_SLOT_CALLED] = rays the filter called, counters[ORT_FILTER_SLOT_WRONG] =
 * calls that disagree with fp64 (must stay 0). */
#ifndef ORT_CULL_TALLY_SMEM
#define ORT_CULL_TALLY_SMEM 1 /* statuses counted in the warp's shared memory (one ATOMS per ray, LSU pipe) instead of five
                                 compare + predicated-add pairs per ray on the ALU pipe, which bounds this kernel: +1 %,
                                 and the five counter registers end the spills at 80 registers */
#endif
#define ORT_CULL_QCAP (64 + 32 * ORT_QUAD) /* < 64 leftovers + the survivors of one stage-A pass */
struct SlimQueue {
    uint2 e[ORT_CULL_QCAP]; /* x: high word of the aim-disc r^2 draw (what stage A tested), y: ray index */
    uint32_t hb[128];       /* list entries on their way to global memory */
#if ORT_CULL_TALLY_SMEM
    unsigned hist[ORT_NSTATUS]; /* this warp's histogram of the statuses the filter proved */
#endif
};
/* VERIFY: a list entry carries the filter's verdict above the ray index (a slice has <= 2^29 rays) */
#define ORT_LIST_ID_BITS 29
#define ORT_LIST_ID_MASK ((1u << ORT_LIST_ID_BITS) - 1u)

/* `count` (<= 32, warp-uniform) entries, one per lane, to the end of the global list */
__device__ __forceinline__ void ort_list_append(uint32_t* __restrict__ list, unsigned* __restrict__ nlist,
                                                unsigned capacity, unsigned long long* __restrict__ counters,
                                                uint32_t value, unsigned count, unsigned lane) {
    unsigned base = 0;
    if (lane == 0) base = atomicAdd(nlist, count);
    base = __shfl_sync(ORT_FULL, base, 0);
    if (lane < count) {
        if (base + lane < capacity) list[base + lane] = value;
        else atomicAdd(counters + ORT_FILTER_SLOT_OVERFLOW, 1ull);
    }
}

template <int K>
__device__ __forceinline__ void ort_tally(unsigned& c, int st) {
    asm("{\n\t.reg .pred p;\n\tsetp.eq.s32 p, %1, %2;\n\t@p add.u32 %0, %0, 1;\n\t}" : "+r"(c) : "r"(st), "n"(K));
}

#ifndef ORT_CULL_MIN_BLOCKS
#define ORT_CULL_MIN_BLOCKS 3 /* 80 registers: two rays per lane; 4 blocks (64 registers) spill and are slower */
#endif
/* list[0 .. *nlist) receives the ray indices (relative to J.first_ray) handed to fp64; entries that
 * would not fit in `capacity` are counted in counters[ORT_FILTER_SLOT_OVERFLOW] instead, which
 * ort_trace reports as an error (the launcher sizes the list 16 sigma above its expectation).
 * A filter pass takes 64 rays, TWO per lane (ortf_filter<OrtfTwo>: packed f32x2 arithmetic). */
template <bool VERIFY>
__global__ void __launch_bounds__(ORT_TPB, ORT_CULL_MIN_BLOCKS)
ort_ring_cull_kernel(const __grid_constant__ OrtfParamsT<OrtfV2> K, const __grid_constant__ DevJob J,
                     uint32_t* __restrict__ list, unsigned* __restrict__ nlist, const unsigned capacity,
                     unsigned long long* __restrict__ counters) {
    extern __shared__ __align__(16) unsigned char ort_smem[];
    SlimQueue& q0 = reinterpret_cast<SlimQueue*>(ort_smem)[threadIdx.x >> 5];
    const unsigned lane = threadIdx.x & 31u;
    const uint32_t nwarps = gridDim.x * ORT_WPB;
    const uint32_t gwarp = blockIdx.x * ORT_WPB + (threadIdx.x >> 5);
    const OrtQuadRange Q = ort_quad_range(J);
    unsigned c9 = 0, c10 = 0, c11 = 0, c12 = 0, c13 = 0, c14 = 0;
#if ORT_CULL_TALLY_SMEM
    ort_hist_clear(q0.hist, lane);
#endif
    int nh = 0;           /* entries parked in q0.hb */
    unsigned below;       /* lanes below this one */
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(below));
    int n0 = 0;
    uint32_t b = gwarp;
    for (;;) {
        const bool emit = n0 < 64 && b < Q.npasses;
        if (!emit && n0 == 0) break;
        if (emit) {
            /* a draw whose high word EQUALS the cut's goes on: it sits on the aperture edge, where the
             * filter hands it to fp64, and ort_l2_enter there makes the exact call */
            c9 += ort_ring_quads_pass<ORT_CULL_QCAP>(J, Q, b, q0.e, n0, lane, [](uint32_t, uint32_t) { return true; });
            b += nwarps;
        } else {
            const int cnt = n0 < 64 ? n0 : 64;
            const int base = n0 - cnt;
            const bool act0 = (int)lane < cnt, act1 = (int)lane + 32 < cnt;
            uint2 e0 = make_uint2(65536u, 0u), e1 = make_uint2(65536u, 0u);
            ORT_ASSERT(base >= 0 && base + cnt <= ORT_CULL_QCAP);
            if (act0) e0 = q0.e[base + lane];
            if (act1) e1 = q0.e[base + 32 + lane];
            n0 = base;
            __syncwarp();
            int sa, sb;
            {
                /* blocks 0 and 1 of both rays (an idle half runs on ray 0: its result is dropped) */
                uint32_t a0[4], b0[4], a1[4], b1[4];
                const OrtRng g0 = ort_make_rng_prod(J, e0.y), g1 = ort_make_rng_prod(J, e1.y);
                ort_block(g0, 0u, a0);
                ort_block(g0, 1u, b0);
                ort_block(g1, 0u, a1);
                ort_block(g1, 1u, b1);
                const OrtfS2 st = ortf_filter<OrtfTwo>(K, OrtfW2{a0[1], a1[1]}, OrtfW2{a0[2], a1[2]}, OrtfW2{a0[3], a1[3]},
                                                       OrtfW2{b0[2], b1[2]}, OrtfW2{b0[3], b1[3]}, OrtfW2{e0.x, e1.x});
                sa = st.a;
                sb = st.b;
            }
            sa = act0 ? sa : -1;
            sb = act1 ? sb : -1;
#if ORT_CULL_TALLY_SMEM
            if (!VERIFY) {
                ort_tally_smem(q0.hist, sa, sa > 0);
                ort_tally_smem(q0.hist, sb, sb > 0);
            }
#else
            if (!VERIFY) {
                /* one compare + one predicated add per status (left to itself the compiler builds
                 * add / conditional move / move triples here) */
                ort_tally<ORT_ST_L2_SPHERE_MISS>(c10, sa);
                ort_tally<ORT_ST_L2_CURVED_REFLECT>(c11, sa);
                ort_tally<ORT_ST_L3_IRIS_BEFORE>(c12, sa);
                ort_tally<ORT_ST_L3_S1_MISS>(c13, sa);
                ort_tally<ORT_ST_L3_APERTURE>(c14, sa);
                ort_tally<ORT_ST_L2_SPHERE_MISS>(c10, sb);
                ort_tally<ORT_ST_L2_CURVED_REFLECT>(c11, sb);
                ort_tally<ORT_ST_L3_IRIS_BEFORE>(c12, sb);
                ort_tally<ORT_ST_L3_S1_MISS>(c13, sb);
                ort_tally<ORT_ST_L3_APERTURE>(c14, sb);
            }
#endif
            /* rays for fp64 (VERIFY: every ray, with the verdict above its index): parked in a warp-private
             * buffer and written to the list 32 at a time (one atomic and one coalesced store per 32 entries) */
            const bool l0 = VERIFY ? act0 : sa == 0, l1 = VERIFY ? act1 : sb == 0;
            const unsigned m0 = __ballot_sync(ORT_FULL, l0), m1 = __ballot_sync(ORT_FULL, l1);
            if (l0) {
                int p = nh + __popc(m0 & below);
                ORT_ASSERT(p >= 0 && p < 128);
                q0.hb[p] = VERIFY ? (e0.y | ((uint32_t)(sa > 0 ? sa - 9 : 0) << ORT_LIST_ID_BITS)) : e0.y;
            }
            nh += __popc(m0);
            if (l1) {
                int p = nh + __popc(m1 & below);
                ORT_ASSERT(p >= 0 && p < 128);
                q0.hb[p] = VERIFY ? (e1.y | ((uint32_t)(sb > 0 ? sb - 9 : 0) << ORT_LIST_ID_BITS)) : e1.y;
            }
            nh += __popc(m1);
            __syncwarp();
            while (nh >= 32) {
                nh -= 32;
                ort_list_append(list, nlist, capacity, counters, q0.hb[nh + lane], 32u, lane);
                __syncwarp();
            }
        }
    }
    if (nh > 0) ort_list_append(list, nlist, capacity, counters, (int)lane < nh ? q0.hb[lane] : 0u, (unsigned)nh, lane);
#if ORT_CULL_TALLY_SMEM
    if (lane == 0 && c9) atomicAdd(q0.hist + ORT_ST_L2_APERTURE, c9);
    ort_hist_flush(q0.hist, lane, counters);
    return;
#endif
    /* per-lane tallies -> one atomic per status and warp */
    c10 = __reduce_add_sync(ORT_FULL, c10);
    c11 = __reduce_add_sync(ORT_FULL, c11);
    c12 = __reduce_add_sync(ORT_FULL, c12);
    c13 = __reduce_add_sync(ORT_FULL, c13);
    c14 = __reduce_add_sync(ORT_FULL, c14);
    unsigned mine = lane == 9 ? c9 : lane == 10 ? c10 : lane == 11 ? c11 : lane == 12 ? c12 : lane == 13 ? c13 : lane == 14 ? c14 : 0u;
    if (mine) atomicAdd(counters + lane, (unsigned long long)mine);
}

/* fp64 stages B and C over the listed rays */
struct SurvShared {
    WarpQueue<double> q;
    unsigned hist[ORT_NSTATUS];
};
template <bool VERIFY>
__global__ void __launch_bounds__(ORT_TPB, ORT_MIN_BLOCKS)
ort_ring_survivors_kernel(const __grid_constant__ DevSceneT<double> S, const __grid_constant__ DevJob J, const uint32_t* __restrict__ list,
                          const unsigned* __restrict__ nlist, const unsigned capacity,
                          unsigned long long* __restrict__ img, unsigned long long* __restrict__ counters) {
    extern __shared__ __align__(16) unsigned char ort_smem[];
    SurvShared& ws = reinterpret_cast<SurvShared*>(ort_smem)[threadIdx.x >> 5];
    WarpQueue<double>& q = ws.q;
    const unsigned lane = threadIdx.x & 31u;
    const uint32_t nwarps = gridDim.x * ORT_WPB;
    const uint32_t gwarp = blockIdx.x * ORT_WPB + (threadIdx.x >> 5);
    const uint32_t total = *nlist < capacity ? *nlist : capacity;
    const uint32_t nbatches = (total + 31u) >> 5;

    ort_hist_clear(ws.hist, lane);
    int n2 = 0;
    uint32_t b = gwarp;
    for (;;) {
        const bool take = n2 < 32 && b < nbatches;
        if (!take && n2 == 0) break;
        OrtRayT<double> r;
        uint32_t id = 0;
        if (take) {
            uint32_t i = b * 32u + lane;
            b += nwarps;
            int st = -1, verdict = 0;
            if (i < total) {
                id = list[i];
                if (VERIFY) { /* the verdict the cull kernel's filter reached, above the ray index */
                    const uint32_t code = id >> ORT_LIST_ID_BITS;
                    verdict = code ? (int)code + 9 : 0;
                    id &= ORT_LIST_ID_MASK;
                }
                OrtRng g = ort_make_rng_prod(J, id);
                const uint32_t hi = ort_aim_hi(g);
                r.px = r.py = r.pz = r.dx = r.dy = r.dz = 0.0;
                st = ort_stage_b<ORT_PHASE_RING, ORT_SRC_POINT>(S, J, g, r, 0u, hi);
            }
            if (VERIFY) {
                ort_tally_smem(ws.hist, ORT_FILTER_SLOT_CALLED, verdict > 0);
                ort_tally_smem(ws.hist, ORT_FILTER_SLOT_WRONG, verdict > 0 && verdict != st);
            }
            ort_q_push(q, n2, st == 0, r, id, lane);
            __syncwarp();
            ort_tally_smem(ws.hist, st, st > 0);
        } else {
            const bool act = ort_q_pop(q, n2, r, id, lane) >= 0;
            __syncwarp();
            int st = -1, xp = 0, yp = 0;
            if (act) {
                OrtRng g = ort_make_rng_prod(J, id);
                st = ort_stage_c(S, J, g, r, &xp, &yp);
            }
            ort_bin(img, st == ORT_ST_BINNED, xp, yp, lane);
            ort_tally_smem(ws.hist, st, st >= 0);
        }
    }
    ort_hist_flush(ws.hist, lane, counters);
}

/* The same path, one thread per ray from source to detector, no compaction: every early exit
 * leaves its lane idle until the slowest lane of the warp is done.  Kept to measure what the
 * compaction buys (warp execution efficiency in ncu) and as a cross-check of the megakernel. */
template <int PHASE, int BOTTLE, typename R>
__global__ void __launch_bounds__(ORT_TPB, ORT_MIN_BLOCKS)
ort_trace_flat_kernel(const __grid_constant__ DevSceneT<R> S, const __grid_constant__ DevJob J,
                      unsigned long long* __restrict__ img, unsigned long long* __restrict__ counters) {
    const unsigned lane = threadIdx.x & 31u;
    const uint32_t nwarps = gridDim.x * ORT_WPB;
    const uint32_t gwarp = blockIdx.x * ORT_WPB + (threadIdx.x >> 5);
    const uint32_t nrays = (uint32_t)J.nrays;
    const uint32_t nbatches = (nrays + 31u) >> 5;
    OrtCounts cnt;
#pragma unroll
    for (int k = 0; k < ORT_NSTATUS; ++k) cnt.c[k] = 0;
    for (uint32_t b = gwarp; b < nbatches; b += nwarps) {
        uint32_t id = b * 32u + lane;
        int st = -1, xp = 0, yp = 0;
        if (id < nrays) {
            OrtRng g = ort_make_rng_prod(J, id);
            OrtRayT<R> r;
            uint32_t wf, wc;
            int nev = 0;
            st = ort_stage_a<PHASE, BOTTLE, ORT_SRC_POINT>(S, J, g, id, r, wf, wc, &nev);
            if (nev) atomicAdd(counters + ORT_SCATTER_EVENTS_SLOT, (unsigned long long)nev);
            if (st == 0) st = ort_stage_b<PHASE, ORT_SRC_POINT, false>(S, J, g, r, wf, wc);
            if (st == 0) st = ort_stage_c(S, J, g, r, &xp, &yp);
        }
        ort_bin(img, st == ORT_ST_BINNED, xp, yp, lane);
        ort_count<0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24>(cnt, st);
    }
    ort_counts_flush(cnt, lane, counters);
}

/* Explicit ray list; SoA in/out, see ort_trace_rays in include/ort.h */
template <typename R>
__global__ void __launch_bounds__(ORT_TPB)
ort_rays_kernel(const __grid_constant__ DevSceneT<R> S, const __grid_constant__ DevJob J,
                const double* __restrict__ pin, const double* __restrict__ din, double* __restrict__ pout,
                double* __restrict__ dout, int32_t* __restrict__ status, int32_t* __restrict__ bin, long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    OrtRng g = ort_make_rng(J, (uint32_t)i);
    OrtRayT<R> r;
    int xp = INT32_MIN, yp = INT32_MIN, x = 0, y = 0;
    const bool have_input = pin != nullptr;
    if (have_input) { /* the ABI stays double; the fp32 variant rounds the start state once */
        r.px = (R)pin[i]; r.py = (R)pin[n + i]; r.pz = (R)pin[2 * n + i];
        r.dx = (R)din[i]; r.dy = (R)din[n + i]; r.dz = (R)din[2 * n + i];
    }
    int st = ort_full_path(S, J, g, have_input, r, &x, &y);
    if (st == ORT_ST_BINNED) { xp = x; yp = y; }
    pout[i] = r.px; pout[n + i] = r.py; pout[2 * n + i] = r.pz;
    dout[i] = r.dx; dout[n + i] = r.dy; dout[2 * n + i] = r.dz;
    status[i] = st;
    bin[i] = xp;
    bin[n + i] = yp;
}

/* Volume image (makeImage3D, src/imageMod.f90:61-90): one thread per ray through the whole path up
 * to L3 (J.stop_after = ORT_STOP_L3), then the transfer to the image plane and the march through
 * ORT_VOL_DEPTH depths with one 32-bit RED per sample.  A diagnostic path: no compaction, the
 * status histogram goes through shared memory. */
__global__ void __launch_bounds__(ORT_TPB)
ort_volume_kernel(const __grid_constant__ DevSceneT<double> S, const __grid_constant__ DevJob J, const double dzs,
                  unsigned* __restrict__ vol, unsigned long long* __restrict__ counters) {
    __shared__ unsigned hist[ORT_NSTATUS];
    if (threadIdx.x < ORT_NSTATUS) hist[threadIdx.x] = 0;
    __syncthreads();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < J.nrays;
         i += (long long)gridDim.x * blockDim.x) {
        OrtRng g = ort_make_rng_prod(J, (uint32_t)i);
        OrtRayT<double> r;
        int x = 0, y = 0;
        int st = ort_full_path(S, J, g, false, r, &x, &y);
        if (st == ORT_ST_STOPPED) {
            double d = ort_div_z(S.img_z - r.pz, r.dz);
            ort_advance(r, d);
            int hits = 0;
            for (int k = 0; k < ORT_VOL_DEPTH; ++k) {
                const double t = k * dzs;
                const double fx = floor(fma(r.dx, t, r.px) * S.inv_binwid), fy = floor(fma(r.dy, t, r.py) * S.inv_binwid);
                if (!(fabs(fx) <= 200.0) || !(fabs(fy) <= 200.0)) break;
                const size_t idx = ((size_t)k * ORT_IMG_N + (size_t)((int)fy + ORT_IMG_HALF)) * ORT_IMG_N +
                                   (size_t)((int)fx + ORT_IMG_HALF);
                ORT_ASSERT(idx < (size_t)ORT_VOL_DEPTH * ORT_IMG_BINS);
                atomicAdd(vol + idx, 1u);
                ++hits;
            }
            st = hits ? ORT_ST_BINNED : ORT_ST_OFF_DETECTOR;
        }
        atomicAdd(&hist[st], 1u);
    }
    __syncthreads();
    if (threadIdx.x < ORT_NSTATUS && hist[threadIdx.x])
        atomicAdd(counters + threadIdx.x, (unsigned long long)hist[threadIdx.x]);
}

__global__ void ort_uniforms_kernel(uint64_t seed, int32_t phase, int64_t ray, int32_t first_slot, int32_t n,
                                    double* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    OrtRng g;
    g.k0 = (uint32_t)seed; g.k1 = (uint32_t)(seed >> 32);
    g.rk = nullptr;
    g.r0 = (uint32_t)(uint64_t)ray; g.r1 = (uint32_t)((uint64_t)ray >> 32);
    g.phase = (uint32_t)phase;
    g.override_u = -1.0;
    out[i] = ort_slot<double>(g, (uint32_t)(first_slot + i));
}

/* FP64 FMA peak: 8 independent DFMA chains per thread, ITERS x 8 x 2 flops per thread */
#define ORT_PEAK_ITERS 16384
__global__ void __launch_bounds__(256) ort_dfma_peak_kernel(double* __restrict__ out, double seed,
                                                            unsigned long long* __restrict__ cycles) {
    double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5,
           a6 = a0 + 6, a7 = a0 + 7;
    const double m = 0.999999, c = 1e-9;
    long long t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < ORT_PEAK_ITERS; ++i) {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
    long long t1 = clock64();
    double s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (s == 123.456) out[0] = s; /* keep the chains alive */
    /* all blocks are resident from the start (8 x 256 threads per SM), so the longest-lived
     * block spans the whole kernel: max(t1 - t0) / kernel time = SM clock under this load */
    if (threadIdx.x == 0) atomicMax(cycles, (unsigned long long)(t1 - t0));
}

/* max error, in units of the last place, of ort_rcp / ort_div / ort_sqrt / ort_rsqrt against the
 * correctly rounded operators, over log-uniform operands in [2^-40, 2^40) */
__device__ __forceinline__ unsigned long long ort_ulp_diff(double a, double b) {
    long long ia = __double_as_longlong(a), ib = __double_as_longlong(b);
    long long d = ia - ib;
    return (unsigned long long)(d < 0 ? -d : d);
}
__global__ void ort_math_selftest_kernel(long long n, unsigned long long* __restrict__ worst) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    OrtRng g;
    g.k0 = 0x5eedu; g.k1 = 0; g.rk = nullptr; g.r0 = (uint32_t)i; g.r1 = (uint32_t)(i >> 32); g.phase = 7; g.override_u = -1.0;
    double u0, u1, u2, u3;
    ort_draw2(g, 8, &u0, &u1);
    ort_draw2(g, 9, &u2, &u3);
    double x = exp2(80.0 * u0 - 40.0) * (1.0 + u1);
    double a = (exp2(80.0 * u2 - 40.0) * (1.0 + u3)) * ((i & 1) ? -1.0 : 1.0);
    atomicMax(worst + 0, ort_ulp_diff(ort_rcp(x), 1.0 / x));
    atomicMax(worst + 1, ort_ulp_diff(ort_div(a, x), a / x));
    atomicMax(worst + 2, ort_ulp_diff(ort_sqrt(x), sqrt(x)));
    atomicMax(worst + 3, ort_ulp_diff(ort_rsqrt(x), 1.0 / sqrt(x)));
}

/* Rule R4 of the ring filter's error bound (ort_filter.cuh): the largest error of the MUFU
 * approximations it uses, over EVERY fp32 argument -- all 2^32 bit patterns are tried.
 *   worst[0..2]  relative error of rcp / rsqrt / sqrt.approx.ftz.f32, |x| (x > 0 for the roots) in [2^-64, 2^64]
 *   worst[3..4]  absolute error of sin / cos.approx.ftz.f32, |x| <= 3.1416 (ortf_sincos_word: the angle is in [-pi, pi))
 * as the bit patterns of non-negative doubles (which order like integers), against fp64 references. */
__global__ void ort_mufu_selftest_kernel(unsigned long long* __restrict__ worst) {
    double w0 = 0.0, w1 = 0.0, w2 = 0.0, w3 = 0.0, w4 = 0.0;
    const unsigned long long total = 1ull << 32;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const float x = __uint_as_float((unsigned)i);
        const float ax = fabsf(x);
        if (!(ax == ax) || ax == INFINITY) continue;
        if (ax >= 5.421010862427522e-20f && ax <= 1.8446744073709552e19f) {
            const double xd = (double)x;
            w0 = fmax(w0, fabs((double)ortf_rcp(x) * xd - 1.0));
            if (x > 0.0f) {
                const double r = sqrt(xd);
                w1 = fmax(w1, fabs((double)ortf_rsqrt(x) * r - 1.0));
                w2 = fmax(w2, fabs((double)ortf_sqrt(x) / r - 1.0));
            }
        }
        if (ax <= 3.1416f) {
            float s, c;
            asm("sin.approx.ftz.f32 %0, %1;" : "=f"(s) : "f"(x));
            asm("cos.approx.ftz.f32 %0, %1;" : "=f"(c) : "f"(x));
            w3 = fmax(w3, fabs((double)s - sin((double)x)));
            w4 = fmax(w4, fabs((double)c - cos((double)x)));
        }
    }
    atomicMax(worst + 0, (unsigned long long)__double_as_longlong(w0));
    atomicMax(worst + 1, (unsigned long long)__double_as_longlong(w1));
    atomicMax(worst + 2, (unsigned long long)__double_as_longlong(w2));
    atomicMax(worst + 3, (unsigned long long)__double_as_longlong(w3));
    atomicMax(worst + 4, (unsigned long long)__double_as_longlong(w4));
}

#endif /* ORT_KERNELS_CUH */
