/*
 * ort_cuda.cu -- device half of the C-ABI declared in include/ort.h: device / stream / NCCL
 * lifetime, scene flattening, kernel launches, the image reduce and the copies.
 *
 * There is no CPU fallback in this file or anywhere in the library: with no CUDA device every
 * compute entry point fails with ORT_ENODEVICE.
 */
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h> /* types and prototypes only: NCCL is dlopen()ed, never linked */

#include <chrono>
#include <cstdlib>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <thread>
#include <utility>
#include <vector>

#include "ort_internal.h"
#include "ort_flatten.h"
#include "ort_kernels.cuh"

/* ------------------------------------------------------------------------------------------
 * error text
 * ---------------------------------------------------------------------------------------- */
static char g_err[1024] = "";
void ort_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}
extern "C" const char* ort_last_error(void) { return g_err; }

#define CK(call)                                                                          \
    do {                                                                                  \
        cudaError_t e_ = (call);                                                          \
        if (e_ != cudaSuccess) {                                                          \
            ort_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
            return (e_ == cudaErrorNoDevice || e_ == cudaErrorInsufficientDriver) ? ORT_ENODEVICE : ORT_ECUDA; \
        }                                                                                 \
    } while (0)

/* ------------------------------------------------------------------------------------------
 * NCCL, loaded at run time so the library has no link-time dependency on it and a process that
 * already carries a libnccl.so.2 (e.g. a torch.distributed launcher) shares that copy.
 * ---------------------------------------------------------------------------------------- */
struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Reduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t,
                           cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi g_nccl;

static int nccl_load() {
    if (g_nccl.handle) return ORT_OK;
    const char* names[] = {"libnccl.so.2", "libnccl.so", "/usr/lib/x86_64-linux-gnu/libnccl.so.2"};
    void* h = nullptr;
    for (const char* n : names) {
        h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) {
        ort_set_error("NCCL not loadable: %s", dlerror());
        return ORT_ENCCL;
    }
#define SYM(field, name)                                              \
    *(void**)(&g_nccl.field) = dlsym(h, name);                        \
    if (!g_nccl.field) {                                              \
        ort_set_error("NCCL symbol %s missing", name);                \
        return ORT_ENCCL;                                             \
    }
    SYM(GetUniqueId, "ncclGetUniqueId")
    SYM(CommInitRank, "ncclCommInitRank")
    SYM(CommInitAll, "ncclCommInitAll")
    SYM(CommDestroy, "ncclCommDestroy")
    SYM(Reduce, "ncclReduce")
    SYM(GroupStart, "ncclGroupStart")
    SYM(GroupEnd, "ncclGroupEnd")
    SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
    g_nccl.handle = h;
    return ORT_OK;
}
#define NK(call)                                                                              \
    do {                                                                                      \
        ncclResult_t r_ = (call);                                                             \
        if (r_ != ncclSuccess) {                                                              \
            ort_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, g_nccl.GetErrorString(r_)); \
            return ORT_ENCCL;                                                                 \
        }                                                                                     \
    } while (0)

/* ------------------------------------------------------------------------------------------
 * library state
 * ---------------------------------------------------------------------------------------- */
/* One in-order lane of launches.  A batched call with few rays per scene (a quick-look sweep) spreads its
 * scenes over ORT_LANES lanes so that the tail of one scene's persistent kernel is filled by the next
 * scene's blocks and launch latencies overlap; large jobs use lane 0 alone. */
#define ORT_LANES 4
struct Lane {
    cudaStream_t stream = nullptr;    /* lane 0: the device's main stream */
    cudaStream_t stream2 = nullptr;   /* the ring loop's fp64 survivors kernels run here, behind the cull kernels */
    cudaEvent_t ev_cull[2] = {nullptr, nullptr}, ev_surv[2] = {nullptr, nullptr}, ev_done = nullptr;
    uint32_t* d_list = nullptr;       /* ring loop: ray indices the fp32 filter hands to fp64, two
                                         buffers of list_cap entries used alternately */
    size_t list_cap = 0;
    unsigned* d_nlist = nullptr;      /* ... and the length of that list, one slot per slice */
    size_t nlist_cap = 0;
};
struct DeviceCtx {
    int dev = -1;
    int num_sms = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_start = nullptr, ev_traced = nullptr, ev_reduced = nullptr, ev_copied = nullptr, ev_zeroed = nullptr;
    unsigned long long* d_buf = nullptr; /* [nscenes*BINS image][nscenes*NSTATUS counters] */
    size_t d_elems = 0;
    ncclComm_t comm = nullptr;
    long long* d_image_cdf = nullptr; /* image source: prefix sums of the ray budget */
    Lane lanes[ORT_LANES];
    /* resident blocks per SM of every kernel launched so far (the attribute set-up and the
     * occupancy query are done once per kernel and device, not once per scene of a batch) */
    std::vector<std::pair<const void*, int>> occ;
    /* grow-only device scratch of the entry points beside ort_trace (explicit ray lists, the
     * volume image, the generator's test entry): kept until ort_finalize instead of a
     * cudaMalloc / cudaFree pair per call */
    void* scratch[3] = {nullptr, nullptr, nullptr};
    size_t scratch_bytes[3] = {0, 0, 0};
};
struct LibState {
    bool ready = false;
    bool rank_mode = false;
    int rank = 0, nranks = 1;
    std::vector<DeviceCtx> devs;
    std::vector<long long> image_cdf; /* host copy, re-uploaded when devices are (re)opened */
    unsigned long long* h_pinned = nullptr;
    size_t h_elems = 0;
};
static LibState g;

int ort_internal_primary_device(void) { return (g.ready && !g.devs.empty()) ? g.devs[0].dev : -1; }

extern "C" int ort_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

static int ctx_open(DeviceCtx& c, int dev) {
    c.dev = dev;
    CK(cudaSetDevice(dev));
    CK(cudaDeviceGetAttribute(&c.num_sms, cudaDevAttrMultiProcessorCount, dev));
    CK(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
    CK(cudaEventCreate(&c.ev_start));
    CK(cudaEventCreate(&c.ev_traced));
    CK(cudaEventCreate(&c.ev_reduced));
    CK(cudaEventCreate(&c.ev_copied));
    CK(cudaEventCreateWithFlags(&c.ev_zeroed, cudaEventDisableTiming));
    for (int l = 0; l < ORT_LANES; ++l) {
        Lane& L = c.lanes[l];
        if (l == 0) L.stream = c.stream;
        else CK(cudaStreamCreateWithFlags(&L.stream, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&L.stream2, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) {
            CK(cudaEventCreateWithFlags(&L.ev_cull[i], cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&L.ev_surv[i], cudaEventDisableTiming));
        }
        CK(cudaEventCreateWithFlags(&L.ev_done, cudaEventDisableTiming));
    }
    return ORT_OK;
}

/* resident blocks per SM of `fn` with `smem` bytes of dynamic shared memory (cached) */
static int ctx_occupancy(DeviceCtx& c, const void* fn, size_t smem, int* occ) {
    for (auto& e : c.occ)
        if (e.first == fn) {
            *occ = e.second;
            return ORT_OK;
        }
    if (smem) CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int o = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, fn, ORT_TPB, smem));
    if (o < 1) o = 1;
    c.occ.emplace_back(fn, o);
    *occ = o;
    return ORT_OK;
}
/* grow-only scratch buffer `slot` of at least `bytes` */
static int ctx_scratch(DeviceCtx& c, int slot, size_t bytes, void** out) {
    if (c.scratch_bytes[slot] < bytes) {
        if (c.scratch[slot]) CK(cudaFree(c.scratch[slot]));
        c.scratch[slot] = nullptr;
        c.scratch_bytes[slot] = 0;
        CK(cudaMalloc(&c.scratch[slot], bytes));
        c.scratch_bytes[slot] = bytes;
    }
    *out = c.scratch[slot];
    return ORT_OK;
}

extern "C" int ort_finalize(void) {
    for (auto& c : g.devs) {
        if (c.dev < 0) continue;
        cudaSetDevice(c.dev);
        if (c.comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c.comm);
        if (c.d_buf) cudaFree(c.d_buf);
        if (c.d_image_cdf) cudaFree(c.d_image_cdf);
        for (int l = 0; l < ORT_LANES; ++l) {
            Lane& L = c.lanes[l];
            if (L.d_list) cudaFree(L.d_list);
            if (L.d_nlist) cudaFree(L.d_nlist);
            for (int i = 0; i < 2; ++i) {
                if (L.ev_cull[i]) cudaEventDestroy(L.ev_cull[i]);
                if (L.ev_surv[i]) cudaEventDestroy(L.ev_surv[i]);
            }
            if (L.ev_done) cudaEventDestroy(L.ev_done);
            if (L.stream2) cudaStreamDestroy(L.stream2);
            if (l > 0 && L.stream) cudaStreamDestroy(L.stream);
        }
        if (c.ev_zeroed) cudaEventDestroy(c.ev_zeroed);
        for (void* p : c.scratch)
            if (p) cudaFree(p);
        if (c.ev_start) cudaEventDestroy(c.ev_start);
        if (c.ev_traced) cudaEventDestroy(c.ev_traced);
        if (c.ev_reduced) cudaEventDestroy(c.ev_reduced);
        if (c.ev_copied) cudaEventDestroy(c.ev_copied);
        if (c.stream) cudaStreamDestroy(c.stream);
    }
    if (g.h_pinned) cudaFreeHost(g.h_pinned);
    std::vector<long long> keep;
    keep.swap(g.image_cdf); /* the image source outlives re-initialisation */
    g = LibState();
    g.image_cdf.swap(keep);
    return ORT_OK;
}

/* The first collective on a new communicator builds its channels (tens of milliseconds per peer): it is done
 * here, on the buffer the first job will reduce, so that no job's reduce time contains it.  Collective:
 * every rank is inside ort_init / ort_init_rank when this runs. */
static int nccl_warm_up() {
    const size_t elems = (size_t)ORT_IMG_BINS + ORT_NSTATUS;
    for (auto& c : g.devs) {
        CK(cudaSetDevice(c.dev));
        if (c.d_elems < elems) {
            if (c.d_buf) cudaFree(c.d_buf);
            c.d_buf = nullptr;
            CK(cudaMalloc(&c.d_buf, elems * sizeof(unsigned long long)));
            c.d_elems = elems;
        }
        CK(cudaMemsetAsync(c.d_buf, 0, elems * sizeof(unsigned long long), c.stream));
    }
    NK(g_nccl.GroupStart());
    for (auto& c : g.devs) NK(g_nccl.Reduce(c.d_buf, c.d_buf, elems, ncclUint64, ncclSum, 0, c.comm, c.stream));
    NK(g_nccl.GroupEnd());
    for (auto& c : g.devs) {
        CK(cudaSetDevice(c.dev));
        CK(cudaStreamSynchronize(c.stream));
    }
    return ORT_OK;
}

extern "C" int ort_init(int ngpus) {
    if (g.ready) ort_finalize();
    int n = ort_device_count();
    if (n <= 0) {
        ort_set_error("no CUDA device visible (this library has no CPU fallback)");
        return ORT_ENODEVICE;
    }
    if (ngpus <= 0 || ngpus > n) ngpus = n;
    g.devs.resize(ngpus);
    for (int i = 0; i < ngpus; ++i) {
        int rc = ctx_open(g.devs[i], i);
        if (rc) return rc;
    }
    if (ngpus > 1) {
        int rc = nccl_load();
        if (rc) return rc;
        std::vector<ncclComm_t> comms(ngpus);
        std::vector<int> list(ngpus);
        for (int i = 0; i < ngpus; ++i) list[i] = i;
        NK(g_nccl.CommInitAll(comms.data(), ngpus, list.data()));
        for (int i = 0; i < ngpus; ++i) g.devs[i].comm = comms[i];
        rc = nccl_warm_up();
        if (rc) return rc;
    }
    g.rank_mode = false;
    g.rank = 0;
    g.nranks = 1;
    g.ready = true;
    return ngpus;
}

extern "C" int ort_synchronize(void) {
    if (!g.ready) return ORT_ENODEVICE;
    for (auto& c : g.devs) {
        CK(cudaSetDevice(c.dev));
        CK(cudaDeviceSynchronize());
    }
    return ORT_OK;
}

extern "C" int ort_nccl_unique_id(void* out128) {
    if (!out128) return ORT_EINVAL;
    int rc = nccl_load();
    if (rc) return rc;
    ncclUniqueId id;
    NK(g_nccl.GetUniqueId(&id));
    memcpy(out128, &id, sizeof id);
    return ORT_OK;
}

extern "C" int ort_init_rank(int device, int rank, int nranks, const void* nccl_id) {
    if (g.ready) ort_finalize();
    if (nranks < 1 || rank < 0 || rank >= nranks) {
        ort_set_error("ort_init_rank: bad rank %d / %d", rank, nranks);
        return ORT_EINVAL;
    }
    int n = ort_device_count();
    if (n <= 0) {
        ort_set_error("no CUDA device visible (this library has no CPU fallback)");
        return ORT_ENODEVICE;
    }
    if (device < 0 || device >= n) {
        ort_set_error("ort_init_rank: device %d not in 0..%d", device, n - 1);
        return ORT_EINVAL;
    }
    g.devs.resize(1);
    int rc = ctx_open(g.devs[0], device);
    if (rc) return rc;
    if (nranks > 1) {
        if (!nccl_id) {
            ort_set_error("ort_init_rank: nranks > 1 needs the ncclUniqueId");
            return ORT_EINVAL;
        }
        rc = nccl_load();
        if (rc) return rc;
        ncclUniqueId id;
        memcpy(&id, nccl_id, sizeof id);
        NK(g_nccl.CommInitRank(&g.devs[0].comm, nranks, id, rank));
        rc = nccl_warm_up();
        if (rc) return rc;
    }
    g.rank_mode = true;
    g.rank = rank;
    g.nranks = nranks;
    g.ready = true;
    return 1;
}

static int validate_job(const ort_job* job) {
    if (!job) {
        ort_set_error("job is NULL");
        return ORT_EINVAL;
    }
    if (job->phase != ORT_PHASE_RING && job->phase != ORT_PHASE_POINT) {
        ort_set_error("job.phase must be 1 (ring) or 2 (point), got %d", job->phase);
        return ORT_EINVAL;
    }
    if (job->precision != 64 && job->precision != 32) {
        ort_set_error("job.precision %d not supported (64 or 32)", job->precision);
        return ORT_EINVAL;
    }
    if (job->nrays < 0 || job->first_ray < 0) {
        ort_set_error("negative ray count / index");
        return ORT_EINVAL;
    }
    if (!(job->image_diameter > 0.0)) {
        ort_set_error("image_diameter must be > 0");
        return ORT_EINVAL;
    }
    if (job->source_kind < ORT_SRC_POINT || job->source_kind > ORT_SRC_IMAGE) {
        ort_set_error("job.source_kind %d unknown", job->source_kind);
        return ORT_EINVAL;
    }
    if (job->uniform_override >= 1.0) {
        ort_set_error("uniform_override must be < 1");
        return ORT_EINVAL;
    }
    return ORT_OK;
}

/* ------------------------------------------------------------------------------------------
 * kernel dispatch
 * ---------------------------------------------------------------------------------------- */
/* image source: the host keeps the prefix sums; each device gets a copy on first use */
static int ensure_image_source(DeviceCtx& c) {
    if (g.image_cdf.empty()) {
        ort_set_error("source_kind = image needs ort_set_image_source() first");
        return ORT_EINVAL;
    }
    if (c.d_image_cdf) return ORT_OK;
    CK(cudaSetDevice(c.dev));
    CK(cudaMalloc(&c.d_image_cdf, g.image_cdf.size() * sizeof(long long)));
    CK(cudaMemcpy(c.d_image_cdf, g.image_cdf.data(), g.image_cdf.size() * sizeof(long long), cudaMemcpyHostToDevice));
    return ORT_OK;
}

extern "C" int ort_set_image_source(const int32_t* budget) {
    for (auto& c : g.devs) {
        if (c.d_image_cdf) {
            cudaSetDevice(c.dev);
            cudaFree(c.d_image_cdf);
            c.d_image_cdf = nullptr;
        }
    }
    g.image_cdf.clear();
    if (!budget) return ORT_OK;
    const size_t n = (size_t)ORT_SRCIMG_N * ORT_SRCIMG_N;
    g.image_cdf.resize(n);
    long long acc = 0;
    for (size_t k = 0; k < n; ++k) {
        acc += budget[k] > 0 ? budget[k] : 0;
        g.image_cdf[k] = acc;
    }
    return ORT_OK;
}

/* ------------------------------------------------------------------------------------------
 * kernel dispatch: <loop, bottle kind, source kind, real type>
 * ---------------------------------------------------------------------------------------- */
template <typename R>
struct Kernels {
    typedef void (*trace_t)(const DevSceneT<R>, const DevJob, unsigned long long*, unsigned long long*);
    /* point loop through a scattering bottle: the kernel that queues rays between scatter events */
    static trace_t pick_scatter(int src) {
        switch (src) {
            case ORT_SRC_IMAGE: return ort_trace_scatter_kernel<ORT_SRC_IMAGE, R>;
            case ORT_SRC_SPOT: return ort_trace_scatter_kernel<ORT_SRC_SPOT, R>;
            default: return ort_trace_scatter_kernel<ORT_SRC_POINT, R>; /* point, crs, isors: point() */
        }
    }
    static trace_t pick(int phase, int bottle_mode, int src, bool flat) {
        if (flat) { /* diagnostic kernel: default sources only */
            if (phase == ORT_PHASE_RING) return ort_trace_flat_kernel<ORT_PHASE_RING, 0, R>;
            switch (bottle_mode) {
                case 0: return ort_trace_flat_kernel<ORT_PHASE_POINT, 0, R>;
                case 1: return ort_trace_flat_kernel<ORT_PHASE_POINT, 1, R>;
                default: return ort_trace_flat_kernel<ORT_PHASE_POINT, 2, R>;
            }
        }
        if (phase == ORT_PHASE_RING) {
            switch (src) {
                case ORT_SRC_CRS: return ort_trace_kernel<ORT_PHASE_RING, 0, ORT_SRC_CRS, R>;
                case ORT_SRC_ISORS: return ort_trace_kernel<ORT_PHASE_RING, 0, ORT_SRC_ISORS, R>;
                default: return ort_trace_kernel<ORT_PHASE_RING, 0, ORT_SRC_POINT, R>; /* point, spot: ring() */
            }
        }
        if (src == ORT_SRC_IMAGE) {
            switch (bottle_mode) {
                case 0: return ort_trace_kernel<ORT_PHASE_POINT, 0, ORT_SRC_IMAGE, R>;
                case 1: return ort_trace_kernel<ORT_PHASE_POINT, 1, ORT_SRC_IMAGE, R>;
                default: return ort_trace_kernel<ORT_PHASE_POINT, 2, ORT_SRC_IMAGE, R>;
            }
        }
        const bool spot = src == ORT_SRC_SPOT; /* crs, isors: point() in the point loop */
        switch (bottle_mode) {
            case 0: return spot ? ort_trace_kernel<ORT_PHASE_POINT, 0, ORT_SRC_SPOT, R> : ort_trace_kernel<ORT_PHASE_POINT, 0, ORT_SRC_POINT, R>;
            case 1: return spot ? ort_trace_kernel<ORT_PHASE_POINT, 1, ORT_SRC_SPOT, R> : ort_trace_kernel<ORT_PHASE_POINT, 1, ORT_SRC_POINT, R>;
            default: return spot ? ort_trace_kernel<ORT_PHASE_POINT, 2, ORT_SRC_SPOT, R> : ort_trace_kernel<ORT_PHASE_POINT, 2, ORT_SRC_POINT, R>;
        }
    }
};
static inline void scene_as(const DevScene& s, DevSceneT<double>& d) { d = s; }
static inline void scene_as(const DevScene& s, DevSceneT<float>& d) { ort_scene_to_float(s, d); }

static const int64_t ORT_CHUNK = (int64_t)1 << 31; /* rays per scene per launch (ids are 32-bit) */

/* The ring loop's four-stage kernel with the fp32 culling filter (ort_kernels.cuh): fp64 jobs with
 * the default ring source on scenes whose L2 flat face lies in the aim plane. */
template <typename R>
static bool ring_filter_applies(const ort_job& job, const DevScene& s, bool flat, unsigned long long* aim_cut,
                                DevFilter* K) {
    if (!(sizeof(R) == sizeof(double) && !flat && job.phase == ORT_PHASE_RING && s.ring_shortcut &&
          (job.source_kind == ORT_SRC_POINT || job.source_kind == ORT_SRC_SPOT) && !(job.flags & ORT_FLAG_NO_FILTER)))
        return false;
    ort_make_filter(s, job.iris_before != 0, *K);
    return K->usable == 2 && ort_ring_aim_cut(s, aim_cut);
}
/* Slice of the ray range per cull/survivors launch pair: long enough that the tail of a launch
 * is < 1 % of it, short enough that the list of ray indices stays a few hundred MB. */
static const int64_t ORT_RING_SLICE = (int64_t)1 << 29;
static_assert(ORT_RING_SLICE <= ((int64_t)1 << ORT_LIST_ID_BITS), "a list entry keeps its ray index in ORT_LIST_ID_BITS bits");

/* The survivors list of one lane: at most the rays that pass stage A land on it -- Binomial(slice, p),
 * p = aim_cut / 2^64, standard deviation <= sqrt(slice) / 2; capacity = expectation + 8 sqrt(slice) >= 16
 * sigma -- two buffers used alternately, plus one length per slice.  Grow-only; called for every scene
 * BEFORE the timed region starts, so that a first call does not time its own cudaMalloc. */
static int ring_reserve(Lane& L, unsigned long long aim_cut, int64_t n, size_t* capacity_out, int64_t* nslices_out) {
    const int64_t slice = n < ORT_RING_SLICE ? n : ORT_RING_SLICE;
    const double p_pass = (double)aim_cut * (1.0 / 18446744073709551616.0);
    double want_cap = (double)slice * p_pass + 8.0 * std::sqrt((double)slice) + 1024.0;
    if (want_cap > (double)slice) want_cap = (double)slice;
    size_t capacity = (size_t)want_cap;
#ifdef ORT_DEBUG
    /* assert-instrumented build only (libort_debug.so): lets a test provoke the overflow report */
    if (const char* e = getenv("ORT_TEST_RING_LIST_CAP")) {
        long v = atol(e);
        if (v > 0 && (size_t)v < capacity) capacity = (size_t)v;
    }
#endif
    const int64_t nslices = (n + ORT_RING_SLICE - 1) / ORT_RING_SLICE;
    if (L.list_cap < capacity) {
        if (L.d_list) CK(cudaFree(L.d_list)); /* (cudaFree waits for the device: nothing still reads the old list) */
        L.d_list = nullptr;
        L.list_cap = 0;
        CK(cudaMalloc(&L.d_list, 2 * capacity * sizeof(uint32_t)));
        L.list_cap = capacity;
    }
    if (L.nlist_cap < (size_t)nslices) {
        if (L.d_nlist) CK(cudaFree(L.d_nlist));
        L.d_nlist = nullptr;
        L.nlist_cap = 0;
        CK(cudaMalloc(&L.d_nlist, (size_t)nslices * sizeof(unsigned)));
        L.nlist_cap = (size_t)nslices;
    }
    *capacity_out = capacity;
    *nslices_out = nslices;
    return ORT_OK;
}

static int enqueue_ring_filter(DeviceCtx& c, Lane& L, const ort_job& job, const DevScene& s, const DevFilter& K,
                               unsigned long long aim_cut, int nscenes, int64_t first, int64_t n, unsigned long long* d_img,
                               unsigned long long* d_cnt, int64_t* launches, bool dry) {
    const bool verify = (job.flags & ORT_FLAG_VERIFY_FILTER) != 0;
    auto cull = verify ? ort_ring_cull_kernel<true> : ort_ring_cull_kernel<false>;
    auto surv = verify ? ort_ring_survivors_kernel<true> : ort_ring_survivors_kernel<false>;
    const size_t smem_cull = (size_t)ORT_WPB * sizeof(SlimQueue);
    const size_t smem_surv = (size_t)ORT_WPB * sizeof(SurvShared);
    int occ_cull = 0, occ_surv = 0;
    int orc = ctx_occupancy(c, (const void*)cull, smem_cull, &occ_cull);
    if (orc == ORT_OK) orc = ctx_occupancy(c, (const void*)surv, smem_surv, &occ_surv);
    if (orc != ORT_OK) return orc;

    size_t capacity = 0;
    int64_t nslices = 0;
    const double p_pass = (double)aim_cut * (1.0 / 18446744073709551616.0);
    int rrc = ring_reserve(L, aim_cut, n, &capacity, &nslices);
    if (rrc != ORT_OK) return rrc;
    if (dry) return ORT_OK; /* kernels loaded, occupancy known, memory reserved */
    /* the previous scene of this lane may still be reading its lists and lengths on stream2 */
    CK(cudaStreamWaitEvent(L.stream, L.ev_surv[0], 0));
    CK(cudaStreamWaitEvent(L.stream, L.ev_surv[1], 0));
    CK(cudaMemsetAsync(L.d_nlist, 0, (size_t)nslices * sizeof(unsigned), L.stream));

    DevSceneT<float> sf;
    ort_scene_to_float(s, sf);
    OrtfParamsT<OrtfV2> fp; /* every constant of the filter, twice: one 64-bit operand for two rays */
    ortf_make_params<OrtfV2>(sf, K, job.iris_before, fp);
    /* cull(k) runs on the lane's stream, survivors(k) on its second one behind it, so that the small
     * fp64 kernel fills the tail of cull(k+1) instead of standing between two cull kernels; the
     * two list buffers alternate, and cull(k+2) waits until survivors(k) has read its buffer */
    int64_t k = 0;
    for (int64_t off = 0; off < n; off += ORT_RING_SLICE, ++k) {
        int64_t m = n - off < ORT_RING_SLICE ? n - off : ORT_RING_SLICE;
        const int buf = (int)(k & 1);
        uint32_t* list = L.d_list + (size_t)buf * L.list_cap;
        DevJob dj;
        ort_make_dev_job(job, nscenes, first + off, m, dj);
        dj.aim_cut = aim_cut;
        int64_t batches = (m + 127) / 128; /* a warp of the cull kernel takes 128 rays per pass */
        int64_t want = (batches + ORT_WPB - 1) / ORT_WPB;
        int grid = c.num_sms * occ_cull;
        int gsz = (int)(want < grid ? (want > 0 ? want : 1) : grid);
        if (k >= 2) CK(cudaStreamWaitEvent(L.stream, L.ev_surv[buf], 0));
        cull<<<gsz, ORT_TPB, smem_cull, L.stream>>>(fp, dj, list, L.d_nlist + k, (unsigned)capacity, d_cnt);
        CK(cudaGetLastError());
        CK(cudaEventRecord(L.ev_cull[buf], L.stream));
        /* the list length is only known on the device: size the grid for the longest list there can
         * be (every ray that passes stage A); blocks that find nothing to do leave at once */
        double expect = (double)m * p_pass;
        int64_t sb = ((int64_t)expect / 32 + ORT_WPB) / ORT_WPB;
        int sgrid = c.num_sms * occ_surv;
        int sgsz = (int)(sb < sgrid ? (sb > 0 ? sb : 1) : sgrid);
        CK(cudaStreamWaitEvent(L.stream2, L.ev_cull[buf], 0));
        surv<<<sgsz, ORT_TPB, smem_surv, L.stream2>>>(s, dj, list, L.d_nlist + k, (unsigned)capacity, d_img, d_cnt);
        CK(cudaGetLastError());
        CK(cudaEventRecord(L.ev_surv[buf], L.stream2));
        *launches += 2;
    }
    /* join: whatever follows on the lane's stream (next scene, reduce, read-back) sees every count */
    for (int64_t j = k > 2 ? k - 2 : 0; j < k; ++j) CK(cudaStreamWaitEvent(L.stream, L.ev_surv[j & 1], 0));
    return ORT_OK;
}

/* what the host works out per scene before anything is launched (once per call) */
struct ScenePlan {
    bool filter = false;              /* ring loop through the culling kernel */
    unsigned long long aim_cut = 0;   /* ... its integer aperture cut */
    DevFilter K;                      /* ... and error-bound constants */
    unsigned long long ring_cut = 0;  /* the cut of the other ring kernels (their own expression) */
};
template <typename R>
static void plan_scenes(const ort_job& job, const std::vector<DevScene>& ds, bool flat, std::vector<ScenePlan>& plan) {
    plan.resize(ds.size());
    for (size_t sc = 0; sc < ds.size(); ++sc) {
        ScenePlan& p = plan[sc];
        p.filter = ring_filter_applies<R>(job, ds[sc], flat, &p.aim_cut, &p.K);
        if (!p.filter && job.phase == ORT_PHASE_RING && ds[sc].ring_shortcut &&
            !ort_ring_aim_cut(ds[sc], &p.ring_cut, sizeof(R) == 4))
            p.ring_cut = ~0ull; /* no draw fails L2's aperture: every word is below the cut, the all-ones word is re-tested */
    }
}

/* one pass over the scenes of a call: the launches, or (dry) only what has to exist before them */
template <typename R>
static int enqueue_scenes(DeviceCtx& c, const ort_job& job, const std::vector<DevScene>& ds, const std::vector<ScenePlan>& plan,
                          int64_t first, int64_t n, bool flat, size_t smem_trace, int nlanes, unsigned long long* d_img,
                          unsigned long long* d_cnt, int64_t* launches, bool dry) {
    const int nscenes = (int)ds.size();
    for (int sc = 0; sc < nscenes; ++sc) {
        Lane& L = c.lanes[sc % nlanes];
        const ScenePlan& P = plan[sc];
        if (P.filter) {
            int rc = enqueue_ring_filter(c, L, job, ds[sc], P.K, P.aim_cut, nscenes, first, n, d_img + (size_t)sc * ORT_IMG_BINS,
                                         d_cnt + (size_t)sc * ORT_NSTATUS, launches, dry);
            if (rc != ORT_OK) return rc;
            continue;
        }
        /* the kernel is chosen per scene: a clear bottle batched with a scattering one runs exactly
         * the arithmetic it runs alone (hoisted 1/R wall normals), so a scene's result does not
         * depend on what it is batched with */
        const int bottle_mode = (job.phase == ORT_PHASE_POINT && job.use_bottle)
                                    ? ((ds[sc].scatter_b || ds[sc].scatter_c) ? 2 : 1) : 0;
        const bool scatter_kernel = bottle_mode == 2 && !flat;
        typename Kernels<R>::trace_t k = scatter_kernel ? Kernels<R>::pick_scatter(job.source_kind)
                                                        : Kernels<R>::pick(job.phase, bottle_mode, job.source_kind, flat);
        const size_t smem = scatter_kernel ? (size_t)ORT_WPB * sizeof(ScatterShared<R>) : smem_trace;
        int occ = 0;
        int orc = ctx_occupancy(c, (const void*)k, smem, &occ);
        if (orc != ORT_OK) return orc;
        if (dry) continue;
        DevSceneT<R> dsr;
        scene_as(ds[sc], dsr);
        const int grid = c.num_sms * occ;
        for (int64_t off = 0; off < n; off += ORT_CHUNK) {
            int64_t m = n - off < ORT_CHUNK ? n - off : ORT_CHUNK;
            DevJob dj;
            ort_make_dev_job(job, nscenes, first + off, m, dj);
            dj.image_cdf = c.d_image_cdf;
            dj.aim_cut = P.ring_cut; /* the ring loop's stage A decides on the high word of the aim draw */
            int64_t batches = (m + 31) / 32;
            int64_t want = (batches + ORT_WPB - 1) / ORT_WPB;
            int gsz = (int)(want < grid ? (want > 0 ? want : 1) : grid);
            k<<<gsz, ORT_TPB, smem, L.stream>>>(dsr, dj, d_img + (size_t)sc * ORT_IMG_BINS,
                                                d_cnt + (size_t)sc * ORT_NSTATUS);
            CK(cudaGetLastError());
            ++*launches;
        }
    }
    return ORT_OK;
}

/* enqueue the trace of rays [first, first+n) of every scene on device ctx; returns launches */
/* what: 1 = only the preparation (memory, kernels loaded), 2 = only the launches, 3 = both */
template <typename R>
static int enqueue_trace_t(DeviceCtx& c, const ort_job& job, const std::vector<DevScene>& ds, int64_t first,
                           int64_t n, int64_t* launches, int what) {
    const int nscenes = (int)ds.size();
    CK(cudaSetDevice(c.dev));
    size_t elems = (size_t)nscenes * (ORT_IMG_BINS + ORT_NSTATUS);
    if (c.d_elems < elems) {
        if (c.d_buf) CK(cudaFree(c.d_buf));
        c.d_buf = nullptr;
        CK(cudaMalloc(&c.d_buf, elems * sizeof(unsigned long long)));
        c.d_elems = elems;
    }
    bool flat = (job.flags & ORT_FLAG_NO_COMPACTION) != 0 && job.source_kind == ORT_SRC_POINT;
    const size_t smem_trace = flat ? 0 : (size_t)ORT_WPB * sizeof(WarpShared<R>);

    const int nlanes = (nscenes > 1 && n <= ((int64_t)1 << 26) && !(job.flags & ORT_FLAG_ONE_LANE))
                           ? (nscenes < ORT_LANES ? nscenes : ORT_LANES) : 1;
    /* everything that is not tracing happens before the timed region, in a dry pass over the scenes: device
     * memory for the ring loop's lists, and the first use of a kernel (the driver loads it then, and the
     * occupancy query is answered from a cache afterwards) */
    unsigned long long* d_img = c.d_buf;
    unsigned long long* d_cnt = c.d_buf + (size_t)nscenes * ORT_IMG_BINS;
    std::vector<ScenePlan> plan;
    plan_scenes<R>(job, ds, flat, plan);
    if (what & 1) {
        int64_t none = 0;
        int rc = enqueue_scenes<R>(c, job, ds, plan, first, n, flat, smem_trace, nlanes, d_img, d_cnt, &none, true);
        if (rc != ORT_OK) return rc;
    }
    if (!(what & 2)) return ORT_OK;

    CK(cudaEventRecord(c.ev_start, c.stream));
    CK(cudaMemsetAsync(c.d_buf, 0, elems * sizeof(unsigned long long), c.stream));
    /* One launch per scene and per <= 2^31-ray chunk; scene and job travel as kernel parameters, so there
     * is nothing to upload between launches and every scene scalar is an immediate constant-bank operand.
     * (Indexing the scenes INSIDE one kernel was measured and rejected: a register-indexed LDC per scalar
     * and a three-register DFMA behind it, ~15 % per ray.)  Large jobs run back to back on the main
     * stream.  When a batched call has few rays per scene -- a quick-look sweep -- the scenes go round
     * robin over ORT_LANES streams instead: the tail of one scene's persistent kernel is filled by the
     * blocks of the next, and the launch latencies overlap. */
    if (nlanes > 1) {
        CK(cudaEventRecord(c.ev_zeroed, c.stream));
        for (int l = 1; l < nlanes; ++l) CK(cudaStreamWaitEvent(c.lanes[l].stream, c.ev_zeroed, 0));
    }
    {
        int rc = enqueue_scenes<R>(c, job, ds, plan, first, n, flat, smem_trace, nlanes, d_img, d_cnt, launches, false);
        if (rc != ORT_OK) return rc;
    }
    for (int l = 1; l < nlanes; ++l) { /* the main stream continues when every lane is done */
        CK(cudaEventRecord(c.lanes[l].ev_done, c.lanes[l].stream));
        CK(cudaStreamWaitEvent(c.stream, c.lanes[l].ev_done, 0));
    }
    CK(cudaEventRecord(c.ev_traced, c.stream));
    return ORT_OK;
}
static int enqueue_trace(DeviceCtx& c, const ort_job& job, const std::vector<DevScene>& ds, int64_t first,
                         int64_t n, int64_t* launches, int what = 3) {
    return job.precision == 32 ? enqueue_trace_t<float>(c, job, ds, first, n, launches, what)
                               : enqueue_trace_t<double>(c, job, ds, first, n, launches, what);
}

/* pinned staging buffer -> the caller's (pageable, often untouched) buffer.  A batched call returns
 * 1.29 MB per scene; above a few MB the copy and the page faults of a fresh destination are spread over
 * a few threads (a 75-scene sweep: 96 MB, ~18 ms single-threaded -- more than its kernels take) */
static void host_copy(void* dst, const void* src, size_t bytes) {
    const size_t chunk = (size_t)4 << 20;
    unsigned hw = std::thread::hardware_concurrency();
    size_t nt = bytes / chunk;
    if (nt > 8) nt = 8;
    if (hw && nt > hw) nt = hw;
    if (nt < 2) {
        memcpy(dst, src, bytes);
        return;
    }
    std::vector<std::thread> th;
    const size_t per = ((bytes / nt) + 4095) & ~(size_t)4095;
    for (size_t t = 0; t < nt; ++t) {
        const size_t lo = t * per, hi = (t + 1 == nt || (t + 1) * per > bytes) ? bytes : (t + 1) * per;
        if (lo >= hi) break;
        th.emplace_back([=] { memcpy((char*)dst + lo, (const char*)src + lo, hi - lo); });
    }
    for (auto& t : th) t.join();
}

extern "C" int ort_trace(const ort_job* job, const ort_scene* scenes, int nscenes, uint64_t* image,
                         int64_t* lost, int64_t* status_hist, ort_timing* timing) {
    auto w0 = std::chrono::steady_clock::now();
    if (!g.ready) {
        ort_set_error("ort_trace: library not initialised (ort_init / ort_init_rank)");
        return ORT_ENODEVICE;
    }
    int rc = validate_job(job);
    if (rc) return rc;
    if (job->uniform_override >= 0.0) {
        ort_set_error("ort_trace: uniform_override is a known-answer-test device of ort_trace_rays only");
        return ORT_EINVAL;
    }
    if (!scenes || nscenes < 1 || nscenes > ORT_MAX_SCENES) {
        ort_set_error("ort_trace: nscenes must be 1..%d", ORT_MAX_SCENES);
        return ORT_EINVAL;
    }
    std::vector<DevScene> ds(nscenes);
    for (int i = 0; i < nscenes; ++i) ort_flatten_scene(scenes[i], *job, ds[i]);

    const int G = (int)g.devs.size();
    if (job->source_kind == ORT_SRC_IMAGE && job->phase == ORT_PHASE_POINT) {
        for (int d = 0; d < G; ++d) {
            rc = ensure_image_source(g.devs[d]);
            if (rc) return rc;
        }
    }
    const size_t elems = (size_t)nscenes * (ORT_IMG_BINS + ORT_NSTATUS);
    int64_t launches = 0;
    /* contiguous ray-index ranges per device; uniforms depend only on the ray
     * index, so the summed image is identical for any G */
    /* every device is prepared before the first one is started: a device's timed region then holds its own
     * launches only, not the host work for the devices after it */
    for (int pass = G > 1 ? 1 : 3; pass <= 3; pass += (pass == 1 ? 1 : 2)) {
        for (int d = 0; d < G; ++d) {
            int64_t lo = job->nrays * d / G, hi = job->nrays * (d + 1) / G;
            rc = enqueue_trace(g.devs[d], *job, ds, job->first_ray + lo, hi - lo, &launches, pass);
            if (rc) return rc;
        }
    }
    /* one reduce of [images | counters] to device 0 / rank 0 */
    bool reduced = false;
    if (G > 1) {
        NK(g_nccl.GroupStart());
        for (int d = 0; d < G; ++d) {
            DeviceCtx& c = g.devs[d];
            NK(g_nccl.Reduce(c.d_buf, c.d_buf, elems, ncclUint64, ncclSum, 0, c.comm, c.stream));
        }
        NK(g_nccl.GroupEnd());
        reduced = true;
    } else if (g.rank_mode && g.nranks > 1 && !(job->flags & ORT_FLAG_NO_REDUCE)) {
        DeviceCtx& c = g.devs[0];
        CK(cudaSetDevice(c.dev));
        NK(g_nccl.Reduce(c.d_buf, c.d_buf, elems, ncclUint64, ncclSum, 0, c.comm, c.stream));
        reduced = true;
    }
    for (int d = 0; d < G; ++d) {
        CK(cudaSetDevice(g.devs[d].dev));
        CK(cudaEventRecord(g.devs[d].ev_reduced, g.devs[d].stream));
    }
    /* results back to the host */
    DeviceCtx& c0 = g.devs[0];
    CK(cudaSetDevice(c0.dev));
    if (g.h_elems < elems) {
        if (g.h_pinned) CK(cudaFreeHost(g.h_pinned));
        g.h_pinned = nullptr;
        CK(cudaHostAlloc(&g.h_pinned, elems * sizeof(unsigned long long), cudaHostAllocDefault));
        g.h_elems = elems;
    }
    size_t img_elems = (size_t)nscenes * ORT_IMG_BINS;
    size_t d2h = 0;
    if (image) {
        CK(cudaMemcpyAsync(g.h_pinned, c0.d_buf, img_elems * 8, cudaMemcpyDeviceToHost, c0.stream));
        d2h += img_elems * 8;
    }
    CK(cudaMemcpyAsync(g.h_pinned + img_elems, c0.d_buf + img_elems, (size_t)nscenes * ORT_NSTATUS * 8,
                       cudaMemcpyDeviceToHost, c0.stream));
    d2h += (size_t)nscenes * ORT_NSTATUS * 8;
    CK(cudaEventRecord(c0.ev_copied, c0.stream));
    for (int d = 0; d < G; ++d) {
        CK(cudaSetDevice(g.devs[d].dev));
        CK(cudaStreamSynchronize(g.devs[d].stream));
    }
    if (image) host_copy(image, g.h_pinned, img_elems * 8);
    bool trapped = false, list_overflow = false;
    for (int s = 0; s < nscenes; ++s) {
        const unsigned long long* h = g.h_pinned + img_elems + (size_t)s * ORT_NSTATUS;
        int64_t l = 0;
        for (int k = 0; k < ORT_NSTATUS; ++k) {
            if (ORT_STATUS_IS_LOST(k)) l += (int64_t)h[k];
            if (status_hist) status_hist[(size_t)s * ORT_NSTATUS + k] = (int64_t)h[k];
        }
        if (lost) lost[s] = l;
        if (h[ORT_ST_L3_S3_MISS] || h[ORT_ST_TAUINT_MISS] || h[ORT_ST_SOURCE_MISS]) trapped = true;
        if (h[ORT_FILTER_SLOT_OVERFLOW]) list_overflow = true;
    }
    if (timing) {
        double tmax = 0.0, rmax = 0.0;
        for (int d = 0; d < G; ++d) {
            float a = 0.f, b = 0.f;
            CK(cudaSetDevice(g.devs[d].dev));
            CK(cudaEventElapsedTime(&a, g.devs[d].ev_start, g.devs[d].ev_traced));
            CK(cudaEventElapsedTime(&b, g.devs[d].ev_traced, g.devs[d].ev_reduced));
            if (a > tmax) tmax = a;
            if (b > rmax) rmax = b;
        }
        timing->trace_seconds = tmax * 1e-3;
        timing->reduce_seconds = reduced ? rmax * 1e-3 : 0.0;
        {
            float cms = 0.f;
            CK(cudaSetDevice(c0.dev));
            CK(cudaEventElapsedTime(&cms, c0.ev_reduced, c0.ev_copied));
            timing->d2h_seconds = cms * 1e-3;
        }
        timing->kernel_launches = launches;
        /* scene + job of every launch travel host -> device as kernel parameters */
        timing->h2d_bytes = (int64_t)(launches * (int64_t)(sizeof(DevScene) + sizeof(DevJob)));
        timing->d2h_bytes = (int64_t)d2h;
        timing->wall_seconds =
            std::chrono::duration<double>(std::chrono::steady_clock::now() - w0).count();
    }
    if (list_overflow) {
        ort_set_error("ort_trace: the ring filter's survivor list overflowed (results incomplete); "
                      "retry with ORT_FLAG_NO_FILTER");
        return ORT_ECUDA;
    }
    if (trapped) {
        ort_set_error("trace hit a reference `error stop` invariant (status 18, 24 or 26); results returned");
        return ORT_ETRACE;
    }
    return ORT_OK;
}

extern "C" int ort_trace_rays(const ort_job* job, const ort_scene* scene, int64_t n, const double* pos_in,
                              const double* dir_in, double* pos_out, double* dir_out, int32_t* status,
                              int32_t* bin_xy) {
    if (!g.ready) {
        ort_set_error("ort_trace_rays: library not initialised (ort_init / ort_init_rank)");
        return ORT_ENODEVICE;
    }
    int rc = validate_job(job);
    if (rc) return rc;
    if (!scene || n < 0 || n >= ((int64_t)1 << 31) || (pos_in == nullptr) != (dir_in == nullptr)) {
        ort_set_error("ort_trace_rays: bad arguments");
        return ORT_EINVAL;
    }
    if (n == 0) return ORT_OK;
    DeviceCtx& c = g.devs[0];
    CK(cudaSetDevice(c.dev));
    DevScene ds;
    ort_flatten_scene(*scene, *job, ds);
    DevJob dj;
    ort_make_dev_job(*job, 1, job->first_ray, n, dj);
    if (job->source_kind == ORT_SRC_IMAGE && job->phase == ORT_PHASE_POINT && !pos_in) {
        rc = ensure_image_source(c);
        if (rc) return rc;
    }
    dj.image_cdf = c.d_image_cdf;
    double *d_in = nullptr, *d_out = nullptr;
    int32_t* d_int = nullptr;
    size_t vb = (size_t)3 * n * sizeof(double);
    int ret = ORT_OK;
    do {
#define CKB(call)                                                                                  \
    {                                                                                              \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            ort_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_));    \
            ret = ORT_ECUDA;                                                                       \
            break;                                                                                 \
        }                                                                                          \
    }
        void* sp = nullptr;
        if (pos_in) {
            if ((ret = ctx_scratch(c, 0, 2 * vb, &sp)) != ORT_OK) break;
            d_in = (double*)sp;
            CKB(cudaMemcpyAsync(d_in, pos_in, vb, cudaMemcpyHostToDevice, c.stream));
            CKB(cudaMemcpyAsync(d_in + 3 * n, dir_in, vb, cudaMemcpyHostToDevice, c.stream));
        }
        if ((ret = ctx_scratch(c, 1, 2 * vb, &sp)) != ORT_OK) break;
        d_out = (double*)sp;
        if ((ret = ctx_scratch(c, 2, (size_t)3 * n * sizeof(int32_t), &sp)) != ORT_OK) break;
        d_int = (int32_t*)sp;
        int grid = (int)((n + ORT_TPB - 1) / ORT_TPB);
        if (job->precision == 32) {
            DevSceneT<float> dsf;
            ort_scene_to_float(ds, dsf);
            ort_rays_kernel<float><<<grid, ORT_TPB, 0, c.stream>>>(dsf, dj, d_in, d_in ? d_in + 3 * n : nullptr, d_out,
                                                                   d_out + 3 * n, d_int, d_int + n, (long long)n);
        } else {
            ort_rays_kernel<double><<<grid, ORT_TPB, 0, c.stream>>>(ds, dj, d_in, d_in ? d_in + 3 * n : nullptr, d_out,
                                                                    d_out + 3 * n, d_int, d_int + n, (long long)n);
        }
        CKB(cudaGetLastError());
        if (pos_out) CKB(cudaMemcpyAsync(pos_out, d_out, vb, cudaMemcpyDeviceToHost, c.stream));
        if (dir_out) CKB(cudaMemcpyAsync(dir_out, d_out + 3 * n, vb, cudaMemcpyDeviceToHost, c.stream));
        if (status) CKB(cudaMemcpyAsync(status, d_int, n * sizeof(int32_t), cudaMemcpyDeviceToHost, c.stream));
        if (bin_xy) CKB(cudaMemcpyAsync(bin_xy, d_int + n, 2 * n * sizeof(int32_t), cudaMemcpyDeviceToHost, c.stream));
        CKB(cudaStreamSynchronize(c.stream));
#undef CKB
    } while (0);
    return ret;
}

extern "C" int ort_uniforms(uint64_t seed, int32_t phase, int64_t ray, int32_t first_slot, int32_t n,
                            double* out) {
    if (!g.ready) {
        ort_set_error("ort_uniforms: library not initialised");
        return ORT_ENODEVICE;
    }
    if (n <= 0 || !out || first_slot < 0) return ORT_EINVAL;
    DeviceCtx& c = g.devs[0];
    CK(cudaSetDevice(c.dev));
    void* sp = nullptr;
    int src = ctx_scratch(c, 1, n * sizeof(double), &sp);
    if (src != ORT_OK) return src;
    double* d = (double*)sp;
    ort_uniforms_kernel<<<(n + 127) / 128, 128, 0, c.stream>>>(seed, phase, ray, first_slot, n, d);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, d, n * sizeof(double), cudaMemcpyDeviceToHost, c.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c.stream);
    if (e != cudaSuccess) {
        ort_set_error("ort_uniforms: %s", cudaGetErrorString(e));
        return ORT_ECUDA;
    }
    return ORT_OK;
}

extern "C" int ort_measure_fp64_peak(double* tflops, double* sm_clock_mhz) {
    if (!g.ready) {
        ort_set_error("ort_measure_fp64_peak: library not initialised");
        return ORT_ENODEVICE;
    }
    DeviceCtx& c = g.devs[0];
    CK(cudaSetDevice(c.dev));
    double* d_out = nullptr;
    unsigned long long* d_cyc = nullptr;
    CK(cudaMalloc(&d_out, sizeof(double)));
    CK(cudaMalloc(&d_cyc, sizeof(unsigned long long)));
    const int grid = c.num_sms * 8, tpb = 256;
    double best = 0.0, mhz = 0.0;
    for (int rep = 0; rep < 6; ++rep) {
        CK(cudaMemsetAsync(d_cyc, 0, sizeof(unsigned long long), c.stream));
        CK(cudaEventRecord(c.ev_start, c.stream));
        ort_dfma_peak_kernel<<<grid, tpb, 0, c.stream>>>(d_out, 1.0 + rep, d_cyc);
        CK(cudaGetLastError());
        CK(cudaEventRecord(c.ev_traced, c.stream));
        CK(cudaStreamSynchronize(c.stream));
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, c.ev_start, c.ev_traced));
        double flops = (double)grid * tpb * ORT_PEAK_ITERS * 8.0 * 2.0;
        double tf = flops / (ms * 1e-3) * 1e-12;
        if (rep >= 1 && tf > best) {
            best = tf;
            unsigned long long cyc = 0;
            CK(cudaMemcpy(&cyc, d_cyc, sizeof cyc, cudaMemcpyDeviceToHost));
            mhz = (double)cyc / (ms * 1e-3) * 1e-6;
        }
    }
    cudaFree(d_out);
    cudaFree(d_cyc);
    if (tflops) *tflops = best;
    if (sm_clock_mhz) *sm_clock_mhz = mhz;
    return ORT_OK;
}

extern "C" int ort_math_selftest(int64_t n, uint64_t max_ulp[4]) {
    if (!g.ready) {
        ort_set_error("ort_math_selftest: library not initialised");
        return ORT_ENODEVICE;
    }
    if (n <= 0 || !max_ulp) return ORT_EINVAL;
    DeviceCtx& c = g.devs[0];
    CK(cudaSetDevice(c.dev));
    unsigned long long* d = nullptr;
    CK(cudaMalloc(&d, 4 * sizeof(unsigned long long)));
    CK(cudaMemsetAsync(d, 0, 4 * sizeof(unsigned long long), c.stream));
    ort_math_selftest_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c.stream>>>((long long)n, d);
    cudaError_t e = cudaGetLastError();
    unsigned long long h[4] = {0, 0, 0, 0};
    if (e == cudaSuccess) e = cudaMemcpyAsync(h, d, sizeof h, cudaMemcpyDeviceToHost, c.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c.stream);
    cudaFree(d);
    if (e != cudaSuccess) {
        ort_set_error("ort_math_selftest: %s", cudaGetErrorString(e));
        return ORT_ECUDA;
    }
    for (int i = 0; i < 4; ++i) max_ulp[i] = h[i];
    return ORT_OK;
}

extern "C" int ort_mufu_selftest(double worst[5], double assumed[5]) {
    if (!g.ready) {
        ort_set_error("ort_mufu_selftest: library not initialised");
        return ORT_ENODEVICE;
    }
    if (!worst) return ORT_EINVAL;
    DeviceCtx& c = g.devs[0];
    CK(cudaSetDevice(c.dev));
    unsigned long long* d = nullptr;
    CK(cudaMalloc(&d, 5 * sizeof(unsigned long long)));
    CK(cudaMemsetAsync(d, 0, 5 * sizeof(unsigned long long), c.stream));
    ort_mufu_selftest_kernel<<<c.num_sms * 8, 256, 0, c.stream>>>(d);
    cudaError_t e = cudaGetLastError();
    unsigned long long h[5] = {0, 0, 0, 0, 0};
    if (e == cudaSuccess) e = cudaMemcpyAsync(h, d, sizeof h, cudaMemcpyDeviceToHost, c.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c.stream);
    cudaFree(d);
    if (e != cudaSuccess) {
        ort_set_error("ort_mufu_selftest: %s", cudaGetErrorString(e));
        return ORT_ECUDA;
    }
    for (int i = 0; i < 5; ++i) memcpy(&worst[i], &h[i], sizeof(double));
    if (assumed) {
        assumed[0] = ORTF_E_RCP; assumed[1] = ORTF_E_RSQ; assumed[2] = ORTF_E_SQRT;
        assumed[3] = assumed[4] = ORTF_E_SIN;
    }
    return ORT_OK;
}

/* ------------------------------------------------------------------------------------------
 * tracker files (reference src/stackMod.f90:38-52, src/main.f90:103-107,144-160,
 * src/optics_system.f90:28-50)
 * ---------------------------------------------------------------------------------------- */
/* makeImage3D on the first device; see include/ort.h */
extern "C" int ort_trace_volume(const ort_job* job, const ort_scene* scene, uint32_t* volume, int64_t* lost,
                                int64_t* status_hist) {
    if (!g.ready) {
        ort_set_error("ort_trace_volume: library not initialised (ort_init / ort_init_rank)");
        return ORT_ENODEVICE;
    }
    int rc = validate_job(job);
    if (rc) return rc;
    if (!scene || !volume || job->precision != 64 || job->uniform_override >= 0.0 ||
        (job->source_kind == ORT_SRC_IMAGE && job->phase == ORT_PHASE_POINT)) {
        ort_set_error("ort_trace_volume: needs a scene, a volume buffer, precision 64 and a generated source "
                      "other than `image`");
        return ORT_EINVAL;
    }
    DeviceCtx& c = g.devs[0];
    CK(cudaSetDevice(c.dev));
    DevScene ds;
    ort_flatten_scene(*scene, *job, ds);
    const size_t nvox = (size_t)ORT_VOL_DEPTH * ORT_IMG_BINS;
    unsigned* d_vol = nullptr;
    unsigned long long* d_cnt = nullptr;
    unsigned long long h_cnt[ORT_NSTATUS];
    void* sp = nullptr;
    rc = ctx_scratch(c, 0, nvox * sizeof(unsigned), &sp); /* 128.6 MB, kept for the next call */
    if (rc != ORT_OK) return rc;
    d_vol = (unsigned*)sp;
    rc = ctx_scratch(c, 2, ORT_NSTATUS * sizeof(unsigned long long), &sp);
    if (rc != ORT_OK) return rc;
    d_cnt = (unsigned long long*)sp;
    cudaError_t e = cudaMemsetAsync(d_vol, 0, nvox * sizeof(unsigned), c.stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(d_cnt, 0, ORT_NSTATUS * sizeof(unsigned long long), c.stream);
    for (int64_t off = 0; e == cudaSuccess && off < job->nrays; off += ORT_CHUNK) {
        int64_t m = job->nrays - off < ORT_CHUNK ? job->nrays - off : ORT_CHUNK;
        ort_job jj = *job;
        jj.stop_after = ORT_STOP_L3;
        DevJob dj;
        ort_make_dev_job(jj, 1, job->first_ray + off, m, dj);
        int64_t want = (m + ORT_TPB - 1) / ORT_TPB;
        int64_t cap = (int64_t)c.num_sms * 8;
        ort_volume_kernel<<<(unsigned)(want < cap ? want : cap), ORT_TPB, 0, c.stream>>>(ds, dj, job->image_diameter / 200.0, d_vol, d_cnt);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(volume, d_vol, nvox * sizeof(unsigned), cudaMemcpyDeviceToHost, c.stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(h_cnt, d_cnt, sizeof h_cnt, cudaMemcpyDeviceToHost, c.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c.stream);
    if (e != cudaSuccess) {
        ort_set_error("ort_trace_volume: %s", cudaGetErrorString(e));
        return ORT_ECUDA;
    }
    int64_t l = 0;
    bool trapped = false;
    for (int k = 0; k < ORT_NSTATUS; ++k) {
        if (ORT_STATUS_IS_LOST(k)) l += (int64_t)h_cnt[k];
        if (status_hist) status_hist[k] = (int64_t)h_cnt[k];
    }
    if (h_cnt[ORT_ST_L3_S3_MISS] || h_cnt[ORT_ST_TAUINT_MISS] || h_cnt[ORT_ST_SOURCE_MISS]) trapped = true;
    if (lost) *lost = l;
    if (trapped) {
        ort_set_error("trace hit a reference `error stop` invariant (status 18, 24 or 26); results returned");
        return ORT_ETRACE;
    }
    return ORT_OK;
}

extern "C" int ort_write_tracks(const ort_job* job, const ort_scene* scene, const char* path) {
    if (!job || !scene || !path) return ORT_EINVAL;
    const int64_t n = job->nrays;
    if (n > 10000) { /* src/setupMod.f90:75 */
        ort_set_error("Too many photons for tracker use!");
        return ORT_EINVAL;
    }
    if (n <= 0) return ORT_OK;
    /* positions after the source, the bottle, L2, L3 and at the image plane: the same kernel
     * stopped at five places (ort_job.stop_after) */
    const int stops[5] = {ORT_STOP_SOURCE, ORT_STOP_BOTTLE, ORT_STOP_L2, ORT_STOP_L3, ORT_STOP_NONE};
    std::vector<double> pos[5], dir((size_t)3 * n);
    std::vector<int32_t> status[5], bins((size_t)2 * n);
    for (int k = 0; k < 5; ++k) {
        pos[k].resize((size_t)3 * n);
        status[k].resize((size_t)n);
        ort_job j = *job;
        j.stop_after = stops[k];
        int rc = ort_trace_rays(&j, scene, n, nullptr, nullptr, pos[k].data(), dir.data(), status[k].data(),
                                bins.data());
        if (rc && rc != ORT_ETRACE) return rc;
    }
    FILE* fh = fopen(path, "w");
    if (!fh) {
        ort_set_error("cannot write %s", path);
        return ORT_EIO;
    }
    auto line = [&](int k, int64_t i) {
        fprintf(fh, "%10.7f %10.7f %10.7f \n", pos[k][i], pos[k][n + i], pos[k][2 * n + i]);
    };
    auto blanks = [&]() { fputs("  \n  \n  \n", fh); };
    const bool bottle = job->phase == ORT_PHASE_POINT && job->use_bottle;
    for (int64_t i = 0; i < n; ++i) {
        const int st = status[4][i];
        const bool lost_bottle = (st >= 1 && st <= 8) || st == ORT_ST_TAUINT_MISS;
        const bool lost_source = st == ORT_ST_SOURCE_MISS;
        const bool lost_lens = st >= ORT_ST_L2_APERTURE && st <= ORT_ST_L3_IRIS_AFTER;
        if (lost_source) { /* the reference aborts here */
            blanks();
        } else if (lost_bottle) { /* src/main.f90:150-155: write what was pushed, then write_empty */
            line(4, i); /* where the ray stopped inside the bottle */
            line(0, i);
            blanks();
            blanks();
        } else if (lost_lens) { /* src/optics_system.f90:31-34,41-44: write_empty */
            blanks();
        } else { /* src/main.f90:107,160: pop everything */
            line(4, i);
            line(3, i);
            line(2, i);
            if (bottle) line(1, i);
            line(0, i);
            blanks();
        }
    }
    fclose(fh);
    return ORT_OK;
}
