/*
 * ort_bpm.cu -- beam-propagation pre-processor on the GPU.
 *
 * Replaces the reference's bpm.py, the numpy script that writes `bessel-normal.dat` -- the
 * 512x512 fp64 intensity map the `image` source samples (src/sourceMod.f90:363-408).  What the
 * script computes once its commented-out blocks are set aside:
 *   ring field exp(-((r - r0)/w)^2)                                   bpm.py:100-121
 *   nz/10 split-step free-space propagations  e <- ifft2(fft2(e) * exp(i arg)),
 *   arg = -dz (k1^2 + k2^2) / 2k on the folded frequency grid          bpm.py:57-80, :106-117, :126-127
 *   thin-lens phase exp(-i k r^2 / 2R)                                 bpm.py:136
 *   out = |e^T|^2 as raw fp64                                          bpm.py:203-205
 * The transforms are cuFFT (Z2Z, in place; the library is dlopen()ed like NCCL so that libort.so
 * keeps no load-time dependency on it); everything around them is three small kernels: field +
 * propagator set-up, the propagator multiply with numpy's 1/N^2 folded in, and the final
 * lens phase + |.|^2 + transpose.  The field never leaves the device between steps.
 */
#include <cuda_runtime.h>
#include <cufft.h> /* types and prototypes only */
#include <dlfcn.h>

#include <cmath>
#include <cstdio>
#include <vector>

#include "ort_internal.h"

namespace {

struct CufftApi {
    void* handle = nullptr;
    cufftResult (*Plan2d)(cufftHandle*, int, int, cufftType) = nullptr;
    cufftResult (*SetStream)(cufftHandle, cudaStream_t) = nullptr;
    cufftResult (*ExecZ2Z)(cufftHandle, cufftDoubleComplex*, cufftDoubleComplex*, int) = nullptr;
    cufftResult (*Destroy)(cufftHandle) = nullptr;
};
CufftApi g_fft;

bool load_cufft() {
    if (g_fft.handle) return true;
    const char* names[] = {"libcufft.so.11", "libcufft.so.12", "libcufft.so", "/usr/local/cuda/lib64/libcufft.so.11"};
    for (const char* n : names) {
        g_fft.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (g_fft.handle) break;
    }
    if (!g_fft.handle) return false;
    g_fft.Plan2d = (decltype(g_fft.Plan2d))dlsym(g_fft.handle, "cufftPlan2d");
    g_fft.SetStream = (decltype(g_fft.SetStream))dlsym(g_fft.handle, "cufftSetStream");
    g_fft.ExecZ2Z = (decltype(g_fft.ExecZ2Z))dlsym(g_fft.handle, "cufftExecZ2Z");
    g_fft.Destroy = (decltype(g_fft.Destroy))dlsym(g_fft.handle, "cufftDestroy");
    return g_fft.Plan2d && g_fft.SetStream && g_fft.ExecZ2Z && g_fft.Destroy;
}

struct BpmGrid {
    double dx, half, dk, dz, k, R, r0, inv_w2;
    int nxy, nmid;
};

/* e[i][j] at x = j dx - xmax/2, y = i dx - xmax/2 (numpy meshgrid order); freq[i][j] = exp(i arg) / N^2 */
__global__ void bpm_setup_kernel(BpmGrid g, double2* __restrict__ e, double2* __restrict__ freq) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
    if (j >= g.nxy) return;
    const double x = j * g.dx - g.half, y = i * g.dx - g.half;
    const double r = sqrt(x * x + y * y);
    const double d = r - g.r0;
    e[(size_t)i * g.nxy + j] = make_double2(exp(-(d * d) * g.inv_w2), 0.0);
    const int fi = i > g.nmid ? g.nxy - i : i, fj = j > g.nmid ? g.nxy - j : j;
    const double k1 = fi * g.dk, k2 = fj * g.dk;
    const double arg = -g.dz * (k1 * k1 + k2 * k2) / (2.0 * g.k);
    double s, c;
    sincos(arg, &s, &c);
    const double norm = 1.0 / ((double)g.nxy * (double)g.nxy);
    freq[(size_t)i * g.nxy + j] = make_double2(c * norm, s * norm);
}
__global__ void bpm_multiply_kernel(size_t n, double2* __restrict__ e, const double2* __restrict__ freq) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const double2 a = e[t], f = freq[t];
    e[t] = make_double2(a.x * f.x - a.y * f.y, a.x * f.y + a.y * f.x);
}
/* out[j][i] = |e[i][j] * exp(-i k r^2 / 2R)|^2 */
__global__ void bpm_finish_kernel(BpmGrid g, const double2* __restrict__ e, double* __restrict__ out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
    if (j >= g.nxy) return;
    const double x = j * g.dx - g.half, y = i * g.dx - g.half;
    double s, c;
    sincos(-g.k * (x * x + y * y) / (2.0 * g.R), &s, &c);
    const double2 a = e[(size_t)i * g.nxy + j];
    const double re = a.x * c - a.y * s, im = a.x * s + a.y * c;
    out[(size_t)j * g.nxy + i] = re * re + im * im;
}

}  // namespace

#define BK(call)                                                                             \
    do {                                                                                     \
        cudaError_t e_ = (call);                                                             \
        if (e_ != cudaSuccess) {                                                             \
            ort_set_error("ort_bpm: %s failed: %s", #call, cudaGetErrorString(e_));          \
            rc = ORT_ECUDA;                                                                  \
            goto done;                                                                       \
        }                                                                                    \
    } while (0)
#define FK(call)                                                                             \
    do {                                                                                     \
        cufftResult r_ = (call);                                                             \
        if (r_ != CUFFT_SUCCESS) {                                                           \
            ort_set_error("ort_bpm: %s failed: cufft error %d", #call, (int)r_);             \
            rc = ORT_ECUDA;                                                                  \
            goto done;                                                                       \
        }                                                                                    \
    } while (0)

extern "C" int ort_bpm_defaults(ort_bpm* p) {
    if (!p) return ORT_EINVAL;
    p->w0 = 582.0 * 4.0;      /* bpm.py:84 */
    p->wavelength = 0.785;    /* :85 */
    p->axicon_deg = 5.0;      /* :87 */
    p->n_axicon = 1.45;       /* :88 */
    p->xymax = 5000.0;        /* :93 */
    p->ring_radius = 1612.0;  /* :120 */
    p->ring_width = 300.0;
    p->nxy = 512;             /* :94 */
    p->nz = 1000;             /* :95 */
    p->steps = -1;            /* nz / 10, :126 */
    p->reserved = 0;
    return ORT_OK;
}
extern "C" int ort_bpm_struct_size(void) { return (int)sizeof(ort_bpm); }

extern "C" int ort_bpm_bessel(const ort_bpm* p, double* intensity) {
    if (!p || !intensity) {
        ort_set_error("ort_bpm_bessel: null argument");
        return ORT_EINVAL;
    }
    if (p->nxy < 2 || p->nxy > 16384 || p->nz < 1 || !(p->wavelength > 0.0) || !(p->xymax > 0.0) ||
        !(p->ring_width > 0.0) || !(p->w0 > 0.0) || !(p->axicon_deg > 0.0) || !(p->n_axicon > 1.0)) {
        ort_set_error("ort_bpm_bessel: parameters out of range");
        return ORT_EINVAL;
    }
    const int dev = ort_internal_primary_device();
    if (dev < 0) {
        ort_set_error("ort_bpm_bessel: library not initialised (ort_init / ort_init_rank)");
        return ORT_ENODEVICE;
    }
    if (!load_cufft()) {
        ort_set_error("ort_bpm_bessel: cannot load cuFFT (libcufft.so.11): %s", dlerror());
        return ORT_ECUDA;
    }
    const int steps = p->steps >= 0 ? p->steps : p->nz / 10;
    BpmGrid g;
    g.k = 2.0 * M_PI / p->wavelength;
    const double k_r = g.k * (p->n_axicon - 1.0) * p->axicon_deg * M_PI / 360.0;
    const double L = 3.0 * (p->w0 * (g.k / k_r));
    g.R = L;
    g.dz = L / p->nz;
    g.nxy = p->nxy;
    g.nmid = p->nxy / 2;
    g.dx = p->xymax / p->nxy;
    g.half = p->xymax / 2;
    g.dk = (2.0 * M_PI / g.dx) / p->nxy;
    g.r0 = p->ring_radius;
    g.inv_w2 = 1.0 / (p->ring_width * p->ring_width);

    int rc = ORT_OK;
    const size_t n = (size_t)p->nxy * p->nxy;
    double2 *d_e = nullptr, *d_f = nullptr;
    double* d_out = nullptr;
    cudaStream_t st = nullptr;
    cufftHandle plan = 0;
    bool have_plan = false;
    const dim3 tb(128), grid2((p->nxy + 127) / 128, p->nxy);
    BK(cudaSetDevice(dev));
    BK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    BK(cudaMalloc(&d_e, n * sizeof(double2)));
    BK(cudaMalloc(&d_f, n * sizeof(double2)));
    BK(cudaMalloc(&d_out, n * sizeof(double)));
    FK(g_fft.Plan2d(&plan, p->nxy, p->nxy, CUFFT_Z2Z));
    have_plan = true;
    FK(g_fft.SetStream(plan, st));
    bpm_setup_kernel<<<grid2, tb, 0, st>>>(g, d_e, d_f);
    BK(cudaGetLastError());
    for (int s = 0; s < steps; ++s) {
        FK(g_fft.ExecZ2Z(plan, (cufftDoubleComplex*)d_e, (cufftDoubleComplex*)d_e, CUFFT_FORWARD));
        bpm_multiply_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, d_e, d_f);
        FK(g_fft.ExecZ2Z(plan, (cufftDoubleComplex*)d_e, (cufftDoubleComplex*)d_e, CUFFT_INVERSE));
    }
    BK(cudaGetLastError());
    bpm_finish_kernel<<<grid2, tb, 0, st>>>(g, d_e, d_out);
    BK(cudaGetLastError());
    BK(cudaMemcpyAsync(intensity, d_out, n * sizeof(double), cudaMemcpyDeviceToHost, st));
    BK(cudaStreamSynchronize(st));
done:
    if (have_plan) g_fft.Destroy(plan);
    if (d_e) cudaFree(d_e);
    if (d_f) cudaFree(d_f);
    if (d_out) cudaFree(d_out);
    if (st) cudaStreamDestroy(st);
    return rc;
}

/* compute and write `bessel-normal.dat`-style raw fp64 (bpm.py:203-205) */
extern "C" int ort_bpm_write_file(const ort_bpm* p, const char* path) {
    if (!p || !path) return ORT_EINVAL;
    std::vector<double> img((size_t)(p->nxy > 0 ? p->nxy : 0) * (p->nxy > 0 ? p->nxy : 0));
    int rc = ort_bpm_bessel(p, img.data());
    if (rc != ORT_OK) return rc;
    FILE* f = fopen(path, "wb");
    if (!f) {
        ort_set_error("ort_bpm_write_file: cannot open %s", path);
        return ORT_EIO;
    }
    size_t w = fwrite(img.data(), sizeof(double), img.size(), f);
    if (fclose(f) != 0 || w != img.size()) {
        ort_set_error("ort_bpm_write_file: short write to %s", path);
        return ORT_EIO;
    }
    return ORT_OK;
}
