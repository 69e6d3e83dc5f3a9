// FFMA vs FFMA2 (fma.rn.f32x2, sm_100a) issue rate: N independent chains per thread, 12 warps per SMSP.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2 ffma2.cu && ./ffma2
#include <cstdio>
#include <cuda_runtime.h>
template <int PACKED, int CH>
__global__ void __launch_bounds__(256) k(float* out, float a, float b, int iters) {
    float x[CH], y[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) { x[c] = threadIdx.x * 1e-3f + c; y[c] = x[c] + 0.5f; }
    unsigned long long A, B;
    asm("mov.b64 %0, {%1, %1};" : "=l"(A) : "f"(a));
    asm("mov.b64 %0, {%1, %1};" : "=l"(B) : "f"(b));
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            if (PACKED) {
                unsigned long long v;
                asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(x[c]), "f"(y[c]));
                asm("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(v) : "l"(A), "l"(B));
                asm("mov.b64 {%0, %1}, %2;" : "=f"(x[c]), "=f"(y[c]) : "l"(v));
            } else {
                x[c] = fmaf(x[c], a, b);
                y[c] = fmaf(y[c], a, b);
            }
        }
    }
    float s = 0;
#pragma unroll
    for (int c = 0; c < CH; ++c) s += x[c] + y[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int PACKED, int CH>
double run(int sms, int iters) {
    float* d; cudaMalloc(&d, sms * 6 * 256 * sizeof(float));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<PACKED, CH><<<sms * 6, 256>>>(d, 0.999f, 1e-3f, 16);
    cudaEventRecord(e0);
    k<PACKED, CH><<<sms * 6, 256>>>(d, 0.999f, 1e-3f, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    cudaFree(d);
    return ms;
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const int iters = 1 << 16, CH = 8;
    for (int rep = 0; rep < 2; ++rep) {
        double ms0 = run<0, CH>(p.multiProcessorCount, iters), ms1 = run<1, CH>(p.multiProcessorCount, iters);
        double fl = (double)p.multiProcessorCount * 6 * 256 * iters * CH * 2;  // fma per launch
        double cyc0 = ms0 * 1e-3 * clk * 1e3, cyc1 = ms1 * 1e-3 * clk * 1e3;
        double smsp = p.multiProcessorCount * 4.0;
        printf("FFMA : %.3f ms  %.2f TFLOP/s  %.3f warp-inst/clk/SMSP\n", ms0, 2 * fl / ms0 * 1e-9, fl / 32 / cyc0 / smsp);
        printf("FFMA2: %.3f ms  %.2f TFLOP/s  %.3f warp-inst/clk/SMSP (each = 2 fma per lane)\n", ms1, 2 * fl / ms1 * 1e-9, fl / 2 / 32 / cyc1 / smsp);
    }
    return 0;
}
