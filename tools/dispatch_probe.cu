// dispatch_probe.cu -- do integer (IMAD) and fp64 (DFMA) warp-instructions overlap on one B200 scheduler, or does each
// hold the dispatch port for its two 16-lane passes?  Times N DFMA, N IMAD, and N DFMA + N IMAD interleaved in one loop.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/dispatch_probe tools/dispatch_probe.cu
#include <cuda_runtime.h>
#include <cstdio>

template <int ND, int NI, int NF>
__global__ void k(double* out, double a, double b, unsigned m, float fa, int iters) {
    double x[8]; unsigned u[8]; float f[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { x[i] = threadIdx.x * 1e-9 + i; u[i] = threadIdx.x + i; f[i] = threadIdx.x * 1e-3f + i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (i < ND) x[i] = fma(x[i], a, b);
            if (i < NI) u[i] = u[i] * m + 12345u;
            if (i < NF) f[i] = fmaf(f[i], fa, 0.5f);
        }
    }
    double s = 0; unsigned t = 0; float g = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { s += x[i]; t += u[i]; g += f[i]; }
    if (s == 12345.678 || t == 77u || g == 3.25f) out[0] = s + t + g;
}

template <int ND, int NI, int NF>
double run(int sms, double* out) {
    const int iters = 20000, warps = 16;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        k<ND, NI, NF><<<sms, warps * 32>>>(out, 0.999999, 1e-7, 1664525u, 0.9999f, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep >= 1 && ms < best) best = ms;
    }
    // cycles per loop iteration per scheduler at 1.9 GHz nominal: 4 warps per scheduler, each issuing (ND + NI + NF) instructions
    return best * 1e-3 * 1.9e9 / iters / 4.0;
}

int main() {
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    double* out; cudaMalloc(&out, 8);
    const int sms = prop.multiProcessorCount;
    printf("# cycles (at 1.9 GHz) per warp per loop iteration on one scheduler (16 warps per SM); 8 instructions of each kind per iteration\n");
    printf("8 DFMA                 : %6.2f\n", run<8, 0, 0>(sms, out));
    printf("8 IMAD                 : %6.2f\n", run<0, 8, 0>(sms, out));
    printf("8 FFMA                 : %6.2f\n", run<0, 0, 8>(sms, out));
    printf("8 DFMA + 8 IMAD        : %6.2f\n", run<8, 8, 0>(sms, out));
    printf("8 DFMA + 8 FFMA        : %6.2f\n", run<8, 0, 8>(sms, out));
    printf("8 IMAD + 8 FFMA        : %6.2f\n", run<0, 8, 8>(sms, out));
    printf("8 DFMA + 8 IMAD + 8 FFMA: %6.2f\n", run<8, 8, 8>(sms, out));
    return 0;
}
