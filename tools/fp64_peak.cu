// fp64_peak.cu -- measures the sustained fp64 vector (DFMA) issue rate of the GPU, the second roofline
// denominator of the plume kernels (MEASURED_PEAKS.json only holds HBM and bf16 numbers).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/fp64_peak tools/fp64_peak.cu
// Prints one JSON line: DFMA thread-instructions per second, TFLOP/s (2 flop per DFMA), SM clock seen.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

template <int ILP>
__global__ void dfma_kernel(double* out, double a, double b, int iters) {
    double x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 1e-9 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i];
    if (s == 12345.678) out[0] = s;
}

int main(int argc, char** argv) {
    int dev = 0;
    cudaSetDevice(dev);
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, dev);
    double* out;
    cudaMalloc(&out, 8);
    const int threads = 256, blocks = prop.multiProcessorCount * 8, iters = 20000;
    constexpr int ILP = 8;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    double best = 0;
    for (int rep = 0; rep < 6; ++rep) {
        cudaEventRecord(e0);
        dfma_kernel<ILP><<<blocks, threads>>>(out, 0.999999, 1e-7, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        double inst = double(blocks) * threads * ILP * double(iters);
        double rate = inst / (ms * 1e-3);
        if (rep >= 1 && rate > best) best = rate;
    }
    // sustained: back-to-back launches for ~2 s
    int launches = 0;
    cudaEventRecord(e0);
    float ms = 0;
    do {
        for (int i = 0; i < 10; ++i) dfma_kernel<ILP><<<blocks, threads>>>(out, 0.999999, 1e-7, iters);
        launches += 10;
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
    } while (ms < 2000.f);
    double sustained = double(launches) * blocks * threads * ILP * double(iters) / (ms * 1e-3);
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, dev);
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"dfma_per_s_burst\": %.4e, \"dfma_per_s_sustained\": %.4e, "
           "\"fp64_tflops_burst\": %.2f, \"fp64_tflops_sustained\": %.2f, \"dfma_per_clk_per_sm_at_max_clock\": %.2f, "
           "\"max_clock_khz\": %d}\n",
           prop.name, prop.multiProcessorCount, best, sustained, 2 * best / 1e12, 2 * sustained / 1e12,
           best / (double(clk) * 1e3) / prop.multiProcessorCount, clk);
    return 0;
}
