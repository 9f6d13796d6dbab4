// fp64_occupancy.cu -- DFMA issue rate of one SM as a function of resident warps and independent chains per thread:
// how much thread- and instruction-level parallelism the B200 fp64 pipe needs before it saturates (the reduce-only
// kernel runs 12 warps per SM with 12 chains per thread).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/fp64_occupancy tools/fp64_occupancy.cu
#include <cuda_runtime.h>
#include <cstdio>

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

template <int ILP>
double run(int warps_per_sm, int sms, double* out) {
    const int iters = 20000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    double best = 0;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        dfma_kernel<ILP><<<sms, warps_per_sm * 32>>>(out, 0.999999, 1e-7, iters);   // one block per SM
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        const double rate = double(sms) * warps_per_sm * 32 * ILP * double(iters) / (ms * 1e-3);
        if (rep >= 1 && rate > best) best = rate;
    }
    return best;
}

int main() {
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    double* out;
    cudaMalloc(&out, 8);
    const int sms = prop.multiProcessorCount;
    printf("# DFMA thread-instructions per second (1e12), one block per SM; rows: warps per SM, columns: independent chains per thread\n");
    printf("# warps   ILP1    ILP2    ILP4    ILP8   ILP12   ILP16\n");
    const int ws[] = {4, 8, 12, 16, 24, 32};
    for (int w : ws) {
        printf("%6d %7.2f %7.2f %7.2f %7.2f %7.2f %7.2f\n", w, run<1>(w, sms, out) / 1e12, run<2>(w, sms, out) / 1e12, run<4>(w, sms, out) / 1e12,
               run<8>(w, sms, out) / 1e12, run<12>(w, sms, out) / 1e12, run<16>(w, sms, out) / 1e12);
    }
    return 0;
}
