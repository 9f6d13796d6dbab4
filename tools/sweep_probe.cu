// sweep_probe.cu -- the recurrence sweep of the plume kernels in isolation (no prologue, no stores): how close the B200
// fp64 pipe gets to its DFMA peak on THIS instruction mix (4 DMUL + 2 DADD + 2 DFMA per evaluation, all operands in
// registers, weights broadcast from shared memory), as a function of samples per thread and warps per SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o build/sweep_probe tools/sweep_probe.cu
#include <cuda_runtime.h>
#include <cstdio>

template <int NS, bool WEIGHTS>
__global__ void sweep_kernel(double* out, int chunks, double seed) {
    __shared__ double2 w[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) w[i] = make_double2(1.0 / (i + 1), 0.5 / (i + 2));
    __syncthreads();
    double e1[NS], e2[NS], r1[NS], r2[NS], q1[NS], q2[NS], jc[NS], num[NS], den[NS], acc[NS];
#pragma unroll
    for (int s = 0; s < NS; ++s) {
        e1[s] = 1.0 + seed * threadIdx.x; e2[s] = 0.5 + seed * s; r1[s] = 0.999; r2[s] = 0.9999; q1[s] = 0.99999; q2[s] = 0.999999;
        jc[s] = 1e-3; num[s] = den[s] = acc[s] = 0.0;
    }
    for (int c = 0; c < chunks; ++c) {
        const int i0 = (c & 15) * 16;
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
            const double2 ww = WEIGHTS ? w[i0 + kk] : make_double2(0.25, 0.125);
#pragma unroll
            for (int s = 0; s < NS; ++s) {
                const double sum = e1[s] + e2[s];
                den[s] = fma(ww.x, sum, den[s]);
                num[s] = fma(ww.y, sum, num[s]);
                acc[s] += sum + jc[s];
                e1[s] *= r1[s]; r1[s] *= q1[s];
                e2[s] *= r2[s]; r2[s] *= q2[s];
            }
        }
    }
    double t = 0;
#pragma unroll
    for (int s = 0; s < NS; ++s) t += num[s] + den[s] + acc[s] + e1[s] + r2[s];
    if (t == 12345.678) out[0] = t;
}

template <int NS, bool WEIGHTS>
double run(int warps, int sms, double* out) {
    const int chunks = 4000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    double best = 0;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        sweep_kernel<NS, WEIGHTS><<<sms, warps * 32>>>(out, chunks, 1e-9);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double rate = double(sms) * warps * 32 * NS * 16.0 * chunks * 9.0 / (ms * 1e-3);   // 9 fp64 instr per evaluation here
        if (rep >= 1 && rate > best) best = rate;
    }
    return best;
}

int main() {
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    double* out; cudaMalloc(&out, 8);
    const int sms = prop.multiProcessorCount;
    printf("# fp64 thread-instructions per second (1e12) of the sweep loop (9 per evaluation: 4 DMUL, 3 DADD, 2 DFMA); DFMA peak 18.5\n");
    printf("# warps  1 sample/thr  2 samples/thr  2 samples no-LDS  4 samples/thr\n");
    const int ws[] = {4, 8, 12, 16, 24};
    for (int w : ws)
        printf("%6d %12.2f %13.2f %16.2f %13.2f\n", w, run<1, true>(w, sms, out) / 1e12, run<2, true>(w, sms, out) / 1e12,
               run<2, false>(w, sms, out) / 1e12, (w <= 12 ? run<4, true>(w, sms, out) : 0.0) / 1e12);
    return 0;
}
