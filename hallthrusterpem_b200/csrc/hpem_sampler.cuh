// hpem_sampler.cuh -- counter-based on-device sampler for the PEM v0 input priors.
//
// The step immediately before the hot path in the reference is amisc's `system.sample_inputs(N)` over the YAML priors
// (/root/reference/scripts/gen_data.py:238, scripts/pem_v0/pem_v0_SPT-100.yml: U(a,b), Uniform(a,b), LogUniform(a,b),
// Normal(mu,sigma), Relative(p) = U(nominal(1-p/100), nominal(1+p/100))).  Here every (sample index, input) pair owns
// a fixed slice of a Philox4x32-10 stream, so a sample's inputs do not depend on which GPU, launch or chunk draws them:
// shards of one global index range reproduce the unsharded run bit for bit.
#pragma once
#include <stdint.h>

#include "hpem_device.cuh"

namespace hpem {

enum PriorKind { PRIOR_CONST = 0, PRIOR_UNIFORM = 1, PRIOR_LOGUNIFORM = 2, PRIOR_NORMAL = 3 };

struct Prior {      // struct hpem_prior + host-precomputed constants
    int32_t kind;
    int32_t reserved;
    double a, b;    // const: a | uniform: [a, b) | loguniform: [a, b) in the variable itself | normal: mean a, std b
    double log_a, log_ratio;   // loguniform: ln a, ln b - ln a (host std::log)
};

struct SamplerParams {
    Prior prior[15];
    unsigned long long seed;
    unsigned long long first_index;   // global index of sample 0 of this launch
};

// Philox4x32-10 (Salmon et al., SC'11): counter (c0..c3), key (k0, k1)
__host__ __device__ inline void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                              uint32_t out[4]) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        const uint32_t n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        const uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// two uniforms in [0, 1) with 53 random bits each, for (sample, pair) -- pair p serves inputs 2p and 2p+1
__host__ __device__ inline void uniform_pair(unsigned long long seed, unsigned long long sample, uint32_t pair, uint32_t stream,
                                             double& u0, double& u1) {
    uint32_t o[4];
    philox4x32_10((uint32_t)sample, (uint32_t)(sample >> 32), pair, stream, (uint32_t)seed, (uint32_t)(seed >> 32), o);
    const uint64_t a = ((uint64_t)o[1] << 32) | o[0], b = ((uint64_t)o[3] << 32) | o[2];
    u0 = (double)(a >> 11) * 0x1.0p-53;
    u1 = (double)(b >> 11) * 0x1.0p-53;
}

__device__ __forceinline__ double apply_prior(const Prior& pr, double u, unsigned long long seed, unsigned long long sample,
                                              uint32_t input) {
    switch (pr.kind) {
        case PRIOR_UNIFORM: return fma(u, pr.b - pr.a, pr.a);
        case PRIOR_LOGUNIFORM: return exp(fma(u, pr.log_ratio, pr.log_a));
        case PRIOR_NORMAL: {   // Box-Muller with a second uniform from stream 1 of the same (sample, input)
            double v0, v1;
            uniform_pair(seed, sample, input, 1u, v0, v1);
            const double r = sqrt(-2.0 * log(1.0 - u));          // 1-u in (0, 1]
            return fma(r * cospi(2.0 * v0), pr.b, pr.a);
        }
        default: return pr.a;
    }
}

// all 15 inputs of one sample
__device__ __forceinline__ void sample_inputs(const SamplerParams& sp, unsigned long long local_index, double x[15]) {
    const unsigned long long sample = sp.first_index + local_index;
#pragma unroll
    for (uint32_t pair = 0; pair < 8; ++pair) {
        double u0, u1;
        uniform_pair(sp.seed, sample, pair, 0u, u0, u1);
        x[2 * pair] = apply_prior(sp.prior[2 * pair], u0, sp.seed, sample, 2 * pair);
        if (2 * pair + 1 < 15) x[2 * pair + 1] = apply_prior(sp.prior[2 * pair + 1], u1, sp.seed, sample, 2 * pair + 1);
    }
}

__global__ void __launch_bounds__(256) sample_inputs_kernel(const SamplerParams sp, long long n, double* o0, double* o1,
                                                            double* o2, double* o3, double* o4, double* o5, double* o6,
                                                            double* o7, double* o8, double* o9, double* o10, double* o11,
                                                            double* o12, double* o13, double* o14) {
    double* out[15] = {o0, o1, o2, o3, o4, o5, o6, o7, o8, o9, o10, o11, o12, o13, o14};
    for (long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x; s < n; s += (long long)gridDim.x * blockDim.x) {
        double x[15];
        sample_inputs(sp, (unsigned long long)s, x);
#pragma unroll
        for (int k = 0; k < 15; ++k)
            if (out[k]) out[k][s] = x[k];
    }
}

}  // namespace hpem
