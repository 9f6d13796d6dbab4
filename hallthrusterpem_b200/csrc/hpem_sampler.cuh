// hpem_sampler.cuh -- counter-based on-device sampler for the PEM v0 input priors.
//
// The step immediately before the hot path in the reference is amisc's `system.sample_inputs(N)` over the YAML priors
// (/root/reference/scripts/gen_data.py:238, scripts/pem_v0/pem_v0_SPT-100.yml: U(a,b), Uniform(a,b), LogUniform(a,b),
// Normal(mu,sigma), Relative(p) = U(nominal(1-p/100), nominal(1+p/100))).  Here every (sample index, input) pair owns
// a fixed slice of a Philox4x32-10 stream, so a sample's inputs do not depend on which GPU, launch or chunk draws them:
// shards of one global index range reproduce the unsharded run bit for bit.
#pragma once
#include <stdint.h>
#include <string.h>

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
    // value = scale * u + offset, then exp() for the inputs in log_mask, Box-Muller for those in normal_mask (host-filled)
    double scale[15], offset[15];
    uint32_t log_mask, normal_mask;
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

// two uniforms in [0, 1) with 53 random bits each for (sample, pair, stream): the second stream of the Normal prior
__host__ __device__ inline void uniform_pair(unsigned long long seed, unsigned long long sample, uint32_t pair, uint32_t stream,
                                             double& u0, double& u1) {
    uint32_t o[4];
    philox4x32_10((uint32_t)sample, (uint32_t)(sample >> 32), pair, stream, (uint32_t)seed, (uint32_t)(seed >> 32), o);
    const uint64_t a = ((uint64_t)o[1] << 32) | o[0], b = ((uint64_t)o[3] << 32) | o[2];
#if defined(__CUDA_ARCH__)
    // (double)(a >> 11) * 2^-53 without the 64-bit integer conversion: the 53 bits are split into a 21-bit and a 32-bit
    // part, each dropped into the low mantissa word of a power of two (2^84, 2^52); the three operations below are exact
    const uint64_t ka = a >> 11, kb = b >> 11;
    const double ha = __hiloint2double(0x45300000, (int)(ka >> 32)) - 0x1.0p84, la = __hiloint2double(0x43300000, (int)(uint32_t)ka) - 0x1.0p52;
    const double hb = __hiloint2double(0x45300000, (int)(kb >> 32)) - 0x1.0p84, lb = __hiloint2double(0x43300000, (int)(uint32_t)kb) - 0x1.0p52;
    u0 = (ha + la) * 0x1.0p-53;
    u1 = (hb + lb) * 0x1.0p-53;
#else
    u0 = (double)(a >> 11) * 0x1.0p-53;
    u1 = (double)(b >> 11) * 0x1.0p-53;
#endif
}

// Three uniforms in [0, 1) with 42 random bits each from ONE Philox call, for (sample, triple) -- triple t serves inputs
// 3t, 3t+1, 3t+2, so the 15 inputs of a sample cost five calls.  Uniform j is k_j 2^-42 with the 42-bit integer
// k_j = (bits 10j..10j+9 of word 3) << 32 | word j.  k_j << 10 is dropped into the mantissa of 1.0 (two integer
// operations per word), and ONE exact subtraction gives u; the affine map of the prior is then the usual fma(u, b - a, a),
// which keeps every draw inside [a, b).  (The first version took two 53-bit uniforms per call: eight calls per sample and
// five fp64 operations per uniform, a quarter of the reduce-only kernel's per-sample work.)
__host__ __device__ inline void triple_from_words(const uint32_t (&o)[4], double (&u)[3]) {
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const uint32_t top = j == 0 ? (o[3] << 10) : (j == 1 ? o[3] : (o[3] >> 10));   // 10 bits of word 3 at bits 10..19
        const uint32_t hi = 0x3FF00000u | (top & 0x000FFC00u) | (o[j] >> 22), lo = o[j] << 10;
#if defined(__CUDA_ARCH__)
        u[j] = __hiloint2double((int)hi, (int)lo) - 1.0;
#else
        const uint64_t bits = ((uint64_t)hi << 32) | lo;
        double d;
        memcpy(&d, &bits, sizeof(d));
        u[j] = d - 1.0;
#endif
    }
}
__host__ __device__ inline void uniform_triple(unsigned long long seed, unsigned long long sample, uint32_t triple, double (&u)[3]) {
    uint32_t o[4];
    philox4x32_10((uint32_t)sample, (uint32_t)(sample >> 32), triple, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), o);
    triple_from_words(o, u);
}

// Normal(mean a, std b): Box-Muller with a second uniform from stream 1 of the same (sample, input).  Out of line: no prior
// of the PEM v0 plume/cathode inputs is normal, and fifteen inlined copies made up a quarter of the reduce-only kernel's code.
__device__ __noinline__ double normal_prior(double a, double b, double u, unsigned long long seed, unsigned long long sample,
                                            uint32_t input) {
    double v0, v1;
    uniform_pair(seed, sample, input, 1u, v0, v1);
    const double r = sqrt(-2.0 * log(1.0 - u));          // 1-u in (0, 1]
    return fma(r * cospi(2.0 * v0), b, a);
}

__device__ __forceinline__ double apply_prior(const Prior& pr, double u, unsigned long long seed, unsigned long long sample,
                                              uint32_t input) {
    switch (pr.kind) {
        case PRIOR_UNIFORM: return fma(u, pr.b - pr.a, pr.a);
        case PRIOR_LOGUNIFORM: return fm_exp(fma(u, pr.log_ratio, pr.log_a));   // branch-free, < 1 ulp (hpem_fastmath.cuh)
        case PRIOR_NORMAL: return normal_prior(pr.a, pr.b, u, seed, sample, input);
        default: return pr.a;
    }
}

// All 15 inputs of N samples (N = 1, or the two samples a thread of the reduce-only kernel owns).  The 5 N Philox calls come
// first, in ONE basic block, so their ten-round dependency chains interleave; then the affine maps; the exponentials of
// the LogUniform inputs (and the rare Normal inputs) last, behind warp-uniform tests of the prior masks.  Values are
// exactly those of apply_prior() applied input by input.
// The generator half and the transform half are separate so that the reduce-only kernel can draw the NEXT batch's words
// next to the latency-bound end of the current batch (integer work beside fp64 dependency chains).
constexpr uint32_t kLogCandidates = (1u << 0) | (1u << 10) | (1u << 11);   // P_b, c4, c5: the LogUniform inputs of PEM v0
template <int N>
__device__ __forceinline__ void sample_words_n(const SamplerParams& sp, const unsigned long long (&local_index)[N], uint32_t (&o)[N][5][4]) {
    // the 5 N calls advance together, one round per trip of a ROLLED loop: the same ten-round chains interleaved, a tenth of
    // the code (the straight-line version was 10 KB of the reduce-only kernel's instruction stream; same speed within 2 %)
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int s = 0; s < N; ++s) {
        const unsigned long long sample = sp.first_index + local_index[s];
#pragma unroll
        for (uint32_t t = 0; t < 5; ++t) {
            o[s][t][0] = (uint32_t)sample;
            o[s][t][1] = (uint32_t)(sample >> 32);
            o[s][t][2] = t;
            o[s][t][3] = 0u;
        }
    }
    uint32_t k0 = (uint32_t)sp.seed, k1 = (uint32_t)(sp.seed >> 32);
#pragma unroll 1
    for (int r = 0; r < 10; ++r) {
#pragma unroll
        for (int s = 0; s < N; ++s) {
#pragma unroll
            for (int t = 0; t < 5; ++t) {
                const uint64_t p0 = (uint64_t)M0 * o[s][t][0], p1 = (uint64_t)M1 * o[s][t][2];
                const uint32_t n0 = (uint32_t)(p1 >> 32) ^ o[s][t][1] ^ k0;
                const uint32_t n2 = (uint32_t)(p0 >> 32) ^ o[s][t][3] ^ k1;
                o[s][t][0] = n0;
                o[s][t][1] = (uint32_t)p1;
                o[s][t][2] = n2;
                o[s][t][3] = (uint32_t)p0;
            }
        }
        k0 += W0;
        k1 += W1;
    }
}
template <int N>
__device__ __forceinline__ void sample_transform_n(const SamplerParams& sp, const unsigned long long (&local_index)[N], const uint32_t (&o)[N][5][4],
                                                   double (&x)[N][15]) {
    double u[N][15];
#pragma unroll
    for (int t = 0; t < 5; ++t) {
#pragma unroll
        for (int s = 0; s < N; ++s) {
            double v[3];
            triple_from_words(o[s][t], v);
            u[s][3 * t] = v[0];
            u[s][3 * t + 1] = v[1];
            u[s][3 * t + 2] = v[2];
        }
    }
#pragma unroll
    for (int k = 0; k < 15; ++k) {
#pragma unroll
        for (int s = 0; s < N; ++s) x[s][k] = fma(u[s][k], sp.scale[k], sp.offset[k]);
    }
    // LogUniform: the inputs that are LogUniform in PEM v0 get their exponentials unconditionally, in ONE basic block (3 N
    // independent chains; fm_exp is branch-free and harmless on any input), selected by the mask; any other input with a
    // LogUniform prior takes a warp-uniform branch of its own
#pragma unroll
    for (int k = 0; k < 15; ++k) {
        if (kLogCandidates & (1u << k)) {
#pragma unroll
            for (int s = 0; s < N; ++s) {
                const double e = fm_exp(x[s][k]);
                x[s][k] = (sp.log_mask & (1u << k)) ? e : x[s][k];
            }
        }
    }
    if (sp.log_mask & ~kLogCandidates) {
#pragma unroll
        for (int k = 0; k < 15; ++k) {
            if (!(kLogCandidates & (1u << k)) && (sp.log_mask & (1u << k))) {
#pragma unroll
                for (int s = 0; s < N; ++s) x[s][k] = fm_exp(x[s][k]);   // branch-free, < 1 ulp (hpem_fastmath.cuh)
            }
        }
    }
    if (sp.normal_mask) {
#pragma unroll
        for (int k = 0; k < 15; ++k) {
            if (sp.normal_mask & (1u << k)) {
#pragma unroll
                for (int s = 0; s < N; ++s)
                    x[s][k] = normal_prior(sp.prior[k].a, sp.prior[k].b, u[s][k], sp.seed, sp.first_index + local_index[s], k);
            }
        }
    }
}
template <int N>
__device__ __forceinline__ void sample_inputs_n(const SamplerParams& sp, const unsigned long long (&local_index)[N], double (&x)[N][15]) {
    uint32_t o[N][5][4];
    sample_words_n<N>(sp, local_index, o);
    sample_transform_n<N>(sp, local_index, o, x);
}
__device__ __forceinline__ void sample_inputs(const SamplerParams& sp, unsigned long long local_index, double x[15]) {
    const unsigned long long idx[1] = {local_index};
    double xx[1][15];
    sample_inputs_n<1>(sp, idx, xx);
#pragma unroll
    for (int k = 0; k < 15; ++k) x[k] = xx[0][k];
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
