// hpem_kernels.cuh -- sm_100a kernels for the fused cathode + plume sample batch.
//
// K1u  eval_uniform_kernel : one THREAD per sample, uniform angle grid, single radius.
//        The two Gaussian beam profiles exp(-(i h / a)^2), i = 0..A-1, are advanced with a two-term
//        multiplicative recurrence (e *= r; r *= q) restarted hierarchically, so the per-(sample, angle)
//        cost is 4 DMUL + 2 DADD + 2 DFMA + 1 DSETP instead of two fp64 exp() calls (~40 fp64-pipe ops).
//        The kernel is then bound by the j_ion store stream (8 B per evaluation) -> HBM roofline.
//        The warp's 32 samples x 16 angles are transposed through a private shared-memory tile so
//        the global stores are full 128-byte row segments (streaming, evict-first).
//        Quadrature sums are thread-local FMAs against weights broadcast from shared memory.
// K1d  eval_direct_kernel  : one WARP per sample, any angle grid, any number of radii; evaluates
//        the reference's expressions in the reference's operation order (divide, square, negate, exp);
//        warp-shuffle reductions for the two Simpson sums.  Fallback + independent cross-check.
//
// Reference lines reproduced: plume.py:95-127,136-140 (per angle / per sample epilogue),
// cathode.py:26-37 and plume.py:40-85 via hpem_device.cuh.
#pragma once
#include <cuda.h>
#include <stdint.h>

#include <type_traits>

#include "hpem_device.cuh"

namespace hpem {

constexpr int kNumInputs = 15;
enum InputId {
    IN_P_b = 0, IN_V_a, IN_T_e, IN_V_vac, IN_Pstar, IN_P_T,
    IN_c0, IN_c1, IN_c2, IN_c3, IN_c4, IN_c5, IN_sigma, IN_I_B0, IN_T
};

struct EvalParams {
    const double* in[kNumInputs];  // device pointers (already offset to the first sample) or nullptr
    double scalar[kNumInputs];     // broadcast value when in[k] == nullptr
    long long n;                   // samples in this launch
    double torr;
    // outputs (nullptr = not wanted)
    double* v_cc;
    double* j_ion;
    double* div_angle;
    double* t_c;
    double* cos_div;
    uint8_t* invalid;
    // grid constants
    int n_angles;
    int n_angles_pad;        // multiple of kChunk, weights zero-padded
    int n_radii;
    const double2* w;        // (wd[i], wn[i]) interleaved, length n_angles_pad (device)
    const double* alpha;     // angle grid (device)
    const double* radii;     // radii (device)
    double h;                // uniform step alpha[1] (uniform kernel only)
    double radius0;          // radii[0]
    bool has_thrust;         // input T supplied
};

constexpr int kChunk = 16;          // angles per staged tile / inner recurrence length
constexpr int kRestartChunks = 16;  // exact exp() restart every kRestartChunks*kChunk angles
constexpr int kTilePitch = kChunk + 1;
constexpr int kThreadsU = 128;
constexpr int kWarpsU = kThreadsU / 32;
constexpr double kInvalidFill = 1e-20;  // plume.py:106

__device__ __forceinline__ double load_in(const EvalParams& p, int k, long long s) {
    return p.in[k] ? __ldg(p.in[k] + s) : p.scalar[k];
}

// ---------------------------------------------------------------------------------------------
// TMA helpers (cp.async.bulk.tensor store, shared::cta -> global)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t smem_src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(map), "r"(c0),
                 "r"(c1), "r"(smem_src)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void tma_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void tma_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------
// K1u: thread per sample, uniform grid, R == 1
// ---------------------------------------------------------------------------------------------
constexpr int kTmaTileBytes = 32 * kChunk * 8;  // 32 samples x 16 angles, dense 128-byte rows (SWIZZLE_128B)
constexpr int kTmaBuffers = 2;

struct BeamState {  // per-thread recurrence state of one Gaussian beam
    double ec, rc, gc;     // chunk-start profile value, chunk-start ratio, chunk-to-chunk factor
    double q, qk, hh;      // per-sample constants exp(-2x), exp(-2Kx), exp(-2K^2 x)
    double x, amp;
};

__device__ __forceinline__ void beam_init(BeamState& b, double h, double a, double amp) {
    const double t = h / a;
    b.x = t * t;                           // profile(i) = exp(-x i^2)
    b.amp = amp;
    b.q = exp(-2.0 * b.x);
    b.qk = exp(-(2.0 * kChunk) * b.x);
    b.hh = exp(-(2.0 * kChunk * kChunk) * b.x);
    b.rc = exp(-b.x);
    b.gc = exp(-double(kChunk * kChunk) * b.x);
    b.ec = 1.0;
}
__device__ __forceinline__ void beam_restart(BeamState& b, int i0) {  // exact values at angle index i0
    const double di = double(i0);
    b.ec = exp(-b.x * (di * di));
    b.rc = exp(-b.x * (2.0 * di + 1.0));
    b.gc = exp(-b.x * (2.0 * kChunk * di + double(kChunk * kChunk)));
}
__device__ __forceinline__ void beam_next_chunk(BeamState& b) {
    b.ec *= b.gc;
    b.gc *= b.hh;
    b.rc *= b.qk;
}

template <bool WANT_PLUME, bool STORE_J, bool USE_TMA>
__global__ void __launch_bounds__(kThreadsU) eval_uniform_kernel(const EvalParams p,
                                                                 const __grid_constant__ CUtensorMap jmap) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    // layout: [staging tiles (1024-byte aligned for the 128B TMA swizzle)] [fused weights]
    unsigned char* smem_al = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    constexpr int kStageBytesPerWarp = USE_TMA ? kTmaBuffers * kTmaTileBytes : 32 * kTilePitch * 8;
    constexpr int kStageBytes = STORE_J ? kWarpsU * kStageBytesPerWarp : 0;
    double2* wsm = reinterpret_cast<double2*>(smem_al + kStageBytes);

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    unsigned char* stage = smem_al + warp * kStageBytesPerWarp;

    if (WANT_PLUME) {
        for (int i = threadIdx.x; i < p.n_angles_pad; i += kThreadsU) wsm[i] = p.w[i];
        __syncthreads();
    }

    const long long s_raw = (long long)blockIdx.x * kThreadsU + threadIdx.x;
    const bool active = s_raw < p.n;
    const long long s = active ? s_raw : p.n - 1;  // inactive lanes shadow the last sample, never store
    const long long warp_s0 = s_raw - lane;
    if (warp_s0 >= p.n) return;  // whole warp out of range (after the only __syncthreads)

    const double p_b = load_in(p, IN_P_b, s);

    if (p.v_cc) {
        const double v = cathode_vcc(p_b, load_in(p, IN_V_a, s), load_in(p, IN_T_e, s), load_in(p, IN_V_vac, s),
                                     load_in(p, IN_Pstar, s), load_in(p, IN_P_T, s), p.torr);
        if (active) p.v_cc[s] = v;
    }
    if (!WANT_PLUME) return;

    const SampleConsts k = plume_sample_consts(p_b, load_in(p, IN_c0, s), load_in(p, IN_c1, s), load_in(p, IN_c2, s),
                                               load_in(p, IN_c3, s), load_in(p, IN_c4, s), load_in(p, IN_c5, s), p.torr);
    double j_cex, base;
    cex_terms(k.density, load_in(p, IN_sigma, s), load_in(p, IN_I_B0, s), p.radius0, j_cex, base);

    BeamState b1, b2;
    beam_init(b1, p.h, k.a1, __dmul_rn(base, k.amp1));   // (base_density * A1), plume.py:99
    beam_init(b2, p.h, k.a2, __dmul_rn(base, k.amp2));   // (base_density * A2), plume.py:100

    const bool known_invalid = (k.a1 <= 0.0);  // plume.py:105 first term
    // With non-negative beam amplitudes and a positive CEX floor every j_ion is > 0 (or NaN), so the per-angle
    // `j_ion <= 0` test of plume.py:105 cannot fire; only warps holding an exceptional sample run the checked loop.
    const bool needs_check = known_invalid || !(b1.amp >= 0.0 && b2.amp >= 0.0 && j_cex > 0.0);
    bool bad = false;
    double num = 0.0, den = 0.0;

    const int A = p.n_angles;
    const int n_chunks = p.n_angles_pad / kChunk;
    const int rows_valid = (int)min((long long)32, p.n - warp_s0);
    const int col = lane & (kChunk - 1);
    const int rsub = lane >> 4;

    auto chunk_loop = [&](auto checked_tag) {
        constexpr bool CHECKED = decltype(checked_tag)::value;
        for (int c = 0; c < n_chunks; ++c) {
            const int i0 = c * kChunk;
            if (c != 0 && (c % kRestartChunks) == 0) {  // exact restart bounds the recurrence error for large A
                beam_restart(b1, i0);
                beam_restart(b2, i0);
            }
            double e1 = b1.amp * b1.ec, e2 = b2.amp * b2.ec;
            double r1 = b1.rc, r2 = b2.rc;
            const int kcount = min(kChunk, A - i0);
            unsigned char* my_row = USE_TMA ? stage + (c & 1) * kTmaTileBytes + lane * (kChunk * 8)
                                            : stage + lane * (kTilePitch * 8);
            auto step = [&](double2 w, double& jout) {
                const double sum = e1 + e2;    // j_beam + j_scat
                const double j = sum + j_cex;  // plume.py:102
                den = fma(w.x, sum, den);
                num = fma(w.y, sum, num);
                if (CHECKED) {
                    bad |= (j <= 0.0);
                    jout = known_invalid ? kInvalidFill : j;
                } else {
                    jout = j;
                }
                e1 *= r1; r1 *= b1.q;
                e2 *= r2; r2 *= b2.q;
            };
            if (kcount == kChunk) {
#pragma unroll
                for (int kk = 0; kk < kChunk; kk += 2) {
                    double ja, jb;
                    step(wsm[i0 + kk], ja);
                    step(wsm[i0 + kk + 1], jb);
                    if (STORE_J) {
                        if (USE_TMA) {
                            *reinterpret_cast<double2*>(my_row + ((((kk >> 1) ^ (lane & 7))) << 4)) = make_double2(ja, jb);
                        } else {
                            reinterpret_cast<double*>(my_row)[kk] = ja;
                            reinterpret_cast<double*>(my_row)[kk + 1] = jb;
                        }
                    }
                }
            } else {
                for (int kk = 0; kk < kcount; ++kk) {
                    double ja;
                    step(wsm[i0 + kk], ja);
                    if (STORE_J) {
                        if (USE_TMA)
                            *reinterpret_cast<double*>(my_row + ((((kk >> 1) ^ (lane & 7))) << 4) + ((kk & 1) << 3)) = ja;
                        else
                            reinterpret_cast<double*>(my_row)[kk] = ja;
                    }
                }
            }
            beam_next_chunk(b1);
            beam_next_chunk(b2);

            if (STORE_J) {
                if (USE_TMA) {
                    // one 32x16 box per warp and chunk; rows >= n and columns >= A are clipped by the tensor map
                    fence_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_2d(&jmap, smem_u32(stage + (c & 1) * kTmaTileBytes), i0, (int)warp_s0);
                        tma_wait_read<kTmaBuffers - 1>();   // the other buffer is free again
                    }
                    __syncwarp();
                } else {
                    __syncwarp();
                    const double* trow = reinterpret_cast<const double*>(stage) + rsub * kTilePitch + col;
                    double* g = p.j_ion + (warp_s0 + rsub) * (long long)A + col + i0;
                    const bool col_ok = col < kcount;
#pragma unroll
                    for (int rr = 0; rr < 16; ++rr) {
                        if (col_ok && (2 * rr + rsub) < rows_valid) __stcs(g, trow[2 * rr * kTilePitch]);
                        g += 2 * (long long)A;
                    }
                    __syncwarp();
                }
            }
        }
    };
    if (__any_sync(0xffffffffu, needs_check))
        chunk_loop(std::true_type{});
    else
        chunk_loop(std::false_type{});

    if (STORE_J && USE_TMA) {
        if (lane == 0) tma_wait_all();
        fence_async_all();
        __syncwarp();
    }

    // per-sample epilogue: plume.py:124-127,137 (NOT masked by `invalid`)
    double cd = num / den;
    if (cd == CUDART_INF) cd = CUDART_NAN;  // plume.py:125
    const bool invalid = known_invalid || bad;
    if (active) {
        if (p.div_angle) p.div_angle[s] = acos(cd);
        if (p.cos_div) p.cos_div[s] = cd;
        if (p.t_c) p.t_c[s] = __dmul_rn(load_in(p, IN_T, s), cd);
        if (p.invalid) p.invalid[s] = invalid ? 1 : 0;
        if (STORE_J && bad && !known_invalid) {
            // rare: a non-positive j_ion found after earlier chunks were already written -> overwrite the
            // row (plume.py:106).  All of this warp's earlier stores are complete and ordered before this point.
            double* row = p.j_ion + s * (long long)A;
            for (int i = 0; i < A; ++i) row[i] = kInvalidFill;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// K1d: warp per sample, any grid, any radii, reference operation order
// ---------------------------------------------------------------------------------------------
constexpr int kThreadsD = 128;
constexpr int kWarpsD = kThreadsD / 32;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <bool WANT_PLUME>
__global__ void __launch_bounds__(kThreadsD) eval_direct_kernel(const EvalParams p) {
    const int lane = threadIdx.x & 31;
    const long long s = (long long)blockIdx.x * kWarpsD + (threadIdx.x >> 5);
    if (s >= p.n) return;

    const double p_b = load_in(p, IN_P_b, s);
    if (p.v_cc && lane == 0) {
        p.v_cc[s] = cathode_vcc(p_b, load_in(p, IN_V_a, s), load_in(p, IN_T_e, s), load_in(p, IN_V_vac, s),
                                load_in(p, IN_Pstar, s), load_in(p, IN_P_T, s), p.torr);
    }
    if (!WANT_PLUME) return;

    const SampleConsts k = plume_sample_consts(p_b, load_in(p, IN_c0, s), load_in(p, IN_c1, s), load_in(p, IN_c2, s),
                                               load_in(p, IN_c3, s), load_in(p, IN_c4, s), load_in(p, IN_c5, s), p.torr);
    const double sigma = load_in(p, IN_sigma, s), i_b0 = load_in(p, IN_I_B0, s);
    const int A = p.n_angles, R = p.n_radii;
    const bool known_invalid = (k.a1 <= 0.0);

    // pass 1: per-radius quadrature + detection of non-positive j_ion over all (angle, radius)
    bool bad = false;
    for (int rho = 0; rho < R; ++rho) {
        double j_cex, base;
        cex_terms(k.density, sigma, i_b0, __ldg(p.radii + rho), j_cex, base);
        const double amp1 = __dmul_rn(base, k.amp1), amp2 = __dmul_rn(base, k.amp2);
        double num = 0.0, den = 0.0;
        for (int i = lane; i < A; i += 32) {
            const double al = __ldg(p.alpha + i);
            const double u1 = al / k.a1, u2 = al / k.a2;
            const double jb = __dmul_rn(amp1, exp(-(u1 * u1)));   // plume.py:99
            const double js = __dmul_rn(amp2, exp(-(u2 * u2)));   // plume.py:100
            const double sum = __dadd_rn(jb, js);
            const double j = __dadd_rn(sum, j_cex);               // plume.py:102
            bad |= (j <= 0.0);
            const double2 w = __ldg(p.w + i);
            den = fma(w.x, sum, den);
            num = fma(w.y, sum, num);
        }
        num = warp_sum(num);
        den = warp_sum(den);
        double cd = num / den;
        if (cd == CUDART_INF) cd = CUDART_NAN;
        if (lane == 0) {
            const long long o = s * R + rho;
            if (p.div_angle) p.div_angle[o] = acos(cd);
            if (p.cos_div) p.cos_div[o] = cd;
            if (p.t_c) p.t_c[o] = __dmul_rn(load_in(p, IN_T, s), cd);
        }
    }
    const bool invalid = known_invalid || __any_sync(0xffffffffu, bad);
    if (p.invalid && lane == 0) p.invalid[s] = invalid ? 1 : 0;

    // pass 2: j_ion (n, A, R), radius fastest
    if (p.j_ion) {
        double* row = p.j_ion + s * (long long)A * R;
        if (invalid) {
            for (int cidx = lane; cidx < A * R; cidx += 32) row[cidx] = kInvalidFill;
        } else if (R == 1) {
            double j_cex, base;
            cex_terms(k.density, sigma, i_b0, __ldg(p.radii), j_cex, base);
            const double amp1 = __dmul_rn(base, k.amp1), amp2 = __dmul_rn(base, k.amp2);
            for (int i = lane; i < A; i += 32) {
                const double al = __ldg(p.alpha + i);
                const double u1 = al / k.a1, u2 = al / k.a2;
                const double jb = __dmul_rn(amp1, exp(-(u1 * u1)));
                const double js = __dmul_rn(amp2, exp(-(u2 * u2)));
                __stcs(row + i, __dadd_rn(__dadd_rn(jb, js), j_cex));
            }
        } else {
            for (int cidx = lane; cidx < A * R; cidx += 32) {
                const int i = cidx / R, rho = cidx - i * R;
                double j_cex, base;
                cex_terms(k.density, sigma, i_b0, __ldg(p.radii + rho), j_cex, base);
                const double al = __ldg(p.alpha + i);
                const double u1 = al / k.a1, u2 = al / k.a2;
                const double jb = __dmul_rn(__dmul_rn(base, k.amp1), exp(-(u1 * u1)));
                const double js = __dmul_rn(__dmul_rn(base, k.amp2), exp(-(u2 * u2)));
                __stcs(row + cidx, __dadd_rn(__dadd_rn(jb, js), j_cex));
            }
        }
    }
}

}  // namespace hpem
