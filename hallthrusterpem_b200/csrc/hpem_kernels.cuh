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
#include <stdint.h>

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
// K1u: thread per sample, uniform grid, R == 1
// ---------------------------------------------------------------------------------------------
template <bool WANT_PLUME, bool STORE_J>
__global__ void __launch_bounds__(kThreadsU) eval_uniform_kernel(const EvalParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double2* wsm = reinterpret_cast<double2*>(smem_raw);
    double* tiles = reinterpret_cast<double*>(smem_raw + size_t(p.n_angles_pad) * sizeof(double2));

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    double* tile = tiles + warp * (32 * kTilePitch);

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
    const double amp1 = __dmul_rn(base, k.amp1);  // (base_density * A1), plume.py:99
    const double amp2 = __dmul_rn(base, k.amp2);  // (base_density * A2), plume.py:100

    // x_b = (h / a_b)^2 ; profile_b(i) = exp(-x_b i^2)
    const double t1 = p.h / k.a1, t2 = p.h / k.a2;
    const double x1 = t1 * t1, x2 = t2 * t2;
    // recurrence constants: r(i) = exp(-x(2i+1)), q = exp(-2x); chunk level: R(c+1) = R(c)*QK,
    // E(c+1) = E(c)*G(c), G(c+1) = G(c)*H
    const double q1 = exp(-2.0 * x1), q2 = exp(-2.0 * x2);
    const double qk1 = exp(-(2.0 * kChunk) * x1), qk2 = exp(-(2.0 * kChunk) * x2);
    const double hh1 = exp(-(2.0 * kChunk * kChunk) * x1), hh2 = exp(-(2.0 * kChunk * kChunk) * x2);
    double rc1 = exp(-x1), rc2 = exp(-x2);
    double gc1 = exp(-double(kChunk * kChunk) * x1), gc2 = exp(-double(kChunk * kChunk) * x2);
    double ec1 = 1.0, ec2 = 1.0;

    const bool known_invalid = (k.a1 <= 0.0);  // plume.py:105 first term
    bool bad = false;
    double num = 0.0, den = 0.0;

    const int A = p.n_angles;
    const int n_chunks = p.n_angles_pad / kChunk;
    const int rows_valid = (int)min((long long)32, p.n - warp_s0);
    const int col = lane & (kChunk - 1);
    const int rsub = lane >> 4;
    double* gp = STORE_J ? p.j_ion + (warp_s0 + rsub) * (long long)A + col : nullptr;
    double* my_tile_row = tile + lane * kTilePitch;

    for (int c = 0; c < n_chunks; ++c) {
        const int i0 = c * kChunk;
        if (c != 0 && (c % kRestartChunks) == 0) {  // exact restart bounds the recurrence error for large A
            const double di = double(i0);
            ec1 = exp(-x1 * (di * di));
            ec2 = exp(-x2 * (di * di));
            rc1 = exp(-x1 * (2.0 * di + 1.0));
            rc2 = exp(-x2 * (2.0 * di + 1.0));
            gc1 = exp(-x1 * (2.0 * kChunk * di + double(kChunk * kChunk)));
            gc2 = exp(-x2 * (2.0 * kChunk * di + double(kChunk * kChunk)));
        }
        double e1 = amp1 * ec1, e2 = amp2 * ec2;
        double r1 = rc1, r2 = rc2;
        const int kcount = min(kChunk, A - i0);
        if (kcount == kChunk) {
#pragma unroll
            for (int kk = 0; kk < kChunk; ++kk) {
                const double2 w = wsm[i0 + kk];
                const double sum = e1 + e2;        // j_beam + j_scat
                const double j = sum + j_cex;      // plume.py:102
                den = fma(w.x, sum, den);
                num = fma(w.y, sum, num);
                bad |= (j <= 0.0);
                if (STORE_J) my_tile_row[kk] = known_invalid ? kInvalidFill : j;
                e1 *= r1; r1 *= q1;
                e2 *= r2; r2 *= q2;
            }
        } else {
            for (int kk = 0; kk < kcount; ++kk) {
                const double2 w = wsm[i0 + kk];
                const double sum = e1 + e2;
                const double j = sum + j_cex;
                den = fma(w.x, sum, den);
                num = fma(w.y, sum, num);
                bad |= (j <= 0.0);
                if (STORE_J) my_tile_row[kk] = known_invalid ? kInvalidFill : j;
                e1 *= r1; r1 *= q1;
                e2 *= r2; r2 *= q2;
            }
        }
        ec1 *= gc1; gc1 *= hh1; rc1 *= qk1;
        ec2 *= gc2; gc2 *= hh2; rc2 *= qk2;

        if (STORE_J) {
            __syncwarp();
            const double* trow = tile + rsub * kTilePitch + col;
            double* g = gp + i0;
            const bool col_ok = col < kcount;
#pragma unroll
            for (int rr = 0; rr < 16; ++rr) {
                if (col_ok && (2 * rr + rsub) < rows_valid) __stcs(g, trow[2 * rr * kTilePitch]);
                g += 2 * (long long)A;
            }
            __syncwarp();
        }
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
            // row (plume.py:106).  The __syncwarp() after the last store-out orders those stores first.
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
