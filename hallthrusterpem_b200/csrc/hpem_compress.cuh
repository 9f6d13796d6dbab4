// hpem_compress.cuh -- SVD compression of the j_ion field quantity.
//
// Downstream of the plume model the reference never trains on the (n, A) field itself: amisc normalises j_ion with
// log10 and projects every sample onto the leading left-singular vectors of a small "compression" sample set
// (/root/reference/scripts/pem_v0/pem_v0_SPT-100.yml:272-280 `norm: log10`, `compression: {method: svd,
// reconstruction_tol: 0.01}`; scripts/gen_data.py:279-290 builds the map with `var.normalize(...)` and
// `var.compression.compute_map(...)`).  amisc itself is un-vendored (uv.lock:14-16, archermarx/amisc v0.8.1); the
// arithmetic restated here is its published SVD compression: latent = U_r^T x, x = log10(j_ion); reconstruction
// j_ion = 10^(U_r z).
//
// K4  latent_kernel        : fused plume model -> log10 -> projection.  One thread per sample runs the recurrence
//                            sweep of K1u; j_ion is never materialised, only (n, rank) latent coefficients leave the SM.
// K4f compress_field_kernel: projection of an already materialised (n, dof) field (one warp per row).
// K5  reconstruct_kernel   : (n, rank) latent coefficients -> (n, dof) field, coalesced streaming stores.
#pragma once
#include "hpem_kernels.cuh"

namespace hpem {

constexpr int kMaxRank = 32;
constexpr int kThreadsC = 128;

struct BasisParams {
    int dof;               // rows of the projection matrix (= n_angles for the fused kernel)
    int rank;              // columns actually used
    int rank_pad;          // row pitch of `basis` (multiple of 4, zero-padded)
    int norm_log10;        // 1: x = log10(field) / field = 10^x ; 0: identity
    const double* basis;   // device, [dof][rank_pad] row-major (projection_matrix, zero-padded columns)
    const double* basis_t; // device, [rank][dof]  (transpose, for the reconstruction)
};

// plume.py:105 second term for the rare samples that can have a non-positive j_ion (negative amplitude or no CEX floor):
// sweep once without storing anything.
__device__ __forceinline__ bool lookahead_nonpositive(BeamState t1, BeamState t2, double j_cex, int A) {
    bool any_bad = false;
    const int n_chunks = (A + kChunk - 1) / kChunk;
    for (int c = 0; c < n_chunks; ++c) {
        const int i0 = c * kChunk;
        if (c != 0 && (c % kRestartChunks) == 0) {
            beam_restart(t1, i0);
            beam_restart(t2, i0);
        }
        double e1 = t1.amp * t1.ec, e2 = t2.amp * t2.ec, r1 = t1.rc, r2 = t2.rc;
        for (int kk = 0; kk < kChunk && i0 + kk < A; ++kk) {
            any_bad |= ((e1 + e2) + j_cex <= 0.0);
            e1 *= r1; r1 *= t1.q;
            e2 *= r2; r2 *= t2.q;
        }
        beam_next_chunk(t1);
        beam_next_chunk(t2);
    }
    return any_bad;
}

// K4: RK = padded rank held in registers (4, 8, 16 or 32)
template <int RK>
__global__ void __launch_bounds__(kThreadsC) latent_kernel(const EvalParams p, const BasisParams bp, double* __restrict__ latent) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    double* usm = reinterpret_cast<double*>(smem_raw);   // [A][RK]
    const int A = p.n_angles;
    for (int i = threadIdx.x; i < A * RK; i += kThreadsC) {
        const int a = i / RK, k = i - a * RK;
        usm[i] = (k < bp.rank_pad) ? bp.basis[a * bp.rank_pad + k] : 0.0;
    }
    __syncthreads();
    const long long s = (long long)blockIdx.x * kThreadsC + threadIdx.x;
    if (s >= p.n) return;

    double x_in[kNumInputs];
#pragma unroll
    for (int q = 0; q < kNumInputs; ++q) x_in[q] = (q == IN_P_b || (q > IN_P_T && q != IN_T)) ? load_in(p, q, s) : 0.0;
    SampleConsts k;
    double j_cex, base;
    BeamState b1, b2;
    const bool fast = plume_prologue(p, x_in, k, j_cex, base, b1, b2);

    // plume.py:105-106: invalid samples return 1e-20 at every angle -- that row is what gets normalised and projected
    bool invalid = (k.a1 <= 0.0);
    const bool ordinary = b1.amp >= 0.0 && b2.amp >= 0.0 && j_cex > 0.0;
    if (!invalid && !ordinary) invalid = lookahead_nonpositive(b1, b2, j_cex, A);
    // Warps whose rows are all ordinary (finite positive amplitudes, a normal CEX floor: every j_ion is a positive normal
    // number) take the branch-free logarithm: sixteen independent log10 per chunk schedule as one block of fp64 work
    // instead of sixteen libdevice calls with a special-case branch each.
    const bool fast_log = fast && bp.norm_log10 && __all_sync(__activemask(), !invalid && ordinary && fm_mid(j_cex) && fm_mid0(b1.amp) && fm_mid0(b2.amp));

    double z[RK];
#pragma unroll
    for (int r = 0; r < RK; ++r) z[r] = 0.0;
    const int n_chunks = (A + kChunk - 1) / kChunk;
    for (int c = 0; c < n_chunks; ++c) {
        const int i0 = c * kChunk;
        if (c != 0 && (c % kRestartChunks) == 0) {
            beam_restart(b1, i0);
            beam_restart(b2, i0);
        }
        double e1 = b1.amp * b1.ec, e2 = b2.amp * b2.ec, r1 = b1.rc, r2 = b2.rc;
        const int kcount = min(kChunk, A - i0);
        if (fast_log && kcount == kChunk) {
#pragma unroll
            for (int kk = 0; kk < kChunk; ++kk) {
                const double x = fm_log10((e1 + e2) + j_cex);
                const double2* u = reinterpret_cast<const double2*>(usm + (i0 + kk) * RK);   // broadcast loads
#pragma unroll
                for (int r = 0; r < RK; r += 2) {
                    const double2 uu = u[r >> 1];
                    z[r] = fma(uu.x, x, z[r]);
                    z[r + 1] = fma(uu.y, x, z[r + 1]);
                }
                e1 *= r1; r1 *= b1.q;
                e2 *= r2; r2 *= b2.q;
            }
        } else
        for (int kk = 0; kk < kcount; ++kk) {
            const double j = invalid ? kInvalidFill : (e1 + e2) + j_cex;
            const double x = bp.norm_log10 ? log10(j) : j;
            const double2* u = reinterpret_cast<const double2*>(usm + (i0 + kk) * RK);   // broadcast loads
#pragma unroll
            for (int r = 0; r < RK; r += 2) {
                const double2 uu = u[r >> 1];
                z[r] = fma(uu.x, x, z[r]);
                z[r + 1] = fma(uu.y, x, z[r + 1]);
            }
            e1 *= r1; r1 *= b1.q;
            e2 *= r2; r2 *= b2.q;
        }
        beam_next_chunk(b1);
        beam_next_chunk(b2);
    }
    double* out = latent + s * (long long)bp.rank;
#pragma unroll
    for (int r = 0; r < RK; ++r)
        if (r < bp.rank) out[r] = z[r];
}

// K4f: one warp per row of a materialised field.  Lane l takes elements l, l+32, ... (coalesced row reads, ONE log10 per
// element), keeps RK partial coefficients in registers against the transposed projection [rank][dof] (coalesced,
// L1-resident), and the warp reduces the RK partials by shuffles.
template <int RK>
__global__ void __launch_bounds__(kThreadsC) compress_field_kernel(const double* __restrict__ field, long long n,
                                                                   const BasisParams bp, double* __restrict__ latent) {
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * (kThreadsC / 32) + (threadIdx.x >> 5);
    if (row >= n) return;
    const double* f = field + row * (long long)bp.dof;
    double z[RK];
#pragma unroll
    for (int r = 0; r < RK; ++r) z[r] = 0.0;
    auto project = [&](int i, double x) {
        const double* u = bp.basis_t + i;
#pragma unroll
        for (int r = 0; r < RK; ++r)
            if (r < bp.rank) z[r] = fma(__ldg(u + (long long)r * bp.dof), x, z[r]);
    };
    // four elements per lane and trip (warp-uniform trip count): when all 128 of them are positive normal numbers (any j_ion
    // the plume model returns) the four logarithms are the branch-free ones and schedule as one block; anything else takes
    // libdevice's log10
    int i = lane;
    for (int i0 = 0; i0 + 128 <= bp.dof; i0 += 128, i += 128) {
        double v[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) v[t] = __ldcs(f + i + 32 * t);
        if (bp.norm_log10) {
            const bool ok = fm_mid(v[0]) && v[0] > 0.0 && fm_mid(v[1]) && v[1] > 0.0 && fm_mid(v[2]) && v[2] > 0.0 && fm_mid(v[3]) && v[3] > 0.0;
            if (__all_sync(0xffffffffu, ok)) {
#pragma unroll
                for (int t = 0; t < 4; ++t) v[t] = fm_log10(v[t]);
            } else {
#pragma unroll
                for (int t = 0; t < 4; ++t) v[t] = log10(v[t]);
            }
        }
#pragma unroll
        for (int t = 0; t < 4; ++t) project(i + 32 * t, v[t]);
    }
    for (; i < bp.dof; i += 32) {
        const double v = __ldcs(f + i);
        project(i, bp.norm_log10 ? log10(v) : v);
    }
#pragma unroll
    for (int r = 0; r < RK; ++r) z[r] = warp_sum(z[r]);
    if (lane == 0) {
        double* out = latent + row * (long long)bp.rank;
#pragma unroll
        for (int r = 0; r < RK; ++r)
            if (r < bp.rank) out[r] = z[r];
    }
}

// K5: a block takes kReconSamples consecutive samples; thread t owns field elements t, t+blockDim, ... with their
// projection rows in registers and walks the block's samples, whose latent rows sit in shared memory (broadcast reads).
// Consecutive threads write consecutive addresses of one row (coalesced streaming stores).
constexpr int kReconSamples = 64;
constexpr int kThreadsR = 128;
template <int RK>
__global__ void __launch_bounds__(kThreadsR) reconstruct_kernel(const double* __restrict__ latent, long long n, const BasisParams bp,
                                                                double* __restrict__ field) {
    __shared__ __align__(16) double zs[kReconSamples * RK];
    const long long s0 = (long long)blockIdx.x * kReconSamples;
    const int ns = (int)min((long long)kReconSamples, n - s0);
    for (int e = threadIdx.x; e < ns * RK; e += kThreadsR) {
        const int s = e / RK, r = e - s * RK;
        zs[e] = (r < bp.rank) ? __ldg(latent + (s0 + s) * bp.rank + r) : 0.0;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < bp.dof; i += kThreadsR) {
        double u[RK];
#pragma unroll
        for (int r = 0; r < RK; ++r) u[r] = (r < bp.rank) ? __ldg(bp.basis_t + (long long)r * bp.dof + i) : 0.0;
        double* out = field + s0 * bp.dof + i;
        for (int s = 0; s < ns; ++s) {
            const double2* zr = reinterpret_cast<const double2*>(zs + s * RK);
            double x = 0.0;
#pragma unroll
            for (int r = 0; r < RK; r += 2) {
                const double2 zz = zr[r >> 1];
                x = fma(u[r], zz.x, x);
                x = fma(u[r + 1], zz.y, x);
            }
            // (libdevice's exp10 stays: a branch-free 10^x through fm_exp measured 25 % slower here -- the kernel is bound by
            //  fp64 issue, not by latency, and exp10's table-free core is shorter than exp(x ln10) with a compensated product)
            __stcs(out + (long long)s * bp.dof, bp.norm_log10 ? exp10(x) : x);
        }
    }
}

}  // namespace hpem
