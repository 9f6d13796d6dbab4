// hpem_moments.cuh -- K2: reduce-only Monte-Carlo pass (no j_ion materialisation): sample moments and histograms.
//
// What the consumers of the reference's outputs compute over the sample axis (np.percentile / means of j_ion per angle,
// tests/test_plume.py:50-52, scripts/gen_data.py:402-404) for sample counts whose j_ion cannot be stored (BASELINE
// configs 4-5: 1e8-1e9 samples x up to 512 angles).
//
// Mapping.  One persistent block per SM; every THREAD owns NS = 2 or 3 samples (s, s + 32, s + 64) and runs their recurrence
// sweeps (same arithmetic as K1u) interleaved: four independent multiply chains per sample, and the per-sample prologue
// of all its samples is one branch-free block (hpem_fastmath.cuh).  The Philox words of a batch are drawn one iteration
// ahead, beside the latency-bound tail of the previous batch.  The two Simpson sums of plume.py:121-122 come from the
// grid's table (hpem_qtable.cuh): one lookup per beam instead of two fused multiply-adds per (sample, angle).  The
// per-angle sums over samples need a transposition (thread = sample for the sweep, thread = angle for the column sums).
// The samples of a thread are combined in registers first,
//        t = sum_u j_u,   q = sum_u j_u^2,
// so the tile that goes through shared memory holds one (t, q) pair per NS evaluations: 16-byte conflict-free
// stores, half (a third) of the shared-memory traffic of a value-per-evaluation tile (the first version was as busy on
// the shared-memory pipe as on the fp64 pipe).  Column sums: 2 lanes per column of 16 rows, added into per-warp
// accumulators; one partial vector per block at the end, merged in block order by moments_finalize_kernel
// (bit-reproducible for a fixed launch geometry).
//
// Histograms: log-linear bins straight from the leading bits of the fp64 pattern (no log), one fire-and-forget reduction
// per lane to the block's private histogram in global memory (L2 atomics; blocks never share a line) -- see hist_add.
//
// Second moments are kept CENTRED: a block's raw sums become (n, S, M2 = Q - S^2/n) and are merged with Chan's
// pairwise update, block after block, call after call, rank after rank -- the variance of the whole population never
// comes from E[x^2] - E[x]^2.  The per-sample scalars are accumulated about a caller-supplied shift.
#pragma once
#include "hpem_kernels.cuh"

namespace hpem {

struct MomentsParams {
    int hist_stride;      // power of two, 0 = no histograms
    int hist_shift;       // log2(hist_stride)
    int want_cathode;     // accumulate V_cc moments (the six cathode inputs are read)
    int hist_sub_bits;
    int hist_min_exp2, hist_max_exp2;
    int n_hist_angles, n_bins;
    long long n_sums;     // doubles in the packed vector
    long long off_angle_sum, off_angle_sumsq, off_hist;
    double shift[3];      // V_cc, div_angle, T_c are accumulated as (x - shift)
    double* partials;     // [gridDim.x][n_part]   n_part = kMomScalars + 2 A   (raw sums of one block)
    double* partial_minmax;  // [gridDim.x][6]  (-min, max) x (V_cc, div_angle, T_c)
    unsigned* hist_partials; // [gridDim.x][n_hist_angles][n_bins]   zero on entry, re-zeroed by the finalize kernel
};
constexpr int kMomScalars = 12;  // n_samples n_invalid n_nonfinite_rows | {n_finite sum M2} x {V_cc div_angle T_c}
#ifndef HPEM_THREADS_M
#define HPEM_THREADS_M 384
#endif
constexpr int kThreadsM = HPEM_THREADS_M;   // upper bound; the launch picks the warp count that fits shared memory
constexpr int kMaxWarpsM = kThreadsM / 32;
constexpr int kHalfPitch = 9;               // double2 per half-tile row (8 angles + 1): odd -> conflict-free 16-byte stores and column loads
constexpr int kTileElems = 2 * 32 * kHalfPitch;   // double2 per warp: two half-tiles of 32 rows

// shared memory of one block with `warps` warps: fused weights + per-warp (t, q) tile + per-warp per-angle accumulators
__host__ __device__ inline size_t moments_smem_bytes(int n_angles_pad, int a_pad, int warps) {
    return size_t(n_angles_pad) * sizeof(double2) + size_t(warps) * kTileElems * sizeof(double2) +
           size_t(warps) * a_pad * sizeof(double2);
}

// recurrence state of one beam of one sample (same arithmetic as K1u's BeamState, so K2's j_ion is K1u's bit for bit)
struct SweepBeam {
    double ec, rc, gc, q, qk, hh;
};

template <bool FAST>
__device__ __forceinline__ void sweep_beam_init(SweepBeam& b, double& x, double h, double a) {
    const double t = m_div<FAST>(h, a);
    x = t * t;
    b.rc = m_exp<FAST>(-x);
    b.q = b.rc * b.rc;
    b.qk = m_exp<FAST>(-(2.0 * kChunk) * x);
    b.gc = m_exp<FAST>(-double(kChunk * kChunk) * x);
    b.hh = b.gc * b.gc;
    b.ec = 1.0;
}
template <bool FAST = false>
__device__ __forceinline__ void sweep_beam_restart(SweepBeam& b, double x, int i0) {
    const double di = double(i0);
    b.ec = m_exp<FAST>(-x * (di * di));
    b.rc = m_exp<FAST>(-x * (2.0 * di + 1.0));
    b.gc = m_exp<FAST>(-x * (2.0 * kChunk * di + double(kChunk * kChunk)));
}
__device__ __forceinline__ void sweep_beam_next(SweepBeam& b) {
    b.ec *= b.gc;
    b.gc *= b.hh;
    b.rc *= b.qk;
}

// One histogram update: log-linear bin of j (octave from the exponent field, 2^sub_bits linear sub-bins from the leading
// mantissa bits; bin 0 = underflow incl. zero / negative -- the shifted pattern is negative --, last bin = overflow incl.
// +inf) and ONE reduction per lane to the block's private histogram row in global memory.  The 32 lanes of a warp look at
// the same angle and a population occupies few bins there; the L2 atomic units absorb that (fire-and-forget RED, no
// return value, blocks never share a line).  Measured against the alternatives on B200, 256 angles, histogram every 8th:
// per-lane shared-memory atomics serialise on the hot bins (44 % of the first version of this kernel), per-warp slot
// buffers cost 4 instructions per angle step (round 1), __match_any_sync aggregation + one RED per group waits ~200
// cycles for MATCH even when its result is consumed five steps later (7.35e11 evals/s); plain per-lane REDs: 7.77e11.
__device__ __forceinline__ void hist_add(unsigned* hrow, double j, bool ok, int h_shift, int h_lo_key, int h_last) {
    const int b = min(max((__double2hiint(j) >> h_shift) - h_lo_key, 0), h_last);
    if (ok) atomicAdd(hrow + b, 1u);
}

// HS: histogram angle stride known at compile time (8, the default), 0 = no histograms, -1 = any power-of-two stride.
// RESTART: the row is longer than kRestartChunks chunks, the recurrences are re-anchored with exact exps (A > 256).
// Launch geometry.  Registers come in blocks of four warps: 12 warps x 168 registers or 8 warps x 255.
//
// NS = samples per thread.  Three samples on 8 warps x 255 registers keep as many samples in flight as two on 12 warps, with
// a third fewer shared-memory stores, column-reduce additions and loop instructions per evaluation: ahead from ~200 angles
// (4e7 samples, histograms: 224 angles +2 %, 256 +3 %, 512 +13 % over two samples on 12 warps), behind below (128: -2.5 %, 91:
// -4 %: the per-sample part dominates there and wants the twelve warps).  Four samples spill (972 vs 1036e9 at 256 angles).
constexpr int kWarpsLongM = 8;
constexpr int kLongChunksM = 13;     // rows of >= 13 chunks (> 192 angles) take the three-sample geometry (when instantiated)
template <bool SAMPLED, int HS, bool RESTART, int NS>
__global__ void __launch_bounds__((RESTART || NS > 2) ? kWarpsLongM * 32 : kThreadsM, 1) moments_kernel(const EvalParams p, const MomentsParams m,
                                                               const __grid_constant__ SamplerParams sp) {
    extern __shared__ __align__(16) unsigned char smem_m[];
    constexpr int kNS = NS;
    const int A = p.n_angles;
    const int n_chunks = (A + kChunk - 1) / kChunk;
    const int a_pad = n_chunks * kChunk;
    const int n_warps = blockDim.x >> 5;
    double2* wsm = reinterpret_cast<double2*>(smem_m);                      // [n_angles_pad]
    double2* tiles = wsm + p.n_angles_pad;                                  // [warps][2][32][kHalfPitch]
    double2* acc_all = tiles + n_warps * kTileElems;                        // [warps][a_pad]
    __shared__ double red[kMaxWarpsM][kMomScalars + 6];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double2* tile = tiles + warp * kTileElems;
    double2* acc = acc_all + warp * a_pad;
    for (int i = threadIdx.x; i < p.n_angles_pad; i += blockDim.x) wsm[i] = p.w[i];
    for (int i = threadIdx.x; i < n_warps * a_pad; i += blockDim.x) acc_all[i] = make_double2(0.0, 0.0);
    __syncthreads();

    // per-thread accumulators of the per-sample scalars: counts as integers, sums about the caller's shift
    int c_samples = 0, c_invalid = 0, c_nonfinite = 0, c_v = 0, c_d = 0, c_t = 0;
    double s_v = 0.0, q_v = 0.0, s_d = 0.0, q_d = 0.0, s_t = 0.0, q_t = 0.0;
    double mm[6] = {-CUDART_INF, -CUDART_INF, -CUDART_INF, -CUDART_INF, -CUDART_INF, -CUDART_INF};
    const bool want_cathode = m.want_cathode != 0;
    const bool want_thrust = p.has_thrust;
    const int h_shift = 20 - m.hist_sub_bits, h_last = m.n_bins - 1;
    const int h_lo_key = ((m.hist_min_exp2 + 1023) << m.hist_sub_bits) - 1;
    const int h_mask = max(m.hist_stride, 1) - 1;
    unsigned* hist_blk = m.hist_partials + (size_t)blockIdx.x * m.n_hist_angles * m.n_bins;

    const long long batch = (long long)n_warps * (32 * kNS);
    // SAMPLED: the Philox words of a batch are drawn one iteration ahead, next to the latency-bound tail of the previous
    // batch (last column reduce, division, arccos), where the integer pipe is idle
    uint32_t words[kNS][5][4];
    auto draw_words = [&](long long w0_next) {
        unsigned long long nidx[kNS];
#pragma unroll
        for (int u = 0; u < kNS; ++u) {
            const long long s_raw = w0_next + u * 32 + lane;
            nidx[u] = (unsigned long long)(s_raw < p.n ? s_raw : p.n - 1);
        }
        sample_words_n<kNS>(sp, nidx, words);
    };
    if (SAMPLED) draw_words((long long)blockIdx.x * batch + warp * (32 * kNS));
    for (long long b0 = (long long)blockIdx.x * batch; b0 < p.n; b0 += (long long)gridDim.x * batch) {
        const long long w0 = b0 + warp * (32 * kNS);
        if (w0 >= p.n) continue;   // warp-uniform; no block-level barrier inside the loop (later batches of this warp are out of range too)

        // ---- the two samples of this thread: inputs ----
        double x_in[kNS][kNumInputs];
        bool active[kNS];
        unsigned long long sidx[kNS];
#pragma unroll
        for (int u = 0; u < kNS; ++u) {
            const long long s_raw = w0 + u * 32 + lane;
            active[u] = s_raw < p.n;
            sidx[u] = (unsigned long long)(active[u] ? s_raw : p.n - 1);   // inactive lanes shadow the last sample, contribute nothing
        }
        if (SAMPLED) {
            sample_transform_n<kNS>(sp, sidx, words, x_in);
        } else {
#pragma unroll
            for (int u = 0; u < kNS; ++u) {
#pragma unroll
                for (int q = 0; q < kNumInputs; ++q) {
                    const bool needed = (q == IN_P_b) || (q <= IN_P_T ? want_cathode : (q == IN_T ? want_thrust : true));
                    x_in[u][q] = needed ? load_in(p, q, (long long)sidx[u]) : 0.0;
                }
            }
        }
        // ---- per-sample prologue, both samples in one basic block ----
        SweepBeam b1[kNS], b2[kNS];
        double bx1[kNS], bx2[kNS], bamp1[kNS], bamp2[kNS];     // recurrence exponents / amplitudes (restarts, row checks)
        double v_cc[kNS], j_cex[kNS], a1v[kNS];
        double num[kNS], den[kNS];
#pragma unroll
        for (int u = 0; u < kNS; ++u) v_cc[u] = num[u] = den[u] = 0.0;   // the Simpson sums of plume.py:121-122
        const bool use_qt = p.qt.rows != nullptr;
        auto prologue = [&](auto fast_tag) {
            constexpr bool FAST = decltype(fast_tag)::value;
#pragma unroll
            for (int u = 0; u < kNS; ++u) {
                if (want_cathode)
                    v_cc[u] = cathode_vcc<FAST>(x_in[u][IN_P_b], x_in[u][IN_V_a], x_in[u][IN_T_e], x_in[u][IN_V_vac],
                                                x_in[u][IN_Pstar], x_in[u][IN_P_T], p.torr);
            }
#pragma unroll
            for (int u = 0; u < kNS; ++u) {
                const SampleConsts k = plume_sample_consts<FAST>(x_in[u][IN_P_b], x_in[u][IN_c0], x_in[u][IN_c1], x_in[u][IN_c2],
                                                                 x_in[u][IN_c3], x_in[u][IN_c4], x_in[u][IN_c5], p.torr);
                double base;
                cex_terms<FAST>(k.density, x_in[u][IN_sigma], x_in[u][IN_I_B0], p.radius0, j_cex[u], base);
                a1v[u] = k.a1;
                bamp1[u] = __dmul_rn(base, k.amp1);
                bamp2[u] = __dmul_rn(base, k.amp2);
                sweep_beam_init<FAST>(b1[u], bx1[u], p.h, k.a1);
                sweep_beam_init<FAST>(b2[u], bx2[u], p.h, k.a2);
                if (FAST && use_qt) {   // both sums from the grid's table: amplitude x N(x) per beam (hpem_qtable.cuh)
                    double nd1, nn1, nd2, nn2;
                    qtable_eval(p.qt, bx1[u], b1[u].rc, nd1, nn1);
                    qtable_eval(p.qt, bx2[u], b2[u].rc, nd2, nn2);
                    den[u] = fma(bamp1[u], nd1, bamp2[u] * nd2);
                    num[u] = fma(bamp1[u], nn1, bamp2[u] * nn2);
                }
            }
        };
        bool nominal = true;
#pragma unroll
        for (int u = 0; u < kNS; ++u) nominal = nominal && prologue_nominal(x_in[u], p.torr, want_cathode, true, p.radius0);
        const bool fast = __all_sync(0xffffffffu, nominal) && !p.no_fastmath;
        if (fast)
            prologue(std::true_type{});
        else
            prologue(std::false_type{});

        bool invalid[kNS], row_ok[kNS];
        double thrust[kNS], j_fill[kNS];
#pragma unroll
        for (int u = 0; u < kNS; ++u) {
            thrust[u] = x_in[u][IN_T];
            if (want_cathode && active[u] && v_cc[u] == v_cc[u]) {
                const double d = v_cc[u] - m.shift[0];
                c_v += 1; s_v += d; q_v = fma(d, d, q_v);
                mm[0] = fmax(mm[0], -v_cc[u]); mm[1] = fmax(mm[1], v_cc[u]);
            }
            const bool known_invalid = (a1v[u] <= 0.0);   // plume.py:105 first term
            // a row is non-finite iff one of its per-sample constants is (then every angle is NaN/inf): it is counted and
            // contributes zeros to the per-angle sums; finite rows never produce a non-finite j_ion
            const double probe = (bamp1[u] + bamp2[u] + j_cex[u]) * 0.0 + (bx1[u] + bx2[u]) * 0.0;
            row_ok[u] = active[u] && (known_invalid ||   // alpha1 <= 0: the row is the finite 1e-20 fill whatever else is NaN
                                      ((probe == 0.0) && !(bx1[u] == CUDART_INF) && !(bx2[u] == CUDART_INF)));
            // plume.py:105-106: a sample with alpha1 <= 0 or any j_ion <= 0 has its whole row replaced by 1e-20.  Whether a
            // non-positive j_ion exists must be known BEFORE the row is accumulated, so the (rare) samples that can have
            // one -- negative amplitude or no CEX floor -- run a look-ahead sweep first.
            invalid[u] = known_invalid;
            if (!known_invalid && !(bamp1[u] >= 0.0 && bamp2[u] >= 0.0 && j_cex[u] > 0.0)) {
                SweepBeam t1 = b1[u], t2 = b2[u];
                bool any_bad = false;
                for (int c = 0; c < n_chunks; ++c) {
                    const int i0 = c * kChunk;
                    if (c != 0 && (c % kRestartChunks) == 0) {
                        sweep_beam_restart(t1, bx1[u], i0);
                        sweep_beam_restart(t2, bx2[u], i0);
                    }
                    double e1 = bamp1[u] * t1.ec, e2 = bamp2[u] * t2.ec, r1 = t1.rc, r2 = t2.rc;
                    for (int kk = 0; kk < kChunk && i0 + kk < A; ++kk) {
                        any_bad |= ((e1 + e2) + j_cex[u] <= 0.0);
                        e1 *= r1; r1 *= t1.q;
                        e2 *= r2; r2 *= t2.q;
                    }
                    sweep_beam_next(t1);
                    sweep_beam_next(t2);
                }
                invalid[u] = any_bad;
            }
            if (!row_ok[u]) {   // non-finite (or inactive shadow) row: exact zeros to the per-angle sums, NaN to cos_div
                b1[u].ec = b2[u].ec = 1.0;
                b1[u].rc = b1[u].gc = b1[u].q = b1[u].qk = b1[u].hh = 1.0;
                b2[u].rc = b2[u].gc = b2[u].q = b2[u].qk = b2[u].hh = 1.0;
                bx1[u] = bx2[u] = 0.0;
                bamp1[u] = bamp2[u] = 0.0;
                j_cex[u] = 0.0;
            }
            j_fill[u] = row_ok[u] ? kInvalidFill : 0.0;
        }
        // warps whose 64 rows are all ordinary (finite, valid) -- virtually all of them -- skip the per-element selects
        bool odd_row = false;
#pragma unroll
        for (int u = 0; u < kNS; ++u) odd_row = odd_row || invalid[u] || !row_ok[u];
        const bool plain = !__any_sync(0xffffffffu, odd_row);
        // The 16-angle chunk goes through shared memory as two half-tiles of 8 angles, [32 rows][kHalfPitch] (t, q) pairs each.
        // While a thread sweeps one half it column-reduces the OTHER one, two rows per recurrence step, so the loads and
        // additions of the reduction are scheduled inside the fp64 stream of the sweep instead of forming a latency-bound
        // phase of their own between two warp barriers.  Seen as doubles a half-tile has 16 columns (8 angles x {t, q}):
        // lane (c16 = lane & 15, h = lane >> 4) sums column c16 over rows 16 h .. 16 h + 15 with 8-byte loads (a half-warp
        // reads 128 contiguous bytes), ONE shuffle joins the two row halves, and lanes 0-15 add to the per-warp accumulators.
        // (A first version used 16-byte loads with lane = (angle, row quarter): two shuffle stages and a 8-lane update per
        //  half cost 26 % of the loop, tools/sweep_probe2.cu.)
        double2* my0 = tile + lane * kHalfPitch;                        // this thread's row in half-tile 0 (angles 0-7 of a chunk)
        double2* my1 = my0 + 32 * kHalfPitch;                           // ... in half-tile 1 (angles 8-15)
        const int c16 = lane & 15, rh = lane >> 4;
        const double* col0 = reinterpret_cast<const double*>(tile) + (rh * 16) * (2 * kHalfPitch) + c16;
        const double* col1 = col0 + 32 * (2 * kHalfPitch);
        double* acc_d = reinterpret_cast<double*>(acc);                 // (sum, sum of squares) interleaved per angle, like (t, q)

        // PLAIN warps (all rows ordinary) of the fast back end already hold the two Simpson sums (table); every other warp
        // accumulates them angle by angle and selects the fill value per element
        const bool tabulated = plain && fast && use_qt;
        if (!tabulated) {
#pragma unroll
            for (int u = 0; u < kNS; ++u) den[u] = num[u] = 0.0;
        }
        auto sweep = [&](auto plain_tag) {
            constexpr bool PLAIN = decltype(plain_tag)::value;
            unsigned* hrow = hist_blk;                       // histogram row of the next histogrammed angle
            double ra = 0.0, rb = 0.0;                        // running column sum of the half-tile being reduced (two chains)
            auto finish_half = [&](int angle_base, bool keep) {
                double s1 = ra + rb;
                s1 += __shfl_xor_sync(0xffffffffu, s1, 16);
                if (keep && rh == 0) acc_d[2 * angle_base + c16] += s1;
                ra = rb = 0.0;
            };
            for (int c = 0; c < n_chunks; ++c) {
                const int i0 = c * kChunk;
                if (RESTART && c != 0 && (c % kRestartChunks) == 0) {
                    if (fast) {      // all 64 samples nominal: twelve branch-free exps as one block (neutralised rows have x = 0: exp(0) = 1)
#pragma unroll
                        for (int u = 0; u < kNS; ++u) {
                            sweep_beam_restart<true>(b1[u], bx1[u], i0);
                            sweep_beam_restart<true>(b2[u], bx2[u], i0);
                        }
                    } else {
#pragma unroll
                        for (int u = 0; u < kNS; ++u) {
                            if (row_ok[u]) {
                                sweep_beam_restart(b1[u], bx1[u], i0);
                                sweep_beam_restart(b2[u], bx2[u], i0);
                            }
                        }
                    }
                }
                double e1[kNS], e2[kNS], r1[kNS], r2[kNS];
#pragma unroll
                for (int u = 0; u < kNS; ++u) {
                    e1[u] = bamp1[u] * b1[u].ec;
                    e2[u] = bamp2[u] * b2[u].ec;
                    r1[u] = b1[u].rc;
                    r2[u] = b2[u].rc;
                }
#pragma unroll
                for (int kk = 0; kk < kChunk; ++kk) {
                    double jv[kNS];
#pragma unroll
                    for (int u = 0; u < kNS; ++u) {
                        const double su = e1[u] + e2[u];     // j_beam + j_scat
                        if (!PLAIN) {
                            const double2 w = wsm[i0 + kk];  // zero beyond A
                            den[u] = fma(w.x, su, den[u]);
                            num[u] = fma(w.y, su, num[u]);
                        }
                        jv[u] = su + j_cex[u];               // the values current_density() returns (plume.py:102)
                        if (!PLAIN) jv[u] = invalid[u] ? j_fill[u] : jv[u];
                    }
                    double tsum = jv[0], qsum = jv[0] * jv[0];
#pragma unroll
                    for (int u = 1; u < kNS; ++u) {
                        tsum += jv[u];
                        qsum = fma(jv[u], jv[u], qsum);
                    }
                    (kk < 8 ? my0 : my1)[kk & 7] = make_double2(tsum, qsum);   // columns >= A are never read back
                    {   // two rows of the other half-tile (chunk c-1's angles 8-15 during steps 0-7, this chunk's 0-7 during 8-15)
                        const double* cc = (kk < 8 ? col1 : col0) + (2 * (kk & 7)) * (2 * kHalfPitch);
                        ra += cc[0];
                        rb += cc[2 * kHalfPitch];
                    }
                    if (HS != 0 && (HS > 0 ? (kk % (HS > 0 ? HS : 1) == 0) : (((i0 + kk) & h_mask) == 0)) && i0 + kk < A) {
#pragma unroll
                        for (int u = 0; u < kNS; ++u) hist_add(hrow, jv[u], PLAIN || row_ok[u], h_shift, h_lo_key, h_last);
                        hrow += m.n_bins;
                    }
#pragma unroll
                    for (int u = 0; u < kNS; ++u) {
                        e1[u] *= r1[u]; r1[u] *= b1[u].q;
                        e2[u] *= r2[u]; r2[u] *= b2[u].q;
                    }
                    if (kk == 7) {          // half-tile 1 of the previous chunk is reduced (garbage at c == 0: dropped); half-tile 0 is complete
                        finish_half(i0 - 8, c > 0);
                        __syncwarp();
                    }
                    if (kk == 15) {         // half-tile 0 of this chunk is reduced; half-tile 1 is complete
                        finish_half(i0, true);
                        __syncwarp();
                    }
                }
#pragma unroll
                for (int u = 0; u < kNS; ++u) {
                    sweep_beam_next(b1[u]);
                    sweep_beam_next(b2[u]);
                }
            }
            // the last chunk's upper half has no sweep to hide behind
#pragma unroll
            for (int rr = 0; rr < 16; rr += 2) {
                ra += col1[rr * (2 * kHalfPitch)];
                rb += col1[(rr + 1) * (2 * kHalfPitch)];
            }
            finish_half((n_chunks - 1) * kChunk + 8, true);
        };
        if (tabulated)
            sweep(std::true_type{});
        else
            sweep(std::false_type{});

        // ---- per-sample epilogue: plume.py:124-127,137 (NOT masked by `invalid`) ----
        // The branch-free division and arccos are evaluated unconditionally, in one basic block with the draw of the next
        // batch's Philox words; the libdevice forms replace them in the rare warps that need IEEE special cases.
        if (SAMPLED) draw_words(w0 + (long long)gridDim.x * batch);
        double cd[kNS], dv[kNS];
#pragma unroll
        for (int u = 0; u < kNS; ++u) {
            cd[u] = fm_div(num[u], den[u]);
            dv[u] = fm_acos(cd[u]);
        }
        bool sums_mid = true;
#pragma unroll
        for (int u = 0; u < kNS; ++u) sums_mid = sums_mid && fm_mid(den[u]) && fm_mid0(num[u]);
        if (!(fast && __all_sync(0xffffffffu, sums_mid))) {
#pragma unroll
            for (int u = 0; u < kNS; ++u) {
                cd[u] = num[u] / den[u];
                if (cd[u] == CUDART_INF) cd[u] = CUDART_NAN;  // plume.py:125
                dv[u] = acos(cd[u]);
            }
        }
#pragma unroll
        for (int u = 0; u < kNS; ++u) {
            if (active[u]) {
                c_samples += 1;
                if (invalid[u]) c_invalid += 1;
                if (!row_ok[u]) c_nonfinite += 1;
                if (dv[u] == dv[u]) {
                    const double d = dv[u] - m.shift[1];
                    c_d += 1; s_d += d; q_d = fma(d, d, q_d);
                    mm[2] = fmax(mm[2], -dv[u]); mm[3] = fmax(mm[3], dv[u]);
                }
                if (want_thrust) {
                    const double tc = __dmul_rn(thrust[u], cd[u]);
                    if (tc == tc) {
                        const double d = tc - m.shift[2];
                        c_t += 1; s_t += d; q_t = fma(d, d, q_t);
                        mm[4] = fmax(mm[4], -tc); mm[5] = fmax(mm[5], tc);
                    }
                }
            }
        }
    }
    // ---- block reduction of the register accumulators, then one partial vector per block ----
    // partial scalars: [0..2] counts, then per scalar {n, sum of (x - shift), sum of (x - shift)^2}
    const double sc[kMomScalars] = {double(c_samples), double(c_invalid), double(c_nonfinite), double(c_v), s_v, q_v,
                                    double(c_d), s_d, q_d, double(c_t), s_t, q_t};
#pragma unroll
    for (int i = 0; i < kMomScalars; ++i) {
        double v = sc[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) red[warp][i] = v;
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        double v = mm[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
        if (lane == 0) red[warp][kMomScalars + i] = v;
    }
    __syncthreads();
    const long long n_part = kMomScalars + 2LL * A;
    double* out = m.partials + (long long)blockIdx.x * n_part;
    if (threadIdx.x < kMomScalars) {
        double v = 0.0;
        for (int w = 0; w < n_warps; ++w) v += red[w][threadIdx.x];
        out[threadIdx.x] = v;
    } else if (threadIdx.x < kMomScalars + 6) {
        double v = -CUDART_INF;
        for (int w = 0; w < n_warps; ++w) v = fmax(v, red[w][threadIdx.x]);
        m.partial_minmax[(long long)blockIdx.x * 6 + (threadIdx.x - kMomScalars)] = v;
    }
    for (int i = threadIdx.x; i < A; i += blockDim.x) {
        double s1 = 0.0, s2 = 0.0;
        for (int w = 0; w < n_warps; ++w) {
            const double2 a = acc_all[w * a_pad + i];
            s1 += a.x;
            s2 += a.y;
        }
        out[kMomScalars + i] = s1;          // sum over the block's samples of j_ion[:, i]
        out[kMomScalars + A + i] = s2;      // ... of j_ion[:, i]^2
    }
}

// Chan / pairwise update of (n, S, M2) with another group (nb, Sb, M2b): M2 = sum of squared deviations from the mean
__device__ __forceinline__ void chan_merge(double& n, double& S, double& M2, double nb, double Sb, double M2b) {
    if (!(nb > 0.0)) return;
    if (!(n > 0.0)) {
        n = nb; S = Sb; M2 = M2b;
        return;
    }
    const double delta = Sb / nb - S / n;
    M2 += M2b + delta * delta * (n * nb / (n + nb));
    S += Sb;
    n += nb;
}

// Merge the per-block partial vectors (raw sums about the shifts) of ONE accumulate call into the caller's packed vector,
// blocks in index order.  One thread per (scalar group | angle | histogram bin); the three counters [0..2] are read here
// (the per-angle row count is sums[0] - sums[2]) and updated by moments_counts_kernel afterwards.
__global__ void moments_finalize_kernel(const double* __restrict__ partials, const double* __restrict__ partial_minmax,
                                        unsigned* __restrict__ hist_partials, int n_blocks, int n_angles, MomentsParams m,
                                        double* __restrict__ sums, double* __restrict__ minmax) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long n_part = kMomScalars + 2LL * n_angles;
    if (i < 3) {
        const int g = (int)i;   // V_cc, div_angle, T_c
        double n = sums[3 + 3 * g], S = sums[4 + 3 * g], M2 = sums[5 + 3 * g];
        for (int b = 0; b < n_blocks; ++b) {
            const double* pb = partials + (long long)b * n_part + 3 + 3 * g;
            const double nb = pb[0], Sd = pb[1], Qd = pb[2];
            if (nb > 0.0) chan_merge(n, S, M2, nb, Sd + nb * m.shift[g], fmax(Qd - Sd * Sd / nb, 0.0));
        }
        sums[3 + 3 * g] = n; sums[4 + 3 * g] = S; sums[5 + 3 * g] = M2;
    } else if (i >= m.off_angle_sum && i < m.off_angle_sum + n_angles) {
        const int a = (int)(i - m.off_angle_sum);
        double n = sums[0] - sums[2], S = sums[m.off_angle_sum + a], M2 = sums[m.off_angle_sumsq + a];
        for (int b = 0; b < n_blocks; ++b) {
            const double* pb = partials + (long long)b * n_part;
            const double nb = pb[0] - pb[2], Sb = pb[kMomScalars + a], Qb = pb[kMomScalars + n_angles + a];
            if (nb > 0.0) chan_merge(n, S, M2, nb, Sb, fmax(Qb - Sb * Sb / nb, 0.0));
        }
        sums[m.off_angle_sum + a] = S;
        sums[m.off_angle_sumsq + a] = M2;
    } else if (i >= m.off_hist && i < m.n_sums) {
        const long long k = i - m.off_hist, per = (long long)m.n_hist_angles * m.n_bins;
        unsigned long long v = 0;
        for (int b = 0; b < n_blocks; ++b) {
            v += hist_partials[(long long)b * per + k];
            hist_partials[(long long)b * per + k] = 0u;
        }
        sums[i] += double(v);
    }
    if (i < 6 && minmax) {
        double v = minmax[i];
        for (int b = 0; b < n_blocks; ++b) v = fmax(v, partial_minmax[(long long)b * 6 + i]);
        minmax[i] = v;
    }
}
__global__ void moments_counts_kernel(const double* __restrict__ partials, int n_blocks, int n_angles, double* __restrict__ sums) {
    const int i = threadIdx.x;
    if (i < 3) {
        const long long n_part = kMomScalars + 2LL * n_angles;
        double v = 0.0;
        for (int b = 0; b < n_blocks; ++b) v += partials[(long long)b * n_part + i];
        sums[i] += v;
    }
}

// Merge `n_parts` packed vectors (each [sums (n_sums) | minmax (6)], `stride` doubles apart) in index order into
// out_sums / out_minmax (overwritten): what every rank runs after the all-gather, so all ranks hold the same bits
// whatever order the collective moved the data in.
__global__ void moments_merge_kernel(const double* __restrict__ parts, long long stride, int n_parts, int n_angles, MomentsParams m,
                                     double* __restrict__ out_sums, double* __restrict__ out_minmax) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 3) {   // counters
        double v = 0.0;
        for (int r = 0; r < n_parts; ++r) v += parts[r * stride + i];
        out_sums[i] = v;
        const int g = (int)i;
        double n = 0.0, S = 0.0, M2 = 0.0;
        for (int r = 0; r < n_parts; ++r) {
            const double* pr = parts + r * stride + 3 + 3 * g;
            chan_merge(n, S, M2, pr[0], pr[1], pr[2]);
        }
        out_sums[3 + 3 * g] = n; out_sums[4 + 3 * g] = S; out_sums[5 + 3 * g] = M2;
    } else if (i >= m.off_angle_sum && i < m.off_angle_sum + n_angles) {
        const int a = (int)(i - m.off_angle_sum);
        double n = 0.0, S = 0.0, M2 = 0.0;
        for (int r = 0; r < n_parts; ++r) {
            const double* pr = parts + r * stride;
            chan_merge(n, S, M2, pr[0] - pr[2], pr[m.off_angle_sum + a], pr[m.off_angle_sumsq + a]);
        }
        out_sums[m.off_angle_sum + a] = S;
        out_sums[m.off_angle_sumsq + a] = M2;
    } else if (i >= m.off_hist && i < m.n_sums) {
        double v = 0.0;
        for (int r = 0; r < n_parts; ++r) v += parts[r * stride + i];   // integer-valued, exact
        out_sums[i] = v;
    }
    if (i < 6) {
        double v = -CUDART_INF;
        for (int r = 0; r < n_parts; ++r) v = fmax(v, parts[r * stride + m.n_sums + i]);
        out_minmax[i] = v;
    }
}

}  // namespace hpem
