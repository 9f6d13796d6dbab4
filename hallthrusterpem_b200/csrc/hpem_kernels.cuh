// hpem_kernels.cuh -- sm_100a kernels for the fused cathode + plume sample batch.
//
// K1u  eval_uniform_kernel : one THREAD per sample, uniform angle grid, single radius (the default path).
//        The two Gaussian beam profiles exp(-(i h / a)^2), i = 0..A-1, are advanced with a two-term
//        multiplicative recurrence (e *= r; r *= q) restarted hierarchically, so the per-(sample, angle)
//        cost is 4 DMUL + 2 DADD + 2 DFMA (+1 DSETP) instead of two fp64 exp() calls (~40 fp64-pipe ops).
//        The kernel is then bound by the j_ion store stream (8 B per evaluation) -> HBM roofline.
//        j_ion leaves through shared-memory staging and TMA tensor stores that always cover whole 32-byte
//        sectors: (n, A) boxes when rows are sector-aligned (kStoreTma), the quad-row view otherwise
//        (kStoreQuad); whole-row bulk copies (kStoreRows) and plain stores (kStoreStg) are the fallbacks.
//        Quadrature sums are thread-local FMAs against weights broadcast from shared memory.
// K1v  eval_lanes4_kernel  : four lanes per sample in the sweep, whole rows per 1-D bulk store.
// K1w  eval_radii_stream_kernel / K1r eval_multi_radius_kernel : several sweep radii.
// K1d  eval_direct_kernel  : one WARP per sample, any angle grid, any number of radii; evaluates
//        the reference's expressions in the reference's operation order (divide, square, negate, exp);
//        warp-shuffle reductions for the two Simpson sums.  Fallback + independent cross-check.
// K2   moments_kernel      : reduce-only Monte-Carlo pass (moments + histograms), optional on-device sampling.
// K3   loglike_kernel      : interpolation to probe angles + Gaussian log-likelihood.
// (K4/K5, the SVD compression of j_ion, live in hpem_compress.cuh.)
//
// Reference lines reproduced: plume.py:95-127,136-140 (per angle / per sample epilogue),
// cathode.py:26-37 and plume.py:40-85 via hpem_device.cuh.
#pragma once
#include <cuda.h>
#include <stdint.h>

#include <type_traits>

#include "hpem_device.cuh"
#include "hpem_qtable.cuh"
#include "hpem_sampler.cuh"

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
    bool bulk_ok;            // j_ion base is 16-byte aligned and bulk (TMA) stores are allowed
    // quad-row store mode (kStoreQuad): per row phase (sample index mod 4) the elements before the first / after the last
    // 32-byte-aligned body element, and the common body length (multiple of 4)
    int q_lead[4], q_tail[4], q_body;
    int l2_hint;             // K1u tensor stores: 0 normal, 1 evict_last, 2 evict_first (see l2_store_policy)
    int no_fastmath;         // 1: always take the libdevice back end for the per-sample part (HPEM_FLAG_NO_FASTMATH)
    QTableRef qt;            // the grid's tabulated Simpson sums (hpem_qtable.cuh); rows == nullptr: sum angle by angle
};

struct JMaps {               // tensor maps of j_ion: [0] the (n, A) view; quad-row mode: one pair per row phase
    CUtensorMap m2[4];       // 2-D, box = 16 columns x rows (clips ragged columns)
    CUtensorMap m3[4];       // 3-D (16 angles, rows, column blocks), box = kTmaCB column blocks
};

constexpr int kChunk = 16;          // angles per staged tile / inner recurrence length
constexpr int kRestartChunks = 16;  // exact exp() restart every kRestartChunks*kChunk angles
constexpr int kTilePitch = kChunk + 1;
// Small blocks: blocks retire and start at a fine grain, which keeps the warps of one SM in different phases (per-sample
// prologue vs store-bound sweep); 1e6 x 200 on B200: 0.336 ms with 128 threads, 0.311 with 64, 0.306 with 32.  One-warp
// blocks win only for K1u's (n, A) tensor-store mode with two staging buffers (long aligned rows); everything else
// (quad-row mode, single-buffer mode for short rows) is 2-6 % faster with two warps per block.
#ifndef HPEM_THREADS_U
#define HPEM_THREADS_U 64
#endif
constexpr int kThreadsU = HPEM_THREADS_U;
constexpr int kWarpsU = kThreadsU / 32;
constexpr double kInvalidFill = 1e-20;  // plume.py:106

__device__ __forceinline__ double load_in(const EvalParams& p, int k, long long s) {
    return p.in[k] ? __ldg(p.in[k] + s) : p.scalar[k];
}

// ---------------------------------------------------------------------------------------------
// TMA helpers (cp.async.bulk.tensor store, shared::cta -> global)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void tma_issue_2d(const CUtensorMap* map, uint32_t smem_src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(map), "r"(c0),
                 "r"(c1), "r"(smem_src)
                 : "memory");
}
__device__ __forceinline__ void tma_issue_3d(const CUtensorMap* map, uint32_t smem_src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(map), "r"(c0),
                 "r"(c1), "r"(c2), "r"(smem_src)
                 : "memory");
}
// L2 eviction policy for the j_ion store stream (K1u).  Measured on B200: when consecutive rows share 128-byte lines (row
// pitch not a multiple of 128 B) evict_last keeps a line in L2 until its other half arrives (A = 91: 0.181 -> 0.174 ms);
// when every row owns its lines evict_first lets finished lines leave early (A = 256 / 512: -1 %); the wrong hint costs
// up to 10 %.  0 = normal, 1 = evict_last, 2 = evict_first.
__device__ __forceinline__ unsigned long long l2_store_policy(int hint) {
    unsigned long long pol;
    if (hint == 1)
        asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    else if (hint == 2)
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    else
        asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void tma_issue_2d(const CUtensorMap* map, uint32_t smem_src, int c0, int c1, unsigned long long pol) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%1, %2}], [%3], %4;" ::"l"(map),
                 "r"(c0), "r"(c1), "r"(smem_src), "l"(pol)
                 : "memory");
}
__device__ __forceinline__ void tma_issue_3d(const CUtensorMap* map, uint32_t smem_src, int c0, int c1, int c2,
                                             unsigned long long pol) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%1, %2, %3}], [%4], %5;" ::"l"(map),
                 "r"(c0), "r"(c1), "r"(c2), "r"(smem_src), "l"(pol)
                 : "memory");
}
__device__ __forceinline__ void tma_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t smem_src, int c0, int c1) {
    tma_issue_2d(map, smem_src, c0, c1);
    tma_commit();
}
__device__ __forceinline__ void bulk_store_1d(double* gdst, uint32_t smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_src), "r"(bytes)
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
#ifndef HPEM_TMA_BUFFERS
#define HPEM_TMA_BUFFERS 2
#endif
constexpr int kTmaBuffers = HPEM_TMA_BUFFERS;
#ifndef HPEM_TMA_CB
#define HPEM_TMA_CB 2
#endif
constexpr int kTmaCB = HPEM_TMA_CB;   // K1u: 16-angle chunks (128-byte column blocks) written by ONE 3-D TMA op
constexpr int kTmaGroupBytes = kTmaCB * kTmaTileBytes;
#ifndef HPEM_ONE_BUFFER_MAX_ANGLES
#define HPEM_ONE_BUFFER_MAX_ANGLES 128
#endif
constexpr int kOneBufferMaxAngles = HPEM_ONE_BUFFER_MAX_ANGLES;   // K1u: single staging buffer per warp up to this angle count
constexpr int kOneBufferMaxAnglesQuad = 192;                      // ... in the quad-row store mode
constexpr int kQuadMinAngles = 64;                                // quad-row store mode from this angle count on
#ifndef HPEM_SMEM_SKEW
#define HPEM_SMEM_SKEW 0          // test hook: shift K1u's shared-memory layout by a multiple of 128 bytes (swizzle vs alignment)
#endif

struct BeamState {  // per-thread recurrence state of one Gaussian beam
    double ec, rc, gc;     // chunk-start profile value, chunk-start ratio, chunk-to-chunk factor
    double q, qk, hh;      // per-sample constants exp(-2x), exp(-2Kx), exp(-2K^2 x)
    double x, amp;
};

template <bool FAST = false>
__device__ __forceinline__ void beam_init(BeamState& b, double h, double a, double amp) {
    const double t = m_div<FAST>(h, a);
    b.x = t * t;                           // profile(i) = exp(-x i^2)
    b.amp = amp;
    // three exps; q = exp(-2x) and hh = exp(-512x) are squares of two of them.  The extra rounding (1 ulp on a factor that
    // is applied <= 16 times between exact re-anchorings) adds ~1.5e-14 to the recurrence's relative error (budget 1e-12).
    b.rc = m_exp<FAST>(-b.x);
    b.q = b.rc * b.rc;
    b.qk = m_exp<FAST>(-(2.0 * kChunk) * b.x);
    b.gc = m_exp<FAST>(-double(kChunk * kChunk) * b.x);
    b.hh = b.gc * b.gc;
    b.ec = 1.0;
}
// same for a sweep whose chunks start at angle index o + 16 c, o in {0, 1, 2, 3} per lane (quad-row store mode: the rows of
// a warp start their 32-byte-aligned body at different angles).  One code path for every offset -- no divergence -- and
// still three exps: the offset-dependent start values are small powers of u = exp(-x) and of qk = exp(-32 x)
// (<= 6 extra roundings, a constant ~5e-16 relative factor on the row).  `u` is returned for the lead elements.
template <bool FAST = false>
__device__ __forceinline__ double beam_init_offset(BeamState& b, double h, double a, double amp, int o) {
    const double t = m_div<FAST>(h, a);
    b.x = t * t;
    b.amp = amp;
    const double u = m_exp<FAST>(-b.x);
    const double u2 = u * u, u4 = u2 * u2;
    b.q = u2;
    b.qk = m_exp<FAST>(-(2.0 * kChunk) * b.x);
    const double g0 = m_exp<FAST>(-double(kChunk * kChunk) * b.x);
    b.hh = g0 * g0;
    // E(o) = u^(o^2), E(o+1)/E(o) = u^(2o+1), E(o+16)/E(o) = g0 * qk^o
    b.ec = (o == 0) ? 1.0 : (o == 1) ? u : (o == 2) ? u4 : u4 * u4 * u;
    b.rc = (o == 0) ? u : (o == 1) ? u2 * u : (o == 2) ? u4 * u : u4 * u2 * u;
    b.gc = (o == 0) ? g0 : (o == 1) ? g0 * b.qk : (o == 2) ? g0 * (b.qk * b.qk) : g0 * (b.qk * b.qk * b.qk);
    return u;
}
template <bool FAST = false>
__device__ __forceinline__ void beam_restart(BeamState& b, int i0) {  // exact values at angle index i0
    const double di = double(i0);
    b.ec = m_exp<FAST>(-b.x * (di * di));
    b.rc = m_exp<FAST>(-b.x * (2.0 * di + 1.0));
    b.gc = m_exp<FAST>(-b.x * (2.0 * kChunk * di + double(kChunk * kChunk)));
}
__device__ __forceinline__ void beam_next_chunk(BeamState& b) {
    b.ec *= b.gc;
    b.gc *= b.hh;
    b.rc *= b.qk;
}

// Plume part of the per-sample prologue (plume.py:40-98 + recurrence start values) for kernels that run one thread per
// sample and may have retired lanes: branch-free back end when every ACTIVE lane of the warp is in the nominal range,
// libdevice otherwise.  Returns which one ran.
__device__ __forceinline__ bool plume_prologue(const EvalParams& p, const double* x_in, SampleConsts& k, double& j_cex,
                                               double& base, BeamState& b1, BeamState& b2) {
    const bool nominal = prologue_nominal(x_in, p.torr, false, true, p.radius0);
    const bool fast = __all_sync(__activemask(), nominal) && !p.no_fastmath;
    if (fast) {
        k = plume_sample_consts<true>(x_in[IN_P_b], x_in[IN_c0], x_in[IN_c1], x_in[IN_c2], x_in[IN_c3], x_in[IN_c4], x_in[IN_c5], p.torr);
        cex_terms<true>(k.density, x_in[IN_sigma], x_in[IN_I_B0], p.radius0, j_cex, base);
        beam_init<true>(b1, p.h, k.a1, __dmul_rn(base, k.amp1));
        beam_init<true>(b2, p.h, k.a2, __dmul_rn(base, k.amp2));
    } else {
        k = plume_sample_consts<false>(x_in[IN_P_b], x_in[IN_c0], x_in[IN_c1], x_in[IN_c2], x_in[IN_c3], x_in[IN_c4], x_in[IN_c5], p.torr);
        cex_terms<false>(k.density, x_in[IN_sigma], x_in[IN_I_B0], p.radius0, j_cex, base);
        beam_init<false>(b1, p.h, k.a1, __dmul_rn(base, k.amp1));
        beam_init<false>(b2, p.h, k.a2, __dmul_rn(base, k.amp2));
    }
    return fast;
}

#ifndef HPEM_MIN_BLOCKS_U
#define HPEM_MIN_BLOCKS_U (384 / HPEM_THREADS_U)      // 12 resident warps per SM with two staging buffers per warp
#endif
// j_ion staging / store modes of K1u
constexpr int kStoreStg = 0;    // 32x16 tile, transposed read-back, plain streaming stores (any A, any alignment)
constexpr int kStoreTma = 1;    // 128B-swizzled sub-tiles, TMA tensor stores (even A, 16-byte aligned base)
constexpr int kStoreRows = 2;   // whole rows of the warp's 32 samples (dense 32 x A tile), ONE contiguous 1-D bulk store
                                // per warp (small odd A when the pair-row mode is switched off)
constexpr int kStoreQuad = 3;   // row pitch not a multiple of 32 bytes (A % 4 != 0, the reference's 91 included).
// Partial 32-byte sectors are what a store stream must avoid on B200: (n, A) tensor boxes over rows that start mid-sector
// run at 4.4 TB/s where sector-aligned rows reach 5.6 (A = 198/202 vs 200), and odd A has no (n, A) tensor map at all
// (rows only 8-byte aligned).  But FOUR consecutive rows always span a multiple of 32 bytes.  So the rows of a warp are
// split by phase (sample index mod 4): phase f skips lead[f] elements to reach a sector boundary, owns a "body" of
// q_body (multiple of 4) elements that IS sector-aligned, and leaves tail[f] elements.  Each phase gets its own tensor
// maps over the quad-row view (n/4 rows of pitch 4A, base = first body element of the phase), lanes are permuted so that
// quarter-warp f holds 8 rows of phase f (one 128B-swizzle atom per column block), and the chunked sweep of a lane simply
// starts at angle lead[f].  The tail of row r and the lead of row r+1 are contiguous in memory and together fill whole
// sectors: they go through a small per-warp shared-memory buffer and are written by 4-8 adjacent lanes per boundary.
constexpr int kBsecSlots = 8;                       // <= 7 tail + 3 lead elements, always 0, 4 or 8 per row boundary
constexpr int kBsecBytes = 32 * kBsecSlots * 8;     // per warp
static_assert(kBsecBytes <= kTmaGroupBytes, "the row-boundary buffer aliases one staging group");

// NBUF staging buffers per warp: 2 overlap the fill of one group with the TMA read of the other (best for long rows);
// 1 halves shared memory and almost doubles the resident warps, which wins while the per-sample prologue dominates
// (A <~ 128: 0.120 ms vs 0.148 ms at 1e6 x 64, B200)
template <bool WANT_PLUME, bool STORE_J, int MODE, int NBUF, int THREADS = kThreadsU>
#ifndef HPEM_RES1
#define HPEM_RES1 640
#endif
__global__ void __launch_bounds__(THREADS, (NBUF == 1 ? HPEM_RES1 : 384) / THREADS)   // 20 resident warps with one staging buffer, 12 with two
eval_uniform_kernel(const EvalParams p, const __grid_constant__ JMaps maps) {
    extern __shared__ __align__(128) unsigned char smem_u[];
    // layout: [staging tiles] [fused weights].  The staging tiles are 128B-swizzled for the TMA engine, which XORs the
    // 16-byte chunk index with ABSOLUTE shared-memory address bits 7..9; the writer below does the same with the row's
    // address, so the tiles only need 128-byte alignment (no 1 KB of slack per block: at 18 KB per block an SM holds 12
    // blocks of the short-row configuration instead of 9).  Quad mode: the row-boundary buffer aliases the staging area
    // once the sweep's tensor stores have been read out.
    unsigned char* smem_al = smem_u + HPEM_SMEM_SKEW;
    constexpr int kThreadsU = THREADS;           // shadow the file-level defaults: this kernel is instantiated per block size
    constexpr int kWarpsU = THREADS / 32;
    constexpr bool QUAD = (MODE == kStoreQuad);
    constexpr bool USE_TMA = (MODE == kStoreTma) || QUAD;
    constexpr bool ROWS = (MODE == kStoreRows);
    constexpr int kGroupBytes = kTmaGroupBytes;
    const int stage_bytes_per_warp = USE_TMA ? NBUF * kGroupBytes
                                             : (ROWS ? ((32 * p.n_angles * 8 + 15) & ~15) : 32 * kTilePitch * 8);
    const int stage_bytes = STORE_J ? kWarpsU * stage_bytes_per_warp : 0;
    double2* wsm = reinterpret_cast<double2*>(smem_al + stage_bytes);

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    unsigned char* stage = smem_al + warp * stage_bytes_per_warp;
    double* bsec = reinterpret_cast<double*>(stage);   // [32 boundaries][8 slots], after the sweep (kBsecBytes <= one staging group)

    // which of the warp's 32 samples this lane owns.  Quad mode: quarter-warp f (lanes 8f .. 8f+7) takes the samples of
    // phase f, lane 8f+q the one in quad-row q -- conflict-free 16-byte stores into one swizzle atom per phase
    const int phase = QUAD ? lane >> 3 : 0;
    const int prl = QUAD ? lane & 7 : lane;      // row of this lane inside its staging sub-tile
    const int lidx = QUAD ? 4 * prl + phase : lane;
    const long long warp_s0 = (long long)blockIdx.x * kThreadsU + warp * 32;
    const long long s_raw = warp_s0 + lidx;
    const bool active = s_raw < p.n;
    const long long s = active ? s_raw : p.n - 1;  // inactive lanes shadow the last sample, never store

    // all per-sample loads are issued up front (15 independent LDG.64 in flight per thread), BEFORE the weights are staged:
    // the block pays one global-memory round trip at its start, not two
    double x_in[kNumInputs];
#ifndef HPEM_LOADS_LATE
#pragma unroll
    for (int q = 0; q < kNumInputs; ++q) {
        const bool needed = (q == IN_P_b) || (q <= IN_P_T ? p.v_cc != nullptr : (q == IN_T ? p.t_c != nullptr : WANT_PLUME));
        x_in[q] = needed ? load_in(p, q, s) : 0.0;
    }
#endif
    if (WANT_PLUME) {
        const int n_w = ((p.n_angles + kChunk - 1) / kChunk) * kChunk;   // the sweep never reads a weight beyond its last chunk
        for (int i = threadIdx.x; i < n_w; i += kThreadsU) wsm[i] = p.w[i];
        __syncthreads();
    }
#ifdef HPEM_LOADS_LATE
#pragma unroll
    for (int q = 0; q < kNumInputs; ++q) {
        const bool needed = (q == IN_P_b) || (q <= IN_P_T ? p.v_cc != nullptr : (q == IN_T ? p.t_c != nullptr : WANT_PLUME));
        x_in[q] = needed ? load_in(p, q, s) : 0.0;
    }
#endif
    if (warp_s0 >= p.n) return;  // whole warp out of range (after the only __syncthreads)

    // per-sample prologue (cathode.py:26-37, plume.py:40-98, recurrence start values) in one of two arithmetic back ends:
    // branch-free (hpem_fastmath.cuh) when all 32 samples of the warp are in the nominal range, libdevice otherwise
    const bool want_cathode = p.v_cc != nullptr;
    const int off = QUAD ? p.q_lead[phase] : 0;  // angle index of the chunked sweep's first element
    SampleConsts k;
    BeamState b1, b2;
    double v_cc = 0.0, j_cex = 0.0, base = 0.0;
    double u1 = 0.0, u2 = 0.0;                   // exp(-x) of the two beams (quad mode: the lead elements need E(1), E(2))
    double num = 0.0, den = 0.0;                 // the Simpson sums of plume.py:121-122
    // Only the variant that stores no j_ion takes the sums from the grid's table (and then needs no sweep at all).  With
    // stores the table loses: two divergent 160-byte lookups per sample are ~45 % more L2 read traffic next to the store
    // stream (B200, 1e6 x 91: 0.163 -> 0.210 ms; x 200: 0.316 -> 0.333 ms), more than the two fused multiply-adds per angle cost.
    // (below ~50 angles the two lookups cost more than the sweep they replace: 1e6 x 17 no-store 0.083 vs 0.092 ms)
    const bool use_qt = !STORE_J && p.qt.rows != nullptr && p.n_angles >= 64;
    auto prologue = [&](auto fast_tag) {
        constexpr bool FAST = decltype(fast_tag)::value;
        if (want_cathode)
            v_cc = cathode_vcc<FAST>(x_in[IN_P_b], x_in[IN_V_a], x_in[IN_T_e], x_in[IN_V_vac], x_in[IN_Pstar], x_in[IN_P_T], p.torr);
        if (WANT_PLUME) {
            k = plume_sample_consts<FAST>(x_in[IN_P_b], x_in[IN_c0], x_in[IN_c1], x_in[IN_c2], x_in[IN_c3], x_in[IN_c4],
                                          x_in[IN_c5], p.torr);
            cex_terms<FAST>(k.density, x_in[IN_sigma], x_in[IN_I_B0], p.radius0, j_cex, base);
            if (QUAD) {
                u1 = beam_init_offset<FAST>(b1, p.h, k.a1, __dmul_rn(base, k.amp1), off);
                u2 = beam_init_offset<FAST>(b2, p.h, k.a2, __dmul_rn(base, k.amp2), off);
            } else {
                beam_init<FAST>(b1, p.h, k.a1, __dmul_rn(base, k.amp1));   // (base_density * A1), plume.py:99
                beam_init<FAST>(b2, p.h, k.a2, __dmul_rn(base, k.amp2));   // (base_density * A2), plume.py:100
                u1 = b1.rc;
                u2 = b2.rc;
            }
            if (FAST && use_qt) {   // both sums from the grid's table: amplitude x N((h / alpha)^2) per beam (hpem_qtable.cuh)
                double nd1, nn1, nd2, nn2;
                qtable_eval(p.qt, b1.x, u1, nd1, nn1);
                qtable_eval(p.qt, b2.x, u2, nd2, nn2);
                den = fma(b1.amp, nd1, b2.amp * nd2);
                num = fma(b1.amp, nn1, b2.amp * nn2);
            }
        }
    };
    const bool nominal = prologue_nominal(x_in, p.torr, want_cathode, WANT_PLUME, p.radius0);
    const bool fast = __all_sync(0xffffffffu, nominal) && !p.no_fastmath;
    if (fast)
        prologue(std::true_type{});
    else
        prologue(std::false_type{});
    if (want_cathode && active) p.v_cc[s] = v_cc;
    if (!WANT_PLUME) return;

    const bool known_invalid = (k.a1 <= 0.0);  // plume.py:105 first term
    // With non-negative beam amplitudes and a positive CEX floor every j_ion is > 0 (or NaN), so the per-angle
    // `j_ion <= 0` test of plume.py:105 cannot fire; only warps holding an exceptional sample run the checked loop.
    const bool needs_check = known_invalid || !(b1.amp >= 0.0 && b2.amp >= 0.0 && j_cex > 0.0);
    bool bad = false;
    // no j_ion wanted: warps of the fast back end whose rows need no validity check hold the two sums already (table) and
    // skip the sweep; every other warp accumulates them angle by angle
    const bool tabulated = fast && use_qt && !__any_sync(0xffffffffu, needs_check);
    if (!tabulated) num = den = 0.0;

    const int A = p.n_angles;
    const int A_sweep = QUAD ? p.q_body : A;     // angles covered by the chunked sweep
    const int n_chunks = (A_sweep + kChunk - 1) / kChunk;
    const int rows_valid = (int)min((long long)32, p.n - warp_s0);
    const int col = lane & (kChunk - 1);
    const int rsub = lane >> 4;
    const int tma_row0 = QUAD ? (int)(warp_s0 >> 2) : (int)warp_s0;
    double tail_e1 = 0.0, tail_e2 = 0.0, tail_r1 = 0.0, tail_r2 = 0.0;   // recurrence state one step past the sweep (quad mode)

    // one element outside the chunked sweep (quad mode): quadrature, validity test, value into the row-boundary buffer
    auto edge_element = [&](double sum, int angle, int slot_row, int slot) {
        const double j = sum + j_cex;
        const double2 w = wsm[angle];
        den = fma(w.x, sum, den);
        num = fma(w.y, sum, num);
        bad |= (j <= 0.0);
        if (STORE_J) bsec[slot_row * kBsecSlots + slot] = known_invalid ? kInvalidFill : j;
    };
    auto chunk_loop = [&](auto checked_tag) {
        constexpr bool CHECKED = decltype(checked_tag)::value;
        for (int c = 0; c < n_chunks; ++c) {
            const int i0 = c * kChunk;
            if (c != 0 && (c % kRestartChunks) == 0) {  // exact restart bounds the recurrence error for large A
                if (fast) {                             // same back end as the prologue (and as K2: its rows match K1u's bit for bit)
                    beam_restart<true>(b1, i0 + off);
                    beam_restart<true>(b2, i0 + off);
                } else {
                    beam_restart<false>(b1, i0 + off);
                    beam_restart<false>(b2, i0 + off);
                }
            }
            double e1 = b1.amp * b1.ec, e2 = b2.amp * b2.ec;
            double r1 = b1.rc, r2 = b2.rc;
            const int kcount = min(kChunk, A_sweep - i0);
            // TMA staging: [buffer][phase (quad mode)][column block within the group][rows][128 B, 16-byte chunks XOR-swizzled by row]
            unsigned char* group_buf = stage + ((c / kTmaCB) % NBUF) * kGroupBytes;
            // kStoreRows: dense [32][A] tile; for odd A the row pitch A*8 bytes walks all 16 bank pairs -> conflict-free
            unsigned char* my_row =
                QUAD ? group_buf + phase * (kTmaCB * 1024) + (c % kTmaCB) * 1024 + prl * (kChunk * 8)
                     : (USE_TMA ? group_buf + (c % kTmaCB) * kTmaTileBytes + lane * (kChunk * 8)
                                : (ROWS ? stage + (size_t(lane) * A + i0) * 8 : stage + lane * (kTilePitch * 8)));
            const double2* wrow = wsm + i0 + off;
            const int swz = USE_TMA ? (int)((smem_u32(my_row) >> 7) & 7u) : 0;   // the TMA engine's 128B swizzle: chunk ^ address bits 7..9
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
                    step(wrow[kk], ja);
                    step(wrow[kk + 1], jb);
                    if (STORE_J) {
                        if (USE_TMA) {
                            *reinterpret_cast<double2*>(my_row + (((kk >> 1) ^ swz) << 4)) = make_double2(ja, jb);
                        } else {
                            reinterpret_cast<double*>(my_row)[kk] = ja;
                            reinterpret_cast<double*>(my_row)[kk + 1] = jb;
                        }
                    }
                }
            } else {
                for (int kk = 0; kk < kcount; ++kk) {
                    double ja;
                    step(wrow[kk], ja);
                    if (STORE_J) {
                        if (USE_TMA)
                            *reinterpret_cast<double*>(my_row + (((kk >> 1) ^ swz) << 4) + ((kk & 1) << 3)) = ja;
                        else
                            reinterpret_cast<double*>(my_row)[kk] = ja;
                    }
                }
            }
            if (QUAD) {
                tail_e1 = e1; tail_e2 = e2;
                tail_r1 = r1; tail_r2 = r2;
            }
            beam_next_chunk(b1);
            beam_next_chunk(b2);

            if (STORE_J) {
                if (USE_TMA) {
                    // kTmaCB chunks are shipped by ONE 3-D TMA op (per phase): rows x (kTmaCB x 128 B) contiguous row pieces
                    // (6.2-6.5 TB/s on B200 against 5.6 TB/s for single 128-byte pieces, tools/store_pattern.cu).
                    // A trailing group that is incomplete or holds the partial last column block goes out as 2-D
                    // boxes of 16 columns, whose tensor map clips columns >= A_sweep.  Rows >= n are clipped by all maps.
                    const bool last_chunk = (c == n_chunks - 1);
                    if ((c % kTmaCB) == kTmaCB - 1 || last_chunk) {
                        fence_async_smem();
                        __syncwarp();
                        const int c_first = c - (c % kTmaCB);
                        const bool whole = c_first + kTmaCB <= A_sweep / kChunk;
                        const unsigned long long pol = l2_store_policy(p.l2_hint);
                        if (QUAD) {
                            // lanes 0-3 ship the four phases in parallel (one instruction sequence instead of four);
                            // bulk-group bookkeeping is per thread, so each of them commits and waits for its own ops
                            if (lane < 4) {
                                const uint32_t src = smem_u32(group_buf + lane * (kTmaCB * 1024));
                                if (whole) {
                                    tma_issue_3d(&maps.m3[lane], src, 0, tma_row0, c_first, pol);
                                } else {
                                    for (int cc = c_first; cc <= c; ++cc)
                                        tma_issue_2d(&maps.m2[lane], src + (cc - c_first) * 1024, cc * kChunk, tma_row0, pol);
                                }
                                tma_commit();
                                tma_wait_read<NBUF - 1>();   // the buffer the next group goes into is free again
                            }
                        } else if (lane == 0) {
                            if (whole) {
                                tma_issue_3d(&maps.m3[0], smem_u32(group_buf), 0, tma_row0, c_first, pol);
                            } else {
                                for (int cc = c_first; cc <= c; ++cc)
                                    tma_issue_2d(&maps.m2[0], smem_u32(group_buf + (cc - c_first) * kTmaTileBytes), cc * kChunk, tma_row0, pol);
                            }
                            tma_commit();
                            tma_wait_read<NBUF - 1>();   // the buffer the next group goes into is free again
                        }
                        __syncwarp();
                    }
                } else if (!ROWS) {
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
    if (tabulated) {
        // nothing to store and nothing to sum: div_angle / T_c / cos_div need no sweep at all
    } else if (__any_sync(0xffffffffu, needs_check)) {
        chunk_loop(std::true_type{});
    } else {
        chunk_loop(std::false_type{});
    }

    if (QUAD) {
        if (STORE_J) {   // the row-boundary buffer aliases the staging area: every tensor store must have read its tile
            if (lane < 4) tma_wait_read<0>();
            __syncwarp();
        }
        if (off > 0) {
            // lead elements (angles 0 .. off-1 <= 2): E(0) = 1, E(1) = u, E(2) = u^4.  They complete the sector(s) that the
            // previous row's tail starts: boundary lidx-1, after that row's tail[phase-1] slots.
            const int t_prev = p.q_tail[(phase + 3) & 3];
            edge_element(b1.amp + b2.amp, 0, lidx - 1, t_prev);
            if (off > 1) edge_element(b1.amp * u1 + b2.amp * u2, 1, lidx - 1, t_prev + 1);
            if (off > 2) {
                const double v1 = u1 * u1, v2 = u2 * u2;
                edge_element(b1.amp * (v1 * v1) + b2.amp * (v2 * v2), 2, lidx - 1, t_prev + 2);
            }
        }
        // tail elements (angles off + q_body .. A-1, at most 7): the recurrence simply continues.  They open the
        // sector(s) that the next row's lead completes: boundary lidx, slots 0 .. tail-1.
        const int n_tail = p.q_tail[phase];
        const int a0 = off + A_sweep;
        for (int t = 0; t < n_tail; ++t) {
            edge_element(tail_e1 + tail_e2, a0 + t, lidx, t);
            tail_e1 *= tail_r1; tail_r1 *= b1.q;
            tail_e2 *= tail_r2; tail_r2 *= b2.q;
        }
        if (STORE_J) {
            // write the row boundaries: boundary r = tail of row r + lead of row r+1, contiguous in memory and a whole
            // number of sectors; 8 lanes per boundary, so each sector is written by adjacent lanes of one instruction
            __syncwarp();
            // element e = 32 it + lane of the [32 boundaries][8 slots] buffer: boundary r = 4 it + lane / 8, slot = lane % 8,
            // so the phase of row r -- and with it tail and count -- is a per-lane constant
            const int f = lane >> 3, slot = lane & 7;
            const int t_r = p.q_tail[f];
            const bool slot_ok = slot < t_r + p.q_lead[(f + 1) & 3];
            double* gp = p.j_ion + (warp_s0 + f + 1) * (long long)A - t_r + slot;
            const double* bp = bsec + lane;
#pragma unroll
            for (int it = 0; it < kBsecSlots; ++it) {
                if (slot_ok && warp_s0 + 4 * it + f < p.n) *gp = bp[it * 32];         // n % 4 == 0: row r+1 exists whenever its lead is non-empty
                gp += 4 * (long long)A;
            }
        }
    }

    if (STORE_J && ROWS) {   // ship the warp's rows: rows_valid * A * 8 contiguous bytes starting at a 16-byte aligned address
        const uint32_t bytes = (uint32_t)rows_valid * (uint32_t)A * 8u;
        double* gdst = p.j_ion + warp_s0 * (long long)A;
        if (p.bulk_ok && (bytes & 15u) == 0) {
            fence_async_smem();
            __syncwarp();
            if (lane == 0) bulk_store_1d(gdst, smem_u32(stage), bytes);
        } else {              // ragged last warp with an odd byte count, or a misaligned output: plain coalesced stores
            __syncwarp();
            const double* t = reinterpret_cast<const double*>(stage);
            for (int e = lane; e < rows_valid * A; e += 32) __stcs(gdst + e, t[e]);
        }
    }
    // rare: a non-positive j_ion found after earlier chunks were already written -> the row is overwritten below
    const bool late_fix = STORE_J && __any_sync(0xffffffffu, bad && !known_invalid);
    if (STORE_J && (USE_TMA || ROWS)) {
        const bool issuer = QUAD ? lane < 4 : lane == 0;
        if (late_fix) {            // order the TMA (async proxy) writes before the generic-proxy rewrite
            if (issuer) tma_wait_all();
            fence_async_all();
        } else if (issuer) {
            tma_wait_read<0>();    // the staging tiles must outlive the TMA reads
        }
        __syncwarp();
    }

    // per-sample epilogue: plume.py:124-127,137 (NOT masked by `invalid`)
    double cd, dv;
    if (fast && __all_sync(0xffffffffu, fm_mid(den) && fm_mid0(num))) {
        cd = fm_div(num, den);
        dv = fm_acos(cd);
    } else {
        cd = num / den;
        if (cd == CUDART_INF) cd = CUDART_NAN;  // plume.py:125
        dv = acos(cd);
    }
    const bool invalid = known_invalid || bad;
    if (active) {
        if (p.div_angle) p.div_angle[s] = dv;
        if (p.cos_div) p.cos_div[s] = cd;
        if (p.t_c) p.t_c[s] = __dmul_rn(x_in[IN_T], cd);
        if (p.invalid) p.invalid[s] = invalid ? 1 : 0;
        if (STORE_J && bad && !known_invalid) {
            // plume.py:106.  All of this warp's earlier stores are complete and ordered before this point.
            double* row = p.j_ion + s * (long long)A;
            for (int i = 0; i < A; ++i) row[i] = kInvalidFill;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// K1r: several sweep radii (j_ion (n, A, R), radius fastest) -- thread per sample, recurrence sweep
// ---------------------------------------------------------------------------------------------
// plume.py:95-102 with R radii: only `decay`, `j_cex` and `base_density` depend on the radius, the Gaussian profiles do
// not.  Each thread keeps base_rho and j_cex_rho of its sample in shared memory ([rho][thread], conflict-free), runs
// ONE recurrence sweep over the angles and emits R values per angle: j = base_rho * (A1 E1 + A2 E2) + j_cex_rho.
// The flattened (angle, radius) columns are staged 16 at a time exactly like K1u (32x16 TMA boxes over an
// (n, A*R) tensor, or plain stores when A*R is odd).  The two Simpson sums are radius-independent up to the factor
// base_rho, so they are accumulated once and scaled per radius at the end (0/0 = NaN for an opaque plume, as in NumPy).
constexpr int kMaxRadiiFast = 48;   // 2 * R * 128 * 8 B of shared memory per block

template <bool USE_TMA>
__global__ void __launch_bounds__(kThreadsU, 192 / kThreadsU) eval_multi_radius_kernel(const EvalParams p,
                                                                         const __grid_constant__ CUtensorMap jmap) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem_al = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    constexpr int kStageBytesPerWarp = USE_TMA ? kTmaBuffers * kTmaTileBytes : 32 * kTilePitch * 8;
    const int R = p.n_radii, A = p.n_angles;
    double* rad_base = reinterpret_cast<double*>(smem_al + kWarpsU * kStageBytesPerWarp);   // [R][kThreadsU]
    double* rad_cex = rad_base + R * kThreadsU;                                              // [R][kThreadsU]
    double2* wsm = reinterpret_cast<double2*>(rad_cex + R * kThreadsU);

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, tid = threadIdx.x;
    unsigned char* stage = smem_al + warp * kStageBytesPerWarp;
    for (int i = tid; i < p.n_angles_pad; i += kThreadsU) wsm[i] = p.w[i];
    __syncthreads();

    const long long s_raw = (long long)blockIdx.x * kThreadsU + tid;
    const bool active = s_raw < p.n;
    const long long s = active ? s_raw : p.n - 1;
    const long long warp_s0 = s_raw - lane;
    if (warp_s0 >= p.n) return;

    double x_in[kNumInputs];
#pragma unroll
    for (int q = 0; q < kNumInputs; ++q) {
        const bool needed = (q == IN_P_b) || (q <= IN_P_T ? p.v_cc != nullptr : (q == IN_T ? p.t_c != nullptr : true));
        x_in[q] = needed ? load_in(p, q, s) : 0.0;
    }
    if (p.v_cc) {
        const double v = cathode_vcc(x_in[IN_P_b], x_in[IN_V_a], x_in[IN_T_e], x_in[IN_V_vac], x_in[IN_Pstar],
                                     x_in[IN_P_T], p.torr);
        if (active) p.v_cc[s] = v;
    }
    const SampleConsts k = plume_sample_consts(x_in[IN_P_b], x_in[IN_c0], x_in[IN_c1], x_in[IN_c2], x_in[IN_c3],
                                               x_in[IN_c4], x_in[IN_c5], p.torr);
    const bool known_invalid = (k.a1 <= 0.0);
    bool needs_check = known_invalid || !(k.amp1 >= 0.0 && k.amp2 >= 0.0);
    for (int rho = 0; rho < R; ++rho) {
        double j_cex, base;
        cex_terms(k.density, x_in[IN_sigma], x_in[IN_I_B0], __ldg(p.radii + rho), j_cex, base);
        rad_base[rho * kThreadsU + tid] = base;
        rad_cex[rho * kThreadsU + tid] = j_cex;
        needs_check |= !(base >= 0.0 && j_cex > 0.0);
    }
    BeamState b1, b2;
    beam_init(b1, p.h, k.a1, k.amp1);   // amplitudes WITHOUT base_density: it is applied per radius
    beam_init(b2, p.h, k.a2, k.amp2);

    bool bad = false;
    double s_num = 0.0, s_den = 0.0;
    const int n_chunks = (A + kChunk - 1) / kChunk;
    const long long row_len = (long long)A * R;
    const int rows_valid = (int)min((long long)32, p.n - warp_s0);
    const int colq = lane & (kChunk - 1), rsub = lane >> 4;
    long long col_done = 0;   // flattened columns already flushed
    int fill = 0;             // columns in the current tile
    int tile_it = 0;

    auto flush = [&](int ncols) {
        unsigned char* tile = USE_TMA ? stage + (tile_it % kTmaBuffers) * kTmaTileBytes : stage;
        if (USE_TMA) {
            fence_async_smem();
            __syncwarp();
            if (lane == 0) {
                tma_store_2d(&jmap, smem_u32(tile), (int)col_done, (int)warp_s0);
                tma_wait_read<kTmaBuffers - 1>();
            }
            __syncwarp();
        } else {
            __syncwarp();
            const double* trow = reinterpret_cast<const double*>(tile) + rsub * kTilePitch + colq;
            double* g = p.j_ion + (warp_s0 + rsub) * row_len + col_done + colq;
            const bool col_ok = colq < ncols;
#pragma unroll
            for (int rr = 0; rr < 16; ++rr) {
                if (col_ok && (2 * rr + rsub) < rows_valid) __stcs(g, trow[2 * rr * kTilePitch]);
                g += 2 * row_len;
            }
            __syncwarp();
        }
        col_done += ncols;
        ++tile_it;
    };
    auto put = [&](int pos, double v) {
        if (USE_TMA) {
            unsigned char* row = stage + (tile_it % kTmaBuffers) * kTmaTileBytes + lane * (kChunk * 8);
            *reinterpret_cast<double*>(row + (((pos >> 1) ^ (lane & 7)) << 4) + ((pos & 1) << 3)) = v;
        } else {
            reinterpret_cast<double*>(stage)[lane * kTilePitch + pos] = v;
        }
    };

    const bool checked = __any_sync(0xffffffffu, needs_check);
    for (int c = 0; c < n_chunks; ++c) {
        const int i0 = c * kChunk;
        if (c != 0 && (c % kRestartChunks) == 0) {
            beam_restart(b1, i0);
            beam_restart(b2, i0);
        }
        double e1 = b1.amp * b1.ec, e2 = b2.amp * b2.ec;
        double r1 = b1.rc, r2 = b2.rc;
        const int kcount = min(kChunk, A - i0);
        for (int kk = 0; kk < kcount; ++kk) {
            const double2 w = wsm[i0 + kk];
            const double g = e1 + e2;                      // A1 E1 + A2 E2
            s_den = fma(w.x, g, s_den);
            s_num = fma(w.y, g, s_num);
            for (int rho = 0; rho < R; ++rho) {
                const double j = fma(rad_base[rho * kThreadsU + tid], g, rad_cex[rho * kThreadsU + tid]);
                if (checked) bad |= (j <= 0.0);
                put(fill, known_invalid ? kInvalidFill : j);
                if (++fill == kChunk) {
                    flush(kChunk);
                    fill = 0;
                }
            }
            e1 *= r1; r1 *= b1.q;
            e2 *= r2; r2 *= b2.q;
        }
        beam_next_chunk(b1);
        beam_next_chunk(b2);
    }
    if (fill > 0) flush(fill);   // TMA clips columns >= A*R

    const bool late_fix = __any_sync(0xffffffffu, bad && !known_invalid);
    if (USE_TMA) {
        if (late_fix) {
            if (lane == 0) tma_wait_all();
            fence_async_all();
        } else if (lane == 0) {
            tma_wait_read<0>();
        }
        __syncwarp();
    }
    if (active) {
        for (int rho = 0; rho < R; ++rho) {
            const double base = rad_base[rho * kThreadsU + tid];
            double cd = __dmul_rn(base, s_num) / __dmul_rn(base, s_den);   // plume.py:122-124 per radius
            if (cd == CUDART_INF) cd = CUDART_NAN;
            const long long o = s * R + rho;
            if (p.div_angle) p.div_angle[o] = acos(cd);
            if (p.cos_div) p.cos_div[o] = cd;
            if (p.t_c) p.t_c[o] = __dmul_rn(x_in[IN_T], cd);
        }
        if (p.invalid) p.invalid[s] = (known_invalid || bad) ? 1 : 0;
        if (bad && !known_invalid) {
            double* row = p.j_ion + s * row_len;
            for (long long i = 0; i < row_len; ++i) row[i] = kInvalidFill;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// K1v: thread per sample for the per-sample work, FOUR lanes per sample for the angle sweep,
//      whole rows staged in shared memory and written with ONE bulk (TMA) copy per 8 samples
// ---------------------------------------------------------------------------------------------
// HBM write efficiency depends on how many contiguous bytes reach L2 together: 128-byte row pieces (K1u's
// 32 samples x 16 angles box) stream at ~5.6 TB/s on B200, 512-byte pieces at ~7.2 TB/s, >= 1 KB at ~7.4 TB/s
// (tools/store_pattern.cu).  j_ion rows of consecutive samples are adjacent in memory, so K1v lets a group of
// 8 samples sweep ALL angles into a dense shared-memory tile (8 x A doubles) and ships the tile as one contiguous
// 8*A*8-byte cp.async.bulk store (works for odd A too: the group start is always 16-byte aligned).
// Angle counts whose tile would exceed 16 KB are processed in panels of 256 angles (2 KB row pieces).
// The sweep uses 4 lanes per sample, lane q taking angles i = q (mod 4) with a stride-4 Gaussian recurrence.  Per warp:
//   phase 1  lane t <-> sample s0+t : loads, cathode, A1/A2, CEX terms, the 16 exps the recurrences start from
//   phase 2  4 groups of 8 samples, lane (r = lane/4, q = lane%4): sweep + quadrature partial sums + tile + store
//   phase 3  lane t <-> sample s0+t : cos_div, arccos, T_c, invalid mask, rare late fix-up
// Only __syncwarp() separates the phases (each warp owns its 32 samples and its shared-memory slices).
constexpr int kL4 = 4;                        // lanes per sample in the sweep
constexpr int kSteps4 = 16;                   // recurrence steps per chunk (chunk = 64 angles)
constexpr int kCols4 = kL4 * kSteps4;
constexpr int kRows4 = 32 / kL4;              // 8 samples per group
#ifndef HPEM_PANEL_MAX
#define HPEM_PANEL_MAX 256
#endif
constexpr int kPanelMax = HPEM_PANEL_MAX;                // angles per staged panel (tile = 8 x 256 x 8 B = 16 KB)
constexpr int kRestartChunks4 = 8;            // exact restart every 512 angles
constexpr int kAnglePad = kCols4;             // weights are zero-padded to a multiple of this
// doubles per sample handed from phase 1 to phase 2: amp1 amp2 j_cex flags x1 x2 + 8 base exps per beam
constexpr int kXch = 22;
constexpr int kXchPitch = 23;                 // odd pitch: conflict-free column access


// exp(-x*m) for the multipliers the stride-4 recurrence is assembled from
__device__ __forceinline__ void beam_base_exps(double x, double* o) {
    o[0] = exp(-x);                                   // u
    o[1] = exp(-8.0 * x);                             // ratio step in q        (2 L)
    o[2] = exp(-16.0 * x);                            // ratio at i = 0         (L^2)
    o[3] = exp(-32.0 * x);                            // q  : ratio growth      (2 L^2)
    o[4] = exp(-128.0 * x);                           // chunk factor step in q (2 L K)
    o[5] = exp(-512.0 * x);                           // qk : ratio growth per chunk (2 L^2 K)
    o[6] = exp(-4096.0 * x);                          // chunk factor at i = 0  ((L K)^2)
    o[7] = exp(-8192.0 * x);                          // hh : chunk factor growth (2 (L K)^2)
}

struct BeamLane {       // per-lane recurrence state of one beam, stride-4 sweep
    double ec, rc, gc;  // chunk-start value / ratio / chunk-to-chunk factor
    double q, qk, hh;
    double x, amp;
};
// lane-specific start values for angle index q in {0,1,2,3} from the per-sample base exps (products of <= 4 factors)
__device__ __forceinline__ void beam_lane_init(BeamLane& b, const double* xr, int q) {
    const double u = xr[0], e8 = xr[1], e16 = xr[2], e128 = xr[4], e4096 = xr[6];
    b.q = xr[3];
    b.qk = xr[5];
    b.hh = xr[7];
    const double u2 = u * u, u4 = u2 * u2, u8 = u4 * u4;
    const double e8_2 = e8 * e8, e128_2 = e128 * e128;
    // exp(-x q^2), exp(-x (8 q + 16)), exp(-x (128 q + 4096))
    b.ec = (q == 0) ? 1.0 : (q == 1) ? u : (q == 2) ? u4 : u8 * u;
    b.rc = e16 * ((q == 0) ? 1.0 : (q == 1) ? e8 : (q == 2) ? e8_2 : e8_2 * e8);
    b.gc = e4096 * ((q == 0) ? 1.0 : (q == 1) ? e128 : (q == 2) ? e128_2 : e128_2 * e128);
}
__device__ __forceinline__ void beam_lane_restart(BeamLane& b, int i0) {  // exact values at angle index i0
    const double di = double(i0);
    b.ec = exp(-b.x * (di * di));
    b.rc = exp(-b.x * (2.0 * kL4 * di + double(kL4 * kL4)));
    b.gc = exp(-b.x * (2.0 * kCols4 * di + double(kCols4 * kCols4)));
}
__device__ __forceinline__ void beam_lane_next(BeamLane& b) {
    b.ec *= b.gc;
    b.gc *= b.hh;
    b.rc *= b.qk;
}

#ifndef HPEM_THREADS_V
#define HPEM_THREADS_V 64
#endif
#ifndef HPEM_MIN_BLOCKS_V
#define HPEM_MIN_BLOCKS_V 8
#endif
constexpr int kThreadsV = HPEM_THREADS_V;
constexpr int kWarpsV = kThreadsV / 32;

// shared memory of K1v: [tiles: kWarpsV x tile_bytes (16-byte aligned)] [exchange] [weights]
__host__ __device__ inline int k1v_panel_cols(int n_angles) { return n_angles <= kPanelMax ? n_angles : kPanelMax; }
__host__ __device__ inline size_t k1v_tile_bytes(int n_angles) { return size_t(kRows4) * k1v_panel_cols(n_angles) * 8; }

template <bool STORE_J>
__global__ void __launch_bounds__(kThreadsV, HPEM_MIN_BLOCKS_V) eval_lanes4_kernel(const EvalParams p) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int A = p.n_angles;
    const int panel_cols = k1v_panel_cols(A);
    const int tile_bytes = STORE_J ? (int)k1v_tile_bytes(A) : 0;
    double* xch_all = reinterpret_cast<double*>(smem_raw + kWarpsV * tile_bytes);
    double2* wsm = reinterpret_cast<double2*>(xch_all + kWarpsV * 32 * kXchPitch);

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    double* tile = reinterpret_cast<double*>(smem_raw + warp * tile_bytes);
    double* xch = xch_all + warp * 32 * kXchPitch;

    for (int i = threadIdx.x; i < p.n_angles_pad; i += kThreadsV) wsm[i] = p.w[i];
    __syncthreads();

    const long long s_raw = (long long)blockIdx.x * kThreadsV + threadIdx.x;
    const bool active = s_raw < p.n;
    const long long s = active ? s_raw : p.n - 1;  // inactive lanes shadow the last sample, never store
    const long long warp_s0 = s_raw - lane;
    if (warp_s0 >= p.n) return;

    // ---------------- phase 1: one lane per sample ----------------
    double thrust = 0.0;
    bool known_invalid;
    {
        double x_in[kNumInputs];
#pragma unroll
        for (int q = 0; q < kNumInputs; ++q) {
            const bool needed = (q == IN_P_b) || (q <= IN_P_T ? p.v_cc != nullptr : (q == IN_T ? p.t_c != nullptr : true));
            x_in[q] = needed ? load_in(p, q, s) : 0.0;
        }
        thrust = x_in[IN_T];
        if (p.v_cc) {
            const double v = cathode_vcc(x_in[IN_P_b], x_in[IN_V_a], x_in[IN_T_e], x_in[IN_V_vac], x_in[IN_Pstar],
                                         x_in[IN_P_T], p.torr);
            if (active) p.v_cc[s] = v;
        }
        const SampleConsts k = plume_sample_consts(x_in[IN_P_b], x_in[IN_c0], x_in[IN_c1], x_in[IN_c2], x_in[IN_c3],
                                                   x_in[IN_c4], x_in[IN_c5], p.torr);
        double j_cex, base;
        cex_terms(k.density, x_in[IN_sigma], x_in[IN_I_B0], p.radius0, j_cex, base);
        const double amp1 = __dmul_rn(base, k.amp1);  // (base_density * A1), plume.py:99
        const double amp2 = __dmul_rn(base, k.amp2);  // (base_density * A2), plume.py:100
        const double t1 = p.h / k.a1, t2 = p.h / k.a2;
        const double x1 = t1 * t1, x2 = t2 * t2;      // profile_b(i) = exp(-x_b i^2)
        known_invalid = (k.a1 <= 0.0);                 // plume.py:105 first term
        // non-negative amplitudes and a positive CEX floor => every j_ion > 0 (or NaN): no per-angle test needed
        const bool needs_check = known_invalid || !(amp1 >= 0.0 && amp2 >= 0.0 && j_cex > 0.0);
        double* xr = xch + lane * kXchPitch;
        xr[0] = amp1;
        xr[1] = amp2;
        xr[2] = j_cex;
        xr[3] = __longlong_as_double((long long)((known_invalid ? 1 : 0) | (needs_check ? 2 : 0)));
        xr[4] = x1;
        xr[5] = x2;
        double ex[8];
        beam_base_exps(x1, ex);
#pragma unroll
        for (int j = 0; j < 8; ++j) xr[6 + j] = ex[j];
        beam_base_exps(x2, ex);
#pragma unroll
        for (int j = 0; j < 8; ++j) xr[14 + j] = ex[j];
    }
    __syncwarp();

    // ---------------- phase 2: four lanes per sample ----------------
    const int r = lane >> 2, q = lane & 3;
    const int qbase = lane & ~3;
    const int n_panels = (A + panel_cols - 1) / panel_cols;
    const bool bulk_rows_ok = (n_panels == 1);                 // whole rows staged: contiguous 1-D bulk store
    const bool bulk_panel_ok = (A % 2 == 0);                   // panel rows are 16-byte aligned only for even A
#pragma unroll 1
    for (int g = 0; g < kL4; ++g) {
        const long long grow0 = warp_s0 + g * kRows4;
        if (grow0 >= p.n) break;  // warp-uniform
        double* xr = xch + (g * kRows4 + r) * kXchPitch;
        BeamLane b1, b2;
        b1.amp = xr[0];
        b2.amp = xr[1];
        const double j_cex = xr[2];
        const int flags = (int)__double_as_longlong(xr[3]);
        b1.x = xr[4];
        b2.x = xr[5];
        beam_lane_init(b1, xr + 6, q);
        beam_lane_init(b2, xr + 14, q);
        const bool fill_invalid = flags & 1;
        const int rows_valid = (int)min((long long)kRows4, p.n - grow0);
        double num = 0.0, den = 0.0;
        bool bad = false;

        auto sweep = [&](auto checked_tag) {
            constexpr bool CHECKED = decltype(checked_tag)::value;
            int chunk = 0;
#pragma unroll 1
            for (int pn = 0; pn < n_panels; ++pn) {
                const int col0 = pn * panel_cols;                  // first angle of the panel (multiple of 64 or 0)
                const int cols = min(panel_cols, A - col0);
                if (STORE_J) {   // the previous bulk store must have finished READING the tile before it is refilled
                    tma_wait_read<0>();
                    __syncwarp();
                }
                double* my = tile + r * cols + q;
                const int n_ch = (cols + kCols4 - 1) / kCols4;
#pragma unroll 1
                for (int c = 0; c < n_ch; ++c, ++chunk) {
                    const int i0 = col0 + c * kCols4;
                    if (chunk != 0 && (chunk % kRestartChunks4) == 0) {
                        beam_lane_restart(b1, i0 + q);
                        beam_lane_restart(b2, i0 + q);
                    }
                    double e1 = b1.amp * b1.ec, e2 = b2.amp * b2.ec;
                    double r1 = b1.rc, r2 = b2.rc;
                    const double2* wp = wsm + i0 + q;
                    double* mine = my + c * kCols4;
                    const int lim = A - i0 - q;                    // step m is a real angle iff kL4*m < lim
                    auto step = [&](int m) {
                        const double2 w = wp[kL4 * m];             // zero beyond A (padding)
                        const double sum = e1 + e2;                // j_beam + j_scat
                        const double j = sum + j_cex;              // plume.py:102
                        den = fma(w.x, sum, den);
                        num = fma(w.y, sum, num);
                        const bool real = kL4 * m < lim;
                        if (CHECKED) bad |= (j <= 0.0) && real;
                        if (STORE_J && real) mine[kL4 * m] = (CHECKED && fill_invalid) ? kInvalidFill : j;
                        e1 *= r1; r1 *= b1.q;
                        e2 *= r2; r2 *= b2.q;
                    };
                    if (lim >= kCols4) {
#pragma unroll
                        for (int m = 0; m < kSteps4; ++m) {
                            const double2 w = wp[kL4 * m];
                            const double sum = e1 + e2;
                            const double j = sum + j_cex;
                            den = fma(w.x, sum, den);
                            num = fma(w.y, sum, num);
                            if (CHECKED) bad |= (j <= 0.0);
                            if (STORE_J) mine[kL4 * m] = (CHECKED && fill_invalid) ? kInvalidFill : j;
                            e1 *= r1; r1 *= b1.q;
                            e2 *= r2; r2 *= b2.q;
                        }
                    } else {
                        const int m_end = min(kSteps4, (A - i0 + kL4 - 1) / kL4);   // warp-uniform
                        for (int m = 0; m < m_end; ++m) step(m);
                    }
                    beam_lane_next(b1);
                    beam_lane_next(b2);
                }
                if (STORE_J) {
                    double* gdst = p.j_ion + grow0 * (long long)A + col0;
                    const uint32_t row_bytes = (uint32_t)cols * 8u;
                    fence_async_smem();
                    __syncwarp();
                    if (p.bulk_ok && bulk_rows_ok && ((rows_valid * row_bytes) & 15u) == 0) {
                        if (lane == 0) bulk_store_1d(gdst, smem_u32(tile), rows_valid * row_bytes);
                    } else if (p.bulk_ok && !bulk_rows_ok && bulk_panel_ok) {
                        if (lane < rows_valid) bulk_store_1d(gdst + lane * (long long)A, smem_u32(tile + lane * cols), row_bytes);
                    } else {   // odd row length with a ragged group or panels: plain coalesced stores
                        for (int rr = 0; rr < rows_valid; ++rr)
                            for (int cc = lane; cc < cols; cc += 32) __stcs(gdst + rr * (long long)A + cc, tile[rr * cols + cc]);
                        __syncwarp();
                    }
                }
            }
        };
        if (__any_sync(0xffffffffu, (flags & 2) != 0))
            sweep(std::true_type{});
        else
            sweep(std::false_type{});

        // quadrature partial sums of the sample's 4 lanes
        num += __shfl_xor_sync(0xffffffffu, num, 1);
        den += __shfl_xor_sync(0xffffffffu, den, 1);
        num += __shfl_xor_sync(0xffffffffu, num, 2);
        den += __shfl_xor_sync(0xffffffffu, den, 2);
        const unsigned badmask = __ballot_sync(0xffffffffu, bad);
        if (q == 0) {
            xr[0] = num;
            xr[1] = den;
            xr[2] = ((badmask >> qbase) & 0xFu) ? 1.0 : 0.0;
        }
    }
    __syncwarp();

    // ---------------- phase 3: one lane per sample ----------------
    const double* xr = xch + lane * kXchPitch;
    const bool bad = xr[2] != 0.0;
    const bool late_fix = STORE_J && __any_sync(0xffffffffu, active && bad && !known_invalid);
    if (STORE_J) {
        if (late_fix) {   // order the bulk (async proxy) writes before the generic-proxy rewrite below
            tma_wait_all();
            fence_async_all();
        } else {
            tma_wait_read<0>();   // the staging tile must outlive the bulk reads
        }
        __syncwarp();
    }
    double cd = xr[0] / xr[1];              // plume.py:124 (NOT masked by `invalid`)
    if (cd == CUDART_INF) cd = CUDART_NAN;  // plume.py:125
    if (active) {
        if (p.div_angle) p.div_angle[s] = acos(cd);
        if (p.cos_div) p.cos_div[s] = cd;
        if (p.t_c) p.t_c[s] = __dmul_rn(thrust, cd);
        if (p.invalid) p.invalid[s] = (known_invalid || bad) ? 1 : 0;
        if (STORE_J && bad && !known_invalid) {  // plume.py:106, rare
            double* row = p.j_ion + s * (long long)A;
            for (int i = 0; i < A; ++i) row[i] = kInvalidFill;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// K1w: several sweep radii, many of them (R >= 8) -- per-sample tables, then ONE contiguous store stream per 8 samples
// ---------------------------------------------------------------------------------------------
// j_ion[s, i, rho] = base(s, rho) * g(s, i) + j_cex(s, rho) with g = A1 E1 + A2 E2 (plume.py:95-102): R outputs cost one
// profile value, so with many radii the kernel is nothing but a store stream -- and a store stream on B200 wants whole
// 32-byte sectors (see kStoreQuad).  The rows of 8 consecutive samples are contiguous in memory and start sector-aligned,
// so a warp builds the two small tables of its 8 samples in shared memory -- g (8 x A; the stride-4 recurrence of K1v on the uniform grid, the reference's own
// divide / square / negate / exp on any other grid) and (base, j_cex) (8 x R) -- and then writes the 8 rows as ONE flat stream: lane l stores
// elements l, l+32, ... (256 contiguous bytes per instruction), whatever the parity of A*R.  Per warp:
//   phase 1  lane t <-> sample s0+t : loads, cathode, A1/A2, divergence angles
//   phase 2  4 groups of 8 samples  : 2a  lane (r, q) = (sample, angle mod 4): g table + the two Simpson sums
//                                     2b  lanes over (sample, radius): CEX terms, cos_div / div_angle / T_c
//                                     2c  lanes over the flat (sample, angle, radius) stream: j_ion
constexpr int kThreadsW = 64;
constexpr int kWarpsW = kThreadsW / 32;
constexpr int kGroupW = 8;
constexpr int kXw = 27;                 // doubles per sample handed from phase 1: amp1 amp2 a1 a2 density sigma I_B0 T flags,
                                        // and (uniform grids) 8 base exps per beam + x1 x2 for the four-lane recurrence of K1v
constexpr int kMinRadiiStream = 8;

__host__ __device__ inline size_t k1w_warp_bytes(int n_angles, int n_radii) {
    return size_t(32) * kXw * 8 + size_t(kGroupW) * n_angles * 8 + size_t(kGroupW) * n_radii * 16 + size_t(kGroupW) * 16;
}
__host__ __device__ inline size_t k1w_smem_bytes(int n_angles, int n_angles_pad, int n_radii) {
    return size_t(n_angles_pad) * 16 + size_t(n_angles) * 8 + size_t(n_radii) * 8 + kWarpsW * k1w_warp_bytes(n_angles, n_radii) + 64;
}

template <bool UNIFORM>
__global__ void __launch_bounds__(kThreadsW) eval_radii_stream_kernel(const EvalParams p) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int A = p.n_angles, R = p.n_radii;
    const long long L = (long long)A * R;
    double2* wsm = reinterpret_cast<double2*>(smem_raw);                 // [n_angles_pad] fused weights
    double* alpha_sm = reinterpret_cast<double*>(wsm + p.n_angles_pad);  // [A]
    double* radii_sm = alpha_sm + A;                                     // [R]
    unsigned char* wbase = reinterpret_cast<unsigned char*>(radii_sm + R);
    wbase += (16 - (smem_u32(wbase) & 15)) & 15;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char* mine = wbase + warp * k1w_warp_bytes(A, R);
    double* xch = reinterpret_cast<double*>(mine);                       // [32][kXw]
    double* gt = xch + 32 * kXw;                                         // [8][A]
    double2* br = reinterpret_cast<double2*>(gt + kGroupW * A);          // [8][R] (base, j_cex)
    double2* sums = br + kGroupW * R;                                    // [8] (num, den)

    for (int i = threadIdx.x; i < p.n_angles_pad; i += kThreadsW) wsm[i] = p.w[i];
    for (int i = threadIdx.x; i < A; i += kThreadsW) alpha_sm[i] = p.alpha[i];
    for (int i = threadIdx.x; i < R; i += kThreadsW) radii_sm[i] = p.radii[i];
    __syncthreads();

    const long long warp_s0 = (long long)blockIdx.x * kThreadsW + warp * 32;
    if (warp_s0 >= p.n) return;
    // ---------------- phase 1: one lane per sample ----------------
    {
        const long long s_raw = warp_s0 + lane;
        const bool active = s_raw < p.n;
        const long long s = active ? s_raw : p.n - 1;
        double x_in[kNumInputs];
#pragma unroll
        for (int q = 0; q < kNumInputs; ++q) {
            const bool needed = (q == IN_P_b) || (q <= IN_P_T ? p.v_cc != nullptr : (q == IN_T ? p.t_c != nullptr : true));
            x_in[q] = needed ? load_in(p, q, s) : 0.0;
        }
        if (p.v_cc) {
            const double v = cathode_vcc(x_in[IN_P_b], x_in[IN_V_a], x_in[IN_T_e], x_in[IN_V_vac], x_in[IN_Pstar],
                                         x_in[IN_P_T], p.torr);
            if (active) p.v_cc[s] = v;
        }
        const SampleConsts k = plume_sample_consts(x_in[IN_P_b], x_in[IN_c0], x_in[IN_c1], x_in[IN_c2], x_in[IN_c3],
                                                   x_in[IN_c4], x_in[IN_c5], p.torr);
        double* xr = xch + lane * kXw;
        xr[0] = k.amp1; xr[1] = k.amp2; xr[2] = k.a1; xr[3] = k.a2; xr[4] = k.density;
        xr[5] = x_in[IN_sigma]; xr[6] = x_in[IN_I_B0]; xr[7] = x_in[IN_T];
        xr[8] = (k.a1 <= 0.0) ? 1.0 : 0.0;          // plume.py:105 first term
        if (UNIFORM) {                               // start values of the stride-4 Gaussian recurrence (see K1v)
            const double t1 = p.h / k.a1, t2 = p.h / k.a2;
            const double x1 = t1 * t1, x2 = t2 * t2;
            double ex[8];
            beam_base_exps(x1, ex);
#pragma unroll
            for (int j = 0; j < 8; ++j) xr[9 + j] = ex[j];
            beam_base_exps(x2, ex);
#pragma unroll
            for (int j = 0; j < 8; ++j) xr[17 + j] = ex[j];
            xr[25] = x1;
            xr[26] = x2;
        }
    }
    __syncwarp();

    // ---------------- phase 2: groups of 8 samples ----------------
#pragma unroll 1
    for (int g = 0; g < 32 / kGroupW; ++g) {
        const long long gs0 = warp_s0 + g * kGroupW;
        if (gs0 >= p.n) break;   // warp-uniform
        const int nvalid = (int)min((long long)kGroupW, p.n - gs0);
        const double* xg = xch + g * kGroupW * kXw;
        // 2a: g(s, i) = A1 exp(-(alpha_i/a1)^2) + A2 exp(-(alpha_i/a2)^2) in the reference's operation order (plume.py:99-100
        // without the radius-dependent factor), and the two fused Simpson sums of plume.py:117-123
        {
            const int r = lane >> 2, q = lane & 3;
            const double* xr = xg + r * kXw;
            const double amp1 = xr[0], amp2 = xr[1], a1 = xr[2], a2 = xr[3];
            double num = 0.0, den = 0.0;
            if (UNIFORM) {       // lane q sweeps angles i = q (mod 4) with the stride-4 recurrence of K1v
                BeamLane b1, b2;
                b1.amp = amp1; b2.amp = amp2;
                b1.x = xr[25]; b2.x = xr[26];
                beam_lane_init(b1, xr + 9, q);
                beam_lane_init(b2, xr + 17, q);
                const int n_ch = (A + kCols4 - 1) / kCols4;
                for (int c = 0; c < n_ch; ++c) {
                    const int i0 = c * kCols4;
                    if (c != 0 && (c % kRestartChunks4) == 0) {
                        beam_lane_restart(b1, i0 + q);
                        beam_lane_restart(b2, i0 + q);
                    }
                    double e1 = b1.amp * b1.ec, e2 = b2.amp * b2.ec, r1 = b1.rc, r2 = b2.rc;
#pragma unroll
                    for (int m = 0; m < kSteps4; ++m) {
                        const int i = i0 + q + kL4 * m;
                        const double v = e1 + e2;
                        const double2 w = wsm[i];            // zero beyond A (padding to a multiple of 64)
                        den = fma(w.x, v, den);
                        num = fma(w.y, v, num);
                        if (i < A) gt[r * A + i] = v;
                        e1 *= r1; r1 *= b1.q;
                        e2 *= r2; r2 *= b2.q;
                    }
                    beam_lane_next(b1);
                    beam_lane_next(b2);
                }
            } else {
                for (int i = q; i < A; i += 4) {
                    const double al = alpha_sm[i];
                    const double t1 = al / a1, t2 = al / a2;
                    const double v = __dadd_rn(__dmul_rn(amp1, exp(-(t1 * t1))), __dmul_rn(amp2, exp(-(t2 * t2))));
                    gt[r * A + i] = v;
                    const double2 w = wsm[i];
                    den = fma(w.x, v, den);
                    num = fma(w.y, v, num);
                }
            }
            num += __shfl_xor_sync(0xffffffffu, num, 1);
            den += __shfl_xor_sync(0xffffffffu, den, 1);
            num += __shfl_xor_sync(0xffffffffu, num, 2);
            den += __shfl_xor_sync(0xffffffffu, den, 2);
            if (q == 0) sums[r] = make_double2(num, den);
        }
        __syncwarp();
        // 2b: per (sample, radius): decay, j_cex, base (plume.py:95-98); cos_div, div_angle, T_c (plume.py:122-127,137)
        bool needs_check = false;
        for (int pr = lane; pr < nvalid * R; pr += 32) {
            const int smp = pr / R, rho = pr - smp * R;
            const double* xr = xg + smp * kXw;
            double j_cex, base;
            cex_terms(xr[4], xr[5], xr[6], radii_sm[rho], j_cex, base);
            br[smp * R + rho] = make_double2(base, j_cex);
            // non-negative amplitudes and base, positive CEX floor => every j_ion > 0 (or NaN): no per-element test needed
            needs_check |= !(xr[0] >= 0.0 && xr[1] >= 0.0 && base >= 0.0 && j_cex > 0.0) || xr[8] != 0.0;
            const double2 sm = sums[smp];
            double cd = __dmul_rn(base, sm.x) / __dmul_rn(base, sm.y);
            if (cd == CUDART_INF) cd = CUDART_NAN;
            const long long o = (gs0 + smp) * R + rho;
            if (p.div_angle) p.div_angle[o] = acos(cd);
            if (p.cos_div) p.cos_div[o] = cd;
            if (p.t_c) p.t_c[o] = __dmul_rn(xr[7], cd);
        }
        needs_check = __any_sync(0xffffffffu, needs_check);
        __syncwarp();
        // 2c: the 8 rows as one flat stream of nvalid * A * R elements.  Element e = (q, rho) with q = e / R the flat
        // (sample, angle) index -- exactly the index into the g table -- advanced incrementally: e += 32.
        unsigned badbits = 0, knownbits = 0;
        for (int smp = 0; smp < nvalid; ++smp) knownbits |= (xg[smp * kXw + 8] != 0.0) ? (1u << smp) : 0u;
        const int total = nvalid * A * R;              // < 2^29 (checked on the host)
        const int dq = 32 / R, drho = 32 - dq * R;     // e += 32  <=>  q += dq, rho += drho (+ one carry)
        if (p.j_ion && !needs_check) {
            double* dst = p.j_ion + gs0 * L;
            int q = lane / R, rho = lane - q * R;
            int smp_end = A, br_base = 0;               // first q of the next sample, offset of the current sample's (base, j_cex) row
#pragma unroll 4
            for (int e = lane; e < total; e += 32) {
                while (q >= smp_end) { smp_end += A; br_base += R; }
                const double2 bj = br[br_base + rho];
                __stcs(dst + e, fma(bj.x, gt[q], bj.y));   // base * (A1 E1 + A2 E2) + j_cex, plume.py:99-102
                q += dq;
                rho += drho;
                if (rho >= R) { rho -= R; ++q; }
            }
        } else {                                        // rare: rows that may hold a non-positive j_ion, or alpha1 <= 0
            double* dst = p.j_ion ? p.j_ion + gs0 * L : nullptr;
            for (int e = lane; e < total; e += 32) {
                const int q = e / R, rho = e - q * R, smp = q / A;
                const double2 bj = br[smp * R + rho];
                const double j = fma(bj.x, gt[q], bj.y);
                if (j <= 0.0) badbits |= 1u << smp;
                if (dst) __stcs(dst + e, ((knownbits >> smp) & 1u) ? kInvalidFill : j);
            }
        }
        badbits = __reduce_or_sync(0xffffffffu, badbits);
        __syncwarp();
        const unsigned late = badbits & ~knownbits;    // plume.py:106 for rows whose non-positive j_ion was found on the way
        if (p.j_ion && late) {
            for (int smp = 0; smp < nvalid; ++smp) {
                if (!((late >> smp) & 1u)) continue;
                double* row = p.j_ion + (gs0 + smp) * L;
                for (long long e = lane; e < L; e += 32) row[e] = kInvalidFill;
            }
        }
        if (p.invalid && lane < nvalid) p.invalid[gs0 + lane] = (((badbits | knownbits) >> lane) & 1u) ? 1 : 0;
        __syncwarp();
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

// ---------------------------------------------------------------------------------------------
// K3: j_ion interpolated to measurement angles + Gaussian log-likelihood, without materialising j_ion
// ---------------------------------------------------------------------------------------------
// The step right after the plume model in the reference's calibration scripts: mirror the sweep to (-90, 90) deg,
// `interp1d` (linear) to the probe angles (scripts/pem_v0/monte_carlo.py:265-270; the same intent is sketched in
// plume.py:142-149), then sum -0.5 ((y - y_hat)/sigma)^2 (scripts/pem_v0/mcmc.py:103).  One thread per sample runs the
// recurrence sweep once; the measurement points are pre-sorted by |angle| on the host so that the points falling in
// grid interval [i-1, i] are a contiguous range that is consumed when the sweep reaches angle i.
struct __align__(16) MeasPoint {   // one probe location, sorted by |theta|; 32 bytes = two broadcast 16-byte loads
    double w;            // (|theta| - alpha[lo]) / (alpha[lo+1] - alpha[lo]), [alpha[lo], alpha[lo+1]] = the grid interval that holds |theta|
    double y;            // measured j_ion
    double inv_sigma;    // 1 / standard deviation
    int orig;            // index in the caller's arrays
    int rel;             // (lo + 1) mod 16: position of alpha[lo] in the 17-angle window of the chunk that consumes the point
};
struct LoglikeParams {
    int m;                    // number of measurement points
    const int* seg_start;     // [A]: points of interval [i, i+1] are seg_start[i] .. seg_start[i+1]-1 (seg_start[A-1] = m)
    const MeasPoint* meas;    // [m] sorted
    double* loglike;          // (n)   or nullptr
    double* y_pred;           // (n, m) in the caller's point order, or nullptr
};
constexpr int kThreadsL = 128;
constexpr int kRowL = kChunk + 1;   // per-thread window: j of the previous chunk's last angle + the 16 angles of this chunk

// One thread per sample.  The sweep runs in fully unrolled 16-angle chunks exactly like K1u (no quadrature: 6 fp64
// instructions per angle) and parks the chunk in a private shared-memory window; the probe points whose interval lies in
// the window -- a contiguous range of the sorted list, the same for every thread -- are then interpolated from it.
__global__ void __launch_bounds__(kThreadsL) loglike_kernel(const EvalParams p, const LoglikeParams lp) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    MeasPoint* msm = reinterpret_cast<MeasPoint*>(smem_raw);                       // [m]
    int* seg = reinterpret_cast<int*>(msm + lp.m);                                 // [A]
    double* win_all = reinterpret_cast<double*>(smem_raw + ((sizeof(MeasPoint) * lp.m + sizeof(int) * p.n_angles + 15) & ~size_t(15)));
    const int A = p.n_angles;
    for (int i = threadIdx.x; i < lp.m; i += kThreadsL) msm[i] = lp.meas[i];
    for (int i = threadIdx.x; i < A; i += kThreadsL) seg[i] = lp.seg_start[i];
    __syncthreads();
    const long long s = (long long)blockIdx.x * kThreadsL + threadIdx.x;
    if (s >= p.n) return;
    double* win = win_all + threadIdx.x * kRowL;          // odd pitch (17 doubles): conflict-free across the warp

    double x_in[kNumInputs];
#pragma unroll
    for (int q = 0; q < kNumInputs; ++q) x_in[q] = (q == IN_P_b || (q > IN_P_T && q != IN_T)) ? load_in(p, q, s) : 0.0;
    SampleConsts k;
    double j_cex, base;
    BeamState b1, b2;
    plume_prologue(p, x_in, k, j_cex, base, b1, b2);
    const int n_chunks = (A + kChunk - 1) / kChunk;

    // plume.py:105-106: invalid samples return j_ion == 1e-20 at every angle -- that is what gets interpolated
    bool invalid = (k.a1 <= 0.0);
    if (!invalid && !(b1.amp >= 0.0 && b2.amp >= 0.0 && j_cex > 0.0)) {   // rare: look ahead for a non-positive j_ion
        BeamState t1 = b1, t2 = b2;
        for (int c = 0; c < n_chunks; ++c) {
            const int i0 = c * kChunk;
            if (c != 0 && (c % kRestartChunks) == 0) {
                beam_restart(t1, i0);
                beam_restart(t2, i0);
            }
            double e1 = t1.amp * t1.ec, e2 = t2.amp * t2.ec, r1 = t1.rc, r2 = t2.rc;
            for (int kk = 0; kk < kChunk && i0 + kk < A; ++kk) {
                invalid |= ((e1 + e2) + j_cex <= 0.0);
                e1 *= r1; r1 *= t1.q;
                e2 *= r2; r2 *= t2.q;
            }
            beam_next_chunk(t1);
            beam_next_chunk(t2);
        }
    }
    if (invalid) {          // constant row whatever else is NaN: zero amplitudes on a neutral recurrence, the fill as the floor
        b1.amp = b2.amp = 0.0;
        b1.ec = b1.rc = b1.gc = b1.q = b1.qk = b1.hh = 1.0;
        b2.ec = b2.rc = b2.gc = b2.q = b2.qk = b2.hh = 1.0;
        b1.x = b2.x = 0.0;
        j_cex = kInvalidFill;
    }

    double ll = 0.0;
    double* pred = lp.y_pred ? lp.y_pred + s * (long long)lp.m : nullptr;
    for (int c = 0; c < n_chunks; ++c) {
        const int i0 = c * kChunk;
        if (c != 0 && (c % kRestartChunks) == 0) {
            beam_restart(b1, i0);
            beam_restart(b2, i0);
        }
        double e1 = b1.amp * b1.ec, e2 = b2.amp * b2.ec, r1 = b1.rc, r2 = b2.rc;
        if (c != 0) win[0] = win[kChunk];                  // j of angle i0 - 1
#pragma unroll
        for (int kk = 0; kk < kChunk; ++kk) {              // angles >= A of the last chunk are computed but never read
            win[kk + 1] = (e1 + e2) + j_cex;
            e1 *= r1; r1 *= b1.q;
            e2 *= r2; r2 *= b2.q;
        }
        beam_next_chunk(b1);
        beam_next_chunk(b2);
        // intervals [lo, lo+1] with both ends in the window: lo = max(i0 - 1, 0) .. min(i0 + 15, A - 1) - 1
        const int lo_first = max(i0 - 1, 0), hi_last = min(i0 + kChunk - 1, A - 1);
        const int q0 = seg[lo_first], q1 = seg[hi_last];   // warp-uniform range of the sorted points
        if (pred) {
            for (int q = q0; q < q1; ++q) {
                const MeasPoint mp = msm[q];                    // broadcast loads
                const double ja = win[mp.rel], jb = win[mp.rel + 1];
                const double yh = fma(mp.w, jb - ja, ja);       // linear interpolation on [alpha[lo], alpha[lo+1]]
                const double r = (mp.y - yh) * mp.inv_sigma;
                ll = fma(r, r, ll);                             // sum of squared residuals; the factor -1/2 (exact) is applied once below
                pred[mp.orig] = yh;
            }
        } else {
#pragma unroll 4
            for (int q = q0; q < q1; ++q) {
                const MeasPoint mp = msm[q];
                const double ja = win[mp.rel], jb = win[mp.rel + 1];
                const double yh = fma(mp.w, jb - ja, ja);
                const double r = (mp.y - yh) * mp.inv_sigma;
                ll = fma(r, r, ll);
            }
        }
    }
    if (lp.loglike) lp.loglike[s] = -0.5 * ll;
}

// log-sum-exp over the trailing axis (the M Monte-Carlo draws of the nuisance parameters behind one calibration vector):
// out[g] = max_m ll[g, m] + log(sum_m exp(ll[g, m] - max))   -- scripts/pem_v0/mcmc.py:101-102.  One warp per group.
__global__ void __launch_bounds__(128) logsumexp_kernel(const double* __restrict__ ll, long long n_groups, int m,
                                                        double* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const long long g = (long long)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (g >= n_groups) return;
    const double* row = ll + g * (long long)m;
    double mx = -CUDART_INF;
    bool any_nan = false;
    for (int i = lane; i < m; i += 32) {
        const double v = __ldg(row + i);
        any_nan |= (v != v);
        mx = fmax(mx, v);                       // fmax skips NaN; np.max propagates it -> handled below
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    any_nan = __any_sync(0xffffffffu, any_nan);
    double sum = 0.0;
    for (int i = lane; i < m; i += 32) sum += exp(__ldg(row + i) - mx);
    sum = warp_sum(sum);
    if (lane == 0) out[g] = any_nan ? CUDART_NAN : mx + log(sum);   // all -inf: -inf + log(nan) = nan, as in NumPy
}

}  // namespace hpem
