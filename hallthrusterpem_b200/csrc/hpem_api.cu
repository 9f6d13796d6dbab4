// hpem_api.cu -- C ABI of libhpem (see include/hpem.h).  Host side: argument checking, grid handle,
// kernel dispatch, and the host-buffer pipeline (H2D -> kernel -> D2H, chunked over samples).
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <new>
#include <vector>

#include "../../include/hpem.h"
#include "hpem_kernels.cuh"
#include "hpem_moments.cuh"
#include "hpem_compress.cuh"

namespace {

thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define HPEM_CUDA(call)                                                                            \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(HPEM_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

struct DeviceGuard {
    int prev = -1;
    bool ok = false;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) == cudaSuccess && cudaSetDevice(dev) == cudaSuccess) ok = true;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

// Which inputs each output group reads (cathode.py:26-31, plume.py:40-49)
constexpr int kCathodeInputs[] = {HPEM_IN_P_b, HPEM_IN_V_a, HPEM_IN_T_e, HPEM_IN_V_vac, HPEM_IN_Pstar, HPEM_IN_P_T};
constexpr int kPlumeInputs[] = {HPEM_IN_P_b, HPEM_IN_c0, HPEM_IN_c1, HPEM_IN_c2, HPEM_IN_c3,
                                HPEM_IN_c4,  HPEM_IN_c5, HPEM_IN_sigma_cex, HPEM_IN_I_B0};

struct Workspace {
    std::mutex mu;
    double* d_in[HPEM_N_INPUTS] = {};
    size_t in_cap = 0;  // samples
    double* d_small[4] = {};  // V_cc, div_angle, T_c, cos_div
    size_t small_cap = 0;     // elements (n * R)
    uint8_t* d_invalid = nullptr;
    size_t invalid_cap = 0;
    double* d_j = nullptr;
    size_t j_cap = 0;  // elements
    cudaStream_t s_compute = nullptr, s_copy = nullptr, s_h2d = nullptr;
    std::vector<cudaEvent_t> events, h2d_events;
    // small-batch path of hpem_eval_host: ONE packed H2D and ONE packed D2H through pinned staging buffers
    unsigned char* h_stage_in = nullptr;
    unsigned char* h_stage_out = nullptr;
    unsigned char* d_pack_in = nullptr;
    unsigned char* d_pack_out = nullptr;
    double* d_partials = nullptr;  // K2 per-block partial vectors
    size_t partials_cap = 0;
    double* d_partial_minmax = nullptr;
    size_t partial_minmax_cap = 0;
    unsigned* d_hist_partials = nullptr;   // K2 per-block histograms; all-zero between calls (the finalize kernel re-zeroes them)
    size_t hist_partials_cap = 0;
    cudaEvent_t moments_done = nullptr;    // end of the last reduce-only pass that used the scratch buffers above
};

}  // namespace

struct hpem_grid {
    int device = 0;
    int n_angles = 0, n_angles_pad = 0, n_radii = 0;
    bool uniform = false;
    double h = 0.0, radius0 = 1.0;
    double2* d_w = nullptr;
    double* d_alpha = nullptr;
    double* d_radii = nullptr;
    std::vector<double> alpha_host;
    std::vector<double2> w_host;   // fused weights (wd_i, wn_i), zero-padded
    double* d_qtable = nullptr;    // uniform grids: tabulated Simpson sums N_d(x), N_n(x) (hpem_qtable.cuh)
    int qt_key_lo = 0, qt_bins = 0;
    size_t smem_tma = 0, smem_tma32 = 0, smem_tma1 = 0, smem_quad = 0, smem_quad1 = 0, smem_stg = 0, smem_nostore = 0, smem_rows = 0;  // dynamic shared memory of the K1u variants
    size_t smem_v_store = 0, smem_v_nostore = 0;          // ... and of K1v
    int sm_count = 148;
    bool smem_ok = false;
    Workspace ws;
};

namespace {

// Opt a kernel in to large dynamic shared memory.  The attribute is per function (not per launch), so it is raised to
// the device's opt-in maximum once instead of being set to each grid's own requirement (a later, smaller grid must not
// lower the limit of a cached larger one).
template <typename K>
int set_smem(K kernel, size_t bytes) {
    int dev = 0, max_optin = 0;
    HPEM_CUDA(cudaGetDevice(&dev));
    HPEM_CUDA(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    cudaFuncAttributes attr;
    HPEM_CUDA(cudaFuncGetAttributes(&attr, kernel));
    const int max_dynamic = max_optin - (int)attr.sharedSizeBytes;   // static __shared__ counts against the same limit
    if (bytes > (size_t)max_dynamic)
        return fail(HPEM_ERR_UNSUPPORTED, "kernel needs %zu bytes of shared memory, device allows %d", bytes, max_dynamic);
    HPEM_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_dynamic));
    return HPEM_OK;
}

bool wants_plume(const hpem_outputs& o) {
    return o.j_ion || o.div_angle || o.T_c || o.cos_div || o.invalid;
}

// Fill EvalParams for samples [first, first+count) of device-resident inputs/outputs.
void fill_params(const hpem_grid& g, const hpem_inputs& in, const hpem_outputs& out, int64_t first, int64_t count,
                 double torr, hpem::EvalParams& p) {
    for (int k = 0; k < HPEM_N_INPUTS; ++k) {
        p.in[k] = in.ptr[k] ? in.ptr[k] + first : nullptr;
        p.scalar[k] = in.scalar[k];
    }
    p.n = count;
    p.torr = torr;
    const int64_t R = g.n_radii;
    p.v_cc = out.V_cc ? out.V_cc + first : nullptr;
    p.j_ion = out.j_ion ? out.j_ion + first * g.n_angles * R : nullptr;
    p.div_angle = out.div_angle ? out.div_angle + first * R : nullptr;
    p.t_c = out.T_c ? out.T_c + first * R : nullptr;
    p.cos_div = out.cos_div ? out.cos_div + first * R : nullptr;
    p.invalid = out.invalid ? out.invalid + first : nullptr;
    p.n_angles = g.n_angles;
    p.n_angles_pad = g.n_angles_pad;
    p.n_radii = g.n_radii;
    p.w = g.d_w;
    p.alpha = g.d_alpha;
    p.radii = g.d_radii;
    p.h = g.h;
    p.radius0 = g.radius0;
    p.has_thrust = out.T_c != nullptr;
    p.bulk_ok = false;
    // rows that own their 128-byte lines may leave L2 early; rows that share lines with their neighbours should stay
    static const int hint_env = []() { const char* v = std::getenv("HPEM_L2_HINT"); return v ? std::atoi(v) : -1; }();
    p.l2_hint = hint_env >= 0 ? hint_env : (((long long)g.n_angles * g.n_radii) % 16 == 0 ? 2 : 1);
    static const int no_fast_env = []() { const char* v = std::getenv("HPEM_NO_FASTMATH"); return v ? std::atoi(v) : 0; }();
    p.no_fastmath = no_fast_env;
    static const int no_qtable_env = []() { const char* v = std::getenv("HPEM_NO_QTABLE"); return v ? std::atoi(v) : 0; }();
    p.qt.rows = no_qtable_env ? nullptr : g.d_qtable;
    p.qt.key_lo = g.qt_key_lo;
    p.qt.n_bins = g.qt_bins;
    p.qt.wd0 = g.w_host.empty() ? 0.0 : g.w_host[0].x;
    p.qt.wn0 = g.w_host.empty() ? 0.0 : g.w_host[0].y;
}

// Host copies of the quadrature tables, one per distinct weight vector: several devices / radii sets share one build.
struct QTableHost {
    std::vector<double> wd, wn, rows;
    int key_lo = 0, n_bins = 0;
};
std::shared_ptr<const QTableHost> qtable_for(int n_angles, const double* wd, const double* wn) {
    static std::mutex mu;
    static std::vector<std::shared_ptr<const QTableHost>> cache;
    std::lock_guard<std::mutex> lock(mu);
    for (const auto& t : cache)
        if ((int)t->wd.size() == n_angles && std::memcmp(t->wd.data(), wd, n_angles * sizeof(double)) == 0 &&
            std::memcmp(t->wn.data(), wn, n_angles * sizeof(double)) == 0)
            return t;
    auto t = std::make_shared<QTableHost>();
    t->wd.assign(wd, wd + n_angles);
    t->wn.assign(wn, wn + n_angles);
    hpem::qtable_build(n_angles, wd, wn, t->rows, t->key_lo, t->n_bins);
    if (cache.size() >= 32) cache.erase(cache.begin());
    cache.push_back(t);
    return t;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point query (no link-time dependency on libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = []() -> EncodeTiledFn {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            return nullptr;
        return reinterpret_cast<EncodeTiledFn>(ptr);
    }();
    return fn;
}

// Tensor map of j_ion viewed as (rows, cols): `n_cols` columns per row, rows `row_pitch` elements apart;
// box = box_rows x box_cols.  (n, A) view: n_cols = row_pitch = A.  Quad-row view (A % 4 != 0): see kStoreQuad.
int make_j_map(double* base, long long n_cols, long long n_rows, long long row_pitch, int box_cols, int box_rows, bool swizzle128,
               CUtensorMap* map) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return fail(HPEM_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    const cuuint64_t dims[2] = {(cuuint64_t)n_cols, (cuuint64_t)n_rows};
    const cuuint64_t strides[1] = {(cuuint64_t)row_pitch * sizeof(double)};
    const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1u, 1u};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(HPEM_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return HPEM_OK;
}

// 3-D view for K1u: (16 angles inside a 128-byte column block, rows, column blocks); one box = kTmaCB column blocks x
// box_rows rows x 16 angles, read from [column block][row][128 B] sub-tiles (128B swizzle).
int make_j_map3(double* base, long long n_cols, long long n_rows, long long row_pitch, int box_rows, CUtensorMap* map) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return fail(HPEM_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    const long long n_blocks = n_cols / hpem::kChunk;
    if (n_blocks < hpem::kTmaCB) {   // never used by the kernel (every group takes the 2-D path)
        std::memset(map, 0, sizeof(*map));
        return HPEM_OK;
    }
    // dimension order (angle within block, row, column block): the box is laid out in shared memory with the FIRST
    // dimension fastest, i.e. [column block][row][16 angles] -- one 128B-swizzled sub-tile per column block
    const cuuint64_t dims[3] = {(cuuint64_t)hpem::kChunk, (cuuint64_t)n_rows, (cuuint64_t)n_blocks};
    const cuuint64_t strides[2] = {(cuuint64_t)row_pitch * sizeof(double), (cuuint64_t)hpem::kChunk * sizeof(double)};
    const cuuint32_t box[3] = {(cuuint32_t)hpem::kChunk, (cuuint32_t)box_rows, (cuuint32_t)hpem::kTmaCB};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(HPEM_ERR_CUDA, "cuTensorMapEncodeTiled (3-D) failed with CUresult %d", (int)r);
    return HPEM_OK;
}

// Row phases of the quad-row store mode (see kStoreQuad): for a 32-byte aligned base, row r starts (r A 8) mod 32 bytes
// into a sector; lead = elements up to the next sector boundary, body = common sector-aligned length, tail = the rest.
struct QuadLayout {
    int lead[4], tail[4], body;
};
QuadLayout quad_layout(int n_angles) {
    QuadLayout q;
    q.body = 1 << 30;
    for (int f = 0; f < 4; ++f) {
        q.lead[f] = (4 - (int)(((long long)f * n_angles) % 4)) % 4;
        q.body = std::min(q.body, (n_angles - q.lead[f]) & ~3);
    }
    q.body = std::max(q.body, 0);
    for (int f = 0; f < 4; ++f) q.tail[f] = n_angles - q.lead[f] - q.body;
    return q;
}

int launch(const hpem_grid& g, const hpem::EvalParams& p, bool plume, bool store_j, uint32_t flags, cudaStream_t st) {
    using namespace hpem;
    if (p.n <= 0) return HPEM_OK;
    const bool use_uniform = g.uniform && g.n_radii == 1 && g.smem_ok && !(flags & HPEM_FLAG_FORCE_DIRECT);
    const bool use_multi = plume && store_j && g.uniform && g.n_radii > 1 && g.n_radii <= kMaxRadiiFast && g.smem_ok &&
                           !(flags & HPEM_FLAG_FORCE_DIRECT);
    const size_t smem_w = k1w_smem_bytes(g.n_angles, g.n_angles_pad, g.n_radii);
    // K1w for many radii, and for a few radii whenever the row length A*R is odd (K1r has no tensor map there and falls back
    // to plain stores: 0.104 vs 0.151 ms at 91 angles x 7 radii); K1r keeps the short even rows (A = 200, R = 3: 0.098 vs 0.130 ms)
    const bool odd_rows = ((long long)g.n_angles * g.n_radii) % 2 == 1;
    const bool use_stream = plume && g.n_radii >= 2 && (g.n_radii >= kMinRadiiStream || odd_rows) && smem_w <= 100 * 1024 && (long long)g.n_angles * g.n_radii < (1LL << 25) &&
                            !(flags & (HPEM_FLAG_FORCE_DIRECT | HPEM_FLAG_LANES1));
    if (use_stream) {   // K1w: many radii -- per-sample tables + one contiguous store stream per 8 samples (any grid)
        const unsigned blocks = (unsigned)((p.n + kThreadsW - 1) / kThreadsW);
        if (g.uniform) {
            int rc = set_smem(eval_radii_stream_kernel<true>, smem_w);
            if (rc != HPEM_OK) return rc;
            eval_radii_stream_kernel<true><<<blocks, kThreadsW, smem_w, st>>>(p);
        } else {
            int rc = set_smem(eval_radii_stream_kernel<false>, smem_w);
            if (rc != HPEM_OK) return rc;
            eval_radii_stream_kernel<false><<<blocks, kThreadsW, smem_w, st>>>(p);
        }
    } else if (use_multi) {   // K1r: a few radii through the recurrence sweep
        const unsigned blocks = (unsigned)((p.n + kThreadsU - 1) / kThreadsU);
        const long long row_len = (long long)g.n_angles * g.n_radii;
        const size_t rad_bytes = size_t(2) * g.n_radii * kThreadsU * sizeof(double);
        const size_t wbytes = size_t(g.n_angles_pad) * sizeof(double2) + 1024;
        const bool tma_ok = (row_len % 2 == 0) && ((reinterpret_cast<uintptr_t>(p.j_ion) & 15u) == 0) &&
                            !(flags & HPEM_FLAG_NO_TMA) && row_len < 2147483647LL;
        CUtensorMap map;
        std::memset(&map, 0, sizeof(map));
        if (tma_ok) {
            int rc = make_j_map(p.j_ion, row_len, p.n, row_len, kChunk, 32, true, &map);
            if (rc != HPEM_OK) return rc;
            const size_t smem = wbytes + rad_bytes + size_t(kWarpsU) * kTmaBuffers * kTmaTileBytes;
            rc = set_smem(eval_multi_radius_kernel<true>, smem);
            if (rc != HPEM_OK) return rc;
            eval_multi_radius_kernel<true><<<blocks, kThreadsU, smem, st>>>(p, map);
        } else {
            const size_t smem = wbytes + rad_bytes + size_t(kWarpsU) * 32 * kTilePitch * sizeof(double);
            int rc = set_smem(eval_multi_radius_kernel<false>, smem);
            if (rc != HPEM_OK) return rc;
            eval_multi_radius_kernel<false><<<blocks, kThreadsU, smem, st>>>(p, map);
        }
    } else if (use_uniform || !plume) {
        const unsigned blocks = (unsigned)((p.n + kThreadsU - 1) / kThreadsU);
        JMaps maps;
        std::memset(&maps, 0, sizeof(maps));
        const int A = g.n_angles;
        // TMA tensor stores over the (n, A) view need sector-aligned rows (A % 4 == 0; 16-byte alignment alone -- A % 2 == 0
        // -- works but writes partial sectors: 4.4 instead of 5.6 TB/s) and an aligned base ...
        const bool tma_ok = store_j && (A % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.j_ion) & 15u) == 0) &&
                            !(flags & HPEM_FLAG_NO_TMA);
        // ... every other angle count (the reference's 91 included) takes the quad-row view: four rows span whole sectors.
        // Needs a sample count that is a multiple of 4 (launch_range() peels the remainder off) and a 32-byte aligned base.
        QuadLayout ql = quad_layout(A);
        const bool rows_fit = (A % 2 == 1) && size_t(32) * A * 8 <= 16 * 1024;   // small odd A: whole-row mode (0.123 vs 0.142 ms at A = 51)
        // (below 64 angles the lead/tail elements of the quad-row mode cost more than the partial sectors they avoid)
        const bool quad_ok = store_j && (A % 4 != 0) && A >= kQuadMinAngles && ql.body >= kChunk && ((reinterpret_cast<uintptr_t>(p.j_ion) & 31u) == 0) &&
                             (p.n % 4 == 0) && !(flags & (HPEM_FLAG_NO_TMA | HPEM_FLAG_NO_QUAD)) && 4LL * A < 2147483647LL;
        // fallbacks without tensor stores: even A through 16-byte aligned (n, A) boxes, small odd A through K1u's whole-row
        // mode (32 x A tile <= 16 KB), larger odd A through the four-lane sweep with whole-row bulk stores (K1v)
        const bool tma16_ok = store_j && (A % 2 == 0) && ((reinterpret_cast<uintptr_t>(p.j_ion) & 15u) == 0) && !(flags & HPEM_FLAG_NO_TMA);
        const bool lanes1 = (flags & HPEM_FLAG_LANES1) ? true : (flags & HPEM_FLAG_LANES4) ? false
                                                                 : (!store_j || A % 2 == 0 || rows_fit || quad_ok);
        // one staging buffer per warp (twice the resident warps) while the per-sample prologue dominates: measured cross-over
        // at ~160 angles for the (n, A) boxes and ~200 for the quad-row mode (tools/variant_angles.py)
        const bool one_buf = A <= (quad_ok ? kOneBufferMaxAnglesQuad : kOneBufferMaxAngles);
        if (!plume) {
            eval_uniform_kernel<false, false, kStoreStg, 2><<<blocks, kThreadsU, 0, st>>>(p, maps);
        } else if (lanes1) {   // K1u: one lane per sample for the angle sweep as well
            if (!store_j) {
                eval_uniform_kernel<true, false, kStoreStg, 2><<<blocks, kThreadsU, g.smem_nostore, st>>>(p, maps);
            } else if (quad_ok) {
                hpem::EvalParams pq = p;
                pq.q_body = ql.body;
                for (int f = 0; f < 4; ++f) {
                    pq.q_lead[f] = ql.lead[f];
                    pq.q_tail[f] = ql.tail[f];
                    double* base = p.j_ion + (long long)f * A + ql.lead[f];    // first body element of phase f: 32-byte aligned
                    int rc = make_j_map(base, ql.body, p.n / 4, 4LL * A, kChunk, 8, true, &maps.m2[f]);
                    if (rc == HPEM_OK) rc = make_j_map3(base, ql.body, p.n / 4, 4LL * A, 8, &maps.m3[f]);
                    if (rc != HPEM_OK) return rc;
                }
                if (one_buf)
                    eval_uniform_kernel<true, true, kStoreQuad, 1><<<blocks, kThreadsU, g.smem_quad1, st>>>(pq, maps);
                else
                    eval_uniform_kernel<true, true, kStoreQuad, 2><<<blocks, kThreadsU, g.smem_quad, st>>>(pq, maps);
            } else if (tma_ok || tma16_ok) {
                int rc = make_j_map(p.j_ion, A, p.n, A, kChunk, 32, true, &maps.m2[0]);
                if (rc == HPEM_OK) rc = make_j_map3(p.j_ion, A, p.n, A, 32, &maps.m3[0]);
                if (rc != HPEM_OK) return rc;
                if (one_buf) {
                    eval_uniform_kernel<true, true, kStoreTma, 1><<<blocks, kThreadsU, g.smem_tma1, st>>>(p, maps);
                } else {   // long aligned rows: one-warp blocks (0.311 -> 0.306 ms at 1e6 x 200)
                    const unsigned blocks32 = (unsigned)((p.n + 31) / 32);
                    eval_uniform_kernel<true, true, kStoreTma, 2, 32><<<blocks32, 32, g.smem_tma32, st>>>(p, maps);
                }
            } else if (rows_fit && !(flags & HPEM_FLAG_NO_TMA)) {   // small odd A: whole rows, one 1-D bulk store per warp
                hpem::EvalParams pr = p;
                pr.bulk_ok = (reinterpret_cast<uintptr_t>(p.j_ion) & 15u) == 0;
                eval_uniform_kernel<true, true, kStoreRows, 2><<<blocks, kThreadsU, g.smem_rows, st>>>(pr, maps);
            } else {
                eval_uniform_kernel<true, true, kStoreStg, 2><<<blocks, kThreadsU, g.smem_stg, st>>>(p, maps);
            }
        } else {               // K1v: four lanes per sample in the sweep, whole rows per bulk store
            const unsigned vblocks = (unsigned)((p.n + kThreadsV - 1) / kThreadsV);
            hpem::EvalParams pv = p;
            pv.bulk_ok = store_j && ((reinterpret_cast<uintptr_t>(p.j_ion) & 15u) == 0) && !(flags & HPEM_FLAG_NO_TMA);
            if (!store_j)
                eval_lanes4_kernel<false><<<vblocks, kThreadsV, g.smem_v_nostore, st>>>(pv);
            else
                eval_lanes4_kernel<true><<<vblocks, kThreadsV, g.smem_v_store, st>>>(pv);
        }
    } else {
        const unsigned blocks = (unsigned)((p.n + kWarpsD - 1) / kWarpsD);
        eval_direct_kernel<true><<<blocks, kThreadsD, 0, st>>>(p);
    }
    g_launches.fetch_add(1, std::memory_order_relaxed);
    HPEM_CUDA(cudaGetLastError());
    return HPEM_OK;
}

// fill_params + launch for samples [first, first + count).  The quad-row store mode needs a sample count that is a
// multiple of 4: other ranges are split into that part and the last 1-3 samples (which take a fallback store mode).
int launch_range(const hpem_grid& g, const hpem_inputs& in, const hpem_outputs& out, int64_t first, int64_t count, double torr,
                 bool plume, uint32_t flags, cudaStream_t st) {
    const bool store_j = out.j_ion != nullptr;
    const bool quad_candidate = plume && store_j && g.uniform && g.n_radii == 1 && g.smem_ok && (g.n_angles % 4 != 0) &&
                                quad_layout(g.n_angles).body >= hpem::kChunk &&
                                g.n_angles >= hpem::kQuadMinAngles &&
                                !(flags & (HPEM_FLAG_FORCE_DIRECT | HPEM_FLAG_NO_TMA | HPEM_FLAG_NO_QUAD | HPEM_FLAG_LANES4));
    if (quad_candidate && (count & 3) && count > 4) {
        int rc = launch_range(g, in, out, first, count & ~(int64_t)3, torr, plume, flags, st);
        if (rc != HPEM_OK) return rc;
        return launch_range(g, in, out, first + (count & ~(int64_t)3), count & 3, torr, plume, flags, st);
    }
    hpem::EvalParams p;
    fill_params(g, in, out, first, count, torr, p);
    if (flags & HPEM_FLAG_NO_FASTMATH) p.no_fastmath = 1;
    if (flags & HPEM_FLAG_NO_QTABLE) p.qt.rows = nullptr;
    return launch(g, p, plume, store_j, flags, st);
}

int check_request(const hpem_grid* g, int64_t n, const hpem_inputs* in, const hpem_outputs* out) {
    if (!g) return fail(HPEM_ERR_INVALID_ARG, "grid handle is NULL");
    if (!in || !out) return fail(HPEM_ERR_INVALID_ARG, "inputs/outputs struct is NULL");
    if (n < 0) return fail(HPEM_ERR_INVALID_ARG, "negative sample count %lld", (long long)n);
    if (n > (int64_t)2147483647 * 64) return fail(HPEM_ERR_INVALID_ARG, "sample count %lld too large for one call", (long long)n);
    if (!out->V_cc && !wants_plume(*out)) return fail(HPEM_ERR_INVALID_ARG, "no output requested");
    return HPEM_OK;
}

template <typename T>
int grow(T*& ptr, size_t& cap, size_t need) {
    if (need <= cap) return HPEM_OK;
    if (ptr) HPEM_CUDA(cudaFree(ptr));
    ptr = nullptr;
    cap = 0;
    const size_t want = need + need / 8;
    HPEM_CUDA(cudaMalloc((void**)&ptr, want * sizeof(T)));
    cap = want;
    return HPEM_OK;
}

}  // namespace

extern "C" {

int hpem_abi_version(void) { return HPEM_ABI_VERSION; }

#ifndef HPEM_SOURCE_HASH
#define HPEM_SOURCE_HASH unstamped
#endif
#define HPEM_STR2(x) #x
#define HPEM_STR(x) HPEM_STR2(x)
// "HPEM_SOURCE_HASH=<16 hex digits>": the loader finds the marker in the file (no dlopen) and compares it with the tree
static const char g_source_hash[] = "HPEM_SOURCE_HASH=" HPEM_STR(HPEM_SOURCE_HASH);
const char* hpem_source_hash(void) { return g_source_hash + 17; }

const char* hpem_last_error(void) { return g_err; }

int64_t hpem_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int hpem_grid_create(int device, int n_angles, const double* alpha, const double* wd, const double* wn, int n_radii,
                     const double* radii, hpem_grid** out) {
    if (!out) return fail(HPEM_ERR_INVALID_ARG, "out handle pointer is NULL");
    *out = nullptr;
    if (n_angles < 2 || n_angles > 8192) return fail(HPEM_ERR_INVALID_ARG, "n_angles must be in [2, 8192], got %d", n_angles);
    if (n_radii < 1 || n_radii > 4096) return fail(HPEM_ERR_INVALID_ARG, "n_radii must be in [1, 4096], got %d", n_radii);
    if (!alpha || !wd || !wn || !radii) return fail(HPEM_ERR_INVALID_ARG, "alpha/wd/wn/radii must be non-NULL host arrays");
    int ndev = 0;
    HPEM_CUDA(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(HPEM_ERR_INVALID_ARG, "device %d out of range (%d visible)", device, ndev);
    DeviceGuard guard(device);
    if (!guard.ok) return fail(HPEM_ERR_CUDA, "cannot select device %d", device);

    hpem_grid* g = new (std::nothrow) hpem_grid();
    if (!g) return fail(HPEM_ERR_CUDA, "out of host memory");
    g->device = device;
    g->n_angles = n_angles;
    g->n_angles_pad = (n_angles + hpem::kAnglePad - 1) / hpem::kAnglePad * hpem::kAnglePad;
    g->n_radii = n_radii;
    g->radius0 = radii[0];
    g->alpha_host.assign(alpha, alpha + n_angles);
    cudaDeviceGetAttribute(&g->sm_count, cudaDevAttrMultiProcessorCount, device);
    g->h = alpha[1];
    // uniform <=> alpha[i] == i*alpha[1] to rounding (np.linspace(0, pi/2, A), plume.py:53)
    bool uni = (alpha[0] == 0.0) && (alpha[1] > 0.0);
    for (int i = 2; i < n_angles && uni; ++i) {
        const double want = double(i) * alpha[1];
        if (!(std::fabs(alpha[i] - want) <= 4.0 * 2.220446049250313e-16 * std::fabs(want))) uni = false;
    }
    g->uniform = uni;

    std::vector<double2> w(g->n_angles_pad);
    for (int i = 0; i < g->n_angles_pad; ++i) w[i] = (i < n_angles) ? make_double2(wd[i], wn[i]) : make_double2(0.0, 0.0);

    auto cleanup = [&](int rc) {
        hpem_grid_destroy(g);
        return rc;
    };
#define HPEM_CUDA_G(call)                                                                                    \
    do {                                                                                                     \
        cudaError_t e_ = (call);                                                                             \
        if (e_ != cudaSuccess)                                                                               \
            return cleanup(fail(HPEM_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_)));             \
    } while (0)
    g->w_host = w;
    HPEM_CUDA_G(cudaMalloc((void**)&g->d_w, w.size() * sizeof(double2)));
    HPEM_CUDA_G(cudaMalloc((void**)&g->d_alpha, n_angles * sizeof(double)));
    HPEM_CUDA_G(cudaMalloc((void**)&g->d_radii, n_radii * sizeof(double)));
    HPEM_CUDA_G(cudaMemcpy(g->d_w, w.data(), w.size() * sizeof(double2), cudaMemcpyHostToDevice));
    HPEM_CUDA_G(cudaMemcpy(g->d_alpha, alpha, n_angles * sizeof(double), cudaMemcpyHostToDevice));
    HPEM_CUDA_G(cudaMemcpy(g->d_radii, radii, n_radii * sizeof(double), cudaMemcpyHostToDevice));
    if (g->uniform && n_angles <= 2048) {   // tabulated Simpson sums (host build: 20-170 ms for 91-512 angles, 0.5 s for 2048; shared by the handles of one process)
        std::shared_ptr<const QTableHost> t = qtable_for(n_angles, wd, wn);
        g->qt_key_lo = t->key_lo;
        g->qt_bins = t->n_bins;
        HPEM_CUDA_G(cudaMalloc((void**)&g->d_qtable, t->rows.size() * sizeof(double)));
        HPEM_CUDA_G(cudaMemcpy(g->d_qtable, t->rows.data(), t->rows.size() * sizeof(double), cudaMemcpyHostToDevice));
    }
#undef HPEM_CUDA_G

    const size_t wbytes = size_t(g->n_angles_pad) * sizeof(double2) + 1024;  // K1r / K1v: + slack for the 1024-byte alignment
    // K1u: weights up to the last 16-angle chunk, no alignment slack (address-based swizzle, see the kernel)
    const size_t wb_u = size_t((n_angles + hpem::kChunk - 1) / hpem::kChunk * hpem::kChunk) * sizeof(double2) + HPEM_SMEM_SKEW;
    g->smem_nostore = wb_u;
    g->smem_stg = wb_u + size_t(hpem::kWarpsU) * 32 * hpem::kTilePitch * sizeof(double);
    g->smem_tma = wb_u + size_t(hpem::kWarpsU) * 2 * hpem::kTmaGroupBytes;
    // long aligned rows, one-warp blocks: 4 KB of unused shared memory per block cap the SM at 9 resident blocks instead of
    // 11 -- the store stream prefers fewer, longer-lived writers (1e6 x 200: 0.3145 -> 0.3107 ms, x 512: 0.709 -> 0.7025 ms)
    g->smem_tma32 = wb_u + size_t(2) * hpem::kTmaGroupBytes + 4096;
    g->smem_tma1 = wb_u + size_t(hpem::kWarpsU) * 1 * hpem::kTmaGroupBytes;
    g->smem_quad = g->smem_tma;      // the row-boundary buffer aliases the staging area
    g->smem_quad1 = g->smem_tma1;
    g->smem_rows = wb_u + size_t(hpem::kWarpsU) * ((size_t(32) * n_angles * 8 + 15) & ~size_t(15));
    const size_t xbytes = size_t(hpem::kWarpsV) * 32 * hpem::kXchPitch * sizeof(double);
    g->smem_v_nostore = wbytes + xbytes;
    g->smem_v_store = wbytes + xbytes + size_t(hpem::kWarpsV) * hpem::k1v_tile_bytes(n_angles);
    g->smem_ok = std::max(std::max(g->smem_stg, g->smem_tma), g->smem_v_store) <= 200 * 1024;
    if (g->smem_ok) {
        int rc = set_smem(hpem::eval_uniform_kernel<true, true, hpem::kStoreTma, 2>, g->smem_tma);
        if (rc == HPEM_OK) rc = set_smem(hpem::eval_uniform_kernel<true, true, hpem::kStoreTma, 1>, g->smem_tma1);
        if (rc == HPEM_OK) rc = set_smem(hpem::eval_uniform_kernel<true, true, hpem::kStoreTma, 2, 32>, g->smem_tma32);
        if (rc == HPEM_OK) rc = set_smem(hpem::eval_uniform_kernel<true, true, hpem::kStoreQuad, 2>, g->smem_quad);
        if (rc == HPEM_OK) rc = set_smem(hpem::eval_uniform_kernel<true, true, hpem::kStoreQuad, 1>, g->smem_quad1);
        if (rc == HPEM_OK) rc = set_smem(hpem::eval_uniform_kernel<true, true, hpem::kStoreStg, 2>, g->smem_stg);
        if (rc == HPEM_OK && g->smem_rows <= 200 * 1024)
            rc = set_smem(hpem::eval_uniform_kernel<true, true, hpem::kStoreRows, 2>, g->smem_rows);
        if (rc == HPEM_OK) rc = set_smem(hpem::eval_uniform_kernel<true, false, hpem::kStoreStg, 2>, g->smem_nostore);
        if (rc == HPEM_OK) rc = set_smem(hpem::eval_lanes4_kernel<true>, g->smem_v_store);
        if (rc == HPEM_OK) rc = set_smem(hpem::eval_lanes4_kernel<false>, g->smem_v_nostore);
        if (rc != HPEM_OK) return cleanup(rc);
    }
    *out = g;
    return HPEM_OK;
}

int hpem_grid_destroy(hpem_grid* g) {
    if (!g) return HPEM_OK;
    DeviceGuard guard(g->device);
    Workspace& ws = g->ws;
    for (auto& p : ws.d_in) if (p) cudaFree(p);
    for (auto& p : ws.d_small) if (p) cudaFree(p);
    if (ws.d_invalid) cudaFree(ws.d_invalid);
    if (ws.d_j) cudaFree(ws.d_j);
    if (ws.h_stage_in) cudaFreeHost(ws.h_stage_in);
    if (ws.h_stage_out) cudaFreeHost(ws.h_stage_out);
    if (ws.d_pack_in) cudaFree(ws.d_pack_in);
    if (ws.d_pack_out) cudaFree(ws.d_pack_out);
    if (ws.d_partials) cudaFree(ws.d_partials);
    if (ws.d_partial_minmax) cudaFree(ws.d_partial_minmax);
    if (ws.d_hist_partials) cudaFree(ws.d_hist_partials);
    if (ws.moments_done) cudaEventDestroy(ws.moments_done);
    for (auto e : ws.events) cudaEventDestroy(e);
    for (auto e : ws.h2d_events) cudaEventDestroy(e);
    if (ws.s_h2d) cudaStreamDestroy(ws.s_h2d);
    if (ws.s_compute) cudaStreamDestroy(ws.s_compute);
    if (ws.s_copy) cudaStreamDestroy(ws.s_copy);
    if (g->d_w) cudaFree(g->d_w);
    if (g->d_alpha) cudaFree(g->d_alpha);
    if (g->d_radii) cudaFree(g->d_radii);
    if (g->d_qtable) cudaFree(g->d_qtable);
    delete g;
    return HPEM_OK;
}

int hpem_grid_is_uniform(const hpem_grid* g) {
    if (!g) return fail(HPEM_ERR_INVALID_ARG, "grid handle is NULL");
    return (g->uniform && g->smem_ok) ? 1 : 0;
}

int hpem_eval(const hpem_grid* g, int64_t n, const hpem_inputs* in, const hpem_outputs* out, double torr_2_pa,
              uint32_t flags, void* stream) {
    int rc = check_request(g, n, in, out);
    if (rc != HPEM_OK) return rc;
    const bool plume = wants_plume(*out);
    DeviceGuard guard(g->device);
    if (!guard.ok) return fail(HPEM_ERR_CUDA, "cannot select device %d", g->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // one launch covers at most 2^31-1 blocks; split huge batches
    const int64_t max_per_launch = (int64_t)1 << 30;
    for (int64_t first = 0; first < n; first += max_per_launch) {
        const int64_t count = std::min(max_per_launch, n - first);
        rc = launch_range(*g, *in, *out, first, count, torr_2_pa, plume, flags, st);
        if (rc != HPEM_OK) return rc;
    }
    return HPEM_OK;
}

int hpem_eval_host(hpem_grid* g, int64_t n, const hpem_inputs* in, const hpem_outputs* out, double torr_2_pa,
                   uint32_t flags) {
    int rc = check_request(g, n, in, out);
    if (rc != HPEM_OK) return rc;
    if (n == 0) return HPEM_OK;
    const bool plume = wants_plume(*out);
    DeviceGuard guard(g->device);
    if (!guard.ok) return fail(HPEM_ERR_CUDA, "cannot select device %d", g->device);
    Workspace& ws = g->ws;
    std::lock_guard<std::mutex> lock(ws.mu);
    // Whatever way this call ends -- also on an error after the first enqueue -- the three workspace streams are drained
    // before the mutex is released and the caller gets its buffers back: no kernel or copy is left in flight on memory
    // the caller may recycle (pinned output pool) or the next call may reuse (d_j, d_in).
    struct Quiesce {
        Workspace& w;
        ~Quiesce() {
            if (w.s_h2d) cudaStreamSynchronize(w.s_h2d);
            if (w.s_compute) cudaStreamSynchronize(w.s_compute);
            if (w.s_copy) cudaStreamSynchronize(w.s_copy);
        }
    } quiesce{ws};
    if (!ws.s_compute) HPEM_CUDA(cudaStreamCreateWithFlags(&ws.s_compute, cudaStreamNonBlocking));
    if (!ws.s_copy) HPEM_CUDA(cudaStreamCreateWithFlags(&ws.s_copy, cudaStreamNonBlocking));
    if (!ws.s_h2d) HPEM_CUDA(cudaStreamCreateWithFlags(&ws.s_h2d, cudaStreamNonBlocking));

    const int64_t A = g->n_angles, R = g->n_radii;
    // which inputs are needed
    bool need[HPEM_N_INPUTS] = {};
    if (out->V_cc) for (int k : kCathodeInputs) need[k] = true;
    if (plume) for (int k : kPlumeInputs) need[k] = true;
    if (out->T_c) need[HPEM_IN_T] = true;

    // ---- small batches (what amisc passes outside of Monte-Carlo runs, down to one sample per call): the per-array
    // cudaMemcpy calls of the pipeline below cost more than the whole evaluation, so the needed inputs are packed into
    // one pinned staging buffer (ONE H2D), the per-sample outputs come back packed (ONE D2H), one stream, one sync.
    {
        constexpr size_t kPackBytes = size_t(2) << 20;
        int n_arrays = 0;
        for (int k = 0; k < HPEM_N_INPUTS; ++k) n_arrays += (need[k] && in->ptr[k]) ? 1 : 0;
        const size_t in_bytes = size_t(n_arrays) * n * sizeof(double);
        const size_t per_r = size_t(n) * R * sizeof(double);
        const size_t out_bytes = (out->V_cc ? size_t(n) * 8 : 0) + (out->div_angle ? per_r : 0) + (out->T_c ? per_r : 0) +
                                 (out->cos_div ? per_r : 0) + (out->invalid ? ((size_t(n) + 7) & ~size_t(7)) : 0);
        const size_t j_bytes = out->j_ion ? size_t(n) * A * R * sizeof(double) : 0;
        if (in_bytes <= kPackBytes && out_bytes <= kPackBytes && j_bytes <= (size_t(64) << 20)) {
            if (!ws.h_stage_in) {
                HPEM_CUDA(cudaHostAlloc((void**)&ws.h_stage_in, kPackBytes, cudaHostAllocDefault));
                HPEM_CUDA(cudaHostAlloc((void**)&ws.h_stage_out, kPackBytes, cudaHostAllocDefault));
                HPEM_CUDA(cudaMalloc((void**)&ws.d_pack_in, kPackBytes));
                HPEM_CUDA(cudaMalloc((void**)&ws.d_pack_out, kPackBytes));
            }
            hpem_inputs din = *in;
            size_t off = 0;
            for (int k = 0; k < HPEM_N_INPUTS; ++k) {
                din.ptr[k] = nullptr;
                if (!need[k] || !in->ptr[k]) continue;
                std::memcpy(ws.h_stage_in + off, in->ptr[k], size_t(n) * sizeof(double));
                din.ptr[k] = reinterpret_cast<const double*>(ws.d_pack_in + off);
                off += size_t(n) * sizeof(double);
            }
            if (off) HPEM_CUDA(cudaMemcpyAsync(ws.d_pack_in, ws.h_stage_in, off, cudaMemcpyHostToDevice, ws.s_compute));
            hpem_outputs dout = {};
            size_t o = 0;
            auto take = [&](bool wanted, size_t bytes) -> unsigned char* {
                if (!wanted) return nullptr;
                unsigned char* ptr = ws.d_pack_out + o;
                o += bytes;
                return ptr;
            };
            dout.V_cc = reinterpret_cast<double*>(take(out->V_cc != nullptr, size_t(n) * 8));
            dout.div_angle = reinterpret_cast<double*>(take(out->div_angle != nullptr, per_r));
            dout.T_c = reinterpret_cast<double*>(take(out->T_c != nullptr, per_r));
            dout.cos_div = reinterpret_cast<double*>(take(out->cos_div != nullptr, per_r));
            dout.invalid = take(out->invalid != nullptr, (size_t(n) + 7) & ~size_t(7));
            if (out->j_ion) {
                rc = grow(ws.d_j, ws.j_cap, (size_t)(n * A * R));
                if (rc != HPEM_OK) return rc;
                dout.j_ion = ws.d_j;
            }
            rc = launch_range(*g, din, dout, 0, n, torr_2_pa, plume, flags, ws.s_compute);
            if (rc != HPEM_OK) return rc;
            if (o) HPEM_CUDA(cudaMemcpyAsync(ws.h_stage_out, ws.d_pack_out, o, cudaMemcpyDeviceToHost, ws.s_compute));
            if (out->j_ion) HPEM_CUDA(cudaMemcpyAsync(out->j_ion, ws.d_j, j_bytes, cudaMemcpyDeviceToHost, ws.s_compute));
            HPEM_CUDA(cudaStreamSynchronize(ws.s_compute));
            const unsigned char* hs = ws.h_stage_out;
            if (out->V_cc) { std::memcpy(out->V_cc, hs, size_t(n) * 8); hs += size_t(n) * 8; }
            if (out->div_angle) { std::memcpy(out->div_angle, hs, per_r); hs += per_r; }
            if (out->T_c) { std::memcpy(out->T_c, hs, per_r); hs += per_r; }
            if (out->cos_div) { std::memcpy(out->cos_div, hs, per_r); hs += per_r; }
            if (out->invalid) std::memcpy(out->invalid, hs, size_t(n));
            return HPEM_OK;
        }
    }

    // super-batches bound the device footprint of j_ion (default cap 16 GiB of the 180 GB HBM)
    const int64_t row_elems = A * R;
    // (HPEM_HOST_BATCH_BYTES / HPEM_HOST_CHUNK_BYTES override the two sizes: tests drive the multi-batch pipeline with them)
    auto env_bytes = [](const char* name, int64_t dflt) {
        const char* v = std::getenv(name);
        const long long x = v ? std::atoll(v) : 0;
        return x > 0 ? (int64_t)x : dflt;
    };
    const int64_t cap_elems = env_bytes("HPEM_HOST_BATCH_BYTES", (int64_t)16 << 30) / 8;
    // (a multiple of 64 samples: every launch then sees a sample at the same position modulo 4 / 32 as a single launch over
    //  the whole batch would, so the quad-row kernel -- whose rounding depends on a row's phase -- gives identical bits)
    const int64_t batch = out->j_ion ? std::min<int64_t>(n, std::max<int64_t>(64, cap_elems / row_elems / 64 * 64)) : n;
    // D2H chunks: ~32 MiB so the copy engine starts as soon as the first rows exist; 128 MiB once the output is large
    // enough for the per-copy gaps to matter more than the start-up (1e6 x 200: 30.3 -> 29.6 ms)
    const int64_t dflt_chunk_bytes = (n * row_elems * 8 >= ((int64_t)512 << 20)) ? ((int64_t)128 << 20) : ((int64_t)32 << 20);
    const int64_t chunk = out->j_ion ? (std::max<int64_t>(1024, env_bytes("HPEM_HOST_CHUNK_BYTES", dflt_chunk_bytes) / (row_elems * 8)) + 63) / 64 * 64 : batch;

    for (int64_t b0 = 0; b0 < n; b0 += batch) {
        const int64_t nb = std::min(batch, n - b0);
        hpem_inputs din = *in;
        for (int k = 0; k < HPEM_N_INPUTS; ++k) din.ptr[k] = nullptr;
        // (re)allocate inputs with one common capacity
        if ((size_t)nb > ws.in_cap) {
            for (int k = 0; k < HPEM_N_INPUTS; ++k) {
                if (ws.d_in[k]) { cudaFree(ws.d_in[k]); ws.d_in[k] = nullptr; }
            }
            ws.in_cap = (size_t)nb + (size_t)nb / 8;
        }
        for (int k = 0; k < HPEM_N_INPUTS; ++k) {
            if (!need[k] || !in->ptr[k]) continue;
            if (!ws.d_in[k]) HPEM_CUDA(cudaMalloc((void**)&ws.d_in[k], ws.in_cap * sizeof(double)));
            din.ptr[k] = ws.d_in[k];
        }
        hpem_outputs dout = {};
        double* host_small[4] = {out->V_cc, out->div_angle, out->T_c, out->cos_div};
        const size_t small_elems[4] = {(size_t)nb, (size_t)(nb * R), (size_t)(nb * R), (size_t)(nb * R)};
        const size_t small_need = (size_t)(nb * R);
        if (small_need > ws.small_cap) {
            for (auto& p : ws.d_small) if (p) { cudaFree(p); p = nullptr; }
            ws.small_cap = small_need + small_need / 8;
        }
        double** dsmall_out[4] = {&dout.V_cc, &dout.div_angle, &dout.T_c, &dout.cos_div};
        for (int q = 0; q < 4; ++q) {
            if (!host_small[q]) continue;
            if (!ws.d_small[q]) HPEM_CUDA(cudaMalloc((void**)&ws.d_small[q], ws.small_cap * sizeof(double)));
            *dsmall_out[q] = ws.d_small[q];
        }
        if (out->invalid) {
            rc = grow(ws.d_invalid, ws.invalid_cap, (size_t)nb);
            if (rc != HPEM_OK) return rc;
            dout.invalid = ws.d_invalid;
        }
        if (out->j_ion) {
            rc = grow(ws.d_j, ws.j_cap, (size_t)(nb * row_elems));
            if (rc != HPEM_OK) return rc;
            dout.j_ion = ws.d_j;
        }

        // H2D of the per-sample inputs in segments (the first one chunk long, then ~16 MiB of inputs each) on their own stream.  The
        // upload of segment k+1 is ISSUED after the kernels and D2H copies of segment k: with pinned inputs the order does
        // not matter (everything is asynchronous, PCIe is full duplex), but a cudaMemcpyAsync from PAGEABLE memory -- what
        // amisc passes -- blocks the host while the driver stages it, and issued up front those 120 B per sample delayed the
        // first kernel by the whole upload (40 ms instead of 32 ms per 1e6 x 200 batch).
        const int64_t n_chunks = (nb + chunk - 1) / chunk;
        // chunks per upload segment: about 16 MiB of input arrays (a pageable upload of that size blocks the host for ~1.5 ms,
        // less than the D2H of the segment before it keeps the GPU's copy engine busy)
        int n_in_arrays = 0;
        for (int k = 0; k < HPEM_N_INPUTS; ++k) n_in_arrays += din.ptr[k] ? 1 : 0;
        const int64_t spc = std::max<int64_t>(1, ((int64_t)16 << 20) / std::max<int64_t>(1, chunk * n_in_arrays * 8));
        const int64_t n_seg = 1 + (std::max<int64_t>(n_chunks - 1, 0) + spc - 1) / spc;
        auto seg_first_chunk = [spc](int64_t sg) { return sg == 0 ? (int64_t)0 : 1 + (sg - 1) * spc; };
        while ((int64_t)ws.h2d_events.size() < n_seg) {
            cudaEvent_t e;
            HPEM_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            ws.h2d_events.push_back(e);
        }
        while ((int64_t)ws.events.size() < n_chunks) {
            cudaEvent_t e;
            HPEM_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            ws.events.push_back(e);
        }
        auto upload = [&](int64_t sg) -> int {
            const int64_t c0 = seg_first_chunk(sg), c1 = std::min(n_chunks, seg_first_chunk(sg + 1));
            const int64_t first = c0 * chunk, count = std::min(nb, c1 * chunk) - first;
            for (int k = 0; k < HPEM_N_INPUTS; ++k) {
                if (!din.ptr[k]) continue;
                HPEM_CUDA(cudaMemcpyAsync(ws.d_in[k] + first, in->ptr[k] + b0 + first, (size_t)count * sizeof(double),
                                          cudaMemcpyHostToDevice, ws.s_h2d));
            }
            HPEM_CUDA(cudaEventRecord(ws.h2d_events[sg], ws.s_h2d));
            return HPEM_OK;
        };
        rc = upload(0);
        if (rc != HPEM_OK) return rc;
        for (int64_t sg = 0; sg < n_seg; ++sg) {
            const int64_t c0 = seg_first_chunk(sg), c1 = std::min(n_chunks, seg_first_chunk(sg + 1));
            HPEM_CUDA(cudaStreamWaitEvent(ws.s_compute, ws.h2d_events[sg], 0));
            for (int64_t c = c0; c < c1; ++c) {
                const int64_t first = c * chunk, count = std::min(chunk, nb - first);
                rc = launch_range(*g, din, dout, first, count, torr_2_pa, plume, flags, ws.s_compute);
                if (rc != HPEM_OK) return rc;
                if (out->j_ion) {
                    HPEM_CUDA(cudaEventRecord(ws.events[c], ws.s_compute));
                    HPEM_CUDA(cudaStreamWaitEvent(ws.s_copy, ws.events[c], 0));
                    HPEM_CUDA(cudaMemcpyAsync(out->j_ion + (b0 + first) * row_elems, ws.d_j + first * row_elems,
                                              (size_t)(count * row_elems) * sizeof(double), cudaMemcpyDeviceToHost, ws.s_copy));
                }
            }
            if (sg + 1 < n_seg) {
                rc = upload(sg + 1);
                if (rc != HPEM_OK) return rc;
            }
        }
        for (int q = 0; q < 4; ++q) {
            if (!host_small[q]) continue;
            const int64_t per = (q == 0) ? 1 : R;
            HPEM_CUDA(cudaMemcpyAsync(host_small[q] + b0 * per, ws.d_small[q], small_elems[q] * sizeof(double),
                                      cudaMemcpyDeviceToHost, ws.s_compute));
        }
        if (out->invalid)
            HPEM_CUDA(cudaMemcpyAsync(out->invalid + b0, ws.d_invalid, (size_t)nb, cudaMemcpyDeviceToHost, ws.s_compute));
        HPEM_CUDA(cudaStreamSynchronize(ws.s_h2d));
        HPEM_CUDA(cudaStreamSynchronize(ws.s_compute));
        HPEM_CUDA(cudaStreamSynchronize(ws.s_copy));
    }
    return HPEM_OK;
}

}  // extern "C"

namespace {

int moments_layout(const hpem_grid* g, const hpem_moments_spec* spec, hpem_moments_layout* lay) {
    if (!g || !spec || !lay) return fail(HPEM_ERR_INVALID_ARG, "NULL argument");
    if (g->n_radii != 1) return fail(HPEM_ERR_UNSUPPORTED, "the reduce-only pass supports a single sweep radius");
    if (!(g->uniform && g->smem_ok)) return fail(HPEM_ERR_UNSUPPORTED, "the reduce-only pass needs the uniform angle grid");
    const int st = spec->hist_angle_stride;
    if (st < 0 || (st & (st - 1)) != 0) return fail(HPEM_ERR_INVALID_ARG, "hist_angle_stride must be 0 or a power of two");
    if (spec->hist_sub_bits < 0 || spec->hist_sub_bits > 6) return fail(HPEM_ERR_INVALID_ARG, "hist_sub_bits must be in [0, 6]");
    if (st > 0 && !(spec->hist_min_exp2 < spec->hist_max_exp2 && spec->hist_min_exp2 > -1000 && spec->hist_max_exp2 < 1000))
        return fail(HPEM_ERR_INVALID_ARG, "need -1000 < hist_min_exp2 < hist_max_exp2 < 1000");
    lay->n_hist_angles = st > 0 ? (g->n_angles + st - 1) / st : 0;
    lay->n_bins = st > 0 ? ((spec->hist_max_exp2 - spec->hist_min_exp2) << spec->hist_sub_bits) + 2 : 0;
    lay->off_angle_sum = hpem::kMomScalars;
    lay->off_angle_sumsq = lay->off_angle_sum + g->n_angles;
    lay->off_hist = lay->off_angle_sumsq + g->n_angles;
    lay->n_sums = lay->off_hist + (int64_t)lay->n_hist_angles * lay->n_bins;
    lay->n_minmax = 6;
    lay->reserved = 0;
    return HPEM_OK;
}

}  // namespace

extern "C" {

int hpem_moments_layout_query(const hpem_grid* g, const hpem_moments_spec* spec, hpem_moments_layout* lay) {
    return moments_layout(g, spec, lay);
}

static void fill_moments_params(const hpem_moments_spec& spec, const hpem_moments_layout& lay, hpem::MomentsParams& m) {
    std::memset(&m, 0, sizeof(m));
    m.hist_stride = spec.hist_angle_stride;
    m.hist_shift = 0;
    while ((1 << m.hist_shift) < m.hist_stride) ++m.hist_shift;
    m.want_cathode = spec.want_cathode;
    m.hist_sub_bits = spec.hist_sub_bits;
    m.hist_min_exp2 = spec.hist_min_exp2;
    m.hist_max_exp2 = spec.hist_max_exp2;
    m.n_hist_angles = lay.n_hist_angles;
    m.n_bins = lay.n_bins;
    m.n_sums = lay.n_sums;
    m.off_angle_sum = lay.off_angle_sum;
    m.off_angle_sumsq = lay.off_angle_sumsq;
    m.off_hist = lay.off_hist;
    for (int k = 0; k < 3; ++k) m.shift[k] = spec.scalar_shift[k];
}

static int moments_run(hpem_grid* g, int64_t n, const hpem_inputs* in, const hpem::SamplerParams* sampler, double torr_2_pa,
                       const hpem_moments_spec* spec, double* sums, double* minmax, void* stream) {
    hpem_moments_layout lay;
    int rc = moments_layout(g, spec, &lay);
    if (rc != HPEM_OK) return rc;
    if ((!in && !sampler) || !sums || !minmax) return fail(HPEM_ERR_INVALID_ARG, "inputs/sums/minmax must be non-NULL");
    if (n < 0) return fail(HPEM_ERR_INVALID_ARG, "negative sample count");
    if (n > (int64_t)4000000000LL) return fail(HPEM_ERR_INVALID_ARG, "at most 4e9 samples per call (32-bit histogram counters per block)");
    for (int k = 0; k < 3; ++k)
        if (!std::isfinite(spec->scalar_shift[k])) return fail(HPEM_ERR_INVALID_ARG, "scalar_shift[%d] must be finite", k);
    if (n == 0) return HPEM_OK;
    DeviceGuard guard(g->device);
    if (!guard.ok) return fail(HPEM_ERR_CUDA, "cannot select device %d", g->device);
    using namespace hpem;
    const int n_chunks = (g->n_angles + kChunk - 1) / kChunk;
    // One persistent block per SM with as many warps (<= 12, two samples per thread) as shared memory allows: per warp a
    // 32 x 16 tile of (t, q) pairs and the per-angle accumulators.
    int dev_smem = 0;
    HPEM_CUDA(cudaDeviceGetAttribute(&dev_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, g->device));
    const size_t smem_budget = (size_t)dev_smem - 4096;   // static shared memory of the kernel + slack
    static const int warps_env = []() { const char* v = std::getenv("HPEM_MOMENTS_WARPS"); return v ? std::atoi(v) : 0; }();
    const bool restart = n_chunks > kRestartChunks;
    const int hs = spec->hist_angle_stride == 0 ? 0 : (spec->hist_angle_stride == 8 ? 8 : -1);
    static const int ns_env = []() { const char* v = std::getenv("HPEM_MOMENTS_NS"); return v ? std::atoi(v) : 0; }();
    // three samples per thread for long rows (instantiated for the two common histogram settings), two otherwise
    const int ns = (hs != -1 && (restart || (ns_env ? ns_env == 3 : n_chunks >= kLongChunksM))) ? 3 : 2;   // restart rows of strides 0 / 8 exist with three samples only
    const int max_warps = (restart || ns == 3) ? kWarpsLongM : kMaxWarpsM;      // the launch bounds of the kernel families
    int warps = warps_env > 0 ? std::min(warps_env, max_warps) : max_warps;
    while (warps > 1 && moments_smem_bytes(g->n_angles_pad, n_chunks * kChunk, warps) > smem_budget) --warps;
    const size_t smem = moments_smem_bytes(g->n_angles_pad, n_chunks * kChunk, warps);
    if (smem > smem_budget)
        return fail(HPEM_ERR_UNSUPPORTED, "%d angles need %zu bytes of shared memory (> %zu)", g->n_angles, smem, smem_budget);
    const int threads = warps * 32;
    const int64_t batches = (n + (int64_t)ns * threads - 1) / ((int64_t)ns * threads);
    const int blocks = (int)std::min<int64_t>(batches, (int64_t)g->sm_count);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    Workspace& ws = g->ws;
    std::lock_guard<std::mutex> lock(ws.mu);
    // the per-block scratch buffers are shared by every caller of this grid handle: a pass on another stream waits for the
    // previous one to finish with them
    if (!ws.moments_done) HPEM_CUDA(cudaEventCreateWithFlags(&ws.moments_done, cudaEventDisableTiming));
    else HPEM_CUDA(cudaStreamWaitEvent(st, ws.moments_done, 0));
    const size_t n_part = (size_t)kMomScalars + 2 * (size_t)g->n_angles;
    const size_t n_hist = (size_t)lay.n_hist_angles * lay.n_bins;
    rc = grow(ws.d_partials, ws.partials_cap, (size_t)g->sm_count * n_part);
    if (rc == HPEM_OK) rc = grow(ws.d_partial_minmax, ws.partial_minmax_cap, (size_t)g->sm_count * 6);
    if (rc != HPEM_OK) return rc;
    if ((size_t)g->sm_count * n_hist > ws.hist_partials_cap) {
        HPEM_CUDA(cudaStreamSynchronize(st));          // a previous pass may still be reading the old buffer
        rc = grow(ws.d_hist_partials, ws.hist_partials_cap, (size_t)g->sm_count * n_hist);
        if (rc != HPEM_OK) return rc;
        HPEM_CUDA(cudaMemsetAsync(ws.d_hist_partials, 0, ws.hist_partials_cap * sizeof(unsigned), st));
    }

    hpem_outputs no_out = {};
    hpem_inputs no_in = {};
    EvalParams p;
    fill_params(*g, in ? *in : no_in, no_out, 0, n, torr_2_pa, p);
    p.has_thrust = spec->want_thrust != 0;
    MomentsParams m;
    fill_moments_params(*spec, lay, m);
    m.partials = ws.d_partials;
    m.partial_minmax = ws.d_partial_minmax;
    m.hist_partials = ws.d_hist_partials;
    SamplerParams sp_zero;
    std::memset(&sp_zero, 0, sizeof(sp_zero));
    const SamplerParams& sp = sampler ? *sampler : sp_zero;
#define HPEM_MOMENTS_GO(S, H, R, N)                                        \
    do {                                                                   \
        rc = set_smem(moments_kernel<S, H, R, N>, smem);                   \
        if (rc != HPEM_OK) return rc;                                      \
        moments_kernel<S, H, R, N><<<blocks, threads, smem, st>>>(p, m, sp); \
    } while (0)
// instantiated: two samples per thread for every (stride, restart) except the long rows of the two common strides, which
// exist with three samples only; three samples for strides 0 and 8
#define HPEM_MOMENTS_GO_N(S, H, R)                                         \
    do {                                                                   \
        if (ns == 3) HPEM_MOMENTS_GO(S, H, R, 3); else HPEM_MOMENTS_GO(S, H, false, 2); \
    } while (0)
#define HPEM_MOMENTS_GO_R(S, H) do { if (restart) HPEM_MOMENTS_GO_N(S, H, true); else HPEM_MOMENTS_GO_N(S, H, false); } while (0)
#define HPEM_MOMENTS_GO_ANY(S) do { if (restart) HPEM_MOMENTS_GO(S, -1, true, 2); else HPEM_MOMENTS_GO(S, -1, false, 2); } while (0)
#define HPEM_MOMENTS_GO_H(S) do { if (hs == 0) HPEM_MOMENTS_GO_R(S, 0); else if (hs == 8) HPEM_MOMENTS_GO_R(S, 8); else HPEM_MOMENTS_GO_ANY(S); } while (0)
    if (sampler) HPEM_MOMENTS_GO_H(true); else HPEM_MOMENTS_GO_H(false);
#undef HPEM_MOMENTS_GO_H
#undef HPEM_MOMENTS_GO_ANY
#undef HPEM_MOMENTS_GO_R
#undef HPEM_MOMENTS_GO_N
#undef HPEM_MOMENTS_GO
    HPEM_CUDA(cudaGetLastError());
    const int fthreads = 128;
    moments_finalize_kernel<<<(unsigned)((lay.n_sums + fthreads - 1) / fthreads), fthreads, 0, st>>>(
        ws.d_partials, ws.d_partial_minmax, ws.d_hist_partials, blocks, g->n_angles, m, sums, minmax);
    moments_counts_kernel<<<1, 32, 0, st>>>(ws.d_partials, blocks, g->n_angles, sums);
    HPEM_CUDA(cudaGetLastError());
    HPEM_CUDA(cudaEventRecord(ws.moments_done, st));
    g_launches.fetch_add(3, std::memory_order_relaxed);
    return HPEM_OK;
}

int hpem_quadrature_table_eval(int n_angles, const double* wd, const double* wn, int64_t n_x, const double* x, double* nd, double* nn) {
    if (n_angles < 2 || n_angles > 8192) return fail(HPEM_ERR_INVALID_ARG, "n_angles must be in [2, 8192], got %d", n_angles);
    if (!wd || !wn || n_x < 0 || (n_x > 0 && (!x || !nd || !nn))) return fail(HPEM_ERR_INVALID_ARG, "NULL argument");
    std::vector<double> rows;
    hpem::QTableRef q;
    hpem::qtable_build(n_angles, wd, wn, rows, q.key_lo, q.n_bins);
    q.rows = rows.data();
    q.wd0 = wd[0];
    q.wn0 = wn[0];
    for (int64_t k = 0; k < n_x; ++k) hpem::qtable_eval(q, x[k], std::exp(-x[k]), nd[k], nn[k]);
    return HPEM_OK;
}

int hpem_moments_merge(int device, const hpem_moments_layout* lay, int n_parts, const double* parts, int64_t part_stride,
                       double* sums, double* minmax, void* stream) {
    if (!lay || !parts || !sums || !minmax) return fail(HPEM_ERR_INVALID_ARG, "NULL argument");
    if (n_parts < 1) return fail(HPEM_ERR_INVALID_ARG, "need at least one part");
    if (part_stride < lay->n_sums + 6) return fail(HPEM_ERR_INVALID_ARG, "part_stride %lld < n_sums + 6", (long long)part_stride);
    const int64_t n_angles = lay->off_angle_sumsq - lay->off_angle_sum;
    if (n_angles < 2 || lay->off_angle_sum != hpem::kMomScalars || lay->off_hist != lay->off_angle_sumsq + n_angles ||
        lay->n_sums != lay->off_hist + (int64_t)lay->n_hist_angles * lay->n_bins)
        return fail(HPEM_ERR_INVALID_ARG, "inconsistent moments layout");
    DeviceGuard guard(device);
    if (!guard.ok) return fail(HPEM_ERR_CUDA, "cannot select device %d", device);
    hpem::MomentsParams m;
    std::memset(&m, 0, sizeof(m));
    m.n_hist_angles = lay->n_hist_angles;
    m.n_bins = lay->n_bins;
    m.n_sums = lay->n_sums;
    m.off_angle_sum = lay->off_angle_sum;
    m.off_angle_sumsq = lay->off_angle_sumsq;
    m.off_hist = lay->off_hist;
    const int threads = 128;
    hpem::moments_merge_kernel<<<(unsigned)((lay->n_sums + threads - 1) / threads), threads, 0, static_cast<cudaStream_t>(stream)>>>(
        parts, part_stride, n_parts, (int)n_angles, m, sums, minmax);
    HPEM_CUDA(cudaGetLastError());
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return HPEM_OK;
}

static int fill_sampler(uint64_t seed, uint64_t first_index, const hpem_prior* priors, hpem::SamplerParams* sp) {
    if (!priors) return fail(HPEM_ERR_INVALID_ARG, "priors must be non-NULL");
    sp->log_mask = sp->normal_mask = 0u;
    for (int k = 0; k < HPEM_N_INPUTS; ++k) {
        const hpem_prior& q = priors[k];
        if (q.kind < HPEM_PRIOR_CONST || q.kind > HPEM_PRIOR_NORMAL) return fail(HPEM_ERR_INVALID_ARG, "prior %d: unknown kind %d", k, q.kind);
        if (q.kind == HPEM_PRIOR_LOGUNIFORM && !(q.a > 0.0 && q.b > 0.0)) return fail(HPEM_ERR_INVALID_ARG, "prior %d: LogUniform needs positive bounds", k);
        sp->prior[k].kind = q.kind;
        sp->prior[k].reserved = 0;
        sp->prior[k].a = q.a;
        sp->prior[k].b = q.b;
        // LogUniform: exp(u (ln b - ln a) + ln a); the two logarithms are per-prior constants, taken once here
        sp->prior[k].log_a = q.kind == HPEM_PRIOR_LOGUNIFORM ? std::log(q.a) : 0.0;
        sp->prior[k].log_ratio = q.kind == HPEM_PRIOR_LOGUNIFORM ? std::log(q.b) - std::log(q.a) : 0.0;
        // value = scale * u + offset (then exp for LogUniform; Normal is drawn separately)
        sp->scale[k] = q.kind == HPEM_PRIOR_UNIFORM ? q.b - q.a : q.kind == HPEM_PRIOR_LOGUNIFORM ? sp->prior[k].log_ratio : 0.0;
        sp->offset[k] = q.kind == HPEM_PRIOR_LOGUNIFORM ? sp->prior[k].log_a : q.a;
        if (q.kind == HPEM_PRIOR_LOGUNIFORM) sp->log_mask |= 1u << k;
        if (q.kind == HPEM_PRIOR_NORMAL) sp->normal_mask |= 1u << k;
    }
    sp->seed = seed;
    sp->first_index = first_index;
    return HPEM_OK;
}

int hpem_moments_accumulate(hpem_grid* g, int64_t n, const hpem_inputs* in, double torr_2_pa, const hpem_moments_spec* spec,
                            double* sums, double* minmax, void* stream) {
    if (!in) return fail(HPEM_ERR_INVALID_ARG, "inputs must be non-NULL");
    return moments_run(g, n, in, nullptr, torr_2_pa, spec, sums, minmax, stream);
}

int hpem_moments_accumulate_sampled(hpem_grid* g, int64_t n, uint64_t seed, uint64_t first_index, const hpem_prior* priors,
                                    double torr_2_pa, const hpem_moments_spec* spec, double* sums, double* minmax, void* stream) {
    hpem::SamplerParams sp;
    int rc = fill_sampler(seed, first_index, priors, &sp);
    if (rc != HPEM_OK) return rc;
    return moments_run(g, n, nullptr, &sp, torr_2_pa, spec, sums, minmax, stream);
}

int hpem_sample_inputs(int device, int64_t n, uint64_t seed, uint64_t first_index, const hpem_prior* priors, double* const* out,
                       void* stream) {
    if (!out) return fail(HPEM_ERR_INVALID_ARG, "out must be non-NULL");
    if (n < 0) return fail(HPEM_ERR_INVALID_ARG, "negative sample count");
    hpem::SamplerParams sp;
    int rc = fill_sampler(seed, first_index, priors, &sp);
    if (rc != HPEM_OK) return rc;
    if (n == 0) return HPEM_OK;
    DeviceGuard guard(device);
    if (!guard.ok) return fail(HPEM_ERR_CUDA, "cannot select device %d", device);
    const unsigned blocks = (unsigned)std::min<int64_t>((n + 255) / 256, 148 * 8);
    hpem::sample_inputs_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        sp, n, out[0], out[1], out[2], out[3], out[4], out[5], out[6], out[7], out[8], out[9], out[10], out[11], out[12], out[13],
        out[14]);
    HPEM_CUDA(cudaGetLastError());
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return HPEM_OK;
}

}  // extern "C"

struct hpem_measurements {
    int device = 0;
    int m = 0, n_angles = 0;
    hpem::MeasPoint* d_meas = nullptr;
    int* d_seg = nullptr;
};

extern "C" {

int hpem_measurements_create(const hpem_grid* g, int m, const double* theta, const double* y, const double* sigma,
                             hpem_measurements** out) {
    if (!out) return fail(HPEM_ERR_INVALID_ARG, "out handle pointer is NULL");
    *out = nullptr;
    if (!g || !theta || !y || !sigma) return fail(HPEM_ERR_INVALID_ARG, "NULL argument");
    if (m < 1 || m > 4096) return fail(HPEM_ERR_INVALID_ARG, "number of measurement points must be in [1, 4096], got %d", m);
    if (g->n_radii != 1 || !g->uniform) return fail(HPEM_ERR_UNSUPPORTED, "log-likelihood needs one radius and the uniform grid");
    const int A = g->n_angles;
    const std::vector<double>& al = g->alpha_host;
    std::vector<hpem::MeasPoint> pts(m);
    std::vector<int> lo(m);
    for (int q = 0; q < m; ++q) {
        const double a = std::fabs(theta[q]);
        if (!(a <= al[A - 1])) return fail(HPEM_ERR_INVALID_ARG, "theta[%d] = %g is outside the sweep [-pi/2, pi/2]", q, theta[q]);
        if (!(sigma[q] > 0.0)) return fail(HPEM_ERR_INVALID_ARG, "sigma[%d] must be positive", q);
        // scipy.interpolate.interp1d: hi = clip(searchsorted(x, x_new, 'left'), 1, len-1), lo = hi - 1
        int hi = (int)(std::lower_bound(al.begin(), al.end(), a) - al.begin());
        hi = std::min(std::max(hi, 1), A - 1);
        lo[q] = hi - 1;
        pts[q].w = (a - al[hi - 1]) / (al[hi] - al[hi - 1]);
        pts[q].y = y[q];
        pts[q].inv_sigma = 1.0 / sigma[q];
        pts[q].orig = q;
        pts[q].rel = hi & (hpem::kChunk - 1);   // (lo + 1) mod 16
    }
    std::vector<int> order(m);
    for (int q = 0; q < m; ++q) order[q] = q;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return lo[a] < lo[b]; });
    std::vector<hpem::MeasPoint> sorted(m);
    std::vector<int> seg(A, m);
    for (int r = 0; r < m; ++r) sorted[r] = pts[order[r]];
    // seg[i] = first sorted point whose interval index is >= i
    int r = 0;
    for (int i = 0; i < A; ++i) {
        while (r < m && lo[order[r]] < i) ++r;
        seg[i] = r;
    }
    DeviceGuard guard(g->device);
    if (!guard.ok) return fail(HPEM_ERR_CUDA, "cannot select device %d", g->device);
    hpem_measurements* h = new (std::nothrow) hpem_measurements();
    if (!h) return fail(HPEM_ERR_CUDA, "out of host memory");
    h->device = g->device;
    h->m = m;
    h->n_angles = A;
    cudaError_t e = cudaMalloc((void**)&h->d_meas, sizeof(hpem::MeasPoint) * m);
    if (e == cudaSuccess) e = cudaMalloc((void**)&h->d_seg, sizeof(int) * A);
    if (e == cudaSuccess) e = cudaMemcpy(h->d_meas, sorted.data(), sizeof(hpem::MeasPoint) * m, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(h->d_seg, seg.data(), sizeof(int) * A, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        hpem_measurements_destroy(h);
        return fail(HPEM_ERR_CUDA, "measurement upload failed: %s", cudaGetErrorString(e));
    }
    *out = h;
    return HPEM_OK;
}

int hpem_measurements_destroy(hpem_measurements* h) {
    if (!h) return HPEM_OK;
    DeviceGuard guard(h->device);
    if (h->d_meas) cudaFree(h->d_meas);
    if (h->d_seg) cudaFree(h->d_seg);
    delete h;
    return HPEM_OK;
}

int hpem_logsumexp(int device, int64_t n_groups, int m, const double* loglike, double* out, void* stream) {
    if (!loglike || !out) return fail(HPEM_ERR_INVALID_ARG, "NULL argument");
    if (n_groups < 0 || m < 1) return fail(HPEM_ERR_INVALID_ARG, "need n_groups >= 0 and m >= 1");
    if (n_groups == 0) return HPEM_OK;
    DeviceGuard guard(device);
    if (!guard.ok) return fail(HPEM_ERR_CUDA, "cannot select device %d", device);
    const int64_t blocks = (n_groups + 3) / 4;
    if (blocks > 2147483647LL) return fail(HPEM_ERR_INVALID_ARG, "too many groups for one call");
    hpem::logsumexp_kernel<<<(unsigned)blocks, 128, 0, static_cast<cudaStream_t>(stream)>>>(loglike, n_groups, m, out);
    HPEM_CUDA(cudaGetLastError());
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return HPEM_OK;
}

int hpem_loglike(const hpem_grid* g, const hpem_measurements* meas, int64_t n, const hpem_inputs* in, double torr_2_pa,
                 double* loglike, double* y_pred, void* stream) {
    if (!g || !meas || !in) return fail(HPEM_ERR_INVALID_ARG, "NULL argument");
    if (!loglike && !y_pred) return fail(HPEM_ERR_INVALID_ARG, "no output requested");
    if (meas->device != g->device || meas->n_angles != g->n_angles)
        return fail(HPEM_ERR_INVALID_ARG, "measurement handle was created for another grid");
    if (n < 0) return fail(HPEM_ERR_INVALID_ARG, "negative sample count");
    if (n == 0) return HPEM_OK;
    DeviceGuard guard(g->device);
    if (!guard.ok) return fail(HPEM_ERR_CUDA, "cannot select device %d", g->device);
    using namespace hpem;
    hpem_outputs no_out = {};
    EvalParams p;
    fill_params(*g, *in, no_out, 0, n, torr_2_pa, p);
    LoglikeParams lp;
    lp.m = meas->m;
    lp.seg_start = meas->d_seg;
    lp.meas = meas->d_meas;
    lp.loglike = loglike;
    lp.y_pred = y_pred;
    const size_t smem = ((sizeof(MeasPoint) * meas->m + sizeof(int) * g->n_angles + 15) & ~size_t(15)) +
                        size_t(kThreadsL) * kRowL * sizeof(double);
    int rc = set_smem(loglike_kernel, smem);
    if (rc != HPEM_OK) return rc;
    const unsigned blocks = (unsigned)((n + kThreadsL - 1) / kThreadsL);
    loglike_kernel<<<blocks, kThreadsL, smem, static_cast<cudaStream_t>(stream)>>>(p, lp);
    HPEM_CUDA(cudaGetLastError());
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return HPEM_OK;
}

}  // extern "C"

// ---- SVD compression of the j_ion field (hpem_compress.cuh) ---------------------------------------------------------
struct hpem_basis {
    int device = 0;
    int dof = 0, rank = 0, rank_pad = 0, norm_log10 = 1;
    double* d_basis = nullptr;    // [dof][rank_pad]
    double* d_basis_t = nullptr;  // [rank][dof]
};

namespace {
hpem::BasisParams basis_params(const hpem_basis& b) {
    hpem::BasisParams bp;
    bp.dof = b.dof;
    bp.rank = b.rank;
    bp.rank_pad = b.rank_pad;
    bp.norm_log10 = b.norm_log10;
    bp.basis = b.d_basis;
    bp.basis_t = b.d_basis_t;
    return bp;
}
}  // namespace

extern "C" {

int hpem_basis_create(int device, int dof, int rank, const double* projection, int norm_log10, hpem_basis** out) {
    if (!out) return fail(HPEM_ERR_INVALID_ARG, "out handle pointer is NULL");
    *out = nullptr;
    if (!projection) return fail(HPEM_ERR_INVALID_ARG, "projection matrix is NULL");
    if (dof < 1 || dof > (1 << 20)) return fail(HPEM_ERR_INVALID_ARG, "dof must be in [1, 2^20], got %d", dof);
    if (rank < 1 || rank > hpem::kMaxRank) return fail(HPEM_ERR_INVALID_ARG, "rank must be in [1, %d], got %d", hpem::kMaxRank, rank);
    int ndev = 0;
    HPEM_CUDA(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(HPEM_ERR_INVALID_ARG, "device %d out of range (%d visible)", device, ndev);
    DeviceGuard guard(device);
    if (!guard.ok) return fail(HPEM_ERR_CUDA, "cannot select device %d", device);
    hpem_basis* b = new (std::nothrow) hpem_basis();
    if (!b) return fail(HPEM_ERR_CUDA, "out of host memory");
    b->device = device;
    b->dof = dof;
    b->rank = rank;
    b->rank_pad = (rank + 3) / 4 * 4;
    b->norm_log10 = norm_log10 ? 1 : 0;
    std::vector<double> padded(size_t(dof) * b->rank_pad, 0.0), transposed(size_t(dof) * rank);
    for (int i = 0; i < dof; ++i)
        for (int k = 0; k < rank; ++k) {
            padded[size_t(i) * b->rank_pad + k] = projection[size_t(i) * rank + k];
            transposed[size_t(k) * dof + i] = projection[size_t(i) * rank + k];
        }
    cudaError_t e = cudaMalloc((void**)&b->d_basis, padded.size() * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc((void**)&b->d_basis_t, transposed.size() * sizeof(double));
    if (e == cudaSuccess) e = cudaMemcpy(b->d_basis, padded.data(), padded.size() * sizeof(double), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(b->d_basis_t, transposed.data(), transposed.size() * sizeof(double), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        hpem_basis_destroy(b);
        return fail(HPEM_ERR_CUDA, "projection upload failed: %s", cudaGetErrorString(e));
    }
    *out = b;
    return HPEM_OK;
}

int hpem_basis_destroy(hpem_basis* b) {
    if (!b) return HPEM_OK;
    DeviceGuard guard(b->device);
    if (b->d_basis) cudaFree(b->d_basis);
    if (b->d_basis_t) cudaFree(b->d_basis_t);
    delete b;
    return HPEM_OK;
}

int hpem_compress(const hpem_grid* g, const hpem_basis* b, int64_t n, const hpem_inputs* in, double torr_2_pa, double* latent,
                  void* stream) {
    if (!g || !b || !in || !latent) return fail(HPEM_ERR_INVALID_ARG, "NULL argument");
    if (b->device != g->device) return fail(HPEM_ERR_INVALID_ARG, "basis and grid live on different devices");
    if (g->n_radii != 1 || !g->uniform) return fail(HPEM_ERR_UNSUPPORTED, "fused compression needs one radius and the uniform grid");
    if (b->dof != g->n_angles) return fail(HPEM_ERR_INVALID_ARG, "projection has %d rows, the grid has %d angles", b->dof, g->n_angles);
    if (n < 0) return fail(HPEM_ERR_INVALID_ARG, "negative sample count");
    if (n == 0) return HPEM_OK;
    DeviceGuard guard(g->device);
    if (!guard.ok) return fail(HPEM_ERR_CUDA, "cannot select device %d", g->device);
    using namespace hpem;
    hpem_outputs no_out = {};
    EvalParams p;
    fill_params(*g, *in, no_out, 0, n, torr_2_pa, p);
    const BasisParams bp = basis_params(*b);
    const int rk = b->rank <= 4 ? 4 : b->rank <= 8 ? 8 : b->rank <= 16 ? 16 : 32;
    const size_t smem = size_t(g->n_angles) * rk * sizeof(double);
    const unsigned blocks = (unsigned)((n + kThreadsC - 1) / kThreadsC);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int rc = HPEM_OK;
#define HPEM_LAUNCH_LATENT(RK)                                              \
    rc = set_smem(latent_kernel<RK>, smem);                                 \
    if (rc != HPEM_OK) return rc;                                           \
    latent_kernel<RK><<<blocks, kThreadsC, smem, st>>>(p, bp, latent)
    switch (rk) {
        case 4: HPEM_LAUNCH_LATENT(4); break;
        case 8: HPEM_LAUNCH_LATENT(8); break;
        case 16: HPEM_LAUNCH_LATENT(16); break;
        default: HPEM_LAUNCH_LATENT(32); break;
    }
#undef HPEM_LAUNCH_LATENT
    HPEM_CUDA(cudaGetLastError());
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return HPEM_OK;
}

int hpem_compress_field(const hpem_basis* b, int64_t n, const double* field, double* latent, void* stream) {
    if (!b || !field || !latent) return fail(HPEM_ERR_INVALID_ARG, "NULL argument");
    if (n < 0) return fail(HPEM_ERR_INVALID_ARG, "negative sample count");
    if (n == 0) return HPEM_OK;
    DeviceGuard guard(b->device);
    if (!guard.ok) return fail(HPEM_ERR_CUDA, "cannot select device %d", b->device);
    using namespace hpem;
    const int rows_per_block = kThreadsC / 32;
    const int64_t blocks = (n + rows_per_block - 1) / rows_per_block;
    if (blocks > 2147483647LL) return fail(HPEM_ERR_INVALID_ARG, "too many rows for one call");
    const BasisParams bp = basis_params(*b);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (b->rank <= 4) compress_field_kernel<4><<<(unsigned)blocks, kThreadsC, 0, st>>>(field, n, bp, latent);
    else if (b->rank <= 8) compress_field_kernel<8><<<(unsigned)blocks, kThreadsC, 0, st>>>(field, n, bp, latent);
    else if (b->rank <= 16) compress_field_kernel<16><<<(unsigned)blocks, kThreadsC, 0, st>>>(field, n, bp, latent);
    else compress_field_kernel<32><<<(unsigned)blocks, kThreadsC, 0, st>>>(field, n, bp, latent);
    HPEM_CUDA(cudaGetLastError());
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return HPEM_OK;
}

int hpem_reconstruct(const hpem_basis* b, int64_t n, const double* latent, double* field, void* stream) {
    if (!b || !field || !latent) return fail(HPEM_ERR_INVALID_ARG, "NULL argument");
    if (n < 0) return fail(HPEM_ERR_INVALID_ARG, "negative sample count");
    if (n == 0) return HPEM_OK;
    DeviceGuard guard(b->device);
    if (!guard.ok) return fail(HPEM_ERR_CUDA, "cannot select device %d", b->device);
    using namespace hpem;
    const int64_t blocks = (n + kReconSamples - 1) / kReconSamples;
    if (blocks > 2147483647LL) return fail(HPEM_ERR_INVALID_ARG, "too many rows for one call");
    const BasisParams bp = basis_params(*b);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (b->rank <= 4) reconstruct_kernel<4><<<(unsigned)blocks, kThreadsR, 0, st>>>(latent, n, bp, field);
    else if (b->rank <= 8) reconstruct_kernel<8><<<(unsigned)blocks, kThreadsR, 0, st>>>(latent, n, bp, field);
    else if (b->rank <= 16) reconstruct_kernel<16><<<(unsigned)blocks, kThreadsR, 0, st>>>(latent, n, bp, field);
    else reconstruct_kernel<32><<<(unsigned)blocks, kThreadsR, 0, st>>>(latent, n, bp, field);
    HPEM_CUDA(cudaGetLastError());
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return HPEM_OK;
}

}  // extern "C"
