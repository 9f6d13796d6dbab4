/*
 * hpem.h -- C ABI of libhpem: the B200 (sm_100a) plume + cathode Monte-Carlo hot path.
 *
 * This is the drop-in boundary for ONE path of JANUS-Institute/HallThrusterPEM (hallmd 0.3.0):
 *
 *     hallmd.models.cathode.cathode_coupling(inputs)                  src/hallmd/models/cathode.py:16-38
 *     hallmd.models.plume.current_density(inputs, sweep_radius=1.0)   src/hallmd/models/plume.py:21-159
 *
 * The reference has no FFI of its own (it is pure NumPy); the binding a maintainer adds is the ctypes
 * stub shown in INTEGRATION.md.  All entry points are `extern "C"`, take plain pointers and sizes,
 * return an int status (0 = ok, < 0 = error, text via hpem_last_error()), never throw, and never
 * change the reference's in-band numeric conventions (invalid samples -> 1e-20 fill, NaN propagation).
 *
 * Ownership: the caller owns every data buffer.  The library owns only the opaque `hpem_grid`
 * (angle grid, fused Simpson weights, radii; a few KB on the device) and, for the *_host entry
 * points, a device workspace cached inside the grid handle.
 *
 * Threading: entry points are re-entrant; a grid handle may be shared by threads for hpem_eval()
 * (read-only use) but hpem_eval_host() serialises on the handle's workspace mutex.
 */
#ifndef HPEM_H_
#define HPEM_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HPEM_ABI_VERSION 3

/* status codes */
#define HPEM_OK 0
#define HPEM_ERR_INVALID_ARG (-1)
#define HPEM_ERR_CUDA (-2)
#define HPEM_ERR_UNSUPPORTED (-3)

/* Named inputs, in the order of the `ptr` / `scalar` arrays of hpem_inputs.
 * cathode.py:26-31 reads the first six; plume.py:40-49 reads P_b and the last nine (T optional). */
enum hpem_input {
    HPEM_IN_P_b = 0,   /* background pressure, Torr          cathode.py:26, plume.py:40 */
    HPEM_IN_V_a,       /* discharge voltage, V               cathode.py:27 */
    HPEM_IN_T_e,       /* cathode electron temperature, eV   cathode.py:28 */
    HPEM_IN_V_vac,     /* vacuum coupling voltage, V         cathode.py:29 */
    HPEM_IN_Pstar,     /* P*, Torr                           cathode.py:30 */
    HPEM_IN_P_T,       /* P_T, Torr                          cathode.py:31 */
    HPEM_IN_c0,        /* plume.py:41 */
    HPEM_IN_c1,        /* plume.py:42 */
    HPEM_IN_c2,        /* plume.py:43 */
    HPEM_IN_c3,        /* plume.py:44 */
    HPEM_IN_c4,        /* plume.py:45 */
    HPEM_IN_c5,        /* plume.py:46 */
    HPEM_IN_sigma_cex, /* plume.py:47 */
    HPEM_IN_I_B0,      /* plume.py:48 */
    HPEM_IN_T,         /* thrust, N (optional)               plume.py:49 */
    HPEM_N_INPUTS
};

/* SoA view of one sample batch.  ptr[k] != NULL : n contiguous float64 values (device memory for
 * hpem_eval, host memory for hpem_eval_host).  ptr[k] == NULL : the value scalar[k] is broadcast to
 * every sample (NumPy scalar broadcasting, tests/test_plume.py:67-77).  Inputs a requested output
 * does not depend on are never read. */
typedef struct hpem_inputs {
    const double *ptr[HPEM_N_INPUTS];
    double scalar[HPEM_N_INPUTS];
} hpem_inputs;

/* Output buffers; any pointer may be NULL (= not wanted).  Shapes for n samples, A angles, R radii:
 *   V_cc      (n)         cathode.py:34-37      requested  => the six cathode inputs are read
 *   j_ion     (n, A, R)   plume.py:102-111      C-order, radius fastest; (n, A) when R == 1
 *   div_angle (n, R)      plume.py:127
 *   T_c       (n, R)      plume.py:137          requires input T
 *   cos_div   (n, R)      plume.py:124-125      extra (the reference keeps it internal)
 *   invalid   (n) uint8   plume.py:105          extra: the whole-sample invalid mask
 */
typedef struct hpem_outputs {
    double *V_cc;
    double *j_ion;
    double *div_angle;
    double *T_c;
    double *cos_div;
    uint8_t *invalid;
} hpem_outputs;

/* flags for hpem_eval / hpem_eval_host */
#define HPEM_FLAG_FORCE_DIRECT 1u /* use the direct kernel (one exp per beam per angle, the reference's  \
                                     operation order) even when the angle grid is uniform */
#define HPEM_FLAG_NO_TMA 2u       /* recurrence kernel: stage j_ion through plain st.global instead of TMA tensor   \
                                     stores (always the case for odd angle counts: rows are not 16-byte aligned) */

#define HPEM_FLAG_LANES1 4u       /* force the recurrence kernel with ONE lane per sample in the angle sweep (K1u,  \
                                     32x16 TMA boxes); default for even angle counts */
#define HPEM_FLAG_LANES4 8u       /* force the recurrence kernel with FOUR lanes per sample in the sweep (K1v,      \
                                     whole rows per bulk store); default for odd angle counts */

#define HPEM_FLAG_NO_QUAD 16u      /* angle counts that are not a multiple of 4: do not use K1u's quad-row tensor     \
                                     stores (diagnostics; falls back to (n, A) boxes / whole-row bulk stores) */

#define HPEM_FLAG_NO_FASTMATH 32u  /* per-sample part (cathode, normalisations, CEX terms, recurrence start values, arccos)  \
                                     through libdevice exp/log/acos and IEEE division even for warps whose samples are all \
                                     in the nominal range (default: the branch-free functions of csrc/hpem_fastmath.cuh;   \
                                     both back ends meet the rel-1e-12 parity rule, they differ in last bits) */

#define HPEM_FLAG_NO_QTABLE 64u    /* uniform grids: accumulate the two Simpson sums of plume.py:121-122 angle by angle instead of  \
                                     taking them from the grid's table (csrc/hpem_qtable.cuh; the two agree to ~4e-16) */

typedef struct hpem_grid hpem_grid; /* opaque */

int hpem_abi_version(void);
/* 16 hex digits: hash of the sources this binary was compiled from (the Python loader refuses a binary that does not
 * belong to its source tree, whatever the file's mtime says). */
const char *hpem_source_hash(void);
/* Thread-local text of the last error raised on the calling thread ("" if none). */
const char *hpem_last_error(void);

/* Build the per-(A, radii) constants on `device` (plume.py:50,53 and the Simpson weights of
 * plume.py:117-123 folded with flip/cos/sin):
 *   alpha[n_angles]  angle grid in radians (host)
 *   wd[n_angles], wn[n_angles]  un-flipped fused quadrature weights (host):
 *        den = sum_i wd[i] * (j_beam+j_scat)[i],  num = sum_i wn[i] * (j_beam+j_scat)[i]
 *   radii[n_radii]   sweep radii in metres (host)
 * If alpha is the uniform grid i*alpha[1] (as np.linspace(0, pi/2, A) is) the recurrence kernel is used. */
int hpem_grid_create(int device, int n_angles, const double *alpha, const double *wd, const double *wn,
                     int n_radii, const double *radii, hpem_grid **out);
int hpem_grid_destroy(hpem_grid *grid);
/* 1 if the grid was recognised as uniform (fast recurrence kernel), 0 otherwise, < 0 on error. */
int hpem_grid_is_uniform(const hpem_grid *grid);

/* Evaluate n samples with DEVICE buffers, asynchronously on `stream` (a cudaStream_t; NULL = legacy
 * default stream).  torr_2_pa is pem_core.constants.TORR_2_PA (un-vendored; 133.322 historically). */
int hpem_eval(const hpem_grid *grid, int64_t n, const hpem_inputs *in, const hpem_outputs *out,
              double torr_2_pa, uint32_t flags, void *stream);

/* Same with HOST buffers (pinned or pageable): chunked H2D -> kernel -> D2H pipeline on internal
 * streams; returns after all outputs have landed in host memory. */
int hpem_eval_host(hpem_grid *grid, int64_t n, const hpem_inputs *in, const hpem_outputs *out,
                   double torr_2_pa, uint32_t flags);

/* ---- reduce-only Monte-Carlo pass: sample moments and histograms without materialising j_ion -------------------
 * (the consumers of the reference's outputs take percentiles / moments over the sample axis:
 *  tests/test_plume.py:50-52, scripts/gen_data.py:402-404).  Single radius only.
 *
 * sums buffer (float64, length layout.n_sums; counts are stored as float64 so ONE ncclSum all-reduce merges ranks):
 *   [0] n_samples  [1] n_invalid (plume.py:105)  [2] n_nonfinite_rows (NaN/inf samples, excluded from angle sums)
 *   [3..5] V_cc: n_finite, sum, M2 = sum of squared deviations from the mean   [6..8] div_angle: same   [9..11] T_c: same
 *   [off_angle_sum + i]   sum over samples of j_ion[:, i]      (values exactly as current_density() returns them,
 *   [off_angle_sumsq + i] M2 of j_ion[:, i] over those samples  i.e. 1e-20 for invalid samples; n = [0] - [2])
 *        Second moments are CENTRED and merged pairwise (Chan et al.) block by block, call by call and rank by rank, so a
 *        variance M2/n never comes from E[x^2] - E[x]^2 over the population.  A packed vector is therefore NOT additive:
 *        merge vectors with hpem_moments_merge (or the same update on the host), not with a plain sum.
 *   [off_hist + a*n_bins + b] histogram count of j_ion[:, a*hist_angle_stride] in bin b:
 *        bin 0: j < 2^hist_min_exp2 (incl. <= 0);  bin n_bins-1: j >= 2^hist_max_exp2;  otherwise log-linear:
 *        b = 1 + floor((log2-octave - hist_min_exp2) * 2^hist_sub_bits + linear sub-bin within the octave)
 * minmax buffer (float64, length 6): (-min, max) of V_cc, div_angle, T_c.
 * Both buffers are ACCUMULATED into (call hpem_moments_accumulate once per chunk of samples); initialise sums to 0
 * and minmax to -inf.  Ranks: all-gather the [sums | minmax] vectors (ONE collective), then hpem_moments_merge. */
typedef struct hpem_moments_spec {
    int32_t hist_angle_stride; /* power of two; histogram every stride-th angle; 0 = no histograms */
    int32_t hist_sub_bits;     /* 2^sub_bits bins per octave, 0..6 */
    int32_t hist_min_exp2;     /* first octave */
    int32_t hist_max_exp2;     /* one past the last octave */
    int32_t want_cathode;      /* accumulate V_cc moments (reads the six cathode inputs) */
    int32_t want_thrust;       /* accumulate T_c moments (reads input T) */
    double scalar_shift[3];    /* V_cc, div_angle, T_c are accumulated about these values inside the kernel (any finite
                                  number near the expected mean; 0 is fine).  The packed vector does not depend on them
                                  beyond rounding. */
} hpem_moments_spec;

typedef struct hpem_moments_layout {
    int64_t n_sums;
    int64_t off_angle_sum;
    int64_t off_angle_sumsq;
    int64_t off_hist;
    int32_t n_hist_angles;
    int32_t n_bins;
    int32_t n_minmax; /* 6 */
    int32_t reserved;
} hpem_moments_layout;

int hpem_moments_layout_query(const hpem_grid *grid, const hpem_moments_spec *spec, hpem_moments_layout *layout);
/* DEVICE buffers, asynchronous on `stream`.  `sums` (layout.n_sums doubles) and `minmax` (6 doubles) are updated. */
int hpem_moments_accumulate(hpem_grid *grid, int64_t n, const hpem_inputs *in, double torr_2_pa,
                            const hpem_moments_spec *spec, double *sums, double *minmax, void *stream);

/* The Simpson sums of a Gaussian profile on the uniform grid (plume.py:117-123 with the fused weights above) as tabulated
 * functions of x = (h / alpha_beam)^2:  N_d(x) = sum_i wd[i] exp(-x i^2),  N_n(x) = sum_i wn[i] exp(-x i^2).  The reduce-only
 * pass uses this table (csrc/hpem_qtable.cuh) instead of accumulating the two sums angle by angle.  This entry point
 * evaluates it ON THE HOST with the arithmetic the kernel uses (no device needed), so the table can be checked against
 * direct sums anywhere: nd[k], nn[k] for x[k], k < n_x. */
int hpem_quadrature_table_eval(int n_angles, const double *wd, const double *wn, int64_t n_x, const double *x,
                               double *nd, double *nn);

/* Merge n_parts packed vectors, part r = [sums (layout.n_sums) | minmax (6)] at parts + r * part_stride (DEVICE memory),
 * in index order into sums / minmax (DEVICE, overwritten; must not alias parts).  Every rank runs this on the all-gathered
 * buffer, so all ranks end up with identical bits whatever order the collective moved the data in. */
int hpem_moments_merge(int device, const hpem_moments_layout *layout, int n_parts, const double *parts, int64_t part_stride,
                       double *sums, double *minmax, void *stream);

/* ---- on-device sampler for the input priors (the step before the path: amisc `sample_inputs`, gen_data.py:238) ----
 * Counter-based (Philox4x32-10): the inputs of global sample index i depend only on (seed, i), never on the shard,
 * chunk or GPU that draws them.  Stream: call t = 0..4 of sample i has counter (i lo, i hi, t, 0) and key (seed lo, seed hi)
 * and serves inputs 3t, 3t+1, 3t+2 (enum hpem_input order); uniform j of a call is k 2^-42 with the 42-bit integer
 * k = (bits 10j..10j+9 of output word 3) << 32 | output word j; Uniform: fma(u, b - a, a), LogUniform:
 * exp(fma(u, ln b - ln a, ln a)), Normal: Box-Muller with a second stream.  NumPy statement: oracle/sampler_oracle.py. */
#define HPEM_PRIOR_CONST 0      /* value a */
#define HPEM_PRIOR_UNIFORM 1    /* U(a, b)            yml: U(a,b), Uniform(a,b), Relative(p) around a nominal */
#define HPEM_PRIOR_LOGUNIFORM 2 /* LogUniform(a, b)   yml:253,261 */
#define HPEM_PRIOR_NORMAL 3     /* Normal(mean a, std b) */
typedef struct hpem_prior {
    int32_t kind;
    int32_t reserved;
    double a, b;
} hpem_prior;

/* Draw samples [first_index, first_index + n) of every input k with out[k] != NULL into DEVICE arrays (n doubles). */
int hpem_sample_inputs(int device, int64_t n, uint64_t seed, uint64_t first_index, const hpem_prior priors[HPEM_N_INPUTS],
                       double *const out[HPEM_N_INPUTS], void *stream);
/* Reduce-only pass whose inputs are drawn on the fly (no input arrays exist at all): equivalent to
 * hpem_sample_inputs + hpem_moments_accumulate, bit for bit. */
int hpem_moments_accumulate_sampled(hpem_grid *grid, int64_t n, uint64_t seed, uint64_t first_index,
                                    const hpem_prior priors[HPEM_N_INPUTS], double torr_2_pa,
                                    const hpem_moments_spec *spec, double *sums, double *minmax, void *stream);

/* ---- j_ion at measurement angles + Gaussian log-likelihood (the step after the path in the reference's calibration
 * scripts: mirror to (-90, 90) deg + linear interp1d, scripts/pem_v0/monte_carlo.py:265-270; sum of
 * -0.5 ((y - y_hat)/sigma)^2, scripts/pem_v0/mcmc.py:103).  Single radius, uniform grid.  j_ion is never materialised. */
typedef struct hpem_measurements hpem_measurements; /* opaque: sorted / pre-weighted probe points on the device */
/* theta[m] in radians, |theta| <= pi/2 (HPEM_ERR_INVALID_ARG otherwise, where interp1d raises); y[m], sigma[m] > 0: HOST arrays */
int hpem_measurements_create(const hpem_grid *grid, int m, const double *theta, const double *y, const double *sigma,
                             hpem_measurements **out);
int hpem_measurements_destroy(hpem_measurements *meas);
/* DEVICE buffers, asynchronous on `stream`: loglike (n) and/or y_pred (n, m) in the caller's point order (NULL = skip) */
int hpem_loglike(const hpem_grid *grid, const hpem_measurements *meas, int64_t n, const hpem_inputs *in,
                 double torr_2_pa, double *loglike, double *y_pred, void *stream);

/* Marginal log-likelihood over the trailing axis of m Monte-Carlo draws per calibration vector:
 * out[g] = max_m ll[g, m] + log(sum_m exp(ll[g, m] - max))  (scripts/pem_v0/mcmc.py:101-102).  DEVICE buffers:
 * loglike (n_groups, m), out (n_groups). */
int hpem_logsumexp(int device, int64_t n_groups, int m, const double *loglike, double *out, void *stream);

/* ---- SVD compression of the j_ion field quantity (the data format downstream of the path: amisc normalises j_ion
 * with log10 and keeps only its projection on the leading left-singular vectors of a compression sample set,
 * scripts/pem_v0/pem_v0_SPT-100.yml:272-280, scripts/gen_data.py:279-290; amisc itself is un-vendored, uv.lock:14-16).
 *   latent = U^T x,  x = log10(j_ion)  (norm_log10 = 1)  or  x = j_ion  (norm_log10 = 0);   j_ion = 10^(U z). */
typedef struct hpem_basis hpem_basis; /* opaque: projection matrix U on the device */
/* projection: HOST array (dof, rank) row-major = amisc's `SVD.projection_matrix`; 1 <= rank <= 32 */
int hpem_basis_create(int device, int dof, int rank, const double *projection, int norm_log10, hpem_basis **out);
int hpem_basis_destroy(hpem_basis *basis);
/* Fused plume model + normalisation + projection: latent (n, rank) DEVICE; j_ion is never materialised.
 * Needs dof == n_angles, one radius and the uniform grid.  Asynchronous on `stream`. */
int hpem_compress(const hpem_grid *grid, const hpem_basis *basis, int64_t n, const hpem_inputs *in, double torr_2_pa,
                  double *latent, void *stream);
/* Projection of a materialised field (n, dof) DEVICE -> latent (n, rank) DEVICE. */
int hpem_compress_field(const hpem_basis *basis, int64_t n, const double *field, double *latent, void *stream);
/* latent (n, rank) DEVICE -> field (n, dof) DEVICE. */
int hpem_reconstruct(const hpem_basis *basis, int64_t n, const double *latent, double *field, void *stream);

/* Number of kernel launches issued by this process through the library (for bench accounting). */
int64_t hpem_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* HPEM_H_ */
