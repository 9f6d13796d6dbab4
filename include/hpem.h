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

#define HPEM_ABI_VERSION 1

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

typedef struct hpem_grid hpem_grid; /* opaque */

int hpem_abi_version(void);
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

/* Number of kernel launches issued by this process through the library (for bench accounting). */
int64_t hpem_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* HPEM_H_ */
