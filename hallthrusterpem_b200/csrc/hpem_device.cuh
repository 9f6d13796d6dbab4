// hpem_device.cuh -- per-sample device functions of the plume + cathode path (sm_100a, fp64).
//
// Reference arithmetic being reproduced (file:line under /root/reference/src/hallmd/models):
//   cathode.py:26-37   cathode coupling voltage
//   plume.py:40,56-61  unit conversion, neutral density, divergence angles
//   plume.py:64-85     beam normalisations A1, A2 (closed form with complex erfi)
//   plume.py:95-98     charge-exchange attenuation
// Separately-rounded multiplies/adds of the reference are kept separately rounded here
// (__dmul_rn/__dadd_rn are never contracted into FMAs) wherever the result feeds a cancellation.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>

#include "hpem_dtable.inc"

namespace hpem {

constexpr double kPi = 3.14159265358979323846;
constexpr double kHalfPi = 1.57079632679489661923;
// Largest x with finite exp(x) in fp64.  scipy's erfi(a/2) evaluates exp((a/2)^2) internally, so the
// reference's normalisation is non-finite (-> NaN amplitudes, plume.py:64-85) once (a/2)^2 exceeds it.
constexpr double kExpOverflow = 0x1.62e42fefa39efp+9;

// 1/y to ~1 ulp for finite y >= 2 (table index only; not used where the reference divides)
__device__ __forceinline__ double rcp_fast(double y) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(y));
    r = fma(r, fma(-y, r, 1.0), r);
    r = fma(r, fma(-y, r, 1.0), r);
    return r;
}

// numer / D(a) with D(a) = 2 pi int_0^{pi/2} exp(-(t/a)^2) sin t dt, the denominator of A1/A2 (plume.py:64-85).
// D is even in a; the result is NaN at a == 0, for NaN input and in the reference's erfi-overflow domain |a| > 53.2835.
// D = q(u) * 2 pi a^2 / (a^2 + 2) (table, see tools/gen_dtable.py), so numer/D = numer (a^2 + 2) / (q 2 pi a^2): one division.
__device__ __forceinline__ double beam_amplitude(double numer, double a) {
    const double aa = fabs(a);
    const double half = aa * 0.5;
    const double hx = half * half;
    const double u = aa * rcp_fast(aa + 2.0);
    const double v = u * (double(HPEM_DTAB_M) / HPEM_DTAB_UMAX);
    int idx = __double2int_rz(v);  // NaN -> 0
    idx = min(max(idx, 0), HPEM_DTAB_M - 1);
    const double t = fma(2.0, v - double(idx), -1.0);
    const double* __restrict__ c = hpem_dtab[idx];
    double q = c[0];
#pragma unroll
    for (int j = 1; j <= HPEM_DTAB_DEG; ++j) q = fma(q, t, c[j]);
    const double a2 = aa * aa;
    double amp = (numer * (a2 + 2.0)) / (q * ((2.0 * kPi) * a2));
    if (!(hx <= kExpOverflow) || aa == 0.0) amp = CUDART_NAN;
    return amp;
}

// cathode.py:26-37.  log(1 + x), not log1p, and the reference's rounding sequence.
__device__ __forceinline__ double cathode_vcc(double p_b, double v_a, double t_e, double v_vac, double p_star,
                                              double p_t, double torr) {
    const double PB = __dmul_rn(p_b, torr);
    const double PS = __dmul_rn(p_star, torr);
    const double PT = __dmul_rn(p_t, torr);
    const double lg = log(__dadd_rn(1.0, PB / PT));
    const double t1 = __dmul_rn(t_e, lg);
    const double t2 = __dmul_rn(t_e / __dadd_rn(PT, PS), PB);
    double v = __dadd_rn(__dadd_rn(v_vac, t1), -t2);
    if (v < 0.0) v = 0.0;     // NaN passes through both clamps (cathode.py:35-37)
    if (v > v_a) v = v_a;
    return v;
}

// exp(-t) whose value is (almost always) the correctly rounded one for small t: j_cex is proportional to
// 1 - exp(-t) (plume.py:96) and amplifies a 1-ulp difference in `decay` by 1/t.  For t < 0.25 form
// 1 + expm1(-t) with a single final rounding.
__device__ __forceinline__ double decay_exp(double neg_t) {
    if (neg_t > -0.25 && neg_t < 0.25) return __dadd_rn(1.0, expm1(neg_t));
    return exp(neg_t);
}

struct SampleConsts {
    double a1, a2;      // divergence angles (plume.py:59-61)
    double amp1, amp2;  // A1, A2 (plume.py:64-85)
    double density;     // n (plume.py:56)
};

__device__ __forceinline__ SampleConsts plume_sample_consts(double p_b, double c0, double c1, double c2, double c3,
                                                            double c4, double c5, double torr) {
    SampleConsts k;
    const double PB = __dmul_rn(p_b, torr);                // plume.py:40
    k.density = __dadd_rn(__dmul_rn(c4, PB), c5);          // plume.py:56
    double a1 = __dadd_rn(__dmul_rn(c2, PB), c3);          // plume.py:59
    if (a1 > kHalfPi) a1 = kHalfPi;                        // plume.py:60 (no lower clip; NaN unchanged)
    k.a1 = a1;
    k.a2 = a1 / c1;                                        // plume.py:61
    k.amp1 = beam_amplitude(__dadd_rn(1.0, -c0), k.a1);    // plume.py:64-76
    k.amp2 = beam_amplitude(c0, k.a2);                     // plume.py:77-85
    return k;
}

// plume.py:95-98 for one radius: decay, j_cex, base = I_B0*decay/r^2
__device__ __forceinline__ void cex_terms(double density, double sigma, double i_b0, double r, double& j_cex,
                                          double& base) {
    const double arg = __dmul_rn(__dmul_rn(-r, density), sigma);                       // (-r*n)*sigma
    const double decay = decay_exp(arg);
    const double r2 = __dmul_rn(r, r);
    j_cex = __dmul_rn(i_b0, __dadd_rn(1.0, -decay)) / __dmul_rn(2.0 * kPi, r2);        // plume.py:96
    base = __dmul_rn(i_b0, decay) / r2;                                                // plume.py:98
}

}  // namespace hpem
