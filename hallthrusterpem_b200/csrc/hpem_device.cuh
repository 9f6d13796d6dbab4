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
#include "hpem_fastmath.cuh"

namespace hpem {

// Two arithmetic back ends for the per-sample part.  FAST = false: libdevice (exp, log, acos, IEEE division with their
// special-case branches) -- used for any warp that holds a sample outside the nominal ranges, and by the direct kernel.
// FAST = true: the branch-free functions of hpem_fastmath.cuh -- used when `prologue_nominal()` holds for all 32 lanes.
template <bool FAST> __device__ __forceinline__ double m_div(double a, double b) { return FAST ? fm_div(a, b) : a / b; }
template <bool FAST> __device__ __forceinline__ double m_exp(double x) { return FAST ? fm_exp(x) : exp(x); }
template <bool FAST> __device__ __forceinline__ double m_log(double x) { return FAST ? fm_log(x) : log(x); }
template <bool FAST> __device__ __forceinline__ double m_acos(double x) { return FAST ? fm_acos(x) : acos(x); }

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
template <bool FAST = false>
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
    double amp = m_div<FAST>(numer * (a2 + 2.0), q * ((2.0 * kPi) * a2));
    if (!(hx <= kExpOverflow) || aa == 0.0) amp = CUDART_NAN;
    return amp;
}

// cathode.py:26-37.  log(1 + x), not log1p, and the reference's rounding sequence.
template <bool FAST = false>
__device__ __forceinline__ double cathode_vcc(double p_b, double v_a, double t_e, double v_vac, double p_star,
                                              double p_t, double torr) {
    const double PB = __dmul_rn(p_b, torr);
    const double PS = __dmul_rn(p_star, torr);
    const double PT = __dmul_rn(p_t, torr);
    const double lg = m_log<FAST>(__dadd_rn(1.0, m_div<FAST>(PB, PT)));
    const double t1 = __dmul_rn(t_e, lg);
    const double t2 = __dmul_rn(m_div<FAST>(t_e, __dadd_rn(PT, PS)), PB);
    double v = __dadd_rn(__dadd_rn(v_vac, t1), -t2);
    if (v < 0.0) v = 0.0;     // NaN passes through both clamps (cathode.py:35-37)
    if (v > v_a) v = v_a;
    return v;
}

// exp(-t) whose value is (almost always) the correctly rounded one for small t: j_cex is proportional to
// 1 - exp(-t) (plume.py:96) and amplifies a 1-ulp difference in `decay` by 1/t.  fm_exp forms 1 + (r + r^2 g(r)) with a
// single final rounding (r = -t for t < 0.34), is valid for every argument (NaN, +-inf, subnormal results) and has no
// branch; ALL kernels and both arithmetic back ends use it, so they agree on `decay` bit for bit.
__device__ __forceinline__ double decay_exp(double neg_t) { return fm_exp(neg_t); }

struct SampleConsts {
    double a1, a2;      // divergence angles (plume.py:59-61)
    double amp1, amp2;  // A1, A2 (plume.py:64-85)
    double density;     // n (plume.py:56)
};

template <bool FAST = false>
__device__ __forceinline__ SampleConsts plume_sample_consts(double p_b, double c0, double c1, double c2, double c3,
                                                            double c4, double c5, double torr) {
    SampleConsts k;
    const double PB = __dmul_rn(p_b, torr);                // plume.py:40
    k.density = __dadd_rn(__dmul_rn(c4, PB), c5);          // plume.py:56
    double a1 = __dadd_rn(__dmul_rn(c2, PB), c3);          // plume.py:59
    if (a1 > kHalfPi) a1 = kHalfPi;                        // plume.py:60 (no lower clip; NaN unchanged)
    k.a1 = a1;
    k.a2 = m_div<FAST>(a1, c1);                            // plume.py:61
    k.amp1 = beam_amplitude<FAST>(__dadd_rn(1.0, -c0), k.a1);    // plume.py:64-76
    k.amp2 = beam_amplitude<FAST>(c0, k.a2);                     // plume.py:77-85
    return k;
}

// plume.py:95-98 for one radius: decay, j_cex, base = I_B0*decay/r^2
template <bool FAST = false>
__device__ __forceinline__ void cex_terms(double density, double sigma, double i_b0, double r, double& j_cex,
                                          double& base) {
    const double arg = __dmul_rn(__dmul_rn(-r, density), sigma);                       // (-r*n)*sigma
    const double decay = decay_exp(arg);
    const double r2 = __dmul_rn(r, r);
    j_cex = m_div<FAST>(__dmul_rn(i_b0, __dadd_rn(1.0, -decay)), __dmul_rn(2.0 * kPi, r2));   // plume.py:96
    base = m_div<FAST>(__dmul_rn(i_b0, decay), r2);                                            // plume.py:98
}

// Is this sample inside the range where the branch-free back end is exact?  Every divisor and every numerator that
// reaches fm_div / fm_log must be a normal number of moderate magnitude (binary exponent within +-100 of 1: all physical
// inputs are within 1e-20 .. 1e22), pressures non-negative (so log(1 + PB/PT) has an argument >= 1), alpha1 > 0 (rows with
// alpha1 <= 0 are the 1e-20 fill, plume.py:105), and the CEX exponent in [-600, 0] (decay is a normal number).  Anything
// else -- NaN/inf inputs, zeros in a divisor, subnormal products, the fuzz tests' extreme rows -- takes the libdevice
// path, whose IEEE special cases are the reference's.  `radius` is the sweep radius (all radii for several).
__device__ __forceinline__ bool fm_mid(double x) {          // finite, exponent within +-100
    const unsigned e = ((unsigned)__double2hiint(x) >> 20) & 0x7ffu;
    return (e - (1023u - 100u)) <= 200u;
}
__device__ __forceinline__ bool fm_mid0(double x) { return fm_mid(x) || x == 0.0; }
__device__ __forceinline__ bool prologue_nominal(const double* x_in, double torr, bool want_cathode, bool want_plume,
                                                 double radius) {
    // indices follow enum hpem_input (include/hpem.h)
    const double PB = __dmul_rn(x_in[0], torr);
    bool ok = fm_mid(torr) && fm_mid0(PB) && PB >= 0.0;
    if (want_cathode) {
        const double PT = __dmul_rn(x_in[5], torr);
        ok = ok && fm_mid(PT) && PT > 0.0 && fm_mid(__dadd_rn(PT, __dmul_rn(x_in[4], torr))) && fm_mid0(x_in[2]);
    }
    if (want_plume) {
        const double a1 = __dadd_rn(__dmul_rn(x_in[8], PB), x_in[9]);
        ok = ok && fm_mid(radius) && radius > 0.0 && fm_mid(a1) && a1 > 0.0 && fm_mid(x_in[7]) && fm_mid0(x_in[6]) && fm_mid0(x_in[13]);
        const double arg = __dmul_rn(__dmul_rn(-radius, __dadd_rn(__dmul_rn(x_in[10], PB), x_in[11])), x_in[12]);
        ok = ok && arg >= -600.0 && arg <= 0.0;
    }
    return ok;
}

}  // namespace hpem
