// hpem_qtable.cuh -- the Simpson sums of a Gaussian beam profile as a tabulated function of ONE variable.
//
// plume.py:117-123 integrates (j_beam + j_scat) cos / cos sin over the angle grid.  With the fused weights of
// quadrature.py (flip, cos, sin and the Simpson / Cartwright coefficients folded into wd_i, wn_i) and the uniform grid
// alpha_i = i h, each beam contributes  amplitude * N(x),  x = (h / alpha_beam)^2,  with
//
//        N_d(x) = sum_i wd_i exp(-x i^2),        N_n(x) = sum_i wn_i exp(-x i^2)
//
// -- two smooth functions of x alone, fixed once the grid is.  The reduce-only kernel (hpem_moments.cuh) spends 2 of its
// ~10 fp64 instructions per (sample, angle) evaluation on these sums; this table replaces them by one lookup per beam:
//
//        N(x) = w_0 + exp(-x) M(x),      M(x) = sum_{i>=1} w_i exp(-x (i^2 - 1))
//
// (exp(-x) is the recurrence's own start value, so it is free; factoring it out keeps M between w_1 and sum w_i, so the
// RELATIVE accuracy of N holds for needle beams whose sums decay like exp(-x)).  M is tabulated on log-linear bins of x
// -- bin index straight from the exponent and the leading kQtSubBits mantissa bits of the fp64 pattern, no logarithm --
// as degree-kQtDeg polynomials in t = x / x_centre - 1 (|t| <= 1/33), Chebyshev-interpolated in long double on the
// host.  Measured against long-double sums: 2.5e-16 relative for 91..512 angles (tests/test_host_cpu.py), i.e. better
// than the rounding of the A-term fused-multiply-add sums it replaces.  Below the first bin M is linear in x to 1e-17,
// above the last (x >= 16) it is w_1 to 1e-20: the bin index is clamped, nothing is ever out of range.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace hpem {

constexpr int kQtSubBits = 4;                 // 16 bins per octave of x
constexpr int kQtDeg = 8;                     // polynomial degree
constexpr int kQtRow = 20;                    // doubles per bin: 1/x_centre, d_0..d_8 (M_d), n_0..n_8 (M_n), pad -> 160 B = 5 sectors
constexpr int kQtHiExp2 = 4;                  // regular bins end at x = 2^4

struct QTableRef {                            // what a kernel needs to evaluate the table
    const double* rows;                       // [n_bins][kQtRow]
    int key_lo;                               // bin = clamp((hi32(x) >> (20 - kQtSubBits)) - key_lo, 0, n_bins - 1)
    int n_bins;
    double wd0, wn0;                          // the i = 0 weights
};

// first regular octave: below 2^lo the quadratic term of M, x^2 sum w_i i^4 / 2 <= (x A^2)^2 / 2 sum w_i, is < 2^-61 of M
inline int qtable_lo_exp2(int n_angles) {
    int l2 = 0;
    while ((1 << l2) < n_angles) ++l2;
    return -(30 + 2 * l2);
}

// Host: build the table for the fused weights (wd_i, wn_i), i < n_angles.  key_lo / n_bins as in QTableRef.
inline void qtable_build(int n_angles, const double* wd, const double* wn, std::vector<double>& rows, int& key_lo, int& n_bins) {
    typedef long double ld;
    const int lo = qtable_lo_exp2(n_angles), nsub = 1 << kQtSubBits, nn = kQtDeg + 1;
    const int n_regular = (kQtHiExp2 - lo) * nsub;
    n_bins = n_regular + 2;
    key_lo = ((lo + 1023) << kQtSubBits) - 1;
    rows.assign(size_t(n_bins) * kQtRow, 0.0);
    const ld pi = 3.14159265358979323846264338327950288L;
    // Chebyshev nodes, T_k at the nodes, and the Chebyshev -> monomial matrix
    ld s[kQtDeg + 1], tk[kQtDeg + 1][kQtDeg + 1], mono[kQtDeg + 1][kQtDeg + 1];
    for (int j = 0; j < nn; ++j) {
        const ld th = pi * (j + 0.5L) / nn;
        s[j] = cosl(th);
        for (int k = 0; k < nn; ++k) tk[k][j] = cosl(k * th);
    }
    for (int k = 0; k < nn; ++k)
        for (int j = 0; j < nn; ++j) mono[k][j] = 0.0L;
    mono[0][0] = 1.0L;
    mono[1][1] = 1.0L;
    for (int k = 2; k < nn; ++k)
        for (int j = 0; j < nn; ++j) mono[k][j] = (j > 0 ? 2.0L * mono[k - 1][j - 1] : 0.0L) - mono[k - 2][j];
    std::vector<ld> kk(n_angles);            // i^2 - 1
    for (int i = 1; i < n_angles; ++i) kk[i] = ld(i) * ld(i) - 1.0L;
    // terms with x (i^2 - 1) > 50 are below 2e-22 of the i = 1 term: the sums stop at i_max(x)
    auto m_sums = [&](ld x, ld& md, ld& mn) {
        md = mn = 0.0L;
        int i_max = n_angles - 1;
        const ld lim = 50.0L / x + 1.0L;
        if (lim < ld(i_max) * ld(i_max)) i_max = int(sqrtl(lim)) + 1;
        if (i_max > n_angles - 1) i_max = n_angles - 1;
        for (int i = i_max; i >= 1; --i) {             // small terms first
            const ld e = expl(-x * kk[i]);
            md += ld(wd[i]) * e;
            mn += ld(wn[i]) * e;
        }
    };
    // small x: M(x) = sum_k mu_k x^k, mu_k = (-1)^k / k! sum_i w_i (i^2 - 1)^k; with x A^2 < 2^-8 twelve terms reach 2^-96
    constexpr int kTaylor = 12;
    ld mu_d[kTaylor], mu_n[kTaylor];
    {
        std::vector<ld> pw(n_angles, 1.0L);
        ld fact = 1.0L;
        for (int k = 0; k < kTaylor; ++k) {
            if (k > 0) fact *= -ld(k);               // (-1)^k k!
            ld sd = 0.0L, sn = 0.0L;
            for (int i = n_angles - 1; i >= 1; --i) {
                sd += ld(wd[i]) * pw[i];
                sn += ld(wn[i]) * pw[i];
                pw[i] *= kk[i];
            }
            mu_d[k] = sd / fact;
            mu_n[k] = sn / fact;
        }
    }
    const ld a2 = ld(n_angles) * ld(n_angles);
    for (int b = 0; b < n_regular; ++b) {
        const int e2 = lo + b / nsub, m = b % nsub;
        const ld width = ldexpl(1.0L, e2) / nsub, xa = ldexpl(1.0L, e2) + m * width, xc = xa + 0.5L * width, delta = 0.5L * width / xc;
        double* row = rows.data() + size_t(b + 1) * kQtRow;
        row[0] = double(1.0L / xc);
        if ((xa + width) * a2 < 0.00390625L) {
            // M(xc (1 + t)) = sum_k mu_k xc^k (1 + t)^k  ->  coefficient of t^j = sum_{k >= j} mu_k xc^k C(k, j)
            for (int f = 0; f < 2; ++f) {
                const ld* mu = f ? mu_n : mu_d;
                for (int j = 0; j < nn; ++j) {
                    ld c = 0.0L, xk = powl(xc, j), binom = 1.0L;      // C(j, j)
                    for (int k = j; k < kTaylor; ++k) {
                        c += mu[k] * xk * binom;
                        xk *= xc;
                        binom = binom * ld(k + 1) / ld(k + 1 - j);   // C(k+1, j)
                    }
                    row[1 + f * nn + j] = double(c);
                }
            }
            continue;
        }
        ld fd[kQtDeg + 1], fn[kQtDeg + 1];
        for (int j = 0; j < nn; ++j) m_sums(xc * (1.0L + delta * s[j]), fd[j], fn[j]);
        for (int f = 0; f < 2; ++f) {
            const ld* fv = f ? fn : fd;
            ld cheb[kQtDeg + 1], poly[kQtDeg + 1];
            for (int k = 0; k < nn; ++k) {
                ld a = 0.0L;
                for (int j = 0; j < nn; ++j) a += fv[j] * tk[k][j];
                cheb[k] = a * (k == 0 ? 1.0L : 2.0L) / nn;
            }
            for (int j = 0; j < nn; ++j) {
                ld a = 0.0L;
                for (int k = 0; k < nn; ++k) a += cheb[k] * mono[k][j];
                poly[j] = a;
            }
            ld scale = 1.0L;
            for (int j = 0; j < nn; ++j) {
                row[1 + f * nn + j] = double(poly[j] * scale);
                scale /= delta;
            }
        }
    }
    // bin 0: x below the first regular bin, M(x) = M(0) + x M'(0) with t = x / x_lo - 1 in [-1, 0)
    {
        const ld x_lo = ldexpl(1.0L, lo);
        ld m0d = 0.0L, m0n = 0.0L, m1d = 0.0L, m1n = 0.0L;
        for (int i = n_angles - 1; i >= 1; --i) {
            m0d += ld(wd[i]);
            m0n += ld(wn[i]);
            m1d -= ld(wd[i]) * kk[i];
            m1n -= ld(wn[i]) * kk[i];
        }
        double* row = rows.data();
        row[0] = double(1.0L / x_lo);
        row[1] = double(m0d + x_lo * m1d);
        row[2] = double(x_lo * m1d);
        row[1 + nn] = double(m0n + x_lo * m1n);
        row[2 + nn] = double(x_lo * m1n);
    }
    // last bin: x >= 2^kQtHiExp2, M = w_1 (the next term is w_2 exp(-3 x) < 1e-20 w_2)
    {
        double* row = rows.data() + size_t(n_bins - 1) * kQtRow;
        row[0] = std::ldexp(1.0, -kQtHiExp2);
        row[1] = n_angles > 1 ? wd[1] : 0.0;
        row[1 + nn] = n_angles > 1 ? wn[1] : 0.0;
    }
}

#if defined(__CUDACC__)
#define HPEM_QT_HD __host__ __device__ __forceinline__
#else
#define HPEM_QT_HD inline
#endif

// N_d(x), N_n(x) from the table; ex = exp(-x).  Identical arithmetic on host and device (explicit fma, fixed order).
HPEM_QT_HD void qtable_eval(const QTableRef& q, double x, double ex, double& nd, double& nn) {
#if defined(__CUDA_ARCH__)
    const int hi = __double2hiint(x);
#else
    int64_t bits;
    std::memcpy(&bits, &x, sizeof(bits));
    const int hi = int(bits >> 32);
#endif
    int bin = (hi >> (20 - kQtSubBits)) - q.key_lo;
    bin = bin < 0 ? 0 : (bin > q.n_bins - 1 ? q.n_bins - 1 : bin);
    double r[kQtRow];
#if defined(__CUDA_ARCH__)
    const double2* row = reinterpret_cast<const double2*>(q.rows) + size_t(bin) * (kQtRow / 2);
#pragma unroll
    for (int k = 0; k < kQtRow / 2; ++k) {
        const double2 v = __ldg(row + k);
        r[2 * k] = v.x;
        r[2 * k + 1] = v.y;
    }
#else
    for (int k = 0; k < kQtRow; ++k) r[k] = q.rows[size_t(bin) * kQtRow + k];
#endif
    const double t = fma(x, r[0], -1.0);
    double md = r[1 + kQtDeg], mn = r[2 + 2 * kQtDeg];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int k = kQtDeg - 1; k >= 0; --k) {
        md = fma(md, t, r[1 + k]);
        mn = fma(mn, t, r[2 + kQtDeg + k]);
    }
    nd = fma(ex, md, q.wd0);
    nn = fma(ex, mn, q.wn0);
}

}  // namespace hpem
