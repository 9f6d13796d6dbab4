// CPU accuracy check of hpem_fastmath.cuh (host build of the same source the kernels use; the hardware reciprocal /
// rsqrt seeds are emulated with 20-bit truncations).  Reference: x87 long double (64-bit mantissa) libm.
//   g++ -O2 -mfma -ffp-contract=off -o /tmp/fastmath_check tools/fastmath_check.cpp && /tmp/fastmath_check
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#ifndef HPEM_CHECK_N
#define HPEM_CHECK_N 2000000
#endif
#include "../hallthrusterpem_b200/csrc/hpem_fastmath.cuh"

static double ulp_err(double got, long double want) {
    if (std::isnan(got) && std::isnan((double)want)) return 0.0;
    if (want == 0.0L) return got == 0.0 ? 0.0 : 1e9;
    int e; std::frexp((double)want, &e);
    const long double ulp = std::ldexp(1.0L, e - 53 < -1074 ? -1074 : e - 53);
    return (double)(std::fabs((long double)got - want) / ulp);
}

int main() {
    std::mt19937_64 rng(12345);
    std::uniform_real_distribution<double> U(0.0, 1.0);
    double worst;
    int fails = 0;
    // exp: all regimes
    const double ranges[][2] = {{-1e-6, 0}, {-1e-3, 0}, {-0.3, 0}, {-2, 0}, {-40, 0}, {-700, 0}, {0, 2}, {0, 700}, {-745, -700}};
    for (auto& rg : ranges) {
        worst = 0;
        for (int i = 0; i < HPEM_CHECK_N; ++i) {
            const double x = rg[0] + (rg[1] - rg[0]) * U(rng);
            const double e = ulp_err(hpem::fm_exp(x), expl((long double)x));
            if (e > worst) worst = e;
        }
        printf("fm_exp  x in [%g, %g]: worst %.4f ulp\n", rg[0], rg[1], worst);
        if (worst > (rg[0] < -708 ? 1.6 : 1.0)) ++fails;
    }
    // how often is 1 - fm_exp(-t) different from 1 - correctly rounded exp(-t), small t (plume.py:96)
    {
        long bad = 0; const int N = 2 * HPEM_CHECK_N;
        for (int i = 0; i < N; ++i) {
            const double t = -std::pow(10.0, -6.0 + 5.5 * U(rng));
            if (hpem::fm_exp(t) != (double)expl((long double)t)) ++bad;
        }
        printf("fm_exp  small |t| in [1e-6, 0.3]: %ld of %d not correctly rounded\n", bad, N);
        if (bad > N / 50) ++fails;
    }
    printf("fm_exp(-inf)=%g fm_exp(-1e4)=%g fm_exp(710)=%g fm_exp(nan)=%g fm_exp(0)=%g\n", hpem::fm_exp(-INFINITY), hpem::fm_exp(-1e4),
           hpem::fm_exp(710.0), hpem::fm_exp(NAN), hpem::fm_exp(0.0));
    if (hpem::fm_exp(-INFINITY) != 0.0 || !std::isinf(hpem::fm_exp(710.0)) || !std::isnan(hpem::fm_exp(NAN)) || hpem::fm_exp(0.0) != 1.0) ++fails;
    // div
    worst = 0;
    for (int i = 0; i < 2 * HPEM_CHECK_N; ++i) {
        const double a = std::ldexp(1.0 + U(rng), (int)(U(rng) * 300) - 150) * (U(rng) < 0.5 ? -1 : 1);
        const double b = std::ldexp(1.0 + U(rng), (int)(U(rng) * 300) - 150) * (U(rng) < 0.5 ? -1 : 1);
        const double e = ulp_err(hpem::fm_div(a, b), (long double)a / (long double)b);
        if (e > worst) worst = e;
    }
    printf("fm_div: worst %.4f ulp;  0/3 = %g\n", worst, hpem::fm_div(0.0, 3.0));
    if (worst > 0.5001 || hpem::fm_div(0.0, 3.0) != 0.0) ++fails;
    // log on [1, 1e6] (cathode.py:34: log(1 + PB/PT)) and a wide range
    const double lr[][2] = {{0, 1e-6}, {0, 1e-2}, {0, 1}, {0, 1e6}};
    for (auto& rg : lr) {
        worst = 0;
        for (int i = 0; i < HPEM_CHECK_N; ++i) {
            const double y = 1.0 + rg[1] * U(rng);
            const double e = ulp_err(hpem::fm_log(y), logl((long double)y));
            if (e > worst) worst = e;
        }
        printf("fm_log  y in 1 + [0, %g]: worst %.4f ulp\n", rg[1], worst);
        if (worst > 1.0) ++fails;
    }
    worst = 0;
    for (int i = 0; i < HPEM_CHECK_N; ++i) {
        const double y = std::ldexp(1.0 + U(rng), (int)(U(rng) * 380) - 190);
        const double e = ulp_err(hpem::fm_log(y), logl((long double)y));
        if (e > worst) worst = e;
    }
    printf("fm_log  wide: worst %.4f ulp; log(1) = %g\n", worst, hpem::fm_log(1.0));
    if (worst > 1.0 || hpem::fm_log(1.0) != 0.0) ++fails;
    // acos
    const double ar[][2] = {{-1, 1}, {0, 0.5}, {0.5, 1}, {0.999, 1}, {0.999999, 1}};
    for (auto& rg : ar) {
        worst = 0;
        for (int i = 0; i < HPEM_CHECK_N; ++i) {
            const double c = rg[0] + (rg[1] - rg[0]) * U(rng);
            const double e = ulp_err(hpem::fm_acos(c), acosl((long double)c));
            if (e > worst) worst = e;
        }
        printf("fm_acos c in [%g, %g]: worst %.4f ulp\n", rg[0], rg[1], worst);
        if (worst > 1.25) ++fails;
    }
    printf("fm_acos(1)=%g fm_acos(-1)=%.17g fm_acos(0)=%.17g fm_acos(1.0000001)=%g fm_acos(nan)=%g\n", hpem::fm_acos(1.0),
           hpem::fm_acos(-1.0), hpem::fm_acos(0.0), hpem::fm_acos(1.0000001), hpem::fm_acos(NAN));
    if (hpem::fm_acos(1.0) != 0.0 || !std::isnan(hpem::fm_acos(1.0000001)) || !std::isnan(hpem::fm_acos(NAN)) ||
        hpem::fm_acos(-1.0) != std::acos(-1.0) || hpem::fm_acos(0.0) != std::acos(0.0)) ++fails;
    // log10 over j_ion's range (latent / compress_field kernels)
    worst = 0;
    for (int i = 0; i < HPEM_CHECK_N; ++i) {
        const double y = std::pow(10.0, -22.0 + 28.0 * U(rng));
        const double e = ulp_err(hpem::fm_log10(y), log10l((long double)y));
        if (e > worst) worst = e;
    }
    printf("fm_log10 y in [1e-22, 1e6]: worst %.4f ulp\n", worst);
    if (worst > 1.8) ++fails;
    // sqrt
    worst = 0;
    for (int i = 0; i < HPEM_CHECK_N; ++i) {
        const double z = std::ldexp(1.0 + U(rng), (int)(U(rng) * 300) - 150);
        const double e = ulp_err(hpem::fm_sqrt(z), sqrtl((long double)z));
        if (e > worst) worst = e;
    }
    printf("fm_sqrt: worst %.4f ulp\n", worst);
    if (worst > 0.5001) ++fails;
    printf(fails ? "FAILED (%d)\n" : "ok\n", fails);
    return fails ? 1 : 0;
}
