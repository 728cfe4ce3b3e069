// Greengard-Vico truncated-kernel spectrum in 2-D, evaluated on the device.
// Gtruncated2D(L, k, s) = (1 + (i pi/2) L H0(Lk) s J1(Ls) - (i pi/2) L k H1(Lk) J0(Ls)) / (s^2 - k^2)   (Functions.jl:40-42)
// on the centred grid s = (2 pi/Lp) |(kx, ky)|, kx = -2n..2n-1 (FastConvolution.jl:185-231).  The two Hankel scalars are
// evaluated once on the host; J0 / J1 come from Chebyshev expansions generated with 50-digit arithmetic
// (scripts/gen_bessel_tables.py -> bessel_tables.cuh), accurate to ~1e-15 of the envelope min(1, sqrt(2/(pi x))) for
// arguments up to 3e4 (CUDA's own j0 / j1 are specified to 5e-12 absolute beyond |x| = 8 - not enough for 1e-12 parity).
#pragma once
#include "fft_engine.cuh"
#include "bessel_tables.cuh"
#include <cmath>

namespace ls {

template <int N>
__device__ __forceinline__ double cheb_dev(const double (&c)[N], double y) {
    double b1 = 0.0, b2 = 0.0;
#pragma unroll
    for (int k = N - 1; k >= 1; --k) {
        const double t = 2.0 * y * b1 - b2 + c[k];
        b2 = b1;
        b1 = t;
    }
    return y * b1 - b2 + c[0];
}
template <int N>
inline double cheb_host(const double (&c)[N], double y) {
    double b1 = 0.0, b2 = 0.0;
    for (int k = N - 1; k >= 1; --k) {
        const double t = 2.0 * y * b1 - b2 + c[k];
        b2 = b1;
        b1 = t;
    }
    return y * b1 - b2 + c[0];
}

// J0(x), J1(x) for x >= 0
__device__ __forceinline__ void bessel_j01(double x, double& j0, double& j1) {
    using namespace lsbessel;
    if (x <= 8.0) {
        const double y = x * x * 0.03125 - 1.0;
        j0 = cheb_dev(C_J0S, y);
        j1 = x * cheb_dev(C_J1S, y);
    } else {
        const double w = 8.0 / x;
        const double y = 2.0 * w * w - 1.0;
        double s, c;
        sincos(x, &s, &c);
        const double r2 = 0.70710678118654752440;
        const double a = sqrt(0.63661977236758134308 / x);        // sqrt(2/(pi x))
        const double cc = r2 * (c + s), ss = r2 * (s - c);          // cos(x - pi/4), sin(x - pi/4)
        j0 = a * (cheb_dev(C_P0, y) * cc - w * cheb_dev(C_Q0, y) * ss);
        j1 = a * (cheb_dev(C_P1, y) * ss + w * cheb_dev(C_Q1, y) * cc);   // cos(x - 3pi/4) = ss, sin(x - 3pi/4) = -cc
    }
}

// H0^(1)(x), H1^(1)(x) on the host (x = L k, one evaluation per operator)
inline void hankel1_01_host(double x, double& h0r, double& h0i, double& h1r, double& h1i) {
    using namespace lsbessel;
    if (x <= 8.0) {
        h0r = std::cyl_bessel_j(0.0, x); h0i = std::cyl_neumann(0.0, x);
        h1r = std::cyl_bessel_j(1.0, x); h1i = std::cyl_neumann(1.0, x);
        return;
    }
    const double w = 8.0 / x, y = 2.0 * w * w - 1.0;
    const long double xl = (long double)x;
    const double s = (double)sinl(xl), c = (double)cosl(xl);
    const double r2 = 0.70710678118654752440;
    const double a = std::sqrt(0.63661977236758134308 / x);
    const double cc = r2 * (c + s), ss = r2 * (s - c);
    const double P0 = cheb_host(H_P0, y), Q0 = w * cheb_host(H_Q0, y), P1 = cheb_host(H_P1, y), Q1 = w * cheb_host(H_Q1, y);
    h0r = a * (P0 * cc - Q0 * ss);  h0i = a * (P0 * ss + Q0 * cc);         // J0, Y0
    h1r = a * (P1 * ss + Q1 * cc);  h1i = a * (-P1 * cc + Q1 * ss);        // J1, Y1 (chi = x - 3 pi/4)
}

struct Gv2dParams {
    double dk;            // 2 pi / Lp
    double L, k;
    double c1r, c1i;      // (i pi/2) L H0(Lk)
    double c2r, c2i;      // (i pi/2) L k H1(Lk)
};

inline Gv2dParams gv2d_params(double L, double Lp, double k) {
    Gv2dParams p;
    p.dk = 2.0 * 3.141592653589793 / Lp;
    p.L = L; p.k = k;
    double h0r, h0i, h1r, h1i;
    hankel1_01_host(L * k, h0r, h0i, h1r, h1i);
    const double f1 = 3.141592653589793 / 2 * L, f2 = 3.141592653589793 / 2 * L * k;
    p.c1r = -f1 * h0i; p.c1i = f1 * h0r;          // i * f1 * (h0r + i h0i)
    p.c2r = -f2 * h1i; p.c2i = f2 * h1r;
    return p;
}

// spectrum value at the centred grid point (ix, iy) of the ne x me grid; rounding follows the oracle's numpy expression
// (kx^2 + ky^2 without FMA contraction).  Q7: the formula is singular (removably) at s == k; a grid frequency that hits k
// exactly is moved by sqrt(eps) relative instead of producing NaN.
__device__ __forceinline__ lsfft::cd gtrunc2d_eval(const Gv2dParams& p, long ix, long iy, long ne, long me) {
    const double kxv = p.dk * (double)(ix - ne / 2);
    const double kyv = p.dk * (double)(iy - me / 2);
    double s = sqrt(__dadd_rn(__dmul_rn(kxv, kxv), __dmul_rn(kyv, kyv)));
    if (s == p.k) s = p.k * (1.0 + 1.4901161193847656e-08);
    double j0, j1;
    bessel_j01(p.L * s, j0, j1);
    const double sj1 = s * j1;
    const double nr = 1.0 + p.c1r * sj1 - p.c2r * j0;
    const double ni = p.c1i * sj1 - p.c2i * j0;
    const double den = s * s - p.k * p.k;
    return make_double2(nr / den, ni / den);
}

}  // namespace ls
