// Greengard-Vico truncated-kernel spectrum in 3-D, evaluated on the device.
// Gtruncated3D(L, k, s) = (-1 + e^{iLk} (cos(Ls) - i k L sinc(Ls/pi))) / (k^2 - s^2)   (Functions.jl:49-51;
// Julia's sinc(x) = sin(pi x)/(pi x)).  The sinc argument goes through pi*(Ls/pi) exactly like the
// oracle's numpy.sinc so that both round the same way at arguments ~1e3.
#pragma once
#include "fft_engine.cuh"

namespace ls {

__device__ __forceinline__ lsfft::cd gtrunc3d_eval(double s, double L, double k, double eLk_re, double eLk_im) {
    const double PI = 3.141592653589793;
    // Q7: removable singularity at s == k (Functions.jl:50 divides by k^2 - s^2); a grid frequency that hits k exactly is
    // moved by sqrt(eps) relative instead of producing NaN
    if (s == k) s = k * (1.0 + 1.4901161193847656e-08);
    const double Ls = L * s;
    const double c = cos(Ls);
    double sinc;
    {
        const double xs = Ls / PI;
        const double yv = PI * xs;
        sinc = (xs == 0.0) ? 1.0 : sin(yv) / yv;
    }
    const double ar = c, ai = -(k * L * sinc);
    const double nr = -1.0 + (eLk_re * ar - eLk_im * ai);
    const double ni = eLk_re * ai + eLk_im * ar;
    const double den = k * k - s * s;
    return make_double2(nr / den, ni / den);
}

// |kappa| on the centred grid (2 pi/Lp)(-N/2 .. N/2-1), summed like numpy: (kx^2 + ky^2) + kz^2, no FMA contraction
__device__ __forceinline__ double gv_radius(double dk, long ix, long iy, long iz, long ne, long me, long le) {
    const double kxv = dk * (double)(ix - ne / 2);
    const double kyv = dk * (double)(iy - me / 2);
    const double kzv = dk * (double)(iz - le / 2);
    return sqrt(__dadd_rn(__dadd_rn(__dmul_rn(kxv, kxv), __dmul_rn(kyv, kyv)), __dmul_rn(kzv, kzv)));
}

int create_op3d_generic(ls_handle* out, int64_t n, int64_t m, int64_t l, int64_t ne, int64_t me, int64_t le,
                        const double* nu, const ls_cdouble* gfft, double omega, double L, double Lp);

}  // namespace ls
