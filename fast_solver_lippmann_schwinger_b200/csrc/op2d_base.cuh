// Common part of the two 2-D operator implementations (power-of-two pruned path, op2d.cu;
// general-size Bluestein path, op2d_generic.cu).
#pragma once
#include "ls_common.cuh"

namespace ls {

struct Op2DBase : HandleBase {
    long n = 0, m = 0, ne = 0, me = 0;
    double omega = 0;
    int quadrule = 0;
    cd* d_b = nullptr; cd* d_y = nullptr;   // staging for host-pointer applies
    int64_t op_size() const override { return n * m; }
};

// general sizes / trapezoidal rule: line DFTs of arbitrary length by Bluestein's algorithm on the
// power-of-two engine.  Returns LS_ERR_UNSUPPORTED when a padded line does not fit the engine.
// gfft: host array in the reference's layout, or nullptr with gfft_dev = the same array already on the device.
int create_op2d_generic(ls_handle* out, int64_t n, int64_t m, int64_t ne, int64_t me, const double* nu,
                        const ls_cdouble* gfft, double omega, int quadrule, const cd* gfft_dev = nullptr);

}  // namespace ls
