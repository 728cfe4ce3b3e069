// Sparse-matrix handle shared by spmv.cu and krylov.cu (the GMRES driver applies As on its own stream).
#pragma once
#include "ls_common.cuh"

namespace ls {

enum { SPM_FORMAT_CSR = 0, SPM_FORMAT_STENCIL = 1 };

struct SpM : HandleBase {
    long nrows = 0, ncols = 0, nnz = 0;
    int format = SPM_FORMAT_CSR;
    // CSR
    int lanes_per_row = 8;
    int* d_rowptr = nullptr;
    int* d_col = nullptr;
    cd* d_val = nullptr;
    // stencil classes: class id per row + per-class (offset, value) lists
    int ncls = 0, nent = 0;
    unsigned char* d_cls = nullptr;
    int* d_cls_ptr = nullptr;
    int* d_cls_off = nullptr;
    cd* d_cls_val = nullptr;
    cd* d_x = nullptr; cd* d_y = nullptr;   // staging for host-pointer calls
    // y <- alpha*A*x + beta*y on device pointers, enqueued on stream s
    int mv_dev(cd alpha, const cd* x, cd beta, cd* y, cudaStream_t s);
};

}  // namespace ls
