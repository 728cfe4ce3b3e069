// Sparsifying-matrix SpMV  y <- alpha*A*x + beta*y   (complex128, banded 9-/27-point stencil matrix).
// Stands behind `M.As*b` at preconditioner.jl:138,142,159,163 and mirrors the signature of
// SparseBLAS.cscmv!('N', alpha, "GXXF", A, x, beta, y) (sparseblas.jl:14-25).
//
// The matrix arrives exactly as Julia holds it (SparseMatrixCSC{ComplexF64,Int64}: 1-based colptr,
// rowval, nzval) and is converted once to CSR with 32-bit column indices.  Two device formats:
//
//  * stencil classes (fast path).  The sparsifying matrix is built by replicating one coefficient
//    vector per boundary class at columns row + offsets (createIndices, Functions.jl:7-29;
//    buildSparseA, SparsifyingMatrix2D.jl:806-884: 9 classes in 2-D, 27 in 3-D).  At create time
//    rows are grouped by their (relative offsets, values) signature; if that yields few classes,
//    the device keeps one byte per row plus the tiny class tables, so a multiply moves
//    ~33 B/row instead of the 216 B/row of CSR.  Detection is exact (bitwise equality of values),
//    so any matrix without that structure simply takes the CSR path.
//  * CSR, a sub-warp of LPR lanes per row (2 lanes for ~9 nnz/row, 8 for ~27: about 4-5 nonzeros per
//    lane keeps the most independent loads in flight), warp-shuffle reduction in a fixed order.
// Both sum each row in increasing column order, like SparseArrays' column-scatter loop.
// Algorithmic bytes (CSR accounting): nnz*20 + 4(N+1) + 32N.
#include "ls_common.cuh"
#include "spmv.cuh"
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <unordered_map>

using namespace ls;

namespace ls {

template <int LPR>
__global__ void __launch_bounds__(256)
k_spmv_csr(const int* __restrict__ rowptr, const int* __restrict__ col, const cd* __restrict__ val,
           const cd* __restrict__ x, cd* y, cd alpha, cd beta, int use_beta, long nrows) {
    const long gt = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long row = gt / LPR;
    const int lane = (int)(gt % LPR);
    double sr = 0.0, si = 0.0;
    if (row < nrows) {
        const int p0 = rowptr[row], p1 = rowptr[row + 1];
        for (int p = p0 + lane; p < p1; p += LPR) {
            const cd a = __ldg(&val[p]);
            const cd xv = __ldg(&x[__ldg(&col[p])]);
            sr += a.x * xv.x - a.y * xv.y;
            si += a.x * xv.y + a.y * xv.x;
        }
    }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) {
        sr += __shfl_xor_sync(0xffffffffu, sr, o, LPR);
        si += __shfl_xor_sync(0xffffffffu, si, o, LPR);
    }
    if (row < nrows && lane == 0) {
        cd r = make_double2(alpha.x * sr - alpha.y * si, alpha.x * si + alpha.y * sr);
        if (use_beta) {
            const cd y0 = y[row];
            r.x += beta.x * y0.x - beta.y * y0.y;
            r.y += beta.x * y0.y + beta.y * y0.x;
        }
        y[row] = r;
    }
}

// one thread per row; class tables (<= SPM_MAX_CLASS_ENTRIES entries) staged in shared memory
__global__ void __launch_bounds__(256)
k_spmv_stencil(const unsigned char* __restrict__ cls, const int* __restrict__ cls_ptr, const int* __restrict__ cls_off,
               const cd* __restrict__ cls_val, int ncls, int nent, const cd* __restrict__ x, cd* y,
               cd alpha, cd beta, int use_beta, long nrows) {
    extern __shared__ __align__(16) unsigned char smraw[];
    cd* sval = reinterpret_cast<cd*>(smraw);
    int* soff = reinterpret_cast<int*>(sval + nent);
    int* sptr = soff + nent;
    for (int i = threadIdx.x; i < nent; i += blockDim.x) { sval[i] = cls_val[i]; soff[i] = cls_off[i]; }
    for (int i = threadIdx.x; i <= ncls; i += blockDim.x) sptr[i] = cls_ptr[i];
    __syncthreads();
    for (long row = (long)blockIdx.x * blockDim.x + threadIdx.x; row < nrows; row += (long)gridDim.x * blockDim.x) {
        const int c = cls[row];
        const int p0 = sptr[c], p1 = sptr[c + 1];
        double sr = 0.0, si = 0.0;
        for (int p = p0; p < p1; ++p) {
            const cd a = sval[p];
            const cd xv = __ldg(&x[row + soff[p]]);
            sr += a.x * xv.x - a.y * xv.y;
            si += a.x * xv.y + a.y * xv.x;
        }
        cd r = make_double2(alpha.x * sr - alpha.y * si, alpha.x * si + alpha.y * sr);
        if (use_beta) {
            const cd y0 = y[row];
            r.x += beta.x * y0.x - beta.y * y0.y;
            r.y += beta.x * y0.y + beta.y * y0.x;
        }
        y[row] = r;
    }
}

int SpM::mv_dev(cd alpha, const cd* xin, cd beta, cd* y, cudaStream_t s) {
    const int use_beta = (beta.x != 0.0 || beta.y != 0.0) ? 1 : 0;
    const int th = 256;
    const cd* x = xin;
    if (halo > 0) {
        // sharded row slab: x is this rank's slab; the first / last `halo` entries travel to the z-neighbours
        if (comm != nullptr) {
            ncclResult_t r = ncclGroupStart();
            if (rank > 0 && r == ncclSuccess) {
                r = ncclSend(xin, (size_t)halo * 2, ncclDouble, rank - 1, comm, s);
                if (r == ncclSuccess) r = ncclRecv(d_xext, (size_t)halo * 2, ncclDouble, rank - 1, comm, s);
            }
            if (rank < P - 1 && r == ncclSuccess) {
                r = ncclSend(xin + (nrows - halo), (size_t)halo * 2, ncclDouble, rank + 1, comm, s);
                if (r == ncclSuccess) r = ncclRecv(d_xext + halo + nrows, (size_t)halo * 2, ncclDouble, rank + 1, comm, s);
            }
            ncclResult_t r2 = ncclGroupEnd();
            if (r == ncclSuccess) r = r2;
            if (r != ncclSuccess) { set_error("SpMV halo exchange failed: %s", ncclGetErrorString(r)); return LS_ERR_NCCL; }
        }
        LS_CUDA_TRY(cudaMemcpyAsync(d_xext + halo, xin, (size_t)nrows * sizeof(cd), cudaMemcpyDeviceToDevice, s));
        x = d_xext;
    }
    if (format == SPM_FORMAT_STENCIL) {
        const size_t smem = (size_t)nent * (sizeof(cd) + sizeof(int)) + (size_t)(ncls + 1) * sizeof(int);
        long blocks = std::min<long>((nrows + th - 1) / th, 148L * 32);
        k_spmv_stencil<<<(unsigned)blocks, th, smem, s>>>(d_cls, d_cls_ptr, d_cls_off, d_cls_val, ncls, nent, x, y,
                                                          alpha, beta, use_beta, nrows);
    } else if (lanes_per_row == 1) {
        long blocks = (nrows + th - 1) / th;
        k_spmv_csr<1><<<(unsigned)blocks, th, 0, s>>>(d_rowptr, d_col, d_val, x, y, alpha, beta, use_beta, nrows);
    } else if (lanes_per_row == 2) {
        long blocks = (nrows * 2 + th - 1) / th;
        k_spmv_csr<2><<<(unsigned)blocks, th, 0, s>>>(d_rowptr, d_col, d_val, x, y, alpha, beta, use_beta, nrows);
    } else if (lanes_per_row == 4) {
        long blocks = (nrows * 4 + th - 1) / th;
        k_spmv_csr<4><<<(unsigned)blocks, th, 0, s>>>(d_rowptr, d_col, d_val, x, y, alpha, beta, use_beta, nrows);
    } else if (lanes_per_row == 8) {
        long blocks = (nrows * 8 + th - 1) / th;
        k_spmv_csr<8><<<(unsigned)blocks, th, 0, s>>>(d_rowptr, d_col, d_val, x, y, alpha, beta, use_beta, nrows);
    } else if (lanes_per_row == 16) {
        long blocks = (nrows * 16 + th - 1) / th;
        k_spmv_csr<16><<<(unsigned)blocks, th, 0, s>>>(d_rowptr, d_col, d_val, x, y, alpha, beta, use_beta, nrows);
    } else {
        long blocks = (nrows * 32 + th - 1) / th;
        k_spmv_csr<32><<<(unsigned)blocks, th, 0, s>>>(d_rowptr, d_col, d_val, x, y, alpha, beta, use_beta, nrows);
    }
    launches++;
    LS_CUDA_TRY(cudaGetLastError());
    return LS_OK;
}

}  // namespace ls

namespace {

constexpr int SPM_MAX_CLASSES = 255;
constexpr int SPM_MAX_CLASS_ENTRIES = 2048;

// groups rows by (relative offsets, values); returns false when the matrix has no small class structure
bool detect_stencil_classes(long nrows, const std::vector<int>& rowptr, const std::vector<int>& col,
                            const std::vector<cd>& val, std::vector<unsigned char>& cls,
                            std::vector<int>& cls_ptr, std::vector<int>& cls_off, std::vector<cd>& cls_val) {
    struct Rep { long row; int id; };
    std::unordered_map<uint64_t, std::vector<Rep>> table;
    cls.assign((size_t)nrows, 0);
    cls_ptr.assign(1, 0);
    cls_off.clear();
    cls_val.clear();
    int ncls = 0;
    auto same = [&](long r0, long r1) {
        const int a0 = rowptr[r0], a1 = rowptr[r1];
        const int len = rowptr[r0 + 1] - a0;
        if (rowptr[r1 + 1] - a1 != len) return false;
        for (int i = 0; i < len; ++i) {
            if ((long)col[a0 + i] - r0 != (long)col[a1 + i] - r1) return false;
            if (memcmp(&val[a0 + i], &val[a1 + i], sizeof(cd)) != 0) return false;
        }
        return true;
    };
    for (long r = 0; r < nrows; ++r) {
        uint64_t hsh = 1469598103934665603ULL;
        auto mix = [&](uint64_t v) { hsh ^= v; hsh *= 1099511628211ULL; };
        for (int p = rowptr[r]; p < rowptr[r + 1]; ++p) {
            mix((uint64_t)((long)col[p] - r));
            uint64_t bits[2];
            memcpy(bits, &val[p], sizeof(bits));
            mix(bits[0]);
            mix(bits[1]);
        }
        mix((uint64_t)(rowptr[r + 1] - rowptr[r]));
        auto& bucket = table[hsh];
        int id = -1;
        for (const Rep& rep : bucket)
            if (same(rep.row, r)) { id = rep.id; break; }
        if (id < 0) {
            if (ncls == SPM_MAX_CLASSES) return false;
            id = ncls++;
            bucket.push_back(Rep{r, id});
            for (int p = rowptr[r]; p < rowptr[r + 1]; ++p) {
                cls_off.push_back((int)((long)col[p] - r));
                cls_val.push_back(val[p]);
            }
            cls_ptr.push_back((int)cls_off.size());
            if ((int)cls_off.size() > SPM_MAX_CLASS_ENTRIES) return false;
        }
        cls[(size_t)r] = (unsigned char)id;
    }
    return true;
}

}  // namespace

extern "C" {

static int spm_build(ls_handle* out, int64_t nrows, int64_t ncols, const int64_t* colptr, const int64_t* rowval,
                     const ls_cdouble* nzval, bool window) {
    LS_REQUIRE(out && colptr && rowval && nzval, LS_ERR_INVALID, "ls_spm_create: null pointer");
    LS_REQUIRE(nrows > 0 && ncols > 0, LS_ERR_INVALID, "ls_spm_create: non-positive size");
    LS_REQUIRE(colptr[0] == 1, LS_ERR_INVALID, "ls_spm_create: colptr must be 1-based (Julia SparseMatrixCSC)");
    const int64_t nnz = colptr[ncols] - 1;
    LS_REQUIRE(nnz >= 0 && nnz < (int64_t)2147483647 && ncols < (int64_t)2147483647, LS_ERR_UNSUPPORTED,
               "ls_spm_create: nnz=%ld exceeds the 32-bit index range of the device format", (long)nnz);
    // CSC (1-based) -> CSR (0-based, int32), columns ascending inside each row: the same summation
    // order per output row as SparseArrays' column-scatter loop
    std::vector<int> rowptr((size_t)nrows + 1, 0);
    for (int64_t c = 0; c < ncols; ++c) {
        LS_REQUIRE(colptr[c + 1] >= colptr[c], LS_ERR_INVALID, "ls_spm_create: colptr not monotone at column %ld", (long)c);
        for (int64_t p = colptr[c] - 1; p < colptr[c + 1] - 1; ++p) {
            const int64_t r = rowval[p] - 1;
            LS_REQUIRE(r >= 0 && r < nrows, LS_ERR_INVALID, "ls_spm_create: row index %ld out of range", (long)(r + 1));
            rowptr[(size_t)r + 1]++;
        }
    }
    for (int64_t r = 0; r < nrows; ++r) rowptr[(size_t)r + 1] += rowptr[(size_t)r];
    std::vector<int> col((size_t)nnz), next(rowptr.begin(), rowptr.end() - 1);
    std::vector<cd> val((size_t)nnz);
    const cd* nz = reinterpret_cast<const cd*>(nzval);
    for (int64_t c = 0; c < ncols; ++c)
        for (int64_t p = colptr[c] - 1; p < colptr[c + 1] - 1; ++p) {
            const int64_t r = rowval[p] - 1;
            const int q = next[(size_t)r]++;
            col[(size_t)q] = (int)c;
            val[(size_t)q] = nz[p];
        }
    SpM* A = new SpM();
    int rc = A->init_base(KIND_SPM);
    if (rc) { delete A; return rc; }
    A->nrows = nrows; A->ncols = ncols; A->nnz = nnz;
    const double avg = (double)nnz / (double)nrows;
    // lanes per row: ~4-5 nonzeros per lane (measured on B200, 9 nnz/row at 2048^2: 1 lane 70 %, 2 lanes 88 %,
    // 4 lanes 76 %, 8 lanes 48 %, 32 lanes 16 % of the CSR-accounting HBM roofline)
    A->lanes_per_row = avg <= 4.0 ? 1 : (avg <= 10.0 ? 2 : (avg <= 20.0 ? 4 : (avg <= 40.0 ? 8 : (avg <= 80.0 ? 16 : 32))));
    if (const char* lp = getenv("LS_SPM_LANES")) { int v = atoi(lp); if (v == 1 || v == 2 || v == 4 || v == 8 || v == 16 || v == 32) A->lanes_per_row = v; }
#define TRY(x) do { rc = (x); if (rc) { delete A; return rc; } } while (0)
    const char* force = getenv("LS_SPM_FORCE_CSR");
    std::vector<unsigned char> cls;
    std::vector<int> cls_ptr, cls_off;
    std::vector<cd> cls_val;
    if (!(force && atoi(force)) && (nrows == ncols || window) &&
        detect_stencil_classes(nrows, rowptr, col, val, cls, cls_ptr, cls_off, cls_val) && !cls_off.empty()) {
        A->format = SPM_FORMAT_STENCIL;
        A->ncls = (int)cls_ptr.size() - 1;
        A->nent = (int)cls_off.size();
        TRY(A->dupload((void**)&A->d_cls, cls.data(), cls.size()));
        TRY(A->dupload((void**)&A->d_cls_ptr, cls_ptr.data(), cls_ptr.size() * sizeof(int)));
        TRY(A->dupload((void**)&A->d_cls_off, cls_off.data(), cls_off.size() * sizeof(int)));
        TRY(A->dupload((void**)&A->d_cls_val, cls_val.data(), cls_val.size() * sizeof(cd)));
    } else {
        A->format = SPM_FORMAT_CSR;
        TRY(A->dupload((void**)&A->d_rowptr, rowptr.data(), rowptr.size() * sizeof(int)));
        TRY(A->dupload((void**)&A->d_col, col.data(), std::max<size_t>(col.size(), 1) * sizeof(int)));
        TRY(A->dupload((void**)&A->d_val, val.data(), std::max<size_t>(val.size(), 1) * sizeof(cd)));
    }
    TRY(A->dmalloc((void**)&A->d_x, (size_t)ncols * sizeof(cd)));
    TRY(A->dmalloc((void**)&A->d_y, (size_t)nrows * sizeof(cd)));
#undef TRY
    *out = reinterpret_cast<ls_handle>(A);
    return LS_OK;
}

int ls_spm_create(ls_handle* out, int64_t nrows, int64_t ncols, const int64_t* colptr, const int64_t* rowval,
                  const ls_cdouble* nzval) {
    return spm_build(out, nrows, ncols, colptr, rowval, nzval, false);
}

int ls_spm_create_dist(ls_handle* out, ls_handle op, int64_t nrows_local, int64_t halo, const int64_t* colptr,
                       const int64_t* rowval, const ls_cdouble* nzval) {
    LS_REQUIRE(out && op, LS_ERR_INVALID, "ls_spm_create_dist: null pointer");
    HandleBase* o = reinterpret_cast<HandleBase*>(op);
    LS_REQUIRE(o->kind == KIND_OP3D, LS_ERR_INVALID, "ls_spm_create_dist: `op` must be a 3-D operator handle");
    LS_REQUIRE(nrows_local == o->op_size(), LS_ERR_INVALID,
               "ls_spm_create_dist: %ld local rows, the operator's slab has %ld", (long)nrows_local, (long)o->op_size());
    LS_REQUIRE(halo > 0 && halo <= nrows_local, LS_ERR_INVALID,
               "ls_spm_create_dist: halo %ld must be in [1, local rows] (neighbour slabs only)", (long)halo);
    int rc = spm_build(out, nrows_local, nrows_local + 2 * halo, colptr, rowval, nzval, true);
    if (rc) return rc;
    SpM* A = reinterpret_cast<SpM*>(*out);
    A->halo = halo;
    A->rank = o->dist_rank(); A->P = o->dist_size(); A->comm = o->nccl_comm();
    rc = A->dmalloc((void**)&A->d_xext, (size_t)A->ncols * sizeof(cd));
    if (rc == LS_OK && cudaMemset(A->d_xext, 0, (size_t)A->ncols * sizeof(cd)) != cudaSuccess) {
        set_error("ls_spm_create_dist: cudaMemset failed");
        rc = LS_ERR_CUDA;
    }
    if (rc) { delete A; *out = nullptr; return rc; }
    return LS_OK;
}

int ls_spm_mv(ls_handle h, ls_cdouble alpha, const ls_cdouble* x, ls_cdouble beta, ls_cdouble* y, int memloc) {
    LS_REQUIRE(h && x && y, LS_ERR_INVALID, "ls_spm_mv: null argument");
    SpM* A = reinterpret_cast<SpM*>(h);
    LS_REQUIRE(A->kind == KIND_SPM, LS_ERR_INVALID, "ls_spm_mv: not a sparse-matrix handle");
    LS_CUDA_TRY(cudaSetDevice(A->device));
    const cd al = make_double2(alpha.re, alpha.im), be = make_double2(beta.re, beta.im);
    if (memloc == LS_MEM_DEVICE) {
        LS_REQUIRE((const void*)x != (const void*)y, LS_ERR_INVALID, "ls_spm_mv: x and y must not alias");
        return A->mv_dev(al, reinterpret_cast<const cd*>(x), be, reinterpret_cast<cd*>(y), A->stream);
    }
    LS_REQUIRE(memloc == LS_MEM_HOST, LS_ERR_INVALID, "ls_spm_mv: unknown memloc %d", memloc);
    LS_CUDA_TRY(cudaMemcpyAsync(A->d_x, x, (size_t)A->x_len() * sizeof(cd), cudaMemcpyHostToDevice, A->stream));
    if (be.x != 0.0 || be.y != 0.0)
        LS_CUDA_TRY(cudaMemcpyAsync(A->d_y, y, (size_t)A->nrows * sizeof(cd), cudaMemcpyHostToDevice, A->stream));
    int rc = A->mv_dev(al, A->d_x, be, A->d_y, A->stream);
    if (rc) return rc;
    LS_CUDA_TRY(cudaMemcpyAsync(y, A->d_y, (size_t)A->nrows * sizeof(cd), cudaMemcpyDeviceToHost, A->stream));
    LS_CUDA_TRY(cudaStreamSynchronize(A->stream));
    return LS_OK;
}

int ls_spm_info(ls_handle h, int64_t* nrows, int64_t* ncols, int64_t* nnz, int* format, int* nclasses) {
    LS_REQUIRE(h, LS_ERR_INVALID, "ls_spm_info: null handle");
    SpM* A = reinterpret_cast<SpM*>(h);
    LS_REQUIRE(A->kind == KIND_SPM, LS_ERR_INVALID, "ls_spm_info: not a sparse-matrix handle");
    if (nrows) *nrows = A->nrows;
    if (ncols) *ncols = A->ncols;
    if (nnz) *nnz = A->nnz;
    if (format) *format = A->format;
    if (nclasses) *nclasses = A->format == SPM_FORMAT_STENCIL ? A->ncls : 0;
    return LS_OK;
}

}  // extern "C"
