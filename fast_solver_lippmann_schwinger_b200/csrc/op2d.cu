// 2-D Lippmann-Schwinger operator  y = b + omega^2 * G (nu .* b)   on one B200.
// Stands behind struct FastM and its `*` / mul! / fastconvolution / FFTconvolution methods
// (reference FastConvolution.jl:11-154).  Three launches per apply:
//   P1  k_fwd_pruned  columns (x, contiguous):  b, nu      -> A  [ne x m]   (slot order in x)
//   P2  k_mid_fused   rows (y): A, Green spectrum          -> C  [m x ne]   (line contiguous)
//   P3  k_inv_pruned  columns (x): C, b                    -> y
// Padding.  The reference pads 4x per dimension (Greengard_Vico) because the truncated-kernel spectrum has
// to be *built* on the 4n grid; the apply itself only ever touches the spatial kernel g = ifft2(GFFT) at the
// lags (-n, n) x (-m, m) (input supported on [0,n), output cropped to [0,n)).  At create time the handle
// therefore computes g on the device (pruned inverse transforms of the given GFFT), wraps those lags onto a
// 2n x 2m grid and transforms back: G2 = fft2(g2).  Every apply then runs with 2x padding - the same linear
// operator to rounding (measured 1e-16 relative against the literal 4x evaluation), a quarter of the
// spectrum, half of the intermediates and about a third of the FP64 work.  LS_FLAG_PAD4 keeps the literal
// 4x evaluation (tests compare the two).
// Algorithmic HBM bytes per apply: 2x padding 24N + 64N + 64N + 64N + 32N = 248*N;
//                                  4x padding 568*N.
#include "ls_common.cuh"
#include "op2d_base.cuh"
#include "line_kernels.cuh"
#include "gv_spectrum2d.cuh"
#ifdef LS_EXPERIMENTS
#include "line_kernels_experiments.cuh"   // measured-and-rejected variants: only with -DLS_EXPERIMENTS (profiles/r1_b_notes.md)
#endif

using namespace ls;
using namespace lsk;

namespace {

struct Op2D : Op2DBase {
    int nr = 4;               // padding factor actually used by the applies (4: literal, 2: compact)
    long pn = 0, pm = 0;      // padded sizes nr*n, nr*m
    double* d_nu = nullptr;
    cd* d_G = nullptr;        // [sx][ry][slot_y] on the pn x pm grid, normalisation folded in
    cd* d_TABn = nullptr; cd* d_TABm = nullptr;   // engine tables (fft_engine.cuh EngTab)
    cd* d_A = nullptr;        // pn x m
    cd* d_C = nullptr;        // P2 output: slot sx of column j at C[(sx/cw)*(cw*m) + cw*j + sx%cw]
    int cw = 1;               // x-slot interleave of C: P3's strided reads then cover whole sectors (default 4: 64-byte runs)
    int apply_dev(const cd* b, cd* y, int mode) override;
};

__global__ void k_scale(cd* a, long n, double s) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        cd v = a[i];
        a[i] = make_double2(v.x * s, v.y * s);
    }
}

// Gd[(sx*4 + ry)*m + sy] = GFFT[(4 fx[sx%n] + sx/n + ne/2) % ne, (4 fy[sy] + ry + me/2) % me] / (ne*me)
__global__ void k_permute_g2d(const cd* __restrict__ gin, cd* __restrict__ gout,
                              const int* __restrict__ fx, const int* __restrict__ fy,
                              long n, long m, long ne, long me, double scale) {
    long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    long total = ne * me;
    if (idx >= total) return;
    long sy = idx % m;
    long ry = (idx / m) % 4;
    long sx = idx / (4 * m);
    long kx = 4L * fx[sx % n] + sx / n;
    long ky = 4L * fy[sy] + ry;
    long ix = (kx + ne / 2) % ne, iy = (ky + me / 2) % me;
    cd v = gin[ix + ne * iy];
    gout[idx] = make_double2(v.x * scale, v.y * scale);
}

// same layout as k_permute_g2d, values evaluated on the device (Gtruncated2D, Functions.jl:40-42) instead of gathered
__global__ void k_fill_g2d(cd* __restrict__ gout, const int* __restrict__ fx, const int* __restrict__ fy,
                           long n, long m, long ne, long me, double scale, Gv2dParams p) {
    long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    long total = ne * me;
    if (idx >= total) return;
    long sy = idx % m;
    long ry = (idx / m) % 4;
    long sx = idx / (4 * m);
    long kx = 4L * fx[sx % n] + sx / n;
    long ky = 4L * fy[sy] + ry;
    long ix = (kx + ne / 2) % ne, iy = (ky + me / 2) % me;
    cd v = gtrunc2d_eval(p, ix, iy, ne, me);
    gout[idx] = make_double2(v.x * scale, v.y * scale);
}
// the centred ne x me array exactly as the reference holds it (general-size path)
__global__ void k_fill_g2d_centred(cd* __restrict__ gout, long ne, long me, Gv2dParams p) {
    long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= ne * me) return;
    gout[idx] = gtrunc2d_eval(p, idx % ne, idx / ne, ne, me);
}

template <int N> int launch_fwd(Op2D* op, const cd* b, const double* nu) {
    constexpr int smem = Smem<N, false>::fwd_bytes;
    static unsigned long long optin = 0;
    LS_CUDA_TRY(smem_optin(k_fwd_pruned<N, false>, smem, optin));
    dim3 grid((unsigned)(op->m / GeoA<N>::LPC));
    LineAddr la{1L << 40, op->n, 0, 1, op->pn, 0, 1};
    la.nr = op->nr;
    op->phase_begin(0);
    k_fwd_pruned<N, false><<<grid, GeoA<N>::THREADS, smem, op->stream>>>(b, nu, op->d_A, op->d_TABn, la, 0);
    op->phase_end();
    op->launches++;
    return LS_OK;
}
template <int N, bool GSM, int MINB, int ASM = 0> int launch_mid_v(Op2D* op) {
    static int extra = -1;      // diagnostic: LS_P2_EXTRA_SMEM pads the request to lower the CTAs/SM
    if (extra < 0) { const char* e = getenv("LS_P2_EXTRA_SMEM"); extra = e ? atoi(e) : 0; }
    const int smem = (GSM ? Smem<N, false>::mid_bytes : Smem<N, false>::mid_bytes_direct) + extra
                     + ASM * GeoA<N>::THREADS * (int)sizeof(cd);
    static unsigned long long optin = 0;
    LS_CUDA_TRY(smem_optin(k_mid_fused<N, false, GSM, MINB, ASM>, smem, optin));
    dim3 grid((unsigned)(op->pn / GeoA<N>::LPC));
    // line = x slot sx; point j at A[sx + pn*j]; output: cw adjacent x-slots interleaved, C[(sx/cw)*(cw*m) + cw*j + sx%cw]
    const long cw = op->cw;
    LineAddr la{cw, 1, cw, op->pn, 1, cw * op->m, cw};
    la.nr = op->nr;
    op->phase_begin(1);
    k_mid_fused<N, false, GSM, MINB, ASM><<<grid, GeoA<N>::THREADS, smem, op->stream>>>(
        op->d_A, op->d_C, op->d_G, op->d_TABm, la, 0);
    op->phase_end();
    op->launches++;
    return LS_OK;
}
#ifdef LS_EXPERIMENTS
template <int N, int ASM> int launch_mid_dual(Op2D* op) {
    const int smem = (3 * GeoA<N>::LPC * N + EngTab<N>::TW1N + ASM * GeoA<N>::THREADS) * (int)sizeof(cd);
    static unsigned long long optin = 0;
    LS_CUDA_TRY(smem_optin(k_mid_fused_dual<N, ASM>, smem, optin));
    dim3 grid((unsigned)(op->ne / GeoA<N>::LPC));
    op->phase_begin(1);
    k_mid_fused_dual<N, ASM><<<grid, GeoA<N>::THREADS, smem, op->stream>>>(
        op->d_A, op->d_C, op->d_G, op->d_TABm, LineAddr{1L << 40, 1, 0, op->ne, op->m, 0, 1}, 0);
    op->phase_end();
    op->launches++;
    return LS_OK;
}
#endif
// 2x padding: the swap kernel (no accumulator registers, three CTAs per SM; see line_kernels.cuh)
template <int N> int launch_mid_swap_op(Op2D* op) {
    const long cw = op->cw;
    LineAddr la{cw, 1, cw, op->pn, 1, cw * op->m, cw};
    la.nr = 2;
    static int gl = -1, minb = -1;   // LS_P2_GLOAD: 0 spectrum held in registers across the last stage, 1 L2 prefetch + load at the multiply
    if (gl < 0) { const char* e = getenv("LS_P2_GLOAD"); gl = e ? atoi(e) : 0; }
    // CTAs per SM (128-thread CTAs).  Measured on B200 (profiles/r2_e_notes.md): at N >= 1024 three CTAs are slower than two
    // (168-register cap -> spills, and 3 x 66 KB of shared memory leaves 30 KB of L1 for the strided line loads:
    // 2048^2 0.297 vs 0.211 ms); at N <= 512 several lines share a CTA and three CTAs win (512^2 0.0189 vs 0.0194 ms).
    if (minb < 0) { const char* e = getenv("LS_P2_MINB"); minb = e ? atoi(e) : 0; }
    const int use_minb = minb ? minb : (N <= 256 ? 3 : 2);      // 16 points per thread from N = 512 up
    constexpr int MB = (GeoA<N>::THREADS <= 128 ? 3 : 1);
    op->phase_begin(1);
    cudaError_t e;
    constexpr int MB2 = (MB == 3 ? 2 : 1);      // 128-thread CTAs: two per SM by default (255 registers, no spills)
    if (use_minb == 2 || MB == 1) e = gl ? lsk::launch_mid_swap<N, MB2, 1>(op->stream, op->pn, op->d_A, op->d_C, op->d_G, op->d_TABm, la)
                                     : lsk::launch_mid_swap<N, MB2, 0>(op->stream, op->pn, op->d_A, op->d_C, op->d_G, op->d_TABm, la);
    else e = gl ? lsk::launch_mid_swap<N, MB, 1>(op->stream, op->pn, op->d_A, op->d_C, op->d_G, op->d_TABm, la)
                : lsk::launch_mid_swap<N, MB, 0>(op->stream, op->pn, op->d_A, op->d_C, op->d_G, op->d_TABm, la);
    op->phase_end();
    op->launches++;
    LS_CUDA_TRY(e);
    return LS_OK;
}
template <int N> int launch_mid(Op2D* op) {
    static int kernel = -1;          // LS_P2_KERNEL: 1 (default) swap kernel, 0 k_mid_fused (accumulators in registers)
    if (kernel < 0) { const char* e = getenv("LS_P2_KERNEL"); kernel = e ? atoi(e) : 1; }
    if (op->nr == 2 && kernel == 1) return launch_mid_swap_op<N>(op);
#ifndef LS_EXPERIMENTS
    return launch_mid_v<N, false, 1>(op);       // spectrum straight from HBM into registers, requested before the last butterfly stage
#else
    static int variant = -1;
    if (op->nr != 4) return launch_mid_v<N, false, 1>(op);      // the experiment kernels below are 4x-padding only
    // variants measured on B200 at 2048^2 (profiles/r1_b_notes.md): 1 = spectrum straight from HBM into
    // registers (0.706 ms), 0 = spectrum staged in shared memory by TMA bulk copies (0.761 ms)
    if (variant < 0) { const char* e = getenv("LS_P2_VARIANT"); variant = e ? atoi(e) : 1; }
    if (variant == 0) return launch_mid_v<N, true, 1>(op);
    // experiment kept for the record: persistent CTAs, next input line prefetched by cp.async (0.714-0.720 ms)
    if constexpr (N == 2048) if (variant == 8) {
        static int ctas = -1;
        if (ctas < 0) { const char* e = getenv("LS_P2_CTAS"); ctas = e ? atoi(e) : 2 * 148; }
        op->phase_begin(1);
        cudaError_t e = launch_mid_persist<N>(op->stream, op->ne, op->d_A, op->d_C, op->d_G, op->d_TABm,
                                              LineAddr{1L << 40, 1, 0, op->ne, op->m, 0, 1}, ctas);
        op->phase_end();
        op->launches++;
        LS_CUDA_TRY(e);
        return LS_OK;
    }
    // experiment kept for the record: one sub-transform per CTA, thread-block cluster of 4, DSMEM reduction
    // (3-4 CTAs/SM, no spills, bit-identical result - but 1.05-1.08 ms at 2048^2: two cluster.sync() per CTA and
    // four times the per-CTA fixed costs outweigh the better pipe overlap)
    if constexpr (N == 2048) if (variant == 6 || variant == 7) {
        op->phase_begin(1);
        const LineAddr la{1L << 40, 1, 0, op->ne, op->m, 0, 1};
        cudaError_t e = (variant == 6) ? launch_mid_cluster<N, false, 3>(op->stream, op->ne, op->d_A, op->d_C, op->d_G, op->d_TABm, la)
                                       : launch_mid_cluster<N, false, 4>(op->stream, op->ne, op->d_A, op->d_C, op->d_G, op->d_TABm, la);
        op->phase_end();
        op->launches++;
        LS_CUDA_TRY(e);
        return LS_OK;
    }
    // experiment kept for the record (profiles/r1_b_notes.md): two sub-transforms in flight per thread;
    // correct, but 255 registers + 590 B of spills make it slower (0.92 ms) - instantiated for 2048 only
    if constexpr (N == 2048) { if (variant == 5) return launch_mid_dual<N, 4>(op); }
    if (variant == 2) return launch_mid_v<N, false, 3, 4>(op);
    if (variant == 3) return launch_mid_v<N, false, 3, 2>(op);
    return launch_mid_v<N, false, 1>(op);
#endif
}
template <int N, int PF> int launch_inv_v(Op2D* op, const cd* bsrc, cd* y, double scale) {
    constexpr int smem = Smem<N, false>::fwd_bytes + (PF == 2 ? Smem<N, false>::LPC * N * (int)sizeof(cd) : 0);
    static unsigned long long optin = 0;
    LS_CUDA_TRY(smem_optin(k_inv_pruned<N, false, PF>, smem, optin));
    dim3 grid((unsigned)(op->m / GeoA<N>::LPC));
    // line = column j; slot sx at C[(sx/cw)*(cw*m) + cw*j + sx%cw]: adjacent lanes read adjacent slots
    const long cw = op->cw;
    int cshift = 0;
    while ((1L << cshift) < cw) ++cshift;
    LineAddr la{1L << 40, cw, 0, 1, op->n, 0, 1};
    la.split_shift = cshift; la.split_stride = cw * op->m;
    la.nr = op->nr;
    op->phase_begin(2);
    k_inv_pruned<N, false, PF><<<grid, GeoA<N>::THREADS, smem, op->stream>>>(op->d_C, bsrc, y, op->d_TABn, scale, la, 0);
    op->phase_end();
    op->launches++;
    return LS_OK;
}
template <int N> int launch_inv(Op2D* op, const cd* bsrc, cd* y, double scale) {
    // LS_P3_STAGE: 1 (default) = next slot block and b staged in shared memory by cp.async while the current block is
    // transformed, 0 = plain loads (2048^2 on B200: 0.131 -> 0.077 ms together with the 4-slot interleave of C)
    static int stage = -1;
    if (stage < 0) { const char* e = getenv("LS_P3_STAGE"); stage = e ? atoi(e) : 1; }
    return stage ? launch_inv_v<N, 2>(op, bsrc, y, scale) : launch_inv_v<N, 0>(op, bsrc, y, scale);
}

#define LS_DISPATCH_N(N_, CALL)                                       \
    switch (N_) {                                                     \
        case 64:   rc = CALL(64); break;                              \
        case 128:  rc = CALL(128); break;                             \
        case 256:  rc = CALL(256); break;                             \
        case 512:  rc = CALL(512); break;                             \
        case 1024: rc = CALL(1024); break;                            \
        case 2048: rc = CALL(2048); break;                            \
        case 4096: rc = CALL(4096); break;                            \
        default: rc = LS_ERR_UNSUPPORTED;                             \
    }

int apply_device(Op2D* op, const cd* b, cd* y, int mode) {
    int rc = LS_OK;
    const double* nu = (mode == LS_APPLY_FASTCONVOLUTION) ? op->d_nu : nullptr;   // Q2: GV FFTconvolution has no nu
#define CALL_FWD(N) launch_fwd<N>(op, b, nu)
    LS_DISPATCH_N(op->n, CALL_FWD);
    if (rc) return rc;
#define CALL_MID(N) launch_mid<N>(op)
    LS_DISPATCH_N(op->m, CALL_MID);
    if (rc) return rc;
    const cd* bsrc = (mode == LS_APPLY_FASTCONVOLUTION) ? b : nullptr;
    double scale = (mode == LS_APPLY_FASTCONVOLUTION) ? op->omega * op->omega : 1.0;
#define CALL_INV(N) launch_inv<N>(op, bsrc, y, scale)
    LS_DISPATCH_N(op->n, CALL_INV);
    if (rc) return rc;
    LS_CUDA_TRY(cudaGetLastError());
    return LS_OK;
}

}  // namespace

int Op2D::apply_dev(const cd* b, cd* y, int mode) { return apply_device(this, b, y, mode); }

namespace {

#define LS_DISPATCH_E(N_, CALL)                                                    \
    switch (N_) {                                                                  \
        case 64:   e = CALL(64); break;                                            \
        case 128:  e = CALL(128); break;                                           \
        case 256:  e = CALL(256); break;                                           \
        case 512:  e = CALL(512); break;                                           \
        case 1024: e = CALL(1024); break;                                          \
        case 2048: e = CALL(2048); break;                                          \
        case 4096: e = CALL(4096); break;                                          \
        default: e = cudaErrorInvalidValue;                                        \
    }

// Compact (2x) spectrum from the reference's 4x one.  g4s: [sx4][ry][sy] slot order on the 4n x 4m grid,
// already scaled by 1/(ne me).  On return op->d_G holds G2 = fft2(g wrapped onto 2n x 2m) in the slot order
// of the nr = 2 kernels, scaled by 1/(2n 2m).
int compact_spectrum(Op2D* op, cd* g4s) {
    const long n = op->n, m = op->m;
    cudaStream_t s = op->stream;
    cudaError_t e = cudaSuccess;
    cd *t1 = nullptr, *g2 = nullptr, *t2 = nullptr, *G2 = nullptr;
    int rc;
    if ((rc = op->dmalloc((void**)&t1, (size_t)(4 * n) * (2 * m) * sizeof(cd)))) return rc;
    // (1) inverse along y, keep lags [0, m) and [-m, 0): T1[sx4*2m + jy2]
    for (int c = 0; c <= 3; c += 3) {
        LineAddr la{1L << 40, 4 * m, 0, 1, 2 * m, 0, 1};
        la.nr = 4; la.cblock = c;
        cd* outp = t1 + (c ? m : 0);
#define S1(N) launch_inv<N, false>(s, 4 * n, g4s, nullptr, outp, op->d_TABm, 1.0, la)
        LS_DISPATCH_E(m, S1);
        LS_CUDA_TRY(e);
    }
    LS_CUDA_TRY(cudaStreamSynchronize(s));
    op->dfree(g4s);
    if ((rc = op->dmalloc((void**)&g2, (size_t)(2 * n) * (2 * m) * sizeof(cd)))) return rc;
    // (2) inverse along x: g2[jx2 + 2n*jy2]
    for (int c = 0; c <= 3; c += 3) {
        LineAddr la{1L << 40, 1, 0, 2 * m, 2 * n, 0, 1};
        la.nr = 4; la.cblock = c;
        cd* outp = g2 + (c ? n : 0);
#define S2(N) launch_inv<N, false>(s, 2 * m, t1, nullptr, outp, op->d_TABn, 1.0, la)
        LS_DISPATCH_E(n, S2);
        LS_CUDA_TRY(e);
    }
    LS_CUDA_TRY(cudaStreamSynchronize(s));
    op->dfree(t1);
    if ((rc = op->dmalloc((void**)&t2, (size_t)(2 * n) * (2 * m) * sizeof(cd)))) return rc;
    if ((rc = op->dmalloc((void**)&G2, (size_t)(2 * n) * (2 * m) * sizeof(cd)))) return rc;
    {   // (3) forward along x on the full 2n-point lines: T2[sx2 + 2n*jy2]
        LineAddr la{1L << 40, 2 * n, 0, 1, 2 * n, 0, 1};
        la.nr = 2; la.full2 = 1;
#define S3(N) launch_fwd<N, false>(s, 2 * m, g2, nullptr, t2, op->d_TABn, la)
        LS_DISPATCH_E(n, S3);
        LS_CUDA_TRY(e);
    }
    {   // (4) forward along y: G2[sx2*2m + (ry*m + sy)]
        LineAddr la{1L << 40, 1, 0, 2 * n, 2 * m, 0, 1};
        la.nr = 2; la.full2 = 1;
#define S4(N) launch_fwd<N, false>(s, 2 * n, t2, nullptr, G2, op->d_TABm, la)
        LS_DISPATCH_E(m, S4);
        LS_CUDA_TRY(e);
    }
    k_scale<<<148 * 8, 256, 0, s>>>(G2, (2 * n) * (2 * m), 1.0 / ((double)(2 * n) * (double)(2 * m)));
    LS_CUDA_TRY(cudaStreamSynchronize(s));
    op->dfree(g2);
    op->dfree(t2);
    op->d_G = G2;
    op->nr = 2;
    return LS_OK;
}

}  // namespace

extern "C" {

static int op2d_create_impl(ls_handle* out, int64_t n, int64_t m, int64_t ne, int64_t me,
                            const double* nu, const ls_cdouble* gfft, double omega, int quadrule, int flags,
                            double L, double Lp);

int ls_op2d_create(ls_handle* out, int64_t n, int64_t m, int64_t ne, int64_t me,
                   const double* nu, const ls_cdouble* gfft, double omega, int quadrule, int flags) {
    LS_REQUIRE(out && nu && gfft, LS_ERR_INVALID, "ls_op2d_create: null pointer");
    return op2d_create_impl(out, n, m, ne, me, nu, gfft, omega, quadrule, flags, 0.0, 0.0);
}

int ls_op2d_create_gv(ls_handle* out, int64_t n, int64_t m, const double* nu, double omega, double L, double Lp, int flags) {
    LS_REQUIRE(out && nu, LS_ERR_INVALID, "ls_op2d_create_gv: null pointer");
    LS_REQUIRE(L > 0 && Lp > 0 && omega > 0, LS_ERR_INVALID, "ls_op2d_create_gv: L, Lp and k must be positive");
    return op2d_create_impl(out, n, m, 4 * n, 4 * m, nu, nullptr, omega, LS_QUAD_GREENGARD_VICO, flags, L, Lp);
}

static int op2d_create_impl(ls_handle* out, int64_t n, int64_t m, int64_t ne, int64_t me,
                            const double* nu, const ls_cdouble* gfft, double omega, int quadrule, int flags,
                            double L, double Lp) {
    LS_REQUIRE(n > 0 && m > 0 && ne > 0 && me > 0, LS_ERR_INVALID, "ls_op2d_create: non-positive size");
    LS_REQUIRE(quadrule == LS_QUAD_TRAPEZOIDAL || quadrule == LS_QUAD_GREENGARD_VICO, LS_ERR_INVALID,
               "ls_op2d_create: unknown quadRule %d", quadrule);
    if (quadrule == LS_QUAD_TRAPEZOIDAL) {
        LS_REQUIRE(ne == 2 * n - 1 && me == 2 * m - 1, LS_ERR_INVALID,
                   "ls_op2d_create: trapezoidal needs ne = 2n-1, me = 2m-1 (FastConvolution.jl:183)");
        return create_op2d_generic(out, n, m, ne, me, nu, gfft, omega, quadrule);
    }
    LS_REQUIRE(ne == 4 * n && me == 4 * m, LS_ERR_INVALID,
               "ls_op2d_create: Greengard_Vico needs ne = 4n, me = 4m (FastConvolution.jl:201)");
    if (!(fft_size_supported(n) && fft_size_supported(m)) || (flags & LS_FLAG_FORCE_GENERIC)) {
        if (gfft) return create_op2d_generic(out, n, m, ne, me, nu, gfft, omega, quadrule);
        // general-size path with the spectrum generated on the device: the centred array as the reference holds it
        LS_REQUIRE(ne + n - 1 <= 4096 && me + m - 1 <= 4096, LS_ERR_UNSUPPORTED,
                   "ls_op2d_create_gv: n=%ld, m=%ld: the general-size GPU path serves 5n - 1 <= 4096", (long)n, (long)m);
        cd* d_g = nullptr;
        LS_CUDA_TRY(cudaMalloc((void**)&d_g, (size_t)ne * me * sizeof(cd)));
        const Gv2dParams gp = gv2d_params(L, Lp, omega);
        k_fill_g2d_centred<<<(unsigned)(((size_t)ne * me + 255) / 256), 256>>>(d_g, ne, me, gp);
        cudaError_t ge = cudaDeviceSynchronize();
        int grc = ge == cudaSuccess ? create_op2d_generic(out, n, m, ne, me, nu, nullptr, omega, quadrule, d_g) : LS_ERR_CUDA;
        if (ge != cudaSuccess) set_error("spectrum generation failed: %s", cudaGetErrorString(ge));
        cudaFree(d_g);
        return grc;
    }

    Op2D* op = new Op2D();
    int rc = op->init_base(KIND_OP2D);
    if (rc) { delete op; return rc; }
    op->n = n; op->m = m; op->ne = ne; op->me = me; op->omega = omega; op->quadrule = quadrule;
    const size_t N = (size_t)n * m, NE = (size_t)ne * me;
#define TRY(x) do { rc = (x); if (rc) { delete op; return rc; } } while (0)
    TRY(op->dupload((void**)&op->d_nu, nu, N * sizeof(double)));
    {
        auto Tn = engine_table((int)n), Tm = engine_table((int)m);
        TRY(op->dupload((void**)&op->d_TABn, Tn.data(), Tn.size() * sizeof(cd)));
        TRY(op->dupload((void**)&op->d_TABm, Tm.data(), Tm.size() * sizeof(cd)));
    }
    {
        // one-time permutation of the spectrum into [x slot][ry][y slot] order, ifftshift folded in
        auto fx = slot_freq((int)n), fy = slot_freq((int)m);
        int *d_fx = nullptr, *d_fy = nullptr;
        cd* d_gin = nullptr;
        TRY(op->dupload((void**)&d_fx, fx.data(), fx.size() * sizeof(int)));
        TRY(op->dupload((void**)&d_fy, fy.data(), fy.size() * sizeof(int)));
        TRY(op->dmalloc((void**)&op->d_G, NE * sizeof(cd)));
        const int th = 256;
        if (gfft) {
            TRY(op->dupload((void**)&d_gin, gfft, NE * sizeof(cd)));
            k_permute_g2d<<<(unsigned)((NE + th - 1) / th), th, 0, op->stream>>>(
                d_gin, op->d_G, d_fx, d_fy, n, m, ne, me, 1.0 / ((double)ne * (double)me));
        } else {
            // Gtruncated2D evaluated straight into the kernel layout: no host Bessel evaluation, no 16 ne me byte upload
            k_fill_g2d<<<(unsigned)((NE + th - 1) / th), th, 0, op->stream>>>(
                op->d_G, d_fx, d_fy, n, m, ne, me, 1.0 / ((double)ne * (double)me), gv2d_params(L, Lp, omega));
        }
        cudaError_t e = cudaStreamSynchronize(op->stream);
        if (e != cudaSuccess) { set_error("spectrum set-up failed: %s", cudaGetErrorString(e)); delete op; return LS_ERR_CUDA; }
        if (d_gin) op->dfree(d_gin);
        op->dfree(d_fx); op->dfree(d_fy);
    }
    op->nr = 4;
    if (!(flags & LS_FLAG_PAD4)) {
        cd* g4s = op->d_G;
        op->d_G = nullptr;
        TRY(compact_spectrum(op, g4s));
    }
    op->pn = op->nr * n; op->pm = op->nr * m;
    {   // x-slot interleave of the P2 -> P3 intermediate (LS_C_INTERLEAVE = 1 / 2 / 4 / 8, diagnostic)
        const char* ev = getenv("LS_C_INTERLEAVE");
        int cw = ev ? atoi(ev) : 4;        // measured at 2048^2: P2 + P3 = 0.326 / 0.313 / 0.304 / 0.31 ms for 1 / 2 / 4 / 8
        if (cw != 1 && cw != 2 && cw != 4 && cw != 8) cw = 4;
        const char* pv = getenv("LS_P2_VARIANT");
        if (op->nr == 4 && pv && atoi(pv) >= 5) cw = 1;      // the 4x-only experiment kernels write plain rows
        op->cw = cw;
    }
    TRY(op->dmalloc((void**)&op->d_A, (size_t)op->pn * m * sizeof(cd)));
    TRY(op->dmalloc((void**)&op->d_C, (size_t)op->pn * m * sizeof(cd)));
    TRY(op->dmalloc((void**)&op->d_b, N * sizeof(cd)));
    TRY(op->dmalloc((void**)&op->d_y, N * sizeof(cd)));
#undef TRY
    *out = reinterpret_cast<ls_handle>(op);
    return LS_OK;
}

int ls_op2d_apply(ls_handle h, const ls_cdouble* b, ls_cdouble* y, int mode, int memloc) {
    LS_REQUIRE(h && b && y, LS_ERR_INVALID, "ls_op2d_apply: null argument");
    Op2DBase* op = reinterpret_cast<Op2DBase*>(h);
    LS_REQUIRE(op->kind == KIND_OP2D, LS_ERR_INVALID, "ls_op2d_apply: not a 2-D operator handle");
    LS_REQUIRE(mode == LS_APPLY_FASTCONVOLUTION || mode == LS_APPLY_FFTCONVOLUTION, LS_ERR_INVALID,
               "ls_op2d_apply: unknown mode %d", mode);
    if (mode == LS_APPLY_FFTCONVOLUTION)
        LS_REQUIRE(op->n == op->m, LS_ERR_INVALID,
                   "FFTconvolution pads (ne,ne) and crops (n,n): square grids only (FastConvolution.jl:139-151)");
    LS_CUDA_TRY(cudaSetDevice(op->device));
    const size_t bytes = (size_t)op->n * op->m * sizeof(cd);
    if (memloc == LS_MEM_DEVICE) {
        return op->apply_dev(reinterpret_cast<const cd*>(b), reinterpret_cast<cd*>(y), mode);
    }
    LS_REQUIRE(memloc == LS_MEM_HOST, LS_ERR_INVALID, "ls_op2d_apply: unknown memloc %d", memloc);
    LS_CUDA_TRY(cudaMemcpyAsync(op->d_b, b, bytes, cudaMemcpyHostToDevice, op->stream));
    int rc = op->apply_dev(op->d_b, op->d_y, mode);
    if (rc) return rc;
    LS_CUDA_TRY(cudaMemcpyAsync(y, op->d_y, bytes, cudaMemcpyDeviceToHost, op->stream));
    LS_CUDA_TRY(cudaStreamSynchronize(op->stream));
    return LS_OK;
}

}  // extern "C"
