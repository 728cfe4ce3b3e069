// General-size 2-D Lippmann-Schwinger operator: any n, m and both quadratures.
//
// Serves what the power-of-two fast path (op2d.cu) cannot: the sizes the reference actually ships
// (examples/example.jl: n = 201, Greengard_Vico padded 804 = 2^2*3*67) and the Duan-Rokhlin
// trapezoidal rule (FastConvolution.jl:64-83: padded 2n-1, crop [n:2n-1]).  Every line DFT of
// arbitrary length Lf is evaluated exactly by Bluestein's identity
//       X_q = conj(c_q) * sum_j [x_j conj(c_j)] c_{q-j},      c_t = exp(i pi t^2 / Lf)
// as a circular convolution of power-of-two length Nb >= Lf + nin - 1 on the same register-resident
// engine (forward -> multiply by the chirp spectrum -> adjoint inverse, all inside one CTA).
// Zero padding is still pruned: only the nin non-zero inputs are read and only the nout kept outputs
// are formed.  Three launches per apply, like the fast path:
//   G1 k_bs_fwd   columns: b, nu          -> A [ne x m]     (natural frequency order)
//   G2 k_bs_mid   rows: A, spectrum       -> C [m x ne]     (forward, x GFFT, inverse, crop)
//   G3 k_bs_inv   columns: C, b           -> y              (inverse, crop, b + omega^2 *)
// In G2 the post-chirp of the forward and the pre-chirp of the inverse cancel exactly.
#include "ls_common.cuh"
#include "op2d_base.cuh"
#include "line_kernels.cuh"

using namespace ls;
using namespace lsk;

namespace {

struct BsDim {            // one padded dimension
    long nin = 0;         // non-zero inputs / kept outputs
    long Lf = 0;          // DFT length (ne or me)
    long o0 = 0;          // first kept output of the inverse
    int Nb = 0;           // power-of-two convolution length
    cd* d_tab = nullptr;  // engine table of size Nb
    cd* d_ch = nullptr;   // chirp c_t, t < Lf
    cd* d_hf = nullptr;   // FFT_Nb of the forward kernel, slot order
    cd* d_hi = nullptr;   // FFT_Nb of the inverse kernel, slot order
};

struct Op2DGeneric : Op2DBase {
    BsDim X, Y;
    double* d_nu = nullptr;
    cd* d_G = nullptr;     // [kx][ky] natural order, shift folded, scaled
    cd* d_A = nullptr;     // ne x m
    cd* d_C = nullptr;     // m x ne
    int apply_dev(const cd* b, cd* y, int mode) override;
};

// natural index of element a of thread t
template <int Nb> struct BsGeo {
    static constexpr int E = Cfg<Nb>::E, T = Nb / E;
    static constexpr int LPC = GeoA<Nb>::LPC, THREADS = GeoA<Nb>::THREADS;
    static constexpr int smem = (LPC * Nb + EngTab<Nb>::TW1N) * (int)sizeof(cd);
};

// circular convolution with the precomputed kernel spectrum H (slot order); v natural -> natural
template <int Nb>
__device__ __forceinline__ void bs_convolve(cd* v, int t, cd* ex, const LayA<Nb>& lay, const TwState<Nb>& tw,
                                            const cd* __restrict__ H) {
    constexpr int E = Cfg<Nb>::E, T = Nb / E;
    fft_fwd<Nb>(v, t, 0, ex, lay, tw);
#pragma unroll
    for (int e = 0; e < E; ++e) v[e] = cmul(v[e], __ldg(&H[t + T * e]));
    fft_inv<Nb>(v, t, 0, ex, lay, tw);
}

// spectrum of a Bluestein kernel: h natural (Nb) -> slot order
template <int Nb>
__global__ void __launch_bounds__(BsGeo<Nb>::THREADS)
k_bs_kernel_spectrum(const cd* __restrict__ h, cd* __restrict__ H, const cd* __restrict__ TAB) {
    constexpr int E = Cfg<Nb>::E, T = Nb / E;
    extern __shared__ __align__(128) cd sm[];
    Map<Nb, false> mp;
    cd* tw1 = sm + GeoA<Nb>::LPC * Nb;
    load_tw1<Nb>(tw1, TAB);
    const TwState<Nb> tw = make_tw<Nb>(mp.t, TAB, tw1);
    __syncthreads();
    cd v[E];
    if (mp.line == 0) {
#pragma unroll
        for (int a = 0; a < E; ++a) v[a] = h[a * T + mp.t];
    } else {
#pragma unroll
        for (int a = 0; a < E; ++a) v[a] = make_double2(0.0, 0.0);
    }
    fft_fwd<Nb>(v, mp.t, 0, sm, mp.lay, tw);
    if (mp.line == 0) {
#pragma unroll
        for (int e = 0; e < E; ++e) H[mp.t + T * e] = v[e];
    }
}

// forward: nin inputs (x nu) -> Lf outputs
template <int Nb>
__global__ void __launch_bounds__(BsGeo<Nb>::THREADS)
k_bs_fwd(const cd* __restrict__ in, const double* __restrict__ nu, cd* __restrict__ out, const cd* __restrict__ TAB,
         const cd* __restrict__ CH, const cd* __restrict__ HF, int nin, int Lf, long nlines, const LineAddr la) {
    constexpr int E = Cfg<Nb>::E, T = Nb / E, LPC = GeoA<Nb>::LPC;
    extern __shared__ __align__(128) cd sm[];
    Map<Nb, false> mp;
    cd* tw1 = sm + LPC * Nb;
    load_tw1<Nb>(tw1, TAB);
    const int t = mp.t;
    const TwState<Nb> tw = make_tw<Nb>(t, TAB, tw1);
    long L = (long)blockIdx.x * LPC + mp.line;
    const bool live = L < nlines;
    if (!live) L = nlines - 1;
    const long ib = line_in(la, L), ob = line_out(la, L);
    cd v[E];
#pragma unroll
    for (int a = 0; a < E; ++a) {
        const int idx = a * T + t;
        cd x = make_double2(0.0, 0.0);
        if (idx < nin) {
            const long off = ib + (long)idx * la.in_es;
            x = in[off];
            if (nu != nullptr) { const double s = nu[off]; x.x *= s; x.y *= s; }
            x = cmulc(x, __ldg(&CH[idx]));
        }
        v[a] = x;
    }
    __syncthreads();
    bs_convolve<Nb>(v, t, sm, mp.lay, tw, HF);
    if (live) {
#pragma unroll
        for (int a = 0; a < E; ++a) {
            const int idx = a * T + t;
            if (idx < Lf) out[ob + (long)idx * la.out_es] = cmulc(v[a], __ldg(&CH[idx]));
        }
    }
}

// middle: nin inputs -> DFT_Lf -> x G -> IDFT_Lf -> outputs [o0, o0 + nin)
template <int Nb>
__global__ void __launch_bounds__(BsGeo<Nb>::THREADS)
k_bs_mid(const cd* __restrict__ in, cd* __restrict__ out, const cd* __restrict__ G, const cd* __restrict__ TAB,
         const cd* __restrict__ CH, const cd* __restrict__ HF, const cd* __restrict__ HI, int nin, int Lf, int o0,
         long nlines, const LineAddr la) {
    constexpr int E = Cfg<Nb>::E, T = Nb / E, LPC = GeoA<Nb>::LPC;
    extern __shared__ __align__(128) cd sm[];
    Map<Nb, false> mp;
    cd* tw1 = sm + LPC * Nb;
    load_tw1<Nb>(tw1, TAB);
    const int t = mp.t;
    const TwState<Nb> tw = make_tw<Nb>(t, TAB, tw1);
    long L = (long)blockIdx.x * LPC + mp.line;
    const bool live = L < nlines;
    if (!live) L = nlines - 1;
    const long ib = line_in(la, L), ob = line_out(la, L);
    cd v[E];
#pragma unroll
    for (int a = 0; a < E; ++a) {
        const int idx = a * T + t;
        cd x = make_double2(0.0, 0.0);
        if (idx < nin) x = cmulc(in[ib + (long)idx * la.in_es], __ldg(&CH[idx]));
        v[a] = x;
    }
    __syncthreads();
    bs_convolve<Nb>(v, t, sm, mp.lay, tw, HF);
    // X_q = conj(c_q) conv_q ; inverse input Y_q c_q = conv_q G_q  (chirps cancel)
    const cd* g = G + L * (long)Lf;
#pragma unroll
    for (int a = 0; a < E; ++a) {
        const int idx = a * T + t;
        v[a] = (idx < Lf) ? cmul(v[a], __ldg(&g[idx])) : make_double2(0.0, 0.0);
    }
    bs_convolve<Nb>(v, t, sm, mp.lay, tw, HI);
    if (live) {
#pragma unroll
        for (int a = 0; a < E; ++a) {
            const int idx = a * T + t;
            if (idx >= o0 && idx < o0 + nin) out[ob + (long)(idx - o0) * la.out_es] = cmul(v[a], __ldg(&CH[idx]));
        }
    }
}

// inverse: Lf inputs -> outputs [o0, o0 + nout), optional combine
template <int Nb>
__global__ void __launch_bounds__(BsGeo<Nb>::THREADS)
k_bs_inv(const cd* __restrict__ in, const cd* bsrc, cd* out, const cd* __restrict__ TAB, const cd* __restrict__ CH,
         const cd* __restrict__ HI, int nout, int Lf, int o0, double scale, long nlines, const LineAddr la) {
    constexpr int E = Cfg<Nb>::E, T = Nb / E, LPC = GeoA<Nb>::LPC;
    extern __shared__ __align__(128) cd sm[];
    Map<Nb, false> mp;
    cd* tw1 = sm + LPC * Nb;
    load_tw1<Nb>(tw1, TAB);
    const int t = mp.t;
    const TwState<Nb> tw = make_tw<Nb>(t, TAB, tw1);
    long L = (long)blockIdx.x * LPC + mp.line;
    const bool live = L < nlines;
    if (!live) L = nlines - 1;
    const long ib = line_in(la, L), ob = line_out(la, L);
    cd v[E];
#pragma unroll
    for (int a = 0; a < E; ++a) {
        const int idx = a * T + t;
        v[a] = (idx < Lf) ? cmul(in[ib + (long)idx * la.in_es], __ldg(&CH[idx])) : make_double2(0.0, 0.0);
    }
    __syncthreads();
    bs_convolve<Nb>(v, t, sm, mp.lay, tw, HI);
    if (live) {
#pragma unroll
        for (int a = 0; a < E; ++a) {
            const int idx = a * T + t;
            if (idx >= o0 && idx < o0 + nout) {
                const long off = ob + (long)(idx - o0) * la.out_es;
                cd r = cscale(cmul(v[a], __ldg(&CH[idx])), scale);
                if (bsrc != nullptr) r = cadd(r, bsrc[off]);
                out[off] = r;
            }
        }
    }
}

// Gd[kx*me + ky] = GFFT[(kx + sx) % ne, (ky + sy) % me] * scale   (sx = ne/2 for Greengard_Vico, 0 otherwise)
__global__ void k_bs_permute_g(const cd* __restrict__ gin, cd* __restrict__ gout, long ne, long me, long sx, long sy,
                               double scale) {
    const long total = ne * me;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const long ky = idx % me, kx = idx / me;
        const cd v = gin[(kx + sx) % ne + ne * ((ky + sy) % me)];
        gout[idx] = make_double2(v.x * scale, v.y * scale);
    }
}

#define BS_DISPATCH(NB_, CALL)                                                          \
    switch (NB_) {                                                                      \
        case 64:   CALL(64); break;                                                     \
        case 128:  CALL(128); break;                                                    \
        case 256:  CALL(256); break;                                                    \
        case 512:  CALL(512); break;                                                    \
        case 1024: CALL(1024); break;                                                   \
        case 2048: CALL(2048); break;                                                   \
        case 4096: CALL(4096); break;                                                   \
        default: set_error("unsupported Bluestein length %d", (int)(NB_)); return LS_ERR_UNSUPPORTED; \
    }

template <class K> cudaError_t bs_attr(K kernel, int smem) {
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
}

// exp(i pi t^2 / Lf) with the angle reduced exactly: t^2 mod 2 Lf in integers
cd chirp(long t, long Lf) {
    const long m = 2 * Lf;
    long r = (long)(((__int128)t * t) % m);
    const long double ang = 3.14159265358979323846264338327950288L * (long double)r / (long double)Lf;
    return make_double2((double)cosl(ang), (double)sinl(ang));
}

int setup_dim(Op2DGeneric* op, BsDim& d, long nin, long Lf, long o0) {
    d.nin = nin; d.Lf = Lf; d.o0 = o0;
    int Nb = 64;
    while (Nb < Lf + nin - 1) Nb *= 2;
    LS_REQUIRE(Nb <= 4096, LS_ERR_UNSUPPORTED,
               "padded length %ld with %ld inputs needs a Bluestein length > 4096 (general-size GPU path serves ne + n - 1 <= 4096)",
               Lf, nin);
    d.Nb = Nb;
    int rc;
    auto tab = engine_table(Nb);
    if ((rc = op->dupload((void**)&d.d_tab, tab.data(), tab.size() * sizeof(cd)))) return rc;
    std::vector<cd> ch((size_t)Lf);
    for (long t = 0; t < Lf; ++t) ch[(size_t)t] = chirp(t, Lf);
    if ((rc = op->dupload((void**)&d.d_ch, ch.data(), ch.size() * sizeof(cd)))) return rc;
    // forward kernel h_t = c_t, t in (-nin, Lf); inverse kernel h'_t = conj(c_t), t in (o0 - Lf, o0 + nin)
    std::vector<cd> hf((size_t)Nb, make_double2(0.0, 0.0)), hi((size_t)Nb, make_double2(0.0, 0.0));
    for (long t = -(nin - 1); t < Lf; ++t) hf[(size_t)((t % Nb + Nb) % Nb)] = chirp(t < 0 ? -t : t, Lf);
    for (long t = o0 - (Lf - 1); t < o0 + nin; ++t) {
        cd c = chirp(t < 0 ? -t : t, Lf);
        hi[(size_t)((t % Nb + Nb) % Nb)] = make_double2(c.x, -c.y);
    }
    cd *d_h = nullptr;
    if ((rc = op->dmalloc((void**)&d.d_hf, (size_t)Nb * sizeof(cd)))) return rc;
    if ((rc = op->dmalloc((void**)&d.d_hi, (size_t)Nb * sizeof(cd)))) return rc;
    if ((rc = op->dmalloc((void**)&d_h, (size_t)Nb * sizeof(cd)))) return rc;
    for (int which = 0; which < 2; ++which) {
        LS_CUDA_TRY(cudaMemcpyAsync(d_h, which ? hi.data() : hf.data(), (size_t)Nb * sizeof(cd), cudaMemcpyHostToDevice, op->stream));
        cd* dst = which ? d.d_hi : d.d_hf;
#define CALLK(NB)                                                                                            \
        {                                                                                                    \
            LS_CUDA_TRY(bs_attr(k_bs_kernel_spectrum<NB>, BsGeo<NB>::smem));                                 \
            k_bs_kernel_spectrum<NB><<<1, BsGeo<NB>::THREADS, BsGeo<NB>::smem, op->stream>>>(d_h, dst, d.d_tab); \
        }
        BS_DISPATCH(Nb, CALLK);
        LS_CUDA_TRY(cudaStreamSynchronize(op->stream));
    }
    op->dfree(d_h);
    return LS_OK;
}

}  // namespace

int Op2DGeneric::apply_dev(const cd* b, cd* y, int mode) {
    const bool full = (mode == LS_APPLY_FASTCONVOLUTION);
    const bool trap = (quadrule == LS_QUAD_TRAPEZOIDAL);
    // Q2: FFTconvolution multiplies by nu in the trapezoidal branch only (FastConvolution.jl:122 vs :141)
    const double* nuptr = (full || trap) ? d_nu : nullptr;
    const long big = 1L << 40;
    {   // G1: columns j: in b[n*j + i]; out A[ne*j + kx]
        LineAddr la{big, n, 0, 1, ne, 0, 1};
        phase_begin(0);
#define CALL1(NB)                                                                                             \
        {                                                                                                     \
            LS_CUDA_TRY(bs_attr(k_bs_fwd<NB>, BsGeo<NB>::smem));                                              \
            const long grid = (m + BsGeo<NB>::LPC - 1) / BsGeo<NB>::LPC;                                      \
            k_bs_fwd<NB><<<(unsigned)grid, BsGeo<NB>::THREADS, BsGeo<NB>::smem, stream>>>(                    \
                b, nuptr, d_A, X.d_tab, X.d_ch, X.d_hf, (int)n, (int)ne, m, la);                              \
        }
        BS_DISPATCH(X.Nb, CALL1);
        phase_end(); launches++;
    }
    {   // G2: rows kx: in A[kx + ne*j]; out C[m*kx + j]
        LineAddr la{big, 1, 0, ne, m, 0, 1};
        phase_begin(1);
#define CALL2(NB)                                                                                             \
        {                                                                                                     \
            LS_CUDA_TRY(bs_attr(k_bs_mid<NB>, BsGeo<NB>::smem));                                              \
            const long grid = (ne + BsGeo<NB>::LPC - 1) / BsGeo<NB>::LPC;                                     \
            k_bs_mid<NB><<<(unsigned)grid, BsGeo<NB>::THREADS, BsGeo<NB>::smem, stream>>>(                    \
                d_A, d_C, d_G, Y.d_tab, Y.d_ch, Y.d_hf, Y.d_hi, (int)m, (int)me, (int)Y.o0, ne, la);          \
        }
        BS_DISPATCH(Y.Nb, CALL2);
        phase_end(); launches++;
    }
    {   // G3: columns j: in C[j + m*kx]; out y[n*j + i]
        LineAddr la{big, 1, 0, m, n, 0, 1};
        phase_begin(2);
#define CALL3(NB)                                                                                             \
        {                                                                                                     \
            LS_CUDA_TRY(bs_attr(k_bs_inv<NB>, BsGeo<NB>::smem));                                              \
            const long grid = (m + BsGeo<NB>::LPC - 1) / BsGeo<NB>::LPC;                                      \
            k_bs_inv<NB><<<(unsigned)grid, BsGeo<NB>::THREADS, BsGeo<NB>::smem, stream>>>(                    \
                d_C, full ? b : nullptr, y, X.d_tab, X.d_ch, X.d_hi, (int)n, (int)ne, (int)X.o0,              \
                full ? omega * omega : 1.0, m, la);                                                           \
        }
        BS_DISPATCH(X.Nb, CALL3);
        phase_end(); launches++;
    }
    LS_CUDA_TRY(cudaGetLastError());
    return LS_OK;
}

namespace ls {

int create_op2d_generic(ls_handle* out, int64_t n, int64_t m, int64_t ne, int64_t me, const double* nu,
                        const ls_cdouble* gfft, double omega, int quadrule) {
    // size check first (no CUDA call needed to refuse)
    LS_REQUIRE(ne + n - 1 <= 4096 && me + m - 1 <= 4096, LS_ERR_UNSUPPORTED,
               "ls_op2d_create: n=%ld, m=%ld (padded %ld x %ld): the general-size GPU path serves ne + n - 1 <= 4096; "
               "larger grids need the power-of-two Greengard_Vico fast path", (long)n, (long)m, (long)ne, (long)me);
    Op2DGeneric* op = new Op2DGeneric();
    int rc = op->init_base(KIND_OP2D);
    if (rc) { delete op; return rc; }
    op->n = n; op->m = m; op->ne = ne; op->me = me; op->omega = omega; op->quadrule = quadrule;
    const bool trap = (quadrule == LS_QUAD_TRAPEZOIDAL);
#define TRY(x) do { rc = (x); if (rc) { delete op; return rc; } } while (0)
    // kept outputs: Greengard_Vico [1:n] (FastConvolution.jl:101), trapezoidal [n:2n-1] (:82)
    TRY(setup_dim(op, op->X, n, ne, trap ? n - 1 : 0));
    TRY(setup_dim(op, op->Y, m, me, trap ? m - 1 : 0));
    const size_t N = (size_t)n * m, NE = (size_t)ne * me;
    TRY(op->dupload((void**)&op->d_nu, nu, N * sizeof(double)));
    {
        cd* d_gin = nullptr;
        TRY(op->dupload((void**)&d_gin, gfft, NE * sizeof(cd)));
        TRY(op->dmalloc((void**)&op->d_G, NE * sizeof(cd)));
        // ifft normalisation 1/(ne me) and the 1/Nb of each of the four circular convolutions
        const double scale = 1.0 / ((double)ne * (double)me) / ((double)op->X.Nb * (double)op->X.Nb)
                             / ((double)op->Y.Nb * (double)op->Y.Nb);
        // fftshift/ifftshift pair (FastConvolution.jl:94,98): for even ne a roll by ne/2; none for trapezoidal
        const long sx = trap ? 0 : ne / 2, sy = trap ? 0 : me / 2;
        k_bs_permute_g<<<148 * 8, 256, 0, op->stream>>>(d_gin, op->d_G, ne, me, sx, sy, scale);
        cudaError_t e = cudaStreamSynchronize(op->stream);
        if (e != cudaSuccess) { set_error("spectrum permutation failed: %s", cudaGetErrorString(e)); delete op; return LS_ERR_CUDA; }
        op->dfree(d_gin);
    }
    TRY(op->dmalloc((void**)&op->d_A, (size_t)ne * m * sizeof(cd)));
    TRY(op->dmalloc((void**)&op->d_C, (size_t)ne * m * sizeof(cd)));
    TRY(op->dmalloc((void**)&op->d_b, N * sizeof(cd)));
    TRY(op->dmalloc((void**)&op->d_y, N * sizeof(cd)));
#undef TRY
    *out = reinterpret_cast<ls_handle>(op);
    return LS_OK;
}

}  // namespace ls
