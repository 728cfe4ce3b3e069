// Bluestein line transforms on the power-of-two engine (shared by the general-size 2-D and 3-D operators).
//
// A line DFT of arbitrary length Lf with nin non-zero inputs (and nout kept outputs of the inverse) is
//       X_q = conj(c_q) * sum_j [x_j conj(c_j)] c_{q-j},      c_t = exp(i pi t^2 / Lf)
// i.e. a circular convolution of power-of-two length Nb >= Lf + nin - 1: chirp multiply -> fft_fwd<Nb> ->
// multiply by the precomputed chirp spectrum (slot order) -> fft_inv<Nb> -> chirp multiply, in one CTA.
#pragma once
#include "ls_common.cuh"
#include "line_kernels.cuh"

namespace lsb {
using namespace ls;
using namespace lsk;

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

template <class K> inline cudaError_t bs_attr(K kernel, int smem) {
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
}

// exp(i pi t^2 / Lf) with the angle reduced exactly: t^2 mod 2 Lf in integers
inline cd chirp(long t, long Lf) {
    const long m = 2 * Lf;
    long r = (long)(((__int128)t * t) % m);
    const long double ang = 3.14159265358979323846264338327950288L * (long double)r / (long double)Lf;
    return make_double2((double)cosl(ang), (double)sinl(ang));
}

inline int setup_dim(ls::HandleBase* op, BsDim& d, long nin, long Lf, long o0) {
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


}  // namespace lsb
