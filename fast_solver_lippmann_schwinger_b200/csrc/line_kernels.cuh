// Pruned-FFT line kernels of the Lippmann-Schwinger apply (shared by the 2-D and 3-D operators).
//
// The reference transforms the full zero-padded array (FastConvolution.jl:89-98).  With the
// Greengard-Vico padding factor 4 only 1/4 of every padded line is non-zero on the way in and
// only 1/4 is kept on the way out, and a length-4N DFT of an N-point signal splits exactly
// into four length-N DFTs of the modulated signal:
//     X[4q + r] = FFT_N( x[j] * w_{4N}^{r j} )[q],        r = 0..3
//     y[j]      = 1/4 * sum_r conj(w_{4N}^{r j}) * IFFT_N( Y[4q + r] )[j],   j < N.
// A padded line is therefore stored as 4 blocks (r) of N "slots" (the engine's output order).
//
// Line addressing: point p of line L sits at base + L*ls + p*es (elements of 16 B).
//   mode A kernels: the threads of one line are consecutive lanes (contiguous lines);
//   mode B kernels: eight adjacent lines are interleaved across lanes (lane%8 = line) so that
//                   strided lines (es large, ls == 1) are read/written as 128-byte segments.
#pragma once
#include "fft_engine.cuh"

namespace lsk {
using namespace lsfft;

template <int N> struct GeoA {
    static constexpr int E = Cfg<N>::E;
    static constexpr int T = N / E;
    static constexpr int LPC = (T >= 128) ? 1 : 128 / T;   // lines per CTA
    static constexpr int THREADS = T * LPC;
};
template <int N> struct GeoB {
    static constexpr int E = Cfg<N>::E;
    static constexpr int T = N / E;
    static constexpr int GPC = (T >= 16) ? 1 : 16 / T;     // groups of 8 lines per CTA
    static constexpr int LPC = 8 * GPC;
    static constexpr int THREADS = 8 * T * GPC;
};

// thread -> (line within CTA, thread within line, smem layout)
template <int N, bool MODE_B> struct Map;
template <int N> struct Map<N, false> {
    typedef GeoA<N> G;
    typedef LayA<N> Lay;
    int line, t;
    Lay lay;
    __device__ __forceinline__ Map() {
        line = threadIdx.x / G::T;
        t = threadIdx.x % G::T;
        lay.base = line * N;
    }
};
template <int N> struct Map<N, true> {
    typedef GeoB<N> G;
    typedef LayB<N> Lay;
    int line, t;
    Lay lay;
    __device__ __forceinline__ Map() {
        int grp = threadIdx.x / (8 * G::T);
        int w = threadIdx.x % (8 * G::T);
        lay.lam = w & 7;
        t = w >> 3;
        line = grp * 8 + lay.lam;
        smoff = grp * 8 * N;
    }
    int smoff;
};
template <int N> __device__ __forceinline__ int sm_group_off(const Map<N, false>&) { return 0; }
template <int N> __device__ __forceinline__ int sm_group_off(const Map<N, true>& m) { return m.smoff; }

// ---- forward, pruned: N inputs -> 4N slots -------------------------------------------
// in  : line L point j at in[L*in_ls + j*in_es], optionally scaled by the real nu (same addressing)
// out : line L slot  s at out[L*out_ls + s*out_es], s = r*N + slot
template <int N, bool MODE_B>
__global__ void __launch_bounds__(MODE_B ? GeoB<N>::THREADS : GeoA<N>::THREADS)
k_fwd_pruned(const cd* __restrict__ in, const double* __restrict__ nu, cd* __restrict__ out,
             const cd* __restrict__ W, const cd* __restrict__ MOD,
             long in_ls, long in_es, long out_ls, long out_es, long line0) {
    typedef Map<N, MODE_B> M;
    constexpr int E = Cfg<N>::E, T = N / E;
    extern __shared__ cd sm[];
    M mp;
    cd* ex = sm + sm_group_off(mp);
    const long L = line0 + (long)blockIdx.x * M::G::LPC + mp.line;
    const int t = mp.t;
    cd x[E];
#pragma unroll
    for (int a = 0; a < E; ++a) {
        long off = L * in_ls + (long)(a * T + t) * in_es;
        cd val = in[off];
        if (nu != nullptr) {
            double s = nu[off];
            val.x *= s;
            val.y *= s;
        }
        x[a] = val;
    }
#pragma unroll 1
    for (int r = 0; r < 4; ++r) {
        cd v[E];
        if (r == 0) {
#pragma unroll
            for (int a = 0; a < E; ++a) v[a] = x[a];
        } else {
#pragma unroll
            for (int a = 0; a < E; ++a) v[a] = cmul(x[a], __ldg(&MOD[(r - 1) * N + a * T + t]));
        }
        fft_fwd<N>(v, t, ex, mp.lay, W);
        cd* o = out + L * out_ls + (long)(r * N + t) * out_es;
#pragma unroll
        for (int e = 0; e < E; ++e) o[(long)(T * e) * out_es] = v[e];
        __syncthreads();   // next forward rewrites the exchange buffer at other addresses
    }
}

// ---- middle, fused: 4x (forward, multiply by the Green's spectrum, inverse) --------------
// in  : line L point j at in[L*in_ls + j*in_es]
// G   : spectrum of line L, block r, slot s at G[(L*4 + r)*N + s]                 (mode A)
//       or at G[((L/8)*4 + r)*8N + s*8 + L%8]                                     (mode B)
// out : line L point j at out[L*out_ls + j*out_es]  (may alias in when strides agree)
template <int N, bool MODE_B>
__global__ void __launch_bounds__(MODE_B ? GeoB<N>::THREADS : GeoA<N>::THREADS)
k_mid_fused(const cd* in, cd* out, const cd* __restrict__ G,
            const cd* __restrict__ W, const cd* __restrict__ MOD,
            long in_ls, long in_es, long out_ls, long out_es, long line0) {
    typedef Map<N, MODE_B> M;
    constexpr int E = Cfg<N>::E, T = N / E;
    extern __shared__ cd sm[];
    M mp;
    cd* ex = sm + sm_group_off(mp);
    cd* xs = sm + M::G::LPC * N + sm_group_off(mp);   // thread-private copy of the input line
    const long L = line0 + (long)blockIdx.x * M::G::LPC + mp.line;
    const int t = mp.t;
    {
        const cd* p = in + L * in_ls + (long)t * in_es;
#pragma unroll
        for (int a = 0; a < E; ++a) xs[mp.lay.phys(a * T + t)] = p[(long)(a * T) * in_es];
    }
    cd acc[E];
#pragma unroll
    for (int a = 0; a < E; ++a) acc[a] = make_double2(0.0, 0.0);
    const cd* g;
    long gstep;
    if (MODE_B) {
        g = G + ((L >> 3) * 4) * (long)(8 * N) + (long)t * 8 + (L & 7);
        gstep = 8;
    } else {
        g = G + (L * 4) * (long)N + t;
        gstep = 1;
    }
#pragma unroll 1
    for (int r = 0; r < 4; ++r) {
        cd v[E];
        if (r == 0) {
#pragma unroll
            for (int a = 0; a < E; ++a) v[a] = xs[mp.lay.phys(a * T + t)];
        } else {
#pragma unroll
            for (int a = 0; a < E; ++a)
                v[a] = cmul(xs[mp.lay.phys(a * T + t)], __ldg(&MOD[(r - 1) * N + a * T + t]));
        }
        fft_fwd<N>(v, t, ex, mp.lay, W);
        const cd* gr = g + (long)r * N * gstep;
#pragma unroll
        for (int e = 0; e < E; ++e) v[e] = cmul(v[e], __ldg(&gr[(long)(T * e) * gstep]));
        fft_inv<N>(v, t, ex, mp.lay, W);
        if (r == 0) {
#pragma unroll
            for (int a = 0; a < E; ++a) acc[a] = v[a];
        } else {
#pragma unroll
            for (int a = 0; a < E; ++a) acc[a] = cfmac(v[a], __ldg(&MOD[(r - 1) * N + a * T + t]), acc[a]);
        }
    }
    cd* o = out + L * out_ls + (long)t * out_es;
#pragma unroll
    for (int a = 0; a < E; ++a) o[(long)(a * T) * out_es] = acc[a];
}

// ---- inverse, pruned: 4N slots -> N outputs, optional identity-plus-contrast combine --
// in  : line L slot s at in[L*in_ls + s*in_es]
// out : line L point j at out[L*out_ls + j*out_es];  if bsrc: out = bsrc + scale*result
template <int N, bool MODE_B>
__global__ void __launch_bounds__(MODE_B ? GeoB<N>::THREADS : GeoA<N>::THREADS)
k_inv_pruned(const cd* __restrict__ in, const cd* bsrc, cd* out,
             const cd* __restrict__ W, const cd* __restrict__ MOD, double scale,
             long in_ls, long in_es, long out_ls, long out_es, long line0) {
    typedef Map<N, MODE_B> M;
    constexpr int E = Cfg<N>::E, T = N / E;
    extern __shared__ cd sm[];
    M mp;
    cd* ex = sm + sm_group_off(mp);
    const long L = line0 + (long)blockIdx.x * M::G::LPC + mp.line;
    const int t = mp.t;
    cd acc[E];
#pragma unroll 1
    for (int r = 0; r < 4; ++r) {
        cd v[E];
        const cd* p = in + L * in_ls + (long)(r * N + t) * in_es;
#pragma unroll
        for (int e = 0; e < E; ++e) v[e] = p[(long)(T * e) * in_es];
        fft_inv<N>(v, t, ex, mp.lay, W);
        if (r == 0) {
#pragma unroll
            for (int a = 0; a < E; ++a) acc[a] = v[a];
        } else {
#pragma unroll
            for (int a = 0; a < E; ++a) acc[a] = cfmac(v[a], __ldg(&MOD[(r - 1) * N + a * T + t]), acc[a]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int a = 0; a < E; ++a) {
        long off = L * out_ls + (long)(a * T + t) * out_es;
        cd res = cscale(acc[a], scale);
        if (bsrc != nullptr) res = cadd(res, bsrc[off]);
        out[off] = res;
    }
}

template <int N, bool MODE_B> constexpr int smem_fwd() {
    return (MODE_B ? GeoB<N>::LPC : GeoA<N>::LPC) * N * (int)sizeof(cd);
}
template <int N, bool MODE_B> constexpr int smem_mid() { return 2 * smem_fwd<N, MODE_B>(); }

}  // namespace lsk
