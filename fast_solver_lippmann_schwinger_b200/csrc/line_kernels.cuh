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
#include <cooperative_groups.h>

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
    static constexpr int LB = LinesB<N>::value;             // lines interleaved across lanes (8; 4 for 512-point lines)
    static constexpr int GPC = (LB * T >= 128) ? 1 : 128 / (LB * T);     // groups of LB lines per CTA
    static constexpr int LPC = LB * GPC;
    static constexpr int THREADS = LB * T * GPC;
};
// resident CTAs per SM asked of ptxas: 128-thread CTAs 3 (168 registers); the 256-thread CTAs of the 4-line mode-B
// geometry 2 (128 registers: E = 8 points per thread there); everything else 1
#ifndef LS_INV512_MINB
#define LS_INV512_MINB 2        // measured (profiles/r2_l_512_variants.log): 512^3 P4 3.67 -> 3.19 ms, 2-D 512^2 apply 0.0353 -> 0.0332 ms
#endif
template <int N, bool MODE_B> struct MinB {
    static constexpr int TH = MODE_B ? GeoB<N>::THREADS : GeoA<N>::THREADS;
    static constexpr int value = (TH <= 128) ? 3 : ((MODE_B && GeoB<N>::LB == 4 && TH == 256) ? 2 : 1);
};
// the inverse pass keeps E accumulators next to the E working values: at 512 points (E = 16 since r2) the 168-register
// cap of three CTAs per SM spills inside the loop
template <int N, bool MODE_B> struct MinBInv {
    static constexpr int value = (N == 512 && MinB<N, MODE_B>::value == 3) ? LS_INV512_MINB : MinB<N, MODE_B>::value;
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
    __device__ __forceinline__ int lay_lam() const { return 0; }
};
template <int N> struct Map<N, true> {
    typedef GeoB<N> G;
    typedef LayB<N> Lay;
    int line, t;
    Lay lay;
    __device__ __forceinline__ Map() {
        constexpr int LB = G::LB;
        int grp = threadIdx.x / (LB * G::T);
        int w = threadIdx.x % (LB * G::T);
        lay.lam = w % LB;
        t = w / LB;
        line = grp * LB + lay.lam;
        smoff = grp * LB * N;
    }
    int smoff;
    __device__ __forceinline__ int lay_lam() const { return lay.lam; }
};
template <int N> __device__ __forceinline__ int sm_group_off(const Map<N, false>&) { return 0; }
template <int N> __device__ __forceinline__ int sm_group_off(const Map<N, true>& m) { return m.smoff; }

// Two-level line addressing: line L = l0 + ldim0*l1 starts at l0*ls0 + l1*ls1; its points/slots
// are `es` elements apart.  (2-D uses one level; 3-D y-lines are indexed by (x slot, z plane).)
struct LineAddr {
    long ldim0;
    long in_ls0, in_ls1, in_es;
    long out_ls0, out_ls1, out_es;
    // slab split of the 4N-slot axis (multi-GPU transposes): slot s lives at
    // (s & (2^shift - 1))*es + (s >> shift)*split_stride.  shift = 62 disables it.
    int split_shift = 62;
    long split_stride = 0;
    // optional second level (x-slot chunks inside a rank's slab, pipelined transposes): bits
    // [split_shift, split2_shift) select the chunk (stride split_stride), bits >= split2_shift the
    // rank (stride split2_stride).  split2_shift = 62 leaves the one-level rule above.
    int split2_shift = 62;
    long split2_stride = 0;
    // padding factor of the line: nr = 4 sub-transforms r = 0..3 (the reference's 4N zero padding) or
    // nr = 2 (r = 0, 2 in units of w_4N: a 2N padding, enough once the kernel is restricted to lags (-N, N))
    int nr = 4;
    int cblock = 0;     // inverse: which N-block [cblock*N, (cblock+1)*N) of the padded output line is kept
    int full2 = 0;      // forward: the input line has 2N points (no zero padding), nr must be 2
};
__device__ __forceinline__ long slot_off(const LineAddr& a, long s, long es) {
    return (s & ((1L << a.split_shift) - 1)) * es
         + ((s & ((1L << a.split2_shift) - 1)) >> a.split_shift) * a.split_stride
         + (s >> a.split2_shift) * a.split2_stride;
}
__device__ __forceinline__ long line_in(const LineAddr& a, long L) { return (L % a.ldim0) * a.in_ls0 + (L / a.ldim0) * a.in_ls1; }
__device__ __forceinline__ long line_out(const LineAddr& a, long L) { return (L % a.ldim0) * a.out_ls0 + (L / a.ldim0) * a.out_ls1; }

// per-CTA shared memory: [exchange: LPC*N][(mid only) x copy: LPC*N][(mid only) spectrum: LPC*N][tw1][mbarrier]
template <int N, bool MODE_B> struct Smem {
    static constexpr int LPC = MODE_B ? GeoB<N>::LPC : GeoA<N>::LPC;
    static constexpr int TW1N = EngTab<N>::TW1N;
    static constexpr int fwd_bytes = (LPC * N + TW1N) * (int)sizeof(cd);
    static constexpr int mid_bytes = (3 * LPC * N + TW1N) * (int)sizeof(cd) + 16;
    static constexpr int mid_bytes_direct = (2 * LPC * N + TW1N) * (int)sizeof(cd) + 16;
};

// ---- forward, pruned: N inputs -> 4N slots -------------------------------------------
// in  : line L point j at in[L*in_ls + j*in_es], optionally scaled by the real nu (same addressing)
// out : line L slot  s at out[L*out_ls + s*out_es], s = r*N + slot
template <int N, bool MODE_B>
__global__ void __launch_bounds__(MODE_B ? GeoB<N>::THREADS : GeoA<N>::THREADS, MinB<N, MODE_B>::value)
k_fwd_pruned(const cd* __restrict__ in, const double* __restrict__ nu, cd* __restrict__ out,
             const cd* __restrict__ TAB, const LineAddr la, long line0) {
    typedef Map<N, MODE_B> M;
    constexpr int E = Cfg<N>::E, T = N / E, LPC = M::G::LPC;
    extern __shared__ __align__(128) cd sm[];
    M mp;
    cd* ex = sm + sm_group_off(mp);
    cd* tw1 = sm + LPC * N;
    load_tw1<N>(tw1, TAB);
    const long L = line0 + (long)blockIdx.x * LPC + mp.line;
    const int t = mp.t;
    const TwState<N> tw = make_tw<N>(t, TAB, tw1);
    const long in_base = line_in(la, L), out_base = line_out(la, L);
    const long in_es = la.in_es, out_es = la.out_es;
    cd x[E];
#pragma unroll
    for (int a = 0; a < E; ++a) {
        long off = in_base + (long)(a * T + t) * in_es;
        cd val = in[off];
        if (nu != nullptr) {
            double s = nu[off];
            val.x *= s;
            val.y *= s;
        }
        x[a] = val;
    }
    __syncthreads();   // tw1 visible
    const int nr = la.nr, rstep = 4 / la.nr;
#pragma unroll 1
    for (int rr = 0; rr < nr; ++rr) {
        const int r = rr * rstep;
        cd v[E];
        if (la.full2) {
            // unpadded 2N-point line (setup only): first DIF stage x[j] +- x[j + N]
#pragma unroll
            for (int a = 0; a < E; ++a) {
                const cd hi = in[in_base + (long)(N + a * T + t) * in_es];
                v[a] = rr == 0 ? cadd(x[a], hi) : csub(x[a], hi);
            }
        } else {
#pragma unroll
            for (int a = 0; a < E; ++a) v[a] = x[a];
        }
        fft_fwd<N>(v, t, r, ex, mp.lay, tw);
        cd* o = out + out_base;
#pragma unroll
        for (int e = 0; e < E; ++e) o[slot_off(la, (long)(rr * N + t + T * e), out_es)] = v[e];
        __syncthreads();   // next forward rewrites the exchange buffer at other addresses
    }
}

// ---- middle, fused: 4x (forward, multiply by the Green's spectrum, inverse) --------------
// in  : line L point j at in[L*in_ls + j*in_es]
// G   : spectrum of line L, block r, slot s at G[(L*4 + r)*N + s]                 (mode A)
//       or at G[((L/8)*4 + r)*8N + s*8 + L%8]                                     (mode B)
//       -> per (unit, r) one contiguous chunk, fetched by TMA bulk copy into shared memory
//          while the previous block's inverse transform runs.
// out : line L point j at out[L*out_ls + j*out_es]  (may alias in when strides agree)
// ASM: how many of the E per-thread accumulators live in (thread-private) shared memory instead of
//      registers - the register diet that lets a third CTA fit on the SM.
template <int N, bool MODE_B, bool GSM = true, int MINB = 1, int ASM = 0>
__global__ void __launch_bounds__(MODE_B ? GeoB<N>::THREADS : GeoA<N>::THREADS, MINB)
k_mid_fused(const cd* in, cd* out, const cd* __restrict__ G, const cd* __restrict__ TAB,
            const LineAddr la, long line0) {
    typedef Map<N, MODE_B> M;
    constexpr int E = Cfg<N>::E, T = N / E, LPC = M::G::LPC;
    constexpr int AR = E - ASM, TH = MODE_B ? GeoB<N>::THREADS : GeoA<N>::THREADS;
    constexpr int LBK = MODE_B ? GeoB<N>::LB : 1;     // lines interleaved in a spectrum chunk
    constexpr int UNIT = LBK * N;                     // points per spectrum chunk
    constexpr int UPC = LPC * N / UNIT;               // chunks per CTA and r
    extern __shared__ __align__(128) cd sm[];
    M mp;
    cd* ex = sm + sm_group_off(mp);
    cd* xs = sm + LPC * N + sm_group_off(mp);   // thread-private copy of the input line
    cd* gb = sm + 2 * LPC * N;                  // spectrum chunk(s) of the current r (GSM only)
    cd* accs = sm + (GSM ? 3 : 2) * LPC * N + threadIdx.x;   // accumulator tail, element k at accs[k*TH]
    cd* tw1 = sm + (GSM ? 3 : 2) * LPC * N + ASM * TH;
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(tw1 + Smem<N, MODE_B>::TW1N);
    const long Lcta = line0 + (long)blockIdx.x * LPC;
    const long L = Lcta + mp.line;
    const int t = mp.t;
    const int nr = la.nr, rstep = 4 / la.nr;
    const cd* gsrc = G + (Lcta / LBK) * (long)nr * UNIT;   // chunk (unit u, rr) at gsrc + (u*nr + rr)*UNIT

    auto issue_g = [&](int rr) {
        mbar_expect_tx(bar, (unsigned)(UPC * UNIT * sizeof(cd)));
#pragma unroll
        for (int u = 0; u < UPC; ++u)
            bulk_g2s(gb + u * UNIT, gsrc + ((long)u * nr + rr) * UNIT, (unsigned)(UNIT * sizeof(cd)), bar);
    };
    if (GSM && threadIdx.x == 0) {
        mbar_init(bar, 1);
        fence_proxy_async();
        issue_g(0);
    }
    load_tw1<N>(tw1, TAB);
    const TwState<N> tw = make_tw<N>(t, TAB, tw1);
    {
        const cd* p = in + line_in(la, L) + (long)t * la.in_es;
#pragma unroll
        for (int a = 0; a < E; ++a) xs[mp.lay.phys(a * T + t)] = p[(long)(a * T) * la.in_es];
    }
    // direct-load variant: pull spectrum chunk r into L2 one r ahead (one request per 128-byte line)
    auto prefetch_g = [&](int rr) {
        if (!GSM && rr < nr) {
            const cd* g = MODE_B ? gsrc + ((long)(mp.line / LBK) * nr + rr) * UNIT + mp.lay_lam()
                                 : gsrc + ((long)mp.line * nr + rr) * UNIT;
            constexpr int gs = LBK;
            if (MODE_B ? (mp.lay_lam() == 0) : ((t & 7) == 0)) {
#pragma unroll
                for (int e = 0; e < E; ++e) prefetch_l2(&g[(t + T * e) * gs]);
            }
        }
    };
    prefetch_g(0);
    __syncthreads();   // tw1 + mbarrier init visible
    cd acc[AR];
    const int goff = MODE_B ? 0 : mp.line * N;     // mode B: lay.phys already interleaves the 8 lines
#pragma unroll 1
    for (int rr = 0; rr < nr; ++rr) {
        const int r = rr * rstep;
        cd v[E];
#pragma unroll
        for (int a = 0; a < E; ++a) v[a] = xs[mp.lay.phys(a * T + t)];
        prefetch_g(rr + 1);
        if (!GSM && MINB >= 3) {
            // three CTAs per SM: registers are the scarce resource, the other CTAs hide the load latency
            fft_fwd<N>(v, t, r, ex, mp.lay, tw);
            const cd* g = MODE_B ? gsrc + ((long)(mp.line / LBK) * nr + rr) * UNIT + mp.lay_lam()
                                 : gsrc + ((long)mp.line * nr + rr) * UNIT;
            constexpr int gs = LBK;
#pragma unroll
            for (int e = 0; e < E; ++e) v[e] = cmul(v[e], __ldg(&g[(t + T * e) * gs]));
        } else if (!GSM) {
            // spectrum values are requested before the last butterfly stage and consumed after it
            cd gv[E];
            fft_fwd<N>(v, t, r, ex, mp.lay, tw, [&]() {
                const cd* g = MODE_B ? gsrc + ((long)(mp.line / LBK) * nr + rr) * UNIT + mp.lay_lam()
                                     : gsrc + ((long)mp.line * nr + rr) * UNIT;
                constexpr int gs = LBK;
#pragma unroll
                for (int e = 0; e < E; ++e) gv[e] = __ldg(&g[(t + T * e) * gs]);
            });
#pragma unroll
            for (int e = 0; e < E; ++e) v[e] = cmul(v[e], gv[e]);
        } else {
        fft_fwd<N>(v, t, r, ex, mp.lay, tw);
        mbar_wait(bar, (unsigned)(rr & 1));
        if (MODE_B) {
            const cd* g = gb + sm_group_off(mp);
#pragma unroll
            for (int e = 0; e < E; ++e) v[e] = cmul(v[e], g[(t + T * e) * LBK + mp.lay_lam()]);
        } else {
            const cd* g = gb + goff;
#pragma unroll
            for (int e = 0; e < E; ++e) v[e] = cmul(v[e], g[t + T * e]);
        }
        }
        fft_inv<N>(v, t, r, ex, mp.lay, tw, [&]() {
            // every thread has consumed gb (the multiply precedes this barrier): refill it
            if (GSM && rr + 1 < nr && threadIdx.x == 0) {
                fence_proxy_async();
                issue_g(rr + 1);
            }
        });
        if (ASM == 0) {
            demod_accumulate<N>(acc, v, r);
        } else if (r == 0) {
#pragma unroll
            for (int a = 0; a < AR; ++a) acc[a] = v[a];
#pragma unroll
            for (int a = AR; a < E; ++a) accs[(a - AR) * TH] = v[a];
        } else {
            acc[0] = cadd(acc[0], v[0]);
#pragma unroll
            for (int a = 1; a < AR; ++a) acc[a] = cfmac(v[a], c64(r * a * (16 / E)), acc[a]);
#pragma unroll
            for (int a = AR; a < E; ++a) accs[(a - AR) * TH] = cfmac(v[a], c64(r * a * (16 / E)), accs[(a - AR) * TH]);
        }
    }
    cd* o = out + line_out(la, L) + (long)t * la.out_es;
#pragma unroll
    for (int a = 0; a < AR; ++a) o[(long)(a * T) * la.out_es] = acc[a];
#pragma unroll
    for (int a = AR; a < E; ++a) o[(long)(a * T) * la.out_es] = accs[(a - AR) * TH];
}

// ---- middle, fused, 2x padding, no accumulator registers ("swap" variant, mode A) ----------------------------
// Same arithmetic and the same bits as k_mid_fused with nr = 2 and direct spectrum loads.  k_mid_fused keeps the
// input line in shared memory AND 16 accumulators in registers (248 registers: two 128-thread CTAs per SM, and the
// pass is latency-bound at 8 warps per SM).  With two sub-transforms only one thread-private buffer is needed: it
// holds the input line x while r = 0 is transformed, then x and the r = 0 result trade places (each thread swaps
// its own 16 elements, no barrier), and the r = 2 result is combined with the held one on the way out.  Without
// the accumulators the kernel fits 168 registers and 66 KB of shared memory: three CTAs per SM.
// GL: how the spectrum chunk reaches the multiply.  0: all E values requested before the last butterfly stage and
// held in registers across it (64 more live registers); 1: the chunk is pulled into L2 by prefetches issued before
// the last butterfly stage and loaded right at the multiply (the other CTAs of the SM cover the L2 latency).
template <int N, int MINB, int GL>
__global__ void __launch_bounds__(GeoA<N>::THREADS, MINB)
k_mid_swap(const cd* in, cd* out, const cd* __restrict__ G, const cd* __restrict__ TAB, const LineAddr la, long line0) {
    typedef Map<N, false> M;
    constexpr int E = Cfg<N>::E, T = N / E, LPC = M::G::LPC, TH = GeoA<N>::THREADS;
    extern __shared__ __align__(128) cd sm[];
    M mp;
    cd* ex = sm;
    cd* hold = sm + LPC * N + threadIdx.x;          // element a of this thread at hold[a*TH]
    cd* tw1 = sm + 2 * LPC * N;
    load_tw1<N>(tw1, TAB);
    const long L = line0 + (long)blockIdx.x * LPC + mp.line;
    const int t = mp.t;
    const TwState<N> tw = make_tw<N>(t, TAB, tw1);
    const cd* gline = G + L * 2L * N;               // chunk rr of line L at gline + rr*N
    cd v[E];
    {
        const cd* p = in + line_in(la, L) + (long)t * la.in_es;
#pragma unroll
        for (int a = 0; a < E; ++a) v[a] = p[(long)(a * T) * la.in_es];
#pragma unroll
        for (int a = 0; a < E; ++a) hold[a * TH] = v[a];
    }
    __syncthreads();   // tw1 visible
#pragma unroll 1
    for (int rr = 0; rr < 2; ++rr) {
        const int r = 2 * rr;
        const cd* g = gline + (long)rr * N;
        if constexpr (GL == 0) {
            cd gv[E];
            fft_fwd<N>(v, t, r, ex, mp.lay, tw, [&]() {
#pragma unroll
                for (int e = 0; e < E; ++e) gv[e] = __ldg(&g[t + T * e]);
            });
#pragma unroll
            for (int e = 0; e < E; ++e) v[e] = cmul(v[e], gv[e]);
        } else {
            fft_fwd<N>(v, t, r, ex, mp.lay, tw, [&]() {
                if ((t & 7) == 0) {       // one request per 128-byte line
#pragma unroll
                    for (int e = 0; e < E; ++e) prefetch_l2(&g[t + T * e]);
                }
            });
#pragma unroll
            for (int e = 0; e < E; ++e) v[e] = cmul(v[e], __ldg(&g[t + T * e]));
        }
        fft_inv<N>(v, t, r, ex, mp.lay, tw);
        if (rr == 0) {
            // x and the r = 0 result trade places
#pragma unroll
            for (int a = 0; a < E; ++a) {
                const cd xa = hold[a * TH];
                hold[a * TH] = v[a];
                v[a] = xa;
            }
        }
    }
    cd* o = out + line_out(la, L) + (long)t * la.out_es;
    o[0] = cadd(hold[0], v[0]);
#pragma unroll
    for (int a = 1; a < E; ++a) o[(long)(a * T) * la.out_es] = cfmac(v[a], c64(2 * a * (16 / E)), hold[a * TH]);
}

// cp.async helpers (inverse pass staging; experiment kernels)
__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int K> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(K) : "memory"); }

// ---- inverse, pruned: 4N slots -> N outputs, optional identity-plus-contrast combine --
// in  : line L slot s at in[L*in_ls + s*in_es]
// out : line L point j at out[L*out_ls + j*out_es];  if bsrc: out = bsrc + scale*result
// PF: contiguous lines only - pull the next slot block (and the b values) into L2 one block ahead
// PF = 1: next block / b values prefetched into L2.  PF = 2: next block / b values staged in shared memory by
// cp.async while the current block is transformed (each thread stages exactly the elements it consumes: no
// extra barrier), which takes two of the kernel's three exposed DRAM latencies off the critical path.
template <int N, bool MODE_B, int PF = 0>
__global__ void __launch_bounds__(MODE_B ? GeoB<N>::THREADS : GeoA<N>::THREADS, MinBInv<N, MODE_B>::value)
k_inv_pruned(const cd* __restrict__ in, const cd* bsrc, cd* out, const cd* __restrict__ TAB, double scale,
             const LineAddr la, long line0) {
    typedef Map<N, MODE_B> M;
    constexpr int E = Cfg<N>::E, T = N / E, LPC = M::G::LPC;
    extern __shared__ __align__(128) cd sm[];
    M mp;
    cd* ex = sm + sm_group_off(mp);
    cd* tw1 = sm + LPC * N;
    load_tw1<N>(tw1, TAB);
    const long L = line0 + (long)blockIdx.x * LPC + mp.line;
    const int t = mp.t;
    const TwState<N> tw = make_tw<N>(t, TAB, tw1);
    __syncthreads();
    const long in_base = line_in(la, L), out_base = line_out(la, L);
    cd acc[E];
    const int nr = la.nr, rstep = 4 / la.nr;
    constexpr int TH = MODE_B ? GeoB<N>::THREADS : GeoA<N>::THREADS;
    cd* stg = sm + LPC * N + EngTab<N>::TW1N + threadIdx.x;      // PF == 2: element e of this thread at stg[TH*e]
#pragma unroll 1
    for (int rr = 0; rr < nr; ++rr) {
        const int r = rr * rstep;
        cd v[E];
        const cd* p = in + in_base;
        if (PF == 2 && rr > 0) {
            cp_async_wait<0>();
#pragma unroll
            for (int e = 0; e < E; ++e) v[e] = stg[TH * e];
        } else {
#pragma unroll
            for (int e = 0; e < E; ++e) v[e] = p[slot_off(la, (long)(rr * N + t + T * e), la.in_es)];
        }
        // pull the next block (or, at the end, the b values of the combine) into L2 while this one is transformed
        // (contiguous lines only: measured on B200, prefetching scattered 16-byte pieces costs more L2
        //  requests than the latency it hides - 2-D P3 0.177 -> 0.223 ms, 3-D P5 0.54 -> 0.44 ms)
        if (PF == 2) {
            if (rr + 1 < nr) {
#pragma unroll
                for (int e = 0; e < E; ++e) cp_async16(&stg[TH * e], &p[slot_off(la, (long)((rr + 1) * N + t + T * e), la.in_es)]);
                cp_async_commit();
            } else if (bsrc != nullptr) {
#pragma unroll
                for (int a = 0; a < E; ++a) cp_async16(&stg[TH * a], &bsrc[out_base + (long)(a * T + t) * la.out_es]);
                cp_async_commit();
            }
        } else if (!PF) {
        } else if (rr + 1 < nr) {
            if (!MODE_B && la.in_es == 1 && (t & 7) == 0) {
#pragma unroll
                for (int e = 0; e < E; ++e) prefetch_l2(&p[slot_off(la, (long)((rr + 1) * N + t + T * e), 1)]);
            }
        } else if (bsrc != nullptr && !MODE_B && la.in_es == 1 && (t & 7) == 0) {
#pragma unroll
            for (int a = 0; a < E; ++a) prefetch_l2(&bsrc[out_base + (long)(a * T + t) * la.out_es]);
        }
        fft_inv<N>(v, t, r, ex, mp.lay, tw);
        if (la.cblock == 0) demod_accumulate<N>(acc, v, r);
        else demod_accumulate_block<N>(acc, v, r, 16 * r * la.cblock);     // kept block c: extra factor w_4^(-r c)
        __syncthreads();
    }
#pragma unroll
    for (int a = 0; a < E; ++a) {
        long off = out_base + (long)(a * T + t) * la.out_es;
        cd res = cscale(acc[a], scale);
        if (bsrc != nullptr) {
            if (PF == 2 && a == 0) cp_async_wait<0>();
            const cd bv = (PF == 2) ? stg[TH * a] : bsrc[off];
            res = make_double2(__fma_rn(acc[a].x, scale, bv.x), __fma_rn(acc[a].y, scale, bv.y));   // one rounding, same bits in every variant
        }
        out[off] = res;
    }
}

}  // namespace lsk

// ---- host-side launchers ----------------------------------------------------------------
namespace lsk {

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device attribute and a process may hold handles on several
// GPUs (ls_set_device): opt in once per (kernel instantiation, device).  `done` = the call site's static bit mask.
template <class K> inline cudaError_t smem_optin(K kernel, int smem, unsigned long long& done) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    const unsigned long long bit = 1ull << (dev & 63);
    if (done & bit) return cudaSuccess;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e == cudaSuccess) done |= bit;
    return e;
}

template <int N, bool B>
inline cudaError_t launch_fwd(cudaStream_t s, long nlines, const cd* in, const double* nu, cd* out, const cd* TAB,
                              const LineAddr& la) {
    constexpr int smem = Smem<N, B>::fwd_bytes;
    constexpr int LPC = Smem<N, B>::LPC, TH = B ? GeoB<N>::THREADS : GeoA<N>::THREADS;
    static unsigned long long optin = 0;
    { cudaError_t e = smem_optin(k_fwd_pruned<N, B>, smem, optin); if (e != cudaSuccess) return e; }
    k_fwd_pruned<N, B><<<(unsigned)(nlines / LPC), TH, smem, s>>>(in, nu, out, TAB, la, 0);
    return cudaPeekAtLastError();
}

template <int N, bool B, bool GSM>
inline cudaError_t launch_mid(cudaStream_t s, long nlines, const cd* in, cd* out, const cd* G, const cd* TAB,
                              const LineAddr& la) {
    constexpr int smem = GSM ? Smem<N, B>::mid_bytes : Smem<N, B>::mid_bytes_direct;
    constexpr int LPC = Smem<N, B>::LPC, TH = B ? GeoB<N>::THREADS : GeoA<N>::THREADS;
    static unsigned long long optin = 0;
    constexpr int MB = (B && GeoB<N>::LB == 4) ? 2 : 1;      // 4-line geometry: two CTAs per SM (128 registers)
    { cudaError_t e = smem_optin(k_mid_fused<N, B, GSM, MB>, smem, optin); if (e != cudaSuccess) return e; }
    k_mid_fused<N, B, GSM, MB><<<(unsigned)(nlines / LPC), TH, smem, s>>>(in, out, G, TAB, la, 0);
    return cudaPeekAtLastError();
}

template <int N, int MINB, int GL>
inline cudaError_t launch_mid_swap(cudaStream_t s, long nlines, const cd* in, cd* out, const cd* G, const cd* TAB,
                                   const LineAddr& la) {
    constexpr int smem = (2 * GeoA<N>::LPC * N + EngTab<N>::TW1N) * (int)sizeof(cd);
    static unsigned long long optin = 0;
    { cudaError_t e = smem_optin(k_mid_swap<N, MINB, GL>, smem, optin); if (e != cudaSuccess) return e; }
    k_mid_swap<N, MINB, GL><<<(unsigned)(nlines / GeoA<N>::LPC), GeoA<N>::THREADS, smem, s>>>(in, out, G, TAB, la, 0);
    return cudaPeekAtLastError();
}

template <int N, bool B, int PF = 0>
inline cudaError_t launch_inv(cudaStream_t s, long nlines, const cd* in, const cd* bsrc, cd* out, const cd* TAB,
                              double scale, const LineAddr& la) {
    constexpr int smem = Smem<N, B>::fwd_bytes + (PF == 2 ? Smem<N, B>::LPC * N * (int)sizeof(cd) : 0);
    constexpr int LPC = Smem<N, B>::LPC, TH = B ? GeoB<N>::THREADS : GeoA<N>::THREADS;
    static unsigned long long optin = 0;
    { cudaError_t e = smem_optin(k_inv_pruned<N, B, PF>, smem, optin); if (e != cudaSuccess) return e; }
    k_inv_pruned<N, B, PF><<<(unsigned)(nlines / LPC), TH, smem, s>>>(in, bsrc, out, TAB, scale, la, 0);
    return cudaPeekAtLastError();
}

}  // namespace lsk
