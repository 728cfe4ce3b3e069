// Measured-and-rejected variants of the fused middle pass, kept for the record (profiles/r1_b_notes.md).
// None of them is on a default path: they are reached only through LS_P2_VARIANT / LS_P3_VARIANT together with
// LS_FLAG_PAD4 (they implement the literal 4x padding only) and produce bit-identical results to k_mid_fused.
//   k_mid_lean       short strided lines without the x copy (3-D z pass)
//   k_mid_persist    persistent CTAs, next input line prefetched by cp.async
//   k_mid_cluster    one sub-transform per CTA, 4-CTA cluster, DSMEM reduction
//   k_mid_fused_dual two sub-transforms software-pipelined per thread
#pragma once
#include "line_kernels.cuh"

namespace lsk {

// ---- middle, fused, lean variant for short strided lines (mode B, 3-D z pass) ------------------------------
// Same arithmetic as k_mid_fused with the spectrum staged by TMA.  The input line group is not copied to
// shared memory: in mode B a line group is read as 128-byte segments, so re-reading it for each of the four
// sub-transforms (L2 hits after the first) is cheap, and without the copy a CTA needs 64 KB + tables instead
// of 96 KB - three CTAs per SM when the registers allow (MINB).
// smem: [exchange: LPC*N][spectrum chunk: LPC*N][tw1][mbarrier]
template <int N, int MINB>
__global__ void __launch_bounds__(GeoB<N>::THREADS, MINB)
k_mid_lean(const cd* in, cd* out, const cd* __restrict__ G, const cd* __restrict__ TAB, const LineAddr la, long line0) {
    typedef Map<N, true> M;
    constexpr int E = Cfg<N>::E, T = N / E, LPC = M::G::LPC;
    constexpr int UNIT = 8 * N, UPC = LPC * N / UNIT;
    extern __shared__ __align__(128) cd sm[];
    M mp;
    cd* ex = sm + sm_group_off(mp);
    cd* gb = sm + LPC * N;
    cd* tw1 = sm + 2 * LPC * N;
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(tw1 + EngTab<N>::TW1N);
    const long Lcta = line0 + (long)blockIdx.x * LPC;
    const long L = Lcta + mp.line;
    const int t = mp.t;
    const cd* gsrc = G + (Lcta >> 3) * 4L * UNIT;
    auto issue_g = [&](int r) {
        mbar_expect_tx(bar, (unsigned)(UPC * UNIT * sizeof(cd)));
#pragma unroll
        for (int u = 0; u < UPC; ++u)
            bulk_g2s(gb + u * UNIT, gsrc + ((long)u * 4 + r) * UNIT, (unsigned)(UNIT * sizeof(cd)), bar);
    };
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        fence_proxy_async();
        issue_g(0);
    }
    load_tw1<N>(tw1, TAB);
    const TwState<N> tw = make_tw<N>(t, TAB, tw1);
    const cd* p = in + line_in(la, L) + (long)t * la.in_es;
    __syncthreads();
    cd acc[E];
#pragma unroll 1
    for (int r = 0; r < 4; ++r) {
        cd v[E];
#pragma unroll
        for (int a = 0; a < E; ++a) v[a] = p[(long)(a * T) * la.in_es];
        fft_fwd<N>(v, t, r, ex, mp.lay, tw);
        mbar_wait(bar, (unsigned)(r & 1));
        const cd* g = gb + sm_group_off(mp);
#pragma unroll
        for (int e = 0; e < E; ++e) v[e] = cmul(v[e], g[(t + T * e) * 8 + mp.lay_lam()]);
        fft_inv<N>(v, t, r, ex, mp.lay, tw, [&]() {
            if (r < 3 && threadIdx.x == 0) {
                fence_proxy_async();
                issue_g(r + 1);
            }
        });
        demod_accumulate<N>(acc, v, r);
    }
    cd* o = out + line_out(la, L) + (long)t * la.out_es;
#pragma unroll
    for (int a = 0; a < E; ++a) o[(long)(a * T) * la.out_es] = acc[a];
}

template <int N, int MINB>
inline cudaError_t launch_mid_lean(cudaStream_t s, long nlines, const cd* in, cd* out, const cd* G, const cd* TAB,
                                   const LineAddr& la) {
    constexpr int smem = (2 * GeoB<N>::LPC * N + EngTab<N>::TW1N) * (int)sizeof(cd) + 16;
    static unsigned long long optin = 0;
    { cudaError_t e = smem_optin(k_mid_lean<N, MINB>, smem, optin); if (e != cudaSuccess) return e; }
    k_mid_lean<N, MINB><<<(unsigned)(nlines / GeoB<N>::LPC), GeoB<N>::THREADS, smem, s>>>(in, out, G, TAB, la, 0);
    return cudaPeekAtLastError();
}

// ---- middle, fused, persistent CTAs with an asynchronously prefetched input line (mode A) -------------
// Same arithmetic as k_mid_fused (spectrum straight from HBM, requested before the last butterfly stage).
// A CTA walks over line groups with stride gridDim.x; while it transforms one group, the next group's
// (strided, 16-byte-granular) input line is already on its way into the other half of a double buffer by
// cp.async - thread-private, so no barrier is needed - and the engine tables are set up once per CTA.
// smem: [exchange: LPC*N][x buffer 0: LPC*N][x buffer 1: LPC*N][tw1]

template <int N>
__global__ void __launch_bounds__(GeoA<N>::THREADS, GeoA<N>::THREADS <= 128 ? 2 : 1)
k_mid_persist(const cd* in, cd* out, const cd* __restrict__ G, const cd* __restrict__ TAB,
              const LineAddr la, long ngroups) {
    typedef Map<N, false> M;
    constexpr int E = Cfg<N>::E, T = N / E, LPC = M::G::LPC;
    extern __shared__ __align__(128) cd sm[];
    M mp;
    cd* ex = sm;
    cd* xbuf0 = sm + LPC * N;
    cd* xbuf1 = sm + 2 * LPC * N;
    cd* tw1 = sm + 3 * LPC * N;
    load_tw1<N>(tw1, TAB);
    const int t = mp.t;
    const TwState<N> tw = make_tw<N>(t, TAB, tw1);
    auto issue_x = [&](long grp, cd* xb) {
        const cd* p = in + line_in(la, grp * LPC + mp.line) + (long)t * la.in_es;
#pragma unroll
        for (int a = 0; a < E; ++a) cp_async16(&xb[mp.lay.phys(a * T + t)], p + (long)(a * T) * la.in_es);
        cp_async_commit();
    };
    long grp = blockIdx.x;
    if (grp < ngroups) issue_x(grp, xbuf0);
    __syncthreads();   // tw1 visible
    int it = 0;
#pragma unroll 1
    for (; grp < ngroups; grp += gridDim.x, ++it) {
        cd* xs = (it & 1) ? xbuf1 : xbuf0;
        const long nxt = grp + gridDim.x;
        if (nxt < ngroups) {
            issue_x(nxt, (it & 1) ? xbuf0 : xbuf1);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        const long L = grp * LPC + mp.line;
        const cd* g = G + (L * 4) * (long)N + t;
        cd acc[E];
#pragma unroll 1
        for (int r = 0; r < 4; ++r) {
            cd v[E];
#pragma unroll
            for (int a = 0; a < E; ++a) v[a] = xs[mp.lay.phys(a * T + t)];
            cd gv[E];
            const cd* gr = g + (long)r * N;
            fft_fwd<N>(v, t, r, ex, mp.lay, tw, [&]() {
#pragma unroll
                for (int e = 0; e < E; ++e) gv[e] = __ldg(&gr[T * e]);
            });
#pragma unroll
            for (int e = 0; e < E; ++e) v[e] = cmul(v[e], gv[e]);
            fft_inv<N>(v, t, r, ex, mp.lay, tw);
            demod_accumulate<N>(acc, v, r);
        }
        cd* o = out + line_out(la, L) + (long)t * la.out_es;
#pragma unroll
        for (int a = 0; a < E; ++a) o[(long)(a * T) * la.out_es] = acc[a];
    }
}

template <int N>
inline cudaError_t launch_mid_persist(cudaStream_t s, long nlines, const cd* in, cd* out, const cd* G, const cd* TAB,
                                      const LineAddr& la, int ctas) {
    constexpr int smem = (3 * GeoA<N>::LPC * N + EngTab<N>::TW1N) * (int)sizeof(cd);
    static unsigned long long optin = 0;
    { cudaError_t e = smem_optin(k_mid_persist<N>, smem, optin); if (e != cudaSuccess) return e; }
    const long ngroups = nlines / GeoA<N>::LPC;
    const long grid = ngroups < ctas ? ngroups : ctas;
    k_mid_persist<N><<<(unsigned)grid, GeoA<N>::THREADS, smem, s>>>(in, out, G, TAB, la, ngroups);
    return cudaPeekAtLastError();
}

// ---- middle, fused, one sub-transform per CTA, four CTAs per line group as a thread-block cluster ------
// The four sub-transforms r = 0..3 of a padded line are independent until the final sum.  Giving
// each its own CTA removes the two register-hungry pieces of k_mid_fused - the persistent copy of
// the input line and the 16 accumulators carried across r - so three to four CTAs fit per SM instead
// of two (the FP64 and shared-memory phases of different CTAs then overlap much better).  The four
// partial results meet through distributed shared memory: every CTA parks its demodulated line in
// its own smem, cluster.sync(), then CTA c sums quarter c of the line over the four ranks in a fixed
// order (deterministic) and stores it.
// grid = 4 * (lines / LPC), cluster (4,1,1); cluster rank = r.
template <int N, bool MODE_B, int MINB>
__global__ void __launch_bounds__(MODE_B ? GeoB<N>::THREADS : GeoA<N>::THREADS, MINB)
k_mid_cluster(const cd* in, cd* out, const cd* __restrict__ G, const cd* __restrict__ TAB,
              const LineAddr la, long line0) {
    namespace cg = cooperative_groups;
    typedef Map<N, MODE_B> M;
    constexpr int E = Cfg<N>::E, T = N / E, LPC = M::G::LPC;
    constexpr int TH = MODE_B ? GeoB<N>::THREADS : GeoA<N>::THREADS;
    constexpr int UNIT = MODE_B ? 8 * N : N;
    extern __shared__ __align__(128) cd sm[];
    cg::cluster_group cluster = cg::this_cluster();
    const int r = (int)cluster.block_rank();
    M mp;
    cd* ex = sm + sm_group_off(mp);
    cd* tw1 = sm + LPC * N;
    load_tw1<N>(tw1, TAB);
    const long Lcta = line0 + (long)(blockIdx.x >> 2) * LPC;
    const long L = Lcta + mp.line;
    const int t = mp.t;
    const TwState<N> tw = make_tw<N>(t, TAB, tw1);
    cd v[E];
    {
        const cd* p = in + line_in(la, L) + (long)t * la.in_es;
#pragma unroll
        for (int a = 0; a < E; ++a) v[a] = p[(long)(a * T) * la.in_es];
    }
    __syncthreads();   // tw1 visible
    const cd* g = MODE_B ? G + (((Lcta >> 3) + (mp.line >> 3)) * 4 + r) * (long)UNIT + mp.lay_lam()
                         : G + ((Lcta + mp.line) * 4 + r) * (long)UNIT;
    constexpr int gs = MODE_B ? 8 : 1;
    {
        cd gv[E];
        fft_fwd<N>(v, t, r, ex, mp.lay, tw, [&]() {
#pragma unroll
            for (int e = 0; e < E; ++e) gv[e] = __ldg(&g[(long)(t + T * e) * gs]);
        });
#pragma unroll
        for (int e = 0; e < E; ++e) v[e] = cmul(v[e], gv[e]);
    }
    fft_inv<N>(v, t, r, ex, mp.lay, tw);
    if (r != 0) {
#pragma unroll
        for (int a = 1; a < E; ++a) v[a] = cmulc(v[a], c64(r * a * (16 / E)));
    }
    __syncthreads();   // everyone is done reading the exchange buffer: reuse it for the partial line
    // park: element index within the CTA = line*N + j (mode A) / (grp*8N + j*8 + lam) (mode B)
#pragma unroll
    for (int a = 0; a < E; ++a) {
        const int j = a * T + t;
        const int idx = MODE_B ? (sm_group_off(mp) + j * 8 + mp.lay_lam()) : (mp.line * N + j);
        sm[idx] = v[a];
    }
    cluster.sync();
    // quarter r of the CTA's LPC*N points: sum the four ranks in the order 0,1,2,3
    const cd* part[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) part[q] = cluster.map_shared_rank(sm, q);
    constexpr int QUART = LPC * N / 4;
    for (int i = threadIdx.x; i < QUART; i += TH) {
        const int idx = r * QUART + i;
        cd s0 = part[0][idx];
        const cd s1 = part[1][idx], s2 = part[2][idx], s3 = part[3][idx];
        s0 = cadd(cadd(cadd(s0, s1), s2), s3);
        int line, j;
        if (MODE_B) { const int w = idx % (8 * N); line = (idx / (8 * N)) * 8 + (w & 7); j = w >> 3; }
        else { line = idx / N; j = idx % N; }
        out[line_out(la, Lcta + line) + (long)j * la.out_es] = s0;
    }
    cluster.sync();    // keep every CTA's shared memory alive until its peers have read it
}

template <int N, bool B, int MINB>
inline cudaError_t launch_mid_cluster(cudaStream_t s, long nlines, const cd* in, cd* out, const cd* G, const cd* TAB,
                                      const LineAddr& la) {
    constexpr int smem = Smem<N, B>::fwd_bytes;
    constexpr int LPC = Smem<N, B>::LPC, TH = B ? GeoB<N>::THREADS : GeoA<N>::THREADS;
    static unsigned long long optin = 0;
    { cudaError_t e = smem_optin(k_mid_cluster<N, B, MINB>, smem, optin); if (e != cudaSuccess) return e; }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(4 * (nlines / LPC)));
    cfg.blockDim = dim3(TH);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 4; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, k_mid_cluster<N, B, MINB>, in, out, G, TAB, la, (long)0);
}

// ---- middle, fused, two sub-transforms in flight (mode A, spectrum straight from HBM) ----------------
// Same arithmetic as k_mid_fused; r = 0,1 then r = 2,3 run pairwise through the stages (fft_*_dual).
// smem: [exchange A: LPC*N][x copy: LPC*N][exchange B: LPC*N][accumulator tail: ASM*TH][tw1]
template <int N, int ASM>
__global__ void __launch_bounds__(GeoA<N>::THREADS, 2)
k_mid_fused_dual(const cd* in, cd* out, const cd* __restrict__ G, const cd* __restrict__ TAB,
                 const LineAddr la, long line0) {
    typedef Map<N, false> M;
    constexpr int E = Cfg<N>::E, T = N / E, LPC = M::G::LPC, TH = GeoA<N>::THREADS, AR = E - ASM;
    extern __shared__ __align__(128) cd sm[];
    M mp;
    cd* xs = sm + LPC * N;
    cd* accs = sm + 3 * LPC * N + threadIdx.x;
    cd* tw1 = sm + 3 * LPC * N + ASM * TH;
    const long L = line0 + (long)blockIdx.x * LPC + mp.line;
    const int t = mp.t;
    LayA<N> layA = mp.lay, layB = mp.lay;
    layB.base += 2 * LPC * N;
    load_tw1<N>(tw1, TAB);
    const TwState<N> tw = make_tw<N>(t, TAB, tw1);
    {
        const cd* p = in + line_in(la, L) + (long)t * la.in_es;
#pragma unroll
        for (int a = 0; a < E; ++a) xs[mp.lay.phys(a * T + t)] = p[(long)(a * T) * la.in_es];
    }
    __syncthreads();
    cd acc[AR > 0 ? AR : 1];
    const cd* g = G + (L * 4) * (long)N + t;
#pragma unroll 1
    for (int rr = 0; rr < 4; rr += 2) {
        cd vA[E], vB[E];
#pragma unroll
        for (int a = 0; a < E; ++a) { vA[a] = xs[mp.lay.phys(a * T + t)]; vB[a] = vA[a]; }
        fft_fwd_dual<N>(vA, vB, t, rr, rr + 1, sm, layA, layB, tw);
        const cd* gA = g + (long)rr * N;
#pragma unroll
        for (int e = 0; e < E; ++e) vA[e] = cmul(vA[e], __ldg(&gA[T * e]));
#pragma unroll
        for (int e = 0; e < E; ++e) vB[e] = cmul(vB[e], __ldg(&gA[N + T * e]));
        fft_inv_dual<N>(vA, vB, t, rr, rr + 1, sm, layA, layB, tw);
        // demodulate + accumulate: rr == 0 initialises with r = 0 (no constants), then r = 1; later r = 2, 3
        if (rr == 0) {
#pragma unroll
            for (int a = 0; a < AR; ++a) acc[a] = (a == 0) ? cadd(vA[0], vB[0]) : cfmac(vB[a], c64(a * (16 / E)), vA[a]);
#pragma unroll
            for (int a = AR; a < E; ++a) accs[(a - AR) * TH] = cfmac(vB[a], c64(a * (16 / E)), vA[a]);
        } else {
#pragma unroll
            for (int a = 0; a < AR; ++a) {
                cd z = (a == 0) ? cadd(acc[0], vA[0]) : cfmac(vA[a], c64(2 * a * (16 / E)), acc[a]);
                acc[a] = (a == 0) ? cadd(z, vB[0]) : cfmac(vB[a], c64(3 * a * (16 / E)), z);
            }
#pragma unroll
            for (int a = AR; a < E; ++a) {
                cd z = cfmac(vA[a], c64(2 * a * (16 / E)), accs[(a - AR) * TH]);
                accs[(a - AR) * TH] = cfmac(vB[a], c64(3 * a * (16 / E)), z);
            }
        }
    }
    cd* o = out + line_out(la, L) + (long)t * la.out_es;
#pragma unroll
    for (int a = 0; a < AR; ++a) o[(long)(a * T) * la.out_es] = acc[a];
#pragma unroll
    for (int a = AR; a < E; ++a) o[(long)(a * T) * la.out_es] = accs[(a - AR) * TH];
}

}  // namespace lsk
