// Device-side unit-test hooks for the line-FFT engine (exported, but not part of the
// reference-facing surface).
#include "ls_common.cuh"
#include "line_kernels.cuh"

using namespace ls;
using namespace lsk;

namespace {

template <int N>
__global__ void __launch_bounds__(GeoA<N>::THREADS)
k_test_fft(const cd* __restrict__ in, cd* __restrict__ out, const cd* __restrict__ TAB, int roundtrip) {
    constexpr int E = Cfg<N>::E, T = N / E;
    extern __shared__ __align__(128) cd sm[];
    Map<N, false> mp;
    cd* tw1 = sm + GeoA<N>::LPC * N;
    load_tw1<N>(tw1, TAB);
    const long L = (long)blockIdx.x * GeoA<N>::LPC + mp.line;
    const int t = mp.t;
    const TwState<N> tw = make_tw<N>(t, TAB, tw1);
    __syncthreads();
    cd v[E];
#pragma unroll
    for (int a = 0; a < E; ++a) v[a] = in[L * N + a * T + t];
    fft_fwd<N>(v, t, 0, sm, mp.lay, tw);
    if (roundtrip) {
        fft_inv<N>(v, t, 0, sm, mp.lay, tw);
#pragma unroll
        for (int a = 0; a < E; ++a) out[L * N + a * T + t] = cscale(v[a], 1.0 / N);
    } else {
#pragma unroll
        for (int e = 0; e < E; ++e) out[L * N + t + T * e] = v[e];   // slot order
    }
}

template <int N> int run_test(long nlines, const cd* d_in, cd* d_out, const cd* d_W, int roundtrip) {
    constexpr int smem = Smem<N, false>::fwd_bytes;
    LS_CUDA_TRY(cudaFuncSetAttribute(k_test_fft<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    k_test_fft<N><<<(unsigned)(nlines / GeoA<N>::LPC), GeoA<N>::THREADS, smem>>>(d_in, d_out, d_W, roundtrip);
    LS_CUDA_TRY(cudaGetLastError());
    LS_CUDA_TRY(cudaDeviceSynchronize());
    return LS_OK;
}

}  // namespace

extern "C" int ls_test_fft_lines(int64_t N, int64_t nlines, const ls_cdouble* in_host, ls_cdouble* out_host,
                                 int inverse_roundtrip) {
    LS_REQUIRE(in_host && out_host, LS_ERR_INVALID, "ls_test_fft_lines: null pointer");
    LS_REQUIRE(fft_size_supported(N), LS_ERR_UNSUPPORTED, "ls_test_fft_lines: unsupported N=%ld", (long)N);
    LS_REQUIRE(nlines > 0 && nlines % 16 == 0, LS_ERR_INVALID, "ls_test_fft_lines: nlines must be a positive multiple of 16");
    const size_t bytes = (size_t)N * nlines * sizeof(cd);
    cd *d_in = nullptr, *d_out = nullptr, *d_W = nullptr;
    auto W = engine_table((int)N);
    LS_CUDA_TRY(cudaMalloc(&d_in, bytes));
    LS_CUDA_TRY(cudaMalloc(&d_out, bytes));
    LS_CUDA_TRY(cudaMalloc(&d_W, W.size() * sizeof(cd)));
    LS_CUDA_TRY(cudaMemcpy(d_in, in_host, bytes, cudaMemcpyHostToDevice));
    LS_CUDA_TRY(cudaMemcpy(d_W, W.data(), W.size() * sizeof(cd), cudaMemcpyHostToDevice));
    int rc = LS_OK;
    switch (N) {
        case 64:   rc = run_test<64>(nlines, d_in, d_out, d_W, inverse_roundtrip); break;
        case 128:  rc = run_test<128>(nlines, d_in, d_out, d_W, inverse_roundtrip); break;
        case 256:  rc = run_test<256>(nlines, d_in, d_out, d_W, inverse_roundtrip); break;
        case 512:  rc = run_test<512>(nlines, d_in, d_out, d_W, inverse_roundtrip); break;
        case 1024: rc = run_test<1024>(nlines, d_in, d_out, d_W, inverse_roundtrip); break;
        case 2048: rc = run_test<2048>(nlines, d_in, d_out, d_W, inverse_roundtrip); break;
        case 4096: rc = run_test<4096>(nlines, d_in, d_out, d_W, inverse_roundtrip); break;
    }
    if (rc == LS_OK) {
        std::vector<cd> tmp((size_t)N * nlines);
        cudaError_t e = cudaMemcpy(tmp.data(), d_out, bytes, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { set_error("copy back failed: %s", cudaGetErrorString(e)); rc = LS_ERR_CUDA; }
        cd* o = reinterpret_cast<cd*>(out_host);
        if (inverse_roundtrip) {
            for (size_t i = 0; i < tmp.size(); ++i) o[i] = tmp[i];
        } else {
            auto f = slot_freq((int)N);   // slot -> natural frequency
            for (long L = 0; L < nlines; ++L)
                for (long s = 0; s < N; ++s) o[L * N + f[s]] = tmp[L * N + s];
        }
    }
    cudaFree(d_in); cudaFree(d_out); cudaFree(d_W);
    return rc;
}

// ---- measured device peaks for the roofline denominators of bench.py ----------------------------------------------
// FP64: 8 independent dependent-chains per thread of DFMA (or DADD) instructions, enough warps to fill every SM.
namespace {
template <bool FMA>
__global__ void __launch_bounds__(256)
k_fp64_peak(double* out, int iters, double b, double c) {
    double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
#pragma unroll 4
    for (int i = 0; i < iters; ++i) {
        if (FMA) {
            a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
            a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
        } else {
            a0 = __dadd_rn(a0, c); a1 = __dadd_rn(a1, c); a2 = __dadd_rn(a2, c); a3 = __dadd_rn(a3, c);
            a4 = __dadd_rn(a4, c); a5 = __dadd_rn(a5, c); a6 = __dadd_rn(a6, c); a7 = __dadd_rn(a7, c);
        }
    }
    const double s = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
    if (s == 123.456) out[0] = s;      // keeps the chains alive
}
__global__ void __launch_bounds__(256)
k_copy_peak(const double2* __restrict__ src, double2* __restrict__ dst, long n) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) dst[i] = src[i];
}
}  // namespace

// dfma_tflops: 2 flops per DFMA lane;  dadd_tinst: 1e12 FP64 lane-instructions per second (DADD; DMUL issues alike);
// copy_gbs: read + write bytes of a 1 GiB device-to-device copy kernel (the HBM roofline this library's kernels see)
extern "C" int ls_test_device_peaks(double* dfma_tflops, double* dadd_tinst, double* copy_gbs) {
    cudaEvent_t e0, e1;
    LS_CUDA_TRY(cudaEventCreate(&e0));
    LS_CUDA_TRY(cudaEventCreate(&e1));
    double* d_out = nullptr;
    LS_CUDA_TRY(cudaMalloc(&d_out, 64));
    int sms = 148;
    { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
    const int blocks = sms * 8, iters = 1 << 15;
    float ms = 0.f, best[2] = {1e30f, 1e30f};
    for (int which = 0; which < 2; ++which)
        for (int rep = 0; rep < 4; ++rep) {
            LS_CUDA_TRY(cudaEventRecord(e0));
            if (which == 0) k_fp64_peak<true><<<blocks, 256>>>(d_out, iters, 1.0000001, 1e-9);
            else k_fp64_peak<false><<<blocks, 256>>>(d_out, iters, 1.0000001, 1e-9);
            LS_CUDA_TRY(cudaEventRecord(e1));
            LS_CUDA_TRY(cudaEventSynchronize(e1));
            LS_CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
            if (rep > 0 && ms < best[which]) best[which] = ms;
        }
    const double lane_ops = (double)blocks * 256.0 * iters * 8.0;
    if (dfma_tflops) *dfma_tflops = 2.0 * lane_ops / (best[0] * 1e-3) / 1e12;
    if (dadd_tinst) *dadd_tinst = lane_ops / (best[1] * 1e-3) / 1e12;
    cudaFree(d_out);
    if (copy_gbs) {
        const long n = 1L << 26;        // 1 GiB of double2
        double2 *a = nullptr, *b = nullptr;
        LS_CUDA_TRY(cudaMalloc(&a, n * sizeof(double2)));
        LS_CUDA_TRY(cudaMalloc(&b, n * sizeof(double2)));
        LS_CUDA_TRY(cudaMemset(a, 0, n * sizeof(double2)));
        float bestc = 1e30f;
        for (int rep = 0; rep < 6; ++rep) {
            LS_CUDA_TRY(cudaEventRecord(e0));
            k_copy_peak<<<sms * 16, 256>>>(a, b, n);
            LS_CUDA_TRY(cudaEventRecord(e1));
            LS_CUDA_TRY(cudaEventSynchronize(e1));
            LS_CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
            if (rep > 0 && ms < bestc) bestc = ms;
        }
        *copy_gbs = 2.0 * n * sizeof(double2) / (bestc * 1e-3) / 1e9;
        cudaFree(a); cudaFree(b);
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return LS_OK;
}
