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
