// Library-level entry points of libls_cuda.so: errors, device selection, raw buffers,
// handle services, and the host-side table builders shared by the operator files.
#include "ls_common.cuh"
#include <cstring>

namespace ls {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

std::vector<int> slot_freq(int N) {
    int E, S, R[3];
    switch (N) {
        case 64:   E = 8;  S = 2; R[0] = 8;  R[1] = 8;  R[2] = 1;  break;
        case 128:  E = 16; S = 2; R[0] = 16; R[1] = 8;  R[2] = 1;  break;
        case 256:  E = 16; S = 2; R[0] = 16; R[1] = 16; R[2] = 1;  break;
        case 512: {
            // 16 x 16 x 2 with the closing radix-2 across lane partners (fft_engine.cuh ShflLast): thread t, register e holds
            // the frequency d0 + 16 d1 + 256 d2 with d0 = t / 2 (stage-0 output), d1 = e (stage-1 output), d2 = t & 1
            std::vector<int> f(512);
            for (int t = 0; t < 32; ++t)
                for (int e = 0; e < 16; ++e) f[t + 32 * e] = (t / 2) + 16 * e + 256 * (t & 1);
            return f;
        }
        case 1024: E = 16; S = 3; R[0] = 16; R[1] = 8;  R[2] = 8;  break;
        case 2048: E = 16; S = 3; R[0] = 16; R[1] = 16; R[2] = 8;  break;
        case 4096: E = 16; S = 3; R[0] = 16; R[1] = 16; R[2] = 16; break;
        default: return std::vector<int>();
    }
    const int T = N / E, RL = R[S - 1];
    std::vector<int> f(N);
    for (int t = 0; t < T; ++t)
        for (int e = 0; e < E; ++e) {
            int u = e / RL, d = e % RL;
            int D = t + T * u;
            int digs[3];
            digs[S - 1] = d;
            for (int i = S - 2; i >= 0; --i) { digs[i] = D % R[i]; D /= R[i]; }
            int k = 0, mult = 1;
            for (int i = 0; i < S; ++i) { k += digs[i] * mult; mult *= R[i]; }
            f[t + T * e] = k;
        }
    return f;
}

static inline cd unit_root(long num, long den) {
    // exp(-2 pi i num/den), octant-reduced in long double
    num %= den;
    long double x = (long double)num / (long double)den;   // in [0,1)
    long double ang = 2.0L * 3.14159265358979323846264338327950288L * x;
    return make_double2((double)cosl(ang), (double)(-sinl(ang)));
}

std::vector<cd> twiddle_table(long N, long count) {
    std::vector<cd> w(count);
    for (long k = 0; k < count; ++k) w[k] = unit_root(k, N);
    // exact values on the axes
    for (long k = 0; k < count; ++k) {
        if ((4 * k) % N == 0) {
            int q = (int)((4 * k) / N) & 3;
            const double re[4] = {1, 0, -1, 0}, im[4] = {0, -1, 0, 1};
            w[k] = make_double2(re[q], im[q]);
        }
    }
    return w;
}

std::vector<cd> modulation_table(long N) {
    std::vector<cd> w(3 * N);
    for (int r = 1; r <= 3; ++r)
        for (long j = 0; j < N; ++j) w[(r - 1) * N + j] = unit_root((long)r * j, 4 * N);
    return w;
}

std::vector<cd> engine_table(int N) {
    int E, S, R0, R1;
    switch (N) {
        case 64:   E = 8;  S = 2; R0 = 8;  R1 = 8;  break;
        case 128:  E = 16; S = 2; R0 = 16; R1 = 8;  break;
        case 256:  E = 16; S = 2; R0 = 16; R1 = 16; break;
        case 512:  E = 16; S = 3; R0 = 16; R1 = 16; break;
        case 1024: E = 16; S = 3; R0 = 16; R1 = 8;  break;
        case 2048: E = 16; S = 3; R0 = 16; R1 = 16; break;
        case 4096: E = 16; S = 3; R0 = 16; R1 = 16; break;
        default: return std::vector<cd>();
    }
    const int T = N / E;
    std::vector<cd> tab;
    for (int t = 0; t < T; ++t) tab.push_back(unit_root(t, 4L * N));
    for (int t = 0; t < T; ++t) tab.push_back(unit_root(t, N));
    if (S == 3) {
        const int Mprev = N / R0, M = Mprev / R1;
        for (int d = 0; d < R1; ++d)
            for (int b = 0; b < M; ++b) tab.push_back(unit_root((long)b * d, Mprev));
    }
    return tab;
}

int upload(void** dptr, const void* host, size_t bytes, cudaStream_t s) {
    LS_CUDA_TRY(cudaMalloc(dptr, bytes));
    LS_CUDA_TRY(cudaMemcpyAsync(*dptr, host, bytes, cudaMemcpyHostToDevice, s));
    LS_CUDA_TRY(cudaStreamSynchronize(s));
    return LS_OK;
}

}  // namespace ls

using namespace ls;

extern "C" {

int ls_version(void) { return 100; }

const char* ls_last_error(void) { return ls::g_err; }

int ls_device_count(int* count) {
    LS_REQUIRE(count, LS_ERR_INVALID, "ls_device_count: null pointer");
    LS_CUDA_TRY(cudaGetDeviceCount(count));
    return LS_OK;
}

int ls_set_device(int device) {
    LS_CUDA_TRY(cudaSetDevice(device));
    return LS_OK;
}

int ls_dev_alloc(void** dptr, size_t bytes) {
    LS_REQUIRE(dptr, LS_ERR_INVALID, "ls_dev_alloc: null pointer");
    cudaError_t e = cudaMalloc(dptr, bytes);
    if (e == cudaErrorMemoryAllocation) {
        cudaGetLastError();
        set_error("ls_dev_alloc: out of device memory (%zu bytes)", bytes);
        return LS_ERR_NOMEM;
    }
    LS_CUDA_TRY(e);
    return LS_OK;
}

int ls_dev_free(void* dptr) {
    LS_CUDA_TRY(cudaFree(dptr));
    return LS_OK;
}

int ls_memcpy_h2d(void* dst, const void* src, size_t bytes) {
    LS_CUDA_TRY(cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice));
    return LS_OK;
}

int ls_memcpy_d2h(void* dst, const void* src, size_t bytes) {
    LS_CUDA_TRY(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost));
    return LS_OK;
}

int ls_host_alloc_pinned(void** hptr, size_t bytes) {
    LS_REQUIRE(hptr, LS_ERR_INVALID, "ls_host_alloc_pinned: null pointer");
    LS_CUDA_TRY(cudaMallocHost(hptr, bytes));
    return LS_OK;
}

int ls_host_free_pinned(void* hptr) {
    LS_CUDA_TRY(cudaFreeHost(hptr));
    return LS_OK;
}

int ls_destroy(ls_handle h) {
    if (!h) return LS_OK;
    HandleBase* b = reinterpret_cast<HandleBase*>(h);
    cudaSetDevice(b->device);
    delete b;
    return LS_OK;
}

int ls_sync(ls_handle h) {
    LS_REQUIRE(h, LS_ERR_INVALID, "ls_sync: null handle");
    HandleBase* b = reinterpret_cast<HandleBase*>(h);
    LS_CUDA_TRY(cudaStreamSynchronize(b->stream));
    return LS_OK;
}

int ls_timer_start(ls_handle h) {
    LS_REQUIRE(h, LS_ERR_INVALID, "ls_timer_start: null handle");
    HandleBase* b = reinterpret_cast<HandleBase*>(h);
    LS_CUDA_TRY(cudaEventRecord(b->ev0, b->stream));
    return LS_OK;
}

int ls_timer_stop(ls_handle h, float* ms) {
    LS_REQUIRE(h && ms, LS_ERR_INVALID, "ls_timer_stop: null argument");
    HandleBase* b = reinterpret_cast<HandleBase*>(h);
    LS_CUDA_TRY(cudaEventRecord(b->ev1, b->stream));
    LS_CUDA_TRY(cudaEventSynchronize(b->ev1));
    LS_CUDA_TRY(cudaEventElapsedTime(ms, b->ev0, b->ev1));
    return LS_OK;
}

int ls_profile_enable(ls_handle h, int on) {
    LS_REQUIRE(h, LS_ERR_INVALID, "ls_profile_enable: null handle");
    HandleBase* b = reinterpret_cast<HandleBase*>(h);
    LS_CUDA_TRY(cudaStreamSynchronize(b->stream));
    for (auto& pe : b->phase_events) { cudaEventDestroy(pe.a); cudaEventDestroy(pe.b); }
    b->phase_events.clear();
    b->profiling = on != 0;
    return LS_OK;
}

int ls_profile_read(ls_handle h, double* ms, int64_t* counts, int nphase) {
    LS_REQUIRE(h && ms && counts && nphase > 0, LS_ERR_INVALID, "ls_profile_read: bad argument");
    HandleBase* b = reinterpret_cast<HandleBase*>(h);
    LS_CUDA_TRY(cudaStreamSynchronize(b->stream));
    for (int i = 0; i < nphase; ++i) { ms[i] = 0.0; counts[i] = 0; }
    for (auto& pe : b->phase_events) {
        float t = 0.f;
        LS_CUDA_TRY(cudaEventElapsedTime(&t, pe.a, pe.b));
        if (pe.phase >= 0 && pe.phase < nphase) { ms[pe.phase] += t; counts[pe.phase]++; }
    }
    return LS_OK;
}

int ls_op_size(ls_handle h, int64_t* N) {
    LS_REQUIRE(h && N, LS_ERR_INVALID, "ls_op_size: null argument");
    int64_t v = reinterpret_cast<HandleBase*>(h)->op_size();
    LS_REQUIRE(v >= 0, LS_ERR_INVALID, "ls_op_size: not an operator handle");
    *N = v;
    return LS_OK;
}

int ls_launch_count(ls_handle h, int64_t* count) {
    LS_REQUIRE(h && count, LS_ERR_INVALID, "ls_launch_count: null argument");
    *count = reinterpret_cast<HandleBase*>(h)->launches;
    return LS_OK;
}

}  // extern "C"
