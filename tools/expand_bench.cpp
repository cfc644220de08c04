// Host-side probe: how fast can (column, value) rows be expanded into a dense matrix?
// Variants of the inner loop of fetch_dense's expansion, T threads.  g++ -O2 -pthread.
#include <immintrin.h>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <thread>
#include <vector>
using clk = std::chrono::steady_clock;

static inline void stream_zero(double *p, int64_t n)
{
    while (n > 0 && ((uintptr_t)p & 15)) { _mm_stream_si64((long long *)p, 0); p++; n--; }
    const __m128d z = _mm_setzero_pd();
    for (; n >= 2; n -= 2, p += 2) _mm_stream_pd(p, z);
    if (n) _mm_stream_si64((long long *)p, 0);
}
static void expand_a(double *o, int64_t n_cols, const int32_t *c, const double *v, int nnz)
{
    int64_t at = 0;
    for (int k = 0; k < nnz; k++) {
        const int64_t j = c[k];
        stream_zero(o + at, j - at);
        long long bits; memcpy(&bits, v + k, 8);
        _mm_stream_si64((long long *)(o + j), bits);
        at = j + 1;
    }
    stream_zero(o + at, n_cols - at);
}
// 32-byte blocks, AVX2
__attribute__((target("avx2"))) static void expand_b(double *o, int64_t n_cols, const int32_t *c, const double *v, int nnz)
{
    int64_t at = 0; int k = 0;
    while (at < n_cols && ((uintptr_t)(o + at) & 31)) {
        double x = 0; if (k < nnz && c[k] == at) x = v[k++];
        long long bits; memcpy(&bits, &x, 8); _mm_stream_si64((long long *)(o + at), bits); at++;
    }
    const __m256d z = _mm256_setzero_pd();
    int64_t next = k < nnz ? c[k] : INT64_MAX;
    for (; at + 4 <= n_cols; at += 4) {
        if (next >= at + 4) { _mm256_stream_pd(o + at, z); continue; }
        alignas(32) double t[4] = {0, 0, 0, 0};
        do { t[next - at] = v[k]; k++; next = k < nnz ? c[k] : INT64_MAX; } while (next < at + 4);
        _mm256_stream_pd(o + at, _mm256_load_pd(t));
    }
    for (; at < n_cols; at++) {
        double x = 0; if (k < nnz && c[k] == at) x = v[k++];
        long long bits; memcpy(&bits, &x, 8); _mm_stream_si64((long long *)(o + at), bits);
    }
}
// 64-byte blocks, AVX-512
__attribute__((target("avx512f"))) static void expand_c(double *o, int64_t n_cols, const int32_t *c, const double *v, int nnz)
{
    int64_t at = 0; int k = 0;
    while (at < n_cols && ((uintptr_t)(o + at) & 63)) {
        double x = 0; if (k < nnz && c[k] == at) x = v[k++];
        long long bits; memcpy(&bits, &x, 8); _mm_stream_si64((long long *)(o + at), bits); at++;
    }
    const __m512d z = _mm512_setzero_pd();
    int64_t next = k < nnz ? c[k] : INT64_MAX;
    for (; at + 8 <= n_cols; at += 8) {
        if (next >= at + 8) { _mm512_stream_pd(o + at, z); continue; }
        alignas(64) double t[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        do { t[next - at] = v[k]; k++; next = k < nnz ? c[k] : INT64_MAX; } while (next < at + 8);
        _mm512_stream_pd(o + at, _mm512_load_pd(t));
    }
    for (; at < n_cols; at++) {
        double x = 0; if (k < nnz && c[k] == at) x = v[k++];
        long long bits; memcpy(&bits, &x, 8); _mm_stream_si64((long long *)(o + at), bits);
    }
}
// row built in a cache-resident buffer, then streamed out
__attribute__((target("avx2"))) static void expand_d(double *o, int64_t n_cols, const int32_t *c, const double *v, int nnz, double *tmp)
{
    memset(tmp, 0, n_cols * 8);
    for (int k = 0; k < nnz; k++) tmp[c[k]] = v[k];
    int64_t at = 0;
    while (at < n_cols && ((uintptr_t)(o + at) & 31)) { long long b; memcpy(&b, tmp + at, 8); _mm_stream_si64((long long *)(o + at), b); at++; }
    for (; at + 4 <= n_cols; at += 4) _mm256_stream_pd(o + at, _mm256_loadu_pd(tmp + at));
    for (; at < n_cols; at++) { long long b; memcpy(&b, tmp + at, 8); _mm_stream_si64((long long *)(o + at), b); }
}

int main(int argc, char **argv)
{
    const int T = argc > 1 ? atoi(argv[1]) : 16;
    const int64_t n_cols = 33428, n_rows = argc > 2 ? atoi(argv[2]) : 24000;
    const int nnz = argc > 3 ? atoi(argv[3]) : 2700;
    std::vector<int32_t> cols((size_t)n_rows * nnz); std::vector<double> vals((size_t)n_rows * nnz);
    std::mt19937 rng(1);
    std::vector<int32_t> perm(n_cols);
    for (int64_t r = 0; r < n_rows; r++) {
        // pick nnz distinct sorted columns: stride + jitter
        double step = (double)n_cols / nnz;
        for (int k = 0; k < nnz; k++) { int32_t j = (int32_t)(k * step + (rng() % (int)step)); cols[r * nnz + k] = j; vals[r * nnz + k] = 1.0 + k; }
    }
    double *out = (double *)aligned_alloc(4096, (size_t)n_rows * n_cols * 8);
    memset(out, 1, (size_t)n_rows * n_cols * 8);
    const bool avx2 = __builtin_cpu_supports("avx2"), avx512 = __builtin_cpu_supports("avx512f");
    printf("threads %d rows %ld nnz/row %d avx2 %d avx512 %d\n", T, (long)n_rows, nnz, avx2, avx512);
    for (int variant = 0; variant < 4; variant++) {
        if ((variant == 1 || variant == 3) && !avx2) continue;
        if (variant == 2 && !avx512) continue;
        for (int rep = 0; rep < 2; rep++) {
            auto t0 = clk::now();
            std::vector<std::thread> th;
            for (int t = 0; t < T; t++)
                th.emplace_back([&, t] {
                    std::vector<double> tmp(n_cols + 8);
                    for (int64_t r = n_rows * t / T; r < n_rows * (t + 1) / T; r++) {
                        double *o = out + r * n_cols; const int32_t *c = &cols[r * nnz]; const double *v = &vals[r * nnz];
                        if (variant == 0) expand_a(o, n_cols, c, v, nnz);
                        else if (variant == 1) expand_b(o, n_cols, c, v, nnz);
                        else if (variant == 2) expand_c(o, n_cols, c, v, nnz);
                        else expand_d(o, n_cols, c, v, nnz, tmp.data());
                    }
                    _mm_sfence();
                });
            for (auto &x : th) x.join();
            double dt = std::chrono::duration<double>(clk::now() - t0).count();
            // verify a row
            int64_t r = n_rows / 2; double s = 0; for (int64_t j = 0; j < n_cols; j++) s += out[r * n_cols + j];
            double want = 0; for (int k = 0; k < nnz; k++) want += vals[r * nnz + k];
            printf("variant %c rep %d: %.1f GB/s  (row check %s)\n", "abcd"[variant], rep, n_rows * n_cols * 8 / dt / 1e9, s == want ? "ok" : "BAD");
        }
    }
    return 0;
}
