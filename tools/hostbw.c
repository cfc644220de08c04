// Host memory write bandwidth probe (used to size the host-side expansion of sparse results):
// T threads each fill their slice of a big buffer, plain memset vs non-temporal stores.
#define _GNU_SOURCE
#include <immintrin.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>
static char *buf; static size_t per; static int mode;
static double now(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }
static void *work(void *a) {
    size_t i = (size_t)a; char *p = buf + i * per;
    if (mode == 0) memset(p, 0, per);
    else { __m128i z = _mm_setzero_si128(); for (size_t k = 0; k < per; k += 16) _mm_stream_si128((__m128i *)(p + k), z); _mm_sfence(); }
    return 0;
}
int main(int argc, char **argv) {
    int T = argc > 1 ? atoi(argv[1]) : (int)sysconf(_SC_NPROCESSORS_ONLN);
    size_t gb = argc > 2 ? atoi(argv[2]) : 8;
    size_t total = gb << 30; per = (total / T) & ~(size_t)4095;
    buf = aligned_alloc(4096, per * T);
    printf("nproc %ld threads %d buffer %zu GB\n", sysconf(_SC_NPROCESSORS_ONLN), T, gb);
    for (int rep = 0; rep < 3; rep++)
        for (mode = 0; mode < 2; mode++) {
            pthread_t th[512]; double t0 = now();
            for (int i = 0; i < T; i++) pthread_create(&th[i], 0, work, (void *)(size_t)i);
            for (int i = 0; i < T; i++) pthread_join(th[i], 0);
            double dt = now() - t0;
            printf("rep %d %s: %.1f GB/s\n", rep, mode ? "stream" : "memset", per * T / dt / 1e9);
        }
    return 0;
}
