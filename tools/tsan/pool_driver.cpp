#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <deque>
#include <functional>
#include <memory>
#include <mutex>
#include <random>
#include <string>
#include <thread>
#include <vector>
#include "scasa_host.h"
namespace scasa { void set_global_error(const std::string &) {} }
#include "pool_snippet.h"
int main() {
    std::mt19937_64 rng(3);
    const int64_t n_cols = 1000;
    ExpandPool pool(6);
    for (int it = 0; it < 200; it++) {
        // two producers (like a context and its twin) post batches concurrently and wait on them
        auto producer = [&](int seed) {
            std::mt19937_64 r(seed);
            for (int b = 0; b < 5; b++) {
                const int64_t n_rows = 1 + r() % 40;
                std::vector<int64_t> start((size_t)n_rows);
                std::vector<int32_t> nnz((size_t)n_rows), cols;
                std::vector<double> vals;
                for (int64_t i = 0; i < n_rows; i++) {
                    start[(size_t)i] = (int64_t)cols.size();
                    int k = 0;
                    for (int64_t j = 0; j < n_cols; j++) if (r() % 13 == 0) { cols.push_back((int32_t)j); vals.push_back((double)(r() % 1000) + 1); k++; }
                    nnz[(size_t)i] = k;
                }
                double *raw = (double *)aligned_alloc(64, (size_t)(n_rows * n_cols + 8) * sizeof(double));
                memset(raw, 0xff, (size_t)(n_rows * n_cols) * sizeof(double));
                auto j1 = pool.post(raw, n_cols, n_rows, start.data(), nnz.data(), cols.data(), vals.data());
                pool.wait(j1);
                // check
                for (int64_t i = 0; i < n_rows; i++) {
                    int64_t p = start[(size_t)i];
                    for (int64_t j = 0; j < n_cols; j++) {
                        double want = 0;
                        if (p < start[(size_t)i] + nnz[(size_t)i] && cols[(size_t)p] == j) want = vals[(size_t)p++];
                        if (raw[i * n_cols + j] != want) { printf("MISMATCH\n"); exit(1); }
                    }
                }
                free(raw);
                // a row-range job (the sparse fetch): every row visited exactly once, also when it overlaps the other producer's jobs
                std::vector<int> seen((size_t)n_rows, 0);
                auto j2 = pool.post_fn(n_rows, [&seen](int64_t a, int64_t e) { for (int64_t i = a; i < e; i++) seen[(size_t)i]++; });
                pool.wait(j2);
                for (int v : seen) if (v != 1) { printf("MISMATCH (row job)\n"); exit(1); }
                pool.wait(pool.post_fn(0, [](int64_t, int64_t) {}));          // an empty job is done at once
            }
        };
        std::thread t1(producer, it * 2 + 1), t2(producer, it * 2 + 2);
        t1.join(); t2.join();
    }
    printf("pool ok\n");
}
