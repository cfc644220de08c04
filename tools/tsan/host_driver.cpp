#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <random>
#include "scasa_cuda.h"
namespace scasa { void set_global_error(const std::string &m) { static thread_local std::string e; e = m; } }
extern "C" const char *scasa_last_error(const scasa_ctx *) { return ""; }
int main() {
    std::mt19937_64 rng(1);
    // ---- writer
    const int64_t n_cells = 5, n_cols = 64 * 32;
    std::vector<double> mat((size_t)(n_cells * n_cols));
    for (auto &v : mat) v = (rng() % 10 < 3) ? (double)(rng() % 100000) / 7.0 : 0.0;
    std::string rn, cn;
    for (int64_t j = 0; j < n_cols; j++) rn += "T" + std::to_string(j) + "\n";
    for (int64_t c = 0; c < n_cells; c++) cn += "C" + std::to_string(c) + "\n";
    std::string ref;
    for (int it = 0; it < 60; it++) {
        int rc = scasa_write_table("/tmp/scasa_tsan/t.txt", mat.data(), n_cells, n_cols, nullptr, rn.c_str(), cn.c_str(), 1 + it % 9);
        if (rc) { printf("write rc %d\n", rc); return 1; }
        FILE *f = fopen("/tmp/scasa_tsan/t.txt", "rb"); std::string s; char buf[65536]; size_t k;
        while ((k = fread(buf, 1, sizeof buf, f)) > 0) s.append(buf, k);
        fclose(f);
        if (it == 0) ref = s; else if (s != ref) { printf("MISMATCH %d\n", it); return 1; }
    }
    // ---- bfh parser
    {
        const int n_tx = 50, n_cb = 40, n_eq = 3000;
        FILE *f = fopen("/tmp/scasa_tsan/bfh.txt", "wb");
        fprintf(f, "%d\n%d\n%d\n", n_tx, n_cb, n_eq);
        for (int i = 0; i < n_tx; i++) fprintf(f, "TX%d\n", i);
        const char *B = "ACGT";
        for (int i = 0; i < n_cb; i++) { for (int k = 0; k < 8; k++) fputc(B[(i >> (2 * (k % 4))) & 3], f); fprintf(f, "%d\n", 0); }
        for (int e = 0; e < n_eq; e++) {
            int k = 1 + rng() % 3;
            fprintf(f, "%d", k);
            for (int i = 0; i < k; i++) fprintf(f, "\t%d", (int)(rng() % n_tx));
            int nb = 1 + rng() % 3;
            fprintf(f, "\t%d\t%d", 7, nb);
            for (int b = 0; b < nb; b++) { int nu = 1 + rng() % 2; fprintf(f, "\t%d\t%d", (int)(rng() % n_cb), nu); for (int u = 0; u < nu; u++) fprintf(f, "\tACGTACGT\t%d", 1 + (int)(rng() % 3)); }
            fprintf(f, "\n");
        }
        fclose(f);
        std::vector<int32_t> ref_id;
        for (int it = 0; it < 12; it++) {
            scasa_bfh *h = nullptr;
            int rc = scasa_bfh_open("/tmp/scasa_tsan/bfh.txt", 1 + it % 6, &h);
            if (rc) { printf("bfh rc %d\n", rc); return 1; }
            int64_t a, b, c, d, e2, g, hh;
            scasa_bfh_sizes(h, &a, &b, &c, &d, &e2, &g, &hh);
            std::vector<int32_t> id((size_t)d), cnt((size_t)d), etx((size_t)e2);
            std::vector<int64_t> eo((size_t)c + 1), cp((size_t)b + 1);
            scasa_bfh_read(h, nullptr, nullptr, eo.data(), etx.data(), cp.data(), id.data(), cnt.data());
            scasa_bfh_close(h);
            if (it == 0) ref_id = id; else if (id != ref_id) { printf("BFH MISMATCH\n"); return 1; }
        }
    }
    // ---- expand_pairs is in scasa_cuda.cu (not linked here)
    printf("ok\n");
    return 0;
}
