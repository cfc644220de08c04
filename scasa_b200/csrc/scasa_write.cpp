// scasa_write.cpp — the two text outputs, byte-compatible with the reference's
//     write.table(isoform_count, "scasa_isoform_counts.txt", quote = F, row.names = T, sep = "\t")   step2:170
//     write.table(gene_count,    "scasa_gene_expression.txt", quote = F, row.names = T, sep = "\t")   step3:52
// from the library's cell-major result matrices: the file is the transpose (one line per transcript
// or gene, one field per cell, step2:166), restricted to the rows that survive step2:167, with a
// header of cell names that has one field fewer than the data lines (row.names = T, col.names = T).
//
// Numbers are written the way R's C code does it (src/library/utils/src/io.c writetable ->
// EncodeElement0 -> formatReal / scientific() with R_print.digits = DBL_DIG = 15): every value on its
// own, the fewest significant digits (<= 15) that reproduce it to 15, fixed notation unless
// scientific is narrower.  scientific() is restated with the same long-double arithmetic.
// Host-only: no CUDA here.
#include "scasa_cuda.h"
#include "scasa_host.h"

#include <atomic>
#include <climits>
#include <cmath>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace {

const long double kTbl[] = {1e0L, 1e1L, 1e2L, 1e3L, 1e4L, 1e5L, 1e6L, 1e7L, 1e8L, 1e9L, 1e10L, 1e11L, 1e12L, 1e13L, 1e14L,
                            1e15L, 1e16L, 1e17L, 1e18L, 1e19L, 1e20L, 1e21L, 1e22L};
constexpr int kDigits = 15;        // R_print.digits inside writetable
constexpr int kKpMax = 22;         // KP_MAX of format.c
constexpr int kDecMinExp = -308;   // R_dec_min_exponent on IEEE doubles

// format.c scientific() for one finite non-zero |x|, digits = 15 (the long double branch)
void r_scientific(double x, int &neg, int &kpower, int &nsig, bool &widens)
{
    const int R = kDigits;
    double r;
    if (x < 0.0) { neg = 1; r = -x; } else { neg = 0; r = x; }
    int kp = (int)std::floor(std::log10(r)) - R + 1;
    long double r_prec = r;
    if (std::abs(kp) < 10) {
        if (kp > 0) r_prec /= kTbl[kp]; else if (kp < 0) r_prec *= kTbl[-kp];
    } else if (kp <= kDecMinExp) r_prec = (r * 1e+303L) / powl(10.0L, kp + 303);
    else r_prec /= powl(10.0L, kp);
    if (r_prec < kTbl[R - 1]) { r_prec *= 10.0L; kp--; }
    double alpha = (double)nearbyintl(r_prec);
    nsig = R;
    for (int j = 1; j <= R; j++) {
        alpha /= 10.0;
        if (alpha == std::floor(alpha)) nsig--; else break;
    }
    if (nsig == 0 && R > 0) { nsig = 1; kp += 1; }
    kpower = kp + R - 1;
    widens = kpower > 0 && kpower <= kKpMax && r < (double)kTbl[kpower];
}

// EncodeReal0(x, w, d, e) after formatReal(&x, 1, ...): the text of one number; returns its length
int r_format_real(double x, char *buf)
{
    if (x == 0.0) { buf[0] = '0'; buf[1] = 0; return 1; }             // also strips the sign of -0
    if (std::isnan(x)) { strcpy(buf, "NaN"); return 3; }              // (R would print NA for NA_real_; results hold no NA)
    if (std::isinf(x)) { strcpy(buf, x > 0 ? "Inf" : "-Inf"); return (int)strlen(buf); }
    int neg, kpower, nsig;
    bool widens;
    r_scientific(x, neg, kpower, nsig, widens);
    int left = kpower + 1;
    if (widens) left--;
    const int sleft = neg + (left <= 0 ? 1 : left);
    int rgt = nsig - left;
    if (rgt < 0) rgt = 0;
    if (rgt > kKpMax) rgt = kKpMax;
    const int wF = sleft + rgt + (rgt != 0);
    const int e = (kpower >= 100 || kpower <= -99) ? 2 : 1;
    const int d = nsig - 1;
    const int wE = neg + (d > 0) + d + 4 + e;
    if (wF <= wE) return snprintf(buf, 64, "%.*f", rgt, x);
    return d ? snprintf(buf, 64, "%#.*e", d, x) : snprintf(buf, 64, "%.*e", d, x);
}

std::vector<const char *> split_lines(const char *s, int64_t n)
{
    std::vector<const char *> out;
    const char *p = s;
    for (int64_t k = 0; k < n; k++) {
        out.push_back(p);
        const char *nl = strchr(p, '\n');
        if (!nl) { out.push_back(p + strlen(p)); return out; }
        p = nl + 1;
    }
    out.push_back(p);
    return out;
}

int wfail(int code, const std::string &m)
{
    scasa::set_global_error(m);
    return code;
}

}  // namespace

extern "C" {

int scasa_format_real15(double x, char *buf, int32_t cap)
{
    char tmp[64];
    const int n = r_format_real(x, tmp);
    if (!buf || cap <= n) return wfail(SCASA_ERR_ARG, "buffer too small");
    memcpy(buf, tmp, (size_t)n + 1);
    return n;
}

int scasa_write_table(const char *path, const double *mat, int64_t n_cells, int64_t n_cols, const uint8_t *keep,
                      const char *row_names, const char *cell_names, int32_t n_threads)
{
    if (!path || !mat || n_cells <= 0 || n_cols <= 0 || !row_names || !cell_names) return wfail(SCASA_ERR_ARG, "bad argument");
    try {
        const std::vector<const char *> rn = split_lines(row_names, n_cols), cn = split_lines(cell_names, n_cells);
        if ((int64_t)rn.size() != n_cols + 1 || (int64_t)cn.size() != n_cells + 1)
            return wfail(SCASA_ERR_ARG, "row_names / cell_names must hold one '\\n'-terminated name per column / cell");
        FILE *fp = fopen(path, "wb");
        if (!fp) return wfail(SCASA_ERR_ARG, std::string("cannot create ") + path);
        std::vector<int64_t> kept;
        for (int64_t j = 0; j < n_cols; j++) if (!keep || keep[j]) kept.push_back(j);
        {   // header: cell names, no field for the row-name column
            std::string h;
            for (int64_t c = 0; c < n_cells; c++) {
                if (c) h.push_back('\t');
                const char *e = cn[(size_t)c + 1];
                if (e > cn[(size_t)c] && e[-1] == '\n') e--;
                h.append(cn[(size_t)c], e);
            }
            h.push_back('\n');
            fwrite(h.data(), 1, h.size(), fp);
        }
        // blocks of kBlock output lines: gather the block's columns (a strided read of the cell-major matrix,
        // done once per block), format, write in order
        constexpr int64_t kBlock = 32;
        const int64_t n_blocks = ((int64_t)kept.size() + kBlock - 1) / kBlock;
        int nthr = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
        nthr = (int)std::max<int64_t>(1, std::min<int64_t>(std::min(nthr, 256), n_blocks));
        std::vector<std::string> ready((size_t)n_blocks);
        std::vector<char> done((size_t)n_blocks, 0);
        std::atomic<int64_t> next{0};
        std::mutex m;
        std::condition_variable cv;
        int64_t written = 0;
        bool io_error = false, draining = false;
        auto work_impl = [&]() {
            std::vector<double> tile((size_t)(kBlock * n_cells));
            char num[64];
            for (;;) {
                const int64_t b = next.fetch_add(1);
                if (b >= n_blocks) return;
                {   // keep at most 4 * nthr blocks ahead of the writer (bounded memory)
                    std::unique_lock<std::mutex> lk(m);
                    cv.wait(lk, [&] { return b < written + 4 * (int64_t)nthr || io_error; });
                    if (io_error) return;
                }
                const int64_t k0 = b * kBlock, k1 = std::min<int64_t>(k0 + kBlock, (int64_t)kept.size());
                for (int64_t c = 0; c < n_cells; c++) {
                    const double *row = mat + c * n_cols;
                    for (int64_t k = k0; k < k1; k++) tile[(size_t)((k - k0) * n_cells + c)] = row[kept[(size_t)k]];
                }
                std::string out;
                out.reserve((size_t)((k1 - k0) * (n_cells * 3 + 64)));
                for (int64_t k = k0; k < k1; k++) {
                    const int64_t j = kept[(size_t)k];
                    const char *e = rn[(size_t)j + 1];
                    if (e > rn[(size_t)j] && e[-1] == '\n') e--;
                    out.append(rn[(size_t)j], e);
                    const double *v = &tile[(size_t)((k - k0) * n_cells)];
                    for (int64_t c = 0; c < n_cells; c++) {
                        out.push_back('\t');
                        if (v[c] == 0.0) out.push_back('0');
                        else out.append(num, (size_t)r_format_real(v[c], num));
                    }
                    out.push_back('\n');
                }
                std::unique_lock<std::mutex> lk(m);
                ready[(size_t)b].swap(out);
                done[(size_t)b] = 1;
                if (!draining) {                 // one writer at a time: the lock is released during fwrite, and a second
                    draining = true;             // thread entering here would take the same block again (empty) and skip one
                    while (written < n_blocks && done[(size_t)written]) {    // whoever completes the next block in line writes
                        std::string s;
                        s.swap(ready[(size_t)written]);
                        lk.unlock();
                        const bool ok = fwrite(s.data(), 1, s.size(), fp) == s.size();
                        lk.lock();
                        if (!ok) io_error = true;
                        written++;
                        cv.notify_all();
                    }
                    draining = false;
                }
                cv.notify_all();
            }
        };
        auto work = [&]() {                           // thread entry: nothing may escape a std::thread
            try {
                work_impl();
            } catch (const std::exception &) {        // allocation failure: stop everyone, report below
                std::lock_guard<std::mutex> lk(m);
                io_error = true;
                cv.notify_all();
            }
        };
        std::vector<std::thread> th;
        try {
            for (int t = 1; t < nthr; t++) th.emplace_back(work);
        } catch (const std::exception &) {}          // fewer threads than asked for: the ones that started and this one do the work
        work();
        for (auto &t : th) t.join();
        const bool bad = (fclose(fp) != 0) | io_error;
        if (bad) return wfail(SCASA_ERR_ARG, std::string("write to ") + path + " failed");
        return SCASA_OK;
    } catch (const std::bad_alloc &) {
        return wfail(SCASA_ERR_NOMEM, "host allocation failed");
    } catch (const std::exception &ex) {
        return wfail(SCASA_ERR_ARG, ex.what());
    }
}

}  // extern "C"
