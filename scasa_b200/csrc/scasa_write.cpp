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
#include <charconv>
#include <climits>
#include <cmath>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <algorithm>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

namespace {

const long double kTbl[] = {1e0L, 1e1L, 1e2L, 1e3L, 1e4L, 1e5L, 1e6L, 1e7L, 1e8L, 1e9L, 1e10L, 1e11L, 1e12L, 1e13L, 1e14L,
                            1e15L, 1e16L, 1e17L, 1e18L, 1e19L, 1e20L, 1e21L, 1e22L};
constexpr int kDigits = 15;        // R_print.digits inside writetable
constexpr int kKpMax = 22;         // KP_MAX of format.c
constexpr int kDecMinExp = -308;   // R_dec_min_exponent on IEEE doubles

// powl(10.0L, k) for the exponents a double can have: the values of the very call format.c makes, computed once
// (powl itself costs ~300 ns, more than the rest of a number's formatting together)
inline long double pow10l_cached(int k)
{
    constexpr int kLo = -700, kHi = 700;
    static const std::vector<long double> tbl = [] {
        std::vector<long double> t((size_t)(kHi - kLo + 1));
        for (int e = kLo; e <= kHi; e++) t[(size_t)(e - kLo)] = powl(10.0L, e);
        return t;
    }();
    return (k >= kLo && k <= kHi) ? tbl[(size_t)(k - kLo)] : powl(10.0L, k);
}

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
    } else if (kp <= kDecMinExp) r_prec = (r * 1e+303L) / pow10l_cached(kp + 303);
    else r_prec /= pow10l_cached(kp);
    if (r_prec < kTbl[R - 1]) { r_prec *= 10.0L; kp--; }
    // nearbyintl (x87 frndint) is slow; below 2^63 llrintl rounds the same way (current mode: to nearest even)
    double alpha = r_prec < 9.0e18L ? (double)llrintl(r_prec) : (double)nearbyintl(r_prec);
    nsig = R;
    for (int j = 1; j <= R; j++) {
        alpha /= 10.0;
        if (alpha == std::floor(alpha)) nsig--; else break;
    }
    if (nsig == 0 && R > 0) { nsig = 1; kp += 1; }
    kpower = kp + R - 1;
    widens = kpower > 0 && kpower <= kKpMax && r < (double)kTbl[kpower];
}

// EncodeReal0(x, w, d, e) after formatReal(&x, 1, ...): the text of one number; returns its length.
// The digits are sprintf's (correctly rounded), produced by std::to_chars, which gives the same characters as
// "%.*f" / "%.*e" several times faster; scientific() only decides how many digits and which notation.
int r_format_real(double x, char *buf)
{
    if (x == 0.0) { buf[0] = '0'; buf[1] = 0; return 1; }             // also strips the sign of -0
    if (std::isnan(x)) { strcpy(buf, "NaN"); return 3; }              // (R would print NA for NA_real_; results hold no NA)
    if (std::isinf(x)) { strcpy(buf, x > 0 ? "Inf" : "-Inf"); return (int)strlen(buf); }
    // whole numbers below 1e15 (the pass-through counts): the digits are exact, only the notation has to be decided
    if (x > 0 && x < 1e15 && x == std::floor(x)) {
        unsigned long long v = (unsigned long long)x;
        char dig[24];
        int L = 0;
        while (v) { dig[L++] = (char)('0' + v % 10); v /= 10; }      // least significant first
        int z = 0;
        while (z < L - 1 && dig[z] == '0') z++;
        const int nsig = L - z, kpower = L - 1;
        const int wF = L, wE = (nsig > 1 ? nsig + 1 : 1) + 4;        // e+XX: two exponent digits below 1e100
        int n = 0;
        if (wF <= wE) {
            for (int i = L - 1; i >= 0; i--) buf[n++] = dig[i];
        } else {
            buf[n++] = dig[L - 1];
            if (nsig > 1) { buf[n++] = '.'; for (int i = L - 2; i >= z; i--) buf[n++] = dig[i]; }
            buf[n++] = 'e'; buf[n++] = '+'; buf[n++] = (char)('0' + kpower / 10); buf[n++] = (char)('0' + kpower % 10);
        }
        buf[n] = 0;
        return n;
    }
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
    std::to_chars_result r = wF <= wE ? std::to_chars(buf, buf + 60, x, std::chars_format::fixed, rgt)
                                      : std::to_chars(buf, buf + 60, x, std::chars_format::scientific, d);
    *r.ptr = 0;
    return (int)(r.ptr - buf);
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

}  // extern "C"

namespace {

// Blocks of output lines are formatted by a pool of threads and written in order by whoever completes
// the next block in line; at most 4 blocks per thread are ahead of the writer (bounded memory).
template <class F> bool write_blocks_in_order(FILE *fp, int64_t n_blocks, int n_threads, F &&format_block)
{
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
        for (;;) {
            const int64_t b = next.fetch_add(1);
            if (b >= n_blocks) return;
            {
                std::unique_lock<std::mutex> lk(m);
                cv.wait(lk, [&] { return b < written + 4 * (int64_t)nthr || io_error; });
                if (io_error) return;
            }
            std::string out;
            format_block(b, out);
            std::unique_lock<std::mutex> lk(m);
            ready[(size_t)b].swap(out);
            done[(size_t)b] = 1;
            if (!draining) {                 // one writer at a time: the lock is released during fwrite, and a second
                draining = true;             // thread entering here would take the same block again (empty) and skip one
                while (written < n_blocks && done[(size_t)written]) {
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
    return !io_error;
}

void append_name(std::string &out, const std::vector<const char *> &names, int64_t j)
{
    const char *e = names[(size_t)j + 1];
    if (e > names[(size_t)j] && e[-1] == '\n') e--;
    out.append(names[(size_t)j], e);
}

bool write_header(FILE *fp, const std::vector<const char *> &cn, int64_t n_cells)
{   // cell names, no field for the row-name column
    std::string h;
    for (int64_t c = 0; c < n_cells; c++) {
        if (c) h.push_back('\t');
        append_name(h, cn, c);
    }
    h.push_back('\n');
    return fwrite(h.data(), 1, h.size(), fp) == h.size();
}

}  // namespace

extern "C" {

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
        bool ok = write_header(fp, cn, n_cells);
        // blocks of kBlock output lines: gather the block's columns (a strided read of the cell-major matrix,
        // done once per block), format, write in order
        constexpr int64_t kBlock = 32;
        const int64_t n_blocks = ((int64_t)kept.size() + kBlock - 1) / kBlock;
        ok = ok && write_blocks_in_order(fp, n_blocks, n_threads, [&](int64_t b, std::string &out) {
            std::vector<double> tile((size_t)(kBlock * n_cells));
            char num[64];
            const int64_t k0 = b * kBlock, k1 = std::min<int64_t>(k0 + kBlock, (int64_t)kept.size());
            for (int64_t c = 0; c < n_cells; c++) {
                const double *row = mat + c * n_cols;
                for (int64_t k = k0; k < k1; k++) tile[(size_t)((k - k0) * n_cells + c)] = row[kept[(size_t)k]];
            }
            out.reserve((size_t)((k1 - k0) * (n_cells * 3 + 64)));
            for (int64_t k = k0; k < k1; k++) {
                append_name(out, rn, kept[(size_t)k]);
                const double *v = &tile[(size_t)((k - k0) * n_cells)];
                for (int64_t c = 0; c < n_cells; c++) {
                    out.push_back('\t');
                    if (v[c] == 0.0) out.push_back('0');
                    else out.append(num, (size_t)r_format_real(v[c], num));
                }
                out.push_back('\n');
            }
        });
        const bool bad = (fclose(fp) != 0) | !ok;
        if (bad) return wfail(SCASA_ERR_ARG, std::string("write to ") + path + " failed");
        return SCASA_OK;
    } catch (const std::bad_alloc &) {
        return wfail(SCASA_ERR_NOMEM, "host allocation failed");
    } catch (const std::exception &ex) {
        return wfail(SCASA_ERR_ARG, ex.what());
    }
}

// CSR by row -> CSC, entries of a column in row order.  The rows are cut into ranges of about equal entry
// count; every range counts its entries per column, the counts give each range its own segment of every
// column (ranges in row order), then every range scatters its entries: no atomics, a fixed result.
int scasa_csr_transpose(int64_t n_rows, int64_t n_cols, const int64_t *rowptr, const int32_t *cols, const double *vals,
                        int64_t *colptr, int32_t *rows_out, double *vals_out, int32_t n_threads)
{
    if (n_rows < 0 || n_cols <= 0 || !rowptr || !colptr || (rowptr[n_rows] > 0 && (!cols || !vals || !rows_out || !vals_out)))
        return wfail(SCASA_ERR_ARG, "bad argument");
    try {
        int nthr = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
        nthr = std::max(1, std::min(nthr, 256));
        const int64_t nnz = rowptr[n_rows];
        const int n_rng = (int)std::max<int64_t>(1, std::min<int64_t>((int64_t)nthr * 4, std::max<int64_t>(1, nnz >> 16)));
        std::vector<int64_t> cut((size_t)n_rng + 1, 0);
        for (int r = 1; r < n_rng; r++) {
            const int64_t want = rowptr[0] + (nnz - rowptr[0]) * r / n_rng;
            cut[(size_t)r] = std::lower_bound(rowptr, rowptr + n_rows + 1, want) - rowptr;
            cut[(size_t)r] = std::max(cut[(size_t)r], cut[(size_t)r - 1]);
            cut[(size_t)r] = std::min(cut[(size_t)r], n_rows);
        }
        cut[(size_t)n_rng] = n_rows;
        std::vector<std::vector<int64_t>> hist((size_t)n_rng);
        std::atomic<int> next{0};
        std::atomic<bool> bad{false};
        auto run = [&](auto &&body) {
            next = 0;
            auto loop = [&] { for (;;) { const int r = next.fetch_add(1); if (r >= n_rng) return; body(r); } };
            std::vector<std::thread> th;
            try { for (int t = 1; t < std::min(nthr, n_rng); t++) th.emplace_back(loop); } catch (const std::exception &) {}
            loop();
            for (auto &t : th) t.join();
        };
        run([&](int r) {
            try {
                hist[(size_t)r].assign((size_t)n_cols, 0);
                for (int64_t k = rowptr[cut[(size_t)r]]; k < rowptr[cut[(size_t)r + 1]]; k++) {
                    const int32_t j = cols[k];
                    if (j < 0 || j >= n_cols) { bad = true; return; }
                    hist[(size_t)r][(size_t)j]++;
                }
            } catch (const std::exception &) { bad = true; }
        });
        if (bad) return wfail(SCASA_ERR_ARG, "column index out of range (or out of memory)");
        int64_t acc = 0;
        for (int64_t j = 0; j < n_cols; j++) {
            colptr[j] = acc;
            for (int r = 0; r < n_rng; r++) { const int64_t h = hist[(size_t)r][(size_t)j]; hist[(size_t)r][(size_t)j] = acc; acc += h; }
        }
        colptr[n_cols] = acc;
        run([&](int r) {
            std::vector<int64_t> &cur = hist[(size_t)r];
            for (int64_t i = cut[(size_t)r]; i < cut[(size_t)r + 1]; i++)
                for (int64_t k = rowptr[i]; k < rowptr[i + 1]; k++) {
                    const int64_t pos = cur[(size_t)cols[k]]++;
                    rows_out[pos] = (int32_t)i;
                    vals_out[pos] = vals[k];
                }
        });
        return SCASA_OK;
    } catch (const std::bad_alloc &) {
        return wfail(SCASA_ERR_NOMEM, "host allocation failed");
    }
}

// write.table of a matrix held column-wise and sparse: source column j (a transcript, or a gene) has the entries
// [colptr[j], colptr[j+1]) = (cell, value), cells ascending.  Output line k is source column out_src[k] under the
// name out_names[k]; cells without an entry, and stored zeros, print as 0 — the bytes scasa_write_table produces
// from the dense matrix, without the dense matrix.
int scasa_write_table_sparse(const char *path, int64_t n_cells, int64_t n_src, const int64_t *colptr, const int32_t *cell_idx,
                             const double *vals, int64_t n_out, const int32_t *out_src, const char *out_names,
                             const char *cell_names, int32_t n_threads, int64_t *bytes_written)
{
    if (!path || n_cells <= 0 || n_src <= 0 || !colptr || n_out < 0 || (n_out > 0 && (!out_src || !out_names)) || !cell_names)
        return wfail(SCASA_ERR_ARG, "bad argument");
    try {
        const std::vector<const char *> rn = split_lines(out_names, n_out), cn = split_lines(cell_names, n_cells);
        if ((int64_t)rn.size() != n_out + 1 || (int64_t)cn.size() != n_cells + 1)
            return wfail(SCASA_ERR_ARG, "out_names / cell_names must hold one '\\n'-terminated name per line / cell");
        for (int64_t k = 0; k < n_out; k++)
            if (out_src[k] < 0 || out_src[k] >= n_src) return wfail(SCASA_ERR_ARG, "out_src out of range");
        FILE *fp = fopen(path, "wb");
        if (!fp) return wfail(SCASA_ERR_ARG, std::string("cannot create ") + path);
        bool ok = write_header(fp, cn, n_cells);
        std::string zeros((size_t)n_cells * 2, '\t');              // "\t0\t0..." : the text of a run of empty cells
        for (int64_t c = 0; c < n_cells; c++) zeros[(size_t)c * 2 + 1] = '0';
        constexpr int64_t kBlock = 32;
        const int64_t n_blocks = (n_out + kBlock - 1) / kBlock;
        ok = ok && write_blocks_in_order(fp, n_blocks, n_threads, [&](int64_t b, std::string &out) {
            char num[64];
            const int64_t k0 = b * kBlock, k1 = std::min<int64_t>(k0 + kBlock, n_out);
            size_t need = 0;
            for (int64_t k = k0; k < k1; k++) need += (size_t)n_cells * 2 + 72 + (size_t)(colptr[out_src[k] + 1] - colptr[out_src[k]]) * 18;
            out.reserve(need);
            for (int64_t k = k0; k < k1; k++) {
                append_name(out, rn, k);
                const int64_t j = out_src[k];
                int64_t c = 0;
                for (int64_t e = colptr[j]; e < colptr[j + 1]; e++) {
                    const int64_t ce = cell_idx[e];
                    if (ce < c || ce >= n_cells) throw std::runtime_error("cells of a column must be ascending and < n_cells");
                    out.append(zeros, 0, (size_t)(ce - c) * 2);
                    out.push_back('\t');
                    const double v = vals[e];
                    if (v == 0.0) out.push_back('0');
                    else out.append(num, (size_t)r_format_real(v, num));
                    c = ce + 1;
                }
                out.append(zeros, 0, (size_t)(n_cells - c) * 2);
                out.push_back('\n');
            }
        });
        const long long pos = (long long)ftello(fp);
        const bool bad = (fclose(fp) != 0) | !ok;
        if (bad) return wfail(SCASA_ERR_ARG, std::string("write to ") + path + " failed");
        if (bytes_written) *bytes_written = pos;
        return SCASA_OK;
    } catch (const std::bad_alloc &) {
        return wfail(SCASA_ERR_NOMEM, "host allocation failed");
    } catch (const std::exception &ex) {
        return wfail(SCASA_ERR_ARG, ex.what());
    }
}

}  // extern "C"
