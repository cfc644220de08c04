// scasa_bfh.cpp — alevin's bfh.txt (salmon alevin --dumpBfh) -> the arrays the quant path takes.
//
// Replaces scasa_quant_step1_1_v1.0.0.R (:16-68), which explodes the file into a data.table with one
// row per (cell, eqclass, transcript) and rbind()s it eqclass by eqclass, plus the ordering step of
// scasa_quant_step2_v1.0.0.R:59-60 (setorder(Barcode), allCells = unique(Barcode)).  Same reading of
// the file, no exploded table:
//
//   line 1-3        number of transcripts, of barcodes, of eqclasses                       step1_1:18
//   next lines      transcript names, then barcode sequences                                :19-20
//   per eqclass     n_tx  tx_idx.. (0-based)  total  n_barcodes  then per barcode:
//                   bc_idx (0-based)  n_umi  (umi_sequence  count) x n_umi                  :41-56
//                   Count of a (cell, eqclass) = n_umi: the script drops every field that
//                   contains one of A, G, T, C and the field after it (:50-51)
//
// Output: cells = the barcodes that occur, in byte order (data.table::setorder is C-locale);
// per cell the (eqclass id, count) pairs in ascending eqclass id (the rbind order); per eqclass its
// transcript indices into the file's transcript list.  Host-only: no CUDA here.
#include "scasa_cuda.h"
#include "scasa_host.h"
#include "scasa_reads.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <climits>
#include <cstdint>
#include <string>
#include <thread>
#include <vector>
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

namespace {

struct Span { const char *b, *e; };
inline Span line_span(const scasa_bfh &f, int64_t k)
{
    const char *b = f.line[(size_t)k], *e = f.line[(size_t)k + 1];
    if (e > b && e[-1] == '\n') e--;
    if (e > b && e[-1] == '\r') e--;
    return {b, e};
}
// names: lines of the bfh file, or the strings a bustools read collected
inline Span tx_name(const scasa_bfh &f, int64_t k)
{
    if (!f.tx_names_v.empty()) { const std::string &s = f.tx_names_v[(size_t)k]; return {s.data(), s.data() + s.size()}; }
    return line_span(f, 3 + k);
}
inline Span cell_name(const scasa_bfh &f, int64_t i)
{
    if (!f.cell_names_v.empty()) { const std::string &s = f.cell_names_v[(size_t)i]; return {s.data(), s.data() + s.size()}; }
    return line_span(f, 3 + f.n_tx + f.cells[(size_t)i]);
}
// next tab-separated field of [p, end); returns false at the end of the line
inline bool next_field(const char *&p, const char *end, Span &out)
{
    if (p >= end) return false;               // also drops a trailing empty field, like strsplit()
    const char *t = (const char *)memchr(p, '\t', (size_t)(end - p));
    out.b = p;
    out.e = t ? t : end;
    p = out.e + 1;                             // one past the end after the last field
    return true;
}
inline bool parse_int(Span s, int64_t &v)
{
    if (s.b == s.e) return false;
    const char *p = s.b;
    bool neg = false;
    if (*p == '-') { neg = true; p++; }
    if (p == s.e) return false;
    int64_t x = 0;
    for (; p < s.e; p++) {
        if (*p < '0' || *p > '9') return false;
        x = x * 10 + (*p - '0');
        if (x > ((int64_t)1 << 40)) return false;
    }
    v = neg ? -x : x;
    return true;
}
// grep("[AGTC]", eq) of step1_1:50
inline bool has_base(Span s)
{
    for (const char *p = s.b; p < s.e; p++)
        if (*p == 'A' || *p == 'G' || *p == 'T' || *p == 'C') return true;
    return false;
}

struct Part {                                  // what one thread parsed from its range of eqclass lines
    std::vector<int64_t> n_tx;                 // per eqclass
    std::vector<int32_t> tx;
    std::vector<int32_t> cb, cnt;              // (barcode, count) pairs in line order
    std::vector<int64_t> n_pairs;              // per eqclass
    std::string err;
};

void parse_range_impl(const scasa_bfh &f, int64_t first_line, int64_t e0, int64_t e1, Part &out);

// thread entry: nothing may escape a std::thread (an allocation failure would otherwise end the process)
void parse_range(const scasa_bfh &f, int64_t first_line, int64_t e0, int64_t e1, Part &out)
{
    try {
        parse_range_impl(f, first_line, e0, e1, out);
    } catch (const std::exception &ex) {
        out.err = std::string("bfh: ") + ex.what();
    }
}

void parse_range_impl(const scasa_bfh &f, int64_t first_line, int64_t e0, int64_t e1, Part &out)
{
    std::vector<Span> rest;
    for (int64_t e = e0; e < e1; e++) {
        Span ln = line_span(f, first_line + e);
        const char *p = ln.b;
        Span fld;
        int64_t k = 0, v = 0;
        auto bad = [&](const char *what) {
            out.err = "bfh: eqclass line " + std::to_string(e + 1) + ": " + what;
        };
        if (!next_field(p, ln.e, fld) || !parse_int(fld, k) || k < 0) { bad("bad transcript count"); return; }
        for (int64_t i = 0; i < k; i++) {
            if (!next_field(p, ln.e, fld) || !parse_int(fld, v) || v < 0 || v >= f.n_tx) { bad("bad transcript index"); return; }
            out.tx.push_back((int32_t)v);
        }
        out.n_tx.push_back(k);
        if (!next_field(p, ln.e, fld) || !parse_int(fld, v)) { bad("bad eqclass count"); return; }
        if (!next_field(p, ln.e, fld) || !parse_int(fld, v)) { bad("bad barcode-group size"); return; }
        // eq2 = eq[-c(pick, pick + 1)]: drop the fields with a base letter and their successors
        rest.clear();
        bool drop_next = false;
        while (next_field(p, ln.e, fld)) {
            const bool umi = has_base(fld);
            if (!umi && !drop_next) rest.push_back(fld);
            drop_next = umi;
        }
        if (rest.size() % 2) { bad("odd number of barcode / count fields"); return; }
        for (size_t i = 0; i < rest.size(); i += 2) {
            int64_t bc = 0, c = 0;
            if (!parse_int(rest[i], bc) || bc < 0 || bc >= f.n_cb) { bad("bad barcode index"); return; }
            if (!parse_int(rest[i + 1], c) || c > INT32_MAX || c < INT32_MIN) { bad("bad UMI count"); return; }
            out.cb.push_back((int32_t)bc);
            out.cnt.push_back((int32_t)c);
        }
        out.n_pairs.push_back((int64_t)rest.size() / 2);
    }
}

int bfh_fail(int code, const std::string &m)
{
    scasa::set_global_error(m);
    return code;
}

}  // namespace

extern "C" {

int scasa_bfh_open(const char *path, int32_t n_threads, scasa_bfh **out)
{
    if (!path || !out) return bfh_fail(SCASA_ERR_ARG, "null argument");
    *out = nullptr;
    try {
        std::unique_ptr<scasa_bfh> f(new scasa_bfh());
        // the file into one malloc'd buffer, read in slices by the threads (a copy out of the page cache each)
        int fd = open(path, O_RDONLY);
        if (fd < 0) return bfh_fail(SCASA_ERR_ARG, std::string("cannot open ") + path);
        struct stat st;
        if (fstat(fd, &st) != 0) { close(fd); return bfh_fail(SCASA_ERR_ARG, std::string("cannot stat ") + path); }
        const int64_t size = (int64_t)st.st_size;
        f->raw = (char *)malloc((size_t)size + 1);
        if (!f->raw) { close(fd); return bfh_fail(SCASA_ERR_NOMEM, "host allocation failed"); }
        f->raw_size = size;
        int nthr = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
        nthr = std::max(1, std::min(nthr, 256));
        {
            const int nr = (int)std::max<int64_t>(1, std::min<int64_t>(nthr, size / ((int64_t)8 << 20) + 1));
            std::vector<char> ok((size_t)nr, 1);
            auto read_slice = [&](int t) {
                int64_t at = size * t / nr;
                const int64_t stop = size * (t + 1) / nr;
                while (at < stop) {
                    const ssize_t got = pread(fd, f->raw + at, (size_t)std::min<int64_t>(stop - at, (int64_t)64 << 20), (off_t)at);
                    if (got <= 0) { ok[(size_t)t] = 0; return; }
                    at += got;
                }
            };
            std::vector<std::thread> th;
            int started = 0;
            try {
                for (; started < nr - 1; started++) th.emplace_back(read_slice, started);
            } catch (const std::exception &) {}
            for (int t = started; t < nr; t++) read_slice(t);
            for (auto &x : th) x.join();
            close(fd);
            for (char k : ok) if (!k) return bfh_fail(SCASA_ERR_ARG, std::string("short read of ") + path);
        }
        f->raw[(size_t)size] = '\n';
        const char *b = f->raw, *end = b + size;
        for (const char *p = b; p < end;) {
            f->line.push_back(p);
            const char *nl = (const char *)memchr(p, '\n', (size_t)(end - p));
            p = nl ? nl + 1 : end;
        }
        f->line.push_back(end);
        const int64_t n_lines = (int64_t)f->line.size() - 1;
        int64_t s[3];
        for (int k = 0; k < 3; k++)
            if (k >= n_lines || !parse_int(line_span(*f, k), s[k]) || s[k] < 0)
                return bfh_fail(SCASA_ERR_ARG, "bfh: the first three lines must be the transcript, barcode and eqclass counts");
        f->n_tx = s[0]; f->n_cb = s[1]; f->n_eq = s[2];
        const int64_t first_eq = 3 + f->n_tx + f->n_cb;
        if (first_eq + f->n_eq > n_lines) return bfh_fail(SCASA_ERR_ARG, "bfh: file is shorter than its header says");
        if (f->n_tx >= INT32_MAX || f->n_cb >= INT32_MAX || f->n_eq >= INT32_MAX)
            return bfh_fail(SCASA_ERR_UNSUPPORTED, "bfh: more than 2^31 transcripts, barcodes or eqclasses");

        nthr = (int)std::max<int64_t>(1, std::min<int64_t>(nthr, f->n_eq / 64 + 1));
        // eqclass ranges of equal size in BYTES (a line holds every cell of its class: lengths differ by orders of magnitude)
        std::vector<int64_t> cut((size_t)nthr + 1, f->n_eq);
        cut[0] = 0;
        {
            const char *b0 = f->line[(size_t)first_eq], *b1 = f->line[(size_t)(first_eq + f->n_eq)];
            for (int t = 1; t < nthr; t++) {
                const char *want = b0 + (b1 - b0) * t / nthr;
                cut[(size_t)t] = std::lower_bound(f->line.begin() + first_eq, f->line.begin() + first_eq + f->n_eq, want) - (f->line.begin() + first_eq);
            }
        }
        std::vector<Part> parts((size_t)nthr);
        {
            std::vector<std::thread> th;
            int started = 0;
            try {
                for (; started < nthr - 1; started++)
                    th.emplace_back(parse_range, std::cref(*f), first_eq, cut[(size_t)started], cut[(size_t)started + 1], std::ref(parts[(size_t)started]));
            } catch (const std::exception &) {}      // fewer threads than asked for: this one parses the rest
            for (int t = started; t < nthr; t++)
                parse_range(*f, first_eq, cut[(size_t)t], cut[(size_t)t + 1], parts[(size_t)t]);
            for (auto &x : th) x.join();
        }
        for (auto &p : parts) if (!p.err.empty()) return bfh_fail(SCASA_ERR_ARG, p.err);

        // eqclass dictionary
        f->eq_off.assign(1, 0);
        f->eq_off.reserve((size_t)f->n_eq + 1);
        int64_t n_pairs = 0;
        for (auto &p : parts) {
            for (int64_t k : p.n_tx) f->eq_off.push_back(f->eq_off.back() + k);
            f->eq_tx.insert(f->eq_tx.end(), p.tx.begin(), p.tx.end());
            n_pairs += (int64_t)p.cb.size();
        }
        // cells: barcodes that occur, ordered by their sequences byte-wise (setorder, step2:59)
        std::vector<int64_t> per_cb((size_t)f->n_cb, 0);
        for (auto &p : parts) for (int32_t cb : p.cb) per_cb[(size_t)cb]++;
        for (int64_t cb = 0; cb < f->n_cb; cb++) if (per_cb[(size_t)cb]) f->cells.push_back((int32_t)cb);
        auto seq = [&](int32_t cb) { return line_span(*f, 3 + f->n_tx + cb); };
        std::stable_sort(f->cells.begin(), f->cells.end(), [&](int32_t x, int32_t y) {
            const Span a = seq(x), c = seq(y);
            const size_t la = (size_t)(a.e - a.b), lc = (size_t)(c.e - c.b);
            const int r = memcmp(a.b, c.b, std::min(la, lc));
            return r ? r < 0 : la < lc;
        });
        // equal sequences are one cell in R (unique(Barcode)): merge them
        f->cell_of_cb.assign((size_t)f->n_cb, -1);
        std::vector<int32_t> uniq;
        for (size_t i = 0; i < f->cells.size(); i++) {
            const Span a = seq(f->cells[i]);
            bool same = false;
            if (!uniq.empty()) {
                const Span c = seq(uniq.back());
                same = (a.e - a.b == c.e - c.b) && !memcmp(a.b, c.b, (size_t)(a.e - a.b));
            }
            if (!same) uniq.push_back(f->cells[i]);
            f->cell_of_cb[(size_t)f->cells[i]] = (int32_t)uniq.size() - 1;
        }
        f->cells.swap(uniq);
        const int64_t n_cells = (int64_t)f->cells.size();
        // CSR by cell, eqclasses ascending inside a cell (a stable counting sort of the pairs in file order)
        f->cell_ptr.assign((size_t)n_cells + 1, 0);
        for (int64_t cb = 0; cb < f->n_cb; cb++)
            if (per_cb[(size_t)cb]) f->cell_ptr[(size_t)f->cell_of_cb[(size_t)cb] + 1] += per_cb[(size_t)cb];
        for (int64_t c = 0; c < n_cells; c++) f->cell_ptr[(size_t)c + 1] += f->cell_ptr[(size_t)c];
        f->eq_id.resize((size_t)n_pairs);
        f->eq_count.resize((size_t)n_pairs);
        // every part scatters its own pairs: its first slot in a cell = the cell's start + what the earlier parts
        // (earlier in the file) hold for that cell, so the order inside a cell stays the file order
        {
            const size_t np_ = parts.size();
            std::vector<std::vector<int64_t>> cur(np_);
            std::vector<int64_t> e_first(np_ + 1, 0);
            for (size_t t = 0; t < np_; t++) e_first[t + 1] = e_first[t] + (int64_t)parts[t].n_pairs.size();
            auto count_part = [&](size_t t) {
                cur[t].assign((size_t)n_cells, 0);
                for (int32_t cb : parts[t].cb) cur[t][(size_t)f->cell_of_cb[(size_t)cb]]++;
            };
            auto scatter_part = [&](size_t t) {
                const Part &p = parts[t];
                int64_t e = e_first[t];
                size_t at = 0;
                for (int64_t np : p.n_pairs) {
                    for (int64_t k = 0; k < np; k++, at++) {
                        const int64_t pos = cur[t][(size_t)f->cell_of_cb[(size_t)p.cb[at]]]++;
                        f->eq_id[(size_t)pos] = (int32_t)e;
                        f->eq_count[(size_t)pos] = p.cnt[at];
                    }
                    e++;
                }
            };
            auto run_all = [&](auto &&fn) {
                std::vector<std::thread> th;
                size_t started = 0;
                try {
                    for (; started + 1 < np_; started++) th.emplace_back(fn, started);
                } catch (const std::exception &) {}
                for (size_t t = started; t < np_; t++) fn(t);
                for (auto &x : th) x.join();
            };
            run_all(count_part);
            for (int64_t c = 0; c < n_cells; c++) {              // counts -> first slots
                int64_t at = f->cell_ptr[(size_t)c];
                for (size_t t = 0; t < np_; t++) { const int64_t m = cur[t][(size_t)c]; cur[t][(size_t)c] = at; at += m; }
            }
            run_all(scatter_part);
        }
        *out = f.release();
        return SCASA_OK;
    } catch (const std::bad_alloc &) {
        return bfh_fail(SCASA_ERR_NOMEM, "host allocation failed");
    } catch (const std::exception &ex) {
        return bfh_fail(SCASA_ERR_ARG, ex.what());
    }
}

void scasa_bfh_close(scasa_bfh *f) { delete f; }

int scasa_bfh_sizes(const scasa_bfh *f, int64_t *n_tx, int64_t *n_cells, int64_t *n_eq, int64_t *n_pairs, int64_t *n_eq_tx,
                    int64_t *tx_name_bytes, int64_t *barcode_bytes)
{
    if (!f) return bfh_fail(SCASA_ERR_ARG, "null argument");
    if (n_tx) *n_tx = f->n_tx;
    if (n_cells) *n_cells = (int64_t)f->cells.size();
    if (n_eq) *n_eq = f->n_eq;
    if (n_pairs) *n_pairs = (int64_t)f->eq_id.size();
    if (n_eq_tx) *n_eq_tx = (int64_t)f->eq_tx.size();
    if (tx_name_bytes) {
        int64_t b = 0;
        for (int64_t k = 0; k < f->n_tx; k++) { Span s = tx_name(*f, k); b += (s.e - s.b) + 1; }
        *tx_name_bytes = b;
    }
    if (barcode_bytes) {
        int64_t b = 0;
        for (int64_t i = 0; i < (int64_t)f->cells.size(); i++) { Span s = cell_name(*f, i); b += (s.e - s.b) + 1; }
        *barcode_bytes = b;
    }
    return SCASA_OK;
}

int scasa_bfh_read(const scasa_bfh *f, char *tx_names, char *barcodes, int64_t *eq_off, int32_t *eq_tx, int64_t *cell_ptr,
                   int32_t *eq_id, int32_t *eq_count)
{
    if (!f) return bfh_fail(SCASA_ERR_ARG, "null argument");
    if (tx_names)
        for (int64_t k = 0; k < f->n_tx; k++) {
            Span s = tx_name(*f, k);
            memcpy(tx_names, s.b, (size_t)(s.e - s.b));
            tx_names += s.e - s.b;
            *tx_names++ = '\n';
        }
    if (barcodes)
        for (int64_t i = 0; i < (int64_t)f->cells.size(); i++) {
            Span s = cell_name(*f, i);
            memcpy(barcodes, s.b, (size_t)(s.e - s.b));
            barcodes += s.e - s.b;
            *barcodes++ = '\n';
        }
    if (eq_off) std::copy(f->eq_off.begin(), f->eq_off.end(), eq_off);
    if (eq_tx) std::copy(f->eq_tx.begin(), f->eq_tx.end(), eq_tx);
    if (cell_ptr) std::copy(f->cell_ptr.begin(), f->cell_ptr.end(), cell_ptr);
    if (eq_id) std::copy(f->eq_id.begin(), f->eq_id.end(), eq_id);
    if (eq_count) std::copy(f->eq_count.begin(), f->eq_count.end(), eq_count);
    return SCASA_OK;
}

}  // extern "C"
