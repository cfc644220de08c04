// scasa_host.cpp — eqclass -> (block, row) map: crpcount_td + hamming_correction, once per
// distinct equivalence class.  Line numbers: /root/reference/scasa/SCRIPTS/Scasa_Rsource.R.
#include "scasa_host.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <atomic>
#include <thread>
#include <immintrin.h>
#include <cstdint>

namespace scasa {

// ---------------------------------------------------------------------------------------------
// format(x) at getOption("digits") = 7 for one numeric column, then as.numeric() of the text.
// This is what apply(current, 1, ...) at :617/:626 does to the numeric columns of `current`
// because its first column is character (as.matrix.data.frame formats every numeric column):
// common number of decimals (or common mantissa length in scientific notation) over the rows
// that are candidates, fixed notation unless scientific is narrower (R's formatReal).
void r_format_roundtrip(const double *v, int n, double *back, uint8_t *is_zero_text)
{
    const int digits = 7;
    int rgt = -1000000, mxsl = -1000000, mxe = -1000000, mne = 1000000, mxns = -1000000, neg = 0;
    bool any = false;
    char buf[64];
    for (int i = 0; i < n; i++) {
        double x = v[i];
        if (!std::isfinite(x)) continue;
        any = true;
        int kp = 0, nsig = 1, neg_i = x < 0;
        bool widens = false;
        if (x != 0.0) {
            snprintf(buf, sizeof buf, "%.*e", digits - 1, std::fabs(x));   // d.dddddde[+-]XX
            kp = atoi(strchr(buf, 'e') + 1);
            nsig = digits;
            for (int k = digits; k >= 2; k--) {                            // strip trailing zeros of the mantissa
                if (buf[k] == '0') nsig--; else break;                     // buf[0]='d', buf[1]='.', buf[2..7]
            }
            widens = kp > 0 && std::fabs(x) < std::pow(10.0, kp);
        }
        int left = kp + 1;
        if (widens) left--;
        int sleft = neg_i + (left <= 0 ? 1 : left);
        int right = nsig - left;
        if (neg_i) neg = 1;
        rgt = std::max(rgt, right);
        mxsl = std::max(mxsl, sleft);
        mxe = std::max(mxe, kp);
        mne = std::min(mne, kp);
        mxns = std::max(mxns, nsig);
    }
    bool fixed = true;
    int d = 0;
    if (any) {
        if (mxsl < 0) mxsl = 1 + neg;
        if (rgt < 0) rgt = 0;
        if (rgt > 350) rgt = 350;
        int wF = mxsl + rgt + (rgt != 0);
        int e = (mxe >= 100 || mne <= -99) ? 2 : 1;
        int dE = mxns - 1;
        int wE = neg + (dE > 0) + dE + 4 + e;
        if (wF <= wE) { fixed = true; d = rgt; } else { fixed = false; d = dE; }
    }
    for (int i = 0; i < n; i++) {
        double x = v[i];
        if (!std::isfinite(x)) { back[i] = x; is_zero_text[i] = 0; continue; }
        if (x == 0.0) x = 0.0;                                              // strip negative zero
        char out[512];
        if (fixed) snprintf(out, sizeof out, "%.*f", d, x); else snprintf(out, sizeof out, "%.*e", d, x);
        back[i] = strtod(out, nullptr);
        is_zero_text[i] = (strcmp(out, "0") == 0);
    }
}

static double r_median(std::vector<double> &v)
{
    if (v.empty()) return NAN;
    for (double x : v) if (std::isnan(x)) return NAN;
    std::sort(v.begin(), v.end());
    size_t n = v.size(), h = n / 2;
    return (n & 1) ? v[h] : (v[h - 1] + v[h]) / 2.0;
}

// ---------------------------------------------------------------------------------------------
void CrpIndex::build(int n_blocks, const int32_t *q_off, const int32_t *p_off, const int64_t *x_off,
                     const double *x, const int32_t *col_tx, const uint8_t *row_pat, const int32_t *name_rank,
                     int hamming_dist)
{
    B_ = n_blocks; hd_ = hamming_dist;
    q_off_.assign(q_off, q_off + B_ + 1);
    p_off_.assign(p_off, p_off + B_ + 1);
    x_off_.assign(x_off, x_off + B_ + 1);
    x_.assign(x, x + x_off[B_]);
    col_tx_.assign(col_tx, col_tx + p_off[B_]);
    pat_.assign(row_pat, row_pat + x_off[B_]);
    name_rank_.assign(name_rank, name_rank + B_);
    supported_.assign((size_t)p_off[B_], 0);
    hits_.clear(); hits_.reserve((size_t)p_off[B_]);
    for (int k = 0; k < B_; k++) {
        int q = q_off_[k + 1] - q_off_[k], p = p_off_[k + 1] - p_off_[k];
        for (int j = 0; j < p; j++) {
            long double s = 0;                                   // colSums(crp) > 0   :468, :490
            for (int r = 0; r < q; r++) s += x_[(size_t)x_off_[k] + (size_t)j * q + r];
            uint8_t sup = (double)s > 0;
            supported_[(size_t)p_off_[k] + j] = sup;
            hits_.push_back(Hit{k, p_off_[k] + j, sup});
        }
        if (q * p == 1 && !supported_[(size_t)p_off_[k]])
            throw std::runtime_error("crp: a 1x1 block with value 0 would count an eqclass twice; not supported");
    }
    std::stable_sort(hits_.begin(), hits_.end(), [&](const Hit &a, const Hit &b) {
        return col_tx_[(size_t)a.col] < col_tx_[(size_t)b.col];
    });
    tx_range_.clear(); tx_range_.reserve(hits_.size() * 2);
    for (size_t i = 0; i < hits_.size();) {
        size_t j = i;
        int32_t t = col_tx_[(size_t)hits_[i].col];
        while (j < hits_.size() && col_tx_[(size_t)hits_[j].col] == t) j++;
        tx_range_[t] = {(int32_t)i, (int32_t)j};
        i = j;
    }
}

// which.max(apply(current, 1, function(x) median(as.numeric(as.character(x[x > 0])))))  :616-617, :625-626
int32_t CrpIndex::tie_break(int block, const std::vector<int> &cand) const
{
    const int q = q_off_[block + 1] - q_off_[block], p = p_off_[block + 1] - p_off_[block];
    const size_t xo = (size_t)x_off_[block];
    const int nc = (int)cand.size();
    std::vector<double> col(nc), back((size_t)nc * p);
    std::vector<uint8_t> zt((size_t)nc * p), z(nc);
    std::vector<double> b(nc);
    for (int j = 0; j < p; j++) {
        for (int i = 0; i < nc; i++) col[i] = x_[xo + (size_t)j * q + cand[i]];
        r_format_roundtrip(col.data(), nc, b.data(), z.data());
        for (int i = 0; i < nc; i++) { back[(size_t)i * p + j] = b[i]; zt[(size_t)i * p + j] = z[i]; }
    }
    int best = -1;
    double bestv = 0;
    std::vector<double> vals;
    std::string s;
    for (int i = 0; i < nc; i++) {
        vals.clear();
        s.assign((size_t)p, '0');
        for (int j = 0; j < p; j++) if (pat_[xo + (size_t)cand[i] * p + j]) s[(size_t)j] = '1';
        if (s != "0") vals.push_back(strtod(s.c_str(), nullptr));   // "0110" > "0" as strings; read back as 110
        for (int j = 0; j < p; j++)
            if (!zt[(size_t)i * p + j]) vals.push_back(back[(size_t)i * p + j]);
        double m = r_median(vals);
        if (std::isnan(m)) continue;                                 // which.max skips NA
        if (best < 0 || m > bestv) { best = i; bestv = m; }
    }
    if (best < 0) best = 0;   // the reference stops with an error here (zero-length replacement)
    return cand[best];
}

// hamming_correction for one observed pattern  :562-648
int32_t CrpIndex::pick_row(int block, const std::vector<uint8_t> &obs) const
{
    const int q = q_off_[block + 1] - q_off_[block], p = p_off_[block + 1] - p_off_[block];
    const size_t xo = (size_t)x_off_[block];
    std::vector<int> dist(q);
    int dmin = 1 << 30;
    for (int r = 0; r < q; r++) {
        int d = 0;
        const uint8_t *pr = &pat_[xo + (size_t)r * p];
        for (int j = 0; j < p; j++) d += (pr[j] != obs[(size_t)j]);
        dist[r] = d;
        dmin = std::min(dmin, d);
        if (d == 0) return r;                                        // :609-610
    }
    std::vector<int> cand;
    for (int r = 0; r < q; r++) if (dist[r] <= hd_) cand.push_back(r);
    if (cand.size() == 1) return cand[0];                            // :620-622
    if (cand.empty())                                                // :624-627
        for (int r = 0; r < q; r++) if (dist[r] <= dmin) cand.push_back(r);
    if (cand.size() == 1) return cand[0];
    return tie_break(block, cand);                                   // :613-618
}

int32_t CrpIndex::map_one(const int32_t *tx_in, int n_in, std::vector<int32_t> &tx) const
{
    tx.assign(tx_in, tx_in + n_in);
    std::sort(tx.begin(), tx.end());
    tx.erase(std::unique(tx.begin(), tx.end()), tx.end());
    const int n = (int)tx.size();
    if (n == 0) return -1;

    struct Cand { int32_t block; int32_t inter; uint8_t sup; };
    Cand small[64];
    std::vector<Cand> big;
    int nc = 0;
    auto add = [&](const Hit &h) {
        Cand *arr = big.empty() ? small : big.data();
        for (int i = 0; i < nc; i++)
            if (arr[i].block == h.block) { arr[i].inter++; arr[i].sup |= h.supported; return; }
        if (big.empty() && nc == 64) big.assign(small, small + 64);
        if (big.empty()) small[nc++] = Cand{h.block, 1, h.supported};
        else { big.push_back(Cand{h.block, 1, h.supported}); nc++; }
    };
    for (int i = 0; i < n; i++) {
        auto it = tx_range_.find(tx[(size_t)i]);
        if (it == tx_range_.end()) continue;                         // transcript in no block (t2c gives NA)
        for (int32_t h = it->second.first; h < it->second.second; h++) add(hits_[(size_t)h]);
    }
    const Cand *arr = big.empty() ? small : big.data();

    // 1x1 blocks (length(crp) == 1) take the classes made of exactly their transcript  :457-464
    if (n == 1)
        for (int i = 0; i < nc; i++) {
            int k = arr[i].block;
            if ((q_off_[k + 1] - q_off_[k]) * (p_off_[k + 1] - p_off_[k]) == 1) return q_off_[k];
        }

    // best block by Jaccard score against the block's full transcript list  :484-505
    int best = -1;
    double bestsc = -1;
    for (int i = 0; i < nc; i++) {
        if (!arr[i].sup) continue;                                   // :487-493
        int k = arr[i].block;
        int pk = p_off_[k + 1] - p_off_[k];
        double sc = (double)arr[i].inter / (double)(pk + n - arr[i].inter);      // :502
        if (best < 0 || sc > bestsc || (sc == bestsc && name_rank_[k] < name_rank_[best])) { best = k; bestsc = sc; }
    }
    if (best < 0) return -1;
    const int q = q_off_[best + 1] - q_off_[best], p = p_off_[best + 1] - p_off_[best];
    if (q * p == 1) return -1;      // best match is a 1x1 block: dropped from every other block (:506), n > 1 here

    std::vector<uint8_t> obs((size_t)p, 0);                          // raw3 pattern  :516-525
    for (int i = 0; i < n; i++) {
        auto it = tx_range_.find(tx[(size_t)i]);
        if (it == tx_range_.end()) continue;
        for (int32_t h = it->second.first; h < it->second.second; h++) {
            const Hit &hh = hits_[(size_t)h];
            if (hh.block == best && hh.supported) obs[(size_t)(hh.col - p_off_[best])] = 1;
        }
    }
    return q_off_[best] + pick_row(best, obs);
}

void CrpIndex::map(int64_t n_eq, const int64_t *eq_off, const int32_t *eq_tx, int32_t *eq_row, int n_threads) const
{
    if (n_threads < 1) n_threads = 1;
    n_threads = (int)std::min<int64_t>(n_threads, std::max<int64_t>(1, n_eq / 1024));
    auto work = [&](int64_t lo, int64_t hi) {
        std::vector<int32_t> scratch;
        for (int64_t e = lo; e < hi; e++)
            eq_row[e] = map_one(eq_tx + eq_off[e], (int)(eq_off[e + 1] - eq_off[e]), scratch);
    };
    if (n_threads == 1) { work(0, n_eq); return; }
    std::atomic<bool> failed{false};
    auto entry = [&](int64_t lo, int64_t hi) {       // thread entry: nothing may escape a std::thread
        try { work(lo, hi); } catch (const std::exception &) { failed = true; }
    };
    std::vector<std::thread> th;
    int64_t per = (n_eq + n_threads - 1) / n_threads;
    for (int t = 0; t < n_threads; t++) {
        int64_t lo = t * per, hi = std::min<int64_t>(n_eq, lo + per);
        if (lo >= hi) continue;
        bool started = false;
        try { th.emplace_back(entry, lo, hi); started = true; } catch (const std::exception &) {}
        if (!started) entry(lo, hi);                 // no more threads: do the range here
    }
    for (auto &t : th) t.join();
    if (failed) throw std::bad_alloc();
}

// ------------------------------------------------------------------------------------------
// Expansion of zero-suppressed result rows (see fetch_dense in scasa_cuda.cu).  The destination is
// written once and not read here, so all stores are non-temporal; the row is produced in aligned
// blocks: an empty block is one store of zeros, a block with values is assembled on the stack.
namespace {

inline void put1(double *p, double x)
{
    long long bits;
    memcpy(&bits, &x, 8);
    _mm_stream_si64((long long *)p, bits);
}

void expand_row_sse2(double *o, int64_t n_cols, const int32_t *c, const double *v, int nnz)
{
    int64_t at = 0;
    int k = 0;
    if (n_cols > 0 && ((uintptr_t)o & 15)) { double x = 0; if (k < nnz && c[k] == 0) x = v[k++]; put1(o, x); at = 1; }
    const __m128d z = _mm_setzero_pd();
    int64_t next = k < nnz ? c[k] : INT64_MAX;
    for (; at + 2 <= n_cols; at += 2) {
        if (next >= at + 2) { _mm_stream_pd(o + at, z); continue; }
        alignas(16) double t[2] = {0, 0};
        do { t[next - at] = v[k]; k++; next = k < nnz ? c[k] : INT64_MAX; } while (next < at + 2);
        _mm_stream_pd(o + at, _mm_load_pd(t));
    }
    for (; at < n_cols; at++) { double x = 0; if (k < nnz && c[k] == at) x = v[k++]; put1(o + at, x); }
}

__attribute__((target("avx2"))) void expand_row_avx2(double *o, int64_t n_cols, const int32_t *c, const double *v, int nnz)
{
    int64_t at = 0;
    int k = 0;
    while (at < n_cols && ((uintptr_t)(o + at) & 31)) { double x = 0; if (k < nnz && c[k] == at) x = v[k++]; put1(o + at, x); at++; }
    const __m256d z = _mm256_setzero_pd();
    int64_t next = k < nnz ? c[k] : INT64_MAX;
    for (; at + 4 <= n_cols; at += 4) {
        if (next >= at + 4) { _mm256_stream_pd(o + at, z); continue; }
        alignas(32) double t[4] = {0, 0, 0, 0};
        do { t[next - at] = v[k]; k++; next = k < nnz ? c[k] : INT64_MAX; } while (next < at + 4);
        _mm256_stream_pd(o + at, _mm256_load_pd(t));
    }
    for (; at < n_cols; at++) { double x = 0; if (k < nnz && c[k] == at) x = v[k++]; put1(o + at, x); }
}

__attribute__((target("avx512f"))) void expand_row_avx512(double *o, int64_t n_cols, const int32_t *c, const double *v, int nnz)
{
    int64_t at = 0;
    int k = 0;
    while (at < n_cols && ((uintptr_t)(o + at) & 63)) { double x = 0; if (k < nnz && c[k] == at) x = v[k++]; put1(o + at, x); at++; }
    const __m512d z = _mm512_setzero_pd();
    int64_t next = k < nnz ? c[k] : INT64_MAX;
    for (; at + 8 <= n_cols; at += 8) {
        if (next >= at + 8) { _mm512_stream_pd(o + at, z); continue; }
        alignas(64) double t[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        do { t[next - at] = v[k]; k++; next = k < nnz ? c[k] : INT64_MAX; } while (next < at + 8);
        _mm512_stream_pd(o + at, _mm512_load_pd(t));
    }
    for (; at < n_cols; at++) { double x = 0; if (k < nnz && c[k] == at) x = v[k++]; put1(o + at, x); }
}

using ExpandRow = void (*)(double *, int64_t, const int32_t *, const double *, int);
ExpandRow pick_expand()
{
    if (const char *env = getenv("SCASA_EXPAND")) {          // testing: force a code path
        if (!strcmp(env, "sse2")) return expand_row_sse2;
        if (!strcmp(env, "avx2") && __builtin_cpu_supports("avx2")) return expand_row_avx2;
    }
    if (__builtin_cpu_supports("avx512f")) return expand_row_avx512;
    if (__builtin_cpu_supports("avx2")) return expand_row_avx2;
    return expand_row_sse2;
}

}  // namespace

void expand_rows(double *out, int64_t n_cols, int64_t r0, int64_t r1, const int64_t *start, const int32_t *nnz,
                 const int32_t *cols, const double *vals)
{
    static const ExpandRow fn = pick_expand();
    for (int64_t r = r0; r < r1; r++) fn(out + r * n_cols, n_cols, cols + start[r], vals + start[r], nnz[r]);
    _mm_sfence();
}

// ---- N3: is DM0 what ccrpfun(CRP[[k]], clim) stops at? (Scasa_Rsource.R:168-193; scasa_quant_step2_v1.0.0.R:127) ----
// Singular values of a small dense block (q x p, column-major) from the eigenvalues of its Gram matrix by
// cyclic Jacobi rotations.  Only the ratios max(sv)/sv against clim (30) matter, for which the squared
// conditioning of the Gram route is harmless.
static void singular_values(const double *A, int q, int p, std::vector<double> &sv)
{
    std::vector<double> G((size_t)p * p);
    for (int i = 0; i < p; i++)
        for (int j = i; j < p; j++) {
            long double s = 0;
            for (int r = 0; r < q; r++) s += (long double)A[(size_t)i * q + r] * A[(size_t)j * q + r];
            G[(size_t)i * p + j] = G[(size_t)j * p + i] = (double)s;
        }
    for (int sweep = 0; sweep < 60; sweep++) {
        double off = 0;
        for (int i = 0; i < p; i++) for (int j = i + 1; j < p; j++) off += G[(size_t)i * p + j] * G[(size_t)i * p + j];
        if (off < 1e-30) break;
        for (int i = 0; i < p; i++)
            for (int j = i + 1; j < p; j++) {
                const double g = G[(size_t)i * p + j];
                if (std::fabs(g) < 1e-300) continue;
                const double theta = (G[(size_t)j * p + j] - G[(size_t)i * p + i]) / (2.0 * g);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
                const double c = 1.0 / std::sqrt(t * t + 1.0), sn = t * c;
                for (int k = 0; k < p; k++) {
                    const double a = G[(size_t)k * p + i], b = G[(size_t)k * p + j];
                    G[(size_t)k * p + i] = c * a - sn * b; G[(size_t)k * p + j] = sn * a + c * b;
                }
                for (int k = 0; k < p; k++) {
                    const double a = G[(size_t)i * p + k], b = G[(size_t)j * p + k];
                    G[(size_t)i * p + k] = c * a - sn * b; G[(size_t)j * p + k] = sn * a + c * b;
                }
            }
    }
    sv.resize((size_t)p);
    for (int i = 0; i < p; i++) sv[(size_t)i] = std::sqrt(std::max(0.0, G[(size_t)i * p + i]));
}

static bool all_separated(const double *A, int q, int p, double clim)   // nsc == ncol(newx), Rsource:175-179
{
    if (p < 1) return true;
    std::vector<double> sv;
    singular_values(A, q, p, sv);
    const double mx = *std::max_element(sv.begin(), sv.end());
    for (double v : sv) if (!(std::fabs(mx / v) < clim)) return false;
    return true;
}

// For every block: (i) the CRP columns without a member are exactly the zero-sum ones (:171-172); (ii) every DM0
// column is the mean of its member CRP columns — a k-means centre (:186-187); (iii) all condition numbers of DM0
// are < clim, so the repeat loop stops at it (:175-179); (iv) where DM0 merged columns, the unmerged block was NOT
// separated (otherwise ccrpfun would have returned it unmerged).  What cannot be checked without R's RNG is that
// k-means would pick this very partition among those that satisfy (ii)-(iii).  Returns the first block that
// fails, or -1.
int64_t dm0_fixed_point_check(int n_blocks, const int32_t *q_off, const int32_t *crp_p_off, const int64_t *crp_x_off,
                              const double *crp_x, const int32_t *dm0_p_off, const int64_t *dm0_x_off, const double *dm0_x,
                              const int32_t *crp_member, double clim, int n_threads)
{
    std::atomic<int64_t> bad{INT64_MAX};
    std::atomic<int> next{0};
    auto work = [&]() {
        std::vector<double> mean, sup;
        for (;;) {
            const int k0 = next.fetch_add(256);
            if (k0 >= n_blocks) return;
            for (int k = k0; k < std::min(n_blocks, k0 + 256); k++) {
                const int q = q_off[k + 1] - q_off[k], pc = crp_p_off[k + 1] - crp_p_off[k], pd = dm0_p_off[k + 1] - dm0_p_off[k];
                const double *C = crp_x + crp_x_off[k], *D = dm0_x + dm0_x_off[k];
                const int32_t *mem = crp_member + crp_p_off[k];
                bool ok = true;
                int n_sup = 0;
                for (int c = 0; c < pc && ok; c++) {
                    long double s = 0;
                    for (int r = 0; r < q; r++) s += C[(size_t)c * q + r];
                    const bool supported = (double)s > 0;
                    n_sup += supported;
                    if (supported != (mem[c] >= 0)) ok = false;
                    if (mem[c] >= 0 && (mem[c] < dm0_p_off[k] || mem[c] >= dm0_p_off[k + 1])) ok = false;
                }
                for (int j = 0; j < pd && ok; j++) {
                    mean.assign((size_t)q, 0.0);
                    int cnt = 0;
                    for (int c = 0; c < pc; c++)
                        if (mem[c] == dm0_p_off[k] + j) { cnt++; for (int r = 0; r < q; r++) mean[(size_t)r] += C[(size_t)c * q + r]; }
                    if (!cnt) { ok = false; break; }
                    for (int r = 0; r < q; r++)
                        if (std::fabs(mean[(size_t)r] / cnt - D[(size_t)j * q + r]) > 1e-12) { ok = false; break; }
                }
                if (ok && pd >= 1 && !all_separated(D, q, pd, clim)) ok = false;
                if (ok && pd < n_sup) {                         // merged: the supported columns themselves must not be separated
                    sup.clear();
                    for (int c = 0; c < pc; c++)
                        if (mem[c] >= 0) sup.insert(sup.end(), C + (size_t)c * q, C + (size_t)(c + 1) * q);
                    if (all_separated(sup.data(), q, n_sup, clim)) ok = false;
                }
                if (!ok) { int64_t cur = bad.load(); while (k < cur && !bad.compare_exchange_weak(cur, k)) {} }
            }
        }
    };
    int nthr = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
    nthr = std::max(1, std::min(nthr, 64));
    std::vector<std::thread> th;
    try { for (int t = 1; t < nthr; t++) th.emplace_back(work); } catch (const std::exception &) {}
    work();
    for (auto &t : th) t.join();
    return bad.load() == INT64_MAX ? -1 : bad.load();
}

}  // namespace scasa
