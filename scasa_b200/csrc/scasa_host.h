// scasa_host.h — host-side index for crpcount_td's eqclass -> (block, row) decision.
//
// crpcount_td (Scasa_Rsource.R:416-557) decides, for every equivalence class of every cell,
// which X-matrix block and which row of it receives the class' UMI count.  The decision reads
// only the class' transcript set, never the cell, so it is taken once per distinct eqclass
// (alevin eqclass ids are global in bfh.txt, scasa_quant_step1_1_v1.0.0.R:60) and the per-cell
// work becomes an integer scatter-add on the GPU (pack_cells_kernel).
#pragma once
#include <cstdint>
#include <string>
#include <unordered_map>
#include <vector>

namespace scasa {

class CrpIndex {
public:
    void build(int n_blocks, const int32_t *q_off, const int32_t *p_off, const int64_t *x_off, const double *x,
               const int32_t *col_tx, const uint8_t *row_pat, const int32_t *name_rank, int hamming_dist);
    bool ready() const { return B_ > 0; }
    // eq_row[e] = global row or -1
    void map(int64_t n_eq, const int64_t *eq_off, const int32_t *eq_tx, int32_t *eq_row, int n_threads) const;
    int32_t map_one(const int32_t *tx, int n, std::vector<int32_t> &scratch) const;

private:
    struct Hit { int32_t block, col; uint8_t supported; };
    int B_ = 0, hd_ = 1;
    std::vector<int32_t> q_off_, p_off_, col_tx_, name_rank_;
    std::vector<int64_t> x_off_;
    std::vector<double> x_;
    std::vector<uint8_t> pat_, supported_;            // per column: colSums(crp) > 0
    std::unordered_map<int32_t, std::pair<int32_t, int32_t>> tx_range_;  // tx -> [first, last) in hits_
    std::vector<Hit> hits_;                           // t2c: blocks (and columns) holding a transcript
    int32_t pick_row(int block, const std::vector<uint8_t> &obs) const;
    int32_t tie_break(int block, const std::vector<int> &cand) const;
};

// R's format() of a numeric column at 7 significant digits followed by as.numeric(): the value
// each element is read back as, and whether its text is exactly "0" (see SURVEY.md §9 Q1).
void r_format_roundtrip(const double *v, int n, double *back, uint8_t *is_zero_text);

// Rows [r0, r1) of a zero-suppressed batch -> dense rows of `out` (row stride n_cols).  Pairs are in
// ascending column order inside a row; every element of the rows is written exactly once with
// non-temporal stores (64-byte blocks with AVX-512, 32-byte with AVX2, else SSE2; chosen at run time).
void expand_rows(double *out, int64_t n_cols, int64_t r0, int64_t r1, const int64_t *start, const int32_t *nnz,
                 const int32_t *cols, const double *vals);

// N3: first block where DM0 is not what ccrpfun(CRP[[k]], clim) stops at, or -1 (see scasa_host.cpp)
int64_t dm0_fixed_point_check(int n_blocks, const int32_t *q_off, const int32_t *crp_p_off, const int64_t *crp_x_off,
                              const double *crp_x, const int32_t *dm0_p_off, const int64_t *dm0_x_off, const double *dm0_x,
                              const int32_t *crp_member, double clim, int n_threads);

// message returned by scasa_last_error(NULL) on the calling thread (for the host-only entry points)
void set_global_error(const std::string &m);

}  // namespace scasa
