// scasa_kernels.cuh — device side of libscasa_cuda (sm_100a).
//
// Data flow of one scasa_run() (everything resident in HBM):
//
//   [eqclass counts] --pack_cells_*_kernel--> Y as CSR by cell (rows ascending, duplicates summed)
//   Y --task_count_kernel--> entries / cells-with-counts per (chunk, EM block) "task"
//     --classify_kernel--> blob size, shared-memory need and work class per task; a scan gives
//                          the blob offsets
//   Y --struct_fill_kernel--> the result rows as CSR by cell (struct Result): the columns that can be
//        non-zero, the pass-through estimates (nrow == 1, Scasa_Rsource.R:979-983: the count itself) and
//        their column sums; the dense cells x transcripts matrix never exists on the device
//   Y --task_scatter_kernel--> one compact blob per task (counts by cell, rows, cell ids, result offsets)
//   blobs --em_task_kernel<W, R>--> the alternating EM (estim_chunk's loop, Rsource:1010-1042): W warps
//        stage one task in shared memory and run it to the end; persistent, atomic work queue; estimates
//        go straight into the CSR values, their per-chunk column sums into Result::colpart
//   result --colsum_em_reduce--> per-column sums (step2:167);  --gene_rows_kernel--> gene sums (step3:32-46)
//        as CSR by cell
//
// Why sparse: a cell with no counts in a block keeps beta = 0 through every iteration of
// estim.BETA and contributes 0 to every sum of estim.X.em, so only (cell, block) pairs with
// counts ("records") are stored.  The convergence tests stay what the reference has: maxima over
// the whole chunk (Rsource:755-756, :1019-1020).
// Poor 2-column blocks add adjust_esp to EVERY cell of the chunk (Rsource:1004); the cells
// without counts then all follow one identical trajectory, so they are stored once as a
// "pseudo cell" with a multiplicity (weights in colSums(BETA) and in the X update).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace scasa {

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;

// ------------------------------------------------------------------------------------------
// digamma(x), x > 0, fp64.  psi(x) = psi(x+9) - sum_{i=0..8} 1/(x+i) for x < 9, and the
// asymptotic series for arguments >= 9 (truncation < 4e-16 there).  The nine reciprocals are
// folded into one rational function of u = x(x+8): (x+i)(x+8-i) = u + i(8-i), all terms
// positive, so there is no cancellation for x -> 0 (psi(1e-8) = -1e8 - 0.577..).
// Replaces R's digamma() (nmath dpsifn) at Scasa_Rsource.R:733,738.
__device__ __forceinline__ double digamma_pos(double x)
{
    double xs = x;
    double rec = 0.0;
    if (x < 9.0) {
        double u = x * (x + 8.0);
        double a = u + 7.0, b = u + 12.0, c = u + 15.0;
        double bc = b * c;
        double D = u * a * bc;                          // prod_{k in {0,7,12,15}} (u+k)
        double N = a * bc + u * (bc + a * (b + c));     // dD/du
        double x4 = x + 4.0;
        // sum_i 1/(x+i) = (2x+8) N/D + 1/(x+4)
        rec = ((2.0 * x + 8.0) * N * x4 + D) / (D * x4);
        xs = x + 9.0;
    }
    double r = 1.0 / xs;
    double r2 = r * r;
    double s = r2 * (1.0 / 12.0 - r2 * (1.0 / 120.0 - r2 * (1.0 / 252.0 - r2 * (1.0 / 240.0 -
               r2 * (1.0 / 132.0 - r2 * (691.0 / 32760.0 - r2 * (1.0 / 12.0)))))));
    return log(xs) - 0.5 * r - s - rec;
}

// gamma0 of estim.BETA's fn1 (Rsource:737-738).  One call site per kernel keeps the code small
// (the first version inlined it PMAX times and stalled on instruction fetch).
__device__ __noinline__ double em_gamma(double beta, double lognorm, int modify)
{
    return modify ? exp(digamma_pos(beta + 0.00000001) - lognorm) : beta;
}

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}

// ------------------------------------------------------------------------------------------
struct XDev {                 // resident X-matrix (DM0) and its derived index maps
    int n_blocks, n_em;
    int64_t n_rows, n_cols;
    const int32_t *q_off, *p_off;
    const int64_t *x_off;
    const double *x;
    const int32_t *em_blocks; // [n_em]    block id
    const uint8_t *poor;      // [n_blocks] poorCondX (Rsource:957-971)
    const int32_t *em_order;  // [n_em]    EM indices, most expensive block shape first (work-queue order)
    const int2 *row_info;     // [n_rows]  {EM index of the row's block or -1, row index inside the block}
    const int4 *row_aux;      // [n_rows]  {first DM0 column of the row's block, poor blocks before it, p | poor << 16 of an EM block
                              //            (0: pass-through), EM index or -1}: all that the per-entry passes need in one load
    const int32_t *em_col0;   // [n_em+1]  columns of the EM blocks before this one (index into the per-chunk column sums)
    const int32_t *poor_rank; // [n_blocks] poor blocks with a smaller block id
    const int32_t *poor_col0; // [n_poor]  first DM0 column of the k-th poor block
    int n_poor, n_emcols;
};

// ---- results -------------------------------------------------------------------------------
// The reference's result is a dense cells x transcripts matrix that is ~94 % zeros.  Which entries can
// be non-zero is known as soon as Y is: a cell's pass-through counts, the p columns of every EM block
// the cell has counts in, and the two columns of every poor block (the +2 pseudo count of Rsource:1004
// reaches every cell of the chunk).  The result is therefore kept as CSR by cell over exactly this
// STRUCTURE, columns ascending; the dense matrix only ever exists on the host, when a caller asks for
// it.  The gene matrix (step3) is CSR by cell too, genes ascending, in rows of the same capacity.
struct Result {
    const int64_t *rs_ptr;    // [n_cells+1] first entry of each cell's row
    int32_t *rs_col;          // [nnz] DM0 column, ascending inside a row
    double *rs_val;           // [nnz]
    uint32_t *ypos;           // [nnz of Y] per Y entry: offset inside the cell's row of the entry's column (pass-through)
                              //            or of the first column of its block (head entry of a run; EM blocks)
    int32_t *poor_pos;        // [n_cells * n_poor] offset inside the cell's row of the k-th poor block's two columns
    double *colsum;           // [n_cols] pass-through columns: sum over cells (exact: the counts are integers)
    double *colpart;          // [n_chunks * n_emcols] EM columns: sum over the chunk's cells, in cell order (colsum_chunk_kernel)
    int32_t *g_col;           // gene CSR: row c occupies [rs_ptr[c], rs_ptr[c] + g_len[c])
    double *g_val;
    int32_t *g_len;           // [n_cells]
};

struct Sample {               // the cells [cell_base, cell_base + n_cells) = chunks [chunk_base, chunk_base + n_chunks)
    int64_t n_cells;          // of the staged sample; every per-cell / per-chunk pointer below is already offset to the range
    int n_chunks;
    int64_t cell_base;        // chunk_off and cell_chunk hold sample-wide values: subtract these
    int chunk_base;
    const int64_t *chunk_off;   // [n_chunks+1]
    const int32_t *cell_chunk;  // [n_cells]
    const int64_t *y_start;     // [n_cells]  first Y entry of the cell
    const int32_t *y_len;       // [n_cells]
    const int32_t *y_row;       // global row, ascending inside a cell
    const double *y_val;
};

// ---- task blob ----------------------------------------------------------------------------
// One (chunk, EM block) task with n records (cells with counts, in cell order; for a poor block
// the last record is the pseudo cell) and E entries (record-major, rows ascending inside a record;
// poor blocks store every row of every record).  The blob is what the EM kernel copies verbatim
// into the head of its shared-memory slot:
//     ey[E] f64 | rcell[n] u32 | rpos[n] u32 | eptr[n+1] u16 | erow[E] u16 | pad to 16
// rcell: cell index inside the chunk, or kPseudo | multiplicity.
// rpos:  where the record's p estimates go: offset of its first column in the result values, relative to
//        the first value of the chunk's first cell.
constexpr uint32_t kPseudo = 0x80000000u;
__host__ __device__ inline int align16(int x) { return (x + 15) & ~15; }
struct BlobLayout { int off_cell, off_pos, off_ptr, off_row, bytes; };
__host__ __device__ inline BlobLayout blob_layout(int n, int E)
{
    BlobLayout L;
    L.off_cell = 8 * E;
    L.off_pos = L.off_cell + 4 * n;
    L.off_ptr = L.off_pos + 4 * n;
    L.off_row = L.off_ptr + 2 * (n + 1);
    L.bytes = align16(L.off_row + 2 * E);
    return L;
}
// shared memory one task needs in the EM kernel: blob image, then
//     ew[E] Xc[qp] Xn[qp] rcs[p] bts[p] nrm[p] lnorm[n] beta[p*n] gam[p*n]  (f64, or f32)
//     erec[E] rptr[q+1] perm[E]                                     (u16)
// rb = bytes of the arithmetic type (8, or 4 with scasa_opts_t.fp32)
__host__ __device__ inline int slot_bytes_of(int q, int p, int n, int E, int rb)
{
    return blob_layout(n, E).bytes + align16(rb * (E + 2 * q * p + 3 * p + n + 2 * p * n)) + align16(2 * (2 * E + q + 1));
}

// work classes: W warps per task, a slot of at most kClassBytes[c] (the last class takes the rest)
constexpr int kClasses = 4;
// one more list for tiny tasks, run two per warp by em_pair_kernel: not poor, and records*columns,
// entries and rows*columns all <= pair_max (<= 16 = lanes of a half-warp; 0 turns the list off)
constexpr int kLists = kClasses + 1;
constexpr int kPairList = kClasses;
constexpr int kPairCounter = 6;
struct ClassCfg { int pairs[kClasses]; int bytes[kClasses]; int real_bytes; int pair_max; };

struct Tasks {                // task t = chunk * n_em + em index
    int32_t *cnt;             // [n_tasks] Y entries; after classify: stored entries E (dense for poor blocks)
    int32_t *runs;            // [n_tasks] cells with counts; after classify: records n (incl. the pseudo cell)
    int32_t *tbytes;          // [n_tasks] blob bytes
    int64_t *toff;            // [n_tasks+1] exclusive scan of tbytes
    char *tdata;              // blobs
    int4 *desc;               // [n_tasks] {block, q | p << 16, records n | entries E << 16, first X value of the block}: what an EM
                              //           group needs to know about a task, in one 16-byte load (non-empty tasks only)
    int32_t *list[kLists];    // task ids per work class (+ the tiny-task list), expensive block shapes first
    unsigned int *counters;   // [0..3] list lengths [4] non-empty tasks [5] largest slot need [6] tiny-task list length [7] error flag
};
// ------------------------------------------------------------------------------------------
// crpcount_td's accumulation (Scasa_Rsource.R:528-531, :643-646 and :461-462): an int32 histogram over the
// X-matrix rows in shared memory (hg38: 41,251 rows = 165 KB in one tile; larger X-matrices in several tiles):
// one CTA per cell adds the UMI counts of the cell's eqclasses into hist[row] and then emits the
// non-zero rows in ascending order.  Integer adds: the result is independent of their order.
// Thread i owns the contiguous rows [i*seg, (i+1)*seg), seg odd (no bank conflicts); the histogram
// is left all-zero for the next cell.
constexpr int kPackEntries = 4096;                                // entries of a cell in registers per pass, over the whole CTA
template <int TN>                                                 // threads per CTA
__global__ void __launch_bounds__(TN) pack_cells_dense_kernel(int64_t n_cells, const int64_t *__restrict__ cell_ptr,
                                                              const int32_t *__restrict__ eq_id,
                                                              const int32_t *__restrict__ eq_count,
                                                              const int32_t *__restrict__ eq_row, int64_t n_eq, int n_rows,
                                                              int n_words, int32_t *__restrict__ y_row,
                                                              double *__restrict__ y_val, int32_t *__restrict__ y_len)
{
    constexpr int kPackBatch = kPackEntries / TN;
    extern __shared__ int pack_sm[];
    int *hist = pack_sm;                                          // [n_words * 32] counts by row of the tile
    unsigned *bits = (unsigned *)(pack_sm + (size_t)n_words * 32);   // [n_words] rows touched by this cell
    __shared__ int wsum[TN / 32];
    const int words_per_thread = (n_words + TN - 1) / TN;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int tile_rows = n_words * 32;
    // X-matrices with more rows than the histogram holds are taken in tiles of tile_rows rows: the cell's entries
    // are read once per tile (from L2 after the first), each tile emits its rows behind the previous one's
    const int n_tiles = (n_rows + tile_rows - 1) / tile_rows;
    for (int i = tid; i < n_words * 33; i += blockDim.x) pack_sm[i] = 0;
    __syncthreads();
    // Entries are taken kPackBatch per thread at a time: all their loads are issued before the first
    // gather through eq_row, all gathers before the first atomic, so a batch costs one memory round
    // trip per level instead of one per entry.  The next cell's first batch is loaded while this
    // cell's rows are emitted.
    constexpr int T = TN;
    int e[kPackBatch], c[kPackBatch], r[kPackBatch];
    auto load_entries = [&](int64_t s, int n, int base) {
#pragma unroll
        for (int k = 0; k < kPackBatch; k++) {
            const int i = base + k * T + tid;
            e[k] = i < n ? eq_id[s + i] : -1;
            c[k] = i < n ? eq_count[s + i] : 0;
        }
    };
    auto load_rows = [&]() {
#pragma unroll
        for (int k = 0; k < kPackBatch; k++) r[k] = (e[k] >= 0 && e[k] < n_eq) ? eq_row[e[k]] : -1;
    };
    int64_t cell = blockIdx.x;
    int64_t s = 0;
    int n = 0;
    if (cell < n_cells) { s = cell_ptr[cell]; n = (int)(cell_ptr[cell + 1] - s); }
    load_entries(s, n, 0);
    load_rows();
    for (; cell < n_cells; cell += gridDim.x) {
        const int64_t cell2 = cell + gridDim.x;
        int64_t s2 = 0;
        int n2 = 0;
        if (cell2 < n_cells) { s2 = cell_ptr[cell2]; n2 = (int)(cell_ptr[cell2 + 1] - s2); }
        int emitted = 0;
        for (int tile = 0; tile < n_tiles; tile++) {
            const int r0 = tile * tile_rows, r1 = min(n_rows, r0 + tile_rows);
            for (int base = 0; base < n; base += kPackBatch * T) {
                if (base > 0 || tile > 0) { load_entries(s, n, base); load_rows(); }
#pragma unroll
                for (int k = 0; k < kPackBatch; k++)
                    if (r[k] >= r0 && r[k] < r1 && c[k] > 0) {                    // UMI counts are >= 1
                        atomicAdd(&hist[r[k] - r0], c[k]);
                        atomicOr(&bits[(r[k] - r0) >> 5], 1u << ((r[k] - r0) & 31));
                    }
            }
            __syncthreads();
            if (tile == n_tiles - 1) load_entries(s2, n2, 0);                     // in flight during the emission
            // thread owns words [tid*wpt, (tid+1)*wpt): rows in ascending order across threads
            const int w0 = tid * words_per_thread, w1 = min(n_words, w0 + words_per_thread);
            int cnt = 0;
            for (int k = w0; k < w1; k++) cnt += __popc(bits[k]);
            int inc = cnt;
            for (int o = 1; o < 32; o <<= 1) { int v = __shfl_up_sync(kFull, inc, o); if (lane >= o) inc += v; }
            if (lane == 31) wsum[w] = inc;
            __syncthreads();
            int base = 0, total = 0;
#pragma unroll
            for (int k = 0; k < TN / 32; k++) { int v = wsum[k]; if (k < w) base += v; total += v; }
            int o = emitted + base + inc - cnt;
            if (cnt)
                for (int k = 0; k < w1 - w0; k++) {
                    unsigned m = bits[w0 + k];
                    bits[w0 + k] = 0;
                    while (m) {
                        const int rr = ((w0 + k) << 5) + __ffs(m) - 1;
                        m &= m - 1;
                        const int v = hist[rr];
                        hist[rr] = 0;
                        y_row[s + o] = rr + r0; y_val[s + o] = (double)v; o++;
                    }
                }
            emitted += total;
            if (tile == n_tiles - 1) load_rows();
            __syncthreads();
        }
        if (tid == 0) y_len[cell] = emitted;
        s = s2; n = n2;
    }
}

// ------------------------------------------------------------------------------------------
// General path (any number of rows): per cell, add the UMI counts of all eqclasses that map to the
// same row.  One CTA per cell: (row, count) pairs
// are bitonic-sorted by row in shared memory, equal rows summed, output rows ascending.
// Integer counts => the result does not depend on the order of the additions.
__global__ void pack_cells_kernel(int64_t n_cells, const int64_t *__restrict__ cell_ptr,
                                  const int32_t *__restrict__ eq_id, const int32_t *__restrict__ eq_count,
                                  const int32_t *__restrict__ eq_row, int64_t n_eq, int cap,
                                  int32_t *__restrict__ y_row, double *__restrict__ y_val,
                                  int32_t *__restrict__ y_len, int *__restrict__ err_flag)
{
    extern __shared__ unsigned long long sm_keys[];   // (row << 32) | count, cap entries (power of two)
    for (int64_t cell = blockIdx.x; cell < n_cells; cell += gridDim.x) {
        const int64_t s = cell_ptr[cell];
        const int n = (int)(cell_ptr[cell + 1] - s);
        if (n > cap) { if (threadIdx.x == 0) { atomicExch(err_flag, 1); y_len[cell] = 0; } continue; }
        int m = 1;
        while (m < n) m <<= 1;
        for (int i = threadIdx.x; i < m; i += blockDim.x) {
            unsigned long long k = 0xffffffffffffffffull;
            if (i < n) {
                int e = eq_id[s + i];
                int r = (e >= 0 && e < n_eq) ? eq_row[e] : -1;
                int c = eq_count[s + i];
                if (r >= 0 && c > 0) k = ((unsigned long long)(unsigned)r << 32) | (unsigned)c;   // UMI counts are >= 1
            }
            sm_keys[i] = k;
        }
        __syncthreads();
        for (int k = 2; k <= m; k <<= 1)
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int i = threadIdx.x; i < m; i += blockDim.x) {
                    int l = i ^ j;
                    if (l > i) {
                        unsigned long long a = sm_keys[i], b = sm_keys[l];
                        bool up = ((i & k) == 0);
                        if ((a > b) == up) { sm_keys[i] = b; sm_keys[l] = a; }
                    }
                }
                __syncthreads();
            }
        // heads of equal-row runs -> output slot = number of heads before (block-wide, in order)
        int base = 0;
        for (int i0 = 0; i0 < m; i0 += blockDim.x) {
            int i = i0 + threadIdx.x;
            bool valid = i < m && (sm_keys[i] >> 32) != 0xffffffffull;
            bool head = valid && (i == 0 || (sm_keys[i - 1] >> 32) != (sm_keys[i] >> 32));
            // in-order rank of the heads of this slab
            unsigned bal = __ballot_sync(kFull, head);
            __shared__ int warp_heads[32];
            int w = threadIdx.x >> 5, ln = threadIdx.x & 31;
            if (ln == 0) warp_heads[w] = __popc(bal);
            __syncthreads();
            int before = 0, tot = 0;
            for (int k = 0; k < (int)(blockDim.x >> 5); k++) {
                int v = warp_heads[k];
                if (k < w) before += v;
                tot += v;
            }
            if (head) {
                int slot = base + before + __popc(bal & ((1u << ln) - 1));
                unsigned long long kk = sm_keys[i];
                unsigned row = (unsigned)(kk >> 32);
                long long sum = 0;
                for (int t = i; t < m && (unsigned)(sm_keys[t] >> 32) == row; t++)
                    sum += (int)(unsigned)(sm_keys[t] & 0xffffffffull);
                y_row[s + slot] = (int)row;
                y_val[s + slot] = (double)sum;
            }
            base += tot;
            __syncthreads();
        }
        if (threadIdx.x == 0) y_len[cell] = base;
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------
// Task extents.  One warp per cell: count this cell's entries and runs per (chunk, EM block), and the size of
// the cell's result row (pass-through entries + p per non-poor EM block with counts + 2 per poor block).
__global__ void task_count_kernel(XDev X, Sample S, Tasks T, int32_t *__restrict__ rs_len)
{
    int64_t cell = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (cell >= S.n_cells) return;
    const int lane = threadIdx.x & 31;
    const int64_t s = S.y_start[cell];
    const int n = S.y_len[cell];
    const int64_t tbase = (int64_t)(S.cell_chunk[cell] - S.chunk_base) * X.n_em;
    int w = 0, carry_e = -2;
    int r_nx = lane < n ? S.y_row[s + lane] : -1;
    for (int i0 = 0; i0 < n; i0 += kWarp) {
        const int i = i0 + lane, r = r_nx;
        if (i + kWarp < n) r_nx = S.y_row[s + i + kWarp];      // the next pass' rows are in flight during this one
        int e = -2, pi = 0;
        if (i < n) { const int4 a = X.row_aux[r]; e = a.w; pi = a.z; }
        int prev_e = __shfl_up_sync(kFull, e, 1);
        if (lane == 0) prev_e = carry_e;
        carry_e = __shfl_sync(kFull, e, min(31, n - 1 - i0));
        if (i >= n) continue;
        if (e < 0) { w++; continue; }
        atomicAdd(&T.cnt[tbase + e], 1);
        if (e != prev_e) {
            atomicAdd(&T.runs[tbase + e], 1);
            if (!(pi >> 16)) w += pi & 0xffff;
        }
    }
    for (int o = 16; o > 0; o >>= 1) w += __shfl_xor_sync(kFull, w, o);
    if (lane == 0) rs_len[cell] = w + 2 * X.n_poor;
}

// The structure of the result rows (see struct Result): columns of every entry a cell's row can hold, the
// pass-through estimates themselves (nrow(X0) == 1: the count is the estimate, Scasa_Rsource.R:979-983) and
// their column sums, zeros everywhere else (the EM kernels overwrite their columns).  One warp per cell
// walks the cell's Y entries in row order = block order = column order; the poor blocks' columns, which
// every cell carries, are merged in by the lane whose entry is the first behind them.
__global__ void struct_fill_kernel(XDev X, Sample S, Result R)
{
    int64_t cell = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (cell >= S.n_cells) return;
    const int lane = threadIdx.x & 31;
    const int64_t s = S.y_start[cell];
    const int n = S.y_len[cell];
    const int64_t base = R.rs_ptr[cell];
    int32_t *ppos = R.poor_pos + cell * X.n_poor;
    auto put_poor = [&](int k, int pos) {
        const int c0 = X.poor_col0[k];
        ppos[k] = pos;
        R.rs_col[base + pos] = c0; R.rs_col[base + pos + 1] = c0 + 1;
        R.rs_val[base + pos] = 0.0; R.rs_val[base + pos + 1] = 0.0;
    };
    int running = 0;           // entries of non-poor blocks placed so far
    int carry_e = -2, carry_pb = 0;
    // two-deep software pipeline: this pass' row records and the rows of the pass after it are in flight
    // while the previous pass is placed (a pass = 32 entries; the chain row -> record -> position would
    // otherwise be paid once per pass)
    int r_nx = -1;
    double y_cur = 0.0, y_nx = 0.0;
    int4 a_cur = make_int4(0, 0, 0, -2);
    if (lane < n) { a_cur = X.row_aux[S.y_row[s + lane]]; y_cur = S.y_val[s + lane]; }
    if (lane + kWarp < n) { r_nx = S.y_row[s + lane + kWarp]; y_nx = S.y_val[s + lane + kWarp]; }
    for (int i0 = 0; i0 < n; i0 += kWarp) {
        const int i = i0 + lane;
        const bool valid = i < n;
        const int4 a = a_cur;
        const double y = y_cur;
        if (i + kWarp < n) { a_cur = X.row_aux[r_nx]; y_cur = y_nx; }
        if (i + 2 * kWarp < n) { r_nx = S.y_row[s + i + 2 * kWarp]; y_nx = S.y_val[s + i + 2 * kWarp]; }
        int e = -2, r_col0 = 0, pb = 0, pi = 0;
        if (valid) { r_col0 = a.x; pb = a.y; pi = a.z; e = a.w; }
        int prev_e = __shfl_up_sync(kFull, e, 1), lo = __shfl_up_sync(kFull, pb, 1);
        if (lane == 0) { prev_e = carry_e; lo = carry_pb; }
        const bool head = valid && (e < 0 || e != prev_e);
        const int p = pi & 0xffff;
        const bool poorb = (pi >> 16) != 0;
        const int w = !valid ? 0 : (e < 0 ? 1 : (head && !poorb ? p : 0));
        int inc = w;
        for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(kFull, inc, o); if (lane >= o) inc += v; }
        const int pos_np = running + inc - w;
        if (valid) {
            for (int k = lo; k < pb; k++) put_poor(k, pos_np + 2 * k);        // poor blocks between the previous entry's block and this one
            const int pos = pos_np + 2 * pb;
            R.ypos[s + i] = (uint32_t)pos;
            if (e < 0) {
                R.rs_col[base + pos] = r_col0; R.rs_val[base + pos] = y;
                atomicAdd(&R.colsum[r_col0], y);                              // integer-valued counts: exact in any order
            } else if (head && !poorb) {
                for (int j = 0; j < p; j++) { R.rs_col[base + pos + j] = r_col0 + j; R.rs_val[base + pos + j] = 0.0; }
            }
        }
        const int last = min(31, n - 1 - i0);
        running += __shfl_sync(kFull, inc, 31);
        carry_e = __shfl_sync(kFull, e, last);
        carry_pb = __shfl_sync(kFull, pb, last);
    }
    for (int k = carry_pb + lane; k < X.n_poor; k += kWarp) put_poor(k, running + 2 * k);   // poor blocks behind the last entry
}

struct EmOpts {
    int maxiter_x, maxiter_bt, modify, fixed_iter;
    double maxerr, lim, adjust_esp;     // maxerr / lim: the outer test (Rsource:1019-1020)
    double maxerr_bt, lim_bt;           // estim.BETA's own (Rsource:726-727; not forwarded by estim_chunk)
};

// Per task: records, stored entries, blob size, work class.  Threads enumerate tasks block-major
// in X.em_order (expensive block shapes first), so the work queues start with the long tasks.
// All-zero tasks finish here: the reference runs one inner and one outer iteration on them and
// returns zeros (the result buffer is pre-zeroed).
__global__ void classify_kernel(XDev X, Sample S, Tasks T, int64_t n_tasks, EmOpts O, ClassCfg C, int32_t *iters,
                                unsigned long long *stats)
{
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    // what this thread adds to the run statistics / which list its task joins (warp-aggregated below)
    unsigned long long a_in = 0, a_out = 0, a_bin = 0, a_bout = 0;
    int cls = -1;
    unsigned need = 0;
    int64_t t = 0;
    if (idx < n_tasks) {
        const int e = X.em_order[(int)(idx / S.n_chunks)], h = (int)(idx % S.n_chunks);
        t = (int64_t)h * X.n_em + e;
        const int b = X.em_blocks[e];
        const int q = X.q_off[b + 1] - X.q_off[b], p = X.p_off[b + 1] - X.p_off[b];
        const int nc = (int)(S.chunk_off[h + 1] - S.chunk_off[h]);
        const int c = T.cnt[t];
        if (c == 0) {
            T.tbytes[t] = 0;
            const int o = O.fixed_iter ? O.maxiter_x : 1;
            const int in = O.fixed_iter ? O.maxiter_x * O.maxiter_bt : 1;
            if (iters) { iters[2 * t] = o; iters[2 * t + 1] = in; }
            a_in = (unsigned long long)in; a_out = (unsigned long long)o;
            a_bin = (unsigned long long)in * (q + 2 * p) * nc; a_bout = (unsigned long long)o * (q + p) * nc;
        } else {
            int n = T.runs[t], E = c;
            if (X.poor[b]) { n += 1; E = n * q; }                 // dense records + the pseudo cell
            if (X.poor[b] && q > 64) atomicExch(&T.counters[7], 1u);   // adjID is kept as a 64-bit mask
            if (E > 65535 || n > 65535) { atomicExch(&T.counters[7], 2u); E = 0; n = 0; }   // u16 indices in the blob
            T.runs[t] = n;
            T.cnt[t] = E;
            T.tbytes[t] = blob_layout(n, E).bytes;
            T.desc[t] = make_int4(b, q | (p << 16), n | (E << 16), (int)X.x_off[b]);
            if (n > 0) {
                need = (unsigned)slot_bytes_of(q, p, n, E, C.real_bytes);
                cls = kClasses - 1;
                for (int k = kClasses - 2; k >= 0; k--)
                    if (n * p <= C.pairs[k] && (int)need <= C.bytes[k]) cls = k;
                if (!X.poor[b] && n * p <= C.pair_max && E <= C.pair_max && q * p <= C.pair_max) cls = kPairList;
            }
        }
    }
    // one atomic per warp and counter instead of one per task
    for (int o = 16; o > 0; o >>= 1) {
        a_in += __shfl_xor_sync(kFull, a_in, o); a_out += __shfl_xor_sync(kFull, a_out, o);
        a_bin += __shfl_xor_sync(kFull, a_bin, o); a_bout += __shfl_xor_sync(kFull, a_bout, o);
        need = max(need, __shfl_xor_sync(kFull, need, o));
    }
    if (lane == 0) {
        if (a_in) { atomicAdd(&stats[0], a_in); atomicAdd(&stats[1], a_out); atomicAdd(&stats[2], a_bin); atomicAdd(&stats[3], a_bout); }
        if (need) atomicMax(&T.counters[5], need);
    }
    const unsigned nonempty = __ballot_sync(kFull, cls >= 0);
    if (lane == 0 && nonempty) atomicAdd(&T.counters[4], (unsigned)__popc(nonempty));
    for (int k = 0; k < kLists; k++) {
        const unsigned m = __ballot_sync(kFull, cls == k);
        if (!m) continue;
        unsigned base = 0;
        if (lane == __ffs(m) - 1) base = atomicAdd(&T.counters[k == kPairList ? kPairCounter : k], (unsigned)__popc(m));
        base = __shfl_sync(kFull, base, __ffs(m) - 1);
        if (cls == k) T.list[k][base + __popc(m & ((1u << lane) - 1u))] = (int32_t)t;
    }
}

// exclusive scan int32 -> int64, three phases (block sums, scan of sums, add back)
constexpr int kScanThreads = 1024;
constexpr int kScanItems = 4;
__global__ void scan_block_sums(const int32_t *__restrict__ in, int64_t n, int64_t *__restrict__ sums)
{
    __shared__ long long sh[32];
    int64_t base = (int64_t)blockIdx.x * kScanThreads * kScanItems;
    long long v = 0;
    for (int k = 0; k < kScanItems; k++) {
        int64_t i = base + (int64_t)threadIdx.x * kScanItems + k;
        if (i < n) v += in[i];
    }
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        long long w = sh[threadIdx.x];
        for (int o = 16; o > 0; o >>= 1) w += __shfl_xor_sync(kFull, w, o);
        if (threadIdx.x == 0) sums[blockIdx.x] = w;
    }
}
__global__ void scan_sums(int64_t *sums, int64_t nb, int64_t *total)
{
    // single CTA, sequential over slabs of 1024
    __shared__ long long sh[kScanThreads];
    __shared__ long long carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int64_t i0 = 0; i0 < nb; i0 += kScanThreads) {
        int64_t i = i0 + threadIdx.x;
        long long v = i < nb ? sums[i] : 0;
        sh[threadIdx.x] = v;
        __syncthreads();
        for (int o = 1; o < kScanThreads; o <<= 1) {
            long long a = threadIdx.x >= o ? sh[threadIdx.x - o] : 0;
            __syncthreads();
            sh[threadIdx.x] += a;
            __syncthreads();
        }
        if (i < nb) sums[i] = carry + sh[threadIdx.x] - v;
        __syncthreads();
        if (threadIdx.x == 0) carry += sh[kScanThreads - 1];
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}
__global__ void scan_add_back(const int32_t *__restrict__ in, int64_t n, const int64_t *__restrict__ sums,
                              const int64_t *__restrict__ total, int64_t *__restrict__ out)
{
    __shared__ long long sh[kScanThreads];
    int64_t base = (int64_t)blockIdx.x * kScanThreads * kScanItems;
    long long loc[kScanItems];
    long long v = 0;
    for (int k = 0; k < kScanItems; k++) {
        int64_t i = base + (int64_t)threadIdx.x * kScanItems + k;
        loc[k] = v;
        if (i < n) v += in[i];
    }
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < kScanThreads; o <<= 1) {
        long long a = threadIdx.x >= o ? sh[threadIdx.x - o] : 0;
        __syncthreads();
        sh[threadIdx.x] += a;
        __syncthreads();
    }
    long long excl = sh[threadIdx.x] - v + sums[blockIdx.x];
    for (int k = 0; k < kScanItems; k++) {
        int64_t i = base + (int64_t)threadIdx.x * kScanItems + k;
        if (i < n) out[i] = excl + loc[k];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) out[n] = *total;
}

// ------------------------------------------------------------------------------------------
// Scatter Y into the task blobs.  One CTA per chunk; cells are visited in order so the records
// of a task are sorted by cell (and their entries by row) without atomics — the order of every
// later floating-point sum is therefore fixed (same chunk => same bits on any GPU count).
// The chunk's per-task cursors, extents and blob offsets live in shared memory (st[]: 3 words per
// EM block), or in a global scratch area when the X-matrix has too many EM blocks for that.
constexpr int kScatterThreads = 512;
constexpr int kScatterBatch = 8;                               // entries per thread of a staged cell
constexpr int kScatterCap = kScatterThreads * kScatterBatch;   // cells with more entries take the unstaged path
template <bool staged>
__global__ void __launch_bounds__(kScatterThreads) task_scatter_kernel(XDev X, Sample S, Tasks T, Result R, uint32_t *scratch)
{
    extern __shared__ uint32_t scatter_sm[];
    const int h = blockIdx.x;
    const int64_t c0 = S.chunk_off[h] - S.cell_base, c1 = S.chunk_off[h + 1] - S.cell_base;
    const int nc = (int)(c1 - c0);
    const int64_t tbase = (int64_t)h * X.n_em;
    uint32_t *st = scratch ? scratch + (size_t)h * 3 * X.n_em : scatter_sm;
    uint32_t *cur = st;                  // records placed << 16 | entries placed
    uint32_t *ext = st + X.n_em;         // records << 16 | entries of the task
    uint32_t *boff = st + 2 * X.n_em;    // blob offset inside the chunk's blob area, in 16-byte units
    const int64_t chunk_base = T.toff[tbase];
    char *cdata = T.tdata + chunk_base;
    // staged path (cursors in shared memory and room for it): per cell, the block index and block row of
    // every entry go to shared memory once, so the run walks below touch no global memory; the
    // loads are batched (all rows, then all row_info gathers) and the next cell's are in flight
    // while this cell is placed.  A cell costs two memory round trips instead of ~16 dependent ones.
    int32_t *se = (int32_t *)(scatter_sm + 3 * (size_t)X.n_em);   // [kScatterCap] EM index of the entry's block, or -1
    uint16_t *srl = (uint16_t *)(se + kScatterCap);                // [kScatterCap] row inside the block
    uint16_t *einfo = srl + kScatterCap;                           // [n_em] rows of the block | poor << 15
    for (int e = threadIdx.x; e < X.n_em; e += blockDim.x) {
        if (staged) {
            const int b = X.em_blocks[e];
            einfo[e] = (uint16_t)((X.q_off[b + 1] - X.q_off[b]) | (X.poor[b] ? 0x8000 : 0));
        }
        cur[e] = 0;
        ext[e] = ((uint32_t)T.runs[tbase + e] << 16) | (uint32_t)T.cnt[tbase + e];
        boff[e] = (uint32_t)((T.toff[tbase + e] - chunk_base) >> 4);
    }
    __syncthreads();

    const int tid = threadIdx.x;
    int row[kScatterBatch], nrow[kScatterBatch];
    uint32_t pos[kScatterBatch], npos[kScatterBatch];
    int2 inf[kScatterBatch];
    double val[kScatterBatch], nval[kScatterBatch];
    auto load_rows = [&](int64_t s, int n) {
#pragma unroll
        for (int k = 0; k < kScatterBatch; k++) {
            const int i = k * kScatterThreads + tid;
            nrow[k] = i < n ? S.y_row[s + i] : -1;
            nval[k] = i < n ? S.y_val[s + i] : 0.0;
            npos[k] = i < n ? R.ypos[s + i] : 0u;
        }
    };
    auto load_info = [&]() {
#pragma unroll
        for (int k = 0; k < kScatterBatch; k++) {
            row[k] = nrow[k]; val[k] = nval[k]; pos[k] = npos[k];
            inf[k] = row[k] >= 0 ? X.row_info[row[k]] : make_int2(-1, 0);
        }
    };
    const int64_t vbase = R.rs_ptr[c0];                            // first result value of the chunk
    int64_t s = 0;
    int n = 0;
    if (nc > 0) { s = S.y_start[c0]; n = S.y_len[c0]; }
    if (staged && n <= kScatterCap) { load_rows(s, n); load_info(); }
    uint32_t crel = 0, crel2 = 0;                                  // the cell's row start relative to vbase
    for (int c = 0; c < nc; c++) {
        int64_t s2 = 0;
        int n2 = 0;
        if (c + 1 < nc) { s2 = S.y_start[c0 + c + 1]; n2 = S.y_len[c0 + c + 1]; crel2 = (uint32_t)(R.rs_ptr[c0 + c + 1] - vbase); }
        const bool next_staged = staged && (c + 1 < nc) && n2 <= kScatterCap;
        if (staged && n <= kScatterCap) {
#pragma unroll
            for (int k = 0; k < kScatterBatch; k++) {
                const int i = k * kScatterThreads + tid;
                if (i < n) { se[i] = inf[k].x; srl[i] = (uint16_t)inf[k].y; }
            }
            __syncthreads();
            if (next_staged) load_rows(s2, n2);
            // phase A: place entries; the first entry of a run also writes the record's cell id and start
#pragma unroll
            for (int k = 0; k < kScatterBatch; k++) {
                const int i = k * kScatterThreads + tid;
                const int e = inf[k].x;
                if (i >= n || e < 0) continue;                // pass-through block: written by struct_fill_kernel
                const uint32_t x = ext[e];
                const int nrec = (int)(x >> 16), E = (int)(x & 0xffff);
                if (nrec == 0) continue;                      // task rejected by classify (error flag is set)
                const int rl = inf[k].y;
                int head = i;
                while (head > 0 && se[head - 1] == e) head--;
                const BlobLayout L = blob_layout(nrec, E);
                char *blob = cdata + ((size_t)boff[e] << 4);
                double *ey = (double *)blob;
                uint32_t *rcell = (uint32_t *)(blob + L.off_cell);
                uint32_t *rpos = (uint32_t *)(blob + L.off_pos);
                uint16_t *eptr = (uint16_t *)(blob + L.off_ptr);
                uint16_t *erow = (uint16_t *)(blob + L.off_row);
                const uint32_t cu = cur[e];
                const int a = (int)(cu >> 16), k0 = (int)(cu & 0xffff);
                const bool is_head = (i == head);
                const unsigned ei = einfo[e];
                if (ei & 0x8000u) {                           // dense record: every row of the block
                    const int q = (int)(ei & 0x7fffu);
                    ey[k0 + rl] = val[k];
                    if (is_head) {
                        int tail = i;
                        while (tail + 1 < n && se[tail + 1] == e) tail++;
                        int kk = head;
                        for (int r = 0; r < q; r++) {
                            erow[k0 + r] = (uint16_t)r;
                            if (kk <= tail && srl[kk] == r) kk++; else ey[k0 + r] = 0.0;
                        }
                        rcell[a] = (uint32_t)c;
                        rpos[a] = crel + pos[k];
                        eptr[a] = (uint16_t)k0;
                    }
                } else {
                    ey[k0 + i - head] = val[k];
                    erow[k0 + i - head] = (uint16_t)rl;
                    if (is_head) { rcell[a] = (uint32_t)c; rpos[a] = crel + pos[k]; eptr[a] = (uint16_t)k0; }
                }
            }
            __syncthreads();
            // phase B: the last entry of each run advances the task's cursors
#pragma unroll
            for (int k = 0; k < kScatterBatch; k++) {
                const int i = k * kScatterThreads + tid;
                const int e = inf[k].x;
                if (i >= n || e < 0) continue;
                if (i != n - 1 && se[i + 1] == e) continue;
                int head = i;
                while (head > 0 && se[head - 1] == e) head--;
                const unsigned ei = einfo[e];
                const int m = (ei & 0x8000u) ? (int)(ei & 0x7fffu) : (i - head + 1);
                cur[e] += (1u << 16) + (uint32_t)m;
            }
            if (next_staged) load_info();
            __syncthreads();
        } else {
            if (next_staged) load_rows(s2, n2);
            // phase A: place entries; the first entry of a run also writes the record's cell id and start
            for (int i = threadIdx.x; i < n; i += blockDim.x) {
                const int2 info = X.row_info[S.y_row[s + i]];
                const int e = info.x;
                if (e < 0) continue;               // pass-through block: written by struct_fill_kernel
                const uint32_t x = ext[e];
                const int nrec = (int)(x >> 16), E = (int)(x & 0xffff);
                if (nrec == 0) continue;           // task rejected by classify (error flag is set)
                const double v = S.y_val[s + i];
                const int rl = info.y;
                int head = i;
                while (head > 0 && X.row_info[S.y_row[s + head - 1]].x == e) head--;
                const BlobLayout L = blob_layout(nrec, E);
                char *blob = cdata + ((size_t)boff[e] << 4);
                double *ey = (double *)blob;
                uint32_t *rcell = (uint32_t *)(blob + L.off_cell);
                uint32_t *rpos = (uint32_t *)(blob + L.off_pos);
                uint16_t *eptr = (uint16_t *)(blob + L.off_ptr);
                uint16_t *erow = (uint16_t *)(blob + L.off_row);
                const uint32_t cu = cur[e];
                const int a = (int)(cu >> 16), k0 = (int)(cu & 0xffff);
                const bool is_head = (i == head);
                const int b = X.em_blocks[e];
                if (X.poor[b]) {                   // dense record: every row of the block
                    const int q = X.q_off[b + 1] - X.q_off[b];
                    ey[k0 + rl] = v;
                    if (is_head) {
                        int tail = i;
                        while (tail + 1 < n && X.row_info[S.y_row[s + tail + 1]].x == e) tail++;
                        int k = head;
                        for (int r = 0; r < q; r++) {
                            erow[k0 + r] = (uint16_t)r;
                            if (k <= tail && X.row_info[S.y_row[s + k]].y == r) k++; else ey[k0 + r] = 0.0;
                        }
                        rcell[a] = (uint32_t)c;
                        rpos[a] = crel + R.ypos[s + i];
                        eptr[a] = (uint16_t)k0;
                    }
                } else {
                    ey[k0 + i - head] = v;
                    erow[k0 + i - head] = (uint16_t)rl;
                    if (is_head) { rcell[a] = (uint32_t)c; rpos[a] = crel + R.ypos[s + i]; eptr[a] = (uint16_t)k0; }
                }
            }
            __syncthreads();
            // phase B: the last entry of each run advances the task's cursors
            for (int i = threadIdx.x; i < n; i += blockDim.x) {
                const int e = X.row_info[S.y_row[s + i]].x;
                if (e < 0) continue;
                const bool tail = (i == n - 1) || (X.row_info[S.y_row[s + i + 1]].x != e);
                if (!tail) continue;
                int head = i;
                while (head > 0 && X.row_info[S.y_row[s + head - 1]].x == e) head--;
                const int b = X.em_blocks[e];
                const int m = X.poor[b] ? (X.q_off[b + 1] - X.q_off[b]) : (i - head + 1);
                cur[e] += (1u << 16) + (uint32_t)m;
            }
            __syncthreads();
            if (next_staged) load_info();
        }
        s = s2; n = n2; crel = crel2;
    }
    // close every blob of the chunk (eptr[n] = E); poor tasks get their pseudo cell: all cells of
    // the chunk without counts in the block
    for (int e = threadIdx.x; e < X.n_em; e += blockDim.x) {
        const uint32_t x = ext[e];
        const int nrec = (int)(x >> 16), E = (int)(x & 0xffff);
        if (nrec == 0) continue;
        const BlobLayout L = blob_layout(nrec, E);
        char *blob = cdata + ((size_t)boff[e] << 4);
        uint16_t *eptr = (uint16_t *)(blob + L.off_ptr);
        eptr[nrec] = (uint16_t)E;
        const int b = X.em_blocks[e];
        if (!X.poor[b]) continue;
        const int q = X.q_off[b + 1] - X.q_off[b];
        double *ey = (double *)blob;
        uint32_t *rcell = (uint32_t *)(blob + L.off_cell);
        uint16_t *erow = (uint16_t *)(blob + L.off_row);
        const int a = nrec - 1, k0 = a * q;
        for (int r = 0; r < q; r++) { ey[k0 + r] = 0.0; erow[k0 + r] = (uint16_t)r; }
        rcell[a] = kPseudo | (uint32_t)(nc - a);
        ((uint32_t *)(blob + L.off_pos))[a] = 0u;                  // the pseudo cell's values go wherever poor_pos says
        eptr[a] = (uint16_t)k0;
    }
}

__device__ __forceinline__ void task_done_stats(int32_t *iters, unsigned long long *stats, int64_t t, int outer,
                                                int inner_total, int q, int p, int nc, int ncell)
{
    if (iters) { iters[2 * t] = outer; iters[2 * t + 1] = inner_total; }
    atomicAdd(&stats[0], (unsigned long long)inner_total);
    atomicAdd(&stats[1], (unsigned long long)outer);
    atomicAdd(&stats[2], (unsigned long long)inner_total * (q + 2 * p) * nc);
    atomicAdd(&stats[3], (unsigned long long)outer * (q + p) * nc);
    atomicAdd(&stats[4], 1ull);
    atomicAdd(&stats[5], (unsigned long long)ncell);
}

// ------------------------------------------------------------------------------------------
// Column sums over cells: the rowSums of step2:167 (rows with sum > 0 survive).  Pass-through columns
// are summed by struct_fill_kernel (integer counts: exact).  EM columns: colsum_chunk_kernel, one CTA per
// chunk, walks the chunk's result rows in cell order and adds every EM entry to its column's slot in
// shared memory (a row holds a column at most once: no conflict inside a cell; a barrier between cells),
// which gives the chunk's sums in cell order (Result::colpart); colsum_em_reduce then adds the chunks in
// chunk order.  The sum is a fixed sequence whatever the GPU count or the queue order.
constexpr int kColsumThreads = 512;
__global__ void __launch_bounds__(kColsumThreads) colsum_chunk_kernel(Sample S, Result R, const int32_t *__restrict__ col_emcol,
                                                                      int n_emcols, int tile)
{
    extern __shared__ double colsum_sm[];                         // [tile] EM columns j0 .. j0 + tile of this pass
    const int h = blockIdx.x;
    const int64_t c0 = S.chunk_off[h] - S.cell_base, c1 = S.chunk_off[h + 1] - S.cell_base;
    double *out = R.colpart + (int64_t)(S.chunk_base + h) * n_emcols;
    constexpr int K = 6;
    int col[K];
    double val[K];
    auto load = [&](int64_t b0, int64_t e0) {
#pragma unroll
        for (int k = 0; k < K; k++) {
            const int64_t i = b0 + threadIdx.x + (int64_t)k * kColsumThreads;
            col[k] = -1; val[k] = 0.0;
            if (i < e0) { col[k] = col_emcol[R.rs_col[i]]; val[k] = R.rs_val[i]; }
        }
    };
    for (int j0 = 0; j0 < n_emcols; j0 += tile) {                 // one pass unless the X-matrix has very many EM columns
        const int jn = min(tile, n_emcols - j0);
        for (int j = threadIdx.x; j < jn; j += blockDim.x) colsum_sm[j] = 0.0;
        __syncthreads();
        // the next cell's entries are loaded while this cell's are added
        int64_t b = c0 < c1 ? R.rs_ptr[c0] : 0, e = c0 < c1 ? R.rs_ptr[c0 + 1] : 0;
        load(b, e);
        for (int64_t c = c0; c < c1; c++) {
            int jj[K];
            double vv[K];
#pragma unroll
            for (int k = 0; k < K; k++) { jj[k] = col[k] - j0; vv[k] = val[k]; }
            const int64_t b0 = b, e0 = e;
            if (c + 1 < c1) { b = e; e = R.rs_ptr[c + 2]; load(b, e); }
#pragma unroll
            for (int k = 0; k < K; k++) if (jj[k] >= 0 && jj[k] < jn) colsum_sm[jj[k]] += vv[k];
            for (int64_t i = b0 + threadIdx.x + (int64_t)K * kColsumThreads; i < e0; i += kColsumThreads) {
                const int j = col_emcol[R.rs_col[i]] - j0;
                if (j >= 0 && j < jn) colsum_sm[j] += R.rs_val[i];
            }
            __syncthreads();
        }
        for (int j = threadIdx.x; j < jn; j += blockDim.x) out[j0 + j] = colsum_sm[j];
        __syncthreads();
    }
}
__global__ void colsum_em_reduce(const double *__restrict__ colpart, int n_chunks, int n_emcols,
                                 const int32_t *__restrict__ emcol_col, double *__restrict__ colsum)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_emcols) return;
    double s = 0;
    for (int h = 0; h < n_chunks; h++) s += colpart[(int64_t)h * n_emcols + j];
    colsum[emcol_col[j]] = s;
}

// ------------------------------------------------------------------------------------------
// scasa_genemat on the sparse rows (scasa_quant_step3_v1.0.0.R:32-46): gene row = column sum over the
// gene's isoform rows, members added in row order (parasum's colSums, step3:27); single-member genes are
// copies (step3:38-40).  One CTA per cell (two per SM, 512 threads).
//   col_gene[j] = {gene of column j or -1, slot of j in the member table or -1 for single-member genes}
//   member table (shared memory): the members of every multi-member gene side by side in column order, a fixed
//   stop mark before the first gene and behind every gene; a slot holds the row entry that carries the member, or empty
//   A  every entry sets its gene's bit in a bitmap over the genes; entries of multi-member genes enter the table
//      and leave their value in shared memory
//   B  a block scan of the bitmap's word counts: an entry's output position is the number of present genes before
//      its own = word prefix + popcount below its bit; no sort
//   C  an entry of a multi-member gene looks back in the table: another present member before it => done.
//      Otherwise it is the gene's first present member and adds the later ones walking forward to the stop mark:
//      every sum is taken by one thread in member = column order, whatever the thread count; no rounds, no barriers.
//   D  (gene, value or sum) goes to the entry's position.  Output rows are ascending in gene id by construction.
constexpr int kGeneRowThreads = 512;
constexpr int kGeneKeep = 7;                 // entries a thread keeps in registers (rows up to 3,584 entries; the rest is re-read)
constexpr int kGeneCap = kGeneKeep * kGeneRowThreads;
template <typename T>                        // T = uint16_t when a row cannot have 65534 entries (smaller shared-memory footprint)
__global__ void __launch_bounds__(kGeneRowThreads, 2) gene_rows_kernel(Result R, int64_t n_cells, const int2 *__restrict__ col_gene,
                                                                       int n_genes, int n_slots, const int32_t *__restrict__ stops,
                                                                       int n_stops, int words_per_thread)
{
    extern __shared__ __align__(16) unsigned char gene_sm_raw[];
    constexpr T kEmpty = (T)~(T)0, kStop = (T)(kEmpty - 1);
    const int n_words = words_per_thread * kGeneRowThreads;                         // >= ceil(n_genes / 32)
    double *vsm = (double *)gene_sm_raw;                                            // [kGeneCap] values of the multi-member entries
    unsigned *bits = (unsigned *)(vsm + kGeneCap);                                  // [n_words] genes present
    int *wpre = (int *)(bits + n_words);                                            // [n_words] present genes before the word
    T *mtab = (T *)(wpre + n_words);                                                // [n_slots] member table
    __shared__ int wsum[kGeneRowThreads / 32];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    for (int i = tid; i < n_words; i += kGeneRowThreads) bits[i] = 0u;
    for (int i = tid; i < n_slots; i += kGeneRowThreads) mtab[i] = kEmpty;
    __syncthreads();
    for (int i = tid; i < n_stops; i += kGeneRowThreads) mtab[stops[i]] = kStop;
    int64_t nx_base = 0, nx_end = 0;                                  // extent of the row this CTA takes next
    if (blockIdx.x < n_cells) { nx_base = R.rs_ptr[blockIdx.x]; nx_end = R.rs_ptr[blockIdx.x + 1]; }
    for (int64_t cell = blockIdx.x; cell < n_cells; cell += gridDim.x) {
        const int64_t base = nx_base;
        const int n = (int)(nx_end - base);
        const int32_t *col = R.rs_col + base;
        const double *val = R.rs_val + base;
        const int64_t nx = cell + gridDim.x;
        if (nx < n_cells) { nx_base = R.rs_ptr[nx]; nx_end = R.rs_ptr[nx + 1]; }
        int2 cg[kGeneKeep];
        double v[kGeneKeep];
#pragma unroll
        for (int k = 0; k < kGeneKeep; k++) {
            const int i = tid + k * kGeneRowThreads;
            cg[k] = make_int2(-1, -1); v[k] = 0.0;
            if (i < n) { cg[k] = col_gene[col[i]]; v[k] = val[i]; }
        }
        const bool long_row = n > kGeneCap;
        __syncthreads();                                              // the previous cell's reset (and the stop marks) are in place
        // ---- A
#pragma unroll
        for (int k = 0; k < kGeneKeep; k++)
            if (cg[k].x >= 0) {
                atomicOr(&bits[cg[k].x >> 5], 1u << (cg[k].x & 31));
                if (cg[k].y >= 0) { const int i = tid + k * kGeneRowThreads; mtab[cg[k].y] = (T)i; vsm[i] = v[k]; }
            }
        if (long_row)
            for (int i = tid + kGeneCap; i < n; i += kGeneRowThreads) {
                const int2 c = col_gene[col[i]];
                if (c.x >= 0) { atomicOr(&bits[c.x >> 5], 1u << (c.x & 31)); if (c.y >= 0) mtab[c.y] = (T)i; }
            }
        __syncthreads();
        if (nx < n_cells) {                                           // the next row on its way into L2 while this one is placed
            const int nn = (int)min(nx_end - nx_base, (int64_t)kGeneCap);
            if (tid * 32 < nn) asm volatile("prefetch.global.L2 [%0];" ::"l"(R.rs_col + nx_base + tid * 32));
            if (tid * 16 < nn) asm volatile("prefetch.global.L2 [%0];" ::"l"(R.rs_val + nx_base + tid * 16));
        }
        // ---- B: thread owns words [tid*wpt, (tid+1)*wpt)
        const int w0 = tid * words_per_thread;
        int cnt = 0;
        for (int k = 0; k < words_per_thread; k++) cnt += __popc(bits[w0 + k]);
        int inc = cnt;
        for (int o = 1; o < 32; o <<= 1) { const int x = __shfl_up_sync(kFull, inc, o); if (lane >= o) inc += x; }
        if (lane == 31) wsum[w] = inc;
        __syncthreads();
        int before = 0, total = 0;
        for (int k = 0; k < kGeneRowThreads / 32; k++) { const int x = wsum[k]; if (k < w) before += x; total += x; }
        {
            int run = before + inc - cnt;
            for (int k = 0; k < words_per_thread; k++) { wpre[w0 + k] = run; run += __popc(bits[w0 + k]); }
        }
        __syncthreads();
        // ---- C + D
        auto emit = [&](int2 c, double x) {
            if (c.x < 0) return;
            if (c.y >= 0) {
                for (int t = c.y - 1;; t--) {                         // an earlier member is present: it owns the sum
                    const T h = mtab[t];
                    if (h == kStop) break;
                    if (h != kEmpty) return;
                }
                for (int t = c.y + 1;; t++) {
                    const T h = mtab[t];
                    if (h == kStop) break;
                    if (h != kEmpty) x += (int64_t)h < kGeneCap ? vsm[h] : val[h];
                }
            }
            const int g = c.x;
            const int64_t o = base + wpre[g >> 5] + __popc(bits[g >> 5] & ((1u << (g & 31)) - 1u));
            R.g_col[o] = g; R.g_val[o] = x;
        };
#pragma unroll
        for (int k = 0; k < kGeneKeep; k++) emit(cg[k], v[k]);
        if (long_row)
            for (int i = tid + kGeneCap; i < n; i += kGeneRowThreads) emit(col_gene[col[i]], val[i]);
        if (tid == 0) R.g_len[cell] = total;
        __syncthreads();
        // ---- reset for the next cell (its first barrier publishes this)
        for (int k = 0; k < words_per_thread; k++) bits[w0 + k] = 0u;
#pragma unroll
        for (int k = 0; k < kGeneKeep; k++) if (cg[k].y >= 0) mtab[cg[k].y] = kEmpty;
        if (long_row)
            for (int i = tid + kGeneCap; i < n; i += kGeneRowThreads) { const int2 c = col_gene[col[i]]; if (c.y >= 0) mtab[c.y] = kEmpty; }
    }
}

// ------------------------------------------------------------------------------------------
// Stored zeros out of a CSR result before it crosses PCIe (EM columns whose estimate is exactly 0: a fifth of the
// entries; the expansion on the host writes zeros anyway and the sparse interfaces drop them).  One warp per row:
// row_nonzeros_kernel counts, an exclusive scan gives the packed row starts, row_compact_kernel moves the entries
// (ballot + popcount keep their order).  len == nullptr: full rows.
__global__ void row_nonzeros_kernel(const int64_t *__restrict__ ptr, const int32_t *__restrict__ len, const double *__restrict__ val,
                                    int64_t n_rows, int32_t *__restrict__ out, bool keep_all)
{
    const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= n_rows) return;
    const int lane = threadIdx.x & 31;
    const int64_t s = ptr[r];
    const int n = len ? len[r] : (int)(ptr[r + 1] - s);
    int cnt = 0;
    if (keep_all) cnt = lane == 0 ? n : 0;                      // rows only closed up (gene rows have the capacity of the isoform rows)
    else for (int i = lane; i < n; i += kWarp) cnt += val[s + i] != 0.0;
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(kFull, cnt, o);
    if (lane == 0) out[r] = cnt;
}
__global__ void row_compact_kernel(const int64_t *__restrict__ ptr, const int32_t *__restrict__ len, const int32_t *__restrict__ col,
                                   const double *__restrict__ val, int64_t n_rows, const int64_t *__restrict__ out_ptr,
                                   int32_t *__restrict__ out_col, double *__restrict__ out_val, bool keep_all)
{
    const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= n_rows) return;
    const int lane = threadIdx.x & 31;
    const int64_t s = ptr[r];
    const int n = len ? len[r] : (int)(ptr[r + 1] - s);
    int64_t o = out_ptr[r];
    for (int i0 = 0; i0 < n; i0 += kWarp) {
        const int i = i0 + lane;
        double v = 0.0;
        int c = 0;
        if (i < n) { v = val[s + i]; c = col[s + i]; }
        const bool mine = i < n && (keep_all || v != 0.0);
        const unsigned keep = __ballot_sync(kFull, mine);
        if (mine) {
            const int64_t d = o + __popc(keep & ((1u << lane) - 1u));
            out_col[d] = c; out_val[d] = v;
        }
        o += __popc(keep);
    }
}

// rows [r0, r0 + n) of a CSR result -> dense rows (testing aid: SCASA_FETCH_DENSE); len == nullptr: full rows
__global__ void expand_rows_kernel(const int64_t *__restrict__ ptr, const int32_t *__restrict__ len, const int32_t *__restrict__ col,
                                   const double *__restrict__ val, int64_t r0, int64_t n, int64_t n_cols, double *__restrict__ out)
{
    for (int64_t r = blockIdx.x; r < n; r += gridDim.x) {
        double *o = out + r * n_cols;
        for (int64_t j = threadIdx.x; j < n_cols; j += blockDim.x) o[j] = 0.0;
        __syncthreads();
        const int64_t b = ptr[r0 + r];
        const int m = len ? len[r0 + r] : (int)(ptr[r0 + r + 1] - b);
        for (int i = threadIdx.x; i < m; i += blockDim.x) o[col[b + i]] = val[b + i];
        __syncthreads();
    }
}

// scasa_genemat as a gather: one thread per (cell, gene), members read straight from the isoform row in
// column order (step3:27).  Every isoform value is read exactly once, so HBM traffic is the two
// matrices; the row being gathered sits in L2 (a few hundred KB per CTA row).  kGeneIlp genes per
// thread keep that many gathers in flight.
constexpr int kGeneThreads = 256;
constexpr int kGeneIlp = 4;
__global__ void __launch_bounds__(kGeneThreads) gene_gather_kernel(const double *__restrict__ iso, int64_t n_cells,
                                                                   int64_t n_cols, const int32_t *__restrict__ gene_ptr,
                                                                   const int32_t *__restrict__ gene_cols, int n_genes,
                                                                   int tiles_per_cell, double *__restrict__ gene)
{
    const int64_t c = blockIdx.x / tiles_per_cell;
    const int tile = blockIdx.x - (int)(c * tiles_per_cell);
    const double *row = iso + c * n_cols;
    double *out = gene + c * (int64_t)n_genes;
    const int g0 = tile * (kGeneThreads * kGeneIlp) + threadIdx.x;
    int lo[kGeneIlp], hi[kGeneIlp];
    double sum[kGeneIlp];
#pragma unroll
    for (int u = 0; u < kGeneIlp; u++) {
        const int g = g0 + u * kGeneThreads;
        lo[u] = g < n_genes ? gene_ptr[g] : 0;
        hi[u] = g < n_genes ? gene_ptr[g + 1] : 0;
    }
#pragma unroll
    for (int u = 0; u < kGeneIlp; u++) sum[u] = lo[u] < hi[u] ? row[gene_cols[lo[u]]] : 0.0;   // the copies (step3:39)
#pragma unroll
    for (int u = 0; u < kGeneIlp; u++)
        for (int m = lo[u] + 1; m < hi[u]; m++) sum[u] += row[gene_cols[m]];                    // parasum (step3:24-30)
#pragma unroll
    for (int u = 0; u < kGeneIlp; u++) {
        const int g = g0 + u * kGeneThreads;
        if (g < n_genes) out[g] = sum[u];
    }
}

}  // namespace scasa

#include "scasa_em.cuh"
