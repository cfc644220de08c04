// scasa_kernels.cuh — device side of libscasa_cuda (sm_100a).
//
// Data flow of one scasa_run() (everything resident in HBM):
//
//   [eqclass counts] --pack_cells_kernel--> Y as CSR by cell (rows ascending, duplicates summed)
//   Y --task_count_kernel--> entries / runs per (chunk, EM block) "task"
//     --classify_kernel--> blob size per task, light / heavy work lists; scans give the offsets
//   Y --task_scatter_kernel--> one blob per task: per cell-with-counts a record
//        {logNorm, m, cell | beta[p] | m x (count, row)}; pass-through blocks (nrow == 1,
//        Scasa_Rsource.R:979-983) are written straight into the result
//   blobs --rounds of beta_pass_kernel + task_step_{small,big}_kernel--> the alternating EM, one
//        batched estim.BETA iteration over all still-running cells per round (see below)
//   result --finish_kernel--> per-column sums (step2:167) and gene sums (step3:32-46)
//
// Why sparse: a cell with no counts in a block keeps beta = 0 through every iteration of
// estim.BETA and contributes 0 to every sum of estim.X.em, so only (cell, block) pairs with
// counts ("runs") are stored.  The convergence tests stay what the reference has: maxima over
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
    const int32_t *row_block; // [n_rows]  block of a global row
    const int32_t *em_index;  // [n_blocks] index among EM blocks, -1 for pass-through
    const int32_t *em_blocks; // [n_em]    block id
    const uint8_t *poor;      // [n_blocks] poorCondX (Rsource:957-971)
    const int32_t *poor_em;   // [n_poor]  EM indices of the poor blocks
    int n_poor;
};

struct Sample {               // what is staged for one scasa_run
    int64_t n_cells;
    int n_chunks;
    const int64_t *chunk_off;   // [n_chunks+1]
    const int32_t *cell_chunk;  // [n_cells]
    const int64_t *y_start;     // [n_cells]  first Y entry of the cell
    const int32_t *y_len;       // [n_cells]
    const int32_t *y_row;       // global row, ascending inside a cell
    const double *y_val;
};

// ---- task blob ----------------------------------------------------------------------------
// blob(t) = [TaskState | Xc | Xn | cs | bts] [record of cell 0] [record of cell 1] ...
// record  = CellHdr | beta[ppad] | Ent[m]          (16-byte aligned throughout)
struct alignas(16) CellHdr { double lnorm; int32_t m; uint32_t cell; };  // cell: bit 31 = pseudo cell, low 16 bits = cell in chunk / multiplicity
struct alignas(16) Ent { double y; int32_t row; int32_t pad; };
constexpr uint32_t kPseudo = 0x80000000u;

__host__ __device__ inline int ppad_of(int p) { return (p + 1) & ~1; }
__host__ __device__ inline int rec_bytes(int p, int m) { return 16 + 8 * ppad_of(p) + 16 * m; }
__host__ __device__ inline int state_bytes(int q, int p) { return 64 + 8 * ((2 * q * p + 2 * p + 1) & ~1); }

constexpr int kStepSlot = 32;       // doubles of X.est + colSums(BETA) per thread of the small step kernel: q*p + p
constexpr int kSmallP = 8;          // widest block the thread-per-task step kernel takes

struct Tasks {                // task t = chunk * n_em + em index
    int32_t *cnt, *runs;      // [n_tasks] Y entries / cells with counts (after classify: records incl. pseudo cell)
    int32_t *tbytes, *save;   // [n_tasks] blob bytes / bytes of the state area in front of the records
    int64_t *toff, *run_off;  // [n_tasks+1] exclusive scans of tbytes / runs
    int64_t *byte_cur, *run_cur;  // [n_tasks] cursors of the scatter
    char *tdata, *scr;        // blobs; scratch of the same shape for the heavy kernel (gamma, w)
    int64_t *rec_pos;         // [total records] byte offset of each record, in run order
    int32_t *small_list, *big_list;   // task ids, block-major
    unsigned int *counters;   // [0] n_small [2] n_big [3] tasks still running
};
// ------------------------------------------------------------------------------------------
// crpcount_td's accumulation (Scasa_Rsource.R:528-531, :643-646 and :461-462), fast path: when the
// X-matrix has few enough rows for an int32 histogram in shared memory (hg38: 41,251 rows = 165 KB)
// one CTA per cell adds the UMI counts of the cell's eqclasses into hist[row] and then emits the
// non-zero rows in ascending order.  Integer adds: the result is independent of their order.
// Thread i owns the contiguous rows [i*seg, (i+1)*seg), seg odd (no bank conflicts); the histogram
// is left all-zero for the next cell.
__global__ void __launch_bounds__(512) pack_cells_dense_kernel(int64_t n_cells, const int64_t *__restrict__ cell_ptr,
                                                               const int32_t *__restrict__ eq_id,
                                                               const int32_t *__restrict__ eq_count,
                                                               const int32_t *__restrict__ eq_row, int64_t n_eq, int n_rows,
                                                               int seg, int32_t *__restrict__ y_row,
                                                               double *__restrict__ y_val, int32_t *__restrict__ y_len)
{
    extern __shared__ int hist[];
    __shared__ int wsum[16];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    for (int i = tid; i < seg * (int)blockDim.x; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    for (int64_t cell = blockIdx.x; cell < n_cells; cell += gridDim.x) {
        const int64_t s = cell_ptr[cell];
        const int n = (int)(cell_ptr[cell + 1] - s);
        for (int i = tid; i < n; i += blockDim.x) {
            const int e = eq_id[s + i];
            const int r = (e >= 0 && e < n_eq) ? eq_row[e] : -1;
            const int c = eq_count[s + i];
            if (r >= 0 && r < n_rows && c != 0) atomicAdd(&hist[r], c);
        }
        __syncthreads();
        const int r0 = tid * seg;
        int cnt = 0;
        for (int k = 0; k < seg; k++) cnt += hist[r0 + k] != 0;
        // block-wide exclusive scan of cnt
        int inc = cnt;
        for (int o = 1; o < 32; o <<= 1) { int v = __shfl_up_sync(kFull, inc, o); if (lane >= o) inc += v; }
        if (lane == 31) wsum[w] = inc;
        __syncthreads();
        int base = 0, total = 0;
        for (int k = 0; k < (int)(blockDim.x >> 5); k++) { int v = wsum[k]; if (k < w) base += v; total += v; }
        int o = base + inc - cnt;
        if (cnt)
            for (int k = 0; k < seg; k++) {
                const int v = hist[r0 + k];
                if (v != 0) { y_row[s + o] = r0 + k; y_val[s + o] = (double)v; o++; hist[r0 + k] = 0; }
            }
        if (tid == 0) y_len[cell] = total;
        __syncthreads();
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
    __shared__ int s_heads;
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
                if (r >= 0 && c != 0) k = ((unsigned long long)(unsigned)r << 32) | (unsigned)c;
            }
            sm_keys[i] = k;
        }
        if (threadIdx.x == 0) s_heads = 0;
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
// Task extents.  One warp per cell: count this cell's entries and runs per (chunk, EM block).
__global__ void task_count_kernel(XDev X, Sample S, Tasks T)
{
    int64_t cell = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (cell >= S.n_cells) return;
    const int lane = threadIdx.x & 31;
    const int64_t s = S.y_start[cell];
    const int n = S.y_len[cell];
    const int64_t tbase = (int64_t)S.cell_chunk[cell] * X.n_em;
    for (int i = lane; i < n; i += kWarp) {
        int b = X.row_block[S.y_row[s + i]];
        int e = X.em_index[b];
        if (e < 0) continue;
        atomicAdd(&T.cnt[tbase + e], 1);
        bool head = (i == 0) || (X.row_block[S.y_row[s + i - 1]] != b);
        if (head) atomicAdd(&T.runs[tbase + e], 1);
    }
}

struct EmOpts {
    int maxiter_x, maxiter_bt, modify, fixed_iter;
    double maxerr, lim, adjust_esp;
};

// Per task: blob size, work list.  Threads enumerate tasks block-major (e outer, chunk inner) so
// that neighbouring list entries — hence the lanes of a warp of the light kernel — mostly work
// on the same block shape.  All-zero tasks finish here: the reference runs one inner and one
// outer iteration on them and returns zeros (the result buffer is pre-zeroed).
__global__ void classify_kernel(XDev X, Sample S, Tasks T, int64_t n_tasks, EmOpts O, int32_t *iters,
                                unsigned long long *stats)
{
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_tasks) return;
    const int e = (int)(idx / S.n_chunks), h = (int)(idx % S.n_chunks);
    const int64_t t = (int64_t)h * X.n_em + e;
    const int b = X.em_blocks[e];
    const int q = X.q_off[b + 1] - X.q_off[b], p = X.p_off[b + 1] - X.p_off[b];
    const int nc = (int)(S.chunk_off[h + 1] - S.chunk_off[h]);
    const int c = T.cnt[t];
    if (c == 0) {
        T.tbytes[t] = 0; T.save[t] = 0;
        const int o = O.fixed_iter ? O.maxiter_x : 1;
        const int in = O.fixed_iter ? O.maxiter_x * O.maxiter_bt : 1;
        if (iters) { iters[2 * t] = o; iters[2 * t + 1] = in; }
        atomicAdd(&stats[0], (unsigned long long)in);
        atomicAdd(&stats[1], (unsigned long long)o);
        atomicAdd(&stats[2], (unsigned long long)in * (q + 2 * p) * nc);
        atomicAdd(&stats[3], (unsigned long long)o * (q + p) * nc);
        return;
    }
    int ncell = T.runs[t], ents = c;
    if (X.poor[b]) { ncell += 1; ents = ncell * q; }      // dense records + the pseudo cell
    if (X.poor[b] && q > 64) { atomicExch((int *)&T.counters[7], 1); }   // adjID is kept as a 64-bit mask
    const bool small = (p <= kSmallP) && (q * p + p <= kStepSlot);
    const int sv = state_bytes(q, p);
    T.save[t] = sv;
    T.tbytes[t] = sv + ncell * (16 + 8 * ppad_of(p)) + 16 * ents;
    T.runs[t] = ncell;
    if (small) T.small_list[atomicAdd(&T.counters[0], 1u)] = (int32_t)t;
    else T.big_list[atomicAdd(&T.counters[2], 1u)] = (int32_t)t;
    atomicAdd(&T.counters[3], 1u);
}

__global__ void cursor_init_kernel(Tasks T, int64_t n_tasks)
{
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tasks) return;
    T.byte_cur[t] = T.toff[t] + T.save[t];
    T.run_cur[t] = T.run_off[t];
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
__global__ void task_scatter_kernel(XDev X, Sample S, Tasks T, double *__restrict__ out)
{
    const int h = blockIdx.x;
    const int64_t c0 = S.chunk_off[h], c1 = S.chunk_off[h + 1];
    const int nc = (int)(c1 - c0);
    const int64_t tbase = (int64_t)h * X.n_em;

    for (int c = 0; c < nc; c++) {
        const int64_t cell = c0 + c;
        const int64_t s = S.y_start[cell];
        const int n = S.y_len[cell];
        // phase A: place entries; the first entry of a run also writes the record header
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const int row = S.y_row[s + i];
            const double v = S.y_val[s + i];
            const int b = X.row_block[row];
            const int e = X.em_index[b];
            if (e < 0) {                       // nrow(X0) == 1: the count is the estimate (Rsource:979-983)
                out[cell * X.n_cols + X.p_off[b]] = v;
                continue;
            }
            const int64_t t = tbase + e;
            const int rl = row - X.q_off[b];
            const int p = X.p_off[b + 1] - X.p_off[b];
            int head = i;
            while (head > 0 && X.row_block[S.y_row[s + head - 1]] == b) head--;
            const int64_t pos = T.byte_cur[t];
            char *rec = T.tdata + pos;
            Ent *en = (Ent *)(rec + 16 + 8 * ppad_of(p));
            const bool is_head = (i == head);
            int tail = i;
            if (is_head) while (tail + 1 < n && X.row_block[S.y_row[s + tail + 1]] == b) tail++;
            if (X.poor[b]) {                   // dense record: every row of the block
                const int q = X.q_off[b + 1] - X.q_off[b];
                en[rl].y = v;
                if (is_head) {
                    int k = head;
                    for (int r = 0; r < q; r++) {
                        en[r].row = r; en[r].pad = 0;
                        if (k <= tail && S.y_row[s + k] - X.q_off[b] == r) k++; else en[r].y = 0.0;
                    }
                    CellHdr hd; hd.lnorm = 0.0; hd.m = q; hd.cell = (uint32_t)c;
                    *(CellHdr *)rec = hd;
                    T.rec_pos[T.run_cur[t]] = pos;
                }
            } else {
                Ent ev; ev.y = v; ev.row = rl; ev.pad = 0;
                en[i - head] = ev;
                if (is_head) {
                    CellHdr hd; hd.lnorm = 0.0; hd.m = tail - head + 1; hd.cell = (uint32_t)c;
                    *(CellHdr *)rec = hd;
                    T.rec_pos[T.run_cur[t]] = pos;
                }
            }
        }
        __syncthreads();
        // phase B: the last entry of each run advances the task's cursors
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const int b = X.row_block[S.y_row[s + i]];
            const int e = X.em_index[b];
            if (e < 0) continue;
            const bool tail = (i == n - 1) || (X.row_block[S.y_row[s + i + 1]] != b);
            if (!tail) continue;
            int head = i;
            while (head > 0 && X.row_block[S.y_row[s + head - 1]] == b) head--;
            const int64_t t = tbase + e;
            const int p = X.p_off[b + 1] - X.p_off[b];
            const int m = X.poor[b] ? (X.q_off[b + 1] - X.q_off[b]) : (i - head + 1);
            T.byte_cur[t] += rec_bytes(p, m);
            T.run_cur[t] += 1;
        }
        __syncthreads();
    }
    // pseudo cell of every non-empty poor task: all cells of the chunk without counts in the block
    for (int k = threadIdx.x; k < X.n_poor; k += blockDim.x) {
        const int e = X.poor_em[k];
        const int64_t t = tbase + e;
        if (T.cnt[t] == 0) continue;
        const int b = X.em_blocks[e];
        const int q = X.q_off[b + 1] - X.q_off[b], p = X.p_off[b + 1] - X.p_off[b];
        const int64_t pos = T.byte_cur[t];
        char *rec = T.tdata + pos;
        Ent *en = (Ent *)(rec + 16 + 8 * ppad_of(p));
        for (int r = 0; r < q; r++) { Ent ev; ev.y = 0.0; ev.row = r; ev.pad = 0; en[r] = ev; }
        CellHdr hd; hd.lnorm = 0.0; hd.m = q; hd.cell = kPseudo | (uint32_t)(nc - (T.runs[t] - 1));
        *(CellHdr *)rec = hd;
        T.rec_pos[T.run_cur[t]] = pos;
    }
}
// ------------------------------------------------------------------------------------------
// Bulk-synchronous EM.  One ROUND = one estim.BETA iteration (Rsource:749-758) for every task
// that is still running, executed as
//     beta_pass_kernel   one thread per cell record: gamma (digamma + exp), mu = X gamma for the
//                        rows with counts, the y/mu ratio and the X^T back-projection, fused;
//                        writes beta in place and raises the task's fail flag
//     task_step_kernel   one thread (small tasks) or one warp (big tasks) per task: reads the
//                        flag = the chunk-wide max(err) < maxerr test; when estim.BETA is over it
//                        runs estim.X.em (:898-922), the outer test (:1019-1020) and either
//                        starts the next estim.BETA call or de-adjusts and writes the result.
// Converged tasks drop out; the host compacts the task / record lists when half of them are gone.
// Sums over cells run in record (= cell) order inside one thread, or as per-lane partial sums in
// cell order plus a fixed xor tree inside one warp: the result does not depend on how tasks are
// scheduled, listed or sharded over GPUs.
struct alignas(16) TaskState {
    int32_t alive, fail, outer, inner_it, inner_total, nadj;
    int32_t ncell, pad0;
    unsigned long long adjmask;
    int32_t pad1[6];
};  // 64 bytes, followed by Xc[qp] Xn[qp] cs[p] bts[p]
static_assert(sizeof(TaskState) == 64, "TaskState layout");


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

struct RoundArgs {
    XDev X;
    Sample S;
    Tasks T;
    EmOpts O;
    double *out;
    int32_t *iters;
    unsigned long long *stats;
    const int32_t *tlist;        // task ids (small or big list)
    const unsigned int *tlist_n;
    int32_t *alive_flags;        // [list length] written by the step kernels for the compaction
    const int32_t *r_task;       // record list: task id / byte offset of the record
    const int64_t *r_off;
    int64_t n_rec;
    int xcap, pcap, qcap;        // warp-per-task kernels: shared memory sizing (max q*p, p, q)
};

// per task: state header, X.est0 = X0, colSums(X); for poor blocks rowSums(Ymat), bestTx and adjID
// (Rsource:989-1004).  One thread per listed task (poor blocks are narrow: p = 2).
__global__ void task_init_kernel(RoundArgs A)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= *A.tlist_n) return;
    const XDev &X = A.X;
    const Tasks &T = A.T;
    const int64_t t = A.tlist[i];
    const int e = (int)(t % X.n_em);
    const int b = X.em_blocks[e];
    const int q = X.q_off[b + 1] - X.q_off[b], p = X.p_off[b + 1] - X.p_off[b];
    const int qp = q * p, ppad = ppad_of(p);
    char *blob = T.tdata + T.toff[t];
    TaskState *st = (TaskState *)blob;
    double *Xc = (double *)(blob + 64), *cs = Xc + 2 * qp;
    const double *X0 = X.x + X.x_off[b];
    for (int k = 0; k < qp; k++) Xc[k] = X0[k];                       // X.est0 = X0  :1011
    for (int j = 0; j < p; j++) {                                     // csum = colSums(X)  :731
        double s = 0;
        for (int r = 0; r < q; r++) s += X0[j * q + r];
        cs[j] = s;
    }
    const int ncell = T.runs[t];
    int nadj = 0;
    unsigned long long adjmask = 0;
    if (X.poor[b]) {
        double *rs = Xc + qp;                                         // Xn as scratch
        for (int r = 0; r < q; r++) rs[r] = 0.0;
        const char *rec = blob + T.save[t];
        for (int a = 0; a < ncell; a++) {
            const CellHdr hd = *(const CellHdr *)rec;
            const Ent *en = (const Ent *)(rec + 16 + 8 * ppad);
            const double mult = (hd.cell & kPseudo) ? (double)(hd.cell & 0xffff) : 1.0;
            for (int k = 0; k < hd.m; k++) rs[en[k].row] += mult * en[k].y;
            rec += rec_bytes(p, hd.m);
        }
        double tot = 0;
        for (int r = 0; r < q; r++) tot += rs[r];
        int best = -1;
        double bestv = -INFINITY;
        for (int j = 0; j < p; j++) {                                 // which.max(cos_sim(X0[,j], rs/sum(rs)))  :991-995
            double ab = 0, aa = 0, bb = 0;
            for (int r = 0; r < q; r++) {
                const double xv = X0[j * q + r], rv = rs[r] / tot;
                ab += xv * rv; aa += xv * xv; bb += rv * rv;
            }
            const double sim = ab / sqrt(aa * bb);
            if (sim == sim && sim > bestv) { bestv = sim; best = j; }
        }
        if (best >= 0)
            for (int r = 0; r < q; r++)
                if (X0[best * q + r] > 0) { adjmask |= 1ull << r; nadj++; }       // adjID  :997
    }
    TaskState s0;
    s0.alive = 1; s0.fail = 0; s0.outer = 0; s0.inner_it = 0; s0.inner_total = 0; s0.nadj = nadj;
    s0.ncell = ncell; s0.pad0 = 0; s0.adjmask = adjmask;
    for (int k = 0; k < 6; k++) s0.pad1[k] = 0;
    *st = s0;
    A.alive_flags[i] = 1;
}

// per record: Ymat[adjID,] += adjust_esp (:1004), beta0 = Ysum / p (:1012-1013), logNorm (:732-733)
__global__ void rec_init_kernel(RoundArgs A)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= A.n_rec) return;
    const XDev &X = A.X;
    const Tasks &T = A.T;
    const int64_t t = A.r_task[i];
    const int b = X.em_blocks[(int)(t % X.n_em)];
    const int p = X.p_off[b + 1] - X.p_off[b];
    const TaskState *st = (const TaskState *)(T.tdata + T.toff[t]);
    const unsigned long long adjmask = st->adjmask;
    char *rec = T.tdata + A.r_off[i];
    const CellHdr hd = *(const CellHdr *)rec;
    double *bt = (double *)(rec + 16);
    Ent *en = (Ent *)(rec + 16 + 8 * ppad_of(p));
    double ys = 0;
    for (int k = 0; k < hd.m; k++) {
        Ent ev = en[k];
        if (adjmask >> ev.row & 1ull) { ev.y += A.O.adjust_esp; en[k].y = ev.y; }
        ys += ev.y;
    }
    ((CellHdr *)rec)->lnorm = digamma_pos(ys + 0.00000001);
    const double b0 = ys / p;
    for (int j = 0; j < p; j++) bt[j] = b0;
}

// One estim.BETA iteration for one cell (fn1 + fn2, Rsource:736-747, and the err of :755).
__global__ void __launch_bounds__(256) beta_pass_kernel(RoundArgs A)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= A.n_rec) return;
    const XDev &X = A.X;
    const Tasks &T = A.T;
    const int64_t t = A.r_task[i];
    char *blob = T.tdata + T.toff[t];
    TaskState *st = (TaskState *)blob;
    if (!st->alive) return;
    const int b = X.em_blocks[(int)(t % X.n_em)];
    const int q = X.q_off[b + 1] - X.q_off[b], p = X.p_off[b + 1] - X.p_off[b];
    const int qp = q * p;
    const double *Xc = (const double *)(blob + 64), *cs = Xc + 2 * qp;
    const int64_t off = A.r_off[i];
    char *rec = T.tdata + off;
    const CellHdr hd = *(const CellHdr *)rec;
    double *bt = (double *)(rec + 16);
    const Ent *en = (const Ent *)(rec + 16 + 8 * ppad_of(p));
    double *g = (double *)(T.scr + off + 16);                       // gamma0 of wide blocks
    double *sa = (double *)(T.scr + off + 16 + 8 * ppad_of(p));     // wide blocks: one weight per entry (stride 2)
    double g0 = 0, g1 = 0, acc0 = 0, acc1 = 0;                      // p <= 2 stays in registers
    if (p <= 2) {
        g0 = em_gamma(bt[0], hd.lnorm, A.O.modify);
        if (p == 2) g1 = em_gamma(bt[1], hd.lnorm, A.O.modify);
        for (int k = 0; k < hd.m; k++) {
            const Ent ev = en[k];
            const double x0 = Xc[ev.row], x1 = (p == 2) ? Xc[q + ev.row] : 0.0;
            const double mu = x0 * g0 + x1 * g1;                    // rowSums(t(t(X)*gamma0))
            const double w = ev.y / (mu > 0 ? mu : 1.0);
            acc0 += x0 * w; acc1 += x1 * w;
        }
    } else {
        for (int j = 0; j < p; j++) g[j] = em_gamma(bt[j], hd.lnorm, A.O.modify);
        for (int k = 0; k < hd.m; k++) {
            const Ent ev = en[k];
            double mu = 0.0;
            for (int j = 0; j < p; j++) mu += Xc[j * q + ev.row] * g[j];
            sa[2 * k] = ev.y / (mu > 0 ? mu : 1.0);
        }
    }
    bool fail = false;
    for (int j = 0; j < p; j++) {
        double s, gj;
        if (p <= 2) { s = j ? acc1 : acc0; gj = j ? g1 : g0; }
        else {
            s = 0.0;
            for (int k = 0; k < hd.m; k++) s += Xc[j * q + en[k].row] * sa[2 * k];
            gj = g[j];
        }
        const double old = bt[j];
        double nb = gj * s / cs[j];
        if (nb != nb) nb = 0.0;                                      // :754
        const double er = fabs(nb - old) / (old > A.O.lim ? old : A.O.lim);
        if (!(er < A.O.maxerr)) fail = true;                         // :755-756
        bt[j] = nb;
    }
    if (fail) st->fail = 1;                                          // every writer stores the same value
}

__device__ __forceinline__ void finish_task(const RoundArgs &A, int64_t t, int i, TaskState *st, int q, int p, int nc)
{
    st->alive = 0;
    A.alive_flags[i] = 0;
    task_done_stats(A.iters, A.stats, t, st->outer, st->inner_total, q, p, nc, st->ncell);
    atomicSub(&A.T.counters[3], 1u);
}

// After a beta pass, per SMALL task (a tile of 8 lanes): estim.BETA's loop control, estim.X.em, the
// outer loop control, the result.  Lanes split the task's records (lane l takes records l, l+8, ..),
// keep private partial sums in shared memory, and the partials are added in lane order.
constexpr int kTile = 8;
__device__ __forceinline__ void write_cell(const RoundArgs &A, int64_t row, int64_t col0, const double *bt, int p,
                                           int nadj, double deduct)
{
    double s = 0;
    if (nadj > 0) for (int j = 0; j < p; j++) s += bt[j];
    double *o = A.out + row * A.X.n_cols + col0;
    for (int j = 0; j < p; j++) { double v = bt[j]; if (nadj > 0) v = v - deduct * (v / s); o[j] = v; }
}

__global__ void __launch_bounds__(128) task_step_small_kernel(RoundArgs A)
{
    extern __shared__ double sm[];
    const int sub = threadIdx.x & (kTile - 1);
    const int tile_in_cta = threadIdx.x / kTile;
    double *mine = sm + (size_t)threadIdx.x * kStepSlot;              // this lane's partial sums: T[qp], BTsum[p]
    double *tile0 = sm + (size_t)tile_in_cta * kTile * kStepSlot;     // lane 0's area: the reduced sums / X.est
    const unsigned tmask = 0xffu << ((threadIdx.x & 31) & ~(kTile - 1));
    const unsigned i = blockIdx.x * (blockDim.x / kTile) + tile_in_cta;
    if (i >= *A.tlist_n) return;
    const XDev &X = A.X;
    const Tasks &T = A.T;
    const EmOpts &O = A.O;
    const int64_t t = A.tlist[i];
    char *td = T.tdata;
    char *blob = td + T.toff[t];
    TaskState *st = (TaskState *)blob;
    if (!st->alive) return;
    const bool fail = st->fail != 0;
    const int inner_it = st->inner_it + 1;
    const int outer = st->outer + 1;
    __syncwarp(tmask);
    if (sub == 0) { st->inner_it = inner_it; st->inner_total += 1; st->fail = 0; }
    if (!((!O.fixed_iter && !fail) || inner_it >= O.maxiter_bt)) return;         // next estim.BETA iteration  :756-757

    const int e = (int)(t % X.n_em), h = (int)(t / X.n_em);
    const int b = X.em_blocks[e];
    const int q = X.q_off[b + 1] - X.q_off[b], p = X.p_off[b + 1] - X.p_off[b];
    const int qp = q * p, ppad = ppad_of(p);
    const int nc = (int)(A.S.chunk_off[h + 1] - A.S.chunk_off[h]);
    double *Xc = (double *)(blob + 64), *cs = Xc + 2 * qp;
    const int ncell = st->ncell;
    const int64_t *rpos = T.rec_pos + T.run_off[t];

    // ---- estim.X.em(X.est0, BT, Ymat)  :898-922
    for (int k = 0; k < qp + p; k++) mine[k] = 0.0;
    for (int a = sub; a < ncell; a += kTile) {
        const char *rec = td + rpos[a];
        const CellHdr hd = *(const CellHdr *)rec;
        const double *bt = (const double *)(rec + 16);
        const Ent *en = (const Ent *)(rec + 16 + 8 * ppad);
        const double mult = (hd.cell & kPseudo) ? (double)(hd.cell & 0xffff) : 1.0;
        for (int j = 0; j < p; j++) mine[qp + j] += mult * bt[j];                // BTsum = colSums(BETA)  :904
        for (int k = 0; k < hd.m; k++) {
            const Ent ev = en[k];
            double mu = 0.0;                                                      // sumx1 = rowSums(x1)  :907
            for (int j = 0; j < p; j++) mu += bt[j] * Xc[j * q + ev.row];
            const double w = mult * ev.y / (mu > 0 ? mu : 0.1);                  // :908-910
            for (int j = 0; j < p; j++) mine[j * q + ev.row] += bt[j] * Xc[j * q + ev.row] * w;
        }
    }
    __syncwarp(tmask);
    for (int k = sub; k < qp + p; k += kTile) {                                   // partials added in lane order
        double s = 0;
        for (int l = 0; l < kTile; l++) s += tile0[l * kStepSlot + k];
        tile0[k] = s;                                                             // slot k is read and written by this lane only
    }
    __syncwarp(tmask);
    bool xfail = false;
    for (int j = sub; j < p; j += kTile) {
        const double bs = tile0[qp + j];
        const double d = bs > 0 ? bs : 0.1;                                       // :917
        double csum = 0.0;
        for (int r = 0; r < q; r++) { const double v = tile0[j * q + r] / d; tile0[j * q + r] = v; csum += v; }
        const double d2 = csum > 0 ? csum : 0.1;                                  // :919
        for (int r = 0; r < q; r++) {
            double v = tile0[j * q + r] / d2;
            const double x0 = Xc[j * q + r];
            if (csum < 0.00001) v = x0;                                           // :920
            tile0[j * q + r] = v;
            const double er = fabs(v - x0) / (fabs(x0) > O.lim ? x0 : O.lim);    // :1019
            if (!(er < O.maxerr)) xfail = true;
        }
    }
    const bool any_x = __ballot_sync(tmask, xfail) != 0;
    if (sub == 0) st->outer = outer;
    if (!((!O.fixed_iter && !any_x) || outer >= O.maxiter_x)) {                   // :1020-1023: next estim.BETA call
        for (int j = sub; j < p; j += kTile) {
            double s = 0;
            for (int r = 0; r < q; r++) { const double v = tile0[j * q + r]; Xc[j * q + r] = v; s += v; }
            cs[j] = s;
        }
        if (sub == 0) st->inner_it = 0;
        return;
    }
    // ---- de-adjust (:1027-1037) and write BT into the result (:1042)
    const int nadj = st->nadj;
    const double deduct = nadj * O.adjust_esp;
    const int64_t cbase = A.S.chunk_off[h];
    const int64_t col0 = X.p_off[b];
    for (int a = sub; a < ncell; a += kTile) {
        const char *rec = td + rpos[a];
        const uint32_t cw = ((const CellHdr *)rec)->cell;
        if (!(cw & kPseudo)) write_cell(A, cbase + (int64_t)cw, col0, (const double *)(rec + 16), p, nadj, deduct);
    }
    const uint32_t lastw = ((const CellHdr *)(td + rpos[ncell - 1]))->cell;
    if ((lastw & kPseudo) && (lastw & 0xffff) != 0) {
        // cells of the chunk without a record of their own take the pseudo cell's values; the real
        // cells are listed in ascending order, so walk the gaps
        const double *bt = (const double *)(td + rpos[ncell - 1] + 16);
        for (int a = sub; a < ncell; a += kTile) {
            const int lo = a == 0 ? 0 : (int)(((const CellHdr *)(td + rpos[a - 1]))->cell) + 1;
            const int hi = a == ncell - 1 ? nc : (int)(((const CellHdr *)(td + rpos[a]))->cell);
            for (int cc = lo; cc < hi; cc++) write_cell(A, cbase + cc, col0, bt, p, nadj, deduct);
        }
    }
    __syncwarp(tmask);
    if (sub == 0) finish_task(A, t, (int)i, st, q, p, nc);
}

// estim.X.em's sums for a WIDE block, one warp: lane = block row, cells visited in order.  Each lane
// owns its row of the accumulator Xn and (for j = lane) its entry of colSums(BETA): no shuffles,
// and every sum over cells runs in cell order.  ysm[q] is a per-warp scratch for one cell's counts.
__device__ __forceinline__ void xupd_rows(const Tasks &T, const int64_t *rpos, int ncell, int q, int p, int ppad,
                                          const double *Xc, double *Xn, double *bts, double *ysm, int lane)
{
    const char *td = T.tdata;
    for (int k = lane; k < q * p; k += kWarp) Xn[k] = 0.0;
    for (int j = lane; j < p; j += kWarp) bts[j] = 0.0;
    __syncwarp();
    for (int a = 0; a < ncell; a++) {
        const int64_t off = rpos[a];
        const CellHdr hd = *(const CellHdr *)(td + off);
        const double *bt = (const double *)(td + off + 16);
        const Ent *en = (const Ent *)(td + off + 16 + 8 * ppad);
        const double mult = (hd.cell & kPseudo) ? (double)(hd.cell & 0xffff) : 1.0;
        for (int r = lane; r < q; r += kWarp) ysm[r] = 0.0;
        __syncwarp();
        for (int k = lane; k < hd.m; k += kWarp) { const Ent ev = en[k]; ysm[ev.row] = ev.y; }
        for (int j = lane; j < p; j += kWarp) bts[j] += mult * bt[j];            // BTsum = colSums(BETA)  :904
        __syncwarp();
        for (int r = lane; r < q; r += kWarp) {
            const double y = ysm[r];
            if (y != 0.0) {
                double mu = 0.0;                                                  // sumx1 = rowSums(x1)  :907
                for (int j = 0; j < p; j++) mu += bt[j] * Xc[j * q + r];
                const double w = mult * y / (mu > 0 ? mu : 0.1);                 // :908-910
                for (int j = 0; j < p; j++) Xn[j * q + r] += bt[j] * Xc[j * q + r] * w;
            }
        }
        __syncwarp();
    }
}

// The same per BIG task (wide block or many cells): one warp, lane = cell, row-by-row xor-tree sums.
__global__ void __launch_bounds__(128) task_step_big_kernel(RoundArgs A)
{
    extern __shared__ double sm[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    double *Xn = sm + (size_t)warp * (A.xcap + A.pcap + A.qcap);
    double *bts = Xn + A.xcap;
    double *ysm = bts + A.pcap;
    const unsigned i = blockIdx.x * (blockDim.x >> 5) + warp;
    if (i >= *A.tlist_n) return;
    const XDev &X = A.X;
    const Tasks &T = A.T;
    const EmOpts &O = A.O;
    const int64_t t = A.tlist[i];
    char *td = T.tdata;
    char *blob = td + T.toff[t];
    TaskState *st = (TaskState *)blob;
    if (!st->alive) return;
    const bool fail = st->fail != 0;
    const int inner_it = st->inner_it + 1;
    __syncwarp();
    if (lane == 0) { st->inner_it = inner_it; st->inner_total += 1; st->fail = 0; }
    if (!((!O.fixed_iter && !fail) || inner_it >= O.maxiter_bt)) return;

    const int e = (int)(t % X.n_em), h = (int)(t / X.n_em);
    const int b = X.em_blocks[e];
    const int q = X.q_off[b + 1] - X.q_off[b], p = X.p_off[b + 1] - X.p_off[b];
    const int qp = q * p, ppad = ppad_of(p);
    const int nc = (int)(A.S.chunk_off[h + 1] - A.S.chunk_off[h]);
    double *Xc = (double *)(blob + 64), *cs = Xc + 2 * qp;
    const int ncell = st->ncell;
    const int64_t *rpos = T.rec_pos + T.run_off[t];

    xupd_rows(T, rpos, ncell, q, p, ppad, Xc, Xn, bts, ysm, lane);
    for (int j = lane; j < p; j += kWarp) {
        const double d = bts[j] > 0 ? bts[j] : 0.1;                  // :917
        double csum = 0.0;
        for (int r = 0; r < q; r++) { double v = Xn[j * q + r] / d; Xn[j * q + r] = v; csum += v; }
        const double d2 = csum > 0 ? csum : 0.1;                     // :919
        for (int r = 0; r < q; r++) Xn[j * q + r] /= d2;
        if (csum < 0.00001)                                          // :920
            for (int r = 0; r < q; r++) Xn[j * q + r] = Xc[j * q + r];
    }
    __syncwarp();
    bool xfail = false;                                              // :1019
    for (int k = lane; k < qp; k += kWarp) {
        const double x0 = Xc[k];
        const double er = fabs(Xn[k] - x0) / (fabs(x0) > O.lim ? x0 : O.lim);
        if (!(er < O.maxerr)) xfail = true;
    }
    const bool any_x = __any_sync(kFull, xfail);
    const int outer = st->outer + 1;
    __syncwarp();
    if (lane == 0) st->outer = outer;
    if (!((!O.fixed_iter && !any_x) || outer >= O.maxiter_x)) {      // :1020-1023
        for (int k = lane; k < qp; k += kWarp) Xc[k] = Xn[k];
        __syncwarp();
        for (int j = lane; j < p; j += kWarp) {
            double s = 0;
            for (int r = 0; r < q; r++) s += Xn[j * q + r];
            cs[j] = s;
        }
        if (lane == 0) st->inner_it = 0;
        return;
    }
    const int nadj = st->nadj;
    const double deduct = nadj * O.adjust_esp;
    const int64_t cbase = A.S.chunk_off[h];
    const int64_t col0 = X.p_off[b];
    for (int a = lane; a < ncell; a += kWarp) {
        const int64_t off = rpos[a];
        const CellHdr hd = *(const CellHdr *)(td + off);
        const double *bt = (const double *)(td + off + 16);
        double s = 0;
        if (nadj > 0) for (int j = 0; j < p; j++) s += bt[j];
        if (!(hd.cell & kPseudo)) {
            double *o = A.out + (cbase + (int64_t)hd.cell) * X.n_cols + col0;
            for (int j = 0; j < p; j++) { double v = bt[j]; if (nadj > 0) v = v - deduct * (v / s); o[j] = v; }
        } else if ((hd.cell & 0xffff) != 0) {
            int ai = 0;
            uint32_t next_real = ((const CellHdr *)(td + rpos[0]))->cell;
            for (int cc = 0; cc < nc; cc++) {
                if (ai < ncell - 1 && next_real == (uint32_t)cc) {
                    ai++;
                    next_real = ((const CellHdr *)(td + rpos[ai]))->cell;
                    continue;
                }
                double *o = A.out + (cbase + cc) * X.n_cols + col0;
                for (int j = 0; j < p; j++) { double v = bt[j]; if (nadj > 0) v = v - deduct * (v / s); o[j] = v; }
            }
        }
    }
    __syncwarp();
    if (lane == 0) finish_task(A, t, (int)i, st, q, p, nc);
}

// ------------------------------------------------------------------------------------------
// Stragglers.  After the bulk rounds a few per cent of the tasks still iterate (slowly converging
// blocks run hundreds of iterations); a global round per iteration would then cost a kernel launch
// for a handful of cells.  This persistent kernel takes over: ONE WARP per task runs it to the end
// from the state the rounds left (X.est0, counters in TaskState; beta, logNorm in the records).
// gamma runs over the flattened (cell, transcript) pairs, the rest is lane = cell.
__global__ void __launch_bounds__(128) em_finish_kernel(RoundArgs A, unsigned int *next_task)
{
    extern __shared__ double sm[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int per_warp = 2 * A.xcap + 2 * A.pcap + A.qcap;
    double *Xc = sm + (size_t)warp * per_warp;   // current X (X.est0), column-major
    double *Xn = Xc + A.xcap;                    // estim.X.em result (X.est)
    double *cs = Xn + A.xcap;                    // colSums(X)
    double *bts = cs + A.pcap;                   // colSums(BETA)
    double *ysm = bts + A.pcap;                  // one cell's counts by row (wide blocks)

    const XDev &X = A.X;
    const Tasks &T = A.T;
    const EmOpts &O = A.O;
    const unsigned total = *A.tlist_n;

    for (;;) {
        unsigned ti = 0;
        if (lane == 0) ti = atomicAdd(next_task, 1u);
        ti = __shfl_sync(kFull, ti, 0);
        if (ti >= total) break;
        const int64_t t = A.tlist[ti];
        char *td = T.tdata, *sc = T.scr;
        TaskState *st = (TaskState *)(td + T.toff[t]);
        if (!st->alive) continue;
        const int e = (int)(t % X.n_em), h = (int)(t / X.n_em);
        const int b = X.em_blocks[e];
        const int q = X.q_off[b + 1] - X.q_off[b];
        const int p = X.p_off[b + 1] - X.p_off[b];
        const int qp = q * p, ppad = ppad_of(p);
        const int nc = (int)(A.S.chunk_off[h + 1] - A.S.chunk_off[h]);
        const int ncell = st->ncell;
        const int n_round = (ncell + 31) & ~31;
        const int64_t *rpos = T.rec_pos + T.run_off[t];
        int outer = st->outer, inner_total = st->inner_total, inner_it0 = st->inner_it;
        const int nadj = st->nadj;
        const double *sx = (const double *)(td + T.toff[t] + 64);
        for (int k = lane; k < qp; k += kWarp) Xc[k] = sx[k];
        __syncwarp();

        for (;;) {
            // ================= estim.BETA(X.est0, Ymat, beta0)  :726-761
            for (int j = lane; j < p; j += kWarp) {                     // csum = colSums(X)
                double s = 0;
                for (int r = 0; r < q; r++) s += Xc[j * q + r];
                cs[j] = s;
            }
            __syncwarp();
            for (int it = inner_it0 + 1;; it++) {
                // stage 1: gamma over the flattened (cell, transcript) pairs
                const int npairs = ncell * p;
                for (int idx = lane; idx < npairs; idx += kWarp) {
                    const int a = idx / p, j = idx - a * p;
                    const int64_t off = rpos[a];
                    const double lnm = ((const CellHdr *)(td + off))->lnorm;
                    const double bv = ((const double *)(td + off + 16))[j];
                    ((double *)(sc + off + 16))[j] = em_gamma(bv, lnm, O.modify);
                }
                __syncwarp();
                // stage 2: lane = cell
                bool fail = false;
                for (int a = lane; a < n_round; a += kWarp) {
                    if (a < ncell) {
                        const int64_t off = rpos[a];
                        const CellHdr hd = *(const CellHdr *)(td + off);
                        double *bt = (double *)(td + off + 16);
                        const double *g = (const double *)(sc + off + 16);
                        const Ent *en = (const Ent *)(td + off + 16 + 8 * ppad);
                        double *w = (double *)(sc + off + 16 + 8 * ppad);     // one double per entry (stride 2)
                        for (int k = 0; k < hd.m; k++) {
                            const Ent ev = en[k];
                            double mu = 0.0;
                            for (int j = 0; j < p; j++) mu += Xc[j * q + ev.row] * g[j];
                            w[2 * k] = ev.y / (mu > 0 ? mu : 1.0);
                        }
                        for (int j = 0; j < p; j++) {
                            double s = 0.0;
                            for (int k = 0; k < hd.m; k++) s += Xc[j * q + en[k].row] * w[2 * k];
                            const double old = bt[j];
                            double nb = g[j] * s / cs[j];
                            if (nb != nb) nb = 0.0;                      // :754
                            const double er = fabs(nb - old) / (old > O.lim ? old : O.lim);
                            if (!(er < O.maxerr)) fail = true;
                            bt[j] = nb;
                        }
                    }
                }
                inner_total++;
                const bool any_fail = __any_sync(kFull, fail);
                if ((!O.fixed_iter && !any_fail) || it >= O.maxiter_bt) break;
            }
            inner_it0 = 0;
            __syncwarp();

            // ================= estim.X.em(X.est0, BT, Ymat)  :898-922
            if (p > kSmallP || qp + p > kStepSlot) {
                xupd_rows(T, rpos, ncell, q, p, ppad, Xc, Xn, bts, ysm, lane);
            } else {
                for (int j = 0; j < p; j++) {                               // BTsum = colSums(BETA)
                    double s = 0;
                    for (int a = lane; a < ncell; a += kWarp) {
                        const int64_t off = rpos[a];
                        const uint32_t cw = ((const CellHdr *)(td + off))->cell;
                        const double mult = (cw & kPseudo) ? (double)(cw & 0xffff) : 1.0;
                        s += mult * ((const double *)(td + off + 16))[j];
                    }
                    s = warp_sum(s);
                    if (lane == 0) bts[j] = s;
                }
                for (int i = lane; i < qp; i += kWarp) Xn[i] = 0.0;
                __syncwarp();
                for (int a0 = 0; a0 < ncell; a0 += kWarp) {
                    const int a = a0 + lane;
                    const bool valid = a < ncell;
                    const double *bt = nullptr;
                    const Ent *en = nullptr;
                    int cur = 0, end = 0;
                    double mult = 1.0;
                    if (valid) {
                        const int64_t off = rpos[a];
                        const CellHdr hd = *(const CellHdr *)(td + off);
                        bt = (const double *)(td + off + 16);
                        en = (const Ent *)(td + off + 16 + 8 * ppad);
                        end = hd.m;
                        mult = (hd.cell & kPseudo) ? (double)(hd.cell & 0xffff) : 1.0;
                    }
                    for (int r = 0; r < q; r++) {
                        const bool has = valid && cur < end && en[cur].row == r;
                        if (!__any_sync(kFull, has)) continue;
                        double w = 0.0;
                        if (has) {
                            const double y = en[cur++].y;
                            double mu = 0.0;                                 // sumx1 = rowSums(x1)
                            for (int j = 0; j < p; j++) mu += bt[j] * Xc[j * q + r];
                            w = mult * y / (mu > 0 ? mu : 0.1);
                        }
                        for (int j = 0; j < p; j++) {
                            const double xr = Xc[j * q + r];
                            if (xr != 0.0) {                                 // uniform across the warp
                                double v = has ? bt[j] * xr * w : 0.0;      // x2 * Y
                                v = warp_sum(v);
                                if (lane == 0) Xn[j * q + r] += v;
                            }
                        }
                    }
                }
                __syncwarp();
            }
            for (int j = lane; j < p; j += kWarp) {
                const double d = bts[j] > 0 ? bts[j] : 0.1;              // :917
                double csum = 0.0;
                for (int r = 0; r < q; r++) { double v = Xn[j * q + r] / d; Xn[j * q + r] = v; csum += v; }
                const double d2 = csum > 0 ? csum : 0.1;                 // :919
                for (int r = 0; r < q; r++) Xn[j * q + r] /= d2;
                if (csum < 0.00001)                                      // :920
                    for (int r = 0; r < q; r++) Xn[j * q + r] = Xc[j * q + r];
            }
            __syncwarp();
            bool xfail = false;                                          // :1019
            for (int i = lane; i < qp; i += kWarp) {
                const double x0 = Xc[i];
                const double er = fabs(Xn[i] - x0) / (fabs(x0) > O.lim ? x0 : O.lim);
                if (!(er < O.maxerr)) xfail = true;
            }
            outer++;
            const bool any_x = __any_sync(kFull, xfail);
            if ((!O.fixed_iter && !any_x) || outer >= O.maxiter_x) break; // :1020
            for (int i = lane; i < qp; i += kWarp) Xc[i] = Xn[i];        // X.est0 = X.est  :1022
            __syncwarp();
        }

        // ---- de-adjust (:1027-1037) and scatter BT into the result (:1042)
        const double deduct = nadj * O.adjust_esp;
        const int64_t cbase = A.S.chunk_off[h];
        const int64_t col0 = X.p_off[b];
        for (int a = lane; a < ncell; a += kWarp) {
            const int64_t off = rpos[a];
            const CellHdr hd = *(const CellHdr *)(td + off);
            const double *bt = (const double *)(td + off + 16);
            double s = 0;
            if (nadj > 0) for (int j = 0; j < p; j++) s += bt[j];
            if (!(hd.cell & kPseudo)) {
                double *o = A.out + (cbase + (int64_t)hd.cell) * X.n_cols + col0;
                for (int j = 0; j < p; j++) { double v = bt[j]; if (nadj > 0) v = v - deduct * (v / s); o[j] = v; }
            } else if ((hd.cell & 0xffff) != 0) {
                int ai = 0;
                uint32_t next_real = ((const CellHdr *)(td + rpos[0]))->cell;
                for (int cc = 0; cc < nc; cc++) {
                    if (ai < ncell - 1 && next_real == (uint32_t)cc) {
                        ai++;
                        next_real = ((const CellHdr *)(td + rpos[ai]))->cell;
                        continue;
                    }
                    double *o = A.out + (cbase + cc) * X.n_cols + col0;
                    for (int j = 0; j < p; j++) { double v = bt[j]; if (nadj > 0) v = v - deduct * (v / s); o[j] = v; }
                }
            }
        }
        if (lane == 0) {
            st->alive = 0;
            task_done_stats(A.iters, A.stats, t, outer, inner_total, q, p, nc, ncell);
        }
        __syncwarp();
    }
}

// ---- list maintenance ---------------------------------------------------------------------
__global__ void gather_ncell_kernel(Tasks T, const int32_t *tlist, const unsigned int *n, int32_t *ncell_out)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < *n) ncell_out[i] = T.runs[tlist[i]];
}
// record list of the listed tasks, in list order then record order
__global__ void fill_records_kernel(Tasks T, const int32_t *tlist, const unsigned int *n, const int64_t *roff,
                                    int64_t base, int32_t *r_task, int64_t *r_off)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= *n) return;
    const int64_t t = tlist[i];
    const int ncell = T.runs[t];
    const int64_t *rpos = T.rec_pos + T.run_off[t];
    int64_t o = base + roff[i];
    for (int a = 0; a < ncell; a++) { r_task[o + a] = (int32_t)t; r_off[o + a] = rpos[a]; }
}
// stable compaction of a task list by the alive flags: out[scan[i]] = in[i]
__global__ void compact_list_kernel(const int32_t *in, const int32_t *alive, const int64_t *pos, const unsigned int *n,
                                    int32_t *out, int32_t *alive_out)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= *n) return;
    if (alive[i]) { out[pos[i]] = in[i]; alive_out[pos[i]] = 1; }
}
__global__ void set_count_kernel(unsigned int *dst, const int64_t *src) { *dst = (unsigned int)*src; }
// ------------------------------------------------------------------------------------------
// Column sums over cells: the rowSums of step2:167 (rows with sum > 0 survive).  Grid = (tiles of
// kTileCells cells) x (column slices).  A thread owns kFinPerThread fixed columns and adds them for
// the cells of its tile in order; tiles are then added in order by colsum_reduce, so the sum is a
// fixed sequence whatever the GPU count.
constexpr int kTileCells = 128;
constexpr int kFinThreads = 512;
constexpr int kFinPerThread = 8;
constexpr int kFinCols = kFinThreads * kFinPerThread;
__global__ void __launch_bounds__(kFinThreads) colsum_kernel(const double *__restrict__ iso, int64_t n_cells,
                                                             int64_t n_cols, double *__restrict__ partial)
{
    const int64_t c0 = (int64_t)blockIdx.x * kTileCells;
    const int64_t c1 = min(c0 + (int64_t)kTileCells, n_cells);
    const int64_t j0 = (int64_t)blockIdx.y * kFinCols + threadIdx.x;
    double cs[kFinPerThread];
#pragma unroll
    for (int k = 0; k < kFinPerThread; k++) cs[k] = 0.0;
    for (int64_t c = c0; c < c1; c++) {
        const double *row = iso + c * n_cols;
#pragma unroll
        for (int k = 0; k < kFinPerThread; k++) {
            const int64_t j = j0 + (int64_t)k * kFinThreads;
            if (j < n_cols) cs[k] += row[j];
        }
    }
#pragma unroll
    for (int k = 0; k < kFinPerThread; k++) {
        const int64_t j = j0 + (int64_t)k * kFinThreads;
        if (j < n_cols) partial[(int64_t)blockIdx.x * n_cols + j] = cs[k];
    }
}
__global__ void colsum_reduce(const double *__restrict__ partial, int n_tiles, int64_t n_cols,
                              double *__restrict__ colsum)
{
    int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_cols) return;
    double s = 0;
    for (int t = 0; t < n_tiles; t++) s += partial[(int64_t)t * n_cols + j];
    colsum[j] = s;
}

// scasa_genemat (step3:32-46).  One CTA walks cells; per cell the isoform row is read coalesced,
// every column puts its value where its gene row is — a single-isoform gene is a copy (step3:39),
// the first isoform of a multi-isoform gene sums the members in column order (step3:27) — into a
// shared-memory image of the gene row, which is then written out coalesced.
// col_action[j]: >= 0 copy to that gene row; <= -2 first member of gene -(a+2); -1 nothing.
__global__ void __launch_bounds__(1024) gene_smem_kernel(const double *__restrict__ iso, int64_t n_cells, int64_t n_cols,
                                                         const int32_t *__restrict__ col_action,
                                                         const int32_t *__restrict__ gene_ptr,
                                                         const int32_t *__restrict__ gene_cols, int n_genes,
                                                         double *__restrict__ gene)
{
    extern __shared__ double grow[];
    for (int64_t c = blockIdx.x; c < n_cells; c += gridDim.x) {
        const double *row = iso + c * n_cols;
        for (int64_t jb = threadIdx.x; jb < n_cols; jb += 4 * (int64_t)blockDim.x) {
            int a[4];
            double v[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {                     // independent loads first
                const int64_t j = jb + (int64_t)u * blockDim.x;
                a[u] = j < n_cols ? col_action[j] : -1;
                v[u] = j < n_cols ? row[j] : 0.0;
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                if (a[u] >= 0) grow[a[u]] = v[u];
                else if (a[u] <= -2) {
                    const int g = -(a[u] + 2);
                    double sum = v[u];                        // the first member is this column
                    for (int m = gene_ptr[g] + 1; m < gene_ptr[g + 1]; m++) sum += row[gene_cols[m]];
                    grow[g] = sum;
                }
            }
        }
        __syncthreads();
        double *out = gene + c * (int64_t)n_genes;
        for (int g = threadIdx.x; g < n_genes; g += blockDim.x) out[g] = grow[g];
        __syncthreads();
    }
}
// any number of genes: one thread per gene row, members read straight from the isoform row
__global__ void gene_direct_kernel(const double *__restrict__ iso, int64_t n_cells, int64_t n_cols,
                                   const int32_t *__restrict__ gene_ptr, const int32_t *__restrict__ gene_cols,
                                   int n_genes, double *__restrict__ gene)
{
    for (int64_t c = blockIdx.x; c < n_cells; c += gridDim.x) {
        const double *row = iso + c * n_cols;
        double *out = gene + c * (int64_t)n_genes;
        for (int g = threadIdx.x; g < n_genes; g += blockDim.x) {
            double sum = 0;
            for (int m = gene_ptr[g]; m < gene_ptr[g + 1]; m++) sum += row[gene_cols[m]];
            out[g] = sum;
        }
    }
}

}  // namespace scasa
