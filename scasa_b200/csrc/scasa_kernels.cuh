// scasa_kernels.cuh — device side of libscasa_cuda (sm_100a).
//
// Data flow of one scasa_run() (all on one stream, everything resident in HBM):
//
//   [eqclass counts] --pack_cells_kernel--> Y as CSR by cell (rows ascending, duplicates summed)
//   Y --task_count_kernel / poor_fix_kernel / scan--> per (chunk, EM block) "task" extents
//   Y --task_scatter_kernel--> task-major entries (cell-in-chunk, row-in-block, count) + run table
//        + pass-through blocks (nrow == 1, Scasa_Rsource.R:979-983) written straight to the result
//   tasks --em_kernel<PMAX>--> alternating EM per task (Scasa_Rsource.R:1010-1040), one warp per
//        task, lanes = cells with counts in the block; results scattered into the dense result
//   result --finish_kernel--> per-column sums (step2:167) and gene sums (step3:32-46)
//
// Why sparse: a cell with no counts in a block keeps beta = 0 through every iteration of
// estim.BETA and contributes 0 to every sum of estim.X.em, so only (cell, block) pairs with
// counts are touched ("runs").  The two convergence tests stay what the reference has: maxima
// over the whole chunk (Rsource:755-756, :1019-1020) = warp-wide votes here.
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

struct Tasks {                // task t = chunk * n_em + em index
    int32_t *cnt, *runs, *bw;            // [n_tasks] entries, runs, runs * p
    int64_t *ent_off, *run_off, *beta_off; // [n_tasks+1] exclusive scans
    int64_t *ent_cur, *run_cur;          // [n_tasks] cursors of the scatter
    int32_t *ent_key;                    // (cell in chunk << 16) | row in block
    double *ent_val;
    int64_t *run_start;                  // [total runs + 1] first entry of a run
    double *beta, *lnorm;                // per run: p betas / logNorm
};

// ------------------------------------------------------------------------------------------
// crpcount_td's accumulation (Scasa_Rsource.R:528-531, :643-646 and :461-462): per cell, add the
// UMI counts of all eqclasses that map to the same row.  One CTA per cell: (row, count) pairs
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

// Poor 2-column blocks get adjust_esp added to *every* cell of the chunk (Rsource:1004), so a
// non-empty poor task is stored dense: all cells x all rows.  Also bw = runs * p.
__global__ void poor_fix_kernel(XDev X, Sample S, Tasks T, int64_t n_tasks)
{
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tasks) return;
    int h = (int)(t / X.n_em), e = (int)(t % X.n_em);
    int b = X.em_blocks[e];
    int p = X.p_off[b + 1] - X.p_off[b];
    int c = T.cnt[t];
    if (c > 0 && X.poor[b]) {
        int nc = (int)(S.chunk_off[h + 1] - S.chunk_off[h]);
        int q = X.q_off[b + 1] - X.q_off[b];
        T.cnt[t] = nc * q;
        T.runs[t] = nc;
    }
    T.bw[t] = T.runs[t] * p;
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
// Scatter Y into task-major order.  One CTA per chunk; cells are visited in order so the
// entries of a task end up sorted by (cell, row) without atomics — the order of every later
// floating-point sum is therefore fixed (same chunk => same bits on any GPU count).
__global__ void task_scatter_kernel(XDev X, Sample S, Tasks T, double *__restrict__ out)
{
    const int h = blockIdx.x;
    const int64_t c0 = S.chunk_off[h], c1 = S.chunk_off[h + 1];
    const int nc = (int)(c1 - c0);
    const int64_t tbase = (int64_t)h * X.n_em;

    // dense pre-fill of non-empty poor tasks
    for (int k = 0; k < X.n_poor; k++) {
        int e = X.poor_em[k];
        int b = X.em_blocks[e];
        int64_t t = tbase + e;
        if (T.cnt[t] == 0) continue;
        int q = X.q_off[b + 1] - X.q_off[b];
        int64_t e0 = T.ent_off[t], r0 = T.run_off[t];
        for (int i = threadIdx.x; i < nc * q; i += blockDim.x) {
            int c = i / q, r = i % q;
            T.ent_key[e0 + i] = (c << 16) | r;
            T.ent_val[e0 + i] = 0.0;
        }
        for (int c = threadIdx.x; c < nc; c += blockDim.x) T.run_start[r0 + c] = e0 + (int64_t)c * q;
    }
    __syncthreads();

    for (int c = 0; c < nc; c++) {
        const int64_t cell = c0 + c;
        const int64_t s = S.y_start[cell];
        const int n = S.y_len[cell];
        // phase A: place entries (cursors are read-only here)
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            int row = S.y_row[s + i];
            double v = S.y_val[s + i];
            int b = X.row_block[row];
            int e = X.em_index[b];
            if (e < 0) {                       // nrow(X0) == 1: the count is the estimate (Rsource:979-983)
                out[cell * X.n_cols + X.p_off[b]] = v;
                continue;
            }
            int64_t t = tbase + e;
            int rl = row - X.q_off[b];
            if (X.poor[b]) {
                int q = X.q_off[b + 1] - X.q_off[b];
                T.ent_val[T.ent_off[t] + (int64_t)c * q + rl] = v;
                continue;
            }
            int head = i;
            while (head > 0 && X.row_block[S.y_row[s + head - 1]] == b) head--;
            int64_t pos = T.ent_cur[t] + (i - head);
            T.ent_key[pos] = (c << 16) | rl;
            T.ent_val[pos] = v;
        }
        __syncthreads();
        // phase B: the last entry of each run closes it
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            int b = X.row_block[S.y_row[s + i]];
            int e = X.em_index[b];
            if (e < 0 || X.poor[b]) continue;
            bool tail = (i == n - 1) || (X.row_block[S.y_row[s + i + 1]] != b);
            if (!tail) continue;
            int head = i;
            while (head > 0 && X.row_block[S.y_row[s + head - 1]] == b) head--;
            int64_t t = tbase + e;
            int64_t cur = T.ent_cur[t];
            T.run_start[T.run_cur[t]] = cur;
            T.ent_cur[t] = cur + (i - head + 1);
            T.run_cur[t] += 1;
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------
struct EmArgs {
    XDev X;
    Sample S;
    Tasks T;
    const int32_t *grp_em;   // EM-block indices of this launch group, heaviest first
    int grp_n;
    int xcap;                // doubles per warp for one copy of X
    double *out;             // n_cells x n_cols
    int32_t *iters;          // [n_chunks][n_em][2] or null
    unsigned int *next_task;
    unsigned long long *stats; // [0] inner iters, [1] outer iters, [2] inner*(q+2p)*n_c, [3] outer*(q+p)*n_c, [4] tasks, [5] active runs
    int maxiter_x, maxiter_bt, modify, fixed_iter;
    double maxerr, lim, adjust_esp;
};

// One warp = one (chunk, block) task; lanes = cells with counts in the block.
//   estim_chunk   Scasa_Rsource.R:986-1040     estim.BETA  :726-761     estim.X.em  :898-922
template <int PMAX>
__global__ void __launch_bounds__(128) em_kernel(EmArgs A)
{
    extern __shared__ double sm[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int per_warp = 2 * A.xcap + 2 * PMAX + (A.xcap + 7) / 8;
    double *Xc = sm + (size_t)warp * per_warp;   // current X (X.est0), column-major
    double *Xn = Xc + A.xcap;                    // estim.X.em result (X.est)
    double *cs = Xn + A.xcap;                    // colSums(X)
    double *bts = cs + PMAX;                     // colSums(BETA)
    unsigned char *adj = (unsigned char *)(bts + PMAX);

    const XDev &X = A.X;
    const Tasks &T = A.T;
    const unsigned total = (unsigned)A.grp_n * (unsigned)A.S.n_chunks;

    for (;;) {
        unsigned ti = 0;
        if (lane == 0) ti = atomicAdd(A.next_task, 1u);
        ti = __shfl_sync(kFull, ti, 0);
        if (ti >= total) break;
        const int e = A.grp_em[ti / A.S.n_chunks];
        const int h = ti % A.S.n_chunks;
        const int64_t t = (int64_t)h * X.n_em + e;
        const int64_t e0 = T.ent_off[t], e1 = T.ent_off[t + 1];
        const int b = X.em_blocks[e];
        const int q = X.q_off[b + 1] - X.q_off[b];
        const int p = X.p_off[b + 1] - X.p_off[b];
        const int nc = (int)(A.S.chunk_off[h + 1] - A.S.chunk_off[h]);
        if (e0 == e1) {
            // all-zero block: the reference still runs one inner and one outer iteration and
            // returns zeros (the result buffer is pre-zeroed)
            if (lane == 0) {
                int o = A.fixed_iter ? A.maxiter_x : 1;
                int in = A.fixed_iter ? A.maxiter_x * A.maxiter_bt : 1;
                if (A.iters) { A.iters[2 * t] = o; A.iters[2 * t + 1] = in; }
                atomicAdd(&A.stats[0], (unsigned long long)in);
                atomicAdd(&A.stats[1], (unsigned long long)o);
                atomicAdd(&A.stats[2], (unsigned long long)in * (q + 2 * p) * nc);
                atomicAdd(&A.stats[3], (unsigned long long)o * (q + p) * nc);
            }
            continue;
        }
        const int64_t r0 = T.run_off[t];
        const int n_act = (int)(T.run_off[t + 1] - r0);
        const int n_round = (n_act + 31) & ~31;
        double *bt = T.beta + T.beta_off[t];
        double *ln = T.lnorm + r0;
        const int64_t *rs_ = T.run_start + r0;
        const int32_t *key = T.ent_key;
        double *val = T.ent_val;
        const double *X0 = X.x + X.x_off[b];
        const int qp = q * p;

        for (int i = lane; i < qp; i += kWarp) Xc[i] = X0[i];          // X.est0 = X0   :1011
        for (int i = lane; i < q; i += kWarp) adj[i] = 0;
        __syncwarp();

        // ---- poor-condition adjustment  :989-1008
        int nadj = 0;
        if (X.poor[b]) {
            // rs = rowSums(Ymat) / sum(Ymat); dense task: entry (a, r) at e0 + a*q + r
            for (int r = 0; r < q; r++) {
                double s = 0;
                for (int a = lane; a < n_act; a += kWarp) s += val[e0 + (int64_t)a * q + r];
                s = warp_sum(s);
                if (lane == 0) Xn[r] = s;
            }
            __syncwarp();
            double tot = 0;
            for (int r = 0; r < q; r++) tot += Xn[r];
            int best = -1;
            double bestv = -INFINITY;
            for (int j = 0; j < p; j++) {                // which.max(cos_sim(X0[,j], rs)), NaN skipped
                double ab = 0, aa = 0, bb = 0;
                for (int r = 0; r < q; r++) {
                    double xv = Xc[j * q + r], rv = Xn[r] / tot;
                    ab += xv * rv; aa += xv * xv; bb += rv * rv;
                }
                double sim = ab / sqrt(aa * bb);
                if (sim == sim && sim > bestv) { bestv = sim; best = j; }
            }
            __syncwarp();
            if (best >= 0) {
                for (int r = 0; r < q; r++) if (Xc[best * q + r] > 0) nadj++;
                if (nadj > 0) {
                    for (int i = lane; i < q; i += kWarp) adj[i] = Xc[best * q + i] > 0;
                    __syncwarp();
                    for (int64_t i = e0 + lane; i < e1; i += kWarp)
                        if (adj[key[i] & 0xffff]) val[i] += A.adjust_esp;   // Ymat[adjID,] += adjust_esp
                }
            }
            __syncwarp();
        }

        // ---- beta0 = Ysum / p, logNorm = digamma(Ysum + 1e-8)   :1012-1013, :732-733
        for (int a = lane; a < n_act; a += kWarp) {
            double ys = 0;
            for (int64_t i = rs_[a]; i < rs_[a + 1]; i++) ys += val[i];
            ln[a] = digamma_pos(ys + 0.00000001);
            double b0 = ys / p;
            for (int j = 0; j < p; j++) bt[(int64_t)a * p + j] = b0;
        }
        __syncwarp();

        int outer = 0, inner_total = 0;
        double bl[PMAX], g[PMAX], acc[PMAX];
        for (;;) {
            // ================= estim.BETA(X.est0, Ymat, beta0)  :726-761
            for (int j = lane; j < p; j += kWarp) {                     // csum = colSums(X)
                double s = 0;
                for (int r = 0; r < q; r++) s += Xc[j * q + r];
                cs[j] = s;
            }
            __syncwarp();
            for (int it = 1;; it++) {
                bool fail = false;
                for (int a = lane; a < n_round; a += kWarp) {
                    if (a < n_act) {
                        const double lnm = ln[a];
#pragma unroll
                        for (int j = 0; j < PMAX; j++)
                            if (j < p) {
                                bl[j] = bt[(int64_t)a * p + j];
                                g[j] = A.modify ? exp(digamma_pos(bl[j] + 0.00000001) - lnm) : bl[j];
                                acc[j] = 0.0;
                            }
                        for (int64_t i = rs_[a]; i < rs_[a + 1]; i++) {
                            const int r = key[i] & 0xffff;
                            const double y = val[i];
                            double mu = 0.0;
#pragma unroll
                            for (int j = 0; j < PMAX; j++)
                                if (j < p) mu += Xc[j * q + r] * g[j];    // rowSums(t(t(X)*gamma0))
                            const double w = y / (mu > 0 ? mu : 1.0);
#pragma unroll
                            for (int j = 0; j < PMAX; j++)
                                if (j < p) acc[j] += Xc[j * q + r] * g[j] * w;
                        }
#pragma unroll
                        for (int j = 0; j < PMAX; j++)
                            if (j < p) {
                                double nb = acc[j] / cs[j];
                                if (nb != nb) nb = 0.0;                      // :754
                                double er = fabs(nb - bl[j]) / (bl[j] > A.lim ? bl[j] : A.lim);
                                if (!(er < A.maxerr)) fail = true;
                                bt[(int64_t)a * p + j] = nb;
                            }
                    }
                }
                inner_total++;
                const bool any_fail = __any_sync(kFull, fail);
                if ((!A.fixed_iter && !any_fail) || it >= A.maxiter_bt) break;
            }
            __syncwarp();

            // ================= estim.X.em(X.est0, BT, Ymat)  :898-922
            for (int j = 0; j < p; j++) {                               // BTsum = colSums(BETA)
                double s = 0;
                for (int a = lane; a < n_act; a += kWarp) s += bt[(int64_t)a * p + j];
                s = warp_sum(s);
                if (lane == 0) bts[j] = s;
            }
            for (int i = lane; i < qp; i += kWarp) Xn[i] = 0.0;
            __syncwarp();
            for (int a0 = 0; a0 < n_act; a0 += kWarp) {
                const int a = a0 + lane;
                const bool valid = a < n_act;
                int64_t cur = 0, end = 0;
                if (valid) {
                    cur = rs_[a]; end = rs_[a + 1];
#pragma unroll
                    for (int j = 0; j < PMAX; j++)
                        if (j < p) bl[j] = bt[(int64_t)a * p + j];
                }
                for (int r = 0; r < q; r++) {
                    const bool has = valid && cur < end && (key[cur] & 0xffff) == r;
                    if (!__any_sync(kFull, has)) continue;
                    double w = 0.0;
                    if (has) {
                        const double y = val[cur++];
                        double mu = 0.0;                                 // sumx1 = rowSums(x1)
#pragma unroll
                        for (int j = 0; j < PMAX; j++)
                            if (j < p) mu += bl[j] * Xc[j * q + r];
                        w = y / (mu > 0 ? mu : 0.1);
                    }
#pragma unroll
                    for (int j = 0; j < PMAX; j++)
                        if (j < p) {
                            const double xr = Xc[j * q + r];
                            if (xr != 0.0) {                              // uniform across the warp
                                double v = has ? bl[j] * xr * w : 0.0;   // x2 * Y
                                v = warp_sum(v);
                                if (lane == 0) Xn[j * q + r] += v;
                            }
                        }
                }
            }
            __syncwarp();
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
                double x0 = Xc[i];
                double er = fabs(Xn[i] - x0) / (fabs(x0) > A.lim ? x0 : A.lim);
                if (!(er < A.maxerr)) xfail = true;
            }
            outer++;
            const bool any_x = __any_sync(kFull, xfail);
            if ((!A.fixed_iter && !any_x) || outer >= A.maxiter_x) break; // :1020
            for (int i = lane; i < qp; i += kWarp) Xc[i] = Xn[i];        // X.est0 = X.est  :1022
            __syncwarp();
        }

        // ---- de-adjust (:1027-1037) and scatter BT into the result (:1042)
        const double deduct = nadj * A.adjust_esp;
        const int64_t cbase = A.S.chunk_off[h];
        const int64_t col0 = X.p_off[b];
        for (int a = lane; a < n_act; a += kWarp) {
            const int64_t cell = cbase + (key[rs_[a]] >> 16);
            double s = 0;
            if (nadj > 0)
                for (int j = 0; j < p; j++) s += bt[(int64_t)a * p + j];
            for (int j = 0; j < p; j++) {
                double v = bt[(int64_t)a * p + j];
                if (nadj > 0) v = v - deduct * (v / s);
                A.out[cell * X.n_cols + col0 + j] = v;
            }
        }
        if (lane == 0) {
            if (A.iters) { A.iters[2 * t] = outer; A.iters[2 * t + 1] = inner_total; }
            atomicAdd(&A.stats[0], (unsigned long long)inner_total);
            atomicAdd(&A.stats[1], (unsigned long long)outer);
            atomicAdd(&A.stats[2], (unsigned long long)inner_total * (q + 2 * p) * nc);
            atomicAdd(&A.stats[3], (unsigned long long)outer * (q + p) * nc);
            atomicAdd(&A.stats[4], 1ull);
            atomicAdd(&A.stats[5], (unsigned long long)n_act);
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------
// Column sums over cells (the rowSums of step2:167) and scasa_genemat (step3:32-46).
// One CTA per tile of kTileCells cells, cells visited in order; thread = fixed set of columns,
// so partial[tile][col] is a sequential sum; tiles are then added in order by colsum_reduce.
constexpr int kTileCells = 128;
__global__ void finish_kernel(const double *__restrict__ iso, int64_t n_cells, int64_t n_cols,
                              double *__restrict__ partial,
                              const int32_t *__restrict__ gene_of_col, const int32_t *__restrict__ gene_ptr,
                              const int32_t *__restrict__ gene_cols, int n_genes, double *__restrict__ gene)
{
    const int64_t c0 = (int64_t)blockIdx.x * kTileCells;
    const int64_t c1 = min(c0 + (int64_t)kTileCells, n_cells);
    for (int64_t j = threadIdx.x; j < n_cols; j += blockDim.x) {
        double s = 0;
        for (int64_t c = c0; c < c1; c++) s += iso[c * n_cols + j];
        partial[(int64_t)blockIdx.x * n_cols + j] = s;
    }
    if (gene == nullptr) return;
    for (int64_t c = c0; c < c1; c++) {
        const double *row = iso + c * n_cols;
        double *grow = gene + c * (int64_t)n_genes;
        for (int gi = threadIdx.x; gi < n_genes; gi += blockDim.x) {
            int k0 = gene_ptr[gi], k1 = gene_ptr[gi + 1];
            double s = 0;                                   // colSums in isoform order (step3:27)
            if (k1 - k0 == 1) s = row[gene_cols[k0]];       // single isoform: a copy (step3:39)
            else for (int k = k0; k < k1; k++) s += row[gene_cols[k]];
            grow[gi] = s;
        }
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

}  // namespace scasa
