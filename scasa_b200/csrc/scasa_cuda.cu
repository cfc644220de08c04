// scasa_cuda.cu — host side of libscasa_cuda: context, staging, launch orchestration, C ABI.
// See include/scasa_cuda.h for the contract and scasa_kernels.cuh for the device code.
#include "scasa_cuda.h"
#include "scasa_kernels.cuh"
#include "scasa_host.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <memory>
#include <mutex>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <numeric>
#include <string>
#include <thread>
#include <vector>

using namespace scasa;

namespace {

thread_local std::string g_err;   // errors raised before a context exists

}  // namespace
void scasa::set_global_error(const std::string &m) { g_err = m; }
namespace {

struct CudaErr { cudaError_t e; const char *what; };
#define CK(call)                                                       \
    do {                                                               \
        cudaError_t e_ = (call);                                       \
        if (e_ != cudaSuccess) throw CudaErr{e_, #call};               \
    } while (0)
struct ArgErr { int code; std::string msg; };

template <class T> struct DBuf {   // grow-only device buffer
    T *p = nullptr;
    size_t cap = 0;
    void ensure(size_t n) {
        if (n <= cap) return;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = n + n / 8 + 16;
        cudaError_t e = cudaMalloc(&p, want * sizeof(T));
        if (e != cudaSuccess) { p = nullptr; throw CudaErr{e, "cudaMalloc"}; }
        cap = want;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class V> void upload(const V &v, cudaStream_t s) {
        ensure(v.size());
        if (!v.empty()) CK(cudaMemcpyAsync(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, s));
    }
};

}  // namespace

// Worker threads of the expansion: one pool per context pair (a context and its scasa_quant twin
// share it, so the two never oversubscribe the cores), kept alive between calls.  Jobs (batches of
// rows) queue up first-in first-out; the workers claim rows of the job at the head in small chunks,
// so a batch is balanced across threads whatever its rows look like.
class ExpandPool {
public:
    struct Job {
        double *dst; int64_t n_cols, n_rows; const int64_t *start; const int32_t *nnz; const int32_t *cols; const double *vals;
        std::atomic<int64_t> next{0};
        int64_t finished = 0;
        bool done = false;
    };
    explicit ExpandPool(int n_threads) : n_threads_(n_threads)
    {
        for (int t = 0; t < n_threads; t++) th_.emplace_back([this] { work(); });
    }
    ~ExpandPool()
    {
        { std::lock_guard<std::mutex> lk(m_); stop_ = true; }
        cv_work_.notify_all();
        for (auto &t : th_) t.join();
    }
    int threads() const { return n_threads_; }
    std::shared_ptr<Job> post(double *dst, int64_t n_cols, int64_t n_rows, const int64_t *start, const int32_t *nnz,
                              const int32_t *cols, const double *vals)
    {
        auto j = std::make_shared<Job>();
        j->dst = dst; j->n_cols = n_cols; j->n_rows = n_rows; j->start = start; j->nnz = nnz; j->cols = cols; j->vals = vals;
        {
            std::lock_guard<std::mutex> lk(m_);
            queue_.push_back(j);
        }
        cv_work_.notify_all();
        return j;
    }
    void wait(const std::shared_ptr<Job> &j)
    {
        if (!j) return;
        std::unique_lock<std::mutex> lk(m_);
        cv_done_.wait(lk, [&] { return j->done; });
    }

private:
    std::mutex m_;
    std::condition_variable cv_work_, cv_done_;
    std::deque<std::shared_ptr<Job>> queue_;
    bool stop_ = false;
    int n_threads_;
    std::vector<std::thread> th_;
    void work()
    {
        for (;;) {
            std::shared_ptr<Job> j;
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_work_.wait(lk, [&] { return stop_ || !queue_.empty(); });
                if (queue_.empty()) return;
                j = queue_.front();
            }
            int64_t mine = 0;
            for (;;) {
                const int64_t r = j->next.fetch_add(8);
                if (r >= j->n_rows) break;
                const int64_t e = std::min(r + 8, j->n_rows);
                scasa::expand_rows(j->dst, j->n_cols, r, e, j->start, j->nnz, j->cols, j->vals);
                mine += e - r;
            }
            std::lock_guard<std::mutex> lk(m_);
            if (!queue_.empty() && queue_.front() == j) queue_.pop_front();      // exhausted: the next job becomes the head
            j->finished += mine;
            if (j->finished == j->n_rows && !j->done) { j->done = true; cv_done_.notify_all(); }
        }
    }
};

struct scasa_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[8] = {};
    int n_sm = 148;
    std::string err;
    scasa_opts_t opts;

    // ---- X-matrix (host copies + resident device copies)
    int B = 0, n_em = 0;
    int max_em_q = 0;                // most rows of an EM block
    int64_t n_rows = 0, n_cols = 0;
    std::vector<int32_t> q_off, p_off, em_blocks, em_index, row_block, poor_em, em_order;
    std::vector<int64_t> x_off;
    std::vector<double> x;
    std::vector<uint8_t> poor;
    int64_t poor_rows = 0;           // sum of q over poor blocks
    DBuf<int32_t> d_q_off, d_p_off, d_em_blocks, d_em_order, d_row_col;
    DBuf<int2> d_row_info;
    DBuf<uint32_t> d_scatter_scratch;
    DBuf<int64_t> d_x_off;
    DBuf<double> d_x;
    DBuf<uint8_t> d_poor;
    scasa::CrpIndex crp;             // eqclass -> row map (host), built when CRP is given

    // ---- staged sample
    bool staged = false, have_eq = false, ran = false;
    int64_t n_cells = 0, nnz = 0, n_eq = 0;
    int n_chunks = 0;
    std::vector<int64_t> chunk_off;
    DBuf<int64_t> d_chunk_off, d_y_start, d_cell_ptr;
    DBuf<int32_t> d_cell_chunk, d_y_len, d_y_row, d_eq_id, d_eq_count, d_eq_row;
    DBuf<double> d_y_val;
    int64_t max_cell_entries = 0;

    // ---- task arrays
    DBuf<int32_t> d_cnt, d_runs, d_tbytes, d_iters, d_list[kLists];
    DBuf<int64_t> d_toff, d_scan_sums, d_totals;
    DBuf<char> d_tdata;
    DBuf<unsigned int> d_next, d_counters;
    cudaStream_t cls_stream[kLists] = {};
    cudaEvent_t ev_fork = nullptr, ev_fill = nullptr, ev_join[kLists] = {}, ev_cls[kLists][2] = {};
    // work classes of the EM kernel: warps per task, pairs and slot limits, resident CTAs per SM
    int cls_warps[kClasses] = {1, 2, 4, 8};
    int cls_pairs[kClasses] = {96, 256, 512, 1 << 30};
    int cls_bytes[kClasses] = {6144, 12288, 16384, 1 << 30};
    int cls_cta_threads[kClasses] = {128, 256, 256, 256};
    int cls_ctas_per_sm[kClasses] = {8, 4, 2, 2};
    int pair_max = 16;                // tiny tasks two per warp (em_pair_kernel); SCASA_EM_PAIR=0 turns it off
    int pair_ctas_per_sm = 8;
    DBuf<unsigned long long> d_stats;
    DBuf<int> d_flag;

    // ---- results
    DBuf<double> d_iso, d_partial, d_colsum, d_colsum_parts, d_gene;
    int run_parts = 0;           // ranges scasa_run cuts the sample into (0 = automatic)
    // zero-suppressed fetch: two batches in flight (device pairs -> pinned staging -> threaded expansion)
    struct FetchBuf {
        DBuf<int32_t> d_cols, d_nnz;
        DBuf<double> d_vals;
        DBuf<int64_t> d_start;
        DBuf<unsigned long long> d_cursor;
        int32_t *h_cols = nullptr, *h_nnz = nullptr;
        double *h_vals = nullptr;
        int64_t *h_start = nullptr;
        unsigned long long *h_cursor = nullptr;
        size_t cap = 0, rows = 0;
        cudaEvent_t done = nullptr;
    } fb[2];
    scasa_ctx *twin = nullptr;   // second context on the same device: scasa_quant overlaps its slices on the two
    scasa_ctx *pool_owner = nullptr;                            // a twin expands with its parent's worker pool
    std::unique_ptr<ExpandPool> pool;                           // host threads of the zero-suppressed fetch
    std::mutex pool_mutex;
    std::vector<int32_t> gene_of_col_host;
    std::atomic<long long> wire_h2d{0}, wire_d2h{0};   // bytes that crossed PCIe since the last scasa_transfer_bytes(reset)
    int host_threads = 0;        // 0 = hardware concurrency
    int fetch_dense_copy = 0;    // 1 = plain dense D2H (SCASA_FETCH_DENSE)
    int n_genes = 0;
    DBuf<int32_t> d_gene_ptr, d_gene_cols;
    scasa_timing_t timing{};
};

namespace {

int fail(scasa_ctx *c, int code, const std::string &m)
{
    if (c) c->err = m; else g_err = m;
    return code;
}

template <class F> int guarded(scasa_ctx *c, F &&f)
{
    try {
        if (c) CK(cudaSetDevice(c->device));
        f();
        return SCASA_OK;
    } catch (const CudaErr &e) {
        char buf[512];
        snprintf(buf, sizeof buf, "CUDA error %d (%s) at %s", (int)e.e, cudaGetErrorString(e.e), e.what);
        cudaGetLastError();
        return fail(c, e.e == cudaErrorMemoryAllocation ? SCASA_ERR_NOMEM : SCASA_ERR_CUDA, buf);
    } catch (const ArgErr &e) {
        return fail(c, e.code, e.msg);
    } catch (const std::bad_alloc &) {
        return fail(c, SCASA_ERR_NOMEM, "host allocation failed");
    } catch (const std::exception &e) {
        return fail(c, SCASA_ERR_ARG, e.what());
    }
}

// poorCondX (Scasa_Rsource.R:957-971): a 2-column block is poor unless some pairwise cosine
// similarity (self pairs included) is < 0.99.  long double sums like R's sum().
bool is_poor(const double *X0, int q, int p)
{
    if (p != 2) return false;
    for (int i = 0; i < 2; i++)
        for (int k = 0; k < 2; k++) {
            long double ab = 0, aa = 0, bb = 0;
            for (int r = 0; r < q; r++) {
                long double a = X0[(size_t)k * q + r], b = X0[(size_t)i * q + r];
                ab += a * b; aa += a * a; bb += b * b;
            }
            double sim = (double)ab / std::sqrt((double)aa * (double)bb);
            if (sim < 0.99) return false;
        }
    return true;
}

XDev xdev(scasa_ctx *c)
{
    XDev X;
    X.n_blocks = c->B; X.n_em = c->n_em; X.n_rows = c->n_rows; X.n_cols = c->n_cols;
    X.q_off = c->d_q_off.p; X.p_off = c->d_p_off.p; X.x_off = c->d_x_off.p; X.x = c->d_x.p;
    X.em_blocks = c->d_em_blocks.p;
    X.poor = c->d_poor.p;
    X.em_order = c->d_em_order.p; X.row_info = c->d_row_info.p; X.row_col = c->d_row_col.p;
    return X;
}
// chunks [h0, h1) of the sample staged in c
Sample sdev(scasa_ctx *c, int h0, int h1)
{
    const int64_t c0 = c->chunk_off[(size_t)h0], c1 = c->chunk_off[(size_t)h1];
    Sample S;
    S.n_cells = c1 - c0; S.n_chunks = h1 - h0; S.cell_base = c0; S.chunk_base = h0;
    S.chunk_off = c->d_chunk_off.p + h0;
    S.cell_chunk = c->d_cell_chunk.p + c0; S.y_start = c->d_y_start.p + c0; S.y_len = c->d_y_len.p + c0;
    S.y_row = c->d_y_row.p; S.y_val = c->d_y_val.p;
    return S;
}
Tasks tdev(scasa_ctx *c)
{
    Tasks T;
    T.cnt = c->d_cnt.p; T.runs = c->d_runs.p; T.tbytes = c->d_tbytes.p;
    T.toff = c->d_toff.p;
    T.tdata = c->d_tdata.p;
    for (int k = 0; k < kLists; k++) T.list[k] = c->d_list[k].p;
    T.counters = c->d_counters.p;
    return T;
}

void exclusive_scan(scasa_ctx *c, const int32_t *in, int64_t n, int64_t *out, int64_t *total_slot, int &launches)
{
    const int64_t per = (int64_t)kScanThreads * kScanItems;
    const int64_t nb = (n + per - 1) / per;
    c->d_scan_sums.ensure((size_t)nb + 1);
    scan_block_sums<<<(unsigned)nb, kScanThreads, 0, c->stream>>>(in, n, c->d_scan_sums.p);
    scan_sums<<<1, kScanThreads, 0, c->stream>>>(c->d_scan_sums.p, nb, total_slot);
    scan_add_back<<<(unsigned)nb, kScanThreads, 0, c->stream>>>(in, n, c->d_scan_sums.p, total_slot, out);
    launches += 3;
}

// column sums (+ gene sums) of a resident isoform matrix
// column sums (+ gene sums) of n_cells rows of an isoform matrix resident at iso; w supplies the
// stream, the scratch for the tile sums and the gene map
void launch_finish(scasa_ctx *w, const double *iso, int64_t n_cells, double *gene, double *colsum, int &launches)
{
    cudaStream_t s = w->stream;
    const int n_tiles = (int)((n_cells + kTileCells - 1) / kTileCells);
    w->d_partial.ensure((size_t)n_tiles * w->n_cols);
    colsum_kernel<<<dim3(n_tiles, (unsigned)((w->n_cols + kFinCols - 1) / kFinCols)), kFinThreads, 0, s>>>(
        iso, n_cells, w->n_cols, w->d_partial.p);
    colsum_reduce<<<(unsigned)((w->n_cols + 255) / 256), 256, 0, s>>>(w->d_partial.p, n_tiles, w->n_cols, colsum);
    launches += 2;
    if (gene) {
        const int tiles = (w->n_genes + kGeneThreads * kGeneIlp - 1) / (kGeneThreads * kGeneIlp);
        if (n_cells * tiles >= ((int64_t)1 << 31)) throw ArgErr{SCASA_ERR_UNSUPPORTED, "too many cells x genes for one launch"};
        gene_gather_kernel<<<(unsigned)(n_cells * tiles), kGeneThreads, 0, s>>>(iso, n_cells, w->n_cols, w->d_gene_ptr.p,
                                                                               w->d_gene_cols.p, w->n_genes, tiles, gene);
        launches++;
    }
    CK(cudaGetLastError());
}

void check_xmat(const scasa_xmat_t *m, const char *what)
{
    if (!m || m->n_blocks <= 0 || !m->q_off || !m->p_off || !m->x_off || !m->x)
        throw ArgErr{SCASA_ERR_ARG, std::string(what) + ": null or empty block list"};
    if (m->q_off[0] != 0 || m->p_off[0] != 0 || m->x_off[0] != 0)
        throw ArgErr{SCASA_ERR_ARG, std::string(what) + ": offsets must start at 0"};
    for (int k = 0; k < m->n_blocks; k++) {
        int64_t q = m->q_off[k + 1] - m->q_off[k], p = m->p_off[k + 1] - m->p_off[k];
        if (q <= 0 || p <= 0 || m->x_off[k + 1] - m->x_off[k] != q * p)
            throw ArgErr{SCASA_ERR_ARG, std::string(what) + ": block " + std::to_string(k) + " has inconsistent offsets"};
    }
}

void stage_common(scasa_ctx *c, int64_t n_cells, int32_t n_chunks, const int64_t *chunk_off)
{
    if (n_cells <= 0 || n_chunks <= 0 || !chunk_off) throw ArgErr{SCASA_ERR_ARG, "no cells / no chunks"};
    if (chunk_off[0] != 0 || chunk_off[n_chunks] != n_cells)
        throw ArgErr{SCASA_ERR_ARG, "chunk_off must cover [0, n_cells]"};
    std::vector<int32_t> cell_chunk((size_t)n_cells);
    for (int h = 0; h < n_chunks; h++) {
        int64_t a = chunk_off[h], b = chunk_off[h + 1];
        if (b < a) throw ArgErr{SCASA_ERR_ARG, "chunk_off must be non-decreasing"};
        if (b - a >= 65536) throw ArgErr{SCASA_ERR_UNSUPPORTED, "a chunk may hold at most 65535 cells"};
        for (int64_t i = a; i < b; i++) cell_chunk[(size_t)i] = h;
    }
    if ((int64_t)n_chunks * c->n_em >= (int64_t)1 << 31)
        throw ArgErr{SCASA_ERR_UNSUPPORTED, "n_chunks * n_em_blocks must stay below 2^31 per call"};
    c->n_cells = n_cells; c->n_chunks = n_chunks;
    c->chunk_off.assign(chunk_off, chunk_off + n_chunks + 1);
    c->d_chunk_off.upload(c->chunk_off, c->stream);
    c->d_cell_chunk.upload(cell_chunk, c->stream);
    c->staged = true; c->ran = false;
}

template <int W, typename R> void launch_em_t(const EmArgs &A, unsigned grid, int threads, size_t smem, cudaStream_t s)
{
    CK(cudaFuncSetAttribute(em_task_kernel<W, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    em_task_kernel<W, R><<<grid, threads, smem, s>>>(A);
}
template <int W> void launch_em(const EmArgs &A, unsigned grid, int threads, size_t smem, cudaStream_t s, bool fp32)
{
    if (fp32) launch_em_t<W, float>(A, grid, threads, smem, s); else launch_em_t<W, double>(A, grid, threads, smem, s);
}

// ---- host side of the zero-suppressed fetch: expand_rows() is in scasa_host.cpp -------------------
// the pool of a context: its own, or its parent's when it is a scasa_quant twin
ExpandPool &expand_pool(scasa_ctx *c, int nthr)
{
    scasa_ctx *owner = c->pool_owner ? c->pool_owner : c;
    std::lock_guard<std::mutex> lk(owner->pool_mutex);
    if (!owner->pool || owner->pool->threads() != nthr) owner->pool.reset(new ExpandPool(nthr));
    return *owner->pool;
}

void fetch_buf_ensure(scasa_ctx *c, scasa_ctx::FetchBuf &b, size_t rows, size_t cap)
{
    if (b.rows >= rows && b.cap >= cap) return;
    if (b.h_cols) { cudaFreeHost(b.h_cols); cudaFreeHost(b.h_vals); cudaFreeHost(b.h_start); cudaFreeHost(b.h_nnz); cudaFreeHost(b.h_cursor); }
    b.h_cols = nullptr; b.rows = 0; b.cap = 0;
    b.d_cols.ensure(cap); b.d_vals.ensure(cap); b.d_start.ensure(rows); b.d_nnz.ensure(rows); b.d_cursor.ensure(1);
    CK(cudaHostAlloc((void **)&b.h_cols, cap * sizeof(int32_t), cudaHostAllocDefault));
    CK(cudaHostAlloc((void **)&b.h_vals, cap * sizeof(double), cudaHostAllocDefault));
    CK(cudaHostAlloc((void **)&b.h_start, rows * sizeof(int64_t), cudaHostAllocDefault));
    CK(cudaHostAlloc((void **)&b.h_nnz, rows * sizeof(int32_t), cudaHostAllocDefault));
    CK(cudaHostAlloc((void **)&b.h_cursor, sizeof(unsigned long long), cudaHostAllocDefault));
    if (!b.done) CK(cudaEventCreateWithFlags(&b.done, cudaEventDisableTiming));
    b.rows = rows; b.cap = cap;
}

// Dense (n_rows x n_cols) device matrix -> dense host matrix, sent as (column, value) pairs in
// batches of rows: while the host threads expand batch k, the GPU compacts and copies batch k+1.
void fetch_dense(scasa_ctx *c, const double *d_mat, int64_t n_rows, int64_t n_cols, double *out)
{
    cudaStream_t s = c->stream;
    if (c->fetch_dense_copy || c->host_threads < 0 || n_rows * n_cols < (1 << 20)) {
        CK(cudaMemcpyAsync(out, d_mat, (size_t)(n_rows * n_cols) * sizeof(double), cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        c->wire_d2h += (long long)(n_rows * n_cols) * 8;
        return;
    }
    int nthr = c->host_threads > 0 ? c->host_threads : (int)std::thread::hardware_concurrency();
    nthr = std::max(1, std::min(nthr, 256));
    const int64_t batch = std::max<int64_t>(64, std::min<int64_t>(n_rows, ((int64_t)48 << 20) / n_cols));   // ~48 M cells of the matrix per batch
    const size_t cap = (size_t)(batch * n_cols / 8 + 1024);     // pairs per batch: 12 % density fits (the results are ~6 % dense); denser batches are copied plainly
    for (auto &b : c->fb) fetch_buf_ensure(c, b, (size_t)batch, cap);
    const int64_t n_batches = (n_rows + batch - 1) / batch;
    auto launch = [&](int64_t k) {
        scasa_ctx::FetchBuf &b = c->fb[k & 1];
        const int64_t r0 = k * batch, nr = std::min(batch, n_rows - r0);
        CK(cudaMemsetAsync(b.d_cursor.p, 0, sizeof(unsigned long long), s));
        const int grid = (int)std::min<int64_t>(nr, (int64_t)c->n_sm * 3);
        compact_rows_kernel<<<grid, kCompThreads, 0, s>>>(d_mat + r0 * n_cols, nr, n_cols, b.d_cursor.p, (unsigned long long)b.cap,
                                                         b.d_start.p, b.d_nnz.p, b.d_cols.p, b.d_vals.p);
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(b.h_cursor, b.d_cursor.p, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(b.h_start, b.d_start.p, nr * sizeof(int64_t), cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(b.h_nnz, b.d_nnz.p, nr * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));       // the pair count sizes the copy (the other batch is being expanded meanwhile)
        const size_t used = (size_t)std::min<unsigned long long>(*b.h_cursor, b.cap);
        if (used) {
            CK(cudaMemcpyAsync(b.h_cols, b.d_cols.p, used * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
            CK(cudaMemcpyAsync(b.h_vals, b.d_vals.p, used * sizeof(double), cudaMemcpyDeviceToHost, s));
        }
        c->wire_d2h += (long long)(used * 12 + (size_t)nr * 12 + 8);
        CK(cudaEventRecord(b.done, s));
    };
    // the expansion of batch k runs on the pool of worker threads (rows claimed in small chunks) while
    // this thread drives the GPU for batch k+1
    ExpandPool &pool = expand_pool(c, nthr);
    std::shared_ptr<ExpandPool::Job> pending;           // the batch being expanded (it reads the other staging buffer)
    auto expand_async = [&](int64_t k) {
        scasa_ctx::FetchBuf &b = c->fb[k & 1];
        const int64_t r0 = k * batch, nr = std::min(batch, n_rows - r0);
        pending = pool.post(out + r0 * n_cols, n_cols, nr, b.h_start, b.h_nnz, b.h_cols, b.h_vals);
    };
    auto join_all = [&]() { pool.wait(pending); pending.reset(); };
    struct Joiner { decltype(join_all) &j; ~Joiner() { j(); } } joiner{join_all};   // an exception must not leave a job reading freed buffers
    const bool trace = getenv("SCASA_FETCH_TRACE") != nullptr;
    double t_wait_gpu = 0, t_join = 0, t_launch = 0;
    unsigned long long pairs = 0;
    auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    double t0 = now();
    launch(0);
    t_launch += now() - t0;
    for (int64_t k = 0; k < n_batches; k++) {
        scasa_ctx::FetchBuf &b = c->fb[k & 1];
        t0 = now();
        CK(cudaEventSynchronize(b.done));
        t_wait_gpu += now() - t0;
        pairs += *b.h_cursor;
        const int64_t r0 = k * batch, nr = std::min(batch, n_rows - r0);
        bool overflow = false;
        for (int64_t r = 0; r < nr; r++) if (b.h_nnz[r] < 0) { overflow = true; break; }
        t0 = now();
        join_all();                                   // batch k-1 (it reads the other buffer, which launch(k+1) reuses)
        t_join += now() - t0;
        if (overflow) {                               // denser than the staging buffer: plain copy of this batch
            CK(cudaMemcpyAsync(out + r0 * n_cols, d_mat + r0 * n_cols, (size_t)(nr * n_cols) * sizeof(double), cudaMemcpyDeviceToHost, s));
            CK(cudaStreamSynchronize(s));
            c->wire_d2h += (long long)(nr * n_cols) * 8;
        } else expand_async(k);
        t0 = now();
        if (k + 1 < n_batches) launch(k + 1);
        t_launch += now() - t0;
    }
    t0 = now();
    join_all();
    t_join += now() - t0;
    if (trace)
        fprintf(stderr, "[scasa fetch] %lld x %lld, %lld batches of %lld rows, %llu pairs (%.1f per row): launch+compact %.3f s, "
                        "wait copy %.3f s, wait expansion %.3f s, %d threads\n", (long long)n_rows, (long long)n_cols,
                (long long)n_batches, (long long)batch, pairs, (double)pairs / (double)n_rows, t_launch, t_wait_gpu, t_join, nthr);
}

}  // namespace

// =============================================================================== C ABI
extern "C" {

const char *scasa_version(void) { return "scasa_b200 0.1 (sm_100a)"; }

int scasa_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int scasa_opts_default(scasa_opts_t *o)
{
    if (!o) return SCASA_ERR_ARG;
    o->maxiter_x = 1000; o->maxiter_bt = 1000; o->maxerr = 0.01; o->lim = 0.01; o->modify = 1;
    o->adjust_esp = 2.0; o->hamming_dist = 1; o->fixed_iter = 0; o->fp32 = 0;
    o->maxerr_bt = 0.01; o->lim_bt = 0.01;
    return SCASA_OK;
}

const char *scasa_last_error(const scasa_ctx *c) { return c ? c->err.c_str() : g_err.c_str(); }

int scasa_set_opts(scasa_ctx *c, const scasa_opts_t *o)
{
    if (!c || !o) return fail(c, SCASA_ERR_ARG, "null argument");
    if (o->maxiter_x < 1 || o->maxiter_bt < 1) return fail(c, SCASA_ERR_ARG, "iteration caps must be >= 1");
    if (o->fp32 != 0 && o->fp32 != 1) return fail(c, SCASA_ERR_ARG, "fp32 must be 0 or 1");
    if (!(o->maxerr_bt >= 0) || !(o->lim_bt >= 0)) return fail(c, SCASA_ERR_ARG, "maxerr_bt / lim_bt must be >= 0");
    if (c->crp.ready() && o->hamming_dist != c->opts.hamming_dist)     // the eqclass index was built with the old value
        return fail(c, SCASA_ERR_STATE, "hamming_dist is fixed when the context is created (it is part of the eqclass index)");
    c->opts = *o;
    return SCASA_OK;
}

int scasa_ctx_create(const scasa_xmat_t *dm0, const scasa_crp_t *crp, const scasa_opts_t *opts, int device,
                     scasa_ctx **out)
{
    if (!out) return fail(nullptr, SCASA_ERR_ARG, "out is null");
    *out = nullptr;
    int ndev = scasa_device_count();
    if (ndev <= 0) return fail(nullptr, SCASA_ERR_CUDA, "no CUDA device: libscasa_cuda has no CPU path");
    if (device < 0 || device >= ndev) return fail(nullptr, SCASA_ERR_ARG, "device ordinal out of range");
    scasa_ctx *c = new (std::nothrow) scasa_ctx();
    if (!c) return fail(nullptr, SCASA_ERR_NOMEM, "host allocation failed");
    c->device = device;
    scasa_opts_default(&c->opts);
    int rc = guarded(c, [&] {
        if (opts) { int r = scasa_set_opts(c, opts); if (r) throw ArgErr{r, c->err}; }
        check_xmat(dm0, "dm0");
        cudaDeviceProp prop;
        CK(cudaGetDeviceProperties(&prop, device));
        c->n_sm = prop.multiProcessorCount;
        CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        for (auto &e : c->ev) CK(cudaEventCreate(&e));

        const int B = dm0->n_blocks;
        c->B = B;
        c->q_off.assign(dm0->q_off, dm0->q_off + B + 1);
        c->p_off.assign(dm0->p_off, dm0->p_off + B + 1);
        c->x_off.assign(dm0->x_off, dm0->x_off + B + 1);
        c->x.assign(dm0->x, dm0->x + dm0->x_off[B]);
        c->n_rows = c->q_off[B]; c->n_cols = c->p_off[B];
        c->em_index.assign(B, -1);
        c->poor.assign(B, 0);
        c->row_block.resize((size_t)c->n_rows);
        for (int k = 0; k < B; k++) {
            int q = c->q_off[k + 1] - c->q_off[k], p = c->p_off[k + 1] - c->p_off[k];
            for (int r = 0; r < q; r++) c->row_block[(size_t)c->q_off[k] + r] = k;
            if (q >= 65536) throw ArgErr{SCASA_ERR_UNSUPPORTED, "a block may have at most 65535 rows"};
            if (q == 1) {
                // Rsource:981 assigns colnames(X0) as the single row name: only p == 1 is valid R
                if (p != 1) throw ArgErr{SCASA_ERR_ARG, "block " + std::to_string(k) + ": one row but several columns"};
                continue;
            }
            c->em_index[k] = (int)c->em_blocks.size();
            c->em_blocks.push_back(k);
            c->max_em_q = std::max<int>(c->max_em_q, (int)(c->q_off[(size_t)k + 1] - c->q_off[(size_t)k]));
            if (is_poor(c->x.data() + c->x_off[k], q, p)) {
                c->poor[k] = 1;
                c->poor_em.push_back(c->em_index[k]);
                c->poor_rows += q;
            }
        }
        c->n_em = (int)c->em_blocks.size();

        // queue order of the EM blocks: expensive shapes first (long tasks start early)
        c->em_order.resize((size_t)c->n_em);
        std::iota(c->em_order.begin(), c->em_order.end(), 0);
        {
            std::vector<int64_t> cost((size_t)c->n_em);
            for (int e = 0; e < c->n_em; e++) {
                int k = c->em_blocks[e];
                int64_t q = c->q_off[k + 1] - c->q_off[k], p = c->p_off[k + 1] - c->p_off[k];
                if (p > 1024) throw ArgErr{SCASA_ERR_UNSUPPORTED, "a block may have at most 1024 columns"};
                cost[(size_t)e] = q * p * p;
            }
            std::stable_sort(c->em_order.begin(), c->em_order.end(),
                             [&](int32_t x, int32_t y) { return cost[(size_t)x] > cost[(size_t)y]; });
        }
        if (getenv("SCASA_FETCH_DENSE")) c->fetch_dense_copy = 1;
        if (const char *env = getenv("SCASA_RUN_PARTS")) c->run_parts = std::max(0, atoi(env));
        // EM work classes; SCASA_EM_CFG="warps:pairs:bytes:cta_threads:ctas_per_sm,..." overrides (tuning)
        if (const char *env = getenv("SCASA_EM_CFG")) {
            int k = 0;
            const char *s = env;
            while (k < kClasses && *s) {
                int w, pr, by, th, cps, used = 0;
                if (sscanf(s, "%d:%d:%d:%d:%d%n", &w, &pr, &by, &th, &cps, &used) != 5) break;
                if ((w == 1 || w == 2 || w == 4 || w == 8) && th >= 32 * w && th <= 256 && th % (32 * w) == 0 && cps >= 1) {
                    c->cls_warps[k] = w; c->cls_pairs[k] = pr; c->cls_bytes[k] = by; c->cls_cta_threads[k] = th;
                    c->cls_ctas_per_sm[k] = cps;
                }
                k++;
                s += used;
                if (*s == ',') s++;
            }
        }
        c->cls_pairs[kClasses - 1] = 1 << 30; c->cls_bytes[kClasses - 1] = 1 << 30;
        if (const char *env = getenv("SCASA_EM_PAIR")) {      // "max[:ctas_per_sm]"
            int m = 16, cps = 8;
            const int got = sscanf(env, "%d:%d", &m, &cps);
            if (got >= 1) c->pair_max = std::max(0, std::min(16, m));
            if (got >= 2 && cps >= 1) c->pair_ctas_per_sm = cps;
        }
        for (int k = 0; k < kLists; k++) {
            {   // the classes with few, long tasks get scheduling priority: their CTAs go first when an SM has room,
                // the many-small-tasks class fills what is left and takes over when they are done
                int lo_p = 0, hi_p = 0;
                CK(cudaDeviceGetStreamPriorityRange(&lo_p, &hi_p));
                const int prio = (k == 0 || k == kPairList || getenv("SCASA_EM_NOPRIO")) ? lo_p : hi_p;
                CK(cudaStreamCreateWithPriority(&c->cls_stream[k], cudaStreamNonBlocking, prio));
            }
            CK(cudaEventCreateWithFlags(&c->ev_join[k], cudaEventDisableTiming));
            CK(cudaEventCreate(&c->ev_cls[k][0]));
            CK(cudaEventCreate(&c->ev_cls[k][1]));
        }
        CK(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&c->ev_fill, cudaEventDisableTiming));

        c->d_q_off.upload(c->q_off, c->stream);
        c->d_p_off.upload(c->p_off, c->stream);
        c->d_x_off.upload(c->x_off, c->stream);
        c->d_x.upload(c->x, c->stream);
        c->d_em_blocks.upload(c->em_blocks, c->stream);
        c->d_poor.upload(c->poor, c->stream);
        c->d_em_order.ensure(c->em_order.size() + 1);
        c->d_em_order.upload(c->em_order, c->stream);
        {
            std::vector<int2> info((size_t)c->n_rows);
            std::vector<int32_t> rcol((size_t)c->n_rows);
            for (int k = 0; k < B; k++)
                for (int r = c->q_off[k]; r < c->q_off[k + 1]; r++) {
                    info[(size_t)r] = make_int2(c->em_index[k], r - c->q_off[k]);
                    rcol[(size_t)r] = c->p_off[k];
                }
            c->d_row_info.upload(info, c->stream);
            c->d_row_col.upload(rcol, c->stream);
        }
        c->d_next.ensure(16); c->d_counters.ensure(8); c->d_stats.ensure(8); c->d_totals.ensure(4); c->d_flag.ensure(1);
        CK(cudaStreamSynchronize(c->stream));

        if (crp) {
            check_xmat(&crp->m, "crp");
            if (crp->m.n_blocks != B) throw ArgErr{SCASA_ERR_ARG, "crp and dm0 differ in block count"};
            for (int k = 0; k <= B; k++)
                if (crp->m.q_off[k] != dm0->q_off[k]) throw ArgErr{SCASA_ERR_ARG, "crp and dm0 differ in rows"};
            if (!crp->col_tx || !crp->row_pat || !crp->name_rank) throw ArgErr{SCASA_ERR_ARG, "crp: null dimnames arrays"};
            c->crp.build(crp->m.n_blocks, crp->m.q_off, crp->m.p_off, crp->m.x_off, crp->m.x, crp->col_tx,
                         crp->row_pat, crp->name_rank, c->opts.hamming_dist);
        }
    });
    if (rc != SCASA_OK) { g_err = c->err; scasa_ctx_destroy(c); return rc; }
    *out = c;
    return SCASA_OK;
}

void scasa_ctx_destroy(scasa_ctx *c)
{
    if (!c) return;
    if (c->twin) { scasa_ctx_destroy(c->twin); c->twin = nullptr; }
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    DBuf<int32_t> *i32[] = {&c->d_q_off, &c->d_p_off, &c->d_em_blocks, &c->d_em_order,
                            &c->d_cell_chunk, &c->d_y_len, &c->d_y_row, &c->d_eq_id, &c->d_eq_count, &c->d_eq_row,
                            &c->d_cnt, &c->d_runs, &c->d_tbytes, &c->d_iters, &c->d_row_col,
                            &c->d_list[0], &c->d_list[1], &c->d_list[2], &c->d_list[3], &c->d_list[4],
                            &c->d_gene_ptr, &c->d_gene_cols};
    for (auto b : i32) b->release();
    DBuf<int64_t> *i64[] = {&c->d_x_off, &c->d_chunk_off, &c->d_y_start, &c->d_cell_ptr, &c->d_toff, &c->d_scan_sums, &c->d_totals};
    for (auto b : i64) b->release();
    DBuf<double> *f64[] = {&c->d_x, &c->d_y_val, &c->d_iso, &c->d_partial,
                           &c->d_colsum, &c->d_colsum_parts, &c->d_gene};
    for (auto b : f64) b->release();
    c->d_poor.release(); c->d_next.release(); c->d_stats.release(); c->d_flag.release();
    c->d_tdata.release(); c->d_counters.release(); c->d_row_info.release(); c->d_scatter_scratch.release();
    for (auto &b : c->fb) {
        b.d_cols.release(); b.d_nnz.release(); b.d_vals.release(); b.d_start.release(); b.d_cursor.release();
        if (b.h_cols) { cudaFreeHost(b.h_cols); cudaFreeHost(b.h_vals); cudaFreeHost(b.h_start); cudaFreeHost(b.h_nnz); cudaFreeHost(b.h_cursor); }
        if (b.done) cudaEventDestroy(b.done);
    }
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    if (c->ev_fill) cudaEventDestroy(c->ev_fill);
    for (int k = 0; k < kLists; k++) {
        if (c->ev_join[k]) cudaEventDestroy(c->ev_join[k]);
        for (auto &e : c->ev_cls[k]) if (e) cudaEventDestroy(e);
        if (c->cls_stream[k]) cudaStreamDestroy(c->cls_stream[k]);
    }
    for (auto &e : c->ev) if (e) cudaEventDestroy(e);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

int64_t scasa_n_rows(const scasa_ctx *c) { return c ? c->n_rows : 0; }
int64_t scasa_n_cols(const scasa_ctx *c) { return c ? c->n_cols : 0; }
int32_t scasa_n_em_blocks(const scasa_ctx *c) { return c ? c->n_em : 0; }
int scasa_em_blocks(const scasa_ctx *c, int32_t *ids)
{
    if (!c || !ids) return SCASA_ERR_ARG;
    std::copy(c->em_blocks.begin(), c->em_blocks.end(), ids);
    return SCASA_OK;
}
int scasa_poor_blocks(const scasa_ctx *c, uint8_t *is)
{
    if (!c || !is) return SCASA_ERR_ARG;
    std::copy(c->poor.begin(), c->poor.end(), is);
    return SCASA_OK;
}

int scasa_chunk_split(int64_t n, int32_t core, int64_t *chunk_off, int64_t cap, int32_t *n_chunks)
{
    // scasa_quant_step2_v1.0.0.R:64-68: chunkNum = round(n/100) (half to even), or `core` when
    // n <= 100; brPos = round(seq(1, n+1, length.out = chunkNum+1)).
    if (n <= 0 || !chunk_off || !n_chunks) return fail(nullptr, SCASA_ERR_ARG, "bad argument");
    int64_t k = (n > 100) ? (int64_t)std::nearbyint((double)n / 100.0) : core;
    if (k < 1) return fail(nullptr, SCASA_ERR_ARG, "core must be >= 1 when n_cells <= 100");
    if (k + 1 > cap) return fail(nullptr, SCASA_ERR_ARG, "chunk_off too small");
    const double from = 1.0, to = (double)n + 1.0;
    const double by = (to - from) / (double)k;
    for (int64_t i = 0; i <= k; i++) {
        double v = (i == 0) ? from : (i == k ? to : from + (double)i * by);
        chunk_off[i] = (int64_t)std::nearbyint(v) - 1;
    }
    *n_chunks = (int32_t)k;
    return SCASA_OK;
}

int scasa_stage_counts(scasa_ctx *c, int64_t n_cells, int32_t n_chunks, const int64_t *chunk_off,
                       const int64_t *y_rowptr, const int32_t *y_rowidx, const double *y_val)
{
    if (!c) return fail(c, SCASA_ERR_ARG, "null context");
    return guarded(c, [&] {
        if (!y_rowptr || (!y_rowidx && y_rowptr[n_cells] > 0) || (!y_val && y_rowptr[n_cells] > 0))
            throw ArgErr{SCASA_ERR_ARG, "null count arrays"};
        stage_common(c, n_cells, n_chunks, chunk_off);
        const int64_t nnz = y_rowptr[n_cells];
        if (y_rowptr[0] != 0 || nnz < 0) throw ArgErr{SCASA_ERR_ARG, "y_rowptr must start at 0"};
        std::vector<int32_t> len((size_t)n_cells);
        for (int64_t i = 0; i < n_cells; i++) {
            int64_t a = y_rowptr[i], b = y_rowptr[i + 1];
            if (b < a || b - a > c->n_rows) throw ArgErr{SCASA_ERR_ARG, "y_rowptr is not a valid CSR pointer"};
            len[(size_t)i] = (int32_t)(b - a);
            for (int64_t k = a; k < b; k++) {
                int32_t r = y_rowidx[k];
                if (r < 0 || r >= c->n_rows || (k > a && r <= y_rowidx[k - 1]))
                    throw ArgErr{SCASA_ERR_ARG, "y_rowidx must be strictly ascending inside a cell and < n_rows"};
            }
        }
        c->nnz = nnz; c->have_eq = false;
        c->d_y_start.ensure((size_t)n_cells); c->d_y_len.ensure((size_t)n_cells);
        c->d_y_row.ensure((size_t)nnz + 1); c->d_y_val.ensure((size_t)nnz + 1);
        CK(cudaMemcpyAsync(c->d_y_start.p, y_rowptr, n_cells * sizeof(int64_t), cudaMemcpyHostToDevice, c->stream));
        CK(cudaMemcpyAsync(c->d_y_len.p, len.data(), n_cells * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
        if (nnz) {
            CK(cudaMemcpyAsync(c->d_y_row.p, y_rowidx, nnz * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
            CK(cudaMemcpyAsync(c->d_y_val.p, y_val, nnz * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        }
        CK(cudaStreamSynchronize(c->stream));
        c->wire_h2d += (long long)(n_cells * 12 + nnz * 12 + (n_chunks + 1) * 8 + n_cells * 4);
    });
}

int scasa_stage_eqcounts(scasa_ctx *c, int64_t n_cells, int32_t n_chunks, const int64_t *chunk_off,
                         const int64_t *cell_ptr, const int32_t *eq_id, const int32_t *eq_count,
                         int64_t n_eq, const int32_t *eq_row)
{
    if (!c) return fail(c, SCASA_ERR_ARG, "null context");
    return guarded(c, [&] {
        if (!cell_ptr || !eq_id || !eq_count || !eq_row || n_eq <= 0) throw ArgErr{SCASA_ERR_ARG, "null eqclass arrays"};
        stage_common(c, n_cells, n_chunks, chunk_off);
        const int64_t nnz = cell_ptr[n_cells];
        if (cell_ptr[0] != 0 || nnz < 0) throw ArgErr{SCASA_ERR_ARG, "cell_ptr must start at 0"};
        int64_t mx = 0;
        for (int64_t i = 0; i < n_cells; i++) {
            int64_t d = cell_ptr[i + 1] - cell_ptr[i];
            if (d < 0) throw ArgErr{SCASA_ERR_ARG, "cell_ptr must be non-decreasing"};
            mx = std::max(mx, d);
        }
        if (mx > 16384) throw ArgErr{SCASA_ERR_UNSUPPORTED, "a cell may list at most 16384 eqclasses"};
        for (int64_t e = 0; e < n_eq; e++)
            if (eq_row[e] >= c->n_rows) throw ArgErr{SCASA_ERR_ARG, "eq_row out of range"};
        c->max_cell_entries = mx; c->nnz = nnz; c->n_eq = n_eq; c->have_eq = true;
        c->d_cell_ptr.ensure((size_t)n_cells + 1); c->d_eq_id.ensure((size_t)nnz + 1);
        c->d_eq_count.ensure((size_t)nnz + 1); c->d_eq_row.ensure((size_t)n_eq);
        c->d_y_start.ensure((size_t)n_cells); c->d_y_len.ensure((size_t)n_cells);
        c->d_y_row.ensure((size_t)nnz + 1); c->d_y_val.ensure((size_t)nnz + 1);
        CK(cudaMemcpyAsync(c->d_cell_ptr.p, cell_ptr, (n_cells + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, c->stream));
        CK(cudaMemcpyAsync(c->d_y_start.p, cell_ptr, n_cells * sizeof(int64_t), cudaMemcpyHostToDevice, c->stream));
        if (nnz) {
            CK(cudaMemcpyAsync(c->d_eq_id.p, eq_id, nnz * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
            CK(cudaMemcpyAsync(c->d_eq_count.p, eq_count, nnz * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
        }
        CK(cudaMemcpyAsync(c->d_eq_row.p, eq_row, n_eq * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        c->wire_h2d += (long long)((n_cells + 1) * 8 + n_cells * 8 + nnz * 8 + n_eq * 4 + (n_chunks + 1) * 8 + n_cells * 4);
    });
}

int scasa_set_gene_map(scasa_ctx *c, const int32_t *gene_of_col, int32_t n_genes)
{
    if (!c) return fail(c, SCASA_ERR_ARG, "null context");
    return guarded(c, [&] {
        if (!gene_of_col || n_genes <= 0) { c->n_genes = 0; c->gene_of_col_host.clear(); return; }
        std::vector<int32_t> ptr((size_t)n_genes + 1, 0), cols;
        for (int64_t j = 0; j < c->n_cols; j++) {
            int g = gene_of_col[j];
            if (g >= n_genes) throw ArgErr{SCASA_ERR_ARG, "gene_of_col out of range"};
            if (g >= 0) ptr[(size_t)g + 1]++;
        }
        for (int g = 0; g < n_genes; g++) ptr[(size_t)g + 1] += ptr[(size_t)g];
        cols.resize((size_t)ptr[(size_t)n_genes]);
        std::vector<int32_t> cur(ptr.begin(), ptr.end() - 1);
        for (int64_t j = 0; j < c->n_cols; j++) {        // members in column order (step3:27)
            int g = gene_of_col[j];
            if (g >= 0) cols[(size_t)cur[(size_t)g]++] = (int32_t)j;
        }
        c->n_genes = n_genes;
        std::vector<int32_t> goc(gene_of_col, gene_of_col + c->n_cols);
        c->gene_of_col_host = goc;
        c->d_gene_ptr.upload(ptr, c->stream);
        c->d_gene_cols.ensure(cols.size() + 1);
        c->d_gene_cols.upload(cols, c->stream);
        CK(cudaStreamSynchronize(c->stream));
    });
}

// What one range of chunks contributed to the run statistics.
struct RangeStats {
    float pack_ms = 0, tasks_ms = 0, em_ms = 0, finish_ms = 0, cls_ms[kLists] = {0, 0, 0, 0, 0};
    unsigned long long st[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    unsigned int cnts[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int64_t blob_bytes = 0;
    int launches = 0;
};

// Chunks [h0, h1) of the sample staged in c: pack -> tasks -> EM -> column / gene sums, with w's
// work buffers, streams and X-matrix tables (w is c itself or its twin).  Results go into c's matrices.
static void run_range(scasa_ctx *w, scasa_ctx *c, int h0, int h1, int part, RangeStats &rs)
{
    cudaStream_t s = w->stream;
    int launches = 0;
    const int n_chunks = h1 - h0;
    const int64_t c0 = c->chunk_off[(size_t)h0], n_cells = c->chunk_off[(size_t)h1] - c0;
    const int64_t n_tasks = (int64_t)n_chunks * w->n_em;
    double *iso = c->d_iso.p + c0 * c->n_cols;
    w->d_cnt.ensure((size_t)n_tasks + 1); w->d_runs.ensure((size_t)n_tasks + 1);
    w->d_tbytes.ensure((size_t)n_tasks + 1); w->d_toff.ensure((size_t)n_tasks + 1);
    for (int k = 0; k < kLists; k++) w->d_list[k].ensure((size_t)n_tasks + 1);

    CK(cudaEventRecord(w->ev[0], s));
    CK(cudaMemsetAsync(w->d_cnt.p, 0, (n_tasks + 1) * sizeof(int32_t), s));
    CK(cudaMemsetAsync(w->d_runs.p, 0, (n_tasks + 1) * sizeof(int32_t), s));
    CK(cudaMemsetAsync(w->d_stats.p, 0, 8 * sizeof(unsigned long long), s));
    CK(cudaMemsetAsync(w->d_next.p, 0, 16 * sizeof(unsigned int), s));
    CK(cudaMemsetAsync(w->d_counters.p, 0, 8 * sizeof(unsigned int), s));
    CK(cudaMemsetAsync(w->d_flag.p, 0, sizeof(int), s));

    // ---- pack: eqclass counts -> Y
    if (c->have_eq) {
        const int wpt = (int)((c->n_rows + 32 * 512 - 1) / (32 * 512));      // bitmap words per thread
        const size_t dense_smem = (size_t)wpt * 512 * 33 * sizeof(int);       // int32 histogram + touched-row bitmap
        if (dense_smem <= 200 * 1024 && !getenv("SCASA_PACK_SORT")) {
            CK(cudaFuncSetAttribute(pack_cells_dense_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dense_smem));
            int grid = (int)std::min<int64_t>(n_cells, (int64_t)w->n_sm);
            pack_cells_dense_kernel<<<grid, 512, dense_smem, s>>>(n_cells, c->d_cell_ptr.p + c0, c->d_eq_id.p, c->d_eq_count.p,
                                                                  c->d_eq_row.p, c->n_eq, (int)c->n_rows, wpt, c->d_y_row.p,
                                                                  c->d_y_val.p, c->d_y_len.p + c0);
        } else {
            int cap = 32;
            while (cap < c->max_cell_entries) cap <<= 1;
            size_t smem = (size_t)cap * sizeof(unsigned long long);
            CK(cudaFuncSetAttribute(pack_cells_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            int grid = (int)std::min<int64_t>(n_cells, (int64_t)w->n_sm * 8);
            pack_cells_kernel<<<grid, 256, smem, s>>>(n_cells, c->d_cell_ptr.p + c0, c->d_eq_id.p, c->d_eq_count.p,
                                                      c->d_eq_row.p, c->n_eq, cap, c->d_y_row.p, c->d_y_val.p,
                                                      c->d_y_len.p + c0, w->d_flag.p);
        }
        CK(cudaGetLastError());
        launches++;
    }
    CK(cudaEventRecord(w->ev[1], s));

    XDev X = xdev(w);
    Sample S = sdev(c, h0, h1);
    // ---- result rows: zeros + pass-through blocks, on a side stream next to the task build
    {
        cudaStream_t fs = w->cls_stream[0];
        CK(cudaStreamWaitEvent(fs, w->ev[1], 0));
        const size_t smem = (size_t)2 * kFillTile * sizeof(double);
        CK(cudaFuncSetAttribute(iso_fill_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const int fill_ctas = getenv("SCASA_FILL_CTAS") ? std::max(1, atoi(getenv("SCASA_FILL_CTAS"))) : 2;
        const int grid = (int)std::min<int64_t>(n_cells, (int64_t)w->n_sm * fill_ctas);
        const int use_bulk = (c->n_cols % 2 == 0) && ((uintptr_t)iso % 16 == 0) && !getenv("SCASA_FILL_PLAIN");
        iso_fill_kernel<<<grid, kFillThreads, smem, fs>>>(X, S, iso, use_bulk);
        CK(cudaGetLastError());
        CK(cudaEventRecord(w->ev_fill, fs));
        launches++;
    }

    // ---- task extents, work classes, blobs
    Tasks T = tdev(w);
    EmOpts O;
    O.maxiter_x = c->opts.maxiter_x; O.maxiter_bt = c->opts.maxiter_bt; O.modify = c->opts.modify;
    O.fixed_iter = c->opts.fixed_iter;
    O.maxerr = c->opts.maxerr; O.lim = c->opts.lim; O.adjust_esp = c->opts.adjust_esp;
    O.maxerr_bt = c->opts.maxerr_bt; O.lim_bt = c->opts.lim_bt;
    ClassCfg CC;
    for (int k = 0; k < kClasses; k++) { CC.pairs[k] = c->cls_pairs[k]; CC.bytes[k] = c->cls_bytes[k]; }
    CC.real_bytes = c->opts.fp32 ? 4 : 8;
    CC.pair_max = c->pair_max;
    int32_t *iters = c->d_iters.p + (int64_t)h0 * w->n_em * 2;
    int64_t tot[2] = {0, 0};
    unsigned int *cnts = rs.cnts;
    if (w->n_em > 0) {
        const int wpb = 8;
        task_count_kernel<<<(unsigned)((n_cells + wpb - 1) / wpb), wpb * 32, 0, s>>>(X, S, T);
        classify_kernel<<<(unsigned)((n_tasks + 255) / 256), 256, 0, s>>>(X, S, T, n_tasks, O, CC, iters, w->d_stats.p);
        launches += 2;
        exclusive_scan(w, w->d_tbytes.p, n_tasks, w->d_toff.p, w->d_totals.p + 0, launches);
        CK(cudaMemcpyAsync(tot, w->d_totals.p, sizeof tot, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(cnts, w->d_counters.p, 8 * sizeof(unsigned int), cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        if (cnts[7] == 1) throw ArgErr{SCASA_ERR_UNSUPPORTED, "a poor-condition (2-column) block may have at most 64 rows"};
        if (cnts[7] == 2) throw ArgErr{SCASA_ERR_UNSUPPORTED, "a (chunk, block) task may hold at most 65535 cells-with-counts and 65535 count entries"};
        if (cnts[5] > 200 * 1024) throw ArgErr{SCASA_ERR_UNSUPPORTED, "a (chunk, block) task needs more than 200 KB of shared memory (chunk too large for this block)"};
        w->d_tdata.ensure((size_t)tot[0] + 64);
        T = tdev(w);
        rs.blob_bytes = tot[0];
        const size_t st_bytes = (size_t)3 * w->n_em * sizeof(uint32_t);
        uint32_t *scratch = nullptr;
        if (st_bytes > 160 * 1024) {           // too many EM blocks for shared memory: cursors in HBM
            w->d_scatter_scratch.ensure((size_t)n_chunks * 3 * w->n_em);
            scratch = w->d_scatter_scratch.p;
        }
        // the per-cell index of the staged path sits behind the cursors (see task_scatter_kernel)
        const size_t staged_bytes = st_bytes + (size_t)kScatterCap * 6 + (size_t)w->n_em * 2;
        const int staged = (!scratch && staged_bytes <= 200 * 1024 && w->max_em_q < 32768 && !getenv("SCASA_SCATTER_PLAIN")) ? 1 : 0;   // rows | poor flag in 16 bits
        const size_t sc_smem = scratch ? 0 : (staged ? staged_bytes : st_bytes);
        if (staged) {
            CK(cudaFuncSetAttribute(task_scatter_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sc_smem));
            task_scatter_kernel<true><<<n_chunks, kScatterThreads, sc_smem, s>>>(X, S, T, scratch);
        } else {
            if (sc_smem) CK(cudaFuncSetAttribute(task_scatter_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sc_smem));
            task_scatter_kernel<false><<<n_chunks, kScatterThreads, sc_smem, s>>>(X, S, T, scratch);
        }
        CK(cudaGetLastError());
        launches++;
    }
    CK(cudaStreamWaitEvent(s, w->ev_fill, 0));
    CK(cudaEventRecord(w->ev[2], s));

    // ---- EM: one persistent kernel per work class, side by side on their own streams
    bool ran_cls[kLists] = {false, false, false, false, false};
    if (w->n_em > 0 && cnts[4] > 0) {
        CK(cudaEventRecord(w->ev_fork, s));
        for (int k = kClasses - 1; k >= 0; k--) {          // long tasks first
            const unsigned n = cnts[k];
            if (n == 0) continue;
            const int W = c->cls_warps[k];
            const int threads = c->cls_cta_threads[k];
            const int groups = threads / (32 * W);
            int slot = (k == kClasses - 1) ? (int)cnts[5] : std::min<int>(c->cls_bytes[k], (int)cnts[5]);
            slot = (slot + 15) & ~15;
            const size_t smem = (size_t)slot * groups;
            if (smem > 220 * 1024) throw ArgErr{SCASA_ERR_UNSUPPORTED, "EM work class needs more shared memory than an SM has"};
            EmArgs A;
            A.X = X; A.S = S; A.T = T; A.O = O; A.out = iso; A.iters = iters; A.stats = w->d_stats.p;
            A.list = w->d_list[k].p; A.list_n = w->d_counters.p + k; A.next = w->d_next.p + k; A.slot_bytes = slot;
            const unsigned grid = (unsigned)std::min<int64_t>((int64_t)w->n_sm * c->cls_ctas_per_sm[k], ((int64_t)n + groups - 1) / groups);
            cudaStream_t cs = w->cls_stream[k];
            CK(cudaStreamWaitEvent(cs, w->ev_fork, 0));
            CK(cudaEventRecord(w->ev_cls[k][0], cs));
            switch (W) {
            case 1: launch_em<1>(A, grid, threads, smem, cs, c->opts.fp32 != 0); break;
            case 2: launch_em<2>(A, grid, threads, smem, cs, c->opts.fp32 != 0); break;
            case 4: launch_em<4>(A, grid, threads, smem, cs, c->opts.fp32 != 0); break;
            default: launch_em<8>(A, grid, threads, smem, cs, c->opts.fp32 != 0); break;
            }
            CK(cudaEventRecord(w->ev_cls[k][1], cs));
            CK(cudaEventRecord(w->ev_join[k], cs));
            CK(cudaStreamWaitEvent(s, w->ev_join[k], 0));
            ran_cls[k] = true;
            launches++;
        }
        if (cnts[kPairCounter] > 0) {                        // tiny tasks, two per warp
            const int k = kPairList, threads = 128, halves = threads / 16;
            const int rb = c->opts.fp32 ? 4 : 8, pm = c->pair_max;
            int slot = 16;
            for (int p = 1; p <= pm; p++)
                for (int q = 1; q * p <= pm; q++)
                    for (int n = 1; n * p <= pm; n++) slot = std::max(slot, slot_bytes_of(q, p, n, std::min(pm, n * q), rb));
            slot = (slot + 15) & ~15;
            const size_t smem = (size_t)slot * halves;
            const unsigned n = cnts[kPairCounter];
            EmArgs A;
            A.X = X; A.S = S; A.T = T; A.O = O; A.out = iso; A.iters = iters; A.stats = w->d_stats.p;
            A.list = w->d_list[k].p; A.list_n = w->d_counters.p + kPairCounter; A.next = w->d_next.p + k; A.slot_bytes = slot;
            const unsigned grid = (unsigned)std::min<int64_t>((int64_t)w->n_sm * c->pair_ctas_per_sm, ((int64_t)n + halves - 1) / halves);
            cudaStream_t cs = w->cls_stream[k];
            CK(cudaStreamWaitEvent(cs, w->ev_fork, 0));
            CK(cudaEventRecord(w->ev_cls[k][0], cs));
            if (c->opts.fp32) {
                CK(cudaFuncSetAttribute(em_pair_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                em_pair_kernel<float><<<grid, threads, smem, cs>>>(A);
            } else {
                CK(cudaFuncSetAttribute(em_pair_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                em_pair_kernel<double><<<grid, threads, smem, cs>>>(A);
            }
            CK(cudaEventRecord(w->ev_cls[k][1], cs));
            CK(cudaEventRecord(w->ev_join[k], cs));
            CK(cudaStreamWaitEvent(s, w->ev_join[k], 0));
            ran_cls[k] = true;
            launches++;
        }
        CK(cudaGetLastError());
    }
    CK(cudaEventRecord(w->ev[3], s));

    // ---- column sums (+ gene sums)
    double *gene = c->n_genes > 0 ? c->d_gene.p + c0 * (int64_t)c->n_genes : nullptr;
    launch_finish(w, iso, n_cells, gene, c->d_colsum_parts.p + (size_t)part * c->n_cols, launches);
    CK(cudaEventRecord(w->ev[4], s));

    int flag = 0;
    CK(cudaMemcpyAsync(rs.st, w->d_stats.p, sizeof rs.st, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(&flag, w->d_flag.p, sizeof flag, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    if (flag) throw ArgErr{SCASA_ERR_UNSUPPORTED, "a cell exceeded the eqclass capacity of the pack kernel"};
    CK(cudaEventElapsedTime(&rs.pack_ms, w->ev[0], w->ev[1]));
    CK(cudaEventElapsedTime(&rs.tasks_ms, w->ev[1], w->ev[2]));
    CK(cudaEventElapsedTime(&rs.em_ms, w->ev[2], w->ev[3]));
    CK(cudaEventElapsedTime(&rs.finish_ms, w->ev[3], w->ev[4]));
    for (int k = 0; k < kLists; k++)
        if (ran_cls[k]) CK(cudaEventElapsedTime(&rs.cls_ms[k], w->ev_cls[k][0], w->ev_cls[k][1]));
    rs.launches = launches;
}

// make the twin context (second set of streams and work buffers on the same GPU) match c
static int sync_twin(scasa_ctx *c)
{
    if (!c->twin) {
        scasa_xmat_t m;
        m.n_blocks = c->B; m.q_off = c->q_off.data(); m.p_off = c->p_off.data(); m.x_off = c->x_off.data(); m.x = c->x.data();
        int rc = scasa_ctx_create(&m, nullptr, &c->opts, c->device, &c->twin);
        if (rc) return fail(c, rc, std::string("twin context: ") + scasa_last_error(nullptr));
    }
    scasa_ctx *t = c->twin;
    t->pool_owner = c;
    t->opts = c->opts; t->host_threads = c->host_threads; t->fetch_dense_copy = c->fetch_dense_copy;
    for (int k = 0; k < kClasses; k++) {
        t->cls_warps[k] = c->cls_warps[k]; t->cls_pairs[k] = c->cls_pairs[k]; t->cls_bytes[k] = c->cls_bytes[k];
        t->cls_cta_threads[k] = c->cls_cta_threads[k]; t->cls_ctas_per_sm[k] = c->cls_ctas_per_sm[k];
    }
    t->pair_max = c->pair_max; t->pair_ctas_per_sm = c->pair_ctas_per_sm;
    if (c->n_genes > 0 && (t->n_genes != c->n_genes || t->gene_of_col_host != c->gene_of_col_host)) {
        int rc = scasa_set_gene_map(t, c->gene_of_col_host.data(), c->n_genes);
        if (rc) return fail(c, rc, std::string("twin context: ") + scasa_last_error(t));
    } else if (c->n_genes <= 0) t->n_genes = 0;
    return SCASA_OK;
}

// The staged sample is cut into ranges of chunks that alternate between this context and its twin,
// each driven by its own host thread: the instruction-bound EM kernels of one range run next to
// the bandwidth-bound pack / row-fill / gene kernels of another.  Chunks are independent, so the
// cut changes no result.  `allow_split = false` (scasa_quant's slices, which already alternate on
// the two contexts) runs everything as one range on c.
static int run_staged(scasa_ctx *c, bool allow_split)
{
    if (!c->staged) return fail(c, SCASA_ERR_STATE, "nothing staged: call scasa_stage_counts / scasa_stage_eqcounts first");
    int parts = 1;
    if (allow_split) {
        parts = c->run_parts > 0 ? c->run_parts : 1;   // measured: no gain (the persistent EM kernels fill the SMs' registers
                                                        // and shared memory, the other range's kernels queue behind them)
        parts = std::max(1, std::min(parts, c->n_chunks));
    }
    if (parts > 1) { int rc = sync_twin(c); if (rc) return rc; }
    // range boundaries: the first and last ranges are half-size so that the two workers fall out of phase
    std::vector<int> cut((size_t)parts + 1, 0);
    {
        std::vector<double> wgt((size_t)parts, 1.0);
        if (parts >= 4) { wgt[0] = 0.5; wgt[(size_t)parts - 1] = 0.5; }
        double tot = 0, acc = 0;
        for (double x : wgt) tot += x;
        for (int k = 0; k < parts; k++) {
            acc += wgt[(size_t)k];
            cut[(size_t)k + 1] = std::max(cut[(size_t)k] + 1, (int)std::llround(c->n_chunks * acc / tot));
        }
        cut[(size_t)parts] = c->n_chunks;
        for (int k = parts - 1; k > 0; k--) cut[(size_t)k] = std::min(cut[(size_t)k], cut[(size_t)k + 1] - 1);
    }
    std::vector<RangeStats> rs((size_t)parts);
    int rcs[2] = {SCASA_OK, SCASA_OK};
    int rc0 = guarded(c, [&] {
        const int64_t n_tasks = (int64_t)c->n_chunks * c->n_em;
        c->d_iso.ensure((size_t)(c->n_cells * c->n_cols));
        c->d_iters.ensure((size_t)n_tasks * 2 + 2);
        c->d_colsum.ensure((size_t)c->n_cols);
        c->d_colsum_parts.ensure((size_t)parts * c->n_cols);
        if (c->n_genes > 0) c->d_gene.ensure((size_t)c->n_cells * c->n_genes);
        CK(cudaEventRecord(c->ev[6], c->stream));
    });
    if (rc0) return rc0;
    auto drive = [&](int which) {
        scasa_ctx *w = which ? c->twin : c;
        for (int k = which; k < parts; k += 2) {
            int rc = guarded(w, [&] { run_range(w, c, cut[(size_t)k], cut[(size_t)k + 1], k, rs[(size_t)k]); });
            if (rc) { rcs[which] = rc; return; }
        }
    };
    if (parts > 1) {                                 // drive() catches everything itself (guarded)
        std::thread other;
        try {
            other = std::thread(drive, 1);
        } catch (const std::exception &e) {
            return fail(c, SCASA_ERR_NOMEM, std::string("cannot start the second host thread: ") + e.what());
        }
        drive(0);
        other.join();
    } else drive(0);
    if (rcs[0]) return rcs[0];
    if (rcs[1]) return fail(c, rcs[1], std::string("twin context: ") + scasa_last_error(c->twin));
    return guarded(c, [&] {
        cudaStream_t s = c->stream;
        colsum_combine<<<(unsigned)((c->n_cols + 255) / 256), 256, 0, s>>>(c->d_colsum_parts.p, parts, c->n_cols, c->d_colsum.p);
        CK(cudaGetLastError());
        CK(cudaEventRecord(c->ev[7], s));
        CK(cudaStreamSynchronize(s));
        scasa_timing_t &tm = c->timing;
        memset(&tm, 0, sizeof tm);
        CK(cudaEventElapsedTime(&tm.total_ms, c->ev[6], c->ev[7]));
        unsigned long long st[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (const RangeStats &r : rs) {
            tm.pack_ms += r.pack_ms; tm.tasks_ms += r.tasks_ms; tm.em_ms += r.em_ms; tm.finish_ms += r.finish_ms;
            for (int k = 0; k < kClasses; k++) { tm.em_class_ms[k] += r.cls_ms[k]; tm.em_class_tasks[k] += r.cnts[k]; }
            tm.em_pair_ms += r.cls_ms[kPairList]; tm.em_pair_tasks += r.cnts[kPairCounter];
            for (int k = 0; k < 8; k++) st[k] += r.st[k];
            tm.n_entries += r.blob_bytes;
            tm.launches += r.launches;
            tm.max_slot_bytes = std::max<int32_t>(tm.max_slot_bytes, (int32_t)r.cnts[5]);
        }
        tm.launches += 1;
        tm.run_parts = parts;
        tm.gene_ms = c->n_genes > 0 ? tm.finish_ms : 0.f;
        tm.inner_iters = (int64_t)st[0]; tm.outer_iters = (int64_t)st[1];
        tm.n_tasks = (int64_t)st[4]; tm.n_active = (int64_t)st[5];
        // sum_tasks [in*8*(q+2p)*n_c + out*8*(q+p)*n_c] + n_cells*8*(sum q + sum p)
        tm.em_alg_bytes = 8.0 * ((double)st[2] + (double)st[3]);
        tm.alg_bytes = tm.em_alg_bytes + (double)c->n_cells * 8.0 * (double)(c->n_rows + c->n_cols);
        c->ran = true;
    });
}

int scasa_run(scasa_ctx *c)
{
    if (!c) return fail(c, SCASA_ERR_ARG, "null context");
    return run_staged(c, true);
}

int scasa_set_run_parts(scasa_ctx *c, int32_t n_parts)
{
    if (!c || n_parts < 0) return fail(c, SCASA_ERR_ARG, "bad argument");
    c->run_parts = n_parts;
    return SCASA_OK;
}

int scasa_last_timing(scasa_ctx *c, scasa_timing_t *out)
{
    if (!c || !out) return fail(c, SCASA_ERR_ARG, "null argument");
    if (!c->ran) return fail(c, SCASA_ERR_STATE, "scasa_run has not completed");
    *out = c->timing;
    return SCASA_OK;
}

static int fetch(scasa_ctx *c, void *dst, const void *src, size_t bytes)
{
    if (!c || !dst) return fail(c, SCASA_ERR_ARG, "null argument");
    if (!c->ran) return fail(c, SCASA_ERR_STATE, "scasa_run has not completed");
    return guarded(c, [&] {
        CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        c->wire_d2h += (long long)bytes;
    });
}

int scasa_fetch_iso(scasa_ctx *c, double *out)
{
    if (!c || !out) return fail(c, SCASA_ERR_ARG, "null argument");
    if (!c->ran) return fail(c, SCASA_ERR_STATE, "scasa_run has not completed");
    return guarded(c, [&] { fetch_dense(c, c->d_iso.p, c->n_cells, c->n_cols, out); });
}
int scasa_fetch_iters(scasa_ctx *c, int32_t *out)
{
    return fetch(c, out, c ? c->d_iters.p : nullptr, c ? (size_t)c->n_chunks * c->n_em * 2 * sizeof(int32_t) : 0);
}
int scasa_fetch_colsum(scasa_ctx *c, double *out)
{
    return fetch(c, out, c ? c->d_colsum.p : nullptr, c ? (size_t)c->n_cols * sizeof(double) : 0);
}
int scasa_fetch_gene(scasa_ctx *c, double *out)
{
    if (!c || !out) return fail(c, SCASA_ERR_ARG, "null argument");
    if (!c->ran) return fail(c, SCASA_ERR_STATE, "scasa_run has not completed");
    if (c->n_genes <= 0) return fail(c, SCASA_ERR_STATE, "no gene map set");
    return guarded(c, [&] { fetch_dense(c, c->d_gene.p, c->n_cells, c->n_genes, out); });
}

int scasa_expand_pairs(double *out, int64_t n_rows, int64_t n_cols, const int64_t *row_start, const int32_t *row_nnz,
                       const int32_t *cols, const double *vals, int32_t n_threads)
{
    if (!out || n_rows < 0 || n_cols <= 0 || !row_start || !row_nnz || (!cols && n_rows > 0) || (!vals && n_rows > 0))
        return fail(nullptr, SCASA_ERR_ARG, "bad argument");
    for (int64_t r = 0; r < n_rows; r++) {
        if (row_nnz[r] < 0 || row_nnz[r] > n_cols || row_start[r] < 0) return fail(nullptr, SCASA_ERR_ARG, "bad row extent");
        const int32_t *c = cols + row_start[r];
        for (int32_t k = 0; k < row_nnz[r]; k++)
            if (c[k] < 0 || c[k] >= n_cols || (k > 0 && c[k] <= c[k - 1]))
                return fail(nullptr, SCASA_ERR_ARG, "columns must be strictly ascending inside a row and < n_cols");
    }
    int nthr = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
    nthr = (int)std::max<int64_t>(1, std::min<int64_t>(std::min(nthr, 256), n_rows));
    std::vector<std::thread> workers;                 // expand_rows does not allocate or throw
    for (int t = 0; t < nthr; t++) {
        const int64_t a = n_rows * t / nthr, e = n_rows * (t + 1) / nthr;
        if (a >= e) continue;
        bool started = false;
        try {
            workers.emplace_back(scasa::expand_rows, out, n_cols, a, e, row_start, row_nnz, cols, vals);
            started = true;
        } catch (const std::exception &) {}
        if (!started) scasa::expand_rows(out, n_cols, a, e, row_start, row_nnz, cols, vals);   // no more threads: do the range here
    }
    for (auto &w : workers) w.join();
    return SCASA_OK;
}

int scasa_transfer_bytes(scasa_ctx *c, int64_t *h2d, int64_t *d2h, int32_t reset)
{
    if (!c) return fail(c, SCASA_ERR_ARG, "null context");
    long long a = c->wire_h2d.load(), b = c->wire_d2h.load();
    if (c->twin) { a += c->twin->wire_h2d.load(); b += c->twin->wire_d2h.load(); }
    if (h2d) *h2d = a;
    if (d2h) *d2h = b;
    if (reset) { c->wire_h2d = 0; c->wire_d2h = 0; if (c->twin) { c->twin->wire_h2d = 0; c->twin->wire_d2h = 0; } }
    return SCASA_OK;
}

int scasa_set_host_threads(scasa_ctx *c, int32_t n_threads)
{
    if (!c || n_threads < -1) return fail(c, SCASA_ERR_ARG, "bad argument");
    c->host_threads = n_threads;
    return SCASA_OK;
}

int scasa_fetch_counts(scasa_ctx *c, int64_t *rowptr, int32_t *rowidx, double *val, int64_t cap, int64_t *nnz_out)
{
    if (!c || !rowptr || !nnz_out) return fail(c, SCASA_ERR_ARG, "null argument");
    if (!c->ran) return fail(c, SCASA_ERR_STATE, "scasa_run has not completed");
    return guarded(c, [&] {
        std::vector<int64_t> start((size_t)c->n_cells);
        std::vector<int32_t> len((size_t)c->n_cells), rows((size_t)c->nnz + 1);
        std::vector<double> vals((size_t)c->nnz + 1);
        CK(cudaMemcpyAsync(start.data(), c->d_y_start.p, c->n_cells * sizeof(int64_t), cudaMemcpyDeviceToHost, c->stream));
        CK(cudaMemcpyAsync(len.data(), c->d_y_len.p, c->n_cells * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
        if (c->nnz) {
            CK(cudaMemcpyAsync(rows.data(), c->d_y_row.p, c->nnz * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
            CK(cudaMemcpyAsync(vals.data(), c->d_y_val.p, c->nnz * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        }
        CK(cudaStreamSynchronize(c->stream));
        int64_t n = 0;
        rowptr[0] = 0;
        for (int64_t i = 0; i < c->n_cells; i++) {
            for (int k = 0; k < len[(size_t)i]; k++) {
                if (n < cap && rowidx && val) { rowidx[n] = rows[(size_t)start[(size_t)i] + k]; val[n] = vals[(size_t)start[(size_t)i] + k]; }
                n++;
            }
            rowptr[i + 1] = n;
        }
        *nnz_out = n;
        if (n > cap && rowidx) throw ArgErr{SCASA_ERR_ARG, "count buffers too small (nnz returned)"};
    });
}

int scasa_estimate(scasa_ctx *c, int64_t n_cells, int32_t n_chunks, const int64_t *chunk_off,
                   const int64_t *y_rowptr, const int32_t *y_rowidx, const double *y_val,
                   double *iso_out, int32_t *iters_out)
{
    int rc = scasa_stage_counts(c, n_cells, n_chunks, chunk_off, y_rowptr, y_rowidx, y_val);
    if (rc) return rc;
    rc = scasa_run(c);
    if (rc) return rc;
    rc = scasa_fetch_iso(c, iso_out);
    if (rc) return rc;
    if (iters_out) rc = scasa_fetch_iters(c, iters_out);
    return rc;
}

// One call for a whole sample, host buffers in and out: the chunks are cut into slices and the
// slices alternate between this context and a twin on the same GPU, each driven by its own host
// thread: while one slice's results are expanded into the caller's matrices (host-memory bound)
// the next slice is staged and computed.  Chunks are independent, so slicing changes no result;
// device memory is sized by the slice, not by the sample.
int scasa_quant(scasa_ctx *c, int64_t n_cells, int32_t n_chunks, const int64_t *chunk_off, const int64_t *cell_ptr,
                const int32_t *eq_id, const int32_t *eq_count, int64_t n_eq, const int32_t *eq_row, double *iso_out,
                double *gene_out, double *colsum_out, int32_t *iters_out, int32_t n_slices)
{
    if (!c) return fail(c, SCASA_ERR_ARG, "null context");
    if (n_cells <= 0 || n_chunks <= 0 || !chunk_off || !cell_ptr || !eq_id || !eq_count || !eq_row || !iso_out)
        return fail(c, SCASA_ERR_ARG, "bad argument");
    if (chunk_off[0] != 0 || chunk_off[n_chunks] != n_cells) return fail(c, SCASA_ERR_ARG, "chunk_off must cover [0, n_cells]");
    if (gene_out && c->n_genes <= 0) return fail(c, SCASA_ERR_STATE, "no gene map set");
    if (n_slices <= 0) n_slices = std::max(2, std::min(8, n_chunks / 128));   // measured: 4-8 slices of >= 100 chunks are best
    n_slices = std::max(1, std::min(n_slices, n_chunks));
    if (n_slices > 1) { int rc = sync_twin(c); if (rc) return rc; }
    scasa_ctx *ctxs[2] = {c, n_slices > 1 ? c->twin : c};
    std::vector<double> colsum_parts;
    if (colsum_out) colsum_parts.assign((size_t)n_slices * c->n_cols, 0.0);
    int rcs[2] = {SCASA_OK, SCASA_OK};
    auto drive = [&](int which) {                    // runs on its own thread for which == 1: nothing may escape
        scasa_ctx *x = ctxs[which];
        try {
        std::vector<int64_t> ptr, coff;
        for (int sl = which; sl < n_slices; sl += 2) {
            const int h0 = (int)((int64_t)n_chunks * sl / n_slices), h1 = (int)((int64_t)n_chunks * (sl + 1) / n_slices);
            if (h1 <= h0) continue;
            const int64_t c0 = chunk_off[h0], c1 = chunk_off[h1];
            coff.resize((size_t)(h1 - h0) + 1);
            for (int h = h0; h <= h1; h++) coff[(size_t)(h - h0)] = chunk_off[h] - c0;
            ptr.resize((size_t)(c1 - c0) + 1);
            const int64_t e0 = cell_ptr[c0];
            for (int64_t i = c0; i <= c1; i++) ptr[(size_t)(i - c0)] = cell_ptr[i] - e0;
            int rc = scasa_stage_eqcounts(x, c1 - c0, h1 - h0, coff.data(), ptr.data(), eq_id + e0, eq_count + e0, n_eq, eq_row);
            if (!rc) rc = run_staged(x, false);
            if (!rc) rc = scasa_fetch_iso(x, iso_out + c0 * c->n_cols);
            if (!rc && gene_out) rc = scasa_fetch_gene(x, gene_out + c0 * (int64_t)c->n_genes);
            if (!rc && colsum_out) rc = scasa_fetch_colsum(x, colsum_parts.data() + (size_t)sl * c->n_cols);
            if (!rc && iters_out) rc = scasa_fetch_iters(x, iters_out + (int64_t)h0 * c->n_em * 2);
            if (rc) { rcs[which] = rc; return; }
        }
        } catch (const std::exception &e) {
            rcs[which] = fail(x, SCASA_ERR_NOMEM, e.what());
        }
    };
    if (n_slices > 1) {
        std::thread other;
        try {
            other = std::thread(drive, 1);
        } catch (const std::exception &e) {
            return fail(c, SCASA_ERR_NOMEM, std::string("cannot start the second host thread: ") + e.what());
        }
        drive(0);
        other.join();
    } else drive(0);
    if (rcs[0]) return rcs[0];
    if (rcs[1]) return fail(c, rcs[1], std::string("twin context: ") + scasa_last_error(c->twin));
    if (colsum_out)                                   // slices added in order: a fixed sequence for a given n_slices
        for (int64_t j = 0; j < c->n_cols; j++) {
            double acc = 0;
            for (int sl = 0; sl < n_slices; sl++) acc += colsum_parts[(size_t)sl * c->n_cols + j];
            colsum_out[j] = acc;
        }
    return SCASA_OK;
}

int scasa_gene_sum(scasa_ctx *c, const int32_t *gene_of_col, int32_t n_genes, const double *iso, int64_t n_cells,
                   double *gene_out)
{
    if (!c || !gene_of_col || !iso || !gene_out || n_cells <= 0 || n_genes <= 0) return fail(c, SCASA_ERR_ARG, "bad argument");
    int rc = scasa_set_gene_map(c, gene_of_col, n_genes);
    if (rc) return rc;
    return guarded(c, [&] {
        cudaStream_t s = c->stream;
        c->d_iso.ensure((size_t)(n_cells * c->n_cols));
        c->d_gene.ensure((size_t)n_cells * n_genes);
        CK(cudaMemcpyAsync(c->d_iso.p, iso, n_cells * c->n_cols * sizeof(double), cudaMemcpyHostToDevice, s));
        int launches = 0;
        c->d_colsum.ensure((size_t)c->n_cols);
        launch_finish(c, c->d_iso.p, n_cells, c->d_gene.p, c->d_colsum.p, launches);
        CK(cudaMemcpyAsync(gene_out, c->d_gene.p, (size_t)n_cells * n_genes * sizeof(double), cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
    });
}

int scasa_eqclass_map(scasa_ctx *c, int64_t n_eq, const int64_t *eq_off, const int32_t *eq_tx, int32_t *eq_row,
                      int32_t n_threads)
{
    if (!c || !eq_off || !eq_row || n_eq < 0) return fail(c, SCASA_ERR_ARG, "bad argument");
    if (!c->crp.ready()) return fail(c, SCASA_ERR_STATE, "context was created without CRP");
    return guarded(c, [&] { c->crp.map(n_eq, eq_off, eq_tx, eq_row, n_threads); });
}

int scasa_eqclass_map_host(const scasa_crp_t *crp, int32_t hamming_dist, int64_t n_eq, const int64_t *eq_off,
                           const int32_t *eq_tx, int32_t *eq_row, int32_t n_threads)
{
    if (!crp || !eq_off || !eq_row || n_eq < 0 || hamming_dist < 0) return fail(nullptr, SCASA_ERR_ARG, "bad argument");
    return guarded(nullptr, [&] {
        check_xmat(&crp->m, "crp");
        if (!crp->col_tx || !crp->row_pat || !crp->name_rank) throw ArgErr{SCASA_ERR_ARG, "crp: null dimnames arrays"};
        scasa::CrpIndex idx;
        idx.build(crp->m.n_blocks, crp->m.q_off, crp->m.p_off, crp->m.x_off, crp->m.x, crp->col_tx, crp->row_pat,
                  crp->name_rank, hamming_dist);
        idx.map(n_eq, eq_off, eq_tx, eq_row, n_threads);
    });
}

void *scasa_host_alloc(int64_t bytes)
{
    void *p = nullptr;
    if (bytes <= 0) return nullptr;
    if (cudaHostAlloc(&p, (size_t)bytes, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
void scasa_host_free(void *p) { if (p) cudaFreeHost(p); }

}  // extern "C"
