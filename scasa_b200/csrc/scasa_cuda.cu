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
#include <functional>
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
#include <sys/mman.h>

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
        std::function<void(int64_t, int64_t)> fn;      // set: fn(first row, end row) instead of the expansion
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
    std::shared_ptr<Job> post_fn(int64_t n_rows, std::function<void(int64_t, int64_t)> fn)
    {
        auto j = std::make_shared<Job>();
        j->dst = nullptr; j->n_cols = 0; j->n_rows = n_rows; j->start = nullptr; j->nnz = nullptr; j->cols = nullptr; j->vals = nullptr;
        j->fn = std::move(fn);
        if (n_rows <= 0) { j->done = true; return j; }
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
                if (j->fn) j->fn(r, e);
                else scasa::expand_rows(j->dst, j->n_cols, r, e, j->start, j->nnz, j->cols, j->vals);
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
    std::vector<int32_t> q_off, p_off, em_blocks, em_index, row_block, poor_em, em_order, poor_blocks_list;
    int n_poor = 0, n_emcols = 0;    // poor blocks; columns of all EM blocks
    std::vector<int64_t> x_off;
    std::vector<double> x;
    std::vector<uint8_t> poor;
    int64_t poor_rows = 0;           // sum of q over poor blocks
    DBuf<int32_t> d_q_off, d_p_off, d_em_blocks, d_em_order, d_em_col0, d_poor_rank, d_poor_col0, d_emcol_col, d_col_emcol;
    DBuf<int2> d_row_info;
    DBuf<int4> d_row_aux;
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
    DBuf<int4> d_desc;
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
    int pair_ctas_per_sm = 12;
    int poison = 0;                  // SCASA_POISON=1 (testing): result and blob buffers start as 0xFF bytes in every run
    DBuf<unsigned long long> d_stats;
    DBuf<int> d_flag;

    // ---- results: CSR by cell over the structure (scasa_kernels.cuh: struct Result); no dense matrix on the device
    DBuf<int32_t> d_rs_len, d_rs_col, d_poor_pos, d_g_col, d_g_len, d_gene_of_col;
    DBuf<int64_t> d_rs_ptr;
    DBuf<double> d_rs_val, d_g_val, d_colsum, d_colpart, d_dense_tmp, d_iso_in, d_gene_out;
    DBuf<uint32_t> d_ypos;
    int64_t rs_nnz = 0;                  // entries of the result structure of the last run
    std::vector<int64_t> rs_ptr_host;    // fetched on demand after a run
    std::vector<int64_t> cmp_ptr_host;   // row starts of the result packed without its stored zeros (row_source)
    DBuf<int32_t> d_cmp_len, d_cmp_col;
    DBuf<int64_t> d_cmp_ptr;
    DBuf<double> d_cmp_val;
    int d2h_compact = 1;                 // drop the stored zeros on the device before a result crosses PCIe (SCASA_D2H_COMPACT=0: off)
    std::vector<int32_t> g_len_host;
    bool have_rs_host = false, have_glen_host = false, ran_genes = false;
    // the trip to the host: two batches in flight (device CSR slice -> pinned staging -> threaded expansion)
    struct FetchBuf {
        int32_t *h_cols = nullptr, *h_nnz = nullptr;
        double *h_vals = nullptr;
        int64_t *h_start = nullptr;
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
    int pack_threads = 512;                    // threads per CTA of pack_cells_dense_kernel (SCASA_PACK_THREADS=512|1024)
    int gene_ctas_per_sm = 2;                  // gene_rows_kernel: resident CTAs per SM (SCASA_GENE_CTAS)
    int gene_n_multi = 0;                      // genes with more than one isoform column
    DBuf<int2> d_col_gene;                     // per column: {gene or -1, slot in the gene kernel's member table or -1}
    DBuf<int32_t> d_gene_stops;                // the member table's stop marks
    int gene_slots = 1;
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
    X.em_order = c->d_em_order.p; X.row_info = c->d_row_info.p; X.row_aux = c->d_row_aux.p;
    X.em_col0 = c->d_em_col0.p; X.poor_rank = c->d_poor_rank.p; X.poor_col0 = c->d_poor_col0.p;
    X.n_poor = c->n_poor; X.n_emcols = c->n_emcols;
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
Result rdev(scasa_ctx *c)
{
    Result R;
    R.rs_ptr = c->d_rs_ptr.p; R.rs_col = c->d_rs_col.p; R.rs_val = c->d_rs_val.p; R.ypos = c->d_ypos.p;
    R.poor_pos = c->d_poor_pos.p; R.colsum = c->d_colsum.p; R.colpart = c->d_colpart.p;
    R.g_col = c->d_g_col.p; R.g_val = c->d_g_val.p; R.g_len = c->d_g_len.p;
    return R;
}
Tasks tdev(scasa_ctx *c)
{
    Tasks T;
    T.cnt = c->d_cnt.p; T.runs = c->d_runs.p; T.tbytes = c->d_tbytes.p;
    T.toff = c->d_toff.p;
    T.tdata = c->d_tdata.p;
    for (int k = 0; k < kLists; k++) T.list[k] = c->d_list[k].p;
    T.desc = c->d_desc.p;
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

// scasa_genemat on a dense isoform matrix resident at iso (scasa_gene_sum: a matrix the caller already holds)
void launch_gene_gather(scasa_ctx *w, const double *iso, int64_t n_cells, double *gene)
{
    const int tiles = (w->n_genes + kGeneThreads * kGeneIlp - 1) / (kGeneThreads * kGeneIlp);
    if (n_cells * tiles >= ((int64_t)1 << 31)) throw ArgErr{SCASA_ERR_UNSUPPORTED, "too many cells x genes for one launch"};
    gene_gather_kernel<<<(unsigned)(n_cells * tiles), kGeneThreads, 0, w->stream>>>(iso, n_cells, w->n_cols, w->d_gene_ptr.p,
                                                                                   w->d_gene_cols.p, w->n_genes, tiles, gene);
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
    c->staged = true; c->ran = false; c->have_rs_host = false; c->have_glen_host = false; c->ran_genes = false;
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
    if (b.h_cols) { cudaFreeHost(b.h_cols); cudaFreeHost(b.h_vals); cudaFreeHost(b.h_start); cudaFreeHost(b.h_nnz); }
    b.h_cols = nullptr; b.rows = 0; b.cap = 0;
    CK(cudaHostAlloc((void **)&b.h_cols, cap * sizeof(int32_t), cudaHostAllocDefault));
    CK(cudaHostAlloc((void **)&b.h_vals, cap * sizeof(double), cudaHostAllocDefault));
    CK(cudaHostAlloc((void **)&b.h_start, rows * sizeof(int64_t), cudaHostAllocDefault));
    CK(cudaHostAlloc((void **)&b.h_nnz, rows * sizeof(int32_t), cudaHostAllocDefault));
    if (!b.done) CK(cudaEventCreateWithFlags(&b.done, cudaEventDisableTiming));
    b.rows = rows; b.cap = cap;
}

// host copies of the row pointers (and gene row lengths) of the last run, fetched once per run
void need_rs_host(scasa_ctx *c, bool genes)
{
    if (!c->have_rs_host) {
        c->rs_ptr_host.resize((size_t)c->n_cells + 1);
        CK(cudaMemcpyAsync(c->rs_ptr_host.data(), c->d_rs_ptr.p, ((size_t)c->n_cells + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        c->wire_d2h += (long long)(c->n_cells + 1) * 8;
        c->have_rs_host = true;
    }
    if (genes && !c->have_glen_host) {
        c->g_len_host.resize((size_t)c->n_cells);
        CK(cudaMemcpyAsync(c->g_len_host.data(), c->d_g_len.p, (size_t)c->n_cells * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        c->wire_d2h += (long long)c->n_cells * 4;
        c->have_glen_host = true;
    }
}

// Where the rows of a result are read from: as the kernels left them (row r = [ptr[r], ptr[r] + len[r]), len == nullptr:
// full rows), or — compact — packed on the device without their stored zeros.
struct RowSrc { const int64_t *ptr; const int32_t *len; const int32_t *d_col; const double *d_val; };
RowSrc row_source(scasa_ctx *c, bool genes, bool compact, bool keep_all = false)
{
    need_rs_host(c, genes);
    RowSrc src{c->rs_ptr_host.data(), genes ? c->g_len_host.data() : nullptr, genes ? c->d_g_col.p : c->d_rs_col.p,
               genes ? c->d_g_val.p : c->d_rs_val.p};
    if (!compact || c->n_cells == 0) return src;
    cudaStream_t s = c->stream;
    const int64_t n = c->n_cells;
    const unsigned grid = (unsigned)((n + 7) / 8);
    int launches = 0;
    c->d_cmp_len.ensure((size_t)n + 1);
    c->d_cmp_ptr.ensure((size_t)n + 2);
    row_nonzeros_kernel<<<grid, 256, 0, s>>>(c->d_rs_ptr.p, genes ? c->d_g_len.p : nullptr, src.d_val, n, c->d_cmp_len.p, keep_all);
    CK(cudaGetLastError());
    exclusive_scan(c, c->d_cmp_len.p, n, c->d_cmp_ptr.p, c->d_totals.p + 2, launches);
    c->cmp_ptr_host.resize((size_t)n + 1);
    CK(cudaMemcpyAsync(c->cmp_ptr_host.data(), c->d_cmp_ptr.p, ((size_t)n + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    c->wire_d2h += (long long)(n + 1) * 8;
    const int64_t total = c->cmp_ptr_host[(size_t)n];
    c->d_cmp_col.ensure((size_t)total + 1);
    c->d_cmp_val.ensure((size_t)total + 1);
    row_compact_kernel<<<grid, 256, 0, s>>>(c->d_rs_ptr.p, genes ? c->d_g_len.p : nullptr, src.d_col, src.d_val, n, c->d_cmp_ptr.p,
                                            c->d_cmp_col.p, c->d_cmp_val.p, keep_all);
    CK(cudaGetLastError());
    return RowSrc{c->cmp_ptr_host.data(), nullptr, c->d_cmp_col.p, c->d_cmp_val.p};
}

// A CSR result (rows of the isoform matrix, or of the gene matrix when `genes`) -> the caller's dense
// host matrix.  The sparse rows cross PCIe as they are, in batches of whole rows; while the host threads
// expand batch k into the caller's matrix (every element written exactly once, streaming stores), batch
// k+1 is on the wire.  The matrix ends up bit-identical to what a dense copy would give.
void fetch_dense(scasa_ctx *c, bool genes, double *out)
{
    cudaStream_t s = c->stream;
    const int64_t n_rows = c->n_cells, n_cols = genes ? c->n_genes : c->n_cols;
    const bool device_expand = c->fetch_dense_copy || c->host_threads < 0;
    const RowSrc src = row_source(c, genes, c->d2h_compact && !device_expand);   // zeros need not cross PCIe: the expansion writes them
    const int64_t *ptr = src.ptr;
    const int32_t *d_col = src.d_col;
    const double *d_val = src.d_val;
    if (device_expand) {           // expansion on the device, plain dense copy (testing aid)
        const int64_t batch = std::max<int64_t>(1, std::min<int64_t>(n_rows, ((int64_t)32 << 20) / n_cols));
        c->d_dense_tmp.ensure((size_t)(batch * n_cols));
        for (int64_t r0 = 0; r0 < n_rows; r0 += batch) {
            const int64_t nr = std::min(batch, n_rows - r0);
            expand_rows_kernel<<<(unsigned)std::min<int64_t>(nr, 4096), 256, 0, s>>>(c->d_rs_ptr.p, genes ? c->d_g_len.p : nullptr, d_col, d_val,
                                                                                   r0, nr, n_cols, c->d_dense_tmp.p);
            CK(cudaGetLastError());
            CK(cudaMemcpyAsync(out + r0 * n_cols, c->d_dense_tmp.p, (size_t)(nr * n_cols) * sizeof(double), cudaMemcpyDeviceToHost, s));
            CK(cudaStreamSynchronize(s));
            c->wire_d2h += (long long)(nr * n_cols) * 8;
        }
        return;
    }
    int nthr = c->host_threads > 0 ? c->host_threads : (int)std::thread::hardware_concurrency();
    nthr = std::max(1, std::min(nthr, 256));
    // batches of whole rows holding at most `cap` entries (the largest row always fits)
    int64_t max_row = 1;
    for (int64_t r = 0; r < n_rows; r++) max_row = std::max(max_row, ptr[r + 1] - ptr[r]);
    const size_t cap = (size_t)std::max<int64_t>(max_row, (int64_t)6 << 20);
    std::vector<int64_t> cut{0};
    for (int64_t r = 0; r < n_rows;) {
        int64_t e = r + 1;
        while (e < n_rows && ptr[e + 1] - ptr[r] <= (int64_t)cap && e - r < 65536) e++;
        cut.push_back(e);
        r = e;
    }
    const int64_t n_batches = (int64_t)cut.size() - 1;
    size_t max_rows = 1;
    for (int64_t k = 0; k < n_batches; k++) max_rows = std::max(max_rows, (size_t)(cut[(size_t)k + 1] - cut[(size_t)k]));
    for (auto &b : c->fb) fetch_buf_ensure(c, b, max_rows, cap);
    auto launch = [&](int64_t k) {
        scasa_ctx::FetchBuf &b = c->fb[k & 1];
        const int64_t r0 = cut[(size_t)k], r1 = cut[(size_t)k + 1];
        const int64_t e0 = ptr[r0], ne = ptr[r1] - e0;
        if (ne) {
            CK(cudaMemcpyAsync(b.h_cols, d_col + e0, (size_t)ne * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
            CK(cudaMemcpyAsync(b.h_vals, d_val + e0, (size_t)ne * sizeof(double), cudaMemcpyDeviceToHost, s));
        }
        for (int64_t r = r0; r < r1; r++) {
            b.h_start[r - r0] = ptr[r] - e0;
            b.h_nnz[r - r0] = src.len ? src.len[r] : (int32_t)(ptr[r + 1] - ptr[r]);
        }
        c->wire_d2h += (long long)ne * 12;
        CK(cudaEventRecord(b.done, s));
    };
    ExpandPool &pool = expand_pool(c, nthr);
    std::shared_ptr<ExpandPool::Job> pending;           // the batch being expanded (it reads the other staging buffer)
    auto join_all = [&]() { pool.wait(pending); pending.reset(); };
    struct Joiner { decltype(join_all) &j; ~Joiner() { j(); } } joiner{join_all};   // an exception must not leave a job reading freed buffers
    const bool trace = getenv("SCASA_FETCH_TRACE") != nullptr;
    double t_wait_gpu = 0, t_join = 0;
    auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    launch(0);
    for (int64_t k = 0; k < n_batches; k++) {
        scasa_ctx::FetchBuf &b = c->fb[k & 1];
        double t0 = now();
        CK(cudaEventSynchronize(b.done));
        t_wait_gpu += now() - t0;
        t0 = now();
        join_all();                                   // batch k-1 (it reads the other buffer, which launch(k+1) reuses)
        t_join += now() - t0;
        const int64_t r0 = cut[(size_t)k], nr = cut[(size_t)k + 1] - r0;
        pending = pool.post(out + r0 * n_cols, n_cols, nr, b.h_start, b.h_nnz, b.h_cols, b.h_vals);
        if (k + 1 < n_batches) launch(k + 1);
    }
    double t0 = now();
    join_all();
    t_join += now() - t0;
    if (trace)
        fprintf(stderr, "[scasa fetch] %lld x %lld, %lld batches, %lld entries (%.1f per row): wait copy %.3f s, wait expansion %.3f s, %d threads\n",
                (long long)n_rows, (long long)n_cols, (long long)n_batches, (long long)ptr[n_rows], (double)ptr[n_rows] / (double)n_rows,
                t_wait_gpu, t_join, nthr);
}

// A CSR result -> sparse rows on the host, through the same pipeline (pinned staging, batches of whole rows, host
// threads): row r's entries go to cols/vals (either may be null) at dst_off[r]; with drop0 the stored zeros of a row
// are left out (its entries move up inside the row's span); count[r] (nullable) = entries written for row r.
// dst_off == nullptr: rows side by side without gaps (isoform rows: the device layout; gene rows: closed up).
void fetch_rows(scasa_ctx *c, const RowSrc &src, int64_t n_rows, int32_t *cols, double *vals, const int64_t *dst_off, int64_t *count,
                bool drop0);
void fetch_sparse(scasa_ctx *c, bool genes, int32_t *cols, double *vals, const int64_t *dst_off, int64_t *count, bool drop0)
{
    const bool packed = drop0 && c->d2h_compact;                  // zeros already left behind on the device
    const RowSrc src = row_source(c, genes, packed);
    fetch_rows(c, src, c->n_cells, cols, vals, dst_off, count, drop0 && !packed);
}
// the transfer itself; src may also be rows parked on the device by an earlier slice (HeldRows)
void fetch_rows(scasa_ctx *c, const RowSrc &src, int64_t n_rows, int32_t *cols, double *vals, const int64_t *dst_off, int64_t *count,
                bool drop0)
{
    cudaStream_t s = c->stream;
    const int64_t *ptr = src.ptr;
    const int32_t *d_col = src.d_col;
    const double *d_val = src.d_val;
    const int32_t *glen = src.len;
    std::vector<int64_t> own_off;
    if (!dst_off) {
        own_off.resize((size_t)n_rows + 1);
        own_off[0] = 0;
        for (int64_t r = 0; r < n_rows; r++) own_off[(size_t)r + 1] = own_off[(size_t)r] + (glen ? glen[r] : ptr[r + 1] - ptr[r]);
        dst_off = own_off.data();
    }
    int nthr = c->host_threads > 0 ? c->host_threads : (int)std::thread::hardware_concurrency();
    nthr = std::max(1, std::min(nthr, 256));
    int64_t max_row = 1;
    for (int64_t r = 0; r < n_rows; r++) max_row = std::max(max_row, ptr[r + 1] - ptr[r]);
    const size_t cap = (size_t)std::max<int64_t>(max_row, (int64_t)6 << 20);
    std::vector<int64_t> cut{0};
    for (int64_t r = 0; r < n_rows;) {
        int64_t e = r + 1;
        while (e < n_rows && ptr[e + 1] - ptr[r] <= (int64_t)cap && e - r < 65536) e++;
        cut.push_back(e);
        r = e;
    }
    const int64_t n_batches = (int64_t)cut.size() - 1;
    size_t max_rows = 1;
    for (int64_t k = 0; k < n_batches; k++) max_rows = std::max(max_rows, (size_t)(cut[(size_t)k + 1] - cut[(size_t)k]));
    for (auto &b : c->fb) fetch_buf_ensure(c, b, max_rows, cap);
    auto launch = [&](int64_t k) {
        scasa_ctx::FetchBuf &b = c->fb[k & 1];
        const int64_t r0 = cut[(size_t)k], r1 = cut[(size_t)k + 1];
        const int64_t e0 = ptr[r0], ne = ptr[r1] - e0;
        if (ne) {
            if (cols) CK(cudaMemcpyAsync(b.h_cols, d_col + e0, (size_t)ne * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
            if (vals || drop0) CK(cudaMemcpyAsync(b.h_vals, d_val + e0, (size_t)ne * sizeof(double), cudaMemcpyDeviceToHost, s));
        }
        c->wire_d2h += (long long)ne * ((cols ? 4 : 0) + (vals || drop0 ? 8 : 0));
        CK(cudaEventRecord(b.done, s));
    };
    ExpandPool &pool = expand_pool(c, nthr);
    std::shared_ptr<ExpandPool::Job> pending;
    auto join_all = [&]() { pool.wait(pending); pending.reset(); };
    struct Joiner { decltype(join_all) &j; ~Joiner() { j(); } } joiner{join_all};
    if (n_batches > 0) launch(0);
    for (int64_t k = 0; k < n_batches; k++) {
        scasa_ctx::FetchBuf &b = c->fb[k & 1];
        CK(cudaEventSynchronize(b.done));
        join_all();                                   // batch k-1 reads the buffer that launch(k+1) reuses
        const int64_t r0 = cut[(size_t)k], nr = cut[(size_t)k + 1] - r0, e0 = ptr[r0];
        const int32_t *hc = b.h_cols;
        const double *hv = b.h_vals;
        pending = pool.post_fn(nr, [=](int64_t a, int64_t e) {
            for (int64_t i = a; i < e; i++) {
                const int64_t r = r0 + i, src = ptr[r] - e0, m = glen ? glen[r] : ptr[r + 1] - ptr[r], d = dst_off[r];
                if (!drop0) {
                    if (cols) memcpy(cols + d, hc + src, (size_t)m * sizeof(int32_t));
                    if (vals) memcpy(vals + d, hv + src, (size_t)m * sizeof(double));
                    if (count) count[r] = m;
                } else {
                    int64_t o = d;
                    for (int64_t t = src; t < src + m; t++)
                        if (hv[t] != 0.0) { if (cols) cols[o] = hc[t]; if (vals) vals[o] = hv[t]; o++; }
                    if (count) count[r] = o - d;
                }
            }
        });
        if (k + 1 < n_batches) launch(k + 1);
    }
    join_all();
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
                if (c->x_off[(size_t)k] > 0x7fffffff) throw ArgErr{SCASA_ERR_UNSUPPORTED, "the X-matrix may hold at most 2^31 values"};
                cost[(size_t)e] = q * p * p;
            }
            std::stable_sort(c->em_order.begin(), c->em_order.end(),
                             [&](int32_t x, int32_t y) { return cost[(size_t)x] > cost[(size_t)y]; });
        }
        if (getenv("SCASA_FETCH_DENSE")) c->fetch_dense_copy = 1;
        if (const char *env = getenv("SCASA_D2H_COMPACT")) c->d2h_compact = atoi(env) != 0;
        if (getenv("SCASA_POISON")) c->poison = 1;
        if (const char *env = getenv("SCASA_GENE_CTAS")) c->gene_ctas_per_sm = std::max(1, atoi(env));
        if (const char *env = getenv("SCASA_PACK_THREADS")) c->pack_threads = atoi(env) == 1024 ? 1024 : 512;
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
            int m = 16, cps = 12;
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
            std::vector<int4> aux((size_t)c->n_rows);
            std::vector<int32_t> poor_rank((size_t)B), poor_col0, em_col0((size_t)c->n_em + 1, 0), emcol_col;
            int np = 0;
            for (int k = 0; k < B; k++) {
                poor_rank[(size_t)k] = np;
                for (int r = c->q_off[k]; r < c->q_off[k + 1]; r++) {
                    info[(size_t)r] = make_int2(c->em_index[k], r - c->q_off[k]);
                    const int e = c->em_index[k], p = c->p_off[k + 1] - c->p_off[k];
                    aux[(size_t)r] = make_int4(c->p_off[k], np, e < 0 ? 0 : (p | (c->poor[k] ? 1 << 16 : 0)), e);
                }
                if (c->poor[k]) { np++; poor_col0.push_back(c->p_off[k]); c->poor_blocks_list.push_back(k); }
            }
            c->n_poor = np;
            for (int e = 0; e < c->n_em; e++) {
                const int k = c->em_blocks[(size_t)e], p = c->p_off[k + 1] - c->p_off[k];
                if (p > 0xffff) throw ArgErr{SCASA_ERR_UNSUPPORTED, "a block may have at most 65535 columns"};
                em_col0[(size_t)e + 1] = em_col0[(size_t)e] + p;
                for (int j = 0; j < p; j++) emcol_col.push_back(c->p_off[k] + j);
            }
            c->n_emcols = em_col0[(size_t)c->n_em];
            c->d_row_info.upload(info, c->stream);
            c->d_row_aux.upload(aux, c->stream);
            c->d_em_col0.upload(em_col0, c->stream);
            c->d_poor_rank.upload(poor_rank, c->stream);
            c->d_poor_col0.ensure(poor_col0.size() + 1); c->d_poor_col0.upload(poor_col0, c->stream);
            c->d_emcol_col.ensure(emcol_col.size() + 1); c->d_emcol_col.upload(emcol_col, c->stream);
            std::vector<int32_t> col_emcol((size_t)c->n_cols, -1);
            for (size_t j = 0; j < emcol_col.size(); j++) col_emcol[(size_t)emcol_col[j]] = (int32_t)j;
            c->d_col_emcol.upload(col_emcol, c->stream);
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
    DBuf<int32_t> *i32[] = {&c->d_q_off, &c->d_p_off, &c->d_em_blocks, &c->d_em_order, &c->d_em_col0, &c->d_poor_rank, &c->d_poor_col0,
                            &c->d_emcol_col, &c->d_col_emcol, &c->d_cell_chunk, &c->d_y_len, &c->d_y_row, &c->d_eq_id, &c->d_eq_count, &c->d_eq_row,
                            &c->d_cnt, &c->d_runs, &c->d_tbytes, &c->d_iters,
                            &c->d_list[0], &c->d_list[1], &c->d_list[2], &c->d_list[3], &c->d_list[4],
                            &c->d_gene_ptr, &c->d_gene_cols, &c->d_rs_len, &c->d_rs_col, &c->d_poor_pos, &c->d_g_col, &c->d_g_len,
                            &c->d_gene_of_col};
    for (auto b : i32) b->release();
    DBuf<int64_t> *i64[] = {&c->d_x_off, &c->d_chunk_off, &c->d_y_start, &c->d_cell_ptr, &c->d_toff, &c->d_scan_sums, &c->d_totals,
                            &c->d_rs_ptr};
    for (auto b : i64) b->release();
    DBuf<double> *f64[] = {&c->d_x, &c->d_y_val, &c->d_rs_val, &c->d_g_val, &c->d_colsum, &c->d_colpart,
                           &c->d_dense_tmp, &c->d_iso_in, &c->d_gene_out};
    for (auto b : f64) b->release();
    c->d_poor.release(); c->d_next.release(); c->d_stats.release(); c->d_flag.release();
    c->d_tdata.release(); c->d_counters.release(); c->d_row_info.release(); c->d_row_aux.release();
    c->d_scatter_scratch.release(); c->d_ypos.release(); c->d_col_gene.release(); c->d_gene_stops.release(); c->d_desc.release();
    c->d_cmp_len.release(); c->d_cmp_col.release(); c->d_cmp_ptr.release(); c->d_cmp_val.release();
    for (auto &b : c->fb) {
        if (b.h_cols) { cudaFreeHost(b.h_cols); cudaFreeHost(b.h_vals); cudaFreeHost(b.h_start); cudaFreeHost(b.h_nnz); }
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
        c->d_gene_of_col.upload(goc, c->stream);
        {
            // member table of gene_rows_kernel: the members of every multi-isoform gene side by side, a stop mark
            // in front of the first gene and behind every gene
            std::vector<int2> cg((size_t)c->n_cols, make_int2(-1, -1));
            std::vector<int32_t> stops{0};
            int n_multi = 0, slot = 1;
            for (int g = 0; g < n_genes; g++) {
                const int nm = ptr[(size_t)g + 1] - ptr[(size_t)g];
                for (int m = ptr[(size_t)g]; m < ptr[(size_t)g + 1]; m++)
                    cg[(size_t)cols[(size_t)m]] = make_int2(g, nm > 1 ? slot++ : -1);
                if (nm > 1) { stops.push_back(slot++); n_multi++; }
            }
            c->gene_n_multi = n_multi; c->gene_slots = slot;
            c->d_gene_stops.upload(stops, c->stream);
            c->d_col_gene.upload(cg, c->stream);
        }
        CK(cudaStreamSynchronize(c->stream));
    });
}

// What one range of chunks contributed to the run statistics.
struct RangeStats {
    float pack_ms = 0, tasks_ms = 0, em_ms = 0, finish_ms = 0, gene_ms = 0, cls_ms[kLists] = {0, 0, 0, 0, 0};
    unsigned long long st[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    unsigned int cnts[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int64_t blob_bytes = 0;
    int launches = 0;
};

// The sample staged in c: pack -> result structure + tasks -> EM -> column sums -> gene sums.
static void run_all(scasa_ctx *c, RangeStats &rs)
{
    scasa_ctx *w = c;
    cudaStream_t s = c->stream;
    int launches = 0;
    const int n_chunks = c->n_chunks;
    const int64_t n_cells = c->n_cells;
    const int64_t n_tasks = (int64_t)n_chunks * w->n_em;
    w->d_cnt.ensure((size_t)n_tasks + 1); w->d_runs.ensure((size_t)n_tasks + 1);
    w->d_tbytes.ensure((size_t)n_tasks + 1); w->d_toff.ensure((size_t)n_tasks + 1);
    for (int k = 0; k < kLists; k++) w->d_list[k].ensure((size_t)n_tasks + 1);
    w->d_desc.ensure((size_t)n_tasks + 1);
    c->d_rs_len.ensure((size_t)n_cells + 1); c->d_rs_ptr.ensure((size_t)n_cells + 1);
    c->d_poor_pos.ensure((size_t)n_cells * std::max(1, c->n_poor));
    c->d_ypos.ensure((size_t)c->nnz + 1);
    c->d_colpart.ensure((size_t)n_chunks * std::max(1, c->n_emcols));
    if (c->n_genes > 0) c->d_g_len.ensure((size_t)n_cells);

    CK(cudaEventRecord(w->ev[0], s));
    CK(cudaMemsetAsync(w->d_cnt.p, 0, (n_tasks + 1) * sizeof(int32_t), s));
    CK(cudaMemsetAsync(w->d_runs.p, 0, (n_tasks + 1) * sizeof(int32_t), s));
    CK(cudaMemsetAsync(w->d_stats.p, 0, 8 * sizeof(unsigned long long), s));
    CK(cudaMemsetAsync(w->d_next.p, 0, 16 * sizeof(unsigned int), s));
    CK(cudaMemsetAsync(w->d_counters.p, 0, 8 * sizeof(unsigned int), s));
    CK(cudaMemsetAsync(w->d_flag.p, 0, sizeof(int), s));
    CK(cudaMemsetAsync(c->d_colsum.p, 0, (size_t)c->n_cols * sizeof(double), s));

    // ---- pack: eqclass counts -> Y
    if (c->have_eq) {
        int n_words = (int)std::min<int64_t>((c->n_rows + 31) / 32, 3 * 512);   // 49,152 rows per tile (198 KB); more rows: several tiles
        if (const char *env = getenv("SCASA_PACK_WPT")) n_words = std::min(n_words, 512 * std::max(1, std::min(3, atoi(env))));   // smaller tiles: more CTAs per SM, more passes
        const size_t dense_smem = (size_t)n_words * 33 * sizeof(int);         // int32 histogram + touched-row bitmap
        if (!getenv("SCASA_PACK_SORT")) {
            const int pack_ctas = (int)std::max<size_t>(1, (size_t)(220 * 1024) / (dense_smem + 1024));
            int grid = (int)std::min<int64_t>(n_cells, (int64_t)w->n_sm * pack_ctas);
            if (c->pack_threads == 1024) {
                CK(cudaFuncSetAttribute(pack_cells_dense_kernel<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dense_smem));
                pack_cells_dense_kernel<1024><<<grid, 1024, dense_smem, s>>>(n_cells, c->d_cell_ptr.p, c->d_eq_id.p, c->d_eq_count.p,
                                                                           c->d_eq_row.p, c->n_eq, (int)c->n_rows, n_words, c->d_y_row.p,
                                                                           c->d_y_val.p, c->d_y_len.p);
            } else {
                CK(cudaFuncSetAttribute(pack_cells_dense_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dense_smem));
                pack_cells_dense_kernel<512><<<grid, 512, dense_smem, s>>>(n_cells, c->d_cell_ptr.p, c->d_eq_id.p, c->d_eq_count.p,
                                                                         c->d_eq_row.p, c->n_eq, (int)c->n_rows, n_words, c->d_y_row.p,
                                                                         c->d_y_val.p, c->d_y_len.p);
            }
        } else {
            int cap = 32;
            while (cap < c->max_cell_entries) cap <<= 1;
            size_t smem = (size_t)cap * sizeof(unsigned long long);
            CK(cudaFuncSetAttribute(pack_cells_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            int grid = (int)std::min<int64_t>(n_cells, (int64_t)w->n_sm * 8);
            pack_cells_kernel<<<grid, 256, smem, s>>>(n_cells, c->d_cell_ptr.p, c->d_eq_id.p, c->d_eq_count.p,
                                                      c->d_eq_row.p, c->n_eq, cap, c->d_y_row.p, c->d_y_val.p,
                                                      c->d_y_len.p, w->d_flag.p);
        }
        CK(cudaGetLastError());
        launches++;
    }
    CK(cudaEventRecord(w->ev[1], s));

    XDev X = xdev(w);
    Sample S = sdev(c, 0, n_chunks);
    // ---- task extents and work classes; size of every result row; the two scans
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
    int32_t *iters = c->d_iters.p;
    int64_t tot[2] = {0, 0};
    unsigned int *cnts = rs.cnts;
    {
        const int wpb = 8;
        task_count_kernel<<<(unsigned)((n_cells + wpb - 1) / wpb), wpb * 32, 0, s>>>(X, S, T, c->d_rs_len.p);
        launches++;
        if (w->n_em > 0) {
            classify_kernel<<<(unsigned)((n_tasks + 255) / 256), 256, 0, s>>>(X, S, T, n_tasks, O, CC, iters, w->d_stats.p);
            launches++;
            exclusive_scan(w, w->d_tbytes.p, n_tasks, w->d_toff.p, w->d_totals.p + 0, launches);
        } else CK(cudaMemsetAsync(w->d_totals.p, 0, sizeof(int64_t), s));
        exclusive_scan(w, c->d_rs_len.p, n_cells, c->d_rs_ptr.p, w->d_totals.p + 1, launches);
        CK(cudaMemcpyAsync(tot, w->d_totals.p, sizeof tot, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(cnts, w->d_counters.p, 8 * sizeof(unsigned int), cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        if (cnts[7] == 1) throw ArgErr{SCASA_ERR_UNSUPPORTED, "a poor-condition (2-column) block may have at most 64 rows"};
        if (cnts[7] == 2) throw ArgErr{SCASA_ERR_UNSUPPORTED, "a (chunk, block) task may hold at most 65535 cells-with-counts and 65535 count entries"};
        if (cnts[5] > 200 * 1024) throw ArgErr{SCASA_ERR_UNSUPPORTED, "a (chunk, block) task needs more than 200 KB of shared memory (chunk too large for this block)"};
    }
    c->rs_nnz = tot[1];
    c->d_rs_col.ensure((size_t)tot[1] + 2); c->d_rs_val.ensure((size_t)tot[1] + 2);
    if (c->n_genes > 0) { c->d_g_col.ensure((size_t)tot[1] + 2); c->d_g_val.ensure((size_t)tot[1] + 2); }
    if (c->poison) {       // SCASA_POISON: every byte the kernels are supposed to write starts as 0xFF (NaN / -1): a test finds what was left
        CK(cudaMemsetAsync(c->d_rs_col.p, 0xFF, ((size_t)tot[1] + 2) * sizeof(int32_t), s));
        CK(cudaMemsetAsync(c->d_rs_val.p, 0xFF, ((size_t)tot[1] + 2) * sizeof(double), s));
        CK(cudaMemsetAsync(c->d_ypos.p, 0xFF, ((size_t)c->nnz + 1) * sizeof(uint32_t), s));
        CK(cudaMemsetAsync(c->d_poor_pos.p, 0xFF, (size_t)n_cells * std::max(1, c->n_poor) * sizeof(int32_t), s));
        CK(cudaMemsetAsync(c->d_colpart.p, 0xFF, (size_t)n_chunks * std::max(1, c->n_emcols) * sizeof(double), s));
        if (c->n_genes > 0) {
            CK(cudaMemsetAsync(c->d_g_col.p, 0xFF, ((size_t)tot[1] + 2) * sizeof(int32_t), s));
            CK(cudaMemsetAsync(c->d_g_val.p, 0xFF, ((size_t)tot[1] + 2) * sizeof(double), s));
            CK(cudaMemsetAsync(c->d_g_len.p, 0xFF, (size_t)n_cells * sizeof(int32_t), s));
        }
    }
    Result R = rdev(c);
    {
        const int wpb = 8;
        struct_fill_kernel<<<(unsigned)((n_cells + wpb - 1) / wpb), wpb * 32, 0, s>>>(X, S, R);
        CK(cudaGetLastError());
        launches++;
    }
    if (w->n_em > 0) {
        w->d_tdata.ensure((size_t)tot[0] + 64);
        if (c->poison) CK(cudaMemsetAsync(w->d_tdata.p, 0xFF, (size_t)tot[0] + 64, s));
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
            task_scatter_kernel<true><<<n_chunks, kScatterThreads, sc_smem, s>>>(X, S, T, R, scratch);
        } else {
            if (sc_smem) CK(cudaFuncSetAttribute(task_scatter_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sc_smem));
            task_scatter_kernel<false><<<n_chunks, kScatterThreads, sc_smem, s>>>(X, S, T, R, scratch);
        }
        CK(cudaGetLastError());
        launches++;
    }
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
            A.X = X; A.S = S; A.T = T; A.O = O; A.R = R; A.iters = iters; A.stats = w->d_stats.p;
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
            A.X = X; A.S = S; A.T = T; A.O = O; A.R = R; A.iters = iters; A.stats = w->d_stats.p;
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

    // ---- column sums of the EM columns (chunks in order), gene sums
    if (c->n_emcols > 0) {
        const int cs_tile = std::min(c->n_emcols, 12288);          // 96 KB of column slots: two CTAs per SM
        const size_t cs_smem = (size_t)cs_tile * sizeof(double);
        CK(cudaFuncSetAttribute(colsum_chunk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cs_smem));
        colsum_chunk_kernel<<<n_chunks, kColsumThreads, cs_smem, s>>>(S, R, c->d_col_emcol.p, c->n_emcols, cs_tile);
        launches++;
        colsum_em_reduce<<<(unsigned)((c->n_emcols + 255) / 256), 256, 0, s>>>(c->d_colpart.p, n_chunks, c->n_emcols, c->d_emcol_col.p,
                                                                              c->d_colsum.p);
        launches++;
    }
    CK(cudaEventRecord(w->ev[4], s));
    if (c->n_genes > 0) {
        const bool small = c->n_cols < 65534;                        // 16-bit entry indices when no row can hold 65534 entries
        const int wpt = ((c->n_genes + 31) / 32 + kGeneRowThreads - 1) / kGeneRowThreads;
        const size_t smem = (size_t)kGeneCap * 8 + (size_t)wpt * kGeneRowThreads * 8 + (size_t)c->gene_slots * (small ? 2 : 4) + 16;
        if (smem > 200 * 1024) throw ArgErr{SCASA_ERR_UNSUPPORTED, "gene map: too many isoforms in multi-isoform genes for the gene kernel's shared memory"};
        const int cps = (int)std::max<size_t>(1, std::min<size_t>((size_t)c->gene_ctas_per_sm, (size_t)(220 * 1024) / (smem + 1024)));
        const int grid = (int)std::min<int64_t>(n_cells, (int64_t)w->n_sm * cps);
        const int n_stops = c->gene_n_multi + 1;
        if (small) {
            CK(cudaFuncSetAttribute(gene_rows_kernel<uint16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            gene_rows_kernel<uint16_t><<<grid, kGeneRowThreads, smem, s>>>(R, n_cells, c->d_col_gene.p, c->n_genes, c->gene_slots, c->d_gene_stops.p, n_stops, wpt);
        } else {
            CK(cudaFuncSetAttribute(gene_rows_kernel<uint32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            gene_rows_kernel<uint32_t><<<grid, kGeneRowThreads, smem, s>>>(R, n_cells, c->d_col_gene.p, c->n_genes, c->gene_slots, c->d_gene_stops.p, n_stops, wpt);
        }
        CK(cudaGetLastError());
        launches++;
        c->ran_genes = true;
    }
    CK(cudaEventRecord(w->ev[5], s));

    int flag = 0;
    CK(cudaMemcpyAsync(rs.st, w->d_stats.p, sizeof rs.st, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(&flag, w->d_flag.p, sizeof flag, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    if (flag) throw ArgErr{SCASA_ERR_UNSUPPORTED, "a cell exceeded the eqclass capacity of the pack kernel"};
    CK(cudaEventElapsedTime(&rs.pack_ms, w->ev[0], w->ev[1]));
    CK(cudaEventElapsedTime(&rs.tasks_ms, w->ev[1], w->ev[2]));
    CK(cudaEventElapsedTime(&rs.em_ms, w->ev[2], w->ev[3]));
    CK(cudaEventElapsedTime(&rs.finish_ms, w->ev[3], w->ev[4]));
    CK(cudaEventElapsedTime(&rs.gene_ms, w->ev[4], w->ev[5]));
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
    t->opts = c->opts; t->host_threads = c->host_threads; t->fetch_dense_copy = c->fetch_dense_copy; t->d2h_compact = c->d2h_compact;
    for (int k = 0; k < kClasses; k++) {
        t->cls_warps[k] = c->cls_warps[k]; t->cls_pairs[k] = c->cls_pairs[k]; t->cls_bytes[k] = c->cls_bytes[k];
        t->cls_cta_threads[k] = c->cls_cta_threads[k]; t->cls_ctas_per_sm[k] = c->cls_ctas_per_sm[k];
    }
    t->pair_max = c->pair_max; t->pair_ctas_per_sm = c->pair_ctas_per_sm; t->poison = c->poison; t->pack_threads = c->pack_threads;
    if (c->n_genes > 0 && (t->n_genes != c->n_genes || t->gene_of_col_host != c->gene_of_col_host)) {
        int rc = scasa_set_gene_map(t, c->gene_of_col_host.data(), c->n_genes);
        if (rc) return fail(c, rc, std::string("twin context: ") + scasa_last_error(t));
    } else if (c->n_genes <= 0) t->n_genes = 0;
    return SCASA_OK;
}

static int run_staged(scasa_ctx *c)
{
    if (!c->staged) return fail(c, SCASA_ERR_STATE, "nothing staged: call scasa_stage_counts / scasa_stage_eqcounts first");
    return guarded(c, [&] {
        const int64_t n_tasks = (int64_t)c->n_chunks * c->n_em;
        c->d_iters.ensure((size_t)n_tasks * 2 + 2);
        c->d_colsum.ensure((size_t)c->n_cols);
        c->ran = false; c->have_rs_host = false; c->have_glen_host = false; c->ran_genes = false;
        CK(cudaEventRecord(c->ev[6], c->stream));
        RangeStats r;
        run_all(c, r);
        cudaStream_t s = c->stream;
        CK(cudaEventRecord(c->ev[7], s));
        CK(cudaStreamSynchronize(s));
        scasa_timing_t &tm = c->timing;
        memset(&tm, 0, sizeof tm);
        CK(cudaEventElapsedTime(&tm.total_ms, c->ev[6], c->ev[7]));
        tm.pack_ms = r.pack_ms; tm.tasks_ms = r.tasks_ms; tm.em_ms = r.em_ms; tm.finish_ms = r.finish_ms + r.gene_ms; tm.gene_ms = r.gene_ms;
        for (int k = 0; k < kClasses; k++) { tm.em_class_ms[k] = r.cls_ms[k]; tm.em_class_tasks[k] = r.cnts[k]; }
        tm.em_pair_ms = r.cls_ms[kPairList]; tm.em_pair_tasks = r.cnts[kPairCounter];
        tm.n_entries = r.blob_bytes;
        tm.launches = r.launches;
        tm.max_slot_bytes = (int32_t)r.cnts[5];
        tm.run_parts = 1;
        tm.inner_iters = (int64_t)r.st[0]; tm.outer_iters = (int64_t)r.st[1];
        tm.n_tasks = (int64_t)r.st[4]; tm.n_active = (int64_t)r.st[5];
        // sum_tasks [in*8*(q+2p)*n_c + out*8*(q+p)*n_c] + n_cells*8*(sum q + sum p)
        tm.em_alg_bytes = 8.0 * ((double)r.st[2] + (double)r.st[3]);
        tm.alg_bytes = tm.em_alg_bytes + (double)c->n_cells * 8.0 * (double)(c->n_rows + c->n_cols);
        tm.result_nnz = c->rs_nnz;
        c->ran = true;
    });
}

int scasa_run(scasa_ctx *c)
{
    if (!c) return fail(c, SCASA_ERR_ARG, "null context");
    return run_staged(c);
}

int scasa_set_run_parts(scasa_ctx *c, int32_t n_parts)
{
    // kept for ABI stability: cutting one run into ranges measured slower on B200 and was removed
    if (!c || n_parts < 0) return fail(c, SCASA_ERR_ARG, "bad argument");
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
    return guarded(c, [&] { fetch_dense(c, false, out); });
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
    if (c->n_genes <= 0 || !c->ran_genes) return fail(c, SCASA_ERR_STATE, "no gene map was set before the run");
    return guarded(c, [&] { fetch_dense(c, true, out); });
}

int scasa_result_nnz(scasa_ctx *c, int64_t *iso_nnz, int64_t *gene_nnz)
{
    if (!c) return fail(c, SCASA_ERR_ARG, "null context");
    if (!c->ran) return fail(c, SCASA_ERR_STATE, "scasa_run has not completed");
    return guarded(c, [&] {
        if (iso_nnz) *iso_nnz = c->rs_nnz;
        if (gene_nnz) {
            *gene_nnz = 0;
            if (c->ran_genes) {
                need_rs_host(c, true);
                int64_t t = 0;
                for (int32_t v : c->g_len_host) t += v;
                *gene_nnz = t;
            }
        }
    });
}

int scasa_fetch_iso_csr(scasa_ctx *c, int64_t *rowptr, int32_t *cols, double *vals)
{
    if (!c || !rowptr) return fail(c, SCASA_ERR_ARG, "null argument");
    if (!c->ran) return fail(c, SCASA_ERR_STATE, "scasa_run has not completed");
    return guarded(c, [&] {
        need_rs_host(c, false);
        std::copy(c->rs_ptr_host.begin(), c->rs_ptr_host.end(), rowptr);
        if ((cols || vals) && c->rs_nnz) fetch_sparse(c, false, cols, vals, rowptr, nullptr, false);
    });
}

int scasa_fetch_gene_csr(scasa_ctx *c, int64_t *rowptr, int32_t *genes, double *vals)
{
    if (!c || !rowptr) return fail(c, SCASA_ERR_ARG, "null argument");
    if (!c->ran) return fail(c, SCASA_ERR_STATE, "scasa_run has not completed");
    if (c->n_genes <= 0 || !c->ran_genes) return fail(c, SCASA_ERR_STATE, "no gene map was set before the run");
    return guarded(c, [&] {
        need_rs_host(c, true);
        rowptr[0] = 0;
        for (int64_t r = 0; r < c->n_cells; r++) rowptr[r + 1] = rowptr[r] + c->g_len_host[(size_t)r];
        if (!genes && !vals) return;
        fetch_sparse(c, true, genes, vals, rowptr, nullptr, false);      // the device rows have the capacity of the isoform rows: closed up on the way
    });
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

// ---- one call for a whole sample -------------------------------------------------------------------------
// What a quant call can deliver per slice of chunks.  Dense matrices are expanded into the caller's buffers;
// the sparse sinks receive the slice's CSR rows as they are on the device (for callers that never want the
// dense form: the text writers).
// big host arrays without the zero fill of std::vector (the first touch happens in the threads that write them).
// Arrays of 256 MB and more are aligned to 2 MB and advised for transparent huge pages: a fresh 4.8 GB result is
// 1.2 M first-touch faults in 4 KB pages (1.1 s per 100k cells measured, the threads queue on the address-space lock)
// and 2,400 in 2 MB pages (0.36 s).  Smaller arrays stay with malloc: on a host with fragmented memory a huge-page fault
// can stall in compaction (measured once: 0.07 -> 0.5 s for the 100 MB arrays of a 5,000-cell sample).  SCASA_THP=0: never.
inline void *big_alloc(size_t bytes)
{
    bytes = std::max<size_t>(bytes, 1);
    static const bool thp = [] { const char *e = getenv("SCASA_THP"); return !e || atoi(e) != 0; }();
    if (!thp || bytes < ((size_t)256 << 20)) return malloc(bytes);
    void *p = nullptr;
    if (posix_memalign(&p, (size_t)2 << 20, bytes)) return nullptr;
#ifdef MADV_HUGEPAGE
    madvise(p, bytes, MADV_HUGEPAGE);
#endif
    return p;
}
extern "C++" {
template <class T> struct RawBuf {
    T *p = nullptr;
    size_t n = 0;
    RawBuf() = default;
    RawBuf(const RawBuf &) = delete;
    RawBuf &operator=(const RawBuf &) = delete;
    RawBuf(RawBuf &&o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
    void alloc(size_t m) { release(); n = m; p = (T *)big_alloc(m * sizeof(T)); if (!p) throw std::bad_alloc(); }
    void release() { free(p); p = nullptr; n = 0; }
    ~RawBuf() { free(p); }
};
}
struct HeldRows {                         // a slice's result rows, packed, parked on the device until the host arrays can be sized
    scasa_ctx *ctx = nullptr;             // the context (device, stream, staging buffers) that owns and later drains them
    DBuf<int32_t> col;
    DBuf<double> val;
    std::vector<int64_t> ptr;             // [rows + 1] row starts inside col / val
    int64_t c0 = 0, c1 = 0;               // cells of the sample
};
struct SparseSink {                       // CSR by cell of the whole sample, assembled from the slices
    std::vector<int64_t> rowptr;          // [n_cells+1]: entries per row at [i+1] while the slices run, offsets afterwards
    std::vector<HeldRows> held;           // per slice
    bool drop0 = false;                   // leave the stored zeros behind
    ~SparseSink() { for (auto &h : held) { if (h.ctx && (h.col.p || h.val.p)) cudaSetDevice(h.ctx->device); h.col.release(); h.val.release(); } }
};
struct QuantOut {
    double *iso = nullptr, *gene = nullptr, *colsum = nullptr;
    int32_t *iters = nullptr;
    SparseSink *iso_csr = nullptr, *gene_csr = nullptr;
};

// Chunks [H0, H1) of the caller's sample on context c (and its twin): cut into n_slices slices that alternate
// between the two contexts, each driven by its own host thread, so that while one slice's results cross PCIe
// and are expanded on the host the next slice is staged and computed.  Chunks are independent, so slicing
// changes no result.  colparts ([n_chunks x n_emcols], rows H0..H1 written) and ptsum ([n_cols], this range's
// exact pass-through sums added under the mutex) are what the caller builds the column sums from.
static int quant_range(scasa_ctx *c, int H0, int H1, const int64_t *chunk_off, const int64_t *cell_ptr, const int32_t *eq_id,
                       const int32_t *eq_count, int64_t n_eq, const int32_t *eq_row, const QuantOut &out, double *colparts,
                       double *ptsum, std::mutex *ptsum_mutex, int32_t n_slices, int sink_slice0)
{
    const int n_chunks = H1 - H0;
    if (n_chunks <= 0) return SCASA_OK;
    if (n_slices <= 0) n_slices = std::max(2, std::min(8, n_chunks / 128));   // measured: 4-8 slices of >= 100 chunks are best
    n_slices = std::max(1, std::min(n_slices, n_chunks));
    if (n_slices > 1) { int rc = sync_twin(c); if (rc) return rc; }
    scasa_ctx *ctxs[2] = {c, n_slices > 1 ? c->twin : c};
    int rcs[2] = {SCASA_OK, SCASA_OK};
    auto drive = [&](int which) {                    // runs on its own thread for which == 1: nothing may escape
        scasa_ctx *x = ctxs[which];
        try {
        std::vector<int64_t> ptr, coff;
        std::vector<double> pt;
        for (int sl = which; sl < n_slices; sl += 2) {
            const int h0 = H0 + (int)((int64_t)n_chunks * sl / n_slices), h1 = H0 + (int)((int64_t)n_chunks * (sl + 1) / n_slices);
            if (h1 <= h0) continue;
            const int64_t c0 = chunk_off[h0], c1 = chunk_off[h1];
            coff.resize((size_t)(h1 - h0) + 1);
            for (int h = h0; h <= h1; h++) coff[(size_t)(h - h0)] = chunk_off[h] - c0;
            ptr.resize((size_t)(c1 - c0) + 1);
            const int64_t e0 = cell_ptr[c0];
            for (int64_t i = c0; i <= c1; i++) ptr[(size_t)(i - c0)] = cell_ptr[i] - e0;
            int rc = scasa_stage_eqcounts(x, c1 - c0, h1 - h0, coff.data(), ptr.data(), eq_id + e0, eq_count + e0, n_eq, eq_row);
            if (!rc) rc = run_staged(x);
            if (!rc && out.iso) rc = scasa_fetch_iso(x, out.iso + c0 * c->n_cols);
            if (!rc && out.gene) rc = scasa_fetch_gene(x, out.gene + c0 * (int64_t)c->n_genes);
            if (!rc && out.iters) rc = scasa_fetch_iters(x, out.iters + (int64_t)h0 * c->n_em * 2);
            for (int which_sink = 0; which_sink < 2 && !rc; which_sink++) {
                SparseSink *kp = which_sink ? out.gene_csr : out.iso_csr;
                if (!kp) continue;
                SparseSink &k = *kp;
                const bool genes = which_sink == 1;
                HeldRows &h = k.held[(size_t)(sink_slice0 + sl)];
                rc = guarded(x, [&] {
                    // pack the rows on the device (without their zeros, or just closed up) and keep them there: the
                    // host arrays are sized when every slice of every GPU has run (drain_sink)
                    row_source(x, genes, true, !k.drop0);
                    CK(cudaStreamSynchronize(x->stream));
                    h.ctx = x; h.c0 = c0; h.c1 = c1;
                    std::swap(h.col, x->d_cmp_col);
                    std::swap(h.val, x->d_cmp_val);
                    h.ptr.swap(x->cmp_ptr_host);
                    for (int64_t i = 0; i < c1 - c0; i++) k.rowptr[(size_t)(c0 + i) + 1] = h.ptr[(size_t)i + 1] - h.ptr[(size_t)i];
                });
            }
            if (!rc && colparts && c->n_emcols > 0)
                rc = fetch(x, colparts + (size_t)h0 * c->n_emcols, x->d_colpart.p, (size_t)(h1 - h0) * c->n_emcols * sizeof(double));
            if (!rc && ptsum) {
                pt.resize((size_t)c->n_cols);
                rc = scasa_fetch_colsum(x, pt.data());
                if (!rc) {
                    std::lock_guard<std::mutex> lk(*ptsum_mutex);
                    for (int64_t j = 0; j < c->n_cols; j++) ptsum[j] += pt[(size_t)j];   // only the pass-through columns are read: integer sums, exact in any order
                }
            }
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
    return SCASA_OK;
}

// The column sums of the whole sample (step2:167) from what the ranges delivered: pass-through columns are exact
// integer sums; EM columns are added chunk by chunk in chunk order — the same sequence scasa_run's own reduction
// follows, so the sums do not depend on the slicing or on the number of GPUs.
static void finish_colsum(const scasa_ctx *c, int n_chunks, const double *colparts, const double *ptsum, double *colsum_out)
{
    for (int64_t j = 0; j < c->n_cols; j++) colsum_out[j] = ptsum[j];
    std::vector<double> acc((size_t)std::max(1, c->n_emcols), 0.0);
    for (int h = 0; h < n_chunks; h++) {
        const double *row = colparts + (size_t)h * c->n_emcols;
        for (int j = 0; j < c->n_emcols; j++) acc[(size_t)j] += row[j];
    }
    for (int e = 0, j = 0; e < c->n_em; e++) {
        const int k = c->em_blocks[(size_t)e];
        for (int col = c->p_off[(size_t)k]; col < c->p_off[(size_t)k + 1]; col++) colsum_out[col] = acc[(size_t)j++];
    }
}

static int quant_multi_impl(scasa_ctx *const *ctxs, int32_t n_ctx, int64_t n_cells, int32_t n_chunks, const int64_t *chunk_off,
                            const int64_t *cell_ptr, const int32_t *eq_id, const int32_t *eq_count, int64_t n_eq, const int32_t *eq_row,
                            const QuantOut &out, int32_t n_slices)
{
    scasa_ctx *c = ctxs ? ctxs[0] : nullptr;
    if (!c || n_ctx < 1) return fail(c, SCASA_ERR_ARG, "null context");
    if (n_cells <= 0 || n_chunks <= 0 || !chunk_off || !cell_ptr || !eq_id || !eq_count || !eq_row)
        return fail(c, SCASA_ERR_ARG, "bad argument");
    if (chunk_off[0] != 0 || chunk_off[n_chunks] != n_cells) return fail(c, SCASA_ERR_ARG, "chunk_off must cover [0, n_cells]");
    if ((out.gene || out.gene_csr) && c->n_genes <= 0) return fail(c, SCASA_ERR_STATE, "no gene map set");
    for (int r = 1; r < n_ctx; r++) {
        if (!ctxs[r] || ctxs[r]->n_cols != c->n_cols || ctxs[r]->n_rows != c->n_rows || ctxs[r]->n_em != c->n_em)
            return fail(c, SCASA_ERR_ARG, "the contexts of a multi-GPU call must hold the same X-matrix");
        ctxs[r]->opts = c->opts;
        if (c->n_genes > 0 && (ctxs[r]->n_genes != c->n_genes || ctxs[r]->gene_of_col_host != c->gene_of_col_host)) {
            int rc = scasa_set_gene_map(ctxs[r], c->gene_of_col_host.data(), c->n_genes);
            if (rc) return fail(c, rc, std::string("context ") + std::to_string(r) + ": " + scasa_last_error(ctxs[r]));
        }
    }
    n_ctx = std::min<int32_t>(n_ctx, n_chunks);
    // contiguous chunk ranges per GPU, balanced by the number of (cell, eqclass) entries
    std::vector<int> cut((size_t)n_ctx + 1, 0);
    {
        const double total = (double)cell_ptr[n_cells];
        int h = 0;
        for (int r = 1; r < n_ctx; r++) {
            const double want = total * r / n_ctx;
            while (h < n_chunks - (n_ctx - r) && (double)cell_ptr[chunk_off[h + 1]] <= want) h++;
            h = std::max(h, cut[(size_t)r - 1] + 1);
            cut[(size_t)r] = h;
        }
        cut[(size_t)n_ctx] = n_chunks;
    }
    std::vector<double> colparts, ptsum;
    std::mutex ptm;
    if (out.colsum) { colparts.assign((size_t)n_chunks * std::max(1, c->n_emcols), 0.0); ptsum.assign((size_t)c->n_cols, 0.0); }
    // slices per rank are fixed up front so that the sparse sinks can be sized before the threads start
    std::vector<int> nsl((size_t)n_ctx), sl0((size_t)n_ctx + 1, 0);
    for (int r = 0; r < n_ctx; r++) {
        const int nch = cut[(size_t)r + 1] - cut[(size_t)r];
        int k = n_slices > 0 ? n_slices : std::max(2, std::min(8, nch / 128));
        nsl[(size_t)r] = std::max(1, std::min(k, nch));
        sl0[(size_t)r + 1] = sl0[(size_t)r] + nsl[(size_t)r];
    }
    for (SparseSink *k : {out.iso_csr, out.gene_csr})
        if (k) {
            k->rowptr.assign((size_t)n_cells + 1, 0);
            k->held = std::vector<HeldRows>((size_t)sl0[(size_t)n_ctx]);
        }
    std::vector<int> rcs((size_t)n_ctx, SCASA_OK);
    auto rank = [&](int r) {
        rcs[(size_t)r] = quant_range(ctxs[r], cut[(size_t)r], cut[(size_t)r + 1], chunk_off, cell_ptr, eq_id, eq_count, n_eq, eq_row, out,
                                     out.colsum ? colparts.data() : nullptr, out.colsum ? ptsum.data() : nullptr, &ptm,
                                     nsl[(size_t)r], sl0[(size_t)r]);
    };
    {
        std::vector<std::thread> th;
        try {
            for (int r = 1; r < n_ctx; r++) th.emplace_back(rank, r);
        } catch (const std::exception &e) {
            for (auto &t : th) t.join();
            return fail(c, SCASA_ERR_NOMEM, std::string("cannot start a host thread per GPU: ") + e.what());
        }
        rank(0);
        for (auto &t : th) t.join();
    }
    for (int r = 0; r < n_ctx; r++)
        if (rcs[(size_t)r]) return r == 0 ? rcs[0] : fail(c, rcs[(size_t)r], std::string("GPU ") + std::to_string(ctxs[r]->device) + ": " + scasa_last_error(ctxs[r]));
    if (out.colsum) finish_colsum(c, n_chunks, colparts.data(), ptsum.data(), out.colsum);
    for (SparseSink *k : {out.iso_csr, out.gene_csr})
        if (k) for (int64_t i = 0; i < n_cells; i++) k->rowptr[(size_t)i + 1] += k->rowptr[(size_t)i];
    return SCASA_OK;
}

int scasa_quant(scasa_ctx *c, int64_t n_cells, int32_t n_chunks, const int64_t *chunk_off, const int64_t *cell_ptr,
                const int32_t *eq_id, const int32_t *eq_count, int64_t n_eq, const int32_t *eq_row, double *iso_out,
                double *gene_out, double *colsum_out, int32_t *iters_out, int32_t n_slices)
{
    if (!c) return fail(c, SCASA_ERR_ARG, "null context");
    if (!iso_out) return fail(c, SCASA_ERR_ARG, "bad argument");
    QuantOut out;
    out.iso = iso_out; out.gene = gene_out; out.colsum = colsum_out; out.iters = iters_out;
    return quant_multi_impl(&c, 1, n_cells, n_chunks, chunk_off, cell_ptr, eq_id, eq_count, n_eq, eq_row, out, n_slices);
}

// The same over several GPUs of one box: ctxs[r] created on device r with the same X-matrix; contiguous chunk
// ranges per GPU, one host thread per GPU (each of which slices its range as scasa_quant does), every GPU
// writes its own rows of the caller's matrices.  No collective: chunks are independent (SURVEY.md §8 e).
// Results are bitwise those of a single GPU, column sums included.
int scasa_quant_multi(scasa_ctx *const *ctxs, int32_t n_ctx, int64_t n_cells, int32_t n_chunks, const int64_t *chunk_off,
                      const int64_t *cell_ptr, const int32_t *eq_id, const int32_t *eq_count, int64_t n_eq, const int32_t *eq_row,
                      double *iso_out, double *gene_out, double *colsum_out, int32_t *iters_out, int32_t n_slices)
{
    if (!ctxs || n_ctx < 1 || !ctxs[0]) return fail(nullptr, SCASA_ERR_ARG, "null context list");
    if (!iso_out) return fail(ctxs[0], SCASA_ERR_ARG, "bad argument");
    QuantOut out;
    out.iso = iso_out; out.gene = gene_out; out.colsum = colsum_out; out.iters = iters_out;
    return quant_multi_impl(ctxs, n_ctx, n_cells, n_chunks, chunk_off, cell_ptr, eq_id, eq_count, n_eq, eq_row, out, n_slices);
}

// ---- eqclass counts in, the two text files out -----------------------------------------------------------
// step2:88-170 + step3:21-52 without the dense matrices: the sparse rows of every slice are collected on the
// host, transposed to transcript-major (step2:166) and streamed through the sparse writer.
namespace {
struct Flat { std::vector<int64_t> rowptr; RawBuf<int32_t> cols; RawBuf<double> vals; };
// The parked slices of a sink -> the caller's contiguous arrays (k.rowptr holds the global offsets by now).  Every
// context drains its own slices on its own thread (a GPU's two contexts, and the GPUs, side by side), each through
// its pinned staging buffers and the host thread pool, straight to the rows' final place; the device memory of a
// sample's parked rows is released when all of them are on the host.
int drain_sink(SparseSink &k, int32_t *cols, double *vals)
{
    std::vector<scasa_ctx *> owners;
    for (auto &h : k.held)
        if (h.ctx && std::find(owners.begin(), owners.end(), h.ctx) == owners.end()) owners.push_back(h.ctx);
    std::vector<int> rcs(owners.size(), SCASA_OK);
    auto drain = [&](size_t oi) {
        scasa_ctx *x = owners[oi];
        for (auto &h : k.held) {
            if (h.ctx != x || rcs[oi]) continue;
            rcs[oi] = guarded(x, [&] {
                const int64_t base = k.rowptr[(size_t)h.c0];
                const RowSrc src{h.ptr.data(), nullptr, h.col.p, h.val.p};
                if (h.c1 > h.c0 && h.ptr.back() > 0) fetch_rows(x, src, h.c1 - h.c0, cols + base, vals + base, h.ptr.data(), nullptr, false);
            });
        }
    };
    std::vector<std::thread> th;
    for (size_t oi = 1; oi < owners.size(); oi++) {
        bool started = false;
        try { th.emplace_back(drain, oi); started = true; } catch (const std::exception &) {}
        if (!started) drain(oi);
    }
    if (!owners.empty()) drain(0);
    for (auto &t : th) t.join();
    // cudaFree waits for the whole device: the parked rows go when nobody is copying any more
    for (auto &h : k.held) {
        if (h.ctx && (h.col.p || h.val.p)) cudaSetDevice(h.ctx->device);
        h.col.release(); h.val.release();
        std::vector<int64_t>().swap(h.ptr);
    }
    for (size_t oi = 0; oi < owners.size(); oi++)
        if (rcs[oi]) return oi == 0 ? rcs[oi] : fail(owners[0], rcs[oi], std::string("GPU ") + std::to_string(owners[oi]->device) + ": " + scasa_last_error(owners[oi]));
    return SCASA_OK;
}
int flatten(SparseSink &k, Flat &f)
{
    const int64_t nnz = k.rowptr.empty() ? 0 : k.rowptr.back();
    f.cols.alloc((size_t)nnz); f.vals.alloc((size_t)nnz);
    const int rc = drain_sink(k, f.cols.p, f.vals.p);
    f.rowptr.swap(k.rowptr);
    return rc;
}
int sink_to_csr(SparseSink &k, int64_t n_cells, scasa_csr_t *out)
{
    out->n_rows = n_cells; out->nnz = k.rowptr.empty() ? 0 : k.rowptr.back();
    out->ptr = (int64_t *)malloc(((size_t)n_cells + 1) * sizeof(int64_t));
    out->col = (int32_t *)big_alloc((size_t)out->nnz * sizeof(int32_t));
    out->val = (double *)big_alloc((size_t)out->nnz * sizeof(double));
    if (!out->ptr || !out->col || !out->val) throw std::bad_alloc();
    std::copy(k.rowptr.begin(), k.rowptr.end(), out->ptr);
    return drain_sink(k, out->col, out->val);
}
double wall_now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
}  // namespace

void scasa_csr_free(scasa_csr_t *m)
{
    if (!m) return;
    free(m->ptr); free(m->col); free(m->val);
    m->ptr = nullptr; m->col = nullptr; m->val = nullptr; m->nnz = 0; m->n_rows = 0;
}

int scasa_quant_csr(scasa_ctx *const *ctxs, int32_t n_ctx, int64_t n_cells, int32_t n_chunks, const int64_t *chunk_off,
                    const int64_t *cell_ptr, const int32_t *eq_id, const int32_t *eq_count, int64_t n_eq, const int32_t *eq_row,
                    int32_t drop_zeros, scasa_csr_t *iso, scasa_csr_t *gene, double *colsum_out, int32_t *iters_out)
{
    if (!ctxs || n_ctx < 1 || !ctxs[0]) return fail(nullptr, SCASA_ERR_ARG, "null context list");
    scasa_ctx *c = ctxs[0];
    if (!iso) return fail(c, SCASA_ERR_ARG, "bad argument");
    if (gene && c->n_genes <= 0) return fail(c, SCASA_ERR_STATE, "gene matrix asked for, but no gene map");
    memset(iso, 0, sizeof *iso);
    if (gene) memset(gene, 0, sizeof *gene);
    try {
        SparseSink iso_sink, gene_sink;
        QuantOut out;
        iso_sink.drop0 = gene_sink.drop0 = drop_zeros != 0;
        out.iso_csr = &iso_sink; out.gene_csr = gene ? &gene_sink : nullptr; out.colsum = colsum_out; out.iters = iters_out;
        const double t0 = wall_now();
        int rc = quant_multi_impl(ctxs, n_ctx, n_cells, n_chunks, chunk_off, cell_ptr, eq_id, eq_count, n_eq, eq_row, out, 0);
        if (rc) return rc;
        const double t1 = wall_now();
        rc = sink_to_csr(iso_sink, n_cells, iso);
        if (!rc && gene) rc = sink_to_csr(gene_sink, n_cells, gene);
        if (rc) { scasa_csr_free(iso); scasa_csr_free(gene); return rc; }
        if (getenv("SCASA_FETCH_TRACE"))
            fprintf(stderr, "[scasa quant_csr] %lld cells: slices %.3f s, transfer %.3f s, %lld + %lld entries\n", (long long)n_cells, t1 - t0,
                    wall_now() - t1, (long long)iso->nnz, (long long)(gene ? gene->nnz : 0));
        return SCASA_OK;
    } catch (const std::bad_alloc &) {
        scasa_csr_free(iso); scasa_csr_free(gene);
        return fail(c, SCASA_ERR_NOMEM, "host allocation failed");
    } catch (const std::exception &e) {
        scasa_csr_free(iso); scasa_csr_free(gene);
        return fail(c, SCASA_ERR_ARG, e.what());
    }
}

int scasa_quant_files(scasa_ctx *const *ctxs, int32_t n_ctx, int64_t n_cells, int32_t n_chunks, const int64_t *chunk_off,
                      const int64_t *cell_ptr, const int32_t *eq_id, const int32_t *eq_count, int64_t n_eq, const int32_t *eq_row,
                      const char *iso_path, const char *gene_path, const char *tx_names, const char *gene_names,
                      const int32_t *gene_rank, const char *cell_names, int32_t n_threads, double *colsum_out,
                      scasa_files_timing_t *timing)
{
    if (!ctxs || n_ctx < 1 || !ctxs[0]) return fail(nullptr, SCASA_ERR_ARG, "null context list");
    scasa_ctx *c = ctxs[0];
    if (!iso_path || !tx_names || !cell_names) return fail(c, SCASA_ERR_ARG, "bad argument");
    if (gene_path && (c->n_genes <= 0 || !gene_names)) return fail(c, SCASA_ERR_STATE, "gene file asked for, but no gene map / gene names");
    scasa_files_timing_t tm;
    memset(&tm, 0, sizeof tm);
    try {
        SparseSink iso_sink, gene_sink;
        QuantOut out;
        std::vector<double> colsum((size_t)c->n_cols);
        iso_sink.drop0 = gene_sink.drop0 = true;                 // the writer prints 0 for an absent entry as for a stored zero
        out.iso_csr = &iso_sink; out.gene_csr = gene_path ? &gene_sink : nullptr; out.colsum = colsum.data();
        double t0 = wall_now();
        int rc = quant_multi_impl(ctxs, n_ctx, n_cells, n_chunks, chunk_off, cell_ptr, eq_id, eq_count, n_eq, eq_row, out, 0);
        if (rc) return rc;
        tm.quant_s = wall_now() - t0;
        if (colsum_out) std::copy(colsum.begin(), colsum.end(), colsum_out);

        // ---- isoforms: transpose, keep rows with rowSums > 0 (step2:167), write
        t0 = wall_now();
        Flat f;
        rc = flatten(iso_sink, f);
        if (rc) return rc;
        std::vector<int64_t> colptr((size_t)c->n_cols + 1);
        RawBuf<int32_t> cell_idx;
        RawBuf<double> tvals;
        cell_idx.alloc(f.cols.n); tvals.alloc(f.vals.n);
        rc = scasa_csr_transpose(n_cells, c->n_cols, f.rowptr.data(), f.cols.p, f.vals.p, colptr.data(), cell_idx.p, tvals.p, n_threads);
        if (rc) return fail(c, rc, scasa_last_error(nullptr));
        Flat().rowptr.swap(f.rowptr); f.cols.release(); f.vals.release();
        tm.transpose_s = wall_now() - t0;
        t0 = wall_now();
        std::vector<int32_t> src;
        std::string names;
        {
            const char *p = tx_names;
            for (int64_t j = 0; j < c->n_cols; j++) {
                const char *nl = strchr(p, '\n');
                if (!nl) return fail(c, SCASA_ERR_ARG, "tx_names must hold one '\\n'-terminated name per isoform column");
                if (colsum[(size_t)j] > 0) { src.push_back((int32_t)j); names.append(p, nl + 1); }
                p = nl + 1;
            }
        }
        rc = scasa_write_table_sparse(iso_path, n_cells, c->n_cols, colptr.data(), cell_idx.p, tvals.p, (int64_t)src.size(),
                                      src.data(), names.c_str(), cell_names, n_threads, &tm.iso_bytes);
        if (rc) return fail(c, rc, scasa_last_error(nullptr));
        tm.iso_rows = (int64_t)src.size();
        tm.write_iso_s = wall_now() - t0;

        // ---- genes (step3:32-46): rows = genes with a surviving isoform; single-isoform genes first, then the others,
        // each group in sort() order of the gene names (gene_rank: the caller's collation)
        if (gene_path) {
            t0 = wall_now();
            Flat g;
            rc = flatten(gene_sink, g);
            if (rc) return rc;
            const int64_t ng = c->n_genes;
            std::vector<int64_t> gptr((size_t)ng + 1);
            RawBuf<int32_t> gcell;
            RawBuf<double> gvals;
            gcell.alloc(g.cols.n); gvals.alloc(g.vals.n);
            rc = scasa_csr_transpose(n_cells, ng, g.rowptr.data(), g.cols.p, g.vals.p, gptr.data(), gcell.p, gvals.p, n_threads);
            if (rc) return fail(c, rc, scasa_last_error(nullptr));
            tm.transpose_s += wall_now() - t0;
            t0 = wall_now();
            std::vector<int32_t> kept_members((size_t)ng, 0);
            for (int64_t j = 0; j < c->n_cols; j++) {
                const int32_t gi = c->gene_of_col_host[(size_t)j];
                if (gi >= 0 && colsum[(size_t)j] > 0) kept_members[(size_t)gi]++;
            }
            std::vector<int32_t> order;
            for (int32_t gi = 0; gi < ng; gi++) if (kept_members[(size_t)gi] > 0) order.push_back(gi);
            std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) {
                const bool ma = kept_members[(size_t)a] > 1, mb = kept_members[(size_t)b] > 1;
                if (ma != mb) return !ma;                                        // singles (mat1) before multis (mat2), step3:43
                return gene_rank ? gene_rank[a] < gene_rank[b] : a < b;
            });
            std::vector<const char *> gn;
            {
                const char *p = gene_names;
                for (int64_t k = 0; k < ng; k++) {
                    gn.push_back(p);
                    const char *nl = strchr(p, '\n');
                    if (!nl) return fail(c, SCASA_ERR_ARG, "gene_names must hold one '\\n'-terminated name per gene");
                    p = nl + 1;
                }
                gn.push_back(p);
            }
            std::string gnames;
            for (int32_t gi : order) gnames.append(gn[(size_t)gi], gn[(size_t)gi + 1]);
            rc = scasa_write_table_sparse(gene_path, n_cells, ng, gptr.data(), gcell.p, gvals.p, (int64_t)order.size(), order.data(),
                                          gnames.c_str(), cell_names, n_threads, &tm.gene_bytes);
            if (rc) return fail(c, rc, scasa_last_error(nullptr));
            tm.gene_rows = (int64_t)order.size();
            tm.write_gene_s = wall_now() - t0;
        }
        if (timing) *timing = tm;
        return SCASA_OK;
    } catch (const std::bad_alloc &) {
        return fail(c, SCASA_ERR_NOMEM, "host allocation failed");
    } catch (const std::exception &e) {
        return fail(c, SCASA_ERR_ARG, e.what());
    }
}

int scasa_gene_sum(scasa_ctx *c, const int32_t *gene_of_col, int32_t n_genes, const double *iso, int64_t n_cells,
                   double *gene_out)
{
    if (!c || !gene_of_col || !iso || !gene_out || n_cells <= 0 || n_genes <= 0) return fail(c, SCASA_ERR_ARG, "bad argument");
    int rc = scasa_set_gene_map(c, gene_of_col, n_genes);
    if (rc) return rc;
    return guarded(c, [&] {
        cudaStream_t s = c->stream;
        c->d_iso_in.ensure((size_t)(n_cells * c->n_cols));
        c->d_gene_out.ensure((size_t)n_cells * n_genes);
        CK(cudaMemcpyAsync(c->d_iso_in.p, iso, n_cells * c->n_cols * sizeof(double), cudaMemcpyHostToDevice, s));
        launch_gene_gather(c, c->d_iso_in.p, n_cells, c->d_gene_out.p);
        CK(cudaMemcpyAsync(gene_out, c->d_gene_out.p, (size_t)n_cells * n_genes * sizeof(double), cudaMemcpyDeviceToHost, s));
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

int scasa_dm0_is_fixed_point(const scasa_xmat_t *crp, const scasa_xmat_t *dm0, const int32_t *crp_member, double clim,
                             int32_t n_threads, int32_t *is_fixed_point, int64_t *first_bad_block)
{
    if (!crp || !dm0 || !crp_member || !is_fixed_point || !(clim > 0)) return fail(nullptr, SCASA_ERR_ARG, "bad argument");
    return guarded(nullptr, [&] {
        check_xmat(crp, "crp");
        check_xmat(dm0, "dm0");
        if (crp->n_blocks != dm0->n_blocks) throw ArgErr{SCASA_ERR_ARG, "crp and dm0 differ in block count"};
        for (int k = 0; k <= crp->n_blocks; k++)
            if (crp->q_off[k] != dm0->q_off[k]) throw ArgErr{SCASA_ERR_ARG, "crp and dm0 differ in rows"};
        const int64_t bad = scasa::dm0_fixed_point_check(crp->n_blocks, crp->q_off, crp->p_off, crp->x_off, crp->x, dm0->p_off, dm0->x_off,
                                                         dm0->x, crp_member, clim, n_threads);
        *is_fixed_point = bad < 0;
        if (first_bad_block) *first_bad_block = bad;
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
