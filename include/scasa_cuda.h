/*
 * scasa_cuda.h — C ABI of libscasa_cuda: Scasa's 2QUANT quantification path on B200 (sm_100a).
 *
 * The reference has no FFI layer: the seam is three R closures sourced into the quant
 * driver (paths relative to /root/reference/scasa/SCRIPTS/):
 *
 *   crpcount_td(mytd)                         Scasa_Rsource.R:416-557  (+ hamming_correction :562-648)
 *       called per cell at scasa_quant_step2_v1.0.0.R:95
 *   estim_chunk(Y, DM0, maxiter.bt=1000)      Scasa_Rsource.R:944-1051 (estim.BETA :726-761, estim.X.em :898-922)
 *       called per chunk at scasa_quant_step2_v1.0.0.R:132
 *   scasa_genemat(scasa_iso, txmap)           scasa_quant_step3_v1.0.0.R:32-46, called at :48
 *
 * plus the chunk split (step2:64-68) and the all-zero-row filter (step2:167).  This header is
 * what an R shim (.C / .Call) binds instead.  Plain C: pointers and sizes
 * only, caller-owned buffers, an opaque context that owns everything on the device.
 *
 * Conventions
 *   - every function returns SCASA_OK (0) or a negative SCASA_ERR_*; scasa_last_error()
 *     gives the message.  Nothing aborts, nothing throws across the boundary.
 *   - "cell-major": matrix[cell * ncol + col] (R's t(matrix) stored column-major).
 *   - global row index: rows of all blocks concatenated, row r of block k = q_off[k] + r.
 *   - global column index: DM0 columns of all blocks concatenated = column order of
 *     estim_chunk's result (Rsource:1042).
 *   - there is no CPU fallback: without a CUDA device scasa_ctx_create fails.
 */
#ifndef SCASA_CUDA_H
#define SCASA_CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SCASA_OK 0
#define SCASA_ERR_ARG (-1)         /* bad argument / malformed input            */
#define SCASA_ERR_CUDA (-2)        /* CUDA runtime error (message has the text) */
#define SCASA_ERR_NOMEM (-3)       /* host or device allocation failed          */
#define SCASA_ERR_UNSUPPORTED (-4) /* input exceeds a documented limit          */
#define SCASA_ERR_STATE (-5)       /* call order (nothing staged / not run yet) */

typedef struct scasa_ctx scasa_ctx;

/* The "same convergence options": estim_chunk's defaults (Rsource:944) and the constants
 * hard-wired next to it.  scasa_opts_default() fills exactly these.  maxerr / lim govern the outer
 * (X) test of Rsource:1019-1020 only. */
typedef struct {
    int32_t maxiter_x;    /* 1000  maxiter.X                              Rsource:944  */
    int32_t maxiter_bt;   /* 1000  maxiter.bt as passed at step2:132                   */
    double maxerr;        /* 0.01                                         Rsource:944  */
    double lim;           /* 0.01                                         Rsource:944  */
    int32_t modify;       /* 1     digamma-modified EM                    Rsource:738  */
    double adjust_esp;    /* 2     pseudo count for poor 2-column blocks  Rsource:973  */
    int32_t hamming_dist; /* 1                                            Rsource:534  */
    int32_t fixed_iter;   /* 0     1 = never break: exactly maxiter_x x maxiter_bt iterations (parity mode) */
    int32_t fp32;         /* 0     1 = the EM iterates in fp32 (counts, logNorm and results stay fp64).  Against the fp64 path
                                   in fixed-iteration mode: estimates within 1e-4 relative except for a handful of entries
                                   (about 3 per million: ill-conditioned splits between transcripts with similar columns),
                                   which stay within 2e-3 while their block total per cell stays within 1e-4
                                   (tests/test_gpu_em.py).  north_star's 1e-4 is therefore NOT met on every entry; parity
                                   claims are made for fp64 only. */
    /* estim_chunk does not forward maxerr / lim to estim.BETA (the call at Rsource:1016 passes maxiter and
     * modify only), so the inner loop runs with estim.BETA's own defaults (Rsource:726-727) whatever
     * maxerr / lim above are.  These two are those defaults; a caller who wants the inner loop tightened
     * as well (BASELINE config C5) sets them explicitly. */
    double maxerr_bt;     /* 0.01                                         Rsource:726  */
    double lim_bt;        /* 0.01                                         Rsource:727  */
} scasa_opts_t;

/* A block list (R list of small dense matrices) flattened; blocks column-major like R. */
typedef struct {
    int32_t n_blocks;
    const int32_t *q_off; /* [n_blocks+1] first global row of each block      */
    const int32_t *p_off; /* [n_blocks+1] first global column of each block   */
    const int64_t *x_off; /* [n_blocks+1] first value of each block in x      */
    const double *x;      /* values, block k column-major q_k x p_k           */
} scasa_xmat_t;

/* CRP (un-merged X-matrix) plus what crpcount_td reads from its dimnames. */
typedef struct {
    scasa_xmat_t m;
    const int32_t *col_tx;    /* [sum p] transcript id of each column (caller's dictionary;
                                 the same ids are used in eqclass transcript sets)            */
    const uint8_t *row_pat;   /* [sum q*p] 0/1 row-name patterns, row-major inside a block     */
    const int32_t *name_rank; /* [n_blocks] rank of names(CRP)[k] under sort() (Rsource:484);
                                 R passes its own order so the session collation is honoured  */
} scasa_crp_t;

int scasa_opts_default(scasa_opts_t *opts);

/* Uploads DM0 (and CRP when the eqclass entry points are wanted) once; they stay resident.
 * device = CUDA ordinal (one context per GPU; one host thread or process per context). */
int scasa_ctx_create(const scasa_xmat_t *dm0, const scasa_crp_t *crp_or_null,
                     const scasa_opts_t *opts_or_null, int device, scasa_ctx **out);
void scasa_ctx_destroy(scasa_ctx *ctx);
const char *scasa_last_error(const scasa_ctx *ctx_or_null);
int scasa_set_opts(scasa_ctx *ctx, const scasa_opts_t *opts);

/* sizes a caller needs to allocate outputs */
int64_t scasa_n_rows(const scasa_ctx *ctx);       /* sum q                         */
int64_t scasa_n_cols(const scasa_ctx *ctx);       /* sum p of DM0                  */
int32_t scasa_n_em_blocks(const scasa_ctx *ctx);  /* blocks with nrow > 1          */
int scasa_em_blocks(const scasa_ctx *ctx, int32_t *block_ids /* [n_em_blocks] */);
int scasa_poor_blocks(const scasa_ctx *ctx, uint8_t *is_poor /* [n_blocks] */); /* poorCondX, Rsource:957-971 */

/* step2:64-68.  chunk_off[0..*n_chunks] (0-based cell offsets).  `core` is only read when
 * n_cells <= 100 (step2:66). */
int scasa_chunk_split(int64_t n_cells, int32_t core, int64_t *chunk_off, int64_t cap,
                      int32_t *n_chunks);

/* ---- estim_chunk over all chunks of a sample (host buffers; H2D, kernels, D2H inside) ----
 * Y as CSR by cell: entries of cell c are y_rowptr[c] .. y_rowptr[c+1]-1, y_rowidx = global
 * row, strictly ascending inside a cell.  iso_out: n_cells x n_cols cell-major.
 * iters_out (nullable): [n_chunks][n_em_blocks][2] = {outer, total inner} iterations.
 * gene_* optional (see scasa_set_gene_map). */
int scasa_estimate(scasa_ctx *ctx, int64_t n_cells, int32_t n_chunks, const int64_t *chunk_off,
                   const int64_t *y_rowptr, const int32_t *y_rowidx, const double *y_val,
                   double *iso_out, int32_t *iters_out);

/* ---- the body of step2 (:88-136) + step3 (:48) for a whole sample in one call ---------------
 * eqclass-level counts in (see scasa_stage_eqcounts), dense matrices out, all host buffers.
 * The chunks are processed in n_slices slices (0 = automatic) that alternate between two
 * contexts on the GPU, so the transfer and host-side expansion of one slice's results overlap
 * the computation of the next; device memory is sized by the slice.  gene_out (needs
 * scasa_set_gene_map), colsum_out and iters_out may be NULL. */
int scasa_quant(scasa_ctx *ctx, int64_t n_cells, int32_t n_chunks, const int64_t *chunk_off,
                const int64_t *cell_ptr, const int32_t *eq_id, const int32_t *eq_count, int64_t n_eq,
                const int32_t *eq_row, double *iso_out, double *gene_out, double *colsum_out,
                int32_t *iters_out, int32_t n_slices);

/* The same over several GPUs of one box — what `registerDoParallel(cores = core)` + the two `foreach ... %dopar%`
 * loops of scasa_quant_step2_v1.0.0.R:82-136 are to the reference: ctxs[r] was created on device r with the same
 * X-matrix (options and the gene map are taken from ctxs[0]).  Contiguous chunk ranges per GPU balanced by their
 * number of (cell, eqclass) entries, one host thread per GPU, every GPU writes its own rows of the caller's
 * matrices; no collective (chunks are independent).  Every output, column sums included, is bitwise what one
 * GPU gives. */
int scasa_quant_multi(scasa_ctx *const *ctxs, int32_t n_ctx, int64_t n_cells, int32_t n_chunks, const int64_t *chunk_off,
                      const int64_t *cell_ptr, const int32_t *eq_id, const int32_t *eq_count, int64_t n_eq,
                      const int32_t *eq_row, double *iso_out, double *gene_out, double *colsum_out,
                      int32_t *iters_out, int32_t n_slices);

/* ---- eqclass counts in, the two text files out (step2:88-170 + step3:21-52) --------------------------------
 * scasa_isoform_counts.txt (README.md:80 calls it scasa_isoform_expression.txt) and scasa_gene_expression.txt,
 * byte for byte what scasa_write_table writes from the dense matrices, but no dense matrix ever exists: the
 * sparse rows of every slice are collected on the host, transposed to transcript-major (step2:166) and streamed
 * through scasa_write_table_sparse.  Rows: isoforms with rowSums > 0 (step2:167) under tx_names (one line per
 * isoform column); genes with a surviving isoform, single-isoform genes first, then the others, each group in
 * the order gene_rank gives (the rank of the gene's name under the caller's sort(); NULL = id order), under
 * gene_names (one line per gene id of scasa_set_gene_map).  gene_path may be NULL. */
typedef struct {
    double quant_s, transpose_s, write_iso_s, write_gene_s;    /* wall seconds */
    int64_t iso_bytes, gene_bytes, iso_rows, gene_rows;
} scasa_files_timing_t;
int scasa_quant_files(scasa_ctx *const *ctxs, int32_t n_ctx, int64_t n_cells, int32_t n_chunks, const int64_t *chunk_off,
                      const int64_t *cell_ptr, const int32_t *eq_id, const int32_t *eq_count, int64_t n_eq,
                      const int32_t *eq_row, const char *iso_path, const char *gene_path, const char *tx_names,
                      const char *gene_names, const int32_t *gene_rank, const char *cell_names, int32_t n_threads,
                      double *colsum_out, scasa_files_timing_t *timing);

/* ---- eqclass counts in, sparse matrices out ------------------------------------------------------------------
 * The results as they are on the device: CSR by cell, columns ascending inside a row.  Read column-wise this IS
 * the compressed-column form of the reference's tx x cells matrix (step2:166): ptr/col/val are the p/i/x slots
 * of a Matrix::dgCMatrix with Dim = c(n_cols, n_cells), so an R caller gets `isoform_count` without a dense
 * 8 * n_cols * n_cells byte matrix ever existing (100k cells: 3.2 GB instead of 26.7 GB).  drop_zeros != 0
 * removes the stored zeros (columns of EM blocks whose estimate is exactly 0) on the way.  The arrays are
 * allocated by the library; release them with scasa_csr_free.  gene (needs scasa_set_gene_map), colsum_out and
 * iters_out may be NULL.  Values, column sums and iteration counts are those of scasa_quant_multi, bit for bit. */
typedef struct {
    int64_t n_rows, nnz;     /* cells, stored entries */
    int64_t *ptr;            /* [n_rows + 1] */
    int32_t *col;            /* [nnz] */
    double *val;             /* [nnz] */
} scasa_csr_t;
int scasa_quant_csr(scasa_ctx *const *ctxs, int32_t n_ctx, int64_t n_cells, int32_t n_chunks, const int64_t *chunk_off,
                    const int64_t *cell_ptr, const int32_t *eq_id, const int32_t *eq_count, int64_t n_eq,
                    const int32_t *eq_row, int32_t drop_zeros, scasa_csr_t *iso, scasa_csr_t *gene,
                    double *colsum_out, int32_t *iters_out);
void scasa_csr_free(scasa_csr_t *m);

/* ---- the same in resident stages (what scasa_estimate is made of) ----------------------- */
int scasa_stage_counts(scasa_ctx *ctx, int64_t n_cells, int32_t n_chunks, const int64_t *chunk_off,
                       const int64_t *y_rowptr, const int32_t *y_rowidx, const double *y_val);
/* eqclass-level counts (the crpcount_td input): per cell (eqclass id, UMI count) plus the
 * eqclass -> global row map from scasa_eqclass_map (row < 0: dropped).  Counts <= 0 are ignored
 * (alevin's count is the number of UMIs, >= 1). */
int scasa_stage_eqcounts(scasa_ctx *ctx, int64_t n_cells, int32_t n_chunks, const int64_t *chunk_off,
                         const int64_t *cell_ptr, const int32_t *eq_id, const int32_t *eq_count,
                         int64_t n_eq, const int32_t *eq_row);
int scasa_run(scasa_ctx *ctx);   /* pack -> EM -> (gene sums); everything on the device */
/* Retired (kept for ABI stability, does nothing): cutting one run into ranges of chunks on two stream
 * sets measured slower on B200. */
int scasa_set_run_parts(scasa_ctx *ctx, int32_t n_parts);
/* The device keeps both results sparse (see scasa_fetch_iso_csr); scasa_fetch_iso / scasa_fetch_gene send
 * the sparse rows over PCIe and host threads expand them into the caller's dense buffer, which ends up
 * bit-identical to a dense copy.  scasa_set_host_threads: threads of that expansion, 0 = all cores,
 * -1 = expand on the device and copy densely instead (testing aid). */
int scasa_set_host_threads(scasa_ctx *ctx, int32_t n_threads);
/* The host half of that, usable on its own (no GPU, no context): rows of (column, value) pairs,
 * columns strictly ascending inside a row, -> dense n_rows x n_cols matrix; every element written. */
int scasa_expand_pairs(double *out, int64_t n_rows, int64_t n_cols, const int64_t *row_start, const int32_t *row_nnz,
                       const int32_t *cols, const double *vals, int32_t n_threads);
int scasa_fetch_iso(scasa_ctx *ctx, double *iso_out);            /* n_cells x n_cols           */
/* The results as the device holds them: CSR by cell over the entries that CAN be non-zero (a cell's
 * pass-through counts, the columns of every EM block it has counts in, the two columns of every poor
 * block), columns ascending; stored zeros are possible.  rowptr [n_cells+1]; cols / vals sized by
 * scasa_result_nnz (either may be NULL).  The gene matrix likewise, genes ascending. */
int scasa_result_nnz(scasa_ctx *ctx, int64_t *iso_nnz, int64_t *gene_nnz);
int scasa_fetch_iso_csr(scasa_ctx *ctx, int64_t *rowptr, int32_t *cols, double *vals);
int scasa_fetch_gene_csr(scasa_ctx *ctx, int64_t *rowptr, int32_t *genes, double *vals);
int scasa_fetch_iters(scasa_ctx *ctx, int32_t *iters_out);       /* [n_chunks][n_em_blocks][2] */
int scasa_fetch_colsum(scasa_ctx *ctx, double *colsum /* [n_cols] */); /* rowSums of step2:167 */
int scasa_fetch_counts(scasa_ctx *ctx, int64_t *y_rowptr /* [n_cells+1] */, int32_t *y_rowidx,
                       double *y_val, int64_t cap, int64_t *nnz);  /* Y built by the pack stage */

/* ---- scasa_genemat ----------------------------------------------------------------------
 * gene_of_col[j] = output gene row of isoform column j, -1 = not summed (NA gene, or row
 * dropped at step2:167).  Single-isoform genes are copies, others sums in column order. */
int scasa_set_gene_map(scasa_ctx *ctx, const int32_t *gene_of_col, int32_t n_genes);
int scasa_fetch_gene(scasa_ctx *ctx, double *gene_out /* n_cells x n_genes cell-major */);
int scasa_gene_sum(scasa_ctx *ctx, const int32_t *gene_of_col, int32_t n_genes, const double *iso,
                   int64_t n_cells, double *gene_out);

/* ---- crpcount_td's decision, once per distinct eqclass ------------------------------------
 * eqclass e = transcript ids eq_tx[eq_off[e] .. eq_off[e+1]-1].  eq_row[e] = global row the
 * eqclass' counts are added to, or -1 (dropped). */
int scasa_eqclass_map(scasa_ctx *ctx, int64_t n_eq, const int64_t *eq_off, const int32_t *eq_tx,
                      int32_t *eq_row, int32_t n_threads);

/* The same without a context or a GPU (host only): builds the index from crp and maps.  For callers that
 * prepare eq_row ahead of the GPU call, and for tests of the host logic on machines without a device. */
int scasa_eqclass_map_host(const scasa_crp_t *crp, int32_t hamming_dist, int64_t n_eq, const int64_t *eq_off,
                           const int32_t *eq_tx, int32_t *eq_row, int32_t n_threads);

/* ---- can `DM0 <- lapply(CRP, ccrpfun, 30)` (scasa_quant_step2_v1.0.0.R:127; serial svd + kmeans(nstart = 100) over
 * 29,351 blocks in the R master, minutes) be skipped in favour of the shipped CCRP?  (host only)
 * crp_member[c] = global DM0 column CRP column c was merged into (the DM0 column whose name lists it), -1 = dropped.
 * *is_fixed_point = 1 when, for every block, DM0 is a point ccrpfun's loop (Scasa_Rsource.R:168-193) stops at:
 * the dropped columns are exactly the zero-sum ones, every DM0 column is the mean of its members (a k-means centre),
 * all condition numbers max(sval)/sval of DM0 are < clim, and wherever columns were merged the unmerged block was not
 * separated.  Which partition k-means picks among those is R's RNG and is not checked.  first_bad_block (nullable)
 * receives the first block that fails, or -1. */
int scasa_dm0_is_fixed_point(const scasa_xmat_t *crp, const scasa_xmat_t *dm0, const int32_t *crp_member, double clim,
                             int32_t n_threads, int32_t *is_fixed_point, int64_t *first_bad_block);

/* ---- alevin bfh.txt -> quant inputs (host only; replaces scasa_quant_step1_1_v1.0.0.R:16-68 and
 * the ordering of scasa_quant_step2_v1.0.0.R:59-60) --------------------------------------------
 * Cells are the barcodes that occur in some eqclass, in byte order of their sequences
 * (data.table::setorder); per cell the (eqclass, count) pairs in ascending eqclass id, count =
 * number of UMIs; per eqclass its transcript indices into the file's own transcript list (map the
 * names to the ids used in scasa_crp_t.col_tx before calling scasa_eqclass_map). */
typedef struct scasa_bfh scasa_bfh;
int scasa_bfh_open(const char *path, int32_t n_threads /* 0 = all cores */, scasa_bfh **out);
void scasa_bfh_close(scasa_bfh *f);
int scasa_bfh_sizes(const scasa_bfh *f, int64_t *n_tx, int64_t *n_cells, int64_t *n_eq, int64_t *n_pairs,
                    int64_t *n_eq_tx, int64_t *tx_name_bytes, int64_t *barcode_bytes);
/* names are written one per line ('\n'-terminated, sizes from scasa_bfh_sizes); any pointer may be NULL */
int scasa_bfh_read(const scasa_bfh *f, char *tx_names, char *barcodes, int64_t *eq_off /* [n_eq+1] */,
                   int32_t *eq_tx /* [n_eq_tx] */, int64_t *cell_ptr /* [n_cells+1] */, int32_t *eq_id /* [n_pairs], 0-based line */,
                   int32_t *eq_count /* [n_pairs] */);

/* ---- kallisto | bustools -> quant inputs (host only; replaces scasa_quant_step1_2_v1.0.0.R) -------
 * bus_txt = output.correct.sort.bus.final.txt (barcode, UMI, eqclass, count, barcode+UMI; tab separated),
 * matrix_ec and transcripts_txt as kallisto bus writes them.  Cells = barcodes with more than
 * libsize_thres records (the wrapper passes 200); repeated (barcode, UMI) pairs are resolved as the
 * script does; the count of a (cell, eqclass) is its number of kept molecules.  The handle is read
 * and closed with scasa_bfh_sizes / scasa_bfh_read / scasa_bfh_close (eq_id = line of matrix.ec). */
int scasa_bus_open(const char *bus_txt, const char *matrix_ec, const char *transcripts_txt, int64_t libsize_thres,
                   scasa_bfh **out);

/* ---- the two text outputs (host only) ----------------------------------------------------------
 * write.table(t(result)[keep, ], file, quote = F, row.names = T, sep = "\t") as at
 * scasa_quant_step2_v1.0.0.R:165-170 and scasa_quant_step3_v1.0.0.R:52: mat is the library's
 * cell-major matrix (n_cells x n_cols); the file has a header of cell names (one field fewer than
 * the data lines) and one line per column j with keep[j] != 0 (keep == NULL: all): its name, then
 * its value in every cell, formatted like R's writetable does (15 significant digits at most).
 * row_names / cell_names: one name per line, '\n'-terminated. */
int scasa_write_table(const char *path, const double *mat, int64_t n_cells, int64_t n_cols, const uint8_t *keep,
                      const char *row_names, const char *cell_names, int32_t n_threads);
/* The same bytes from a sparse, column-wise matrix (no dense matrix anywhere): source column j (a transcript, or a
 * gene) holds the entries [colptr[j], colptr[j+1]) = (cell, value), cells ascending; output line k is source
 * column out_src[k] under the k-th line of out_names; absent cells and stored zeros print as 0.
 * bytes_written (nullable) receives the file size. */
int scasa_write_table_sparse(const char *path, int64_t n_cells, int64_t n_src, const int64_t *colptr, const int32_t *cell_idx,
                             const double *vals, int64_t n_out, const int32_t *out_src, const char *out_names,
                             const char *cell_names, int32_t n_threads, int64_t *bytes_written);
/* CSR by row -> CSC with the entries of a column in row order (host threads, deterministic): turns the cell-major
 * results of scasa_fetch_iso_csr / scasa_fetch_gene_csr into the transcript-major form the text files need
 * (the t() of scasa_quant_step2_v1.0.0.R:166). */
int scasa_csr_transpose(int64_t n_rows, int64_t n_cols, const int64_t *rowptr, const int32_t *cols, const double *vals,
                        int64_t *colptr /* [n_cols+1] */, int32_t *rows_out, double *vals_out, int32_t n_threads);
/* the text R's write.table gives one number; returns its length (buf needs 32 bytes) */
int scasa_format_real15(double x, char *buf, int32_t cap);

/* ---- measurement ---------------------------------------------------------------------------
 * device time (CUDA events on the library's stream) of the last scasa_run, per phase */
typedef struct {
    float total_ms, pack_ms, tasks_ms, em_ms, finish_ms, gene_ms;
    float em_class_ms[4];   /* EM kernel time per work class (1, 2, 4, 8 warps per task); they overlap */
    int64_t em_class_tasks[4];
    int64_t n_tasks;        /* non-empty (chunk, EM block) tasks            */
    int64_t n_entries;      /* bytes of the task blobs built for the EM     */
    int64_t n_active;       /* cell records processed (cells with counts, + 1 pseudo cell per poor task) */
    int64_t inner_iters;    /* sum over tasks of estim.BETA iterations      */
    int64_t outer_iters;    /* sum over tasks of AEM iterations             */
    double alg_bytes;       /* SURVEY.md 8(d) streaming-model bytes of the run */
    double em_alg_bytes;    /* ... of the EM kernels alone (per-iteration terms) */
    int32_t launches;       /* kernels launched by the last run             */
    int32_t max_slot_bytes; /* largest shared-memory slot a task needed     */
    int32_t run_parts;      /* ranges of chunks the run was cut into; with more than one the phase times above are
                               sums over ranges that ran side by side (they exceed total_ms) */
    float em_pair_ms;       /* EM kernel time of the tiny-task kernel (two tasks per warp); overlaps the classes above */
    int64_t em_pair_tasks;  /* tasks it ran (not counted in em_class_tasks) */
    int64_t result_nnz;     /* entries of the sparse isoform result (structure incl. stored zeros) */
} scasa_timing_t;
int scasa_last_timing(scasa_ctx *ctx, scasa_timing_t *out);

/* bytes this context (and its scasa_quant twin) moved over PCIe since the last reset */
int scasa_transfer_bytes(scasa_ctx *ctx, int64_t *h2d, int64_t *d2h, int32_t reset);

void *scasa_host_alloc(int64_t bytes); /* pinned host memory for fast H2D/D2H */
void scasa_host_free(void *p);
int scasa_device_count(void);
const char *scasa_version(void);

#ifdef __cplusplus
}
#endif
#endif /* SCASA_CUDA_H */
