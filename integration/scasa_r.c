/* scasa_r.c — the .Call shim between Scasa's R scripts and libscasa_cuda.
 *
 *     R CMD SHLIB scasa_r.c -I<repo>/include -L<dir of libscasa_cuda.so> -lscasa_cuda
 *
 * Built on the user's machine (R's headers are not part of this repository; tests/test_integration.py
 * checks the file against include/scasa_cuda.h with stand-in headers).  All logic lives behind the C ABI:
 * this file only converts R vectors to the plain arrays the library takes.  R has no 64-bit integer
 * type, so offsets arrive as doubles (exact up to 2^53) and are converted here. */
#include <R.h>
#include <Rinternals.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "scasa_cuda.h"

#define SCASA_R_MAX_GPUS 16
typedef struct { int n; scasa_ctx *ctx[SCASA_R_MAX_GPUS]; } scasa_r_handle;   /* one context per GPU of the box */

static void handle_fin(SEXP p)
{
    scasa_r_handle *h = (scasa_r_handle *)R_ExternalPtrAddr(p);
    if (h) {
        for (int i = 0; i < h->n; i++) scasa_ctx_destroy(h->ctx[i]);
        free(h);
    }
    R_ClearExternalPtr(p);
}

static scasa_r_handle *handle_of(SEXP p)
{
    scasa_r_handle *h = (scasa_r_handle *)R_ExternalPtrAddr(p);
    if (!h || h->n < 1) error("scasa context was destroyed");
    return h;
}

static int64_t *as_i64(SEXP v)          /* numeric vector -> transient int64 array (freed by R at the end of .Call) */
{
    const R_xlen_t n = XLENGTH(v);
    int64_t *out = (int64_t *)R_alloc((size_t)n + 1, sizeof(int64_t));
    const double *x = REAL(v);
    for (R_xlen_t i = 0; i < n; i++) out[i] = (int64_t)x[i];
    return out;
}

static void xmat_of(SEXP l, scasa_xmat_t *m)   /* list(q_off, p_off, x_off, x) as flat() of the patch builds it */
{
    m->n_blocks = (int32_t)(XLENGTH(VECTOR_ELT(l, 0)) - 1);
    m->q_off = INTEGER(VECTOR_ELT(l, 0));
    m->p_off = INTEGER(VECTOR_ELT(l, 1));
    m->x_off = as_i64(VECTOR_ELT(l, 2));
    m->x = REAL(VECTOR_ELT(l, 3));
}

/* N3: TRUE when the shipped CCRP is a point `lapply(CRP, ccrpfun, clim)` (step2:127) stops at, so the serial
 * recompute can be skipped.  crp / dm0 = flat(CRP) / flat(CCRP); member = 0-based global DM0 column of every CRP
 * column (-1 = dropped), see scasa_quant_step2_patch.R. */
SEXP scasa_r_dm0_is_fixed_point(SEXP crp, SEXP dm0, SEXP member, SEXP clim)
{
    scasa_xmat_t c, d;
    int32_t ok = 0;
    xmat_of(crp, &c);
    xmat_of(dm0, &d);
    if (scasa_dm0_is_fixed_point(&c, &d, INTEGER(member), asReal(clim), 0, &ok, NULL)) error("%s", scasa_last_error(NULL));
    return ScalarLogical(ok);
}

/* dm0 = list(q_off, p_off, x_off, x); crp = list(q_off, p_off, x_off, x, col_tx, row_pat, name_rank): the flattened
 * block lists prepared in R (scasa_quant_step2_patch.R).  n_gpus: contexts on devices 0 .. n_gpus-1 (capped by the
 * devices present) — the library shards the chunks over them inside one call, like step2's fork pool does over cores. */
SEXP scasa_r_create(SEXP dm0, SEXP crp, SEXP n_gpus)
{
    scasa_xmat_t d;
    scasa_crp_t c;
    xmat_of(dm0, &d);
    xmat_of(crp, &c.m);
    c.col_tx = INTEGER(VECTOR_ELT(crp, 4));
    c.row_pat = RAW(VECTOR_ELT(crp, 5));
    c.name_rank = INTEGER(VECTOR_ELT(crp, 6));
    int n = asInteger(n_gpus), have = scasa_device_count();
    if (n < 1) n = 1;
    if (n > have) n = have;
    if (n > SCASA_R_MAX_GPUS) n = SCASA_R_MAX_GPUS;
    if (have < 1) error("no CUDA device: libscasa_cuda has no CPU path");
    scasa_r_handle *h = (scasa_r_handle *)calloc(1, sizeof *h);
    if (!h) error("out of memory");
    for (int i = 0; i < n; i++) {
        /* only the first context needs CRP (the eqclass -> row map is taken once, on the host) */
        if (scasa_ctx_create(&d, i == 0 ? &c : NULL, NULL, i, &h->ctx[i])) {
            for (int k = 0; k < i; k++) scasa_ctx_destroy(h->ctx[k]);
            free(h);
            error("%s", scasa_last_error(NULL));
        }
        h->n = i + 1;
    }
    SEXP p = PROTECT(R_MakeExternalPtr(h, R_NilValue, R_NilValue));
    R_RegisterCFinalizerEx(p, handle_fin, TRUE);
    UNPROTECT(1);
    return p;
}

/* common argument conversion of the two quant calls */
typedef struct {
    int64_t n, E, *cp, *eo, *ch;
    int32_t nch, *row;
} quant_in;

static void quant_prepare(scasa_r_handle *h, SEXP cell_ptr, SEXP eq_off, SEXP eq_tx, SEXP core, quant_in *q)
{
    q->n = (int64_t)XLENGTH(cell_ptr) - 1;
    q->E = (int64_t)XLENGTH(eq_off) - 1;
    q->cp = as_i64(cell_ptr);
    q->eo = as_i64(eq_off);
    const int nc = asInteger(core);
    const int64_t cap = (q->n / 50 + 8 > nc + 2) ? q->n / 50 + 8 : nc + 2;   /* n <= 100: `core` chunks (step2:66) */
    q->ch = (int64_t *)R_alloc((size_t)cap, sizeof(int64_t));
    q->row = (int32_t *)R_alloc((size_t)q->E + 1, sizeof(int32_t));
    q->nch = 0;
    if (scasa_chunk_split(q->n, nc, q->ch, cap, &q->nch)) error("%s", scasa_last_error(NULL));   /* host-only entry: global message */
    if (scasa_eqclass_map(h->ctx[0], q->E, q->eo, INTEGER(eq_tx), q->row, 0 /* all cores */)) error("%s", scasa_last_error(h->ctx[0]));
}

/* The quant call that replaces step2:88-136 (crpcount_td over all cells + estim_chunk over all chunks) and, when
 * gene_of_col is not NULL, step3:48 (scasa_genemat) as well.
 * cell_ptr (numeric, n+1), eq_id / eq_count (integer, one per (cell, eqclass)), eq_off (numeric, E+1), eq_tx (integer,
 * transcript ids of the eqclasses), core (step2's chunk rule for n <= 100), gene_of_col (integer, 0-based gene id per
 * isoform row, NA -> -1) or NULL, n_genes.
 * Returns list(iso = tx x cells, gene = genes x cells or NULL, colsum = rowSums of iso): column-major (tx, cell) is the
 * library's cell-major layout, so no transposition happens anywhere. */
SEXP scasa_r_quant(SEXP p, SEXP cell_ptr, SEXP eq_id, SEXP eq_count, SEXP eq_off, SEXP eq_tx, SEXP core, SEXP gene_of_col,
                   SEXP n_genes)
{
    scasa_r_handle *h = handle_of(p);
    quant_in q;
    quant_prepare(h, cell_ptr, eq_off, eq_tx, core, &q);
    const int want_gene = gene_of_col != R_NilValue, ng = want_gene ? asInteger(n_genes) : 0;
    if (scasa_set_gene_map(h->ctx[0], want_gene ? INTEGER(gene_of_col) : NULL, ng)) error("%s", scasa_last_error(h->ctx[0]));
    const int ncol = (int)scasa_n_cols(h->ctx[0]);
    SEXP out = PROTECT(allocVector(VECSXP, 3));
    SEXP iso = PROTECT(allocMatrix(REALSXP, ncol, (int)q.n));
    SEXP gene = PROTECT(want_gene ? allocMatrix(REALSXP, ng, (int)q.n) : R_NilValue);
    SEXP colsum = PROTECT(allocVector(REALSXP, ncol));
    if (scasa_quant_multi(h->ctx, h->n, q.n, q.nch, q.ch, q.cp, INTEGER(eq_id), INTEGER(eq_count), q.E, q.row, REAL(iso),
                          want_gene ? REAL(gene) : NULL, REAL(colsum), NULL /* iters */, 0 /* slices: auto */))
        error("%s", scasa_last_error(h->ctx[0]));
    SET_VECTOR_ELT(out, 0, iso);
    SET_VECTOR_ELT(out, 1, gene);
    SET_VECTOR_ELT(out, 2, colsum);
    UNPROTECT(4);
    return out;
}

/* The same with sparse results: list(iso = list(p, i, x), gene = list(p, i, x) or NULL, colsum).  p / i / x are the slots of
 * a Matrix::dgCMatrix with Dim = c(columns, cells) — `new("dgCMatrix", p = , i = , x = , Dim = , Dimnames = )` — i.e. the
 * reference's tx x cells `isoform_count` (step2:166) and the gene matrix of step 3 without their zeros: no dense matrix on
 * either side of the call.  Stored zeros are dropped.  R's integer slots limit a matrix to 2^31 - 1 entries. */
static SEXP csr_to_list(const scasa_csr_t *m)
{
    SEXP l = PROTECT(allocVector(VECSXP, 3));
    SEXP p = PROTECT(allocVector(INTSXP, (R_xlen_t)m->n_rows + 1));
    SEXP i = PROTECT(allocVector(INTSXP, (R_xlen_t)m->nnz));
    SEXP x = PROTECT(allocVector(REALSXP, (R_xlen_t)m->nnz));
    for (int64_t c = 0; c <= m->n_rows; c++) INTEGER(p)[c] = (int)m->ptr[c];
    if (m->nnz) {
        memcpy(INTEGER(i), m->col, (size_t)m->nnz * sizeof(int));
        memcpy(REAL(x), m->val, (size_t)m->nnz * sizeof(double));
    }
    SET_VECTOR_ELT(l, 0, p);
    SET_VECTOR_ELT(l, 1, i);
    SET_VECTOR_ELT(l, 2, x);
    UNPROTECT(4);
    return l;
}
SEXP scasa_r_quant_sparse(SEXP p, SEXP cell_ptr, SEXP eq_id, SEXP eq_count, SEXP eq_off, SEXP eq_tx, SEXP core, SEXP gene_of_col,
                          SEXP n_genes)
{
    scasa_r_handle *h = handle_of(p);
    quant_in q;
    quant_prepare(h, cell_ptr, eq_off, eq_tx, core, &q);
    const int want_gene = gene_of_col != R_NilValue, ng = want_gene ? asInteger(n_genes) : 0;
    if (scasa_set_gene_map(h->ctx[0], want_gene ? INTEGER(gene_of_col) : NULL, ng)) error("%s", scasa_last_error(h->ctx[0]));
    scasa_csr_t iso, gene;
    SEXP colsum = PROTECT(allocVector(REALSXP, (R_xlen_t)scasa_n_cols(h->ctx[0])));
    if (scasa_quant_csr(h->ctx, h->n, q.n, q.nch, q.ch, q.cp, INTEGER(eq_id), INTEGER(eq_count), q.E, q.row, 1 /* drop zeros */,
                        &iso, want_gene ? &gene : NULL, REAL(colsum), NULL /* iters */))
        error("%s", scasa_last_error(h->ctx[0]));
    if (iso.nnz > 2147483647LL || (want_gene && gene.nnz > 2147483647LL)) {
        scasa_csr_free(&iso);
        if (want_gene) scasa_csr_free(&gene);
        error("more than 2^31 - 1 entries: a dgCMatrix cannot hold them; use scasa_r_quant_files or fewer cells per call");
    }
    SEXP out = PROTECT(allocVector(VECSXP, 3));
    SET_VECTOR_ELT(out, 0, csr_to_list(&iso));
    scasa_csr_free(&iso);
    if (want_gene) { SET_VECTOR_ELT(out, 1, csr_to_list(&gene)); scasa_csr_free(&gene); }
    SET_VECTOR_ELT(out, 2, colsum);
    UNPROTECT(2);
    return out;
}

/* The same, but straight to scasa_isoform_counts.txt and scasa_gene_expression.txt (step2:165-170, step3:48-52): R never
 * holds the dense matrices.  tx_names / gene_names / cell_names: one string each, names joined and terminated by "\n";
 * gene_rank: 0-based rank of every gene name under R's sort() (the session collation decides the row order, step3:36).
 * gene_path, gene_of_col may be NULL.  Returns the rowSums of the isoform matrix (rows with 0 are not written). */
SEXP scasa_r_quant_files(SEXP p, SEXP cell_ptr, SEXP eq_id, SEXP eq_count, SEXP eq_off, SEXP eq_tx, SEXP core, SEXP gene_of_col,
                         SEXP n_genes, SEXP iso_path, SEXP gene_path, SEXP tx_names, SEXP gene_names, SEXP gene_rank,
                         SEXP cell_names)
{
    scasa_r_handle *h = handle_of(p);
    quant_in q;
    quant_prepare(h, cell_ptr, eq_off, eq_tx, core, &q);
    const int want_gene = gene_of_col != R_NilValue && gene_path != R_NilValue;
    if (scasa_set_gene_map(h->ctx[0], want_gene ? INTEGER(gene_of_col) : NULL, want_gene ? asInteger(n_genes) : 0))
        error("%s", scasa_last_error(h->ctx[0]));
    SEXP colsum = PROTECT(allocVector(REALSXP, (R_xlen_t)scasa_n_cols(h->ctx[0])));
    if (scasa_quant_files(h->ctx, h->n, q.n, q.nch, q.ch, q.cp, INTEGER(eq_id), INTEGER(eq_count), q.E, q.row,
                          CHAR(STRING_ELT(iso_path, 0)), want_gene ? CHAR(STRING_ELT(gene_path, 0)) : NULL,
                          CHAR(STRING_ELT(tx_names, 0)), want_gene ? CHAR(STRING_ELT(gene_names, 0)) : NULL,
                          want_gene && gene_rank != R_NilValue ? INTEGER(gene_rank) : NULL, CHAR(STRING_ELT(cell_names, 0)),
                          0 /* all cores */, REAL(colsum), NULL))
        error("%s", scasa_last_error(h->ctx[0]));
    UNPROTECT(1);
    return colsum;
}

/* scasa_genemat (step3:24-46) for a matrix already in R: gene_of_col = 0-based gene row of each isoform
 * row (NA -> -1), n_genes rows out.  iso is tx x cells as returned above. */
SEXP scasa_r_genemat(SEXP p, SEXP iso, SEXP gene_of_col, SEXP n_genes)
{
    scasa_r_handle *h = handle_of(p);
    const int ng = asInteger(n_genes), n = ncols(iso);
    SEXP out = PROTECT(allocMatrix(REALSXP, ng, n));
    if (scasa_gene_sum(h->ctx[0], INTEGER(gene_of_col), ng, REAL(iso), n, REAL(out))) error("%s", scasa_last_error(h->ctx[0]));
    UNPROTECT(1);
    return out;
}
