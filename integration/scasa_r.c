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
#include "scasa_cuda.h"

static void ctx_fin(SEXP p)
{
    scasa_ctx_destroy((scasa_ctx *)R_ExternalPtrAddr(p));
    R_ClearExternalPtr(p);
}

static int64_t *as_i64(SEXP v)          /* numeric vector -> transient int64 array (freed by R at the end of .Call) */
{
    const R_xlen_t n = XLENGTH(v);
    int64_t *out = (int64_t *)R_alloc((size_t)n + 1, sizeof(int64_t));
    const double *x = REAL(v);
    for (R_xlen_t i = 0; i < n; i++) out[i] = (int64_t)x[i];
    return out;
}

/* dm0 = list(q_off, p_off, x_off, x); crp = list(q_off, p_off, x_off, x, col_tx, row_pat, name_rank):
 * the flattened block lists prepared in R (scasa_quant_step2_patch.R) */
SEXP scasa_r_create(SEXP dm0, SEXP crp, SEXP device)
{
    scasa_xmat_t d;
    d.n_blocks = (int32_t)(XLENGTH(VECTOR_ELT(dm0, 0)) - 1);
    d.q_off = INTEGER(VECTOR_ELT(dm0, 0));
    d.p_off = INTEGER(VECTOR_ELT(dm0, 1));
    d.x_off = as_i64(VECTOR_ELT(dm0, 2));
    d.x = REAL(VECTOR_ELT(dm0, 3));
    scasa_crp_t c;
    c.m.n_blocks = (int32_t)(XLENGTH(VECTOR_ELT(crp, 0)) - 1);
    c.m.q_off = INTEGER(VECTOR_ELT(crp, 0));
    c.m.p_off = INTEGER(VECTOR_ELT(crp, 1));
    c.m.x_off = as_i64(VECTOR_ELT(crp, 2));
    c.m.x = REAL(VECTOR_ELT(crp, 3));
    c.col_tx = INTEGER(VECTOR_ELT(crp, 4));
    c.row_pat = RAW(VECTOR_ELT(crp, 5));
    c.name_rank = INTEGER(VECTOR_ELT(crp, 6));
    scasa_ctx *ctx = NULL;
    if (scasa_ctx_create(&d, &c, NULL, asInteger(device), &ctx)) error("%s", scasa_last_error(NULL));
    SEXP p = PROTECT(R_MakeExternalPtr(ctx, R_NilValue, R_NilValue));
    R_RegisterCFinalizerEx(p, ctx_fin, TRUE);
    UNPROTECT(1);
    return p;
}

/* The quant call that replaces step2:88-136 (crpcount_td over all cells + estim_chunk over all chunks).
 * cell_ptr (numeric, n+1), eq_id / eq_count (integer, one per (cell, eqclass)), eq_off (numeric, E+1),
 * eq_tx (integer, transcript ids of the eqclasses), core (step2's chunk rule for n <= 100).
 * Returns the tx x cells matrix: column-major (tx, cell) is the library's cell-major layout. */
SEXP scasa_r_quant(SEXP p, SEXP cell_ptr, SEXP eq_id, SEXP eq_count, SEXP eq_off, SEXP eq_tx, SEXP core)
{
    scasa_ctx *ctx = (scasa_ctx *)R_ExternalPtrAddr(p);
    if (!ctx) error("scasa context was destroyed");
    const int64_t n = (int64_t)XLENGTH(cell_ptr) - 1, E = (int64_t)XLENGTH(eq_off) - 1;
    int64_t *cp = as_i64(cell_ptr), *eo = as_i64(eq_off);
    const int64_t cap = n / 50 + 8;
    int64_t *ch = (int64_t *)R_alloc((size_t)cap, sizeof(int64_t));
    int32_t *row = (int32_t *)R_alloc((size_t)E + 1, sizeof(int32_t));
    int32_t nch = 0;
    SEXP iso = PROTECT(allocMatrix(REALSXP, (int)scasa_n_cols(ctx), (int)n));
    if (scasa_chunk_split(n, asInteger(core), ch, cap, &nch) ||
        scasa_eqclass_map(ctx, E, eo, INTEGER(eq_tx), row, 0 /* all cores */) ||
        scasa_quant(ctx, n, nch, ch, cp, INTEGER(eq_id), INTEGER(eq_count), E, row, REAL(iso),
                    NULL /* gene_out: after scasa_r_gene_map */, NULL /* colsum */, NULL /* iters */, 0 /* slices: auto */))
        error("%s", scasa_last_error(ctx));
    UNPROTECT(1);
    return iso;
}

/* scasa_genemat (step3:24-46) for a matrix already in R: gene_of_col = 0-based gene row of each isoform
 * row (NA -> -1), n_genes rows out.  iso is tx x cells as returned above. */
SEXP scasa_r_genemat(SEXP p, SEXP iso, SEXP gene_of_col, SEXP n_genes)
{
    scasa_ctx *ctx = (scasa_ctx *)R_ExternalPtrAddr(p);
    if (!ctx) error("scasa context was destroyed");
    const int ng = asInteger(n_genes), n = ncols(iso);
    SEXP out = PROTECT(allocMatrix(REALSXP, ng, n));
    if (scasa_gene_sum(ctx, INTEGER(gene_of_col), ng, REAL(iso), n, REAL(out))) error("%s", scasa_last_error(ctx));
    UNPROTECT(1);
    return out;
}
