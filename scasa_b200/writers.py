"""The two text outputs of 2QUANT, byte-compatible with the reference's ``write.table`` calls
(scasa_quant_step2_v1.0.0.R:165-170, scasa_quant_step3_v1.0.0.R:51-52); host side of
``scasa_write_table``."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import lib as _l


def format_real15(x: float) -> str:
    """The text ``write.table`` gives one number (R_print.digits = 15)."""
    buf = C.create_string_buffer(64)
    n = _l.load().scasa_format_real15(float(x), buf, 64)
    if n < 0:
        raise _l.ScasaError(n, "scasa_format_real15")
    return buf.value.decode()


def write_table(path: str, mat: np.ndarray, row_names, cell_names, keep=None, n_threads: int = 0) -> None:
    """``write.table(t(mat)[keep, ], path, quote=F, row.names=T, sep="\\t")`` for a cell-major matrix
    ``mat`` (n_cells x n_cols): one line per kept column."""
    L = _l.load()
    mat = np.ascontiguousarray(mat, np.float64)
    n_cells, n_cols = mat.shape
    assert len(row_names) == n_cols and len(cell_names) == n_cells
    rn = ("\n".join(row_names) + "\n").encode()
    cn = ("\n".join(cell_names) + "\n").encode()
    k = None
    if keep is not None:
        k = np.ascontiguousarray(keep, np.uint8)
        assert len(k) == n_cols
    rc = L.scasa_write_table(path.encode(), _l.ptr(mat, C.c_double), n_cells, n_cols,
                             _l.ptr(k, C.c_uint8) if k is not None else None, rn, cn, n_threads)
    if rc:
        raise _l.ScasaError(rc, L.scasa_last_error(None).decode())


def csr_transpose(rowptr, cols, vals, n_cols: int, n_threads: int = 0):
    """CSR by row -> (colptr, rows, vals) with the entries of a column in row order (scasa_csr_transpose)."""
    L = _l.load()
    rowptr = np.ascontiguousarray(rowptr, np.int64)
    cols = np.ascontiguousarray(cols, np.int32)
    vals = np.ascontiguousarray(vals, np.float64)
    colptr = np.empty(n_cols + 1, np.int64)
    rows_out, vals_out = np.empty(max(len(cols), 1), np.int32), np.empty(max(len(cols), 1), np.float64)
    rc = L.scasa_csr_transpose(len(rowptr) - 1, n_cols, _l.ptr(rowptr, C.c_int64), _l.ptr(cols, C.c_int32), _l.ptr(vals, C.c_double),
                               _l.ptr(colptr, C.c_int64), _l.ptr(rows_out, C.c_int32), _l.ptr(vals_out, C.c_double), n_threads)
    if rc:
        raise _l.ScasaError(rc, L.scasa_last_error(None).decode())
    return colptr, rows_out[: len(cols)], vals_out[: len(cols)]


def write_table_sparse(path: str, n_cells: int, colptr, cell_idx, vals, out_src, out_names, cell_names, n_threads: int = 0) -> int:
    """The bytes of ``write_table`` from a column-wise sparse matrix (scasa_write_table_sparse): line k = source
    column out_src[k] under out_names[k].  Returns the file size."""
    L = _l.load()
    colptr = np.ascontiguousarray(colptr, np.int64)
    cell_idx = np.ascontiguousarray(cell_idx, np.int32)
    vals = np.ascontiguousarray(vals, np.float64)
    out_src = np.ascontiguousarray(out_src, np.int32)
    assert len(out_names) == len(out_src) and len(cell_names) == n_cells
    rn = ("\n".join(out_names) + "\n").encode()
    cn = ("\n".join(cell_names) + "\n").encode()
    size = C.c_int64(0)
    rc = L.scasa_write_table_sparse(path.encode(), n_cells, len(colptr) - 1, _l.ptr(colptr, C.c_int64), _l.ptr(cell_idx, C.c_int32),
                                    _l.ptr(vals, C.c_double), len(out_src), _l.ptr(out_src, C.c_int32), rn, cn, n_threads, C.byref(size))
    if rc:
        raise _l.ScasaError(rc, L.scasa_last_error(None).decode())
    return size.value


def gene_layout(gene_of_col, gene_names, keep=None, collate=None):
    """Row layout of scasa_genemat (scasa_quant_step3_v1.0.0.R:34-43) for gene ids as the library holds them
    (``scasa_set_gene_map``: gene_of_col[j] = id of isoform column j, -1 = NA): the genes that still have an
    isoform row after step2:167 (``keep`` = colsum > 0; None = all), single-isoform genes first (mat1), then the
    others (mat2), each group in sort() order of the gene names — ``collate`` is the sort key (default: bytes =
    LC_COLLATE=C; R sorts in the session's collation).  Returns (order, n_kept_members): the gene
    ids in output order and, per gene id, how many of its isoforms survived."""
    gene_of_col = np.asarray(gene_of_col)
    keep = np.ones(len(gene_of_col), bool) if keep is None else np.asarray(keep, bool)
    n_genes = len(gene_names)
    ok = (gene_of_col >= 0) & keep
    cnt = np.bincount(gene_of_col[ok], minlength=n_genes)
    key = collate or (lambda s: s.encode())
    singles = sorted(np.flatnonzero(cnt == 1), key=lambda g: key(gene_names[g]))
    multis = sorted(np.flatnonzero(cnt > 1), key=lambda g: key(gene_names[g]))
    return np.asarray(singles + multis, np.int32), cnt.astype(np.int32)


def write_isoform_counts(path: str, iso: np.ndarray, colsum: np.ndarray, xm, barcodes, n_threads: int = 0) -> None:
    """scasa_isoform_counts.txt: transcripts (DM0 column names) x cells, rows with rowSums > 0 (step2:167)."""
    write_table(path, iso, xm.dm0_colnames(), barcodes, keep=np.asarray(colsum) > 0, n_threads=n_threads)


def write_gene_expression(path: str, gene: np.ndarray, gene_of_col, gene_names, colsum, barcodes, n_threads: int = 0,
                          collate=None) -> None:
    """scasa_gene_expression.txt in the reference's layout.  ``gene`` (n_cells x n_genes) holds the gene sums for the
    ids of ``gene_of_col`` over ALL isoform columns (an isoform that step2:167 drops is all zero, so it changes no
    sum); the rows written are the genes that still have an isoform after the filter, single-isoform genes first,
    then the others, each group sorted by name (``gene_layout``; step3:34-43)."""
    order, _ = gene_layout(gene_of_col, gene_names, np.asarray(colsum) > 0, collate)
    sub = np.ascontiguousarray(np.asarray(gene)[:, order])
    write_table(path, sub, [gene_names[g] for g in order], barcodes, n_threads=n_threads)
