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


def write_isoform_counts(path: str, iso: np.ndarray, colsum: np.ndarray, xm, barcodes, n_threads: int = 0) -> None:
    """scasa_isoform_counts.txt: transcripts (DM0 column names) x cells, rows with rowSums > 0 (step2:167)."""
    write_table(path, iso, xm.dm0_colnames(), barcodes, keep=np.asarray(colsum) > 0, n_threads=n_threads)


def write_gene_expression(path: str, gene: np.ndarray, gene_names, barcodes, n_threads: int = 0) -> None:
    """scasa_gene_expression.txt: genes x cells in scasa_genemat's row order (step3:43)."""
    write_table(path, gene, gene_names, barcodes, n_threads=n_threads)
