"""Host-side mirror of the reference's quant interface on top of libscasa_cuda.

The names are the reference's (paths relative to /root/reference/scasa/SCRIPTS/):

* ``chunk_split``           scasa_quant_step2_v1.0.0.R:64-68
* ``Scasa.crpcount_td``     Scasa_Rsource.R:416-557   (per cell; list over blocks of count vectors)
* ``Scasa.estim_chunk``     Scasa_Rsource.R:944-1051  (per chunk; cells x isoform-columns matrix)
* ``Scasa.scasa_genemat``   scasa_quant_step3_v1.0.0.R:32-46
* ``Scasa.quant``           the body of step2 (:88-136) + step3 (:48) in one call on all chunks

Everything numeric happens in the CUDA library; this file only converts between the
reference's shapes (lists of per-block matrices) and the flat arrays of the C ABI.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import lib as _l
from .xmatrix import XMatrix


def chunk_split(n_cells: int, core: int = 4) -> np.ndarray:
    """0-based cell offsets of the chunks step2 would form (``core`` matters only for n <= 100)."""
    L = _l.load()
    cap = max(n_cells // 50 + 8, core + 2)
    br = np.zeros(cap, np.int64)
    k = C.c_int32(0)
    rc = L.scasa_chunk_split(n_cells, core, _l.ptr(br, C.c_int64), cap, C.byref(k))
    if rc:
        raise _l.ScasaError(rc, L.scasa_last_error(None).decode())
    return br[: k.value + 1].copy()


def eqclass_map_host(xm: XMatrix, eq_off, eq_tx, n_threads: int = 8, hamming_dist: int = 1) -> np.ndarray:
    """crpcount_td's decision per distinct eqclass (global row or -1) without a GPU: scasa_eqclass_map_host."""
    L = _l.load()
    keep = [np.ascontiguousarray(a, t) for a, t in ((xm.q_off, np.int32), (xm.crp_p_off, np.int32), (xm.crp_x_off, np.int64),
                                                    (xm.crp_x, np.float64), (xm.crp_tx, np.int32), (xm.crp_pat, np.uint8),
                                                    (xm.name_rank, np.int32))]
    crp = _l.Crp()
    crp.m.n_blocks = len(keep[0]) - 1
    crp.m.q_off, crp.m.p_off = _l.ptr(keep[0], C.c_int32), _l.ptr(keep[1], C.c_int32)
    crp.m.x_off, crp.m.x = _l.ptr(keep[2], C.c_int64), _l.ptr(keep[3], C.c_double)
    crp.col_tx, crp.row_pat, crp.name_rank = _l.ptr(keep[4], C.c_int32), _l.ptr(keep[5], C.c_uint8), _l.ptr(keep[6], C.c_int32)
    eq_off = np.ascontiguousarray(eq_off, np.int64)
    eq_tx = np.ascontiguousarray(eq_tx, np.int32)
    out = np.empty(len(eq_off) - 1, np.int32)
    rc = L.scasa_eqclass_map_host(C.byref(crp), hamming_dist, len(out), _l.ptr(eq_off, C.c_int64), _l.ptr(eq_tx, C.c_int32),
                                  _l.ptr(out, C.c_int32), n_threads)
    if rc:
        raise _l.ScasaError(rc, L.scasa_last_error(None).decode())
    return out


@dataclass
class QuantResult:
    iso: np.ndarray            # (n_cells, Σp) — estim_chunk's result for all chunks, cell order
    gene: np.ndarray | None    # (n_cells, n_genes)
    colsum: np.ndarray         # (Σp,) sums over cells: rows with colsum > 0 survive step2:167
    iters: np.ndarray          # (n_chunks, n_em_blocks, 2) outer / total inner iterations
    timing: dict


class Scasa:
    """One GPU context with the X-matrix resident (``load(design.matrix)`` + ``DM0``)."""

    def __init__(self, xm: XMatrix, device: int = 0, with_crp: bool = True, **opts):
        self.xm = xm
        self._L = _l.load()
        self._ctx = C.c_void_p()
        self._keep = []
        dm0 = self._xmat(xm.q_off, xm.dm0_p_off, xm.dm0_x_off, xm.dm0_x)
        crp = None
        if with_crp:
            crp = _l.Crp()
            crp.m = self._xmat(xm.q_off, xm.crp_p_off, xm.crp_x_off, xm.crp_x)
            crp.col_tx = _l.ptr(self._hold(xm.crp_tx, np.int32), C.c_int32)
            crp.row_pat = _l.ptr(self._hold(xm.crp_pat, np.uint8), C.c_uint8)
            crp.name_rank = _l.ptr(self._hold(xm.name_rank, np.int32), C.c_int32)
        o = _l.default_opts(**opts)
        rc = self._L.scasa_ctx_create(C.byref(dm0), C.byref(crp) if crp is not None else None, C.byref(o),
                                      device, C.byref(self._ctx))
        if rc:
            raise _l.ScasaError(rc, self._L.scasa_last_error(None).decode())
        self.n_rows = int(self._L.scasa_n_rows(self._ctx))
        self.n_cols = int(self._L.scasa_n_cols(self._ctx))
        self.n_em = int(self._L.scasa_n_em_blocks(self._ctx))
        self.em_blocks = np.zeros(self.n_em, np.int32)
        self._L.scasa_em_blocks(self._ctx, _l.ptr(self.em_blocks, C.c_int32))
        self.n_cells = 0
        self.n_chunks = 0
        self.n_genes = 0

    # ------------------------------------------------------------------ plumbing
    def _hold(self, a, dt):
        a = np.ascontiguousarray(a, dt)
        self._keep.append(a)
        return a

    def _xmat(self, q_off, p_off, x_off, x) -> _l.XMat:
        m = _l.XMat()
        m.n_blocks = len(q_off) - 1
        m.q_off = _l.ptr(self._hold(q_off, np.int32), C.c_int32)
        m.p_off = _l.ptr(self._hold(p_off, np.int32), C.c_int32)
        m.x_off = _l.ptr(self._hold(x_off, np.int64), C.c_int64)
        m.x = _l.ptr(self._hold(x, np.float64), C.c_double)
        return m

    def _ck(self, rc: int) -> None:
        if rc:
            raise _l.ScasaError(rc, self._L.scasa_last_error(self._ctx).decode())

    def close(self) -> None:
        if self._ctx:
            self._L.scasa_ctx_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_opts(self, **opts) -> None:
        o = _l.default_opts(**opts)
        self._ck(self._L.scasa_set_opts(self._ctx, C.byref(o)))

    def poor_blocks(self) -> np.ndarray:
        f = np.zeros(self.xm.n_blocks, np.uint8)
        self._ck(self._L.scasa_poor_blocks(self._ctx, _l.ptr(f, C.c_uint8)))
        return f.astype(bool)

    # ------------------------------------------------------------------ staged pipeline
    def stage_counts(self, chunk_off, y_rowptr, y_rowidx, y_val) -> None:
        chunk_off = np.ascontiguousarray(chunk_off, np.int64)
        y_rowptr = np.ascontiguousarray(y_rowptr, np.int64)
        y_rowidx = np.ascontiguousarray(y_rowidx, np.int32)
        y_val = np.ascontiguousarray(y_val, np.float64)
        self.n_cells, self.n_chunks = len(y_rowptr) - 1, len(chunk_off) - 1
        self._ck(self._L.scasa_stage_counts(self._ctx, self.n_cells, self.n_chunks, _l.ptr(chunk_off, C.c_int64),
                                            _l.ptr(y_rowptr, C.c_int64), _l.ptr(y_rowidx, C.c_int32),
                                            _l.ptr(y_val, C.c_double)))

    def stage_eqcounts(self, chunk_off, cell_ptr, eq_id, eq_count, eq_row) -> None:
        chunk_off = np.ascontiguousarray(chunk_off, np.int64)
        cell_ptr = np.ascontiguousarray(cell_ptr, np.int64)
        eq_id = np.ascontiguousarray(eq_id, np.int32)
        eq_count = np.ascontiguousarray(eq_count, np.int32)
        eq_row = np.ascontiguousarray(eq_row, np.int32)
        self.n_cells, self.n_chunks = len(cell_ptr) - 1, len(chunk_off) - 1
        self._ck(self._L.scasa_stage_eqcounts(self._ctx, self.n_cells, self.n_chunks, _l.ptr(chunk_off, C.c_int64),
                                              _l.ptr(cell_ptr, C.c_int64), _l.ptr(eq_id, C.c_int32),
                                              _l.ptr(eq_count, C.c_int32), len(eq_row), _l.ptr(eq_row, C.c_int32)))

    def set_gene_map(self, gene_of_col, n_genes: int) -> None:
        if gene_of_col is None:                                   # no gene sums in the following runs
            self._ck(self._L.scasa_set_gene_map(self._ctx, None, 0))
            self.n_genes = 0
            return
        g = np.ascontiguousarray(gene_of_col, np.int32)
        assert len(g) == self.n_cols
        self._ck(self._L.scasa_set_gene_map(self._ctx, _l.ptr(g, C.c_int32), n_genes))
        self.n_genes = n_genes

    def run(self) -> dict:
        self._ck(self._L.scasa_run(self._ctx))
        return self.timing()

    def timing(self) -> dict:
        t = _l.Timing()
        self._ck(self._L.scasa_last_timing(self._ctx, C.byref(t)))
        return t.as_dict()

    def transfer_bytes(self, reset: bool = True):
        """(host->device, device->host) bytes moved over PCIe since the last reset."""
        a, b = C.c_int64(0), C.c_int64(0)
        self._ck(self._L.scasa_transfer_bytes(self._ctx, C.byref(a), C.byref(b), int(reset)))
        return a.value, b.value

    def set_run_parts(self, n: int) -> None:
        """Ranges of chunks scasa_run overlaps on two stream sets (0 = automatic, 1 = no overlap)."""
        self._ck(self._L.scasa_set_run_parts(self._ctx, n))

    def set_host_threads(self, n: int) -> None:
        """Threads that re-expand the zero-suppressed results on the host (0 = all cores, -1 = plain dense copy)."""
        self._ck(self._L.scasa_set_host_threads(self._ctx, n))

    def fetch_iso(self, out: np.ndarray | None = None) -> np.ndarray:
        if out is None:
            out = np.empty((self.n_cells, self.n_cols), np.float64)
        self._ck(self._L.scasa_fetch_iso(self._ctx, _l.ptr(out, C.c_double)))
        return out

    def fetch_gene(self, out: np.ndarray | None = None) -> np.ndarray:
        if out is None:
            out = np.empty((self.n_cells, self.n_genes), np.float64)
        self._ck(self._L.scasa_fetch_gene(self._ctx, _l.ptr(out, C.c_double)))
        return out

    def result_nnz(self):
        """(entries of the sparse isoform result, entries of the sparse gene result)."""
        a, b = C.c_int64(0), C.c_int64(0)
        self._ck(self._L.scasa_result_nnz(self._ctx, C.byref(a), C.byref(b)))
        return a.value, b.value

    def fetch_iso_csr(self):
        """The isoform result as the device holds it: CSR by cell (rowptr, cols, vals), columns ascending,
        over the entries that can be non-zero (stored zeros possible)."""
        nnz, _ = self.result_nnz()
        rowptr = np.empty(self.n_cells + 1, np.int64)
        cols, vals = np.empty(max(nnz, 1), np.int32), np.empty(max(nnz, 1), np.float64)
        self._ck(self._L.scasa_fetch_iso_csr(self._ctx, _l.ptr(rowptr, C.c_int64), _l.ptr(cols, C.c_int32), _l.ptr(vals, C.c_double)))
        return rowptr, cols[:nnz], vals[:nnz]

    def fetch_gene_csr(self):
        """The gene result (scasa_genemat) as CSR by cell: (rowptr, genes, vals), genes ascending."""
        _, nnz = self.result_nnz()
        rowptr = np.empty(self.n_cells + 1, np.int64)
        genes, vals = np.empty(max(nnz, 1), np.int32), np.empty(max(nnz, 1), np.float64)
        self._ck(self._L.scasa_fetch_gene_csr(self._ctx, _l.ptr(rowptr, C.c_int64), _l.ptr(genes, C.c_int32), _l.ptr(vals, C.c_double)))
        return rowptr, genes[:nnz], vals[:nnz]

    def fetch_iters(self) -> np.ndarray:
        out = np.empty((self.n_chunks, self.n_em, 2), np.int32)
        self._ck(self._L.scasa_fetch_iters(self._ctx, _l.ptr(out, C.c_int32)))
        return out

    def fetch_colsum(self) -> np.ndarray:
        out = np.empty(self.n_cols, np.float64)
        self._ck(self._L.scasa_fetch_colsum(self._ctx, _l.ptr(out, C.c_double)))
        return out

    def fetch_counts(self):
        """Y as built by the pack stage: CSR (rowptr, rowidx, val)."""
        rowptr = np.zeros(self.n_cells + 1, np.int64)
        nnz = C.c_int64(0)
        self._L.scasa_fetch_counts(self._ctx, _l.ptr(rowptr, C.c_int64), None, None, 0, C.byref(nnz))
        idx = np.zeros(max(nnz.value, 1), np.int32)
        val = np.zeros(max(nnz.value, 1), np.float64)
        self._ck(self._L.scasa_fetch_counts(self._ctx, _l.ptr(rowptr, C.c_int64), _l.ptr(idx, C.c_int32),
                                            _l.ptr(val, C.c_double), len(idx), C.byref(nnz)))
        return rowptr, idx[: nnz.value], val[: nnz.value]

    # ------------------------------------------------------------------ reference-shaped calls
    def eqclass_map(self, eq_off, eq_tx, n_threads: int = 8) -> np.ndarray:
        """crpcount_td's decision per distinct eqclass: global row or -1."""
        eq_off = np.ascontiguousarray(eq_off, np.int64)
        eq_tx = np.ascontiguousarray(eq_tx, np.int32)
        out = np.empty(len(eq_off) - 1, np.int32)
        self._ck(self._L.scasa_eqclass_map(self._ctx, len(out), _l.ptr(eq_off, C.c_int64), _l.ptr(eq_tx, C.c_int32),
                                           _l.ptr(out, C.c_int32), n_threads))
        return out

    def crpcount_td(self, transcript, count, eqclass) -> list:
        """One cell's table (columns Transcript [ids], Count, eqClass) -> list over blocks of
        count vectors, like crpcount_td(mytd)."""
        transcript, count, eqclass = (np.asarray(a) for a in (transcript, count, eqclass))
        ids, inv = np.unique(eqclass, return_inverse=True)
        order = np.argsort(inv, kind="stable")
        eq_off = np.concatenate([[0], np.cumsum(np.bincount(inv, minlength=len(ids)))]).astype(np.int64)
        eq_tx = transcript[order].astype(np.int32)
        first = np.zeros(len(ids), np.int64)
        first[inv[::-1]] = np.arange(len(inv))[::-1]
        eq_cnt = count[first].astype(np.int32)
        eq_row = self.eqclass_map(eq_off, eq_tx, 1)
        self.stage_eqcounts([0, 1], [0, len(ids)], np.arange(len(ids), dtype=np.int32), eq_cnt, eq_row)
        self.run()
        rowptr, idx, val = self.fetch_counts()
        dense = np.zeros(self.n_rows)
        dense[idx] = val
        return [dense[self.xm.q_off[k]:self.xm.q_off[k + 1]] for k in range(self.xm.n_blocks)]

    def estim_chunk(self, Y: list, **opts) -> np.ndarray:
        """estim_chunk(Y, DM0, ...) for ONE chunk: Y = list over blocks of (q_k, n) arrays."""
        if opts:
            self.set_opts(**opts)
        n = Y[0].shape[1]
        dense = np.concatenate([np.asarray(y, np.float64).reshape(-1, n) for y in Y], axis=0).T   # (n, Σq)
        return self.estimate_dense(dense, [0, n])[0]

    def estimate_dense(self, y_dense: np.ndarray, chunk_off):
        rowptr, idx, val = dense_to_csr(y_dense)
        return self.estimate(chunk_off, rowptr, idx, val)

    def estimate(self, chunk_off, y_rowptr, y_rowidx, y_val):
        """scasa_estimate: (iso, iters)."""
        self.stage_counts(chunk_off, y_rowptr, y_rowidx, y_val)
        self.run()
        return self.fetch_iso(), self.fetch_iters()

    def scasa_genemat(self, iso: np.ndarray, gene_of_col, n_genes: int) -> np.ndarray:
        iso = np.ascontiguousarray(iso, np.float64)
        g = np.ascontiguousarray(gene_of_col, np.int32)
        out = np.empty((iso.shape[0], n_genes), np.float64)
        self._ck(self._L.scasa_gene_sum(self._ctx, _l.ptr(g, C.c_int32), n_genes, _l.ptr(iso, C.c_double),
                                        iso.shape[0], _l.ptr(out, C.c_double)))
        self.n_genes = n_genes
        return out

    def quant(self, cell_ptr, eq_id, eq_count, eq_off, eq_tx, core: int = 4, gene_of_col=None, n_genes: int = 0,
              n_threads: int = 8, n_slices: int = 0, iso_out=None, gene_out=None, want_iters: bool = True,
              peers=()) -> QuantResult:
        """step2's two foreach bodies (+ step3's gene sums) for a whole sample: one scasa_quant call
        (host buffers in and out; chunks processed in slices that overlap compute and transfer).
        peers: further Scasa contexts on other GPUs of the box (same X-matrix): one scasa_quant_multi call,
        contiguous chunk ranges per GPU, results bitwise those of one GPU."""
        cell_ptr = np.ascontiguousarray(cell_ptr, np.int64)
        eq_id = np.ascontiguousarray(eq_id, np.int32)
        eq_count = np.ascontiguousarray(eq_count, np.int32)
        n_cells = len(cell_ptr) - 1
        chunks = chunk_split(n_cells, core)
        eq_row = self.eqclass_map(eq_off, eq_tx, n_threads)
        if gene_of_col is not None:
            self.set_gene_map(gene_of_col, n_genes)
        self.n_cells, self.n_chunks = n_cells, len(chunks) - 1
        iso = iso_out if iso_out is not None else np.empty((n_cells, self.n_cols), np.float64)
        gene = None
        if self.n_genes:
            gene = gene_out if gene_out is not None else np.empty((n_cells, self.n_genes), np.float64)
        colsum = np.empty(self.n_cols, np.float64)
        iters = np.empty((self.n_chunks, self.n_em, 2), np.int32) if want_iters else None
        tail = (n_cells, self.n_chunks, _l.ptr(chunks, C.c_int64), _l.ptr(cell_ptr, C.c_int64),
                _l.ptr(eq_id, C.c_int32), _l.ptr(eq_count, C.c_int32), len(eq_row),
                _l.ptr(eq_row, C.c_int32), _l.ptr(iso, C.c_double),
                _l.ptr(gene, C.c_double) if gene is not None else None,
                _l.ptr(colsum, C.c_double), _l.ptr(iters, C.c_int32) if iters is not None else None, n_slices)
        if peers:
            arr = (C.c_void_p * (1 + len(peers)))(self._ctx, *[p._ctx for p in peers])
            self._ck(self._L.scasa_quant_multi(arr, len(arr), *tail))
        else:
            self._ck(self._L.scasa_quant(self._ctx, *tail))
        return QuantResult(iso, gene, colsum, iters, {"n_slices": n_slices, "n_gpus": 1 + len(peers)})


def quant_files(contexts, cell_ptr, eq_id, eq_count, eq_off, eq_tx, iso_path: str, gene_path, tx_names, gene_names, barcodes,
                core: int = 4, n_threads: int = 0, gene_rank=None):
    """scasa_quant_files: eqclass counts in, scasa_isoform_counts.txt and scasa_gene_expression.txt out, over the
    given contexts (one per GPU; the first one's gene map and options are used).  Returns (colsum, timing dict)."""
    g0 = contexts[0]
    L = g0._L
    cell_ptr = np.ascontiguousarray(cell_ptr, np.int64)
    eq_id = np.ascontiguousarray(eq_id, np.int32)
    eq_count = np.ascontiguousarray(eq_count, np.int32)
    n_cells = len(cell_ptr) - 1
    chunks = chunk_split(n_cells, core)
    eq_row = g0.eqclass_map(eq_off, eq_tx, n_threads or 8)
    if gene_rank is None and gene_path:
        order = sorted(range(len(gene_names)), key=lambda g: gene_names[g].encode())
        gene_rank = np.empty(len(gene_names), np.int32)
        gene_rank[order] = np.arange(len(gene_names), dtype=np.int32)
    gr = np.ascontiguousarray(gene_rank, np.int32) if gene_rank is not None else None
    colsum = np.empty(g0.n_cols, np.float64)
    tm = _l.FilesTiming()
    arr = (C.c_void_p * len(contexts))(*[c._ctx for c in contexts])
    rc = L.scasa_quant_files(arr, len(contexts), n_cells, len(chunks) - 1, _l.ptr(chunks, C.c_int64), _l.ptr(cell_ptr, C.c_int64),
                             _l.ptr(eq_id, C.c_int32), _l.ptr(eq_count, C.c_int32), len(eq_row), _l.ptr(eq_row, C.c_int32),
                             iso_path.encode(), gene_path.encode() if gene_path else None,
                             ("\n".join(tx_names) + "\n").encode(),
                             ("\n".join(gene_names) + "\n").encode() if gene_path else None,
                             _l.ptr(gr, C.c_int32) if gr is not None else None, ("\n".join(barcodes) + "\n").encode(), n_threads,
                             _l.ptr(colsum, C.c_double), C.byref(tm))
    if rc:
        raise _l.ScasaError(rc, L.scasa_last_error(g0._ctx).decode())
    return colsum, {k: getattr(tm, k) for k, _ in tm._fields_}


class SparseRows:
    """A scasa_csr_t: CSR by cell over the output columns == the p / i / x slots of the dgCMatrix of the reference's
    (columns x cells) matrix.  `ptr`, `col`, `val` are views of library memory that live as long as this object."""

    def __init__(self, lib, c: "_l.Csr"):
        self._L, self._c = lib, c
        self.n_rows, self.nnz = int(c.n_rows), int(c.nnz)
        self.ptr = np.ctypeslib.as_array(c.ptr, (self.n_rows + 1,))
        self.col = np.ctypeslib.as_array(c.col, (max(self.nnz, 1),))[: self.nnz]
        self.val = np.ctypeslib.as_array(c.val, (max(self.nnz, 1),))[: self.nnz]

    def to_dense(self, n_cols: int) -> np.ndarray:
        out = np.zeros((self.n_rows, n_cols))
        out[np.repeat(np.arange(self.n_rows), np.diff(self.ptr)), self.col] = self.val
        return out

    def close(self):
        if self._c is not None:
            self.ptr = self.col = self.val = None
            self._L.scasa_csr_free(C.byref(self._c))
            self._c = None

    __del__ = close


def quant_csr(contexts, cell_ptr, eq_id, eq_count, eq_off, eq_tx, core: int = 4, drop_zeros: bool = True, want_gene: bool = True,
              want_iters: bool = False, n_threads: int = 0):
    """scasa_quant_csr: eqclass counts in (host), sparse isoform and gene matrices out (host), over the given contexts
    (one per GPU).  Returns (iso: SparseRows, gene: SparseRows | None, colsum, iters | None)."""
    g0 = contexts[0]
    L = g0._L
    cell_ptr = np.ascontiguousarray(cell_ptr, np.int64)
    eq_id = np.ascontiguousarray(eq_id, np.int32)
    eq_count = np.ascontiguousarray(eq_count, np.int32)
    n_cells = len(cell_ptr) - 1
    chunks = chunk_split(n_cells, core)
    eq_row = g0.eqclass_map(eq_off, eq_tx, n_threads or 8)
    colsum = np.empty(g0.n_cols, np.float64)
    iters = np.zeros((len(chunks) - 1, g0.n_em, 2), np.int32) if want_iters else None
    iso, gene = _l.Csr(), _l.Csr()
    arr = (C.c_void_p * len(contexts))(*[c._ctx for c in contexts])
    rc = L.scasa_quant_csr(arr, len(contexts), n_cells, len(chunks) - 1, _l.ptr(chunks, C.c_int64), _l.ptr(cell_ptr, C.c_int64),
                           _l.ptr(eq_id, C.c_int32), _l.ptr(eq_count, C.c_int32), len(eq_row), _l.ptr(eq_row, C.c_int32),
                           1 if drop_zeros else 0, C.byref(iso), C.byref(gene) if want_gene else None, _l.ptr(colsum, C.c_double),
                           _l.ptr(iters, C.c_int32) if iters is not None else None)
    if rc:
        raise _l.ScasaError(rc, L.scasa_last_error(g0._ctx).decode())
    return SparseRows(L, iso), (SparseRows(L, gene) if want_gene else None), colsum, iters


def dense_to_csr(y_dense: np.ndarray):
    """(n_cells, Σq) dense counts -> CSR by cell (rows ascending)."""
    y_dense = np.asarray(y_dense)
    nz = y_dense != 0
    rowptr = np.concatenate([[0], np.cumsum(nz.sum(axis=1))]).astype(np.int64)
    cells, rows = np.nonzero(nz)
    return rowptr, rows.astype(np.int32), y_dense[cells, rows].astype(np.float64)
