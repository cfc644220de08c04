"""ctypes front end of the CPU oracle (oracle/scasa_oracle*.c).  TEST INFRASTRUCTURE.

Imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs — never by scasa_b200 (the product fails loudly without its
CUDA library instead of falling back to this).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libscasa_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith(".c")]
    stale = (not os.path.exists(_SO)) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs)
    if force or stale:
        subprocess.check_call(["make", "-s", "-C", _HERE] + (["-B"] if force else []))
    return _SO


class Opts(C.Structure):
    """estim_chunk's defaults, Scasa_Rsource.R:944 (+ adjust_esp :973)."""
    _fields_ = [("maxiter_x", C.c_int), ("maxiter_bt", C.c_int), ("maxerr", C.c_double),
                ("lim", C.c_double), ("modify", C.c_int), ("adjust_esp", C.c_double),
                ("fixed_iter", C.c_int), ("maxerr_bt", C.c_double), ("lim_bt", C.c_double)]

    @classmethod
    def make(cls, maxiter_x=1000, maxiter_bt=1000, maxerr=0.01, lim=0.01, modify=1,
             adjust_esp=2.0, fixed_iter=0, maxerr_bt=0.01, lim_bt=0.01):
        """maxerr / lim reach the outer test only (Rsource:1019-1020); estim.BETA keeps its own defaults
        (Rsource:726-727, not forwarded at :1016) unless maxerr_bt / lim_bt say otherwise."""
        return cls(maxiter_x, maxiter_bt, maxerr, lim, modify, adjust_esp, fixed_iter, maxerr_bt, lim_bt)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _lib.oracle_digamma.restype = C.c_double
        _lib.oracle_digamma.argtypes = [C.c_double]
        _lib.oracle_keep_rows.restype = C.c_int64
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def digamma(x):
    f = lib().oracle_digamma
    return np.array([f(float(v)) for v in np.atleast_1d(x)])


def chunk_split(n: int, core: int = 4) -> np.ndarray:
    """0-based chunk offsets, scasa_quant_step2_v1.0.0.R:64-68."""
    br = np.zeros(max(n // 50 + 8, core + 2), np.int64)
    k = lib().oracle_chunk_split(C.c_int64(n), C.c_int(core), _p(br, C.c_int64), C.c_int64(len(br)))
    if k < 0:
        raise RuntimeError("chunk_split: buffer too small")
    return br[: k + 1].copy()


def estim_chunks(xm, y_dense: np.ndarray, chunk_off, opts: Opts | None = None, nthreads: int = 1,
                 x_override=None):
    """estim_chunk (Scasa_Rsource.R:944-1051) on every chunk.

    xm: XMatrix (DM0 part used); y_dense: (n_cells, Σq) float64 cell-major.
    Returns (iso (n_cells, Σp), iters (n_chunks, B, 2))."""
    opts = opts or Opts.make()
    y_dense = np.ascontiguousarray(y_dense, np.float64)
    n_cells = y_dense.shape[0]
    assert y_dense.shape[1] == xm.n_rows
    chunk_off = np.ascontiguousarray(chunk_off, np.int64)
    n_chunks = len(chunk_off) - 1
    out = np.zeros((n_cells, xm.n_cols), np.float64)
    iters = np.zeros((n_chunks, xm.n_blocks, 2), np.int32)
    x = xm.dm0_x if x_override is None else np.ascontiguousarray(x_override, np.float64)
    rc = lib().oracle_estim_chunks(
        C.c_int(xm.n_blocks), _p(xm.q_off, C.c_int32), _p(xm.dm0_p_off, C.c_int32),
        _p(xm.dm0_x_off, C.c_int64), _p(x, C.c_double), C.c_int(n_chunks), _p(chunk_off, C.c_int64),
        _p(y_dense, C.c_double), C.byref(opts), _p(out, C.c_double), _p(iters, C.c_int32),
        C.c_int(nthreads))
    if rc:
        raise RuntimeError(f"oracle_estim_chunks failed ({rc})")
    return out, iters


def estim_block(X0: np.ndarray, Ymat: np.ndarray, opts: Opts | None = None):
    """estim_chunk on a single ad-hoc block: X0 (q,p), Ymat (q,n) → (BT (n,p), (outer, inner))."""
    opts = opts or Opts.make()
    q, p = X0.shape
    n = Ymat.shape[1]
    q_off = np.array([0, q], np.int32)
    p_off = np.array([0, p], np.int32)
    x_off = np.array([0, q * p], np.int64)
    x = np.ascontiguousarray(X0.T, np.float64).reshape(-1)        # column-major
    y = np.ascontiguousarray(Ymat.T, np.float64)                  # cell-major (n, q)
    out = np.zeros((n, p), np.float64)
    it = np.zeros(2, np.int32)
    rc = lib().oracle_estim_chunk(C.c_int(1), _p(q_off, C.c_int32), _p(p_off, C.c_int32),
                                  _p(x_off, C.c_int64), _p(x, C.c_double), C.c_int(n),
                                  _p(y, C.c_double), C.byref(opts), _p(out, C.c_double), _p(it, C.c_int32))
    if rc:
        raise RuntimeError("oracle_estim_chunk failed")
    return out, (int(it[0]), int(it[1]))


def keep_rows(iso: np.ndarray) -> np.ndarray:
    """step2:167 — which isoform rows survive (rowSums > 0)."""
    iso = np.ascontiguousarray(iso, np.float64)
    keep = np.zeros(iso.shape[1], np.uint8)
    lib().oracle_keep_rows(_p(iso, C.c_double), C.c_int64(iso.shape[0]), C.c_int64(iso.shape[1]),
                           _p(keep, C.c_uint8))
    return keep.astype(bool)


def gene_layout(col_gene_names, keep):
    """Row layout of scasa_genemat (scasa_quant_step3_v1.0.0.R:34-43), LC_COLLATE=C.

    col_gene_names[j]: gene name of isoform column j (None = NA, dropped by tapply).
    Returns (gene_of_col int32[ncol], n_members int32[G], gene_names list[G]): singles
    (sorted by name) first, then multis (sorted by name)."""
    members: dict = {}
    for j, g in enumerate(col_gene_names):
        if keep[j] and g is not None:
            members.setdefault(g, []).append(j)
    singles = sorted((g for g, v in members.items() if len(v) == 1), key=lambda s: s.encode())
    multis = sorted((g for g, v in members.items() if len(v) > 1), key=lambda s: s.encode())
    names = singles + multis
    gene_of_col = np.full(len(col_gene_names), -1, np.int32)
    n_members = np.zeros(len(names), np.int32)
    for gi, g in enumerate(names):
        n_members[gi] = len(members[g])
        gene_of_col[members[g]] = gi
    return gene_of_col, n_members, names


def gene_sum(iso: np.ndarray, gene_of_col, n_members) -> np.ndarray:
    iso = np.ascontiguousarray(iso, np.float64)
    gene_of_col = np.ascontiguousarray(gene_of_col, np.int32)
    n_members = np.ascontiguousarray(n_members, np.int32)
    out = np.zeros((iso.shape[0], len(n_members)), np.float64)
    lib().oracle_gene_sum(_p(iso, C.c_double), C.c_int64(iso.shape[0]), C.c_int64(iso.shape[1]),
                          _p(gene_of_col, C.c_int32), C.c_int32(len(n_members)),
                          _p(n_members, C.c_int32), _p(out, C.c_double))
    return out


# ------------------------------------------------------------------ crpcount_td / hamming_correction
class CrpHandle:
    """oracle_crp_create over an XMatrix' CRP part (keeps the arrays alive)."""

    def __init__(self, xm, name_rank=None):
        L = lib()
        L.oracle_crp_create.restype = C.c_void_p
        self.xm = xm
        self._a = [np.ascontiguousarray(xm.q_off, np.int32), np.ascontiguousarray(xm.crp_p_off, np.int32),
                   np.ascontiguousarray(xm.crp_x_off, np.int64), np.ascontiguousarray(xm.crp_x, np.float64),
                   np.ascontiguousarray(xm.crp_tx, np.int32), np.ascontiguousarray(xm.crp_pat, np.uint8),
                   np.ascontiguousarray(xm.name_rank if name_rank is None else name_rank, np.int32)]
        self.n_tx = int(max(len(xm.tx_names), int(self._a[4].max(initial=0)) + 1))
        a = self._a
        self.h = C.c_void_p(L.oracle_crp_create(
            C.c_int(xm.n_blocks), _p(a[0], C.c_int32), _p(a[1], C.c_int32), _p(a[2], C.c_int64), _p(a[3], C.c_double),
            _p(a[4], C.c_int32), _p(a[5], C.c_uint8), _p(a[6], C.c_int32), C.c_int(self.n_tx)))

    def __del__(self):
        try:
            lib().oracle_crp_free(self.h)
        except Exception:
            pass


def crpcount_td(handle: CrpHandle, transcript, count, eqclass, hamming_dist: int = 1) -> np.ndarray:
    """crpcount_td(mytd) for one cell: table columns Transcript (ids), Count, eqClass (rows of one
    class contiguous, as in Sample_eqClass).  Returns the dense count vector over all rows (Σq)."""
    tx = np.ascontiguousarray(transcript, np.int32)
    cnt = np.ascontiguousarray(count, np.int32)
    eq = np.ascontiguousarray(eqclass, np.int32)
    y = np.zeros(handle.xm.n_rows, np.float64)
    lib().oracle_crpcount_td(handle.h, C.c_int(len(tx)), _p(tx, C.c_int32), _p(cnt, C.c_int32), _p(eq, C.c_int32),
                             C.c_int(hamming_dist), _p(y, C.c_double))
    return y


def crpcount_cells(handle_or_xm, eq_off, eq_tx, cell_ptr, eq_id, eq_count, hamming_dist: int = 1,
                   nthreads: int = 1) -> np.ndarray:
    """crpcount_td for every cell of a sample given as per-cell (eqclass id, count) lists plus the
    eqclass dictionary.  Returns dense Y (n_cells, Σq)."""
    h = handle_or_xm if isinstance(handle_or_xm, CrpHandle) else CrpHandle(handle_or_xm)
    eq_off = np.ascontiguousarray(eq_off, np.int64)
    eq_tx = np.ascontiguousarray(eq_tx, np.int32)
    cell_ptr = np.ascontiguousarray(cell_ptr, np.int64)
    eq_id = np.ascontiguousarray(eq_id, np.int32)
    eq_count = np.ascontiguousarray(eq_count, np.int32)
    n = len(cell_ptr) - 1
    y = np.zeros((n, h.xm.n_rows), np.float64)
    lib().oracle_crpcount_cells(h.h, C.c_int64(n), _p(cell_ptr, C.c_int64), _p(eq_id, C.c_int32),
                                _p(eq_count, C.c_int32), _p(eq_off, C.c_int64), _p(eq_tx, C.c_int32),
                                C.c_int(hamming_dist), C.c_int(nthreads), _p(y, C.c_double))
    return y


def eqclass_rows(handle: CrpHandle, eq_off, eq_tx, nthreads: int = 1) -> np.ndarray:
    """Where crpcount_td puts each eqclass when it is the only class of a cell (count 1): the global
    row, or -1 if it is dropped.  This is the per-eqclass view the GPU library's map must equal."""
    n = len(eq_off) - 1
    y = crpcount_cells(handle, eq_off, eq_tx, np.arange(n + 1), np.arange(n), np.ones(n), nthreads=nthreads)
    out = np.full(n, -1, np.int32)
    cells, rows = np.nonzero(y)
    assert len(np.unique(cells)) == len(cells), "an eqclass was counted in two rows"
    out[cells] = rows
    return out


def format_roundtrip(v):
    v = np.ascontiguousarray(v, np.float64)
    back = np.zeros_like(v)
    z = np.zeros(len(v), np.uint8)
    lib().oracle_format_roundtrip(_p(v, C.c_double), C.c_int(len(v)), _p(back, C.c_double), _p(z, C.c_uint8))
    return back, z.astype(bool)
