"""Reader / writer of the text dumps tools/make_golden.R produces (tests/golden/r_dumps/).

`load(dir)` is what tests/test_r_golden.py consumes.  `write_from_oracle(dir, ...)` writes the same files
from the CPU oracle: it exists so that the consumer itself is exercised on every test run (a stand-in
dump in a temporary directory) — it is NOT a substitute for running R, and nothing writes it into
tests/golden/r_dumps/."""
from __future__ import annotations

import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
R_DUMPS = os.path.join(HERE, "golden", "r_dumps")
GOLD = os.path.join(HERE, "golden", "quant_c0_48cells.npz")


def have_r_dumps(d: str = R_DUMPS) -> bool:
    return os.path.exists(os.path.join(d, "estim_chunk.tsv")) and os.path.exists(os.path.join(d, "y_counts.tsv"))


def _triplets(path, shape, cols=(0, 1)):
    a = np.zeros(shape)
    if os.path.getsize(path) == 0:
        return a
    t = np.loadtxt(path, delimiter="\t", skiprows=1, ndmin=2, dtype=np.float64)
    if t.size:
        a[t[:, cols[0]].astype(np.int64) - 1, t[:, cols[1]].astype(np.int64) - 1] = t[:, 2]
    return a


def _lines(path):
    with open(path) as f:
        return [ln.rstrip("\n") for ln in f]


def load(d: str, n_cells: int, n_rows: int, n_cols: int) -> dict:
    out = {"y": _triplets(os.path.join(d, "y_counts.tsv"), (n_cells, n_rows)),
           "est": _triplets(os.path.join(d, "estim_chunk.tsv"), (n_cells, n_cols)),
           "colnames": _lines(os.path.join(d, "estim_colnames.txt"))}
    p = os.path.join(d, "fixed_aem.tsv")
    if os.path.exists(p):
        out["fixed"] = _triplets(p, (n_cells, n_cols))
        t = np.loadtxt(p, delimiter="\t", skiprows=1, ndmin=2)
        out["fixed_cols"] = np.unique(t[:, 1].astype(np.int64) - 1) if t.size else np.zeros(0, np.int64)
    if os.path.exists(os.path.join(d, "gene.tsv")):
        out["iso_rownames"] = _lines(os.path.join(d, "iso_rownames.txt"))
        out["gene_rownames"] = _lines(os.path.join(d, "gene_rownames.txt"))
        out["gene"] = _triplets(os.path.join(d, "gene.tsv"), (len(out["gene_rownames"]), n_cells)).T   # (cells, genes)
    p = os.path.join(d, "hamming.tsv")
    if os.path.exists(p):
        rows = [ln.split("\t") for ln in _lines(p)[1:]]
        out["hamming"] = [(int(b) - 1, o, a) for b, o, a in rows]
    return out


def _write_triplets(path, header, a, transpose=False):
    r, c = np.nonzero(a)
    with open(path, "w") as f:
        f.write(header + "\n")
        for i, j in zip(r, c):
            f.write(f"{(j if transpose else i) + 1}\t{(i if transpose else j) + 1}\t{a[i, j]:.17g}\n")


def write_from_oracle(d: str, xm, oracle, hamming_stride: int = 97) -> None:
    """Stand-in dumps (same files, same format) from the CPU oracle and the literal."""
    from oracle.crpcount_literal import CrpLiteral, hamming_assign
    from .hamming_sweep import sweep_classes
    os.makedirs(d, exist_ok=True)
    with np.load(GOLD) as z:
        g = {k: z[k] for k in z.files}
    n = int(g["n_cells"])
    h = oracle.CrpHandle(xm)
    y = oracle.crpcount_cells(h, g["eq_off"], g["eq_tx"], g["cell_ptr"], g["eq_id"], g["eq_cnt"], nthreads=4)
    _write_triplets(os.path.join(d, "y_counts.tsv"), "cell\trow\tvalue", y)
    est, _ = oracle.estim_chunks(xm, y, [0, n], None, nthreads=4)
    _write_triplets(os.path.join(d, "estim_chunk.tsv"), "cell\tcol\tvalue", est)
    names = xm.dm0_colnames()
    with open(os.path.join(d, "estim_colnames.txt"), "w") as f:
        f.write("\n".join(names) + "\n")
    fixed, _ = oracle.estim_chunks(xm, y, [0, n], oracle.Opts.make(fixed_iter=1, maxiter_x=4, maxiter_bt=8), nthreads=4)
    keep_cols = np.zeros(xm.n_cols, bool)
    poor = poor_blocks(xm, oracle)
    for k in xm.em_blocks():
        if not poor[k]:
            keep_cols[xm.dm0_p_off[k]:xm.dm0_p_off[k + 1]] = True
    _write_triplets(os.path.join(d, "fixed_aem.tsv"), "cell\tcol\tvalue", fixed * keep_cols)
    keep = oracle.keep_rows(est)
    goc, nmem, gnames = oracle.gene_layout([xm.gene_names[i] if i >= 0 else None for i in xm.gene_of_col], keep)
    gene = oracle.gene_sum(est, goc, nmem)
    with open(os.path.join(d, "iso_rownames.txt"), "w") as f:
        f.write("\n".join(nm for nm, k in zip(names, keep) if k) + "\n")
    with open(os.path.join(d, "gene_rownames.txt"), "w") as f:
        f.write("\n".join(gnames) + "\n")
    _write_triplets(os.path.join(d, "gene.tsv"), "gene\tcell\tvalue", gene, transpose=True)
    eq_off, eq_tx, blocks = sweep_classes(xm, stride=hamming_stride)
    lit = CrpLiteral(xm)
    with open(os.path.join(d, "hamming.tsv"), "w") as f:
        f.write("block\tobserved\tassigned\n")
        for i, k in enumerate(blocks):
            cols = xm.crp_tx[xm.crp_p_off[k]:xm.crp_p_off[k + 1]]
            have = set(int(t) for t in eq_tx[eq_off[i]:eq_off[i + 1]])
            obs = "".join("1" if int(t) in have else "0" for t in cols)
            f.write(f"{int(k) + 1}\t{obs}\t{hamming_assign([obs], lit.block(int(k)))[0]}\n")


def poor_blocks(xm, oracle) -> np.ndarray:
    import ctypes as C
    L = oracle.lib()
    L.oracle_is_poor.argtypes = [C.POINTER(C.c_double), C.c_int, C.c_int]
    out = np.zeros(xm.n_blocks, bool)
    for k in range(xm.n_blocks):
        q, p = xm.q(k), xm.p(k)
        if q > 1 and p == 2:
            x = np.ascontiguousarray(xm.dm0_x[xm.dm0_x_off[k]:xm.dm0_x_off[k + 1]])
            out[k] = bool(L.oracle_is_poor(x.ctypes.data_as(C.POINTER(C.c_double)), q, p))
    return out
