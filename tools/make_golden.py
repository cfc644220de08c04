#!/usr/bin/env python
"""Golden vectors for the parity tests (tests/golden/quant_c0_48cells.npz).

The reference ships no expected outputs and R is not installed, so these vectors are produced by
the CPU oracle (oracle/) on a small seeded case of config C1's generator — 48 cells, one chunk —
on the shipped X-matrix fixture.  They freeze the oracle's answers: a later change to the oracle
or to the CUDA library that moves any number shows up as a diff against this file.

    python tools/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from scasa_b200 import synth  # noqa: E402
from scasa_b200.xmatrix import XMatrix  # noqa: E402


def sparse(a):
    r, c = np.nonzero(a)
    return r.astype(np.int32), c.astype(np.int32), a[r, c]


def main():
    xm = XMatrix.load_npz(os.path.join(ROOT, "tests", "golden", "xmatrix_hg38_alevin.npz"))
    n = 48
    rowptr, rows, vals = synth.CellGenerator(xm, seed=21).counts_csr(0, n)
    ed = synth.eqclass_dictionary(xm, 21)
    cell_ptr, eq_id, eq_cnt = synth.counts_to_eqcounts(rowptr, rows, vals, 21)
    used = np.unique(eq_id)                                       # only the classes these cells use
    remap = np.full(len(ed.eq_off) - 1, -1, np.int32)
    remap[used] = np.arange(len(used), dtype=np.int32)
    eq_off = np.concatenate([[0], np.cumsum(np.diff(ed.eq_off)[used])]).astype(np.int64)
    eq_tx = np.concatenate([ed.eq_tx[ed.eq_off[e]:ed.eq_off[e + 1]] for e in used]).astype(np.int32)
    eq_id = remap[eq_id]
    h = O.CrpHandle(xm)
    eq_row = O.eqclass_rows(h, eq_off, eq_tx, nthreads=8)
    y = O.crpcount_cells(h, eq_off, eq_tx, cell_ptr, eq_id, eq_cnt, nthreads=8)
    chunks = np.array([0, n], np.int64)
    out = dict(n_cells=n, cell_ptr=cell_ptr, eq_id=eq_id, eq_cnt=eq_cnt, eq_off=eq_off, eq_tx=eq_tx, eq_row=eq_row)
    out["y_r"], out["y_c"], out["y_v"] = sparse(y)
    for name, opts in (("fixed", O.Opts.make(fixed_iter=1, maxiter_x=4, maxiter_bt=8)), ("conv", O.Opts.make())):
        iso, iters = O.estim_chunks(xm, y, chunks, opts, nthreads=4)
        out[f"{name}_r"], out[f"{name}_c"], out[f"{name}_v"] = sparse(iso)
        out[f"{name}_iters"] = iters[0][xm.em_blocks()]
        if name == "conv":
            keep = O.keep_rows(iso)
            goc, nmem, names = O.gene_layout([xm.gene_names[g] for g in xm.gene_of_col], keep)
            gene = O.gene_sum(iso, goc, nmem)
            out["keep"] = np.packbits(keep)
            out["gene_of_col"], out["n_members"] = goc, nmem
            out["gene_r"], out["gene_c"], out["gene_v"] = sparse(gene)
    path = os.path.join(ROOT, "tests", "golden", "quant_c0_48cells.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path) // 1024, "KiB", "classes", len(used), "Y nnz", len(out["y_v"]))


if __name__ == "__main__":
    main()
