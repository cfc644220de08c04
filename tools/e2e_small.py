"""End-to-end latency of one sample of the BASELINE shapes C1 (200 cells) and C2 (3,955 cells): context creation
(X-matrix upload), eqclass map, quant into host matrices; cold (first call in the process) and warm."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from scasa_b200 import synth
from scasa_b200.api import Scasa
from bench import load_x, gene_map
xm = load_x()
goc, n_genes = gene_map(xm)
for n, seed in ((200, 1), (3955, 2)):
    rowptr, rows, vals = synth.CellGenerator(xm, seed=seed).counts_csr(0, n)
    ed = synth.eqclass_dictionary(xm, seed)
    cell_ptr, eq_id, eq_cnt = synth.counts_to_eqcounts(rowptr, rows, vals, seed)
    for rep in range(3):
        t0 = time.perf_counter()
        g = Scasa(xm, device=0)
        t1 = time.perf_counter()
        r = g.quant(cell_ptr, eq_id, eq_cnt, ed.eq_off, ed.eq_tx, gene_of_col=goc, n_genes=n_genes)
        t2 = time.perf_counter()
        r2 = g.quant(cell_ptr, eq_id, eq_cnt, ed.eq_off, ed.eq_tx, gene_of_col=goc, n_genes=n_genes)
        t3 = time.perf_counter()
        g.close()
        print(f"n={n} rep {rep}: ctx {1e3 * (t1 - t0):.1f} ms, first quant {1e3 * (t2 - t1):.1f} ms, second quant {1e3 * (t3 - t2):.1f} ms", flush=True)
