#!/usr/bin/env python
"""Wall time of scasa_quant_csr (host eqclass counts in, sparse host matrices out) next to scasa_quant (dense out).
    SCASA_FETCH_TRACE=1 python tools/csr_probe.py [cells]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from scasa_b200 import synth
from scasa_b200.api import Scasa, quant_csr
from scasa_b200.xmatrix import XMatrix
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
xm = XMatrix.load_npz(os.path.join(ROOT, "tests", "golden", "xmatrix_hg38_alevin.npz"))
dev = torch.device("cuda", 0)
rowptr, rows, vals = synth.CellGenerator(xm, seed=4, device=dev).counts_csr(0, n)
ed = synth.eqclass_dictionary(xm, 4)
cell_ptr, eq_id, eq_cnt = synth.counts_to_eqcounts(rowptr, rows, vals, 4, device=dev)
cell_ptr, eq_id, eq_cnt = (np.ascontiguousarray(a) for a in (cell_ptr, eq_id, eq_cnt))
g = Scasa(xm, device=0)
g.set_gene_map(np.asarray(xm.gene_of_col, np.int32), len(xm.gene_names))
g.set_host_threads(os.cpu_count() or 8)
for drop in (True, False):
    for rep in range(3):
        t0 = time.perf_counter()
        iso, gene, colsum, _ = quant_csr([g], cell_ptr, eq_id, eq_cnt, ed.eq_off, ed.eq_tx, drop_zeros=drop, n_threads=os.cpu_count() or 8)
        t1 = time.perf_counter()
        nz = (iso.nnz, gene.nnz)
        iso.close(); gene.close()
        print(f"quant_csr drop_zeros={drop}: {t1 - t0:.3f} s ({n / (t1 - t0):,.0f} cells/s), entries {nz}, free {time.perf_counter() - t1:.3f} s", flush=True)
iso_out = np.empty((n, xm.n_cols)); gene_out = np.empty((n, len(xm.gene_names)))
for rep in range(3):
    t0 = time.perf_counter()
    g.quant(cell_ptr, eq_id, eq_cnt, ed.eq_off, ed.eq_tx, n_threads=os.cpu_count() or 8, iso_out=iso_out, gene_out=gene_out, want_iters=False)
    print(f"quant (dense): {time.perf_counter() - t0:.3f} s", flush=True)
