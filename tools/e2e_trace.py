"""One traced Scasa.quant call (SCASA_FETCH_TRACE=1): where the end-to-end time goes."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from scasa_b200 import lib as _l, synth
from scasa_b200.api import Scasa
from bench import load_x, gene_map
n = 100000
xm = load_x()
dev = torch.device("cuda", 0)
gen = synth.CellGenerator(xm, seed=4, device=dev)
rowptr, rows, vals = gen.counts_csr(0, n)
ed = synth.eqclass_dictionary(xm, 4)
cell_ptr, eq_id, eq_cnt = synth.counts_to_eqcounts(rowptr, rows, vals, 4, device=dev)
del gen; torch.cuda.empty_cache()
goc, n_genes = gene_map(xm)
g = Scasa(xm, device=0)
g.set_gene_map(goc, n_genes)
h_ptr = _l.pinned_empty(n + 1, np.int64); h_ptr[:] = cell_ptr
h_id = _l.pinned_empty(len(eq_id), np.int32); h_id[:] = eq_id
h_cnt = _l.pinned_empty(len(eq_cnt), np.int32); h_cnt[:] = eq_cnt
h_iso = np.empty((n, xm.n_cols)); h_gene = np.empty((n, n_genes))
ns = int(sys.argv[1]) if len(sys.argv) > 1 else 0
for rep in range(3):
    if rep == 2:
        print("=== traced call", flush=True)
    t0 = time.perf_counter()
    g.quant(h_ptr, h_id, h_cnt, ed.eq_off, ed.eq_tx, n_threads=16, iso_out=h_iso, gene_out=h_gene, want_iters=False, n_slices=ns)
    print("quant", time.perf_counter() - t0, flush=True)
