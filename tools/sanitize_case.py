"""A small run of every kernel of the library (eqclass counts -> Y -> EM -> genes -> sparse fetch),
for compute-sanitizer.  ~150 cells on a 600-block resampled X-matrix (all block shapes, poor blocks included)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from scasa_b200 import synth
from scasa_b200.api import Scasa
from scasa_b200.xmatrix import XMatrix
xm = XMatrix.load_npz(os.path.join(ROOT, "tests", "golden", "xmatrix_hg38_alevin.npz"))
small = synth.resample_xmatrix(xm, 600, seed=77)
wide = synth.resample_xmatrix(xm, 40, seed=78, min_q=8, min_p=4)
for name, m, n in (("mixed", small, 150), ("wide", wide, 130)):
    rowptr, rows, vals = synth.CellGenerator(m, seed=5).counts_csr(0, n)
    vals = np.minimum(vals, 300.0)
    ed = synth.eqclass_dictionary(m, 5)
    cell_ptr, eq_id, eq_cnt = synth.counts_to_eqcounts(rowptr, rows, vals, 5)
    goc = (np.arange(m.n_cols) // 2).astype(np.int32)            # two isoforms per gene
    for fp32 in (0, 1):
        g = Scasa(m, device=0, fp32=fp32, maxiter_x=30, maxiter_bt=30)
        r = g.quant(cell_ptr, eq_id, eq_cnt, ed.eq_off, ed.eq_tx, gene_of_col=goc, n_genes=int(goc.max()) + 1, n_slices=2)
        eq_row = g.eqclass_map(ed.eq_off, ed.eq_tx, 2)
        from scasa_b200.api import chunk_split
        g.stage_eqcounts(chunk_split(n, 4), cell_ptr, eq_id, eq_cnt, eq_row)
        g.run()
        iso = g.fetch_iso()
        assert np.array_equal(iso, r.iso)
        print(name, "fp32" if fp32 else "fp64", "ok", float(r.iso.sum()), float(r.gene.sum()), flush=True)
        g.close()
