#!/usr/bin/env python
"""One short, fixed workload for ncu: N cells of the C4 generator, W warm-up runs + K timed runs.
Prints the library's phase timings and the iteration-count distribution of the AEM tasks."""
import os
import sys
import json
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from scasa_b200 import synth  # noqa: E402
from scasa_b200.api import Scasa, chunk_split  # noqa: E402
from scasa_b200.xmatrix import XMatrix  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
runs = int(sys.argv[2]) if len(sys.argv) > 2 else 2
xm = XMatrix.load_npz(os.path.join(ROOT, "tests", "golden", "xmatrix_hg38_alevin.npz"))
dev = torch.device("cuda", 0)
rowptr, rows, vals = synth.CellGenerator(xm, seed=4, device=dev).counts_csr(0, n)
ed = synth.eqclass_dictionary(xm, 4)
cell_ptr, eq_id, eq_cnt = synth.counts_to_eqcounts(rowptr, rows, vals, 4, device=dev)
g = Scasa(xm, device=0)
g.set_gene_map(np.asarray(xm.gene_of_col, np.int32), len(xm.gene_names))
eq_row = g.eqclass_map(ed.eq_off, ed.eq_tx, 8)
g.stage_eqcounts(chunk_split(n, 4), cell_ptr, eq_id, eq_cnt, eq_row)
for _ in range(runs):
    t = g.run()
print(json.dumps(t))
it = g.fetch_iters()
outer, inner = it[..., 0].ravel(), it[..., 1].ravel()
for name, v in (("outer", outer), ("inner", inner)):
    qs = np.percentile(v, [50, 90, 99, 99.9, 100])
    print(name, "mean %.2f" % v.mean(), "p50/p90/p99/p99.9/max", qs.tolist(), "at cap", int((v >= 1000).sum()))
heavy = np.argsort(-inner)[:10]
print("heaviest tasks (em idx, inner, outer):", [(int(h % it.shape[1]), int(inner[h]), int(outer[h])) for h in heavy])
