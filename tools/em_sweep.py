#!/usr/bin/env python
"""Device time of the C4 step under different tuning switches (read when a context is created).
    python tools/em_sweep.py [cells] "KEY=VAL[;KEY=VAL]" ...    (each argument = one configuration; "-" = defaults)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from scasa_b200 import synth
from scasa_b200.api import Scasa, chunk_split
from scasa_b200.xmatrix import XMatrix
args = sys.argv[1:]
n = int(args.pop(0)) if args and args[0].isdigit() else 100000
xm = XMatrix.load_npz(os.path.join(ROOT, "tests", "golden", "xmatrix_hg38_alevin.npz"))
dev = torch.device("cuda", 0)
rowptr, rows, vals = synth.CellGenerator(xm, seed=4, device=dev).counts_csr(0, n)
ed = synth.eqclass_dictionary(xm, 4)
cell_ptr, eq_id, eq_cnt = synth.counts_to_eqcounts(rowptr, rows, vals, 4, device=dev)
torch.cuda.empty_cache()
chunks = chunk_split(n, 4)
goc = np.asarray(xm.gene_of_col, np.int32)
eq_row = None
for cfg in args or ["-"]:
    keys = []
    if cfg != "-":
        for kv in cfg.split(";"):
            k, v = kv.split("=", 1); os.environ[k] = v; keys.append(k)
    g = Scasa(xm, device=0)
    g.set_gene_map(goc, len(xm.gene_names))
    if eq_row is None:
        eq_row = g.eqclass_map(ed.eq_off, ed.eq_tx, os.cpu_count() or 8)
    g.stage_eqcounts(chunks, cell_ptr, eq_id, eq_cnt, eq_row)
    ts = [g.run() for _ in range(5)][2:]
    f = lambda k: float(np.mean([t[k] for t in ts]))
    print(f"{cfg:48s} total {f('total_ms'):6.2f} pack {f('pack_ms'):5.2f} tasks {f('tasks_ms'):5.2f} em {f('em_ms'):6.2f} "
          f"finish {f('finish_ms'):5.2f} (gene {f('gene_ms'):5.2f}) cls {[round(x, 1) for x in ts[-1]['em_class_ms']]} pair {ts[-1]['em_pair_ms']:.1f}", flush=True)
    g.close()
    for k in keys: os.environ.pop(k, None)
