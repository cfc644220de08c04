#!/usr/bin/env python
"""The single-call multi-GPU entries measured in ONE process (what the R master would do): scasa_quant_multi
(dense host matrices out) and scasa_quant_files (text files out) over 1, 2, 4, 8 contexts = GPUs, same sample
(strong scaling).  python tools/multi_quant_bench.py [cells = 100000] [files_cells = 20000]"""
import json, os, sys, time, tempfile, shutil
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from scasa_b200 import lib as _l, synth
from scasa_b200.api import Scasa, quant_csr, quant_files
from scasa_b200.xmatrix import XMatrix
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
nf = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
xm = XMatrix.load_npz(os.path.join(ROOT, "tests", "golden", "xmatrix_hg38_alevin.npz"))
dev = torch.device("cuda", 0)
rowptr, rows, vals = synth.CellGenerator(xm, seed=4, device=dev).counts_csr(0, n)
ed = synth.eqclass_dictionary(xm, 4)
cell_ptr, eq_id, eq_cnt = synth.counts_to_eqcounts(rowptr, rows, vals, 4, device=dev)
torch.cuda.empty_cache()
goc, ng = np.asarray(xm.gene_of_col, np.int32), len(xm.gene_names)
n_dev = _l.load().scasa_device_count()
ctx = [Scasa(xm, device=d, with_crp=(d == 0)) for d in range(n_dev)]
ctx[0].set_gene_map(goc, ng)
h_iso = np.empty((n, xm.n_cols)); h_gene = np.empty((n, ng))
threads = os.cpu_count() or 16
ctx[0].set_host_threads(threads)
out = {"cells": n, "files_cells": nf, "host_threads": threads, "gpus_present": n_dev, "dense": {}, "csr": {}, "files": {}}
ref = None
for k in [g for g in (1, 2, 4, 8) if g <= n_dev]:
    ts = []
    for rep in range(3):
        t0 = time.perf_counter()
        r = ctx[0].quant(cell_ptr, eq_id, eq_cnt, ed.eq_off, ed.eq_tx, n_threads=threads, iso_out=h_iso, gene_out=h_gene,
                         want_iters=False, peers=ctx[1:k])
        ts.append(time.perf_counter() - t0)
    chk = (float(h_iso.sum()), float(h_gene.sum()), float(r.colsum.sum()))
    if ref is None:
        ref = chk
    out["dense"][k] = {"s": min(ts[1:]), "cells_per_s": n / min(ts[1:]), "same_sums_as_1_gpu": chk == ref}
    print(k, "GPUs dense:", out["dense"][k], flush=True)
ref = None
for k in [g for g in (1, 2, 4, 8) if g <= n_dev]:                  # the same call with sparse host results (scasa_quant_csr)
    ts = []
    for rep in range(3):
        t0 = time.perf_counter()
        iso, gene, colsum, _ = quant_csr(ctx[:k], cell_ptr, eq_id, eq_cnt, ed.eq_off, ed.eq_tx, drop_zeros=True, n_threads=threads)
        ts.append(time.perf_counter() - t0)
        chk = (iso.nnz, gene.nnz, float(iso.val.sum()), float(gene.val.sum()), float(colsum.sum()))
        iso.close(); gene.close()
    if ref is None:
        ref = chk
    out["csr"][k] = {"s": min(ts[1:]), "cells_per_s": n / min(ts[1:]), "entries": [chk[0], chk[1]], "same_as_1_gpu": chk == ref}
    print(k, "GPUs sparse:", out["csr"][k], flush=True)
if nf <= 0:
    for c in ctx:
        c.close()
    print(json.dumps(out))
    sys.exit(0)
tmp = tempfile.mkdtemp(prefix="scasa_multi_")
try:
    bc = synth.barcodes(nf, 4)
    nz = int(cell_ptr[nf])
    tx = xm.dm0_colnames()
    sizes = None
    for k in [g for g in (1, 2, 4, 8) if g <= n_dev]:
        ts = []
        for rep in range(2):
            t0 = time.perf_counter()
            _, tm = quant_files(ctx[:k], cell_ptr[: nf + 1], eq_id[:nz], eq_cnt[:nz], ed.eq_off, ed.eq_tx, os.path.join(tmp, f"iso{k}.txt"),
                                os.path.join(tmp, f"gene{k}.txt"), tx, xm.gene_names, bc, n_threads=threads)
            ts.append(time.perf_counter() - t0)
        same = True
        if k > 1:
            same = open(os.path.join(tmp, "iso1.txt"), "rb").read() == open(os.path.join(tmp, f"iso{k}.txt"), "rb").read() and \
                open(os.path.join(tmp, "gene1.txt"), "rb").read() == open(os.path.join(tmp, f"gene{k}.txt"), "rb").read()
        out["files"][k] = {"s": min(ts), "cells_per_s": nf / min(ts), "phases_s": {p: tm[p] for p in ("quant_s", "transpose_s", "write_iso_s", "write_gene_s")},
                           "text_bytes": tm["iso_bytes"] + tm["gene_bytes"], "files_identical_to_1_gpu": same}
        print(k, "GPUs files:", out["files"][k], flush=True)
finally:
    shutil.rmtree(tmp, ignore_errors=True)
for c in ctx:
    c.close()
print(json.dumps(out))
