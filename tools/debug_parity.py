"""Debug helper: where does the GPU EM differ from the oracle (C1, fixed-iteration mode)?"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from scasa_b200 import synth
from scasa_b200.api import Scasa, chunk_split
from scasa_b200.xmatrix import XMatrix
from oracle import oracle as O
xm = XMatrix.load_npz(os.path.join(ROOT, "tests", "golden", "xmatrix_hg38_alevin.npz"))
rowptr, rows, vals = synth.CellGenerator(xm, seed=1).counts_csr(0, 200)
chunks = chunk_split(200, 4)
mx, mb = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (4, 8)
fp32 = int(os.environ.get('FP32', '0'))
g = Scasa(xm, device=0, fixed_iter=1, maxiter_x=mx, maxiter_bt=mb, fp32=fp32)
iso, iters = g.estimate(chunks, rowptr, rows, vals)
ref, _ = O.estim_chunks(xm, synth.csr_to_dense(rowptr, rows, vals, xm.n_rows), chunks,
                        O.Opts.make(fixed_iter=1, maxiter_x=mx, maxiter_bt=mb), nthreads=4)
print("nan gpu", np.isnan(iso).sum(), "nan ref", np.isnan(ref).sum())
poorcol = np.repeat(g.poor_blocks(), np.diff(xm.dm0_p_off))
err = np.abs(iso - ref) / np.maximum(np.abs(ref), 1e-9)
err[:, poorcol] = np.abs(iso - ref)[:, poorcol]
err = np.nan_to_num(err, nan=9.0)
bad = np.argwhere(err > 1e-6)
print("bad entries", len(bad), "max", err.max())
for fl in (1e-9, 1e-6, 1e-3):
    e2 = np.abs(iso - ref) / np.maximum(np.abs(ref), fl); e2[:, poorcol] = 0
    print("floor", fl, "max rel err non-poor", np.nanmax(e2), " poor abs max", np.nanmax(np.abs(iso - ref)[:, poorcol]))
col_block = np.repeat(np.arange(xm.n_blocks), np.diff(xm.dm0_p_off))
e3 = np.abs(iso - ref) / np.maximum(np.abs(ref), 1e-3); e3[:, poorcol] = 0
worst = np.dstack(np.unravel_index(np.argsort(-e3, axis=None)[:8], e3.shape))[0]
for c, j in worst:
    k = col_block[j]
    print(f"worst non-poor: err {e3[c, j]:.3e} block {k} q={xm.q(k)} p={xm.p(k)} cell {c} col {j - xm.dm0_p_off[k]} gpu {iso[c, j]!r} ref {ref[c, j]!r}")
seen = set()
for c, j in bad[:400]:
    k = col_block[j]
    if k in seen:
        continue
    seen.add(k)
    q, p = xm.q(k), xm.p(k)
    h = np.searchsorted(chunks, c, side="right") - 1
    print(f"block {k} q={q} p={p} poor={bool(g.poor_blocks()[k])} chunk {h} cell {c} col {j - xm.dm0_p_off[k]} gpu {iso[c, j]!r} ref {ref[c, j]!r}")
    cs = slice(chunks[h], chunks[h + 1])
    js = slice(xm.dm0_p_off[k], xm.dm0_p_off[k + 1])
    rows_with = np.nonzero(np.abs(ref[cs, js]).sum(1) + np.abs(np.nan_to_num(iso[cs, js])).sum(1))[0]
    np.set_printoptions(linewidth=220, precision=4, suppress=True)
    for rr in rows_with[:12]:
        print("   cell", rr, "gpu", iso[cs, js][rr], "ref", ref[cs, js][rr])
    if len(seen) >= 4:
        break
