#!/usr/bin/env python
"""Where the EM work is: per (chunk, block) task the active cells, entries and iterations, split by
block class (narrow = q*p+p <= 32 and p <= 8, else wide).  Needs a GPU (runs the library once)."""
import os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from scasa_b200 import synth
from scasa_b200.api import Scasa, chunk_split
from scasa_b200.xmatrix import XMatrix

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
xm = XMatrix.load_npz(os.path.join(ROOT, "tests", "golden", "xmatrix_hg38_alevin.npz"))
rowptr, rows, vals = synth.CellGenerator(xm, seed=4, device="cuda").counts_csr(0, n)
g = Scasa(xm, device=0)
chunks = chunk_split(n, 4)
iso, it = g.estimate(chunks, rowptr, rows, vals)
em = g.em_blocks
q = np.diff(xm.q_off)[em]; p = np.diff(xm.dm0_p_off)[em]
narrow = (q * p + p <= 32) & (p <= 8)
poor = g.poor_blocks()[em]
# active cells per task
row_block = np.repeat(np.arange(xm.n_blocks), np.diff(xm.q_off))
em_index = np.full(xm.n_blocks, -1); em_index[em] = np.arange(len(em))
cells = np.repeat(np.arange(n), np.diff(rowptr))
e = em_index[row_block[rows]]
keep = e >= 0
chunk_of = np.searchsorted(chunks, cells[keep], side="right") - 1
pair = np.unique(np.stack([chunk_of, e[keep], cells[keep]]), axis=1)       # distinct (chunk, em, cell)
nact = np.zeros((len(chunks) - 1, len(em)), np.int64)
np.add.at(nact, (pair[0], pair[1]), 1)
ents = np.zeros_like(nact); np.add.at(ents, (chunk_of, e[keep]), 1)
inner, outer = it[..., 1].astype(np.int64), it[..., 0].astype(np.int64)
nonempty = nact > 0
ncell = np.where(poor[None, :] & nonempty, nact + 1, nact)
for name, sel in (("narrow", narrow & ~poor), ("narrow-poor", narrow & poor), ("wide", ~narrow & ~poor), ("wide-poor", ~narrow & poor)):
    m = nonempty & sel[None, :]
    if not m.any():
        continue
    dig = (ncell * p[None, :] * inner)[m].sum()          # digamma evaluations
    print(f"{name:12s} blocks {sel.sum():5d} tasks {m.sum():8d} cells/task {ncell[m].mean():6.1f} p {np.broadcast_to(p, nact.shape)[m].mean():4.1f} "
          f"inner mean {inner[m].mean():6.1f} p99 {np.percentile(inner[m], 99):5.0f} max {inner[m].max():4d} outer mean {outer[m].mean():5.1f} "
          f"cell-steps {(ncell * inner)[m].sum() / 1e6:8.2f}M digammas {dig / 1e6:8.2f}M")
m = nonempty
cs = (ncell * p[None, :] * inner)
for thr in (2, 4, 8, 16, 32, 64, 128):
    print(f"  digammas in tasks with inner > {thr:3d}: {cs[m & (inner > thr)].sum() / cs[m].sum():.3f}   tasks {((inner > thr) & m).sum() / m.sum():.4f}")

# ---- lane use by task shape: a warp instruction of the pair stages (S1, S3) covers 32 (cell, transcript) pairs, of
# the entry stage (S2) 32 entries; "issue slots" = inner iterations x passes, "lanes" = the useful share of them
np_ = ncell * p[None, :]
E = np.where(poor[None, :], ncell * q[None, :], ents)
qp = np.broadcast_to((q * p)[None, :], nact.shape)
tiny = nonempty & ~poor[None, :] & (np_ <= 16) & (E <= 16) & (qp <= 16)
def slots(width, mask):
    s1 = np.ceil(np_[mask] / width) * inner[mask]
    s2 = np.ceil(E[mask] / width) * inner[mask]
    used = ((np_[mask] * 2 + E[mask]) * inner[mask]).sum()
    return (2 * s1 + s2).sum(), used
print("\nshape class                          tasks   share of issue slots   lanes used of the width")
rows_ = []
for name, mask, width in (
        ("tiny (pair kernel, 16 lanes)", tiny, 16),
        ("np<=16, E<=32, qp<=32 not tiny", nonempty & ~tiny & ~poor[None, :] & (np_ <= 16) & (E <= 32) & (qp <= 32), 32),
        ("np<=16 other (incl. poor)", nonempty & ~tiny & (np_ <= 16) & ~(~poor[None, :] & (E <= 32) & (qp <= 32)), 32),
        ("np 17..32", nonempty & (np_ > 16) & (np_ <= 32), 32),
        ("np 33..64", nonempty & (np_ > 32) & (np_ <= 64), 32),
        ("np 65..96", nonempty & (np_ > 64) & (np_ <= 96), 32),
        ("np 97..256 (2 warps)", nonempty & (np_ > 96) & (np_ <= 256), 64),
        ("np 257..512 (4 warps)", nonempty & (np_ > 256) & (np_ <= 512), 128),
        ("np > 512 (8 warps)", nonempty & (np_ > 512), 256)):
    if not mask.any():
        continue
    s, used = slots(width, mask)
    # warp instructions: every one of the W warps issues each pass; a 16-lane task shares its warp with the other half
    rows_.append((name, int(mask.sum()), s * (width / 32.0), used / (s * width)))
tot = sum(r[2] for r in rows_)
for name, cnt, w, frac in rows_:
    print(f"{name:34s} {cnt:8d}   {w / tot:6.3f}                 {frac:5.2f}")
