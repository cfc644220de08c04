"""Seeded synthetic inputs of the BASELINE.json shapes (SURVEY.md §8 d).

No real eqclass data ships with the reference (its Test_Dataset is a download), so every
config is synthetic:

* cells: library size L_c ~ LogNormal(ln 4000, 0.6); block k is expressed in a cell with
  probability pi_k ~ Beta(0.3, 2.0) (fixed per block); expressed blocks share the library
  ~ Exp(1); within a block transcript usage ~ Dirichlet(0.3); Y_rc ~ Poisson(sum_j X_rj beta_cj).
* eqclass level (the crpcount_td input): every Y count is attributed to an equivalence class
  whose transcript set is the row's membership pattern, 5 % of the mass to a variant at Hamming
  distance 1, 1 % at distance 2 and 2 % to a class that also contains a transcript of another
  block.

Cells are generated in fixed batches of ``BATCH`` cells, each with its own generator seeded by
(seed, batch index), so any rank that generates a batch-aligned cell range gets the same cells
(shard-invariant).  Uses torch only as an array library (CPU here, CUDA in bench.py).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch

from .xmatrix import XMatrix

BATCH = 1000


@dataclass
class EqDict:
    """Distinct equivalence classes: class e = transcripts eq_tx[eq_off[e]:eq_off[e+1]]."""
    eq_off: np.ndarray
    eq_tx: np.ndarray
    n_variants: int = 4


def block_expression(xm: XMatrix, seed: int) -> np.ndarray:
    return np.random.default_rng(seed + 7919).beta(0.3, 2.0, xm.n_blocks)


def _x_sparse(xm: XMatrix, device) -> torch.Tensor:
    """DM0 as one sparse (Σq x Σp) matrix."""
    rows, cols, vals = [], [], []
    q = np.diff(xm.q_off)
    p = np.diff(xm.dm0_p_off)
    for k in range(xm.n_blocks):
        m = xm.dm0_block(k)
        r, c = np.nonzero(m)
        rows.append(r + xm.q_off[k])
        cols.append(c + xm.dm0_p_off[k])
        vals.append(m[r, c])
    idx = torch.tensor(np.stack([np.concatenate(rows), np.concatenate(cols)]), dtype=torch.int64)
    v = torch.tensor(np.concatenate(vals), dtype=torch.float32)
    del q, p
    return torch.sparse_coo_tensor(idx, v, (xm.n_rows, xm.n_cols), check_invariants=False).coalesce().to(device)


class CellGenerator:
    def __init__(self, xm: XMatrix, seed: int, device="cpu", lib_scale: float = 1.0):
        self.xm, self.seed, self.device, self.lib_scale = xm, seed, torch.device(device), lib_scale
        self.pi = torch.tensor(block_expression(xm, seed), dtype=torch.float32, device=self.device)
        self.X = _x_sparse(xm, self.device)
        p = np.diff(xm.dm0_p_off)
        self.col_block = torch.tensor(np.repeat(np.arange(xm.n_blocks), p), dtype=torch.int64, device=self.device)

    def _batch(self, bi: int, n: int):
        """Dense Poisson counts (n, Σq) of batch bi (first n cells of it)."""
        g = torch.Generator(device=self.device)
        g.manual_seed((self.seed * 1000003 + bi) & 0x7FFFFFFF)
        dev = self.device
        B, P = self.xm.n_blocks, self.xm.n_cols
        lib = torch.exp(np.log(4000.0) + 0.6 * torch.randn(BATCH, generator=g, device=dev)) * self.lib_scale
        z = torch.rand(BATCH, B, generator=g, device=dev) < self.pi
        w = -torch.log1p(-torch.rand(BATCH, B, generator=g, device=dev)) * z
        share = w / w.sum(dim=1, keepdim=True).clamp_min(1e-30)
        # Dirichlet(0.3) inside each block = normalised Gamma(0.3) draws
        gam = torch._standard_gamma(torch.full((BATCH, P), 0.3, device=dev), generator=g).clamp_min(1e-30)
        bsum = torch.zeros(BATCH, B, device=dev).index_add_(1, self.col_block, gam)
        usage = gam / bsum.index_select(1, self.col_block).clamp_min(1e-30)
        beta = lib[:, None] * share.index_select(1, self.col_block) * usage
        rate = torch.sparse.mm(self.X, beta.t()).t()
        y = torch.poisson(rate, generator=g)
        return y[:n]

    def counts_csr(self, cell_start: int, n_cells: int):
        """Y as CSR by cell for cells [cell_start, cell_start + n_cells); cell_start % BATCH == 0."""
        assert cell_start % BATCH == 0
        ptr = [np.zeros(1, np.int64)]
        rows, vals = [], []
        done, base = 0, 0
        while done < n_cells:
            n = min(BATCH, n_cells - done)
            y = self._batch((cell_start + done) // BATCH, n)
            nz = y.nonzero()
            v = y[nz[:, 0], nz[:, 1]]
            cnt = torch.bincount(nz[:, 0], minlength=n)
            ptr.append((base + torch.cumsum(cnt, 0)).cpu().numpy().astype(np.int64))
            rows.append(nz[:, 1].to(torch.int32).cpu().numpy())
            vals.append(v.to(torch.float64).cpu().numpy())
            base += int(cnt.sum())
            done += n
        return np.concatenate(ptr), np.concatenate(rows), np.concatenate(vals)


def csr_to_dense(rowptr, rowidx, val, n_rows: int) -> np.ndarray:
    n = len(rowptr) - 1
    out = np.zeros((n, n_rows))
    cells = np.repeat(np.arange(n), np.diff(rowptr))
    out[cells, rowidx] = val
    return out


def barcodes(n: int, seed: int) -> list:
    """n distinct sorted 16-mers."""
    rng = np.random.default_rng(seed + 104729)
    s = set()
    while len(s) < n:
        for row in rng.integers(0, 4, (n - len(s), 16)):
            s.add("".join("ACGT"[i] for i in row))
    return sorted(s)


# --------------------------------------------------------------------------------------------
def eqclass_dictionary(xm: XMatrix, seed: int) -> EqDict:
    """4 classes per X-matrix row: the row's own pattern, a Hamming-1 and a Hamming-2 variant
    (inside the block) and one that also holds a transcript of another block."""
    rng = np.random.default_rng(seed + 15485863)
    n_tx = len(xm.tx_names)
    off = [0]
    tx_out = []
    for k in range(xm.n_blocks):
        p = xm.p_crp(k)
        cols = xm.crp_tx[xm.crp_p_off[k]:xm.crp_p_off[k + 1]]
        pats = xm.crp_patterns(k)
        for r in range(xm.q(k)):
            base = pats[r].astype(bool)
            if not base.any():
                base = base.copy()
                base[0] = True
            variants = [base]
            for nflip in (1, 2):
                v = base.copy()
                if p > 1:
                    for j in rng.choice(p, size=min(nflip, p), replace=False):
                        v[j] = ~v[j]
                    if not v.any():
                        v = base.copy()
                variants.append(v)
            for vi, v in enumerate(variants):
                tx_out.append(cols[v])
                off.append(off[-1] + int(v.sum()))
            other = int(rng.integers(0, n_tx))
            span = np.unique(np.concatenate([cols[base], [other]]))
            tx_out.append(span)
            off.append(off[-1] + len(span))
    return EqDict(np.asarray(off, np.int64), np.concatenate(tx_out).astype(np.int32))


def counts_to_eqcounts(rowptr, rowidx, val, seed: int, device="cpu"):
    """Split every Y count over the 4 classes of its row (92/5/1/2 %).  Returns the per-cell CSR
    (cell_ptr, eq_id, eq_count) with eq ids ascending inside a cell (bfh order)."""
    g = torch.Generator(device=device)
    g.manual_seed((seed * 7 + 13) & 0x7FFFFFFF)
    y = torch.as_tensor(val, device=device).to(torch.float32)
    r = torch.as_tensor(rowidx, device=device).to(torch.int64)

    def binom(n, pr):
        return torch.binomial(n, torch.full_like(n, pr), generator=g)

    n1 = binom(y, 0.05)
    n2 = binom(y - n1, 0.01 / 0.95)
    n3 = binom(y - n1 - n2, 0.02 / 0.94)
    n0 = y - n1 - n2 - n3
    cnt = torch.stack([n0, n1, n2, n3], dim=1)                      # (nnz, 4)
    ids = r[:, None] * 4 + torch.arange(4, device=device)[None, :]
    keep = cnt > 0
    per_entry = keep.sum(dim=1)
    cells = torch.repeat_interleave(torch.arange(len(rowptr) - 1, device=device),
                                    torch.as_tensor(np.diff(rowptr), device=device))
    per_cell = torch.zeros(len(rowptr) - 1, dtype=torch.int64, device=device).index_add_(0, cells, per_entry)
    cell_ptr = np.concatenate([[0], torch.cumsum(per_cell, 0).cpu().numpy()]).astype(np.int64)
    return cell_ptr, ids[keep].to(torch.int32).cpu().numpy(), cnt[keep].to(torch.int32).cpu().numpy()


# --------------------------------------------------------------------------------------------
def resample_xmatrix(xm: XMatrix, n_blocks: int, seed: int, min_q: int = 1, min_p: int = 1) -> XMatrix:
    """A synthetic X-matrix of ``n_blocks`` blocks drawn with replacement from the shipped ones
    (SURVEY.md §8 d: config C3 = "GRCh38.106-like", ≈ 93k blocks; config C5 = heavy multimapping,
    ``min_q = 8, min_p = 4``).  CRP and DM0 pieces are carried along; every drawn block gets fresh
    transcript ids, so blocks share no transcript."""
    rng = np.random.default_rng(seed + 32452843)
    q_all = np.diff(xm.q_off)
    p_all = np.diff(xm.dm0_p_off)
    pool = np.nonzero((q_all >= min_q) & (p_all >= min_p))[0]
    pick = pool[rng.integers(0, len(pool), n_blocks)]
    return assemble_xmatrix(xm, pick)


def heavy_multimapping_xmatrix(xm: XMatrix, seed: int, min_q: int = 8, min_p: int = 4) -> XMatrix:
    """BASELINE config C5 (SURVEY.md §8 d): the shipped block list with every EM block (nrow > 1) replaced by a
    draw from the shipped blocks with q >= 8 rows and p >= 4 merged columns (up to 53 x 14 / 50 x 20);
    pass-through blocks stay where they are."""
    rng = np.random.default_rng(seed + 49979687)
    q_all = np.diff(xm.q_off)
    p_all = np.diff(xm.dm0_p_off)
    pool = np.nonzero((q_all >= min_q) & (p_all >= min_p))[0]
    pick = np.arange(xm.n_blocks)
    em = np.nonzero(q_all > 1)[0]
    pick[em] = pool[rng.integers(0, len(pool), len(em))]
    return assemble_xmatrix(xm, pick)


def assemble_xmatrix(xm: XMatrix, pick: np.ndarray) -> XMatrix:
    """The X-matrix whose i-th block is a copy of block pick[i] of xm, with fresh transcript ids."""
    n_blocks = len(pick)
    q_all = np.diff(xm.q_off)
    p_all = np.diff(xm.dm0_p_off)
    q = q_all[pick]
    pd = p_all[pick]
    pc = np.diff(xm.crp_p_off)[pick]
    q_off = np.concatenate([[0], np.cumsum(q)]).astype(np.int32)
    dm0_p_off = np.concatenate([[0], np.cumsum(pd)]).astype(np.int32)
    crp_p_off = np.concatenate([[0], np.cumsum(pc)]).astype(np.int32)
    dm0_x_off = np.concatenate([[0], np.cumsum(q.astype(np.int64) * pd)]).astype(np.int64)
    crp_x_off = np.concatenate([[0], np.cumsum(q.astype(np.int64) * pc)]).astype(np.int64)
    dm0_x = np.empty(dm0_x_off[-1])
    crp_x = np.empty(crp_x_off[-1])
    crp_pat = np.empty(crp_x_off[-1], np.uint8)
    crp_member = np.empty(crp_p_off[-1], np.int32)
    for i, k in enumerate(pick):
        dm0_x[dm0_x_off[i]:dm0_x_off[i + 1]] = xm.dm0_x[xm.dm0_x_off[k]:xm.dm0_x_off[k + 1]]
        crp_x[crp_x_off[i]:crp_x_off[i + 1]] = xm.crp_x[xm.crp_x_off[k]:xm.crp_x_off[k + 1]]
        crp_pat[crp_x_off[i]:crp_x_off[i + 1]] = xm.crp_pat[xm.crp_x_off[k]:xm.crp_x_off[k + 1]]
        m = xm.crp_member[xm.crp_p_off[k]:xm.crp_p_off[k + 1]]
        crp_member[crp_p_off[i]:crp_p_off[i + 1]] = np.where(m >= 0, m - xm.dm0_p_off[k] + dm0_p_off[i], -1)
    crp_tx = np.arange(crp_p_off[-1], dtype=np.int32)
    tx_names = [f"TX{t}" for t in range(crp_p_off[-1])]
    return XMatrix(q_off, crp_p_off, crp_x_off, crp_x, crp_tx, crp_pat, dm0_p_off, dm0_x_off, dm0_x, crp_member,
                   np.arange(n_blocks, dtype=np.int32), tx_names)


def write_bfh(path: str, tx_names, barcodes, eq_off, eq_tx, cell_ptr, eq_id, eq_count) -> int:
    """The sample as an alevin ``bfh.txt`` (`salmon alevin --dumpBfh`; format per scasa_quant_step1_1_v1.0.0.R:16-68):
    three count lines, transcript names, barcodes, then one line per equivalence class
    ``n_tx, tx idx.., total reads, n_bc, (bc idx, n_umi, (umi, reads) x n_umi) x n_bc``.  The count of a
    (cell, eqclass) is its number of UMIs (step1_1:50-56), so a count c becomes c UMIs with one read each.
    Only classes that occur in some cell are written (their ids are renumbered).  Returns the file size."""
    eq_id = np.asarray(eq_id)
    eq_count = np.asarray(eq_count)
    n_cells = len(cell_ptr) - 1
    cell_of = np.repeat(np.arange(n_cells), np.diff(cell_ptr))
    order = np.argsort(eq_id, kind="stable")
    ids, start = np.unique(eq_id[order], return_index=True)
    start = np.append(start, len(order))
    cells, cnts = cell_of[order].tolist(), eq_count[order].tolist()
    umi = "\tACGTACGTAC\t1"
    with open(path, "w") as f:
        f.write(f"{len(tx_names)}\n{len(barcodes)}\n{len(ids)}\n")
        f.write("\n".join(tx_names) + "\n")
        f.write("\n".join(barcodes) + "\n")
        buf = []
        for k, e in enumerate(ids.tolist()):
            a, b = int(start[k]), int(start[k + 1])
            tx = eq_tx[eq_off[e]:eq_off[e + 1]].tolist()
            body = "".join(f"\t{c}\t{n}" + umi * n for c, n in zip(cells[a:b], cnts[a:b]))
            buf.append(f"{len(tx)}\t" + "\t".join(map(str, tx)) + f"\t{sum(cnts[a:b])}\t{b - a}" + body)
            if len(buf) >= 4096:
                f.write("\n".join(buf) + "\n")
                buf = []
        if buf:
            f.write("\n".join(buf) + "\n")
        return f.tell()
