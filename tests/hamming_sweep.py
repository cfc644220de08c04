"""The sweep SURVEY.md §9 Q1 asks for: every shipped EM block x every Hamming-1 and Hamming-2
perturbation of its row patterns (over the block's supported columns: an observed pattern never has a
1 in a column with colSums(crp) == 0, Rsource:468-474), turned into eqclass transcript sets.

Shared by tests/test_crpcount_literal.py (literal vs C oracle, CPU) and tests/test_gpu_counts.py
(literal vs scasa_eqclass_map)."""
from __future__ import annotations

import os
from itertools import combinations

import numpy as np


def sweep_classes(xm, stride: int = 1, offset: int = 0):
    """(eq_off int64[n+1], eq_tx int32[], block int32[n]): the perturbed classes, every `stride`-th one."""
    eq_tx, sizes, blocks = [], [], []
    i = 0
    for k in xm.em_blocks():
        pat = xm.crp_patterns(k).astype(bool)
        sup = np.nonzero(xm.crp_block(k).sum(axis=0) > 0)[0]
        tx = xm.crp_tx[xm.crp_p_off[k]:xm.crp_p_off[k + 1]]
        seen = set()
        flips = [()] + [(a,) for a in sup] + list(combinations(sup, 2))
        for r in range(pat.shape[0]):
            base = np.zeros(pat.shape[1], bool)
            base[sup] = pat[r, sup]
            for f in flips:
                o = base.copy()
                for a in f:
                    o[a] = not o[a]
                key = o.tobytes()
                if not o.any() or key in seen:
                    continue
                seen.add(key)
                if (i - offset) % stride == 0:
                    t = tx[o]
                    eq_tx.append(t)
                    sizes.append(len(t))
                    blocks.append(k)
                i += 1
    eq_off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    return eq_off, np.concatenate(eq_tx).astype(np.int32), np.asarray(blocks, np.int32)


_LIT = None


def _init(npz_path):
    global _LIT
    from oracle.crpcount_literal import CrpLiteral
    from scasa_b200.xmatrix import XMatrix
    _LIT = CrpLiteral(XMatrix.load_npz(npz_path))


def _rows(args):
    eq_off, eq_tx = args
    from oracle import crpcount_literal as CL
    names = _LIT.xm.tx_names
    t0 = CL.STATS["tie_breaks"]
    rows = [_LIT.eqclass_row([names[t] for t in eq_tx[eq_off[e]:eq_off[e + 1]]]) for e in range(len(eq_off) - 1)]
    return rows, CL.STATS["tie_breaks"] - t0


def literal_rows(npz_path: str, eq_off, eq_tx, procs: int = 0):
    """oracle.crpcount_literal.CrpLiteral.eqclass_row for every class, over a process pool.
    Returns (rows int32[n], number of classes the median tie-break decided)."""
    import multiprocessing as mp
    n = len(eq_off) - 1
    procs = procs or min(os.cpu_count() or 1, 32)
    cuts = np.linspace(0, n, procs * 8 + 1).astype(int)
    jobs = []
    for a, b in zip(cuts[:-1], cuts[1:]):
        if b > a:
            jobs.append((eq_off[a:b + 1] - eq_off[a], eq_tx[eq_off[a]:eq_off[b]]))
    ctx = mp.get_context("fork")
    with ctx.Pool(procs, initializer=_init, initargs=(npz_path,)) as pool:
        parts = pool.map(_rows, jobs)
    return np.asarray([r for p, _ in parts for r in p], np.int32), int(sum(t for _, t in parts))
