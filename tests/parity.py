"""Shared comparison of a GPU result with the oracle's, task by task ((chunk, EM block)).

Nothing is skipped:
  * tasks whose iteration counts agree: estimates within `rtol` relative (north_star: 1e-6 in fp64);
    poor-condition blocks (Rsource:987-1008, :1027-1037 subtract `deductVal * x / sum(x)` from two
    nearly equal numbers) on the absolute scale `rtol * max(1, block total of the cell)`;
  * tasks whose iteration counts differ (a 1-ulp difference flipped a `max(err) < maxerr` break,
    SURVEY.md §7 iii): one more or one fewer iteration of a loop that stops when a step moves less than
    maxerr in the reference's own measure |a - b| / max(b, lim) (Rsource:755, :1019) — so the two
    results must agree within 2 * maxerr in that measure;
  * pass-through columns (nrow(X0) == 1, Rsource:979-983) are copies: bit-exact.
"""
from __future__ import annotations

import numpy as np


def compare_tasks(iso, ref, iters, ref_iters, chunks, em_blocks, p_off, q_off, poor, rtol=1e-6, maxerr=0.01, lim=0.01,
                  min_same=0.995):
    """iso/ref: (n_cells, Σp); iters/ref_iters: (n_chunks, n_em, 2) for the same em_blocks.
    Returns a dict of what was seen; raises AssertionError on the first violated bound."""
    pt = np.repeat(np.diff(q_off) == 1, np.diff(p_off))
    assert np.array_equal(iso[:, pt], ref[:, pt]), "pass-through columns differ"
    same = np.all(iters == ref_iters, axis=2)
    assert same.mean() >= min_same, f"iteration counts agree on {same.mean():.4f} of the tasks"
    worst_same, worst_poor, worst_diff, n_diff, n_poor = 0.0, 0.0, 0.0, 0, 0
    for h in range(len(chunks) - 1):
        cs = slice(int(chunks[h]), int(chunks[h + 1]))
        for e, k in enumerate(em_blocks):
            a, b = iso[cs, p_off[k]:p_off[k + 1]], ref[cs, p_off[k]:p_off[k + 1]]
            if not a.size:
                continue
            if same[h, e]:
                if poor[k]:
                    scale = np.maximum(1.0, np.abs(b).sum(axis=1, keepdims=True))
                    err = float(np.max(np.abs(a - b) / scale))
                    worst_poor = max(worst_poor, err)
                    n_poor += 1
                    assert err <= rtol, f"poor block {k} chunk {h}: abs err / max(1, total) = {err:.3e}"
                else:
                    err = float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-9)))
                    worst_same = max(worst_same, err)
                    assert err < rtol, f"block {k} chunk {h}: rel err {err:.3e} with equal iteration counts"
            else:
                n_diff += 1
                scale = np.maximum(np.abs(b), lim)
                if poor[k]:
                    scale = np.maximum(scale, np.abs(b).sum(axis=1, keepdims=True) * lim)
                err = float(np.max(np.abs(a - b) / scale))
                worst_diff = max(worst_diff, err)
                assert err <= 2 * maxerr, (f"block {k} chunk {h}: iterations {iters[h, e]} vs {ref_iters[h, e]}, "
                                           f"|a-b|/max(b, lim) = {err:.3e} > 2*maxerr")
    assert np.all(np.isfinite(iso))
    return {"same": float(same.mean()), "worst_same": worst_same, "worst_poor": worst_poor, "worst_diff": worst_diff,
            "n_diff": n_diff, "n_poor": n_poor}
