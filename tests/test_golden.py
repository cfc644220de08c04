"""Golden vectors (tests/golden/quant_c0_48cells.npz, made by tools/make_golden.py):
CPU: the oracle still reproduces them.  GPU: the library reproduces them with no oracle in the loop."""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "quant_c0_48cells.npz")


def dense(d, name, shape):
    a = np.zeros(shape)
    a[d[f"{name}_r"], d[f"{name}_c"]] = d[f"{name}_v"]
    return a


def test_oracle_reproduces_golden(xm, oracle):
    d = np.load(GOLD)
    n = int(d["n_cells"])
    h = oracle.CrpHandle(xm)
    assert np.array_equal(oracle.eqclass_rows(h, d["eq_off"], d["eq_tx"], nthreads=4), d["eq_row"])
    y = oracle.crpcount_cells(h, d["eq_off"], d["eq_tx"], d["cell_ptr"], d["eq_id"], d["eq_cnt"], nthreads=4)
    assert np.array_equal(y, dense(d, "y", (n, xm.n_rows)))
    iso, iters = oracle.estim_chunks(xm, y, [0, n], oracle.Opts.make(fixed_iter=1, maxiter_x=4, maxiter_bt=8), nthreads=4)
    assert np.array_equal(iso, dense(d, "fixed", (n, xm.n_cols)))
    assert np.array_equal(iters[0][xm.em_blocks()], d["fixed_iters"])


@pytest.mark.gpu
def test_gpu_reproduces_golden(xm):
    from scasa_b200.api import Scasa
    d = np.load(GOLD)
    n = int(d["n_cells"])
    g = Scasa(xm, device=0)
    try:
        # index maps: bit-exact
        assert np.array_equal(g.eqclass_map(d["eq_off"], d["eq_tx"], 4), d["eq_row"])
        for mode, opts in (("fixed", dict(fixed_iter=1, maxiter_x=4, maxiter_bt=8)), ("conv", {})):
            g.set_opts(**opts)
            g.stage_eqcounts([0, n], d["cell_ptr"], d["eq_id"], d["eq_cnt"], d["eq_row"])
            if mode == "conv":
                g.set_gene_map(d["gene_of_col"], len(d["n_members"]))
            g.run()
            rowptr, idx, val = g.fetch_counts()
            y = np.zeros((n, xm.n_rows))
            y[np.repeat(np.arange(n), np.diff(rowptr)), idx] = val
            assert np.array_equal(y, dense(d, "y", (n, xm.n_rows)))             # counts: bit-exact
            iso, want = g.fetch_iso(), dense(d, mode, (n, xm.n_cols))
            iters = g.fetch_iters()[0]
            same = np.all(iters == d[f"{mode}_iters"], axis=1)
            assert same.all() if mode == "fixed" else same.mean() >= 0.995
            cols_ok = np.repeat(np.isin(np.arange(xm.n_blocks), g.em_blocks[same]) | (np.diff(xm.q_off) == 1),
                                np.diff(xm.dm0_p_off))
            err = np.abs(iso - want) / np.maximum(np.abs(want), 1e-6)
            assert err[:, cols_ok].max() < 1e-6                                  # north_star: 1e-6 relative in fp64
            if mode == "conv" and same.all():
                keep = np.unpackbits(d["keep"])[: xm.n_cols].astype(bool)
                assert np.array_equal(g.fetch_colsum() > 0, keep)                # step2:167 row filter
                gene, gwant = g.fetch_gene(), dense(d, "gene", (n, len(d["n_members"])))
                assert np.max(np.abs(gene - gwant) / np.maximum(np.abs(gwant), 1e-6)) < 1e-6
    finally:
        g.close()
