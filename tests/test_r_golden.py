"""Parity against R ITSELF: the dumps tools/make_golden.R writes by running the reference's own
crpcount_td / hamming_correction / estim_chunk / estim.BETA / estim.X.em / scasa_genemat on the committed
48-cell case (tests/golden/r_dumps/).  When the dumps are present the CPU oracle, the string-level literal,
the library's host map and (on a B200) the CUDA path are all compared with them; when they are absent the
R tests skip with a message that says how to produce them — and the comparison code itself still runs on a
stand-in written from the oracle (tests/r_dumps.py), so it cannot rot.

Bounds: counts, row maps, column and gene-row order bit-exact; fixed-iteration estimates 1e-6 relative
(north_star); convergence-mode estimates 1e-6 relative on >= 99 % of the non-zero entries and never more than
2 * maxerr apart in the reference's own measure |a-b| / max(b, lim) (a flipped break, SURVEY.md §7 iii);
gene sums 1e-9 relative where the isoform estimates agree."""
import os

import numpy as np
import pytest

from . import r_dumps as RD

NPZ = os.path.join(RD.HERE, "golden", "xmatrix_hg38_alevin.npz")
NEED_R = pytest.mark.skipif(not RD.have_r_dumps(), reason="no R dumps: run `python tools/export_golden_inputs.py && "
                            "Rscript tools/make_golden.R scasa=<reference>/scasa` (PARITY UNPINNED until then)")


def _gold():
    with np.load(RD.GOLD) as z:
        return {k: z[k] for k in z.files}


def check_against_dumps(d: dict, xm, oracle, gpu=None) -> dict:
    """Everything that can be compared with one set of dumps.  gpu: a scasa_b200.api.Scasa or None."""
    from scasa_b200.api import eqclass_map_host
    from oracle.crpcount_literal import CrpLiteral, hamming_assign
    g = _gold()
    n = int(g["n_cells"])
    rep = {}
    # ---- crpcount_td: Y bit-exact (oracle, and the library's eqclass -> row map + integer sums)
    h = oracle.CrpHandle(xm)
    y = oracle.crpcount_cells(h, g["eq_off"], g["eq_tx"], g["cell_ptr"], g["eq_id"], g["eq_cnt"], nthreads=4)
    assert np.array_equal(y, d["y"]), "oracle crpcount_td differs from R"
    eq_row = eqclass_map_host(xm, g["eq_off"], g["eq_tx"], 4)
    y_lib = np.zeros_like(y)
    cell_of = np.repeat(np.arange(n), np.diff(g["cell_ptr"]))
    ok = eq_row[g["eq_id"]] >= 0
    np.add.at(y_lib, (cell_of[ok], eq_row[g["eq_id"]][ok]), g["eq_cnt"][ok])
    assert np.array_equal(y_lib, d["y"]), "library eqclass map + counts differ from R"
    # ---- column order of estim_chunk's result
    assert d["colnames"] == xm.dm0_colnames(), "column order differs from R's cbind order (Rsource:1042)"
    # ---- estim_chunk, reference convergence options
    est, _ = oracle.estim_chunks(xm, y, [0, n], None, nthreads=4)

    def conv_check(a, b, what):
        nzm = (a != 0) | (b != 0)
        rel = np.abs(a - b)[nzm] / np.maximum(np.abs(b[nzm]), 1e-9)
        frac = float((rel < 1e-6).mean())
        meas = float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 0.01)))
        assert frac >= 0.99, f"{what}: only {frac:.4f} of the non-zero estimates within 1e-6 of R"
        assert meas <= 0.02, f"{what}: |a-b|/max(b, lim) = {meas:.3e} > 2*maxerr against R"
        return frac, meas
    rep["oracle_conv"] = conv_check(est, d["est"], "oracle estim_chunk")
    # ---- fixed-iteration AEM on the non-poor EM blocks
    if "fixed" in d:
        cols = d["fixed_cols"]
        fx, _ = oracle.estim_chunks(xm, y, [0, n], oracle.Opts.make(fixed_iter=1, maxiter_x=4, maxiter_bt=8), nthreads=4)
        err = np.abs(fx[:, cols] - d["fixed"][:, cols]) / np.maximum(np.abs(d["fixed"][:, cols]), 1e-9)
        rep["oracle_fixed"] = float(err.max())
        assert err.max() < 1e-6, f"oracle fixed-iteration AEM: {err.max():.3e} relative against R"
    # ---- row filter, gene layout and gene sums
    if "gene" in d:
        names = xm.dm0_colnames()
        keep = oracle.keep_rows(d["est"])
        assert d["iso_rownames"] == [nm for nm, k in zip(names, keep) if k], "rows kept by step2:167 differ"
        goc, nmem, gnames = oracle.gene_layout([xm.gene_names[i] if i >= 0 else None for i in xm.gene_of_col], keep)
        if d["gene_rownames"] != gnames:                 # R's tapply order is the session collation
            assert sorted(d["gene_rownames"]) == sorted(gnames), "gene sets differ from R"
            perm = [gnames.index(x) for x in d["gene_rownames"]]
            rep["gene_order"] = "same genes, R session collation differs from C"
        else:
            perm = list(range(len(gnames)))
            rep["gene_order"] = "identical"
        gene = oracle.gene_sum(d["est"], goc, nmem)[:, perm]
        gerr = np.abs(gene - d["gene"]) / np.maximum(np.abs(d["gene"]), 1e-9)
        rep["gene"] = float(gerr.max())
        assert gerr.max() < 1e-9, f"gene sums: {gerr.max():.3e} relative against R"
    # ---- hamming_correction, case by case: literal, and the library through single-class cells
    if "hamming" in d:
        lit = CrpLiteral(xm)
        bad = [(k, o, a) for k, o, a in d["hamming"] if hamming_assign([o], lit.block(k))[0] != a]
        assert not bad, (len(bad), bad[:5])
        rep["hamming_cases"] = len(d["hamming"])
    # ---- the CUDA path
    if gpu is not None:
        from .parity import compare_tasks
        gpu.set_opts()
        gpu.stage_eqcounts([0, n], g["cell_ptr"], g["eq_id"], g["eq_cnt"], gpu.eqclass_map(g["eq_off"], g["eq_tx"], 4))
        gpu.run()
        rowptr, idx, val = gpu.fetch_counts()
        yg = np.zeros_like(y)
        yg[np.repeat(np.arange(n), np.diff(rowptr)), idx] = val
        assert np.array_equal(yg, d["y"]), "GPU pack differs from R's crpcount_td"
        rep["gpu_conv"] = conv_check(gpu.fetch_iso(), d["est"], "GPU estim_chunk")
        if "fixed" in d:
            gpu.set_opts(fixed_iter=1, maxiter_x=4, maxiter_bt=8)
            gpu.run()
            cols = d["fixed_cols"]
            err = np.abs(gpu.fetch_iso()[:, cols] - d["fixed"][:, cols]) / np.maximum(np.abs(d["fixed"][:, cols]), 1e-9)
            rep["gpu_fixed"] = float(err.max())
            assert err.max() < 1e-6, f"GPU fixed-iteration AEM: {err.max():.3e} relative against R"
            gpu.set_opts()
        del compare_tasks
    return rep


def test_dump_consumer_on_a_stand_in(xm, oracle, tmp_path):
    """The comparison code above, run on dumps written from the oracle in R's dump format (NOT R output)."""
    RD.write_from_oracle(str(tmp_path), xm, oracle, hamming_stride=211)
    d = RD.load(str(tmp_path), 48, xm.n_rows, xm.n_cols)
    rep = check_against_dumps(d, xm, oracle)
    assert rep["oracle_conv"][0] == 1.0 and rep["hamming_cases"] > 1000 and rep["gene_order"] == "identical"
    # and it does notice a difference
    d["est"][3, np.nonzero(d["est"][3])[0][:40]] *= 1.05
    with pytest.raises(AssertionError):
        check_against_dumps(d, xm, oracle)


@NEED_R
def test_oracle_and_host_logic_against_r(xm, oracle):
    d = RD.load(RD.R_DUMPS, 48, xm.n_rows, xm.n_cols)
    print("against R:", check_against_dumps(d, xm, oracle))


@NEED_R
@pytest.mark.gpu
def test_cuda_path_against_r(xm, oracle):
    from scasa_b200.api import Scasa
    d = RD.load(RD.R_DUMPS, 48, xm.n_rows, xm.n_cols)
    g = Scasa(xm, device=0)
    try:
        print("CUDA path against R:", check_against_dumps(d, xm, oracle, gpu=g))
    finally:
        g.close()
