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
    (SURVEY.md §7 iii), so: iteration counts agree on >= 99.5 % of (chunk, block) tasks and,
    where they agree, estimates agree to 1e-6."""
    from scasa_b200.api import chunk_split
    from scasa_b200.synth import csr_to_dense
    rowptr, rows, vals = c1
    chunks = chunk_split(200, 4)
    gpu.set_opts()
    iso, iters = gpu.estimate(chunks, rowptr, rows, vals)
    ref, ref_it = oracle.estim_chunks(xm, csr_to_dense(rowptr, rows, vals, xm.n_rows), chunks, None, nthreads=4)
    ref_it = ref_it[:, gpu.em_blocks, :]
    same = np.all(iters == ref_it, axis=2)
    assert same.mean() >= 0.995, same.mean()
    p_off = xm.dm0_p_off
    poor = gpu.poor_blocks()
    worst = 0.0
    for h in range(2):
        cs = slice(chunks[h], chunks[h + 1])
        for e, k in enumerate(gpu.em_blocks):
            if same[h, e] and not poor[k]:
                a, b = iso[cs, p_off[k]:p_off[k + 1]], ref[cs, p_off[k]:p_off[k + 1]]
                worst = max(worst, rel_err(a, b))
    assert worst < RTOL, worst
    assert np.all(np.isfinite(iso))


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
            same = np.all(iters == ref_it, axis=2)
            if fixed:
                assert same.all()
            assert same.mean() > 0.9
            for h in range(len(chunk_off) - 1):
                cs = slice(chunk_off[h], chunk_off[h + 1])
                for e, k in enumerate(g.em_blocks):
                    if same[h, e]:
                        a, b = iso[cs, p_off[k]:p_off[k + 1]], ref[cs, p_off[k]:p_off[k + 1]]
                        assert np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-6)) < RTOL, (fixed, h, k)
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
