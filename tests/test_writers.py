"""N2 (SURVEY.md §8 f): write.table-compatible text output (step2:170, step3:52)."""
import math
import os

import numpy as np
import pytest

# R session facts (R >= 3.x, write.table / format(x, digits = 15) on single values)
R_KNOWN = [
    (1.0, "1"), (1.5, "1.5"), (0.1, "0.1"), (1 / 3, "0.333333333333333"), (2 / 3, "0.666666666666667"),
    (1e5, "1e+05"), (123456.0, "123456"), (1e-4, "1e-04"), (0.001, "0.001"), (0.00012345, "0.00012345"),
    (1e15, "1e+15"), (123456789012.0, "123456789012"), (100.0, "100"), (1e-15, "1e-15"), (0.5, "0.5"),
    (-2.5, "-2.5"), (math.pi, "3.14159265358979"), (math.e, "2.71828182845905"), (1e300, "1e+300"),
    (0.1 + 0.2, "0.3"), (100000.5, "100000.5"), (1234567.1, "1234567.1"), (3e-10, "3e-10"),
    (1.25e-7, "1.25e-07"), (-1e-300, "-1e-300"), (2.0 ** 52, "4503599627370496"), (0.0, "0"), (-0.0, "0"),
    (float("inf"), "Inf"), (float("-inf"), "-Inf"), (float("nan"), "NaN"), (17.0, "17"), (99999.0, "99999"),
    (1e22, "1e+22"), (123456.7890123, "123456.7890123"),
]


def py_format15(x: float) -> str:
    """Independent restatement with correctly rounded decimal conversion (Python's %e) instead of
    R's long double arithmetic: formatReal for one value with digits = 15."""
    if x == 0:
        return "0"
    if math.isnan(x):
        return "NaN"
    if math.isinf(x):
        return "Inf" if x > 0 else "-Inf"
    neg = 1 if x < 0 else 0
    r = abs(x)
    mant, exp = ("%.14e" % r).split("e")
    kp = int(exp)
    digits = mant.replace(".", "").rstrip("0")
    nsig = max(len(digits), 1)
    widens = 0 < kp <= 22 and r < float(10 ** kp)
    left = kp + 1 - (1 if widens else 0)
    sleft = neg + (1 if left <= 0 else left)
    rgt = min(max(nsig - left, 0), 22)
    wF = sleft + rgt + (1 if rgt else 0)
    e = 2 if (kp >= 100 or kp <= -99) else 1
    d = nsig - 1
    wE = neg + (1 if d > 0 else 0) + d + 4 + e
    if wF <= wE:
        return "%.*f" % (rgt, x)
    return "%.*e" % (d, x)


def test_format_real15_known_r_outputs():
    from scasa_b200.writers import format_real15
    for x, want in R_KNOWN:
        assert format_real15(x) == want, (x, format_real15(x), want)
        assert py_format15(x) == want, (x, py_format15(x), want)


def test_format_real15_random_values_agree_with_independent_restatement():
    from scasa_b200.writers import format_real15
    rng = np.random.default_rng(3)
    xs = np.concatenate([
        rng.normal(size=4000) * 10.0 ** rng.integers(-12, 13, 4000),
        rng.integers(1, 10 ** 9, 2000).astype(float),
        rng.integers(1, 2000, 2000) / 8.0,
        rng.gamma(0.5, 3.0, 3000),
        np.round(rng.gamma(2.0, 40.0, 2000), 3),
        10.0 ** rng.integers(-30, 30, 200),
        rng.integers(1, 1000, 3000).astype(float) * 10.0 ** rng.integers(0, 13, 3000),      # whole numbers with trailing zeros
        np.arange(1.0, 1200.0),                                                            # the counts themselves
    ])
    for x in xs:
        a = format_real15(float(x))
        assert a == py_format15(float(x)), (repr(float(x)), a, py_format15(float(x)))
        if np.isfinite(x) and x != 0:
            assert abs(float(a) - x) <= 5e-15 * abs(x)          # 15 significant digits are kept


@pytest.mark.parametrize("threads", [1, 3])
def test_write_table_layout(tmp_path, threads):
    from scasa_b200.writers import write_table
    rng = np.random.default_rng(5)
    n_cells, n_cols = 9, 75                                    # > 2 blocks of 32 lines
    mat = rng.gamma(0.4, 5.0, (n_cells, n_cols)) * (rng.random((n_cells, n_cols)) < 0.3)
    mat[2, 5] = 1e-7
    mat[3, 6] = 123456789.25
    keep = mat.sum(0) > 0
    keep[11] = False
    rows = [f"NM_{i} NM_{i + 1}" if i % 7 == 0 else f"NM_{i}" for i in range(n_cols)]     # merged names contain a space
    cells = [f"ACGT{i:04d}" for i in range(n_cells)]
    path = str(tmp_path / "t.txt")
    write_table(path, mat, rows, cells, keep=keep, n_threads=threads)
    got = open(path, "rb").read().decode()
    want = ["\t".join(cells)]
    for j in range(n_cols):
        if keep[j]:
            want.append("\t".join([rows[j]] + [py_format15(float(v)) for v in mat[:, j]]))
    assert got == "\n".join(want) + "\n"
    write_table(path, mat, rows, cells, n_threads=threads)      # no filter: every column
    assert len(open(path).read().split("\n")) == n_cols + 2


def test_write_table_threads_never_reorder_or_drop_blocks(tmp_path):
    """Many small blocks, more threads than cores, repeated: the block in line is written by exactly one
    thread (a second writer entering while the first is inside fwrite used to skip a block)."""
    from scasa_b200.writers import write_table
    rng = np.random.default_rng(8)
    n_cells, n_cols = 3, 40 * 32
    mat = np.round(rng.gamma(0.4, 5.0, (n_cells, n_cols)), 2)
    rows = [f"T{i}" for i in range(n_cols)]
    cells = ["A", "C", "G"]
    path = str(tmp_path / "t.txt")
    write_table(path, mat, rows, cells, n_threads=1)
    want = open(path, "rb").read()
    assert want.count(b"\n") == n_cols + 1
    for it in range(150):
        write_table(path, mat, rows, cells, n_threads=2 + it % 7)
        assert open(path, "rb").read() == want, it


# ---------------------------------------------------------------------------------------------- sparse path
def _random_sparse(rng, n_cells, n_cols, density):
    dense = rng.gamma(0.3, 40.0, (n_cells, n_cols)) * (rng.random((n_cells, n_cols)) < density)
    dense[rng.random((n_cells, n_cols)) < 0.01] = 1e-7 * rng.random()          # small values: scientific notation
    dense[:, 3] = 0.0                                                            # an all-zero column
    mask = (dense != 0) | (rng.random((n_cells, n_cols)) < 0.02)                 # stored zeros too
    rowptr = np.concatenate([[0], np.cumsum(mask.sum(axis=1))]).astype(np.int64)
    cells, cols = np.nonzero(mask)
    return dense, rowptr, cols.astype(np.int32), dense[cells, cols]


def test_csr_transpose_is_the_stable_counting_sort():
    from scasa_b200.writers import csr_transpose
    rng = np.random.default_rng(3)
    for n_cells, n_cols, nthr in ((1, 5, 1), (37, 11, 3), (900, 260, 8), (5000, 40, 16)):
        dense, rowptr, cols, vals = _random_sparse(rng, n_cells, n_cols, 0.15)
        colptr, rows, tv = csr_transpose(rowptr, cols, vals, n_cols, nthr)
        cell = np.repeat(np.arange(n_cells), np.diff(rowptr))
        order = np.argsort(cols, kind="stable")                                  # column-major, cells ascending inside a column
        assert np.array_equal(colptr, np.concatenate([[0], np.cumsum(np.bincount(cols, minlength=n_cols))]))
        assert np.array_equal(rows, cell[order]) and np.array_equal(tv, vals[order])
    # empty matrix
    colptr, rows, tv = csr_transpose(np.zeros(4, np.int64), np.zeros(0, np.int32), np.zeros(0), 6, 2)
    assert np.array_equal(colptr, np.zeros(7)) and len(rows) == 0


def test_sparse_writer_writes_the_bytes_of_the_dense_writer(tmp_path):
    """scasa_write_table_sparse from the transposed sparse rows == scasa_write_table from the dense matrix,
    byte for byte: kept rows, a row order that is a permutation (the gene layout), stored zeros, empty cells."""
    from scasa_b200.writers import csr_transpose, write_table, write_table_sparse
    rng = np.random.default_rng(5)
    n_cells, n_cols = 333, 190
    dense, rowptr, cols, vals = _random_sparse(rng, n_cells, n_cols, 0.08)
    names = [f"NM_{j:06d} NM_{j + 7}" if j % 9 == 0 else f"NR_{j}" for j in range(n_cols)]
    cells = [f"CELL{c:05d}" for c in range(n_cells)]
    colptr, rows, tv = csr_transpose(rowptr, cols, vals, n_cols, 4)
    keep = dense.sum(axis=0) > 0
    a, b = str(tmp_path / "dense.txt"), str(tmp_path / "sparse.txt")
    write_table(a, dense, names, cells, keep=keep, n_threads=3)
    src = np.flatnonzero(keep)
    size = write_table_sparse(b, n_cells, colptr, rows, tv, src, [names[j] for j in src], cells, n_threads=5)
    assert open(a, "rb").read() == open(b, "rb").read() and size == os.path.getsize(b)
    perm = rng.permutation(src)[:77]                                              # any order / subset of source columns
    write_table(a, np.ascontiguousarray(dense[:, perm]), [names[j] for j in perm], cells, n_threads=2)
    write_table_sparse(b, n_cells, colptr, rows, tv, perm, [names[j] for j in perm], cells, n_threads=2)
    assert open(a, "rb").read() == open(b, "rb").read()


def test_gene_layout_is_scasa_genemats(oracle, xm):
    """scasa_b200.writers.gene_layout against the oracle's restatement of step3:34-43 on the shipped gene map, with
    a filter that removes isoforms (genes disappear, multi-isoform genes become single-isoform ones)."""
    from scasa_b200.writers import gene_layout
    rng = np.random.default_rng(8)
    keep = rng.random(xm.n_cols) < 0.6
    names_of_col = [xm.gene_names[g] for g in xm.gene_of_col]
    want_goc, want_nmem, want_names = oracle.gene_layout(names_of_col, keep)
    order, cnt = gene_layout(xm.gene_of_col, xm.gene_names, keep)
    assert [xm.gene_names[g] for g in order] == want_names
    assert np.array_equal(cnt[order], want_nmem)
    # every kept column lands in the row the oracle assigns
    pos = np.full(len(xm.gene_names), -1)
    pos[order] = np.arange(len(order))
    assert np.array_equal(np.where(keep, pos[xm.gene_of_col], -1), want_goc)
    moved = (np.bincount(xm.gene_of_col, minlength=len(cnt)) > 1) & (cnt == 1)
    assert moved.sum() > 100                                                      # multis that became singles are exercised
