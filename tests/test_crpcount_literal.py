"""oracle/crpcount_literal.py (string-level transcription of Rsource:416-648) against the C oracle
(oracle/scasa_oracle_counts.c): two independent readings of crpcount_td + hamming_correction must
agree on every perturbed class of the sweep and on whole cells.  CPU only."""
import os

import numpy as np

from .hamming_sweep import literal_rows, sweep_classes

NPZ = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "xmatrix_hg38_alevin.npz")
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "quant_c0_48cells.npz")


def test_r_format_known_outputs():
    """format() results as R prints them (digits = 7): common decimals per vector, fixed unless wider
    than scientific, zeros of an all-zero vector stay "0"."""
    from oracle.crpcount_literal import r_format, r_string_gt_zero
    assert r_format([0.5, 0.25]) == ["0.50", "0.25"]
    assert r_format([0.0, 0.0]) == ["0", "0"]
    assert r_format([0.0, 0.1234567891]) == ["0.0000000", "0.1234568"]
    assert r_format([1e-5]) == ["1e-05"]
    assert r_format([0.5, 1.234567e-5]) == ["5.000000e-01", "1.234567e-05"]
    assert r_format([1.0, 0.5]) == ["1.0", "0.5"]
    assert r_format([1.0 / 3.0]) == ["0.3333333"]
    assert r_format([0.1, 0.12]) == ["0.10", "0.12"]
    assert r_format([0.5, 0.0001234567]) == ["0.5000000000", "0.0001234567"]
    assert r_format([100000.0, 1.0]) == ["1e+05", "1e+00"]
    assert r_format([123456.0]) == ["123456"]
    # `x > 0` on the character row: "0" is not, "0.0000000" and any pattern with a 1 are
    assert not r_string_gt_zero("0")
    assert r_string_gt_zero("0.0000000") and r_string_gt_zero("0110") and r_string_gt_zero("1e-05")


def test_hamming_literal_hand_case():
    """The tie-break by hand: rows 110 / 011 / 101 with values, observed 111 is at distance 1 of all
    three; medians over (pattern as a number, formatted values): row 110 -> median(110, .5, .2, 0) etc."""
    from oracle.crpcount_literal import Block, hamming_assign
    b = Block(["110", "011", "101"], ["a", "b", "c"], [[0.5, 0.2, 0.0], [0.0, 0.8, 0.3], [0.5, 0.0, 0.7]])
    # character rows: ("110","0.5","0.2","0.0") -> median(110,.5,.2,0) = .35 ; ("011","0.0","0.8","0.3") -> median(11,0,.8,.3)=.55 ;
    # ("101","0.5","0.0","0.7") -> median(101,.5,0,.7) = .6  => row 101
    assert hamming_assign(["111"], b) == ["101"]
    assert hamming_assign(["110"], b) == ["110"]                      # exact
    # 100: distance 1 to 110 and 101 only -> median rule between them -> .35 vs: columns formatted over the two candidate rows:
    # col b = (0.2, 0.0) -> "0.2","0.0"; col c = (0.0, 0.7) -> "0.0","0.7": 110 -> median(110,.5,.2,0)=.35, 101 -> median(101,.5,0,.7)=.6
    assert hamming_assign(["100"], b) == ["101"]
    # a column that is zero in all candidates prints as "0" and drops out of the median
    b2 = Block(["10", "01", "11"], ["a", "b"], [[0.9, 0.0], [0.0, 0.0], [0.1, 1.0]])
    # observed 00 (cannot occur in crpcount_td, fine here): distance 1 to 10 and 01 -> candidates rows 0,1:
    # col a = ("0.9","0.0"), col b = ("0","0") -> row 10: median(10, .9) = 5.45 ; row 01: median(1, 0) = .5 -> "10"
    assert hamming_assign(["00"], b2) == ["10"]


def test_sweep_literal_equals_c_oracle(xm, oracle):
    """Every 7th class of the full Hamming-1/-2 sweep over all shipped EM blocks (~58k classes, most of
    them decided by the median rule): the string-level literal and the C oracle pick the same row."""
    eq_off, eq_tx, _ = sweep_classes(xm, stride=7)
    lit, ties = literal_rows(NPZ, eq_off, eq_tx)
    assert ties > 40000                                   # the tie-break is what is being exercised
    h = oracle.CrpHandle(xm)
    n = len(eq_off) - 1
    got = np.empty(n, np.int32)
    for a in range(0, n, 20000):                          # eqclass_rows is dense per class: bound its memory
        b = min(n, a + 20000)
        got[a:b] = oracle.eqclass_rows(h, eq_off[a:b + 1] - eq_off[a], eq_tx[eq_off[a]:eq_off[b]], nthreads=8)
    bad = np.nonzero(lit != got)[0]
    assert len(bad) == 0, (len(bad), bad[:5], lit[bad[:5]], got[bad[:5]])


def test_whole_cells_literal_equals_golden(xm):
    """crpcount_td over whole cells of the committed 48-cell case (~2,150 classes each): the literal
    reproduces the golden Y of all 48 cells bit for bit (block selection by Jaccard score, size-1
    classes into 1x1 blocks, sums by pattern)."""
    from oracle.crpcount_literal import CrpLiteral
    with np.load(GOLD) as z:
        d = {k: z[k] for k in z.files}                    # an NpzFile re-reads the array on every access
    lit = CrpLiteral(xm)
    n = int(d["n_cells"])
    y = np.zeros((n, xm.n_rows))
    y[d["y_r"], d["y_c"]] = d["y_v"]
    for cell in range(n):
        rows = []
        for k in range(d["cell_ptr"][cell], d["cell_ptr"][cell + 1]):
            e = int(d["eq_id"][k])
            for t in d["eq_tx"][d["eq_off"][e]:d["eq_off"][e + 1]]:
                rows.append((xm.tx_names[t], int(d["eq_cnt"][k]), e))
        res = lit.crpcount_td(rows)
        got = np.zeros(xm.n_rows)
        for k, v in res.items():
            got[xm.q_off[k]:xm.q_off[k + 1]] = v
        assert np.array_equal(got, y[cell]), cell


def test_sweep_library_host_map_equals_literal(xm):
    """The PRODUCT's eqclass -> row decision (CrpIndex in scasa_host.cpp, through the host-only C-ABI entry
    scasa_eqclass_map_host: no GPU involved) on EVERY class of the sweep (~400k, ~85 % decided by the median tie-break) against the literal."""
    from scasa_b200.api import eqclass_map_host
    eq_off, eq_tx, blocks = sweep_classes(xm, stride=1)
    want, ties = literal_rows(NPZ, eq_off, eq_tx)
    assert len(blocks) > 350000 and ties > 250000
    got = eqclass_map_host(xm, eq_off, eq_tx, 8)
    bad = np.nonzero(got != want)[0]
    assert len(bad) == 0, (len(bad), bad[:5], got[bad[:5]], want[bad[:5]], blocks[bad[:5]])
