"""N1 (SURVEY.md §8 f): the bfh.txt reader against the literal restatement of step1_1 + step2:59-60."""
import os

import numpy as np
import pytest


def _write_bfh(path, rng, n_tx=40, n_cb=30, n_eq=120, crlf=False, trailing_newline=True):
    tx = [f"NM_{1000 + i}.{i % 3}" for i in range(n_tx)]
    cb = ["".join(rng.choice(list("ACGT"), 16)) for _ in range(n_cb)]
    cb[3] = cb[7]                                  # a repeated sequence: one cell in R (unique(Barcode))
    lines = [str(n_tx), str(n_cb), str(n_eq)] + tx + cb
    for _ in range(n_eq):
        k = int(rng.integers(1, 6))
        tids = sorted(rng.choice(n_tx, k, replace=False).tolist())
        nb = int(rng.integers(1, 5))
        bcs = rng.choice(n_cb - 2, nb, replace=False).tolist()      # the last two barcodes never occur
        f = [str(k)] + [str(t) for t in tids]
        body, total = [], 0
        for b in bcs:
            numi = int(rng.integers(1, 4))
            body += [str(b), str(numi)]
            for _u in range(numi):
                c = int(rng.integers(1, 9))
                total += c
                body += ["".join(rng.choice(list("ACGT"), 10)), str(c)]
        f += [str(total), str(nb)] + body
        lines.append("\t".join(f))
    nl = "\r\n" if crlf else "\n"
    with open(path, "w", newline="") as fh:
        fh.write(nl.join(lines) + (nl if trailing_newline else ""))
    return tx, cb


@pytest.mark.parametrize("crlf,trailing", [(False, True), (True, True), (False, False)])
def test_bfh_reader_equals_literal_restatement(tmp_path, crlf, trailing):
    from scasa_b200.bfh import read_bfh
    from oracle import bfh_literal
    rng = np.random.default_rng(11)
    path = str(tmp_path / "bfh.txt")
    _write_bfh(path, rng, crlf=crlf, trailing_newline=trailing)
    tx_all, cells, per_cell, eq_tx = bfh_literal.quant_inputs(path)
    for threads in (1, 3):
        b = read_bfh(path, threads)
        assert b.tx_names == tx_all
        assert b.barcodes == cells
        assert b.barcodes == sorted(set(b.barcodes), key=str.encode)
        assert len(b.cell_ptr) == len(cells) + 1 and b.cell_ptr[-1] == len(b.eq_id)
        for c, bc in enumerate(cells):
            ids = b.eq_id[b.cell_ptr[c]:b.cell_ptr[c + 1]] + 1          # R's eqClass is the 1-based line
            cnt = b.eq_count[b.cell_ptr[c]:b.cell_ptr[c + 1]]
            want = sorted(per_cell[bc].items())
            assert list(ids) == [e for e, _ in want]
            assert list(cnt) == [v for _, v in want]
        for e, txs in eq_tx.items():
            got = [b.tx_names[t] for t in b.eq_tx[b.eq_off[e - 1]:b.eq_off[e]]]
            assert got == txs


def test_bfh_reader_does_not_depend_on_the_thread_count(tmp_path):
    """More eqclasses, lines of very different lengths (the ranges are cut by bytes), a file read in slices: the
    arrays are the same for 1, 2, 5 and 16 threads, and eqclass ids ascend inside every cell (the rbind order)."""
    from scasa_b200.bfh import read_bfh
    rng = np.random.default_rng(5)
    path = str(tmp_path / "bfh.txt")
    n_tx, n_cb, n_eq = 300, 400, 3000
    tx = [f"NM_{i}" for i in range(n_tx)]
    cb = ["".join(rng.choice(list("ACGT"), 16)) for _ in range(n_cb)]
    lines = [str(n_tx), str(n_cb), str(n_eq)] + tx + cb
    for e in range(n_eq):
        k = int(rng.integers(1, 8))
        tids = sorted(rng.choice(n_tx, k, replace=False).tolist())
        nb = int(rng.integers(200, 390)) if e % 97 == 0 else int(rng.integers(1, 6))      # a few classes hold most cells
        body, total = [], 0
        for b in rng.choice(n_cb, nb, replace=False).tolist():
            numi = int(rng.integers(1, 4))
            body += [str(b), str(numi)]
            for _u in range(numi):
                c = int(rng.integers(1, 9))
                total += c
                body += ["".join(rng.choice(list("ACGT"), 10)), str(c)]
        lines.append("\t".join([str(k)] + [str(t) for t in tids] + [str(total), str(nb)] + body))
    with open(path, "w", newline="") as fh:
        fh.write("\n".join(lines) + "\n")
    a = read_bfh(path, 1)
    assert len(a.eq_off) == n_eq + 1 and a.cell_ptr[-1] == len(a.eq_id) > 15000
    for c in range(len(a.barcodes)):
        assert np.all(np.diff(a.eq_id[a.cell_ptr[c]:a.cell_ptr[c + 1]]) > 0)
    for threads in (2, 5, 16):
        b = read_bfh(path, threads)
        assert a.barcodes == b.barcodes and a.tx_names == b.tx_names
        for k in ("eq_off", "eq_tx", "cell_ptr", "eq_id", "eq_count"):
            assert np.array_equal(getattr(a, k), getattr(b, k)), (threads, k)


def test_bfh_errors(tmp_path):
    from scasa_b200 import lib
    from scasa_b200.bfh import read_bfh
    p = str(tmp_path / "bad.txt")
    open(p, "w").write("2\n1\n1\nA\nB\nACGT\n1\t5\t3\t1\t0\t1\tACGT\t3\n")       # transcript index 5 of 2
    with pytest.raises(lib.ScasaError):
        read_bfh(p)
    with pytest.raises(lib.ScasaError):
        read_bfh(str(tmp_path / "missing.txt"))
    open(p, "w").write("2\n1\n3\nA\nB\nACGT\n")                                  # shorter than the header says
    with pytest.raises(lib.ScasaError):
        read_bfh(p)


def _write_bus(tmp_path, rng, n_tx=30, n_ec=60, n_cells=12, sort=True):
    tx = [f"ENST{i:05d}.{i % 4}" for i in range(n_tx)]
    (tmp_path / "transcripts.txt").write_text("\n".join(tx) + "\n")
    ecs = []
    for e in range(n_ec):
        k = 1 if e < n_tx // 2 else int(rng.integers(2, 6))
        ecs.append(sorted(rng.choice(n_tx, k, replace=False).tolist()))
    (tmp_path / "matrix.ec").write_text("".join(f"{e}\t{','.join(map(str, t))}\n" for e, t in enumerate(ecs)))
    recs = []
    for c in range(n_cells):
        bc = "".join(rng.choice(list("ACGT"), 16))
        n_mol = int(rng.integers(3, 40))                          # some cells fall under the library-size threshold
        for _m in range(n_mol):
            umi = "".join(rng.choice(list("ACGT"), 6))            # short UMIs: collisions happen
            kind = rng.random()
            if kind < 0.6:
                recs.append((bc, umi, int(rng.integers(0, n_ec)), int(rng.integers(1, 5))))
            elif kind < 0.8:                                      # several records, unique largest count
                es = rng.choice(n_ec, 3, replace=False)
                for i, e in enumerate(es):
                    recs.append((bc, umi, int(e), 5 - i))
            else:                                                 # tie on the count: the eqclass size decides, or nothing does
                es = rng.choice(n_ec, 3, replace=False)
                for e in es:
                    recs.append((bc, umi, int(e), 2))
    recs = sorted(set(recs)) if sort else [recs[i] for i in rng.permutation(len(recs))]
    (tmp_path / "bus.txt").write_text("".join(f"{b}\t{u}\t{e}\t{c}\t{b}{u}\n" for b, u, e, c in recs))
    return str(tmp_path / "bus.txt"), str(tmp_path / "matrix.ec"), str(tmp_path / "transcripts.txt")


@pytest.mark.parametrize("sort", [True, False])
def test_bustools_front_end_equals_literal_restatement(tmp_path, sort):
    """N4: step1_2's library-size filter, duplicated (barcode, UMI) rules and per-cell eqclass counts."""
    from scasa_b200.bfh import read_bus
    from oracle import bus_literal
    rng = np.random.default_rng(23)
    paths = _write_bus(tmp_path, rng, sort=sort)
    for thres in (0, 10, 25):
        ref, cells, per_cell, eq_tx = bus_literal.quant_inputs(*paths, thres)
        b = read_bus(*paths, libsize_thres=thres)
        assert b.tx_names == ref
        assert b.barcodes == cells, thres
        for c, bc in enumerate(cells):
            ids = b.eq_id[b.cell_ptr[c]:b.cell_ptr[c + 1]]
            cnt = b.eq_count[b.cell_ptr[c]:b.cell_ptr[c + 1]]
            want = sorted(per_cell[bc].items())
            assert list(ids) == [e for e, _ in want] and list(cnt) == [v for _, v in want], (thres, bc)
        for e, txs in eq_tx.items():
            assert [b.tx_names[t] for t in b.eq_tx[b.eq_off[e]:b.eq_off[e + 1]]] == txs
    assert len(read_bus(*paths, libsize_thres=10 ** 6).barcodes) == 0
