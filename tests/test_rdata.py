"""The .RData reader on a stream written here byte by byte (R's XDR serialisation, version 3)."""
import gzip
import struct

import numpy as np


def _i(v):
    return struct.pack(">i", v)


def _chars(s, flags=0x00040009):      # CHARSXP, UTF8 bit in gp
    b = s.encode()
    return _i(flags) + _i(len(b)) + b


def _sym(name):
    return _i(1) + _chars(name)


def _strvec(v):
    return _i(16) + _i(len(v)) + b"".join(_chars(s) for s in v)


def _intvec(v):
    return _i(13) + _i(len(v)) + b"".join(_i(x) for x in v)


def _pairlist(items, nil=True):
    out = b""
    for tag, val in items:
        out += _i(2 | (1 << 10)) + tag + val
    return out + (_i(254) if nil else b"")


def test_reads_a_matrix_with_dimnames_and_altrep(tmp_path):
    from scasa_b200.rdata import load_rdata
    vals = [0.6, 0.4, 0.0, 1.0]
    dimnames = _i(19) + _i(2) + _strvec(["10", "11"]) + _strvec(["NM_1", "NM_2"])
    attrs = _pairlist([(_sym("dim"), _intvec([2, 2])), (_sym("dimnames"), dimnames)])
    mat = _i(14 | (1 << 9)) + _i(4) + b"".join(struct.pack(">d", x) for x in vals) + attrs
    # a compact integer sequence 1:5 as ALTREP: info = (class sym, package sym, type), state = c(n, start, step)
    info = _pairlist([(_sym("a"), _sym("compact_intseq")), (_sym("b"), _sym("base")), (_sym("c"), _intvec([13]))])
    state = _i(14) + _i(3) + b"".join(struct.pack(">d", x) for x in (5.0, 1.0, 1.0))
    alt = _i(238) + info + state + _i(254)
    top = _pairlist([(_sym("X"), mat), (_sym("seq"), alt), (_i(0x2FF), _strvec(["again"]))])  # 0x2FF: REFSXP back-reference to symbol 2 ("dim")
    enc = b"UTF-8"
    payload = b"RDX3\nX\n" + _i(3) + _i(0x030600) + _i(0x030500) + _i(len(enc)) + enc + top
    path = tmp_path / "t.RData"
    path.write_bytes(gzip.compress(payload))
    ws = load_rdata(str(path))
    m = ws["X"]
    assert m.dim == (2, 2) and m.dimnames == [["10", "11"], ["NM_1", "NM_2"]]
    assert np.array_equal(m.as_matrix(), [[0.6, 0.0], [0.4, 1.0]])       # column-major payload
    assert list(ws["seq"].value) == [1, 2, 3, 4, 5]
    assert ws["dim"].value == ["again"]                                  # tag resolved through the reference table


def test_bustools_xmatrix_loads_through_the_same_reader():
    """N4 (SURVEY.md §8 f): the kallisto/bustools X-matrix is the same RDX3 layout; reader and packing are
    generic.  Sizes are the ones the survey measured (29,088 blocks, 4,563 EM blocks, sum q_em 18,507,
    sum p_em 9,475, largest block 57 x 21).  Runs only where the reference tree is present."""
    import os
    import pytest
    ref = "/root/reference/scasa/REFERENCE"
    if not os.path.exists(os.path.join(ref, "Xmatrix_hg38_bustools.RData")):
        pytest.skip("reference tree not present (GPU box)")
    from scasa_b200.xmatrix import XMatrix
    xm = XMatrix.from_rdata(os.path.join(ref, "Xmatrix_hg38_bustools.RData"),
                            os.path.join(ref, "isoform_groups_UCSC_hg38_bustools.RData"))
    q, p = np.diff(xm.q_off), np.diff(xm.dm0_p_off)
    em = q > 1
    assert (xm.n_blocks, int(em.sum()), int(q[em].sum()), int(p[em].sum())) == (29088, 4563, 18507, 9475)
    assert (int(q.max()), int(p.max())) == (57, 21)
    cs = np.concatenate([xm.dm0_block(k).sum(0) for k in range(0, xm.n_blocks, 7)])
    assert np.allclose(cs, 1.0, atol=1e-12)                              # DM0 columns are stochastic (K3/K4)
    assert xm.gene_of_col is not None and (np.asarray(xm.gene_of_col) >= 0).all()   # every column has a gene (K7)
    # row patterns: 0/1, as many columns as the CRP block (K2)
    for k in (0, 100, 20000):
        pat = xm.crp_patterns(k)
        assert pat.shape == (xm.q(k), xm.p_crp(k)) and set(np.unique(pat)) <= {0, 1}
