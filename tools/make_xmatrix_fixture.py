#!/usr/bin/env python
"""Pack the reference's shipped X-matrix + gene map into tests/golden/*.npz.

Run in the build container (needs /root/reference, which the GPU box does not have):

    python tools/make_xmatrix_fixture.py [alevin|bustools]

The output is the flat layout of scasa_b200/xmatrix.py (numbers and names only,
no reference source).  While packing it re-checks the relations SURVEY.md §4
lists as K1-K8, so a reader bug cannot silently produce a wrong fixture.
"""
import sys
import numpy as np

sys.path.insert(0, ".")
from scasa_b200.rdata import load_rdata  # noqa: E402
from scasa_b200.xmatrix import XMatrix  # noqa: E402

REF = "/root/reference/scasa/REFERENCE"


def main(mapper: str = "alevin") -> None:
    xp = f"{REF}/Xmatrix_hg38_{mapper}.RData"
    gp = f"{REF}/isoform_groups_UCSC_hg38_{mapper}.RData"
    xm = XMatrix.from_rdata(xp, gp)
    ws = load_rdata(xp)
    # K1: names(CRP)[k] == paste(colnames(CRP[[k]]), collapse=" ")
    assert xm.block_names() == list(ws["CRP"].names)
    # K4: CCRP column names == members joined in CRP column order
    want = [n for m in ws["CCRP"].value for n in m.dimnames[1]]
    assert xm.dm0_colnames() == want
    # K7: every DM0 column has a gene
    assert (xm.gene_of_col >= 0).all()
    out = f"tests/golden/xmatrix_hg38_{mapper}.npz"
    xm.save_npz(out)
    back = XMatrix.load_npz(out)
    for k in XMatrix._ARRAYS:
        assert np.array_equal(getattr(xm, k), getattr(back, k)), k
    assert back.tx_names == xm.tx_names and back.gene_names == xm.gene_names
    print(out, "blocks", xm.n_blocks, "rows", xm.n_rows, "crp cols", int(xm.crp_p_off[-1]),
          "dm0 cols", xm.n_cols, "em blocks", len(xm.em_blocks()), "genes", len(xm.gene_names))


if __name__ == "__main__":
    main(*sys.argv[1:])
