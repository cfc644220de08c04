#!/usr/bin/env python
"""Inputs of tools/make_golden.R as plain text, from the committed fixtures.

    python tools/export_golden_inputs.py [outdir = tests/golden/r_case] [hamming_stride = 1]

Writes
  sample_eqclass.tsv   the committed 48-cell case (tests/golden/quant_c0_48cells.npz) as the table
                       scasa_quant_step1_1_v1.0.0.R builds: Barcode, Transcript, Count, eqClass — one
                       line per (cell, eqclass, transcript); cells are 16-mer barcodes in sorted order
  hamming_cases.tsv    block (1-based position in CRP) and observed 0/1 pattern for the sweep of
                       tests/hamming_sweep.py: every EM block x all Hamming-1 / Hamming-2 perturbations
Nothing here is read by the library; the files are not committed (they are derived from the fixtures).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from scasa_b200.xmatrix import XMatrix  # noqa: E402
from tests.hamming_sweep import sweep_classes  # noqa: E402


def barcode(i: int) -> str:
    """16-mer over ACGT whose byte order is the order of i (data.table::setorder is C-locale)."""
    return "".join("ACGT"[(i >> (2 * (15 - k))) & 3] for k in range(16))


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tests", "golden", "r_case")
    stride = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    os.makedirs(out, exist_ok=True)
    xm = XMatrix.load_npz(os.path.join(ROOT, "tests", "golden", "xmatrix_hg38_alevin.npz"))
    with np.load(os.path.join(ROOT, "tests", "golden", "quant_c0_48cells.npz")) as z:
        d = {k: z[k] for k in z.files}
    n = int(d["n_cells"])
    with open(os.path.join(out, "sample_eqclass.tsv"), "w") as f:
        f.write("Barcode\tTranscript\tCount\teqClass\n")
        for c in range(n):
            bc = barcode(c * 7919 + 13)
            for k in range(int(d["cell_ptr"][c]), int(d["cell_ptr"][c + 1])):
                e = int(d["eq_id"][k])
                for t in d["eq_tx"][d["eq_off"][e]:d["eq_off"][e + 1]]:
                    f.write(f"{bc}\t{xm.tx_names[t]}\t{int(d['eq_cnt'][k])}\t{e + 1}\n")
    eq_off, eq_tx, blocks = sweep_classes(xm, stride=stride)
    with open(os.path.join(out, "hamming_cases.tsv"), "w") as f:
        f.write("block\tpat\n")
        for i, k in enumerate(blocks):
            cols = xm.crp_tx[xm.crp_p_off[k]:xm.crp_p_off[k + 1]]
            have = set(int(t) for t in eq_tx[eq_off[i]:eq_off[i + 1]])
            f.write(f"{int(k) + 1}\t{''.join('1' if int(t) in have else '0' for t in cols)}\n")
    print(out, "cells", n, "hamming cases", len(blocks))


if __name__ == "__main__":
    main()
