"""TEST INFRASTRUCTURE — literal Python restatement of the reference's bfh.txt reader.

Follows /root/reference/scasa/SCRIPTS/scasa_quant_step1_1_v1.0.0.R line by line (:16-68: readLines,
the three counts, transcript and barcode lists, per eqclass strsplit / grep("[AGTC]") / the pairs of
(barcode index, count)) and builds the same exploded table the script rbind()s — one row per
(barcode, transcript, count, eqclass) — then applies scasa_quant_step2_v1.0.0.R:59-60
(setorder by Barcode in C-locale byte order, allCells = unique).  Pure Python loops: small files only.
Only tests/ may import this.
"""
import re


def read_bfh_table(path):
    """The data.table of step1_1:37-66 as a list of rows (Barcode, Transcript, Count, eqClass)."""
    with open(path, "rb") as f:
        dat = f.read().decode().split("\n")
    if dat and dat[-1] == "":
        dat.pop()                                              # readLines: a final newline ends the last line
    dat = [x[:-1] if x.endswith("\r") else x for x in dat]
    s = [int(x) for x in dat[0:3]]; dat = dat[3:]              # :18
    tx_all = dat[0:s[0]]; dat = dat[s[0]:]                     # :19
    cb_all = dat[0:s[1]]; dat = dat[s[1]:]                     # :20
    rows = []
    for i in range(1, s[2] + 1):                               # :37
        eq = dat[i - 1].split("\t")                            # :41
        while eq and eq[-1] == "":
            eq.pop()                                           # strsplit drops a trailing empty string
        txps_size = int(eq[0]); eq = eq[1:]                    # :43
        tid_set = [int(x) + 1 for x in eq[0:txps_size]]; eq = eq[txps_size:]   # :44-45
        eq = eq[1:]                                            # :46 eqclass count
        eq = eq[1:]                                            # :47 number of cells
        pick = [k for k, x in enumerate(eq) if re.search("[AGTC]", x)]          # :50
        drop = set(pick) | {k + 1 for k in pick}
        eq2 = [x for k, x in enumerate(eq) if k not in drop]   # :51
        bc_list = [int(x) + 1 for x in eq2[0::2]]              # :53-55
        bc_count = [int(x) for x in eq2[1::2]]                 # :56
        for b, c in zip(bc_list, bc_count):                    # :59-62
            for t in tid_set:
                rows.append((cb_all[b - 1], tx_all[t - 1], c, i))
    return tx_all, rows


def quant_inputs(path):
    """What step2 derives from that table: cells in setorder() order, per cell the eqclasses (ascending
    id, as crpcount_td re-indexes them, Rsource:438-441) with their Count, per eqclass its transcripts."""
    tx_all, rows = read_bfh_table(path)
    rows = sorted(rows, key=lambda r: r[0].encode())           # step2:59, stable like setorder
    cells = []
    per_cell = {}
    eq_tx = {}
    for bc, tx, cnt, e in rows:
        if bc not in per_cell:
            per_cell[bc] = {}
            cells.append(bc)                                   # step2:60
        per_cell[bc].setdefault(e, cnt)
        eq_tx.setdefault(e, [])
        if tx not in eq_tx[e]:
            eq_tx[e].append(tx)
    return tx_all, cells, per_cell, eq_tx
