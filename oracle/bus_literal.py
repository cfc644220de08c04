"""TEST INFRASTRUCTURE — literal Python restatement of the reference's bustools front end.

Follows /root/reference/scasa/SCRIPTS/scasa_quant_step1_2_v1.0.0.R statement by statement (the data.table /
data.frame operations are written out with lists and dicts, in the same order): the library-size filter
(:30-32), matrix.ec / transcripts.txt (:34-46), the duplicated Barcode_UMI cases (:48-98), Count <- 1 and the
per-cell table(eqClass) (:100-119), then scasa_quant_step2_v1.0.0.R:59-60 (setorder by Barcode, unique).
Pure Python: small files only.  Only tests/ may import this.
"""


def quant_inputs(bus_txt, matrix_ec, transcripts_txt, libsize_thres):
    bus = []                                                   # :28-29 fread + colnames
    for line in open(bus_txt).read().split("\n"):
        if line:
            b, u, e, c, bu = line.split("\t")
            bus.append({"Barcode": b, "UMI_Code": u, "eqClass": int(e), "Count": int(c), "Barcode_UMI": bu})
    libsize = {}                                               # :30 bus_output[, .N, by = Barcode]
    for r in bus:
        libsize[r["Barcode"]] = libsize.get(r["Barcode"], 0) + 1
    keep = {b for b, n in libsize.items() if n > libsize_thres}          # :31
    bus = [r for r in bus if r["Barcode"] in keep]                       # :32
    eq_rows = []                                               # :34-46 eq exploded to (eqClass, Transcript_ID)
    ref = [ln.split()[0] for ln in open(transcripts_txt).read().split("\n") if ln.strip()]
    for line in open(matrix_ec).read().split("\n"):
        if line:
            ec, txs = line.split("\t")
            for t in txs.split(","):
                eq_rows.append((ec, ref[int(t)]))
    cell_umi = {}                                              # :48 .N by Barcode_UMI
    for r in bus:
        cell_umi[r["Barcode_UMI"]] = cell_umi.get(r["Barcode_UMI"], 0) + 1
    eq_size = {}                                               # :51 table(eq$eqClass)
    for ec, _ in eq_rows:
        eq_size[ec] = eq_size.get(ec, 0) + 1
    for r in bus:                                              # :52-57
        r["eqSize"] = eq_size.get(str(r["eqClass"]))
        r["Duplicate_Freq"] = cell_umi[r["Barcode_UMI"]]
    dup = [r for r in bus if r["Duplicate_Freq"] > 1]          # :58
    bus = [r for r in bus if r["Duplicate_Freq"] == 1]         # :59
    groups = {}
    for r in dup:
        groups.setdefault(r["Barcode_UMI"], []).append(r)
    max_count_num = {k: sum(x["Count"] == max(y["Count"] for y in g) for x in g) for k, g in groups.items()}   # :63
    case1 = []                                                 # :65-71
    for k, g in groups.items():
        if max_count_num[k] == 1:
            mx = max(x["Count"] for x in g)
            case1 += [x for x in g if x["Count"] == mx]
    case2 = []                                                 # :74-93
    for k, g in groups.items():
        if max_count_num[k] > 1:
            mx = max(x["Count"] for x in g)
            g2 = [x for x in g if x["Count"] == mx]            # :79-80
            sizes = [x["eqSize"] for x in g2]
            if any(s is None for s in sizes):
                continue                                       # NA max(eqSize): which(V1 == 1) drops the group
            if sum(s == max(sizes) for s in sizes) == 1:       # :82-84
                case2 += [x for x in g2 if x["eqSize"] == max(sizes)]          # :87-90
    bus = bus + case1 + case2                                  # :93-98
    for r in bus:
        r["Count"] = 1                                         # :100
    known = {ec for ec, _ in eq_rows}
    per_cell = {}                                              # :103-119 fun(): table(eqClass) joined to eq
    for r in bus:
        if str(r["eqClass"]) in known:
            d = per_cell.setdefault(r["Barcode"], {})
            d[r["eqClass"]] = d.get(r["eqClass"], 0) + 1
    cells = sorted(per_cell, key=str.encode)                   # step2:59-60
    eq_tx = {}
    for ec, tx in eq_rows:
        eq_tx.setdefault(int(ec), []).append(tx)
    return ref, cells, per_cell, eq_tx
