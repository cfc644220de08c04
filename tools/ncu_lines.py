#!/usr/bin/env python
"""Per-source-line share of executed warp instructions, stall samples and lanes per instruction from
    ncu -i X.ncu-rep --page source --csv --print-source cuda,sass --kernel-name regex:<k> --launch-count 1 > lines.csv
Usage: python tools/ncu_lines.py lines.csv [top N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 45
hi = next(i for i, r in enumerate(rows[:12]) if "Instructions Executed" in r)
hdr = rows[hi]
col = lambda name, start=0: hdr.index(name, start)
i_line, i_src, i_addr = 0, 1, 2
i_inst, i_thr, i_samp = col("Instructions Executed"), col("Thread Instructions Executed"), col("# Samples")
lines, n_sass = [], 0
for r in rows[hi + 1:]:
    if len(r) != len(hdr):
        continue
    if r[i_addr] == "-":                       # a CUDA source line: totals of its SASS
        lines.append([r[i_line], r[i_src].strip(), int(r[i_inst] or 0), int(r[i_thr] or 0), int(r[i_samp] or 0), 0])
    else:
        n_sass += 1
        if lines:
            lines[-1][5] += 1
tot_i = sum(x[2] for x in lines) or 1
tot_s = sum(x[4] for x in lines) or 1
print(f"# {rows[0][1] if len(rows[0]) > 1 else ''}")
print(f"# {n_sass} SASS instructions, {tot_i} executed warp instructions, {tot_s} stall samples, "
      f"{sum(x[3] for x in lines) / tot_i:.1f} lanes per instruction")
for ln, src, inst, thr, samp, ns in sorted(lines, key=lambda x: -x[2])[:top]:
    print(f"line {ln:>4s} sass {ns:4d} inst {100 * inst / tot_i:5.1f}% samp {100 * samp / tot_s:5.1f}% lanes {thr / max(inst, 1):5.1f} | {src[:100]}")
