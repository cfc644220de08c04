#!/usr/bin/env python
"""bench.py — cells/s of the 2QUANT hot path (eqclass counts -> Y -> AEM -> isoform + gene
matrices) on N B200s, next to the CPU restatement of the reference timed on the host.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--cells C]

Workload (BASELINE.json config 4, "100k-cell 10x-scale synthetic atlas on refMrna X-matrix"):
`--cells` synthetic cells PER GPU (weak scaling; default 100000 = 1000 chunks of 100,
scasa_quant_step2_v1.0.0.R:64-68) on the shipped hg38 alevin X-matrix (tests/golden fixture),
reference convergence options (maxerr .01, lim .01, 1000/1000, digamma on).  One step = one
pass of the whole path over those cells.  Ranks are independent (no collective on the path);
torch.distributed is only used for the barrier and the max-over-ranks time.

One JSON line on stdout (rank 0).  `value`: inputs resident in HBM, device-only.  `e2e`: the
same through scasa_b200.api.Scasa.quant-style calls with pinned HOST buffers — eqclass->row map
on the host, H2D of the counts, kernels, D2H of the isoform and gene matrices — every step.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

B_IN = 255472.0    # algorithmic bytes per cell per inner EM iteration (shipped alevin X)
B_OUT = 191176.0   # per cell per outer X update
B_IO = 597432.0    # per cell, one-time: dense Y in + estimates out


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cells", type=int, default=100000, help="cells per GPU")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--e2e-cells", type=int, default=0, help="cells per GPU for the e2e leg (0 = same, capped by host RAM)")
    ap.add_argument("--cpu-cells", type=int, default=0, help="cells of the cpu_baseline sample (0 = 100 per host thread, max 3200)")
    ap.add_argument("--seed", type=int, default=4)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--fp32", action="store_true", help="EM iterates in fp32 (scasa_opts_t.fp32); the default and the parity path is fp64")
    return ap.parse_args()


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons while the timed region runs (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.samples = index, threading.Event(), []

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                if len(f) >= 6:
                    self.samples.append(f)
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self) -> dict:
        self.stop_flag.set()
        self.join(timeout=6)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        mx = [int(s[1]) for s in self.samples if s[1].isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.samples)}


def load_x():
    from scasa_b200.xmatrix import XMatrix
    return XMatrix.load_npz(os.path.join(ROOT, "tests", "golden", "xmatrix_hg38_alevin.npz"))


def gene_map(xm):
    """scasa_genemat's row layout with every isoform row kept (singles, then multis, C collation)."""
    names = [xm.gene_names[g] for g in xm.gene_of_col]
    members: dict = {}
    for j, g in enumerate(names):
        members.setdefault(g, []).append(j)
    order = sorted((g for g in members if len(members[g]) == 1), key=str.encode) + \
        sorted((g for g in members if len(members[g]) > 1), key=str.encode)
    goc = np.full(xm.n_cols, -1, np.int32)
    for gi, g in enumerate(order):
        goc[members[g]] = gi
    return goc, len(order)


def cpu_reference(xm, args, nthreads: int, steps: int = 1, warmup: int = 0):
    """The reference's CPU path, restated (oracle/): crpcount + estim_chunk + gene sums on a
    bounded sample of the same workload, all host threads over chunks (≙ registerDoParallel)."""
    from oracle import oracle as O
    from scasa_b200 import synth
    from scasa_b200.api import chunk_split  # pure host arithmetic in the library (no GPU needed)
    O.build()
    n = args.cpu_cells or min(3200, 100 * nthreads)
    gen = synth.CellGenerator(xm, seed=args.seed, device="cpu")
    rowptr, rows, vals = gen.counts_csr(0, n)
    chunks = O.chunk_split(n, nthreads)
    del chunk_split
    nthreads = max(1, min(nthreads, len(chunks) - 1))
    goc, n_genes = gene_map(xm)
    nmem = np.bincount(goc[goc >= 0], minlength=n_genes).astype(np.int32)
    have_counts = True
    ed = synth.eqclass_dictionary(xm, args.seed)
    cell_ptr, eq_id, eq_cnt = synth.counts_to_eqcounts(rowptr, rows, vals, args.seed)
    handle = O.CrpHandle(xm)      # the reference rebuilds c2t/t2c per cell (Rsource:423-427); the port does it once
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        if have_counts:
            y = O.crpcount_cells(handle, ed.eq_off, ed.eq_tx, cell_ptr, eq_id, eq_cnt, nthreads=nthreads)
        else:
            y = synth.csr_to_dense(rowptr, rows, vals, xm.n_rows)
        iso, _ = O.estim_chunks(xm, y, chunks, None, nthreads=nthreads)
        O.gene_sum(iso, goc, nmem)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    t = float(np.mean(times))
    sample = f"{n} cells of the same generator ({len(chunks) - 1} chunks), " + \
        ("crpcount_td + estim_chunk + gene sums" if have_counts else "estim_chunk + gene sums (Y given)")
    import shutil
    rscript = shutil.which("Rscript")                 # probe at run time
    return {"value": n / t, "unit": "cells/s", "cores": nthreads, "kind": "port", "sample": sample,
            "seconds": t, "rscript": rscript,
            "note": ("C restatement of the R functions (Rscript probe: not installed on this box)" if not rscript else
                     "C restatement of the R functions (Rscript is on this box, but the reference's scripts are not: "
                     "nothing under /root/reference travels)") +
                    "; README-derived R figures for the whole quant stage: 0.045-0.6 cells/s"}


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank != 0:
            return 0
        xm = load_x()
        nthreads = os.cpu_count() or 1
        r = cpu_reference(xm, args, nthreads, steps=max(1, min(args.steps, 3)), warmup=min(args.warmup, 1))
        line = {"impl": "reference", "metric": "cells/sec EM quant (2QUANT hot path)", "value": r["value"],
                "unit": "cells/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": r["seconds"] * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": "C4: 100k-cell synthetic atlas on the shipped hg38 alevin X-matrix",
                           "cells_per_gpu": args.cells, "sample": r["sample"]},
                "cpu_baseline": r,
                "e2e": {"value": r["value"], "unit": "cells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        print(json.dumps({"error": "no CUDA device: bench.py has no CPU path for the b200 arm"}))
        return 2
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    from scasa_b200 import lib as _l, synth
    from scasa_b200.api import Scasa, chunk_split

    xm = load_x()
    n = args.cells
    dev = torch.device("cuda", local)
    # every rank takes its own batch-aligned cell range of one global synthetic atlas
    start = rank * ((n + synth.BATCH - 1) // synth.BATCH) * synth.BATCH
    gen = synth.CellGenerator(xm, seed=args.seed, device=dev)
    rowptr, rows, vals = gen.counts_csr(start, n)
    ed = synth.eqclass_dictionary(xm, args.seed)
    cell_ptr, eq_id, eq_cnt = synth.counts_to_eqcounts(rowptr, rows, vals, args.seed + rank, device=dev)
    del gen
    torch.cuda.empty_cache()
    goc, n_genes = gene_map(xm)
    chunks = chunk_split(n, 4)

    g = Scasa(xm, device=local, fp32=1 if args.fp32 else 0)
    g.set_gene_map(goc, n_genes)
    nthreads = max(1, (os.cpu_count() or 8) // max(world, 1))
    eq_row = g.eqclass_map(ed.eq_off, ed.eq_tx, nthreads)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ------------------------------------------------------------------ value: resident inputs
    g.stage_eqcounts(chunks, cell_ptr, eq_id, eq_cnt, eq_row)
    for _ in range(max(args.warmup, 3)):
        g.run()
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    t0 = time.perf_counter()
    dev_ms, tms = 0.0, []
    for _ in range(args.steps):
        t = g.run()                       # synchronises its stream before returning
        dev_ms += t["total_ms"]
        tms.append(t)
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.summary()
    tt = torch.tensor([wall, dev_ms / 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    wall_max, dev_max = float(tt[0]), float(tt[1])
    step_s = wall_max / args.steps
    t = tms[-1]

    # ------------------------------------------------------------------ e2e: host buffers every step
    e2e = None
    if not args.no_e2e:
        import psutil
        ne = args.e2e_cells or n
        out_bytes = lambda m: m * (xm.n_cols + n_genes) * 8
        avail = psutil.virtual_memory().available / max(world, 1)
        while ne > 1000 and out_bytes(ne) > 0.5 * avail:
            ne //= 2
        ne = min(ne, n)
        # pinned host buffers for the inputs, pageable ones for the two dense results; if the box cannot hold
        # that much (several ranks share one host) fall back to a smaller sample rather than to no number
        while True:
            bufs = []
            try:
                nnz_e = int(cell_ptr[ne])
                h_ptr = _l.pinned_empty(ne + 1, np.int64); bufs.append(h_ptr)
                h_id = _l.pinned_empty(nnz_e, np.int32); bufs.append(h_id)
                h_cnt = _l.pinned_empty(nnz_e, np.int32); bufs.append(h_cnt)
                # the results are written by host threads (zero-suppressed transfer), so the caller's matrices
                # need not be pinned: plain pageable memory, like the R matrices of the reference
                h_iso = np.empty((ne, xm.n_cols), np.float64)
                h_gene = np.empty((ne, n_genes), np.float64)
                break
            except Exception:
                for a_ in bufs:
                    _l.pinned_free(a_)
                if ne <= 1000:
                    raise
                ne //= 2
        h_ptr[:] = cell_ptr[: ne + 1]
        h_id[:] = eq_id[:nnz_e]
        h_cnt[:] = eq_cnt[:nnz_e]
        g.set_host_threads(nthreads)

        def e2e_step():
            # the call a user makes: host eqclass counts in, dense host matrices out
            r = g.quant(h_ptr, h_id, h_cnt, ed.eq_off, ed.eq_tx, core=4, n_threads=nthreads, iso_out=h_iso,
                        gene_out=h_gene, want_iters=False)
            return float(r.iso[0].sum() + r.gene[-1].sum() + r.colsum[0])

        e2e_step()
        barrier()
        g.transfer_bytes(reset=True)
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            e2e_step()
        barrier()
        te = time.perf_counter() - t0
        wire_h2d, wire_d2h = g.transfer_bytes(reset=True)
        tt = torch.tensor([te], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        te = float(tt[0]) / args.e2e_steps
        e2e = {"value": ne * world / te, "unit": "cells/s", "cells_per_gpu": ne, "s_per_step": te,
               "h2d_bytes_per_step": int(wire_h2d // args.e2e_steps), "d2h_bytes_per_step": int(wire_d2h // args.e2e_steps),
               "host_input_bytes": int(h_ptr.nbytes + h_id.nbytes + h_cnt.nbytes + eq_row.nbytes),
               "host_result_bytes": int(h_iso.nbytes + h_gene.nbytes),
               "includes": "host eqclass->row map, H2D counts, pack+EM+gene kernels, D2H isoform+gene matrices "
                           "(zero-suppressed on the wire, dense in the caller's pageable buffers)",
               "api": "scasa_b200.api.Scasa.quant -> scasa_quant (chunks in slices, two contexts: transfer of one slice "
                      "overlaps the computation of the next)", "host_threads": nthreads}
        for a in (h_ptr, h_id, h_cnt):
            _l.pinned_free(a)
        del h_iso, h_gene

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    em_bytes = t["em_alg_bytes"]
    em_ms = float(np.mean([x["em_ms"] for x in tms]))          # average launch duration over the timed steps
    fin_ms = float(np.mean([x["finish_ms"] for x in tms]))
    fin_bytes = n * (xm.n_cols * 8 * 2 + n_genes * 8)           # colsum reads iso; gene_gather reads iso and writes genes
    roofline = {"bound": "hbm", "kernel": "em_task_kernel<W> + em_pair_kernel (5 persistent launches per step: W = 1/2/4/8 warps per task and the tiny tasks two per warp, side by side on 5 streams; em_ms = fork to join)",
                "achieved": em_bytes / (em_ms * 1e-3) / 1e9, "peak": peak,
                "peak_source": "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)",
                "unit": "GB/s",
                # dram__bytes_read.sum + dram__bytes_write.sum of the four em_task_kernel launches of one step, from the
                # ncu capture of this command committed under profiles/ (r1h_launches_100k.txt); only valid for the default workload
                "traffic": 4.6e9 if (n == 100000 and not os.environ.get("SCASA_EM_CFG")) else None,
                "model": "SURVEY.md 8(d) streaming model: sum over (chunk, block) tasks of inner*8*(q+2p)*n_c + "
                         "outer*8*(q+p)*n_c with the iteration counts the kernel actually ran",
                "alg_bytes_per_launch_set": em_bytes, "em_ms": em_ms,
                "note": "a task stays in shared memory for all its iterations, so the kernel moves `traffic` bytes, not the "
                        "streaming-model bytes: frac > 1 is expected; ncu shows it instruction-issue bound (DESIGN.md 4-5)",
                "hbm_phase": {"kernels": "colsum_kernel + colsum_reduce + gene_gather_kernel (stream the result matrices)",
                              "bytes": fin_bytes, "ms": fin_ms, "achieved": fin_bytes / (fin_ms * 1e-3) / 1e9,
                              "frac": fin_bytes / (fin_ms * 1e-3) / 1e9 / peak},
                "whole_step": {"alg_bytes": t["alg_bytes"], "ms": t["total_ms"],
                               "achieved": t["alg_bytes"] / (t["total_ms"] * 1e-3) / 1e9},
                "compulsory": {"bytes": n * B_IO, "achieved": n * B_IO / (t["total_ms"] * 1e-3) / 1e9}}
    roofline["frac"] = roofline["achieved"] / peak
    line = {"metric": "cells/sec EM quant (2QUANT hot path)", "value": n * world / step_s, "unit": "cells/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": step_s * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32" if args.fp32 else "f64", "data": "synthetic",
            "config": {"workload": "C4: 100k-cell synthetic atlas on the shipped hg38 alevin X-matrix "
                                   "(29,351 blocks, 3,960 EM blocks), reference convergence options",
                       "cells_per_gpu": n, "chunks_per_gpu": len(chunks) - 1, "seed": args.seed,
                       "l2": "inputs larger than L2 (counts %.1f GB, results %.1f GB per GPU)" %
                             ((eq_id.nbytes + eq_cnt.nbytes) / 1e9, n * (xm.n_cols + n_genes) * 8 / 1e9),
                       "parallelism": f"cells sharded by chunk over {world} GPU(s), no collective"},
            "device_ms_per_step": dev_max * 1e3 / args.steps,
            "phases_ms": {k: t[k] for k in ("pack_ms", "tasks_ms", "em_ms", "finish_ms")},
            "em_class_ms": t["em_class_ms"], "em_class_tasks": t["em_class_tasks"],
            "em_pair_ms": t["em_pair_ms"], "em_pair_tasks": t["em_pair_tasks"], "max_slot_bytes": t["max_slot_bytes"],
            "work": {"tasks": t["n_tasks"], "active_cell_blocks": t["n_active"], "em_entries": t["n_entries"],
                     "inner_iters": t["inner_iters"], "outer_iters": t["outer_iters"]},
            "gpu_launches": int(sum(x["launches"] for x in tms)),
            "clocks": clocks, "roofline": roofline}
    if e2e:
        line["e2e"] = e2e
    if not args.no_cpu:
        line["cpu_baseline"] = cpu_reference(xm, args, os.cpu_count() or 1)
    print(json.dumps(line))
    g.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
