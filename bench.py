#!/usr/bin/env python
"""bench.py — cells/s of the 2QUANT hot path (eqclass counts -> Y -> AEM -> isoform + gene
results) on N B200s, next to the CPU restatement of the reference timed on the host.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config C1|C2|C3|C4|C5] [--cells C]

Workloads = BASELINE.json `configs` (SURVEY.md §8 d), all synthetic and seeded:
  C1  200 cells, shipped hg38 alevin X-matrix, reference options            (the reference's own CPU-runnable case)
  C2  3,955 cells (40 chunks of 98-99), same X
  C3  10,000 cells on a "GRCh38.106-like" X-matrix: 93,400 blocks / ~250k transcripts resampled from the shipped blocks
  C4  100,000 cells PER GPU on the shipped X (default; the configuration the metric is quoted on; weak scaling)
  C5  convergence stress: every EM block replaced by a wide one (q >= 8, p >= 4), library x4, maxerr 1e-6 on the
      outer AND the inner test, caps 1000/1000, 125,000 cells per GPU (1M over 8)
One step = one pass of the whole path over those cells.  Ranks are independent (no collective on the path);
torch.distributed is only used for the barrier and the max-over-ranks time.

One JSON line on stdout (rank 0).  `value`: inputs resident in HBM, results left in HBM in the library's
sparse form (CSR by cell, scasa_fetch_iso_csr) — no host copies in the timed region.  `e2e`: the same through
Scasa.quant with pinned HOST buffers — eqclass->row map on the host, H2D of the counts, kernels, D2H of the
isoform and gene results expanded into the caller's DENSE matrices (what the R call site receives) — every step.
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

CONFIGS = {
    "C1": dict(cells=200, x="shipped", seed=1, desc="C1: 200 synthetic cells on the shipped hg38 alevin X-matrix, reference convergence options"),
    "C2": dict(cells=3955, x="shipped", seed=2, desc="C2: 3,955 synthetic cells (40 chunks of 98-99) on the shipped hg38 alevin X-matrix"),
    "C3": dict(cells=10000, x="grch38_106_like", seed=3,
               desc="C3: 10,000 synthetic cells on a GRCh38.106-like X-matrix (93,400 blocks, ~250k transcripts, resampled from the shipped blocks)"),
    "C4": dict(cells=100000, x="shipped", seed=4,
               desc="C4: 100k-cell synthetic atlas on the shipped hg38 alevin X-matrix (29,351 blocks, 3,960 EM blocks), reference convergence options"),
    "C5": dict(cells=125000, x="heavy", seed=5, lib_scale=4.0, opts=dict(maxerr=1e-6, maxerr_bt=1e-6, lim=0.01, maxiter_x=1000, maxiter_bt=1000),
               desc="C5: convergence stress, every EM block wide (q >= 8, p >= 4; up to 53x14 / 50x20), library x4, maxerr 1e-6 "
                    "(outer and inner test), caps 1000/1000, 125k cells per GPU"),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="C4", choices=sorted(CONFIGS))
    ap.add_argument("--cells", type=int, default=0, help="cells per GPU (0 = the config's)")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--e2e-cells", type=int, default=0, help="cells per GPU for the e2e leg (0 = same, capped by host RAM)")
    ap.add_argument("--cpu-cells", type=int, default=0, help="cells of the cpu_baseline sample (0 = 100 per host thread, max 3200)")
    ap.add_argument("--seed", type=int, default=0, help="0 = the config's")
    ap.add_argument("--files-cells", type=int, default=5000, help="cells per GPU of the e2e_files leg (bfh.txt in, two text files out); 0 = skip")
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
            self.stop_flag.wait(0.1)

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


def load_x(kind: str = "shipped"):
    from scasa_b200 import synth
    from scasa_b200.xmatrix import XMatrix
    xm = XMatrix.load_npz(os.path.join(ROOT, "tests", "golden", "xmatrix_hg38_alevin.npz"))
    if kind == "grch38_106_like":
        return synth.resample_xmatrix(xm, 93400, seed=106)        # 29,351 x 250,000 / 78,545 blocks (SURVEY.md §8 d)
    if kind == "heavy":
        return synth.heavy_multimapping_xmatrix(xm, seed=5)
    return xm


def gene_map(xm):
    """scasa_genemat's row layout with every isoform row kept (singles, then multis, C collation).  The
    synthetic X-matrices carry no annotation: there one gene per block (a block's transcripts share reads,
    i.e. a gene or a paralog family), rows in a seeded order that stands in for the sort by name."""
    if xm.gene_of_col is None:
        p = np.diff(xm.dm0_p_off)
        order = np.random.default_rng(12345).permutation(xm.n_blocks)
        single, multi = order[p[order] == 1], order[p[order] > 1]
        rank = np.empty(xm.n_blocks, np.int32)
        rank[np.concatenate([single, multi])] = np.arange(xm.n_blocks, dtype=np.int32)
        return np.repeat(rank, p).astype(np.int32), int(xm.n_blocks)
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


def cpu_reference(xm, cfg, args, nthreads: int, steps: int = 1, warmup: int = 0):
    """The reference's CPU path, restated (oracle/): crpcount + estim_chunk + gene sums on a
    bounded sample of the same workload, all host threads over chunks (≙ registerDoParallel)."""
    from oracle import oracle as O
    from scasa_b200 import synth
    O.build()
    seed = args.seed or cfg["seed"]
    n = args.cpu_cells or min(3200, 100 * nthreads, cfg["cells"] if cfg["cells"] < 1000 else 1 << 30)
    if cfg["x"] == "grch38_106_like" and not args.cpu_cells:
        n = min(n, 800)                                       # three times the blocks: slower per cell
    # C5 keeps 100 cells per thread: a chunk of ~100 cells is the unit of work (cells of a chunk are coupled through the X
    # update), and one such chunk of wide blocks at maxerr 1e-6 takes a core about a minute — fewer cells would mean
    # smaller chunks, which converge much faster and are not the workload
    gen = synth.CellGenerator(xm, seed=seed, device="cpu", lib_scale=cfg.get("lib_scale", 1.0))
    rowptr, rows, vals = gen.counts_csr(0, n)
    chunks = O.chunk_split(n, nthreads if n <= 100 else 4)
    nthreads = max(1, min(nthreads, len(chunks) - 1))
    goc, n_genes = gene_map(xm)
    nmem = np.bincount(goc[goc >= 0], minlength=n_genes).astype(np.int32)
    ed = synth.eqclass_dictionary(xm, seed)
    cell_ptr, eq_id, eq_cnt = synth.counts_to_eqcounts(rowptr, rows, vals, seed)
    handle = O.CrpHandle(xm)      # the reference rebuilds c2t/t2c per cell (Rsource:423-427); the port does it once
    opts = O.Opts.make(**cfg.get("opts", {}))
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        y = O.crpcount_cells(handle, ed.eq_off, ed.eq_tx, cell_ptr, eq_id, eq_cnt, nthreads=nthreads)
        iso, _ = O.estim_chunks(xm, y, chunks, opts, nthreads=nthreads)
        O.gene_sum(iso, goc, nmem)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    t = float(np.mean(times))
    import shutil
    rscript = shutil.which("Rscript")                 # probe at run time
    return {"value": n / t, "unit": "cells/s", "cores": nthreads, "kind": "port",
            "sample": f"{n} cells of the same generator ({len(chunks) - 1} chunks), crpcount_td + estim_chunk + gene sums",
            "seconds": t, "steps_run": steps, "warmup_run": warmup, "rscript": rscript,
            "note": ("C restatement of the R functions (Rscript probe: not installed on this box)" if not rscript else
                     "C restatement of the R functions (Rscript is on this box, but the reference's scripts are not: "
                     "nothing under /root/reference travels)") +
                    "; the port builds crpcount_td's c2t/t2c maps once, R rebuilds them per cell; "
                    "README-derived R figures for the whole quant stage: 0.045-0.6 cells/s"}


def committed_profile(config: str, fp32: bool) -> dict:
    """ncu numbers of the dominant kernels from the capture committed under profiles/ (a number taken under a
    profiler is never a bench value; these explain the bench value).  {} when there is none for this config."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "em_physical.json")))
        return d.get(config + ("_fp32" if fp32 else ""), {})
    except Exception:
        return {}


def main():
    args = parse()
    cfg = dict(CONFIGS[args.config])
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    n = args.cells or cfg["cells"]
    seed = args.seed or cfg["seed"]
    cfg["cells"] = n

    if args.impl == "reference":
        if rank != 0:
            return 0
        xm = load_x(cfg["x"])
        nthreads = os.cpu_count() or 1
        steps, warm = max(1, min(args.steps, 3)), min(args.warmup, 1)     # bounded: the whole arm ends within minutes
        r = cpu_reference(xm, cfg, args, nthreads, steps=steps, warmup=warm)
        line = {"impl": "reference", "metric": "cells/sec EM quant (2QUANT hot path)", "value": r["value"],
                "unit": "cells/s", "n_gpus": args.gpus, "steps": steps, "warmup": warm,
                "steps_requested": args.steps, "warmup_requested": args.warmup,
                "ms_per_step": r["seconds"] * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": cfg["desc"], "cells_per_gpu": n, "sample": r["sample"]},
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

    xm = load_x(cfg["x"])
    dev = torch.device("cuda", local)
    # every rank takes its own batch-aligned cell range of one global synthetic atlas
    start = rank * ((n + synth.BATCH - 1) // synth.BATCH) * synth.BATCH
    gen = synth.CellGenerator(xm, seed=seed, device=dev, lib_scale=cfg.get("lib_scale", 1.0))
    rowptr, rows, vals = gen.counts_csr(start, n)
    ed = synth.eqclass_dictionary(xm, seed)
    cell_ptr, eq_id, eq_cnt = synth.counts_to_eqcounts(rowptr, rows, vals, seed + rank, device=dev)
    del gen
    torch.cuda.empty_cache()
    goc, n_genes = gene_map(xm)
    chunks = chunk_split(n, 4)

    g = Scasa(xm, device=local, fp32=1 if args.fp32 else 0, **cfg.get("opts", {}))
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
    warm = max(args.warmup, 3)
    for _ in range(warm):
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
    iso_nnz, gene_nnz = g.result_nnz()

    # ------------------------------------------------------------------ e2e: host buffers every step
    e2e = e2e_csr = None
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
                # the results are written by host threads (sparse rows on the wire), so the caller's matrices
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
               "includes": "host eqclass->row map, H2D counts, pack+EM+gene kernels, D2H isoform+gene results "
                           "(sparse rows on the wire, expanded into the caller's dense pageable matrices by host threads)",
               "api": "scasa_b200.api.Scasa.quant -> scasa_quant (chunks in slices, two contexts: transfer of one slice "
                      "overlaps the computation of the next)", "host_threads": nthreads}
        del h_iso, h_gene
        # the same call with sparse results (scasa_quant_csr: the dgCMatrix slots of the reference's matrices)
        from scasa_b200.api import quant_csr

        def csr_step():
            iso_s, gene_s, cs_, _ = quant_csr([g], h_ptr, h_id, h_cnt, ed.eq_off, ed.eq_tx, core=4, drop_zeros=True, n_threads=nthreads)
            out = (iso_s.nnz, gene_s.nnz, float(iso_s.val[:1].sum() + gene_s.val[-1:].sum() + cs_[0]))
            iso_s.close(); gene_s.close()
            return out

        csr_step()
        barrier()
        g.transfer_bytes(reset=True)
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            nz_i, nz_g, _ = csr_step()
        barrier()
        tc = time.perf_counter() - t0
        wire_h2d, wire_d2h = g.transfer_bytes(reset=True)
        tt = torch.tensor([tc], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        tc = float(tt[0]) / args.e2e_steps
        e2e_csr = {"value": ne * world / tc, "unit": "cells/s", "cells_per_gpu": ne, "s_per_step": tc,
                   "h2d_bytes_per_step": int(wire_h2d // args.e2e_steps), "d2h_bytes_per_step": int(wire_d2h // args.e2e_steps),
                   "host_result_bytes": int((nz_i + nz_g) * 12 + 2 * (ne + 1) * 8),
                   "result_entries": {"isoform": int(nz_i), "gene": int(nz_g)},
                   "includes": "as e2e, but the results stay sparse on the host (stored zeros dropped): "
                               "CSR by cell = the p/i/x slots of the dgCMatrix of the reference's tx x cells matrices",
                   "api": "scasa_b200.api.quant_csr -> scasa_quant_csr", "host_threads": nthreads}
        for a in (h_ptr, h_id, h_cnt):
            _l.pinned_free(a)

    # ------------------------------------------------------------------ e2e_files: bfh.txt in, the two text files out
    e2e_files = None
    if not args.no_e2e and args.files_cells > 0:
        import shutil
        import tempfile
        from scasa_b200.api import quant_files
        from scasa_b200.bfh import read_bfh
        nf = min(args.files_cells, n)
        tmp = tempfile.mkdtemp(prefix=f"scasa_bench_{rank}_")
        try:
            bc = synth.barcodes(nf, seed + rank)
            nz = int(cell_ptr[nf])
            bfh_path = os.path.join(tmp, "bfh.txt")
            bfh_bytes = synth.write_bfh(bfh_path, xm.tx_names, bc, ed.eq_off, ed.eq_tx, cell_ptr[: nf + 1], eq_id[:nz], eq_cnt[:nz])
            tx_names = xm.dm0_colnames()
            gnames = list(xm.gene_names) if xm.gene_names is not None else [f"GENE{i}" for i in range(n_genes)]
            if xm.gene_of_col is not None:
                g.set_gene_map(np.asarray(xm.gene_of_col, np.int32), len(gnames))      # ids as the annotation has them; the writer orders the rows

            def files_step():
                b = read_bfh(bfh_path, nthreads)                                       # step1_1 + step2:59-60
                return quant_files([g], b.cell_ptr, b.eq_id, b.eq_count, b.eq_off, b.tx_ids(xm),
                                   os.path.join(tmp, "scasa_isoform_counts.txt"), os.path.join(tmp, "scasa_gene_expression.txt"),
                                   tx_names, gnames, b.barcodes, core=4, n_threads=nthreads)

            files_step()
            barrier()
            t0 = time.perf_counter()
            _, ftm = files_step()
            barrier()
            tf = time.perf_counter() - t0
            tt = torch.tensor([tf], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            tf = float(tt[0])
            out_bytes_f = ftm["iso_bytes"] + ftm["gene_bytes"]
            e2e_files = {"value": nf * world / tf, "unit": "cells/s", "cells_per_gpu": nf, "s_per_step": tf,
                         "bfh_bytes": int(bfh_bytes), "text_bytes_written": int(out_bytes_f), "text_gb_per_s": out_bytes_f / 1e9 / max(ftm["write_iso_s"] + ftm["write_gene_s"], 1e-9),
                         "phases_s": {k: ftm[k] for k in ("quant_s", "transpose_s", "write_iso_s", "write_gene_s")},
                         "rows_written": {"isoform": ftm["iso_rows"], "gene": ftm["gene_rows"]},
                         "includes": "scasa_bfh_open (parse alevin bfh.txt), host eqclass->row map, scasa_quant_files: H2D, kernels, sparse D2H, "
                                     "host transposition, scasa_isoform_counts.txt + scasa_gene_expression.txt written (write.table bytes); "
                                     "no dense matrix anywhere",
                         "dir": "tempfile.mkdtemp() on the box's local disk", "host_threads": nthreads}
            g.set_gene_map(goc, n_genes)
        finally:
            shutil.rmtree(tmp, ignore_errors=True)

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
    prof = committed_profile(args.config, args.fp32) if (not args.cells and not os.environ.get("SCASA_EM_CFG")) else {}
    b_io = 8.0 * (xm.n_rows + xm.n_cols)
    roofline = {"bound": "hbm",
                "kernel": "em_task_kernel<W> + em_pair_kernel (5 persistent launches per step: W = 1/2/4/8 warps per task and the "
                          "tiny tasks two per warp, side by side on 5 streams; em_ms = fork to join, CUDA events on the library's stream)",
                "achieved": em_bytes / (em_ms * 1e-3) / 1e9, "peak": peak,
                "peak_source": "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)",
                "unit": "GB/s",
                # dram__bytes_read.sum + dram__bytes_write.sum of the EM launches of one step, from the ncu capture of this
                # config committed under profiles/ (profiles/em_physical.json names the file); null when there is none
                "traffic": prof.get("em_dram_bytes"),
                "model": "SURVEY.md 8(d) streaming model: sum over (chunk, block) tasks of inner*8*(q+2p)*n_c + "
                         "outer*8*(q+p)*n_c with the iteration counts the kernels actually ran",
                "alg_bytes_per_launch_set": em_bytes, "em_ms": em_ms,
                "note": "NOMINAL: a task stays in shared memory for all its iterations, so the kernels move `traffic` bytes, not the "
                        "streaming-model bytes; frac > 1 says the model does not describe them.  `physical` does.",
                "physical": ({"dram_gbs": prof["em_dram_bytes"] / (em_ms * 1e-3) / 1e9, "dram_frac_of_peak": prof["em_dram_bytes"] / (em_ms * 1e-3) / 1e9 / peak,
                              "fp64_pipe_pct": prof.get("fp64_pipe_pct"), "issue_active_pct": prof.get("issue_active_pct"),
                              "lanes_per_inst": prof.get("lanes_per_inst"), "bound": "instruction issue", "source": prof.get("source")}
                             if prof.get("em_dram_bytes") else None),
                "whole_step": {"alg_bytes": t["alg_bytes"], "ms": t["total_ms"],
                               "achieved": t["alg_bytes"] / (t["total_ms"] * 1e-3) / 1e9,
                               "frac": t["alg_bytes"] / (t["total_ms"] * 1e-3) / 1e9 / peak,
                               "dram_bytes": prof.get("step_dram_bytes")},
                "compulsory": {"bytes": n * b_io, "achieved": n * b_io / (t["total_ms"] * 1e-3) / 1e9}}
    roofline["frac"] = roofline["achieved"] / peak
    line = {"metric": "cells/sec EM quant (2QUANT hot path)", "value": n * world / step_s, "unit": "cells/s",
            "n_gpus": world, "steps": args.steps, "warmup": warm, "ms_per_step": step_s * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32" if args.fp32 else "f64", "data": "synthetic",
            "config": {"workload": cfg["desc"], "config": args.config,
                       "cells_per_gpu": n, "chunks_per_gpu": len(chunks) - 1, "seed": seed,
                       "x_matrix": {"blocks": int(xm.n_blocks), "rows": int(xm.n_rows), "columns": int(xm.n_cols), "em_blocks": int(g.n_em)},
                       "options": cfg.get("opts", "reference defaults (maxerr .01, lim .01, 1000/1000, digamma on)"),
                       "value_excludes": "host<->device copies: inputs resident in HBM, results left in HBM as CSR by cell "
                                         "(%.2f GB isoform + %.2f GB gene entries per GPU); `e2e` includes them" %
                                         (iso_nnz * 12 / 1e9, gene_nnz * 12 / 1e9),
                       "l2": "inputs larger than L2 (counts %.2f GB per GPU)" % ((eq_id.nbytes + eq_cnt.nbytes) / 1e9)
                             if (eq_id.nbytes + eq_cnt.nbytes) > 200e6 else "small workload: per-step working set is rebuilt by the step's own kernels (task blobs, results); inputs %.1f MB" % ((eq_id.nbytes + eq_cnt.nbytes) / 1e6),
                       "parallelism": f"cells sharded by chunk over {world} GPU(s), no collective"},
            "device_ms_per_step": dev_max * 1e3 / args.steps,
            "phases_ms": {k: t[k] for k in ("pack_ms", "tasks_ms", "em_ms", "finish_ms", "gene_ms")},
            "em_class_ms": t["em_class_ms"], "em_class_tasks": t["em_class_tasks"],
            "em_pair_ms": t["em_pair_ms"], "em_pair_tasks": t["em_pair_tasks"], "max_slot_bytes": t["max_slot_bytes"],
            "work": {"tasks": t["n_tasks"], "active_cell_blocks": t["n_active"], "em_entries": t["n_entries"],
                     "inner_iters": t["inner_iters"], "outer_iters": t["outer_iters"],
                     "result_entries": {"isoform": iso_nnz, "gene": gene_nnz}},
            "gpu_launches": int(sum(x["launches"] for x in tms)),
            "clocks": clocks, "roofline": roofline}
    if e2e:
        line["e2e"] = e2e
        line["e2e_csr"] = e2e_csr
    if e2e_files:
        line["e2e_files"] = e2e_files
    if not args.no_cpu:
        line["cpu_baseline"] = cpu_reference(xm, cfg, args, os.cpu_count() or 1)
    print(json.dumps(line))
    g.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
