#!/usr/bin/env python
"""Regenerates INTEGRATION.md with the shipped integration/ files embedded verbatim (tests/test_integration.py checks it)."""
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rd = lambda f: open(os.path.join(ROOT, "integration", f)).read().strip()
MD = r'''# INTEGRATION — calling libscasa_cuda from Scasa's R scripts

The reference has no plugin or FFI layer: `scasa_quant_step2_v1.0.0.R` sources `Scasa_Rsource.R` and calls two
closures inside `foreach … %dopar%` loops (`crpcount_td` at `step2:95`, `estim_chunk` at `step2:132`), and
`scasa_quant_step3_v1.0.0.R` calls `scasa_genemat` (`step3:48`). The drop-in keeps the driver, its inputs
(`Sample_eqClass.RData`, `Xmatrix_*.RData`, `isoform_groups_*.RData`) and its outputs
(`scasa_isoform_counts.{RDS,txt}`, `scasa_gene_counts.RDS`, `scasa_gene_expression.txt`) and replaces
`step2:88-167` and `step3:48` by **one call from the R master process** into `libscasa_cuda.so`
(never from a forked `%dopar%` worker: a CUDA context does not survive `fork()`). The library shards the chunks over
the GPUs of the box inside that call — what `registerDoParallel(cores = core)` (`step2:82`) does over CPU cores.

R and its headers are not present in the build image, so the shim below is not compiled or run here; it is
deliberately thin — every decision lives behind the C ABI (`include/scasa_cuda.h`), which *is* tested
(`tests/`, through `ctypes`), and the shim is type-checked against that header with stand-in R declarations.

## 1. C ABI ↔ reference interface

| C entry point | replaces | notes |
|---|---|---|
| `scasa_ctx_create(dm0, crp, opts, device, &ctx)` | `load(design.matrix)` + `DM0` (`step2:39,127`) being passed around | uploads the flattened blocks once, one context per GPU; `ccrpfun` itself stays in R (R's RNG/kmeans), its result is the `dm0` argument |
| `scasa_dm0_is_fixed_point(crp, dm0, member, clim, …)` | the serial `DM0 <- lapply(CRP, ccrpfun, 30)` (`step2:127`; svd + `kmeans(nstart=100)` over 29,351 blocks) | host only: when the shipped `CCRP` is a point `ccrpfun`'s loop stops at (it is, for all 29,351 blocks: `tests/test_oracle.py`) the recompute is skipped |
| `scasa_chunk_split(n, core, …)` | `step2:64-68` | identical rounding (`round` half-to-even, `seq(length.out=)`) |
| `scasa_eqclass_map(ctx, E, eq_off, eq_tx, eq_row, threads)`, `scasa_eqclass_map_host(crp, …)` | the block/row decision inside `crpcount_td` + `hamming_correction` (`Rsource:453-546`, `:562-648`) | once per distinct eqclass instead of once per cell and block; the `_host` form needs no GPU |
| `scasa_quant` / `scasa_quant_multi(ctxs, n_ctx, …)` | the two `foreach` bodies `step2:88-115`, `:128-136`, the stacking `step2:139-165` and (with a gene map) `step3:48` | one call for a whole sample over 1..N GPUs; rows in barcode order (= `order(rownames)`), columns in block order like `cbind(final, BT)`; dense host matrices out; results bitwise independent of the GPU count |
| `scasa_quant_files(ctxs, n_ctx, …, iso_path, gene_path, …)` | the same plus `step2:165-170` and `step3:51-52` | eqclass counts in, `scasa_isoform_counts.txt` and `scasa_gene_expression.txt` out: no dense matrix in R or anywhere else |
| `scasa_quant_csr(ctxs, n_ctx, …, drop_zeros, &iso, &gene, …)` + `scasa_csr_free` | the same as `scasa_quant_multi`, results sparse | host eqclass counts in, `scasa_csr_t` out: CSR by cell = the `p`/`i`/`x` slots of the `Matrix::dgCMatrix` of the reference's tx × cells matrices (`step2:166`), so R gets `isoform_count` and the gene matrix without a dense matrix on either side (100k cells: ≈ 3 GB instead of 26.7 + 19.9 GB); bit for bit the values of the dense call (`SCASA_SPARSE=1` in the step-2 patch) |
| `scasa_stage_eqcounts` + `scasa_run` + `scasa_fetch_iso` / `scasa_fetch_iso_csr` / `scasa_fetch_gene(_csr)` | the stages of the above | the device keeps both results sparse (CSR by cell); the `_csr` forms hand that over as it is |
| `scasa_bfh_open` / `_sizes` / `_read` / `_close` | `scasa_quant_step1_1_v1.0.0.R:16-68` (bfh.txt → exploded data.table → `Sample_eqClass.RData`) and `step2:59-60` | optional: feeds `scasa_quant` directly, no `.RData` round trip |
| `scasa_bus_open` (+ the `scasa_bfh_*` accessors) | `scasa_quant_step1_2_v1.0.0.R:28-119` (kallisto/bustools text → `Sample_eqClass.RData`) | optional: library-size filter, duplicated (barcode, UMI) rules and per-cell eqclass counts as in the script |
| `scasa_write_table`, `scasa_write_table_sparse`, `scasa_csr_transpose`, `scasa_format_real15` | `write.table(..., quote=F, row.names=T, sep="\t")` at `step2:170`, `step3:52` | same bytes as R writes (15 significant digits, fixed unless scientific is narrower), threaded; the sparse form never needs the dense matrix |
| `scasa_estimate` / `scasa_stage_counts` | `estim_chunk(Y, DM0, maxiter.bt=1000)` when `Y` already exists | `Y` as CSR by cell |
| `scasa_fetch_colsum` | `rowSums(isoform_count) > 0` (`step2:167`) | summed on the GPU in cell / chunk order |
| `scasa_set_gene_map` + `scasa_fetch_gene`, `scasa_gene_sum` | `scasa_genemat` / `parasum` (`step3:24-46`) | gene ids over ALL isoform rows; the row layout after the `step2:167` filter (single-isoform genes first, each group in `sort()` order) is applied by the caller / by `scasa_quant_files` |
| `scasa_opts_t` | `estim_chunk`'s defaults (`Rsource:944`), `adjust_esp` (`:973`), `hamming_dist` (`:534`), `estim.BETA`'s own `maxerr`/`lim` (`:726-727`) | `scasa_opts_default` fills exactly these; `maxerr`/`lim` reach the outer test only, as in R (`Rsource:1016` does not forward them) |
| return codes + `scasa_last_error` | R errors | the shim turns non-zero into `stop(msg)` |

Limits (checked, reported as `SCASA_ERR_UNSUPPORTED`): ≤ 65,535 rows and ≤ 1,024 columns per block, ≤ 65,535 cells per
chunk, ≤ 16,384 eqclasses per cell, a poor-condition block ≤ 64 rows, a (chunk, block) task ≤ 200 KB of shared memory.
Gene sums on the device keep a member table in shared memory: (isoform columns of multi-isoform genes + number of such
genes) × 2 bytes (4 when the X-matrix has ≥ 65,534 columns) + n_genes / 4 + 29 KB must stay ≤ 200 KB — about 85 k such
columns (42 k for the wide X-matrices); the shipped hg38 annotation needs 65 KB, the 106 k-column synthetic C3 case 174 KB.
Beyond that `scasa_run` reports it and `scasa_gene_sum` (any size, from a dense isoform matrix) remains.

## 2. The R shim (`integration/scasa_r.c`: `.Call`, built on the user's machine with `R CMD SHLIB`)

R's headers are not in this repository; `tests/test_integration.py` compiles the file against `include/scasa_cuda.h`
with stand-in declarations (`tests/mock_r/`), so a change of the ABI that the shim does not follow fails the CPU suite.

```c
@@SHIM@@
```

## 3. The patched call site in `scasa_quant_step2_v1.0.0.R` (`integration/scasa_quant_step2_patch.R`)

Set `SCASA_GPUS=<n>` to shard over n GPUs, `SCASA_SPARSE=1` to get `isoform_count` and the gene matrix as `dgCMatrix`
(the form Seurat-style downstream code reads; no dense matrix is built) and `SCASA_FILES=1` to let the library write the two text files itself
(R's `write.table` of a 33k x 100k matrix takes far longer than the quantification). To fuse step 3, `load(txgroup)`
(`step3:22`) before the patch runs.

```r
@@STEP2@@
saveRDS(isoform_count, ...); write.table(isoform_count, ...)      # step2:169-170 unchanged (matrix mode)
```

## 4. The patched call site in `scasa_quant_step3_v1.0.0.R` (`integration/scasa_quant_step3_patch.R`)

```r
@@STEP3@@
```

## 5. Pinning parity to R (`tools/make_golden.R`)

No R in the build image means every parity claim of this repository is against our *reading* of the R code. One command
on a machine with R changes that: `python tools/export_golden_inputs.py && Rscript tools/make_golden.R scasa=<reference>/scasa`
runs the reference's own `crpcount_td`, `hamming_correction`, `estim_chunk`, `estim.BETA`/`estim.X.em` and `scasa_genemat`
on the committed 48-cell case and writes text dumps under `tests/golden/r_dumps/`; `tests/test_r_golden.py` then
compares the CPU oracle, the string-level literal, the library's host map and (on a B200) the CUDA path with them.

### Without the `.RData` round trip (optional)
`scasa_bfh_open(path, threads, &f)` reads alevin's `bfh.txt` the way `step1_1` does (Count = number of UMIs,
cells = barcodes that occur, sorted like `setorder(Barcode)`) and `scasa_bfh_read` hands back exactly the
arrays `scasa_quant` / `scasa_quant_files` take (`cell_ptr`, `eq_id`, `eq_count`, `eq_off`, `eq_tx` as indices into the
file's transcript list); `scasa_bus_open` does the same for the kallisto/bustools text. `scasa_b200/bfh.py` is the
ctypes mirror.
'''
md = MD.replace("@@SHIM@@", rd("scasa_r.c")).replace("@@STEP2@@", rd("scasa_quant_step2_patch.R")).replace("@@STEP3@@", rd("scasa_quant_step3_patch.R"))
open(os.path.join(ROOT, "INTEGRATION.md"), "w").write(md)
print("INTEGRATION.md", len(md))
