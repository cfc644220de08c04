## scasa_quant_step2_patch.R — what replaces lines 88-167 of scasa_quant_step2_v1.0.0.R (the two %dopar% loops over
## crpcount_td and estim_chunk, the serial ccrpfun recompute, the stacking and the row filter) when libscasa_cuda is used.
## Everything before (loading Sample_eqClass.RData and the X-matrix, step2:35-59) stays as it is; step2:169-170
## (saveRDS / write.table) stay too in matrix mode and are replaced by the library's writer in file mode; in sparse mode
## saveRDS stays and write.table needs as.matrix(isoform_count) (or SCASA_FILES=1 for the text file).  See INTEGRATION.md.
dyn.load("scasa_r.so")
n_gpus    <- as.integer(Sys.getenv("SCASA_GPUS", "1"))          # GPUs of this box to shard the chunks over
file_mode <- Sys.getenv("SCASA_FILES", "0") == "1"              # 1: write the two text files directly, no dense matrix in R
sparse    <- Sys.getenv("SCASA_SPARSE", "0") == "1"             # 1: isoform_count / gene matrix as Matrix::dgCMatrix

flat <- function(L) list(q_off = c(0L, cumsum(sapply(L, nrow))), p_off = c(0L, cumsum(sapply(L, ncol))),
                         x_off = as.numeric(c(0, cumsum(sapply(L, length)))), x = unlist(L, use.names = FALSE))
nl   <- function(x) paste0(paste(x, collapse = "\n"), "\n")

## step2:126-127: `DM0 <- lapply(CRP, ccrpfun, 30)` is minutes of serial svd + kmeans; the shipped CCRP is its result.
## The library checks that (every CCRP column is the mean of its member CRP columns, zero-sum columns dropped, all
## condition numbers < 30, merged blocks were not separated) and only otherwise is the recompute run.
crp_flat <- flat(CRP)
dm0_off  <- c(0L, cumsum(sapply(CCRP, ncol)))
member   <- unlist(lapply(seq_along(CRP), function(k) {
  mem <- strsplit(as.character(colnames(CCRP[[k]])), " ", fixed = TRUE)
  idx <- rep(seq_along(mem), lengths(mem)); names(idx) <- unlist(mem)
  m <- idx[colnames(CRP[[k]])]
  ifelse(is.na(m), -1L, dm0_off[k] + m - 1L)
}), use.names = FALSE)
DM0 <- if (.Call("scasa_r_dm0_is_fixed_point", crp_flat, flat(CCRP), as.integer(member), 30)) CCRP else lapply(CRP, ccrpfun, 30)

## transcript ids: CRP's own dictionary first, so that every CRP transcript has an id of its own whether or not the
## sample contains it; transcripts the sample has and CRP does not follow
txCRP <- unlist(lapply(CRP, colnames), use.names = FALSE)
txAll <- unique(c(txCRP, sample_eqclass$Transcript))
crp <- c(crp_flat, list(
  col_tx    = match(txCRP, txAll) - 1L,
  row_pat   = as.raw(unlist(lapply(CRP, function(m) t(do.call(rbind, strsplit(rownames(m), "")) == "1")))),
  name_rank = rank(names(CRP), ties.method = "first") - 1L))      # R's own collation (Rsource:484)
ctx <- .Call("scasa_r_create", flat(DM0), crp, n_gpus)

setorder(sample_eqclass, Barcode)                                 # step2:59
allCells <- unique(sample_eqclass$Barcode)
eq  <- unique(sample_eqclass[, .(eqClass, Transcript)])           # the eqclass dictionary
setorder(eq, eqClass)
ce  <- unique(sample_eqclass[, .(Barcode, eqClass, Count)])       # one row per (cell, eqclass), cells in barcode order
ids <- sort(unique(eq$eqClass))
cell_ptr <- as.numeric(c(0, cumsum(table(factor(ce$Barcode, allCells)))))
eq_off   <- as.numeric(c(0, cumsum(table(factor(eq$eqClass, ids)))))
eq_id    <- match(ce$eqClass, ids) - 1L
eq_tx    <- match(eq$Transcript, txAll) - 1L

## gene map for step 3 (scasa_quant_step3_v1.0.0.R:34-36), over ALL isoform rows: an isoform that step2:167 drops is
## all zero, so it changes no gene sum; which genes are written, and in which order, is decided after the filter
tx     <- unlist(lapply(DM0, function(m) as.character(colnames(m))), use.names = FALSE)
have_g <- exists("genes.tx.map.all.final")                        # load(txgroup) before this patch to fuse step 3
if (have_g) {
  genes      <- genes.tx.map.all.final[tx]
  gene_names <- unique(genes[!is.na(genes)])
  gene_of    <- match(genes, gene_names) - 1L; gene_of[is.na(gene_of)] <- -1L
  gene_rank  <- rank(gene_names, ties.method = "first") - 1L      # sort() order of the names in this session's collation
} else { gene_of <- NULL; gene_names <- character(0); gene_rank <- NULL }

if (file_mode) {                                                  # step2:165-170 + step3:48-52 inside the library
  iso_txt <- paste0(workdir, "/scasa_isoform_counts.txt")
  colsum <- .Call("scasa_r_quant_files", ctx, cell_ptr, eq_id, ce$Count, eq_off, eq_tx, core, gene_of, length(gene_names),
                  iso_txt, if (have_g) paste0(workdir, "/scasa_gene_expression.txt") else NULL,
                  nl(tx), if (have_g) nl(gene_names) else NULL, gene_rank, nl(allCells))
  file.copy(iso_txt, paste0(workdir, "/scasa_isoform_expression.txt"), overwrite = TRUE)   # the name README.md:80 uses
} else if (sparse) {                                              # the same matrices, sparse: the form Seurat & co. read anyway
  library(Matrix)
  res <- .Call("scasa_r_quant_sparse", ctx, cell_ptr, eq_id, ce$Count, eq_off, eq_tx, core, gene_of, length(gene_names))
  dgc <- function(m, rn) new("dgCMatrix", p = m[[1]], i = m[[2]], x = m[[3]], Dim = c(length(rn), length(allCells)),
                             Dimnames = list(rn, allCells))
  isoform_count <- dgc(res[[1]], tx)[which(res[[3]] > 0), ]       # step2:167
  if (have_g) {
    kept <- tapply(res[[3]] > 0, factor(genes, gene_names), sum)
    saveRDS(list(gene = dgc(res[[2]], gene_names), kept = kept), paste0(workdir, "/scasa_gene_sums_all.RDS"))
  }
} else {
  res <- .Call("scasa_r_quant", ctx, cell_ptr, eq_id, ce$Count, eq_off, eq_tx, core, gene_of, length(gene_names))
  iso <- res[[1]]
  dimnames(iso) <- list(tx, allCells)                             # step2:163-166 (already tx x cells, barcode order)
  isoform_count <- iso[which(res[[3]] > 0), ]                     # step2:167: res[[3]] = rowSums(iso), summed on the GPU
  if (have_g) {                                                   # hand-off to the step-3 patch
    gene_all <- res[[2]]; dimnames(gene_all) <- list(gene_names, allCells)
    kept <- tapply(res[[3]] > 0, factor(genes, gene_names), sum)  # surviving isoforms per gene
    saveRDS(list(gene = gene_all, kept = kept), paste0(workdir, "/scasa_gene_sums_all.RDS"))
  }
}
