#!/usr/bin/env Rscript
# tools/make_golden.R — pin parity to R itself.
#
# Runs the REFERENCE's own functions (sourced, unmodified, from the reference checkout) on the committed
# 48-cell case and writes plain-text dumps that tests/test_r_golden.py compares with the CPU oracle, the
# string-level literal and the CUDA library.  Until someone with R runs this, every parity claim in this
# repository means "equal to our reading of the R code" (DESIGN.md: PARITY UNPINNED).
#
#   python tools/export_golden_inputs.py                       # writes tests/golden/r_case/*.tsv
#   Rscript tools/make_golden.R scasa=/path/to/scasa [indir=tests/golden/r_case] [outdir=tests/golden/r_dumps]
#                               [hamming=1] [hamming_stride=1] [ccrpfun=0]
#   python -m pytest tests/test_r_golden.py                    # and -m gpu on a B200
#
# scasa= is the directory that holds SCRIPTS/ and REFERENCE/ (the reference repository's scasa/).
# Needs R >= 3.6 with data.table.  About 20 s per cell for crpcount_td (48 cells), minutes for the rest;
# the Hamming sweep is ~400k calls of hamming_correction (hamming_stride=10 takes a tenth).
#
# What is dumped (all indices 1-based; numbers with 17 significant digits, which round-trips doubles):
#   session.txt           R version, collation locale, data.table version
#   y_counts.tsv          crpcount_td(myeq) per cell (Rsource:416-557):  cell, global row, value   (non-zeros)
#   estim_colnames.txt    colnames of estim_chunk's result (Rsource:1042)
#   estim_chunk.tsv       estim_chunk(Y, CCRP, maxiter.bt = 1000) on all 48 cells as one chunk (step2:132): cell, col, value
#   fixed_aem.tsv         4 outer x 8 inner iterations, no break, of the reference's estim.BETA / estim.X.em sequenced as
#                         at Rsource:1011-1024, for every non-poor block with nrow > 1: cell, col, value
#   iso_rownames.txt      rownames of the isoform matrix after step2:165-167 (order, transpose, rowSums > 0)
#   gene_rownames.txt     rownames of scasa_genemat's result (step3:32-46)
#   gene.tsv              gene row, cell, value
#   hamming.tsv           hamming_correction (Rsource:562-648) on every case of hamming_cases.tsv: block, observed, assigned
#   ccrpfun.txt           (ccrpfun=1) max |lapply(CRP, ccrpfun, 30) - CCRP| and whether the column names agree (step2:127)

args <- commandArgs(trailingOnly = TRUE)
opt <- list(scasa = NA, indir = "tests/golden/r_case", outdir = "tests/golden/r_dumps", hamming = "1",
            hamming_stride = "1", ccrpfun = "0")
for (a in args) {
  kv <- strsplit(a, "=", fixed = TRUE)[[1]]
  if (length(kv) == 2) opt[[kv[1]]] <- kv[2]
}
if (is.na(opt$scasa)) stop("usage: Rscript tools/make_golden.R scasa=<dir with SCRIPTS/ and REFERENCE/> [indir=] [outdir=]")

suppressPackageStartupMessages(library(data.table))
options(stringsAsFactors = FALSE)
source(file.path(opt$scasa, "SCRIPTS", "Scasa_Rsource.R"))
load(file.path(opt$scasa, "REFERENCE", "Xmatrix_hg38_alevin.RData"))               # CRP, CCRP, CRPCOUNT (step2:39)
load(file.path(opt$scasa, "REFERENCE", "isoform_groups_UCSC_hg38_alevin.RData"))   # genes.tx.map.all.final (step3:22)
dir.create(opt$outdir, recursive = TRUE, showWarnings = FALSE)
out <- function(name) file.path(opt$outdir, name)
num <- function(x) formatC(x, digits = 17, format = "g")

writeLines(c(R.version.string, paste("LC_COLLATE", Sys.getlocale("LC_COLLATE")),
             paste("data.table", as.character(packageVersion("data.table")))), out("session.txt"))

# scasa_genemat / parasum live in the step-3 script, which runs when sourced: take only the two definitions
for (e in as.list(parse(file.path(opt$scasa, "SCRIPTS", "scasa_quant_step3_v1.0.0.R")))) {
  if (is.call(e) && length(e) >= 3 && is.name(e[[1]]) && as.character(e[[1]]) %in% c("=", "<-") && is.name(e[[2]]) &&
      as.character(e[[2]]) %in% c("parasum", "scasa_genemat")) eval(e)
}

# ---- 1. crpcount_td per cell, exactly as step2:59-60 and :93-97 call it ------------------------------------
sample_eqclass <- fread(file.path(opt$indir, "sample_eqclass.tsv"), sep = "\t", header = TRUE,
                        colClasses = c("character", "character", "integer", "integer"))
sample_eqclass[, Weight := 1]                       # step1_1 writes Weight = 1; crpcount_td overwrites it (:446)
setorder(sample_eqclass, Barcode)
allCells <- unique(sample_eqclass$Barcode)
n <- length(allCells)
cat("cells:", n, "\n")
res_chunk <- vector("list", n)
con <- file(out("y_counts.tsv"), "w")
writeLines("cell\trow\tvalue", con)
for (ci in seq_len(n)) {
  myeq <- sample_eqclass[which(sample_eqclass$Barcode == allCells[ci]), ]
  mycrp <- crpcount_td(myeq)
  res_chunk[[ci]] <- mycrp
  v <- unlist(mycrp, use.names = FALSE)             # length sum(q): the blocks' count vectors in list order
  nz <- which(v != 0)
  if (length(nz) > 0) writeLines(paste(ci, nz, num(v[nz]), sep = "\t"), con)
  cat(" crpcount_td cell", ci, "non-zero rows", length(nz), "\n")
}
close(con)

# ---- 2. Y as step2:100-110 builds it, then estim_chunk as at step2:132 ---------------------------------------
Y <- res_chunk[[1]]
for (i in 2:n) Y <- Map(cbind, Y, res_chunk[[i]])
Y <- lapply(Y, function(x) { colnames(x) <- allCells; x })
est <- estim_chunk(Y, CCRP, maxiter.bt = 1000)
writeLines(colnames(est), out("estim_colnames.txt"))
nz <- which(est != 0, arr.ind = TRUE)
write.table(data.frame(cell = nz[, 1], col = nz[, 2], value = num(est[nz])), out("estim_chunk.tsv"),
            sep = "\t", quote = FALSE, row.names = FALSE)

# ---- 3. fixed-iteration AEM with the reference's estim.BETA / estim.X.em (no break; maxerr = -1 never holds) ----
cos_sim <- function(a, b) sum(a * b) / sqrt(sum(a^2) * sum(b^2))
p_k <- sapply(CCRP, ncol)
q_k <- sapply(CCRP, nrow)
col_off <- c(0, cumsum(p_k))
is_poor <- function(X0) {                            # the scan of Rsource:957-971
  if (ncol(X0) != 2) return(FALSE)
  for (i in 1:ncol(X0)) { sim <- apply(X0, 2, cos_sim, X0[, i]); if (sum(sim < 0.99)) return(FALSE) }
  TRUE
}
con <- file(out("fixed_aem.tsv"), "w")
writeLines("cell\tcol\tvalue", con)
for (m in which(q_k > 1)) {
  X0 <- CCRP[[m]]
  if (is_poor(X0)) next
  Ymat <- Y[[m]]
  if (sum(Ymat) == 0) next
  X.est0 <- X0
  Ysum <- colSums(Ymat)
  beta0 <- rep(Ysum / ncol(X0), rep(ncol(X0), length(Ysum)))
  for (i in 1:4) {
    BT <- estim.BETA(X.est0, Ymat, beta0 = beta0, maxiter = 8, maxerr = -1, modify = TRUE)
    X.est <- estim.X.em(X.est0, BT, Ymat)
    X.est0 <- X.est
    beta0 <- c(t(BT))
  }
  nz <- which(BT != 0, arr.ind = TRUE)
  if (nrow(nz) > 0) writeLines(paste(nz[, 1], col_off[m] + nz[, 2], num(BT[nz]), sep = "\t"), con)
}
close(con)

# ---- 4. assemble (step2:165-167) and gene sums (step3:48) -----------------------------------------------------
isoform_count <- est[order(rownames(est)), ]
isoform_count <- t(isoform_count)
isoform_count <- isoform_count[which(rowSums(isoform_count) > 0), ]
writeLines(rownames(isoform_count), out("iso_rownames.txt"))
gene_count <- scasa_genemat(isoform_count, genes.tx.map.all.final)
writeLines(rownames(gene_count), out("gene_rownames.txt"))
nz <- which(gene_count != 0, arr.ind = TRUE)
cell_of <- match(colnames(gene_count), allCells)
write.table(data.frame(gene = nz[, 1], cell = cell_of[nz[, 2]], value = num(gene_count[nz])), out("gene.tsv"),
            sep = "\t", quote = FALSE, row.names = FALSE)

# ---- 5. hamming_correction on every case of the sweep ---------------------------------------------------------
if (opt$hamming == "1" && file.exists(file.path(opt$indir, "hamming_cases.tsv"))) {
  cases <- fread(file.path(opt$indir, "hamming_cases.tsv"), sep = "\t", header = TRUE, colClasses = c("integer", "character"))
  pick <- seq(1, nrow(cases), by = as.integer(opt$hamming_stride))
  con <- file(out("hamming.tsv"), "w")
  writeLines("block\tobserved\tassigned", con)
  last <- -1
  for (i in pick) {
    k <- cases$block[i]
    if (k != last) { crp <- CRP[[k]]; crp1 <- data.frame(pat = rownames(crp), crp); last <- k }   # Rsource:532
    raw4 <- data.frame(pat = cases$pat[i], sample1 = 1)
    r <- hamming_correction(raw4, crp1, hamming_dist = 1)                                          # Rsource:534
    writeLines(paste(k, cases$pat[i], r$pat[1], sep = "\t"), con)
  }
  close(con)
}

# ---- 6. is the shipped CCRP what step2:127 recomputes? ---------------------------------------------------------
if (opt$ccrpfun == "1") {
  DM0 <- lapply(CRP, ccrpfun, 30)
  dmax <- 0
  same_names <- TRUE
  for (k in seq_along(DM0)) {
    if (!identical(dim(DM0[[k]]), dim(CCRP[[k]]))) { dmax <- Inf; same_names <- FALSE; next }
    dmax <- max(dmax, max(abs(DM0[[k]] - CCRP[[k]])))
    if (!identical(as.character(colnames(DM0[[k]])), as.character(colnames(CCRP[[k]])))) same_names <- FALSE
  }
  writeLines(c(paste("max_abs_diff", num(dmax)), paste("same_colnames", same_names)), out("ccrpfun.txt"))
}
cat("done:", opt$outdir, "\n")
