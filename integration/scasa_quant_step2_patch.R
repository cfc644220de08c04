## scasa_quant_step2_patch.R — what replaces lines 88-136 of scasa_quant_step2_v1.0.0.R (the %dopar% loops over
## crpcount_td and estim_chunk) when libscasa_cuda is used.  Everything before (loading Sample_eqClass.RData and the
## X-matrix, step2:35-59) and after (saveRDS / write.table, step2:169-170) stays as it is.  See INTEGRATION.md.
dyn.load("scasa_r.so")

flat <- function(L) list(q_off = c(0L, cumsum(sapply(L, nrow))), p_off = c(0L, cumsum(sapply(L, ncol))),
                         x_off = as.numeric(c(0, cumsum(sapply(L, length)))), x = unlist(L, use.names = FALSE))

DM0 <- lapply(CRP, ccrpfun, 30)                                   # unchanged (step2:127); or keep DM0 = CCRP (INTEGRATION.md)
txAll <- unique(sample_eqclass$Transcript)                        # any consistent transcript dictionary
crp <- c(flat(CRP), list(
  col_tx    = match(unlist(lapply(CRP, colnames)), txAll, nomatch = length(txAll) + 1L) - 1L,
  row_pat   = as.raw(unlist(lapply(CRP, function(m) t(do.call(rbind, strsplit(rownames(m), "")) == "1")))),
  name_rank = rank(names(CRP), ties.method = "first") - 1L))      # R's own collation (Rsource:484)
ctx <- .Call("scasa_r_create", flat(DM0), crp, 0L)

setorder(sample_eqclass, Barcode)                                 # step2:59
allCells <- unique(sample_eqclass$Barcode)
eq  <- unique(sample_eqclass[, .(eqClass, Transcript)])           # the eqclass dictionary
setorder(eq, eqClass)
ce  <- unique(sample_eqclass[, .(Barcode, eqClass, Count)])       # one row per (cell, eqclass)
ids <- sort(unique(eq$eqClass))
iso <- .Call("scasa_r_quant", ctx,
             as.numeric(c(0, cumsum(table(factor(ce$Barcode, allCells))))),
             match(ce$eqClass, ids) - 1L, ce$Count,
             as.numeric(c(0, cumsum(table(factor(eq$eqClass, ids))))),
             match(eq$Transcript, txAll) - 1L, core)
dimnames(iso) <- list(unlist(lapply(DM0, colnames)), allCells)    # step2:163-166 (already tx x cells, barcode order)
isoform_count <- iso[which(rowSums(iso) > 0), ]                   # step2:167
