## scasa_quant_step3_patch.R — what replaces line 48 of scasa_quant_step3_v1.0.0.R
##     gene_count = scasa_genemat(isoform_count, genes.tx.map.all.final)
## when step 2 ran with libscasa_cuda and the gene map loaded (scasa_quant_step2_patch.R): the gene sums were taken on
## the GPU next to the isoform estimates; what is left of scasa_genemat (step3:32-46) is its ROW LAYOUT after the
## step2:167 filter: genes with one surviving isoform first (mat1), then the others (mat2), each group in the order
## tapply gives (sort() of the gene names in this session's collation).  Lines 21-22 and 49-52 of the script stay.
fused <- paste0(workdir, "/scasa_gene_sums_all.RDS")
if (file.exists(paste0(workdir, "/scasa_gene_expression.txt")) && !file.exists(fused)) {
  quit(save = "no")                                               # file mode of the step-2 patch already wrote the gene file
} else if (file.exists(fused)) {
  g <- readRDS(fused)
  kept   <- g$kept[!is.na(g$kept) & g$kept > 0]
  gnames <- sort(names(kept))                                     # tapply's group order (step3:36)
  gnames <- c(gnames[kept[gnames] == 1], gnames[kept[gnames] > 1])   # rbind(mat1, t(mat2)), step3:43
  gene_count <- g$gene[gnames, , drop = FALSE]
} else {
  gene_count <- scasa_genemat(isoform_count, genes.tx.map.all.final)   # step 2 ran without the library: unchanged
}
