// scasa_reads.h — what the read-side parsers (alevin bfh.txt, bustools text) hand to the quant path.
#pragma once
#include <cstdint>
#include <cstdlib>
#include <string>
#include <vector>

struct scasa_bfh {
    std::vector<char> text;                    // the whole file (bustools reader)
    char *raw = nullptr;                       // the whole file (bfh reader: malloc, no zero fill, read in slices by threads)
    int64_t raw_size = 0;
    ~scasa_bfh() { free(raw); }
    scasa_bfh() = default;
    scasa_bfh(const scasa_bfh &) = delete;
    scasa_bfh &operator=(const scasa_bfh &) = delete;
    std::vector<const char *> line;            // line starts; line[k+1] - line[k] spans line k incl. its newline
    int64_t n_tx = 0, n_cb = 0, n_eq = 0;
    std::vector<int64_t> eq_off;               // [n_eq+1]
    std::vector<int32_t> eq_tx;
    std::vector<int32_t> cell_of_cb;           // [n_cb] rank among the barcodes that occur, -1 otherwise
    std::vector<int32_t> cells;                // barcode index of each cell, in byte order of the sequences
    std::vector<int64_t> cell_ptr;             // [n_cells+1]
    std::vector<int32_t> eq_id, eq_count;
    // bustools front end: names are collected strings instead of lines of `text`
    std::vector<std::string> tx_names_v, cell_names_v;
};

