// scasa_bus.cpp — kallisto | bustools output -> the arrays the quant path takes (the second mapper).
//
// Replaces scasa_quant_step1_2_v1.0.0.R, which turns
//     output.correct.sort.bus.final.txt   barcode, UMI, eqclass, count, barcode+UMI  (tab separated, no header;
//                                          `bustools text` plus an awk column, SCASA_WRAP_V1.0.0.pl:383-385)
//     matrix.ec                            eqclass id <tab> comma-separated transcript numbers (0-based)
//     transcripts.txt                      transcript names, one per line
// into the same exploded data.table step1_1 makes for alevin.  Same decisions here, no exploded table:
//
//   cells        barcodes with more than libsize_thres records (:30-32; the wrapper passes 200)
//   molecules    a (barcode, UMI) seen once is kept (:56-57).  Seen several times: keep the record with the
//                largest Count if it is unique (:62-69); if several records tie on Count, keep the one with the
//                largest eqclass size among them if that is unique, otherwise drop the molecule (:71-88)
//   counts       every kept record counts 1 (:100); Count of a (cell, eqclass) = its number of kept records
//                (:105-116, table(eqClass))
//
// The result goes into the same handle as a bfh read (scasa_bfh_sizes / _read / _close): cells in byte order
// (step2:59-60), per cell the (eqclass, count) pairs in ascending eqclass line index.  Host-only.
#include "scasa_cuda.h"
#include "scasa_host.h"
#include "scasa_reads.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <numeric>
#include <string>
#include <unordered_map>
#include <vector>

namespace {

int bus_fail(int code, const std::string &m)
{
    scasa::set_global_error(m);
    return code;
}

bool read_file(const std::string &path, std::vector<char> &out)
{
    FILE *fp = fopen(path.c_str(), "rb");
    if (!fp) return false;
    fseek(fp, 0, SEEK_END);
    const long long size = ftell(fp);
    fseek(fp, 0, SEEK_SET);
    out.resize((size_t)size + 1);
    const size_t got = size ? fread(out.data(), 1, (size_t)size, fp) : 0;
    fclose(fp);
    out[(size_t)size] = '\n';
    return (long long)got == size;
}

struct Rec { const char *bc; const char *umi; int32_t bc_len, umi_len, ec_line, count; };

inline int cmp_span(const char *a, int la, const char *b, int lb)
{
    const int r = memcmp(a, b, (size_t)std::min(la, lb));
    return r ? r : la - lb;
}

}  // namespace

extern "C" int scasa_bus_open(const char *bus_txt, const char *matrix_ec, const char *transcripts_txt, int64_t libsize_thres,
                              scasa_bfh **out)
{
    if (!bus_txt || !matrix_ec || !transcripts_txt || !out) return bus_fail(SCASA_ERR_ARG, "null argument");
    *out = nullptr;
    try {
        std::unique_ptr<scasa_bfh> f(new scasa_bfh());
        std::vector<char> ec_text, tx_text;
        if (!read_file(bus_txt, f->text)) return bus_fail(SCASA_ERR_ARG, std::string("cannot read ") + bus_txt);
        if (!read_file(matrix_ec, ec_text)) return bus_fail(SCASA_ERR_ARG, std::string("cannot read ") + matrix_ec);
        if (!read_file(transcripts_txt, tx_text)) return bus_fail(SCASA_ERR_ARG, std::string("cannot read ") + transcripts_txt);

        // transcripts.txt: names, numbered from 0 (:38-42; read.table takes the first field of a line)
        for (const char *p = tx_text.data(), *end = p + tx_text.size() - 1; p < end;) {
            const char *nl = (const char *)memchr(p, '\n', (size_t)(end - p));
            const char *e = nl ? nl : end;
            const char *q = e;
            if (q > p && q[-1] == '\r') q--;
            const char *sp = p;
            while (sp < q && *sp != ' ' && *sp != '\t') sp++;
            if (sp > p) f->tx_names_v.emplace_back(p, sp);
            p = e + 1;
        }
        f->n_tx = (int64_t)f->tx_names_v.size();

        // matrix.ec: one eqclass per line; the id need not be the line number, so ids are mapped to lines
        std::unordered_map<int64_t, int32_t> ec_line;
        f->eq_off.assign(1, 0);
        for (const char *p = ec_text.data(), *end = p + ec_text.size() - 1; p < end;) {
            const char *nl = (const char *)memchr(p, '\n', (size_t)(end - p));
            const char *e = nl ? nl : end;
            if (e > p) {
                char *q = nullptr;
                const long long id = strtoll(p, &q, 10);
                if (q == p) return bus_fail(SCASA_ERR_ARG, "matrix.ec: bad eqclass id");
                while (q < e && (*q == '\t' || *q == ' ')) q++;
                while (q < e) {
                    char *r = nullptr;
                    const long long t = strtoll(q, &r, 10);
                    if (r == q) break;
                    if (t < 0 || t >= f->n_tx) return bus_fail(SCASA_ERR_ARG, "matrix.ec: transcript number out of range");
                    f->eq_tx.push_back((int32_t)t);
                    q = r;
                    if (q < e && *q == ',') q++;
                }
                ec_line[id] = (int32_t)(f->eq_off.size() - 1);
                f->eq_off.push_back((int64_t)f->eq_tx.size());
            }
            p = e + 1;
        }
        f->n_eq = (int64_t)f->eq_off.size() - 1;

        // the bus records
        std::vector<Rec> rec;
        for (const char *p = f->text.data(), *end = p + f->text.size() - 1; p < end;) {
            const char *nl = (const char *)memchr(p, '\n', (size_t)(end - p));
            const char *e = nl ? nl : end;
            if (e > p) {
                const char *t1 = (const char *)memchr(p, '\t', (size_t)(e - p));
                const char *t2 = t1 ? (const char *)memchr(t1 + 1, '\t', (size_t)(e - t1 - 1)) : nullptr;
                const char *t3 = t2 ? (const char *)memchr(t2 + 1, '\t', (size_t)(e - t2 - 1)) : nullptr;
                if (!t3) return bus_fail(SCASA_ERR_ARG, "bus text: a line with fewer than four tab-separated fields");
                Rec r;
                r.bc = p; r.bc_len = (int32_t)(t1 - p);
                r.umi = t1 + 1; r.umi_len = (int32_t)(t2 - t1 - 1);
                const long long ec = strtoll(t2 + 1, nullptr, 10);
                r.count = (int32_t)strtoll(t3 + 1, nullptr, 10);
                auto it = ec_line.find(ec);
                r.ec_line = it == ec_line.end() ? -1 : it->second;
                rec.push_back(r);
            }
            p = e + 1;
        }
        // group by (barcode, UMI): bustools sort leaves the file in that order; sort (stably) if it is not
        auto key_less = [&](const Rec &a, const Rec &b) {
            const int c = cmp_span(a.bc, a.bc_len, b.bc, b.bc_len);
            return c ? c < 0 : cmp_span(a.umi, a.umi_len, b.umi, b.umi_len) < 0;
        };
        if (!std::is_sorted(rec.begin(), rec.end(), key_less)) std::stable_sort(rec.begin(), rec.end(), key_less);

        std::vector<std::pair<int32_t, int32_t>> cell_ec;      // kept records of the current cell: eqclass line
        std::vector<int32_t> ecs;
        f->cell_ptr.assign(1, 0);
        const size_t n = rec.size();
        for (size_t b0 = 0; b0 < n;) {
            size_t b1 = b0 + 1;
            while (b1 < n && !cmp_span(rec[b1].bc, rec[b1].bc_len, rec[b0].bc, rec[b0].bc_len)) b1++;
            if ((int64_t)(b1 - b0) > libsize_thres) {                         // libsize = records of the barcode  :30-32
                ecs.clear();
                for (size_t u0 = b0; u0 < b1;) {
                    size_t u1 = u0 + 1;
                    while (u1 < b1 && !cmp_span(rec[u1].umi, rec[u1].umi_len, rec[u0].umi, rec[u0].umi_len)) u1++;
                    int keep = -1;
                    if (u1 - u0 == 1) keep = (int)u0;                          // :56-57
                    else {
                        int32_t mx = rec[u0].count;
                        for (size_t k = u0; k < u1; k++) mx = std::max(mx, rec[k].count);
                        int n_mx = 0;
                        for (size_t k = u0; k < u1; k++) n_mx += rec[k].count == mx;
                        if (n_mx == 1) {                                       // case 1  :62-69
                            for (size_t k = u0; k < u1; k++) if (rec[k].count == mx) keep = (int)k;
                        } else {                                               // case 2  :71-88
                            int64_t best = -1;
                            int n_best = 0;
                            for (size_t k = u0; k < u1; k++) {
                                if (rec[k].count != mx) continue;
                                const int64_t sz = rec[k].ec_line >= 0 ? f->eq_off[(size_t)rec[k].ec_line + 1] - f->eq_off[(size_t)rec[k].ec_line] : -1;
                                if (sz > best) { best = sz; n_best = 1; keep = (int)k; }
                                else if (sz == best) n_best++;
                            }
                            if (n_best != 1) keep = -1;                        // case 3: the molecule is dropped
                        }
                    }
                    if (keep >= 0 && rec[(size_t)keep].ec_line >= 0) ecs.push_back(rec[(size_t)keep].ec_line);
                    u0 = u1;
                }
                if (!ecs.empty()) {                                            // table(eqClass)  :105-116
                    std::sort(ecs.begin(), ecs.end());
                    for (size_t i = 0; i < ecs.size();) {
                        size_t j = i;
                        while (j < ecs.size() && ecs[j] == ecs[i]) j++;
                        f->eq_id.push_back(ecs[i]);
                        f->eq_count.push_back((int32_t)(j - i));
                        i = j;
                    }
                    f->cell_names_v.emplace_back(rec[b0].bc, rec[b0].bc + rec[b0].bc_len);
                    f->cell_ptr.push_back((int64_t)f->eq_id.size());
                }
            }
            b0 = b1;
        }
        f->cells.resize(f->cell_names_v.size());
        std::iota(f->cells.begin(), f->cells.end(), 0);
        if (f->tx_names_v.empty()) return bus_fail(SCASA_ERR_ARG, "transcripts.txt is empty");
        if (f->cell_names_v.empty()) {          // keep the accessors on the string path
            f->cells.clear();
            f->cell_ptr.assign(1, 0);
        }
        f->text.clear(); f->text.shrink_to_fit();
        *out = f.release();
        return SCASA_OK;
    } catch (const std::bad_alloc &) {
        return bus_fail(SCASA_ERR_NOMEM, "host allocation failed");
    } catch (const std::exception &ex) {
        return bus_fail(SCASA_ERR_ARG, ex.what());
    }
}
