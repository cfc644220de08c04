/*
 * scasa_oracle_counts.c — CPU restatement of crpcount_td + hamming_correction.  TEST INFRASTRUCTURE
 * (same rules as scasa_oracle.c: only tests/, smoke() and bench.py's CPU legs may load it).
 *
 * PARITY UNPINNED: no R here and the reference ships no fixtures; in particular the tie-break
 * of hamming_correction (Scasa_Rsource.R:616-617, :625-626) relies on R coercing a data.frame
 * row to character and is restated from R's documented formatReal rules.
 *
 * The restatement keeps the reference's structure: it works per CELL, walks every block of the
 * X-matrix for that cell (Rsource:453), rebuilds the per-block eqclass list, scores candidate
 * blocks, builds the pattern strings, sums by pattern and corrects by Hamming distance.  The
 * only liberty taken is a per-cell transcript -> table-rows index instead of R's vector scans
 * (`rawcount$Transcript == tx`), which changes no result.
 *
 * Line numbers: /root/reference/scasa/SCRIPTS/Scasa_Rsource.R.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

typedef struct {
    int B;
    const int32_t *q_off, *p_off;
    const int64_t *x_off;
    const double *x;
    const int32_t *col_tx;
    const uint8_t *row_pat;
    const int32_t *name_rank;
    int n_tx;             /* transcript ids are < n_tx */
    uint8_t *supported;   /* per column: colSums(crp) > 0 */
    /* t2c (:423-427): for a transcript, the blocks whose NAME lists it, sorted by name */
    int32_t *t2c_off;     /* [n_tx+1] */
    int32_t *t2c_block;
} oracle_crp;

static const oracle_crp *g_sort_crp;
static int cmp_block_by_name(const void *a, const void *b)
{
    int32_t ka = *(const int32_t *)a, kb = *(const int32_t *)b;
    return g_sort_crp->name_rank[ka] - g_sort_crp->name_rank[kb];
}

oracle_crp *oracle_crp_create(int B, const int32_t *q_off, const int32_t *p_off, const int64_t *x_off,
                              const double *x, const int32_t *col_tx, const uint8_t *row_pat,
                              const int32_t *name_rank, int n_tx)
{
    oracle_crp *c = (oracle_crp *)calloc(1, sizeof *c);
    c->B = B; c->q_off = q_off; c->p_off = p_off; c->x_off = x_off; c->x = x;
    c->col_tx = col_tx; c->row_pat = row_pat; c->name_rank = name_rank; c->n_tx = n_tx;
    int ncol = p_off[B];
    c->supported = (uint8_t *)calloc(ncol > 0 ? ncol : 1, 1);
    c->t2c_off = (int32_t *)calloc((size_t)n_tx + 2, sizeof(int32_t));
    c->t2c_block = (int32_t *)calloc(ncol > 0 ? ncol : 1, sizeof(int32_t));
    for (int k = 0; k < B; k++) {
        int q = q_off[k + 1] - q_off[k], p = p_off[k + 1] - p_off[k];
        for (int j = 0; j < p; j++) {
            long double s = 0;
            for (int r = 0; r < q; r++) s += x[x_off[k] + (int64_t)j * q + r];
            c->supported[p_off[k] + j] = (double)s > 0;
            int t = col_tx[p_off[k] + j];
            if (t >= 0 && t < n_tx) c->t2c_off[t + 1]++;
        }
    }
    for (int t = 0; t < n_tx; t++) c->t2c_off[t + 1] += c->t2c_off[t];
    int32_t *cur = (int32_t *)malloc(sizeof(int32_t) * ((size_t)n_tx + 1));
    memcpy(cur, c->t2c_off, sizeof(int32_t) * ((size_t)n_tx + 1));
    for (int k = 0; k < B; k++)
        for (int j = p_off[k]; j < p_off[k + 1]; j++) {
            int t = col_tx[j];
            if (t >= 0 && t < n_tx) c->t2c_block[cur[t]++] = k;
        }
    free(cur);
    g_sort_crp = c;   /* tapply(names(t2c), t2c, c) keeps name order; candidates are sort()ed at :484 */
    for (int t = 0; t < n_tx; t++)
        qsort(c->t2c_block + c->t2c_off[t], (size_t)(c->t2c_off[t + 1] - c->t2c_off[t]), sizeof(int32_t),
              cmp_block_by_name);
    return c;
}

void oracle_crp_free(oracle_crp *c)
{
    if (!c) return;
    free(c->supported); free(c->t2c_off); free(c->t2c_block); free(c);
}

/* ---- format(x) with 7 significant digits on a numeric column, then as.numeric(): what
 * as.matrix.data.frame does to `current` inside apply() at :617.  R's formatReal: each value
 * needs nsig significant digits (<= 7, trailing zeros dropped) and has decimal exponent kp;
 * fixed notation shows max(nsig - kp - 1) decimals and is used unless wider than scientific. */
void oracle_format_roundtrip(const double *v, int n, double *back, uint8_t *is_zero_text)
{
    int have = 0, neg = 0, rgt = 0, mxsl = 1, mxe = -99999, mne = 99999, mxns = 1;
    int first = 1;
    for (int i = 0; i < n; i++) {
        double a = fabs(v[i]);
        if (!isfinite(a)) continue;
        int kp = 0, nsig = 1;
        if (a > 0) {
            char m[40];
            snprintf(m, sizeof m, "%.6e", a);              /* 7 significant digits, correctly rounded */
            kp = atoi(strchr(m, 'e') + 1);
            nsig = 7;
            while (nsig > 1 && m[nsig] == '0') nsig--;     /* m = "d.dddddd": digit i (1-based) at m[i] for i>=2 */
        }
        int left = kp + 1;
        if (kp > 0 && a < pow(10.0, kp)) left--;           /* rounding to 7 digits widened the number */
        int sleft = (v[i] < 0) + (left <= 0 ? 1 : left);
        int right = nsig - left;
        if (v[i] < 0) neg = 1;
        if (first) { rgt = right; mxsl = sleft; mxe = mne = kp; mxns = nsig; first = 0; }
        else {
            if (right > rgt) rgt = right;
            if (sleft > mxsl) mxsl = sleft;
            if (kp > mxe) mxe = kp;
            if (kp < mne) mne = kp;
            if (nsig > mxns) mxns = nsig;
        }
        have = 1;
    }
    int use_fixed = 1, dec = 0;
    if (have) {
        if (rgt < 0) rgt = 0;
        int wF = mxsl + rgt + (rgt != 0);
        int e = (mxe >= 100 || mne <= -99) ? 2 : 1;
        int d = mxns - 1;
        int wE = neg + (d > 0) + d + 4 + e;
        if (wF <= wE) { use_fixed = 1; dec = rgt; } else { use_fixed = 0; dec = d; }
    }
    for (int i = 0; i < n; i++) {
        if (!isfinite(v[i])) { back[i] = v[i]; is_zero_text[i] = 0; continue; }
        char out[400];
        double xv = v[i] == 0 ? 0.0 : v[i];
        snprintf(out, sizeof out, use_fixed ? "%.*f" : "%.*e", dec, xv);
        back[i] = atof(out);
        is_zero_text[i] = strcmp(out, "0") == 0;
    }
}

static int cmp_double(const void *a, const void *b)
{
    double x = *(const double *)a, y = *(const double *)b;
    return (x > y) - (x < y);
}

/* median(as.numeric(as.character(x[x > 0]))) over one row of `current` (pat + the crp values),
 * for each candidate row; returns the index (into cand) of which.max. */
static int which_max_median(const oracle_crp *c, int k, const int *cand, int ncand)
{
    int q = c->q_off[k + 1] - c->q_off[k], p = c->p_off[k + 1] - c->p_off[k];
    const double *X = c->x + c->x_off[k];
    const uint8_t *P = c->row_pat + c->x_off[k];
    double *col = (double *)malloc(sizeof(double) * ncand);
    double *back = (double *)malloc(sizeof(double) * (size_t)ncand * p);
    uint8_t *zt = (uint8_t *)malloc((size_t)ncand * p);
    double *b1 = (double *)malloc(sizeof(double) * ncand);
    uint8_t *z1 = (uint8_t *)malloc(ncand);
    for (int j = 0; j < p; j++) {
        for (int i = 0; i < ncand; i++) col[i] = X[(size_t)j * q + cand[i]];
        oracle_format_roundtrip(col, ncand, b1, z1);
        for (int i = 0; i < ncand; i++) { back[(size_t)i * p + j] = b1[i]; zt[(size_t)i * p + j] = z1[i]; }
    }
    int best = -1;
    double bestv = 0;
    double *vals = (double *)malloc(sizeof(double) * (p + 1));
    char *s = (char *)malloc((size_t)p + 1);
    for (int i = 0; i < ncand; i++) {
        int nv = 0;
        for (int j = 0; j < p; j++) s[j] = P[(size_t)cand[i] * p + j] ? '1' : '0';
        s[p] = 0;
        if (strcmp(s, "0") != 0) vals[nv++] = atof(s);    /* "0110" > "0" is TRUE for strings; reads back as 110 */
        for (int j = 0; j < p; j++)
            if (!zt[(size_t)i * p + j]) vals[nv++] = back[(size_t)i * p + j];
        if (nv == 0) continue;                             /* median(numeric(0)) = NA, skipped by which.max */
        qsort(vals, nv, sizeof(double), cmp_double);
        double med = (nv & 1) ? vals[nv / 2] : (vals[nv / 2 - 1] + vals[nv / 2]) / 2.0;
        if (isnan(med)) continue;
        if (best < 0 || med > bestv) { best = i; bestv = med; }
    }
    free(col); free(back); free(zt); free(b1); free(z1); free(vals); free(s);
    return best < 0 ? 0 : best;
}

/* hamming_correction (:562-648) for one observed pattern of block k: the block row it is moved to */
static int hamming_row(const oracle_crp *c, int k, const uint8_t *obs, int hamming_dist)
{
    int q = c->q_off[k + 1] - c->q_off[k], p = c->p_off[k + 1] - c->p_off[k];
    const uint8_t *P = c->row_pat + c->x_off[k];
    int *d = (int *)malloc(sizeof(int) * q), *cand = (int *)malloc(sizeof(int) * q);
    int dmin = 1 << 30, row = -1;
    for (int r = 0; r < q; r++) {                          /* :567-570 */
        int s = 0;
        for (int j = 0; j < p; j++) s += obs[j] != P[(size_t)r * p + j];
        d[r] = s;
        if (s < dmin) dmin = s;
    }
    for (int r = 0; r < q && row < 0; r++) if (d[r] == 0) row = r;     /* :609-610 */
    if (row < 0) {
        int close = 0;
        for (int r = 0; r < q; r++) if (d[r] <= hamming_dist) cand[close++] = r;   /* :612 */
        if (close == 1) row = cand[0];                                              /* :620-622 */
        else if (close > 1) row = cand[which_max_median(c, k, cand, close)];        /* :613-618 */
        else {                                                                      /* :624-627 */
            int m = 0;
            for (int r = 0; r < q; r++) if (d[r] <= dmin) cand[m++] = r;
            row = cand[which_max_median(c, k, cand, m)];
        }
    }
    free(d); free(cand);
    return row;
}

typedef struct {   /* per-thread scratch for one cell */
    int32_t *tx_head, *row_next;   /* transcript -> chain of table rows */
    int32_t *touched; int ntouched;
    int cap_rows;
} cell_index;

/* crpcount_td for one cell.  Table rows i = 0..n-1: (tx[i], count[i], eqclass[i]) in the order of
 * sample_eqclass (by eqclass id, then the class' transcript order).  y[sum q] receives the counts. */
int oracle_crpcount_td(const oracle_crp *c, int n, const int32_t *tx, const int32_t *count, const int32_t *eqclass,
                       int hamming_dist, double *y)
{
    const int B = c->B;
    memset(y, 0, sizeof(double) * (size_t)c->q_off[B]);
    if (n == 0) return 0;
    /* :438-446 eqclass re-indexing and Weight = 1 / (rows of the class) */
    int32_t emax = 0;
    for (int i = 0; i < n; i++) if (eqclass[i] > emax) emax = eqclass[i];
    int32_t *eq_n = (int32_t *)calloc((size_t)emax + 2, sizeof(int32_t));
    for (int i = 0; i < n; i++) eq_n[eqclass[i]]++;
    /* rows of one eqclass are contiguous in sample_eqclass; keep the first row of each for lookups */
    int32_t *eq_first = (int32_t *)malloc(sizeof(int32_t) * ((size_t)emax + 2));
    for (int32_t e = 0; e <= emax; e++) eq_first[e] = -1;
    for (int i = n - 1; i >= 0; i--) eq_first[eqclass[i]] = i;
    /* index: transcript -> rows (ascending) */
    int32_t *head = (int32_t *)malloc(sizeof(int32_t) * (size_t)c->n_tx);
    int32_t *next = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
    for (int i = 0; i < n; i++) if (tx[i] >= 0 && tx[i] < c->n_tx) head[tx[i]] = -1;
    for (int i = n - 1; i >= 0; i--)
        if (tx[i] >= 0 && tx[i] < c->n_tx) { next[i] = head[tx[i]]; head[tx[i]] = i; }
    /* mark which transcripts occur in this cell at all */
    uint8_t *present = (uint8_t *)calloc((size_t)c->n_tx, 1);
    for (int i = 0; i < n; i++) if (tx[i] >= 0 && tx[i] < c->n_tx) present[tx[i]] = 1;

    int32_t *eq_list = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
    uint8_t *eq_seen = (uint8_t *)calloc((size_t)emax + 2, 1);
    int32_t *cand = (int32_t *)malloc(sizeof(int32_t) * 4096);
    int cand_cap = 4096;

    for (int j = 0; j < B; j++) {                                              /* :453 */
        const int q = c->q_off[j + 1] - c->q_off[j], p = c->p_off[j + 1] - c->p_off[j];
        const int32_t *cols = c->col_tx + c->p_off[j];
        if (q * p == 1) {                                                      /* :457-464 */
            int t = cols[0];
            double s = 0;
            if (t >= 0 && t < c->n_tx && present[t])
                for (int i = head[t]; i >= 0; i = next[i])
                    if (1.0 / eq_n[eqclass[i]] > 0.9) s += count[i];           /* Weight > 0.9 */
            y[c->q_off[j]] = s;
            continue;
        }
        /* :468-475 raw.c = rows whose transcript is a supported column, column by column; eq = unique */
        int neq = 0;
        for (int cc = 0; cc < p; cc++) {
            if (!c->supported[c->p_off[j] + cc]) continue;
            int t = cols[cc];
            if (t < 0 || t >= c->n_tx || !present[t]) continue;
            for (int i = head[t]; i >= 0; i = next[i])
                if (!eq_seen[eqclass[i]]) { eq_seen[eqclass[i]] = 1; eq_list[neq++] = eqclass[i]; }
        }
        for (int i = 0; i < neq; i++) eq_seen[eq_list[i]] = 0;
        if (neq == 0) continue;                                                /* :542-543 all zero */
        /* :477-512 keep an eqclass only if this block is its best-matching block */
        int nkeep = 0;
        for (int ei = 0; ei < neq; ei++) {
            const int e = eq_list[ei];
            const int f = eq_first[e], m = eq_n[e];
            const int32_t *mytx = tx + f;                                      /* the class' transcripts */
            /* eq_candidates = sort(unique(unlist(t2c[mytx])))  :484 */
            int nc = 0;
            for (int a = 0; a < m; a++) {
                int t = mytx[a];
                if (t < 0 || t >= c->n_tx) continue;
                for (int h = c->t2c_off[t]; h < c->t2c_off[t + 1]; h++) {
                    int blk = c->t2c_block[h], dup = 0;
                    for (int z = 0; z < nc; z++) if (cand[z] == blk) { dup = 1; break; }
                    if (dup) continue;
                    if (nc == cand_cap) { cand_cap *= 2; cand = (int32_t *)realloc(cand, sizeof(int32_t) * cand_cap); }
                    cand[nc++] = blk;
                }
            }
            /* insertion sort by name rank (qsort's global comparator is not thread safe) */
            for (int a = 1; a < nc; a++) {
                int32_t v = cand[a]; int b = a - 1;
                while (b >= 0 && c->name_rank[cand[b]] > c->name_rank[v]) { cand[b + 1] = cand[b]; b--; }
                cand[b + 1] = v;
            }
            int best = -1;
            double bestsc = 0;
            int nuniq = 0;                                                     /* length(unique(mytx)) */
            for (int a = 0; a < m; a++) {
                int dup = 0;
                for (int b = 0; b < a; b++) if (mytx[b] == mytx[a]) { dup = 1; break; }
                nuniq += !dup;
            }
            for (int z = 0; z < nc; z++) {
                const int blk = cand[z];
                const int pb = c->p_off[blk + 1] - c->p_off[blk];
                const int32_t *bc = c->col_tx + c->p_off[blk];
                int keep = 0, inter = 0;
                for (int cc = 0; cc < pb; cc++) {
                    int in = 0;
                    for (int a = 0; a < m; a++) if (mytx[a] == bc[cc]) { in = 1; break; }
                    if (in) { inter++; if (c->supported[c->p_off[blk] + cc]) keep = 1; }   /* :487-493 */
                }
                if (!keep) continue;
                double sc = (double)inter / (double)(pb + nuniq - inter);      /* :502 */
                if (best < 0 || sc > bestsc) { best = blk; bestsc = sc; }      /* which.max: first maximum */
            }
            if (best == j) eq_list[nkeep++] = e;                                /* :506-508 */
        }
        if (nkeep == 0) continue;
        /* :516-531 patterns of the kept classes over ALL columns of the block, counts summed by pattern;
         * :534 hamming_correction; :536-540 merge onto the block's rows */
        uint8_t *obs = (uint8_t *)malloc((size_t)p);
        for (int ei = 0; ei < nkeep; ei++) {
            const int e = eq_list[ei];
            const int f = eq_first[e], m = eq_n[e];
            memset(obs, 0, (size_t)p);
            int a1 = 0;                                                         /* apply(raw2, 1, max) */
            for (int cc = 0; cc < p; cc++) {
                if (!c->supported[c->p_off[j] + cc]) continue;
                for (int a = 0; a < m; a++)
                    if (tx[f + a] == cols[cc]) {
                        if (count[f + a] > 0) obs[cc] = 1;                      /* raw3 = ifelse(raw2 > 0, 1, 0) */
                        if (count[f + a] > a1) a1 = count[f + a];
                    }
            }
            int row = hamming_row(c, j, obs, hamming_dist);
            y[c->q_off[j] + row] += a1;
        }
        free(obs);
    }
    free(eq_n); free(eq_first); free(head); free(next); free(present); free(eq_list); free(eq_seen); free(cand);
    return 0;
}

/* ---- many cells given as (eqclass id, count) lists + the eqclass dictionary ---------------------- */
typedef struct {
    const oracle_crp *c;
    int64_t n_cells; const int64_t *cell_ptr; const int32_t *eq_id, *eq_count;
    const int64_t *eq_off; const int32_t *eq_tx; int hamming; double *y;
    volatile int64_t next;
} cells_job;

static void *cells_worker(void *arg)
{
    cells_job *J = (cells_job *)arg;
    const int64_t nrow = J->c->q_off[J->c->B];
    int cap = 1 << 16;
    int32_t *tx = (int32_t *)malloc(sizeof(int32_t) * cap), *cnt = (int32_t *)malloc(sizeof(int32_t) * cap),
            *eq = (int32_t *)malloc(sizeof(int32_t) * cap);
    for (;;) {
        int64_t cell = __atomic_fetch_add(&J->next, 1, __ATOMIC_RELAXED);
        if (cell >= J->n_cells) break;
        int n = 0;
        for (int64_t k = J->cell_ptr[cell]; k < J->cell_ptr[cell + 1]; k++) {
            int32_t e = J->eq_id[k];
            for (int64_t a = J->eq_off[e]; a < J->eq_off[e + 1]; a++) {
                if (n == cap) {
                    cap *= 2;
                    tx = (int32_t *)realloc(tx, sizeof(int32_t) * cap);
                    cnt = (int32_t *)realloc(cnt, sizeof(int32_t) * cap);
                    eq = (int32_t *)realloc(eq, sizeof(int32_t) * cap);
                }
                tx[n] = J->eq_tx[a]; cnt[n] = J->eq_count[k]; eq[n] = e; n++;
            }
        }
        /* local dense eqclass ids keep the scratch arrays small (R re-indexes too, :438-441) */
        int32_t last = -1, id = -1;
        for (int i = 0; i < n; i++) { if (eq[i] != last) { last = eq[i]; id++; } eq[i] = id; }
        oracle_crpcount_td(J->c, n, tx, cnt, eq, J->hamming, J->y + (size_t)cell * nrow);
    }
    free(tx); free(cnt); free(eq);
    return NULL;
}

int oracle_crpcount_cells(const oracle_crp *c, int64_t n_cells, const int64_t *cell_ptr, const int32_t *eq_id,
                          const int32_t *eq_count, const int64_t *eq_off, const int32_t *eq_tx, int hamming,
                          int nthreads, double *y)
{
    cells_job J = {c, n_cells, cell_ptr, eq_id, eq_count, eq_off, eq_tx, hamming, y, 0};
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    pthread_t th[256];
    for (int t = 1; t < nthreads; t++) pthread_create(&th[t], NULL, cells_worker, &J);
    cells_worker(&J);
    for (int t = 1; t < nthreads; t++) pthread_join(th[t], NULL);
    return 0;
}
