/*
 * scasa_oracle.c — CPU restatement of Scasa's 2QUANT estimation path.  TEST INFRASTRUCTURE.
 *
 * This file is the checker for the CUDA path, never the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load it.  libscasa_cuda never links or calls it.
 *
 * PARITY UNPINNED: the reference (R scripts, /root/reference/scasa/SCRIPTS) ships no
 * tests, no golden outputs, and R is not installed in the build image, so this
 * restatement could not be checked against an execution of the reference.  It is
 * pinned only (a) against an independent literal numpy transcription
 * (oracle/estim_literal.py, which keeps R's shapes incl. the Kronecker form), (b) on
 * analytic cases (tests/test_oracle.py) and (c) on the data relations K1-K8 of the
 * shipped X-matrix.
 *
 * Every function cites the R lines it follows (paths relative to scasa/SCRIPTS/):
 *   Rsource = Scasa_Rsource.R, step2 = scasa_quant_step2_v1.0.0.R,
 *   step3 = scasa_quant_step3_v1.0.0.R.
 *
 * Third-party arithmetic restated here: base R's digamma() = nmath psigamma(x,0) =
 * dpsifn (D. E. Amos, ACM TOMS 9 (1983) Algorithm 610), not vendored by the reference
 * and not pinned to an R version (docker image nghiavtr/scasa:v1.0.1 is the only pin).
 * R's sum/rowSums/colSums accumulate in x87 long double: ACC below does the same on
 * x86-64.
 *
 * Layouts (all caller-owned):
 *   X-matrix  q_off[B+1], p_off[B+1], x_off[B+1], x[]  — blocks column-major (R's).
 *   Y         cell-major dense: y[c * nrow_total + q_off[k] + r] = Y[[k]][r, c].
 *   result    cell-major dense: out[c * ncol_total + p_off[k] + j] = final[c, col].
 */
#include <math.h>
#include <float.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

typedef long double ACC; /* R_xlen_t loops in R's colSums/rowSums/rsum use LDOUBLE */

/* ------------------------------------------------------------------ digamma */
/* R nmath/polygamma.c dpsifn() specialised to n = 0, kode = 1, m = 1, then
 * digamma(x) = -ans (psigamma).  Call sites: Rsource:733, Rsource:738. */
double oracle_digamma(double x)
{
    static const double bvalues[] = { /* Bernoulli numbers B0, B1, B2, B4, ... */
        1.0, -0.5, 1.0 / 6.0, -1.0 / 30.0, 1.0 / 42.0, -1.0 / 30.0, 5.0 / 66.0,
        -691.0 / 2730.0, 7.0 / 6.0, -3617.0 / 510.0, 43867.0 / 798.0,
        -174611.0 / 330.0, 854513.0 / 138.0, -236364091.0 / 2730.0, 8553103.0 / 6.0,
        -23749461029.0 / 870.0, 8615841276005.0 / 14322.0, -7709321041217.0 / 510.0,
        2577687858367.0 / 6.0, -1.371165520508833e13, 4.883323189735932e14,
        -1.929657934194006e16};
    if (isnan(x)) return x;
    if (x <= 0.0) return NAN; /* never reached on this path: arguments are >= 1e-8 */
    const double wdtol = fmax(DBL_EPSILON * 0.5, 0.5e-18);
    const double r1m5 = 0.30102999566398119521;
    const int nxlim = (-(DBL_MIN_EXP) < DBL_MAX_EXP) ? -(DBL_MIN_EXP) : DBL_MAX_EXP;
    const double elim = 2.302 * (nxlim * r1m5 - 3.0);
    double xln = log(x);
    if (fabs(xln) > elim) return (xln > 0) ? xln : -INFINITY; /* t = (fn+1)*xln, fn = 0 */
    if (x < wdtol) return -1.0 / x;              /* ans = x^(-n-1) */
    double rln = r1m5 * DBL_MANT_DIG;
    if (rln > 18.06) rln = 18.06;
    double fln = fmax(rln, 3.0) - 3.0;
    double yint = 3.50 + 0.40 * fln;
    int mx = (int)yint + 1;                      /* xm = yint + slope*fn, fn = 0 */
    double xmin = mx;
    double xdmy = x, xdmln = xln, xinc = 0.0;
    if (x < xmin) {
        int nx = (int)x;
        xinc = xmin - nx;
        xdmy = x + xinc;
        xdmln = log(xdmy);
    }
    /* asymptotic expansion at xdmy (label L10 in dpsifn; tss = exp(-0) = 1) */
    double tt = 0.5 / xdmy;
    double t1 = tt;
    double tst = wdtol * tt;
    double rxsq = 1.0 / (xdmy * xdmy);
    double ta = 0.5 * rxsq;
    double t = ta;                               /* (fn + 1) * ta */
    double s = t * bvalues[2];
    if (fabs(s) >= tst) {
        double tk = 2.0;
        for (int k = 4; k <= 22; k++) {
            t = t * ((tk + 1.0) / (tk + 1.0)) * (tk / (tk + 2.0)) * rxsq;
            double trm = t * bvalues[k - 1];
            if (fabs(trm) < tst) break;
            s += trm;
            tk += 2.0;
        }
    }
    s = s + t1;
    if (xinc != 0.0) {                           /* backward recurrence, label L20 */
        int nx = (int)xinc;
        for (int i = 1; i <= nx; i++) s += 1.0 / (x + (nx - i));
    }
    return -(s - xdmln);
}

/* ------------------------------------------------------------------ A1 chunk split */
static double r_round(double x) { return nearbyint(x); } /* half-to-even under FE_TONEAREST */

/* step2:64-68.  br[0..chunkNum] receives 0-based start positions (brPos - 1).
 * Returns chunkNum.  `core` only matters when n <= 100 (step2:66). */
int oracle_chunk_split(int64_t n, int core, int64_t *br, int64_t br_cap)
{
    const double chunkSize = 100.0;
    double chunkNum_d = r_round((double)n / chunkSize);
    int64_t chunkNum = (n > 100) ? (int64_t)chunkNum_d : core;
    if (chunkNum + 1 > br_cap) return -1;
    /* seq(1, n+1, length.out = chunkNum+1): ends exact, interior from + i*by */
    double from = 1.0, to = (double)n + 1.0;
    int64_t lout = chunkNum + 1;
    double by = (lout > 1) ? (to - from) / (double)(lout - 1) : 0.0;
    for (int64_t i = 0; i < lout; i++) {
        double v = (i == 0) ? from : (i == lout - 1 ? to : from + (double)i * by);
        br[i] = (int64_t)r_round(v) - 1;
    }
    return (int)chunkNum;
}

/* ------------------------------------------------------------------ A5 estim.BETA */
/* Rsource:726-761.  X q×p column-major, Ymat q×n column-major, beta0/beta n×p stored
 * cell-major (R: matrix(beta0, nrow=n, byrow=TRUE), :750).  Returns the number of
 * iterations executed.  fixed_iter != 0 disables the break at :756 (parity mode). */
int oracle_estim_beta(const double *X, int q, int p, const double *Ymat, int n,
                      double *beta0, double *beta, int maxiter, double maxerr,
                      double lim, int modify, int fixed_iter, double *work)
{
    double *csum = work;           /* p */
    double *logNorm = csum + p;    /* n */
    double *gamma0 = logNorm + n;  /* p */
    double *tmp = gamma0 + p;      /* q*p */
    double *rsum = tmp + (size_t)q * p; /* q */
    for (int j = 0; j < p; j++) {  /* csum = colSums(X)  :731 */
        ACC a = 0;
        for (int r = 0; r < q; r++) a += X[(size_t)j * q + r];
        csum[j] = (double)a;
    }
    for (int c = 0; c < n; c++) {  /* totCount = colSums(Ymat); logNorm = digamma(.+1e-8)  :732-733 */
        ACC a = 0;
        for (int r = 0; r < q; r++) a += Ymat[(size_t)c * q + r];
        logNorm[c] = oracle_digamma((double)a + 0.00000001);
    }
    int it;
    for (it = 1; it <= maxiter; it++) {          /* :749 */
        for (int c = 0; c < n; c++) {
            const double *bt = beta0 + (size_t)c * p;
            const double *yi = Ymat + (size_t)c * q;
            /* fn1 :736-742 */
            for (int j = 0; j < p; j++)
                gamma0[j] = modify ? exp(oracle_digamma(bt[j] + 0.00000001) - logNorm[c]) : bt[j];
            for (int r = 0; r < q; r++) {
                ACC a = 0;
                for (int j = 0; j < p; j++) {
                    tmp[(size_t)j * q + r] = X[(size_t)j * q + r] * gamma0[j];
                    a += tmp[(size_t)j * q + r];
                }
                rsum[r] = (double)a;
            }
            /* fn2 :744-747: bt = c(t(X2) %*% yi) / csum */
            for (int j = 0; j < p; j++) {
                double a = 0.0;
                for (int r = 0; r < q; r++) {
                    double x2 = tmp[(size_t)j * q + r] / (rsum[r] > 0 ? rsum[r] : 1.0);
                    a += x2 * yi[r];
                }
                double b = a / csum[j];
                if (isnan(b)) b = 0.0;           /* :754 */
                beta[(size_t)c * p + j] = b;
            }
        }
        double mx = -INFINITY;                    /* :755-756 */
        for (size_t i = 0; i < (size_t)n * p; i++) {
            double e = fabs(beta[i] - beta0[i]) / (beta0[i] > lim ? beta0[i] : lim);
            if (e > mx || isnan(e)) mx = e;
        }
        if (!fixed_iter && mx < maxerr) break;
        if (it < maxiter) memcpy(beta0, beta, sizeof(double) * (size_t)n * p); /* :757 */
    }
    return it > maxiter ? maxiter : it;
}

/* ------------------------------------------------------------------ A6 estim.X.em */
/* Rsource:898-922 (the second definition; the copy at :691-715 is identical and
 * overridden).  BETA cell-major n×p.  Xout q×p column-major. */
void oracle_estim_x_em(const double *Xem, int q, int p, const double *BETA, int n,
                       const double *Ymat, double *Xout, double *work)
{
    double *BTsum = work;               /* p */
    ACC *tmpacc = (ACC *)malloc(sizeof(ACC) * (size_t)q * p);
    for (int j = 0; j < p; j++) {       /* colSums(BETA)  :904 */
        ACC a = 0;
        for (int c = 0; c < n; c++) a += BETA[(size_t)c * p + j];
        BTsum[j] = (double)a;
    }
    for (size_t i = 0; i < (size_t)q * p; i++) tmpacc[i] = 0;
    /* x1[(c,r), j] = BETA[c,j] * X[r,j]  (kronecker, :906); rows indexed (c-1)*q + r = c(Ymat) */
    for (int c = 0; c < n; c++) {
        for (int r = 0; r < q; r++) {
            ACC s = 0;                  /* sumx1 = rowSums(x1)  :907 */
            for (int j = 0; j < p; j++) s += BETA[(size_t)c * p + j] * Xem[(size_t)j * q + r];
            double sumx1 = (double)s;
            double pos = sumx1 > 0 ? sumx1 : 0.1;      /* :908 */
            double y = Ymat[(size_t)c * q + r];
            for (int j = 0; j < p; j++) {
                double x1 = BETA[(size_t)c * p + j] * Xem[(size_t)j * q + r];
                double x2 = x1 / pos;                  /* :909 */
                tmpacc[(size_t)j * q + r] += x2 * y;   /* :910, colSums over cells :913 */
            }
        }
    }
    for (int j = 0; j < p; j++) {
        double d = BTsum[j] > 0 ? BTsum[j] : 0.1;      /* :917 */
        for (int r = 0; r < q; r++) Xout[(size_t)j * q + r] = (double)tmpacc[(size_t)j * q + r] / d;
        ACC cs = 0;                                    /* :918 */
        for (int r = 0; r < q; r++) cs += Xout[(size_t)j * q + r];
        double csum = (double)cs;
        double d2 = csum > 0 ? csum : 0.1;             /* :919 */
        for (int r = 0; r < q; r++) Xout[(size_t)j * q + r] /= d2;
        if (csum < 0.00001)                            /* :920 */
            for (int r = 0; r < q; r++) Xout[(size_t)j * q + r] = Xem[(size_t)j * q + r];
    }
    free(tmpacc);
}

/* ------------------------------------------------------------------ A4 estim_chunk */
static double cos_sim(const double *a, const double *b, int q) /* Rsource:950 */
{
    ACC ab = 0, aa = 0, bb = 0;
    for (int r = 0; r < q; r++) { ab += a[r] * b[r]; aa += a[r] * a[r]; bb += b[r] * b[r]; }
    return (double)ab / sqrt((double)aa * (double)bb);
}

/* poorCondX, Rsource:957-971: a 2-column block is "poor" unless some pairwise
 * cosine similarity (including the self pairs) is < 0.99. */
int oracle_is_poor(const double *X0, int q, int p)
{
    if (p != 2) return 0;
    for (int i = 0; i < p; i++)
        for (int k = 0; k < p; k++)
            if (cos_sim(X0 + (size_t)k * q, X0 + (size_t)i * q, q) < 0.99) return 0;
    return 1;
}

typedef struct {
    int maxiter_x, maxiter_bt;
    double maxerr, lim;
    int modify;
    double adjust_esp;
    int fixed_iter;
    /* estim_chunk does NOT forward maxerr / lim to estim.BETA (Rsource:1016): the inner loop always runs with
     * estim.BETA's own defaults (Rsource:726-727), 0.01 / 0.01.  Kept as fields so a test can tighten both. */
    double maxerr_bt, lim_bt;
} oracle_opts;

/* One chunk.  y: n × nrow_total cell-major.  out: n × ncol_total cell-major.
 * iters (optional): [B][2] = {outer, total inner} per block.  Rsource:944-1051. */
int oracle_estim_chunk(int n_blocks, const int32_t *q_off, const int32_t *p_off,
                       const int64_t *x_off, const double *x, int n, const double *y,
                       const oracle_opts *o, double *out, int32_t *iters)
{
    const int64_t nrow = q_off[n_blocks], ncol = p_off[n_blocks];
    int qmax = 1, pmax = 1;
    for (int m = 0; m < n_blocks; m++) {
        int q = q_off[m + 1] - q_off[m], p = p_off[m + 1] - p_off[m];
        if (q > qmax) qmax = q;
        if (p > pmax) pmax = p;
    }
    size_t qp = (size_t)qmax * pmax;
    double *Ymat = (double *)malloc(sizeof(double) * (size_t)qmax * n);
    double *Xest0 = (double *)malloc(sizeof(double) * qp);
    double *Xest = (double *)malloc(sizeof(double) * qp);
    double *beta0 = (double *)malloc(sizeof(double) * (size_t)n * pmax);
    double *BT = (double *)malloc(sizeof(double) * (size_t)n * pmax);
    double *work = (double *)malloc(sizeof(double) * (2 * (size_t)pmax + n + qp + qmax + 8));
    int *adjID = (int *)malloc(sizeof(int) * qmax);
    double *rs = (double *)malloc(sizeof(double) * qmax);
    if (!Ymat || !Xest0 || !Xest || !beta0 || !BT || !work || !adjID || !rs) return -1;

    for (int m = 0; m < n_blocks; m++) {                       /* :975 */
        const int q = q_off[m + 1] - q_off[m], p = p_off[m + 1] - p_off[m];
        const double *X0 = x + x_off[m];
        if (iters) iters[2 * m] = iters[2 * m + 1] = 0;
        if (q == 1) {                                          /* :979-983 pass-through */
            for (int c = 0; c < n; c++)
                out[(size_t)c * ncol + p_off[m]] = y[(size_t)c * nrow + q_off[m]];
            continue;
        }
        for (int c = 0; c < n; c++)
            for (int r = 0; r < q; r++) Ymat[(size_t)c * q + r] = y[(size_t)c * nrow + q_off[m] + r];

        int needAdjustX = 0, nadj = 0;                         /* :988-1008 */
        if (oracle_is_poor(X0, q, p)) {
            ACC tot = 0;
            for (int r = 0; r < q; r++) {
                ACC a = 0;
                for (int c = 0; c < n; c++) a += Ymat[(size_t)c * q + r];
                rs[r] = (double)a;
                tot += rs[r];
            }
            for (int r = 0; r < q; r++) rs[r] = rs[r] / (double)tot;
            int bestTx = -1;
            double best = -INFINITY;
            for (int j = 0; j < p; j++) {                      /* which.max skips NaN */
                double s = cos_sim(X0 + (size_t)j * q, rs, q);
                if (!isnan(s) && s > best) { best = s; bestTx = j; }
            }
            if (bestTx >= 0) {
                for (int r = 0; r < q; r++)
                    if (X0[(size_t)bestTx * q + r] > 0) adjID[nadj++] = r;
                if (nadj > 0) {
                    needAdjustX = 1;
                    for (int c = 0; c < n; c++)
                        for (int k = 0; k < nadj; k++) Ymat[(size_t)c * q + adjID[k]] += o->adjust_esp;
                }
            }
        }
        memcpy(Xest0, X0, sizeof(double) * (size_t)q * p);     /* :1011 */
        for (int c = 0; c < n; c++) {                          /* :1012-1013 */
            ACC a = 0;
            for (int r = 0; r < q; r++) a += Ymat[(size_t)c * q + r];
            for (int j = 0; j < p; j++) beta0[(size_t)c * p + j] = (double)a / p;
        }
        int i, inner = 0;
        for (i = 1; i <= o->maxiter_x; i++) {                  /* :1015 */
            inner += oracle_estim_beta(Xest0, q, p, Ymat, n, beta0, BT, o->maxiter_bt, o->maxerr_bt,
                                       o->lim_bt, o->modify, o->fixed_iter, work);
            oracle_estim_x_em(Xest0, q, p, BT, n, Ymat, Xest, work);
            double mx = -INFINITY;                             /* :1019 */
            for (size_t k = 0; k < (size_t)q * p; k++) {
                double e = fabs(Xest[k] - Xest0[k]) / (fabs(Xest0[k]) > o->lim ? Xest0[k] : o->lim);
                if (e > mx || isnan(e)) mx = e;
            }
            if (!o->fixed_iter && mx < o->maxerr) break;        /* :1020 */
            memcpy(Xest0, Xest, sizeof(double) * (size_t)q * p);     /* :1022 */
            memcpy(beta0, BT, sizeof(double) * (size_t)n * p);       /* :1023 */
        }
        if (iters) { iters[2 * m] = i > o->maxiter_x ? o->maxiter_x : i; iters[2 * m + 1] = inner; }
        if (needAdjustX) {                                     /* :1027-1037 */
            double deductVal = nadj * o->adjust_esp;
            for (int c = 0; c < n; c++) {
                ACC s = 0;
                for (int j = 0; j < p; j++) s += BT[(size_t)c * p + j];
                for (int j = 0; j < p; j++) {
                    double pr = BT[(size_t)c * p + j] / (double)s;
                    BT[(size_t)c * p + j] = BT[(size_t)c * p + j] - deductVal * pr;
                }
            }
        }
        for (int c = 0; c < n; c++)                            /* :1042 */
            for (int j = 0; j < p; j++) out[(size_t)c * ncol + p_off[m] + j] = BT[(size_t)c * p + j];
    }
    free(Ymat); free(Xest0); free(Xest); free(beta0); free(BT); free(work); free(adjID); free(rs);
    return 0;
}

/* All chunks of a sample: chunk i = cells [chunk_off[i], chunk_off[i+1]).  Mirrors the
 * foreach at step2:128-136 (one worker per chunk; pthreads ≙ registerDoParallel(cores)). */
typedef struct {
    int n_blocks; const int32_t *q_off, *p_off; const int64_t *x_off; const double *x;
    int n_chunks; const int64_t *chunk_off; const double *y; const oracle_opts *o;
    double *out; int32_t *iters; volatile int next; volatile int rc;
} chunk_job;

static void *chunk_worker(void *arg)
{
    chunk_job *j = (chunk_job *)arg;
    const int64_t nrow = j->q_off[j->n_blocks], ncol = j->p_off[j->n_blocks];
    for (;;) {
        int i = __atomic_fetch_add(&j->next, 1, __ATOMIC_RELAXED);
        if (i >= j->n_chunks) break;
        int64_t c0 = j->chunk_off[i], c1 = j->chunk_off[i + 1];
        int r = oracle_estim_chunk(j->n_blocks, j->q_off, j->p_off, j->x_off, j->x, (int)(c1 - c0),
                                   j->y + (size_t)c0 * nrow, j->o, j->out + (size_t)c0 * ncol,
                                   j->iters ? j->iters + (size_t)i * 2 * j->n_blocks : NULL);
        if (r) j->rc = r;
    }
    return NULL;
}

int oracle_estim_chunks(int n_blocks, const int32_t *q_off, const int32_t *p_off,
                        const int64_t *x_off, const double *x, int n_chunks,
                        const int64_t *chunk_off, const double *y, const oracle_opts *o,
                        double *out, int32_t *iters, int nthreads)
{
    chunk_job job = {n_blocks, q_off, p_off, x_off, x, n_chunks, chunk_off, y, o, out, iters, 0, 0};
    if (nthreads < 1) nthreads = 1;
    if (nthreads > n_chunks) nthreads = n_chunks > 0 ? n_chunks : 1;
    if (nthreads > 256) nthreads = 256;
    pthread_t th[256];
    for (int t = 1; t < nthreads; t++) pthread_create(&th[t], NULL, chunk_worker, &job);
    chunk_worker(&job);
    for (int t = 1; t < nthreads; t++) pthread_join(th[t], NULL);
    return job.rc;
}

/* ------------------------------------------------------------------ A7 row filter */
/* step2:166-167: after transposing to tx × cells keep rows with rowSums > 0.
 * iso: n_cells × ncol cell-major.  keep[ncol] ← 0/1.  Returns the number kept. */
int64_t oracle_keep_rows(const double *iso, int64_t n_cells, int64_t ncol, uint8_t *keep)
{
    int64_t kept = 0;
    for (int64_t j = 0; j < ncol; j++) {
        ACC a = 0;
        for (int64_t c = 0; c < n_cells; c++) a += iso[(size_t)c * ncol + j];
        keep[j] = ((double)a > 0) ? 1 : 0;
        kept += keep[j];
    }
    return kept;
}

/* ------------------------------------------------------------------ A8 gene sums */
/* step3:24-46 with the grouping already resolved to integers by the caller:
 * gene_of_col[j] = output gene row of isoform column j (or -1: NA gene, dropped by
 * tapply :36, or a row removed at step2:167).  The order of gene rows (singles first,
 * then multis, each sorted by name, :38-43) is part of that integer map and is built
 * by oracle_gene_order() below.  colSums (:27) accumulate in long double in isoform
 * row order. gene: n_cells × n_genes cell-major. */
void oracle_gene_sum(const double *iso, int64_t n_cells, int64_t ncol, const int32_t *gene_of_col,
                     int32_t n_genes, const int32_t *n_members, double *gene)
{
    ACC *acc = (ACC *)malloc(sizeof(ACC) * (size_t)(n_genes > 0 ? n_genes : 1));
    for (int64_t c = 0; c < n_cells; c++) {
        for (int32_t g = 0; g < n_genes; g++) acc[g] = 0;
        for (int64_t j = 0; j < ncol; j++) {
            int32_t g = gene_of_col[j];
            if (g < 0) continue;
            if (n_members[g] == 1) gene[(size_t)c * n_genes + g] = iso[(size_t)c * ncol + j]; /* :39 copy */
            else acc[g] += iso[(size_t)c * ncol + j];                                          /* :27 */
        }
        for (int32_t g = 0; g < n_genes; g++)
            if (n_members[g] != 1) gene[(size_t)c * n_genes + g] = (double)acc[g];
    }
    free(acc);
}
