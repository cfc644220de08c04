// scasa_em.cuh — the alternating EM of estim_chunk (Scasa_Rsource.R:1010-1042), one (chunk, block)
// task at a time, run to completion on chip.
//
// A GROUP of W warps (W = 1, 2, 4 or 8 by task size) takes a task from an atomic work queue, copies
// its blob into its shared-memory slot, builds the per-task index (record of each entry, entries by
// row) and then iterates without touching HBM again:
//
//   estim.BETA (Rsource:726-761), one iteration =
//     S1  thread per (cell, transcript) pair:  gamma0 = exp(digamma(beta + 1e-8) - logNorm)   :738
//     S2  thread per count entry (cell, row):  mu = sum_j X[row,j] gamma0[j];  w = y / mu      :739-741
//     S3  thread per pair:                     beta' = gamma0[j] * sum_rows X[row,j] w / csum[j],
//                                              err vote over the whole chunk                    :745,:755-756
//   estim.X.em (Rsource:898-922) =
//     thread per entry:  w = y / (sum_j X[row,j] beta[j])                                        :905-910
//     thread per X element (row, j): sum of beta[j] X w over the cells with counts in that row,
//     in cell order;  column normalisation, revert, and the outer test                           :911-920, :1019
//
// Nothing in a task's arithmetic depends on which group runs it, on the queue order or on the
// number of GPUs: every sum is taken in a fixed order (cell order, or row order) by one thread.
#pragma once

namespace scasa {

// gamma0 of estim.BETA's fn1 (Rsource:737-738) without the logarithm:
//   digamma(x) = log(xs) - 1/(2 xs) - series(1/xs^2) - sum_{i<9} 1/(x+i),   xs = x + 9 for x < 9
//   => exp(digamma(x) - L) = xs * exp(-(1/(2 xs) + series + rec + L)),
// one division serving both 1/xs and the folded reciprocal sum (see digamma_pos).
// exp(y) for the argument range of em_gamma_fast (y <= ~0.6; very negative for dying transcripts): Cody-Waite
// reduction by ln 2, degree-13 Taylor polynomial on |r| <= 0.347 (truncation 4e-18), scaling through the exponent
// field.  Below -700 (results under 1e-304) the library exp takes over, so underflow to denormals and to an exact 0
// is what R's exp gives (Rsource:738: gamma = 0 is how a transcript dies).  ~25 instructions instead of ~60 executed
// of the library's 135; differs from it by at most 1-2 ulp.
__device__ __forceinline__ double em_exp(double y)
{
    if (y < -700.0) return y < -746.0 ? 0.0 : exp(y);
    const double kd = rint(y * 1.4426950408889634074);
    double r = fma(kd, -6.93147180369123816490e-01, y);
    r = fma(kd, -1.90821492927058770002e-10, r);
    double p = 1.0 / 6227020800.0;
    p = fma(p, r, 1.0 / 479001600.0);
    p = fma(p, r, 1.0 / 39916800.0);
    p = fma(p, r, 1.0 / 3628800.0);
    p = fma(p, r, 1.0 / 362880.0);
    p = fma(p, r, 1.0 / 40320.0);
    p = fma(p, r, 1.0 / 5040.0);
    p = fma(p, r, 1.0 / 720.0);
    p = fma(p, r, 1.0 / 120.0);
    p = fma(p, r, 1.0 / 24.0);
    p = fma(p, r, 1.0 / 6.0);
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    const int k = (int)kd;                                    // >= -1010 here: the result is a normal number
    return __hiloint2double(__double2hiint(p) + k * 1048576, __double2loint(p));
}
__device__ __forceinline__ float em_exp(float y) { return expf(y); }

template <typename R> __device__ __forceinline__ R em_gamma_fast(R beta, R lnorm)
{
    const R x = beta + R(0.00000001);
    R xs = x, num = R(0), den = R(1);
    if (x < R(9)) {
        const R u = x * (x + R(8));
        const R a = u + R(7), b = u + R(12), c = u + R(15);
        const R bc = b * c;
        const R D = u * a * bc;
        const R N = a * bc + u * (bc + a * (b + c));
        const R x4 = x + R(4);
        num = (R(2) * x + R(8)) * N * x4 + D;
        den = D * x4;
        xs = x + R(9);
    }
    const R inv = R(1) / (xs * den);
    const R r = den * inv;
    const R rec = num * xs * inv;
    const R r2 = r * r;
    const R s = r2 * (R(1.0 / 12.0) - r2 * (R(1.0 / 120.0) - r2 * (R(1.0 / 252.0) - r2 * (R(1.0 / 240.0) -
                r2 * (R(1.0 / 132.0) - r2 * (R(691.0 / 32760.0) - r2 * R(1.0 / 12.0)))))));
    return xs * em_exp(-(R(0.5) * r + s + rec + lnorm));
}
__device__ __forceinline__ double rabs(double x) { return fabs(x); }
__device__ __forceinline__ float rabs(float x) { return fabsf(x); }

struct EmArgs {
    XDev X;
    Sample S;
    Tasks T;
    EmOpts O;
    Result R;                     // estimates go into R.rs_val
    int32_t *iters;
    unsigned long long *stats;
    const int32_t *list;          // task ids of this work class
    const unsigned int *list_n;
    unsigned int *next;           // queue head
    int slot_bytes;               // shared memory per group
};

template <int W> struct Group {
    static constexpr int kThreads = 32 * W;
    int gl;    // thread index inside the group
    int bar;   // named barrier of the group (W > 1)
    __device__ __forceinline__ void sync() const
    {
        if (W == 1) __syncwarp();
        else asm volatile("bar.sync %0, %1;" ::"r"(bar), "n"(kThreads) : "memory");
    }
    // group-wide OR, and a barrier with memory ordering: the callers read other lanes' shared-memory
    // writes right after it (a vote alone orders nothing, hence the __syncwarp for W == 1)
    __device__ __forceinline__ bool any(bool v) const
    {
        if (W == 1) { const bool r = __any_sync(kFull, v); __syncwarp(); return r; }
        int out;
        asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.u32 q, %1, 0;\n\tbar.red.or.pred p, %2, %3, q;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(out) : "r"((int)v), "r"(bar), "n"(kThreads) : "memory");
        return out != 0;
    }
};

// R = double: the parity path.  R = float (scasa_opts_t.fp32): shared-memory state and arithmetic in
// fp32, counts and results still fp64 in HBM; estimates agree with the fp64 path to ~1e-4 relative.
template <int W, typename R> __global__ void __launch_bounds__(256, 4) em_task_kernel(EmArgs A)
{
    extern __shared__ __align__(16) unsigned char em_smem[];
    constexpr int GT = 32 * W;
    const int lane = threadIdx.x & 31;
    const int grp = threadIdx.x / GT;
    Group<W> G;
    G.gl = threadIdx.x - grp * GT;
    G.bar = 1 + grp;
    const int gl = G.gl;
    unsigned char *slot = em_smem + (size_t)grp * A.slot_bytes;
    __shared__ unsigned int s_ticket[8];
    const XDev &X = A.X;
    const EmOpts &O = A.O;
    const unsigned n_list = *A.list_n;

    for (;;) {
        // ---- next task
        unsigned ti = 0;
        if (W == 1) {
            if (lane == 0) ti = atomicAdd(A.next, 1u);
            ti = __shfl_sync(kFull, ti, 0);
        } else {
            if (gl == 0) s_ticket[grp] = atomicAdd(A.next, 1u);
            G.sync();
            ti = s_ticket[grp];
            G.sync();
        }
        if (ti >= n_list) break;
        const int64_t t = A.list[ti];
        const int h = (int)(t / X.n_em);
        const int4 td = A.T.desc[t];                                  // block, shape and extents in one load
        const int64_t blob_off = A.T.toff[t];
        const int64_t cb0 = A.S.chunk_off[h], cb1 = A.S.chunk_off[h + 1];
        const int b = td.x;
        const int q = td.y & 0xffff, p = td.y >> 16;
        const int qp = q * p;
        const int n = td.z & 0xffff, E = (int)((unsigned)td.z >> 16);
        const int np = n * p;
        const bool poor = X.poor[b] != 0;
        const int nc = (int)(cb1 - cb0);
        const BlobLayout BL = blob_layout(n, E);

        // ---- carve the slot
        double *ey = (double *)slot;
        const uint32_t *rcell = (const uint32_t *)(slot + BL.off_cell);
        const uint32_t *rpos = (const uint32_t *)(slot + BL.off_pos);
        const uint16_t *eptr = (const uint16_t *)(slot + BL.off_ptr);
        const uint16_t *erow = (const uint16_t *)(slot + BL.off_row);
        R *ew = (R *)(slot + BL.bytes);
        R *Xc = ew + E, *Xn = Xc + qp, *rcs = Xn + qp, *bts = rcs + p, *nrm = bts + p, *lnorm = nrm + p;
        R *beta = lnorm + n, *gam = beta + np;
        uint16_t *erec = (uint16_t *)(gam + np), *rptr = erec + E, *perm = rptr + q + 1;
        int *cur = (int *)Xn;                        // q ints of scratch until the first X update

        // ---- stage: blob image, X.est0 = X0 (:1011)
        {
            const uint4 *src = (const uint4 *)(A.T.tdata + blob_off);
            uint4 *dst = (uint4 *)slot;
            for (int i = gl; i < (BL.bytes >> 4); i += GT) dst[i] = __ldcs(src + i);
            const double *X0 = X.x + td.w;
            for (int k = gl; k < qp; k += GT) Xc[k] = (R)X0[k];
        }
        G.sync();
        // a poor task's last record is the pseudo cell: all pmult cells of the chunk without counts in the block
        const uint32_t last_cw = rcell[n - 1];
        const int ap = (last_cw & kPseudo) ? n - 1 : -1;
        const R pmult = (last_cw & kPseudo) ? (R)(last_cw & 0xffff) : R(0);
        const int n_real = ap >= 0 ? n - 1 : n;
        for (int a = gl; a < n; a += GT) {
            const int k1 = eptr[a + 1];
            for (int k = eptr[a]; k < k1; k++) erec[k] = (uint16_t)a;
        }
        for (int r = gl; r < q; r += GT) cur[r] = 0;
        G.sync();
        // entries by row, stable in cell order (one warp: counting sort with match_any)
        if (gl < 32) {
            for (int k0 = 0; k0 < E; k0 += 32) {
                const int k = k0 + lane;
                const bool valid = k < E;
                const unsigned act = __ballot_sync(kFull, valid);
                if (valid) {
                    const int r = erow[k];
                    const unsigned peers = __match_any_sync(act, r);
                    if (lane == __ffs(peers) - 1) cur[r] += __popc(peers);
                }
                __syncwarp();
            }
            if (lane == 0) {
                int s = 0;
                for (int r = 0; r < q; r++) { const int c = cur[r]; rptr[r] = (uint16_t)s; cur[r] = s; s += c; }
                rptr[q] = (uint16_t)s;
            }
            __syncwarp();
            for (int k0 = 0; k0 < E; k0 += 32) {
                const int k = k0 + lane;
                const bool valid = k < E;
                const unsigned act = __ballot_sync(kFull, valid);
                int r = 0;
                unsigned peers = 0;
                if (valid) {
                    r = erow[k];
                    peers = __match_any_sync(act, r);
                    perm[cur[r] + __popc(peers & ((1u << lane) - 1u))] = (uint16_t)k;
                }
                __syncwarp();
                if (valid && lane == __ffs(peers) - 1) cur[r] += __popc(peers);
                __syncwarp();
            }
        }
        G.sync();

        // ---- poor blocks: rowSums(Ymat), bestTx, adjID, Ymat[adjID,] += adjust_esp (:989-1004)
        int nadj = 0;
        if (poor) {
            R *rs = Xn;
#pragma unroll 1
            for (int r = gl; r < q; r += GT) {
                double s = 0;
                const int i1 = rptr[r + 1];
#pragma unroll 1
                for (int i = rptr[r]; i < i1; i++) {
                    const int k = perm[i];
                    s += (erec[k] == ap ? (double)pmult : 1.0) * ey[k];
                }
                rs[r] = (R)s;
            }
            G.sync();
            const double *X0 = X.x + td.w;
            double tot = 0;
#pragma unroll 1
            for (int r = 0; r < q; r++) tot += (double)rs[r];
            int best = -1;
            double bestv = -INFINITY;
#pragma unroll 1
            for (int j = 0; j < p; j++) {                               // which.max(cos_sim(X0[,j], rs/sum(rs)))  :991-995
                double ab = 0, aa = 0, bb = 0;
#pragma unroll 1
                for (int r = 0; r < q; r++) {
                    const double xv = X0[j * q + r], rv = (double)rs[r] / tot;
                    ab += xv * rv; aa += xv * xv; bb += rv * rv;
                }
                const double sim = ab / sqrt(aa * bb);
                if (sim == sim && sim > bestv) { bestv = sim; best = j; }
            }
            unsigned long long adjmask = 0;
            if (best >= 0)
#pragma unroll 1
                for (int r = 0; r < q; r++)
                    if (X0[best * q + r] > 0) { adjmask |= 1ull << r; nadj++; }   // adjID  :997
            G.sync();
#pragma unroll 1
            for (int k = gl; k < E; k += GT)
                if ((adjmask >> erow[k]) & 1ull) ey[k] += O.adjust_esp;
            G.sync();
        }

        // ---- 1/colSums(X) (:731, :745), logNorm (:732-733), beta0 = colSums(Ymat)/p (:1012-1013)
        for (int j = gl; j < p; j += GT) {
            R s = 0;
#pragma unroll 1
            for (int r = 0; r < q; r++) s += Xc[j * q + r];
            rcs[j] = R(1) / s;
        }
        const R lim = (R)O.lim, maxerr = (R)O.maxerr, lim_bt = (R)O.lim_bt, maxerr_bt = (R)O.maxerr_bt;
        for (int a = gl; a < n; a += GT) {
            double ys = 0;
            const int k1 = eptr[a + 1];
#pragma unroll 1
            for (int k = eptr[a]; k < k1; k++) ys += ey[k];
            lnorm[a] = (R)digamma_pos(ys + 0.00000001);
            const R b0 = (R)(ys / p);
#pragma unroll 1
            for (int j = 0; j < p; j++) beta[j * n + a] = b0;
        }
        G.sync();

        // the pair index idx = j*n + a and the X element index idx = j*q + r walk in steps of GT
        const int j_first = gl / n, a_first = gl - j_first * n;
        const int dj = GT / n, da = GT - dj * n;
        const int xj_first = gl / q, xr_first = gl - xj_first * q;
        const int dxj = GT / q, dxr = GT - dxj * q;

        int outer = 0, inner_total = 0;
        for (;;) {
            // ================= estim.BETA(X.est0, Ymat, beta0)  :726-761
            for (int it = 1;; it++) {
#pragma unroll 1
                for (int idx = gl, a = a_first; idx < np; idx += GT) {          // S1
                    gam[idx] = O.modify ? em_gamma_fast<R>(beta[idx], lnorm[a]) : beta[idx];
                    a += da;
                    if (a >= n) a -= n;
                }
                G.sync();
#pragma unroll 1
                for (int k = gl; k < E; k += GT) {                               // S2
                    const int r = erow[k], a = erec[k];
                    R mu = 0;                                                    // rowSums(t(t(X)*gamma0))  :739-740
#pragma unroll 1
                    for (int j = 0; j < p; j++) mu += Xc[j * q + r] * gam[j * n + a];
                    ew[k] = (R)ey[k] / (mu > 0 ? mu : R(1));                     // :741
                }
                G.sync();
                bool fail = false;
#pragma unroll 1
                for (int idx = gl, a = a_first, j = j_first; idx < np; idx += GT) {   // S3
                    R s = 0;
                    const int k1 = eptr[a + 1];
#pragma unroll 1
                    for (int k = eptr[a]; k < k1; k++) s += Xc[j * q + erow[k]] * ew[k];
                    const R old = beta[idx];
                    const R gs = gam[idx] * s;
                    R nb = gs == R(0) ? R(0) : gs * rcs[j];                      // :745 (division by csum as a reciprocal)
                    if (nb != nb) nb = R(0);                                     // :754
                    const R den = old > lim_bt ? old : lim_bt;                   // :755
                    if (!(rabs(nb - old) < maxerr_bt * den)) fail = true;
                    beta[idx] = nb;
                    a += da; j += dj;
                    if (a >= n) { a -= n; j++; }
                }
                inner_total++;
                const bool any_fail = G.any(fail);                               // :756
                if ((!O.fixed_iter && !any_fail) || it >= O.maxiter_bt) break;
            }

            // ================= estim.X.em(X.est0, BT, Ymat)  :898-922
            for (int j = gl; j < p; j += GT) {                                   // BTsum = colSums(BETA)  :904
                R s = 0;
#pragma unroll 1
                for (int a = 0; a < n_real; a++) s += beta[j * n + a];
                if (ap >= 0) s += pmult * beta[j * n + ap];                      // the pseudo cell stands for pmult cells
                bts[j] = s > 0 ? s : R(0.1);                                     // :917 (may be denormal: no reciprocal)
            }
#pragma unroll 1
            for (int k = gl; k < E; k += GT) {
                const int r = erow[k], a = erec[k];
                R mu = 0;                                                         // sumx1 = rowSums(x1)  :907
#pragma unroll 1
                for (int j = 0; j < p; j++) mu += beta[j * n + a] * Xc[j * q + r];
                const R mult = a == ap ? pmult : R(1);
                ew[k] = mult * (R)ey[k] / (mu > 0 ? mu : R(0.1));                // :908-910
            }
            G.sync();
#pragma unroll 1
            for (int idx = gl, j = xj_first, r = xr_first; idx < qp; idx += GT) {   // X.est = colSums(x2 * Y) / BTsum  :911-917
                const R x = Xc[idx];
                R acc = 0;
                const int i1 = rptr[r + 1];
#pragma unroll 1
                for (int i = rptr[r]; i < i1; i++) {
                    const int k = perm[i];
                    acc += beta[j * n + erec[k]] * x * ew[k];
                }
                Xn[idx] = acc == R(0) ? R(0) : acc / bts[j];                     // zero numerators would take the slow division path
                r += dxr; j += dxj;
                if (r >= q) { r -= q; j++; }
            }
            G.sync();
            for (int j = gl; j < p; j += GT) {                                   // csum = colSums(X.est)  :918
                R csum = 0;
#pragma unroll 1
                for (int r = 0; r < q; r++) csum += Xn[j * q + r];
                // :919-920: scale by 1/csum (0.1 when 0); a negative scale marks "revert to X.est0"
                const R sc = csum < R(0.00001) ? R(-1) : R(1) / (csum > 0 ? csum : R(0.1));
                nrm[j] = sc;
                if (sc >= 0) {                                                   // 1/colSums of the normalised column, for the next estim.BETA call
                    R s2 = 0;
#pragma unroll 1
                    for (int r = 0; r < q; r++) s2 += Xn[j * q + r] * sc;
                    rcs[j] = R(1) / s2;
                }
            }
            G.sync();
            bool xfail = false;                                                  // :1019
#pragma unroll 1
            for (int idx = gl, j = xj_first, r = xr_first; idx < qp; idx += GT) {
                const R x0 = Xc[idx], sc = nrm[j];
                const R v = sc < 0 ? x0 : Xn[idx] * sc;
                Xc[idx] = v;                                                     // X.est0 = X.est (:1022); unused if the loop ends here
                const R num = rabs(v - x0), den = rabs(x0) > lim ? x0 : lim;
                const bool ok = den > 0 ? (num < maxerr * den) : (num / den < maxerr);
                if (!ok) xfail = true;
                r += dxr; j += dxj;
                if (r >= q) { r -= q; j++; }
            }
            outer++;
            const bool any_x = G.any(xfail);
            if ((!O.fixed_iter && !any_x) || outer >= O.maxiter_x) break;        // :1020
        }

        // ---- de-adjust (:1027-1037) and write BT into the result rows (:1042)
        const double deduct = nadj * O.adjust_esp;
        const int64_t cbase = cb0 - A.S.cell_base;
        double *vals = A.R.rs_val + A.R.rs_ptr[cbase];                           // the chunk's first value; rpos is relative to it
#pragma unroll 1
        for (int a = gl; a < n; a += GT) {
            const uint32_t cw = rcell[a];
            double s = 0;
            if (nadj > 0) for (int j = 0; j < p; j++) s += (double)beta[j * n + a];
            if (!(cw & kPseudo)) {
                double *o = vals + rpos[a];
#pragma unroll 1
                for (int j = 0; j < p; j++) { double v = (double)beta[j * n + a]; if (nadj > 0) v = v - deduct * (v / s); o[j] = v; }
            }
        }
        if (n > 0 && (rcell[n - 1] & kPseudo) && (rcell[n - 1] & 0xffff) != 0) {
            // cells of the chunk without a record of their own take the pseudo cell's values; the real
            // cells are listed in ascending order, so walk the gaps.  Every cell's row holds the two
            // columns of every poor block (struct_fill_kernel): poor_pos says where.
            const int ap = n - 1;
            const int kp = X.poor_rank[b];
            double s = 0, v0 = (double)beta[ap], v1 = (double)beta[n + ap];      // a poor block has two columns
            if (nadj > 0) { s = v0 + v1; v0 = v0 - deduct * (v0 / s); v1 = v1 - deduct * (v1 / s); }
#pragma unroll 1
            for (int cc = gl; cc < nc; cc += GT) {                               // a lane per cell of the chunk: is it a real record?
                int lo = 0, hi = ap;
                while (lo < hi) { const int mid = (lo + hi) >> 1; if ((int)rcell[mid] < cc) lo = mid + 1; else hi = mid; }
                if (lo < ap && (int)rcell[lo] == cc) continue;
                double *o = A.R.rs_val + A.R.rs_ptr[cbase + cc] + A.R.poor_pos[(cbase + cc) * X.n_poor + kp];
                o[0] = v0; o[1] = v1;
            }
        }
        if (gl == 0) task_done_stats(A.iters, A.stats, t, outer, inner_total, q, p, nc, n);
        G.sync();
    }
}

// ------------------------------------------------------------------------------------------
// Tiny tasks (not poor, n*p <= 16, E <= 16, q*p <= 16: more than half of all tasks) two per warp.
// Each half-warp owns a task and a shared-memory slot; every stage fits one pass of its 16 lanes.
// The warp runs a loop whose body is LOAD / one estim.BETA iteration / estim.X.em / OUT, each guarded
// by the half's own phase: halves in the same phase execute it together (the common case for the
// inner iteration), a half in another phase waits at the end of the block.  Same arithmetic, in
// the same order, as em_task_kernel: a task's result does not depend on which kernel ran it.
template <typename R> __global__ void __launch_bounds__(128, 8) em_pair_kernel(EmArgs A)
{
    extern __shared__ __align__(16) unsigned char em_smem[];
    enum { LOAD = 0, INNER = 1, XUPD = 2, OUT = 3, DONE = 4 };
    const int lane = threadIdx.x & 31;
    const int hl = lane & 15;                                  // lane inside the half
    const unsigned hmask = (lane & 16) ? 0xffff0000u : 0x0000ffffu;
    const int hlead = lane & 16;                               // first lane of the half
    unsigned char *slot = em_smem + (size_t)(threadIdx.x >> 4) * A.slot_bytes;
    const XDev &X = A.X;
    const EmOpts &O = A.O;
    const unsigned n_list = *A.list_n;
    const R lim = (R)O.lim, maxerr = (R)O.maxerr, lim_bt = (R)O.lim_bt, maxerr_bt = (R)O.maxerr_bt;

    int phase = LOAD;
    int64_t t = 0;
    int q = 1, p = 1, n = 1, E = 1, np = 1, qp = 1, nc = 0;
    int outer = 0, inner_total = 0, it = 0;
    double *vals = nullptr;                                    // first result value of the task's chunk
    // slot arrays (set in LOAD)
    double *ey = nullptr;
    const uint32_t *rpos = nullptr;
    const uint16_t *eptr = nullptr, *erow = nullptr;
    R *ew = nullptr, *Xc = nullptr, *Xn = nullptr, *rcs = nullptr, *bts = nullptr, *nrm = nullptr, *lnorm = nullptr, *beta = nullptr,
      *gam = nullptr;
    uint16_t *erec = nullptr, *rptr = nullptr, *perm = nullptr;

    for (;;) {
        // ================= LOAD: next task of the queue into this half's slot
        if (phase == LOAD) {
            unsigned ti = 0;
            if (hl == 0) ti = atomicAdd(A.next, 1u);
            ti = __shfl_sync(hmask, ti, hlead);
            if (ti >= n_list) phase = DONE;
            else {
                t = A.list[ti];
                const int hh = (int)(t / X.n_em);
                const int4 td = A.T.desc[t];                   // block, shape and extents in one load
                const int64_t blob_off = A.T.toff[t];
                const int64_t c0 = A.S.chunk_off[hh], c1 = A.S.chunk_off[hh + 1];
                nc = (int)(c1 - c0);
                vals = A.R.rs_val + A.R.rs_ptr[c0 - A.S.cell_base];
                q = td.y & 0xffff; p = td.y >> 16;
                qp = q * p;
                n = td.z & 0xffff; E = (int)((unsigned)td.z >> 16);
                np = n * p;
                const BlobLayout BL = blob_layout(n, E);
                ey = (double *)slot;
                rpos = (const uint32_t *)(slot + BL.off_pos);
                eptr = (const uint16_t *)(slot + BL.off_ptr);
                erow = (const uint16_t *)(slot + BL.off_row);
                ew = (R *)(slot + BL.bytes);
                Xc = ew + E; Xn = Xc + qp; rcs = Xn + qp; bts = rcs + p; nrm = bts + p; lnorm = nrm + p;
                beta = lnorm + n; gam = beta + np;
                erec = (uint16_t *)(gam + np); rptr = erec + E; perm = rptr + q + 1;
                {
                    const uint4 *src = (const uint4 *)(A.T.tdata + blob_off);
                    uint4 *dst = (uint4 *)slot;
#pragma unroll 1
                    for (int i = hl; i < (BL.bytes >> 4); i += 16) dst[i] = __ldcs(src + i);
                    const double *X0 = X.x + td.w;
                    if (hl < qp) Xc[hl] = (R)X0[hl];
                }
                __syncwarp(hmask);
                if (hl < n) {
                    const int k1 = eptr[hl + 1];
#pragma unroll 1
                    for (int k = eptr[hl]; k < k1; k++) erec[k] = (uint16_t)hl;
                }
                // entries by row, stable in cell order: rank of an entry = entries with a smaller (row, index)
                {
                    const int myrow = hl < E ? erow[hl] : 0x7fffffff;
                    int rank = 0, below = 0;
#pragma unroll 1
                    for (int k = 0; k < E; k++) {
                        const int rk = __shfl_sync(hmask, myrow, hlead + k);
                        rank += (rk < myrow) || (rk == myrow && k < hl);
                        below += rk < hl;                                   // entries in rows before row hl (for rptr)
                    }
                    if (hl < E) perm[rank] = (uint16_t)hl;
                    if (hl < q) rptr[hl] = (uint16_t)below;
                    if (hl == 0) rptr[q] = (uint16_t)E;
                }
                if (hl < p) {                                               // 1/colSums(X)  :731, :745
                    R s0 = 0;
#pragma unroll 1
                    for (int r = 0; r < q; r++) s0 += Xc[hl * q + r];
                    rcs[hl] = R(1) / s0;
                }
                if (hl < n) {                                               // logNorm :732-733, beta0 :1012-1013
                    double ys = 0;
                    const int k1 = eptr[hl + 1];
#pragma unroll 1
                    for (int k = eptr[hl]; k < k1; k++) ys += ey[k];
                    lnorm[hl] = (R)digamma_pos(ys + 0.00000001);
                    const R b0 = (R)(ys / p);
#pragma unroll 1
                    for (int j = 0; j < p; j++) beta[j * n + hl] = b0;
                }
                __syncwarp(hmask);
                outer = 0; inner_total = 0; it = 0;
                phase = INNER;
            }
        }
        if (__all_sync(kFull, phase == DONE)) break;

        // ================= one iteration of estim.BETA  :726-761
        if (phase == INNER) {
            const int pj = hl / n, pa = hl - pj * n;                        // pair hl = (transcript pj, cell pa)
            if (hl < np) gam[hl] = O.modify ? em_gamma_fast<R>(beta[hl], lnorm[pa]) : beta[hl];      // S1
            __syncwarp(hmask);
            if (hl < E) {                                                    // S2
                const int r = erow[hl], a = erec[hl];
                R mu = 0;
#pragma unroll 1
                for (int j = 0; j < p; j++) mu += Xc[j * q + r] * gam[j * n + a];
                ew[hl] = (R)ey[hl] / (mu > 0 ? mu : R(1));
            }
            __syncwarp(hmask);
            bool fail = false;
            if (hl < np) {                                                   // S3
                R s0 = 0;
                const int k1 = eptr[pa + 1];
#pragma unroll 1
                for (int k = eptr[pa]; k < k1; k++) s0 += Xc[pj * q + erow[k]] * ew[k];
                const R old = beta[hl];
                const R gs = gam[hl] * s0;
                R nb = gs == R(0) ? R(0) : gs * rcs[pj];
                if (nb != nb) nb = R(0);
                const R den = old > lim_bt ? old : lim_bt;
                if (!(rabs(nb - old) < maxerr_bt * den)) fail = true;
                beta[hl] = nb;
            }
            it++;
            inner_total++;
            const bool any_fail = __any_sync(hmask, fail);
            __syncwarp(hmask);                                               // beta[] is read across lanes next
            if ((!O.fixed_iter && !any_fail) || it >= O.maxiter_bt) phase = XUPD;
        }

        // ================= estim.X.em  :898-922, outer test :1019-1020
        if (phase == XUPD) {
            if (hl < p) {                                                    // BTsum  :904, :917
                R s0 = 0;
#pragma unroll 1
                for (int a = 0; a < n; a++) s0 += beta[hl * n + a];
                bts[hl] = s0 > 0 ? s0 : R(0.1);
            }
            if (hl < E) {
                const int r = erow[hl], a = erec[hl];
                R mu = 0;
#pragma unroll 1
                for (int j = 0; j < p; j++) mu += beta[j * n + a] * Xc[j * q + r];
                ew[hl] = (R)ey[hl] / (mu > 0 ? mu : R(0.1));                 // mult = 1: no pseudo cell here
            }
            __syncwarp(hmask);
            const int xj = hl / q, xr = hl - xj * q;                         // element hl = (row xr, transcript xj)
            if (hl < qp) {
                const R x = Xc[hl];
                R acc = 0;
                const int i1 = rptr[xr + 1];
#pragma unroll 1
                for (int i = rptr[xr]; i < i1; i++) {
                    const int k = perm[i];
                    acc += beta[xj * n + erec[k]] * x * ew[k];
                }
                Xn[hl] = acc == R(0) ? R(0) : acc / bts[xj];
            }
            __syncwarp(hmask);
            if (hl < p) {
                R csum = 0;
#pragma unroll 1
                for (int r = 0; r < q; r++) csum += Xn[hl * q + r];
                const R sc = csum < R(0.00001) ? R(-1) : R(1) / (csum > 0 ? csum : R(0.1));
                nrm[hl] = sc;
                if (sc >= 0) {
                    R s2 = 0;
#pragma unroll 1
                    for (int r = 0; r < q; r++) s2 += Xn[hl * q + r] * sc;
                    rcs[hl] = R(1) / s2;
                }
            }
            __syncwarp(hmask);
            bool xfail = false;
            if (hl < qp) {
                const R x0 = Xc[hl], sc = nrm[xj];
                const R v = sc < 0 ? x0 : Xn[hl] * sc;
                Xc[hl] = v;
                const R num = rabs(v - x0), den = rabs(x0) > lim ? x0 : lim;
                const bool ok = den > 0 ? (num < maxerr * den) : (num / den < maxerr);
                if (!ok) xfail = true;
            }
            outer++;
            const bool any_x = __any_sync(hmask, xfail);
            __syncwarp(hmask);
            if ((!O.fixed_iter && !any_x) || outer >= O.maxiter_x) phase = OUT;
            else { it = 0; phase = INNER; }
        }

        // ================= BT into the result rows  :1042
        if (phase == OUT) {
            if (hl < n) {
                double *o = vals + rpos[hl];
#pragma unroll 1
                for (int j = 0; j < p; j++) o[j] = (double)beta[j * n + hl];
            }
            if (hl == 0) task_done_stats(A.iters, A.stats, t, outer, inner_total, q, p, nc, n);
            __syncwarp(hmask);
            phase = LOAD;
        }
    }
}

}  // namespace scasa
