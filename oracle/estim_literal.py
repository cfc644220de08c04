"""Literal numpy transcription of estim.BETA / estim.X.em / estim_chunk.  TEST INFRASTRUCTURE.

Second, independent statement of the reference arithmetic, written to keep R's
*shapes* (the mapply over cells, the Kronecker form of estim.X.em, c(t(BT)) ...)
rather than a closed form, and using SciPy's digamma (Cephes psi) instead of the C
oracle's restatement of R's dpsifn.  It exists to pin oracle/scasa_oracle.c
(tests/test_oracle.py); it is slow and is never used to check the GPU directly at
size.  PARITY UNPINNED against R itself (R is not available in the build image).

Line numbers refer to scasa/SCRIPTS/Scasa_Rsource.R.
"""
import numpy as np
from scipy.special import digamma


def estim_BETA(X, Ymat, beta0, maxiter=100, maxerr=0.01, lim=0.01, modify=True, fixed_iter=False):
    """:726-761.  X (q,p); Ymat (q,n); beta0 vector of length n*p (cell-major).  Returns (n,p)."""
    q, n = X.shape[0], Ymat.shape[1]
    Y_list = [Ymat[:, c] for c in range(n)]                       # :728
    csum = X.sum(axis=0)                                          # :731
    totCount = Ymat.sum(axis=0)                                   # :732
    logNorm = digamma(totCount + 0.00000001)                      # :733

    def fn1(bt, logNorm_i):                                       # :736-742
        gamma0 = bt
        if modify:
            gamma0 = np.exp(digamma(bt + 0.00000001) - logNorm_i)
        tmp = (X.T * gamma0[:, None]).T                           # t(t(X)*gamma0)
        rsum = tmp.sum(axis=1)
        return tmp / np.where(rsum > 0, rsum, 1.0)[:, None]

    def fn2(Xi, yi):                                              # :744-747
        with np.errstate(divide="ignore", invalid="ignore"):
            return (Xi.T @ yi) / csum

    beta0 = np.asarray(beta0, float).copy()
    niter = 0
    for _ in range(maxiter):                                      # :749
        niter += 1
        BETA = beta0.reshape(n, -1)                               # byrow=TRUE :750
        X2_list = [fn1(BETA[c], logNorm[c]) for c in range(n)]    # :752
        beta = np.concatenate([fn2(X2_list[c], Y_list[c]) for c in range(n)])  # p×n col-major == cell-major
        beta = np.where(np.isnan(beta), 0.0, beta)                # :754
        err = np.abs(beta - beta0) / np.where(beta0 > lim, beta0, lim)   # :755
        if not fixed_iter and err.max() < maxerr:                 # :756
            break
        beta0 = beta
    return beta.reshape(n, -1), niter                             # :760


def estim_X_em(X_em, BETA, Ymat):
    """:898-922, Kronecker form kept."""
    X0 = X_em
    Y = Ymat.T.reshape(-1)          # c(Ymat): column-major = cell-major, index c*q + r   :900
    q, n, p = Ymat.shape[0], Ymat.shape[1], BETA.shape[1]
    BTsum = np.repeat(BETA.sum(axis=0), q)                        # :904
    x1 = np.stack([np.kron(BETA[:, j], X_em[:, j]) for j in range(p)], axis=1)   # :906  (n*q, p)
    sumx1 = x1.sum(axis=1)                                        # :907
    sumx1_pos = np.where(sumx1 > 0, sumx1, 0.1)                   # :908
    x2 = x1 / sumx1_pos[:, None]                                  # :909
    Ycell = x2 * Y[:, None]                                       # :910
    tmp = np.concatenate([Ycell[:, j].reshape(n, q).sum(axis=0) for j in range(p)])  # :912-914
    X_new = (tmp / np.where(BTsum > 0, BTsum, 0.1)).reshape(p, q).T                 # :917 matrix(., ncol=p)
    csum = X_new.sum(axis=0)                                      # :918
    X_new = (X_new.T / np.where(csum > 0, csum, 0.1)[:, None]).T  # :919
    for j in range(p):                                            # :920
        if csum[j] < 0.00001:
            X_new[:, j] = X0[:, j]
    return X_new


def cos_sim(a, b):                                                # :950
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.sum(a * b) / np.sqrt(np.sum(a ** 2) * np.sum(b ** 2))


def estim_chunk(Y, DM0, maxiter_X=1000, maxerr=0.01, lim=0.01, maxiter_bt=1000, modify=True,
                fixed_iter=False, return_iters=False):
    """:944-1051.  Y, DM0: lists of (q_k, n) and (q_k, p_k) arrays.  Returns (n, Σp)."""
    poorCondX = []
    for m, X0 in enumerate(DM0):                                  # :957-971
        if X0.shape[1] == 2:
            isPoor = True
            for i in range(2):
                sim = np.array([cos_sim(X0[:, k], X0[:, i]) for k in range(2)])
                if (sim < 0.99).sum():
                    isPoor = False
                    break
            if isPoor:
                poorCondX.append(m)
    adjust_esp = 2
    final, iters = [], []
    for m in range(len(Y)):                                       # :975
        Ymat = np.array(Y[m], float)
        X0 = DM0[m]
        if X0.shape[0] == 1:                                      # :979-983
            final.append(Ymat.T)
            iters.append((0, 0))
            continue
        needAdjustX = False
        if m in poorCondX:                                        # :989-1008
            rs = Ymat.sum(axis=1)
            with np.errstate(divide="ignore", invalid="ignore"):
                rs = rs / rs.sum()
            sim = np.array([cos_sim(X0[:, j], rs) for j in range(X0.shape[1])])
            if not np.all(np.isnan(sim)):
                bestTx = int(np.nanargmax(sim))
                adjID = np.nonzero(X0[:, bestTx] > 0)[0]
                if len(adjID) > 0:
                    needAdjustX = True
                    Ymat[adjID, :] += adjust_esp
        X_est0 = X0.copy()                                        # :1011
        Ysum = Ymat.sum(axis=0)
        p = X0.shape[1]
        beta0 = np.repeat(Ysum / p, p)                            # :1013
        n_out = n_in = 0
        for _ in range(maxiter_X):                                # :1015
            n_out += 1
            BT, k = estim_BETA(X_est0, Ymat, beta0, maxiter=maxiter_bt, modify=modify, fixed_iter=fixed_iter)   # :1016: maxerr / lim are NOT forwarded
            n_in += k
            X_est = estim_X_em(X_est0, BT, Ymat)
            err = np.abs(X_est - X_est0) / np.where(np.abs(X_est0) > lim, X_est0, lim)
            if not fixed_iter and err.max() < maxerr:
                break
            X_est0 = X_est
            beta0 = BT.reshape(-1)                                # c(t(BT))
        if needAdjustX:                                           # :1027-1037
            deductVal = len(adjID) * adjust_esp
            with np.errstate(divide="ignore", invalid="ignore"):
                BT = np.stack([x - deductVal * (x / x.sum()) for x in BT])
        final.append(BT)
        iters.append((n_out, n_in))
    out = np.concatenate(final, axis=1)                           # :1042
    return (out, iters) if return_iters else out
