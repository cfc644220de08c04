"""Packed form of Scasa's X-matrix (the block-diagonal list ``CRP`` / ``CCRP``).

The reference keeps the X-matrix as an R list of small dense matrices, one per
transcript cluster ("CRP", scasa/SCRIPTS/scasa_quant_step2_v1.0.0.R:39):
rows = equivalence-class membership patterns (row names are 0/1 strings over
the block's columns), columns = transcripts, values column-stochastic.  ``DM0``
is its paralog-merged version, ``lapply(CRP, ccrpfun, 30)``
(scasa_quant_step2_v1.0.0.R:127; Scasa_Rsource.R:168-193), which equals the
shipped ``CCRP``.

Here every block is flattened into offset arrays, which is exactly the layout
``include/scasa_cuda.h`` takes (``scasa_xmat_t`` / ``scasa_crp_t``):

* ``q_off[B+1]``      first global row of each block (rows are shared by CRP and DM0)
* ``crp_p_off[B+1]``  first CRP column of each block;  ``crp_x`` column-major values
* ``crp_tx[Σp]``      transcript id of each CRP column (index into ``tx_names``)
* ``crp_pat[Σ q·p]``  0/1 row patterns, row-major inside a block (the row *names*,
  which differ from the support of the values in 4,206 rows)
* ``dm0_p_off[B+1]``, ``dm0_x``   the same for DM0
* ``crp_member[Σp]``  DM0 column (global index) a CRP column was merged into, -1 if dropped
* ``name_rank[B]``    rank of ``names(CRP)[k]`` under ``sort()`` (C collation here; the R
  caller passes its own ``order(names(CRP))`` so the session locale is honoured)
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from .rdata import load_rdata


@dataclass
class XMatrix:
    q_off: np.ndarray        # int32 [B+1]
    crp_p_off: np.ndarray    # int32 [B+1]
    crp_x_off: np.ndarray    # int64 [B+1]
    crp_x: np.ndarray        # float64, column-major per block
    crp_tx: np.ndarray       # int32 [Σp_crp]
    crp_pat: np.ndarray      # uint8 [Σ q·p_crp], row-major per block
    dm0_p_off: np.ndarray    # int32 [B+1]
    dm0_x_off: np.ndarray    # int64 [B+1]
    dm0_x: np.ndarray        # float64, column-major per block
    crp_member: np.ndarray   # int32 [Σp_crp]
    name_rank: np.ndarray    # int32 [B]
    tx_names: list           # distinct transcript names, id order
    gene_of_col: np.ndarray | None = None   # int32 [Σp_dm0] gene id of each DM0 column, -1 = unmapped
    gene_names: list | None = None

    # ------------------------------------------------------------------ sizes
    @property
    def n_blocks(self) -> int:
        return len(self.q_off) - 1

    @property
    def n_rows(self) -> int:
        return int(self.q_off[-1])

    @property
    def n_cols(self) -> int:
        """Σp of DM0 = number of output (isoform) columns."""
        return int(self.dm0_p_off[-1])

    def q(self, k: int) -> int:
        return int(self.q_off[k + 1] - self.q_off[k])

    def p(self, k: int) -> int:
        return int(self.dm0_p_off[k + 1] - self.dm0_p_off[k])

    def p_crp(self, k: int) -> int:
        return int(self.crp_p_off[k + 1] - self.crp_p_off[k])

    def dm0_block(self, k: int) -> np.ndarray:
        """DM0[[k]] as a (q, p) array."""
        q, p = self.q(k), self.p(k)
        o = int(self.dm0_x_off[k])
        return self.dm0_x[o:o + q * p].reshape(p, q).T

    def crp_block(self, k: int) -> np.ndarray:
        q, p = self.q(k), self.p_crp(k)
        o = int(self.crp_x_off[k])
        return self.crp_x[o:o + q * p].reshape(p, q).T

    def crp_patterns(self, k: int) -> np.ndarray:
        """Row-name patterns of block k as a (q, p_crp) 0/1 array."""
        q, p = self.q(k), self.p_crp(k)
        o = int(self.crp_x_off[k])
        return self.crp_pat[o:o + q * p].reshape(q, p)

    def dm0_colnames(self) -> list:
        """colnames of every DM0 column: members joined by ' ' in CRP column order
        (Scasa_Rsource.R:189-190)."""
        out = [[] for _ in range(self.n_cols)]
        for c, m in enumerate(self.crp_member):
            if m >= 0:
                out[m].append(self.tx_names[self.crp_tx[c]])
        return [" ".join(v) for v in out]

    def block_names(self) -> list:
        """names(CRP): CRP column names joined by ' ' (xmatrix step4:469)."""
        return [" ".join(self.tx_names[t] for t in self.crp_tx[self.crp_p_off[k]:self.crp_p_off[k + 1]])
                for k in range(self.n_blocks)]

    def dm0_is_ccrpfun_fixed_point(self, clim: float = 30.0, n_threads: int = 0):
        """N3 (SURVEY.md §8 f): can ``DM0 <- lapply(CRP, ccrpfun, clim)`` (scasa_quant_step2_v1.0.0.R:127,
        serial minutes of svd + kmeans in the R master) be skipped in favour of the shipped ``CCRP``?
        Calls the library's host-only ``scasa_dm0_is_fixed_point`` (the same entry the R shim uses): True when
        every block of DM0 is a point ccrpfun's loop (Scasa_Rsource.R:168-193) stops at.  Returns
        (is_fixed_point, first_bad_block or -1)."""
        import ctypes as C
        from . import lib as _l
        L = _l.load()
        keep = [np.ascontiguousarray(a, t) for a, t in ((self.q_off, np.int32), (self.crp_p_off, np.int32), (self.crp_x_off, np.int64),
                                                        (self.crp_x, np.float64), (self.dm0_p_off, np.int32), (self.dm0_x_off, np.int64),
                                                        (self.dm0_x, np.float64), (self.crp_member, np.int32))]
        crp, dm0 = _l.XMat(), _l.XMat()
        crp.n_blocks = dm0.n_blocks = self.n_blocks
        crp.q_off = dm0.q_off = _l.ptr(keep[0], C.c_int32)
        crp.p_off, crp.x_off, crp.x = _l.ptr(keep[1], C.c_int32), _l.ptr(keep[2], C.c_int64), _l.ptr(keep[3], C.c_double)
        dm0.p_off, dm0.x_off, dm0.x = _l.ptr(keep[4], C.c_int32), _l.ptr(keep[5], C.c_int64), _l.ptr(keep[6], C.c_double)
        ok, bad = C.c_int32(0), C.c_int64(-1)
        rc = L.scasa_dm0_is_fixed_point(C.byref(crp), C.byref(dm0), _l.ptr(keep[7], C.c_int32), float(clim), n_threads,
                                        C.byref(ok), C.byref(bad))
        if rc:
            raise _l.ScasaError(rc, L.scasa_last_error(None).decode())
        return bool(ok.value), int(bad.value)

    def em_blocks(self) -> np.ndarray:
        """Blocks that go through the AEM loop: nrow(X0) > 1 (Scasa_Rsource.R:979)."""
        return np.nonzero(np.diff(self.q_off) > 1)[0].astype(np.int32)

    # ------------------------------------------------------------- persistence
    _ARRAYS = ("q_off", "crp_p_off", "crp_x_off", "crp_x", "crp_tx", "crp_pat",
               "dm0_p_off", "dm0_x_off", "dm0_x", "crp_member", "name_rank")

    def save_npz(self, path: str) -> None:
        d = {k: getattr(self, k) for k in self._ARRAYS}
        d["crp_pat"] = np.packbits(self.crp_pat)
        d["n_pat"] = np.int64(len(self.crp_pat))
        d["tx_names"] = np.frombuffer("\n".join(self.tx_names).encode(), dtype=np.uint8)
        if self.gene_of_col is not None:
            d["gene_of_col"] = self.gene_of_col
            d["gene_names"] = np.frombuffer("\n".join(self.gene_names).encode(), dtype=np.uint8)
        np.savez_compressed(path, **d)

    @classmethod
    def load_npz(cls, path: str) -> "XMatrix":
        z = np.load(path)
        kw = {k: z[k] for k in cls._ARRAYS}
        kw["crp_pat"] = np.unpackbits(z["crp_pat"])[: int(z["n_pat"])]
        kw["tx_names"] = z["tx_names"].tobytes().decode().split("\n")
        if "gene_of_col" in z:
            kw["gene_of_col"] = z["gene_of_col"]
            kw["gene_names"] = z["gene_names"].tobytes().decode().split("\n")
        return cls(**kw)

    # --------------------------------------------------------------- from R
    @classmethod
    def from_rdata(cls, xmatrix_path: str, groups_path: str | None = None, collate=None) -> "XMatrix":
        """Flatten ``CRP`` and ``CCRP`` of an ``Xmatrix_*.RData``; optionally attach the
        tx→gene map ``genes.tx.map.all.final`` of an ``isoform_groups_*.RData``
        (scasa_quant_step3_v1.0.0.R:22,35).

        ``name_rank`` (the order ``sort()`` puts the block names in, which breaks Jaccard ties at Rsource:484-505)
        is taken over ``names(CRP)`` with ``collate`` as the sort key; the default, bytes, is ``LC_COLLATE=C``.
        R sorts in the session's collation (en_US ignores spaces and punctuation at the first level), so a
        caller that must match such a session passes its own key; the R shim passes R's own ranks."""
        ws = load_rdata(xmatrix_path)
        crp_l, ccrp_l = ws["CRP"].value, ws["CCRP"].value
        crp_names = list(ws["CRP"].names) if getattr(ws["CRP"], "names", None) is not None else None
        B = len(crp_l)
        q_off = np.zeros(B + 1, np.int32)
        cp_off = np.zeros(B + 1, np.int32)
        cx_off = np.zeros(B + 1, np.int64)
        dp_off = np.zeros(B + 1, np.int32)
        dx_off = np.zeros(B + 1, np.int64)
        tx_id: dict = {}
        crp_x, crp_tx, crp_pat, dm0_x, member = [], [], [], [], []
        for k in range(B):
            c, m = crp_l[k], ccrp_l[k]
            q, p = c.dim
            q2, p2 = m.dim
            rn, cn = c.dimnames
            rn2, cn2 = m.dimnames
            if q2 != q or list(rn2) != list(rn):
                raise ValueError(f"block {k}: CCRP rows differ from CRP rows")
            q_off[k + 1] = q_off[k] + q
            cp_off[k + 1] = cp_off[k] + p
            cx_off[k + 1] = cx_off[k] + q * p
            dp_off[k + 1] = dp_off[k] + p2
            dx_off[k + 1] = dx_off[k] + q * p2
            crp_x.append(np.asarray(c.value, np.float64))
            dm0_x.append(np.asarray(m.value, np.float64))
            for t in cn:
                crp_tx.append(tx_id.setdefault(t, len(tx_id)))
            pat = np.frombuffer("".join(rn).encode(), np.uint8) - ord("0")
            if pat.size != q * p or pat.max(initial=0) > 1:
                raise ValueError(f"block {k}: row names are not 0/1 patterns of width ncol")
            crp_pat.append(pat)
            where = {}
            for j, name in enumerate(cn2):
                for t in name.split(" "):
                    where[t] = int(dp_off[k]) + j
            member.extend(where.get(t, -1) for t in cn)
        names = [" ".join(c.dimnames[1]) for c in crp_l]
        if crp_names is not None and crp_names != names:          # K1: t2c / c2t are built from names(CRP) (Rsource:423-427)
            raise ValueError("names(CRP) are not the blocks' column names joined by ' ' (SURVEY.md K1)")
        key = collate or (lambda s: s.encode())
        order = sorted(range(B), key=lambda i: key(names[i]))
        rank = np.empty(B, np.int32)
        rank[order] = np.arange(B, dtype=np.int32)
        xm = cls(q_off, cp_off, cx_off, np.concatenate(crp_x), np.asarray(crp_tx, np.int32),
                 np.concatenate(crp_pat).astype(np.uint8), dp_off, dx_off, np.concatenate(dm0_x),
                 np.asarray(member, np.int32), rank, list(tx_id))
        if groups_path is not None:
            g = load_rdata(groups_path)["genes.tx.map.all.final"]
            gmap = dict(zip(g.names, g.value))
            gid: dict = {}
            goc = np.full(xm.n_cols, -1, np.int32)
            for j, name in enumerate(xm.dm0_colnames()):
                gname = gmap.get(name)
                if gname is not None:
                    goc[j] = gid.setdefault(gname, len(gid))
            xm.gene_of_col, xm.gene_names = goc, list(gid)
        return xm
