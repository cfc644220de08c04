"""String-level literal transcription of crpcount_td + hamming_correction.  TEST INFRASTRUCTURE.

A second, independent reading of /root/reference/scasa/SCRIPTS/Scasa_Rsource.R:416-648 that keeps
R's own data shapes: transcripts and patterns are *strings*, the per-block tables are lists of
rows, the Hamming tie-break goes through a character matrix exactly as `apply()` on a data.frame
with a character column does (as.matrix.data.frame -> format() with 7 significant digits ->
string comparison with "0" -> as.numeric -> median -> which.max).  It shares no code with
oracle/scasa_oracle_counts.c (integer ids, bit patterns, a numeric restatement of the tie-break)
nor with the product (scasa_b200/csrc/scasa_host.cpp), so a misreading in one of them shows up
as a disagreement in tests/test_crpcount_literal.py.

PARITY UNPINNED: no R in the image; tools/make_golden.R dumps the same quantities from R itself
(`hamming.tsv`, `y_counts.tsv`) and tests/test_r_golden.py compares all three implementations
against them when the dumps are present.

Only liberty taken for speed: `for (j in nID)` (:453) visits only the blocks that list one of the
cell's transcripts in their *name*; every other block yields an all-zero vector in R too
(`raw.c` empty at :473-475, `pick` empty at :461).

R semantics restated here (each next to the line that needs it):
  * format(<numeric>) with getOption("digits") = 7      -> r_format()
  * "<string>" > 0  (0 is coerced to "0"; C collation)   -> r_string_gt_zero()
  * median(), which.max() with NA                        -> r_median(), r_which_max()
  * sort(<character>) / unique() / tapply(..., c)        -> python sorted(key=collate) / dict order
"""
from __future__ import annotations

import math


# ----------------------------------------------------------------------------------------------
# R's format.default() on a double vector, digits = 7, nsmall = 0, scipen = 0
# (src/main/format.c: formatReal + scientific; src/main/paste.c: do_format pads to the common
# width, right-justified).  Returns the strings R would put into the character matrix.
def _sig7(a: float):
    """(kpower, nsig) of |x| = a > 0 at 7 significant digits: decimal exponent of the rounded
    value and the number of significant digits left after dropping trailing zeros."""
    m, e = ("%.6e" % a).split("e")          # d.dddddd, correctly rounded to 7 digits
    digits = m.replace(".", "")
    nsig = 7
    while nsig > 1 and digits[nsig - 1] == "0":
        nsig -= 1
    return int(e), nsig


def r_format(x: list) -> list:
    """format(x) for a list of finite doubles (the CRP values: 0 <= x <= 1, but written generally)."""
    neg = any(v < 0 for v in x)
    mxsl, rgt, mxe, mne, mxns = None, None, None, None, None
    for v in x:
        a = abs(v)
        if a == 0.0:
            kp, nsig = 0, 1
        else:
            kp, nsig = _sig7(a)
        left = kp + 1
        if kp > 0 and a < 10.0 ** kp:        # "roundingwidens": 7-digit rounding carried into a new digit
            left -= 1
        sleft = (1 if v < 0 else 0) + (1 if left <= 0 else left)
        right = nsig - kp - 1                # digits right of the point needed in fixed notation
        if mxsl is None:
            mxsl, rgt, mxe, mne, mxns = sleft, right, kp, kp, nsig
        else:
            mxsl, rgt = max(mxsl, sleft), max(rgt, right)
            mxe, mne, mxns = max(mxe, kp), min(mne, kp), max(mxns, nsig)
    if mxsl is None:
        return []
    rgt = max(rgt, 0)
    w_fixed = mxsl + rgt + (1 if rgt else 0)
    e_width = 2 if (mxe >= 100 or mne <= -99) else 1
    d_sci = mxns - 1
    w_sci = (1 if neg else 0) + (1 if d_sci > 0 else 0) + d_sci + 4 + e_width
    out = []
    for v in x:
        v = 0.0 if v == 0 else v             # no "-0"
        if w_fixed <= w_sci:
            s = "%.*f" % (rgt, v)
            out.append(s.rjust(w_fixed))
        else:
            s = "%.*e" % (d_sci, v)          # C and R both print at least two exponent digits
            out.append(s.rjust(w_sci))
    return out


def r_string_gt_zero(s: str) -> bool:
    """`s > 0` for a character s: 0 becomes "0" and the strings are compared (C collation = code
    points; the CRP strings are digits, '.', 'e', '-', '+': locale-invariant apart from padding,
    which cannot occur for values in [0, 1])."""
    return s > "0"


def r_median(v: list):
    """median(v); None stands for NA (length-0 input)."""
    if not v:
        return None
    s = sorted(v)
    n = len(s)
    if n % 2:
        return s[n // 2]
    return (s[n // 2 - 1] + s[n // 2]) / 2.0


def r_which_max(v: list):
    """which.max: index of the first maximum, NA / NaN skipped; None if there is none."""
    best, bi = None, None
    for i, a in enumerate(v):
        if a is None or (isinstance(a, float) and math.isnan(a)):
            continue
        if best is None or a > best:
            best, bi = a, i
    return bi


# ----------------------------------------------------------------------------------------------
class Block:
    """One element of the R list CRP: a q x p numeric matrix with dimnames."""

    def __init__(self, rownames: list, colnames: list, values: list):
        self.rownames = rownames            # 0/1 pattern strings
        self.colnames = colnames            # transcript names
        self.values = values                # values[r][j]
        self.name = " ".join(colnames)      # names(CRP)[k]  

    def colsums_positive(self) -> list:
        return [sum(row[j] for row in self.values) > 0 for j in range(len(self.colnames))]


def blocks_from_xmatrix(xm, which=None) -> dict:
    """{block index: Block} straight from the packed fixture (strings rebuilt from ids / bits)."""
    out = {}
    ks = range(xm.n_blocks) if which is None else which
    for k in ks:
        q, p = xm.q(k), xm.p_crp(k)
        pat = xm.crp_patterns(k)
        val = xm.crp_block(k)
        cn = [xm.tx_names[t] for t in xm.crp_tx[xm.crp_p_off[k]:xm.crp_p_off[k + 1]]]
        rn = ["".join("1" if b else "0" for b in pat[r]) for r in range(q)]
        out[int(k)] = Block(rn, cn, [[float(val[r, j]) for j in range(p)] for r in range(q)])
    return out


STATS = {"tie_breaks": 0}      # how often the median rule decided (tests report it)


# ----------------------------------------------------------------------------------------------
def hamming_assign(raw_pats: list, crp: Block, hamming_dist: int = 1) -> list:
    """raw_corrected$pat of hamming_correction (Rsource:562-629): for every observed pattern
    string the block row name it is moved to.  crp_data = data.frame(pat = rownames(crp), crp)."""
    crp_pats = crp.rownames
    raw_current = [[int(ch) for ch in s] for s in raw_pats]                 # :563
    crp_current = [[int(ch) for ch in s] for s in crp_pats]                 # :564
    compare_hamming = [[sum(1 for a, b in zip(x, y) if a != b) for y in crp_current] for x in raw_current]   # :567-575
    corrected = list(raw_pats)                                              # :604

    def best_by_median(rows: list) -> str:
        STATS["tie_breaks"] += 1
        # current <- crp_data[rows, ]; apply(current, 1, function(x) median(as.numeric(as.character(x[x > 0]))))
        # apply() -> as.matrix(<data.frame with a character column>) -> every numeric column goes
        # through format() on the selected rows only; the matrix is character.
        cols = [r_format([crp.values[r][j] for r in rows]) for j in range(len(crp.colnames))]
        meds = []
        for i, r in enumerate(rows):
            x = [crp_pats[r]] + [cols[j][i] for j in range(len(crp.colnames))]   # one row of the character matrix
            kept = [s for s in x if r_string_gt_zero(s)]                          # x[x > 0]
            meds.append(r_median([float(s) for s in kept]))                      # as.numeric(as.character(.))
        return crp_pats[rows[r_which_max(meds)]]                                 # row.names(current)[which.max(.)]

    for i, d in enumerate(compare_hamming):                                  # :606
        zero = [r for r, v in enumerate(d) if v == 0]
        if len(zero) > 0:                                                    # :609-610
            assert len(zero) == 1, "duplicate row names in a block"
            corrected[i] = crp_pats[zero[0]]
            continue
        close = [r for r, v in enumerate(d) if v <= hamming_dist]           # :612
        if len(close) > 1:                                                   # :613-618
            corrected[i] = best_by_median(close)
        if len(close) == 1:                                                  # :620-622
            corrected[i] = crp_pats[close[0]]
        if len(close) == 0:                                                  # :624-627
            mn = min(d)
            corrected[i] = best_by_median([r for r, v in enumerate(d) if v <= mn])
    return corrected


def hamming_correction(raw4: list, crp: Block, hamming_dist: int = 1) -> dict:
    """raw4 = [(pat, sample1)] -> {pat: sum of sample1} (Rsource:643-646: tapply(sample1, pat, sum))."""
    corrected = hamming_assign([p for p, _ in raw4], crp, hamming_dist)
    a: dict = {}
    for pat, (_, s) in zip(corrected, raw4):
        a[pat] = a.get(pat, 0) + s
    return a


# ----------------------------------------------------------------------------------------------
class CrpLiteral:
    """The global `CRP` plus what crpcount_td derives from it on every call (:423-427)."""

    def __init__(self, xm, collate=None):
        self.xm = xm
        self.collate = collate or (lambda s: s.encode())       # sort(): LC_COLLATE=C by default
        self._blocks: dict = {}
        # c2t = sapply(names(CRP), strsplit, split = ' ');  t2c = tapply(names(t2c), t2c, c)
        self.names = xm.block_names()
        self.t2c: dict = {}
        for k, nm in enumerate(self.names):
            for t in nm.split(" "):
                self.t2c.setdefault(t, []).append(k)           # block order == names(CRP) order inside a group
        self.index_of = {nm: k for k, nm in enumerate(self.names)}

    def block(self, k: int) -> Block:
        b = self._blocks.get(k)
        if b is None:
            b = blocks_from_xmatrix(self.xm, [k])[k]
            self._blocks[k] = b
        return b

    def c2t(self, name: str) -> list:
        return name.split(" ")

    def crpcount_td(self, mytd: list, hamming_dist: int = 1) -> dict:
        """mytd: rows (Transcript: str, Count: int, eqClass: int) of ONE cell, in table order.
        Returns {block index: [count per block row]} for the blocks with a non-zero vector."""
        rawcount = [{"Transcript": str(t), "Count": c, "eqClass": e} for t, c, e in mytd]
        myeqID = sorted(set(r["eqClass"] for r in rawcount))                  # :438-441
        match = {e: i + 1 for i, e in enumerate(myeqID)}
        for r in rawcount:
            r["eqClass"] = match[r["eqClass"]]
        table: dict = {}
        for r in rawcount:
            table[r["eqClass"]] = table.get(r["eqClass"], 0) + 1               # :444
        for r in rawcount:
            r["Weight"] = 1.0 / table[r["eqClass"]]                            # :445-446
        visit = sorted({k for r in rawcount for k in self.t2c.get(r["Transcript"], [])})
        # row lookups by value (what `rawcount$Transcript == t` / `$eqClass == e` select, in table order)
        by_tx: dict = {}
        by_eq: dict = {}
        for r in rawcount:
            by_tx.setdefault(r["Transcript"], []).append(r)
            by_eq.setdefault(r["eqClass"], []).append(r)
        result = {}
        for j in visit:                                                        # :453 (see module docstring)
            crp = self.block(j)
            if len(crp.rownames) * len(crp.colnames) == 1:                     # length(crp) == 1  :457
                tx = crp.colnames
                pick = [r for t in tx for r in by_tx.get(t, []) if r["Weight"] > 0.9]           # :461
                sample1 = sum(r["Count"] for r in pick) if pick else 0
                if sample1:
                    result[j] = [sample1]
                continue
            tx = crp.colnames                                                  # :466
            pick = crp.colsums_positive()                                      # :468
            tx1 = [t for t, ok in zip(tx, pick) if ok]                         # :469
            raw_c = []
            for t in tx1:                                                      # :473-474
                raw_c.extend(by_tx.get(t, []))
            eq = []
            for r in raw_c:                                                    # :475 unique(), first appearance
                if r["eqClass"] not in eq:
                    eq.append(r["eqClass"])
            if eq:                                                             # :477
                current_name = crp.name
                # :479 myRawcount = rawcount[rawcount$eqClass %in% eq, ] (only ever filtered by eqClass below)
                rm = []
                for i, e in enumerate(eq):                                     # :481
                    mytx = sorted((r["Transcript"] for r in by_eq[e]), key=self.collate)   # :482-483
                    cand = []
                    for t in mytx:                                             # unlist(t2c[mytx])
                        cand.extend(self.names[k] for k in self.t2c.get(t, []))
                    eq_candidates = sorted(set(cand), key=self.collate)        # :484
                    keep = []
                    for nm in eq_candidates:                                   # :488-492
                        acrp = self.block(self.index_of[nm])
                        acrp_tx = [t for t, ok in zip(acrp.colnames, acrp.colsums_positive()) if ok]
                        if sum(1 for t in acrp_tx if t in mytx) > 0:
                            keep.append(nm)
                    eq_candidates = keep                                       # :493
                    myscore = []
                    for nm in eq_candidates:                                   # :497-504
                        x = self.c2t(nm)
                        sc = sum(1 for t in x if t in mytx) / len(set(x) | set(mytx))
                        myscore.append(sc)
                    best = eq_candidates[r_which_max(myscore)]                 # :505
                    if current_name != best:                                   # :506
                        rm.append(i)
                eq = [e for i, e in enumerate(eq) if i not in rm]              # :508
                if eq:
                    raw_c = [r for r in raw_c if r["eqClass"] in eq]          # :510
            if not eq:                                                         # :542-543
                continue
            raw2 = {e: {t: 0 for t in tx} for e in eq}                         # :516-518
            for r in raw_c:                                                    # :519-523
                raw2[r["eqClass"]][r["Transcript"]] = r["Count"]
            pat = {e: "".join("1" if raw2[e][t] > 0 else "0" for t in tx) for e in eq}   # :524-525
            a1 = {e: max(raw2[e][t] for t in tx) for e in eq}                  # :528
            a2: dict = {}
            for e in eq:                                                       # :529 tapply(a1, pat, sum)
                a2[pat[e]] = a2.get(pat[e], 0) + a1[e]
            raw4 = [(p, a2[p]) for p in sorted(a2)]                            # :531 (tapply sorts the group names)
            corrected = hamming_correction(raw4, crp, hamming_dist)            # :534
            y = [corrected.get(rn, 0) for rn in crp.rownames]                  # :536-540 merge(all.x), NA -> 0, crp row order
            if any(y):
                result[j] = y
        return result

    # ---- the per-eqclass view the product's scasa_eqclass_map must equal -----------------------
    def eqclass_row(self, tx_names: list) -> int:
        """Global row a class with this transcript set adds its count to when it is the only class of
        a cell (count 1), or -1 if crpcount_td drops it."""
        res = self.crpcount_td([(t, 1, 1) for t in tx_names])
        hits = [(k, r) for k, y in res.items() for r, v in enumerate(y) if v]
        assert len(hits) <= 1, f"a class was counted in {len(hits)} rows: {hits}"
        if not hits:
            return -1
        k, r = hits[0]
        return int(self.xm.q_off[k]) + r
