"""Reader for R's ``save()`` / ``saveRDS()`` files (RDX3 / RDS, XDR serialisation).

Scasa's inputs are R workspaces: the X-matrix (``CRP``, ``CCRP``, ``CRPCOUNT``;
written at scasa/SCRIPTS/Xmatrix_Gen/scasa_simulate_xmatrix_step4_v1.0.0.R:477,
loaded at scasa/SCRIPTS/scasa_quant_step2_v1.0.0.R:39) and the isoform/gene map
(``genes.tx.map.all.final``; aux/categorise_isoforms.R:152, loaded at
scasa/SCRIPTS/scasa_quant_step3_v1.0.0.R:22).  R is not part of this project's
toolchain, so the harness reads those files itself.  Host-side tooling only:
nothing on the CUDA path depends on it (the R caller hands the library flat
arrays).

Only what those files contain is decoded: NULL, symbols, pairlists, character,
logical, integer, real and generic vectors, attributes, reference back-pointers,
external pointers (skipped payload) and the three ALTREP classes base R emits
(``compact_intseq``, ``compact_realseq``, ``wrap_*``, ``deferred_string``).
"""
from __future__ import annotations

import gzip
import bz2
import lzma
import struct
from dataclasses import dataclass, field
from typing import Any

import numpy as np

# SEXP type codes (R internals, Rinternals.h) and serialize.c pseudo-types.
NILSXP, SYMSXP, LISTSXP, CLOSXP, ENVSXP, PROMSXP, LANGSXP = 0, 1, 2, 3, 4, 5, 6
CHARSXP, LGLSXP, INTSXP, REALSXP, CPLXSXP, STRSXP = 9, 10, 13, 14, 15, 16
DOTSXP, VECSXP, EXPRSXP, BCODESXP, EXTPTRSXP, WEAKREFSXP, RAWSXP = 17, 19, 20, 21, 22, 23, 24
REFSXP, NILVALUE_SXP, GLOBALENV_SXP, UNBOUNDVALUE_SXP, MISSINGARG_SXP = 255, 254, 253, 252, 251
BASENAMESPACE_SXP, NAMESPACESXP, PACKAGESXP, PERSISTSXP = 250, 249, 248, 247
EMPTYENV_SXP, BASEENV_SXP, ATTRLANGSXP, ATTRLISTSXP, ALTREP_SXP = 242, 241, 240, 239, 238

NA_INTEGER = -2147483648


class RDataError(ValueError):
    pass


@dataclass
class RObject:
    """A decoded R value plus its attributes (``names``, ``dim``, ``dimnames`` ...)."""

    value: Any
    attrs: dict = field(default_factory=dict)
    sexptype: int = NILSXP

    # conveniences used by the X-matrix loader -------------------------------
    @property
    def names(self):
        n = self.attrs.get("names")
        return None if n is None else n.value

    @property
    def dim(self):
        d = self.attrs.get("dim")
        return None if d is None else tuple(int(x) for x in d.value)

    @property
    def dimnames(self):
        d = self.attrs.get("dimnames")
        if d is None:
            return None
        return [None if e is None or e.value is None else e.value for e in d.value]

    def as_matrix(self) -> np.ndarray:
        """Column-major R matrix → C-contiguous numpy array of shape ``dim``."""
        d = self.dim
        if d is None or len(d) != 2:
            raise RDataError("not a matrix")
        return np.asarray(self.value).reshape(d[1], d[0]).T


class _Reader:
    def __init__(self, buf: bytes, pos: int = 0):
        self.buf = memoryview(buf)
        self.pos = pos
        self.refs: list = []

    # primitive XDR reads -----------------------------------------------------
    def i32(self) -> int:
        v = struct.unpack_from(">i", self.buf, self.pos)[0]
        self.pos += 4
        return v

    def length(self) -> int:
        n = self.i32()
        if n == -1:
            hi, lo = self.i32(), self.i32()
            n = (hi << 32) + (lo & 0xFFFFFFFF)
        return n

    def raw(self, n: int) -> bytes:
        b = self.buf[self.pos:self.pos + n].tobytes()
        if len(b) != n:
            raise RDataError("truncated stream")
        self.pos += n
        return b

    # items ---------------------------------------------------------------------
    def item(self) -> RObject | None:
        flags = self.i32()
        t = flags & 0xFF
        has_attr = bool(flags & (1 << 9))
        has_tag = bool(flags & (1 << 10))

        if t == NILVALUE_SXP:
            return None
        if t in (GLOBALENV_SXP, EMPTYENV_SXP, BASEENV_SXP, BASENAMESPACE_SXP,
                 UNBOUNDVALUE_SXP, MISSINGARG_SXP):
            return RObject(None, sexptype=t)
        if t == REFSXP:
            idx = flags >> 8
            if idx == 0:
                idx = self.i32()
            return self.refs[idx - 1]
        if t == SYMSXP:
            ch = self.item()
            sym = RObject(ch.value if ch is not None else None, sexptype=SYMSXP)
            self.refs.append(sym)
            return sym
        if t in (NAMESPACESXP, PACKAGESXP, PERSISTSXP):
            self.i32()  # always 0
            n = self.i32()
            info = [self.item() for _ in range(n)]
            obj = RObject([None if s is None else s.value for s in info], sexptype=t)
            self.refs.append(obj)
            return obj
        if t == ENVSXP:
            self.i32()  # locked
            env = RObject({}, sexptype=ENVSXP)
            self.refs.append(env)
            self.item()  # enclos
            frame = self.item()
            self.item()  # hashtab
            self.item()  # attrib
            if frame is not None and frame.sexptype == LISTSXP:
                env.value = dict(frame.value)
            return env
        if t in (LISTSXP, LANGSXP, CLOSXP, PROMSXP, DOTSXP, ATTRLISTSXP, ATTRLANGSXP):
            # pairlist: walked iteratively, kept as a list of (tag, value)
            out = []
            attrs = {}
            while True:
                if t in (ATTRLISTSXP, ATTRLANGSXP) or has_attr:
                    a = self.item()
                    if a is not None:
                        attrs.update(_pairs_to_dict(a))
                tag = None
                if has_tag:
                    tg = self.item()
                    tag = tg.value if tg is not None else None
                car = self.item()
                out.append((tag, car))
                flags = self.i32()
                t = flags & 0xFF
                has_attr = bool(flags & (1 << 9))
                has_tag = bool(flags & (1 << 10))
                if t == NILVALUE_SXP:
                    break
                if t not in (LISTSXP, LANGSXP, ATTRLISTSXP, ATTRLANGSXP):
                    # dotted tail: rewind and read as a plain item
                    self.pos -= 4
                    out.append((None, self.item()))
                    break
            return RObject(out, attrs, LISTSXP)
        if t == EXTPTRSXP:
            obj = RObject(None, sexptype=EXTPTRSXP)
            self.refs.append(obj)
            self.item()  # prot
            self.item()  # tag
            if has_attr:
                obj.attrs = _pairs_to_dict(self.item())
            return obj
        if t == ALTREP_SXP:
            info = self.item()
            state = self.item()
            attr = self.item()
            obj = _altrep(info, state)
            if attr is not None:
                obj.attrs.update(_pairs_to_dict(attr))
            return obj

        # plain vectors -------------------------------------------------------
        if t == CHARSXP:
            n = self.i32()
            if n == -1:
                obj = RObject(None, sexptype=CHARSXP)  # NA_character_
            else:
                b = self.raw(n)
                enc = "latin-1" if flags & (1 << 14) else "utf-8"  # LATIN1_MASK in gp
                try:
                    s = b.decode(enc)
                except UnicodeDecodeError:
                    s = b.decode("latin-1")
                obj = RObject(s, sexptype=CHARSXP)
        elif t in (LGLSXP, INTSXP):
            n = self.length()
            v = np.frombuffer(self.buf, dtype=">i4", count=n, offset=self.pos).astype(np.int32)
            self.pos += 4 * n
            obj = RObject(v, sexptype=t)
        elif t == REALSXP:
            n = self.length()
            v = np.frombuffer(self.buf, dtype=">f8", count=n, offset=self.pos).astype(np.float64)
            self.pos += 8 * n
            obj = RObject(v, sexptype=t)
        elif t == CPLXSXP:
            n = self.length()
            v = np.frombuffer(self.buf, dtype=">c16", count=n, offset=self.pos).astype(np.complex128)
            self.pos += 16 * n
            obj = RObject(v, sexptype=t)
        elif t == RAWSXP:
            n = self.length()
            obj = RObject(self.raw(n), sexptype=t)
        elif t == STRSXP:
            n = self.length()
            obj = RObject(self._strvec(n), sexptype=t)
        elif t in (VECSXP, EXPRSXP):
            n = self.length()
            obj = RObject([self.item() for _ in range(n)], sexptype=t)
        else:
            raise RDataError(f"unsupported SEXP type {t} at byte {self.pos}")

        if has_attr:
            a = self.item()
            if a is not None:
                obj.attrs = _pairs_to_dict(a)
        return obj

    def _strvec(self, n: int) -> list:
        """STRSXP payload: n CHARSXP items.  Inlined fast path (millions of strings)."""
        out = [None] * n
        buf, pos = self.buf, self.pos
        unpack = struct.unpack_from
        for i in range(n):
            flags, ln = unpack(">ii", buf, pos)
            pos += 8
            if (flags & 0xFF) != CHARSXP:
                # not a plain CHARSXP (never seen in practice) → generic path
                self.pos = pos - 8
                it = self.item()
                out[i] = None if it is None else it.value
                pos = self.pos
                continue
            if ln == -1:
                continue
            b = bytes(buf[pos:pos + ln])
            pos += ln
            try:
                out[i] = b.decode("latin-1" if flags & (1 << 14) else "utf-8")
            except UnicodeDecodeError:
                out[i] = b.decode("latin-1")
        self.pos = pos
        return out


def _pairs_to_dict(pl: RObject | None) -> dict:
    if pl is None:
        return {}
    return {k: v for k, v in pl.value if k is not None}


def _format_g15(x: float) -> str:
    """as.character(double) for the integer-valued reals deferred_string wraps."""
    if x == int(x) and abs(x) < 1e15:
        return str(int(x))
    return repr(x)


def _altrep(info: RObject, state: RObject | None) -> RObject:
    cls = info.value[0][1].value  # class symbol
    if cls == "compact_intseq":
        n, start, step = (float(x) for x in state.value)
        n = int(n)
        v = (int(start) + int(step) * np.arange(n, dtype=np.int64)).astype(np.int32)
        return RObject(v, sexptype=INTSXP)
    if cls == "compact_realseq":
        n, start, step = (float(x) for x in state.value)
        return RObject(start + step * np.arange(int(n), dtype=np.float64), sexptype=REALSXP)
    if cls.startswith("wrap_"):
        # state = list(x, meta); the wrapped vector may carry the attributes itself
        inner = state.value[0]
        if isinstance(inner, tuple):
            inner = inner[1]
        return RObject(inner.value, dict(inner.attrs), inner.sexptype)
    if cls == "deferred_string":
        # state = pairlist (arg, scipen): as.character() of a numeric/integer vector
        arg = state.value[0][1]
        vals = arg.value
        if arg.sexptype == INTSXP:
            out = [None if int(x) == NA_INTEGER else str(int(x)) for x in vals]
        else:
            out = [_format_g15(float(x)) for x in vals]
        return RObject(out, dict(arg.attrs), STRSXP)
    if cls == "mmap_real" or cls == "mmap_integer":
        raise RDataError("mmap ALTREP objects cannot be read")
    raise RDataError(f"unsupported ALTREP class {cls!r}")


def _decompress(raw: bytes) -> bytes:
    if raw[:2] == b"\x1f\x8b":
        return gzip.decompress(raw)
    if raw[:3] == b"BZh":
        return bz2.decompress(raw)
    if raw[:6] == b"\xfd7zXZ\x00":
        return lzma.decompress(raw)
    return raw


def _read_header(r: _Reader) -> None:
    fmt = r.raw(2)
    if fmt != b"X\n":
        raise RDataError(f"only XDR serialisation is supported (got {fmt!r})")
    version = r.i32()
    r.i32()  # writer R version
    r.i32()  # minimal reader version
    if version == 3:
        n = r.i32()
        r.raw(n)  # native encoding name
    elif version != 2:
        raise RDataError(f"unsupported serialisation version {version}")


def load_rdata(path: str) -> dict:
    """``load(path)``: returns {object name: RObject}."""
    with open(path, "rb") as fh:
        buf = _decompress(fh.read())
    if buf[:5] not in (b"RDX3\n", b"RDX2\n"):
        raise RDataError(f"{path}: not an RDX2/RDX3 workspace (magic {buf[:5]!r})")
    r = _Reader(buf, 5)
    _read_header(r)
    top = r.item()
    if top is None:
        return {}
    return {k: v for k, v in top.value}


def read_rds(path: str) -> RObject:
    """``readRDS(path)``."""
    with open(path, "rb") as fh:
        buf = _decompress(fh.read())
    r = _Reader(buf, 0)
    _read_header(r)
    return r.item()
