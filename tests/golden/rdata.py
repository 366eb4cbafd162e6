"""Minimal pure-Python reader for R's XDR serialisation (RDX2 / RDX3 .RData / .rda files).

Test infrastructure only: used by ``make_fixtures.py`` to turn the reference's bundled
datasets (``/root/reference/data/*.RData``; described in ``/root/reference/R/data.R``)
into small ``.npz`` fixtures that can travel to the GPU box.  Supports exactly the SEXP
types those three files contain.
"""
import bz2
import gzip
import lzma
import struct

import numpy as np

NILVALUE, GLOBALENV, REFSXP, NAMESPACESXP, ALTREP = 254, 253, 255, 249, 238
SYMSXP, LISTSXP, CHARSXP, LGLSXP, INTSXP, REALSXP, STRSXP, VECSXP, S4SXP = 1, 2, 9, 10, 13, 14, 16, 19, 25
NA_INT = -2147483648


class RObj:
    """A parsed R value: ``value`` plus its ``attr`` dict (names, dim, class, levels ...)."""

    def __init__(self, value, attr=None):
        self.value = value
        self.attr = attr or {}

    def __repr__(self):
        return f"RObj({type(self.value).__name__}, attr={list(self.attr)})"


class _Reader:
    def __init__(self, buf):
        self.b = buf
        self.p = 0
        self.refs = []

    def i32(self):
        v = struct.unpack_from(">i", self.b, self.p)[0]
        self.p += 4
        return v

    def length(self):
        n = self.i32()
        if n == -1:
            hi, lo = self.i32(), self.i32()
            n = (hi << 32) + lo
        return n

    def raw(self, n):
        v = self.b[self.p:self.p + n]
        self.p += n
        return v

    def item(self):
        flags = self.i32()
        ty = flags & 0xFF
        has_attr = bool(flags & (1 << 9))
        has_tag = bool(flags & (1 << 10))
        if ty == NILVALUE:
            return None
        if ty == GLOBALENV:
            return "<globalenv>"
        if ty == REFSXP:
            idx = flags >> 8
            if idx == 0:
                idx = self.i32()
            return self.refs[idx - 1]
        if ty == NAMESPACESXP:
            self.i32()
            n = self.i32()
            info = [self.item() for _ in range(n)]
            self.refs.append(("<namespace>", info))
            return self.refs[-1]
        if ty == SYMSXP:
            name = self.item()
            self.refs.append(name)
            return name
        if ty == CHARSXP:
            n = self.i32()
            return None if n == -1 else self.raw(n).decode("utf-8", "replace")
        if ty == LISTSXP:
            # pairlist: iterate instead of recursing on the tail
            out = []
            while True:
                attr = self.item() if has_attr else None
                tag = self.item() if has_tag else None
                car = self.item()
                out.append((tag, car))
                flags = self.i32()
                ty2 = flags & 0xFF
                if ty2 == NILVALUE:
                    break
                if ty2 != LISTSXP:
                    raise ValueError(f"unexpected pairlist tail type {ty2}")
                has_attr = bool(flags & (1 << 9))
                has_tag = bool(flags & (1 << 10))
            return out
        if ty == ALTREP:
            info = self.item()
            state = self.item()
            self.item()  # attributes
            cls = info[0][1] if isinstance(info, list) else info
            if cls == "compact_intseq":
                n, start, step = (int(x) for x in state.value)
                return RObj(np.arange(start, start + n * step, step, dtype=np.int32))
            if cls in ("wrap_real", "wrap_integer", "wrap_string", "wrap_logical"):
                return state.value[0] if isinstance(state, RObj) and isinstance(state.value, list) else state
            raise ValueError(f"unsupported ALTREP class {cls}")
        if ty in (LGLSXP, INTSXP):
            n = self.length()
            v = np.frombuffer(self.raw(4 * n), dtype=">i4").astype(np.int32)
        elif ty == REALSXP:
            n = self.length()
            v = np.frombuffer(self.raw(8 * n), dtype=">f8").astype(np.float64)
        elif ty == STRSXP:
            n = self.length()
            v = [self.item() for _ in range(n)]
        elif ty == VECSXP:
            n = self.length()
            v = [self.item() for _ in range(n)]
        elif ty == S4SXP:
            v = "<S4>"
        elif ty == 22:  # EXTPTRSXP (e.g. data.table's .internal.selfref): protected value + tag
            self.refs.append("<extptr>")
            self.item()
            self.item()
            v = "<extptr>"
        else:
            raise ValueError(f"unsupported SEXP type {ty} at byte {self.p}")
        attr = {}
        if has_attr:
            for tag, val in self.item():
                attr[tag] = val
        return RObj(v, attr)


def _decompress(path):
    raw = open(path, "rb").read()
    if raw[:2] == b"\x1f\x8b":
        return gzip.decompress(raw)
    if raw[:3] == b"BZh":
        return bz2.decompress(raw)
    if raw[:6] == b"\xfd7zXZ\x00":
        return lzma.decompress(raw)
    return raw


def read_rdata(path):
    """Return ``{name: RObj}`` for the top-level objects saved in an .RData/.rda file."""
    buf = _decompress(path)
    if buf[:5] not in (b"RDX2\n", b"RDX3\n"):
        raise ValueError("not an RDX2/RDX3 file")
    r = _Reader(buf)
    r.p = 5
    if r.raw(2) != b"X\n":
        raise ValueError("only XDR serialisation is supported")
    version = r.i32()
    r.i32()
    r.i32()
    if version == 3:
        r.raw(r.i32())  # native encoding
    top = r.item()
    return {tag: val for tag, val in top}


def r_list(obj):
    """Named R list -> dict."""
    names = obj.attr.get("names")
    return dict(zip(names.value, obj.value))


def r_matrix(obj):
    """R matrix (column-major) -> numpy array [nrow, ncol]."""
    nr, nc = (int(x) for x in obj.attr["dim"].value)
    return np.asarray(obj.value).reshape((nc, nr)).T.copy()


def r_factor_or_strings(obj):
    """Return (codes int32 0-based, levels list) following R's ``as.factor`` (sorted unique)."""
    if "levels" in obj.attr:
        return np.asarray(obj.value, dtype=np.int32) - 1, list(obj.attr["levels"].value)
    vals = list(obj.value)
    levels = sorted(set(vals))
    lut = {v: i for i, v in enumerate(levels)}
    return np.array([lut[v] for v in vals], dtype=np.int32), levels
