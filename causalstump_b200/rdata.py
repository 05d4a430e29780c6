"""Reader for R workspace files (`save()` format "RDX2", XDR serialisation, version 2) -- SURVEY.md 8(f)-4.

The reference keeps its simulation fixtures as `*.RData` (the real IHDP covariates `X`, the treatment
indicator `Trt` and the stored 100-replicate `results` matrices of test/sec5-3_Hill2011.R, e.g.
/root/reference/test/H11_100_27_august.RData).  R is not available where this package is built, so this
module decodes the subset of R's serialisation format those files use, in pure Python / NumPy:

  gzip | "RDX2\\n" | "X\\n" | int32 version, writer R version, min reader version | one pairlist of
  (symbol -> value) bindings.

Supported SEXP types: NULL, symbols, pairlists (with tags / attributes), language objects, closures,
environments (skipped structurally), character / logical / integer / real / complex / raw vectors,
generic vectors (lists, data.frames), expression vectors, byte code is NOT supported (the fixtures hold
data only).  Attributes (names, dim, dimnames, class, levels, row.names) are returned next to the values.

    ws = read_rdata(path)            # dict name -> RObject | ndarray
    X  = as_matrix(ws["X"])          # data.frame / matrix -> 2-D float64 array, column names
"""
from __future__ import annotations

import gzip
import struct

import numpy as np

NILSXP, SYMSXP, LISTSXP, CLOSXP, ENVSXP, PROMSXP, LANGSXP = 0, 1, 2, 3, 4, 5, 6
CHARSXP, LGLSXP, INTSXP, REALSXP, CPLXSXP, STRSXP, DOTSXP, VECSXP, EXPRSXP, RAWSXP = 9, 10, 13, 14, 15, 16, 17, 19, 20, 24
REFSXP, NILVALUE_SXP, GLOBALENV_SXP, UNBOUNDVALUE_SXP, MISSINGARG_SXP, BASENAMESPACE_SXP = 255, 254, 253, 252, 251, 250
NAMESPACESXP, PACKAGESXP, PERSISTSXP, EMPTYENV_SXP, BASEENV_SXP, ATTRLANGSXP, ATTRLISTSXP = 249, 248, 247, 246, 245, 244, 243
NA_INTEGER = -2147483648


class RObject:
    """A decoded R value with attributes: `.value` (ndarray, list or dict) and `.attr` (dict)."""

    def __init__(self, value, attr=None, kind=""):
        self.value, self.attr, self.kind = value, attr or {}, kind

    def __repr__(self):
        return "RObject(%s, attr=%s)" % (self.kind, sorted(self.attr))


class _Reader:
    def __init__(self, buf):
        self.b, self.o, self.refs = buf, 0, []

    def i32(self):
        v = struct.unpack_from(">i", self.b, self.o)[0]
        self.o += 4
        return v

    def length(self):
        n = self.i32()
        if n == -1:  # long vector: two int32 words
            hi, lo = self.i32(), self.i32()
            n = (hi << 32) + (lo & 0xFFFFFFFF)
        return n

    def take(self, dtype, n):
        a = np.frombuffer(self.b, dtype=dtype, count=n, offset=self.o)
        self.o += a.nbytes
        return a

    def attributes(self):
        a = self.item()
        return a if isinstance(a, dict) else {}

    def item(self):
        flags = self.i32()
        t = flags & 0xFF
        has_attr, has_tag = bool(flags & 0x200), bool(flags & 0x400)
        if t in (NILVALUE_SXP, NILSXP):
            return None
        if t in (GLOBALENV_SXP, EMPTYENV_SXP, BASEENV_SXP, BASENAMESPACE_SXP, UNBOUNDVALUE_SXP, MISSINGARG_SXP):
            return RObject(None, kind="env%d" % t)
        if t == REFSXP:
            idx = flags >> 8
            if idx == 0:
                idx = self.i32()
            return self.refs[idx - 1]
        if t == SYMSXP:
            name = self.item()
            self.refs.append(name)
            return name
        if t in (NAMESPACESXP, PACKAGESXP, PERSISTSXP):
            self.i32()  # always 0
            n = self.i32()
            v = RObject([self.item() for _ in range(n)], kind="namespace")
            self.refs.append(v)
            return v
        if t == ENVSXP:
            self.i32()  # locked
            v = RObject(None, kind="environment")
            self.refs.append(v)
            for _ in range(4):  # enclosure, frame, hash table, attributes
                self.item()
            return v
        if t in (LISTSXP, LANGSXP, CLOSXP, PROMSXP, DOTSXP, ATTRLANGSXP, ATTRLISTSXP):
            # one cons cell: [attributes] [tag] car cdr.  Chains of LISTSXP cells are walked iteratively; a
            # pairlist whose cells all carry tags (attribute lists, the top-level bindings) becomes a dict.
            cells = []
            while True:
                if has_attr or t in (ATTRLANGSXP, ATTRLISTSXP):
                    self.attributes()
                tag = self.item() if has_tag else None
                cells.append((tag, self.item()))
                if t in (CLOSXP, PROMSXP):
                    cells.append((None, self.item()))  # body / expression
                    return RObject([c[1] for c in cells], kind="closure" if t == CLOSXP else "promise")
                nflags = struct.unpack_from(">i", self.b, self.o)[0]
                nt = nflags & 0xFF
                if nt not in (LISTSXP, ATTRLISTSXP, LANGSXP, ATTRLANGSXP, DOTSXP):
                    tail = self.item()  # NILVALUE_SXP ends the chain
                    if tail is not None:
                        cells.append((None, tail))
                    break
                self.o += 4
                flags, t = nflags, nt
                has_attr, has_tag = bool(flags & 0x200), bool(flags & 0x400)
            if cells and all(k is not None for k, _ in cells):
                return dict(cells)
            return [v for _, v in cells]
        if t == 22:  # EXTPTRSXP: reference-tracked, protected value + tag, then attributes
            v = RObject(None, kind="externalptr")
            self.refs.append(v)
            self.item(); self.item()
            if has_attr:
                v.attr = self.attributes()
            return v
        if t == 23:  # WEAKREFSXP
            v = RObject(None, kind="weakref")
            self.refs.append(v)
            if has_attr:
                v.attr = self.attributes()
            return v
        if t == 25:  # S4SXP: attributes only
            return RObject(None, self.attributes() if has_attr else {}, "S4")
        if t in (7, 8):  # SPECIALSXP / BUILTINSXP: name of the primitive
            n = self.i32()
            s = bytes(self.b[self.o:self.o + n]).decode()
            self.o += n
            return RObject(s, kind="builtin")
        if t == CHARSXP:
            n = self.i32()
            if n == -1:
                return None
            s = bytes(self.b[self.o:self.o + n])
            self.o += n
            return s.decode("utf-8", "replace")
        if t in (LGLSXP, INTSXP):
            v = self.take(">i4", self.length()).astype(np.int32)
        elif t == REALSXP:
            v = self.take(">f8", self.length()).astype(np.float64)
        elif t == CPLXSXP:
            v = self.take(">c16", self.length()).astype(np.complex128)
        elif t == RAWSXP:
            v = self.take("u1", self.length()).copy()
        elif t == STRSXP:
            v = [self.item() for _ in range(self.length())]
        elif t in (VECSXP, EXPRSXP):
            v = [self.item() for _ in range(self.length())]
        else:
            raise NotImplementedError("R SEXP type %d at offset %d is not supported" % (t, self.o - 4))
        attr = self.attributes() if has_attr else {}
        if not attr and t in (LGLSXP, INTSXP, REALSXP, CPLXSXP, RAWSXP):
            return v
        kind = {LGLSXP: "logical", INTSXP: "integer", REALSXP: "numeric", STRSXP: "character", VECSXP: "list"}.get(t, str(t))
        return RObject(v, attr, kind)


def read_rdata(path_or_bytes):
    """-> dict name -> value of every object saved in the workspace file."""
    raw = path_or_bytes if isinstance(path_or_bytes, (bytes, bytearray)) else open(path_or_bytes, "rb").read()
    if raw[:2] == b"\x1f\x8b":
        raw = gzip.decompress(raw)
    if raw[:5] != b"RDX2\n":
        raise ValueError("not an RDX2 workspace file (magic %r)" % raw[:5])
    if raw[5:7] != b"X\n":
        raise ValueError("only the XDR (binary big-endian) serialisation is supported")
    rd = _Reader(memoryview(raw))
    rd.o = 7
    version, _writer, _minreader = rd.i32(), rd.i32(), rd.i32()
    if version != 2:
        raise ValueError("serialisation version %d is not supported (expected 2)" % version)
    top = rd.item()
    if not isinstance(top, dict):
        raise ValueError("workspace does not start with a tagged pairlist")
    return top


def _strings(x):
    if x is None:
        return None
    return list(x.value) if isinstance(x, RObject) else list(x)


def as_matrix(obj):
    """data.frame or matrix -> (2-D float64 array, column names or None).  Factors become their integer codes."""
    if isinstance(obj, np.ndarray):
        return obj.astype(np.float64).reshape(-1, 1), None
    cls = _strings(obj.attr.get("class")) or []
    if "data.frame" in cls:
        cols = []
        for c in obj.value:
            v = c.value if isinstance(c, RObject) else c
            v = np.asarray(v)
            col = v.astype(np.float64)
            if v.dtype == np.int32:
                col[v == NA_INTEGER] = np.nan
            cols.append(col)
        return np.column_stack(cols), _strings(obj.attr.get("names"))
    dim = obj.attr.get("dim")
    if dim is None:
        return np.asarray(obj.value, dtype=np.float64).reshape(-1, 1), None
    dim = [int(d) for d in (dim.value if isinstance(dim, RObject) else dim)]
    a = np.asarray(obj.value, dtype=np.float64).reshape(dim, order="F")
    dn = obj.attr.get("dimnames")
    names = None
    if dn is not None:
        dnl = dn.value if isinstance(dn, RObject) else dn
        if len(dnl) > 1 and dnl[1] is not None:
            names = _strings(dnl[1])
    return a, names
