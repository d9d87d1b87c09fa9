"""Minimal reader/writer of the NetCDF classic formats (CDF-1, CDF-2 "64-bit offset", CDF-5 "64-bit data").

MPAS init and restart files (``x1.*.init.nc``, ``restart.*.nc``) are written by PIO/PnetCDF as CDF-2 or CDF-5
(``io_type="pnetcdf,cdf5"``, src/framework/mpas_io.F); no NetCDF library exists in this image and scipy reads
CDF-1/2 only, so the classic file layout (header: dimensions, global attributes, variables; fixed-size data,
then interleaved records) is restated here from the format specification.  Row f3 of SURVEY.md §8: it lets a
host-less run start from an MPAS file; it is not on the timed path.

    dims, attrs, variables = ncio.read(path)          # variables[name] = Var(dims=(..), data=ndarray, attrs={})
    ncio.write(path, dims, attrs, variables, version=5, unlimited="Time")
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field

import numpy as np

NC_DIMENSION, NC_VARIABLE, NC_ATTRIBUTE = 0x0A, 0x0B, 0x0C
# nc_type -> big-endian numpy dtype
_TYPES = {1: ">i1", 2: "S1", 3: ">i2", 4: ">i4", 5: ">f4", 6: ">f8", 7: ">u1", 8: ">u2", 9: ">u4", 10: ">i8", 11: ">u8"}
_CODES = {"int8": 1, "bytes8": 2, "int16": 3, "int32": 4, "float32": 5, "float64": 6,
          "uint8": 7, "uint16": 8, "uint32": 9, "int64": 10, "uint64": 11}


@dataclass
class Var:
    dims: tuple
    data: np.ndarray
    attrs: dict = field(default_factory=dict)


def _pad4(n):
    return (n + 3) // 4 * 4


class _Reader:
    def __init__(self, buf):
        self.b, self.o = buf, 0
        if bytes(buf[:3]) != b"CDF" or buf[3] not in (1, 2, 5):
            raise ValueError("not a NetCDF classic file (CDF-1/2/5)")
        self.version = buf[3]
        self.o = 4

    def u32(self):
        v = struct.unpack_from(">I", self.b, self.o)[0]; self.o += 4; return v

    def u64(self):
        v = struct.unpack_from(">Q", self.b, self.o)[0]; self.o += 8; return v

    def nonneg(self):                      # "NON_NEG": 64-bit in CDF-5, 32-bit otherwise
        return self.u64() if self.version == 5 else self.u32()

    def offset(self):
        return self.u32() if self.version == 1 else self.u64()

    def name(self):
        n = self.nonneg()
        s = bytes(self.b[self.o:self.o + n]).decode(); self.o += _pad4(n); return s

    def values(self, nc_type, n):
        dt = np.dtype(_TYPES[nc_type])
        a = np.frombuffer(self.b, dtype=dt, count=n, offset=self.o)
        self.o += _pad4(n * dt.itemsize)
        return a

    def attrs(self):
        tag = self.u32(); n = self.nonneg()
        if tag == 0 and n == 0:
            return {}
        if tag != NC_ATTRIBUTE:
            raise ValueError("bad attribute list")
        out = {}
        for _ in range(n):
            nm = self.name(); t = self.u32(); cnt = self.nonneg()
            v = self.values(t, cnt)
            out[nm] = b"".join(v.tolist()).decode(errors="replace") if t == 2 else (v.astype(v.dtype.newbyteorder("=")) if cnt != 1 else v[0].item())
        return out


def read(path, only=None):
    """-> (dims {name: length; the unlimited one holds numrecs}, global attrs, {name: Var}); ``only`` restricts the
    variables whose data are materialised.  Data are returned in native byte order, C-ordered as stored."""
    buf = np.memmap(path, dtype=np.uint8, mode="r")
    r = _Reader(buf)
    numrecs = r.nonneg()
    if numrecs == (0xFFFFFFFFFFFFFFFF if r.version == 5 else 0xFFFFFFFF):
        numrecs = None                                  # streaming: derived from the file size below
    tag = r.u32(); n = r.nonneg()
    dim_names, dim_len, unlimited = [], [], None
    if tag == NC_DIMENSION:
        for i in range(n):
            dim_names.append(r.name()); L = r.nonneg(); dim_len.append(L)
            if L == 0:
                unlimited = i
    gattrs = r.attrs()
    tag = r.u32(); n = r.nonneg()
    metas = []
    if tag == NC_VARIABLE:
        for _ in range(n):
            nm = r.name(); nd = r.nonneg()
            dimids = [r.nonneg() for _ in range(nd)]
            va = r.attrs(); t = r.u32(); vsize = r.nonneg(); begin = r.offset()
            metas.append((nm, dimids, va, t, vsize, begin))
    rec_vars = [m for m in metas if m[1] and m[1][0] == unlimited]
    # per-record size of each record variable from its dimensions and type, padded to 4 bytes -- NOT from the header's vsize,
    # which saturates at 2^32-1 in CDF-1/2 for variables of 4 GiB or more per record
    def _rec_bytes(m):
        nbytes = int(np.prod([dim_len[d] for d in m[1][1:]], dtype=np.int64)) * np.dtype(_TYPES[m[3]]).itemsize
        return nbytes if len(rec_vars) == 1 else (nbytes + 3) // 4 * 4          # a single record variable is not padded
    recsize = sum(_rec_bytes(m) for m in rec_vars)
    if numrecs is None:
        numrecs = (len(buf) - min(m[5] for m in rec_vars)) // recsize if rec_vars else 0
    out = {}
    for nm, dimids, va, t, vsize, begin in metas:
        if only is not None and nm not in only:
            continue
        dt = np.dtype(_TYPES[t])
        if dimids and dimids[0] == unlimited:
            shape = [dim_len[d] for d in dimids[1:]]
            cnt = int(np.prod(shape, dtype=np.int64))
            a = np.empty([numrecs] + shape, dtype=dt)
            for k in range(numrecs):
                a[k] = np.frombuffer(buf, dtype=dt, count=cnt, offset=begin + k * recsize).reshape(shape)
        else:
            shape = [dim_len[d] for d in dimids]
            cnt = int(np.prod(shape, dtype=np.int64))
            a = np.frombuffer(buf, dtype=dt, count=cnt, offset=begin).reshape(shape)
        a = a.astype(dt.newbyteorder("=")) if t != 2 else np.array(a)
        out[nm] = Var(tuple(dim_names[d] for d in dimids), a, va)
    dims = {nm: (numrecs if i == unlimited else L) for i, (nm, L) in enumerate(zip(dim_names, dim_len))}
    return dims, gattrs, out


class _Writer:
    def __init__(self, version):
        self.v, self.parts = version, []

    def u32(self, x): self.parts.append(struct.pack(">I", x))
    def u64(self, x): self.parts.append(struct.pack(">Q", x))
    def nonneg(self, x): self.u64(x) if self.v == 5 else self.u32(x)
    def offset(self, x): self.u32(x) if self.v == 1 else self.u64(x)

    def name(self, s):
        b = s.encode(); self.nonneg(len(b)); self.parts.append(b + b"\0" * (_pad4(len(b)) - len(b)))

    def attrs(self, d):
        if not d:
            self.u32(0); self.nonneg(0); return
        self.u32(NC_ATTRIBUTE); self.nonneg(len(d))
        for k, v in d.items():
            self.name(k)
            if isinstance(v, str):
                b = v.encode(); self.u32(2); self.nonneg(len(b)); self.parts.append(b + b"\0" * (_pad4(len(b)) - len(b)))
            else:
                a = np.atleast_1d(np.asarray(v))
                if a.dtype.kind == "i" and a.dtype.itemsize == 8 and self.v != 5:
                    a = a.astype(np.int32)
                code = _code_of(a.dtype)
                raw = a.astype(_TYPES[code]).tobytes()
                self.u32(code); self.nonneg(a.size); self.parts.append(raw + b"\0" * (_pad4(len(raw)) - len(raw)))

    def size(self): return sum(len(p) for p in self.parts)


def _code_of(dt):
    dt = np.dtype(dt)
    if dt.kind == "S":
        return 2
    key = dt.name
    if key not in _CODES:
        raise ValueError(f"unsupported dtype {dt}")
    return _CODES[key]


def write(path, dims, attrs, variables, version=5, unlimited=None):
    """``dims``: {name: length} in file order; the ``unlimited`` one (if any) takes its length from the record
    variables.  ``variables``: {name: Var}; a variable is a record variable iff its first dimension is ``unlimited``."""
    if version not in (1, 2, 5):
        raise ValueError("version must be 1, 2 or 5")
    names = list(dims)
    ids = {n: i for i, n in enumerate(names)}
    numrecs = 0
    for v in variables.values():
        if v.dims and v.dims[0] == unlimited:
            numrecs = max(numrecs, v.data.shape[0])
    metas = []
    for nm, v in variables.items():
        code = _code_of(v.data.dtype)
        if code > 6 and version != 5:
            raise ValueError(f"{nm}: dtype {v.data.dtype} needs CDF-5")
        isz = np.dtype(_TYPES[code]).itemsize
        rec = bool(v.dims) and v.dims[0] == unlimited
        shape = v.data.shape[1:] if rec else v.data.shape
        want = tuple(dims[d] for d in (v.dims[1:] if rec else v.dims))
        if tuple(shape) != want:
            raise ValueError(f"{nm}: shape {v.data.shape} does not match dims {v.dims} = {want}")
        vsize = _pad4(int(np.prod(shape, dtype=np.int64)) * isz)
        metas.append([nm, v, code, rec, vsize, 0])
    rec_metas = [m for m in metas if m[3]]
    if len(rec_metas) == 1:
        m = rec_metas[0]
        recsize = int(np.prod(m[1].data.shape[1:], dtype=np.int64)) * np.dtype(_TYPES[m[2]]).itemsize
    else:
        recsize = sum(m[4] for m in rec_metas)

    def header(with_offsets):
        w = _Writer(version)
        w.parts.append(b"CDF" + bytes([version]))
        w.nonneg(numrecs)
        if names:
            w.u32(NC_DIMENSION); w.nonneg(len(names))
            for n in names:
                w.name(n); w.nonneg(0 if n == unlimited else dims[n])
        else:
            w.u32(0); w.nonneg(0)
        w.attrs(attrs)
        if metas:
            w.u32(NC_VARIABLE); w.nonneg(len(metas))
            for nm, v, code, rec, vsize, begin in metas:
                w.name(nm); w.nonneg(len(v.dims))
                for d in v.dims:
                    w.nonneg(ids[d])
                w.attrs(v.attrs); w.u32(code)
                w.nonneg(min(vsize, 0xFFFFFFFF) if version != 5 else vsize)
                w.offset(begin if with_offsets else 0)
        else:
            w.u32(0); w.nonneg(0)
        return w

    off = header(False).size()
    for m in metas:
        if not m[3]:
            m[5] = off; off += m[4]
    rec_begin = off
    for m in rec_metas:
        m[5] = off; off += m[4] if len(rec_metas) > 1 else recsize
    if version == 1 and off + max(0, numrecs - 1) * recsize > 0x7FFFFFFF:
        raise ValueError("file too large for CDF-1")
    with open(path, "wb") as f:
        f.write(b"".join(header(True).parts))
        for nm, v, code, rec, vsize, begin in metas:
            if not rec:
                raw = np.ascontiguousarray(v.data).astype(_TYPES[code]).tobytes()
                f.write(raw + b"\0" * (vsize - len(raw)))
        for k in range(numrecs):
            for nm, v, code, rec, vsize, begin in rec_metas:
                raw = np.ascontiguousarray(v.data[k]).astype(_TYPES[code]).tobytes()
                f.write(raw + b"\0" * ((vsize if len(rec_metas) > 1 else recsize) - len(raw)))
    return rec_begin
