"""Field table shared by the CUDA library and the oracle (include/mpasb_fields.def)."""
from __future__ import annotations

import os
import re
from dataclasses import dataclass

_DEF = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "include", "mpasb_fields.def")


@dataclass(frozen=True)
class FieldDef:
    name: str
    loc: str        # CELL | EDGE | VERTEX | LEVS
    inner: str
    levels: int
    type: str       # REAL | INT
    target: str     # CELL | EDGE | VERTEX | LOCAL | NONE


def _parse():
    out = {}
    pat = re.compile(r"^F\(\s*(\w+)\s*,\s*(\w+)\s*,\s*(\w+)\s*,\s*(\d)\s*,\s*(\w+)\s*,\s*(\w+)\s*\)")
    with open(_DEF) as f:
        for line in f:
            m = pat.match(line.strip())
            if m:
                n, loc, inner, lev, typ, tgt = m.groups()
                out[n] = FieldDef(n, loc, inner, int(lev), typ, tgt)
    return out


FIELDS = _parse()


def host_shape(fd: FieldDef, dims) -> tuple:
    """C-order numpy shape of the dense host array (== Fortran shape reversed)."""
    nl = dims.nVertLevels
    inner = {
        "ONE": (), "NL": (nl,), "NL1": (nl + 1,), "ME": (dims.maxEdges,), "ME2": (dims.maxEdges2,),
        "VD": (dims.vertexDegree,), "TWO": (2,), "F15": (15,), "NL1_ME": (dims.maxEdges, nl + 1),
        "S_NL": (nl, dims.num_scalars), "NL_TWO": (2, nl), "THREE_ME": (dims.maxEdges, 3),
    }[fd.inner]
    outer = {"CELL": (dims.nCells + 1,), "EDGE": (dims.nEdges + 1,), "VERTEX": (dims.nVertices + 1,), "LEVS": ()}[fd.loc]
    return outer + inner
