"""MPAS-Atmosphere init files (``x1.<nCells>.init.nc``) <-> the block dict the dycore consumes.

The variable set is the reference's ``invariant`` + ``input`` streams (src/core_atmosphere/Registry.xml:450-577)
restricted to what the dry/moist dycore reads: mesh geometry and connectivity, the vertical grid, terrain metrics
``zgrid, zz, zxu, zb, zb3``, the advection/deformation coefficients ``deriv_two, defc_a, defc_b``, the reconstruction
vectors, and the uncoupled state ``u, w, rho, theta, rho_base, theta_base, surface_pressure`` plus one variable per
scalar (``qv`` ...).  On disk every array is in the file's C order (the reverse of the Fortran declaration, e.g.
``zb(nVertLevelsP1, TWO, nEdges)`` is stored as ``[nEdges][TWO][nVertLevelsP1]``), indices are 1-based and there is no
garbage element; in memory (``mesh.py`` conventions) indices are 0-based and every per-cell/edge/vertex array carries
the trailing garbage row.  ``read_init_file`` ends with ``init_block`` (atm_mpas_init_block, mpas_atm_core.F:368-602),
so the returned dict can be handed to ``Dycore`` / ``OracleDycore`` directly.
"""
from __future__ import annotations

import numpy as np

from . import ncio
from .init_block import default_config, init_block
from .mesh import _GARBAGE_OF

# name -> dimension names of the stored array (without Time); per-element arrays lose/gain the garbage row
_MESH = {
    "latCell": ("nCells",), "lonCell": ("nCells",), "xCell": ("nCells",), "yCell": ("nCells",), "zCell": ("nCells",),
    "indexToCellID": ("nCells",),
    "latEdge": ("nEdges",), "lonEdge": ("nEdges",), "xEdge": ("nEdges",), "yEdge": ("nEdges",), "zEdge": ("nEdges",),
    "indexToEdgeID": ("nEdges",),
    "latVertex": ("nVertices",), "lonVertex": ("nVertices",), "xVertex": ("nVertices",), "yVertex": ("nVertices",),
    "zVertex": ("nVertices",), "indexToVertexID": ("nVertices",),
    "cellsOnEdge": ("nEdges", "TWO"), "nEdgesOnCell": ("nCells",), "nEdgesOnEdge": ("nEdges",),
    "edgesOnCell": ("nCells", "maxEdges"), "edgesOnEdge": ("nEdges", "maxEdges2"), "weightsOnEdge": ("nEdges", "maxEdges2"),
    "dvEdge": ("nEdges",), "dcEdge": ("nEdges",), "angleEdge": ("nEdges",), "areaCell": ("nCells",), "areaTriangle": ("nVertices",),
    "cellsOnCell": ("nCells", "maxEdges"), "verticesOnCell": ("nCells", "maxEdges"), "verticesOnEdge": ("nEdges", "TWO"),
    "edgesOnVertex": ("nVertices", "vertexDegree"), "cellsOnVertex": ("nVertices", "vertexDegree"),
    "kiteAreasOnVertex": ("nVertices", "vertexDegree"), "fEdge": ("nEdges",), "fVertex": ("nVertices",), "meshDensity": ("nCells",),
    "zgrid": ("nCells", "nVertLevelsP1"), "zz": ("nCells", "nVertLevels"), "zxu": ("nEdges", "nVertLevels"),
    "zb": ("nEdges", "TWO", "nVertLevelsP1"), "zb3": ("nEdges", "TWO", "nVertLevelsP1"),
    "deriv_two": ("nEdges", "TWO", "FIFTEEN"), "defc_a": ("nCells", "maxEdges"), "defc_b": ("nCells", "maxEdges"),
    "rdzw": ("nVertLevels",), "dzu": ("nVertLevels",), "rdzu": ("nVertLevels",), "fzm": ("nVertLevels",), "fzp": ("nVertLevels",),
    "u_init": ("nVertLevels",), "v_init": ("nVertLevels",), "qv_init": ("nVertLevels",), "t_init": ("nCells", "nVertLevels"),
    "hx": ("nCells",),
}
_STATE = {       # stored with a leading Time dimension of length 1
    "u": ("nEdges", "nVertLevels"), "w": ("nCells", "nVertLevelsP1"), "rho": ("nCells", "nVertLevels"),
    "theta": ("nCells", "nVertLevels"), "rho_base": ("nCells", "nVertLevels"), "theta_base": ("nCells", "nVertLevels"),
    "surface_pressure": ("nCells",),
}
_SCALARS = ("nCells", "nVertLevels")
_SCALAR_NAMES = ("qv", "qc", "qr", "qi", "qs", "qg")          # Registry.xml var_array "scalars", moist species first
_COUNTS = {"nCells": "nCells", "nEdges": "nEdges", "nVertices": "nVertices"}
_ONES = ("areaCell", "areaTriangle", "dcEdge", "dvEdge", "meshDensity")      # garbage-row value 1 (mesh.add_garbage_slots)


def scalar_names(num_scalars):
    return [(_SCALAR_NAMES[s] if s < len(_SCALAR_NAMES) else f"tracer{s - len(_SCALAR_NAMES) + 1}") for s in range(num_scalars)]


def write_init_file(d: dict, path: str, version: int = 5) -> None:
    """Write the case dict as an MPAS init file (CDF-5 by default, as PnetCDF's ``cdf5`` io_type does)."""
    nC, nE, nV, nz = d["nCells"], d["nEdges"], d["nVertices"], d["nVertLevels"]
    dims = {"Time": 0, "nCells": nC, "nEdges": nE, "nVertices": nV, "maxEdges": d["maxEdges"], "maxEdges2": d["maxEdges2"],
            "TWO": 2, "THREE": 3, "vertexDegree": d["vertexDegree"], "FIFTEEN": 15, "R3": 3,
            "nVertLevels": nz, "nVertLevelsP1": nz + 1, "StrLen": 64}
    out = {}
    for name, dn in _MESH.items():
        if name not in d:
            continue
        a = np.asarray(d[name])
        if dn[0] in _COUNTS:
            a = a[: d[_COUNTS[dn[0]]]]
        if name in _GARBAGE_OF:
            a = (a + 1).astype(np.int32)                     # 1-based; the garbage index n becomes n + 1 as in the reference
        elif a.dtype.kind == "i":
            a = a.astype(np.int32)
        out[name] = ncio.Var(dn, np.ascontiguousarray(a))
    for name in ("nominalMinDc", "cf1", "cf2", "cf3"):
        out[name] = ncio.Var((), np.asarray(float(d[name])))
    for name, dn in _STATE.items():
        out[name] = ncio.Var(("Time",) + dn, np.ascontiguousarray(np.asarray(d[name])[: d[_COUNTS[dn[0]]]][None]))
    for s, nm in enumerate(scalar_names(d["num_scalars"])):
        out[nm] = ncio.Var(("Time",) + _SCALARS, np.ascontiguousarray(d["scalars"][:nC, :, s][None]))
    xt = np.zeros((1, 64), dtype="S1"); stamp = b"0000-01-01_00:00:00"
    xt[0, : len(stamp)] = np.frombuffer(stamp, dtype="S1")
    out["xtime"] = ncio.Var(("Time", "StrLen"), xt)
    attrs = {"on_a_sphere": "YES" if d.get("on_a_sphere", True) else "NO", "sphere_radius": float(d["sphere_radius"]),
             "is_periodic": "NO", "model_name": "mpas", "core_name": "init_atmosphere", "mesh_spec": "1.0",
             "Conventions": "MPAS", "source": "mpas_model_b200.initfile"}
    ncio.write(path, dims, attrs, out, version=version, unlimited="Time")


def read_init_file(path: str, dt: float | None = None, time_index: int = 0, derive: str = "numpy", **cfg_overrides):
    """-> (block dict, cfg) ready for ``Dycore``/``OracleDycore``.  Scalars found in the file (``qv, qc, ...`` then
    ``tracerN``) become the planes of ``scalars`` in that order; ``qv`` is the only moist species assumed.
    ``derive``: who runs the mesh part of atm_mpas_init_block -- "numpy" (init_block.py, here) or "library" (nothing is derived
    here; the caller hands the raw fields to ``Dycore.atm_mpas_init_block``, i.e. mpasb_init_block)."""
    dims, attrs, v = ncio.read(path)
    nC, nE, nV, nz = dims["nCells"], dims["nEdges"], dims["nVertices"], dims["nVertLevels"]
    d = dict(nCells=nC, nEdges=nE, nVertices=nV, maxEdges=dims["maxEdges"], maxEdges2=dims["maxEdges2"],
             vertexDegree=dims["vertexDegree"], nVertLevels=nz, sphere_radius=float(attrs.get("sphere_radius", 6371229.0)),
             on_a_sphere=str(attrs.get("on_a_sphere", "YES")).strip().upper().startswith("Y"))
    count = {"nCells": nC, "nEdges": nE, "nVertices": nV}

    def with_garbage(name, a, lead):
        if lead not in count:
            return a
        if name in _GARBAGE_OF:
            a = a.astype(np.int32) - 1                      # 0-based; n + 1 -> n (the garbage slot)
            fill = count[_GARBAGE_OF[name]]
        else:
            fill = 1 if name in _ONES else 0
        return np.concatenate([a, np.full((1,) + a.shape[1:], fill, dtype=a.dtype)])

    for name, dn in _MESH.items():
        if name in v:
            d[name] = with_garbage(name, np.array(v[name].data), dn[0])
    for name in ("nominalMinDc", "cf1", "cf2", "cf3"):
        d[name] = float(np.asarray(v[name].data).reshape(-1)[0])
    for name, dn in _STATE.items():
        d[name] = with_garbage(name, np.array(v[name].data[time_index]), dn[0])
    names = [n for n in _SCALAR_NAMES if n in v]
    t = 1
    while f"tracer{t}" in v:
        names.append(f"tracer{t}"); t += 1
    if "qv" not in names:
        raise ValueError("init file has no qv")
    sc = np.zeros((nC + 1, nz, len(names)))
    for s, nm in enumerate(names):
        sc[:nC, :, s] = v[nm].data[time_index]
    d.update(scalars=sc, num_scalars=len(names), index_qv=names.index("qv"), moist_start=names.index("qv"), moist_end=names.index("qv"))
    if dt is None:
        dt = 6.0 * round(d["nominalMinDc"] / 1000.0)
    cfg = default_config(d["nominalMinDc"], dt)
    cfg.update(cfg_overrides)
    if derive == "library":
        for n in ("specZoneMaskCell", "specZoneMaskEdge"):           # a global mesh: no lateral boundary zones
            d[n] = np.zeros((nC if n.endswith("Cell") else nE) + 1)
        for n in ("bdyMaskCell", "bdyMaskEdge"):
            d[n] = np.zeros((nC if n.endswith("Cell") else nE) + 1, dtype=np.int32)
    else:
        init_block(d, cfg)
    return d, cfg
