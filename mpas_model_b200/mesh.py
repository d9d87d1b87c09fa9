"""Quasi-uniform icosahedral SCVT mesh generator in the MPAS mesh specification.

The reference never computes its horizontal mesh fields: ``cellsOnEdge``,
``weightsOnEdge``, ``kiteAreasOnVertex``, ``angleEdge`` ... are read from an
externally generated ``x1.<nCells>.grid.nc`` (init "input" stream,
src/core_init_atmosphere/Registry.xml) and no such file ships with it
(SURVEY.md Appendix C).  This module generates the ``x1.10*4^n+2`` family
(recursive bisection of the icosahedron + a fixed number of Lloyd iterations)
with every field following the conventions implied by the reference's usage
sites, which are cited next to each field below.

All index arrays are 0-based here; entry ``n`` (one past the last element) is
the reference's "garbage" slot ``n+1`` (src/framework/mpas_block_creator.F:1464-1531).
Arrays carry the garbage slot as a trailing row, i.e. a Fortran array
``(maxEdges, nCells+1)`` is a C-order numpy array ``[nCells+1, maxEdges]``.
Geometry is on the unit sphere, exactly like a grid file; ``jw_init`` scales it
by ``sphere_radius`` (src/core_init_atmosphere/mpas_init_atm_cases.F:554-568).
"""
from __future__ import annotations

import numpy as np

MAX_EDGES = 6          # icosahedral meshes: pentagons and hexagons only
MAX_EDGES2 = 2 * MAX_EDGES
VERTEX_DEGREE = 3

LEVEL_OF = {12: 0, 42: 1, 162: 2, 642: 3, 2562: 4, 10242: 5, 40962: 6,
            163842: 7, 655362: 8, 2621442: 9}


def _normalize(p):
    return p / np.linalg.norm(p, axis=-1, keepdims=True)


def _arc(a, b):
    """Great-circle distance on the unit sphere (chord form, as
    mpas_atm_advection.F:506-529 arc_length)."""
    c = np.linalg.norm(b - a, axis=-1)
    return 2.0 * np.arcsin(np.minimum(1.0, 0.5 * c))


def _tri_area(a, b, c):
    """Unsigned spherical triangle area (Van Oosterom & Strackee)."""
    num = np.abs(np.einsum("ij,ij->i", a, np.cross(b, c)))
    den = 1.0 + np.einsum("ij,ij->i", a, b) + np.einsum("ij,ij->i", b, c) \
        + np.einsum("ij,ij->i", c, a)
    return 2.0 * np.arctan2(num, den)


def _icosahedron():
    t = (1.0 + np.sqrt(5.0)) / 2.0
    v = np.array([(-1, t, 0), (1, t, 0), (-1, -t, 0), (1, -t, 0),
                  (0, -1, t), (0, 1, t), (0, -1, -t), (0, 1, -t),
                  (t, 0, -1), (t, 0, 1), (-t, 0, -1), (-t, 0, 1)], dtype=np.float64)
    f = np.array([(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11),
                  (1, 5, 9), (5, 11, 4), (11, 10, 2), (10, 7, 6), (7, 1, 8),
                  (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9),
                  (4, 9, 5), (2, 4, 11), (6, 2, 10), (8, 6, 7), (9, 8, 1)], dtype=np.int64)
    v = _normalize(v)
    # a small fixed rotation keeps generators off the poles and off lon = 0
    # (the reference special-cases zc == 1.0, mpas_atm_advection.F:145)
    ax, ay = 0.31, 0.17
    rx = np.array([[1, 0, 0], [0, np.cos(ax), -np.sin(ax)], [0, np.sin(ax), np.cos(ax)]])
    ry = np.array([[np.cos(ay), 0, np.sin(ay)], [0, 1, 0], [-np.sin(ay), 0, np.cos(ay)]])
    v = v @ (rx @ ry).T
    # make every face counter-clockwise seen from outside
    n = np.cross(v[f[:, 1]] - v[f[:, 0]], v[f[:, 2]] - v[f[:, 0]])
    flip = np.einsum("ij,ij->i", n, v[f[:, 0]]) < 0
    f[flip] = f[flip][:, [0, 2, 1]]
    return v, f


def _subdivide(p, tri):
    n = p.shape[0]
    nt = tri.shape[0]
    e = np.concatenate([tri[:, [0, 1]], tri[:, [1, 2]], tri[:, [2, 0]]])
    e = np.sort(e, axis=1)
    key = e[:, 0] * n + e[:, 1]
    uniq, inv = np.unique(key, return_inverse=True)
    a, b = uniq // n, uniq % n
    mid = _normalize(p[a] + p[b])
    m = n + inv
    m01, m12, m20 = m[:nt], m[nt:2 * nt], m[2 * nt:]
    v0, v1, v2 = tri[:, 0], tri[:, 1], tri[:, 2]
    new = np.concatenate([np.stack([v0, m01, m20], 1), np.stack([v1, m12, m01], 1),
                          np.stack([v2, m20, m12], 1), np.stack([m01, m12, m20], 1)])
    return np.concatenate([p, mid]), new


def _circumcenters(p, tri):
    a, b, c = p[tri[:, 0]], p[tri[:, 1]], p[tri[:, 2]]
    v = _normalize(np.cross(b - a, c - a))
    s = np.sign(np.einsum("ij,ij->i", v, a))
    return v * s[:, None]


def _lloyd(p, tri, iters):
    """Lloyd iterations with the (icosahedral) Delaunay topology held fixed:
    each generator moves to the centroid of its Voronoi cell, assembled
    order-free from the two planar sub-triangles of every kite."""
    for _ in range(iters):
        v = _circumcenters(p, tri)
        cen = np.zeros_like(p)
        for j in range(3):
            c = tri[:, j]
            pc = p[c]
            ma = _normalize(pc + p[tri[:, (j + 1) % 3]])
            mb = _normalize(pc + p[tri[:, (j + 2) % 3]])
            for q0, q1 in ((ma, v), (v, mb)):
                w = 0.5 * np.linalg.norm(np.cross(q0 - pc, q1 - pc), axis=1)
                np.add.at(cen, c, w[:, None] * (pc + q0 + q1) / 3.0)
        p = _normalize(cen)
    return p


def _morton3(p, bits=20):
    q = np.clip(((p + 1.0) * 0.5 * (2 ** bits - 1)).astype(np.uint64), 0, 2 ** bits - 1)
    code = np.zeros(p.shape[0], dtype=np.uint64)
    for b in range(bits):
        for d in range(3):
            code |= ((q[:, d] >> np.uint64(b)) & np.uint64(1)) << np.uint64(3 * b + d)
    return code


def _latlon(p):
    lat = np.arcsin(np.clip(p[:, 2], -1.0, 1.0))
    lon = np.arctan2(p[:, 1], p[:, 0])
    lon = np.where(lon < 0.0, lon + 2.0 * np.pi, lon)
    return lat, lon


def _pad_rows(a, fill):
    """Append the garbage-slot row."""
    pad = np.full((1,) + a.shape[1:], fill, dtype=a.dtype)
    return np.concatenate([a, pad])


def _delaunay_on_sphere(p):
    """Spherical Delaunay triangulation = facets of the convex hull, oriented counter-clockwise seen from outside."""
    from scipy.spatial import ConvexHull
    tri = ConvexHull(p).simplices.astype(np.int64)
    a, b, c = p[tri[:, 0]], p[tri[:, 1]], p[tri[:, 2]]
    flip = np.einsum("ij,ij->i", np.cross(b - a, c - a), a + b + c) < 0
    tri[flip] = tri[flip][:, [0, 2, 1]]
    return tri


def generate(n_cells: int = 2562, lloyd_iters: int = 12, reorder: str = "morton", jitter: float = 0.0, seed: int = 0) -> dict:
    """Build the ``x1.<n_cells>`` mesh.  Returns a dict of numpy arrays named as
    in the MPAS grid file, 0-based connectivity, garbage slot included.

    ``jitter`` > 0 (a fraction of the nominal cell spacing) displaces the generators at random and re-triangulates:
    an irregular Voronoi mesh with 5-, 6-, 7- (and occasionally 8-) sided cells like the reference's variable-resolution
    meshes, used by the tests to exercise the paths that a pentagon/hexagon-only mesh never takes (``maxEdges`` > 6)."""
    if n_cells not in LEVEL_OF:
        raise ValueError(f"n_cells must be 10*4^n+2, got {n_cells}")
    p, tri = _icosahedron()
    for _ in range(LEVEL_OF[n_cells]):
        p, tri = _subdivide(p, tri)
    p = _lloyd(p, tri, lloyd_iters)
    if jitter > 0.0:
        rng = np.random.default_rng(seed)
        spacing = np.sqrt(2.0 * (4.0 * np.pi / p.shape[0]) / np.sqrt(3.0))
        p = _normalize(p + jitter * spacing * rng.uniform(-1.0, 1.0, p.shape))
        tri = _delaunay_on_sphere(p)
    MAX_EDGES = max(6, int(np.bincount(tri.reshape(-1)).max()))      # shadows the module constant: widest cell of this mesh
    MAX_EDGES2 = 2 * MAX_EDGES

    nC = p.shape[0]
    if reorder == "morton":
        perm = np.argsort(_morton3(p), kind="stable")        # new -> old
        inv = np.empty(nC, dtype=np.int64)
        inv[perm] = np.arange(nC)
        p = p[perm]
        tri = inv[tri]
    elif reorder != "none":
        raise ValueError(reorder)
    # rotate each triangle so its smallest cell id is first (keeps CCW), then
    # number vertices by that id: vertex numbering follows cell numbering
    r = np.argmin(tri, axis=1)
    tri = np.stack([tri[np.arange(len(tri)), (r + j) % 3] for j in range(3)], 1)
    tri = tri[np.lexsort((tri[:, 2], tri[:, 1], tri[:, 0]))]
    nV = tri.shape[0]
    xv = _circumcenters(p, tri)

    # ---- edges: cellsOnEdge(1,e) < cellsOnEdge(2,e); u > 0 points cell1 -> cell2
    # (mpas_atm_core.F:1201-1217).  verticesOnEdge(1)->(2) is k x n: the triangle
    # that holds the directed pair (c1,c2) counter-clockwise lies LEFT of the
    # normal and is vertex 2 (mpas_atm_core.F:1187-1199).
    a = tri.reshape(-1)                                       # (t,j) flattened
    b = np.roll(tri, -1, axis=1).reshape(-1)
    lo, hi = np.minimum(a, b), np.maximum(a, b)
    uniq, e_of = np.unique(lo * nC + hi, return_inverse=True)
    nE = uniq.shape[0]
    assert nE == 3 * nC - 6 and nV == 2 * nC - 4
    cellsOnEdge = np.stack([uniq // nC, uniq % nC], 1).astype(np.int64)
    t_of = np.repeat(np.arange(nV), 3)
    verticesOnEdge = np.empty((nE, 2), dtype=np.int64)
    fwd = a < b
    verticesOnEdge[e_of[fwd], 1] = t_of[fwd]
    verticesOnEdge[e_of[~fwd], 0] = t_of[~fwd]
    edgesOnVertex = e_of.reshape(nV, 3)          # edge j joins cellsOnVertex j, j+1
    cellsOnVertex = tri.copy()

    xe = _normalize(p[cellsOnEdge[:, 0]] + p[cellsOnEdge[:, 1]])
    dcEdge = _arc(p[cellsOnEdge[:, 0]], p[cellsOnEdge[:, 1]])
    dvEdge = _arc(xv[verticesOnEdge[:, 0]], xv[verticesOnEdge[:, 1]])

    # ---- per-cell counter-clockwise rings (mpas_atm_advection.F:907-935 needs
    # edge i to join verticesOnCell(i), verticesOnCell(i+1), CCW)
    inc_c = np.concatenate([cellsOnEdge[:, 0], cellsOnEdge[:, 1]])
    inc_n = np.concatenate([cellsOnEdge[:, 1], cellsOnEdge[:, 0]])
    inc_e = np.concatenate([np.arange(nE), np.arange(nE)])
    inc_v = np.concatenate([verticesOnEdge[:, 0], verticesOnEdge[:, 1]])  # CW-side vertex
    pc = p[inc_c]
    zhat = np.array([0.0, 0.0, 1.0])
    east = _normalize(np.cross(zhat, pc))
    north = np.cross(pc, east)
    d = p[inc_n] - pc
    ang = np.arctan2(np.einsum("ij,ij->i", d, north), np.einsum("ij,ij->i", d, east))
    order = np.lexsort((ang, inc_c))
    inc_c, inc_n, inc_e, inc_v = inc_c[order], inc_n[order], inc_e[order], inc_v[order]
    nEdgesOnCell = np.bincount(inc_c, minlength=nC).astype(np.int64)
    start = np.concatenate([[0], np.cumsum(nEdgesOnCell)[:-1]])
    slot = np.arange(inc_c.shape[0]) - start[inc_c]
    edgesOnCell = np.full((nC, MAX_EDGES), nE, dtype=np.int64)
    cellsOnCell = np.full((nC, MAX_EDGES), nC, dtype=np.int64)
    verticesOnCell = np.full((nC, MAX_EDGES), nV, dtype=np.int64)
    edgesOnCell[inc_c, slot] = inc_e
    cellsOnCell[inc_c, slot] = inc_n
    verticesOnCell[inc_c, slot] = inc_v

    # ---- kites: kiteAreasOnVertex(j,v) = area(dual triangle v  ∩  cell cellsOnVertex(j,v))
    # (mpas_atm_time_integration.F:6581-6593)
    kite = np.zeros((nV, 3))
    for j in range(3):
        c = p[tri[:, j]]
        m_next = xe[edgesOnVertex[:, j]]                 # towards cell j+1
        m_prev = xe[edgesOnVertex[:, (j + 2) % 3]]       # towards cell j-1
        kite[:, j] = _tri_area(c, m_next, xv) + _tri_area(c, xv, m_prev)
    areaTriangle = kite.sum(1)
    areaCell = np.zeros(nC)
    np.add.at(areaCell, tri.reshape(-1), kite.reshape(-1))

    # ---- angleEdge: angle of the normal from local east (TI:5633-5634)
    east_e = _normalize(np.cross(zhat, xe))
    north_e = np.cross(xe, east_e)
    nrm = p[cellsOnEdge[:, 1]] - p[cellsOnEdge[:, 0]]
    nrm = nrm - np.einsum("ij,ij->i", nrm, xe)[:, None] * xe
    angleEdge = np.arctan2(np.einsum("ij,ij->i", nrm, north_e), np.einsum("ij,ij->i", nrm, east_e))

    # ---- TRiSK tangential reconstruction (Thuburn et al. 2009), consumed at
    # TI:6624-6631 and TI:5418-5428: v_e = sum_j weightsOnEdge(j,e) u(edgesOnEdge(j,e))
    # is the velocity along k x n.
    kite_of = {}
    kcv = np.zeros((nC, MAX_EDGES))                      # kite area of (cell, its vertex slot)
    for j in range(3):
        # locate vertex v in verticesOnCell[tri[:, j]]
        c = tri[:, j]
        hit = verticesOnCell[c] == np.arange(nV)[:, None]
        s = np.argmax(hit, axis=1)
        assert hit.any(axis=1).all()
        kcv[c, s] = kite[:, j]
    del kite_of
    edgesOnEdge = np.full((nE, MAX_EDGES2), nE, dtype=np.int64)
    weightsOnEdge = np.zeros((nE, MAX_EDGES2))
    nEdgesOnEdge = np.zeros(nE, dtype=np.int64)
    ar = np.arange(nE)
    for side in (0, 1):
        c = cellsOnEdge[:, side]
        ne = nEdgesOnCell[c]
        i0 = np.argmax(edgesOnCell[c] == ar[:, None], axis=1)
        sum_r = np.zeros(nE)
        for j in range(1, MAX_EDGES):
            live = j < ne
            ii = (i0 + j) % ne
            eoe = edgesOnCell[c, ii]
            sum_r = sum_r + kcv[c, ii] / areaCell[c]
            s_eoe = np.where(cellsOnEdge[np.minimum(eoe, nE - 1), 0] == c, 1.0, -1.0)
            w = s_eoe * (0.5 - sum_r) * dvEdge[np.minimum(eoe, nE - 1)] / dcEdge
            if side == 1:
                w = -w
            col = nEdgesOnEdge + (j - 1)
            edgesOnEdge[ar[live], col[live]] = eoe[live]
            weightsOnEdge[ar[live], col[live]] = w[live]
        nEdgesOnEdge = nEdgesOnEdge + (ne - 1)

    latC, lonC = _latlon(p)
    latE, lonE = _latlon(xe)
    latV, lonV = _latlon(xv)
    nominalMinDc = float(np.sqrt(2.0 * (4.0 * np.pi / nC) / np.sqrt(3.0)))

    i32 = np.int32
    m = dict(
        nCells=nC, nEdges=nE, nVertices=nV, maxEdges=MAX_EDGES, maxEdges2=MAX_EDGES2,
        vertexDegree=VERTEX_DEGREE, sphere_radius=1.0, on_a_sphere=True,
        nominalMinDc=nominalMinDc,
        xCell=p[:, 0].copy(), yCell=p[:, 1].copy(), zCell=p[:, 2].copy(), latCell=latC, lonCell=lonC,
        xEdge=xe[:, 0].copy(), yEdge=xe[:, 1].copy(), zEdge=xe[:, 2].copy(), latEdge=latE, lonEdge=lonE,
        xVertex=xv[:, 0].copy(), yVertex=xv[:, 1].copy(), zVertex=xv[:, 2].copy(), latVertex=latV, lonVertex=lonV,
        dcEdge=dcEdge, dvEdge=dvEdge, angleEdge=angleEdge,
        areaCell=areaCell, areaTriangle=areaTriangle, kiteAreasOnVertex=kite,
        meshDensity=np.ones(nC),
        nEdgesOnCell=nEdgesOnCell.astype(i32), nEdgesOnEdge=nEdgesOnEdge.astype(i32),
        cellsOnEdge=cellsOnEdge.astype(i32), verticesOnEdge=verticesOnEdge.astype(i32),
        edgesOnCell=edgesOnCell.astype(i32), cellsOnCell=cellsOnCell.astype(i32),
        verticesOnCell=verticesOnCell.astype(i32),
        edgesOnVertex=edgesOnVertex.astype(i32), cellsOnVertex=cellsOnVertex.astype(i32),
        edgesOnEdge=edgesOnEdge.astype(i32), weightsOnEdge=weightsOnEdge,
        indexToCellID=np.arange(1, nC + 1, dtype=i32),
        indexToEdgeID=np.arange(1, nE + 1, dtype=i32),
        indexToVertexID=np.arange(1, nV + 1, dtype=i32),
    )
    return add_garbage_slots(m)


_CELL_FIELDS = ("xCell", "yCell", "zCell", "latCell", "lonCell", "areaCell", "meshDensity",
                "nEdgesOnCell", "edgesOnCell", "cellsOnCell", "verticesOnCell", "indexToCellID")
_EDGE_FIELDS = ("xEdge", "yEdge", "zEdge", "latEdge", "lonEdge", "dcEdge", "dvEdge", "angleEdge",
                "nEdgesOnEdge", "cellsOnEdge", "verticesOnEdge", "edgesOnEdge", "weightsOnEdge",
                "indexToEdgeID")
_VERTEX_FIELDS = ("xVertex", "yVertex", "zVertex", "latVertex", "lonVertex", "areaTriangle",
                  "kiteAreasOnVertex", "edgesOnVertex", "cellsOnVertex", "indexToVertexID")
_GARBAGE_OF = {"edgesOnCell": "nEdges", "cellsOnCell": "nCells", "verticesOnCell": "nVertices",
               "cellsOnEdge": "nCells", "verticesOnEdge": "nVertices", "edgesOnEdge": "nEdges",
               "edgesOnVertex": "nEdges", "cellsOnVertex": "nCells"}


def add_garbage_slots(m: dict) -> dict:
    """Append the trailing garbage element to every per-cell/edge/vertex array.
    Connectivity stored in the garbage row points at the garbage slot of its
    target; real-valued garbage is 0 (areas 1 so that inverses stay finite)."""
    out = dict(m)
    for names in (_CELL_FIELDS, _EDGE_FIELDS, _VERTEX_FIELDS):
        for k in names:
            a = m[k]
            if k in _GARBAGE_OF:
                fill = m[_GARBAGE_OF[k]]
            elif k in ("areaCell", "areaTriangle", "dcEdge", "dvEdge", "meshDensity"):
                fill = 1.0
            else:
                fill = 0
            out[k] = _pad_rows(a, fill)
    return out
