"""Init-time coefficients of the cell-centre velocity reconstruction.

Restates, vectorised over cells with the reference's operation order inside each cell,

* ``mpas_initialize_vectors``   src/operators/mpas_vector_operations.F:652-771
  (``localVerticalUnitVectors``, ``edgeNormalVectors``, ``cellTangentPlane``),
* ``mpas_init_reconstruct``     src/operators/mpas_vector_reconstruction.F:60-177
  (``coeffs_reconstruct(3, maxEdges, nCells)``),
* ``mpas_rbf_interp_func_3D_plane_vec_const_dir_comp_coeffs``
                                src/operators/mpas_rbf_interpolation.F:1079-1145,
  ``mpas_set_up_vector_dirichlet_rbf_matrix_and_rhs`` :1527-1559, ``evaluate_rbf`` :1369-1376
  (inverse multiquadric), ``mpas_legs`` :1670-1698 and ``elgs`` :1782-1846 (Gaussian
  elimination with scaled partial pivoting).

The step itself only consumes ``coeffs_reconstruct`` (``mpas_reconstruct_2d``, :205-330): it is an
input of the C ABI like every other init-time mesh field (row M).
"""
from __future__ import annotations

import numpy as np


def _unit(v):
    """mpas_unit_vec_in_r3 (mpas_vector_operations.F:40-60): v / sqrt(sum(v**2))."""
    return v / np.sqrt((v[..., 0] ** 2 + v[..., 1] ** 2) + v[..., 2] ** 2)[..., None]


def _dot3(a, b):
    return (a[..., 0] * b[..., 0] + a[..., 1] * b[..., 1]) + a[..., 2] * b[..., 2]


def mpas_initialize_vectors(d: dict) -> None:
    """Spherical, non-periodic branch of mpas_vector_operations.F:652-771.  Adds (C order)
    ``localVerticalUnitVectors [nCells+1, 3]``, ``edgeNormalVectors [nEdges+1, 3]``,
    ``cellTangentPlane [nCells+1, 2, 3]``."""
    nC, nE = d["nCells"], d["nEdges"]
    xc = np.stack([d["xCell"], d["yCell"], d["zCell"]], axis=1)[: nC + 1].astype(np.float64)
    xe = np.stack([d["xEdge"], d["yEdge"], d["zEdge"]], axis=1)[: nE + 1].astype(np.float64)
    lv = np.zeros((nC + 1, 3))
    lv[:nC] = _unit(xc[:nC])
    c1, c2 = d["cellsOnEdge"][:nE, 0], d["cellsOnEdge"][:nE, 1]
    en = np.zeros((nE + 1, 3))
    v = xc[c2] - xc[c1]                                     # interior edge
    b1 = c1 == nC                                           # cell1 outside the block (:713-722)
    b2 = c2 == nC                                           # cell2 outside the block (:724-733)
    v[b1] = xc[c2[b1]] - xe[:nE][b1]
    v[b2] = xe[:nE][b2] - xc[c1[b2]]
    both = b1 & b2                                          # outermost halo edges of a block: never referenced by an owned cell
    v[both] = (1.0, 0.0, 0.0)
    en[:nE] = _unit(v)
    e0 = d["edgesOnCell"][:nC, 0]
    rhat = lv[:nC]
    ndr = _dot3(en[e0], rhat)
    xhat = _unit(en[e0] - ndr[:, None] * rhat)
    yhat = _unit(np.cross(rhat, xhat))
    tp = np.zeros((nC + 1, 2, 3))
    tp[:nC, 0], tp[:nC, 1] = xhat, yhat
    d["localVerticalUnitVectors"], d["edgeNormalVectors"], d["cellTangentPlane"] = lv, en, tp


def _elgs(A):
    """elgs (mpas_rbf_interpolation.F:1782-1846) on a batch A[B, N, N] (in place); returns INDX[B, N]."""
    B, N, _ = A.shape
    bi = np.arange(B)
    indx = np.tile(np.arange(N), (B, 1))
    c = np.abs(A).max(axis=2)
    for j in range(N - 1):
        pi1 = np.zeros(B)
        k = np.full(B, j)
        for i in range(j, N):
            r = indx[:, i]
            pi = np.abs(A[bi, r, j]) / c[bi, r]
            upd = pi > pi1
            pi1 = np.where(upd, pi, pi1)
            k = np.where(upd, i, k)
        itmp = indx[bi, j].copy()
        indx[bi, j] = indx[bi, k]
        indx[bi, k] = itmp
        rj = indx[:, j]
        for i in range(j + 1, N):
            ri = indx[:, i]
            pj = A[bi, ri, j] / A[bi, rj, j]
            A[bi, ri, j] = pj
            A[bi, ri, j + 1:] = A[bi, ri, j + 1:] - pj[:, None] * A[bi, rj, j + 1:]
    return indx


def _legs(A, indx, b):
    """Forward/back substitution of mpas_legs (:1683-1697) with the factors of ``_elgs``."""
    B, N, _ = A.shape
    bi = np.arange(B)
    b = b.copy()
    for i in range(N - 1):
        for j in range(i + 1, N):
            b[bi, indx[:, j]] = b[bi, indx[:, j]] - A[bi, indx[:, j], i] * b[bi, indx[:, i]]
    x = np.zeros((B, N))
    x[:, N - 1] = b[bi, indx[:, N - 1]] / A[bi, indx[:, N - 1], N - 1]
    for i in range(N - 2, -1, -1):
        xi = b[bi, indx[:, i]]
        for j in range(i + 1, N):
            xi = xi - A[bi, indx[:, i], j] * x[:, j]
        x[:, i] = xi / A[bi, indx[:, i], i]
    return x


def _rbf_plane_vec_const_dir_coeffs(src, uvec, dest, alpha, plane):
    """mpas_rbf_interp_func_3D_plane_vec_const_dir_comp_coeffs for a batch of cells with the same
    pointCount n: src, uvec [B, n, 3]; dest [B, 3]; alpha [B]; plane [B, 2, 3] -> coefficients [B, n, 3]."""
    B, n, _ = src.shape
    ps = np.stack([_dot3(src, plane[:, None, 0]), _dot3(src, plane[:, None, 1])], axis=2)       # planarSourcePoints
    pu = np.stack([_dot3(uvec, plane[:, None, 0]), _dot3(uvec, plane[:, None, 1])], axis=2)     # planarUnitVectors
    pd = np.stack([_dot3(dest, plane[:, 0]), _dot3(dest, plane[:, 1])], axis=1)                 # planarDestinationPoint
    a2 = (alpha ** 2)[:, None, None]
    diff = ps[:, :, None, :] - ps[:, None, :, :]
    rsq = (diff[..., 0] ** 2 + diff[..., 1] ** 2) / a2
    rbf = 1.0 / np.sqrt(1.0 + rsq)
    udot = pu[:, :, None, 0] * pu[:, None, :, 0] + pu[:, :, None, 1] * pu[:, None, :, 1]
    M = np.zeros((B, n + 2, n + 2))
    M[:, :n, :n] = rbf * udot
    dd = pd[:, None, :] - ps
    rsq_d = (dd[..., 0] ** 2 + dd[..., 1] ** 2) / a2[:, :, 0]
    rhs = np.zeros((B, n + 2, 2))
    rhs[:, :n, :] = (1.0 / np.sqrt(1.0 + rsq_d))[:, :, None] * pu
    M[:, :n, n:n + 2] = pu
    M[:, n:n + 2, :n] = np.transpose(pu, (0, 2, 1))
    rhs[:, n, 0] = 1.0
    rhs[:, n + 1, 1] = 1.0
    indx = _elgs(M)
    co1 = _legs(M, indx, rhs[:, :, 0])[:, :n]
    co2 = _legs(M, indx, rhs[:, :, 1])[:, :n]
    return plane[:, None, 0, :] * co1[:, :, None] + plane[:, None, 1, :] * co2[:, :, None]


def mpas_init_reconstruct(d: dict, include_halos: bool = False) -> None:
    """``coeffs_reconstruct`` [nCells+1, maxEdges, 3] (Fortran (3, maxEdges, nCells+1)); cells beyond the
    owned prefix stay zero unless ``include_halos`` (mpas_vector_reconstruction.F:115-119)."""
    if "edgeNormalVectors" not in d:
        mpas_initialize_vectors(d)
    nC = d["nCells"]
    ncell = nC if include_halos else int(d.get("nCellsSolve", nC))
    mx = d["maxEdges"]
    co = np.zeros((nC + 1, mx, 3))
    xc = np.stack([d["xCell"], d["yCell"], d["zCell"]], axis=1).astype(np.float64)
    xe = np.stack([d["xEdge"], d["yEdge"], d["zEdge"]], axis=1).astype(np.float64)
    ne = d["nEdgesOnCell"][:ncell]
    for n in np.unique(ne):
        cells = np.nonzero(ne == n)[0]
        e = d["edgesOnCell"][cells, :n]
        src = xe[e]
        uvec = d["edgeNormalVectors"][e]
        dest = xc[cells]
        diff = dest[:, None, :] - src
        r = np.sqrt((diff[..., 0] ** 2 + diff[..., 1] ** 2) + diff[..., 2] ** 2)
        alpha = np.zeros(len(cells))
        for i in range(n):
            alpha = alpha + r[:, i]
        alpha = alpha / n
        co[cells, :n, :] = _rbf_plane_vec_const_dir_coeffs(src, uvec, dest, alpha, d["cellTangentPlane"][cells])
    d["coeffs_reconstruct"] = co
