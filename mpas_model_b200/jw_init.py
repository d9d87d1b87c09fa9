"""Jablonowski-Williamson baroclinic-wave initial state (init_atmosphere case 1/2)
and the advection / deformation coefficient tables, restated in numpy.

Follows, line by line but vectorised over cells/edges:
  * init_atm_case_jw                     src/core_init_atmosphere/mpas_init_atm_cases.F:426-1231
  * init_atm_calc_flux_zonal             ... :1234-1282
  * init_atm_recompute_geostrophic_wind  ... :1286-1381
  * atm_initialize_advection_rk          src/core_init_atmosphere/mpas_atm_advection.F:21-394
  * atm_initialize_deformation_weights   ... :744-939
  * sphere_angle / arc_length / arc_bisect ... :403-564
Constants: src/framework/mpas_constants.F:43-56.

This is the input generator for every BASELINE.json config (no initial states
ship with the reference); it is host-side setup, as it is in the reference.
"""
from __future__ import annotations

import os

import numpy as np

# src/framework/mpas_constants.F:43-56
PII = 3.141592653589793
A_EARTH = 6371229.0
OMEGA = 7.29212e-5
GRAVITY = 9.80616
RGAS = 287.0
RV = 461.6
CP = 7.0 * RGAS / 2.0
RVORD = RV / RGAS
CV = CP - RGAS
P0 = 1.0e5
PRANDTL = 1.0


# ----------------------------------------------------------------------------
# spherical helpers (mpas_atm_advection.F:403-564), vectorised: inputs [...,3]
# ----------------------------------------------------------------------------
def arc_length(a, b):
    r = np.sqrt((a * a).sum(-1))
    c = np.sqrt(((b - a) ** 2).sum(-1))
    return r * 2.0 * np.arcsin(c / (2.0 * r))


def sphere_angle(a, b, c):
    """Signed angle between arcs AB and AC (mpas_atm_advection.F:403-447)."""
    la = arc_length(b, c)
    lb = arc_length(a, c)
    lc = arc_length(a, b)
    d = np.cross(b - a, c - a)
    s = 0.5 * (la + lb + lc)
    with np.errstate(invalid="ignore", divide="ignore"):
        q = (np.sin(s - lb) * np.sin(s - lc)) / (np.sin(lb) * np.sin(lc))
    sin_angle = np.sqrt(np.minimum(1.0, np.maximum(0.0, q)))
    ang = 2.0 * np.arcsin(np.maximum(np.minimum(sin_angle, 1.0), -1.0))
    return np.where((d * a).sum(-1) >= 0.0, ang, -ang)


def arc_bisect(a, b):
    r = np.sqrt((a * a).sum(-1))
    c = 0.5 * (a + b)
    d = np.sqrt((c * c).sum(-1))
    return r[..., None] * c / d[..., None]


def sphere_distance(lat1, lon1, lat2, lon2, radius):
    """src/core_init_atmosphere/mpas_init_atm_static.F:2241-2255"""
    arg1 = np.sqrt(np.sin(0.5 * (lat2 - lat1)) ** 2
                   + np.cos(lat1) * np.cos(lat2) * np.sin(0.5 * (lon2 - lon1)) ** 2)
    return 2.0 * radius * np.arcsin(arg1)


def _cell_rings(m):
    """Yield (cells, ne) groups of interior cells with the same edge count."""
    nC = m["nCells"]
    nec = m["nEdgesOnCell"][:nC]
    for ne in np.unique(nec):
        yield np.nonzero(nec == ne)[0], int(ne)


# ----------------------------------------------------------------------------
def initialize_advection_rk(m, sphere_radius):
    """deriv_two(15,2,nEdges): second-derivative stencil weights from a
    least-squares quadratic fit on the tangent plane of every cell
    (mpas_atm_advection.F:21-394, polynomial_order = 2, unit weights)."""
    nC, nE = m["nCells"], m["nEdges"]
    xc = np.stack([m["xCell"], m["yCell"], m["zCell"]], 1) / sphere_radius
    xv = np.stack([m["xVertex"], m["yVertex"], m["zVertex"]], 1) / sphere_radius
    deriv_two = np.zeros((nE + 1, 2, 15))
    pole = np.array([0.0, 0.0, 1.0])
    for cells, ne in _cell_rings(m):
        n = ne + 1
        nb = m["cellsOnCell"][cells, :ne]                 # [g, ne]
        c0 = xc[cells]                                    # [g, 3]
        cn = xc[nb]                                       # [g, ne, 3]
        theta_abs = np.where(c0[:, 2] == 1.0, PII / 2.0,
                             PII / 2.0 - sphere_angle(c0, cn[:, 0], pole[None, :]))
        c0b = np.broadcast_to(c0[:, None, :], cn.shape)
        thetav = sphere_angle(c0b, cn, np.roll(cn, -1, axis=1))          # [g, ne]
        dl = sphere_radius * arc_length(c0b, cn)
        thetat = np.empty_like(thetav)
        thetat[:, 0] = theta_abs
        for i in range(1, ne):
            thetat[:, i] = thetat[:, i - 1] + thetav[:, i - 1]
        xp = np.cos(thetat) * dl
        yp = np.sin(thetat) * dl
        # quadratic fit, rows = (cell itself, neighbours); the reference inverts
        # A^T A by scaled-pivot Gaussian elimination (poly_fit_2/MIGS, :567-741),
        # which is invariant to the column scaling applied here for conditioning.
        L = dl.mean(axis=1)[:, None]
        xs, ys = xp / L, yp / L
        amat = np.zeros((len(cells), n, 6))
        amat[:, 0, 0] = 1.0
        amat[:, 1:, 0] = 1.0
        amat[:, 1:, 1] = xs
        amat[:, 1:, 2] = ys
        amat[:, 1:, 3] = xs * xs
        amat[:, 1:, 4] = xs * ys
        amat[:, 1:, 5] = ys * ys
        at = np.swapaxes(amat, 1, 2)
        bmat = np.linalg.solve(at @ amat, at)             # [g, 6, n]
        bmat[:, 3:6, :] /= (L * L)[:, :, None]
        # angle of every edge of the cell in the tangent-plane frame
        eoc = m["edgesOnCell"][cells, :ne]
        v1 = xv[m["verticesOnEdge"][eoc, 0]]
        v2 = xv[m["verticesOnEdge"][eoc, 1]]
        xec = arc_bisect(v1, v2)
        thetae = sphere_angle(c0b, cn, xec) + thetat
        cos2t, sin2t = np.cos(thetae), np.sin(thetae)
        costsint = cos2t * sin2t
        cos2t, sin2t = cos2t ** 2, sin2t ** 2
        d2 = (2.0 * cos2t[:, :, None] * bmat[:, None, 3, :]
              + 2.0 * costsint[:, :, None] * bmat[:, None, 4, :]
              + 2.0 * sin2t[:, :, None] * bmat[:, None, 5, :])          # [g, ne, n]
        side = np.where(m["cellsOnEdge"][eoc, 0] == cells[:, None], 0, 1)
        deriv_two[eoc, side, :n] = d2
    return deriv_two


def initialize_deformation_weights(m, sphere_radius):
    """defc_a, defc_b (mpas_atm_advection.F:744-939)."""
    nC = m["nCells"]
    xc = np.stack([m["xCell"], m["yCell"], m["zCell"]], 1) / sphere_radius
    xv = np.stack([m["xVertex"], m["yVertex"], m["zVertex"]], 1) / sphere_radius
    defc_a = np.zeros((nC + 1, m["maxEdges"]))
    defc_b = np.zeros((nC + 1, m["maxEdges"]))
    pole = np.array([0.0, 0.0, 1.0])
    for cells, ne in _cell_rings(m):
        c0 = xc[cells]
        vn = xv[m["verticesOnCell"][cells, :ne]]
        theta_abs = np.where(c0[:, 2] == 1.0, PII / 2.0,
                             PII / 2.0 - sphere_angle(c0, vn[:, 0], pole[None, :]))
        c0b = np.broadcast_to(c0[:, None, :], vn.shape)
        thetav = sphere_angle(c0b, vn, np.roll(vn, -1, axis=1))
        dl_sphere = sphere_radius * arc_length(c0b, vn)
        thetat = np.empty_like(thetav)
        thetat[:, 0] = theta_abs
        for i in range(1, ne):
            thetat[:, i] = thetat[:, i - 1] + thetav[:, i - 1]
        xp = np.cos(thetat) * dl_sphere
        yp = np.sin(thetat) * dl_sphere
        xq, yq = np.roll(xp, -1, axis=1), np.roll(yp, -1, axis=1)
        dx, dy = xq - xp, yq - yp
        area_cell = (0.25 * (xp + xq) * (yq - yp) - 0.25 * (yp + yq) * (xq - xp)).sum(1)[:, None]
        th = np.arctan2(dy, dx) - PII / 2.0
        dl = np.sqrt(dx ** 2 + dy ** 2)
        sint2, cost2 = np.sin(th) ** 2, np.cos(th) ** 2
        sint_cost = np.sin(th) * np.cos(th)
        a = dl * (cost2 - sint2) / area_cell
        b = dl * 2.0 * sint_cost / area_cell
        eoc = m["edgesOnCell"][cells, :ne]
        flip = m["cellsOnEdge"][eoc, 0] != cells[:, None]
        defc_a[cells, :ne] = np.where(flip, -a, a)
        defc_b[cells, :ne] = np.where(flip, -b, b)
    return defc_a, defc_b


# ----------------------------------------------------------------------------
def _recompute_geostrophic_wind(u_2d, rho_2d, pp_2d, qv_2d, lat_2d, zz_2d, zx_2d,
                                cf1, cf2, cf3, fzm, fzp, rdzw, nz1, nlat, dlat, rad):
    """mpas_init_atm_cases.F:1286-1381.  2-D arrays are [nlat, nz1]."""
    rdx = 1.0 / (dlat * rad)
    pgrad = rdx * (pp_2d[1:] / zz_2d[1:] - pp_2d[:-1] / zz_2d[:-1])       # [nlat-1, nz1]
    dpzx = np.zeros((nlat - 1, nz1 + 1))
    s = pp_2d[1:] + pp_2d[:-1]
    dpzx[:, 0] = 0.5 * zx_2d[:, 0] * (cf1 * s[:, 0] + cf2 * s[:, 1] + cf3 * s[:, 2])
    dpzx[:, 1:nz1] = 0.5 * zx_2d[:, 1:] * (fzm[1:] * s[:, 1:] + fzp[1:] * s[:, :-1])
    pgrad = pgrad - rdzw * (dpzx[:, 1:] - dpzx[:, :-1])
    u = 0.5 * (u_2d[:-1] + u_2d[1:])
    rho_m = rho_2d[:-1] + rho_2d[1:]
    ru = u * rho_m * 0.5
    phi = (lat_2d[:-1] + lat_2d[1:]) / 2.0
    f = 2.0 * OMEGA * np.sin(phi)
    qtot = 0.5 * (qv_2d[:-1] + qv_2d[1:])
    for _ in range(50):
        with np.errstate(divide="ignore", invalid="ignore"):
            new = -(1.0 / (1.0 + qtot) * pgrad + (np.tan(phi) / rad)[:, None] * u * ru) / f[:, None]
        ru = np.where(f[:, None] == 0.0, 0.0, new)
        u = ru * 2.0 / rho_m
    out = u_2d.copy()
    out[1:nlat - 1] = (ru[:-1] + ru[1:]) * 0.5
    out[0] = (3.0 * out[1] - out[2]) * 0.5
    out[nlat - 1] = (3.0 * out[nlat - 2] - out[nlat - 3]) * 0.5
    return out


def _jw_columns(lat, zgrid, zz, dzw, dzu, fzm, fzp, qv, chunk=2048):
    """Iterative hydrostatic balance of the JW temperature profile on columns.
    lat [n]; zgrid [n, nz]; zz [n, nz1]; returns dict of [n, nz1] arrays.
    (mpas_init_atm_cases.F:803-886 for the (lat,z) slice and :912-1021 per cell.)
    Columns are independent: they are processed in cache-sized chunks on a thread pool
    (numpy releases the GIL inside its loops); the arithmetic per column is unchanged."""
    n = zgrid.shape[0]
    keys = ("ppb", "pb", "rb", "tb", "pp", "rr", "tt")
    if n <= chunk:
        return _jw_columns_chunk(lat, zgrid, zz, dzw, dzu, fzm, fzp, qv)
    out = {k: np.empty_like(zz) for k in keys}

    def work(s):
        e = slice(s, min(n, s + chunk))
        r = _jw_columns_chunk(lat[e], zgrid[e], zz[e], dzw, dzu, fzm, fzp, qv[e])
        for k in keys:
            out[k][e] = r[k]

    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=min(32, os.cpu_count() or 1)) as ex:
        list(ex.map(work, range(0, n, chunk)))
    return out


def _jw_columns_chunk(lat, zgrid, zz, dzw, dzu, fzm, fzp, qv):
    # level-major [nz1, n] work arrays: every level slice is contiguous
    u0, t0b, t0, delta_t, dtdz, znut = 35.0, 250.0, 288.0, 4.8e5, 0.005, 0.2
    n, nz = zgrid.shape
    nz1 = nz - 1
    zgrid = np.ascontiguousarray(zgrid.T)
    zz = np.ascontiguousarray(zz.T)
    qv = np.ascontiguousarray(qv.T)
    ztemp = 0.5 * (zgrid[1:] + zgrid[:-1])
    ppb = P0 * np.exp(-GRAVITY * ztemp / (RGAS * t0b))
    pb = (ppb / P0) ** (RGAS / CP)
    rb = ppb / (RGAS * t0b * zz)
    tb = t0b / pb
    pp = np.zeros_like(ppb)
    rr = np.zeros_like(ppb)
    phi = lat[None, :]
    lat_term_a = (-2.0 * np.sin(phi) ** 6 * (np.cos(phi) ** 2 + 1.0 / 3.0) + 10.0 / 63.0)
    lat_term_b = (1.6 * np.cos(phi) ** 3 * (np.sin(phi) ** 2 + 2.0 / 3.0) - PII / 4.0) * A_EARTH * OMEGA
    tt = None
    ppi = np.empty_like(pp)
    cz = (dzu[1:nz1] * GRAVITY)[:, None]
    fp_, fm_ = fzp[1:nz1][:, None], fzm[1:nz1][:, None]
    for _itr in range(10):
        eta = (ppb + pp) / P0
        etav = (eta - 0.252) * PII / 2.0
        teta = t0 * eta ** (RGAS * dtdz / GRAVITY)
        teta = np.where(eta >= znut, teta, teta + delta_t * np.maximum(znut - eta, 0.0) ** 5)
        temperature = teta + 0.75 * eta * PII * u0 / RGAS * np.sin(etav) * np.sqrt(np.cos(etav)) * (
            lat_term_a * 2.0 * u0 * np.cos(etav) ** 1.5 + lat_term_b) / (1.0 + 0.61 * qv)
        tt = temperature * (1.0 + 1.61 * qv)
        for _itrp in range(25):
            rr = (pp / (RGAS * zz) - rb * (tt - t0b)) / tt
            ppi[0] = P0 - 0.5 * dzw[0] * GRAVITY * (1.25 * (rr[0] + rb[0]) * (1.0 + qv[0])
                                                    - 0.25 * (rr[1] + rb[1]) * (1.0 + qv[1]))
            ppi[0] = ppi[0] - ppb[0]
            rq = rr + (rr + rb) * qv
            term = cz * (rq[:-1] * fp_ + rq[1:] * fm_)
            for k in range(nz1 - 1):
                np.subtract(ppi[k], term[k], out=ppi[k + 1])
            pp = 0.2 * ppi + 0.8 * pp
    return {k: np.ascontiguousarray(v.T) for k, v in
            dict(ppb=ppb, pb=pb, rb=rb, tb=tb, pp=pp, rr=rr, tt=tt).items()}


def _flux_zonal(u_2d, lat_2d, lat1_in, lat2_in, dvEdge, a, u0, nz1, nlat, chunk=65536):
    """init_atm_calc_flux_zonal (mpas_init_atm_cases.F:1234-1282) for all edges.
    u_2d is [nlat, nz1]; returns [nEdges, nz1]."""
    nE = lat1_in.shape[0]
    out = np.empty((nE, nz1))
    dlat0 = lat_2d[1] - lat_2d[0]
    for s in range(0, nE, chunk):
        e = slice(s, min(nE, s + chunk))
        l1i, l2i = lat1_in[e], lat2_in[e]
        lat1, lat2 = np.abs(l1i), np.abs(l2i)
        swap = lat2 <= lat1
        lat1, lat2 = np.where(swap, np.abs(l2i), lat1), np.where(swap, np.abs(l1i), lat2)
        fz = np.zeros((lat1.shape[0], nz1))
        dlat_last = np.full(lat1.shape[0], dlat0)
        i_lo = np.maximum(np.floor(lat1 / dlat0).astype(np.int64) - 2, 0)
        span = int(np.max(np.ceil((lat2 - lat1) / dlat0))) + 5
        for j in range(span):
            i = np.minimum(i_lo + j, nlat - 2)
            valid = (i_lo + j) <= nlat - 2
            la, lb = lat_2d[i], lat_2d[i + 1]
            hit = valid & (lat1 <= lb) & (lat2 >= la)
            dlat = lb - la
            da = (np.maximum(lat1, la) - la) / dlat
            db = (np.minimum(lat2, lb) - la) / dlat
            w1 = (db - da) - 0.5 * (db - da) ** 2
            w2 = 0.5 * (db - da) ** 2
            w1 = np.where(hit, w1, 0.0)
            w2 = np.where(hit, w2, 0.0)
            fz += w1[:, None] * u_2d[i] + w2[:, None] * u_2d[i + 1]
            dlat_last = np.where(hit, dlat, dlat_last)
        sgn = np.copysign(1.0, l2i - l1i)
        out[e] = sgn[:, None] * fz * (dlat_last * a / dvEdge[e] / u0)[:, None]
    return out


# ----------------------------------------------------------------------------
def init_atm_case_jw(m: dict, n_vert_levels: int = 26, init_case: int = 2,
                     coef_3rd_order: float = 0.25, theta_adv_order: int = 3,
                     num_scalars: int = 1) -> dict:
    """Returns a new dict: the mesh scaled to the Earth's radius plus every
    field init_atmosphere case 2 writes to its output stream.  ``init_case`` 1
    is the unperturbed (steady) state, 2 adds the Gaussian u perturbation."""
    d = dict(m)
    R = A_EARTH
    nC, nE, nV = m["nCells"], m["nEdges"], m["nVertices"]
    # :554-568
    for k in ("xCell", "yCell", "zCell", "xVertex", "yVertex", "zVertex", "xEdge", "yEdge", "zEdge",
              "dvEdge", "dcEdge"):
        d[k] = m[k] * R
    for k in ("areaCell", "areaTriangle", "kiteAreasOnVertex"):
        d[k] = m[k] * R ** 2.0
    d["nominalMinDc"] = m["nominalMinDc"] * R
    d["sphere_radius"] = R
    nz1 = n_vert_levels
    nz = nz1 + 1
    d["nVertLevels"] = nz1

    d["deriv_two"] = initialize_advection_rk(d, R)
    d["defc_a"], d["defc_b"] = initialize_deformation_weights(d, R)

    u0 = 35.0
    etavs = (1.0 - 0.252) * PII / 2.0
    latC = d["latCell"][:nC]

    def hx_of(phi):
        return u0 / GRAVITY * np.cos(etavs) ** 1.5 * (
            (-2.0 * np.sin(phi) ** 6 * (np.cos(phi) ** 2 + 1.0 / 3.0) + 10.0 / 63.0) * u0 * np.cos(etavs) ** 1.5
            + (1.6 * np.cos(phi) ** 3 * (np.sin(phi) ** 2 + 2.0 / 3.0) - PII / 4.0) * R * OMEGA)

    hx = hx_of(latC)
    # vertical grid (:675-740)
    str_, zt = 1.5, 45000.0
    dz = zt / float(nz1)
    kk = np.arange(nz, dtype=np.float64)
    sh = (kk * dz / zt) ** str_
    zw = kk * dz
    ah = 1.0 - np.cos(0.5 * PII * kk * dz / zt) ** 6
    dzw = zw[1:] - zw[:-1]
    rdzw = 1.0 / dzw
    dzu = np.zeros(nz1)
    rdzu = np.zeros(nz1)
    fzp = np.zeros(nz1)
    fzm = np.zeros(nz1)
    dzu[1:] = 0.5 * (dzw[1:] + dzw[:-1])
    rdzu[1:] = 1.0 / dzu[1:]
    fzp[1:] = 0.5 * dzw[1:] / dzu[1:]          # linear_interpolation
    fzm[1:] = 0.5 * dzw[:-1] / dzu[1:]
    cof1 = (2.0 * dzu[1] + dzu[2]) / (dzu[1] + dzu[2]) * dzw[0] / dzu[1]
    cof2 = dzu[1] / (dzu[1] + dzu[2]) * dzw[0] / dzu[2]
    cf1 = fzp[1] + cof1
    cf2 = fzm[1] - cof1 - cof2
    cf3 = cof2

    def zgrid_of(hx_):
        return (1.0 - ah) * (sh * (zt - hx_[:, None]) + hx_[:, None]) + ah * sh * zt

    zgrid = zgrid_of(hx)                                            # [nC, nz]
    zz = (zw[1:] - zw[:-1]) / (zgrid[:, 1:] - zgrid[:, :-1])        # [nC, nz1]
    c1, c2 = d["cellsOnEdge"][:nE, 0], d["cellsOnEdge"][:nE, 1]
    dcE, dvE = d["dcEdge"][:nE], d["dvEdge"][:nE]
    zxu = 0.5 * (zgrid[c2, :-1] - zgrid[c1, :-1] + zgrid[c2, 1:] - zgrid[c1, 1:]) / dcE[:, None]

    # ---- (lat, z) slice for the balanced zonal wind (:789-903)
    nlat = 721
    dlat = 0.5 * PII / float(nlat - 1)
    lat_2d = np.arange(nlat, dtype=np.float64) * dlat
    zgrid_2d = zgrid_of(hx_of(lat_2d))
    zz_2d = (zw[1:] - zw[:-1]) / (zgrid_2d[:, 1:] - zgrid_2d[:, :-1])
    qv_2d = np.zeros((nlat, nz1))
    col2 = _jw_columns(lat_2d, zgrid_2d, zz_2d, dzw, dzu, fzm, fzp, qv_2d)
    rho_2d = col2["rr"] + col2["rb"]
    etavs_2d = ((col2["ppb"] + col2["pp"]) / P0 - 0.252) * PII / 2.0
    u_2d = u0 * (np.sin(2.0 * lat_2d) ** 2)[:, None] * np.cos(etavs_2d) ** 1.5
    zx_2d = (zgrid_2d[1:, :-1] - zgrid_2d[:-1, :-1]) / (dlat * R)   # [nlat-1, nz1] (uses level k only, :896)
    u_2d = _recompute_geostrophic_wind(u_2d, rho_2d, col2["pp"], qv_2d, lat_2d, zz_2d, zx_2d,
                                       cf1, cf2, cf3, fzm, fzp, rdzw, nz1, nlat, dlat, R)

    # ---- cell columns (:912-1029)
    qv = np.zeros((nC, nz1))
    col = _jw_columns(latC, zgrid, zz, dzw, dzu, fzm, fzp, qv)
    ppb, pb, rb, tb, pp, rr, tt = (col[k] for k in ("ppb", "pb", "rb", "tb", "pp", "rr", "tt"))
    p = ((ppb + pp) / P0) ** (RGAS / CP)
    t = tt / p
    rho_zz = rb + rr
    surface_pressure = 0.5 * dzw[0] * GRAVITY * (1.25 * (rr[:, 0] + rb[:, 0]) * (1.0 + qv[:, 0])
                                                  - 0.25 * (rr[:, 1] + rb[:, 1]) * (1.0 + qv[:, 1]))
    surface_pressure = surface_pressure + pp[:, 0] + ppb[:, 0]

    # ---- edge-normal wind (:1042-1100)
    lat_pert, lon_pert = 40.0 * PII / 180.0, 20.0 * PII / 180.0
    vtx1, vtx2 = d["verticesOnEdge"][:nE, 0], d["verticesOnEdge"][:nE, 1]
    lat1, lat2 = d["latVertex"][vtx1], d["latVertex"][vtx2]
    latE, lonE = d["latEdge"][:nE], d["lonEdge"][:nE]
    if init_case == 2:
        r_pert = sphere_distance(latE, lonE, lat_pert, lon_pert, 1.0) / 0.1
        u_pert = 1.0 * np.exp(-r_pert ** 2) * (lat2 - lat1) * R / dvE
    elif init_case == 3:
        u_pert = 1.0 * np.cos(9.0 * (lonE - lon_pert)) * (
            0.5 * (lat2 - lat1) - 0.125 * (np.sin(4.0 * lat2) - np.sin(4.0 * lat1))) * R / dvE
    else:
        u_pert = np.zeros(nE)
    flux_zonal = _flux_zonal(u_2d, lat_2d, lat1, lat2, dvE, R, u0, nz1, nlat)
    u = u0 * flux_zonal / (0.5 * (rb[c1] + rb[c2] + rr[c1] + rr[c2])) + u_pert[:, None]
    ru = 0.5 * (rho_zz[c1] + rho_zz[c2]) * u
    alpha_grid = 0.0
    fEdge = 2.0 * OMEGA * (-np.cos(lonE) * np.cos(latE) * np.sin(alpha_grid) + np.sin(latE) * np.cos(alpha_grid))
    latV, lonV = d["latVertex"][:nV], d["lonVertex"][:nV]
    fVertex = 2.0 * OMEGA * (-np.cos(lonV) * np.cos(latV) * np.sin(alpha_grid) + np.sin(latV) * np.cos(alpha_grid))

    # ---- zb, zb3 (:1116-1164), theta_adv_order 3
    coc = d["cellsOnCell"]
    nec = d["nEdgesOnCell"]
    zg = np.concatenate([zgrid, np.zeros((1, nz))])                 # garbage cell row
    dt2 = d["deriv_two"][:nE]
    d2 = [None, None]
    for s, cc in ((0, c1), (1, c2)):
        acc = dt2[:, s, 0:1] * zg[cc]
        for i in range(m["maxEdges"]):
            live = (i < nec[cc])[:, None]
            acc = acc + np.where(live, dt2[:, s, i + 1:i + 2] * zg[coc[cc, i]], 0.0)
        d2[s] = acc
    if theta_adv_order == 2:
        z_edge = (zgrid[c1] + zgrid[c2]) / 2.0
        z_edge3 = np.zeros_like(z_edge)
    else:
        z_edge = 0.5 * (zgrid[c1] + zgrid[c2]) - (dcE[:, None] ** 2) * (d2[0] + d2[1]) / 12.0
        z_edge3 = -(dcE[:, None] ** 2) * (d2[0] - d2[1]) / 12.0 if theta_adv_order == 3 else np.zeros_like(z_edge)
    areaC = d["areaCell"]
    zb = np.zeros((nE + 1, 2, nz))
    zb3 = np.zeros((nE + 1, 2, nz))
    zb[:nE, 0] = (z_edge - zgrid[c1]) * (dvE / areaC[c1])[:, None]
    zb[:nE, 1] = (z_edge - zgrid[c2]) * (dvE / areaC[c2])[:, None]
    zb3[:nE, 0] = z_edge3 * (dvE / areaC[c1])[:, None]
    zb3[:nE, 1] = z_edge3 * (dvE / areaC[c2])[:, None]
    # the reference fills levels 1..nVertLevels only (:1123); level nz stays 0
    zb[:, :, nz1] = 0.0
    zb3[:, :, nz1] = 0.0

    # ---- rw, w from the terrain-following omega = 0 condition (:1167-1197)
    rw = np.zeros((nC, nz))
    flux = fzm[1:] * ru[:, 1:] + fzp[1:] * ru[:, :-1]               # k = 2..nz1  -> [nE, nz1-1]
    zzf = fzm[1:] * zz[:, 1:] + fzp[1:] * zz[:, :-1]                 # [nC, nz1-1]
    sg = np.copysign(1.0, ru[:, 1:])
    t2 = zzf[c2] * zb[:nE, 1, 1:nz1] * flux
    t1 = -zzf[c1] * zb[:nE, 0, 1:nz1] * flux
    if theta_adv_order == 3:
        t2 = t2 - sg * coef_3rd_order * zzf[c2] * zb3[:nE, 1, 1:nz1] * flux
        t1 = t1 + sg * coef_3rd_order * zzf[c1] * zb3[:nE, 0, 1:nz1] * flux
    # accumulate in edge order, as the reference's edge loop does
    np.add.at(rw[:, 1:nz1], c2, t2)
    np.add.at(rw[:, 1:nz1], c1, t1)
    w = np.zeros((nC, nz))
    w[:, 1:nz1] = rw[:, 1:nz1] / (fzp[1:] * rho_zz[:, :-1] + fzm[1:] * rho_zz[:, 1:])

    rho = rho_zz * zz
    theta = t / (1.0 + 1.61 * qv)

    def pad(a):
        return np.concatenate([a, np.zeros((1,) + a.shape[1:], dtype=a.dtype)])

    scalars = np.zeros((nC + 1, nz1, num_scalars))
    d.update(
        hx=pad(hx), zgrid=pad(zgrid), zz=pad(zz), zxu=pad(zxu), zb=zb, zb3=zb3,
        rdzw=rdzw, dzu=dzu, rdzu=rdzu, fzm=fzm, fzp=fzp, cf1=float(cf1), cf2=float(cf2), cf3=float(cf3),
        fEdge=pad(fEdge), fVertex=pad(fVertex),
        u=pad(u), w=pad(w), rho=pad(rho), theta=pad(theta), scalars=scalars,
        rho_base=pad(rb), theta_base=pad(tb), surface_pressure=pad(surface_pressure),
        # fields the init file also carries but the model re-derives (kept for checks)
        rho_zz_init=pad(rho_zz), theta_m_init=pad(t), ru_init=pad(ru), rw_init=pad(rw),
        pressure_p_init=pad(pp), pressure_base_init=pad(ppb), exner_init=pad(p), exner_base_init=pad(pb),
        u_init=np.zeros(nz1), v_init=np.zeros(nz1), qv_init=np.zeros(nz1), t_init=np.zeros((nC + 1, nz1)),
        num_scalars=num_scalars, index_qv=0, moist_start=0, moist_end=0,
    )
    return d
