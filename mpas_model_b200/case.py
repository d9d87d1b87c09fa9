"""Build a complete JW baroclinic-wave case (mesh + initial state + derived
mesh fields + namelist) for one of the BASELINE.json configs."""
from __future__ import annotations

import numpy as np

from . import init_block, jw_init, mesh


def make_case(n_cells: int, n_levels: int, dt: float | None = None, init_case: int = 2,
              num_scalars: int = 1, lloyd_iters: int = 12, jitter: float = 0.0, derive: bool = True, **cfg_overrides):
    """Returns (block, cfg).  ``dt`` defaults to the reference's rule of thumb
    of 6 s per km of nominal grid distance.  ``derive=False`` leaves out the init-time derived
    fields (``init_block``): a case that is only going to be decomposed gets them per block."""
    m = mesh.generate(n_cells, lloyd_iters=lloyd_iters, jitter=jitter)
    d = jw_init.init_atm_case_jw(m, n_levels, init_case=init_case, num_scalars=num_scalars)
    if dt is None:
        dt = 6.0 * round(d["nominalMinDc"] / 1000.0)
    cfg = init_block.default_config(d["nominalMinDc"], dt)
    cfg.update(cfg_overrides)
    if derive:
        init_block.init_block(d, cfg)
    if num_scalars > 1:
        add_passive_tracers(d)
    return d, cfg


def add_passive_tracers(d):
    """Smooth positive analytic blobs (cosine bells) for scalars 2..S."""
    nC, S = d["nCells"], d["num_scalars"]
    lat, lon = d["latCell"][:nC], d["lonCell"][:nC]
    nz = d["nVertLevels"]
    prof = np.sin(np.pi * (np.arange(nz) + 0.5) / nz) ** 2
    for s in range(1, S):
        lat0 = np.pi / 3.0 * np.sin(1.7 * s)
        lon0 = 2.0 * np.pi * ((0.37 * s) % 1.0)
        r = jw_init.sphere_distance(lat, lon, lat0, lon0, 1.0)
        bell = np.where(r < 0.6, 0.5 * (1.0 + np.cos(np.pi * r / 0.6)), 0.0)
        d["scalars"][:nC, :, s] = 1.0e-3 * (0.1 + bell[:, None] * prof[None, :])


N_RELAX_ZONE, N_SPEC_ZONE = 5, 2          # mpas_atm_boundaries.F:35-38


def make_regional(d: dict, cfg: dict, lat0: float = 0.6, lon0: float = 0.4, radius: float = 1.7, seed: int = 7,
                  lbc_interval: float = 10800.0):
    """Turn a global case into a limited-area one for the regional path (config_apply_lbcs, TI:7198-7910): everything
    further than ``radius`` (radians) from (lat0, lon0) becomes the outermost specified-zone ring (bdyMaskCell = 7), and
    the rings 6, 5 ... 1 are the successive cell layers inside it, as the limited-area mesh tool marks them; an edge takes
    the smaller mask of its two cells.  specZoneMask* as mpas_atm_boundaries.F:729-730.  The driving fields are the
    case's own state plus a smooth offset (time level 2 = state at the end of the LBC interval) and a small tendency
    (time level 1), so that every relaxation and specified-zone branch does non-trivial arithmetic.
    Returns (block, cfg, seconds to the end of the LBC interval at the start of the first step)."""
    d = dict(d)
    cfg = dict(cfg, config_apply_lbcs=True)
    nC, nE = d["nCells"], d["nEdges"]
    lat, lon = d["latCell"][:nC], d["lonCell"][:nC]
    dist = jw_init.sphere_distance(lat, lon, lat0, lon0, 1.0)
    mask = np.zeros(nC + 1, dtype=np.int32)
    mask[:nC][dist > radius] = N_RELAX_ZONE + N_SPEC_ZONE
    coc, ne = d["cellsOnCell"], d["nEdgesOnCell"]
    for ring in range(N_RELAX_ZONE + N_SPEC_ZONE - 1, 0, -1):
        outer = mask[:nC] == ring + 1
        nxt = np.zeros(nC, dtype=bool)
        for j in range(coc.shape[1]):
            nb = coc[:nC, j]
            ok = (j < ne[:nC]) & (nb < nC)
            nxt |= ok & outer[np.minimum(nb, nC - 1)]
        mask[:nC][nxt & (mask[:nC] == 0)] = ring
    assert (mask[:nC] == 0).sum() > 0 and all((mask[:nC] == m).any() for m in range(1, 8)), "region too small for 7 rings"
    c1, c2 = d["cellsOnEdge"][:nE, 0], d["cellsOnEdge"][:nE, 1]
    emask = np.zeros(nE + 1, dtype=np.int32)
    emask[:nE] = np.minimum(mask[c1], mask[c2])
    d["bdyMaskCell"], d["bdyMaskEdge"] = mask, emask
    d["specZoneMaskCell"] = (mask > N_RELAX_ZONE).astype(np.float64)
    d["specZoneMaskEdge"] = (emask > N_RELAX_ZONE).astype(np.float64)
    rho, th, u = d["rho_zz_init"], d["theta_m_init"], d["u"]         # what atm_init_coupled_diagnostics derives from the state
    wob_c = 1.0 + 1.0e-3 * np.sin(3.0 * d["latCell"])[:, None] * np.cos(2.0 * d["lonCell"])[:, None]
    wob_e = 1.0 + 1.0e-2 * np.sin(2.0 * d["latEdge"])[:, None] * np.cos(3.0 * d["lonEdge"])[:, None]
    d["lbc_rho_zz_2"] = rho * wob_c
    d["lbc_rtheta_m_2"] = rho * th * wob_c * (1.0 + 5.0e-4)
    d["lbc_u_2"] = u * wob_e
    d["lbc_ru_2"] = d["ru_init"] * wob_e
    d["lbc_scalars_2"] = d["scalars"] * 1.05 + 1.0e-5
    # tendencies over the interval (per second): smooth, a 0.1 % drift of the driving state over three hours
    ten_c = 1.0e-7 * np.cos(2.0 * d["latCell"]) * np.sin(d["lonCell"] + float(seed))
    ten_e = 1.0e-7 * np.cos(2.0 * d["latEdge"]) * np.sin(d["lonEdge"] + float(seed))
    for n, ten in (("rho_zz", ten_c), ("rtheta_m", ten_c), ("scalars", ten_c), ("u", ten_e), ("ru", ten_e)):
        st = d["lbc_" + n + "_2"]
        d["lbc_" + n] = st * ten.reshape((-1,) + (1,) * (st.ndim - 1))
    return d, cfg, lbc_interval
