"""Init-time derived mesh fields the dycore kernels consume.

numpy restatement of atm_mpas_init_block's mesh part:
  * atm_compute_signs            src/core_atmosphere/mpas_atm_core.F:1151-1238
  * inverses                     ... :456-470
  * atm_adv_coef_compression     ... :1285-1430
  * atm_couple_coef_3rd_order    ... :1433-1452
  * atm_compute_mesh_scaling     ... :1091-1148
  * atm_compute_damping_coefs    ... :1241-1282
Works on a whole mesh or on one decomposed block (anything pointing at the
garbage slot is treated as "outside the block", as the reference does).
"""
from __future__ import annotations

import numpy as np

from .jw_init import PII


def default_config(len_disp: float, dt: float) -> dict:
    """Namelist defaults, src/core_atmosphere/Registry.xml:63-395."""
    return dict(
        config_dt=dt, config_time_integration_order=2, config_split_dynamics_transport=True,
        config_dynamics_split_steps=3, config_number_of_sub_steps=2,
        config_horiz_mixing="2d_smagorinsky", config_visc4_2dsmag=0.05, config_smagorinsky_coef=0.125,
        config_del4u_div_factor=10.0, config_h_mom_eddy_visc2=0.0, config_h_mom_eddy_visc4=0.0,
        config_v_mom_eddy_visc2=0.0, config_h_theta_eddy_visc2=0.0, config_h_theta_eddy_visc4=0.0,
        config_v_theta_eddy_visc2=0.0, config_len_disp=len_disp, config_mix_full=True,
        config_coef_3rd_order=0.25, config_epssm=0.1, config_smdiv=0.1, config_apvm_upwinding=0.5,
        config_scalar_advection=True, config_monotonic=True, config_positive_definite=False,
        config_h_ScaleWithMesh=True, config_zd=22000.0, config_xnutr=0.2,
        config_mpas_cam_coef=0.0, config_number_cam_damping_levels=4,
        config_rayleigh_damp_u=False, config_rayleigh_damp_u_timescale_days=5.0,
        config_number_rayleigh_damp_u_levels=6, config_apply_lbcs=False, config_relax_zone_divdamp_coef=6.0,
    )


def compute_signs(d: dict) -> None:
    nC, nE, nV = d["nCells"], d["nEdges"], d["nVertices"]
    eov = d["edgesOnVertex"]
    sgn = np.where(d["verticesOnEdge"][eov, 1] == np.arange(nV + 1)[:, None], 1.0, -1.0)
    sgn[eov >= nE] = 0.0
    sgn[nV] = 0.0
    d["edgesOnVertex_sign"] = sgn

    eoc = d["edgesOnCell"]
    mx = d["maxEdges"]
    live = np.arange(mx)[None, :] < d["nEdgesOnCell"][:, None]
    own1 = d["cellsOnEdge"][eoc, 0] == np.arange(nC + 1)[:, None]
    s = np.where(own1, 1.0, -1.0)
    s[(eoc >= nE) | ~live] = 0.0
    s[nC] = 0.0
    d["edgesOnCell_sign"] = s
    nz = d["zb"].shape[2]
    side = np.where(own1, 0, 1)
    zb_cell = d["zb"][eoc, side]                          # [nC+1, mx, nz]
    zb3_cell = d["zb3"][eoc, side]
    dead = ((eoc >= nE) | ~live)
    zb_cell[dead] = 0.0
    zb3_cell[dead] = 0.0
    d["zb_cell"] = zb_cell
    d["zb3_cell"] = zb3_cell

    voc = d["verticesOnCell"]
    cov = d["cellsOnVertex"][voc]                         # [nC+1, mx, 3]
    hit = cov == np.arange(nC + 1)[:, None, None]
    k = np.argmax(hit, axis=2)
    k[(voc >= nV) | ~live] = 0
    d["kiteForCell"] = k.astype(np.int32)


def adv_coef_compression(d: dict) -> None:
    nC, nE = d["nCells"], d["nEdges"]
    coe, coc, nec = d["cellsOnEdge"][:nE], d["cellsOnCell"], d["nEdgesOnCell"]
    c1, c2 = coe[:, 0].astype(np.int64), coe[:, 1].astype(np.int64)
    mx = d["maxEdges"]
    lst = np.full((nE, 20), -1, dtype=np.int64)
    lst[:, 0], lst[:, 1] = c1, c2
    n = np.full(nE, 2, dtype=np.int64)
    ar = np.arange(nE)
    for i in range(mx):
        cand = coc[c1, i]
        add = (i < nec[c1]) & (cand != c2)
        lst[ar[add], n[add]] = cand[add]
        n = n + add
    for i in range(mx):
        cand = coc[c2, i].astype(np.int64)
        add = (i < nec[c2]) & ~(lst == cand[:, None]).any(axis=1)
        lst[ar[add], n[add]] = cand[add]
        n = n + add
    active = (c1 < nC) | (c2 < nC)
    adv = np.zeros((nE, 15))
    adv3 = np.zeros((nE, 15))
    dt2 = d["deriv_two"][:nE]

    def pos(target):
        # the LAST match, as the reference's `do j=1,n; if (cell_list(j) == ...) j_in = j` (it differs from the first only where
        # several neighbours of a block's outer halo cell point at the garbage cell)
        return 14 - np.argmax((lst[:, :15] == target[:, None])[:, ::-1], axis=1)

    for side, cc, sg3 in ((0, c1, 1.0), (1, c2, -1.0)):
        j = pos(cc)
        adv[ar, j] += dt2[:, side, 0]
        adv3[ar, j] += sg3 * dt2[:, side, 0]
        for i in range(mx):
            live = i < nec[cc]
            j = pos(coc[cc, i].astype(np.int64))
            adv[ar[live], j[live]] += dt2[live, side, i + 1]
            adv3[ar[live], j[live]] += sg3 * dt2[live, side, i + 1]
    dc2 = (d["dcEdge"][:nE] ** 2)[:, None]
    adv = -dc2 * adv / 12.0
    adv3 = -dc2 * adv3 / 12.0
    adv[ar, pos(c1)] += 0.5
    adv[ar, pos(c2)] += 0.5
    dv = d["dvEdge"][:nE, None]
    adv, adv3 = dv * adv, dv * adv3
    valid = np.arange(15)[None, :] < n[:, None]
    adv[~valid] = 0.0
    adv3[~valid] = 0.0
    nadv = np.where(active, n, 0)
    cells = np.where(valid & active[:, None], lst[:, :15], nC)
    adv[~active] = 0.0
    adv3[~active] = 0.0
    z15 = np.zeros((1, 15))
    d["nAdvCellsForEdge"] = np.concatenate([nadv, [0]]).astype(np.int32)
    d["advCellsForEdge"] = np.concatenate([cells, np.full((1, 15), nC)]).astype(np.int32)
    d["adv_coefs"] = np.concatenate([adv, z15])
    d["adv_coefs_3rd"] = np.concatenate([adv3, z15])


def init_block(d: dict, cfg: dict) -> dict:
    """Adds every row-M field to ``d`` (in place) and returns it."""
    nC, nE, nV = d["nCells"], d["nEdges"], d["nVertices"]
    compute_signs(d)
    d["invAreaCell"] = 1.0 / d["areaCell"]
    d["invDvEdge"] = 1.0 / d["dvEdge"]
    d["invDcEdge"] = 1.0 / d["dcEdge"]
    d["invAreaTriangle"] = 1.0 / d["areaTriangle"]
    adv_coef_compression(d)
    d["adv_coefs_3rd"] = cfg["config_coef_3rd_order"] * d["adv_coefs_3rd"]
    d["zb3_cell"] = cfg["config_coef_3rd_order"] * d["zb3_cell"]
    # mesh scaling
    c1, c2 = d["cellsOnEdge"][:, 0], d["cellsOnEdge"][:, 1]
    md = d["meshDensity"]
    s2 = np.ones(nE + 1)
    s4 = np.ones(nE + 1)
    if cfg["config_h_ScaleWithMesh"]:
        s2[:nE] = 1.0 / ((md[c1[:nE]] + md[c2[:nE]]) / 2.0) ** 0.25
        s4[:nE] = 1.0 / ((md[c1[:nE]] + md[c2[:nE]]) / 2.0) ** 0.75
    d["meshScalingDel2"], d["meshScalingDel4"] = s2, s4
    # Rayleigh damping profile for w
    zg = d["zgrid"]
    nz1 = zg.shape[1] - 1
    zt = zg[:, nz1:nz1 + 1]
    z = 0.5 * (zg[:, :-1] + zg[:, 1:])
    zd = cfg["config_zd"]
    with np.errstate(invalid="ignore", divide="ignore"):
        dss = cfg["config_xnutr"] * np.sin(0.5 * PII * (z - zd) / (zt - zd)) ** 2.0
    dss = np.where(z > zd, dss, 0.0) / md[:, None] ** 0.25
    dss[nC] = 0.0
    d["dss"] = dss
    d["specZoneMaskCell"] = np.zeros(nC + 1)
    d["specZoneMaskEdge"] = np.zeros(nE + 1)
    d["bdyMaskCell"] = np.zeros(nC + 1, dtype=np.int32)
    d["bdyMaskEdge"] = np.zeros(nE + 1, dtype=np.int32)
    # relaxation-zone scalings of a regional run (mpas_atm_core.F:1131-1146)
    d["meshScalingRegionalCell"] = np.ones(nC + 1)
    d["meshScalingRegionalEdge"] = np.ones(nE + 1)
    if cfg["config_h_ScaleWithMesh"]:
        d["meshScalingRegionalEdge"][:nE] = 1.0 / ((md[c1[:nE]] + md[c2[:nE]]) / 2.0) ** 0.25
        d["meshScalingRegionalCell"][:nC] = 1.0 / md[:nC] ** 0.25
    # mpas_atm_core.F:534-535: mpas_rbf_interp_initialize (-> mpas_initialize_vectors), mpas_init_reconstruct -> coeffs_reconstruct (owned cells)
    from .reconstruct import mpas_init_reconstruct
    mpas_init_reconstruct(d)
    return d
