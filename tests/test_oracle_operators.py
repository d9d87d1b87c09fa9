"""Analytic pins of the oracle's diagnostic operators (atm_compute_solve_diagnostics, TI:6337-6773, and mpas_reconstruct):
for a solid-body rotation the relative vorticity is 2 w0 sin(lat), the divergence vanishes, the kinetic energy is
U^2 / 2 and the tangential velocity is known in closed form.  Independent of the reference's text: a wrong sign or
orientation convention in the mesh, in edgesOnCell_sign / edgesOnVertex_sign or in the operators shows up here, and the
errors must shrink with the mesh spacing."""
import numpy as np
import pytest

from mpas_model_b200.case import make_case
from oracle.oracle import OracleDycore


def _errors(n_cells):
    d, cfg = make_case(n_cells, 10)
    nC, nE, nV = d["nCells"], d["nEdges"], d["nVertices"]
    a = d["sphere_radius"]
    w0 = 2.0 * np.pi / (12.0 * 86400.0)                               # one revolution in 12 days: U = 38.6 m/s at the equator
    xe = np.stack([d["xEdge"], d["yEdge"], d["zEdge"]], 1)[:nE]
    vel = np.cross(np.array([0.0, 0.0, w0]), xe)                      # V = w0 k x r
    en = d["edgeNormalVectors"][:nE]
    et = np.cross(xe / np.linalg.norm(xe, axis=1)[:, None], en)       # k x n: the direction of the tangential velocity v
    u = np.zeros_like(d["u"]); u[:nE] = np.einsum("ij,ij->i", vel, en)[:, None]
    o = OracleDycore(d, cfg)
    o.set_array("u", u, 1)
    o.set_array("rho_zz", np.ones(o.shape("rho_zz")), 1)
    o.atm_init_solve_diagnostics(cfg["config_dt"])                    # time level 1, rk_step absent: v is reconstructed too
    o.mpas_reconstruct(1, False)
    U = w0 * a
    zeta = 2.0 * w0 * np.sin(d["latVertex"][:nV])
    ke = 0.5 * (U * np.cos(d["latCell"][:nC])) ** 2
    vt = np.einsum("ij,ij->i", vel, et)
    def both(err, scale):                       # (max norm, rms) of an error field, relative to the field's scale
        return np.abs(err).max() / scale, float(np.sqrt(np.mean(err ** 2))) / scale

    return {
        "vorticity": both(o.get_array("vorticity")[:nV, 0] - zeta, 2.0 * w0),
        "divergence": both(o.get_array("divergence")[:nC, 0], 2.0 * w0),
        "ke": both(o.get_array("ke")[:nC, 0] - ke, 0.5 * U * U),
        "v": both(o.get_array("v")[:nE, 0] - vt, U),
        "zonal": both(o.get_array("uReconstructZonal")[:nC, 0] - U * np.cos(d["latCell"][:nC]), U),
        "meridional": both(o.get_array("uReconstructMeridional")[:nC, 0], U),
    }


def test_solid_body_rotation_diagnostics_converge():
    coarse, fine = _errors(642), _errors(2562)
    print(coarse, fine)
    # max norm: bounded (the kinetic-energy blend and the TRiSK tangential wind do not converge point-wise at the twelve
    # pentagons, a known property of the scheme); rms: shrinks with the spacing (factor 2 in h between the two meshes)
    for name, bound in (("vorticity", 0.02), ("divergence", 0.01), ("ke", 0.03), ("v", 0.02), ("zonal", 0.01), ("meridional", 0.01)):
        assert fine[name][0] < bound, (name, fine[name])
        assert fine[name][1] < 0.7 * coarse[name][1], (name, coarse[name], fine[name])


def _advect_bell(n_cells, hours=24.0):
    """A cosine bell carried by a solid-body rotation through the oracle's transport routines alone (advance_scalars on RK
    stages 1-2, the monotonic routine on stage 3; TI:1546-1592), compared with the analytically rotated bell."""
    d, cfg = make_case(n_cells, 10, num_scalars=2)
    nC, nE, nl = d["nCells"], d["nEdges"], d["nVertLevels"]
    w0 = 2.0 * np.pi / (12.0 * 86400.0)
    xe = np.stack([d["xEdge"], d["yEdge"], d["zEdge"]], 1)[:nE]
    en = d["edgeNormalVectors"][:nE]
    o = OracleDycore(d, cfg)
    ru = np.zeros(o.shape("ruAvg")); ru[:nE] = np.einsum("ij,ij->i", np.cross([0.0, 0.0, w0], xe), en)[:, None]
    o.set_array("ruAvg", ru)
    o.set_array("wwAvg", np.zeros(o.shape("wwAvg")))
    for lev in (1, 2):
        o.set_array("rho_zz", np.ones(o.shape("rho_zz")), lev)
    lat, lon = d["latCell"][:nC], d["lonCell"][:nC]

    def bell(lon0):
        r = np.arccos(np.clip(np.sin(0.3) * np.sin(lat) + np.cos(0.3) * np.cos(lat) * np.cos(lon - lon0), -1.0, 1.0))
        return np.where(r < 0.7, 0.5 * (1.0 + np.cos(np.pi * r / 0.7)), 0.0)

    q = np.zeros(o.shape("scalars")); q[:nC, :, 1] = bell(1.0)[:, None]
    o.set_array("scalars", q, 1)
    dt = 0.25 * d["nominalMinDc"] / (w0 * d["sphere_radius"])                  # Courant number 0.25 at the equator
    n_steps = int(round(hours * 3600.0 / dt))
    area = d["areaCell"][:nC]
    m0 = (q[:nC, 0, 1] * area).sum()
    for _ in range(n_steps):
        o.set_array("scalars", o.get_array("scalars", 1), 2)                     # atm_rk_integration_setup for the scalars
        o.k("advance_scalars", dt / 2.0, 1)
        o.k("advance_scalars", dt / 2.0, 2)
        o.k("advance_scalars_mono", float(dt))
        o.mpas_pool_shift_time_levels()
        o.set_array("rho_zz", np.ones(o.shape("rho_zz")), 1)                     # the shift swapped the two (identical) density levels
    got = o.get_array("scalars", 1)[:nC, 0, 1]
    want = bell(1.0 + w0 * n_steps * dt)
    return {"l2": float(np.sqrt(((got - want) ** 2 * area).sum() / (want ** 2 * area).sum())), "min": got.min(), "max": got.max(),
            "mass": abs((got * area).sum() - m0) / m0, "steps": n_steps}


def test_cosine_bell_in_solid_body_rotation():
    coarse, fine = _advect_bell(2562), _advect_bell(10242)
    print(coarse, fine)
    for r in (coarse, fine):
        assert r["mass"] < 1e-3 and r["min"] >= -1e-15 and r["max"] <= 1.0 + 1e-12        # conservative (to the discrete divergence of the wind), monotone
    assert fine["l2"] < 0.08 and fine["l2"] < 0.5 * coarse["l2"], (coarse, fine)          # converges: the 3rd-order flux coefficients are right
