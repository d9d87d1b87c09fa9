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
