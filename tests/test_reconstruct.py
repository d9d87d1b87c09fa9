"""CPU checks of the velocity reconstruction: the init-time coefficients
(mpas_init_reconstruct, mpas_vector_reconstruction.F:60-177 through the RBF routines of
mpas_rbf_interpolation.F) and the oracle's mpas_reconstruct / atm_compute_output_diagnostics."""
import numpy as np

from oracle.oracle import OracleDycore


def test_coefficients_reproduce_constant_fields(tiny_case):
    """The RBF system carries a constant basis (mpas_rbf_interpolation.F:1120-1127): a vector that is constant
    in the cell's tangent plane, sampled through the planar edge normals, is reconstructed exactly."""
    d, _ = tiny_case
    nC = d["nCells"]
    co, en, tp = d["coeffs_reconstruct"], d["edgeNormalVectors"], d["cellTangentPlane"]
    assert co.shape == (nC + 1, d["maxEdges"], 3) and not co[nC].any()
    w = np.array([0.7, -0.4])
    worst = 0.0
    for c in range(nC):
        n = d["nEdgesOnCell"][c]
        e = d["edgesOnCell"][c, :n]
        planar = np.stack([en[e] @ tp[c, 0], en[e] @ tp[c, 1]], axis=1)
        rec = (co[c, :n, :] * (planar @ w)[:, None]).sum(0)
        worst = max(worst, np.abs(rec - (w[0] * tp[c, 0] + w[1] * tp[c, 1])).max())
        assert not co[c, n:].any()
    assert worst < 1e-13
    # the reconstructed vector lies in the tangent plane
    rhat = d["localVerticalUnitVectors"][:nC]
    assert np.abs(np.einsum("cej,cj->ce", co[:nC], rhat)).max() < 1e-13


def test_oracle_reconstruct_and_output_diagnostics(small_case):
    d, cfg = small_case
    nC = d["nCells"]
    o = OracleDycore(d, cfg)
    o.atm_init_coupled_diagnostics()
    o.mpas_reconstruct(1, False)
    uz, um = o.get_array("uReconstructZonal")[:nC], o.get_array("uReconstructMeridional")[:nC]
    # JW case 2: a zonal jet of 35 m/s with a 1 m/s perturbation; the meridional wind is the perturbation's only
    assert 33.0 < uz.max() < 37.0 and np.abs(um).max() < 1.5
    x, y, z = (o.get_array("uReconstruct" + c)[:nC] for c in "XYZ")
    rhat = d["localVerticalUnitVectors"][:nC]
    radial = x * rhat[:, None, 0] + y * rhat[:, None, 1] + z * rhat[:, None, 2]
    assert np.abs(radial).max() < 1e-12                                 # no radial component
    assert np.allclose(uz ** 2 + um ** 2, x ** 2 + y ** 2 + z ** 2, rtol=1e-12, atol=1e-12)   # a rotation
    o.atm_compute_output_diagnostics(1)
    rvord = 461.6 / 287.0
    qv = o.get_array("scalars")[..., d["index_qv"]]
    assert np.array_equal(o.get_array("theta"), o.get_array("theta_m") / (1.0 + rvord * qv))
    assert np.array_equal(o.get_array("rho"), o.get_array("rho_zz") * o.get_array("zz"))
    assert np.array_equal(o.get_array("pressure"), o.get_array("pressure_base") + o.get_array("pressure_p"))
    # the step itself ends with mpas_reconstruct on time level 2 (TI:1606)
    dt = cfg["config_dt"]
    o.atm_init_solve_diagnostics(dt)
    o.atm_srk3(dt)
    after_step = o.get_array("uReconstructZonal").copy()
    o.mpas_reconstruct(2, False)
    assert np.array_equal(after_step, o.get_array("uReconstructZonal")) and not np.array_equal(after_step[:nC], uz)


def test_blocks_reconstruct_like_the_single_block(tiny_case):
    """Owned cells of a decomposed run carry the coefficients and values of the undecomposed mesh."""
    from mpas_model_b200 import decomp
    d, cfg = tiny_case
    part = decomp.partition_rcb(d, 3)
    blocks, _ = decomp.decompose_case(d, cfg, part)
    ref = OracleDycore(d, cfg)
    ref.mpas_reconstruct(1, False)
    for r, b in blocks.items():
        n = b["nCellsSolve"]
        gid = b["indexToCellID"][:n] - 1
        assert np.allclose(b["coeffs_reconstruct"][:n], d["coeffs_reconstruct"][gid], rtol=0, atol=1e-14)
        assert not b["coeffs_reconstruct"][n:].any()                       # owned cells only (includeHalos absent)
        o = OracleDycore(b, cfg, rank=r)
        o.mpas_reconstruct(1, False)
        for name in ("uReconstructZonal", "uReconstructMeridional"):
            assert np.allclose(o.get_array(name)[:n], ref.get_array(name)[gid], rtol=0, atol=1e-12), (r, name)
