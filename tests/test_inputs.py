"""CPU checks of the generated inputs: mesh conventions, JW
initial state, init-time derived fields."""
import numpy as np
import pytest

from mpas_model_b200 import mesh as meshmod


@pytest.fixture(scope="module")
def m():
    return meshmod.generate(642)


def test_euler_and_areas(m):
    nC, nE, nV = m["nCells"], m["nEdges"], m["nVertices"]
    assert nE == 3 * nC - 6 and nV == 2 * nC - 4
    assert (m["nEdgesOnCell"][:nC] == 5).sum() == 12 and (m["nEdgesOnCell"][:nC] == 6).sum() == nC - 12
    four_pi = 4 * np.pi
    assert abs(m["areaCell"][:nC].sum() - four_pi) < 1e-11
    assert abs(m["areaTriangle"][:nV].sum() - four_pi) < 1e-11
    assert np.allclose(m["kiteAreasOnVertex"][:nV].sum(1), m["areaTriangle"][:nV], rtol=1e-13)


def test_ring_orientation(m):
    """edge i of a cell joins verticesOnCell(i), verticesOnCell(i+1), counter-clockwise."""
    nC = m["nCells"]
    eoc, voc, coe, voe = m["edgesOnCell"], m["verticesOnCell"], m["cellsOnEdge"], m["verticesOnEdge"]
    for c in range(nC):
        n = m["nEdgesOnCell"][c]
        for i in range(n):
            e = eoc[c, i]
            cw, ccw = (voe[e, 0], voe[e, 1]) if coe[e, 0] == c else (voe[e, 1], voe[e, 0])
            assert voc[c, i] == cw and voc[c, (i + 1) % n] == ccw
            assert m["cellsOnCell"][c, i] == (coe[e, 1] if coe[e, 0] == c else coe[e, 0])


def _edge_frames(m):
    nE = m["nEdges"]
    P = np.stack([m["xEdge"], m["yEdge"], m["zEdge"]], 1)[:nE]
    pc = np.stack([m["xCell"], m["yCell"], m["zCell"]], 1)
    coe = m["cellsOnEdge"][:nE]
    n = pc[coe[:, 1]] - pc[coe[:, 0]]
    n -= (n * P).sum(1)[:, None] * P
    n /= np.linalg.norm(n, axis=1)[:, None]
    return P, n, np.cross(P, n)


def test_operator_identities(m):
    nC, nE, nV = m["nCells"], m["nEdges"], m["nVertices"]
    coe, voe, eov = m["cellsOnEdge"], m["verticesOnEdge"], m["edgesOnVertex"][:nV]
    phi = np.sin(3 * m["latCell"]) * np.cos(2 * m["lonCell"])
    g = np.concatenate([(phi[coe[:nE, 1]] - phi[coe[:nE, 0]]) / m["dcEdge"][:nE], [0.0]])
    dc = np.concatenate([m["dcEdge"][:nE], [0.0]])
    sv = np.where(voe[eov, 1] == np.arange(nV)[:, None], 1.0, -1.0)
    curl = (sv * dc[eov] * g[eov]).sum(1) / m["areaTriangle"][:nV]
    assert np.abs(curl).max() < 1e-11                                   # curl(grad) = 0


def test_trisk_weights(m):
    """Tangential reconstruction of a solid-body rotation and the TRiSK antisymmetry."""
    nE = m["nEdges"]
    P, n, t = _edge_frames(m)
    pv = np.stack([m["xVertex"], m["yVertex"], m["zVertex"]], 1)
    voe = m["verticesOnEdge"][:nE]
    assert (((pv[voe[:, 1]] - pv[voe[:, 0]]) * t).sum(1) > 0).all()      # vertex 1 -> 2 is k x n
    U = np.cross(np.array([0.3, -0.2, 1.0]), P)
    u = np.concatenate([(U * n).sum(1), [0.0]])
    v = (m["weightsOnEdge"][:nE] * u[m["edgesOnEdge"][:nE]]).sum(1)
    vt = (U * t).sum(1)
    assert np.abs(v - vt).max() / np.abs(vt).max() < 0.03
    east = np.cross([0, 0, 1.0], P); east /= np.linalg.norm(east, axis=1)[:, None]
    assert np.abs(np.cos(m["angleEdge"][:nE]) - (n * east).sum(1)).max() < 1e-12
    W = {}
    for e in range(nE):
        for j in range(m["nEdgesOnEdge"][e]):
            W[(e, int(m["edgesOnEdge"][e, j]))] = m["weightsOnEdge"][e, j]
    worst = max(abs(w * m["dcEdge"][e] / m["dvEdge"][e2] + W[(e2, e)] * m["dcEdge"][e2] / m["dvEdge"][e]) for (e, e2), w in W.items())
    assert worst < 1e-13


def test_jw_state_and_derived_fields(tiny_case):
    d, cfg = tiny_case
    nC, nE = d["nCells"], d["nEdges"]
    assert np.allclose(d["surface_pressure"][:nC], 1.0e5, rtol=1e-9)
    assert 30.0 < np.abs(d["u"][:nE]).max() < 40.0 and np.abs(d["w"][:nC]).max() < 1e-2
    assert d["theta"][:nC].min() > 200 and d["rho"][:nC].min() > 0
    # advection coefficients: consistency (sum = dvEdge) and upwind part sums to zero
    assert np.abs(d["adv_coefs"][:nE].sum(1) / d["dvEdge"][:nE] - 1).max() < 1e-12
    assert np.abs(d["adv_coefs_3rd"][:nE].sum(1)).max() / d["dvEdge"][:nE].mean() < 1e-12
    assert set(np.unique(d["nAdvCellsForEdge"][:nE])) <= {9, 10}
    # second-derivative stencil annihilates constants; signs and kites are consistent
    assert np.abs(d["deriv_two"][:nE].sum(2)).max() * d["dcEdge"][:nE].mean() ** 2 < 1e-10
    k = d["kiteForCell"][:nC]
    voc = d["verticesOnCell"][:nC]
    live = np.arange(d["maxEdges"])[None, :] < d["nEdgesOnCell"][:nC, None]
    assert (d["cellsOnVertex"][voc, k][live] == np.broadcast_to(np.arange(nC)[:, None], voc.shape)[live]).all()
    kite_sum = np.where(live, d["kiteAreasOnVertex"][voc, k], 0.0).sum(1)
    assert np.allclose(kite_sum, d["areaCell"][:nC], rtol=1e-12)
