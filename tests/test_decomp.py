"""Decomposition, exchange lists and the N-block oracle (CPU).

No reference partition files or halo lists ship with the reference, so the restatement in
mpas_model_b200/decomp.py is checked against (i) the orderings the reference's bootstrap code
produces by construction, (ii) the reference's own halo known-answer pattern
(src/core_test/mpas_halo_testing.F:153-190: after an exchange every halo slot holds its owner's
global id) and (iii) bit equality of a 4-block run with the single-block run."""
import numpy as np
import pytest

from mpas_model_b200 import decomp
from oracle import oracle as orc

KINDS = (("cells", 0), ("edges", 1), ("vertices", 2))


@pytest.fixture(scope="module")
def parts(small_case):
    d, cfg = small_case
    part = decomp.partition_rcb(d, 4)
    blocks, ex = decomp.decompose_case(d, cfg, part)
    return d, cfg, part, blocks, ex


def test_partition_file_roundtrip(tmp_path, parts):
    d, cfg, part, blocks, ex = parts
    f = tmp_path / "x1.2562.graph.info.part.4"
    decomp.write_partition_file(str(f), part)
    assert np.array_equal(decomp.read_partition_file(str(f)), part)
    counts = np.bincount(part)
    assert len(counts) == 4 and counts.max() - counts.min() <= 2


def test_local_orderings(parts):
    d, cfg, part, blocks, ex = parts
    nC = d["nCells"]
    for r, b in blocks.items():
        L = b["lists"]
        cb = L["cell_bounds"]
        cells = L["cells"]
        assert (part[cells[:cb[0]]] == r).all() and (part[cells[cb[0]:]] != r).all()
        for lo, hi in ((0, cb[0]), (cb[0], cb[1]), (cb[1], cb[2])):
            assert (np.diff(cells[lo:hi]) > 0).all()                      # each layer ascending in global id
        # halo 1 = exactly the cells adjacent to owned cells; halo 2 = adjacent to halo 1
        owned = set(cells[:cb[0]].tolist())
        ring1 = {int(c) for o in owned for c in d["cellsOnCell"][o, :d["nEdgesOnCell"][o]]} - owned
        assert ring1 == set(cells[cb[0]:cb[1]].tolist())
        eb, edges = L["edge_bounds"], L["edges"]
        own_e = edges[:eb[0]]
        assert (part[d["cellsOnEdge"][own_e, 0]] == r).all()               # owned iff cellsOnEdge(1) is owned
        e_of_owned = {int(e) for o in owned for e in d["edgesOnCell"][o, :d["nEdgesOnCell"][o]]}
        assert e_of_owned == set(edges[:eb[1]].tolist())                   # owned + layer 1 = all edges of owned cells
        assert len(set(edges.tolist())) == len(edges)
        vb, verts = L["vertice_bounds"], L["vertices"]
        assert (part[d["cellsOnVertex"][verts[:vb[0]], 0]] == r).all()
        # local connectivity round-trips to global ids
        gid = np.concatenate([cells, [nC]])
        loc = b["cellsOnEdge"][:b["nEdges"]]
        ge = d["cellsOnEdge"][edges]
        inside = loc < b["nCells"]
        assert (gid[loc][inside] == ge[inside]).all()
    # every global cell / edge / vertex is owned exactly once
    for kind, key in (("cells", "cell_bounds"), ("edges", "edge_bounds"), ("vertices", "vertice_bounds")):
        owned_all = np.concatenate([b["lists"][kind][:b["lists"][key][0]] for b in blocks.values()])
        assert len(owned_all) == len(set(owned_all.tolist())) == {"cells": d["nCells"], "edges": d["nEdges"], "vertices": d["nVertices"]}[kind]


def test_exchange_lists_are_sorted_and_symmetric(parts):
    d, cfg, part, blocks, ex = parts
    for r in blocks:
        for kind, _ in KINDS:
            e = ex[r][kind]
            gid = blocks[r]["lists"][kind]
            for qi, q in enumerate(e["neighbors"]):
                eq = ex[q][kind]
                ri = eq["neighbors"].index(r)
                for l in range(e["n_layers"]):
                    s = e["send"][qi][l]
                    assert (np.diff(gid[s]) > 0).all()                    # ascending global id (DM:1893-1913)
                    rq = eq["recv"][ri][l]
                    assert len(s) == len(rq)
                    assert np.array_equal(gid[s], blocks[q]["lists"][kind][rq])     # same elements, same order
                    pos, loc = eq["recv_ref"][ri][l]
                    assert (np.diff(loc) > 0).all() and sorted(pos.tolist()) == list(range(len(pos)))


def _mk_oracles(blocks, ex, cfg):
    os_ = []
    for r in sorted(blocks):
        o = orc.OracleDycore(blocks[r], cfg, rank=r)
        for kind, k in KINDS:
            o.set_halo_lists(k, ex[r][kind])
        os_.append(o)
    return os_


def test_halo_known_answer(parts):
    """mpas_halo_testing.F:153-190 pattern: owned = global id, halo = -1, exchange, halo == owner's id."""
    d, cfg, part, blocks, ex = parts
    os_ = _mk_oracles(blocks, ex, cfg)
    nl = d["nVertLevels"]
    for o, r in zip(os_, sorted(blocks)):
        b = blocks[r]
        for name, ids, nsolve, lev in (("theta_m", b["indexToCellID"], b["nCellsSolve"], 1), ("u", b["indexToEdgeID"], b["nEdgesSolve"], 1)):
            a = np.repeat(ids.astype(np.float64)[:, None], nl, axis=1) + np.arange(nl)[None, :] * 1e-3
            a[nsolve:] = -1.0
            o.set_array(name, a, lev)
    orc.exchange(os_, "dynamics:theta_m,scalars,pressure_p,rtheta_p")     # cells, layers 1-2
    orc.exchange(os_, "initialization:u")                                 # edges, layers 1-3
    for o, r in zip(os_, sorted(blocks)):
        b = blocks[r]
        for name, ids, n in (("theta_m", b["indexToCellID"], b["nCells"]), ("u", b["indexToEdgeID"], b["nEdges"])):
            a = o.get_array(name)
            want = np.repeat(ids.astype(np.float64)[:, None], nl, axis=1) + np.arange(nl)[None, :] * 1e-3
            assert np.array_equal(a[:n], want[:n]), (r, name)


COUPLED = dict(config_split_dynamics_transport=False, config_number_of_sub_steps=6)     # scalars inside the dynamics RK loop


@pytest.mark.parametrize("overrides", [{}, COUPLED], ids=["split_transport", "coupled_transport"])
def test_four_blocks_equal_one_block_bit_for_bit(parts, overrides):
    d, cfg, part, blocks, ex = parts
    cfg = dict(cfg, **overrides)
    dt = cfg["config_dt"]
    one = orc.OracleDycore(d, cfg)
    one.atm_init_coupled_diagnostics(); one.atm_init_solve_diagnostics(dt)
    os_ = _mk_oracles(blocks, ex, cfg)
    orc.exchange(os_, "initialization:u")                                 # mpas_atm_core.F:250
    for o in os_:
        o.atm_init_coupled_diagnostics(); o.atm_init_solve_diagnostics(dt)
    orc.exchange(os_, "initialization:pv_edge,ru,rw")                     # mpas_atm_core.F:288
    for _ in range(2):
        one.atm_srk3(dt); one.mpas_pool_shift_time_levels()
        orc.step(os_, dt)
        for o in os_:
            o.mpas_pool_shift_time_levels()
    for o, r in zip(os_, sorted(blocks)):
        L = blocks[r]["lists"]
        for name, kind, key in (("u", "edges", "edge_bounds"), ("w", "cells", "cell_bounds"), ("rho_zz", "cells", "cell_bounds"),
                                ("theta_m", "cells", "cell_bounds"), ("scalars", "cells", "cell_bounds")):
            n = L[key][0]
            assert np.array_equal(o.get_array(name)[:n], one.get_array(name)[L[kind][:n]]), (r, name)


def test_irregular_mesh_blocks_equal_one_block():
    """Variable cell degree (maxEdges = 7) through the decomposition: three blocks = one block, bit for bit."""
    from mpas_model_b200.case import make_case
    d, cfg = make_case(2562, 10, num_scalars=1, jitter=0.2)
    assert d["maxEdges"] == 7
    part = decomp.partition_rcb(d, 3)
    blocks, ex = decomp.decompose_case(d, cfg, part)
    dt = cfg["config_dt"]
    one = orc.OracleDycore(d, cfg)
    one.atm_init_coupled_diagnostics(); one.atm_init_solve_diagnostics(dt)
    os_ = _mk_oracles(blocks, ex, cfg)
    orc.exchange(os_, "initialization:u")
    for o in os_:
        o.atm_init_coupled_diagnostics(); o.atm_init_solve_diagnostics(dt)
    orc.exchange(os_, "initialization:pv_edge,ru,rw")
    one.atm_srk3(dt); orc.step(os_, dt)
    for o, r in zip(os_, sorted(blocks)):
        L = blocks[r]["lists"]
        for name, kind, key in (("u", "edges", "edge_bounds"), ("w", "cells", "cell_bounds"), ("theta_m", "cells", "cell_bounds"),
                                ("scalars", "cells", "cell_bounds")):
            n = L[key][0]
            assert np.array_equal(o.get_array(name, 2)[:n], one.get_array(name, 2)[L[kind][:n]]), (r, name)


@pytest.mark.parametrize("seed", [1, 2])
def test_arbitrary_partition_vectors(tiny_case, seed):
    """The decomposition takes any cell -> block vector (the reference reads METIS files, mpas_block_decomp.F:101-137), not only
    the coordinate bisection used for the benchmarks: blocks made of scattered patches, with every other block as neighbour,
    still give the single-block result bit for bit, and the halo known-answer pattern holds."""
    d, cfg = tiny_case
    rng = np.random.default_rng(seed)
    nC = d["nCells"]
    # patches: a few random seed cells per block, every cell joins the block of its nearest seed
    seeds = rng.choice(nC, size=9, replace=False)
    xyz = np.stack([d["xCell"][:nC], d["yCell"][:nC], d["zCell"][:nC]], 1)
    part = (np.argmax(xyz @ xyz[seeds].T, axis=1) % 3).astype(np.int64)
    assert len(np.unique(part)) == 3
    blocks, ex = decomp.decompose_case(d, cfg, part)
    os_ = _mk_oracles(blocks, ex, cfg)
    nl = d["nVertLevels"]
    for o, r in zip(os_, sorted(blocks)):
        b = blocks[r]
        a = np.repeat(b["indexToCellID"].astype(np.float64)[:, None], nl, axis=1)
        a[b["nCellsSolve"]:] = -1.0
        o.set_array("theta_m", a, 1)
    orc.exchange(os_, "dynamics:theta_m,scalars,pressure_p,rtheta_p")
    for o, r in zip(os_, sorted(blocks)):
        b = blocks[r]
        want = np.repeat(b["indexToCellID"].astype(np.float64)[:, None], nl, axis=1)
        assert np.array_equal(o.get_array("theta_m")[: b["nCells"]], want[: b["nCells"]]), r
        o.load_block(b)
    dt = cfg["config_dt"]
    one = orc.OracleDycore(d, cfg)
    one.atm_init_coupled_diagnostics(); one.atm_init_solve_diagnostics(dt)
    orc.exchange(os_, "initialization:u")
    for o in os_:
        o.atm_init_coupled_diagnostics(); o.atm_init_solve_diagnostics(dt)
    orc.exchange(os_, "initialization:pv_edge,ru,rw")
    one.atm_srk3(dt); orc.step(os_, dt)
    for o, r in zip(os_, sorted(blocks)):
        L = blocks[r]["lists"]
        for name, kind, key in (("u", "edges", "edge_bounds"), ("w", "cells", "cell_bounds"), ("rho_zz", "cells", "cell_bounds"),
                                ("theta_m", "cells", "cell_bounds"), ("scalars", "cells", "cell_bounds")):
            n = L[key][0]
            assert np.array_equal(o.get_array(name, 2)[:n], one.get_array(name, 2)[L[kind][:n]]), (r, name)
