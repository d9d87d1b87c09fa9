"""NetCDF classic (CDF-1/2/5) reader/writer and the MPAS init-file <-> block-dict mapping.
No NetCDF library exists in the image; CDF-1/2 files are cross-checked against scipy's reader/writer."""
import numpy as np
import pytest

from mpas_model_b200 import initfile, ncio


def _sample(rng):
    xt = np.zeros((2, 8), dtype="S1")
    xt[0] = np.frombuffer(b"2000-01-", dtype="S1"); xt[1] = np.frombuffer(b"2000-02-", dtype="S1")
    return {
        "xCell": ncio.Var(("nCells",), rng.standard_normal(7)),
        "cellsOnEdge": ncio.Var(("nCells", "TWO"), rng.integers(1, 8, (7, 2)).astype(np.int32)),
        "theta": ncio.Var(("Time", "nCells", "nVertLevels"), rng.standard_normal((2, 7, 3))),
        "qv": ncio.Var(("Time", "nCells", "nVertLevels"), rng.standard_normal((2, 7, 3)).astype(np.float32)),
        "cf1": ncio.Var((), np.asarray(2.0)),
        "xtime": ncio.Var(("Time", "StrLen"), xt),
    }


@pytest.mark.parametrize("version", [1, 2, 5])
def test_ncio_round_trip(tmp_path, version):
    rng = np.random.default_rng(version)
    dims = {"Time": 0, "nCells": 7, "nVertLevels": 3, "TWO": 2, "StrLen": 8}
    vars_ = _sample(rng)
    attrs = {"on_a_sphere": "YES", "sphere_radius": 6371229.0, "n": np.int32(5)}
    p = str(tmp_path / f"t{version}.nc")
    ncio.write(p, dims, attrs, vars_, version=version, unlimited="Time")
    assert open(p, "rb").read(4) == b"CDF" + bytes([version])
    d, a, v = ncio.read(p)
    assert d == {"Time": 2, "nCells": 7, "nVertLevels": 3, "TWO": 2, "StrLen": 8}
    assert a["on_a_sphere"] == "YES" and a["sphere_radius"] == 6371229.0 and a["n"] == 5
    for k, x in vars_.items():
        assert v[k].dims == x.dims and v[k].data.dtype == x.data.dtype and np.array_equal(v[k].data, x.data), k
    only = ncio.read(p, only={"theta"})[2]
    assert list(only) == ["theta"]


@pytest.mark.parametrize("version", [1, 2])
def test_ncio_against_scipy(tmp_path, version):
    from scipy.io import netcdf_file
    rng = np.random.default_rng(10 + version)
    vars_ = _sample(rng)
    p = str(tmp_path / "ours.nc")
    ncio.write(p, {"Time": 0, "nCells": 7, "nVertLevels": 3, "TWO": 2, "StrLen": 8}, {"sphere_radius": 6371229.0}, vars_,
               version=version, unlimited="Time")
    f = netcdf_file(p, "r", mmap=False)                         # scipy reads what we wrote
    for k in ("xCell", "cellsOnEdge", "theta", "qv"):
        assert np.array_equal(f.variables[k][:], vars_[k].data), k
    assert f.sphere_radius == 6371229.0 and f.variables["theta"].isrec and f.dimensions["Time"] is None
    f.close()
    q = str(tmp_path / "scipy.nc")                              # and we read what scipy wrote
    f = netcdf_file(q, "w", version=version)
    f.createDimension("Time", None); f.createDimension("n", 4)
    a = f.createVariable("a", "d", ("Time", "n")); a[0] = np.arange(4.0); a[1] = 2 * np.arange(4.0)
    b = f.createVariable("b", "i", ("n",)); b[:] = np.arange(4)
    c = f.createVariable("c", "f", ("Time",)); c[0] = 1.5; c[1] = 2.5
    f.title = "x"; f.close()
    d, at, v = ncio.read(q)
    assert d == {"Time": 2, "n": 4} and at == {"title": "x"}
    assert np.array_equal(v["a"].data, [[0, 1, 2, 3], [0, 2, 4, 6]]) and np.array_equal(v["b"].data, np.arange(4))
    assert np.array_equal(v["c"].data, [1.5, 2.5])


def test_ncio_rejects_other_files(tmp_path):
    p = tmp_path / "x.nc"
    p.write_bytes(b"\x89HDF\r\n\x1a\n" + b"\0" * 64)            # NetCDF-4/HDF5 is not a classic file
    with pytest.raises(ValueError):
        ncio.read(str(p))


def test_init_file_round_trip_and_step(tmp_path):
    """case -> x1.642.init.nc (CDF-5, 1-based, no garbage rows) -> block dict: every field the library consumes is
    recovered exactly on real elements, and one oracle step from the file equals one from the generated case bit for bit."""
    from mpas_model_b200.case import make_case
    from mpas_model_b200.fields import FIELDS
    from oracle.oracle import OracleDycore
    d, cfg = make_case(642, 10, num_scalars=2)
    p = str(tmp_path / "x1.642.init.nc")
    initfile.write_init_file(d, p)
    dims, attrs, v = ncio.read(p, only={"cellsOnEdge", "zb", "u", "qv", "tracer1"})
    assert dims["nCells"] == 642 and dims["Time"] == 1 and attrs["on_a_sphere"] == "YES"
    assert v["cellsOnEdge"].data.min() >= 1 and v["cellsOnEdge"].data.max() <= 642          # 1-based on disk
    assert v["zb"].dims == ("nEdges", "TWO", "nVertLevelsP1") and v["u"].dims == ("Time", "nEdges", "nVertLevels")
    d2, cfg2 = initfile.read_init_file(p, dt=cfg["config_dt"])
    assert cfg2 == cfg and d2["num_scalars"] == 2 and d2["index_qv"] == 0
    n_of = {"CELL": d["nCells"], "EDGE": d["nEdges"], "VERTEX": d["nVertices"]}
    for name, fd in FIELDS.items():
        if name in d and isinstance(d[name], np.ndarray):
            n = n_of.get(fd.loc)
            a, b = (d[name], d2[name]) if n is None else (d[name][:n], d2[name][:n])
            assert np.array_equal(a, b), name
    dt = cfg["config_dt"]
    o1, o2 = OracleDycore(d, cfg), OracleDycore(d2, cfg2)
    for o in (o1, o2):
        o.atm_init_coupled_diagnostics(); o.atm_init_solve_diagnostics(dt); o.atm_srk3(dt)
    for name in ("u", "w", "rho_zz", "theta_m", "scalars"):
        assert np.array_equal(o1.get_array(name, 2), o2.get_array(name, 2)), name


def test_compare_tool_against_a_restart_file(tmp_path):
    """tools/compare_with_reference_output.py, the one-command pin against a real MPAS run: here the 'reference' restart
    file is written from the oracle's own state after two steps, so the comparison must come out exactly zero."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location(
        "cmp_tool", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "compare_with_reference_output.py"))
    tool = importlib.util.module_from_spec(spec); spec.loader.exec_module(tool)
    init = str(tmp_path / "x1.642.init.nc")
    tool.main(["write-init", init, "--cells", "642", "--levels", "10"])
    d, cfg, b = tool.run_from_init(init, 2)
    nC, nE, nz = d["nCells"], d["nEdges"], d["nVertLevels"]
    fields = {"u": ncio.Var(("Time", "nEdges", "nVertLevels"), b.get_array("u", 1)[:nE][None]),
              "w": ncio.Var(("Time", "nCells", "nVertLevelsP1"), b.get_array("w", 1)[:nC][None]),
              "rho_zz": ncio.Var(("Time", "nCells", "nVertLevels"), b.get_array("rho_zz", 1)[:nC][None]),
              "theta_m": ncio.Var(("Time", "nCells", "nVertLevels"), b.get_array("theta_m", 1)[:nC][None]),
              "qv": ncio.Var(("Time", "nCells", "nVertLevels"), b.get_array("scalars", 1)[:nC, :, 0][None])}
    rst = str(tmp_path / "restart.nc")
    ncio.write(rst, {"Time": 0, "nCells": nC, "nEdges": nE, "nVertLevels": nz, "nVertLevelsP1": nz + 1}, {}, fields, version=5, unlimited="Time")
    assert tool.main(["compare", init, rst, "--steps", "2"]) == 0.0
    assert tool.main(["compare", init, rst, "--steps", "1"]) > 1e-6          # and it does notice a different state


@pytest.mark.gpu
def test_gpu_step_started_from_an_init_file(tmp_path):
    """Row f3 on the GPU: x1.642.init.nc (written in the reference's on-disk conventions) -> initfile.read_init_file (which ends
    with atm_mpas_init_block's derivations) -> the CUDA library through the C ABI -> two steps; compared with the oracle started
    from the generated case (not from the file): rel-L2 <= 1e-11 per step, and bit-identical between file start and case
    start on the GPU itself."""
    from mpas_model_b200.case import make_case
    from mpas_model_b200.dycore import Dycore
    from oracle.oracle import OracleDycore
    d, cfg = make_case(642, 10, num_scalars=2)
    p = str(tmp_path / "x1.642.init.nc")
    initfile.write_init_file(d, p)
    d2, cfg2 = initfile.read_init_file(p, dt=cfg["config_dt"])
    d3, cfg3 = initfile.read_init_file(p, dt=cfg["config_dt"], derive="library")   # raw mesh fields only: the library derives the rest
    assert "adv_coefs" not in d3 and "zb_cell" not in d3 and "coeffs_reconstruct" not in d3
    dt = cfg["config_dt"]
    o, g_file, g_case, g_lib = OracleDycore(d, cfg), Dycore(d2, cfg2), Dycore(d, cfg), Dycore(d3, cfg3)
    g_lib.atm_mpas_init_block(d3, cfg3)                              # mpasb_init_block (C++ in the library)
    for b in (o, g_file, g_case, g_lib):
        b.atm_init_coupled_diagnostics(); b.atm_init_solve_diagnostics(dt)
        for _ in range(2):
            b.atm_srk3(dt); b.mpas_pool_shift_time_levels()
    for name in ("u", "w", "rho_zz", "theta_m", "scalars"):
        a, r = g_file.get_array(name, 1), o.get_array(name, 1)
        assert np.array_equal(a, g_case.get_array(name, 1)), name
        assert np.array_equal(a, g_lib.get_array(name, 1)), name
        assert np.linalg.norm((a - r).ravel()) <= 2e-11 * np.linalg.norm(r.ravel()), name
    assert g_file.kernel_launch_count() > 0
    for b in (o, g_file, g_case, g_lib):
        b.close()
