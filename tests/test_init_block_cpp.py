"""atm_mpas_init_block's mesh part in the library (mpasb_init_block / mpasb_init_block_host, csrc/init_block_host.inl;
SURVEY.md §8 rows M and f3).  The C++ host routine is checked on the CPU against the numpy restatement every other test
uses (bit for bit) and against the reference's own init routines as transliterated by oracle/f2cpp.py; on the GPU a handle
whose derived fields come from mpasb_init_block steps bit-identically to one that was handed the numpy-derived fields."""
import ctypes as C

import numpy as np
import pytest

POW_BASED = ("meshScalingDel2", "meshScalingDel4", "meshScalingRegionalCell", "meshScalingRegionalEdge", "dss")


def _nonuniform(d):
    nC = d["nCells"]
    d["meshDensity"] = np.concatenate([0.4 + 0.6 * np.cos(d["latCell"][:nC]) ** 2, [1.0]])
    return d


def _same(name, got, want):
    got = got.reshape(np.shape(want))
    if name in POW_BASED:            # meshDensity ** 0.25 / 0.75: numpy's vectorised pow and libm's are one ulp apart
        return np.allclose(got, want, rtol=5e-16, atol=0)
    return np.array_equal(got, want)


@pytest.mark.parametrize("jitter", [0.0, 0.2], ids=["icosahedral", "irregular"])
def test_library_init_block_equals_the_numpy_restatement(jitter):
    """... and mpas_init_reconstruct's coeffs_reconstruct (reconstruct.py, which init_block.init_block ends with)."""
    from mpas_model_b200 import decomp, init_block
    from mpas_model_b200.case import make_case
    from mpas_model_b200.dycore import INIT_BLOCK_OUT_INT, INIT_BLOCK_OUT_REAL, init_block_host
    d, cfg = make_case(642, 10, num_scalars=1, derive=False, jitter=jitter)
    _nonuniform(d)
    init_block.init_block(d, cfg)
    got = init_block_host(d, cfg)
    for n in INIT_BLOCK_OUT_REAL + INIT_BLOCK_OUT_INT + ("coeffs_reconstruct",):       # (the last: mpas_init_reconstruct, from the coordinates)
        assert _same(n, got[n], d[n]), n
    assert float(np.abs(got["meshScalingDel2"] - 1.0).max()) > 0.05 and got["dss"].max() > 0.0 and np.abs(got["coeffs_reconstruct"]).max() > 0.1
    # one block of a decomposition: connectivity pointing at the garbage slot, edges without an owned cell
    blocks, _ = decomp.decompose_case(d, cfg, decomp.partition_rcb(d, 4))
    for r in (0, 3):
        got = init_block_host(blocks[r], cfg)
        for n in INIT_BLOCK_OUT_REAL + INIT_BLOCK_OUT_INT + ("coeffs_reconstruct",):
            assert _same(n, got[n], blocks[r][n]), (r, n)


def test_library_init_block_equals_the_reference_init_routines():
    """Bit for bit against mpas_atm_core.F:1091-1452 as transliterated from the reference's source (pow included: both are libm)."""
    ref = pytest.importorskip("oracle.ref")
    if not ref.build():
        pytest.skip("oracle/_ref is not built (the reference tree is needed at build time)")
    from mpas_model_b200 import init_block
    from mpas_model_b200.case import make_case
    from mpas_model_b200.dycore import init_block_host
    d, cfg = make_case(642, 10, num_scalars=1, derive=False, jitter=0.2)
    _nonuniform(d)
    got = init_block_host(d, cfg)
    full = dict(d)
    init_block.init_block(full, cfg)                       # (shapes of the derived fields; values are zeroed below)
    derived_real = ("edgesOnCell_sign", "edgesOnVertex_sign", "zb_cell", "zb3_cell", "meshScalingDel2", "meshScalingDel4", "dss",
                    "adv_coefs", "adv_coefs_3rd")
    raw = dict(full)
    for k in derived_real + ("kiteForCell", "advCellsForEdge", "nAdvCellsForEdge"):
        raw[k] = np.zeros_like(full[k])
    r = ref.RefDycore(raw, cfg)
    for routine in ("compute_mesh_scaling", "compute_signs", "compute_damping_coefs", "adv_coef_compression", "couple_coef_3rd_order"):
        r.k(routine)                                       # order of atm_mpas_init_block, mpas_atm_core.F:573-586
    for k in derived_real:
        assert np.array_equal(r.a[(k, 1)].reshape(got[k].shape), got[k]), k
    nC, nadv = d["nCells"], got["nAdvCellsForEdge"]
    assert np.array_equal(r.a[("nAdvCellsForEdge", 1)], nadv)
    assert np.array_equal(r.a[("kiteForCell", 1)][:nC], got["kiteForCell"][:nC] + 1)
    used = np.arange(15)[None, :] < nadv[:, None]
    assert np.array_equal(r.a[("advCellsForEdge", 1)][used], got["advCellsForEdge"][used] + 1)
    r.close()


def test_single_precision_library_and_error_returns(tiny_case):
    from mpas_model_b200 import dycore
    d, cfg = tiny_case
    got = dycore.init_block_host(d, cfg, precision="single")
    for n in dycore.INIT_BLOCK_OUT_REAL:
        assert got[n].dtype == np.float32
        want = np.asarray(d[n])              # (fp32 sums of cancelling deriv_two terms: bounded relative to the field's magnitude)
        assert np.allclose(got[n].reshape(want.shape), want, rtol=2e-6, atol=2e-6 * np.abs(want).max()), n
    for n in dycore.INIT_BLOCK_OUT_INT:
        assert np.array_equal(got[n].reshape(np.shape(d[n])), d[n]), n
    # a missing input and an output the routine does not derive are refused (1), a null table is a usage error (2)
    lib = dycore._load_lib()
    dims, config = dycore.make_dims(d), dycore.make_config(cfg, d)
    keep, k, names, ptrs = dycore._init_block_inputs(d, np.float64)
    args = (C.byref(dims), C.byref(config), C.c_int(1), C.c_double(cfg["config_zd"]), C.c_double(cfg["config_xnutr"]))
    assert lib.mpasb_init_block_host(*args, C.c_int(5), names, ptrs, C.c_int(0), None, None) == 1             # only the first five inputs
    raw = {n: v for n, v in d.items() if n not in dycore.INIT_BLOCK_IN_COORDS}                                  # without coordinates: no coeffs_reconstruct
    keep2, k2, names2, ptrs2 = dycore._init_block_inputs(raw, np.float64)
    co = np.zeros((d["nCells"] + 1, d["maxEdges"], 3))
    assert lib.mpasb_init_block_host(*args, C.c_int(k2), names2, ptrs2, C.c_int(1), (C.c_char_p * 1)(b"coeffs_reconstruct"),
                                     (C.c_void_p * 1)(co.ctypes.data)) == 1
    buf = np.zeros(4)
    onames, optrs = (C.c_char_p * 1)(b"theta_m"), (C.c_void_p * 1)(buf.ctypes.data)
    assert lib.mpasb_init_block_host(*args, C.c_int(k), names, ptrs, C.c_int(1), onames, optrs) == 1
    assert lib.mpasb_init_block_host(*args, C.c_int(k), None, None, C.c_int(0), None, None) == 2


@pytest.mark.gpu
def test_gpu_handle_derives_its_own_mesh_fields(small_case):
    """mpasb_init_block on a handle that was given only the raw mesh fields: the derived real fields read back equal the numpy
    ones, and two steps are bit-identical to a handle that was handed the numpy-derived fields."""
    from mpas_model_b200.dycore import INIT_BLOCK_OUT_INT, INIT_BLOCK_OUT_REAL, Dycore
    d, cfg = small_case
    raw = {k: v for k, v in d.items() if k not in INIT_BLOCK_OUT_REAL + INIT_BLOCK_OUT_INT + ("coeffs_reconstruct",)}
    dt = cfg["config_dt"]
    g_lib, g_np = Dycore(raw, cfg), Dycore(d, cfg)
    g_lib.atm_mpas_init_block(d, cfg)
    for n in INIT_BLOCK_OUT_REAL + ("coeffs_reconstruct",):
        assert _same(n, g_lib.get_array(n, 1), np.asarray(d[n])), n
    for b in (g_lib, g_np):
        b.atm_init_coupled_diagnostics(); b.atm_init_solve_diagnostics(dt)
        for _ in range(2):
            b.atm_srk3(dt); b.mpas_pool_shift_time_levels()
    for name in ("u", "w", "rho_zz", "theta_m", "scalars", "uReconstructZonal", "uReconstructMeridional"):
        assert np.array_equal(g_lib.get_array(name, 1), g_np.get_array(name, 1)), name
    g_lib.close(); g_np.close()
