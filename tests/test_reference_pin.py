"""The oracle pinned against the reference's OWN statements.

oracle/_ref/libtiref.so is the reference's mpas_atm_time_integration.F (every *_work routine of the step, the wrappers
that pick their arguments out of the pools, and the pool-based routines) transliterated statement by statement to C++
by oracle/f2cpp.py at build time, from the source as it lies under /root/reference -- nothing in it is restated by hand
(oracle/ref_harness.cpp only holds the pool plumbing and the one call atm_srk3 makes per routine).  The hand-written
oracle (oracle/dycore_oracle.cpp), which every GPU parity test compares against, must reproduce it BIT FOR BIT:
routine by routine on identical inputs, and over free-running steps.
"""
import numpy as np
import pytest

from tests.util import real_fields, srk3_stepwise, sync_all

ref = pytest.importorskip("oracle.ref")
if not ref.build():
    pytest.skip("oracle/_ref is not built (the reference tree is needed at build time)", allow_module_level=True)

STATE = ("u", "w", "rho_zz", "theta_m", "scalars")


def _differing(a, b):
    """Names of the real fields (every time level, scratch arrays included) that are not bit-identical."""
    return [f"{n}@{lev}" for n, lev in real_fields() if not np.array_equal(a.get_array(n, lev), b.get_array(n, lev), equal_nan=True)]


def _walk(d, cfg, precision="double"):
    from oracle.oracle import OracleDycore
    o, r = OracleDycore(d, cfg, precision=precision), ref.RefDycore(d, cfg, precision=precision)
    dt = cfg["config_dt"]
    for b in (o, r):
        b.atm_init_coupled_diagnostics(); b.atm_init_solve_diagnostics(dt)
    assert _differing(o, r) == []
    labels = []

    def after(label):
        bad = _differing(o, r)
        assert bad == [], (label, bad)
        labels.append(label)
        sync_all(o, r)

    srk3_stepwise([o, r], cfg, dt, after, reconstruct=False)
    log = r.exchange_log()
    o.close(); r.close()
    return labels, log


def test_every_routine_bit_identical_to_the_reference_source(tiny_case):
    d, cfg = tiny_case
    labels, log = _walk(d, cfg)
    assert len(labels) >= 70                                   # 3 x (vert_imp_coefs + 3 stages) + transport
    # the exchanges the reference's monotonic transport asks for (TI:4155, 4568): one scalars_old, one scale per scalar
    assert log == "dynamics:scalars_old;" + "dynamics:scale;" * d["num_scalars"]


def test_55_levels_two_scalars():
    from mpas_model_b200.case import make_case
    d, cfg = make_case(642, 55, num_scalars=2)
    _walk(d, cfg)


def test_irregular_mesh():
    """pentagons, hexagons and heptagons (maxEdges = 7, stencils of up to 12 cells): the select-case branches of the
    transport routines (TI:3677, 4367) take their `case default` arms."""
    from mpas_model_b200.case import make_case
    d, cfg = make_case(642, 10, num_scalars=2, jitter=0.2)
    assert d["nAdvCellsForEdge"].max() > 10
    _walk(d, cfg)


VARIANTS = {
    "order3_substeps4": dict(config_time_integration_order=3, config_number_of_sub_steps=4),
    "fixed_mixing": dict(config_horiz_mixing="2d_fixed", config_h_mom_eddy_visc2=1.0e4, config_h_theta_eddy_visc2=1.0e4,
                         config_h_mom_eddy_visc4=1.0e13, config_h_theta_eddy_visc4=1.0e13),
    "vertical_mixing": dict(config_v_mom_eddy_visc2=10.0, config_v_theta_eddy_visc2=10.0, config_mix_full=False),
    "rayleigh_u_and_cam_damping": dict(config_rayleigh_damp_u=True, config_number_rayleigh_damp_u_levels=4,
                                       config_mpas_cam_coef=2.0, config_number_cam_damping_levels=3),
    "no_apvm_not_monotonic": dict(config_apvm_upwinding=0.0, config_monotonic=False, config_epssm=0.2, config_smdiv=0.2),
    "coupled_transport": dict(config_split_dynamics_transport=False, config_number_of_sub_steps=6),
}


@pytest.mark.parametrize("variant", sorted(VARIANTS))
def test_namelist_variants(tiny_case, variant):
    d, cfg0 = tiny_case
    _walk(d, dict(cfg0, **VARIANTS[variant]))


def test_single_precision_restatement(tiny_case):
    """PRECISION=single: liboracle_sp.so against the transliteration compiled with RKIND = float (REF_SINGLE: every
    default-kind literal a float, as in the reference's single build)."""
    d, cfg = tiny_case
    _walk(d, cfg, precision="single")


def test_free_running_steps(small_case):
    """No synchronisation between routines: three steps of the transliterated reference (driven in atm_srk3's order)
    and of the oracle's own atm_srk3 end in bit-identical states -- the oracle's orchestration is pinned as well."""
    from oracle.oracle import OracleDycore
    d, cfg = small_case
    o, r = OracleDycore(d, cfg), ref.RefDycore(d, cfg)
    dt = cfg["config_dt"]
    for b in (o, r):
        b.atm_init_coupled_diagnostics(); b.atm_init_solve_diagnostics(dt)
    for _ in range(3):
        o.atm_srk3(dt)
        srk3_stepwise([r], cfg, dt, reconstruct=False)
        o.mpas_pool_shift_time_levels(); r.mpas_pool_shift_time_levels()
    for n in STATE:
        assert np.array_equal(o.get_array(n, 1), r.get_array(n, 1)), n
    assert float(np.abs(o.get_array("w", 1)).max()) > 0.0
    o.close(); r.close()


def test_threaded_reference_equals_serial(small_case):
    """The transliteration keeps the work routines' !$OMP BARRIER / MASTER directives; entered by 4 threads with the index
    ranges of mpas_atm_threading.F:100-111 (the reference's OpenMP build) it gives bit-identical results to one thread."""
    d, cfg = small_case
    r1, r4 = ref.RefDycore(d, cfg, threads=1), ref.RefDycore(d, cfg, threads=4)
    dt = cfg["config_dt"]
    for b in (r1, r4):
        b.atm_init_coupled_diagnostics(); b.atm_init_solve_diagnostics(dt)
        srk3_stepwise([b], cfg, dt, reconstruct=False)
    assert _differing(r1, r4) == []
    r1.close(); r4.close()


@pytest.mark.parametrize("jitter", [0.0, 0.2], ids=["icosahedral", "irregular"])
def test_init_block_against_the_reference_init_routines(jitter):
    """SURVEY.md §8 row M.  mpas_model_b200/init_block.py feeds BOTH the oracle and the CUDA path, so no parity test between
    those two can see an error in it.  Here its outputs are compared bit for bit with the reference's own init-time routines
    (mpas_atm_core.F: atm_compute_mesh_scaling :1091, atm_compute_signs :1151, atm_compute_damping_coefs :1241,
    atm_adv_coef_compression :1285, atm_couple_coef_3rd_order :1433), transliterated by oracle/f2cpp.py, on a mesh with a
    non-uniform meshDensity so that the del2/del4 scalings are not trivially one."""
    from mpas_model_b200 import init_block
    from mpas_model_b200.case import make_case
    d_raw, cfg = make_case(642, 10, num_scalars=1, derive=False, jitter=jitter)
    nC = d_raw["nCells"]
    d_raw["meshDensity"] = np.concatenate([0.4 + 0.6 * np.cos(d_raw["latCell"][:nC]) ** 2, [1.0]])
    d = dict(d_raw)
    init_block.init_block(d, cfg)
    derived_real = ("edgesOnCell_sign", "edgesOnVertex_sign", "zb_cell", "zb3_cell", "meshScalingDel2", "meshScalingDel4", "dss",
                    "adv_coefs", "adv_coefs_3rd")
    derived_int = ("kiteForCell", "advCellsForEdge", "nAdvCellsForEdge")
    raw = dict(d)
    for k in derived_real + derived_int:
        raw[k] = np.zeros_like(d[k])
    r = ref.RefDycore(raw, cfg)
    for routine in ("compute_mesh_scaling", "compute_signs", "compute_damping_coefs", "adv_coef_compression", "couple_coef_3rd_order"):
        r.k(routine)                                       # order of atm_mpas_init_block, mpas_atm_core.F:573-586
    assert float(np.abs(d["meshScalingDel2"] - 1.0).max()) > 0.05 and d["dss"].max() > 0.0
    for k in derived_real:
        got = r.a[(k, 1)].reshape(np.shape(d[k]))
        if k.startswith("meshScalingDel") or k == "dss":   # meshDensity ** 0.25, ** 0.75: numpy's vectorised pow vs libm's, one ulp apart
            assert np.allclose(got, d[k], rtol=5e-16, atol=0), k
        else:
            assert np.array_equal(got, d[k]), k
    nadv = d["nAdvCellsForEdge"]
    assert np.array_equal(r.a[("nAdvCellsForEdge", 1)], nadv)
    assert np.array_equal(r.a[("kiteForCell", 1)][:nC], d["kiteForCell"][:nC] + 1)           # the reference's arrays are 1-based
    used = np.arange(15)[None, :] < nadv[:, None]
    assert np.array_equal(r.a[("advCellsForEdge", 1)][used], d["advCellsForEdge"][used] + 1)
    r.close()


def test_constants_come_from_the_reference():
    """The physical constants inside the generated code are the parameter statements of src/framework/mpas_constants.F."""
    import os
    text = open(os.path.join(ref.REF_DIR, "ti_ref.inc")).read()
    for line in ("static const real gravity = RL(9.80616);", "static const real rgas = RL(287.0);",
                 "static const real cp = ((RL(7.0) * rgas) / RL(2.0));", "static const real p0 = RL(1.0e5);"):
        assert line in text, line
