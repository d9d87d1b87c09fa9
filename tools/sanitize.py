"""Small driver for compute-sanitizer: steps of every code path on a tiny mesh -- default (relaxed arithmetic, persistent
kernels, alternating sweep, programmatic dependent launch), strict arithmetic, generic kernels, TMA-staged flux and split
tendency kernels, Coriolis partial sums, coupled transport, the regional path (lateral boundary conditions), batched
asynchronous transfers, async summary, reconstruct / output diagnostics.
  compute-sanitizer --tool memcheck python tools/sanitize.py        compute-sanitizer --tool racecheck python tools/sanitize.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mpas_model_b200.case import make_case, make_regional
from mpas_model_b200.dycore import Dycore

SWITCHES = ("MPASB_GENERIC_KERNELS", "MPASB_TMA_FLUX", "MPASB_SPLIT_CELL_F", "MPASB_STRICT", "MPASB_COR", "MPASB_SNAKE", "MPASB_PDL", "MPASB_AC9",
            "MPASB_MONO_BATCH")
# `python tools/sanitize.py new`: only the paths added in the last session of round 2 (the default now stages the flux weights
# with cp.async.bulk + mbarrier and runs the monotonic transport batched over scalars; the TMA-pipelined column solve; the
# scalar-at-a-time transport; a block of a decomposition, i.e. halo columns, with its mesh fields derived by mpasb_init_block)
ONLY_NEW = len(sys.argv) > 1 and sys.argv[1] == "new"
variants = [("default", {}, {}, False),
            ("k9: TMA-pipelined column solve", {"MPASB_AC9": "1"}, {}, False),
            ("monotonic transport one scalar at a time", {"MPASB_MONO_BATCH": "0"}, {}, False),
            ("block 1 of 4, mesh fields by mpasb_init_block", {}, {}, "block"),
            ("strict", {"MPASB_STRICT": "1"}, {}, False),
            ("generic", {"MPASB_GENERIC_KERNELS": "1"}, {}, False),
            ("tma flux + split cell_f (strict path)", {"MPASB_STRICT": "1", "MPASB_TMA_FLUX": "1", "MPASB_SPLIT_CELL_F": "1"}, {}, False),
            ("coriolis partial sums, no snake, no pdl", {"MPASB_COR": "1", "MPASB_SNAKE": "0", "MPASB_PDL": "0"}, {}, False),
            ("coupled transport, order 3", {}, dict(config_split_dynamics_transport=False, config_number_of_sub_steps=6, config_time_integration_order=3), False),
            ("regional", {}, {}, True)]
if ONLY_NEW:
    variants = variants[:4]
for label, env, over, regional in variants:
    for k in SWITCHES:
        os.environ.pop(k, None)
    os.environ.update(env)
    d, cfg = make_case(642, 10, num_scalars=2, **over)
    t_end = 0.0
    block = regional == "block"
    regional = regional is True
    if regional:
        d, cfg, t_end = make_regional(d, cfg)
    if block:
        from mpas_model_b200 import decomp
        from mpas_model_b200.dycore import INIT_BLOCK_OUT_INT, INIT_BLOCK_OUT_REAL
        d = decomp.decompose_case(d, cfg, decomp.partition_rcb(d, 4))[0][1]
        full = d
        d = {k: v for k, v in d.items() if k not in INIT_BLOCK_OUT_REAL + INIT_BLOCK_OUT_INT}
    g = Dycore(d, cfg)
    if block:
        g.atm_mpas_init_block(full, cfg)
    dt = cfg["config_dt"]
    g.atm_init_coupled_diagnostics(); g.atm_init_solve_diagnostics(dt)
    for step in range(2):
        if regional:
            g.set_lbc_time(t_end - step * dt)
        g.atm_srk3(dt); g.summarize_timestep_async(); g.mpas_pool_shift_time_levels()
        print(label, g.summarize_timestep_fetch()[0][:4], flush=True)
    # batched asynchronous transfers around a step
    names = ("u", "w", "rho_zz", "theta_m", "scalars")
    host = {n: np.ascontiguousarray(g.get_array(n, 1)) for n in names}
    out = {n: np.empty_like(host[n]) for n in names}
    g.set_fields_async([(n, 1, host[n]) for n in names])
    if regional:
        g.set_lbc_time(t_end - 2 * dt)
    g.atm_srk3(dt)
    g.get_fields_async([(n, 2, out[n]) for n in names])
    g.wait_downloads(0)
    assert all(np.isfinite(out[n]).all() for n in names), label
    g.mpas_reconstruct(1, True); g.atm_compute_output_diagnostics(1); g.synchronize()
    g.close()
print("done")
