"""Small driver for compute-sanitizer: one step of every code path on a tiny mesh (column-warp kernels, generic kernels,
TMA-staged flux, split tendency kernels, coupled transport, async summary, reconstruct/output diagnostics)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mpas_model_b200.case import make_case
from mpas_model_b200.dycore import Dycore

variants = [({}, {}), ({"MPASB_GENERIC_KERNELS": "1"}, {}), ({"MPASB_TMA_FLUX": "1", "MPASB_SPLIT_CELL_F": "1"}, {}),
            ({}, dict(config_split_dynamics_transport=False, config_number_of_sub_steps=6, config_time_integration_order=3))]
for env, over in variants:
    for k in ("MPASB_GENERIC_KERNELS", "MPASB_TMA_FLUX", "MPASB_SPLIT_CELL_F"):
        os.environ.pop(k, None)
    os.environ.update(env)
    d, cfg = make_case(642, 10, num_scalars=2, **over)
    g = Dycore(d, cfg)
    dt = cfg["config_dt"]
    g.atm_init_coupled_diagnostics(); g.atm_init_solve_diagnostics(dt)
    for _ in range(2):
        g.atm_srk3(dt); g.summarize_timestep_async(); g.mpas_pool_shift_time_levels()
        print(env, over, g.summarize_timestep_fetch()[0][:4])
    g.mpas_reconstruct(1, True); g.atm_compute_output_diagnostics(1); g.synchronize()
    g.close()
print("done")
