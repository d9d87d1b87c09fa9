"""Regenerates tests/golden/oracle_fingerprint_x1.642_L10.json (run from the repo root)."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from mpas_model_b200.case import make_case  # noqa: E402
from oracle.oracle import OracleDycore  # noqa: E402

d, cfg = make_case(642, 10, num_scalars=1)
o = OracleDycore(d, cfg)
dt = cfg["config_dt"]
o.atm_init_coupled_diagnostics(); o.atm_init_solve_diagnostics(dt)
for _ in range(2):
    o.atm_srk3(dt); o.mpas_pool_shift_time_levels()
out = {}
for name in ("u", "w", "rho_zz", "theta_m"):
    a = o.get_array(name).ravel()
    idx = int(np.argmax(np.abs(a)))
    out[name] = {"l2": float(np.linalg.norm(a)), "probe_index": idx, "probe_value": float(a[idx])}
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_fingerprint_x1.642_L10.json")
with open(path, "w") as f:
    json.dump(out, f, indent=1)
print(out)
