"""Regenerates tests/golden/jw_wave_x1.10242_L26.json: the Jablonowski-Williamson baroclinic wave (BASELINE.json
configs[0]: x1.10242, 26 levels, fp64) integrated for 9 simulated days by the CPU oracle (3 minutes on 8 cores).
Run from the repo root:  python tests/golden/make_jw_wave.py"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from mpas_model_b200.case import make_case  # noqa: E402
from oracle.oracle import OracleDycore  # noqa: E402


def run(backend, d, cfg, days=9):
    """Daily diagnostics of the wave: lowest-level pressure extrema (hPa), max |meridional wind| (m/s), w extrema."""
    dt = cfg["config_dt"]
    nC = d["nCells"]
    backend.atm_init_coupled_diagnostics(); backend.atm_init_solve_diagnostics(dt)
    per_day = int(round(86400.0 / dt))
    out = []
    for day in range(days + 1):
        if day:
            for _ in range(per_day):
                backend.atm_srk3(dt); backend.mpas_pool_shift_time_levels()
        backend.atm_compute_output_diagnostics(1)
        backend.mpas_reconstruct(1, False)
        p = backend.get_array("pressure")[:nC, 0] / 100.0
        v = backend.get_array("uReconstructMeridional")[:nC]
        w = backend.get_array("w")[:nC]
        out.append({"day": day, "p_low_min_hPa": float(p.min()), "p_low_max_hPa": float(p.max()),
                    "v_abs_max": float(np.abs(v).max()), "w_min": float(w.min()), "w_max": float(w.max())})
    return out


if __name__ == "__main__":
    d, cfg = make_case(10242, 26)
    rows = run(OracleDycore(d, cfg), d, cfg)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "jw_wave_x1.10242_L26.json")
    with open(path, "w") as f:
        json.dump({"mesh": "x1.10242", "levels": 26, "dt": cfg["config_dt"], "days": rows}, f, indent=1)
    for r in rows:
        print(r)
