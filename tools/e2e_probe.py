"""Which part of the pipelined request (upload / step / download) is not overlapped?  python tools/e2e_probe.py [cells levels]"""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from mpas_model_b200.case import make_case
from mpas_model_b200.dycore import Dycore
import bench

n = int(sys.argv[1]) if len(sys.argv) > 1 else 40962
nl = int(sys.argv[2]) if len(sys.argv) > 2 else 55
d, cfg = make_case(n, nl); dt = cfg["config_dt"]
g = Dycore(d, cfg); g.atm_init_coupled_diagnostics(); g.atm_init_solve_diagnostics(dt)
for _ in range(3): g.atm_srk3(dt); g.mpas_pool_shift_time_levels()
M = 3
members = []
for m in range(M):
    host = {}
    for name, lev in set(bench.E2E_FIELDS) | set(bench.E2E_OUT):
        host[(name, lev)] = torch.empty(tuple(g.shape(name)), dtype=torch.float64).pin_memory().numpy()
    for (name, lev) in bench.E2E_FIELDS: g._get_real(name, lev, host[(name, lev)])
    members.append(host)

def run(up, down, steps=12, do_step=True):
    g.synchronize(); t0 = time.perf_counter()
    for k in range(steps):
        host = members[k % M]
        if down and k >= M: g.wait_downloads(M - 1)
        if up: g.set_fields_async([(a, l, host[(a, l)]) for (a, l) in bench.E2E_FIELDS])
        if do_step: g.atm_srk3(dt)
        if down: g.get_fields_async([(a, l, host[(a, l)]) for (a, l) in bench.E2E_OUT])
        if not up and do_step: g.mpas_pool_shift_time_levels()
    if down: g.wait_downloads(0)
    g.synchronize()
    return 1e3 * (time.perf_counter() - t0) / steps

for label, kw in (("step only", dict(up=False, down=False)), ("upload + step", dict(up=True, down=False)),
                  ("step + download", dict(up=False, down=True)), ("upload + step + download", dict(up=True, down=True)),
                  ("upload only", dict(up=True, down=False, do_step=False)), ("download only", dict(up=False, down=True, do_step=False)),
                  ("upload + download, no step", dict(up=True, down=True, do_step=False))):
    run(**kw, steps=4)
    print(f"{label:32s} {run(**kw):8.2f} ms/request", flush=True)
