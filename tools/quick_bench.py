import sys, time, numpy as np
sys.path.insert(0, '.')
from mpas_model_b200.case import make_case
from mpas_model_b200.dycore import Dycore
n=int(sys.argv[1]); nl=int(sys.argv[2]); K=int(sys.argv[3])
t=time.time(); d,cfg=make_case(n,nl); print("case s", time.time()-t, flush=True)
g=Dycore(d,cfg); dt=cfg['config_dt']
g.atm_init_coupled_diagnostics(); g.atm_init_solve_diagnostics(dt)
for _ in range(3): g.atm_srk3(dt); g.mpas_pool_shift_time_levels()
g.synchronize()
l0=g.kernel_launch_count(); g.timer_start()
for _ in range(K): g.atm_srk3(dt); g.mpas_pool_shift_time_levels()
ms=g.timer_stop(); l1=g.kernel_launch_count()
print("ms/step", ms/K, "launches/step", (l1-l0)/K, "minmax", g.summarize_timestep())
C = n*nl*8
print("B_step model GB", (2197+107)*C/1e9, " achieved GB/s", (2197+107)*C/(ms/K*1e-3)/1e9)
g.set_profile(True)
for _ in range(3): g.atm_srk3(dt); g.mpas_pool_shift_time_levels()
rows=g.get_profile(); g.set_profile(False)
tot=sum(ms for n_,ms,c in rows if not n_.startswith('k:'))
for n_,ms_,c in sorted(rows, key=lambda r:-r[1]):
    print(f"{n_:45s} {ms_/3:9.3f} ms/step  {c//3:4d} calls/step  {ms_/c*1000:9.1f} us/call")
print("sum routines ms/step", tot/3)
