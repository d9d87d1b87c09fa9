#!/usr/bin/env python
"""Pin the oracle (and through it the CUDA path) against output of the REAL MPAS-Atmosphere build.

Nobody can build the reference in this image (no Fortran compiler, MPI or NetCDF), so the oracle's parity is
"unpinned" (DESIGN.md §2).  On a machine that has the Fortran build this closes the gap in three commands:

  1. python tools/compare_with_reference_output.py write-init x1.10242.init.nc --cells 10242 --levels 26
        writes this repository's JW case as an MPAS init file (CDF-5; use --cdf 2 for 64-bit-offset),
  2. run the reference on it:  atmosphere_model with config_init_case-independent namelist.atmosphere
        &nhyd_model  config_dt = <printed dt>, config_time_integration_order = 2, config_dynamics_split_steps = 3,
                     config_number_of_sub_steps = 2, config_horiz_mixing = '2d_smagorinsky', config_len_disp = <printed>, /
        &physics     config_physics_suite = 'none' /          (or a core built without physics, src/core_atmosphere/Makefile:9-11)
        &io          config_pio_num_iotasks = 0 /  and in streams.atmosphere io_type="pnetcdf,cdf5" for the restart stream,
        restart_interval = N * config_dt,
  3. python tools/compare_with_reference_output.py compare x1.10242.init.nc restart.<date>.nc --steps N [--backend oracle|cuda]
        steps this repository's oracle (or the CUDA library) N times from the same init file and prints the relative
        L2 difference of u, w, rho_zz, theta_m and qv against the reference's restart fields, plus the
        `global min, max w/u` values to hold against the reference's log (mpas_atm_time_integration.F:8304, 8319).

The north-star bars are 1e-11 after one step and 1e-6 after one simulated day.
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mpas_model_b200 import initfile, ncio  # noqa: E402

RESTART_VARS = {"u": ("u", 3), "w": ("w", 3), "rho_zz": ("rho_zz", 3), "theta_m": ("theta_m", 3), "qv": ("scalars", 0)}


def write_init(args):
    from mpas_model_b200.case import make_case
    d, cfg = make_case(args.cells, args.levels, num_scalars=1)
    initfile.write_init_file(d, args.path, version=args.cdf)
    print(f"wrote {args.path}: nCells={d['nCells']} nVertLevels={d['nVertLevels']} config_dt={cfg['config_dt']:g} "
          f"config_len_disp={cfg['config_len_disp']:.6f}")


def run_from_init(path, steps, backend="oracle", dt=None):
    d, cfg = initfile.read_init_file(path, dt=dt)
    if backend == "cuda":
        from mpas_model_b200.dycore import Dycore
        b = Dycore(d, cfg)
    else:
        from oracle.oracle import OracleDycore
        b = OracleDycore(d, cfg)
    step = cfg["config_dt"]
    b.atm_init_coupled_diagnostics(); b.atm_init_solve_diagnostics(step)
    for _ in range(steps):
        b.atm_srk3(step); b.mpas_pool_shift_time_levels()
    return d, cfg, b


def compare(args):
    d, cfg, b = run_from_init(args.init, args.steps, args.backend, args.dt)
    dims, attrs, v = ncio.read(args.reference, only=set(RESTART_VARS))
    nC, nE = d["nCells"], d["nEdges"]
    worst = 0.0
    for name, (field, _) in RESTART_VARS.items():
        if name not in v:
            print(f"{name:8s} not in {args.reference}")
            continue
        ref = np.asarray(v[name].data[-1], dtype=np.float64)
        mine = b.get_array(field, 1)
        mine = mine[:nE] if field == "u" else mine[:nC]
        if field == "scalars":
            mine = mine[..., d["index_qv"]]
        num, den = np.linalg.norm((mine - ref).ravel()), np.linalg.norm(ref.ravel())
        rel = num / den if den > 0 else num
        worst = max(worst, rel)
        print(f"{name:8s} rel-L2 {rel:.3e}   max |diff| {np.abs(mine - ref).max():.3e}")
    mm = b.summarize_timestep()
    print(f"global min, max w {mm[0]:.10g} {mm[1]:.10g}\nglobal min, max u {mm[2]:.10g} {mm[3]:.10g}")
    print(f"worst rel-L2 {worst:.3e} after {args.steps} step(s) of {cfg['config_dt']:g} s")
    return worst


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    sub = ap.add_subparsers(dest="cmd", required=True)
    w = sub.add_parser("write-init"); w.add_argument("path"); w.add_argument("--cells", type=int, default=10242)
    w.add_argument("--levels", type=int, default=26); w.add_argument("--cdf", type=int, default=5, choices=(2, 5))
    c = sub.add_parser("compare"); c.add_argument("init"); c.add_argument("reference"); c.add_argument("--steps", type=int, default=1)
    c.add_argument("--backend", default="oracle", choices=("oracle", "cuda")); c.add_argument("--dt", type=float, default=None)
    args = ap.parse_args(argv)
    return write_init(args) if args.cmd == "write-init" else compare(args)


if __name__ == "__main__":
    main()
