#!/usr/bin/env python
"""bench.py -- JW baroclinic-wave dycore step throughput on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    torchrun --nproc-per-node N ... bench.py --gpus N ...      (N > 1, one rank per GPU)

A "step" is one atm_srk3 call (mpas_atm_time_integration.F:803) over the whole mesh.
N = 1 workload: BASELINE.json configs[1], JW wave on x1.40962 (120 km), 55 levels, fp64,
dt = 720 s.  N > 1: the mesh is partitioned with one block per GPU (weak scaling is not
natural for a fixed global mesh family, so N = 2/4 use x1.163842 and N = 8 uses x1.655362 as
BASELINE.json names them: per-GPU work is 40962 / 81921 / 40961 / 81920 columns, i.e. weak scaling
within a factor of two; cell_columns_per_s is the figure comparable across N).

Prints ONE JSON line (rank 0).  `value` is the device-resident whole-job rate in cell-column
updates/s (nCells x steps/s: the BASELINE throughput figure that is comparable across the
per-N meshes; `steps_per_s` and `sdpd` are in the same line); `e2e` is the
same step driven through the C ABI with HOST buffers (restart-state upload, diagnostics
recompute, step, state download inside the timed region).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {1: (40962, 55), 2: (163842, 55), 4: (163842, 55), 8: (655362, 55)}
# restart-stream state that a host keeps between steps (Registry.xml stream "restart")
E2E_FIELDS = (("u", 1), ("w", 1), ("rho_zz", 1), ("theta_m", 1), ("scalars", 1), ("ru", 1), ("rw", 1),
              ("rtheta_p", 1), ("rho_p", 1), ("exner", 1), ("pressure_p", 1))
E2E_OUT = (("u", 2), ("w", 2), ("rho_zz", 2), ("theta_m", 2), ("scalars", 2), ("ru", 1), ("rw", 1),
           ("rtheta_p", 1), ("rho_p", 1), ("exner", 1), ("pressure_p", 1))


def bind_to_gpu_numa_node(local_rank):
    """Pin this rank to the CPUs next to its GPU (NVML's ideal-CPU mask), so that the pinned host buffers of the e2e leg are
    first-touched on the GPU's own NUMA node instead of every rank's on node 0.  Returns what was done (for the JSON line)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        before = os.sched_getaffinity(0)
        pynvml.nvmlDeviceSetCpuAffinity(h)
        return {"bound": True, "cpus": len(os.sched_getaffinity(0)), "cpus_before": len(before), "_restore": before}
    except Exception as e:                                # no NVML, a container cpuset without those CPUs, ...
        return {"bound": False, "why": str(e)[:80]}


def model_bytes_per_step(n_cells, n_levels, n_scalars, real_bytes=8):
    """B_step = (2197 + 107 S) * nVertLevels * nCells * sizeof(real)."""
    return (2197 + 107 * n_scalars) * n_levels * n_cells * real_bytes


class ClockSampler(threading.Thread):
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, gpu=0):
        super().__init__(daemon=True)
        self.gpu, self.rows, self._halt = gpu, [], threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._halt.wait(0.2)

    def stop(self):
        self._halt.set()
        self.join(timeout=6)
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        reasons = sorted({n for r in self.rows for n, v in zip(names, r[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# algorithmic bytes per launch of the kernels worth a roofline line, in units of
# C = nVertLevels * nCells * 8 B (SURVEY.md §8a per-routine model; DESIGN.md §4 per kernel)
KERNEL_MODEL_C = {
    # acoustic cell solve: 1 E (ru_p or tend_u) + 21 C read, 5 C written on a later small step (29 C); the first
    # small step of a stage does not read rho_pp, rtheta_pp, rw_p, wwAvg (25 C); 9 first + 3 later per step
    "k:k3_acoustic_cell": (9 * 25.0 + 3 * 29.0) / 12,
    "k:k_acoustic_cell": 29.0,
    # edge tendency: rk 1 reads 5 E + 8 C + 1 V and writes 3 E (34 C); rk 2,3 read 5 E + 3 C, write 1 E (21 C)
    "k:k2_dt_edge_b": (3 * 34.0 + 6 * 21.0) / 9,
    "k:k2_dt_edge_b<false>": (3 * 34.0 + 6 * 21.0) / 9,
    # with the Coriolis sum moved to k8_coriolis_cell the edge kernel reads 2 partial sums (6 C) instead of gathering u and
    # pv_edge over edgesOnEdge (their own-edge reads stay): same model bytes + 6 C; k8: u, pv_edge (2 E) read, 6 C written
    "k:k2_dt_edge_b<true>": (3 * 34.0 + 6 * 21.0) / 9 + 6.0,
    "k:k8_coriolis_cell": 12.0,
    "k:k_dt_edge_b": (3 * 34.0 + 6 * 21.0) / 9,
    # cell tendency after the per-edge flux split: 2 E fluxes + ru, ru_save (rk > 1) + 13 C read, 3-5 C written
    "k:k2_dt_cell_f": 24.0,
    "k:k_dt_cell_f": 22.0,
    "k:k2_dt_edge_flux": 11.0,       # ru (E), w, theta_m read; 2 E written
    "k:k2_recover_cell2": 19.0,      # zb_cell + zb3_cell (12 C), ru (E), rho_zz, w r/w
    # relaxed-arithmetic path: the column solve as a warp-level scan (same operands as k3_acoustic_cell), the cell-centred
    # flux sweep (ru: E, w, theta_m read; 2 C written) and the cell tendency without the two per-edge flux arrays
    "k:k6_acoustic_cell": (9 * 25.0 + 3 * 29.0) / 12,
    "k:k5_flux_cell": 7.0,
    "k:k2_dt_cell_f<true>": 14.0,
    "k:k7_dt_cell_f": 14.0,
    "k:k2_dt_cell_f<false>": 24.0,
}


NCU_STEP_CSV = os.path.join("profiles", "r2_ncu_step_metrics_ae.csv")


def kernel_traffic(kernel, n_cells, n_lev, rb):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel`, averaged over the launches of one step in the
    committed ncu capture NCU_STEP_CSV (tools/ncu_step_metrics.sh: x1.40962 x 55 levels, fp64; template arguments of the
    kernel name ignored).  Only valid for that workload; None otherwise."""
    path = os.path.join(ROOT, NCU_STEP_CSV)
    if not (os.path.exists(path) and (n_cells, n_lev, rb) == (40962, 55, 8)):
        return None
    import csv
    tot, launches = 0.0, set()
    with open(path, errors="ignore") as f:
        rows = [r for r in csv.reader(f) if len(r) > 10]
    hdr = rows[0]
    ik, im, iv, iid = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
    iu = hdr.index("Metric Unit")
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    base = lambda n: n.split("(")[0].replace("void ", "").split("<")[0].strip()
    for r in rows[1:]:
        if base(r[ik]) == base(kernel) and r[im] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            tot += float(r[iv].replace(",", "")) * scale.get(r[iu], 1.0); launches.add(r[iid])
    return tot / len(launches) if launches else None


def workload_for(args):
    n_cells, n_lev = WORKLOADS.get(args.gpus, WORKLOADS[1])
    return (args.cells or n_cells), (args.levels or n_lev)


def line_common(args, n_cells, n_lev, dt, world):
    return {
        "metric": "JW wave dycore cell-column updates/sec (= steps/sec x nCells; steps_per_s and sdpd alongside)",
        "unit": "cell-columns/s", "n_gpus": world, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic (generated icosahedral SCVT mesh + JW case 2 initial state)",
        "config": {"workload": f"JW baroclinic wave x1.{n_cells} {n_lev} levels fp64 dt={dt:g}s S={args.scalars}",
                   "cells_per_gpu": n_cells // world,
                   "decomposition": "single block" if world == 1 else f"{world} blocks (recursive coordinate bisection), one per GPU, reference halo lists",
                   "scaling_note": "BASELINE.json names one mesh per GPU count (x1.40962 @1, x1.163842 @2/4, x1.655362 @8), so "
                                   "`value` is the one BASELINE throughput figure that is comparable across N: cell-column "
                                   "updates/s = nCells x steps/s; steps_per_s and sdpd are reported beside it",
                   "l2": "no flush: every step streams the block's fields (>= 3.5 GB at 40962 cells x 55 levels), >> 126 MB L2",
                   "namelist": "reference defaults (SRK3 order 2, 3 dynamics substeps, 2 acoustic substeps, monotonic split transport)"
                               if getattr(args, "order", 2) == 2 else
                               "reference defaults except config_time_integration_order = 3 (acoustic substeps 1, 1, 2 per stage, a second "
                               "vertical-coefficient pass; the byte model of step_roofline is the order-2 one)",
                   "timed_region": "K x (atm_srk3 + mpas_pool_shift_time_levels), fields resident; excluded: mesh/state generation, "
                                   "upload, init-time diagnostics, the min/max summary (taken once after the loop), H2D/D2H (those are in e2e)"},
        # outside `config` so that both arms print the same config object
        "implementation": "B200 CUDA library (libmpasb.so) through the C ABI" if args.impl != "reference"
                          else "the reference's CPU dycore: oracle/_ref (its Fortran source transliterated to C++ and compiled here) at N = 1, "
                               "the C++/OpenMP restatement (oracle/) for decomposed N > 1 runs; not the gfortran build",
    }


def reference_arm(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path.  The Fortran build is
    impossible in this image (no Fortran compiler/MPI/NetCDF), so this times the C++ restatement
    (oracle/, OpenMP over the same cell/edge ranges) on the box's host cores, on the same workload
    and decomposition as the GPU arm."""
    if rank != 0:
        return
    from oracle import oracle as orc
    from oracle.oracle import OracleDycore
    n_threads = int(os.environ.get("MPASB_REF_THREADS", os.cpu_count()))
    orc.set_threads(n_threads)                    # torchrun forces OMP_NUM_THREADS=1 into its ranks; the baseline uses every core
    n_cells, n_lev = workload_for(args)
    blocks = None
    extrapolated = False
    kind = "port"
    if args.gpus > 1:
        # the N-block decomposition itself, every block in this process in lock step with in-process halo exchanges
        # (the oracle's "virtual ranks") -- the whole mesh is stepped, nothing is extrapolated.  Only the 8-block
        # x1.655362 case (55 GB of oracle state, minutes per step) falls back to ONE block x 1/N, labelled as such.
        from mpas_model_b200 import multigpu as mg
        prefix = mg.prepare_blocks(n_cells, n_lev, args.scalars, args.gpus)
        whole = args.gpus <= 4 or os.environ.get("MPASB_REF_WHOLE_MESH")
        ranks = range(args.gpus) if whole else range(1)
        blocks = []
        for r in ranks:
            rec = mg.load_block(prefix, r)
            ob = OracleDycore(rec["block"], rec["cfg"], rank=r)
            for halo_kind, k in mg.KINDS:
                ob.set_halo_lists(k, rec["ex"][halo_kind])
            blocks.append(ob)
            cfg = rec["cfg"]
        frac = 1.0 if whole else 1.0 / args.gpus
        extrapolated = not whole
        sample_what = (f"all {args.gpus} blocks of the decomposition in lock step (in-process halo exchanges)" if whole else
                       f"EXTRAPOLATED: block 0 of {args.gpus} only ({rec['block']['nCellsSolve']} owned cells + halo, halo values frozen), rate scaled by 1/{args.gpus}")
    else:
        from mpas_model_b200.case import make_case
        from oracle import ref as oref
        d, cfg = make_case(n_cells, n_lev, num_scalars=args.scalars, **(dict(config_time_integration_order=3) if args.order == 3 else {}))
        frac = 1.0
        sample_what = "the whole mesh"
        if oref.build() and not os.environ.get("MPASB_REF_PORT"):
            # oracle/_ref: the reference's own mpas_atm_time_integration.F, transliterated to C++ at build time (oracle/f2cpp.py),
            # entered by all host threads with the index ranges and barriers of its MPAS_OPENMP build
            blocks = [oref.RefDycore(d, cfg, threads=n_threads)]
            kind = "reference"
            sample_what += " (reference source transliterated Fortran -> C++, OpenMP threading of mpas_atm_threading.F; without the trailing mpas_reconstruct)"
        else:
            blocks = [OracleDycore(d, cfg)]
    dt = cfg["config_dt"]
    multi = len(blocks) > 1

    def one_step():
        if multi:
            orc.step(blocks, dt)
        else:
            blocks[0].atm_srk3(dt)
        for b in blocks:
            b.mpas_pool_shift_time_levels()

    if multi:
        orc.exchange(blocks, "initialization:u")
    for b in blocks:
        b.atm_init_coupled_diagnostics(); b.atm_init_solve_diagnostics(dt)
    if multi:
        orc.exchange(blocks, "initialization:pv_edge,ru,rw")
    t0 = time.time(); one_step(); t1 = time.time() - t0
    warm = max(0, min(args.warmup - 1, int(20.0 / max(t1, 1e-3))))
    for _ in range(warm):
        one_step()
    steps = max(1, min(args.steps, int(100.0 / max(t1, 1e-3))))
    t0 = time.time()
    for _ in range(steps):
        one_step()
    el = time.time() - t0
    sps = frac * steps / el                       # whole-mesh steps/s
    v = n_cells * sps                             # cell-column updates/s (the line's unit)
    cores = n_threads
    line = line_common(args, n_cells, n_lev, dt, args.gpus)
    line.update({
        "impl": "reference", "value": v, "steps": steps, "warmup": warm + 1, "ms_per_step": 1e3 / sps,
        "cpu_baseline": {"value": v, "unit": "cell-columns/s", "cores": cores, "kind": kind,
                         "sample": f"{steps} full atm_srk3 steps on {sample_what}", "extrapolated": extrapolated},
        "e2e": {"value": v, "unit": "cell-columns/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "steps_per_s": sps, "sdpd": dt * sps, "cell_columns_per_s": v,
    })
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--cells", type=int, default=0, help="override the mesh (10*4^n+2 cells)")
    ap.add_argument("--levels", type=int, default=0)
    ap.add_argument("--scalars", type=int, default=1)
    ap.add_argument("--precision", default="double", choices=("double", "single"), help="RKIND of the library build")
    ap.add_argument("--order", type=int, default=2, choices=(2, 3), help="config_time_integration_order (SURVEY.md §8d: order 3 as a second line; N = 1 only)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-members", type=int, default=3, help="independent host-resident instances alternating in the e2e leg (1..4)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        reference_arm(args, rank, world)
        return

    import torch
    from mpas_model_b200.dycore import Dycore

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the dycore step has no CPU path")
    if world != args.gpus:
        raise SystemExit(f"bench.py: --gpus {args.gpus} needs {args.gpus} ranks (torch.distributed.run), got WORLD_SIZE={world}")
    n_cells, n_lev = workload_for(args)
    torch.cuda.set_device(local_rank)
    numa = bind_to_gpu_numa_node(local_rank)          # before any pinned allocation: first touch decides the NUMA node
    dist = None
    if world > 1:
        from mpas_model_b200 import multigpu as mg
        dist = mg.init_distributed("gloo")
        g, d, cfg, _ex, _ = mg.setup_rank(n_cells, n_lev, args.scalars, rank, world, local_rank, precision=args.precision)
        dt = cfg["config_dt"]
    else:
        from mpas_model_b200.case import make_case
        over = dict(config_time_integration_order=3) if args.order == 3 else {}
        d, cfg = make_case(n_cells, n_lev, num_scalars=args.scalars, **over)
        dt = cfg["config_dt"]
        g = Dycore(d, cfg, device=0, precision=args.precision)
        g.atm_init_coupled_diagnostics(); g.atm_init_solve_diagnostics(dt)

    def barrier():
        g.synchronize(); torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    def sum_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t[0])

    # ---------------- device-resident throughput
    STATE = ("u", "w", "rho_zz", "theta_m", "scalars")
    snap = {}                                      # GPU state after warm-up step 1 and after the last warm-up step (parity check below)
    for k in range(1, args.warmup + 1):
        g.atm_srk3(dt); g.mpas_pool_shift_time_levels()
        if world == 1 and not args.no_cpu_baseline and k in (1, args.warmup):
            snap[k] = {n: g.get_array(n, 1).astype(np.float64) for n in STATE}
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:                                  # one nvidia-smi poller per job is enough (the line reports rank 0's GPU)
        sampler.start()
    l0 = g.kernel_launch_count()
    g.timer_start()
    for k in range(args.steps):
        g.atm_srk3(dt)
        if k + 1 < args.steps:
            g.mpas_pool_shift_time_levels()
    ms = g.timer_stop()
    barrier()
    launches = int(sum_over_ranks(g.kernel_launch_count() - l0))
    ms_per_step = max_over_ranks(ms) / args.steps
    steps_per_s = 1e3 / ms_per_step
    value = n_cells * steps_per_s                 # cell-column updates/s over all ranks (the line's unit)
    mm = g.summarize_timestep()
    if dist is not None:
        t = torch.tensor([-mm[0], mm[1], -mm[2], mm[3]], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)              # the reference's mpas_dmpar_min/max_real (TI:8300-8318)
        mm = (-float(t[0]), float(t[1]), -float(t[2]), float(t[3]))
    minmax = mm                                     # of the state the LAST timed step produced (time level 2, before the shift)
    g.mpas_pool_shift_time_levels()

    # ---------------- end to end through the C ABI with pinned host buffers (every rank moves its own block)
    # A request = restart-stream state of one model instance in (host -> device), the diagnostics a restart read is followed
    # by, one atm_srk3 step, the new state out (device -> host) and the step's logged min/max.  TWO independent instances
    # (think two ensemble members) alternate, so that while instance A steps on the GPU, B's result is on its way down and
    # A's next input on its way up (mpasb_set_fields_async / mpasb_get_fields_async: separate copy streams, events, one host
    # wait per request).  Every request still pays its full upload, step and download; `serial_ms_per_step` is the same
    # request with nothing overlapped (one instance, blocking set_field / get_field per array).
    e2e = None
    if not args.no_e2e:
        tdt = torch.float64 if g.rdtype == np.float64 else torch.float32
        members = []
        M = max(1, min(4, args.e2e_members))
        for m in range(M):
            host = {}
            for name, lev in set(E2E_FIELDS) | set(E2E_OUT):
                host[(name, lev)] = torch.empty(tuple(g.shape(name)), dtype=tdt).pin_memory().numpy()
            for (name, lev) in E2E_FIELDS:
                g._get_real(name, lev, host[(name, lev)])
            members.append(host)
        h2d = sum(members[0][k].nbytes for k in E2E_FIELDS)
        d2h = sum(members[0][k].nbytes for k in E2E_OUT) + 32
        e2e_steps = max(4, min(args.steps, 20)) // 2 * 2

        def serial_step(host):
            for (name, lev) in E2E_FIELDS:                 # host pools -> device (pinned)
                g._set_real(name, lev, host[(name, lev)])
            g.atm_init_solve_diagnostics(dt)                # what a restart read is followed by (mpas_atm_core.F:524)
            if dist is not None:
                g.exchange_halo_group("initialization:pv_edge,ru,rw")      # mpas_atm_core.F:288
            g.atm_srk3(dt)
            for (name, lev) in E2E_OUT:                     # device -> host pools
                g._get_real(name, lev, host[(name, lev)])
            out = g.summarize_timestep()                    # the step's logged result (TI:8304,8319)
            for name in ("u", "w", "rho_zz", "theta_m", "scalars"):     # mpas_pool_shift_time_levels on the host: swap pointers
                host[(name, 1)], host[(name, 2)] = host[(name, 2)], host[(name, 1)]
            return out

        def submit(host):
            g.set_fields_async([(n, l, host[(n, l)]) for (n, l) in E2E_FIELDS])
            g.atm_init_solve_diagnostics_async(dt)
            if dist is not None:
                g.exchange_halo_group_async("initialization:pv_edge,ru,rw")      # enqueued only: the host keeps running ahead
            g.atm_srk3(dt)
            g.get_fields_async([(n, l, host[(n, l)]) for (n, l) in E2E_OUT])
            g.summarize_timestep_async()

        def retire(host):
            """the result of this member's request in flight is on the host: log line + host-side time-level shift"""
            out, _ = g.summarize_timestep_fetch(scalars=False)
            for name in ("u", "w", "rho_zz", "theta_m", "scalars"):
                host[(name, 1)], host[(name, 2)] = host[(name, 2)], host[(name, 1)]
            return out

        serial_step(members[0])
        barrier()
        t0 = time.perf_counter()
        ns = max(3, e2e_steps // 4)
        for _ in range(ns):
            serial_step(members[0])
        barrier()
        serial_s = max_over_ranks((time.perf_counter() - t0) / ns)

        # pipelined: request k belongs to member k % M; a member's next request needs its previous result on the host
        # (its output arrays are the next input), so request k - M is retired before request k is submitted
        for m in range(1, M):                               # every member has stepped once before the clock starts
            serial_step(members[m])
        submit(members[0])
        g.wait_downloads(0); retire(members[0])
        barrier()
        t0 = time.perf_counter()
        for k in range(e2e_steps):
            if k >= M:
                g.wait_downloads(M - 1)                     # request k - M is complete on the host ...
                retire(members[(k - M) % M])                # ... log line, host-side time-level shift
            submit(members[k % M])
        for j in range(max(0, e2e_steps - M), e2e_steps):
            g.wait_downloads(e2e_steps - 1 - j); retire(members[j % M])
        barrier()
        e2e_s = max_over_ranks((time.perf_counter() - t0) / e2e_steps)
        e2e = {"value": n_cells / e2e_s, "unit": "cell-columns/s", "steps_per_s": 1.0 / e2e_s, "h2d_bytes_per_step": int(sum_over_ranks(h2d)),
               "d2h_bytes_per_step": int(sum_over_ranks(d2h)), "ms_per_step": 1e3 * e2e_s, "serial_ms_per_step": 1e3 * serial_s,
               "members": M,
               "mode": f"{M} independent host-resident instances (ensemble members) taking turns on one device handle; the upload of "
                       "a later request, the step, and the download of an earlier one overlap (copy streams + events); every request "
                       "moves its full state both ways and waits for its own previous result; serial_ms_per_step is one instance, "
                       "nothing overlapped"}
    clocks = sampler.stop() if rank == 0 else None
    if numa.get("_restore"):                            # the CPU baseline below uses every core again
        os.sched_setaffinity(0, numa.pop("_restore"))

    # ---------------- per-kernel timing (CUDA events on the launching stream) for the roofline object
    barrier()
    g.set_profile(True)
    nprof = 3
    for _ in range(nprof):
        g.atm_srk3(dt); g.mpas_pool_shift_time_levels()
    rows = g.get_profile(); g.set_profile(False)
    barrier()
    krows = {n: (m, c) for n, m, c in rows if n.startswith("k:")}
    ksum = sum(m for m, _ in krows.values())
    dom = max(krows.items(), key=lambda kv: kv[1][0])
    dom_name, (dom_ms, dom_cnt) = dom
    peak, peak_src = measured_peak_gbs()
    rb = 8 if args.precision == "double" else 4
    C = n_lev * d["nCells"] * rb                      # this rank's block (owned + halo columns)
    model_c = KERNEL_MODEL_C.get(dom_name)
    dom_us = 1e3 * dom_ms / dom_cnt
    achieved = (model_c * C / (dom_us * 1e-6) / 1e9) if model_c else None
    roofline = {"bound": "hbm", "kernel": dom_name[2:], "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": (achieved / peak) if achieved else None,
                "traffic": kernel_traffic(dom_name[2:], n_cells, n_lev, rb) if world == 1 else None,
                "traffic_source": f"replayed from the committed ncu capture {NCU_STEP_CSV} (dram__bytes_read.sum + dram__bytes_write.sum per "
                                  "launch, averaged over the launches of one step of this workload); not measured in this run",
                "avg_launch_us": dom_us, "share_of_step": dom_ms / ksum,
                "algorithmic_bytes_per_launch": (model_c * C) if model_c else None, "peak_source": peak_src}
    B_step = model_bytes_per_step(n_cells, n_lev, args.scalars, rb)
    step_gbs = B_step / (ms_per_step * 1e-3) / 1e9

    # ---------------- CPU baseline: the oracle on the host cores, bounded sample (rank 0, N = 1 only); the same oracle
    # run is the parity check of this very bench run: GPU state after warm-up steps 1 and W against the oracle's
    cpu = None
    parity = None
    if not args.no_cpu_baseline and world == 1:
        from oracle import oracle as orc
        from oracle.oracle import OracleDycore
        n_threads = int(os.environ.get("MPASB_REF_THREADS", os.cpu_count()))
        orc.set_threads(n_threads)
        o = OracleDycore(d, cfg, precision=args.precision)
        o.atm_init_coupled_diagnostics(); o.atm_init_solve_diagnostics(dt)

        def rel_l2_all(k):
            out = {}
            for nme in STATE:
                a, r = snap[k][nme], o.get_array(nme, 1).astype(np.float64)
                out[nme] = float(np.linalg.norm((a - r).ravel()) / max(np.linalg.norm(r.ravel()), 1e-300))
            return out

        o.atm_srk3(dt); o.mpas_pool_shift_time_levels()
        parity = {"against": "CPU oracle (oracle/, same precision) stepped from the same initial state",
                  "after_steps": {}}
        if 1 in snap:
            parity["after_steps"]["1"] = rel_l2_all(1)
        t0 = time.time(); n = 0
        while n < 3 or (time.time() - t0 < 10.0 and n < 20) or (n + 1 < args.warmup <= 12):
            o.atm_srk3(dt); o.mpas_pool_shift_time_levels(); n += 1
            if n + 1 == args.warmup and args.warmup in snap and args.warmup > 1:
                t1 = time.time()
                parity["after_steps"][str(args.warmup)] = rel_l2_all(args.warmup)
                t0 += time.time() - t1              # the comparison is not part of the CPU timing
        el = time.time() - t0
        parity["parity_rel_l2"] = max(parity["after_steps"]["1"].values()) if "1" in parity["after_steps"] else None
        parity["bar"] = "1e-11 after one RK3 step (fp64); fp32: compared with the fp32 restatement"
        cpu = {"value": n_cells * n / el, "unit": "cell-columns/s", "steps_per_s": n / el, "cores": n_threads,
               "kind": "port", "sample": f"{n} full atm_srk3 steps of the same workload (C++/OpenMP restatement, not the Fortran build)"}

    if rank != 0:
        return
    line = line_common(args, n_cells, n_lev, dt, world)
    if world > 1:
        line["halo_exchange"] = ("CUDA-IPC put/get kernels over NVLink" if getattr(g, "p2p_on", False) else "pack -> NCCL send/recv -> unpack") \
            + ("" if os.environ.get("MPASB_NO_OVERLAP") else ", 3 of the exchange groups overlapped with compute on a priority stream")
    if args.precision == "single":
        line["dtype"] = "f32"
        line["config"]["workload"] = line["config"]["workload"].replace("fp64", "fp32 (PRECISION=single build)")
    line.update({
        "value": value, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
        "e2e": e2e, "host_affinity": numa, "gpu_launches": launches, "clocks": clocks, "roofline": roofline,
        "step_roofline": {"model_bytes_per_step": B_step, "achieved_gbs": step_gbs,
                          "frac_of_measured_peak": step_gbs / (peak * world), "frac_of_nominal_8TBs": step_gbs / (8000.0 * world)},
        "cpu_baseline": cpu, "parity": parity,
        "parity_rel_l2": parity["parity_rel_l2"] if parity else None,
        "steps_per_s": steps_per_s, "sdpd": dt * steps_per_s, "cell_columns_per_s": value,
        "minmax_w_u": list(minmax),
        "kernel_ms_per_step": {k[2:]: round(v[0] / nprof, 4) for k, v in sorted(krows.items(), key=lambda kv: -kv[1][0])},
        "kernel_launches_per_step": {k[2:]: v[1] // nprof for k, v in sorted(krows.items(), key=lambda kv: -kv[1][0])},
        "kernel_ms_sum": round(ksum / nprof, 4),
    })
    print(json.dumps(line))


if __name__ == "__main__":
    main()
