#!/bin/bash
# round 2, call h: all column-warp kernels persistent + next-column L2 prefetch; one-field-per-warp flux sweep
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2h_pytest.log
timeout 300 python bench.py --steps 30 --no-e2e --no-cpu-baseline > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err
MPASB_PERSIST=0 timeout 300 python bench.py --steps 30 --no-e2e --no-cpu-baseline > gpurun_out/r2h_bench_nopersist.json 2> gpurun_out/r2h_bench_nopersist.err
MPASB_PF_NEXT=0 timeout 300 python bench.py --steps 30 --no-e2e --no-cpu-baseline > gpurun_out/r2h_bench_nopf.json 2> gpurun_out/r2h_bench_nopf.err
MPASB_FLUX_BOTH=1 timeout 300 python bench.py --steps 30 --no-e2e --no-cpu-baseline > gpurun_out/r2h_bench_fluxboth.json 2> gpurun_out/r2h_bench_fluxboth.err
tail -4 gpurun_out/r2h_pytest.log
