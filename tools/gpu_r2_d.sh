#!/bin/bash
# round 2, call d: parity suite in both arithmetic modes; bench relaxed (scan solve + cell-centred flux), without the scan, strict
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -s > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d_pytest.log
timeout 400 python bench.py --steps 30 > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err
MPASB_NO_SCAN=1 timeout 300 python bench.py --steps 30 --no-e2e --no-cpu-baseline > gpurun_out/r2d_bench_noscan.json 2> gpurun_out/r2d_bench_noscan.err
MPASB_STRICT=1 timeout 300 python bench.py --steps 30 --no-e2e --no-cpu-baseline > gpurun_out/r2d_bench_strict.json 2> gpurun_out/r2d_bench_strict.err
tail -5 gpurun_out/r2d_pytest.log
