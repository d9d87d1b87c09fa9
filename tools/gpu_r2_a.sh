#!/bin/bash
# round 2, call a: full GPU parity suite, N=1 bench (default pitch) and the 128-byte column pitch experiment
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_gpu.txt 2>&1
nproc >> gpurun_out/r2a_gpu.txt; free -g | head -2 >> gpurun_out/r2a_gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
timeout 600 python bench.py > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
MPASB_LDK_ALIGN=16 timeout 600 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/r2a_bench_pitch512.json 2> gpurun_out/r2a_bench_pitch512.err
tail -3 gpurun_out/r2a_pytest.log
