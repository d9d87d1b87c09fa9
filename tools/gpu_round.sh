#!/bin/bash
# One GPU-box visit: parity tests, bench line, per-kernel table, ncu launch list.  Usage: tools/gpu_round.sh <tag>
tag=${1:-x}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > gpurun_out/${tag}_gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -5 gpurun_out/${tag}_pytest.log
timeout 600 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
timeout 300 python tools/quick_bench.py 40962 55 20 > gpurun_out/${tag}_kernels.txt 2>&1
head -12 gpurun_out/${tag}_kernels.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${tag}_ncu_bench.log 2>&1; echo "ncu rc=$?"
