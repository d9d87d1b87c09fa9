#!/bin/bash
# round 2, call i: state after reverting the generic persistent conversion; reference arm with oracle/_ref
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2i_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2i_pytest.log
timeout 400 python bench.py > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err
timeout 400 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/r2i_bench_ref.json 2> gpurun_out/r2i_bench_ref.err
tail -3 gpurun_out/r2i_pytest.log
