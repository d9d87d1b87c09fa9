#!/bin/bash
# round 2, call b: full GPU parity suite (no -x) after the correctly rounded pow
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -s > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
timeout 300 python bench.py --steps 20 --no-e2e > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
tail -5 gpurun_out/r2b_pytest.log
