#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_regional.py tests/test_parity_gpu.py -m gpu -q -k "regional or every_routine_in_sequence or one_step" > gpurun_out/r2x_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2x_pytest.log; tail -3 gpurun_out/r2x_pytest.log
timeout 300 python tools/quick_bench.py 40962 55 20 > gpurun_out/r2x_kernels.txt 2>&1; head -3 gpurun_out/r2x_kernels.txt; grep -E "smlstep|recover_cell2" gpurun_out/r2x_kernels.txt
timeout 300 python bench.py --steps 40 --no-e2e --no-cpu-baseline > gpurun_out/r2x_bench.json 2> gpurun_out/r2x_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r2x_bench.json')); print('ms/step', d['ms_per_step'])"
