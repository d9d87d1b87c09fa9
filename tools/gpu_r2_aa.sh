#!/bin/bash
# round 2, call aa: chunked persistent schedule (k5/k6/k7), wider blocks / persistent chunks for the edge tendency
mkdir -p gpurun_out
L=$PWD/mpas_model_b200/csrc
for v in "" _vA _vB _vC _vD _vE _vF; do
  echo "=== base$v"
  MPASB_LIB=$L/libmpasb$v.so timeout 200 python tools/quick_bench.py 40962 55 20 > gpurun_out/aa_kernels$v.txt 2>&1
  grep -E "^ms/step" gpurun_out/aa_kernels$v.txt
  grep -E "k:(k2_dt_edge_b|k5_flux|k6_ac|k7_dt|k2_diag_edge|k2_recover_edge|k2_smlstep|k2_recover_cell)" gpurun_out/aa_kernels$v.txt | awk '{printf "%s %s | ", $1, $2} END {print ""}'
done
MPASB_LIB=$L/libmpasb_vE.so timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "one_step or every_routine_in_sequence or irregular or bench_configuration" > gpurun_out/aa_pytest_vE.log 2>&1; echo "pytest rc=$?" >> gpurun_out/aa_pytest_vE.log; tail -3 gpurun_out/aa_pytest_vE.log
