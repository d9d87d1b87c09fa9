#!/bin/bash
# round 2, call ab: TMA-pipelined column solve (k9_acoustic_cell), TMA-staged flux weights (k5, FX_TMA)
mkdir -p gpurun_out
L=$PWD/mpas_model_b200/csrc
MPASB_AC9=1 timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "one_step or every_routine_in_sequence or one_block_of or 55_levels or bench_configuration" > gpurun_out/ab_pytest_ac9.log 2>&1; echo "pytest rc=$?" >> gpurun_out/ab_pytest_ac9.log; tail -3 gpurun_out/ab_pytest_ac9.log
run() { # name, env..., lib
  echo "=== $1"; shift
  env "$@" timeout 200 python tools/quick_bench.py 40962 55 20 > gpurun_out/ab_k_$N.txt 2>&1
  grep -E "^ms/step" gpurun_out/ab_k_$N.txt
  grep -E "k:(k2_dt_edge_b|k5_flux|k6_ac|k9_ac|k7_dt|k2_diag_edge|k2_recover_edge)" gpurun_out/ab_k_$N.txt | awk '{printf "%s %s | ", $1, $2} END {print ""}'
}
N=base run base MPASB_LIB=$L/libmpasb.so
N=ac9 run ac9 MPASB_AC9=1 MPASB_LIB=$L/libmpasb.so
N=ac9_62 run ac9_62 MPASB_AC9=1 MPASB_LIB=$L/libmpasb_vH.so
N=fxtma run fxtma MPASB_LIB=$L/libmpasb_vG.so
N=base2 run base2 MPASB_LIB=$L/libmpasb.so
N=ac9b run ac9b MPASB_AC9=1 MPASB_LIB=$L/libmpasb.so
MPASB_LIB=$L/libmpasb_vG.so timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "one_step or every_routine_in_sequence or irregular" > gpurun_out/ab_pytest_fxtma.log 2>&1; echo "pytest rc=$?" >> gpurun_out/ab_pytest_fxtma.log; tail -3 gpurun_out/ab_pytest_fxtma.log
