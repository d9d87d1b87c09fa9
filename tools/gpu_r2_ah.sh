#!/bin/bash
# round 2, call ah: edge tendency at 32 registers (64 / 56 resident warps, 152 B of stack)
mkdir -p gpurun_out
L=$PWD/mpas_model_b200/csrc
for v in "" _eb8 _eb8u2 _eb7 ""; do
  echo "=== base$v"
  MPASB_LIB=$L/libmpasb$v.so timeout 200 python tools/quick_bench.py 40962 55 20 > gpurun_out/ah_k$v.txt 2>&1
  grep -E "^ms/step" gpurun_out/ah_k$v.txt | cut -c1-30
  grep -E "k:(k2_dt_edge_b)" gpurun_out/ah_k$v.txt
done
