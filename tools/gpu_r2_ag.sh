#!/bin/bash
# round 2, call ag: edge tendency -- spills against occupancy (the stall samples of call ae show 41 local-memory instructions per
# edge beside 40 global loads, on the pipe that bounds the kernel)
mkdir -p gpurun_out
L=$PWD/mpas_model_b200/csrc
for v in "" _eb5 _eb4 _eb49 _ebu2 _ebu2_5 _ebu10_4 ""; do
  echo "=== base$v"
  MPASB_LIB=$L/libmpasb$v.so timeout 200 python tools/quick_bench.py 40962 55 20 > gpurun_out/ag_k$v.txt 2>&1
  grep -E "^ms/step" gpurun_out/ag_k$v.txt | cut -c1-30
  grep -E "k:(k2_dt_edge_b)" gpurun_out/ag_k$v.txt
done
