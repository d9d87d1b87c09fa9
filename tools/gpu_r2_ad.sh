#!/bin/bash
# round 2, call ad: what does a change of the shared-memory carve-out between consecutive kernels cost?
mkdir -p gpurun_out
L=$PWD/mpas_model_b200/csrc
tools/micro/carveout_switch > gpurun_out/ad_carveout_switch.txt 2>&1; cat gpurun_out/ad_carveout_switch.txt
run() { # name, env...
  echo "=== $1"; N=$1; shift
  env "$@" timeout 200 python tools/quick_bench.py 40962 55 20 > gpurun_out/ad_k_$N.txt 2>&1
  grep -E "^ms/step" gpurun_out/ad_k_$N.txt | cut -c1-40
  grep -E "k:(k2_dt_edge_b|k5_flux|k6_ac|k9_ac|k7_dt|k3_vert|k2_diag_edge|k2_recover_edge)|sum routines" gpurun_out/ad_k_$N.txt | awk '{printf "%s %s | ", $1, $2} END {print ""}'
}
run base
run carve100 MPASB_CARVEOUT=100
run carve50 MPASB_CARVEOUT=50
run carve25 MPASB_CARVEOUT=25
run carve0 MPASB_CARVEOUT=0
run ac9 MPASB_AC9=1
run ac9_carve100 MPASB_AC9=1 MPASB_CARVEOUT=100
run base_nopdl MPASB_PDL=0
run ac9_nopdl MPASB_AC9=1 MPASB_PDL=0
