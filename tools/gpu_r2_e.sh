#!/bin/bash
# round 2, call e: ncu --set full of the heavy kernels of one step (relaxed path); raw metrics exported as CSV on the box
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none \
  -k regex:"k5_flux_cell|k6_acoustic_cell|k2_dt_edge_b|k2_dt_cell_f|k2_diag_edge|k2_recover_cell2|k2_smlstep_pert" \
  -s 140 -c 22 -o /tmp/r2e_full -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2e_ncu.log 2>&1
ncu -i /tmp/r2e_full.ncu-rep --page raw --csv > gpurun_out/r2e_raw.csv 2>/dev/null
ls -la /tmp/r2e_full.ncu-rep gpurun_out/
