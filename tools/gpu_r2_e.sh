#!/bin/bash
# round 2, call e: ncu --set full of the heavy kernels of one step (relaxed path)
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:"k5_flux_cell|k6_acoustic_cell|k2_dt_edge_b|k2_dt_cell_f|k2_diag_edge|k2_recover_cell2|k2_smlstep_pert" \
  -s 600 -c 14 -o gpurun_out/r2e_full -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2e_ncu.log 2>&1
ls -la gpurun_out/r2e_full.ncu-rep
tail -3 gpurun_out/r2e_ncu.log
