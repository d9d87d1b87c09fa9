#!/bin/bash
# Per-launch key metrics of ONE whole step (every kernel), as CSV small enough to travel back.
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,launch__occupancy_limit_registers,launch__occupancy_limit_shared_mem,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio
mkdir -p gpurun_out
timeout 900 ncu --metrics $M --clock-control none -s 330 -c 170 --csv --log-file gpurun_out/${1:-u}_step_metrics.csv python tools/quick_bench.py 40962 55 1 > gpurun_out/${1:-u}_ncu.log 2>&1
echo rc=$?; ls -la gpurun_out
