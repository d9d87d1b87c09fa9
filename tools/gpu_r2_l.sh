#!/bin/bash
# round 2, call l: dd fusion on/off, gather prefetch, k7 block shape
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2l_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2l_pytest.log
B="python bench.py --steps 30 --no-e2e --no-cpu-baseline"
timeout 300 $B > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err
MPASB_NO_DD_FUSE=1 timeout 300 $B > gpurun_out/r2l_bench_nodd.json 2> gpurun_out/r2l_bench_nodd.err
MPASB_PF_NEXT=2 timeout 300 $B > gpurun_out/r2l_bench_pf2.json 2> gpurun_out/r2l_bench_pf2.err
MPASB_CF7=1 timeout 300 $B > gpurun_out/r2l_bench_cf7.json 2> gpurun_out/r2l_bench_cf7.err
MPASB_CF7=1 MPASB_PF_NEXT=2 timeout 300 $B > gpurun_out/r2l_bench_cf7_pf2.json 2> gpurun_out/r2l_bench_cf7_pf2.err
tail -3 gpurun_out/r2l_pytest.log
