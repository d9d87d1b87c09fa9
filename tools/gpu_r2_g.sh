#!/bin/bash
# round 2, call g: next-cell L2 prefetch in the persistent kernels on/off
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "init_routines or async or one_step or every_routine_in_sequence" > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2g_pytest.log
for pf in 0 1; do
  MPASB_PF_NEXT=$pf timeout 300 python bench.py --steps 30 --no-e2e --no-cpu-baseline > gpurun_out/r2g_bench_pf$pf.json 2> gpurun_out/r2g_bench_pf$pf.err
done
tail -3 gpurun_out/r2g_pytest.log
