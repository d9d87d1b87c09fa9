#!/bin/bash
# round 2, call f: parity suite; bench with persistent flux / column-solve / cell-tendency kernels and their variants
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -s > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_pytest.log
timeout 400 python bench.py --steps 30 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err
for v in 0 2; do
  MPASB_AC6=$v timeout 300 python bench.py --steps 30 --no-e2e --no-cpu-baseline > gpurun_out/r2f_bench_ac6_$v.json 2> gpurun_out/r2f_bench_ac6_$v.err
done
MPASB_OLD_CELL_F=1 timeout 300 python bench.py --steps 30 --no-e2e --no-cpu-baseline > gpurun_out/r2f_bench_oldf.json 2> gpurun_out/r2f_bench_oldf.err
tail -5 gpurun_out/r2f_pytest.log
