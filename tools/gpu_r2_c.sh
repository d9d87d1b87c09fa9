#!/bin/bash
# round 2, call c: parity suite in both arithmetic modes, bench relaxed vs strict, bulk-L2-prefetch sweep
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -s > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log
timeout 400 python bench.py --steps 30 > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err
MPASB_STRICT=1 timeout 300 python bench.py --steps 30 --no-e2e --no-cpu-baseline > gpurun_out/r2c_bench_strict.json 2> gpurun_out/r2c_bench_strict.err
for pf in 100 222 444 888; do
  MPASB_PF=$pf timeout 300 python bench.py --steps 30 --no-e2e --no-cpu-baseline > gpurun_out/r2c_bench_pf$pf.json 2> gpurun_out/r2c_bench_pf$pf.err
done
tail -5 gpurun_out/r2c_pytest.log
