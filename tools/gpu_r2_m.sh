#!/bin/bash
# round 2, call m: cell-centred Coriolis partial sums on/off
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2m_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2m_pytest.log
B="python bench.py --steps 30 --no-e2e --no-cpu-baseline"
timeout 300 $B > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err
MPASB_NO_COR=1 timeout 300 $B > gpurun_out/r2m_bench_nocor.json 2> gpurun_out/r2m_bench_nocor.err
tail -3 gpurun_out/r2m_pytest.log
