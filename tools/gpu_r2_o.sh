#!/bin/bash
# round 2, call o: programmatic dependent launch on/off; regional free-running test
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r2o_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2o_pytest.log
tail -3 gpurun_out/r2o_pytest.log
B="python bench.py --steps 40 --no-e2e --no-cpu-baseline"
timeout 300 $B > gpurun_out/r2o_bench.json 2> gpurun_out/r2o_bench.err
MPASB_PDL=0 timeout 300 $B > gpurun_out/r2o_bench_nopdl.json 2> gpurun_out/r2o_bench_nopdl.err
timeout 300 $B > gpurun_out/r2o_bench_b.json 2> gpurun_out/r2o_bench_b.err
python - <<'PY'
import json
for f in ("r2o_bench", "r2o_bench_nopdl", "r2o_bench_b"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json")); print(f, d["ms_per_step"], d["parity_rel_l2"])
    except Exception as e: print(f, "failed", e)
PY
