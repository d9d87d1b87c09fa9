#!/bin/bash
# round 2, call p: alternating sweep direction (snake) and programmatic dependent launch, on/off
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2p_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2p_pytest.log
tail -3 gpurun_out/r2p_pytest.log
B="python bench.py --steps 40 --no-e2e --no-cpu-baseline"
timeout 300 $B > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err
MPASB_SNAKE=0 timeout 300 $B > gpurun_out/r2p_bench_nosnake.json 2> gpurun_out/r2p_bench_nosnake.err
MPASB_PDL=0 timeout 300 $B > gpurun_out/r2p_bench_nopdl.json 2> gpurun_out/r2p_bench_nopdl.err
MPASB_PDL=0 MPASB_SNAKE=0 timeout 300 $B > gpurun_out/r2p_bench_neither.json 2> gpurun_out/r2p_bench_neither.err
python - <<'PY'
import json
for f in ("r2p_bench", "r2p_bench_nosnake", "r2p_bench_nopdl", "r2p_bench_neither"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json")); print(f, round(d["ms_per_step"], 3), d["minmax_w_u"][1])
        print("   ", {k: v for k, v in list(d["kernel_ms_per_step"].items())[:12]})
    except Exception as e: print(f, "failed", e)
PY
