#!/bin/bash
mkdir -p gpurun_out
timeout 300 python bench.py --steps 20 --no-cpu-baseline > gpurun_out/r2s_bench.json 2> gpurun_out/r2s_bench.err
timeout 300 python bench.py --steps 20 --no-cpu-baseline --e2e-members 2 > gpurun_out/r2s_bench_m2.json 2> gpurun_out/r2s_bench_m2.err
timeout 300 python -m pytest tests/test_parity_gpu.py -q -k "summarize or batched" > gpurun_out/r2s_pytest.log 2>&1; tail -2 gpurun_out/r2s_pytest.log
python - <<'PY'
import json
for f in ("r2s_bench", "r2s_bench_m2"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json")); e = d["e2e"]
        print(f, round(d["ms_per_step"], 3), e["members"], round(e["ms_per_step"], 2), round(e["serial_ms_per_step"], 2))
    except Exception as ex: print(f, "failed", ex)
PY
