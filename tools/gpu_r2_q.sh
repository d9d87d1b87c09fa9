#!/bin/bash
# round 2, call q: whole GPU suite (regional free-running test, file-start test), full bench line with the 3-member e2e pipeline, 2-member line beside it
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2q_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2q_pytest.log
tail -3 gpurun_out/r2q_pytest.log
timeout 600 python bench.py > gpurun_out/r2q_bench.json 2> gpurun_out/r2q_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --steps 20 --no-cpu-baseline --e2e-members 2 > gpurun_out/r2q_bench_m2.json 2> gpurun_out/r2q_bench_m2.err
timeout 300 python bench.py --steps 20 --no-cpu-baseline --e2e-members 4 > gpurun_out/r2q_bench_m4.json 2> gpurun_out/r2q_bench_m4.err
python - <<'PY'
import json
for f in ("r2q_bench", "r2q_bench_m2", "r2q_bench_m4"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json")); e = d["e2e"]
        print(f, round(d["ms_per_step"], 3), e["members"], round(e["ms_per_step"], 2), round(e["serial_ms_per_step"], 2), d.get("host_affinity"), d["parity_rel_l2"])
    except Exception as ex: print(f, "failed", ex)
PY
