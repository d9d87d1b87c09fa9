#!/bin/bash
# round 2, call ac: whole GPU suite with the batched monotonic transport, mpasb_init_block, TMA-staged flux weights;
# 21-scalar configuration batched against one scalar at a time
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/ac_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/ac_pytest.log; tail -5 gpurun_out/ac_pytest.log
B="python bench.py --steps 20 --no-e2e --no-cpu-baseline --scalars 21"
timeout 400 $B > gpurun_out/ac_bench_s21.json 2> gpurun_out/ac_bench_s21.err
MPASB_MONO_BATCH=0 timeout 400 $B > gpurun_out/ac_bench_s21_unbatched.json 2> gpurun_out/ac_bench_s21_unbatched.err
python - <<'PY'
import json
for f in ("ac_bench_s21", "ac_bench_s21_unbatched"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json")); print(f, round(d["ms_per_step"], 3), d.get("gpu_launches"))
        print("   ", {k: v for k, v in d["kernel_ms_per_step"].items() if "mono" in k or "scalars" in k})
    except Exception as e: print(f, "failed", e)
PY
