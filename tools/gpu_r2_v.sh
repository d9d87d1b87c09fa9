#!/bin/bash
# round 2, call v: compile-time sweep directions, regional path on the column-warp kernels; same-box comparison with the build of
# commit ce8ebde (before the alternating sweep / dependent launch work), checked out under build/old
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2v_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2v_pytest.log; tail -3 gpurun_out/r2v_pytest.log
B="python bench.py --steps 40 --no-e2e --no-cpu-baseline"
timeout 300 $B > gpurun_out/r2v_bench.json 2> gpurun_out/r2v_bench.err
(cd build/old && timeout 300 $B > ../../gpurun_out/r2v_bench_old.json 2> ../../gpurun_out/r2v_bench_old.err)
MPASB_PDL=0 timeout 300 $B > gpurun_out/r2v_bench_nopdl.json 2> gpurun_out/r2v_bench_nopdl.err
timeout 300 $B > gpurun_out/r2v_bench_b.json 2> gpurun_out/r2v_bench_b.err
(cd build/old && timeout 300 $B > ../../gpurun_out/r2v_bench_old_b.json 2> ../../gpurun_out/r2v_bench_old_b.err)
python - <<'PY'
import json
for f in ("r2v_bench", "r2v_bench_old", "r2v_bench_nopdl", "r2v_bench_b", "r2v_bench_old_b"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json")); print(f, round(d["ms_per_step"], 3))
        print("   ", {k: v for k, v in list(d["kernel_ms_per_step"].items())[:14]})
    except Exception as e: print(f, "failed", e)
PY
