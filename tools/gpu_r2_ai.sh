#!/bin/bash
# round 2, call ai (2 GPUs): multi-rank parity and the 2-GPU bench line with the final build
N=2
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multigpu.py -m gpu -q > gpurun_out/r2ai_pytest_multigpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2ai_pytest_multigpu.log; tail -3 gpurun_out/r2ai_pytest_multigpu.log
run() { timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 30 --warmup 3 "$@"; }
run > gpurun_out/r2ai_bench_2gpu.json 2> gpurun_out/r2ai_bench_2gpu.err; echo "bench rc=$?"
python - <<'PY'
import json
f = "gpurun_out/r2ai_bench_2gpu.json"
try:
    d = json.loads(open(f).read().strip().splitlines()[-1]); e = d.get("e2e") or {}
    print(f, round(d["ms_per_step"], 3), round(d["value"]), e.get("ms_per_step"), e.get("serial_ms_per_step"), d["step_roofline"]["frac_of_measured_peak"])
except Exception as ex: print(f, "unreadable", ex)
PY
tail -2 gpurun_out/r2ai_bench_2gpu.err
