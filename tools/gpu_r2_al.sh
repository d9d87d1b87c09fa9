#!/bin/bash
# round 2, call al (8 GPUs): BASELINE config 4b -- x1.655362 x 55 levels, PRECISION=single build, one block per GPU
N=8
mkdir -p gpurun_out
run() { timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 "$@"; }
run --precision single --no-cpu-baseline > gpurun_out/r2al_bench_8gpu_fp32.json 2> gpurun_out/r2al_bench_8gpu_fp32.err; echo "bench fp32 rc=$?"
python - <<'PY'
import json
f = "gpurun_out/r2al_bench_8gpu_fp32.json"
try:
    d = json.loads(open(f).read().strip().splitlines()[-1]); e = d.get("e2e") or {}
    print(f, round(d["ms_per_step"], 3), round(d["value"]), e.get("ms_per_step"), d["step_roofline"]["frac_of_measured_peak"], d["config"]["workload"])
except Exception as ex: print(f, "unreadable", ex)
PY
tail -2 gpurun_out/r2al_bench_8gpu_fp32.err
