#!/bin/bash
# round 2, call u (4 GPUs): multi-rank parity with the partial damping fold, 4-GPU bench with / without it, e2e with the enqueued init exchange
N=4
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multigpu.py -m gpu -q > gpurun_out/r2u_pytest_multigpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2u_pytest_multigpu.log; tail -3 gpurun_out/r2u_pytest_multigpu.log
run() { timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 30 --warmup 3 "$@"; }
run > gpurun_out/r2u_bench_4gpu.json 2> gpurun_out/r2u_bench_4gpu.err; echo "bench rc=$?"
MPASB_NO_DD_PARTIAL=1 run --no-e2e > gpurun_out/r2u_bench_4gpu_nopartial.json 2> gpurun_out/r2u_bench_4gpu_nopartial.err; echo "bench (full damping kernel) rc=$?"
run --no-e2e > gpurun_out/r2u_bench_4gpu_b.json 2> gpurun_out/r2u_bench_4gpu_b.err; echo "bench (again) rc=$?"
python - <<'PY'
import json
for v in ("", "_nopartial", "_b"):
    f = "gpurun_out/r2u_bench_4gpu%s.json" % v
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); e = d.get("e2e") or {}
        print(f, round(d["ms_per_step"], 3), round(d["value"]), e.get("ms_per_step"), e.get("serial_ms_per_step"))
    except Exception as ex: print(f, "unreadable", ex)
PY
