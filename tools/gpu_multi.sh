#!/bin/bash
# N-GPU visit: NCCL parity test, then bench.py with and without exchange/compute overlap.  Usage: tools/gpu_multi.sh <tag> <N>
tag=$1; N=$2
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multigpu.py -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; tail -3 gpurun_out/${tag}_pytest.log
run() { timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 "$@"; }
run > gpurun_out/${tag}_bench_${N}gpu.json 2> gpurun_out/${tag}_bench_${N}gpu.err; echo "bench rc=$?"
MPASB_NO_OVERLAP=1 run --no-e2e > gpurun_out/${tag}_bench_${N}gpu_nooverlap.json 2> gpurun_out/${tag}_bench_${N}gpu_nooverlap.err; echo "bench(no overlap) rc=$?"
python - <<PY
import json
for f in ("gpurun_out/${tag}_bench_${N}gpu.json", "gpurun_out/${tag}_bench_${N}gpu_nooverlap.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); print(f, d["ms_per_step"], d["value"], d.get("e2e"))
    except Exception as e: print(f, "unreadable", e)
PY
