#!/bin/bash
# N-GPU visit: parity test (2 GPUs), then bench.py with the exchange variants.  Usage: tools/gpu_multi.sh <tag> <N>
tag=$1; N=$2
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multigpu.py -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; tail -3 gpurun_out/${tag}_pytest.log
MPASB_P2P=0 timeout 600 python -m pytest tests/test_multigpu.py -m gpu -x -q 2>&1 | tail -1
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 "$@"; }
run > gpurun_out/${tag}_bench_${N}gpu.json 2> gpurun_out/${tag}_bench_${N}gpu.err; echo "bench (p2p + overlap) rc=$?"
MPASB_P2P=0 run --no-e2e > gpurun_out/${tag}_bench_${N}gpu_nccl.json 2> gpurun_out/${tag}_bench_${N}gpu_nccl.err; echo "bench (nccl + overlap) rc=$?"
MPASB_NO_OVERLAP=1 run --no-e2e > gpurun_out/${tag}_bench_${N}gpu_p2p_nooverlap.json 2> gpurun_out/${tag}_bench_${N}gpu_p2p_nooverlap.err; echo "bench (p2p, no overlap) rc=$?"
python - <<PY
import json
for v in ("", "_nccl", "_p2p_nooverlap"):
    f = "gpurun_out/${tag}_bench_${N}gpu%s.json" % v
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); print(f, d["ms_per_step"], d["value"], d.get("e2e"))
    except Exception as e: print(f, "unreadable", e)
PY
tail -3 gpurun_out/${tag}_bench_${N}gpu.err
