#!/bin/bash
# round 2, call aj (2 GPUs): the NCCL case of the multi-rank parity test again (call ai lost it to a port collision in the launcher)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multigpu.py -m gpu -q -k "2gpus" > gpurun_out/r2aj_pytest_multigpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2aj_pytest_multigpu.log; tail -3 gpurun_out/r2aj_pytest_multigpu.log
