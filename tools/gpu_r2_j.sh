#!/bin/bash
# round 2, call j (4 GPUs): multi-rank parity matrix (2 and 4 ranks; IPC put/get and NCCL; with and without overlap)
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2j_gpus.txt
timeout 1500 python -m pytest tests/test_multigpu.py -m gpu -v -k "2gpus or 4gpus" > gpurun_out/r2j_pytest_multigpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2j_pytest_multigpu.log
tail -12 gpurun_out/r2j_pytest_multigpu.log
