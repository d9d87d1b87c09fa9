#!/bin/bash
# round 2, call k (8 GPUs): 8-rank parity, IPC put/get and NCCL
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2k_gpus.txt
timeout 900 python -m pytest tests/test_multigpu.py -m gpu -v -k "8gpus" > gpurun_out/r2k_pytest_multigpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2k_pytest_multigpu.log
tail -8 gpurun_out/r2k_pytest_multigpu.log
