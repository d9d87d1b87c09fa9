#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/e2e_probe.py > gpurun_out/r2r_e2e_probe.txt 2>&1
cat gpurun_out/r2r_e2e_probe.txt
