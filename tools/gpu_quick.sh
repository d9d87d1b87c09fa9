#!/bin/bash
# quick A/B on the GPU box: parity of the step + per-kernel table for the default library and for each variant given
mkdir -p gpurun_out
tag=$1; shift
timeout 900 python -m pytest tests/test_parity_gpu.py -x -q -k "one_step or every_routine or fused or reconstruct" > gpurun_out/${tag}_pytest.log 2>&1; tail -3 gpurun_out/${tag}_pytest.log
timeout 300 python tools/quick_bench.py 40962 55 20 > gpurun_out/${tag}_kernels.txt 2>&1; grep -E "ms/step |$KPAT|Error|error" gpurun_out/${tag}_kernels.txt | head -8
for v in "$@"; do
  echo "=== variant $v"
  MPASB_LIB=$PWD/mpas_model_b200/csrc/libmpasb_$v.so timeout 300 python tools/quick_bench.py 40962 55 20 > gpurun_out/${tag}_kernels_$v.txt 2>&1
  grep -E "ms/step |$KPAT|Error|error" gpurun_out/${tag}_kernels_$v.txt | head -8
done
