#!/bin/bash
# round 2, call am: the final build once more -- whole GPU suite (the file-start and init-block tests now derive coeffs_reconstruct in
# the library) and compute-sanitizer over every code path
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/am_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/am_pytest.log; tail -3 gpurun_out/am_pytest.log
timeout 400 compute-sanitizer --tool memcheck python tools/sanitize.py > gpurun_out/am_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -2 gpurun_out/am_sanitizer_memcheck.log
timeout 400 compute-sanitizer --tool racecheck python tools/sanitize.py > gpurun_out/am_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -2 gpurun_out/am_sanitizer_racecheck.log
