#!/bin/bash
# round 2, call ak: compute-sanitizer over the code paths added in the last session (tools/sanitize.py new)
mkdir -p gpurun_out
timeout 420 compute-sanitizer --tool memcheck python tools/sanitize.py new > gpurun_out/ak_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/ak_sanitizer_memcheck.log
timeout 420 compute-sanitizer --tool racecheck python tools/sanitize.py new > gpurun_out/ak_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -3 gpurun_out/ak_sanitizer_racecheck.log
