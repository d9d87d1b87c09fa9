#!/bin/bash
# round 2, call n: regional (LBC) GPU tests + the whole GPU suite, full bench line, per-kernel table, ncu launch list and step metrics
tools/gpu_round.sh r2n
tools/ncu_step_metrics.sh r2n
