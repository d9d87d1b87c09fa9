#!/bin/bash
# round 2, last 1-GPU visit with the final build: whole GPU suite, full bench line, per-kernel table, ncu launch list, per-launch
# metrics of one step, --set full capture of the four heaviest kernels with source-level stall samples of the top two
tag=${1:-r2y}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > gpurun_out/${tag}_gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log; tail -3 gpurun_out/${tag}_pytest.log
timeout 600 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 10 > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err; echo "reference arm rc=$?"
timeout 300 python tools/quick_bench.py 40962 55 20 > gpurun_out/${tag}_kernels.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${tag}_ncu_bench.log 2>&1; echo "ncu launches rc=$?"
tools/ncu_step_metrics.sh ${tag} > /dev/null 2>&1; echo "ncu step metrics rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:"k5_flux_cell|k6_acoustic_cell|k2_dt_edge_b|k7_dt_cell_f" -s 40 -c 8 -o /tmp/${tag}_full -f \
  python tools/quick_bench.py 40962 55 1 > gpurun_out/${tag}_ncu_full.log 2>&1; echo "ncu full rc=$?"
ncu -i /tmp/${tag}_full.ncu-rep --page raw --csv > gpurun_out/${tag}_full_raw.csv 2>/dev/null
ncu -i /tmp/${tag}_full.ncu-rep --page source --csv -k regex:k2_dt_edge_b > gpurun_out/${tag}_source_edge_b.csv 2>/dev/null
ncu -i /tmp/${tag}_full.ncu-rep --page source --csv -k regex:k6_acoustic_cell > gpurun_out/${tag}_source_k6.csv 2>/dev/null
python - <<PY
import json
for f in ("${tag}_bench", "${tag}_bench_ref"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json")); e = d.get("e2e") or {}
        print(f, round(d["ms_per_step"], 3), e.get("ms_per_step"), d.get("parity_rel_l2"), (d.get("roofline") or {}).get("kernel"), (d.get("roofline") or {}).get("frac"))
    except Exception as ex: print(f, "failed", ex)
PY
ls -la gpurun_out/${tag}_*
