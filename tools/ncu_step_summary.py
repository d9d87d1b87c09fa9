"""Per-kernel averages of the per-launch metric table written by tools/ncu_step_metrics.sh (one whole step under ncu).
  python tools/ncu_step_summary.py gpurun_out/<tag>_step_metrics.csv [nCells nLevels] > profiles/<tag>_ncu_step_metrics_summary.txt"""
import collections, csv, sys

path = sys.argv[1]
ncells, nlev = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (40962, 55)
C = ncells * nlev * 8 / 1e6
rows = [r for r in csv.reader(open(path, errors="ignore")) if len(r) > 10]
hdr = rows[0]
ik, im, iv, iu, iid = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("ID")
per = collections.OrderedDict()
for r in rows[1:]:
    key = (r[iid], r[ik].split("(")[0])
    v = float(r[iv].replace(",", "") or 0)
    u = r[iu]
    if u in ("ns", "nsecond"): v /= 1e3
    if u == "Kbyte": v *= 1e3
    if u == "Mbyte": v *= 1e6
    if u == "Gbyte": v *= 1e9
    per.setdefault(key, {})[r[im]] = v
agg = collections.defaultdict(list)
for (_, name), m in per.items():
    agg[name].append(m)
def avg(ms, k): return sum(m.get(k, 0.0) for m in ms) / len(ms)
print(f"per-launch averages of one whole step under ncu ({path}); MB = dram__bytes_read.sum + dram__bytes_write.sum; C = {C:.2f} MB;")
print("dram% = gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed (a pure copy, k_segments, reaches ~69 %); times are cold-cache and serialised")
print(f"{'kernel':26s} {'n':>3s} {'us':>7s} {'MB':>7s} {'C':>5s} {'GB/s':>6s} {'dram%':>6s} {'occ%':>5s} {'regs':>4s} {'L1pipe%':>7s} {'L1hit':>5s} {'L2hit':>5s} {'issue%':>6s} {'lsb/iss':>7s}")
tot_us = tot_mb = n_all = 0
for name, ms in sorted(agg.items(), key=lambda kv: -sum(m.get("gpu__time_duration.sum", 0) for m in kv[1])):
    us = avg(ms, "gpu__time_duration.sum"); mb = (avg(ms, "dram__bytes_read.sum") + avg(ms, "dram__bytes_write.sum")) / 1e6
    tot_us += us * len(ms); tot_mb += mb * len(ms); n_all += len(ms)
    print(f"{name[:26]:26s} {len(ms):3d} {us:7.1f} {mb:7.1f} {mb / C:5.1f} {mb / us * 1e3 if us else 0:6.0f} "
          f"{avg(ms, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):6.1f} {avg(ms, 'sm__warps_active.avg.pct_of_peak_sustained_active'):5.1f} "
          f"{avg(ms, 'launch__registers_per_thread'):4.0f} {avg(ms, 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed'):7.1f} "
          f"{avg(ms, 'l1tex__t_sector_hit_rate.pct'):5.1f} {avg(ms, 'lts__t_sector_hit_rate.pct'):5.1f} {avg(ms, 'smsp__issue_active.avg.pct_of_peak_sustained_active'):6.1f} "
          f"{avg(ms, 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio'):7.2f}")
print(f"{n_all} launches: {tot_us / 1e3:.2f} ms, {tot_mb / 1e3:.1f} GB of DRAM traffic ({tot_mb / C:.0f} C)")
