"""Summaries of ncu output for profiles/.
  python tools/ncu_summary.py launches <launches.csv>          per-kernel totals of a gpu__time_duration launch list
  python tools/ncu_summary.py full <report.ncu-rep>            key metrics per captured kernel of a --set full report"""
import collections, csv, io, subprocess, sys

def launches(path):
    rows = [r for r in csv.reader(open(path, errors="ignore")) if len(r) > 10]
    hdr = rows[0]; ik = hdr.index("Kernel Name"); iv = hdr.index("Metric Value"); iu = hdr.index("Metric Unit")
    tot = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        v = float(r[iv].replace(",", "")); v = v / 1e3 if r[iu] in ("ns", "nsecond") else v
        name = r[ik].split("(")[0]; tot[name][0] += 1; tot[name][1] += v
    s = sum(v for _, v in tot.values())
    print(f"{sum(n for n, _ in tot.values())} launches, total {s/1e3:.3f} ms (cold-cache, serialised: compare shares)")
    for k, (n, v) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:30s} {n:4d} launches {v/1e3:9.3f} ms {100*v/s:6.2f} % {v/n:9.1f} us/launch")

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]

def full(path):
    # a .ncu-rep report, or the `ncu -i report --page raw --csv` text already made on the GPU box
    out = open(path, errors="ignore").read() if path.endswith(".csv") else subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out))); hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in data:
        print("----"); print("  Kernel Name", r[idx["Kernel Name"]])
        for w in WANT:
            if w in idx: print(f"  {w:82s} {r[idx[w]]:>16s} {units[idx[w]]}")

if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
