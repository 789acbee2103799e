"""Summaries of ncu outputs for profiles/ (development aid).
  ncu_summary.py launches <csv>            -> per-kernel totals / shares of a `--metrics gpu__time_duration.sum` launch list
  ncu_summary.py full <ncu-rep> <out.csv>  -> selected metrics of a `--set full` capture, one row per launch"""
import collections
import csv
import io
import re
import subprocess
import sys

mode = sys.argv[1]
if mode == "launches":
    rows = [r for r in csv.reader(open(sys.argv[2])) if len(r) > 10]
    h = rows[0]
    ik, iv, iu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    tot, cnt = collections.Counter(), collections.Counter()
    for r in rows[1:]:
        try:
            v = float(r[iv].replace(",", ""))
        except ValueError:
            continue
        scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[iu], 1e-6)
        name = re.sub(r"\(.*", "", r[ik])
        tot[name] += v * scale
        cnt[name] += 1
    total = sum(tot.values())
    print(f"{len(rows) - 1} launches, {total:.1f} ms of device time")
    for k, v in tot.most_common(12):
        print(f"  {v:9.2f} ms  {100 * v / total:5.1f} %  {cnt[k]:5d} launches  {v / cnt[k]:8.3f} ms/launch  {k[:90]}")
else:
    raw = subprocess.run(["ncu", "-i", sys.argv[2], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h = rows[0]
    want = ["launch__grid_size", "launch__block_size", "launch__registers_per_thread", "gpu__time_duration.sum",
            "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum"]
    with open(sys.argv[3], "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["Kernel Name"] + want)
        w.writerow([""] + [rows[1][h.index(k)] if k in h else "" for k in want])
        for r in rows[2:]:
            w.writerow([r[h.index("Kernel Name")]] + [r[h.index(k)] if k in h else "" for k in want])
            print(r[h.index("Kernel Name")][:60], " ".join(f"{k.split('.')[0][-18:]}={r[h.index(k)]}" for k in want[3:12] if k in h))
