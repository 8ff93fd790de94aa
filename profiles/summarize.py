"""Turns ncu outputs (gpurun_out/) into the small tracked summaries under profiles/.

  python profiles/summarize.py launches gpurun_out/launches_r1.csv profiles/r1_launches_summary.csv
  python profiles/summarize.py raw gpurun_out/prof_iterate_r1.ncu-rep profiles/r1_iterate_full.csv
"""
import collections
import csv
import io
import subprocess
import sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__cluster_size",
        "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_tensor.sum", "lts__t_bytes.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]


def launches(src, dst):
    lines = [l for l in open(src) if not l.startswith("==")]
    per = collections.defaultdict(list)
    for row in csv.DictReader(io.StringIO("".join(lines))):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = row["Kernel Name"].split("(")[0].replace("void ", "")
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(row["Metric Unit"], 1.0)
        per[(name, row.get("Grid Size", ""))].append(v)
    total = sum(sum(v) for v in per.values())
    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "grid", "launches", "total_us", "share_of_listed_time", "avg_us", "min_us", "max_us"])
        for (name, grid), v in sorted(per.items(), key=lambda kv: -sum(kv[1])):
            w.writerow([name, grid, len(v), "%.1f" % sum(v), "%.4f" % (sum(v) / total), "%.2f" % (sum(v) / len(v)),
                        "%.2f" % min(v), "%.2f" % max(v)])
    print("wrote", dst)


def raw(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    cols = [i for i, h in enumerate(hdr) if h in KEEP or h in ("Kernel Name", "ID")
            or (h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio")
                and "not_issued" not in h)]
    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + ["launch_%d" % k for k in range(len(data))])
        for i in cols:
            w.writerow([hdr[i], units[i]] + [r[i] for r in data])
    print("wrote", dst)


if __name__ == "__main__":
    {"launches": launches, "raw": raw}[sys.argv[1]](sys.argv[2], sys.argv[3])
