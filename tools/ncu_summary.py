"""Summarise ncu captures into the tracked profiles/ directory (the .ncu-rep files are scratch).

    python tools/ncu_summary.py launches gpurun_out/launches_r1.csv profiles/r1_launches.txt
    python tools/ncu_summary.py kernel gpurun_out/prof_k1_r1.ncu-rep profiles/r1_k1_ncu.txt [traffic.json|-] [name substring]
"""
import collections
import csv
import json
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.avg.per_second", "lts__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"]


def launches(src, dst):
    lines = [l for l in open(src) if not l.startswith("==")]
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
        name = row["Kernel Name"][:90]
        tot[name] += v
        cnt[name] += 1
    s = sum(tot.values())
    with open(dst, "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)\n")
        f.write("# command: python bench.py --steps 2 --warmup 1   (3 rollouts of 34 steps)\n")
        f.write("%-92s %6s %12s %10s %7s\n" % ("kernel", "n", "total_us", "avg_us", "share"))
        for k, v in sorted(tot.items(), key=lambda x: -x[1]):
            f.write("%-92s %6d %12.1f %10.1f %6.1f%%\n" % (k, cnt[k], v, v / cnt[k], 100 * v / s))
        f.write("total_us %.1f\n" % s)


def kernel(rep, dst, traffic=None, match=None):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    vals = next(r for r in rows[2:] if match is None or match in r[ki])   # first captured launch whose name matches
    d = {h: (units[i], vals[i]) for i, h in enumerate(hdr)}
    with open(dst, "w") as f:
        f.write("# ncu --set full --clock-control none, one launch; source: %s\n" % rep)
        f.write("kernel: %s\n" % d.get("Kernel Name", ("", ""))[1])
        for k in KEYS:
            if k in d:
                f.write("%-80s %-16s %s\n" % (k, d[k][0], d[k][1]))
    if traffic:
        def tobytes(key):
            u, v = d[key]
            v = float(v.replace(",", ""))
            return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
        json.dump({"kernel": d["Kernel Name"][1], "dram_bytes_per_launch": tobytes("dram__bytes_read.sum") + tobytes("dram__bytes_write.sum"),
                   "dram_read": tobytes("dram__bytes_read.sum"), "dram_write": tobytes("dram__bytes_write.sum"),
                   "source": "ncu --set full --clock-control none, bench.py --steps 2 --warmup 1"}, open(traffic, "w"), indent=1)


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        kernel(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 and sys.argv[4] != "-" else None,
               sys.argv[5] if len(sys.argv) > 5 else None)
