#!/usr/bin/env python
"""Summarise ncu output brought back in gpurun_out/ into small tracked files under profiles/.

    python tools/ncu_summary.py launches gpurun_out/r1_launches.csv profiles/r1_launches.md
    python tools/ncu_summary.py full gpurun_out/r1_conv_tc_kernel.ncu-rep profiles/r1_conv_tc_kernel.json

`launches` reads the `--metrics gpu__time_duration.sum` CSV (per-launch device time, cold cache, serialised:
compare SHARES); `full` reads one `--set full` report through `ncu -i ... --page raw --csv`.
"""
import csv
import json
import subprocess
import sys
from collections import OrderedDict, defaultdict

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__waves_per_multiprocessor", "launch__shared_mem_per_block_dynamic",
    "sm__cycles_elapsed.max", "sm__cycles_elapsed.max.per_second",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.pct", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
]


def launches(src, dst):
    rows = list(csv.reader(open(src, errors="replace")))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = defaultdict(lambda: [0, 0.0])
    for r in data:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui], 1.0)
        name = r[ki].replace("void ", "").replace("<unnamed>::", "")
        name = name.split("(")[0][:70]
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v for _, v in agg.values())
    with open(dst, "w") as f:
        f.write("| kernel | launches | total ms | avg ms | share |\n|---|---:|---:|---:|---:|\n")
        for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("| `%s` | %d | %.3f | %.4f | %.1f%% |\n" % (k, n, v, v / n, 100 * v / tot))
        f.write("\nsource: %s (ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache serialised launches)\n" % src)
    print(open(dst).read())


def full(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    res = []
    for vals in rows[2:]:
        d = OrderedDict()
        d["kernel"] = vals[hdr.index("Kernel Name")][:120]
        for i, h in enumerate(hdr):
            if h in KEEP:
                d[h] = "%s %s" % (vals[i], units[i])
        res.append(d)
    json.dump({"source": src, "launches": res}, open(dst, "w"), indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
