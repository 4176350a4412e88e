#!/usr/bin/env python
"""Turns the ncu outputs a gpurun call brought back into the small tracked files under profiles/.

    python tools/ncu_summary.py launches <launches.csv> <out.txt> "<command line / note>"
    python tools/ncu_summary.py kernel   <report.ncu-rep> <out.txt> "<command line / note>" [traffic.json n L]

`launches`: per-kernel totals and shares of an `ncu --metrics gpu__time_duration.sum --csv` log.
`kernel`  : the metrics DESIGN.md quotes from one `ncu --set full` capture (first kernel in the report);
            with the optional arguments also rewrites profiles/umma_traffic.json, which bench.py reads."""
import collections
import csv
import io
import json
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_read.sum.per_second", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "l1tex__m_xbar2l1tex_read_bytes.sum.per_second", "l1tex__m_xbar2l1tex_read_bytes.sum.pct_of_peak_sustained_elapsed",
        "launch__block_size", "launch__grid_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.avg.per_second",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active"]


def launches(src, dst, note):
    lines = [l for l in open(src) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v *= {"ns": 1e-6, "nsecond": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3}[unit]
        k = row["Kernel Name"]
        agg[k][0] += 1
        agg[k][1] += v
    total = sum(v[1] for v in agg.values())
    mine = sum(v[1] for k, v in agg.items() if "btb::" in k)
    with open(dst, "w") as f:
        f.write("# ncu launch list summary: %s\n" % note)
        f.write("# per-launch times are cold-cache and serialised: compare SHARES, not absolute times\n")
        f.write("%-100s %6s %12s %7s\n" % ("kernel", "count", "total_ms", "share"))
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("%-100s %6d %12.3f %6.1f%%\n" % (k[:100], v[0], v[1], 100 * v[1] / total))
        f.write("total %.1f ms over %d launches\n" % (total, sum(v[0] for v in agg.values())))
        f.write("this library's kernels (btb::*): %.1f ms = %.1f %% of all launches; shares among them:\n" % (mine, 100 * mine / total))
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            if "btb::" in k:
                f.write("  %-98s %6.1f%%\n" % (k[:98], 100 * v[1] / mine))


def kernel(rep, dst, note, traffic=None):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    name = vals[hdr.index("Kernel Name")]
    got = {}
    for h, u, v in zip(hdr, units, vals):
        key = h.split("TriageCompute.")[-1]
        if key in WANT:
            got[key] = (u, v)
    with open(dst, "w") as f:
        f.write("# %s\n# kernel: %s\n" % (note, name))
        for k in sorted(got):
            f.write("%-90s %-16s %s\n" % (k, got[k][0], got[k][1]))
        rd = float(got["dram__bytes_read.sum"][1]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Tbyte": 1e12, "byte": 1.0, "Kbyte": 1e3}[got["dram__bytes_read.sum"][0]]
        wr = float(got["dram__bytes_write.sum"][1]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Tbyte": 1e12, "byte": 1.0, "Kbyte": 1e3}[got["dram__bytes_write.sum"][0]]
        f.write("derived: DRAM traffic %.2f GB read + %.2f GB written per launch\n" % (rd / 1e9, wr / 1e9))
    if traffic:
        path, n, L = traffic
        json.dump({"n": int(n), "L": int(L), "n_gpus": 1, "kernel": name, "dram_bytes_per_launch": rd + wr,
                   "dram_bytes_read": rd, "dram_bytes_write": wr, "source": note}, open(path, "w"), indent=1)


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], sys.argv[4])
    else:
        kernel(sys.argv[2], sys.argv[3], sys.argv[4], sys.argv[5:8] if len(sys.argv) >= 8 else None)
