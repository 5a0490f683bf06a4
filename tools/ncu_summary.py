#!/usr/bin/env python
"""Summaries of ncu outputs for profiles/: (1) launch-list CSV -> per-kernel time share,
(2) .ncu-rep -> key metrics per captured kernel (via `ncu -i ... --page raw --csv`)."""
import collections
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.avg", "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum",
        "smsp__cycles_active.avg", "sm__inst_executed_pipe_xu.sum", "launch__grid_size", "launch__block_size"]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = [i for i, r in enumerate(rows) if r[0] == "ID"][0]
    h, data = rows[hdr], rows[hdr + 1:]
    ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in data:
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] == "ns" else v * 1e3 if r[ui] == "ms" else v
        agg[r[ki][:90]][0] += 1
        agg[r[ki][:90]][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"# {path}: {len(data)} launches, {tot/1e3:.3f} ms total (cold-cache, serialised: compare shares)")
    for k, v in sorted(agg.items(), key=lambda x: -x[1][1])[:20]:
        print(f"{v[1]/1e3:10.3f} ms {v[0]:4d}x {100*v[1]/tot:5.1f}%  avg {v[1]/v[0]/1e3:8.3f} ms  {k}")


def report(path, keys=KEYS):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h = rows[0]
    units = rows[1]
    name_i = h.index("Kernel Name")
    for r in rows[2:]:
        print(f"## {r[name_i][:100]}")
        for k in keys:
            for i, col in enumerate(h):
                if col == k or col.startswith(k):
                    print(f"   {col:90s} {r[i]:>18s} {units[i]}")


if __name__ == "__main__":
    for p in sys.argv[1:]:
        (report if p.endswith(".ncu-rep") else launches)(p)
