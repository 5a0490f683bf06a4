#!/usr/bin/env python
"""profiles/r02_traffic.json from the ncu --set full summaries under profiles/: per-launch DRAM traffic and tensor-pipe
share of the dominant kernel, stamped with the hash of the CUDA sources they were captured from (bench.py reports them
only while that hash still matches).  Run right after the capture, before touching csrc/."""
import json, os, re, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench


def grab(path):
    txt = open(os.path.join(ROOT, "profiles", path)).read()
    def val(key):
        m = re.search(r"^\s*" + re.escape(key) + r"\s+([0-9.]+)\s+(\S+)", txt, re.M)
        v, unit = float(m.group(1)), m.group(2)
        return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(unit, 1.0)
    def val_ms(key):
        m = re.search(r"^\s*" + re.escape(key) + r"\s+([0-9.]+)\s+(\S+)", txt, re.M)
        return float(m.group(1)) * {"ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}.get(m.group(2), 1.0)
    return {"traffic": val("dram__bytes_read.sum") + val("dram__bytes_write.sum"),
            "tensor_pipe_pct": float(re.search(r"sm__pipe_tensor_cycles_active\S*\s+([0-9.]+)", txt).group(1)),
            "kernel_ms_under_ncu": val_ms("gpu__time_duration.sum"), "file": "profiles/" + path}


out = {"source_sha": bench._source_sha(), "G:full:tc": grab("r02_bp_tc_full_summary.txt"), "G:lowres:tc": grab("r02_bp_lr_full_summary.txt")}
json.dump(out, open(os.path.join(ROOT, "profiles", "r02_traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
