"""Builds lib/libgwbp.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python 3dgs-gradient-backprojection_b200/build.py [--force] [--verbose]

One translation unit per stage; project.cu is compiled with -fmad=false because its results are
compared bit-exact with the oracle (see the header of that file)."""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
OBJDIR = os.path.join(HERE, "build")
SO = os.path.join(LIBDIR, "libgwbp.so")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC,-ffp-contract=off", "--expt-relaxed-constexpr",
          "-Xptxas", "-v"]
SOURCES = {
    "api.cu": [],
    "project.cu": ["-fmad=false", "-prec-div=true", "-prec-sqrt=true"],
    "binning.cu": [],
    "composite_simt.cu": [],
    "finalize.cu": [],
    "sh.cu": [],
    "backproject_tc.cu": [],
    "backproject_lr.cu": [],
    "render_tc.cu": [],
}


def _stale(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, experiments: bool = False) -> str:
    """experiments=True builds lib/libgwbp_exp.so with -DGWBP_EXPERIMENTS (timing knobs read from GWBP_*_DEBUG
    environment variables: they alter results and exist in no other build); tools/ use it, the package never does
    unless GWBP_LIB_VARIANT=exp is set."""
    global OBJDIR, SO
    if experiments:
        OBJDIR = os.path.join(HERE, "build", "exp")
        SO = os.path.join(LIBDIR, "libgwbp_exp.so")
    else:
        OBJDIR = os.path.join(HERE, "build")
        SO = os.path.join(LIBDIR, "libgwbp.so")
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(OBJDIR, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(HERE, "..", "include", "gwbp.h"))
    headers.append(os.path.abspath(__file__))
    jobs = []
    for src, extra in SOURCES.items():
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJDIR, src.replace(".cu", ".o"))
        if force or _stale(o, [s] + headers):
            jobs.append((["nvcc", "-c", s, "-o", o] + ARCH + COMMON + extra + (["-DGWBP_EXPERIMENTS"] if experiments else []), o))

    def run(job):
        cmd, o = job
        r = subprocess.run(cmd, capture_output=True, text=True)
        log = r.stdout + r.stderr
        with open(o + ".log", "w") as f:
            f.write(" ".join(cmd) + "\n" + log)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {o}:\n{log}")
        if verbose:
            print(log)
        return o

    if jobs:
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
            list(ex.map(run, jobs))
    objs = [os.path.join(OBJDIR, s.replace(".cu", ".o")) for s in SOURCES]
    if force or jobs or _stale(SO, objs):
        cmd = ["nvcc", "-shared", "-o", SO] + objs + ARCH + ["-cudart", "static"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, experiments="--experiments" in sys.argv))
