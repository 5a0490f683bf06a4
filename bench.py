#!/usr/bin/env python
"""Benchmark of the back-projection hot path (BASELINE.json metric: back-projected views/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config G] [--kernel auto|simt|tc]
    python bench.py --impl reference ...          # the CPU port of the path, timed on host cores

A "step" is ONE VIEW: project + bin + sort + fused composite/contract/accumulate of one camera
against a feature map resident in HBM.  Workload at N=1: BASELINE config[1]
("G": 5.8M Gaussians, 1297x840, 512-d LSeg-shaped features, synthetic, seed 0).  With N>1 every
rank back-projects its own K views (weak scaling, views r, r+N, ...) into its own full
accumulators and ONE all-reduce of (num, den) closes the timed region.

Prints ONE JSON line (rank 0).  Keys follow the driver contract; `roofline` describes the fused
kernel (HBM-bound: algorithmic bytes / CUDA-event time) and carries `stages` (every stage of a view:
ms by CUDA events, algorithmic bytes, fraction of the measured HBM peak) and `view` (the whole view
against SURVEY.md §8d's byte count); `cpu_baseline` the oracle port on host cores, `e2e` the same
metric through the public API with HOST feature maps (H2D copy + D2H result read inside the timed
region); `shim` the cost of the reference's own loop (3 x rasterization + 2 x backward per view)
on the drop-in operator.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "back-projected views/s"
UNIT = "views/s"


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), float(p.get("bf16_tflops_sustained", p.get("bf16_tflops", 0))), "measured"
    except Exception:
        return 6650.0, 1400.0, "fallback"


def _source_sha() -> str:
    """Identity of the CUDA sources: profiles/*_traffic.json is only trusted for the code it was captured from
    (the GPU box has no .git, so the commit hash is not available there)."""
    import hashlib

    h = hashlib.sha256()
    d = os.path.join(ROOT, "3dgs-gradient-backprojection_b200", "csrc")
    for name in sorted(os.listdir(d)):
        if name.endswith((".cu", ".cuh")):
            with open(os.path.join(d, name), "rb") as f:
                h.update(name.encode() + b"\0" + f.read())
    return h.hexdigest()[:16]


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.proc = None
        self.idx = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle port (plain C, OpenMP) on host cores
# ------------------------------------------------------------------------------------------------
def _cpu_views(cfg, n_views_max: float, budget_s: float, seed=0):
    """Back-project up to n_views_max views of `cfg` with the C oracle; stop when budget_s is spent
    (at least one view).  Returns (views_done, seconds, threads)."""
    import numpy as np
    import torch

    import gwbp
    from oracle import c_oracle

    c_oracle.build()
    c_oracle.set_num_threads(os.cpu_count() or 1)  # torchrun exports OMP_NUM_THREADS=1: ask for every host core
    S = gwbp.scene
    sc = S.make_scene(cfg["n"], seed)
    vm, K = S.make_cameras(cfg["views"], cfg["width"], cfg["height"], seed)
    W, H, d = cfg["width"], cfg["height"], cfg["d"]
    enc = 240 if min(W, H) >= 480 else 24
    feats = S.make_feature_map_torch(0, d, H, W, "cpu", seed, enc_res=enc).numpy()  # permuted planar view
    num = np.zeros((sc.n, d), np.float64)  # calloc: only touched rows are ever committed
    den = np.zeros(sc.n, np.float64)
    done, t0 = 0, time.perf_counter()
    while done < n_views_max:
        v = c_oracle.View(sc.means, sc.quats, sc.scales, sc.opacities, vm[done % cfg["views"]], K, W, H)
        v.backproject(feats, num, den)
        v.close()
        done += 1
        if time.perf_counter() - t0 > budget_s:
            break
    return done, time.perf_counter() - t0, c_oracle.num_threads()


def _time_shim(gwbp, torch, dev, sc, vm, K, W, H, d, pool, n_views):
    """The reference's loop body, unmodified in shape (backproject.py:115-151): rasterization(colors_feats[N,D]) ->
    (out*feats).sum().backward() -> clone/zero_ -> rasterization(colors_feats_0[N,3]) -> out.sum().backward() ->
    accumulate.  INTEGRATION.md §1's zero-edit path: what it costs per view on this engine."""
    t = lambda a: torch.from_numpy(a).to(dev)  # noqa: E731
    means, quats, scales, opac = t(sc.means), t(sc.quats), t(sc.scales), t(sc.opacities)
    Kt = t(K)[None]
    n = sc.n
    gaussian_features = torch.zeros(n, d, device=dev)
    gaussian_denoms = torch.ones(n, device=dev) * 1e-12
    colors_feats = torch.zeros(n, d, device=dev, requires_grad=True)
    colors_feats_0 = torch.zeros(n, 3, device=dev, requires_grad=True)

    def one(v):
        nonlocal gaussian_features, gaussian_denoms
        viewmat = t(vm[v])[None]
        feats = pool[v % len(pool)]
        out, _, _ = gwbp.rasterization(means, quats, scales, opac, colors_feats, viewmat, Kt, width=W, height=H)
        (out[0] * feats).sum().backward()
        copy = colors_feats.grad.clone()
        colors_feats.grad.zero_()
        out0, _, _ = gwbp.rasterization(means, quats, scales, opac, colors_feats_0, viewmat, Kt, width=W, height=H)
        out0[0].sum().backward()
        gaussian_features += copy
        gaussian_denoms += colors_feats_0.grad[:, 0]
        colors_feats_0.grad.zero_()
        del out, out0, copy

    one(0)  # warm-up (allocations, lazy module loads)
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for v in range(1, 1 + n_views):
        one(v)
    torch.cuda.synchronize(dev)
    secs = (time.perf_counter() - t0) / n_views
    gwbp.rasterization_cache_clear()
    return {"ms_per_view": 1e3 * secs, "views_per_s": 1.0 / secs, "views": n_views,
            "what": "2 x rasterization(colors=[N,D] / [N,3] zeros, requires_grad) + 2 x backward + clone/zero_/+= per view "
                    "(backproject.py:115-151 verbatim) on the drop-in operator; projection + binning shared between the "
                    "two calls of a view (scene/view cache); the RGB render that feeds the encoder is excluded as in every "
                    "BASELINE config"}


def run_reference(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # the reference's own stack cannot run here or on the box (gsplat-1.4.0 is un-vendored, CUDA-only and
    # not installable offline: SURVEY.md §0.3) -> the CPU arm is the oracle port, all host threads.
    for _ in range(min(args.warmup, 1)):
        _cpu_views(dict(cfg, n=min(cfg["n"], 200_000)), 1, 0.0)
    done, secs, threads = _cpu_views(cfg, args.steps, args.cpu_budget)
    val = done / secs
    sample = (f"{done} view(s) of config {args.config} ({cfg['n']} Gaussians, {cfg['width']}x{cfg['height']}, "
              f"D={cfg['d']}), feature map resident in host RAM; bounded to ~{args.cpu_budget:.0f} s")
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
            "steps": done, "warmup": min(args.warmup, 1), "ms_per_step": 1e3 * secs / done,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"config {args.config}", **cfg},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args, cfg):
    import numpy as np
    import torch
    import torch.distributed as dist

    import gwbp

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU port")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # Rank 0 must print exactly ONE JSON line on stdout.  Libraries (NCCL's version banner, torchrun) write to
    # fd 1 from native code, so everything else is routed to stderr and the JSON goes to the saved descriptor.
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if world > 1:
        # the caller's NCCL_DEBUG is honoured: fd 1 already points at stderr, so NCCL's INFO lines cannot reach the
        # JSON line on the saved stdout descriptor
        dist.init_process_group("nccl", device_id=dev)
    S = gwbp.scene
    W, H, d, V = cfg["width"], cfg["height"], cfg["d"], cfg["views"]
    sc = S.make_scene(cfg["n"], 0)
    vm, K = S.make_cameras(V, W, H, 0)
    t = lambda a: torch.from_numpy(a).to(dev)  # noqa: E731
    bp = gwbp.BackProjector(t(sc.means), t(sc.quats), t(sc.scales), t(sc.opacities), d, kernel=args.kernel,
                            collect_stats=True)
    if args.overlap_pack >= 0:
        bp.overlap_pack = bool(args.overlap_pack)
    enc = cfg.get("enc", 240 if min(W, H) >= 480 else 24)  # encoder resolution (LSeg 240 x 240; config D: 64 x 64 DINOv2 tokens)
    lmode = cfg.get("mode", "bilinear")                    # how the reference up-samples it (backproject.py:110-112 / :245-249)
    nearest = 1 if lmode == "nearest" else 0
    pool_n = max(1, min(V, args.pool))
    pool = [S.make_feature_map_torch(v, d, H, W, dev, 0, enc_res=enc, mode=lmode) for v in range(pool_n)]
    fmap_bytes = H * W * d * 4
    fpack_dev_bytes = gwbp.fpack_bytes(W, H, d)
    my_view = lambda i: (rank + i * world) % V  # noqa: E731

    low_pool = None
    if args.features == "lowres":  # hand over the encoder-resolution map, upsample fused in the pack pass
        low_pool = [torch.nn.functional.normalize(torch.randn(d, enc, enc, device=dev), dim=0).permute(1, 2, 0)
                    for _ in range(pool_n)]

    def step(i):
        v = my_view(i)
        if low_pool is not None:
            return bp.add_view_lowres(vm[v], K, W, H, low_pool[v % pool_n], mode=lmode)
        return bp.add_view(vm[v], K, W, H, pool[v % pool_n])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for i in range(args.warmup):
        step(i)
    px, collective, peer_note = None, args.collective, None
    if world > 1:  # warm-up of the closing exchange too (NCCL channel setup / buffer registration / IPC mappings are lazy)
        if collective in ("auto", "peer"):
            px = gwbp.dist.peer_exchange_for(bp)  # collective: every rank maps every rank's accumulators (CUDA IPC)
            if px is None:
                if collective == "peer":
                    raise RuntimeError("--collective peer: peer memory could not be set up on this box")
                peer_note = "peer memory unavailable on this box: NCCL reduce-scatter instead"
                collective = "reduce_scatter"
            else:
                collective = "peer"
        if collective == "peer":
            px.reduce_finalize()
        elif collective == "allreduce":
            gwbp.dist.allreduce_accumulators(bp.num, bp.den)
        else:
            gwbp.dist.reduce_scatter_accumulators(bp.num, bp.den)
    bp.reset()
    barrier()

    # ---- timed region: K views (+ the closing all-reduce when N > 1) -------------------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    bp.kernel_events = []
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    barrier()
    launches0 = int(gwbp._lib.lib().gwbp_launch_count())
    e0.record()
    for i in range(args.steps):
        step(args.warmup + i)
    e1.record()
    launches = int(gwbp._lib.lib().gwbp_launch_count()) - launches0  # kernels libgwbp.so launched in the timed loop
    if world > 1:
        if collective == "peer":  # sparse pull over NVLink peer memory + finalise, ONE kernel per rank (dist.PeerExchange)
            px.reduce_finalize()
        elif collective == "allreduce":
            gwbp.dist.allreduce_accumulators(bp.num, bp.den)
        else:  # every rank keeps (and would finalise / save) its own rows of the feature field
            gwbp.dist.reduce_scatter_accumulators(bp.num, bp.den)
    e2.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms_total = e0.elapsed_time(e2)
    ms_views = e0.elapsed_time(e1)
    ms_kernel = sum(a.elapsed_time(b) for a, b in bp.kernel_events) / max(1, len(bp.kernel_events))
    bp.kernel_events = None
    st = bp.stats()
    last = bp.last_view
    if world > 1:
        tt = torch.tensor([ms_total, ms_views], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_total, ms_views = tt.tolist()
    value = world * args.steps / (ms_total * 1e-3)

    # ---- e2e: public API, HOST feature maps, H2D + D2H inside the timed region ---------------
    e2e = None
    if args.e2e_steps > 0:
        k2 = min(args.steps, args.e2e_steps)
        host = [torch.empty(d, H, W, dtype=torch.float32).pin_memory() for _ in range(2)]
        for hbuf, src in zip(host, pool):
            hbuf.copy_(src.permute(2, 0, 1))  # the reference's planar layout (backproject.py:110-113)
        for i in range(2):  # untimed: allocates the two staging buffers
            bp.add_view_host(vm[my_view(i)], K, W, H, host[i % 2])
        bp.flush()
        torch.cuda.synchronize(dev)
        bp.reset()
        barrier()
        t0 = time.perf_counter()
        d2h = 0
        for i in range(k2):
            # public host-facing call: pinned host map -> copy stream -> staging buffer, overlapped with the
            # back-projection of the previous view (BackProjector.add_view_host)
            bp.add_view_host(vm[my_view(i)], K, W, H, host[i % 2])
            res = bp._stats.cpu()  # the step's result read: running counters (rows, entries walked)
            d2h = res.numel() * 8
        bp.flush()
        res = bp._stats.cpu()
        torch.cuda.synchronize(dev)
        secs = time.perf_counter() - t0
        tt = torch.tensor([secs], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e = {"value": world * k2 / float(tt.item()), "unit": UNIT, "h2d_bytes_per_step": fmap_bytes + 100,
               "d2h_bytes_per_step": d2h, "steps": k2,
               "note": "feature map uploaded from pinned host memory every view, overlapped with the previous view's "
                       "kernels (PCIe-bound: 2.23 GB/view); the reference keeps it on the GPU, where the encoder "
                       "produces it"}
        del host
        bp._stage = None
        # extra (not the headline): the same loop fed with the ENCODER-resolution map the reference's driver
        # actually has (backproject.py:109: [512,h,w] before F.interpolate); the upsample is fused on the GPU
        try:
            hlow = [torch.nn.functional.normalize(torch.randn(d, enc, enc), dim=0).pin_memory() for _ in range(2)]
            for i in range(2):
                bp.add_view_host(vm[my_view(i)], K, W, H, hlow[i % 2], lowres_mode=lmode)
            bp.flush()
            torch.cuda.synchronize(dev)
            bp.reset()
            barrier()
            t0 = time.perf_counter()
            for i in range(k2):
                bp.add_view_host(vm[my_view(i)], K, W, H, hlow[i % 2], lowres_mode=lmode)
                res = bp._stats.cpu()
            bp.flush()
            res = bp._stats.cpu()
            torch.cuda.synchronize(dev)
            tt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            e2e["lowres_variant"] = {"value": world * k2 / float(tt.item()), "unit": UNIT,
                                     "h2d_bytes_per_step": d * enc * enc * 4 + 100, "d2h_bytes_per_step": d2h,
                                     "note": f"host hands over the {enc}x{enc} encoder map (what the reference has before "
                                             "F.interpolate, backproject.py:108-112); back-projected against that map "
                                             "directly by the adjoint kernel, no [H,W,D] tensor (SURVEY 8f row 3)"}
        except Exception as ex:  # never let the extra measurement break the contract line
            e2e["lowres_variant"] = {"error": str(ex)[:200]}

    # ---- per-stage timing (separate, untimed loop: CUDA events at the stage boundaries inside the library) --------
    stage_ms = None
    if rank == 0 and args.stage_views > 0:
        import ctypes as C

        lib = gwbp._lib.lib()
        nst = len(gwbp._lib.PROFILE_STAGES)
        acc = [0.0] * nst
        buf = (C.c_float * nst)()
        lib.gwbp_profile_enable(1)
        for i in range(args.stage_views):
            step(args.warmup + i)
            gwbp._lib.check(lib.gwbp_profile_read(buf, nst), "gwbp_profile_read")  # waits for this view
            for j in range(nst):
                acc[j] += max(0.0, float(buf[j]))
        lib.gwbp_profile_enable(0)
        stage_ms = [a / args.stage_views for a in acc]

    # ---- zero-edit shim: the reference's own loop body (backproject.py:115-151) on the drop-in `rasterization` ----
    shim = None
    if rank == 0 and args.shim_views > 0:
        try:
            shim = _time_shim(gwbp, torch, dev, sc, vm, K, W, H, d, pool, args.shim_views)
        except Exception as ex:  # the extra measurement never breaks the contract line
            shim = {"error": str(ex)[:200]}

    if rank == 0:
        hbm, tf, src = _peaks()
        k = max(1, args.steps)
        rows, walked = st.get("rows_nonzero", 0) / k, st.get("entries_walked", 0) / k
        # algorithmic bytes of the fused kernel: walked entries x (4 B id + 32 B record)
        # + the feature map once + one (D+1)-float accumulator update per non-zero row
        lr_adjoint = args.features == "lowres" and args.kernel != "simt" and bool(
            gwbp._lib.lib().gwbp_lowres_adjoint_supported(W, H, enc, enc, d, nearest))
        in_bytes = float(enc * enc * d * 4) if args.features == "lowres" else float(fmap_bytes)  # the map as supplied
        algo_bytes = walked * 36.0 + (in_bytes if lr_adjoint else fmap_bytes) + rows * (d + 1) * 4.0
        achieved = algo_bytes / (ms_kernel * 1e-3) / 1e9 if ms_kernel > 0 else 0.0
        # DRAM traffic / tensor-pipe share of the dominant kernel come from an `ncu --set full` capture and are only
        # reported when that capture was taken from THESE sources (profiles/r02_traffic.json records their hash)
        traffic = tensor_pct = None
        traffic_note = "no capture of these sources"
        try:
            with open(os.path.join(ROOT, "profiles", "r02_traffic.json")) as f:
                cap = json.load(f)
            ent = cap.get(f"{args.config}:{args.features}:{'simt' if args.kernel == 'simt' else 'tc'}")
            if ent and cap.get("source_sha") == _source_sha():
                traffic, tensor_pct = ent.get("traffic"), ent.get("tensor_pipe_pct")
                traffic_note = f"ncu --set full capture of source {cap['source_sha']} ({ent.get('file', '')})"
            elif ent:
                traffic_note = f"capture is of source {cap.get('source_sha')}, this is {_source_sha()}: not reported"
        except Exception:
            pass
        n_g, n_vis, n_is = cfg["n"], last.n_vis, last.n_isects
        tiles = ((W + 15) // 16) * ((H + 15) // 16)
        # B_view = 44 N + 40 n_vis + 68 I + HWD s_F + 4 R (D+1)   (I = the list this engine builds)
        view_bytes = 44.0 * n_g + 40.0 * n_vis + 68.0 * n_is + in_bytes + 4.0 * rows * (d + 1)
        ms_view = ms_views / args.steps
        view = {"algorithmic_bytes": view_bytes, "ms": ms_view, "achieved": view_bytes / (ms_view * 1e-3) / 1e9,
                "formula": "SURVEY 8d: 44N + 40n_vis + 68I + 4HWD + 4R(D+1), measured n_vis, I, R"}
        view["frac"] = view["achieved"] / hbm
        stages = None
        if stage_ms is not None:
            adjoint = args.features == "lowres" and bool(gwbp._lib.lib().gwbp_lowres_adjoint_supported(W, H, enc, enc, d, nearest))
            low_bytes = (enc * enc * d * 4.0) if args.features == "lowres" else float(fmap_bytes)
            sb = {"project": 44.0 * n_g + 40.0 * n_vis,                  # SURVEY 8d "Project" (projection + tile test + ordered compaction: one kernel)
                  "count_scan_and_readback": 0.0,                        # host read-back of the view's totals (the one sync per view)
                  "compact": 0.0,                                        # fused into project_pack_kernel: nothing runs here any more
                  "depth_sort": 16.0 * n_vis,                            # one logical pass over (key, value) pairs
                  "tile_binning": 24.0 * n_is + 4.0 * tiles,             # SURVEY 8d "Bin+sort": 12 I written + 12 I read
                  # map read + packed operand written (adjoint low-res path: fp32 map read + 2 x bf16 copy written)
                  "feature_relayout": 2.0 * low_bytes if adjoint else low_bytes + float(fpack_dev_bytes),
                  "backproject": algo_bytes}
            stages = []
            for name, ms in zip(gwbp._lib.PROFILE_STAGES, stage_ms):
                ach = sb[name] / (ms * 1e-3) / 1e9 if ms > 0 and sb[name] > 0 else None
                stages.append({"stage": name, "ms": ms, "algorithmic_bytes": sb[name], "achieved_gbs": ach,
                               "frac": ach / hbm if ach is not None else None})
            stages.append({"stage": "sum", "ms": sum(stage_ms), "note": f"{args.stage_views} views, CUDA events at the stage "
                           "boundaries inside the library (one sync per view to read them)"})
        roofline = {"bound": "hbm", "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm,
                    "traffic": traffic, "tensor_pipe_pct": tensor_pct, "traffic_source": traffic_note, "peak_source": src,
                    "kernel": "bp_simt_kernel" if args.kernel == "simt" else
                              "bp_lr_kernel (fused composite + weight down-sample GEMM + low-res contraction + accumulate)" if lr_adjoint
                              else "bp_tc_kernel (fused composite + tcgen05 contraction + accumulate)",
                    "kernel_ms": ms_kernel, "algorithmic_bytes_per_launch": algo_bytes,
                    "rows_nonzero_per_view": rows, "entries_walked_per_view": walked,
                    "n_vis": last.n_vis, "n_isects": last.n_isects, "stages": stages, "view": view}
        cpu = None
        if world == 1 and args.cpu_budget > 0:
            done, secs, threads = _cpu_views(cfg, 1e9, args.cpu_budget)
            cpu = {"value": done / secs, "unit": UNIT, "cores": threads, "kind": "port",
                   "sample": f"{done} view(s) of config {args.config} at full size with the C oracle "
                             f"(oracle/oracle.c, OpenMP), ~{secs:.0f} s"}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": f"config {args.config}" + (" --features lowres (encoder-resolution maps, adjoint kernel)"
                                                                  if args.features == "lowres" else ""),
                           **cfg, "kernel": args.kernel, "features": args.features,
                           "l2": f"{pool_n} feature maps of {fmap_bytes / 1e9:.2f} GB cycled: every view's input "
                                 "is far larger than the 126 MB L2",
                           "parallelism": f"views sharded over {world} GPU(s), one closing exchange of (num, den): " + {
                               "peer": "sparse reduce-scatter over NVLink peer memory fused with the finalise (one kernel per rank)",
                               "reduce_scatter": "NCCL reduce-scatter", "allreduce": "NCCL all-reduce",
                               "auto": "none (1 GPU)"}[collective]},
                "clocks": clocks, "e2e": e2e, "gpu_launches": launches,
                "ms_views": ms_views, "exchange_ms": ms_total - ms_views, "exchange": collective if world > 1 else None, "exchange_note": peer_note,
                "roofline": roofline, "cpu_baseline": cpu, "shim": shim}
        real_stdout.write(json.dumps(line) + "\n")
        real_stdout.flush()
    if world > 1:
        dist.destroy_process_group()


def run_query(args):
    """BASELINE configs[3] (segment.py query path): per view, forward 512-d feature render + text cosine mask
    (segment.py:209-224) at garden scale.  A "step" is ONE frame: project + bin + sort + render + mask.  `value` = the
    reference's order of operations (D-channel render -> normalise -> scores -> compare); `linear_path` = the same mask
    from a P-channel render of the per-Gaussian scores.  One JSON line, same contract."""
    import torch

    import gwbp

    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    S = gwbp.scene
    cfg = dict(S.CONFIGS["G"])
    if args.d:
        cfg["d"] = args.d
    W, H, d, V = cfg["width"], cfg["height"], cfg["d"], cfg["views"]
    sc = S.make_scene(cfg["n"], 0)
    vm, K = S.make_cameras(V, W, H, 0)
    t = lambda a: torch.from_numpy(a).to(dev)  # noqa: E731
    scene = gwbp.PackedScene(t(sc.means), t(sc.quats), t(sc.scales), t(sc.opacities))
    g = torch.Generator(device=dev).manual_seed(0)
    feats = torch.nn.functional.normalize(torch.randn(sc.n, d, device=dev, generator=g), dim=1)  # a finished field
    text = t(S.make_text_queries(3, d, 0))
    scores = gwbp.gaussian_scores(feats, text)
    steps = min(args.steps, V)

    def timed(fn, n, warm):
        for i in range(warm):
            fn(i)
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = None
        for i in range(n):
            r = fn(warm + i)
        e1.record()
        torch.cuda.synchronize(dev)
        return e0.elapsed_time(e1) / n, r

    sampler = ClockSampler(dev.index or 0)
    sampler.start()
    l0 = int(gwbp._lib.lib().gwbp_launch_count())
    ms_exact, m_exact = timed(lambda i: gwbp.render_mask_2d(scene, feats, text, 1, vm[i % V], K, W, H, exact_render=True),
                              steps, args.warmup)
    launches = int(gwbp._lib.lib().gwbp_launch_count()) - l0
    clocks = sampler.stop()
    ms_lin, m_lin = timed(lambda i: gwbp.render_mask_2d(scene, feats, text, 1, vm[i % V], K, W, H, exact_render=False,
                                                        scores=scores), steps, args.warmup)
    last = (args.warmup + steps - 1) % V
    diff_paths = int((m_exact != m_lin).sum())
    ms_3d, _ = timed(lambda i: gwbp.get_mask3d(feats, text, 1), 5, 2)
    view = gwbp.View(scene, gwbp.make_camera(vm[last], K, W, H), tile_cull=True)
    ms_tc, _ = timed(lambda i: view.render(feats, None, gwbp.KERNEL_TC), 5, 2)
    # end to end: the frame's mask is read back to the host every step (segment.py:226 `.cpu()`), text prompts uploaded
    text_host = text.cpu().pin_memory()
    t0 = time.perf_counter()
    k2 = min(steps, max(args.e2e_steps, 1))
    for i in range(k2):
        tx = text_host.to(dev, non_blocking=True)
        m = gwbp.render_mask_2d(scene, feats, tx, 1, vm[i % V], K, W, H, exact_render=True)
        mh = m.cpu()
    torch.cuda.synchronize(dev)
    e2e_fps = k2 / (time.perf_counter() - t0)
    # parity of the masks against the CPU oracle on ONE frame (bounded CPU work)
    parity = None
    if args.cpu_budget > 0:
        from oracle import c_oracle, gsplat_oracle

        c_oracle.set_num_threads(os.cpu_count() or 1)
        cv = c_oracle.View(sc.means, sc.quats, sc.scales, sc.opacities, vm[last], K, W, H)
        tc0 = time.perf_counter()
        r_o, a_o = cv.render(feats.cpu().numpy())
        cpu_s = time.perf_counter() - tc0
        m_o, s_o = gsplat_oracle.mask2d(r_o, text.cpu().numpy(), 1)
        margin = abs(s_o[..., 0] - s_o[..., 1:].max(-1))
        cover = a_o > 1e-3
        me, ml = m_exact.cpu().numpy(), m_lin.cpu().numpy()
        parity = {"frame": int(last), "pixels": int(W * H), "covered_pixels": int(cover.sum()),
                  "differing_exact_vs_oracle": int(((me != m_o) & cover).sum()),
                  "differing_exact_vs_oracle_margin_gt_1e-4": int(((me != m_o) & cover & (margin > 1e-4)).sum()),
                  "differing_linear_vs_oracle_margin_gt_1e-4": int(((ml != m_o) & cover & (margin > 1e-4)).sum()),
                  "oracle_frame_seconds": cpu_s, "oracle_threads": c_oracle.num_threads()}
        cpu = {"value": 1.0 / cpu_s, "unit": "frames/s", "cores": c_oracle.num_threads(), "kind": "port",
               "sample": f"the render of ONE frame at full size with the C oracle ({cpu_s:.1f} s; mask compare excluded)"}
    else:
        cpu = None
    hbm, _, src = _peaks()
    out_bytes = W * H * d * 4.0
    line = {"metric": "2-D query masks/s (forward 512-d feature render + text cosine mask per view)", "value": 1e3 / ms_exact,
            "unit": "frames/s", "n_gpus": 1, "steps": steps, "warmup": args.warmup, "ms_per_step": ms_exact,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "config Q (BASELINE configs[3]): segment.py query path at garden scale", **cfg,
                       "prompts": 3, "n_pos": 1, "l2": "every frame writes and re-reads a 2.2 GB render"},
            "clocks": clocks, "gpu_launches": launches,
            "e2e": {"value": e2e_fps, "unit": "frames/s", "h2d_bytes_per_step": int(text_host.numel() * 4 + 100),
                    "d2h_bytes_per_step": int(W * H), "steps": k2},
            "linear_path": {"value": 1e3 / ms_lin, "unit": "frames/s", "ms_per_step": ms_lin,
                            "pixels_differing_from_exact_path": diff_paths,
                            "what": "P-channel render of per-Gaussian scores f_g . t_j (computed once per query), same compare"},
            "mask3d": {"ms": ms_3d, "gbs": sc.n * d * 4 / ms_3d / 1e6, "frac": sc.n * d * 4 / ms_3d / 1e6 / hbm},
            "roofline": {"bound": "hbm", "kernel": "render_tc_kernel (tcgen05 forward feature render)", "kernel_ms": ms_tc,
                         "algorithmic_bytes_per_launch": out_bytes, "achieved": out_bytes / ms_tc / 1e6, "peak": hbm,
                         "unit": "GB/s", "frac": out_bytes / ms_tc / 1e6 / hbm, "traffic": None, "peak_source": src,
                         "note": "floor = the 4 HWD bytes of the render that must reach DRAM"},
            "parity": parity, "cpu_baseline": cpu}
    real_stdout.write(json.dumps(line) + "\n")
    real_stdout.flush()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=185)  # one garden-sized job per GPU
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="G")
    ap.add_argument("--kernel", default="auto", choices=["auto", "simt", "tc"])
    ap.add_argument("--pool", type=int, default=8, help="distinct resident feature maps cycled through")
    ap.add_argument("--features", default="full", choices=["full", "lowres"],
                    help="full: [H,W,D] map resident in HBM (the BASELINE metric); lowres: encoder-resolution map, "
                         "back-projected directly by the adjoint kernel bp_lr_kernel (not the headline)")
    ap.add_argument("--collective", default="auto", choices=["auto", "peer", "allreduce", "reduce_scatter"],
                    help="closing exchange of (num, den) for N > 1: peer = sparse reduce-scatter over NVLink peer memory "
                         "fused with the finalise, one hand-written kernel per rank (auto: peer when CUDA IPC mappings "
                         "can be set up, else reduce_scatter); reduce_scatter = NCCL, every rank ends with the global "
                         "sums of ITS rows; allreduce = NCCL, every rank ends with the full field")
    ap.add_argument("--stage-views", type=int, default=6, help="views of the per-stage profiling loop (0 = skip)")
    ap.add_argument("--shim-views", type=int, default=2,
                    help="views of the zero-edit shim loop (3 x rasterization + 2 x backward per view; 0 = skip)")
    ap.add_argument("--d", type=int, default=0, help="override the config's feature width (experiments only)")
    ap.add_argument("--e2e-steps", type=int, default=12)
    ap.add_argument("--cpu-budget", type=float, default=20.0, help="seconds of CPU-oracle work (0 = skip)")
    ap.add_argument("--overlap-pack", type=int, default=-1,
                    help="1/0: force the feature re-layout onto / off a second stream next to projection + binning "
                         "(-1 = the BackProjector default)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    import gwbp

    if not os.path.exists(gwbp._lib.LIB_PATH) and int(os.environ.get("LOCAL_RANK", "0")) == 0:
        # harness convenience only (the library itself never builds or falls back): a checkout without built artefacts
        import __graft_entry__

        __graft_entry__.build()
    for _ in range(600):  # the other ranks of a torchrun launch wait for rank 0's build
        if os.path.exists(gwbp._lib.LIB_PATH):
            break
        time.sleep(0.5)
    if args.config == "Q":
        if int(os.environ.get("RANK", "0")) == 0:
            run_query(args)
        return
    cfg = dict(gwbp.scene.CONFIGS[args.config])
    if args.d:
        cfg["d"] = args.d
    if args.impl == "reference":
        run_reference(args, cfg)
    else:
        run_ours(args, cfg)


if __name__ == "__main__":
    main()
