#!/usr/bin/env python3
"""bench.py -- FOTO frame-pairs/s at 388x584 on N B200s (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference [--gpus N] [--steps K] ...    # CPU reference arm (oracle port)

One "step" = one pass of the hot path over one batch: every rank solves `--pairs-per-gpu`
synthetic 388x584 pairs (Nt=4, r=1, tol=0.1, eps=1e-3, max_it=100: the reference CLI defaults,
main.py:38-42) with benamou_brenier.solve semantics (cg_parity Poisson back-end).  Pairs are
independent, so ranks never communicate on the data path (weak scaling); the only collectives
are the barriers and the max-over-ranks of the device time.

  value  : pairs/s, inputs resident in HBM, timed with CUDA events on the library's stream
  e2e    : pairs/s through the host-buffer C-ABI call the shim modules make (pinned host
           inputs, H2D + solve + D2H inside the timed region, wall clock)
  roofline: the persistent CG kernel (K2a), algorithmic bytes 88 B/cell/CG-iteration
            over its CUDA-event time, against MEASURED_PEAKS.json
  cpu_baseline: the C oracle (a port of the reference's algorithm) on a bounded sample
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "optical-flow-optimal-transport_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

H, W, NT = 388, 584, 4
PARAMS = dict(r=1.0, convergence_tol=0.1, reg_epsilon=1e-3, max_it=100)
P = H * W
N = NT * P
CG_KERNEL_NAMES = {0: "cg_stream_kernel", 1: "cg_onchip_kernel", 3: "cg_fused_kernel"}   # foto_stats.cg_variant
CG_BYTES_PER_CELL_ITER = 88      # 11 fp64 words per cell per CG iteration
RHS_BYTES_PER_CELL = 56
PROX_BYTES_PER_CELL = 80
CPU_SAMPLE_OUTER = 2             # outer iterations timed on the CPU (of the 9 the config needs)
EXPECTED_OUTER = 9               # measured with the reference on pair 0


def pair_seed(rank, i):
    """Config-1 pairs: a seeded texture and its (0.4, 0.7)-pixel translate; a different seed for
    every pair of every rank (pair 0 of rank 0 is exactly BASELINE.md's pair)."""
    return 1000 * rank + i


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.path = None

    def __enter__(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None
        return self

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()

    def summary(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        try:
            rows = [l.strip().split(", ") for l in open(self.path) if l.strip()]
            sm = [float(r[0]) for r in rows]
            out["sm_mhz"] = statistics.median(sm)
            out["sm_max_mhz"] = float(rows[0][1])
            out["power_w_max"] = max(float(r[2]) for r in rows)
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            out["reasons"] = [n for i, n in enumerate(names) if any(r[3 + i].strip() == "Active" for r in rows)]
            out["samples"] = len(rows)
        except Exception as e:      # sampling is best effort
            out["error"] = str(e)
        finally:
            if self.path and os.path.exists(self.path):
                os.unlink(self.path)
        return out


def dist_env():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


# ------------------------------------------------------------------------------ CPU oracle legs
def _cpu_sample_worker(arg):
    """One bounded CPU sample: the first n_outer outer iterations of one pair."""
    seed, n_outer = arg if isinstance(arg, tuple) else (arg, CPU_SAMPLE_OUTER)
    import oracle
    from foto_b200 import synth
    f0, f1 = synth.make_pair(H, W, seed=pair_seed(0, seed))
    t0 = time.perf_counter()
    kw = dict(PARAMS)
    kw["max_it"] = n_outer
    kw["convergence_tol"] = 0.0
    _, _, _, info = oracle.solve(f0, f1, NT, W, H, return_info=True, **kw)
    return time.perf_counter() - t0, int(info["n_outer"])


def cpu_baseline_single():
    import oracle
    oracle.build()
    dt, outer = _cpu_sample_worker(0)
    pairs_per_s = (outer / EXPECTED_OUTER) / dt
    return {"value": pairs_per_s, "unit": "pairs/s", "cores": 1, "kind": "port",
            "sample": f"C oracle (oracle/foto_oracle.c), first {outer} of {EXPECTED_OUTER} outer iterations of pair 0 "
                      f"in {dt:.1f} s, scaled by {EXPECTED_OUTER}/{outer}; the unmodified Python reference needs "
                      f"259.7 s/pair on one core (BASELINE.md section 2)"}


def run_reference(args):
    """--impl reference: the oracle port on all host cores, one pair per process per step."""
    rank, _, world = dist_env()
    if rank != 0:
        return
    import multiprocessing as mp
    import oracle
    oracle.build()
    avail = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    cores = min(avail, 64)           # the path is DRAM-bound on the host: more processes only add contention
    times = []
    with mp.get_context("fork").Pool(cores) as pool:
        for step in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            res = pool.map(_cpu_sample_worker, [(i, 1) for i in range(cores)])   # 1 of ~9 outer iterations each
            dt = time.perf_counter() - t0
            if step >= args.warmup:
                times.append(dt)
            outer = res[0][1]
    total = sum(times)
    value = args.steps * cores * (outer / EXPECTED_OUTER) / total
    sample = (f"per step, {cores} processes (of {avail} usable cores) each run the first {outer} of {EXPECTED_OUTER} outer iterations of one "
              f"388x584 pair with the C oracle (port of the reference's algorithm; the reference itself is pure "
              f"Python and cannot travel to the GPU box); pairs/s scaled by {EXPECTED_OUTER}/{outer}")
    line = {"impl": "reference", "metric": "FOTO frame-pairs/s at 388x584", "value": value, "unit": "pairs/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "foto_388x584_nt4_cli_defaults", "poisson_backend": "cg_parity"},
            "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), file=_OUT, flush=True)


# ------------------------------------------------------------------------------ CUDA arm
def run_b200(args):
    import torch
    import foto_b200
    from foto_b200 import synth
    rank, local_rank, world = dist_env()
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")     # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    B = args.pairs_per_gpu
    ctx = foto_b200.Context(local_rank)
    if args.cg_variant is not None:
        ctx.set_cg_variant(args.cg_variant)

    # synthetic pairs, distinct per rank; resident in HBM before the timed region
    pairs = [synth.make_pair(H, W, seed=pair_seed(rank, i)) for i in range(B)]
    h0 = torch.empty((B, P), dtype=torch.float64).pin_memory()
    h1 = torch.empty((B, P), dtype=torch.float64).pin_memory()
    for i, (a, b) in enumerate(pairs):
        h0[i] = torch.from_numpy(a); h1[i] = torch.from_numpy(b)
    d0, d1 = h0.to(dev), h1.to(dev)
    du, dv, dm = (torch.empty((B, P), dtype=torch.float64, device=dev) for _ in range(3))
    hu, hv, hm = (torch.empty((B, P), dtype=torch.float64).pin_memory() for _ in range(3))
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)       # > 126 MB L2
    torch.cuda.synchronize()

    from foto_b200 import shard

    def barrier():
        shard.barrier()
        torch.cuda.synchronize()

    outer_counts = []

    def step_device():
        for i in range(B):
            info = ctx.solve_dev(d0[i].data_ptr(), d1[i].data_ptr(), NT, W, H, du[i].data_ptr(), dv[i].data_ptr(),
                                 dm[i].data_ptr(), **PARAMS)
            outer_counts.append(info["n_outer"])

    def step_host():
        n0, n1, nu, nv, nm = (t.numpy() for t in (h0, h1, hu, hv, hm))
        for i in range(B):
            ctx.solve_host(n0[i], n1[i], NT, W, H, nu[i], nv[i], nm[i], **PARAMS)

    def l2_flush():
        flush.zero_()
        torch.cuda.synchronize()

    # ---- device-resident throughput ("value")
    for _ in range(args.warmup):
        step_device(); l2_flush()
    ctx.set_profiling(True); ctx.reset_stats()
    barrier()
    with ClockSampler(local_rank) as clk:
        ctx.event_record(0)
        for _ in range(args.steps):
            step_device(); l2_flush()
        ctx.event_record(1)
        dev_ms = ctx.event_elapsed_ms()
    barrier()
    stats = ctx.stats()
    clocks = clk.summary()
    ctx.set_profiling(False)

    # ---- end to end through the host-buffer C ABI ("e2e")
    for _ in range(min(args.warmup, 2)):
        step_host()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_host(); l2_flush()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    barrier()

    # ---- GN (config 2) on the same pairs, device resident -- auxiliary figure
    gn_iters = []
    for _ in range(2):
        ctx.gn_solve_dev(d0[0].data_ptr(), d1[0].data_ptr(), W, H, 0.1, 0.2, du[0].data_ptr(), dv[0].data_ptr(), dm[0].data_ptr())
    barrier()
    ctx.event_record(0)
    for i in range(B):
        gn_iters.append(ctx.gn_solve_dev(d0[i].data_ptr(), d1[i].data_ptr(), W, H, 0.1, 0.2, du[i].data_ptr(),
                                         dv[i].data_ptr(), dm[i].data_ptr())["iters"])
    ctx.event_record(1)
    gn_ms = ctx.event_elapsed_ms()
    barrier()

    # ---- opt-in exact Poisson back-end (dct_exact): same pairs, device resident -- auxiliary figure
    for i in range(B):
        ctx.solve_dev(d0[i].data_ptr(), d1[i].data_ptr(), NT, W, H, du[i].data_ptr(), dv[i].data_ptr(), dm[i].data_ptr(),
                      backend=foto_b200.POISSON_DCT_EXACT, **PARAMS)
    barrier()
    ctx.event_record(0)
    for i in range(B):
        ctx.solve_dev(d0[i].data_ptr(), d1[i].data_ptr(), NT, W, H, du[i].data_ptr(), dv[i].data_ptr(), dm[i].data_ptr(),
                      backend=foto_b200.POISSON_DCT_EXACT, **PARAMS)
    ctx.event_record(1)
    dct_ms = ctx.event_elapsed_ms()
    barrier()

    # device times: max over ranks (no-op without a process group)
    dev_ms, e2e_ms, gn_ms, dct_ms = shard.max_over_ranks([dev_ms, e2e_s * 1e3, gn_ms, dct_ms], device=dev)

    if rank == 0:
        peak, peak_src = peaks()
        total_pairs = world * B * args.steps
        value = total_pairs / (dev_ms / 1e3)
        cg_bytes = CG_BYTES_PER_CELL_ITER * stats["cg_cells"]
        cg_s = stats["cg_ms"] / 1e3
        achieved = cg_bytes / cg_s / 1e9 if cg_s > 0 else 0.0
        traffic, traffic_src = None, None            # DRAM bytes per launch from the committed ncu --set full capture
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            t = json.load(open(tpath)).get(CG_KERNEL_NAMES.get(stats["cg_variant"], "cg_stream_kernel"))
            if t:
                traffic, traffic_src = t["dram_bytes_per_launch"], t["source"]
        per_launch_iters = stats["cg_iterations"] / max(stats["cg_launches"], 1)
        line = {
            "metric": "FOTO frame-pairs/s at 388x584", "value": value, "unit": "pairs/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "foto_388x584_nt4_cli_defaults", "pairs_per_gpu_per_step": B,
                       "params": PARAMS, "poisson_backend": "cg_parity",
                       "cg_variant": {0: "streaming", 1: "on-chip (textbook recurrences, 2 all-reduces/iteration)",
                                      3: "on-chip single-reduction (Chronopoulos-Gear arrangement)"}.get(stats["cg_variant"], "?"),
                       "l2": "512 MB device memset between steps (working set 87 MB/pair < 126 MB L2)",
                       "outer_iterations_per_pair": statistics.mean(outer_counts) if outer_counts else None},
            "outer_iters_per_s": world * stats["cg_launches"] / (dev_ms / 1e3),
            "cg_iters_per_s": world * stats["cg_iterations"] / (dev_ms / 1e3),
            "e2e": {"value": total_pairs / (e2e_ms / 1e3), "unit": "pairs/s",
                    "h2d_bytes_per_step": B * 2 * P * 8, "d2h_bytes_per_step": B * 3 * P * 8,
                    "timing": "wall clock around foto_solve_host calls, pinned host buffers"},
            "gpu_launches": int(stats["launches"]),
            "clocks": clocks,
            "roofline": {"kernel": CG_KERNEL_NAMES.get(stats["cg_variant"], "cg_stream_kernel"),
                         "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": CG_BYTES_PER_CELL_ITER * N * per_launch_iters,
                         "avg_launch_ms": stats["cg_ms"] / max(stats["cg_launches"], 1),
                         "cg_iterations_per_launch": per_launch_iters,
                         "us_per_cg_iteration": 1e3 * stats["cg_ms"] / max(stats["cg_iterations"], 1),
                         "share_of_step": stats["cg_ms"] / dev_ms,
                         "note": "working set (4 CG vectors = 29 MB) is L2/SM resident at this size, so the "
                                 "algorithmic-bytes rate is not bounded by HBM; see DESIGN.md",
                         "other_kernels": {
                             "rhs_K1_GBs": RHS_BYTES_PER_CELL * stats["rhs_cells"] / max(stats["rhs_ms"], 1e-9) / 1e6,
                             "prox_dual_K3_GBs": PROX_BYTES_PER_CELL * stats["prox_cells"] / max(stats["prox_ms"], 1e-9) / 1e6}},
            "aux": {"gn_solves_per_s": world * B / (gn_ms / 1e3), "gn_pcg_iterations": statistics.mean(gn_iters),
                    "gn_config": "GN 388x584 alpha=0.1 lambda=0.2, PCG rtol 1e-13 (config 2)",
                    "foto_dct_exact_pairs_per_s": world * B / (dct_ms / 1e3),
                    "foto_dct_exact_note": "opt-in exact Poisson back-end; differs from the reference's truncated CG by "
                                           "~5e-7 relative (parity-gated against the tight oracle only)"},
        }
        if world == 1 and not args.no_hd:
            line["roofline_streaming_hd"] = hd_roofline(ctx, torch, dev, peak)
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_single()
        print(json.dumps(line), file=_OUT, flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


def hd_roofline(ctx, torch, dev, peak):
    """HBM-bound evidence: one outer ALG2 iteration on a 1080x1920x16 grid (3.2 GB working set, 25x the
    L2), streaming CG kernel, CUDA events per kernel.  Not part of `value`."""
    from foto_b200 import synth
    h, w, Nt = 1080, 1920, 16
    P_, N_ = h * w, 16 * h * w
    f0, f1 = synth.make_pair(h, w, seed=0)
    a = torch.from_numpy(f0).to(dev); b = torch.from_numpy(f1).to(dev)
    o = [torch.empty(P_, dtype=torch.float64, device=dev) for _ in range(3)]
    kw = dict(r=1.0, convergence_tol=0.0, reg_epsilon=1e-3, max_it=1)
    ctx.solve_dev(a.data_ptr(), b.data_ptr(), Nt, w, h, o[0].data_ptr(), o[1].data_ptr(), o[2].data_ptr(), **kw)
    ctx.set_profiling(True); ctx.reset_stats()
    ctx.solve_dev(a.data_ptr(), b.data_ptr(), Nt, w, h, o[0].data_ptr(), o[1].data_ptr(), o[2].data_ptr(), **kw)
    st = ctx.stats(); ctx.set_profiling(False)
    gbs = lambda bytes_per_cell, cells, ms: bytes_per_cell * cells / max(ms, 1e-9) / 1e6
    k1 = gbs(RHS_BYTES_PER_CELL, st["rhs_cells"], st["rhs_ms"])
    k2 = gbs(CG_BYTES_PER_CELL_ITER, st["cg_cells"], st["cg_ms"])
    k3 = gbs(PROX_BYTES_PER_CELL, st["prox_cells"], st["prox_ms"])
    return {"grid": [Nt, h, w], "cells": N_, "working_set_GB": 12 * N_ * 8 / 1e9, "peak": peak, "unit": "GB/s",
            "K1_rhs": {"achieved": k1, "frac": k1 / peak, "ms": st["rhs_ms"]},
            "K2a_cg_stream": {"achieved": k2, "frac": k2 / peak, "us_per_cg_iteration": 1e3 * st["cg_ms"] / max(st["cg_iterations"], 1),
                              "cg_iterations": st["cg_iterations"], "bytes_per_cell_iteration": CG_BYTES_PER_CELL_ITER},
            "K3_prox_dual": {"achieved": k3, "frac": k3 / peak, "ms": st["prox_ms"]}}


_OUT = sys.stdout


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--pairs-per-gpu", type=int, default=4)
    ap.add_argument("--cg-variant", type=int, default=None, help="-1 auto, 0 streaming, 1 on-chip textbook, 2 on-chip single-reduction")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-hd", action="store_true", help="skip the 1080x1920x16 streaming-roofline measurement")
    args = ap.parse_args()
    # stdout carries exactly one JSON line: libraries that print there (NCCL's version banner under NCCL_DEBUG=VERSION)
    # are sent to stderr at the file-descriptor level, the JSON goes to the saved descriptor
    global _OUT
    sys.stdout.flush()
    _OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
