#!/usr/bin/env python3
"""bench.py -- FOTO frame-pairs/s at 388x584 on N B200s (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference [--gpus N] [--steps K] ...    # CPU reference arm (oracle port, full solves)

One "step" = one pass of the hot path over one batch of N x `--pairs-per-gpu` synthetic 388x584 pairs (Nt=4, r=1,
tol=0.1, eps=1e-3, max_it=100: the reference CLI defaults, main.py:38-42) with benamou_brenier.solve semantics
(cg_parity Poisson back-end).  Pairs are independent, so ranks never communicate on the data path (weak scaling);
every rank holds the whole batch and draws pair indices from one shared queue (foto_b200.shard.WorkQueue, an atomic
counter in the process group's store), because outer and CG iteration counts are data dependent; the only
collectives are the barriers, the max-over-ranks of the device time and the gather of per-rank statistics.

  value   : pairs/s, inputs resident in HBM, timed with CUDA events on the library's stream (max over ranks)
  e2e     : pairs/s through the host-buffer C-ABI call the shim modules make (pinned host inputs, H2D + solve + D2H
            inside the timed region, wall clock)
  roofline: the HBM statement the path supports: the streaming kernels K1 / K2a / K3 on a 1080x1920x16 volume
            (3.2 GB, 25x the L2), algorithmic bytes of SURVEY.md section 8(d) over CUDA-event time, worst kernel
            as the headline.  The dominant kernel of the config-1 solve (cg_fused_kernel, state resident in shared
            memory and registers) is not HBM bound; `onchip` reports it against its own limits.
  cpu_baseline: the C oracle (a port of the reference's algorithm), one full solve on one core
"""
import argparse
import json
import os
import statistics
import sys
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
CG_KERNEL_NAMES = {0: "cg_stream_kernel", 3: "cg_fused_kernel"}   # foto_stats.cg_variant
CG_BYTES_PER_CELL_ITER = 88      # 11 fp64 words per cell per CG iteration
RHS_BYTES_PER_CELL = 56
PROX_BYTES_PER_CELL = 80
FLUSH_EVERY = 4                  # solves between two L2 flushes (512 MB memset)
REFERENCE_BUDGET_S = 300.0       # wall budget of the timed steps of --impl reference (full solves, ~65 s per step)


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
    """SM clocks / throttle reasons of the GPUs in use during the timed region (B200_PROFILING.md's clocks line),
    read through NVML from one thread of rank 0 (a `nvidia-smi -lms` process per rank stalls every rank's launches
    while it holds the driver lock)."""

    def __init__(self, indices, period_s=0.2):
        self.indices, self.period = list(indices), period_s
        self.rows, self._stop, self._thr, self.err = [], threading.Event(), None, None

    def _run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            hs = [nv.nvmlDeviceGetHandleByIndex(i) for i in self.indices]
            mx = [nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM) for h in hs]
            while not self._stop.is_set():
                for h, m in zip(hs, mx):
                    self.rows.append((nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM), m, nv.nvmlDeviceGetPowerUsage(h) / 1e3,
                                      nv.nvmlDeviceGetCurrentClocksEventReasons(h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons")
                                      else nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)))
                self._stop.wait(self.period)
        except Exception as e:          # sampling is best effort
            self.err = f"{type(e).__name__}: {e}"

    def __enter__(self):
        if self.indices:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thr:
            self._thr.join(timeout=5)

    def summary(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.rows:
            out["sm_mhz"] = float(statistics.median(r[0] for r in self.rows))
            out["sm_min_mhz"] = float(min(r[0] for r in self.rows))
            out["sm_max_mhz"] = float(max(r[1] for r in self.rows))
            out["power_w_max"] = max(r[2] for r in self.rows)
            bits = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
            out["reasons"] = [n for n, b in bits.items() if any(r[3] & b for r in self.rows)]
            out["samples"] = len(self.rows)
            out["gpus_sampled"] = len(self.indices)
        if self.err:
            out["error"] = self.err
        return out


def dist_env():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


# ------------------------------------------------------------------------------ CPU oracle legs
def _cpu_full_solve(seed_index):
    """One complete reference solve on the CPU: the ALG2 loop to its stopping rule + flow extraction."""
    import oracle
    from foto_b200 import synth
    f0, f1 = synth.make_pair(H, W, seed=pair_seed(0, seed_index))
    t0 = time.perf_counter()
    _, _, _, info = oracle.solve(f0, f1, NT, W, H, return_info=True, **PARAMS)
    return time.perf_counter() - t0, int(info["n_outer"])


def cpu_baseline_single():
    import oracle
    oracle.build()
    dt, outer = _cpu_full_solve(0)
    return {"value": 1.0 / dt, "unit": "pairs/s", "cores": 1, "kind": "port",
            "sample": f"C oracle (oracle/foto_oracle.c), one full solve of pair 0 ({outer} outer iterations + flow "
                      f"extraction) in {dt:.1f} s on one core; the unmodified Python reference needs 259.7 s/pair on "
                      f"one core (BASELINE.md section 2)"}


def run_reference(args):
    """--impl reference: the oracle port on all host cores, one FULL solve of one pair per process per step
    (nothing extrapolated).  A step costs about a minute, so the number of timed steps is bounded by
    REFERENCE_BUDGET_S and reported in `steps`."""
    rank, _, world = dist_env()
    if rank != 0:
        return
    import multiprocessing as mp
    import oracle
    oracle.build()
    avail = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    cores = min(avail, 64)           # the path is DRAM-bound on the host: more processes only add contention
    times, outers = [], []
    with mp.get_context("fork").Pool(cores) as pool:
        t_begin = time.perf_counter()
        for step in range(max(args.steps, 1)):
            t0 = time.perf_counter()
            res = pool.map(_cpu_full_solve, list(range(cores)))
            dt = time.perf_counter() - t0
            times.append(dt); outers += [o for _, o in res]
            if time.perf_counter() - t_begin + dt > REFERENCE_BUDGET_S:
                break
    steps = len(times)
    total = sum(times)
    value = steps * cores / total
    sample = (f"{steps} timed step(s) (of {args.steps} requested; wall budget {REFERENCE_BUDGET_S:.0f} s, no warm-up: nothing to warm "
              f"on the CPU); per step {cores} processes (of {avail} usable cores) each run one FULL solve of one 388x584 pair "
              f"with the C oracle (port of the reference's algorithm: ALG2 loop to its stopping rule, mean {statistics.mean(outers):.2f} "
              f"outer iterations, + flow extraction); the reference itself is pure Python (259.7 s/pair on one core) and cannot "
              f"travel to the GPU box")
    line = {"impl": "reference", "metric": "FOTO frame-pairs/s at 388x584", "value": value, "unit": "pairs/s",
            "n_gpus": args.gpus, "steps": steps, "steps_requested": args.steps, "warmup": 0,
            "ms_per_step": 1e3 * total / steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "foto_388x584_nt4_cli_defaults", "poisson_backend": "cg_parity", "pairs_per_step": cores,
                       "params": PARAMS},
            "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), file=_OUT, flush=True)


# ------------------------------------------------------------------------------ CUDA arm
def run_b200(args):
    import torch
    import foto_b200
    from foto_b200 import shard, synth
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
    NB = world * B                       # pairs per step, whole job
    ctx = foto_b200.Context(local_rank)
    if args.cg_variant is not None:
        ctx.set_cg_variant(args.cg_variant)

    # the whole batch is resident on every GPU (pinned host copy + HBM copy) before any timed region
    # (every rank synthesises its own B pairs; the others arrive by one all-gather, untimed)
    h0 = torch.empty((NB, P), dtype=torch.float64).pin_memory()
    h1 = torch.empty((NB, P), dtype=torch.float64).pin_memory()
    for i in range(B):
        a, b = synth.make_pair(H, W, seed=pair_seed(rank, i))
        h0[rank * B + i] = torch.from_numpy(a); h1[rank * B + i] = torch.from_numpy(b)
    d0, d1 = h0.to(dev), h1.to(dev)
    if world > 1:
        for d, h in ((d0, h0), (d1, h1)):
            dist.all_gather_into_tensor(d, d[rank * B:(rank + 1) * B].clone())
            h.copy_(d)
    du, dv, dm = (torch.empty((NB, P), dtype=torch.float64, device=dev) for _ in range(3))
    hu, hv, hm = (torch.empty((NB, P), dtype=torch.float64).pin_memory() for _ in range(3))
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)       # > 126 MB L2
    torch.cuda.synchronize()

    def barrier():
        shard.barrier()
        torch.cuda.synchronize()

    def l2_flush():
        flush.zero_()

    qn = [0]

    def run_steps(n_steps, solve_one):
        """n_steps passes over the batch, pairs drawn from the shared queue; L2 flushed after every FLUSH_EVERY solves of this rank."""
        qn[0] += 1
        q = shard.WorkQueue(f"bench{qn[0]}", n_steps * NB)
        done = 0
        while True:
            i = q.next()
            if i is None:
                break
            solve_one(i % NB)
            done += 1
            if done % FLUSH_EVERY == 0:
                l2_flush()
        return done

    log = {"pairs": 0, "outer": 0, "cg": 0}

    def solve_dev(i):
        info = ctx.solve_dev(d0[i].data_ptr(), d1[i].data_ptr(), NT, W, H, du[i].data_ptr(), dv[i].data_ptr(), dm[i].data_ptr(), **PARAMS)
        log["pairs"] += 1; log["outer"] += int(info["n_outer"]); log["cg"] += int(info["cg_iters"].sum())

    n0, n1, nu, nv, nm = (t.numpy() for t in (h0, h1, hu, hv, hm))

    def solve_host(i):
        ctx.solve_host(n0[i], n1[i], NT, W, H, nu[i], nv[i], nm[i], **PARAMS)

    # ---- device-resident throughput ("value")
    run_steps(args.warmup, solve_dev)
    torch.cuda.synchronize()
    log.update(pairs=0, outer=0, cg=0)
    ctx.set_profiling(True); ctx.reset_stats()
    barrier()
    with ClockSampler(range(world) if rank == 0 else []) as clk:
        ctx.event_record(0)
        run_steps(args.steps, solve_dev)
        ctx.event_record(1)
        dev_ms_rank = ctx.event_elapsed_ms()
    barrier()
    stats = ctx.stats()
    clocks = clk.summary()
    ctx.set_profiling(False)

    # ---- end to end through the host-buffer C ABI ("e2e")
    run_steps(min(args.warmup, 1), solve_host)
    barrier()
    t0 = time.perf_counter()
    e2e_pairs = run_steps(args.steps, solve_host)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    barrier()

    # ---- GN (config 2) on this rank's pairs, device resident -- auxiliary figure
    mine = list(range(rank * B, rank * B + B))
    gn_iters = []
    for _ in range(2):
        ctx.gn_solve_dev(d0[mine[0]].data_ptr(), d1[mine[0]].data_ptr(), W, H, 0.1, 0.2, du[0].data_ptr(), dv[0].data_ptr(), dm[0].data_ptr())
    barrier()
    ctx.event_record(0)
    for i in mine:
        gn_iters.append(ctx.gn_solve_dev(d0[i].data_ptr(), d1[i].data_ptr(), W, H, 0.1, 0.2, du[i].data_ptr(),
                                         dv[i].data_ptr(), dm[i].data_ptr())["iters"])
    ctx.event_record(1)
    gn_ms = ctx.event_elapsed_ms()
    barrier()

    # ---- opt-in exact Poisson back-end (dct_exact): same pairs, device resident -- auxiliary figure
    for i in mine:
        ctx.solve_dev(d0[i].data_ptr(), d1[i].data_ptr(), NT, W, H, du[i].data_ptr(), dv[i].data_ptr(), dm[i].data_ptr(),
                      backend=foto_b200.POISSON_DCT_EXACT, **PARAMS)
    barrier()
    ctx.event_record(0)
    for i in mine:
        ctx.solve_dev(d0[i].data_ptr(), d1[i].data_ptr(), NT, W, H, du[i].data_ptr(), dv[i].data_ptr(), dm[i].data_ptr(),
                      backend=foto_b200.POISSON_DCT_EXACT, **PARAMS)
    ctx.event_record(1)
    dct_ms = ctx.event_elapsed_ms()
    barrier()

    # device times: max over ranks (no-op without a process group); per-rank record for the scaling analysis
    dev_ms, e2e_ms, gn_ms, dct_ms = shard.max_over_ranks([dev_ms_rank, e2e_s * 1e3, gn_ms, dct_ms], device=dev)
    per_rank = shard.gather_objects({"rank": rank, "ms": dev_ms_rank, "pairs": log["pairs"], "outer_iterations": log["outer"],
                                     "cg_iterations": log["cg"], "cg_ms": stats["cg_ms"], "e2e_ms": e2e_s * 1e3, "e2e_pairs": e2e_pairs,
                                     "launches": int(stats["launches"])})

    if rank == 0:
        peak, peak_src = peaks()
        total_pairs = NB * args.steps
        assert sum(r["pairs"] for r in per_rank) == total_pairs, per_rank
        value = total_pairs / (dev_ms / 1e3)
        tot_cg = sum(r["cg_iterations"] for r in per_rank); tot_outer = sum(r["outer_iterations"] for r in per_rank)
        cg_s = stats["cg_ms"] / 1e3
        us_iter = 1e3 * stats["cg_ms"] / max(stats["cg_iterations"], 1)
        kname = CG_KERNEL_NAMES.get(stats["cg_variant"], "cg_stream_kernel")
        traffic = {}
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            traffic = json.load(open(tpath))
        sm_mhz = clocks.get("sm_mhz") or 1965.0
        onchip = {
            # launches that did work = outer iterations (the look-ahead also enqueues one launch per solve that returns at once)
            "kernel": kname, "share_of_step": stats["cg_ms"] / dev_ms_rank, "avg_launch_ms": stats["cg_ms"] / max(log["outer"], 1),
            "cg_iterations_per_launch": stats["cg_iterations"] / max(log["outer"], 1),
            "launches_total": int(stats["cg_launches"]), "launches_with_work": log["outer"],
            "us_per_cg_iteration": us_iter, "cycles_per_cg_iteration": us_iter * sm_mhz,
            "algorithmic_88B_rate_GBs": CG_BYTES_PER_CELL_ITER * stats["cg_cells"] / cg_s / 1e9 if cg_s > 0 else 0.0,
            "algorithmic_88B_rate_note": "algorithmic bytes of SURVEY.md 8(d) over kernel time; the state (x, r, p, s, w) never leaves "
                                         "shared memory / registers, so this is NOT HBM traffic and is not compared with the HBM peak",
        }
        t = traffic.get(kname)
        if t:
            onchip["dram_bytes_per_launch_ncu"] = t["dram_bytes_per_launch"]
            onchip["ncu_source"] = t["source"]
            if "lsu_shared_wavefronts_per_iteration_per_sm" in t:
                wf = t["lsu_shared_wavefronts_per_iteration_per_sm"]; ar = t["allreduce_floor_cycles"]
                cyc = onchip["cycles_per_cg_iteration"]
                onchip["bounds"] = {
                    "shared_memory_pipe": {"wavefronts_per_iteration_per_sm": wf, "peak_wavefronts_per_clk": 1,
                                           "frac_of_iteration": wf / cyc},
                    "grid_allreduce_floor_cycles": ar, "grid_allreduce_floor_frac_of_iteration": ar / cyc,
                    "note": "critical path of an iteration = stencil + block reduce + one grid all-reduce (two L2 traversals across the "
                            "dies, tools/ubench_allreduce2.cu) + vector update; the x update and the tile-edge export / import run in "
                            "the shadow of the all-reduce (profiles/r2_onchip_phase_cycles.log: 3 113 + 4 272 + 2 343 cycles); the "
                            "shared-memory pipe (ncu wavefronts, 1 per clock per SM) is the busiest unit"}
        line = {
            "metric": "FOTO frame-pairs/s at 388x584", "value": value, "unit": "pairs/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "foto_388x584_nt4_cli_defaults", "pairs_per_gpu_per_step": B,
                       "params": PARAMS, "poisson_backend": "cg_parity",
                       "cg_variant": {0: "streaming (textbook recurrences)",
                                      3: "on-chip single-reduction (Chronopoulos-Gear arrangement)"}.get(stats["cg_variant"], "?"),
                       "sharding": "by pair, one shared work queue over all ranks (no data-path collective)",
                       "l2": f"512 MB device memset after every {FLUSH_EVERY} solves (working set 87 MB/pair < 126 MB L2)",
                       "outer_iterations_per_pair": tot_outer / max(total_pairs, 1)},
            "outer_iters_per_s": tot_outer / (dev_ms / 1e3),
            "cg_iters_per_s": tot_cg / (dev_ms / 1e3),
            "e2e": {"value": total_pairs / (e2e_ms / 1e3), "unit": "pairs/s",
                    "h2d_bytes_per_step": NB * 2 * P * 8, "d2h_bytes_per_step": NB * 3 * P * 8,
                    "timing": "wall clock around foto_solve_host calls, pinned host buffers, max over ranks"},
            "gpu_launches": sum(r["launches"] for r in per_rank),
            "clocks": clocks,
            "per_rank": per_rank,
            "onchip": onchip,
            "aux": {"gn_solves_per_s": world * B / (gn_ms / 1e3), "gn_pcg_iterations": statistics.mean(gn_iters),
                    "gn_config": "GN 388x584 alpha=0.1 lambda=0.2, PCG rtol 1e-13 (config 2)",
                    "foto_dct_exact_pairs_per_s": world * B / (dct_ms / 1e3),
                    "foto_dct_exact_note": "opt-in exact Poisson back-end; differs from the reference's truncated CG by "
                                           "~5e-7 relative (parity-gated against the tight oracle only)"},
        }
        if not args.no_hd:
            hd = hd_roofline(ctx, torch, dev, peak)
            worst = min(("K1_rhs", "K2a_cg_stream", "K3_prox_dual"), key=lambda k: hd[k]["frac"])
            tk = {"K1_rhs": "k_rhs", "K2a_cg_stream": "cg_stream_kernel", "K3_prox_dual": hd["K3_prox_dual"]["kernel"]}[worst]
            tr = traffic.get(tk, {})
            line["roofline"] = {"bound": "hbm", "kernel": tk, "achieved": hd[worst]["achieved"], "peak": peak, "unit": "GB/s",
                                "frac": hd[worst]["achieved"] / peak,
                                "traffic": tr.get("dram_bytes_per_launch"), "traffic_source": tr.get("source"),
                                "peak_source": peak_src,
                                "algorithmic_bytes_per_launch": hd[worst]["algorithmic_bytes_per_launch"],
                                "avg_launch_ms": hd[worst]["ms"],
                                "what": "worst of the three streaming kernels of one ALG2 iteration on a 1080x1920x16 volume "
                                        "(3.2 GB working set, 25x the L2); the config-1 solve itself is SM-resident, see `onchip`",
                                "streaming_hd": hd}
        else:
            line["roofline"] = {"bound": "hbm", "kernel": None, "achieved": None, "peak": peak, "unit": "GB/s", "frac": None,
                                "traffic": None, "note": "--no-hd: HBM-bound streaming kernels not measured in this run"}
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_single()
        print(json.dumps(line), file=_OUT, flush=True)
    barrier()
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


def hd_roofline(ctx, torch, dev, peak):
    """HBM-bound evidence: outer ALG2 iterations on a 1080x1920x16 grid (3.2 GB working set, 25x the
    L2), streaming CG kernel, CUDA events per kernel.  Not part of `value`."""
    from foto_b200 import synth
    h, w, Nt = 1080, 1920, 16
    P_, N_ = h * w, 16 * h * w
    f0, f1 = synth.make_pair(h, w, seed=0)
    a = torch.from_numpy(f0).to(dev); b = torch.from_numpy(f1).to(dev)
    o = [torch.empty(P_, dtype=torch.float64, device=dev) for _ in range(3)]
    kw = dict(r=1.0, convergence_tol=0.0, reg_epsilon=1e-3, max_it=1)
    ctx.set_cg_variant(0)
    ctx.solve_dev(a.data_ptr(), b.data_ptr(), Nt, w, h, o[0].data_ptr(), o[1].data_ptr(), o[2].data_ptr(), **kw)
    ctx.set_profiling(True); ctx.reset_stats()
    ctx.solve_dev(a.data_ptr(), b.data_ptr(), Nt, w, h, o[0].data_ptr(), o[1].data_ptr(), o[2].data_ptr(), **kw)
    st = ctx.stats(); ctx.set_profiling(False)
    ctx.set_cg_variant(-1)
    gbs = lambda bytes_per_cell, cells, ms: bytes_per_cell * cells / max(ms, 1e-9) / 1e6
    k1 = gbs(RHS_BYTES_PER_CELL, st["rhs_cells"], st["rhs_ms"])
    k2 = gbs(CG_BYTES_PER_CELL_ITER, st["cg_cells"], st["cg_ms"])
    k3 = gbs(PROX_BYTES_PER_CELL, st["prox_cells"], st["prox_ms"])
    k3_name = "k_prox_dual" if os.environ.get("FOTO_K3") == "legacy" else "k_prox_dual_tma"
    return {"grid": [Nt, h, w], "cells": N_, "working_set_GB": 12 * N_ * 8 / 1e9, "peak": peak, "unit": "GB/s",
            "K1_rhs": {"kernel": "k_rhs", "achieved": k1, "frac": k1 / peak, "ms": st["rhs_ms"],
                       "algorithmic_bytes_per_launch": RHS_BYTES_PER_CELL * N_},
            "K2a_cg_stream": {"kernel": "cg_stream_kernel", "achieved": k2, "frac": k2 / peak, "ms": st["cg_ms"],
                              "us_per_cg_iteration": 1e3 * st["cg_ms"] / max(st["cg_iterations"], 1),
                              "cg_iterations": st["cg_iterations"], "bytes_per_cell_iteration": CG_BYTES_PER_CELL_ITER,
                              "algorithmic_bytes_per_launch": CG_BYTES_PER_CELL_ITER * N_ * st["cg_iterations"]},
            "K3_prox_dual": {"kernel": k3_name, "achieved": k3, "frac": k3 / peak, "ms": st["prox_ms"],
                             "algorithmic_bytes_per_launch": PROX_BYTES_PER_CELL * N_}}


_OUT = sys.stdout


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--pairs-per-gpu", type=int, default=16,
                    help="pairs per GPU per step; 16 keeps the end-of-queue idle time (at most one solve, 28 ms) near 1 %% of a 5-step run")
    ap.add_argument("--cg-variant", type=int, default=None, help="-1 auto, 0 streaming, 2 on-chip single-reduction")
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
