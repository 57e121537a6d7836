#!/usr/bin/env python3
"""Time-slab FOTO solve of ONE volume over the ranks of a torchrun job (config 5 shape, scaled by args).
    python -m torch.distributed.run --nproc-per-node G --master-addr 127.0.0.1 tools/run_slab.py H W NT MAX_IT [--check]
--check: rank 0 also solves the volume alone and (dct_exact) requires bit-identical u, v, m.
--cg: the reference's truncated CG as the Poisson back-end (cg_parity) instead of the exact DCT solve; --check then
      reports the CG iteration counts of both runs and the largest relative difference.
--one-gpu: every rank uses GPU 0 and the process group is gloo (exchanges staged through the host): the 2-rank
           decomposition checked on a single-GPU box."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "optical-flow-optimal-transport_b200"))
import numpy as np, torch, torch.distributed as dist
import foto_b200
from foto_b200 import synth, slab

h, w, Nt, max_it = (int(x) for x in sys.argv[1:5])
check = "--check" in sys.argv
poisson = "cg_parity" if "--cg" in sys.argv else "dct_exact"
rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
one_gpu = "--one-gpu" in sys.argv
if one_gpu:
    local = 0
torch.cuda.set_device(local)
if world > 1 and one_gpu:
    dist.init_process_group("gloo")
elif world > 1:
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
f0, f1 = synth.make_pair(h, w, seed=0)
d0, d1 = torch.from_numpy(f0).to(dev), torch.from_numpy(f1).to(dev)
kw = dict(r=1.0, convergence_tol=0.0 if not check else 0.1, reg_epsilon=1e-3, max_it=max_it)
s = slab.SlabSolver(Nt, w, h)
s.solve(d0, d1, poisson=poisson, **dict(kw, max_it=1))    # warm-up (NCCL channels, DCT tables)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
u, v, m, info = s.solve(d0, d1, poisson=poisson, **kw)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
out = {"poisson": poisson, "mode": "time-slab" + (" (one GPU, gloo, host-staged exchanges)" if one_gpu else ""), "ranks": world, "grid": [Nt, h, w], "cells": Nt * h * w, "planes_per_rank": [b - a for a, b in s.geom["t"]],
       "outer": info["n_outer"], "seconds": dt, "outer_iters_per_s": info["n_outer"] / dt, "crit_last": float(info["crit"][-1])}
if rank == 0 and check:
    ctx = foto_b200.Context(local)
    ou, ov, om = (torch.empty(h * w, dtype=torch.float64, device=dev) for _ in range(3))
    if poisson == "cg_parity":
        ctx.set_cg_variant(0)                             # the streaming kernel: the same recurrences
    ref = ctx.solve_dev(d0.data_ptr(), d1.data_ptr(), Nt, w, h, ou.data_ptr(), ov.data_ptr(), om.data_ptr(),
                        backend=foto_b200.POISSON_CG_PARITY if poisson == "cg_parity" else foto_b200.POISSON_DCT_EXACT, **kw)
    out["single_gpu_outer"] = ref["n_outer"]
    if poisson == "cg_parity":
        out["cg_iters"] = info["cg_iters"].tolist(); out["single_gpu_cg_iters"] = ref["cg_iters"].tolist()
        den = [float(t.abs().max()) for t in (ou, ov, om)]
        out["max_rel_diff"] = float(max(float((a - b).abs().max()) / d for a, b, d in zip((u, v, m), (ou, ov, om), den)))
    out["bit_identical"] = bool(torch.equal(u, ou) and torch.equal(v, ov) and torch.equal(m, om))
    out["max_abs_diff"] = float(max((u - ou).abs().max(), (v - ov).abs().max(), (m - om).abs().max()))
if rank == 0:
    print(json.dumps(out), flush=True)
if world > 1:
    dist.destroy_process_group()
if rank == 0 and check and poisson == "dct_exact" and not out["bit_identical"]:
    sys.exit(3)
