#!/usr/bin/env python3
"""BASELINE.json config 5 shape on ONE B200: 2160x3840 pair, Nt = 32 (265 M cells, 25 GB of solver state --
the whole volume fits the 180 GB of one GPU, so no time-slab decomposition is needed for capacity).
A few outer ALG2 iterations with each Poisson back-end, per-kernel times and HBM fractions."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "optical-flow-optimal-transport_b200"))
import numpy as np, torch, foto_b200
from foto_b200 import synth
h, w = 2160, 3840
Nt = int(sys.argv[1]) if len(sys.argv) > 1 else 32
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
f0, f1 = synth.make_pair(h, w, seed=0)
d0 = torch.from_numpy(f0).cuda(); d1 = torch.from_numpy(f1).cuda()
o = [torch.empty(h * w, dtype=torch.float64, device="cuda") for _ in range(3)]
ctx = foto_b200.Context(0); ctx.set_profiling(True)
N = Nt * h * w
for name, backend, max_it in (("dct_exact", foto_b200.POISSON_DCT_EXACT, 5), ("cg_parity", foto_b200.POISSON_CG_PARITY, 1)):
    ctx.reset_stats(); ctx.event_record(0)
    info = ctx.solve_dev(d0.data_ptr(), d1.data_ptr(), Nt, w, h, o[0].data_ptr(), o[1].data_ptr(), o[2].data_ptr(),
                         r=1.0, convergence_tol=0.0, reg_epsilon=1e-3, max_it=max_it, backend=backend)
    ctx.event_record(1); ms = ctx.event_elapsed_ms(); st = ctx.stats()
    u = o[0].cpu().numpy()
    out = {"config": 5, "gpus": 1, "grid": [Nt, h, w], "cells": N, "solver_state_GB": 12 * N * 8 / 1e9, "backend": name,
           "outer": info["n_outer"], "seconds": ms / 1e3, "crit": [float(c) for c in info["crit"]], "finite": bool(np.isfinite(u).all()),
           "K1_rhs_GBs": 56 * st["rhs_cells"] / st["rhs_ms"] / 1e6, "K3_prox_GBs": 80 * st["prox_cells"] / st["prox_ms"] / 1e6,
           "poisson_ms_per_solve": st["cg_ms"] / st["cg_launches"]}
    if name == "cg_parity":
        out.update({"cg_iterations": int(st["cg_iterations"]), "cg_ms_per_iteration": st["cg_ms"] / st["cg_iterations"],
                    "cg_GBs_algorithmic": 88 * st["cg_cells"] / st["cg_ms"] / 1e6, "cg_frac_of_hbm_peak": 88 * st["cg_cells"] / st["cg_ms"] / 1e6 / peak})
    else:
        flop = 4.0 * (w + h) * N * 2 / 2      # 2 N (Nx + Ny) multiply-adds forward, the same back
        out.update({"poisson_TFLOPs": 2 * 2.0 * (w + h) * N / (st["cg_ms"] / st["cg_launches"] / 1e3) / 1e12})
    print(json.dumps(out), flush=True)
