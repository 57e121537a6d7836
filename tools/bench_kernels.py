#!/usr/bin/env python3
"""Per-kernel algorithmic GB/s on one grid (default: HD 1080x1920, Nt=16 -- far larger than L2).
One outer ALG2 iteration: K1 (rhs), K2a (streaming CG, runs to rtol 1e-6 or 1000 iterations), K3.
Usage: bench_kernels.py [h w Nt] [variant]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "optical-flow-optimal-transport_b200"))
import numpy as np, torch, foto_b200
from foto_b200 import synth

h, w, Nt = (int(x) for x in sys.argv[1:4]) if len(sys.argv) >= 4 else (1080, 1920, 16)
variant = int(sys.argv[4]) if len(sys.argv) > 4 else 0
P = h * w; N = Nt * P
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
f0, f1 = synth.make_pair(h, w, seed=0)
d0 = torch.from_numpy(f0).cuda(); d1 = torch.from_numpy(f1).cuda()
du, dv, dm = (torch.empty(P, dtype=torch.float64, device="cuda") for _ in range(3))
ctx = foto_b200.Context(0); ctx.set_cg_variant(variant)
kw = dict(r=1.0, convergence_tol=0.0, reg_epsilon=1e-3, max_it=2)
ctx.solve_dev(d0.data_ptr(), d1.data_ptr(), Nt, w, h, du.data_ptr(), dv.data_ptr(), dm.data_ptr(), **kw)   # warm-up
ctx.set_profiling(True); ctx.reset_stats()
info = ctx.solve_dev(d0.data_ptr(), d1.data_ptr(), Nt, w, h, du.data_ptr(), dv.data_ptr(), dm.data_ptr(), **kw)
st = ctx.stats()
out = {"grid": [Nt, h, w], "cells": N, "working_set_MB": 12 * N * 8 / 1e6, "cg_iters": info["cg_iters"].tolist(),
       "cg_variant": st["cg_variant"], "peak_GBs": peak,
       "K1_rhs": {"ms": st["rhs_ms"] / 2, "GBs": 56 * st["rhs_cells"] / st["rhs_ms"] / 1e6},
       "K2a_cg": {"us_per_iter": 1e3 * st["cg_ms"] / st["cg_iterations"], "GBs_algorithmic_88B": 88 * st["cg_cells"] / st["cg_ms"] / 1e6,
                  "GBs_moved_80B": 80 * st["cg_cells"] / st["cg_ms"] / 1e6},
       "K3_prox_dual": {"ms": st["prox_ms"] / 2, "GBs": 80 * st["prox_cells"] / st["prox_ms"] / 1e6},
       "K4_flow_ms": st["flow_ms"]}
for k in ("K1_rhs", "K3_prox_dual"):
    out[k]["frac_of_peak"] = out[k]["GBs"] / peak
out["K2a_cg"]["frac_of_peak_algorithmic"] = out["K2a_cg"]["GBs_algorithmic_88B"] / peak
print(json.dumps(out))
