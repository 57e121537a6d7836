#!/usr/bin/env python3
"""BASELINE.json config 4: one synthetic 1080x1920 pair, Nt in {8,16,32}, iteration-count sweep with
tol = 0 (run to max_it): criterion vs iterations vs time.  Usage: run_config4.py [backend] [max_its...]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "optical-flow-optimal-transport_b200"))
import numpy as np, torch, foto_b200
from foto_b200 import synth
name = sys.argv[1] if len(sys.argv) > 1 else "dct_exact"
backend = {"cg_parity": foto_b200.POISSON_CG_PARITY, "dct_exact": foto_b200.POISSON_DCT_EXACT}[name]
sweep = [int(x) for x in sys.argv[2:]] or [10, 25, 50, 100, 200]
h, w = 1080, 1920
f0, f1 = synth.make_pair(h, w, seed=0)
d0 = torch.from_numpy(f0).cuda(); d1 = torch.from_numpy(f1).cuda()
du, dv, dm = (torch.empty(h * w, dtype=torch.float64, device="cuda") for _ in range(3))
ctx = foto_b200.Context(0)
for Nt in (8, 16, 32):
    for max_it in sweep:
        ctx.event_record(0)
        info = ctx.solve_dev(d0.data_ptr(), d1.data_ptr(), Nt, w, h, du.data_ptr(), dv.data_ptr(), dm.data_ptr(),
                             r=1.0, convergence_tol=0.0, reg_epsilon=1e-3, max_it=max_it, backend=backend)
        ctx.event_record(1)
        ms = ctx.event_elapsed_ms()
        u = du.cpu().numpy(); v = dv.cpu().numpy()
        print(json.dumps({"config": 4, "backend": name, "Nt": Nt, "max_it": max_it, "outer": info["n_outer"], "seconds": ms / 1e3,
                          "outer_iters_per_s": info["n_outer"] / (ms / 1e3), "crit_last": float(info["crit"][-1]),
                          "u_mean": float(u.mean()), "v_mean": float(v.mean()), "cells": Nt * h * w}), flush=True)
