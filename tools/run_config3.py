#!/usr/bin/env python3
"""BASELINE.json config 3: 64 synthetic Middlebury-shape pairs (8 sequences x 8 brightness
perturbations) sharded by pair over the visible GPUs (work queue, no collective).
Usage: run_config3.py [n_gpus] [backend: cg_parity|dct_exact]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "optical-flow-optimal-transport_b200"))
import numpy as np, foto_b200
from foto_b200 import synth

n_gpus = int(sys.argv[1]) if len(sys.argv) > 1 else foto_b200.device_count()
backend = {"cg_parity": foto_b200.POISSON_CG_PARITY, "dct_exact": foto_b200.POISSON_DCT_EXACT}[sys.argv[2] if len(sys.argv) > 2 else "cg_parity"]
kw = dict(r=1.0, convergence_tol=0.1, reg_epsilon=1e-3, max_it=100)
groups = {}
for s, (name, (h, w)) in enumerate(synth.MIDDLEBURY_SHAPES.items()):
    shift = (0.4 + 0.05 * s, 0.7 - 0.05 * s)
    f0, f1 = synth.make_pair(h, w, seed=s, shift=shift)
    for p in range(8):
        g = groups.setdefault((h, w), ([], [], []))
        g[0].append(f0); g[1].append(f1 if p == 0 else synth.perturb_brightness(f1, h, w, seed=12345 + p)); g[2].append(f"{name}/{p}")
devices = list(range(n_gpus))
foto_b200.solve_batch(np.stack(groups[(388, 584)][0][:n_gpus]), np.stack(groups[(388, 584)][1][:n_gpus]), 4, 584, 388,
                      devices=devices, backend=backend, **kw)                      # warm-up: contexts, kernels
t0 = time.perf_counter()
total_outer = 0; n_pairs = 0; checks = {}
for (h, w), (a, b, names) in groups.items():
    us, vs, ms, outer = foto_b200.solve_batch(np.stack(a), np.stack(b), 4, w, h, devices=devices, backend=backend, **kw)
    total_outer += int(outer.sum()); n_pairs += len(names)
    checks[f"{h}x{w}"] = {"pairs": len(names), "outer_min_max": [int(outer.min()), int(outer.max())],
                          "finite": bool(np.isfinite(us).all() and np.isfinite(ms).all()), "u_abs_mean": float(np.abs(us).mean())}
dt = time.perf_counter() - t0
print(json.dumps({"config": 3, "n_gpus": n_gpus, "backend": sys.argv[2] if len(sys.argv) > 2 else "cg_parity", "pairs": n_pairs,
                  "seconds": dt, "pairs_per_s": n_pairs / dt, "outer_iters_per_s": total_outer / dt, "total_outer": total_outer,
                  "shapes": checks, "timing": "wall clock incl. H2D/D2H through foto_solve_batch (host threads, one per GPU)"}))
