#!/usr/bin/env python3
"""BASELINE.json config 3: 64 synthetic Middlebury-shape pairs (8 sequences x 8 illumination perturbations,
foto_b200.synth.config3_pairs) sharded by pair over the visible GPUs (foto_solve_batch: one host thread per GPU, one
atomic work queue, no collective), longest shapes first.  Results are checked against the oracle goldens
(tests/golden/config3_oracle.npz) when present.
Usage: run_config3.py [n_gpus] [backend: cg_parity|dct_exact]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "optical-flow-optimal-transport_b200"))
import numpy as np, foto_b200
from foto_b200 import synth

n_gpus = int(sys.argv[1]) if len(sys.argv) > 1 else foto_b200.device_count()
bname = sys.argv[2] if len(sys.argv) > 2 else "cg_parity"
backend = {"cg_parity": foto_b200.POISSON_CG_PARITY, "dct_exact": foto_b200.POISSON_DCT_EXACT}[bname]
kw = dict(synth.CONFIG3_PARAMS)
groups = {}
for name, h, w, f0, f1 in synth.config3_pairs():
    g = groups.setdefault((h, w), ([], [], []))
    g[0].append(f0); g[1].append(f1); g[2].append(name)
devices = list(range(n_gpus))
h0, w0 = 388, 584
foto_b200.solve_batch(np.stack(groups[(h0, w0)][0][:n_gpus]), np.stack(groups[(h0, w0)][1][:n_gpus]), 4, w0, h0,
                      devices=devices, backend=backend, **dict(kw, max_it=2))     # warm-up: contexts, kernels
gpath = os.path.join(ROOT, "tests", "golden", "config3_oracle.npz")
gold = np.load(gpath) if (os.path.exists(gpath) and bname == "cg_parity") else None
t0 = time.perf_counter()
total_outer = 0; n_pairs = 0; checks = {}; worst = 0.0; compared = 0; outer_mismatch = 0
for (h, w), (a, b, names) in sorted(groups.items(), key=lambda kv: -kv[0][0] * kv[0][1]):
    t1 = time.perf_counter()
    us, vs, ms, outer = foto_b200.solve_batch(np.stack(a), np.stack(b), 4, w, h, devices=devices, backend=backend, **kw)
    dt = time.perf_counter() - t1
    total_outer += int(outer.sum()); n_pairs += len(names)
    checks[f"{h}x{w}"] = {"pairs": len(names), "seconds": dt, "outer_min_max": [int(outer.min()), int(outer.max())],
                          "finite": bool(np.isfinite(us).all() and np.isfinite(ms).all())}
    if gold is not None:
        sub = np.arange(0, h * w, int(gold["sub_stride"]))
        for i, name in enumerate(names):
            if f"{name}/u" not in gold.files:
                continue
            compared += 1
            outer_mismatch += int(outer[i] != int(gold[f"{name}/n_outer"]))
            for comp, arr in (("u", us), ("v", vs), ("m", ms)):
                ref = gold[f"{name}/{comp}"]
                worst = max(worst, float(np.abs(arr[i][sub] - ref).max() / max(np.abs(ref).max(), 1e-300)))
dt = time.perf_counter() - t0
print(json.dumps({"config": 3, "n_gpus": n_gpus, "backend": bname, "pairs": n_pairs,
                  "seconds": dt, "pairs_per_s": n_pairs / dt, "outer_iters_per_s": total_outer / dt, "total_outer": total_outer,
                  "shapes": checks,
                  "vs_oracle": None if gold is None else {"pairs_compared": compared, "outer_count_mismatches": outer_mismatch,
                                                           "worst_rel_err_uvm": worst},
                  "timing": "wall clock incl. H2D/D2H through foto_solve_batch (host threads, one per GPU, shared work queue)"}))
