#!/usr/bin/env python3
"""dct_exact Poisson solve at HD size: ms per solve (CUDA events around the K2b launches) after a warm-up that builds
the tables.  Usage: bench_dct_hd.py [h w Nt [iters]]   (env FOTO_DCT_DENSE=1: dense transforms instead of the folded ones; FOTO_DCT_LEVELS=1|2: folding levels)"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "optical-flow-optimal-transport_b200"))
import torch, foto_b200
from foto_b200 import synth
h, w, Nt = (int(x) for x in sys.argv[1:4]) if len(sys.argv) >= 4 else (1080, 1920, 16)
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 20
f0, f1 = synth.make_pair(h, w, seed=0)
d0 = torch.from_numpy(f0).cuda(); d1 = torch.from_numpy(f1).cuda()
o = [torch.empty(h * w, dtype=torch.float64, device="cuda") for _ in range(3)]
ctx = foto_b200.Context(0)
kw = dict(r=1.0, convergence_tol=0.0, reg_epsilon=1e-3, backend=foto_b200.POISSON_DCT_EXACT)
ctx.solve_dev(d0.data_ptr(), d1.data_ptr(), Nt, w, h, *[t.data_ptr() for t in o], max_it=1, **kw)
ctx.set_profiling(True); ctx.reset_stats()
ctx.event_record(0)
info = ctx.solve_dev(d0.data_ptr(), d1.data_ptr(), Nt, w, h, *[t.data_ptr() for t in o], max_it=iters, **kw)
ctx.event_record(1)
ms = ctx.event_elapsed_ms(); st = ctx.stats()
flop_dense = 4.0 * (w + h) * Nt * h * w          # 2 transforms each way, 2 n flop per element and axis
print(json.dumps({"grid": [Nt, h, w], "variant": {k: os.environ.get(k) for k in ("FOTO_DCT_DENSE", "FOTO_DCT_LEVELS")},
                  "outer": info["n_outer"], "ms_per_outer": ms / info["n_outer"], "poisson_ms_per_solve": st["cg_ms"] / st["cg_launches"],
                  "rhs_ms": st["rhs_ms"] / st["cg_launches"], "prox_ms": st["prox_ms"] / st["cg_launches"],
                  "dense_equivalent_TFLOPs": flop_dense / (st["cg_ms"] / st["cg_launches"] * 1e-3) / 1e12,
                  "checksum": float(o[0].abs().sum())}))
