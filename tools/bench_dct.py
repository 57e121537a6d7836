#!/usr/bin/env python3
"""FOTO pairs/s with the dct_exact Poisson back-end (opt-in fast mode) on 388x584 pairs."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "optical-flow-optimal-transport_b200"))
import numpy as np, torch, foto_b200
from foto_b200 import synth
h, w, Nt = 388, 584, 4
P = h * w
ctx = foto_b200.Context(0)
pairs = [synth.make_pair(h, w, seed=i) for i in range(4)]
d0 = [torch.from_numpy(a).cuda() for a, _ in pairs]; d1 = [torch.from_numpy(b).cuda() for _, b in pairs]
du, dv, dm = (torch.empty(P, dtype=torch.float64, device="cuda") for _ in range(3))
kw = dict(r=1.0, convergence_tol=0.1, reg_epsilon=1e-3, max_it=100)
for backend, name in ((foto_b200.POISSON_DCT_EXACT, "dct_exact"), (foto_b200.POISSON_CG_PARITY, "cg_parity")):
    for i in range(4):
        ctx.solve_dev(d0[i].data_ptr(), d1[i].data_ptr(), Nt, w, h, du.data_ptr(), dv.data_ptr(), dm.data_ptr(), backend=backend, **kw)
    ctx.set_profiling(True); ctx.reset_stats()
    ctx.event_record(0)
    outer = 0
    reps = 5 if backend == foto_b200.POISSON_DCT_EXACT else 1
    for _ in range(reps):
        for i in range(4):
            outer += ctx.solve_dev(d0[i].data_ptr(), d1[i].data_ptr(), Nt, w, h, du.data_ptr(), dv.data_ptr(), dm.data_ptr(), backend=backend, **kw)["n_outer"]
    ctx.event_record(1)
    ms = ctx.event_elapsed_ms(); st = ctx.stats()
    print(json.dumps({"backend": name, "pairs_per_s": 4 * reps / (ms / 1e3), "ms_per_pair": ms / (4 * reps), "outer_per_pair": outer / (4 * reps),
                      "poisson_ms_per_solve": st["cg_ms"] / st["cg_launches"], "rhs_ms": st["rhs_ms"] / st["cg_launches"],
                      "prox_ms": st["prox_ms"] / st["cg_launches"],
                      "poisson_GFLOPs": (3.5e9 / (st["cg_ms"] / st["cg_launches"] / 1e3) / 1e9) if name == "dct_exact" else None}))
    ctx.set_profiling(False)
