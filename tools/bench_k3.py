#!/usr/bin/env python3
"""K3 (grad_st + stepB + stepC + criterion) alone: register-marching kernel (FOTO_K3=legacy) against the TMA-staged
kernel, algorithmic 80 B/cell, CUDA events around the launches.  Usage: bench_k3.py [h w Nt] [reps]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "optical-flow-optimal-transport_b200"))
import numpy as np, torch, foto_b200
from foto_b200 import synth

h, w, Nt = (int(x) for x in sys.argv[1:4]) if len(sys.argv) >= 4 else (1080, 1920, 16)
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 6
P = h * w
pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
peak = json.load(open(pk))["hbm_gbs"] if os.path.exists(pk) else 6650.0
f0, f1 = synth.make_pair(h, w, seed=0)
d0 = torch.from_numpy(f0).cuda(); d1 = torch.from_numpy(f1).cuda()
o = [torch.empty(P, dtype=torch.float64, device="cuda") for _ in range(3)]
ctx = foto_b200.Context(0)
kw = dict(r=1.0, convergence_tol=0.0, reg_epsilon=1e-3, max_it=reps, backend=foto_b200.POISSON_DCT_EXACT)
res = {}
for mode in ("legacy", "tma", "legacy", "tma"):
    os.environ["FOTO_K3"] = mode
    ctx.solve_dev(d0.data_ptr(), d1.data_ptr(), Nt, w, h, *[t.data_ptr() for t in o], **dict(kw, max_it=1))
    ctx.set_profiling(True); ctx.reset_stats()
    info = ctx.solve_dev(d0.data_ptr(), d1.data_ptr(), Nt, w, h, *[t.data_ptr() for t in o], **kw)
    st = ctx.stats(); ctx.set_profiling(False)
    gbs = 80 * st["prox_cells"] / st["prox_ms"] / 1e6
    k1 = 56 * st["rhs_cells"] / st["rhs_ms"] / 1e6
    res[mode] = torch.stack(o).cpu().numpy()
    print(json.dumps({"grid": [Nt, h, w], "K3": mode, "ms": st["prox_ms"] / reps, "GBs": gbs, "frac": gbs / peak,
                      "K1_GBs": k1, "K1_frac": k1 / peak, "crit_last": float(info["crit"][-1])}), flush=True)
print("max |legacy - tma| over u, v, m:", float(np.abs(res["legacy"] - res["tma"]).max()))
