"""Sweep FOTO_ONCHIP_GRID tile grids for the on-chip CG kernel on one pair (us per CG iteration).
usage: [SWEEP_SHAPE=388,584] python tools/sweep_grid.py [gy,gx ...]   (0,0 = planner's choice)"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "optical-flow-optimal-transport_b200"))
import torch, foto_b200
from foto_b200 import synth
grids = sys.argv[1:] or ["7,21", "4,37", "12,12", "6,24", "8,18", "9,16", "11,13", "10,14", "5,29", "3,49"]
h, w = (int(v) for v in os.environ.get("SWEEP_SHAPE", "388,584").split(","))
Nt = 4
f0, f1 = synth.make_pair(h, w, seed=0)
d0 = torch.from_numpy(f0).cuda(); d1 = torch.from_numpy(f1).cuda()
o = [torch.empty(h * w, dtype=torch.float64, device="cuda") for _ in range(3)]
kw = dict(r=1.0, convergence_tol=0.1, reg_epsilon=1e-3, max_it=100)
for cfg in ["0"]:
    ctx = foto_b200.Context(0); ctx.set_cg_variant(2)
    for g in grids:
        os.environ["FOTO_ONCHIP_GRID"] = g
        try:
            ctx.solve_dev(d0.data_ptr(), d1.data_ptr(), Nt, w, h, *[t.data_ptr() for t in o], **kw)
            ctx.set_profiling(True); ctx.reset_stats()
            ctx.solve_dev(d0.data_ptr(), d1.data_ptr(), Nt, w, h, *[t.data_ptr() for t in o], **kw)
            st = ctx.stats()
            print(f"config {cfg} grid {g}: {1e3 * st['cg_ms'] / st['cg_iterations']:.3f} us/iter, checksum {float(o[0].abs().sum()):.12e}", flush=True)
        except Exception as e:
            print(f"config {cfg} grid {g}: {type(e).__name__}: {e}", flush=True)
    ctx.close()
