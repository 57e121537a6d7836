#!/usr/bin/env python3
"""Per-phase cycle breakdown of the on-chip CG kernel (CTA 0) + per-iteration time of both
CG variants on one 388x584 pair.  Usage: python tools/prof_onchip.py [h w Nt]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "optical-flow-optimal-transport_b200"))
import numpy as np
import torch
import foto_b200
from foto_b200 import synth

h, w, Nt = (int(x) for x in sys.argv[1:4]) if len(sys.argv) >= 4 else (388, 584, 4)
P = h * w
f0, f1 = synth.make_pair(h, w, seed=0)
d0 = torch.from_numpy(f0).cuda(); d1 = torch.from_numpy(f1).cuda()
du, dv, dm = (torch.empty(P, dtype=torch.float64, device="cuda") for _ in range(3))
kw = dict(r=1.0, convergence_tol=0.1, reg_epsilon=1e-3, max_it=100)
res = {}
for variant, cfg in ((0, None), (2, None)):
    ctx = foto_b200.Context(0)
    ctx.set_cg_variant(variant)
    ctx.solve_dev(d0.data_ptr(), d1.data_ptr(), Nt, w, h, du.data_ptr(), dv.data_ptr(), dm.data_ptr(), **kw)
    ctx.set_profiling(True); ctx.reset_stats()
    info = ctx.solve_dev(d0.data_ptr(), d1.data_ptr(), Nt, w, h, du.data_ptr(), dv.data_ptr(), dm.data_ptr(), **kw)
    st = ctx.stats()
    if variant >= 1:          # second run with the in-kernel phase counters on (slightly intrusive)
        ctx.onchip_prof(True)
        ctx.solve_dev(d0.data_ptr(), d1.data_ptr(), Nt, w, h, du.data_ptr(), dv.data_ptr(), dm.data_ptr(), **kw)
    res[variant] = du.cpu().numpy().copy()
    print(f"variant {variant} config {cfg}: outer {info['n_outer']} cg {info['cg_iters'].tolist()} "
          f"cg_ms {st['cg_ms']:.3f} us/iter {1e3 * st['cg_ms'] / st['cg_iterations']:.3f}")
    if variant >= 1:
        c = ctx.onchip_prof(False)
        c = c[c[:, 6] > 0].astype(float)
        it = c[:, 6:7]
        names = ["stencil", "reduce+export+x+import", "all-reduce wait", "p,s,r + ring update", "import rounds (thread 32)", "import cycles (thread 32)"]
        per = c[:, :6] / it
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        np.savetxt(os.path.join(ROOT, "gpurun_out", f"phase_cycles_variant{variant}_cfg{cfg}.csv"), per, fmt="%.0f", delimiter=",")
        print(f"   {len(c)} CTAs, cycles/iteration      CTA0      min   median      max   argmax")
        for k, n in enumerate(names):
            col = per[:, k]
            print(f"   {n:20s} {col[0]:9.0f} {col.min():8.0f} {np.median(col):8.0f} {col.max():8.0f} {int(col.argmax()):6d}")
        tot = per[:, :4].sum(axis=1)
        print(f"   {'total':20s} {tot[0]:9.0f} {tot.min():8.0f} {np.median(tot):8.0f} {tot.max():8.0f}")
        comp = per[:, [0, 3]].sum(axis=1)
        print(f"   compute (no barriers, no x): min {comp.min():.0f} median {np.median(comp):.0f} max {comp.max():.0f} at CTA {int(comp.argmax())}")
    if variant >= 1:
        print("   max |u_onchip - u_stream| / max|u| =", float(np.max(np.abs(res[0] - res[variant])) / np.max(np.abs(res[0]))))
    ctx.close()
