#!/usr/bin/env python3
"""A few dct_exact FOTO solves on one 388x584 pair (for ncu launch lists)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "optical-flow-optimal-transport_b200"))
import torch, foto_b200
from foto_b200 import synth
h, w, Nt = 388, 584, 4
f0, f1 = synth.make_pair(h, w, seed=0)
d0 = torch.from_numpy(f0).cuda(); d1 = torch.from_numpy(f1).cuda()
du, dv, dm = (torch.empty(h * w, dtype=torch.float64, device="cuda") for _ in range(3))
ctx = foto_b200.Context(0)
for _ in range(2):
    info = ctx.solve_dev(d0.data_ptr(), d1.data_ptr(), Nt, w, h, du.data_ptr(), dv.data_ptr(), dm.data_ptr(),
                         r=1.0, convergence_tol=0.0, reg_epsilon=1e-3, max_it=3, backend=foto_b200.POISSON_DCT_EXACT)
print(info["n_outer"], float(du.abs().max()))
