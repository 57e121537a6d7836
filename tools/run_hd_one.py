#!/usr/bin/env python3
"""One outer ALG2 iteration at 1080x1920xNt with the streaming CG (for ncu captures of HBM-bound kernels).
Usage: run_hd_one.py [cg|dct] [Nt=8] [eps=1e-1]   (`cg 16 1e-3` is the launch bench.py's roofline times)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "optical-flow-optimal-transport_b200"))
import torch, foto_b200
from foto_b200 import synth
h, w, Nt = 1080, 1920, int(sys.argv[2]) if len(sys.argv) > 2 else 8
eps = float(sys.argv[3]) if len(sys.argv) > 3 else 1e-1
backend = foto_b200.POISSON_DCT_EXACT if len(sys.argv) > 1 and sys.argv[1] == "dct" else foto_b200.POISSON_CG_PARITY
f0, f1 = synth.make_pair(h, w, seed=0)
d0 = torch.from_numpy(f0).cuda(); d1 = torch.from_numpy(f1).cuda()
o = [torch.empty(h * w, dtype=torch.float64, device="cuda") for _ in range(3)]
ctx = foto_b200.Context(0); ctx.set_cg_variant(0)
info = ctx.solve_dev(d0.data_ptr(), d1.data_ptr(), Nt, w, h, o[0].data_ptr(), o[1].data_ptr(), o[2].data_ptr(),
                     r=1.0, convergence_tol=0.0, reg_epsilon=eps, max_it=1, backend=backend)
print(info["cg_iters"].tolist(), float(o[0].abs().max()))
