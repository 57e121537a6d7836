#!/usr/bin/env python3
"""Small instances of every entry point, for `compute-sanitizer --tool memcheck` (one tool per call)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "optical-flow-optimal-transport_b200"))
import numpy as np, foto_b200
from foto_b200 import synth
rng = np.random.default_rng(0)
h, w, Nt = 21, 30, 3
f0, f1 = synth.make_pair(h, w, seed=1)
kw = dict(r=1.0, convergence_tol=0.1, reg_epsilon=1e-2, max_it=2)
for variant in (-1, 0):
    foto_b200.set_default_cg_variant(variant)
    u, v, m, info = foto_b200.solve(f0, f1, Nt, w, h, **kw)
    print("solve variant", variant, info["cg_iters"].tolist(), float(np.abs(u).max()))
foto_b200.set_default_cg_variant(-1)
u, v, m, info = foto_b200.solve(f0, f1, Nt, w, h, backend=foto_b200.POISSON_DCT_EXACT, **kw)
print("dct", float(np.abs(u).max()))
u, v, m, info = foto_b200.solve(f0, f1, Nt, w, h, r=0.7, convergence_tol=0.1, reg_epsilon=1e-2, max_it=2)
print("general r", info["cg_iters"].tolist())
gu, gv, gm, gi = foto_b200.gn_solve(f0, f1, w, h, 0.1, 0.2)
print("gn", gi)
N = Nt * h * w
print("stepB", foto_b200.stepB(rng.standard_normal(3 * N), Nt, w, h).shape)
print("flow", foto_b200.flow_from_phi(rng.standard_normal(N), Nt, w, h)[0].shape)
print("warp", foto_b200.warp_apply(f0, u * 3, v * 3, w, h, m).shape)
for op in ("grad_st", "div_st", "laplacian_st", "grad", "div", "grad_forward"):
    n_out, n_in = foto_b200.op_shape(op, Nt, w, h)
    x = rng.standard_normal(n_in)
    foto_b200.op_apply(op, "N", Nt, w, h, 1, 1, 1, x); foto_b200.op_apply(op, "D", Nt, w, h, 1, 1, 1, rng.standard_normal(n_out), transpose=True)
print("ops ok")
print("metrics", foto_b200.flow_metrics(u, v, u + 0.1, v - 0.1)); print("flo", foto_b200.pack_flo(u, v).shape)
us, vs, ms, outer = foto_b200.solve_batch(np.stack([f0, f1]), np.stack([f1, f0]), Nt, w, h, devices=[0], **kw)
print("batch", outer.tolist())
