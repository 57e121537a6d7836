"""Achieved parity of the FOTO solve against the reference goldens, per Poisson kernel (needs a GPU).
Prints max relative error of u, v, m and whether the CG iteration counts agree.  usage: python tools/parity_report.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "optical-flow-optimal-transport_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import foto_b200
from conftest import load_golden, relerr

NAMES = ["foto_24x32", "foto_48x64", "foto_97x146", "foto_37x53_nt5", "foto_40x56_nt16_runsh", "foto_31x29_nt2", "foto_squares32", "foto_388x584"]
VARIANTS = [("auto (single-reduction on-chip, Nt <= 8 or 16)  ", -1), ("streaming (textbook recurrences)", 0)]
for name in NAMES:
    g = load_golden(name)
    h, w, Nt = map(int, g["dims"])
    f0, f1 = (g["f0"], g["f1"]) if "f0" in g.files else (None, None)
    if f0 is None:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from test_gpu_parity import _frames
        f0, f1 = _frames(g)
    if "params" in g.files:
        r, tol, eps, max_it = g["params"]
    else:
        r, tol, eps, max_it = 1.0, 0.1, 1e-3, 100
    for label, var in VARIANTS:
        foto_b200.set_default_cg_variant(var)
        u, v, m, info = foto_b200.solve(f0, f1, Nt, w, h, r=float(r), convergence_tol=float(tol), reg_epsilon=float(eps), max_it=int(max_it))
        if "sub" in g.files:
            sub = g["sub"]; u, v, m = u[sub], v[sub], m[sub]
        same = list(info["cg_iters"]) == list(g["cg_iters"])
        print(f"{name:24s} {label:44s} u {relerr(u, g['u']):.1e} v {relerr(v, g['v']):.1e} m {relerr(m, g['m']):.1e}  CG counts equal: {same}", flush=True)
