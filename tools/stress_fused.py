"""Stress the spin-based edge exchange of the single-reduction kernels: many solves over shapes and seeds, results must be
bitwise reproducible and finite.  usage: python tools/stress_fused.py [repeats]"""
import sys, time
sys.path.insert(0, "optical-flow-optimal-transport_b200")
import numpy as np, torch, foto_b200
from foto_b200 import synth
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
ctx = foto_b200.Context(0)
shapes = [(388, 584, 4), (97, 146, 4), (480, 640, 4), (61, 83, 6), (40, 56, 16), (128, 192, 8), (33, 200, 2), (380, 420, 4), (24, 32, 5)]
t0 = time.time(); n = 0
for (h, w, Nt) in shapes:
    f0, f1 = synth.make_pair(h, w, seed=h + w)
    d0 = torch.from_numpy(f0).cuda(); d1 = torch.from_numpy(f1).cuda()
    o = [torch.empty(h * w, dtype=torch.float64, device="cuda") for _ in range(3)]
    ref = None
    for r in range(reps):
        info = ctx.solve_dev(d0.data_ptr(), d1.data_ptr(), Nt, w, h, *[t.data_ptr() for t in o], max_it=3, convergence_tol=0.0)
        cur = torch.stack(o).cpu().numpy()
        assert np.isfinite(cur).all()
        if ref is None: ref = (cur.copy(), list(info["cg_iters"]))
        else:
            assert list(info["cg_iters"]) == ref[1], (h, w, Nt, r)
            assert np.array_equal(cur, ref[0]), (h, w, Nt, r)
        n += 1
    g = ctx.gn_solve_dev(d0.data_ptr(), d1.data_ptr(), w, h, 0.1, 0.2, *[t.data_ptr() for t in o])
    print(h, w, Nt, "ok, kernel", ctx.stats()["cg_variant"], "cg", ref[1], "gn", g["iters"], flush=True)
print(f"{n} FOTO solves, {time.time() - t0:.1f} s, all bitwise reproducible")
