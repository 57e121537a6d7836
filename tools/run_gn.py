"""Time the GN solve (K5 + K6) at one size with both PCG kernels.  usage: run_gn.py [h w [reps]]"""
import sys, time
import numpy as np, torch
sys.path.insert(0, "optical-flow-optimal-transport_b200")
import foto_b200
from foto_b200 import synth

h, w = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (388, 584)
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
f0, f1 = synth.make_pair(h, w, seed=7)
ctx = foto_b200.Context(0)
d0 = torch.tensor(f0.ravel(), dtype=torch.float64, device="cuda")
d1 = torch.tensor(f1.ravel(), dtype=torch.float64, device="cuda")
out = [torch.empty(h * w, dtype=torch.float64, device="cuda") for _ in range(3)]
res = {}
for name, var in (("streaming", 0), ("auto", -1)):
    ctx.set_cg_variant(var)
    try:
        for i in range(reps + 1):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            r = ctx.gn_solve_dev(d0.data_ptr(), d1.data_ptr(), w, h, 0.1, 0.2, *[o.data_ptr() for o in out]); it, info = r["iters"], r["info"]
            torch.cuda.synchronize(); dt = time.perf_counter() - t0
        res[name] = torch.stack(out).cpu().numpy().copy()
        print(f"{name}: {dt*1e3:.2f} ms, {it} iterations, info {info}, {dt*1e6/max(it,1):.2f} us/iter", flush=True)
    except Exception as e:
        print(name, "failed:", e)
if len(res) == 2:
    a, b = res["streaming"], res["auto"]
    print("rel diff auto vs streaming:", np.abs(a - b).max() / np.abs(a).max())
