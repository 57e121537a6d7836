"""One GN solve at 388x584 (for ncu captures).  usage: run_gn_one.py [variant]"""
import sys
sys.path.insert(0, "optical-flow-optimal-transport_b200")
import torch, foto_b200
from foto_b200 import synth
h, w = 388, 584
f0, f1 = synth.make_pair(h, w, seed=7)
ctx = foto_b200.Context(0); ctx.set_cg_variant(int(sys.argv[1]) if len(sys.argv) > 1 else -1)
d0 = torch.from_numpy(f0).cuda(); d1 = torch.from_numpy(f1).cuda()
o = [torch.empty(h * w, dtype=torch.float64, device="cuda") for _ in range(3)]
print(ctx.gn_solve_dev(d0.data_ptr(), d1.data_ptr(), w, h, 0.1, 0.2, *[t.data_ptr() for t in o]))
