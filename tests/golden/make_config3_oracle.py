#!/usr/bin/env python3
"""Oracle results for every pair that carries a published number: the 64 pairs of BASELINE.json config 3 and the
bench seeds (bench.pair_seed, 8 ranks x 16 pairs: pairs 0-3 of every rank in bench_seeds_oracle.npz, pairs 4-15 in
bench_seeds_oracle_b16.npz with a coarser sub-sample).  CPU only, test infrastructure: runs oracle/foto_oracle.c (the C
restatement of the reference, pinned to the reference's own outputs by tests/test_oracle_golden.py) with
`multiprocessing`, and writes compact goldens the `-m gpu` tests compare the CUDA path with:

    tests/golden/config3_oracle.npz   per pair: n_outer, cg_iters[n_outer], crit[n_outer], u/v/m on every 101st
                                      pixel, and (sum, sum|.|, max|.|, l2) of each full field
    tests/golden/bench_seeds_oracle.npz, bench_seeds_oracle_b16.npz   the same for the bench pairs (b16: every 1009th pixel)

    python tests/golden/make_config3_oracle.py [--procs 7] [--only config3|bench|bench16]
"""
import argparse, multiprocessing as mp, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "optical-flow-optimal-transport_b200"))
import numpy as np
from foto_b200 import synth

SUB = 101


def stats(a):
    return np.array([a.sum(), np.abs(a).sum(), np.abs(a).max(), np.sqrt((a * a).sum())])


def work(job):
    import oracle
    key, h, w, f0, f1, kw = job[:6]
    stride = job[6] if len(job) > 6 else SUB
    t0 = time.time()
    u, v, m, info = oracle.solve(f0, f1, 4, w, h, return_info=True, **kw)
    sub = np.arange(0, h * w, stride)
    out = {f"{key}/dims": np.array([h, w, 4]), f"{key}/n_outer": np.int32(info["n_outer"]),
           f"{key}/cg_iters": info["cg_iters"].astype(np.int32), f"{key}/crit": info["crit"],
           f"{key}/seconds": np.float64(time.time() - t0)}
    for n, a in (("u", u), ("v", v), ("m", m)):
        out[f"{key}/{n}"] = a[sub]; out[f"{key}/{n}_stats"] = stats(a)
    print(f"  {key}: {info['n_outer']} outer, cg {info['cg_iters'][:3].tolist()}..{info['cg_iters'][-1]}, {time.time() - t0:.0f} s", flush=True)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--procs", type=int, default=7)
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    import oracle
    oracle.build()
    kw = dict(synth.CONFIG3_PARAMS)
    sets = {}
    if a.only in ("", "bench"):
        import bench
        sets["bench_seeds_oracle"] = [(f"rank{r}/pair{i}", bench.H, bench.W, *synth.make_pair(bench.H, bench.W, seed=bench.pair_seed(r, i)), dict(bench.PARAMS))
                                      for r in range(8) for i in range(4)]
    if a.only in ("", "bench16"):
        import bench
        sets["bench_seeds_oracle_b16"] = [(f"rank{r}/pair{i}", bench.H, bench.W, *synth.make_pair(bench.H, bench.W, seed=bench.pair_seed(r, i)),
                                           dict(bench.PARAMS), 1009) for r in range(8) for i in range(4, 16)]
    if a.only in ("", "config3"):
        # longest jobs first (perturbed pairs run to max_it)
        jobs = [(n, h, w, f0, f1, kw) for (n, h, w, f0, f1) in synth.config3_pairs()]
        jobs.sort(key=lambda j: (j[0].endswith("/0"), -j[1] * j[2]))
        sets["config3_oracle"] = jobs
    with mp.get_context("fork").Pool(a.procs) as pool:
        for name, jobs in sets.items():
            t0 = time.time()
            merged = {"sub_stride": np.int32(jobs[0][6] if len(jobs[0]) > 6 else SUB), "params": np.array([jobs[0][5][k] for k in ("r", "convergence_tol", "reg_epsilon", "max_it")])}
            for out in pool.imap_unordered(work, jobs, chunksize=1):
                merged.update(out)
            path = os.path.join(ROOT, "tests", "golden", name + ".npz")
            np.savez_compressed(path, **merged)
            print(f"wrote {path} ({os.path.getsize(path) / 1e6:.2f} MB) in {time.time() - t0:.0f} s", flush=True)


if __name__ == "__main__":
    main()
