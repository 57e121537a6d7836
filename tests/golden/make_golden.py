#!/usr/bin/env python3
"""Generate the golden vectors under tests/golden/ by RUNNING THE UNMODIFIED REFERENCE.

The reference (a pure-Python repo) ships no tests and no golden vectors,
so the only way to pin the oracle and the CUDA path is to import the reference here
(`/root/reference`, read-only, present in the build container only) and record what it
computes on small seeded inputs.  The outputs are committed; this script is committed so
they can be regenerated.  Nothing under tests/ reads /root/reference at test time.

    python tests/golden/make_golden.py            # small cases  (~2 min)
    python tests/golden/make_golden.py --full     # + 388x584 FOTO/GN (~6 min, 5 GB RAM)

Reference entry points exercised (file:line in /root/reference):
  operators.py:5-191          every 1-D builder and Kronecker operator
  benamou_brenier.py:93-149   stepB
  benamou_brenier.py:26-91    solve_benamou_brenier_step (scipy cg, rtol 1e-6)
  benamou_brenier.py:151-271  solve
  classical.py:68-130         GLLOpticalFlow.assemble/process (SuperLU)
  utils.py:44-99,148-183      reconstructTrajectory / opticalflow_from_benamoubrenier
  utils.py:186-248            apply_opticalflow
"""
import argparse
import contextlib
import io
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("FOTO_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
sys.path.insert(1, os.path.join(ROOT, "optical-flow-optimal-transport_b200"))

import scipy  # noqa: E402
import scipy.sparse.linalg  # noqa: E402
import operators as ref_ops  # noqa: E402
import benamou_brenier as ref_bb  # noqa: E402
import classical as ref_gn  # noqa: E402
import utils as ref_utils  # noqa: E402
from foto_b200 import synth  # noqa: E402

assert os.path.realpath(ref_bb.__file__).startswith(os.path.realpath(REF)), ref_bb.__file__

_real_cg = scipy.sparse.linalg.cg
CG_LOG = []          # iteration count of every cg() call the reference makes
CG_RTOL = [None]     # None = as the reference passes it (1e-6); else override ("tight" oracle)


def _counting_cg(A, b, **kw):
    n = [0]

    def cb(_xk):
        n[0] += 1
    if CG_RTOL[0] is not None:
        kw["rtol"] = CG_RTOL[0]
        kw["maxiter"] = 100000
    x, info = _real_cg(A, b, callback=cb, **kw)
    CG_LOG.append(n[0])
    return x, info


ref_bb.cg = _counting_cg   # same algorithm; only counts iterations (and optional rtol override)


def u8(f):
    q = np.round(np.asarray(f) * 255.0)
    assert np.array_equal(q / 255.0, f), "input is not on the 8-bit lattice"
    return q.astype(np.uint8)


def run_solve(f0, f1, Nt, w, h, **kw):
    CG_LOG.clear()
    buf = io.StringIO()
    t0 = time.time()
    with contextlib.redirect_stdout(buf):
        u, v, m = ref_bb.solve(f0, f1, Nt, w, h, **kw)
    dt = time.time() - t0
    crit = [float(line.split(" ")[0]) for line in buf.getvalue().splitlines()
            if line.endswith(")") and "/" in line and not line.startswith("WARNING")]
    return dict(u=u, v=v, m=m, crit=np.array(crit), cg_iters=np.array(CG_LOG, dtype=np.int32),
                seconds=np.float64(dt))


def save(name, **arrs):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrs)
    print(f"  wrote {name}.npz ({os.path.getsize(path)/1024:.1f} KiB)")


# ----------------------------------------------------------------------------- operators
def gold_operators():
    rng = np.random.default_rng(101)
    out = {}
    builders = ["grad_1d_forward_weird", "grad_1d_backward_weird", "grad_1d_central_weird",
                "grad_1d_central", "grad_1d_forward", "grad_1d_backward", "lap1d"]
    for b in builders:
        for bc in "ND":
            for (n, h) in [(5, 1.0), (6, 0.5), (2, 2.0)]:
                out[f"dense/{b}/{bc}/{n}/{h}"] = np.asarray(getattr(ref_ops, b)(n, h, bc).todense())
    grids = [(3, 7, 5, 1.0, 1.0, 1.0), (4, 16, 9, 0.5, 2.0, 1.5), (2, 2, 2, 1.0, 1.0, 1.0)]
    for (Nt, Nx, Ny, dt, dx, dy) in grids:
        N = Nt * Nx * Ny
        P = Nx * Ny
        key = f"{Nt}x{Ny}x{Nx}/{dt}/{dx}/{dy}"
        xN = rng.standard_normal(N); x3N = rng.standard_normal(3 * N)
        xP = rng.standard_normal(P); x2P = rng.standard_normal(2 * P)
        out[f"in/{key}/N"] = xN; out[f"in/{key}/3N"] = x3N
        out[f"in/{key}/P"] = xP; out[f"in/{key}/2P"] = x2P
        for bc in "ND":
            G = ref_ops.grad_st(Nt, Nx, Ny, dt, dx, dy, bc)
            D = ref_ops.div_st(Nt, Nx, Ny, dt, dx, dy, bc)
            L = ref_ops.laplacian_st(Nt, Nx, Ny, dt, dx, dy, bc)
            out[f"grad_st/{bc}/{key}"] = G @ xN
            out[f"grad_st.T/{bc}/{key}"] = G.transpose() @ x3N
            out[f"div_st/{bc}/{key}"] = D @ x3N
            out[f"div_st.T/{bc}/{key}"] = D.transpose() @ xN
            out[f"laplacian_st/{bc}/{key}"] = L @ xN
            g = ref_ops.grad(Nx, Ny, dx, dy, bc)
            gf = ref_ops.grad_forward(Nx, Ny, dx, dy, bc)
            d = ref_ops.div(Nx, Ny, dx, dy, bc)
            out[f"grad/{bc}/{key}"] = g @ xP
            out[f"grad.T/{bc}/{key}"] = g.transpose() @ x2P
            out[f"grad_forward/{bc}/{key}"] = gf @ xP
            out[f"grad_forward.T/{bc}/{key}"] = gf.transpose() @ x2P
            out[f"div/{bc}/{key}"] = d @ x2P
            out[f"div.T/{bc}/{key}"] = d.transpose() @ xP
    save("operators", **out)


# ----------------------------------------------------------------------------- stepB
def gold_stepB():
    rng = np.random.default_rng(202)
    # crafted cells: inside K, boundary, single root, triple root (alpha < -1 with small rho),
    # beta = 0, alpha = -1, large magnitudes, negative betas in all quadrants
    crafted = np.array([
        [-1.0, 0.5, 0.5], [-0.25, 0.5, 0.5], [-0.5, 1.0, 0.0], [0.0, 0.0, 0.0],
        [0.3, 0.0, 0.0], [1.0, 0.0, 0.0], [-1.0, 3.0, -4.0], [-1.0, 0.0, 2.0],
        [-10.0, 5.0, 0.0], [-10.0, 3.0, 4.0], [-10.0, -3.0, -4.0], [-4.5, 2.0, 2.5],
        [-50.0, 10.1, 0.0], [-2.0, 1.5, -0.2], [-1.5, 0.9, 0.5], [0.5, -2.0, 3.0],
        [2.0, 1e-8, -1e-8], [1e-9, 1e-9, 0.0], [-1e-9, 1e-3, 1e-3], [100.0, -250.0, 40.0],
        [-3.0, 0.0, 2.6], [-3.0, 2.6, 0.0], [0.0, 1e-160, 0.0], [5.0, 0.0, -7.0],
    ])
    rnd = rng.standard_normal((600, 3)) * np.array([3.0, 2.0, 2.0])
    rnd[:200, 0] -= 6.0       # push a third of the random cells towards the triple-root wedge
    cells = np.vstack([crafted, rnd])
    n = cells.shape[0]
    Nt, Ny, Nx = 2, 2, n // 4
    assert Nt * Ny * Nx == n
    p = np.concatenate([cells[:, 0], cells[:, 1], cells[:, 2]])
    q = ref_bb.stepB(p, Nt, Nx, Ny)
    a, b1, b2 = cells[:, 0], cells[:, 1], cells[:, 2]
    rho2 = b1 ** 2 + b2 ** 2
    inside = 2 * a + rho2 <= 0
    single = (~inside) & (-32 * (a + 1) ** 3 - 108 * rho2 < 0)
    triple = (~inside) & (~single)
    print(f"  stepB census: inside {inside.sum()}, single-root {single.sum()}, triple-root {triple.sum()}")
    assert inside.sum() > 5 and single.sum() > 5 and triple.sum() > 5
    save("stepB", p=p, q=q, dims=np.array([Nt, Nx, Ny]))


# ----------------------------------------------------------------------------- stepA
def gold_stepA():
    rng = np.random.default_rng(303)
    out = {}
    for tag, (Nt, Ny, Nx, r, eps) in {"a": (4, 24, 32, 1.0, 1e-3), "b": (5, 17, 23, 0.7, 1e-2),
                                       "c": (2, 9, 11, 2.0, 1e-1)}.items():
        N = Nt * Nx * Ny
        mu = rng.standard_normal(3 * N); q = rng.standard_normal(3 * N)
        rho0 = rng.random(Nx * Ny); rhoT = rng.random(Nx * Ny)
        L = ref_ops.laplacian_st(Nt, Nx, Ny, 1, 1, 1, bc="N")
        D = ref_ops.div_st(Nt, Nx, Ny, 1, 1, 1, bc="N")
        A = -r * L + r * eps * scipy.sparse.eye(N)
        CG_LOG.clear()
        phi = ref_bb.solve_benamou_brenier_step(mu, q, rho0, rhoT, r, A, D, Nt, Nx, Ny, 1, 1, 1)
        # the right-hand side alone (CG-independent): the same formula, re-evaluated by the
        # reference's own operator; kept so that K1 can be tested without the solver
        F = D @ (mu - r * q)
        P = Nx * Ny
        F[:P] -= rho0 - mu[:P] + r * q[:P]
        F[(Nt - 1) * P:N] += rhoT - mu[(Nt - 1) * P:N] + r * q[(Nt - 1) * P:N]
        out.update({f"{tag}/dims": np.array([Nt, Nx, Ny]), f"{tag}/r_eps": np.array([r, eps]),
                    f"{tag}/mu": mu, f"{tag}/q": q, f"{tag}/rho0": rho0, f"{tag}/rhoT": rhoT,
                    f"{tag}/phi": phi, f"{tag}/F": F, f"{tag}/cg_iters": np.array(CG_LOG, dtype=np.int32)})
        print(f"  stepA {tag}: cg iterations {CG_LOG}")
    save("stepA", **out)


# ----------------------------------------------------------------------------- end to end
FOTO_CASES = {
    # tag: (h, w, Nt, seed, kwargs)
    "foto_24x32": (24, 32, 4, 1, dict(r=1.0, convergence_tol=0.1, reg_epsilon=1e-3, max_it=100)),
    "foto_48x64": (48, 64, 4, 2, dict(r=1.0, convergence_tol=0.1, reg_epsilon=1e-3, max_it=100)),
    "foto_37x53_nt5": (37, 53, 5, 3, dict(r=0.7, convergence_tol=0.05, reg_epsilon=1e-2, max_it=30)),
    "foto_40x56_nt16_runsh": (40, 56, 16, 4, dict(r=1.0, convergence_tol=0.01, reg_epsilon=1e-2, max_it=10)),
    "foto_97x146": (97, 146, 4, 0, dict(r=1.0, convergence_tol=0.1, reg_epsilon=1e-3, max_it=100)),
    "foto_31x29_nt2": (31, 29, 2, 5, dict(r=1.0, convergence_tol=0.1, reg_epsilon=1e-3, max_it=8)),
}


def gold_foto():
    for tag, (h, w, Nt, seed, kw) in FOTO_CASES.items():
        f0, f1 = synth.make_pair(h, w, seed=seed)
        res = run_solve(f0, f1, Nt, w, h, **kw)
        print(f"  {tag}: {len(res['crit'])} outer, cg {res['cg_iters'].tolist()}, {float(res['seconds']):.1f}s")
        save(tag, f0_u8=u8(f0), f1_u8=u8(f1), dims=np.array([h, w, Nt]),
             params=np.array([kw["r"], kw["convergence_tol"], kw["reg_epsilon"], kw["max_it"]]), **res)
    # the author's commented-out two-squares fixture (main.py:55-65)
    f0, f1 = synth.two_squares(32)
    kw = dict(r=1.0, convergence_tol=0.1, reg_epsilon=1e-3, max_it=12)
    res = run_solve(f0, f1, 4, 32, 32, **kw)
    print(f"  foto_squares32: {len(res['crit'])} outer, cg {res['cg_iters'].tolist()}")
    save("foto_squares32", f0_u8=u8(f0), f1_u8=u8(f1), dims=np.array([32, 32, 4]),
         params=np.array([1.0, 0.1, 1e-3, 12]), **res)
    # "tight" oracle: same reference code with the inner CG run to rtol 1e-13 (exact-solve limit;
    # the spsolve the author left commented at benamou_brenier.py:84).  Gate for dct_exact.
    CG_RTOL[0] = 1e-13
    for tag in ("foto_24x32", "foto_48x64", "foto_37x53_nt5"):
        h, w, Nt, seed, kw = FOTO_CASES[tag]
        f0, f1 = synth.make_pair(h, w, seed=seed)
        res = run_solve(f0, f1, Nt, w, h, **kw)
        print(f"  {tag}_tight: {len(res['crit'])} outer, cg {res['cg_iters'].tolist()}")
        save(tag + "_tight", f0_u8=u8(f0), f1_u8=u8(f1), dims=np.array([h, w, Nt]),
             params=np.array([kw["r"], kw["convergence_tol"], kw["reg_epsilon"], kw["max_it"]]), **res)
    CG_RTOL[0] = None


GN_CASES = {
    "gn_24x32": (24, 32, 1, 0.1, 0.2),
    "gn_48x64": (48, 64, 2, 0.1, 0.2),
    "gn_37x53": (37, 53, 3, 0.5, 1.0),
    "gn_97x146": (97, 146, 0, 0.1, 0.2),
}


def run_gn(f0, f1, w, h, alpha, lam):
    t0 = time.time()
    s = ref_gn.GLLOpticalFlow(w, h)
    s.setAlpha(alpha); s.setLambda(lam)
    s.assemble(f0, f1)
    A, b = s.A, s.b
    u, v, m = s.process()
    return dict(u=u, v=v, m=m, seconds=np.float64(time.time() - t0)), A, b


def gold_gn():
    rng = np.random.default_rng(404)
    for tag, (h, w, seed, alpha, lam) in GN_CASES.items():
        f0, f1 = synth.make_pair(h, w, seed=seed)
        if tag == "gn_37x53":
            f1 = synth.perturb_brightness(f1, h, w, seed=7)
        res, A, b = run_gn(f0, f1, w, h, alpha, lam)
        extra = {}
        if h * w <= 48 * 64:   # pin the operator itself: A @ x on a random x, and b
            x = rng.standard_normal(3 * h * w)
            extra = dict(x_probe=x, Ax_probe=A @ x, b=b)
        print(f"  {tag}: |u|max {np.abs(res['u']).max():.3f} mean u {res['u'].mean():.3f} v {res['v'].mean():.3f}")
        save(tag, f0_u8=u8(f0), f1_u8=u8(f1), dims=np.array([h, w]), params=np.array([alpha, lam]),
             **res, **extra)


# ----------------------------------------------------------------------------- flow / warp
def gold_flow():
    rng = np.random.default_rng(505)
    out = {}
    for tag, (Nt, Ny, Nx, amp) in {"a": (5, 20, 31, 0.6), "b": (2, 7, 9, 2.0), "c": (9, 16, 12, 0.25), "d": (4, 33, 47, 1.0)}.items():
        from scipy import ndimage
        phi = ndimage.gaussian_filter(rng.standard_normal((Nt, Ny, Nx)), (0.5, 1.5, 1.5)) * amp * 4
        phi = phi.ravel()
        grad = ref_ops.grad(Nx, Ny, 1, 1, bc="N")
        div = ref_ops.div(Nx, Ny, 1, 1, bc="D")
        u, v, m = ref_utils.opticalflow_from_benamoubrenier(phi, Nt, Nx, Ny, grad, div)
        print(f"  flow {tag}: |u|max {np.abs(u).max():.2f} |v|max {np.abs(v).max():.2f}")
        out.update({f"{tag}/dims": np.array([Nt, Nx, Ny]), f"{tag}/phi": phi,
                    f"{tag}/u": u, f"{tag}/v": v, f"{tag}/m": m})
    save("flow", **out)


def gold_warp():
    rng = np.random.default_rng(606)
    out = {}
    for tag, (h, w, amp) in {"a": (19, 27, 1.5), "b": (12, 8, 9.0), "c": (2, 2, 1.0), "d": (33, 21, 40.0)}.items():
        f1 = np.round(rng.random(h * w) * 255) / 255
        u = rng.standard_normal(h * w) * amp
        v = rng.standard_normal(h * w) * amp
        u[:: 7] = np.round(u[:: 7])          # exact-integer displacements too
        v[:: 5] = np.round(v[:: 5])
        m = rng.standard_normal(h * w) * 0.1
        out.update({f"{tag}/dims": np.array([h, w]), f"{tag}/f1": f1, f"{tag}/u": u, f"{tag}/v": v,
                    f"{tag}/m": m,
                    f"{tag}/out_m": ref_utils.apply_opticalflow(f1.copy(), u, v, w, h, m),
                    # m = 0 stands in for "no luminosity": with the installed numpy the reference's
                    # default m=np.array([None]) path raises TypeError (utils.py:202-203), and its
                    # only caller always passes m (main.py:110)
                    f"{tag}/out_m0": ref_utils.apply_opticalflow(f1.copy(), u, v, w, h, np.zeros(h * w))})
    save("warp", **out)


# ----------------------------------------------------------------------------- full size
def gold_full():
    h, w, Nt = 388, 584, 4
    f0, f1 = synth.make_pair(h, w, seed=0)
    sub = np.arange(0, h * w, 13)          # every 13th pixel (13 is coprime with w = 584)
    kw = dict(r=1.0, convergence_tol=0.1, reg_epsilon=1e-3, max_it=100)
    print("  FOTO 388x584 (about 260 s) ...", flush=True)
    res = run_solve(f0, f1, Nt, w, h, **kw)
    print(f"  foto_388x584: {len(res['crit'])} outer, cg {res['cg_iters'].tolist()}, {float(res['seconds']):.1f}s")
    stats = {k + "_stats": np.array([res[k].sum(), np.abs(res[k]).sum(), np.abs(res[k]).max(),
                                     np.sqrt((res[k] ** 2).sum())]) for k in "uvm"}
    save("foto_388x584", f0_u8=u8(f0), f1_u8=u8(f1), dims=np.array([h, w, Nt]),
         params=np.array([1.0, 0.1, 1e-3, 100]), sub=sub, u=res["u"][sub], v=res["v"][sub], m=res["m"][sub],
         crit=res["crit"], cg_iters=res["cg_iters"], seconds=res["seconds"], **stats)
    print("  GN 388x584 (about 60 s, 4.5 GB) ...", flush=True)
    g, _, _ = run_gn(f0, f1, w, h, 0.1, 0.2)
    stats = {k + "_stats": np.array([g[k].sum(), np.abs(g[k]).sum(), np.abs(g[k]).max(),
                                     np.sqrt((g[k] ** 2).sum())]) for k in "uvm"}
    save("gn_388x584", dims=np.array([h, w]), params=np.array([0.1, 0.2]), sub=sub,
         u=g["u"][sub], v=g["v"][sub], m=g["m"][sub], seconds=g["seconds"], **stats)


# ----------------------------------------------------------------------------- config-3 shapes, dataset tools, metrics
def _full_case(tag, name, h, w, f0, f1, kw):
    print(f"  {tag}: reference FOTO {h}x{w} ...", flush=True)
    res = run_solve(f0, f1, 4, w, h, **kw)
    sub = np.arange(0, h * w, 13)
    stats = {k + "_stats": np.array([res[k].sum(), np.abs(res[k]).sum(), np.abs(res[k]).max(),
                                     np.sqrt((res[k] ** 2).sum())]) for k in "uvm"}
    print(f"  {tag} ({name}): {len(res['crit'])} outer, cg {res['cg_iters'].tolist()}, {float(res['seconds']):.1f}s")
    save(tag, f0_u8=u8(f0), f1_u8=u8(f1), dims=np.array([h, w, 4]),
         params=np.array([kw["r"], kw["convergence_tol"], kw["reg_epsilon"], kw["max_it"]]), sub=sub,
         u=res["u"][sub], v=res["v"][sub], m=res["m"][sub], crit=res["crit"], cg_iters=res["cg_iters"],
         seconds=res["seconds"], **stats)


def gold_config3ref():
    """One pair per config-3 shape that config 1 does not cover, from the unmodified reference: Grove2/0 (480x640, the
    shape that runs the large on-chip CG variant with x in global memory) and Venus/0 (380x420)."""
    kw = dict(synth.CONFIG3_PARAMS)
    for seq, tag in (("Grove2", "foto_480x640_grove2"), ("Venus", "foto_380x420_venus")):
        (name, h, w, f0, f1), = synth.config3_pairs(sequences=[seq], perturbations=[0])
        _full_case(tag, name, h, w, f0, f1, kw)


def gold_lum():
    """Outputs of the reference's dataset tools themselves (bin/create_lum_dataset.py, bin/normalize_image.py, run as
    scripts on PNG files) and of the metric functions (utils.py:294-354)."""
    import subprocess
    import tempfile
    from PIL import Image
    rng = np.random.default_rng(707)
    out = {}
    env = dict(os.environ, PYTHONPATH=REF)
    with tempfile.TemporaryDirectory() as tmp:
        for tag, (h, w) in {"a": (40, 56), "b": (97, 146), "c": (33, 21)}.items():
            a = rng.integers(0, 256, (h, w), dtype=np.uint8); b = rng.integers(0, 200, (h, w), dtype=np.uint8)
            pa, pb, po, po2 = (os.path.join(tmp, n) for n in ("a.png", "b.png", "o.png", "o2.png"))
            Image.fromarray(a, "L").save(pa); Image.fromarray(b, "L").save(pb)
            out[f"{tag}/a"] = a; out[f"{tag}/b"] = b
            for seed in (1, 12346, 12352):
                subprocess.check_call([sys.executable, os.path.join(REF, "bin", "create_lum_dataset.py"), pa, po, str(seed)], env=env)
                out[f"{tag}/lum/{seed}"] = np.asarray(Image.open(po).convert("L"))
            subprocess.check_call([sys.executable, os.path.join(REF, "bin", "normalize_image.py"), pa, pb, po, po2], env=env)
            out[f"{tag}/norm1"] = np.asarray(Image.open(po).convert("L")); out[f"{tag}/norm2"] = np.asarray(Image.open(po2).convert("L"))
    save("lum", **out)
    # metrics: EE / AE with the > 50 px and NaN filters, IE
    out = {}
    for tag, (h, w, amp) in {"a": (19, 27, 1.5), "b": (31, 44, 30.0), "c": (8, 8, 0.0)}.items():
        n = h * w
        u = rng.standard_normal(n) * amp; v = rng.standard_normal(n) * amp
        ug = u + rng.standard_normal(n) * 0.3; vg = v + rng.standard_normal(n) * 0.3
        ug[::11] += 80.0                                  # endpoint errors beyond the 50 px filter
        if tag == "c":
            ug[:] = u; vg[:] = v                          # identical flows: arccos argument rounds above 1 -> NaN filter
        I = rng.random(n); IGT = rng.random(n)
        with np.errstate(invalid="ignore"):
            ee = ref_utils.EE(w, h, u, v, ug, vg); ae = ref_utils.AE(w, h, u, v, ug, vg)
        out.update({f"{tag}/dims": np.array([h, w]), f"{tag}/u": u, f"{tag}/v": v, f"{tag}/ug": ug, f"{tag}/vg": vg,
                    f"{tag}/I": I, f"{tag}/IGT": IGT, f"{tag}/EE": np.array(ee), f"{tag}/AE": np.array(ae),
                    f"{tag}/IE": np.float64(ref_utils.IE(w, h, I, IGT))})
        print(f"  metrics {tag}: EE {ee}, AE {ae}")
    save("metrics", **out)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--full", action="store_true", help="also the 388x584 cases")
    ap.add_argument("--only", default="", help="comma list: operators,stepB,stepA,foto,gn,flow,warp,lum,full,config3ref")
    a = ap.parse_args()
    print(f"numpy {np.__version__}, scipy {scipy.__version__}, reference at {REF}")
    todo = a.only.split(",") if a.only else ["operators", "stepB", "stepA", "flow", "warp", "lum", "gn", "foto"]
    if a.full:
        todo += [t for t in ("full", "config3ref") if t not in todo]
    for name in todo:
        print(f"[{name}]", flush=True)
        globals()["gold_" + name]()
