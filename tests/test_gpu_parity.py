"""GPU: parity of the CUDA path (through the C ABI / ctypes, exactly what the shim modules call)
against (a) golden vectors recorded from the unmodified reference and (b) the CPU oracle on the
same seeded inputs.  Tolerances: bit-exact for index/warp/stencil-order work; the north star's
1e-9 relative per flow component and 1e-6 px endpoint error for the fp64 solvers."""
import os
import sys

import numpy as np
import pytest

from conftest import PKG, load_golden, relerr, epe_max

import foto_b200
from foto_b200 import synth

pytestmark = pytest.mark.gpu

ULP = 2.3e-16
DEGENERATE = {"foto_31x29_nt2", "foto_squares32"}     # see tests/test_oracle_golden.py
DEGENERATE_TOL = 5e-8


@pytest.fixture(params=["onchip_or_auto", "streaming"], autouse=True)
def cg_variant(request):
    """Every test runs twice: auto (single-reduction on-chip CG when Nt <= 8 or 16 and the grid fits one tile per
    SM, else streaming) and the streaming CG (textbook recurrences) forced."""
    foto_b200.set_default_cg_variant(0 if request.param == "streaming" else -1)
    yield request.param
    foto_b200.set_default_cg_variant(-1)


def _frames(g):
    return g["f0_u8"].astype(np.float64).ravel() / 255, g["f1_u8"].astype(np.float64).ravel() / 255


# ----------------------------------------------------------------------------- operators
@pytest.mark.parametrize("op", ["grad_st", "div_st", "laplacian_st", "grad", "div", "grad_forward"])
def test_operator_apply_vs_reference(op):
    g = load_golden("operators")
    vec_in = {"grad_st": "N", "div_st": "3N", "laplacian_st": "N", "grad": "P", "div": "2P", "grad_forward": "P"}[op]
    vec_out = {"grad_st": "3N", "div_st": "N", "laplacian_st": "N", "grad": "2P", "div": "P", "grad_forward": "2P"}[op]
    n = 0
    for k in g.files:
        parts = k.split("/")
        if parts[0] not in (op, op + ".T"):
            continue
        bc, dims, dt, dx, dy = parts[1], parts[2], float(parts[3]), float(parts[4]), float(parts[5])
        Nt, Ny, Nx = map(int, dims.split("x"))
        tr = parts[0].endswith(".T")
        x = g[f"in/{dims}/{parts[3]}/{parts[4]}/{parts[5]}/" + (vec_out if tr else vec_in)]
        y = foto_b200.op_apply(op, bc, Nt, Nx, Ny, dt, dx, dy, x, transpose=tr)
        assert np.max(np.abs(y - g[k])) <= 16 * ULP * max(1.0, np.max(np.abs(g[k]))), k
        n += 1
    assert n >= 6


# ----------------------------------------------------------------------------- stepB
def test_stepB_all_branches_vs_reference():
    g = load_golden("stepB")
    Nt, Nx, Ny = map(int, g["dims"])
    q = foto_b200.stepB(g["p"], Nt, Nx, Ny)
    assert np.all(np.abs(q - g["q"]) <= 1e-12 * np.maximum(1.0, np.abs(g["q"])))


def test_stepB_vs_oracle_random_and_idempotent(oracle):
    rng = np.random.default_rng(7)
    Nt, Ny, Nx = 3, 37, 41
    n = Nt * Ny * Nx
    p = rng.standard_normal(3 * n) * np.repeat([3.0, 2.0, 2.0], n)
    p[:n // 3] -= 5.0
    q = foto_b200.stepB(p, Nt, Nx, Ny)
    qo = oracle.stepB(p, Nt, Nx, Ny)
    assert np.all(np.abs(q - qo) <= 1e-12 * np.maximum(1.0, np.abs(qo)))
    # a projection is idempotent; every output lies in K = {a + |b|^2/2 <= 0}
    q2 = foto_b200.stepB(q, Nt, Nx, Ny)
    assert np.max(np.abs(q2 - q)) < 1e-12
    a, b1, b2 = q[:n], q[n:2 * n], q[2 * n:]
    assert np.max(a + 0.5 * (b1 ** 2 + b2 ** 2)) < 1e-12


# ----------------------------------------------------------------------------- stepA
@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_stepA_vs_reference(tag):
    g = load_golden("stepA")
    Nt, Nx, Ny = map(int, g[f"{tag}/dims"]); r, eps = g[f"{tag}/r_eps"]
    args = (g[f"{tag}/mu"], g[f"{tag}/q"], g[f"{tag}/rho0"], g[f"{tag}/rhoT"])
    F = foto_b200.rhs(*args, r, Nt, Nx, Ny)
    np.testing.assert_array_equal(F, g[f"{tag}/F"])               # K1 is bit-identical to coo_matvec
    phi, iters, info = foto_b200.stepA(*args, r, eps, Nt, Nx, Ny)
    assert info == 0 and iters == int(g[f"{tag}/cg_iters"][0])
    assert relerr(phi, g[f"{tag}/phi"]) < 1e-10
    # size-independent property: the returned phi satisfies the stopping rule ||A phi - F|| < 1e-6 ||F||
    L = foto_b200.op_apply("laplacian_st", "N", Nt, Nx, Ny, 1, 1, 1, phi)
    res = (-r * L + r * eps * phi) - F
    assert np.linalg.norm(res) < 1.0001e-6 * np.linalg.norm(F)


# ----------------------------------------------------------------------------- FOTO end to end
FOTO = ["foto_24x32", "foto_48x64", "foto_37x53_nt5", "foto_40x56_nt16_runsh", "foto_31x29_nt2",
        "foto_squares32", "foto_97x146"]


@pytest.mark.parametrize("name", FOTO)
def test_foto_solve_vs_reference(name):
    g = load_golden(name)
    h, w, Nt = map(int, g["dims"]); r, tol, eps, max_it = g["params"]
    f0, f1 = _frames(g)
    u, v, m, info = foto_b200.solve(f0, f1, Nt, w, h, r=r, convergence_tol=tol, reg_epsilon=eps, max_it=int(max_it))
    assert info["n_outer"] == len(g["crit"])
    np.testing.assert_array_equal(info["cg_iters"], g["cg_iters"])
    np.testing.assert_allclose(info["crit"], g["crit"], rtol=1e-7)
    t = DEGENERATE_TOL if name in DEGENERATE else 1e-9
    assert relerr(u, g["u"]) < t and relerr(v, g["v"]) < t and relerr(m, g["m"]) < t
    assert epe_max(u, v, g["u"], g["v"]) < 1e-6


@pytest.mark.parametrize("name", ["foto_24x32", "foto_48x64", "foto_37x53_nt5"])
def test_foto_tight_backend_vs_tight_reference(name):
    g = load_golden(name + "_tight")
    h, w, Nt = map(int, g["dims"]); r, tol, eps, max_it = g["params"]
    f0, f1 = _frames(g)
    u, v, m, info = foto_b200.solve(f0, f1, Nt, w, h, r=r, convergence_tol=tol, reg_epsilon=eps,
                                    max_it=int(max_it), backend=foto_b200.POISSON_CG_TIGHT)
    assert info["n_outer"] == len(g["crit"])
    assert relerr(u, g["u"]) < 1e-9 and relerr(v, g["v"]) < 1e-9 and relerr(m, g["m"]) < 1e-9


def test_foto_full_size_vs_reference():
    """Config 1: 388x584, CLI defaults.  The golden keeps every 13th pixel plus whole-field sums."""
    g = load_golden("foto_388x584")
    h, w, Nt = map(int, g["dims"])
    f0, f1 = _frames(g)
    u, v, m, info = foto_b200.solve(f0, f1, Nt, w, h, r=1.0, convergence_tol=0.1, reg_epsilon=1e-3, max_it=100)
    assert info["n_outer"] == len(g["crit"]) == 9
    np.testing.assert_array_equal(info["cg_iters"], g["cg_iters"])
    sub = g["sub"]
    for name, full in (("u", u), ("v", v), ("m", m)):
        assert relerr(full[sub], g[name]) < 1e-9, name
        st = np.array([full.sum(), np.abs(full).sum(), np.abs(full).max(), np.sqrt((full ** 2).sum())])
        np.testing.assert_allclose(st, g[name + "_stats"], rtol=1e-9, atol=1e-9 * g[name + "_stats"][1])
    assert epe_max(u[sub], v[sub], g["u"], g["v"]) < 1e-6


def test_foto_vs_oracle_other_shape(oracle):
    h, w, Nt = 61, 83, 6
    f0, f1 = synth.make_pair(h, w, seed=11, shift=(0.3, -0.6))
    kw = dict(r=1.3, convergence_tol=0.05, reg_epsilon=5e-3, max_it=12)
    u, v, m, info = foto_b200.solve(f0, f1, Nt, w, h, **kw)
    uo, vo, mo, io = oracle.solve(f0, f1, Nt, w, h, return_info=True, **kw)
    assert info["n_outer"] == io["n_outer"]
    np.testing.assert_array_equal(info["cg_iters"], io["cg_iters"])
    assert relerr(u, uo) < 1e-9 and relerr(v, vo) < 1e-9 and relerr(m, mo) < 1e-9


# ----------------------------------------------------------------------------- flow / warp
@pytest.mark.parametrize("tag", ["a", "b", "c", "d"])
def test_flow_extraction_bit_exact(tag):
    g = load_golden("flow")
    Nt, Nx, Ny = map(int, g[f"{tag}/dims"])
    u, v, m = foto_b200.flow_from_phi(g[f"{tag}/phi"], Nt, Nx, Ny)
    np.testing.assert_array_equal(u, g[f"{tag}/u"])
    np.testing.assert_array_equal(v, g[f"{tag}/v"])
    np.testing.assert_array_equal(m, g[f"{tag}/m"])


def test_flow_zero_potential_gives_zero_flow():
    u, v, m = foto_b200.flow_from_phi(np.zeros(4 * 388 * 584), 4, 584, 388)
    assert not u.any() and not v.any() and not m.any()


@pytest.mark.parametrize("tag", ["a", "b", "c", "d"])
def test_warp_bit_exact(tag):
    g = load_golden("warp")
    h, w = map(int, g[f"{tag}/dims"])
    out = foto_b200.warp_apply(g[f"{tag}/f1"], g[f"{tag}/u"], g[f"{tag}/v"], w, h, g[f"{tag}/m"])
    np.testing.assert_array_equal(out, g[f"{tag}/out_m"])
    out0 = foto_b200.warp_apply(g[f"{tag}/f1"], g[f"{tag}/u"], g[f"{tag}/v"], w, h, None)
    np.testing.assert_array_equal(out0, g[f"{tag}/out_m0"])


def test_warp_full_size_vs_oracle_and_identity(oracle):
    h, w = 388, 584
    rng = np.random.default_rng(3)
    f = np.round(rng.random(h * w) * 255) / 255
    u, v, m = rng.standard_normal(h * w) * 3, rng.standard_normal(h * w) * 3, rng.standard_normal(h * w) * 0.1
    np.testing.assert_array_equal(foto_b200.warp_apply(f, u, v, w, h, m), oracle.warp_apply(f, u, v, w, h, m))
    np.testing.assert_array_equal(foto_b200.warp_apply(f, np.zeros(h * w), np.zeros(h * w), w, h, None), f)


# ----------------------------------------------------------------------------- GN
@pytest.mark.parametrize("name", ["gn_24x32", "gn_48x64", "gn_37x53", "gn_97x146"])
def test_gn_vs_reference(name):
    g = load_golden(name)
    h, w = map(int, g["dims"]); alpha, lam = g["params"]
    f0, f1 = _frames(g)
    if "x_probe" in g.files:
        y, b = foto_b200.gn_system(f0, f1, w, h, alpha, lam, g["x_probe"])
        assert relerr(y, g["Ax_probe"]) < 1e-14
        np.testing.assert_array_equal(b, g["b"])
    u, v, m, info = foto_b200.gn_solve(f0, f1, w, h, alpha, lam)
    assert info["info"] == 0
    assert relerr(u, g["u"]) < 1e-9 and relerr(v, g["v"]) < 1e-9 and relerr(m, g["m"]) < 1e-9
    assert epe_max(u, v, g["u"], g["v"]) < 1e-6


def test_gn_full_size_vs_reference():
    """Config 2: 388x584, alpha 0.1, lambda 0.2 against the reference's SuperLU solution."""
    g = load_golden("gn_388x584")
    gf = load_golden("foto_388x584")
    h, w = map(int, g["dims"])
    f0, f1 = _frames(gf)
    u, v, m, info = foto_b200.gn_solve(f0, f1, w, h, 0.1, 0.2)
    assert info["info"] == 0
    sub = g["sub"]
    for name, full in (("u", u), ("v", v), ("m", m)):
        assert relerr(full[sub], g[name]) < 1e-9, name
        st = np.array([full.sum(), np.abs(full).sum(), np.abs(full).max(), np.sqrt((full ** 2).sum())])
        np.testing.assert_allclose(st, g[name + "_stats"], rtol=1e-9, atol=1e-9 * g[name + "_stats"][1])
    # size-independent property: the residual of the returned solution is tiny
    x = np.concatenate([u, v, m])
    y, b = foto_b200.gn_system(f0, f1, w, h, 0.1, 0.2, x)
    assert np.linalg.norm(y - b) < 1e-11 * np.linalg.norm(b)


# ----------------------------------------------------------------------------- batch / determinism
def test_batch_equals_single_bitwise():
    h, w, Nt = 48, 64, 4
    pairs = synth.make_batch(5, h, w, base_seed=20)
    f0s = np.stack([p[0] for p in pairs]); f1s = np.stack([p[1] for p in pairs])
    kw = dict(r=1.0, convergence_tol=0.1, reg_epsilon=1e-3, max_it=6)
    us, vs, ms, outer = foto_b200.solve_batch(f0s, f1s, Nt, w, h, devices=list(range(foto_b200.device_count())), **kw)
    for i in range(len(pairs)):
        u, v, m, info = foto_b200.solve(f0s[i], f1s[i], Nt, w, h, **kw)
        np.testing.assert_array_equal(us[i], u); np.testing.assert_array_equal(vs[i], v)
        np.testing.assert_array_equal(ms[i], m); assert outer[i] == info["n_outer"]
    gu, gv, gm, it = foto_b200.gn_solve_batch(f0s, f1s, w, h, 0.1, 0.2, devices=[0])
    u, v, m, info = foto_b200.gn_solve(f0s[2], f1s[2], w, h, 0.1, 0.2)
    np.testing.assert_array_equal(gu[2], u); assert it[2] == info["iters"]


# ----------------------------------------------------------------------------- shim (reference API)
def test_shim_modules_reference_api(capsys, monkeypatch):
    monkeypatch.syspath_prepend(os.path.join(PKG, "shim"))
    for name in ("operators", "utils", "benamou_brenier", "classical"):
        sys.modules.pop(name, None)
    import benamou_brenier as bb, classical, operators, utils  # noqa: E401
    g = load_golden("foto_24x32")
    h, w, Nt = map(int, g["dims"])
    f0, f1 = _frames(g)
    u, v, m = bb.solve(f0, f1, Nt, w, h, r=1.0, convergence_tol=0.1, reg_epsilon=1e-3, max_it=100)
    lines = [l for l in capsys.readouterr().out.splitlines() if l.endswith("/100)")]
    assert len(lines) == len(g["crit"]) and lines[0].endswith("(1/100)")
    assert abs(float(lines[-1].split(" ")[0]) - g["crit"][-1]) < 1e-7 * g["crit"][-1]
    assert relerr(u, g["u"]) < 1e-9 and relerr(m, g["m"]) < 1e-9
    # stepA / stepB / operators through the reference-named functions
    gs = load_golden("stepA")
    Nt2, Nx2, Ny2 = map(int, gs["a/dims"]); r, eps = gs["a/r_eps"]
    from scipy import sparse
    A = -r * operators.laplacian_st(Nt2, Nx2, Ny2, 1, 1, 1, bc='N') + r * eps * sparse.eye(Nt2 * Nx2 * Ny2)
    D = operators.div_st(Nt2, Nx2, Ny2, 1, 1, 1, bc='N')
    phi = bb.solve_benamou_brenier_step(gs["a/mu"], gs["a/q"], gs["a/rho0"], gs["a/rhoT"], r, A, D,
                                        Nt2, Nx2, Ny2, 1, 1, 1)
    assert relerr(phi, gs["a/phi"]) < 1e-10
    np.testing.assert_array_equal(D @ (gs["a/mu"] - r * gs["a/q"]),
                                  foto_b200.op_apply("div_st", "N", Nt2, Nx2, Ny2, 1, 1, 1, gs["a/mu"] - r * gs["a/q"]))
    gb = load_golden("stepB")
    q = bb.stepB(gb["p"], *map(int, gb["dims"]))
    assert np.all(np.abs(q - gb["q"]) <= 1e-12 * np.maximum(1.0, np.abs(gb["q"])))
    # GN class
    gg = load_golden("gn_24x32")
    hh, ww = map(int, gg["dims"]); f0, f1 = _frames(gg)
    s = classical.GLLOpticalFlow(ww, hh); s.setAlpha(0.1); s.setLambda(0.2)
    uu, vv, mm = s.assemble(f0, f1).process()
    assert relerr(uu, gg["u"]) < 1e-9 and relerr(mm, gg["m"]) < 1e-9
    np.testing.assert_array_equal(s.b, gg["b"])
    assert relerr(s.A @ gg["x_probe"], gg["Ax_probe"]) < 1e-14
    # flow + warp through utils
    gfl = load_golden("flow")
    Nt3, Nx3, Ny3 = map(int, gfl["a/dims"])
    u3, v3, m3 = utils.opticalflow_from_benamoubrenier(gfl["a/phi"], Nt3, Nx3, Ny3,
                                                       operators.grad(Nx3, Ny3, 1, 1, bc='N'),
                                                       operators.div(Nx3, Ny3, 1, 1, bc='D'))
    np.testing.assert_array_equal(u3, gfl["a/u"]); np.testing.assert_array_equal(m3, gfl["a/m"])
    gw = load_golden("warp")
    h4, w4 = map(int, gw["a/dims"])
    np.testing.assert_array_equal(utils.apply_opticalflow(gw["a/f1"], gw["a/u"], gw["a/v"], w4, h4, gw["a/m"]), gw["a/out_m"])
    for name in ("operators", "utils", "benamou_brenier", "classical"):
        sys.modules.pop(name, None)


# ----------------------------------------------------------------------------- CLI pipeline
def test_cli_pipeline_flo_output(tmp_path, monkeypatch, oracle, capsys):
    """The sequence main.py runs (main.py:52-53,93,110-111,136-138): PNG -> solve -> warp -> IE -> .flo,
    through the shim modules; the .flo payload must equal the oracle's flow cast to float32."""
    from PIL import Image
    monkeypatch.syspath_prepend(os.path.join(PKG, "shim"))
    for name in ("operators", "utils", "benamou_brenier", "classical"):
        sys.modules.pop(name, None)
    import benamou_brenier as bb, utils  # noqa: E401
    h, w, Nt = 40, 56, 4
    f0, f1 = synth.make_pair(h, w, seed=9)
    for name, f in (("f0.png", f0), ("f1.png", f1)):
        Image.fromarray(np.uint8(np.round(255 * f)).reshape(h, w), "L").save(str(tmp_path / name))
    a, ww, hh = utils.openGrayscaleImage(str(tmp_path / "f0.png"))
    b, _, _ = utils.openGrayscaleImage(str(tmp_path / "f1.png"))
    assert (ww, hh) == (w, h)
    np.testing.assert_array_equal(a, f0)
    u, v, m = bb.solve(a, b, Nt, ww, hh, r=1.0, convergence_tol=0.1, reg_epsilon=1e-3, max_it=100)
    rec = np.clip(utils.apply_opticalflow(a, u, v, ww, hh, m), 0, 1)
    ie = utils.IE(ww, hh, rec, b)
    utils.saveFlo(ww, hh, u, v, str(tmp_path / "out.flo"))
    uo, vo, mo = oracle.solve(f0, f1, Nt, w, h, r=1.0, convergence_tol=0.1, reg_epsilon=1e-3, max_it=100)
    ie_o = utils.IE(w, h, np.clip(oracle.warp_apply(f0, uo, vo, w, h, mo), 0, 1), f1)
    assert abs(ie - ie_o) < 1e-8 * max(1.0, ie_o)
    w2, h2, uf, vf = utils.openFlo(str(tmp_path / "out.flo"))
    assert (w2, h2) == (w, h)
    np.testing.assert_allclose(uf, uo.astype(np.float32), rtol=0, atol=1e-7 * np.abs(uo).max())
    np.testing.assert_allclose(vf, vo.astype(np.float32), rtol=0, atol=1e-7 * np.abs(vo).max())
    for name in ("operators", "utils", "benamou_brenier", "classical"):
        sys.modules.pop(name, None)


# ----------------------------------------------------------------------------- dct_exact back-end (K2b)
@pytest.mark.parametrize("name", ["foto_24x32", "foto_48x64", "foto_37x53_nt5"])
def test_dct_exact_vs_tight_reference(name):
    """Exact DCT Poisson solve against the reference run with its inner CG at rtol 1e-13."""
    g = load_golden(name + "_tight")
    h, w, Nt = map(int, g["dims"]); r, tol, eps, max_it = g["params"]
    f0, f1 = _frames(g)
    u, v, m, info = foto_b200.solve(f0, f1, Nt, w, h, r=r, convergence_tol=tol, reg_epsilon=eps,
                                    max_it=int(max_it), backend=foto_b200.POISSON_DCT_EXACT)
    assert info["n_outer"] == len(g["crit"])
    np.testing.assert_allclose(info["crit"], g["crit"], rtol=1e-7)
    assert relerr(u, g["u"]) < 1e-9 and relerr(v, g["v"]) < 1e-9 and relerr(m, g["m"]) < 1e-9
    assert epe_max(u, v, g["u"], g["v"]) < 1e-6


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_stepA_dct_exact_solves_the_system(tag):
    g = load_golden("stepA")
    Nt, Nx, Ny = map(int, g[f"{tag}/dims"]); r, eps = g[f"{tag}/r_eps"]
    args = (g[f"{tag}/mu"], g[f"{tag}/q"], g[f"{tag}/rho0"], g[f"{tag}/rhoT"])
    phi, iters, info = foto_b200.stepA(*args, r, eps, Nt, Nx, Ny, backend=foto_b200.POISSON_DCT_EXACT)
    assert iters == 0 and info == 0
    F = g[f"{tag}/F"]
    L = foto_b200.op_apply("laplacian_st", "N", Nt, Nx, Ny, 1, 1, 1, phi)
    assert np.linalg.norm((-r * L + r * eps * phi) - F) < 1e-11 * np.linalg.norm(F)
    phi_t, _, _ = foto_b200.stepA(*args, r, eps, Nt, Nx, Ny, backend=foto_b200.POISSON_CG_TIGHT)
    assert relerr(phi, phi_t) < 1e-9
    # the reference's truncated CG differs from the exact solve at the 1e-6 level, by construction
    assert 1e-9 < relerr(phi, g[f"{tag}/phi"]) < 1e-3


@pytest.mark.parametrize("h,w,Nt", [(48, 64, 5), (388, 584, 4), (132, 260, 3)])
def test_dct_folded_transforms_equal_dense(h, w, Nt, monkeypatch):
    """Nx, Ny multiples of 4: the x and y transforms use the even / odd symmetry of the DCT matrix (half the flops, spectrum
    in permuted order), twice along axes whose length is a multiple of 8 (FOTO_DCT_LEVELS).  Same Poisson solution as the dense
    transforms (FOTO_DCT_DENSE=1) to rounding."""
    rng = np.random.default_rng(h * w)
    N = Nt * h * w
    mu, q = rng.standard_normal(3 * N), rng.standard_normal(3 * N)
    rho0, rhoT = rng.random(h * w), rng.random(h * w)
    res = {}
    for dense, levels in (("1", "1"), (None, "1"), (None, "2")):      # dense, one folding level, two (axes of length 8k)
        if dense: monkeypatch.setenv("FOTO_DCT_DENSE", dense)
        else: monkeypatch.delenv("FOTO_DCT_DENSE", raising=False)
        monkeypatch.setenv("FOTO_DCT_LEVELS", levels)
        res[(dense, levels)] = foto_b200.stepA(mu, q, rho0, rhoT, 1.0, 1e-3, Nt, w, h, backend=foto_b200.POISSON_DCT_EXACT)[0]
    assert relerr(res[(None, "1")], res[("1", "1")]) < 1e-13 and relerr(res[(None, "2")], res[("1", "1")]) < 1e-13
    phi = res[(None, "2")]
    L = foto_b200.op_apply("laplacian_st", "N", Nt, w, h, 1, 1, 1, phi)
    F = foto_b200.rhs(mu, q, rho0, rhoT, 1.0, Nt, w, h)
    assert np.linalg.norm((-L + 1e-3 * phi) - F) < 1e-11 * np.linalg.norm(F)


def test_dct_exact_full_size_matches_tight_cg():
    g = load_golden("foto_388x584")
    h, w, Nt = map(int, g["dims"])
    f0, f1 = _frames(g)
    kw = dict(r=1.0, convergence_tol=0.1, reg_epsilon=1e-3, max_it=100)
    u, v, m, info = foto_b200.solve(f0, f1, Nt, w, h, backend=foto_b200.POISSON_DCT_EXACT, **kw)
    ut, vt, mt, it = foto_b200.solve(f0, f1, Nt, w, h, backend=foto_b200.POISSON_CG_TIGHT, **kw)
    assert info["n_outer"] == it["n_outer"]
    assert relerr(u, ut) < 1e-9 and relerr(v, vt) < 1e-9 and relerr(m, mt) < 1e-9
    # and stays within the 1e-6 px endpoint-error contract of the as-shipped reference
    sub = g["sub"]
    assert epe_max(u[sub], v[sub], g["u"], g["v"]) < 1e-6


# ----------------------------------------------------------------------------- next-tier rows
def test_flo_egress_byte_identical(tmp_path, monkeypatch):
    monkeypatch.syspath_prepend(os.path.join(PKG, "shim"))
    sys.modules.pop("utils", None)
    import utils
    rng = np.random.default_rng(5)
    w, h = 37, 23
    u, v = rng.standard_normal(w * h) * 3, rng.standard_normal(w * h) * 3
    utils.saveFlo(w, h, u, v, str(tmp_path / "ref.flo"))          # numpy astype(float32), as the reference does
    foto_b200.save_flo(w, h, u, v, str(tmp_path / "gpu.flo"))
    assert open(tmp_path / "ref.flo", "rb").read() == open(tmp_path / "gpu.flo", "rb").read()
    sys.modules.pop("utils", None)


def test_flow_metrics_match_numpy(monkeypatch):
    monkeypatch.syspath_prepend(os.path.join(PKG, "shim"))
    sys.modules.pop("utils", None)
    import utils
    rng = np.random.default_rng(6)
    n = 388 * 584
    u, v = rng.standard_normal(n), rng.standard_normal(n)
    ug, vg = u + 0.1 * rng.standard_normal(n), v + 0.1 * rng.standard_normal(n)
    ug[::1000] += 100.0                                           # EE > 50: filtered out
    aee, sdee, aae, sdae = foto_b200.flow_metrics(u, v, ug, vg)
    r_aee, r_sdee = utils.EE(584, 388, u, v, ug, vg)
    r_aae, r_sdae = utils.AE(584, 388, u, v, ug, vg)
    assert abs(aee - r_aee) < 1e-12 and abs(sdee - r_sdee) < 1e-9
    assert abs(aae - r_aae) < 1e-12 and abs(sdae - r_sdae) < 1e-9
    sys.modules.pop("utils", None)


def test_stepA_large_grid_streaming_property():
    """A grid too large for the on-chip CG variant (2 M cells): the streaming kernel must return a phi
    that satisfies scipy's stopping rule, whatever variant the default context prefers."""
    rng = np.random.default_rng(8)
    Nt, Ny, Nx = 4, 540, 960
    N = Nt * Ny * Nx
    mu = rng.standard_normal(3 * N); q = rng.standard_normal(3 * N)
    rho0 = rng.random(Ny * Nx); rhoT = rng.random(Ny * Nx)
    F = foto_b200.rhs(mu, q, rho0, rhoT, 1.0, Nt, Nx, Ny)
    phi, iters, info = foto_b200.stepA(mu, q, rho0, rhoT, 1.0, 1e-3, Nt, Nx, Ny)
    assert info == 0 and 50 < iters < 1000
    L = foto_b200.op_apply("laplacian_st", "N", Nt, Nx, Ny, 1, 1, 1, phi)
    assert np.linalg.norm((-L + 1e-3 * phi) - F) < 1.0001e-6 * np.linalg.norm(F)
    phi_d, _, _ = foto_b200.stepA(mu, q, rho0, rhoT, 1.0, 1e-3, Nt, Nx, Ny, backend=foto_b200.POISSON_DCT_EXACT)
    assert np.linalg.norm((-foto_b200.op_apply("laplacian_st", "N", Nt, Nx, Ny, 1, 1, 1, phi_d) + 1e-3 * phi_d) - F) < 1e-11 * np.linalg.norm(F)


def test_identical_frames_zero_rhs(oracle):
    """rho0 == rhoT: the right-hand side is pure rounding noise of the time interpolation (1e-17), CG
    converges on it, the flow is exactly zero and the loop stops after one outer iteration."""
    h, w, Nt = 30, 44, 4
    f0, _ = synth.make_pair(h, w, seed=3)
    u, v, m, info = foto_b200.solve(f0, f0, Nt, w, h, r=1.0, convergence_tol=0.1, reg_epsilon=1e-3, max_it=20)
    uo, vo, mo, io = oracle.solve(f0, f0, Nt, w, h, r=1.0, convergence_tol=0.1, reg_epsilon=1e-3, max_it=20, return_info=True)
    assert info["n_outer"] == io["n_outer"] == 1
    assert list(info["cg_iters"]) == list(io["cg_iters"]) and list(info["cg_info"]) == [0]
    assert not u.any() and not v.any() and not m.any() and not uo.any()
    # an exactly zero right-hand side: scipy returns b (= 0) with info 0 and no iteration
    z = np.zeros(3 * Nt * h * w)
    phi, iters, cinfo = foto_b200.stepA(z, z, np.zeros(h * w), np.zeros(h * w), 1.0, 1e-3, Nt, w, h)
    assert iters == 0 and cinfo == 0 and not phi.any()


def test_cg_maxiter_warning_path(capsys, monkeypatch):
    """A tiny eps makes the 1000-iteration budget bind: info = maxiter, the shim prints the
    reference's WARNING (benamou_brenier.py:86-87) and carries on."""
    monkeypatch.syspath_prepend(os.path.join(PKG, "shim"))
    for name in ("operators", "utils", "benamou_brenier", "classical"):
        sys.modules.pop(name, None)
    import benamou_brenier as bb
    h, w, Nt = 128, 192, 4
    f0, f1 = synth.make_pair(h, w, seed=4)
    u, v, m, info = foto_b200.solve(f0, f1, Nt, w, h, r=1.0, convergence_tol=0.1, reg_epsilon=1e-12, max_it=2)
    assert list(info["cg_info"]) == [1000, 1000] and list(info["cg_iters"]) == [1000, 1000]
    bb.solve(f0, f1, Nt, w, h, r=1.0, convergence_tol=0.1, reg_epsilon=1e-12, max_it=1)
    out = capsys.readouterr().out
    assert "WARNING: CG did not converge in 1000 iterations." in out
    for name in ("operators", "utils", "benamou_brenier", "classical"):
        sys.modules.pop(name, None)


# ----------------------------------------------------------------------------- odd shapes (tile decomposition)
_SHAPES = [(2, 2, 2), (2, 3, 5), (3, 2, 7), (5, 4, 3), (4, 13, 2), (3, 31, 17), (7, 19, 23), (4, 64, 96),
           (16, 23, 41), (9, 150, 11), (2, 7, 300), (6, 97, 101), (33, 12, 14), (4, 200, 240), (70, 9, 8), (8, 45, 52)]


@pytest.mark.parametrize("Nt,Ny,Nx", _SHAPES)
def test_stepA_odd_shapes_vs_oracle(oracle, Nt, Ny, Nx):
    """Grids from 8 cells to 190 k cells, thin in every direction: exercises 1-wide tiles, tiles with
    no neighbours, Nt > 64 (streaming only) and every boundary case of the on-chip decomposition."""
    rng = np.random.default_rng(Nt * 10007 + Ny * 101 + Nx)
    N = Nt * Ny * Nx
    mu = rng.standard_normal(3 * N); q = rng.standard_normal(3 * N)
    rho0 = rng.random(Ny * Nx); rhoT = rng.random(Ny * Nx)
    r, eps = 1.0 if (Nt + Ny) % 2 else 0.8, 1e-2
    phi, iters, info = foto_b200.stepA(mu, q, rho0, rhoT, r, eps, Nt, Nx, Ny)
    phio, iterso, infoo = oracle.stepA(mu, q, rho0, rhoT, r, eps, Nt, Nx, Ny)
    assert info == infoo == 0
    assert abs(iters - iterso) <= 1          # random right-hand sides: the count may tip at the threshold
    if iters == iterso:
        assert relerr(phi, phio) < 1e-9
    L = foto_b200.op_apply("laplacian_st", "N", Nt, Nx, Ny, 1, 1, 1, phi)
    assert np.linalg.norm((-r * L + r * eps * phi) - foto_b200.rhs(mu, q, rho0, rhoT, r, Nt, Nx, Ny)) < \
        1.0001e-6 * np.linalg.norm(oracle.rhs(mu, q, rho0, rhoT, r, Nt, Nx, Ny))


@pytest.mark.parametrize("Nt,Ny,Nx", [(3, 5, 7), (4, 33, 2), (2, 2, 64), (5, 40, 57)])
def test_solve_and_gn_odd_shapes_vs_oracle(oracle, Nt, Ny, Nx):
    rng = np.random.default_rng(Ny * 7 + Nx)
    f0 = np.round(rng.random(Ny * Nx) * 255) / 255; f1 = np.round(np.clip(f0 + 0.05 * rng.standard_normal(Ny * Nx), 0, 1) * 255) / 255
    kw = dict(r=1.0, convergence_tol=0.05, reg_epsilon=1e-2, max_it=4)
    u, v, m, info = foto_b200.solve(f0, f1, Nt, Nx, Ny, **kw)
    uo, vo, mo, io = oracle.solve(f0, f1, Nt, Nx, Ny, return_info=True, **kw)
    assert info["n_outer"] == io["n_outer"]
    np.testing.assert_array_equal(info["cg_iters"], io["cg_iters"])
    scale = max(np.abs(uo).max(), np.abs(vo).max(), 1e-300)
    # white-noise frames on tiny grids are ill-conditioned inputs: flipping the last bit of one frame moves
    # the oracle's own answer by 2e-8 relative (measured), so this indexing test uses 1e-6, not 1e-9
    assert np.abs(u - uo).max() < 1e-6 * scale and np.abs(v - vo).max() < 1e-6 * scale
    gu, gv, gm, gi = foto_b200.gn_solve(f0, f1, Nx, Ny, 0.1, 0.2)
    ou, ov, om = oracle.gn_solve(f0, f1, Nx, Ny, 0.1, 0.2)
    assert relerr(gu, ou) < 1e-9 and relerr(gv, ov) < 1e-9 and relerr(gm, om) < 1e-9


def test_deterministic_bitwise_repeat():
    """Fixed-order reductions everywhere (no floating-point atomics): two runs give identical bits."""
    h, w, Nt = 97, 146, 4
    f0, f1 = synth.make_pair(h, w, seed=0)
    kw = dict(r=1.0, convergence_tol=0.1, reg_epsilon=1e-3, max_it=3)
    a = foto_b200.solve(f0, f1, Nt, w, h, **kw)
    b = foto_b200.solve(f0, f1, Nt, w, h, **kw)
    for x, y in zip(a[:3], b[:3]):
        np.testing.assert_array_equal(x, y)
    ga = foto_b200.gn_solve(f0, f1, w, h, 0.1, 0.2); gb = foto_b200.gn_solve(f0, f1, w, h, 0.1, 0.2)
    np.testing.assert_array_equal(ga[0], gb[0])


def test_run_sh_parameters_nt16_on_chip(oracle):
    """The author's production parameters (run.sh:114): Nt=16, eps=1e-2, tol=0.01, on a half-resolution
    frame; 16 time planes per tile in the on-chip CG kernel."""
    h, w, Nt = 97, 146, 16
    f0, f1 = synth.make_pair(h, w, seed=0)
    kw = dict(r=1.0, convergence_tol=0.01, reg_epsilon=1e-2, max_it=6)
    u, v, m, info = foto_b200.solve(f0, f1, Nt, w, h, **kw)
    uo, vo, mo, io = oracle.solve(f0, f1, Nt, w, h, return_info=True, **kw)
    assert info["n_outer"] == io["n_outer"] == 6
    np.testing.assert_array_equal(info["cg_iters"], io["cg_iters"])
    assert relerr(u, uo) < 1e-9 and relerr(v, vo) < 1e-9 and relerr(m, mo) < 1e-9


# ----------------------------------------------------------------------------- time-slab mode (config 5)
def test_slab_single_rank_equals_whole_volume():
    """world = 1: the slab driver (halo-aware K1/K3, split DCT pieces, torch stream) must reproduce the
    whole-volume dct_exact solve bit for bit."""
    import torch
    from foto_b200 import slab
    h, w, Nt = 48, 64, 5
    f0, f1 = synth.make_pair(h, w, seed=2)
    dev = torch.device("cuda", 0)
    kw = dict(r=1.0, convergence_tol=0.1, reg_epsilon=1e-3, max_it=6)
    s = slab.SlabSolver(Nt, w, h, device=dev)
    u, v, m, info = s.solve(torch.from_numpy(f0).to(dev), torch.from_numpy(f1).to(dev), **kw)
    uo, vo, mo, io = foto_b200.solve(f0, f1, Nt, w, h, backend=foto_b200.POISSON_DCT_EXACT, **kw)
    assert info["n_outer"] == io["n_outer"]
    np.testing.assert_array_equal(u.cpu().numpy(), uo)
    np.testing.assert_array_equal(m.cpu().numpy(), mo)


@pytest.mark.skipif(foto_b200.device_count() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("ranks,h,w,Nt", [(2, 48, 64, 5), (2, 61, 83, 4)])
def test_slab_two_ranks_bit_identical(ranks, h, w, Nt):
    import subprocess
    from conftest import ROOT
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={ranks}", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tools", "run_slab.py"), str(h), str(w), str(Nt), "6", "--check"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-3000:]
    import json
    res = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert res["bit_identical"] and res["outer"] == res["single_gpu_outer"]


@pytest.mark.parametrize("ranks,h,w,Nt", [(2, 48, 64, 5), (3, 61, 83, 7)])
def test_slab_ranks_share_one_gpu_bit_identical(ranks, h, w, Nt):
    """The same decomposition with every rank on GPU 0 (gloo group, exchanges staged through the host): runs on a
    single-GPU box, same kernels, same slab geometry, uneven splits with 3 ranks."""
    import json
    import subprocess
    from conftest import ROOT
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={ranks}", "--master-addr", "127.0.0.1",
           "--master-port", "29534", os.path.join(ROOT, "tools", "run_slab.py"), str(h), str(w), str(Nt), "6", "--check", "--one-gpu"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-3000:]
    res = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert res["bit_identical"] and res["outer"] == res["single_gpu_outer"] and res["ranks"] == ranks


def test_allreduce_placement_does_not_change_results(monkeypatch):
    """The on-chip CG kernel keeps its all-reduce words in 2 KB granules it has classified by die (cg_fused.cu,
    place_allreduce).  Where the words live must never change a bit of the answer: forced placements (FOTO_AR_PLACE), the
    single-copy broadcast and the automatic choice give identical flows and CG counts."""
    h, w, Nt = 97, 146, 4
    f0, f1 = synth.make_pair(h, w, seed=5)
    kw = dict(r=1.0, convergence_tol=0.1, reg_epsilon=1e-3, max_it=4)
    ref = foto_b200.solve(f0, f1, Nt, w, h, **kw)
    for env in ({"FOTO_AR_PLACE": "0,1,2"}, {"FOTO_AR_PLACE": "29,30,31"}, {"FOTO_AR_PLACE": "5,17,9", "FOTO_AR_ONECOPY": "1"}, {"FOTO_AR_ONECOPY": "1"}):
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        out = foto_b200.solve(f0, f1, Nt, w, h, **kw)
        for k in env:
            monkeypatch.delenv(k)
        np.testing.assert_array_equal(out[3]["cg_iters"], ref[3]["cg_iters"])
        for a, b in zip(out[:3], ref[:3]):
            np.testing.assert_array_equal(a, b)


@pytest.mark.parametrize("ranks,h,w,Nt", [(1, 48, 64, 5), (2, 48, 64, 5), (3, 61, 83, 7)])
def test_slab_cg_parity_matches_single_gpu(ranks, h, w, Nt):
    """The reference's truncated CG as the slab Poisson back-end (foto_slab_cg_dev: stepwise kernels, one boundary plane of r
    and two all-reduces per CG iteration) against the one-GPU streaming kernel: same outer count, CG counts equal (or +-1 in
    the rare flipped solve) and the flow to 1e-9 (5e-6 with a flip).  Ranks share GPU 0 over gloo, so it runs anywhere."""
    import json
    import subprocess
    from conftest import ROOT
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={ranks}", "--master-addr", "127.0.0.1",
           "--master-port", "29535", os.path.join(ROOT, "tools", "run_slab.py"), str(h), str(w), str(Nt), "4", "--check", "--one-gpu", "--cg"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-3000:]
    res = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert res["outer"] == res["single_gpu_outer"] and res["ranks"] == ranks
    a, b = np.array(res["cg_iters"]), np.array(res["single_gpu_cg_iters"])
    assert np.all(np.abs(a - b) <= 1), (a, b)
    assert res["max_rel_diff"] < (1e-9 if np.array_equal(a, b) else 5e-6), res


@pytest.mark.parametrize("h,w", [(48, 64), (388, 584), (130, 210)])
def test_gn_folded_preconditioner_equals_dense(h, w, monkeypatch):
    """The spectral GN preconditioner with even / odd folded transforms (automatic from 1 M pixels on, forced here with
    FOTO_GN_FOLD) against the dense transforms: same solution to 1e-11 and the same PCG iteration count (+-2); 130x210
    exercises half sizes that need padding to multiples of 4."""
    f0, f1 = synth.make_pair(h, w, seed=3 * h + w)
    res = {}
    for fold in ("0", "1"):
        monkeypatch.setenv("FOTO_GN_FOLD", fold)
        res[fold] = foto_b200.gn_solve(f0, f1, w, h, 0.1, 0.2)
    monkeypatch.delenv("FOTO_GN_FOLD")
    assert abs(res["0"][3]["iters"] - res["1"][3]["iters"]) <= 2 and res["1"][3]["info"] == 0
    for a, b in zip(res["0"][:3], res["1"][:3]):
        assert relerr(a, b) < 1e-11


def test_gn_large_image_streaming_property(cg_variant):
    """720x1280 does not fit the on-chip GN kernel: the spectral solver (auto) and the streaming Jacobi-PCG (forced)
    must return a solution of A x = b (residual checked with the library's own K5/K6 operator, which the goldens
    pin); forcing the on-chip kernel must fail loudly, not fall back."""
    h, w = 720, 1280
    f0, f1 = synth.make_pair(h, w, seed=12)
    u, v, m, info = foto_b200.gn_solve(f0, f1, w, h, 0.1, 0.2, rtol=1e-10)
    assert info["info"] == 0 and (info["iters"] > 100 if cg_variant == "streaming" else 10 < info["iters"] < 300)
    y, b = foto_b200.gn_system(f0, f1, w, h, 0.1, 0.2, np.concatenate([u, v, m]))
    assert np.linalg.norm(y - b) < 2e-10 * np.linalg.norm(b)
    ctx = foto_b200.Context(0)
    ctx.set_cg_variant(2)
    out = [np.empty(h * w) for _ in range(3)]
    with pytest.raises(ValueError, match="does not fit"):
        ctx.gn_solve_host(f0, f1, w, h, 0.1, 0.2, *out)
    ctx.close()


def test_gn_onchip_matches_streaming_many_shapes():
    """The two Jacobi-PCG kernels run the same recurrence with different summation orders: same iteration count
    (+-1) and the same solution to 1e-11 over tile shapes that exercise uneven splits and 1-pixel-wide tiles."""
    ctx = foto_b200.Context(0)
    for (h, w) in [(2, 2), (3, 5), (17, 149), (149, 17), (64, 64), (150, 600), (388, 584), (431, 571)]:
        f0, f1 = synth.make_pair(h, w, seed=h * 1000 + w)
        res = {}
        for var in (0, 2):
            ctx.set_cg_variant(var)
            out = [np.empty(h * w) for _ in range(3)]
            it = ctx.gn_solve_host(f0, f1, w, h, 0.1, 0.2, *out)
            res[var] = (it, np.concatenate(out))
        assert abs(res[0][0] - res[2][0]) <= 1, (h, w, res[0][0], res[2][0])
        scale = np.abs(res[0][1]).max() + 1e-300
        assert np.abs(res[0][1] - res[2][1]).max() <= 1e-11 * scale, (h, w)
    ctx.close()


def test_gn_spectral_preconditioner_many_shapes():
    """The spectral solver (fp64 CG, TF32 tensor-core DCT preconditioner; auto from 64 x 64 pixels on) against the
    streaming Jacobi-PCG: same solution to 1e-10 relative (both converge to rtol 1e-13), an order of magnitude fewer
    iterations, on shapes with odd sizes and sizes that are not multiples of the GEMM tiles."""
    ctx = foto_b200.Context(0)
    for (h, w) in [(2, 2), (3, 5), (17, 149), (149, 17), (64, 64), (97, 146), (150, 600), (388, 584), (431, 571), (480, 640)]:
        f0, f1 = synth.make_pair(h, w, seed=h * 1000 + w)
        if (h, w) == (97, 146):
            f1 = synth.perturb_brightness(f1, h, w, seed=5)
        res = {}
        for var in (0, 3):
            ctx.set_cg_variant(var)
            out = [np.empty(h * w) for _ in range(3)]
            it = ctx.gn_solve_host(f0, f1, w, h, 0.1, 0.2, *out)
            res[var] = (it, np.concatenate(out))
        scale = np.abs(res[0][1]).max() + 1e-300
        assert np.abs(res[0][1] - res[3][1]).max() <= 1e-10 * scale, (h, w, res[0][0], res[3][0])
        if h * w >= 64 * 64:
            assert res[3][0] * 5 < res[0][0], (h, w, res[0][0], res[3][0])
    ctx.set_cg_variant(-1)
    out = [np.empty(388 * 584) for _ in range(3)]
    f0, f1 = synth.make_pair(388, 584, seed=0)
    assert ctx.gn_solve_host(f0, f1, 584, 388, 0.1, 0.2, *out) < 200           # auto takes the spectral solver here
    ctx.close()


def test_cg_kernel_selection(cg_variant):
    """Which Poisson kernel ran (stats.cg_variant): auto takes the single-reduction on-chip kernel for the
    truncated cg_parity solve at Nt = 2..8 or 16 when the grid fits, the streaming kernel for other Nt, for cg_tight
    and when the grid does not fit; forcing a kernel that does not fit (or the removed variant 1) fails loudly."""
    import torch
    ctx = foto_b200.Context(0)
    expect = {"onchip_or_auto": (3, 0, 0), "streaming": (0, 0, 0)}[cg_variant]
    ctx.set_cg_variant(0 if cg_variant == "streaming" else -1)
    def run(h, w, Nt, backend=foto_b200.POISSON_CG_PARITY):
        f0, f1 = synth.make_pair(h, w, seed=2)
        d0 = torch.from_numpy(f0).cuda(); d1 = torch.from_numpy(f1).cuda()
        o = [torch.empty(h * w, dtype=torch.float64, device="cuda") for _ in range(3)]
        ctx.solve_dev(d0.data_ptr(), d1.data_ptr(), Nt, w, h, *[t.data_ptr() for t in o], max_it=1, backend=backend)
        return ctx.stats()["cg_variant"]
    assert run(97, 146, 4) == expect[0]
    assert run(97, 146, 5) == expect[0]                             # patch shapes exist for Nt = 2..8 and 16
    assert run(97, 146, 8) == expect[0]
    assert run(40, 56, 16) == expect[0]
    assert run(97, 146, 9) == expect[1]
    assert run(97, 146, 4, foto_b200.POISSON_CG_TIGHT) == expect[2]
    assert run(540, 960, 4) == 0                                   # 2 M cells: nothing on-chip fits
    ctx.set_cg_variant(2)
    with pytest.raises(ValueError, match="does not fit"):
        run(97, 146, 9)
    assert run(97, 146, 4) == 3
    with pytest.raises(ValueError, match="cg variant"):
        ctx.set_cg_variant(1)
    ctx.close()


def test_large_single_reduction_variant_480x640():
    """480x640x4 (half of the Middlebury sequences) does not fit the 512-thread single-reduction kernel; auto takes its
    large variant (384 threads x 24 cell slots, x in global memory).  Same CG iteration count and phi as the streaming kernel."""
    import torch
    h, w, Nt = 480, 640, 4
    f0, f1 = synth.make_pair(h, w, seed=5)
    d0 = torch.from_numpy(f0).cuda(); d1 = torch.from_numpy(f1).cuda()
    res = {}
    for var in (-1, 0):
        ctx = foto_b200.Context(0); ctx.set_cg_variant(var)
        o = [torch.empty(h * w, dtype=torch.float64, device="cuda") for _ in range(3)]
        info = ctx.solve_dev(d0.data_ptr(), d1.data_ptr(), Nt, w, h, *[t.data_ptr() for t in o], max_it=2, convergence_tol=0.0)
        res[var] = (ctx.stats()["cg_variant"], list(info["cg_iters"]), torch.stack(o).cpu().numpy())
        ctx.close()
    assert res[-1][0] == 3
    assert res[0][0] == 0
    assert res[-1][1] == res[0][1]
    assert relerr(res[-1][2], res[0][2]) < 1e-10


def test_stepA_full_size_true_residual(cg_variant):
    """388x584x4 (config 1): whichever Poisson kernel runs, the returned phi satisfies scipy's stopping rule with its TRUE
    residual (the single-reduction kernel carries A p by recurrence, so its recursive residual must not have drifted),
    and the two kernels agree on the iteration count and on phi to 1e-9."""
    rng = np.random.default_rng(21)
    Nt, Ny, Nx = 4, 388, 584
    N = Nt * Ny * Nx
    f0, f1 = synth.make_pair(Ny, Nx, seed=0)
    mu = np.zeros(3 * N); q = 0.01 * rng.standard_normal(3 * N)
    for n in range(Nt):
        mu[n * Ny * Nx:(n + 1) * Ny * Nx] = (1 - n / (Nt - 1)) * f0 + n / (Nt - 1) * f1
    F = foto_b200.rhs(mu, q, f0, f1, 1.0, Nt, Nx, Ny)
    phi, iters, info = foto_b200.stepA(mu, q, f0, f1, 1.0, 1e-3, Nt, Nx, Ny)
    assert info == 0 and 100 < iters < 1000
    L = foto_b200.op_apply("laplacian_st", "N", Nt, Nx, Ny, 1, 1, 1, phi)
    assert np.linalg.norm((-L + 1e-3 * phi) - F) < 1.0001e-6 * np.linalg.norm(F)
    foto_b200.set_default_cg_variant(0)
    phi_s, iters_s, _ = foto_b200.stepA(mu, q, f0, f1, 1.0, 1e-3, Nt, Nx, Ny)
    assert iters == iters_s and relerr(phi, phi_s) < 1e-9
