"""CPU: pin the oracle (oracle/foto_oracle.c) against golden vectors that were produced by
running the unmodified reference (tests/golden/make_golden.py).  No GPU, no /root/reference."""
import numpy as np
import pytest

from conftest import load_golden, relerr, epe_max

ULP = 2.3e-16


def test_1d_builders_match_reference_dense(oracle):
    g = load_golden("operators")
    keys = [k for k in g.files if k.startswith("dense/")]
    assert len(keys) == 7 * 2 * 3
    for k in keys:
        _, kind, bc, n, h = k.split("/")
        np.testing.assert_array_equal(oracle.tri_dense(kind, int(n), float(h), bc), g[k], err_msg=k)


@pytest.mark.parametrize("op,vec_in", [("grad_st", "N"), ("div_st", "3N"), ("laplacian_st", "N"),
                                       ("grad", "P"), ("div", "2P"), ("grad_forward", "P")])
def test_operator_apply_matches_reference(oracle, op, vec_in):
    g = load_golden("operators")
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
        y = oracle.op_apply(op, bc, Nt, Nx, Ny, dt, dx, dy, x, transpose=tr)
        assert y.shape == g[k].shape
        assert np.max(np.abs(y - g[k])) <= 16 * ULP * max(1.0, np.max(np.abs(g[k]))), k
        n += 1
    assert n >= 6


def test_stepB_all_branches(oracle):
    g = load_golden("stepB")
    Nt, Nx, Ny = g["dims"]
    q = oracle.stepB(g["p"], int(Nt), int(Nx), int(Ny))
    # closed forms with cancellation: the reference's own scalar vs vector evaluation differs
    # by ~3e-16; 1e-12 relative per cell is the oracle gate
    assert np.all(np.abs(q - g["q"]) <= 1e-12 * np.maximum(1.0, np.abs(g["q"])))


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_stepA_rhs_and_cg(oracle, tag):
    g = load_golden("stepA")
    Nt, Nx, Ny = map(int, g[f"{tag}/dims"]); r, eps = g[f"{tag}/r_eps"]
    args = (g[f"{tag}/mu"], g[f"{tag}/q"], g[f"{tag}/rho0"], g[f"{tag}/rhoT"])
    F = oracle.rhs(*args, r, Nt, Nx, Ny)
    np.testing.assert_array_equal(F, g[f"{tag}/F"])     # same IEEE operations in coo_matvec order
    phi, iters, info = oracle.stepA(*args, r, eps, Nt, Nx, Ny)
    assert info == 0
    assert iters == int(g[f"{tag}/cg_iters"][0])
    assert relerr(phi, g[f"{tag}/phi"]) < 1e-10


FOTO = ["foto_24x32", "foto_48x64", "foto_37x53_nt5", "foto_40x56_nt16_runsh", "foto_31x29_nt2",
        "foto_squares32", "foto_97x146"]


# Inputs with an exact symmetry (Nt = 2: the first right-hand side is exactly t-symmetric; the
# two-squares fixture: exact zeros/ones) leave the reference's truncated CG (rtol 1e-6) at the
# mercy of rounding noise: which t-antisymmetric modes get seeded depends on the summation order
# inside scipy's csr_matvec and BLAS ddot.  The oracle mimics the csr_matvec order (that alone
# moved the Nt = 2 first-iteration count from 166 to the reference's 180) but cannot mimic
# BLAS, so these two cases are held to 5e-8 relative instead of 1e-9 (still < 1e-6 px EPE).
DEGENERATE = {"foto_31x29_nt2", "foto_squares32"}
DEGENERATE_TOL = 5e-8


@pytest.mark.parametrize("name", FOTO)
def test_foto_solve_end_to_end(oracle, name):
    g = load_golden(name)
    h, w, Nt = map(int, g["dims"]); r, tol, eps, max_it = g["params"]
    f0 = g["f0_u8"].astype(np.float64).ravel() / 255; f1 = g["f1_u8"].astype(np.float64).ravel() / 255
    u, v, m, info = oracle.solve(f0, f1, Nt, w, h, r=r, convergence_tol=tol, reg_epsilon=eps,
                                 max_it=int(max_it), return_info=True)
    assert info["n_outer"] == len(g["crit"])
    np.testing.assert_array_equal(info["cg_iters"], g["cg_iters"])
    np.testing.assert_allclose(info["crit"], g["crit"], rtol=1e-7)
    tol = DEGENERATE_TOL if name in DEGENERATE else 1e-9
    assert relerr(u, g["u"]) < tol and relerr(v, g["v"]) < tol and relerr(m, g["m"]) < tol
    assert epe_max(u, v, g["u"], g["v"]) < 1e-6


@pytest.mark.parametrize("name", ["foto_24x32", "foto_48x64", "foto_37x53_nt5"])
def test_foto_tight_oracle(oracle, name):
    """Reference with its inner CG run to rtol 1e-13 (the exact-solve limit) -- gate for dct_exact."""
    g = load_golden(name + "_tight")
    h, w, Nt = map(int, g["dims"]); r, tol, eps, max_it = g["params"]
    f0 = g["f0_u8"].astype(np.float64).ravel() / 255; f1 = g["f1_u8"].astype(np.float64).ravel() / 255
    u, v, m, info = oracle.solve(f0, f1, Nt, w, h, r=r, convergence_tol=tol, reg_epsilon=eps,
                                 max_it=int(max_it), cg_rtol=1e-13, cg_maxiter=100000, return_info=True)
    assert info["n_outer"] == len(g["crit"])
    assert relerr(u, g["u"]) < 1e-9 and relerr(v, g["v"]) < 1e-9 and relerr(m, g["m"]) < 1e-9


@pytest.mark.parametrize("tag", ["a", "b", "c", "d"])
def test_flow_extraction(oracle, tag):
    g = load_golden("flow")
    Nt, Nx, Ny = map(int, g[f"{tag}/dims"])
    u, v, m = oracle.flow_from_phi(g[f"{tag}/phi"], Nt, Nx, Ny)
    # same IEEE operations in the same order as the reference's scalar loop => bit-exact u, v
    np.testing.assert_array_equal(u, g[f"{tag}/u"])
    np.testing.assert_array_equal(v, g[f"{tag}/v"])
    np.testing.assert_array_equal(m, g[f"{tag}/m"])


@pytest.mark.parametrize("tag", ["a", "b", "c", "d"])
def test_warp_bit_exact(oracle, tag):
    g = load_golden("warp")
    h, w = map(int, g[f"{tag}/dims"])
    out = oracle.warp_apply(g[f"{tag}/f1"], g[f"{tag}/u"], g[f"{tag}/v"], w, h, g[f"{tag}/m"])
    np.testing.assert_array_equal(out, g[f"{tag}/out_m"])
    out0 = oracle.warp_apply(g[f"{tag}/f1"], g[f"{tag}/u"], g[f"{tag}/v"], w, h, None)
    np.testing.assert_array_equal(out0, g[f"{tag}/out_m0"])


@pytest.mark.parametrize("name", ["gn_24x32", "gn_48x64", "gn_37x53", "gn_97x146"])
def test_gn_solve(oracle, name):
    g = load_golden(name)
    h, w = map(int, g["dims"]); alpha, lam = g["params"]
    f0 = g["f0_u8"].astype(np.float64).ravel() / 255; f1 = g["f1_u8"].astype(np.float64).ravel() / 255
    if "x_probe" in g.files:
        y, b = oracle.gn_system(f0, f1, w, h, alpha, lam, g["x_probe"])
        assert relerr(y, g["Ax_probe"]) < 1e-14
        np.testing.assert_array_equal(b, g["b"])
    u, v, m, info = oracle.gn_solve(f0, f1, w, h, alpha, lam, return_info=True)
    assert info["info"] == 0
    assert relerr(u, g["u"]) < 1e-9 and relerr(v, g["v"]) < 1e-9 and relerr(m, g["m"]) < 1e-9


def test_gn_direct_variant_agrees(oracle):
    from oracle import gn_direct
    g = load_golden("gn_24x32")
    h, w = map(int, g["dims"]); alpha, lam = g["params"]
    f0 = g["f0_u8"].astype(np.float64).ravel() / 255; f1 = g["f1_u8"].astype(np.float64).ravel() / 255
    u, v, m = gn_direct.gn_solve(f0, f1, w, h, alpha, lam)
    assert relerr(u, g["u"]) < 1e-11 and relerr(v, g["v"]) < 1e-11 and relerr(m, g["m"]) < 1e-11
