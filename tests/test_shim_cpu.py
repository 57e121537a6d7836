"""CPU: host-side logic of the shim modules (operator algebra, Kronecker inspection path,
.flo round trip, metrics, error behaviour) -- no kernels are launched here."""
import io
import os
import sys

import numpy as np
import pytest
from scipy import sparse

from conftest import PKG, load_golden

SHIM = os.path.join(PKG, "shim")


@pytest.fixture()
def shim(monkeypatch):
    monkeypatch.syspath_prepend(SHIM)
    for name in ("operators", "utils", "benamou_brenier", "classical"):
        sys.modules.pop(name, None)
    import operators, utils, benamou_brenier, classical  # noqa: E401
    yield dict(operators=operators, utils=utils, bb=benamou_brenier, gn=classical)
    for name in ("operators", "utils", "benamou_brenier", "classical"):
        sys.modules.pop(name, None)


def test_operator_inspection_path_matches_reference(shim):
    ops = shim["operators"]
    g = load_golden("operators")
    n = 0
    for k in g.files:
        parts = k.split("/")
        name = parts[0].replace(".T", "")
        if name not in ("grad_st", "div_st", "laplacian_st", "grad", "div", "grad_forward"):
            continue
        bc, dims, dt, dx, dy = parts[1], parts[2], float(parts[3]), float(parts[4]), float(parts[5])
        Nt, Ny, Nx = map(int, dims.split("x"))
        op = getattr(ops, name)(Nt, Nx, Ny, dt, dx, dy, bc) if name.endswith("_st") else getattr(ops, name)(Nx, Ny, dx, dy, bc)
        if parts[0].endswith(".T"):
            op = op.transpose()
        M = op.tosparse()
        assert M.shape == op.shape
        three_d = name.endswith("_st")
        big, small = ("3N", "N") if three_d else ("2P", "P")
        n_small = Nt * Nx * Ny if three_d else Nx * Ny
        vec = small if (op.shape[1] == n_small and not (name == "laplacian_st" and False)) else big
        if op.shape[0] == op.shape[1]:
            vec = small
        elif op.shape[1] > op.shape[0]:
            vec = big
        x = g[f"in/{dims}/{parts[3]}/{parts[4]}/{parts[5]}/" + vec]
        assert np.max(np.abs(M @ x - g[k])) <= 1e-14 * max(1.0, np.max(np.abs(g[k]))), k
        n += 1
    assert n >= 40


def test_operator_algebra(shim):
    ops = shim["operators"]
    L = ops.laplacian_st(3, 4, 5, 1, 1, 1, "N")
    A = -2.0 * L + 2.0 * 1e-3 * sparse.eye(60)
    assert A.shape == (60, 60) and A.ident == pytest.approx(2e-3) and A.terms[0][0] == -2.0
    assert shim["bb"]._eps_from_A(A, 2.0) == pytest.approx(1e-3)
    G = ops.grad_st(3, 3, 3, 1, 1, 1, "N")
    assert G.shape == (81, 27) and G.T.shape == (27, 81) and (-G).terms[0][0] == -1.0
    D = ops.div_st(3, 3, 3, 1, 1, 1, "D")
    assert np.sum(-G.transpose().todense() - D.todense()) == 0.0       # what the reference's test.py prints
    with pytest.raises(NotImplementedError):
        ops.grad(4, 4, 1, 1, "X")
    with pytest.raises(NotImplementedError):
        ops.lap1d(4, 1, "Q")


def test_flo_roundtrip_and_metrics(shim, tmp_path):
    ut = shim["utils"]
    rng = np.random.default_rng(0)
    w, h = 7, 5
    u, v = rng.standard_normal(w * h), rng.standard_normal(w * h)
    path = str(tmp_path / "a.flo")
    ut.saveFlo(w, h, u, v, path)
    raw = open(path, "rb").read()
    assert len(raw) == 12 + 8 * w * h
    assert np.frombuffer(raw[:4], np.float32)[0] == np.float32(202021.25)
    assert tuple(np.frombuffer(raw[4:12], np.int32)) == (w, h)
    np.testing.assert_array_equal(np.frombuffer(raw[12:], np.float32).reshape(-1, 2)[:, 0], u.astype(np.float32))
    w2, h2, u2, v2 = ut.openFlo(path)
    assert (w2, h2) == (w, h)
    np.testing.assert_array_equal(u2, u.astype(np.float32)); np.testing.assert_array_equal(v2, v.astype(np.float32))
    ee, sd = ut.EE(w, h, u, v, u + 3.0, v + 4.0)
    assert ee == pytest.approx(5.0) and sd == pytest.approx(0.0, abs=1e-12)
    assert ut.IE(w, h, np.ones(w * h), np.zeros(w * h)) == pytest.approx(255.0)
    ae, _ = ut.AE(w, h, u, v, u, v)
    assert ae == pytest.approx(0.0, abs=1e-6)


def test_reference_error_behaviour(shim):
    bb, gn = shim["bb"], shim["gn"]
    with pytest.raises(ZeroDivisionError):
        bb.solve(np.ones(16), np.ones(16), 1, 4, 4)
    s = gn.GLLOpticalFlow(4, 4)
    assert s.NAME == "GLL" and s.LUMINOSITY is True and s.alpha == 0.1
    with pytest.raises(AttributeError):
        s.assemble(np.ones(16), np.ones(16))           # setLambda never called (classical.py:88)


def test_scalar_trajectory_helper_matches_reference_golden(shim):
    ut = shim["utils"]
    g = load_golden("flow")
    Nt, Nx, Ny = map(int, g["b/dims"])
    phi = g["b/phi"].reshape(Nt, Ny, Nx)
    un = np.zeros((Nt, Ny, Nx)); vn = np.zeros((Nt, Ny, Nx))
    un[:, :, 1:-1] = 0.5 * phi[:, :, 2:] - 0.5 * phi[:, :, :-2]
    vn[:, 1:-1, :] = 0.5 * phi[:, 2:, :] - 0.5 * phi[:, :-2, :]
    un = un.reshape(Nt, -1); vn = vn.reshape(Nt, -1)
    for (x, y) in [(0, 0), (3, 2), (Nx - 1, Ny - 1)]:
        du, dv = ut.reconstructTrajectory(x, y, un, vn, Nx, Ny, Nt)
        assert du == g["b/u"][y * Nx + x] and dv == g["b/v"][y * Nx + x]


def test_dataset_tools_match_the_reference_tools_byte_for_byte():
    """synth.perturb_brightness / normalize_pair against PNGs written by the reference's own
    bin/create_lum_dataset.py and bin/normalize_image.py (tests/golden/lum.npz, make_golden.py --only lum)."""
    from foto_b200 import synth
    g = load_golden("lum")
    n = 0
    for tag in ("a", "b", "c"):
        a, b = g[f"{tag}/a"], g[f"{tag}/b"]
        h, w = a.shape
        for key in [k for k in g.files if k.startswith(f"{tag}/lum/")]:
            seed = int(key.split("/")[-1])
            mine = synth.perturb_brightness(a.ravel() / 255, h, w, seed)
            np.testing.assert_array_equal(mine, g[key].ravel() / 255)      # what utils.openGrayscaleImage returns for that PNG
            n += 1
        m1, m2 = synth.normalize_pair(a.ravel() / 255, b.ravel() / 255)
        np.testing.assert_array_equal(m1, g[f"{tag}/norm1"].ravel() / 255)
        np.testing.assert_array_equal(m2, g[f"{tag}/norm2"].ravel() / 255)
    assert n == 9


def test_metrics_match_reference_goldens(shim):
    """EE / AE (with the > 50 px and NaN filters) and IE against the reference's utils.py:294-354."""
    g = load_golden("metrics")
    for tag in ("a", "b", "c"):
        h, w = map(int, g[f"{tag}/dims"])
        u, v, ug, vg = (g[f"{tag}/{k}"] for k in ("u", "v", "ug", "vg"))
        with np.errstate(invalid="ignore"):
            np.testing.assert_allclose(shim["utils"].EE(w, h, u, v, ug, vg), g[f"{tag}/EE"], rtol=1e-13, atol=1e-15)
            np.testing.assert_allclose(shim["utils"].AE(w, h, u, v, ug, vg), g[f"{tag}/AE"], rtol=1e-13, atol=1e-15)
        np.testing.assert_allclose(shim["utils"].IE(w, h, g[f"{tag}/I"], g[f"{tag}/IGT"]), g[f"{tag}/IE"], rtol=1e-14)
