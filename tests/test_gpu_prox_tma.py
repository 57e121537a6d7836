"""GPU: the TMA-staged stepB/stepC kernel (K3, csrc/prox_tma.cu -- the kernel every volume of 2 M cells or more with even
Nx takes, i.e. configs 4 and 5 and the launch bench.py's roofline times) against (a) the reference's goldens, (b) the CPU
oracle and (c) the register-marching kernel it replaces (csrc/foto_kernels.cu: k_prox_dual), on tiles that do and do not
divide the image.  FOTO_K3=tma forces it below its size threshold, FOTO_K3=legacy switches it off."""
import numpy as np
import pytest

from conftest import load_golden, relerr, epe_max

import foto_b200
from foto_b200 import synth

pytestmark = pytest.mark.gpu


def _frames(g):
    return g["f0_u8"].astype(np.float64).ravel() / 255, g["f1_u8"].astype(np.float64).ravel() / 255


@pytest.mark.parametrize("name", ["foto_24x32", "foto_48x64", "foto_97x146", "foto_40x56_nt16_runsh"])
def test_tma_kernel_vs_reference(name, monkeypatch):
    """Whole ALG2 solves with K3 = k_prox_dual_tma against the unmodified reference's results (benamou_brenier.py:205-258):
    same outer and CG iteration counts, crit trace, u / v / m within 1e-9.  All four goldens have even Nx."""
    monkeypatch.setenv("FOTO_K3", "tma")
    g = load_golden(name)
    h, w, Nt = map(int, g["dims"]); r, tol, eps, max_it = g["params"]
    assert w % 2 == 0
    f0, f1 = _frames(g)
    u, v, m, info = foto_b200.solve(f0, f1, Nt, w, h, r=r, convergence_tol=tol, reg_epsilon=eps, max_it=int(max_it))
    assert info["n_outer"] == len(g["crit"])
    np.testing.assert_array_equal(info["cg_iters"], g["cg_iters"])
    np.testing.assert_allclose(info["crit"], g["crit"], rtol=1e-7)
    assert relerr(u, g["u"]) < 1e-9 and relerr(v, g["v"]) < 1e-9 and relerr(m, g["m"]) < 1e-9
    assert epe_max(u, v, g["u"], g["v"]) < 1e-6


# tile of the TMA kernel: 8 rows x 64 columns; shapes with partial tiles in x, in y, in both, one tile, Nt = 2 (no interior plane)
SHAPES = [(3, 37, 70), (2, 9, 66), (5, 40, 258), (4, 8, 64), (7, 64, 130), (4, 101, 2)]


@pytest.mark.parametrize("Nt,Ny,Nx", SHAPES)
def test_tma_kernel_equals_register_kernel(Nt, Ny, Nx, monkeypatch):
    """Same arithmetic per cell in both kernels: mu and q after every iteration are the same words, so u, v, m of a fixed
    number of ALG2 iterations are bit-identical; the criterion is a sum over blocks in another order (1e-13)."""
    f0, f1 = synth.make_pair(Ny, Nx, seed=Nt * 1000 + Ny)
    kw = dict(r=1.0, convergence_tol=0.0, reg_epsilon=1e-2, max_it=4, backend=foto_b200.POISSON_DCT_EXACT)
    res = {}
    for mode in ("legacy", "tma"):
        monkeypatch.setenv("FOTO_K3", mode)
        res[mode] = foto_b200.solve(f0, f1, Nt, Nx, Ny, **kw)
    a, b = res["legacy"], res["tma"]
    assert a[3]["n_outer"] == b[3]["n_outer"] == 4
    np.testing.assert_allclose(b[3]["crit"], a[3]["crit"], rtol=1e-12)
    for x, y in zip(a[:3], b[:3]):
        np.testing.assert_array_equal(x, y)


def test_tma_kernel_vs_oracle_odd_tiles(oracle, monkeypatch):
    """A shape no tile divides, against the CPU oracle.  Both sides solve the Poisson problems tightly (CG to 1e-13), so
    that what is compared at 1e-9 is stepB / stepC and not where a truncated CG happened to stop (on a 45x78 grid that
    alone is worth 1e-9, see test_solve_and_gn_odd_shapes_vs_oracle)."""
    monkeypatch.setenv("FOTO_K3", "tma")
    Nt, Ny, Nx = 3, 45, 78
    f0, f1 = synth.make_pair(Ny, Nx, seed=11)
    kw = dict(r=1.0, convergence_tol=0.1, reg_epsilon=1e-2, max_it=6)
    u, v, m, info = foto_b200.solve(f0, f1, Nt, Nx, Ny, backend=foto_b200.POISSON_CG_TIGHT, **kw)
    uo, vo, mo, io = oracle.solve(f0, f1, Nt, Nx, Ny, return_info=True, cg_rtol=1e-13, cg_maxiter=100000, **kw)
    assert info["n_outer"] == io["n_outer"]
    np.testing.assert_allclose(info["crit"], io["crit"], rtol=1e-9)
    assert relerr(u, uo) < 1e-9 and relerr(v, vo) < 1e-9 and relerr(m, mo) < 1e-9


def test_tma_kernel_is_the_default_above_two_million_cells(monkeypatch):
    """560x960x4 = 2.15 M cells (threshold: 2 097 152): auto takes the TMA kernel (no FOTO_K3), and the result is the register
    kernel's."""
    Nt, Ny, Nx = 4, 560, 960
    f0, f1 = synth.make_pair(Ny, Nx, seed=5)
    kw = dict(r=1.0, convergence_tol=0.0, reg_epsilon=1e-3, max_it=3, backend=foto_b200.POISSON_DCT_EXACT)
    monkeypatch.delenv("FOTO_K3", raising=False)
    auto = foto_b200.solve(f0, f1, Nt, Nx, Ny, **kw)
    monkeypatch.setenv("FOTO_K3", "legacy")
    legacy = foto_b200.solve(f0, f1, Nt, Nx, Ny, **kw)
    for x, y in zip(auto[:3], legacy[:3]):
        np.testing.assert_array_equal(x, y)
    np.testing.assert_allclose(auto[3]["crit"], legacy[3]["crit"], rtol=1e-12)


def test_which_kernel_ran(monkeypatch):
    """foto_stats.prox_variant proves the selection the tests above rely on: forced TMA on a small even-Nx volume, forced
    register kernel, register kernel for odd Nx whatever is asked, and TMA by default from 2 M cells on."""
    import torch

    def variant(Nt, Ny, Nx, mode):
        if mode is None: monkeypatch.delenv("FOTO_K3", raising=False)
        else: monkeypatch.setenv("FOTO_K3", mode)
        f0, f1 = synth.make_pair(Ny, Nx, seed=2)
        ctx = foto_b200.Context(0)
        d0 = torch.from_numpy(f0).cuda(); d1 = torch.from_numpy(f1).cuda()
        o = [torch.empty(Ny * Nx, dtype=torch.float64, device="cuda") for _ in range(3)]
        ctx.solve_dev(d0.data_ptr(), d1.data_ptr(), Nt, Nx, Ny, *[t.data_ptr() for t in o], max_it=1, backend=foto_b200.POISSON_DCT_EXACT)
        v = ctx.stats()["prox_variant"]
        ctx.close()
        return v

    assert variant(3, 37, 70, "tma") == 1
    assert variant(3, 37, 70, "legacy") == 0
    assert variant(3, 37, 70, None) == 0            # small volume: register kernel
    assert variant(3, 37, 71, "tma") == 0           # odd Nx: TMA strides must be multiples of 16 bytes
    assert variant(4, 540, 960, None) == 0          # 2 073 600 cells: just under the threshold of 2 097 152
    assert variant(4, 560, 960, None) == 1
    assert variant(4, 560, 960, "legacy") == 0
