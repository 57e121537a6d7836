"""Drop-in for the reference's `benamou_brenier` module (benamou_brenier.py:26-271).

Same function names, parameters, defaults, printed lines and exception types; the arithmetic
runs in libfoto_b200.so (hand-written sm_100a CUDA) through the C ABI of include/foto_b200.h.

Poisson back-end for stepA: environment variable FOTO_POISSON = "cg_parity" (default: the
scipy-cg recurrence the reference runs, rtol 1e-6, maxiter 1000, x0 = 0) or "cg_tight"
(rtol 1e-13: the exact-solve limit, i.e. the spsolve the author left commented at
benamou_brenier.py:84) or "dct_exact" (the same exact solve by separable DCT, ~50x faster; like
cg_tight it differs from the reference's truncated CG by ~5e-7 relative).
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import foto_b200  # noqa: E402

import operators  # noqa: E402,F401  (the reference module imports it too)
import utils      # noqa: E402,F401


def _backend():
    name = os.environ.get("FOTO_POISSON", "cg_parity")
    try:
        return {"cg_parity": foto_b200.POISSON_CG_PARITY, "cg_tight": foto_b200.POISSON_CG_TIGHT,
                "dct_exact": foto_b200.POISSON_DCT_EXACT}[name]
    except KeyError:
        raise ValueError(f"FOTO_POISSON={name!r}: expected 'cg_parity', 'cg_tight' or 'dct_exact'")


def _eps_from_A(A, r):
    """The reference passes A = -r*L_st + r*eps*I (benamou_brenier.py:201-203); recover eps."""
    if isinstance(A, operators.StencilOperator):
        return A.ident / r
    d0 = A[0, 0] if hasattr(A, "__getitem__") else None
    if d0 is None:
        raise TypeError("A must be the operator built as -r*laplacian_st + r*eps*I")
    return float(d0) / r - 3.0        # corner cell: L_ii = -3


def solve_benamou_brenier_step(mu, q, rho0, rhoT, r, A, div, Nt, Nx, Ny, dt, dx, dy):
    """stepA (benamou_brenier.py:26-91): phi = CG(A, div(mu - r q) + time-boundary terms)."""
    if not (dt == 1 and dx == 1 and dy == 1):
        raise NotImplementedError("the CUDA stepA is built for dt = dx = dy = 1, the only spacing "
                                  "benamou_brenier.solve uses (benamou_brenier.py:185-187)")
    phi, iters, info = foto_b200.stepA(mu, q, rho0, rhoT, r, _eps_from_A(A, r), Nt, Nx, Ny, backend=_backend())
    if info > 0:
        print(f"WARNING: CG did not converge in {info} iterations.")
    elif info < 0:
        raise RuntimeError("CG solver failed due to illegal input or breakdown.")
    return phi


def stepB(p, Nt, Nx, Ny):
    """Projection onto the paraboloid K (benamou_brenier.py:93-149)."""
    return foto_b200.stepB(p, Nt, Nx, Ny)


def solve(rho0, rhoT, Nt, Nx, Ny, r=1, convergence_tol=0.3, reg_epsilon=1e-3, max_it=100):
    """ALG2 iteration + flow extraction (benamou_brenier.py:151-271).  Returns (u, v, m)."""
    if Nt == 1:
        raise ZeroDivisionError("division by zero")        # n / (Nt - 1), benamou_brenier.py:194
    if max_it < 1:
        raise UnboundLocalError("cannot access local variable 'phi' where it is not associated with a value")
    u, v, m, info = foto_b200.solve(rho0, rhoT, Nt, Nx, Ny, r=r, convergence_tol=convergence_tol,
                                    reg_epsilon=reg_epsilon, max_it=max_it, backend=_backend())
    for i in range(info["n_outer"]):
        if info["cg_info"][i] > 0:
            print(f"WARNING: CG did not converge in {info['cg_info'][i]} iterations.")
        elif info["cg_info"][i] < 0:
            raise RuntimeError("CG solver failed due to illegal input or breakdown.")
        print(str(np.float64(info["crit"][i])) + " (" + str(i + 1) + "/" + str(max_it) + ")")
    return u, v, m
