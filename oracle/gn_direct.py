"""GN oracle, direct-solve variant (TEST INFRASTRUCTURE ONLY).

The reference solves its 3P x 3P system with scipy.sparse.linalg.spsolve (SuperLU,
classical.py:126).  scipy is a third-party dependency that IS present in this image, so this
file re-assembles the same matrix from the matrix-free definition (SURVEY.md appendix A:
A = diag(alpha, alpha, lambda) (x) (-Lap_Neumann) + g g^T, g = (fx, fy, -f2); classical.py:90-110)
and calls the same solver.  Used only at small sizes to cross-check oracle_gn_solve.
"""
import numpy as np
from scipy import sparse
from scipy.sparse.linalg import spsolve


def _neumann_lap_1d(n):
    main = -2.0 * np.ones(n); main[0] = main[-1] = -1.0
    return sparse.diags([np.ones(n - 1), main, np.ones(n - 1)], [-1, 0, 1], format="csr")


def assemble(f1, f2, w, h, alpha, lam):
    f1 = np.asarray(f1, dtype=np.float64).reshape(h, w); f2 = np.asarray(f2, dtype=np.float64).reshape(h, w)
    fx = np.zeros((h, w)); fy = np.zeros((h, w))
    fx[:, 1:-1] = 0.5 * (f2[:, 2:] - f2[:, :-2])
    fy[1:-1, :] = 0.5 * (f2[2:, :] - f2[:-2, :])
    ft = (f2 - f1).ravel(); fx = fx.ravel(); fy = fy.ravel(); g2 = f2.ravel()
    neg_lap = -(sparse.kron(sparse.eye(h), _neumann_lap_1d(w)) + sparse.kron(_neumann_lap_1d(h), sparse.eye(w)))
    d = sparse.diags
    A = sparse.bmat([[alpha * neg_lap + d(fx * fx), d(fx * fy), d(-fx * g2)],
                     [d(fy * fx), alpha * neg_lap + d(fy * fy), d(-fy * g2)],
                     [d(-g2 * fx), d(-g2 * fy), lam * neg_lap + d(g2 * g2)]]).tocsr()
    b = np.concatenate([-fx * ft, -fy * ft, g2 * ft])
    return A, b


def gn_solve(f1, f2, w, h, alpha, lam):
    A, b = assemble(f1, f2, w, h, alpha, lam)
    x = spsolve(A.tocsc(), b)
    P = w * h
    return x[:P], x[P:2 * P], x[2 * P:]
