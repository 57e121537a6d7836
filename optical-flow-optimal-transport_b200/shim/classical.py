"""Drop-in for the reference's `classical` module (classical.py:25-130): the
Gennert-Negahdaripour solver.  The 3P x 3P system is never assembled; `assemble` keeps the two
frames, `A` is a matrix-free handle (A @ x runs on the GPU) and `process` runs the persistent
Jacobi-PCG kernel of libfoto_b200.so in place of the SuperLU factorisation.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import foto_b200  # noqa: E402

import operators  # noqa: E402,F401


class _GNMatrix:
    """Matrix-free stand-in for the reference's `self.A` (classical.py:106-108)."""

    def __init__(self, f1, f2, w, h, alpha, lam):
        self._args = (f1, f2, w, h, alpha, lam)
        self.shape = (3 * w * h, 3 * w * h)

    def dot(self, x):
        return foto_b200.gn_system(*self._args, x)[0]

    __matmul__ = dot


class GLLOpticalFlow(object):
    """Gennert and Negahdaripour Optical Flow Estimator."""
    NAME = "GLL"
    LUMINOSITY = True

    def __init__(self, w=0, h=0):
        self.w = w
        self.h = h
        self.alpha = 0.1

    def setAlpha(self, alpha):
        self.alpha = alpha

    def setLambda(self, lambdap):
        self.lambdap = lambdap

    def assemble(self, f1, f2):
        w, h = self.w, self.h
        alpha = self.alpha
        lambdap = self.lambdap              # AttributeError if setLambda was never called (classical.py:88)
        self._f1 = np.ascontiguousarray(f1, dtype=np.float64).reshape(-1)
        self._f2 = np.ascontiguousarray(f2, dtype=np.float64).reshape(-1)
        self.A = _GNMatrix(self._f1, self._f2, w, h, alpha, lambdap)
        self._params = (alpha, lambdap)
        self._b = None
        return self

    @property
    def b(self):
        if self._b is None:
            self._b = foto_b200.gn_system(self._f1, self._f2, self.w, self.h, *self._params,
                                          np.zeros(3 * self.w * self.h))[1]
        return self._b

    def process(self):
        u, v, m, info = foto_b200.gn_solve(self._f1, self._f2, self.w, self.h, *self._params)
        if info["info"] != 0:
            print(f"WARNING: PCG did not converge in {info['iters']} iterations.")
        return [u, v, m]
