"""Drop-in for the reference's `operators` module (operators.py:5-191), matrix-free.

1-D builders return small scipy CSR matrices assembled on the host from the library's
tridiagonal rows (foto_tri_coeffs); they are assembly helpers, never on the hot path.
The Kronecker operators (grad_st, div_st, laplacian_st, grad, grad_forward, div) return a
`StencilOperator`: `op @ x` launches the CUDA stencil (foto_op_apply), and the object supports
what the reference's callers do with the scipy matrices it replaces -- unary minus, scalar
multiples, sums, `.transpose()`/`.T`, `.shape`, `.todense()` (benamou_brenier.py:64,201-213;
classical.py:102-104; utils.py:165,181; test.py:5-15).
"""
import os
import sys

import numpy as np
from scipy import sparse

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import foto_b200  # noqa: E402


def _check_bc(bc):
    if bc not in ("N", "D"):
        raise NotImplementedError("These boundary conditions are not implemented")


def _tri(kind, n, h, bc):
    _check_bc(bc)
    lo, di, up = foto_b200.tri_coeffs(kind, n, h, bc)
    m = sparse.diags([lo[1:], di, up[:-1]], [-1, 0, 1], shape=(n, n), format="csr")
    return m


def grad_1d_forward_weird(n, h, bc):
    return _tri("grad_1d_forward_weird", n, h, bc)


def grad_1d_backward_weird(n, h, bc):
    return _tri("grad_1d_backward_weird", n, h, bc)


def grad_1d_central_weird(n, h, bc):
    return _tri("grad_1d_central_weird", n, h, bc)


def grad_1d_central(n, h, bc):
    return _tri("grad_1d_central", n, h, bc)


def grad_1d_forward(n, h, bc):
    return _tri("grad_1d_forward", n, h, bc)


def grad_1d_backward(n, h, bc):
    return _tri("grad_1d_backward", n, h, bc)


def lap1d(N, dx, bc):
    return _tri("lap1d", N, dx, bc)


class StencilOperator:
    """sum_i coef_i * Op_i(^T)  +  ident * I, applied on the GPU term by term."""

    __array_priority__ = 1000      # make numpy defer to our __rmul__/__rmatmul__

    def __init__(self, shape, terms=(), ident=0.0):
        self.shape = tuple(shape)
        self.terms = list(terms)   # (coef, op, bc, Nt, Nx, Ny, dt, dx, dy, transposed)
        self.ident = float(ident)

    # -- algebra ---------------------------------------------------------------------
    def _scaled(self, s):
        s = float(s)
        return StencilOperator(self.shape, [(s * t[0],) + t[1:] for t in self.terms], s * self.ident)

    def __neg__(self):
        return self._scaled(-1.0)

    def __mul__(self, s):
        if np.isscalar(s):
            return self._scaled(s)
        return NotImplemented

    __rmul__ = __mul__

    def __truediv__(self, s):
        return self._scaled(1.0 / float(s))

    def __add__(self, other):
        if isinstance(other, StencilOperator):
            if other.shape != self.shape:
                raise ValueError("inconsistent shapes")
            return StencilOperator(self.shape, self.terms + other.terms, self.ident + other.ident)
        if sparse.issparse(other) and other.shape == self.shape and self.shape[0] == self.shape[1]:
            d = other.diagonal()
            if other.nnz <= other.shape[0] and np.all(d == d[0]):      # c * identity (sparse.eye)
                return StencilOperator(self.shape, self.terms, self.ident + float(d[0]))
        return NotImplemented

    __radd__ = __add__

    def __sub__(self, other):
        return self + (-other)

    def transpose(self):
        terms = [t[:9] + (not t[9],) for t in self.terms]
        return StencilOperator(self.shape[::-1], terms, self.ident)

    @property
    def T(self):
        return self.transpose()

    # -- application -----------------------------------------------------------------
    def dot(self, x):
        x = np.asarray(x, dtype=np.float64)
        if x.ndim == 2 and x.shape[1] != 1:
            return np.stack([self.dot(x[:, j]) for j in range(x.shape[1])], axis=1)
        flat = x.reshape(-1)
        if flat.size != self.shape[1]:
            raise ValueError(f"dimension mismatch: operator is {self.shape}, vector has {flat.size}")
        out = np.zeros(self.shape[0])
        for (coef, op, bc, Nt, Nx, Ny, dt, dx, dy, tr) in self.terms:
            out += coef * foto_b200.op_apply(op, bc, Nt, Nx, Ny, dt, dx, dy, flat, transpose=tr)
        if self.ident != 0.0:
            out += self.ident * flat
        return out.reshape(x.shape) if x.ndim == 2 else out

    def __matmul__(self, x):
        if isinstance(x, StencilOperator):
            raise NotImplementedError("operator-operator products are not materialised; apply them in sequence")
        return self.dot(x)

    def tosparse(self):
        """Host-side Kronecker assembly (small grids only; for inspection, like test.py does)."""
        acc = sparse.csr_matrix(self.shape)
        for (coef, op, bc, Nt, Nx, Ny, dt, dx, dy, tr) in self.terms:
            m = _assemble(op, bc, Nt, Nx, Ny, dt, dx, dy)
            acc = acc + coef * (m.transpose() if tr else m)
        if self.ident != 0.0:
            acc = acc + self.ident * sparse.eye(self.shape[0])
        return acc.tocsr()

    def todense(self):
        return self.tosparse().todense()

    def toarray(self):
        return self.tosparse().toarray()


def _assemble(op, bc, Nt, Nx, Ny, dt, dx, dy):
    kind = {"grad_st": "grad_1d_central_weird", "div_st": "grad_1d_central_weird", "laplacian_st": "lap1d",
            "grad": "grad_1d_central", "div": "grad_1d_central", "grad_forward": "grad_1d_forward"}[op]
    Dx, Dy = _tri(kind, Nx, dx, bc), _tri(kind, Ny, dy, bc)
    Ix, Iy = sparse.eye(Nx), sparse.eye(Ny)
    x, y = sparse.kron(Iy, Dx), sparse.kron(Dy, Ix)
    if op in ("grad", "grad_forward"):
        return sparse.vstack([x, y]).tocsr()
    if op == "div":
        return sparse.hstack([x, y]).tocsr()
    Dt, It, Ixy = _tri(kind, Nt, dt, bc), sparse.eye(Nt), sparse.eye(Nx * Ny)
    t, x, y = sparse.kron(Dt, Ixy), sparse.kron(It, x), sparse.kron(It, y)
    if op == "grad_st":
        return sparse.vstack([t, x, y]).tocsr()
    if op == "div_st":
        return sparse.hstack([t, x, y]).tocsr()
    return (t + x + y).tocsr()


def _make(op, bc, Nt, Nx, Ny, dt, dx, dy):
    _check_bc(bc)
    from foto_b200.lib import op_shape
    shape = op_shape(op, Nt, Nx, Ny)
    return StencilOperator(shape, [(1.0, op, bc, int(Nt), int(Nx), int(Ny), float(dt), float(dx), float(dy), False)])


def grad_st(Nt, Nx, Ny, dt, dx, dy, bc):
    return _make("grad_st", bc, Nt, Nx, Ny, dt, dx, dy)


def div_st(Nt, Nx, Ny, dt, dx, dy, bc):
    return _make("div_st", bc, Nt, Nx, Ny, dt, dx, dy)


def laplacian_st(Nt, Nx, Ny, dt, dx, dy, bc):
    return _make("laplacian_st", bc, Nt, Nx, Ny, dt, dx, dy)


def grad(Nx, Ny, dx, dy, bc):
    return _make("grad", bc, 1, Nx, Ny, 1.0, dx, dy)


def grad_forward(Nx, Ny, dx, dy, bc='N'):
    return _make("grad_forward", bc, 1, Nx, Ny, 1.0, dx, dy)


def div(Nx, Ny, dx, dy, bc):
    return _make("div", bc, 1, Nx, Ny, 1.0, dx, dy)
