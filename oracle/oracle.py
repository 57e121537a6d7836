"""ctypes front-end of oracle/libfoto_oracle.so (see foto_oracle.c for the file:line map
into the reference).  TEST INFRASTRUCTURE ONLY -- never imported by the product path."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libfoto_oracle.so")
_lib = None

OPS = {"grad_st": 0, "div_st": 1, "laplacian_st": 2, "grad": 3, "div": 4, "grad_forward": 5}
KINDS = {"grad_1d_forward_weird": 0, "grad_1d_backward_weird": 1, "grad_1d_central_weird": 2,
         "grad_1d_central": 3, "grad_1d_forward": 4, "grad_1d_backward": 5, "lap1d": 6}
BC = {"N": 0, "D": 1}

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def build(force=False):
    src = os.path.join(_HERE, "foto_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "libfoto_oracle.so"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
    return _lib


def _a(x):
    return np.ascontiguousarray(x, dtype=np.float64)


def _p(x):
    return x.ctypes.data_as(_dp)


def tri_coeffs(kind, n, h, bc):
    lo, di, up = np.zeros(n), np.zeros(n), np.zeros(n)
    rc = lib().oracle_tri_coeffs(KINDS[kind], n, C.c_double(h), BC[bc], _p(lo), _p(di), _p(up))
    if rc:
        raise ValueError("bad 1-D builder arguments")
    return lo, di, up


def tri_dense(kind, n, h, bc):
    lo, di, up = tri_coeffs(kind, n, h, bc)
    return np.diag(di) + np.diag(lo[1:], -1) + np.diag(up[:-1], 1)


def op_apply(op, bc, Nt, Nx, Ny, dt, dx, dy, x, transpose=False):
    P = Nx * Ny
    N = P * (Nt if OPS[op] <= 2 else 1)
    n_in, n_out = {0: (N, 3 * N), 1: (3 * N, N), 2: (N, N), 3: (P, 2 * P), 4: (2 * P, P), 5: (P, 2 * P)}[OPS[op]]
    if transpose:
        n_in, n_out = n_out, n_in
    x = _a(x)
    assert x.size == n_in, (x.size, n_in)
    out = np.empty(n_out)
    rc = lib().oracle_op_apply(OPS[op], BC[bc], Nt, Nx, Ny, C.c_double(dt), C.c_double(dx), C.c_double(dy),
                               int(transpose), _p(x), _p(out))
    if rc:
        raise ValueError("bad operator arguments")
    return out


def stepB(p, Nt, Nx, Ny):
    p = _a(p); n = Nt * Nx * Ny
    q = np.empty(3 * n)
    lib().oracle_stepB(_p(p), C.c_long(n), _p(q))
    return q


def rhs(mu, q, rho0, rhoT, r, Nt, Nx, Ny):
    F = np.empty(Nt * Nx * Ny)
    mu, q, rho0, rhoT = _a(mu), _a(q), _a(rho0), _a(rhoT)
    lib().oracle_rhs(_p(mu), _p(q), _p(rho0), _p(rhoT), C.c_double(r), Nt, Nx, Ny, _p(F))
    return F


def stepA(mu, q, rho0, rhoT, r, eps, Nt, Nx, Ny, rtol=1e-6, maxiter=1000):
    phi = np.empty(Nt * Nx * Ny); it = C.c_int(0)
    mu, q, rho0, rhoT = _a(mu), _a(q), _a(rho0), _a(rhoT)
    info = lib().oracle_stepA(_p(mu), _p(q), _p(rho0), _p(rhoT), C.c_double(r), C.c_double(eps), Nt, Nx, Ny,
                              C.c_double(rtol), maxiter, _p(phi), C.byref(it))
    return phi, it.value, info


def solve(rho0, rhoT, Nt, Nx, Ny, r=1, convergence_tol=0.3, reg_epsilon=1e-3, max_it=100,
          cg_rtol=1e-6, cg_maxiter=1000, return_info=False):
    P = Nx * Ny
    rho0, rhoT = _a(rho0), _a(rhoT)
    u, v, m = np.empty(P), np.empty(P), np.empty(P)
    crit = np.zeros(max_it); cgi = np.zeros(max_it, dtype=np.int32); n_outer = C.c_int(0)
    phi = np.empty(Nt * P)
    rc = lib().oracle_solve(_p(rho0), _p(rhoT), Nt, Nx, Ny, C.c_double(r), C.c_double(convergence_tol),
                            C.c_double(reg_epsilon), max_it, C.c_double(cg_rtol), cg_maxiter,
                            _p(u), _p(v), _p(m), _p(crit), C.byref(n_outer), cgi.ctypes.data_as(_ip), _p(phi))
    if rc:
        raise ValueError("bad solve arguments")
    if return_info:
        k = n_outer.value
        return u, v, m, dict(crit=crit[:k].copy(), cg_iters=cgi[:k].copy(), n_outer=k, phi=phi)
    return u, v, m


def flow_from_phi(phi, Nt, Nx, Ny):
    P = Nx * Ny
    phi = _a(phi)
    u, v, m = np.empty(P), np.empty(P), np.empty(P)
    rc = lib().oracle_flow_from_phi(_p(phi), Nt, Nx, Ny, _p(u), _p(v), _p(m))
    if rc:
        raise ValueError("bad flow arguments")
    return u, v, m


def warp_apply(f1, u, v, w, h, m=None):
    f1, u, v = _a(f1), _a(u), _a(v)
    out = np.empty(w * h)
    mp = None
    if m is not None:
        m = _a(m); mp = _p(m)
    lib().oracle_warp_apply(_p(f1), _p(u), _p(v), w, h, mp, _p(out))
    return out


def gn_system(f1, f2, w, h, alpha, lam, x):
    """(A @ x, b) of the GN system, matrix-free."""
    P = w * h
    f1, f2, x = _a(f1), _a(f2), _a(x)
    fx, fy, ft = np.empty(P), np.empty(P), np.empty(P)
    lib().oracle_gn_coeffs(_p(f1), _p(f2), w, h, _p(fx), _p(fy), _p(ft))
    y = np.empty(3 * P); b = np.empty(3 * P)
    lib().oracle_gn_apply(_p(fx), _p(fy), _p(f2), w, h, C.c_double(alpha), C.c_double(lam), _p(x), _p(y))
    lib().oracle_gn_rhs(_p(fx), _p(fy), _p(f2), _p(ft), w, h, _p(b))
    return y, b


def gn_solve(f1, f2, w, h, alpha, lam, rtol=1e-14, maxiter=20000, return_info=False):
    P = w * h
    f1, f2 = _a(f1), _a(f2)
    u, v, m = np.empty(P), np.empty(P), np.empty(P); it = C.c_int(0)
    info = lib().oracle_gn_solve(_p(f1), _p(f2), w, h, C.c_double(alpha), C.c_double(lam), C.c_double(rtol),
                                 maxiter, _p(u), _p(v), _p(m), C.byref(it))
    if info < 0:
        raise ValueError("bad GN arguments")
    if return_info:
        return u, v, m, dict(iters=it.value, info=info)
    return u, v, m
