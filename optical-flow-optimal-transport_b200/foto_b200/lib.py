"""ctypes bindings for include/foto_b200.h."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.environ.get("FOTO_B200_LIB") or os.path.join(_HERE, "libfoto_b200.so")   # override: A/B experiments only
_CSRC = os.path.join(os.path.dirname(_HERE), "csrc")

POISSON_CG_PARITY, POISSON_CG_TIGHT, POISSON_DCT_EXACT = 0, 1, 2
OPS = {"grad_st": 0, "div_st": 1, "laplacian_st": 2, "grad": 3, "div": 4, "grad_forward": 5}
KINDS = {"grad_1d_forward_weird": 0, "grad_1d_backward_weird": 1, "grad_1d_central_weird": 2,
         "grad_1d_central": 3, "grad_1d_forward": 4, "grad_1d_backward": 5, "lap1d": 6}
BC = {"N": 0, "D": 1}
ERR_ARG, ERR_CUDA, ERR_NODEV, ERR_BREAKDOWN, ERR_TIMEOUT, ERR_NOTIMPL = -1, -2, -3, -4, -5, -6

# every symbol include/foto_b200.h declares (tests check that the .so exports exactly these)
EXPORTS = [
    "foto_last_error", "foto_version", "foto_device_count",
    "foto_ctx_create", "foto_ctx_destroy", "foto_ctx_device", "foto_ctx_set_profiling",
    "foto_ctx_reset_stats", "foto_ctx_get_stats", "foto_ctx_set_cg_variant",
    "foto_ctx_event_record", "foto_ctx_event_elapsed_ms", "foto_set_default_cg_variant", "foto_debug_onchip_prof",
    "foto_solve_dev", "foto_gn_solve_dev", "foto_solve_host", "foto_gn_solve_host",
    "foto_solve", "foto_stepB", "foto_stepA", "foto_rhs", "foto_flow_from_phi", "foto_op_apply",
    "foto_tri_coeffs", "foto_gn_solve", "foto_gn_system", "foto_warp_apply",
    "foto_solve_batch", "foto_gn_solve_batch", "foto_pack_flo", "foto_flow_metrics",
    "foto_ctx_set_stream", "foto_slab_rhs_dev", "foto_slab_prox_dev", "foto_dct_xy_dev", "foto_dct_t_solve_dev",
    "foto_flow_dev", "foto_ingest_u8_dev", "foto_pack_flo_dev", "foto_flow_metrics_dev", "foto_warp_dev",
    "foto_solve_batch_u8", "foto_slab_pack_dev", "foto_slab_cg_dev", "foto_slab_cg_state_words",
]

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


class FotoError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libfoto_b200 error {code}: {msg}")
        self.code = code


class Stats(C.Structure):
    _fields_ = [("launches", C.c_longlong), ("cg_launches", C.c_longlong), ("cg_iterations", C.c_longlong),
                ("cg_cells", C.c_longlong), ("cg_ms", C.c_double), ("rhs_ms", C.c_double),
                ("prox_ms", C.c_double), ("flow_ms", C.c_double), ("rhs_cells", C.c_longlong),
                ("prox_cells", C.c_longlong), ("gn_launches", C.c_longlong), ("gn_iterations", C.c_longlong),
                ("gn_pixels", C.c_longlong), ("gn_ms", C.c_double), ("cg_variant", C.c_int), ("prox_variant", C.c_int)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


_lib = None


def library_path():
    return _SO


def build(force=False):
    """Compile libfoto_b200.so in-tree with nvcc for sm_100a (no GPU needed to build)."""
    srcs = [os.path.join(_CSRC, f) for f in os.listdir(_CSRC) if f.endswith((".cu", ".cuh"))]
    srcs.append(os.path.join(os.path.dirname(os.path.dirname(_HERE)), "include", "foto_b200.h"))
    stale = not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs)
    if force or stale:
        subprocess.check_call(["make", "-s", "-C", _CSRC] + (["-B"] if force else []))
    return _SO


def lib():
    """Load the CUDA library.  No fallback: a missing .so is an error, not a reason to use numpy."""
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            raise ImportError(f"{_SO} is missing: build it with `make -C {_CSRC}` "
                              "(or __graft_entry__.build()); there is no CPU fallback")
        _lib = C.CDLL(_SO)
        _lib.foto_last_error.restype = C.c_char_p
        _lib.foto_ctx_destroy.restype = None
    return _lib


def last_error():
    return lib().foto_last_error().decode(errors="replace")


def _check(rc):
    if rc == 0:
        return
    msg = last_error()
    if rc == ERR_NOTIMPL:
        raise NotImplementedError(msg)
    if rc == ERR_ARG:
        raise ValueError(msg)
    raise FotoError(rc, msg)


def set_default_cg_variant(variant):
    """-1 auto, 0 streaming (textbook CG recurrences), 2 on-chip single-reduction, 3 GN only: fp64 CG with the spectral
    (DCT) preconditioner on the tensor cores.  Auto: Poisson solve on-chip when the grid fits, else streaming; GN solve
    spectral from 64 x 64 pixels on, else on-chip / streaming Jacobi-PCG."""
    _check(lib().foto_set_default_cg_variant(int(variant)))


def device_count():
    n = lib().foto_device_count()
    return max(n, 0)


def _a(x, n=None):
    x = np.ascontiguousarray(x, dtype=np.float64).reshape(-1)
    if n is not None and x.size != n:
        raise ValueError(f"expected {n} values, got {x.size}")
    return x


def _out(x, n):
    """A caller-provided output buffer: must be a C-contiguous float64 numpy array of exactly n values."""
    if not (isinstance(x, np.ndarray) and x.dtype == np.float64 and x.flags["C_CONTIGUOUS"] and x.size == n
            and x.flags["WRITEABLE"]):
        raise ValueError(f"output buffer must be a writable C-contiguous float64 numpy array of {n} values")
    return x


def _p(x):
    return x.ctypes.data_as(_dp)


def _d(x):
    return C.c_double(float(x))


def solve(rho0, rhoT, Nt, Nx, Ny, r=1.0, convergence_tol=0.3, reg_epsilon=1e-3, max_it=100,
          backend=POISSON_CG_PARITY):
    """benamou_brenier.solve on the GPU.  Returns (u, v, m, info)."""
    P = int(Nx) * int(Ny)
    rho0, rhoT = _a(rho0, P), _a(rhoT, P)
    max_it = int(max_it)
    u, v, m = np.empty(P), np.empty(P), np.empty(P)
    crit = np.zeros(max(max_it, 1)); cgi = np.zeros(max(max_it, 1), dtype=np.int32)
    cginfo = np.zeros(max(max_it, 1), dtype=np.int32); n_outer = C.c_int(0)
    _check(lib().foto_solve(_p(rho0), _p(rhoT), int(Nt), int(Nx), int(Ny), _d(r), _d(convergence_tol),
                            _d(reg_epsilon), max_it, int(backend), _p(u), _p(v), _p(m), _p(crit),
                            C.byref(n_outer), cgi.ctypes.data_as(_ip), cginfo.ctypes.data_as(_ip)))
    k = n_outer.value
    return u, v, m, dict(n_outer=k, crit=crit[:k].copy(), cg_iters=cgi[:k].copy(), cg_info=cginfo[:k].copy())


def stepB(p, Nt, Nx, Ny):
    n = int(Nt) * int(Nx) * int(Ny)
    p = _a(p, 3 * n)
    q = np.empty(3 * n)
    _check(lib().foto_stepB(_p(p), int(Nt), int(Nx), int(Ny), _p(q)))
    return q


def rhs(mu, q, rho0, rhoT, r, Nt, Nx, Ny):
    P = int(Nx) * int(Ny); N = P * int(Nt)
    mu, q, rho0, rhoT = _a(mu, 3 * N), _a(q, 3 * N), _a(rho0, P), _a(rhoT, P)
    F = np.empty(N)
    _check(lib().foto_rhs(_p(mu), _p(q), _p(rho0), _p(rhoT), _d(r), int(Nt), int(Nx), int(Ny), _p(F)))
    return F


def stepA(mu, q, rho0, rhoT, r, eps, Nt, Nx, Ny, backend=POISSON_CG_PARITY):
    P = int(Nx) * int(Ny); N = P * int(Nt)
    mu, q, rho0, rhoT = _a(mu, 3 * N), _a(q, 3 * N), _a(rho0, P), _a(rhoT, P)
    phi = np.empty(N); it = C.c_int(0); info = C.c_int(0)
    _check(lib().foto_stepA(_p(mu), _p(q), _p(rho0), _p(rhoT), _d(r), _d(eps), int(Nt), int(Nx), int(Ny),
                            int(backend), _p(phi), C.byref(it), C.byref(info)))
    return phi, it.value, info.value


def flow_from_phi(phi, Nt, Nx, Ny):
    P = int(Nx) * int(Ny)
    phi = _a(phi, P * int(Nt))
    u, v, m = np.empty(P), np.empty(P), np.empty(P)
    _check(lib().foto_flow_from_phi(_p(phi), int(Nt), int(Nx), int(Ny), _p(u), _p(v), _p(m)))
    return u, v, m


def op_shape(op, Nt, Nx, Ny):
    P = int(Nx) * int(Ny)
    N = P * (int(Nt) if OPS[op] <= 2 else 1)
    return {0: (3 * N, N), 1: (N, 3 * N), 2: (N, N), 3: (2 * P, P), 4: (P, 2 * P), 5: (2 * P, P)}[OPS[op]]


def op_apply(op, bc, Nt, Nx, Ny, dt, dx, dy, x, transpose=False):
    if bc not in BC:
        raise NotImplementedError("These boundary conditions are not implemented")
    n_out, n_in = op_shape(op, Nt, Nx, Ny)
    if transpose:
        n_out, n_in = n_in, n_out
    x = _a(x, n_in)
    out = np.empty(n_out)
    _check(lib().foto_op_apply(OPS[op], BC[bc], int(Nt), int(Nx), int(Ny), _d(dt), _d(dx), _d(dy),
                               int(bool(transpose)), _p(x), _p(out)))
    return out


def tri_coeffs(kind, n, h, bc):
    if bc not in BC:
        raise NotImplementedError("These boundary conditions are not implemented")
    lo, di, up = np.zeros(n), np.zeros(n), np.zeros(n)
    _check(lib().foto_tri_coeffs(KINDS[kind], int(n), _d(h), BC[bc], _p(lo), _p(di), _p(up)))
    return lo, di, up


def gn_solve(f1, f2, w, h, alpha, lam, rtol=0.0, max_it=0):
    P = int(w) * int(h)
    f1, f2 = _a(f1, P), _a(f2, P)
    u, v, m = np.empty(P), np.empty(P), np.empty(P); it = C.c_int(0); info = C.c_int(0)
    _check(lib().foto_gn_solve(_p(f1), _p(f2), int(w), int(h), _d(alpha), _d(lam), _d(rtol), int(max_it),
                               _p(u), _p(v), _p(m), C.byref(it), C.byref(info)))
    return u, v, m, dict(iters=it.value, info=info.value)


def gn_system(f1, f2, w, h, alpha, lam, x):
    P = int(w) * int(h)
    f1, f2, x = _a(f1, P), _a(f2, P), _a(x, 3 * P)
    y, b = np.empty(3 * P), np.empty(3 * P)
    _check(lib().foto_gn_system(_p(f1), _p(f2), int(w), int(h), _d(alpha), _d(lam), _p(x), _p(y), _p(b)))
    return y, b


def warp_apply(f1, u, v, w, h, m=None):
    P = int(w) * int(h)
    f1, u, v = _a(f1, P), _a(u, P), _a(v, P)
    out = np.empty(P)
    mp = None
    if m is not None:
        m = _a(m, P); mp = _p(m)
    _check(lib().foto_warp_apply(_p(f1), _p(u), _p(v), int(w), int(h), mp, _p(out)))
    return out


def pack_flo(u, v):
    """float32 interleaved (u, v) payload of a Middlebury .flo file, packed on the GPU."""
    u = _a(u); v = _a(v, u.size)
    out = np.empty(2 * u.size, dtype=np.float32)
    _check(lib().foto_pack_flo(_p(u), _p(v), int(u.size), out.ctypes.data_as(C.POINTER(C.c_float))))
    return out


def save_flo(w, h, u, v, pathname):
    """utils.saveFlo (utils.py:273-292) with the payload packed on the GPU; byte-identical files."""
    payload = pack_flo(u, v)
    with open(pathname, "wb") as f:
        np.array([202021.25], dtype=np.float32).tofile(f)
        np.array([w, h], dtype=np.int32).tofile(f)
        payload.tofile(f)


def flow_metrics(u, v, uGT, vGT):
    """(AEE, SDEE, AAE, SDAE) as utils.EE / utils.AE compute them (EE <= 50 and non-NaN AE filters)."""
    u = _a(u); n = u.size
    out = np.empty(6)
    _check(lib().foto_flow_metrics(_p(u), _p(_a(v, n)), _p(_a(uGT, n)), _p(_a(vGT, n)), int(n), _p(out)))
    aee = out[0] / out[2]; aae = out[3] / out[5]
    sdee = np.sqrt(max(out[1] / out[2] - aee * aee, 0.0)); sdae = np.sqrt(max(out[4] / out[5] - aae * aae, 0.0))
    return aee, sdee, aae, sdae


def _devices(devices):
    if devices is None:
        devices = list(range(max(device_count(), 1)))
    arr = np.ascontiguousarray(devices, dtype=np.int32)
    return arr, arr.ctypes.data_as(_ip), int(arr.size)


def solve_batch(rho0s, rhoTs, Nt, Nx, Ny, r=1.0, convergence_tol=0.3, reg_epsilon=1e-3, max_it=100,
                backend=POISSON_CG_PARITY, devices=None):
    """Independent pairs of one shape, sharded over `devices` by a work queue (no collective)."""
    P = int(Nx) * int(Ny)
    rho0s = np.ascontiguousarray(rho0s, dtype=np.float64).reshape(-1, P)
    rhoTs = np.ascontiguousarray(rhoTs, dtype=np.float64).reshape(-1, P)
    n = rho0s.shape[0]
    if rhoTs.shape[0] != n:
        raise ValueError("rho0s and rhoTs must hold the same number of pairs")
    us, vs, ms = np.empty((n, P)), np.empty((n, P)), np.empty((n, P))
    outer = np.zeros(n, dtype=np.int32)
    keep, dp, nd = _devices(devices)
    _check(lib().foto_solve_batch(n, _p(rho0s), _p(rhoTs), int(Nt), int(Nx), int(Ny), _d(r), _d(convergence_tol),
                                  _d(reg_epsilon), int(max_it), int(backend), dp, nd, _p(us), _p(vs), _p(ms),
                                  outer.ctypes.data_as(_ip)))
    return us, vs, ms, outer


def solve_batch_u8(f0s_u8, f1s_u8, Nt, Nx, Ny, r=1.0, convergence_tol=0.3, reg_epsilon=1e-3, max_it=100,
                   backend=POISSON_CG_PARITY, devices=None, want_m=True):
    """Batched ingest + solve + .flo egress: 8-bit grey frames in (n, P), float32 .flo payloads out (n, 2P)
    [+ m (n, P)], sharded over `devices`.  `save_flo_payload` writes a payload as a .flo file."""
    P = int(Nx) * int(Ny)
    f0s = np.ascontiguousarray(f0s_u8, dtype=np.uint8).reshape(-1, P)
    f1s = np.ascontiguousarray(f1s_u8, dtype=np.uint8).reshape(-1, P)
    n = f0s.shape[0]
    if f1s.shape[0] != n:
        raise ValueError("both frame stacks must hold the same number of images")
    flo = np.empty((n, 2 * P), dtype=np.float32)
    ms = np.empty((n, P)) if want_m else None
    outer = np.zeros(n, dtype=np.int32)
    keep, dp, nd = _devices(devices)
    u8p = C.POINTER(C.c_ubyte)
    _check(lib().foto_solve_batch_u8(n, f0s.ctypes.data_as(u8p), f1s.ctypes.data_as(u8p), int(Nt), int(Nx), int(Ny), _d(r),
                                     _d(convergence_tol), _d(reg_epsilon), int(max_it), int(backend), dp, nd,
                                     flo.ctypes.data_as(C.POINTER(C.c_float)), _p(ms) if want_m else None,
                                     outer.ctypes.data_as(_ip)))
    return flo, ms, outer


def save_flo_payload(w, h, payload, pathname):
    """utils.saveFlo's file (utils.py:273-292) from a packed float32 payload: magic, w, h, interleaved (u, v)."""
    with open(pathname, "wb") as f:
        np.array([202021.25], dtype=np.float32).tofile(f)
        np.array([w, h], dtype=np.int32).tofile(f)
        np.ascontiguousarray(payload, dtype=np.float32).tofile(f)


def gn_solve_batch(f1s, f2s, w, h, alpha, lam, rtol=0.0, max_it=0, devices=None):
    P = int(w) * int(h)
    f1s = np.ascontiguousarray(f1s, dtype=np.float64).reshape(-1, P)
    f2s = np.ascontiguousarray(f2s, dtype=np.float64).reshape(-1, P)
    n = f1s.shape[0]
    us, vs, ms = np.empty((n, P)), np.empty((n, P)), np.empty((n, P))
    iters = np.zeros(n, dtype=np.int32)
    keep, dp, nd = _devices(devices)
    _check(lib().foto_gn_solve_batch(n, _p(f1s), _p(f2s), int(w), int(h), _d(alpha), _d(lam), _d(rtol),
                                     int(max_it), dp, nd, _p(us), _p(vs), _p(ms), iters.ctypes.data_as(_ip)))
    return us, vs, ms, iters


class Context:
    """Device-resident API: pointers are raw device addresses (e.g. torch_tensor.data_ptr())."""

    def __init__(self, device=0):
        self._h = C.c_void_p()
        _check(lib().foto_ctx_create(int(device), C.byref(self._h)))
        self.device = int(device)

    def close(self):
        if self._h:
            lib().foto_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_profiling(self, on=True):
        _check(lib().foto_ctx_set_profiling(self._h, int(bool(on))))

    def set_cg_variant(self, variant):
        _check(lib().foto_ctx_set_cg_variant(self._h, int(variant)))

    def reset_stats(self):
        _check(lib().foto_ctx_reset_stats(self._h))

    def stats(self):
        s = Stats()
        _check(lib().foto_ctx_get_stats(self._h, C.byref(s)))
        return s.as_dict()

    def onchip_prof(self, enable=True):
        out = (C.c_longlong * (8 * 1024))()
        _check(lib().foto_debug_onchip_prof(self._h, int(bool(enable)), out))
        return np.array(out, dtype=np.int64).reshape(1024, 8)

    # ---- time-slab building blocks (raw device pointers; see foto_b200/slab.py) ----------
    def set_stream(self, cuda_stream=None):
        """Launch on the given CUDA stream handle (0 = legacy default stream); None: the context's own stream."""
        own = cuda_stream is None
        _check(lib().foto_ctx_set_stream(self._h, C.c_void_p(0 if own else int(cuda_stream)), int(own)))

    def slab_rhs(self, d_mu, d_q, cs, d_rho0, d_rhoT, r, gNt, n0, nloc, Nx, Ny, d_F):
        vp = C.c_void_p
        _check(lib().foto_slab_rhs_dev(self._h, vp(d_mu), vp(d_q), C.c_ulonglong(cs), vp(d_rho0), vp(d_rhoT), _d(r),
                                       int(gNt), int(n0), int(nloc), int(Nx), int(Ny), vp(d_F)))

    def slab_prox(self, d_phi, d_mu, d_q, cs, r, gNt, n0, nloc, Nx, Ny, d_out2):
        vp = C.c_void_p
        _check(lib().foto_slab_prox_dev(self._h, vp(d_phi), vp(d_mu), vp(d_q), C.c_ulonglong(cs), _d(r), int(gNt), int(n0),
                                        int(nloc), int(Nx), int(Ny), vp(d_out2)))

    def dct_xy(self, d_in, d_out, d_tmp, nplanes, gNt, Ny, Nx, inverse):
        vp = C.c_void_p
        _check(lib().foto_dct_xy_dev(self._h, vp(d_in), vp(d_out), vp(d_tmp), int(nplanes), int(gNt), int(Ny), int(Nx),
                                     int(bool(inverse))))

    def dct_t_solve(self, d_in, d_out, gNt, Ny, Nx, y_off, ny_loc, r, eps):
        vp = C.c_void_p
        _check(lib().foto_dct_t_solve_dev(self._h, vp(d_in), vp(d_out), int(gNt), int(Ny), int(Nx), int(y_off), int(ny_loc),
                                          _d(r), _d(eps)))

    def slab_cg(self, op, gNt, n0, nloc, Ny, Nx, r, eps, rtol, it, maxiter, d_b, d_x, d_r, d_p_old, d_p_new, d_q, d_state):
        vp = C.c_void_p
        _check(lib().foto_slab_cg_dev(self._h, int(op), int(gNt), int(n0), int(nloc), int(Ny), int(Nx), _d(r), _d(eps), _d(rtol),
                                      int(it), int(maxiter), vp(d_b), vp(d_x), vp(d_r), vp(d_p_old), vp(d_p_new), vp(d_q), vp(d_state)))

    def slab_pack(self, direction, nloc, Ny, Nx, world, d_in, d_out):
        vp = C.c_void_p
        _check(lib().foto_slab_pack_dev(self._h, int(direction), int(nloc), int(Ny), int(Nx), int(world), vp(d_in), vp(d_out)))

    def flow_dev(self, d_phi, Nt, Nx, Ny, d_u, d_v, d_m):
        vp = C.c_void_p
        _check(lib().foto_flow_dev(self._h, vp(d_phi), int(Nt), int(Nx), int(Ny), vp(d_u), vp(d_v), vp(d_m)))

    # ---- device-resident ingest / egress (raw device pointers) --------------------------------
    def ingest_u8(self, d_u8, n, d_out):
        _check(lib().foto_ingest_u8_dev(self._h, C.c_void_p(d_u8), int(n), C.c_void_p(d_out)))

    def pack_flo_dev(self, d_u, d_v, n, d_out_f32):
        vp = C.c_void_p
        _check(lib().foto_pack_flo_dev(self._h, vp(d_u), vp(d_v), int(n), vp(d_out_f32)))

    def flow_metrics_dev(self, d_u, d_v, d_ug, d_vg, n, d_out6):
        vp = C.c_void_p
        _check(lib().foto_flow_metrics_dev(self._h, vp(d_u), vp(d_v), vp(d_ug), vp(d_vg), int(n), vp(d_out6)))

    def warp_dev(self, d_f1, d_u, d_v, w, h, d_m, d_out, d_igt=None, d_ie=None):
        vp = C.c_void_p
        _check(lib().foto_warp_dev(self._h, vp(d_f1), vp(d_u), vp(d_v), int(w), int(h), vp(d_m) if d_m else None, vp(d_out),
                                   vp(d_igt) if d_igt else None, vp(d_ie) if d_ie else None))

    def event_record(self, which):
        _check(lib().foto_ctx_event_record(self._h, int(which)))

    def event_elapsed_ms(self):
        ms = C.c_double(0.0)
        _check(lib().foto_ctx_event_elapsed_ms(self._h, C.byref(ms)))
        return ms.value

    def solve_host(self, rho0, rhoT, Nt, Nx, Ny, u, v, m, r=1.0, convergence_tol=0.3, reg_epsilon=1e-3,
                   max_it=100, backend=POISSON_CG_PARITY):
        """Host-buffer solve into preallocated numpy outputs (H2D + solve + D2H on this context)."""
        n_outer = C.c_int(0)
        P = int(Nx) * int(Ny)
        rho0, rhoT = _a(rho0, P), _a(rhoT, P)
        for o in (u, v, m):
            _out(o, P)
        _check(lib().foto_solve_host(self._h, _p(rho0), _p(rhoT), int(Nt), int(Nx), int(Ny), _d(r),
                                     _d(convergence_tol), _d(reg_epsilon), int(max_it), int(backend),
                                     _p(u), _p(v), _p(m), None, C.byref(n_outer), None, None))
        return n_outer.value

    def gn_solve_host(self, f1, f2, w, h, alpha, lam, u, v, m, rtol=0.0, max_it=0):
        it = C.c_int(0); info = C.c_int(0)
        P = int(w) * int(h)
        f1, f2 = _a(f1, P), _a(f2, P)
        for o in (u, v, m):
            _out(o, P)
        _check(lib().foto_gn_solve_host(self._h, _p(f1), _p(f2), int(w), int(h), _d(alpha), _d(lam), _d(rtol),
                                        int(max_it), _p(u), _p(v), _p(m), C.byref(it), C.byref(info)))
        return it.value

    def solve_dev(self, d_rho0, d_rhoT, Nt, Nx, Ny, d_u, d_v, d_m, r=1.0, convergence_tol=0.3,
                  reg_epsilon=1e-3, max_it=100, backend=POISSON_CG_PARITY):
        max_it = int(max_it)
        crit = np.zeros(max(max_it, 1)); cgi = np.zeros(max(max_it, 1), dtype=np.int32)
        cginfo = np.zeros(max(max_it, 1), dtype=np.int32); n_outer = C.c_int(0)
        vp = C.c_void_p
        _check(lib().foto_solve_dev(self._h, vp(d_rho0), vp(d_rhoT), int(Nt), int(Nx), int(Ny), _d(r),
                                    _d(convergence_tol), _d(reg_epsilon), max_it, int(backend), vp(d_u), vp(d_v),
                                    vp(d_m), _p(crit), C.byref(n_outer), cgi.ctypes.data_as(_ip),
                                    cginfo.ctypes.data_as(_ip)))
        k = n_outer.value
        return dict(n_outer=k, crit=crit[:k].copy(), cg_iters=cgi[:k].copy(), cg_info=cginfo[:k].copy())

    def gn_solve_dev(self, d_f1, d_f2, w, h, alpha, lam, d_u, d_v, d_m, rtol=0.0, max_it=0):
        it = C.c_int(0); info = C.c_int(0)
        vp = C.c_void_p
        _check(lib().foto_gn_solve_dev(self._h, vp(d_f1), vp(d_f2), int(w), int(h), _d(alpha), _d(lam), _d(rtol),
                                       int(max_it), vp(d_u), vp(d_v), vp(d_m), C.byref(it), C.byref(info)))
        return dict(iters=it.value, info=info.value)
