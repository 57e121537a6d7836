"""foto_b200 -- ctypes front-end of libfoto_b200.so (hand-written sm_100a CUDA).

The library is the product; this package only marshals numpy arrays across the C ABI of
include/foto_b200.h and maps its status codes onto the exception types the reference raises.
There is no CPU fallback: importing works anywhere (so that symbols can be inspected), but
every compute call needs a CUDA device and fails loudly without one.
"""
from .lib import (  # noqa: F401
    FotoError, lib, library_path, build, device_count, last_error, set_default_cg_variant,
    solve, stepB, stepA, rhs, flow_from_phi, op_apply, tri_coeffs, gn_solve, gn_system,
    warp_apply, op_shape, solve_batch, gn_solve_batch, solve_batch_u8, save_flo_payload, pack_flo, save_flo, flow_metrics,
    Context, Stats,
    POISSON_CG_PARITY, POISSON_CG_TIGHT, POISSON_DCT_EXACT, OPS, KINDS, BC, EXPORTS,
)
