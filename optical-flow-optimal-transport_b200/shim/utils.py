"""Drop-in for the reference's `utils` module (utils.py:25-354).

On the GPU: `opticalflow_from_benamoubrenier` (K4) and `apply_opticalflow` (K7).  Image and
.flo I/O and the error metrics are host-side reporting code that the north star leaves
unchanged in behaviour; they are re-stated here compactly with numpy so that the unmodified
reference `main.py` runs against this directory.
"""
import math
import os
import sys

import numpy as np
from PIL import Image

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import foto_b200  # noqa: E402


def openGrayscaleImage(inputPathname):
    """(flat float64 image in [0,1], width, height)   (utils.py:25-42)"""
    f = np.asarray(Image.open(inputPathname).convert('L'))
    return f.flatten() / 255, f.shape[1], f.shape[0]


def reconstructTrajectory(xStart, yStart, u, v, Nx, Ny, Nt):
    """One particle through the velocity slices u[n], v[n] (utils.py:44-99).  Scalar helper kept
    for API completeness; the solver traces all pixels at once in the K4 kernel."""
    x, y = xStart, yStart
    for n in range(Nt - 1):
        ix = max(0, min(Nx - 2, int(x))); iy = max(0, min(Ny - 2, int(y)))
        dX, dY = x - ix, y - iy
        wts = ((1 - dY) * (1 - dX), dX * (1 - dY), dY * dX, (1 - dX) * dY)
        idx = (iy * Nx + ix, iy * Nx + ix + 1, (iy + 1) * Nx + ix + 1, (iy + 1) * Nx + ix)
        x += (wts[0] * u[n, idx[0]] + wts[1] * u[n, idx[1]] + wts[2] * u[n, idx[2]] + wts[3] * u[n, idx[3]])
        y += (wts[0] * v[n, idx[0]] + wts[1] * v[n, idx[1]] + wts[2] * v[n, idx[2]] + wts[3] * v[n, idx[3]])
    return [x - xStart, y - yStart]


def _is_standard(op, name, bc):
    terms = getattr(op, "terms", None)
    return (terms is not None and len(terms) == 1 and terms[0][0] == 1.0 and terms[0][1] == name
            and terms[0][2] == bc and terms[0][7] == 1.0 and terms[0][8] == 1.0 and not terms[0][9]
            and getattr(op, "ident", 0.0) == 0.0)


def opticalflow_from_benamoubrenier(phi, Nt, Nx, Ny, grad, div):
    """(u, v, m) from the potential phi (utils.py:148-183): trajectories through grad(phi_n),
    m = -div(u, v).  `grad`/`div` must be operators.grad(Nx,Ny,1,1,'N') / operators.div(Nx,Ny,1,1,'D'),
    the pair benamou_brenier.solve passes (benamou_brenier.py:269-271): the kernel fuses them."""
    if not (_is_standard(grad, "grad", "N") and _is_standard(div, "div", "D")):
        raise NotImplementedError("the fused flow-extraction kernel implements grad(...,'N') and div(...,'D') "
                                  "with unit spacing, as benamou_brenier.solve uses them")
    return foto_b200.flow_from_phi(phi, Nt, Nx, Ny)


def apply_opticalflow(f1, u, v, w, h, m=np.array([None])):
    """Backward bilinear warp of (1+m)*f1 by (u, v) (utils.py:186-248)."""
    m = np.asarray(m)
    lum = None if (m.dtype == object or m.size != w * h) else m
    return foto_b200.warp_apply(f1, u, v, w, h, lum)


def openFlo(pathname):
    """Middlebury .flo reader: (w, h, u, v)   (utils.py:250-271)"""
    with open(pathname, 'rb') as f:
        magic = np.fromfile(f, np.float32, count=1)[0]
        if 202021.25 != magic:
            print('Magic number incorrect. Invalid .flo file')
        w = np.fromfile(f, np.int32, count=1)[0]
        h = np.fromfile(f, np.int32, count=1)[0]
        data = np.fromfile(f, np.float32).reshape((h, w, 2))
    return w, h, data[..., 0].flatten(), data[..., 1].flatten()


def saveFlo(w, h, u, v, pathname):
    """Middlebury .flo writer: float32 magic, int32 w, h, interleaved float32 (u, v) (utils.py:273-292)"""
    with open(pathname, 'wb') as f:
        np.array([202021.25], dtype=np.float32).tofile(f)
        np.array([w, h], dtype=np.int32).tofile(f)
        np.stack([np.asarray(u), np.asarray(v)], axis=1).astype(np.float32).tofile(f)


def _mean_std(vals):
    mean = np.sum(vals) / len(vals)
    return mean, np.sqrt(np.sum((vals - mean) ** 2) / len(vals))


def EE(w, h, u, v, uGT, vGT):
    """Endpoint error mean / std over pixels with EE <= 50 (utils.py:294-315)"""
    ee = np.sqrt((u - uGT) ** 2 + (v - vGT) ** 2)
    return _mean_std(ee[ee <= 50])


def AE(w, h, u, v, uGT, vGT):
    """Angular error mean / std over non-NaN pixels (utils.py:317-338)"""
    ae = np.arccos((1.0 + u * uGT + v * vGT) / (np.sqrt(1.0 + u ** 2 + v ** 2) * np.sqrt(1.0 + uGT ** 2 + vGT ** 2)))
    return _mean_std(ae[~np.isnan(ae)])


def IE(w, h, I, IGT):
    """Interpolation error (utils.py:340-354)"""
    return np.sqrt(np.sum((255 * I - 255 * IGT) ** 2) / (w * h))
