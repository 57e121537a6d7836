"""Seeded synthetic frame pairs (numpy/scipy only, no reference code needed).

The generator is the one SURVEY.md §8(d) fixes for configs 1-5: a Gaussian-filtered
uniform-noise texture, frame 2 = cubic-spline sub-pixel shift of frame 1, both on
the 8-bit lattice k/255 that the reference's image loader produces
(reference utils.py:39-42).  The brightness perturbation follows the recipe of the
reference's dataset tool (bin/create_lum_dataset.py:23-56: two random rectangles and
two random discs of +-0.25 brightness, clipped, re-quantised) but is vectorised and
driven by a numpy Generator, so it is a recipe-alike, not a byte-for-byte clone of
that tool's `random` stream.
"""
import numpy as np

# nominal Middlebury "other-data" shapes (h, w) used for config 3
MIDDLEBURY_SHAPES = {
    "Dimetrodon": (388, 584), "Hydrangea": (388, 584), "RubberWhale": (388, 584),
    "Grove2": (480, 640), "Grove3": (480, 640), "Urban2": (480, 640),
    "Urban3": (480, 640), "Venus": (380, 420),
}


def _quantise(a):
    return np.round(np.clip(a, 0.0, 1.0) * 255.0) / 255.0


def make_pair(h, w, seed=0, shift=(0.4, 0.7), sigma=3.0):
    """Return (f0, f1) as flat float64 arrays of length h*w (row-major, x fastest).

    True flow of the pair is (u, v) = (shift[1], shift[0]).
    """
    from scipy import ndimage
    rng = np.random.default_rng(seed)
    base = ndimage.gaussian_filter(rng.random((h + 16, w + 16)), sigma)
    base = (base - base.min()) / (base.max() - base.min())
    moved = ndimage.shift(base, shift, order=3, mode="nearest")
    f0 = _quantise(base[8:-8, 8:-8])
    f1 = _quantise(moved[8:-8, 8:-8])
    return np.ascontiguousarray(f0).ravel(), np.ascontiguousarray(f1).ravel()


def perturb_brightness(f, h, w, seed):
    """Two rectangles + two discs of uniform(-0.25, 0.25) brightness, clipped, 8-bit."""
    rng = np.random.default_rng(seed)
    img = np.array(f, dtype=np.float64).reshape(h, w).copy()
    yy, xx = np.mgrid[0:h, 0:w]
    for _ in range(2):
        lx = int(rng.integers(10, w)); ly = int(rng.integers(10, h))
        cx = int(rng.integers(lx // 2, w - lx // 2 + 1))
        cy = int(rng.integers(ly // 2, h - ly // 2 + 1))
        val = rng.uniform(-0.25, 0.25)
        y0, y1 = max(0, cy - ly // 2), min(h, cy + ly // 2)
        x0, x1 = max(0, cx - lx // 2), min(w, cx + lx // 2)
        img[y0:y1, x0:x1] += val
    for _ in range(2):
        rad = int(rng.integers(10, min(w, h) + 1)) / 2.0
        cx = int(rng.integers(int(rad), int(w - rad) + 1))
        cy = int(rng.integers(int(rad), int(h - rad) + 1))
        val = rng.uniform(-0.25, 0.25)
        img[(xx - cx) ** 2 + (yy - cy) ** 2 < rad ** 2] += val
    return _quantise(img).ravel()


def normalize_pair(f0, f1):
    """Joint mass / peak normalisation of the reference's bin/normalize_image.py:20-26: each frame is
    scaled to unit mass, then both are divided by the larger peak and re-quantised to 8 bits."""
    a = np.asarray(f0, dtype=np.float64) / np.sum(f0)
    b = np.asarray(f1, dtype=np.float64) / np.sum(f1)
    scale = max(a.max(), b.max())
    return _quantise(a / scale), _quantise(b / scale)


def two_squares(n=32):
    """The reference author's commented-out fixture (main.py:55-65): two shifted unit
    squares on an n x n grid.  Exercises the 'inside K' branch of the projection."""
    f0 = np.zeros((n, n)); f1 = np.zeros((n, n))
    f0[n // 6:3 * n // 6, n // 6:3 * n // 6] = 1.0
    f1[2 * n // 6:4 * n // 6, 2 * n // 6:4 * n // 6] = 1.0
    return f0.ravel(), f1.ravel()


def make_batch(n_pairs, h=388, w=584, base_seed=0):
    """n_pairs pairs of one shape: sequence s = i // 8 (own texture and shift),
    perturbation p = i % 8 applied to frame 2 (p == 0: unperturbed)."""
    pairs = []
    for i in range(n_pairs):
        s, p = divmod(i, 8)
        shift = (0.4 + 0.05 * s, 0.7 - 0.05 * s)
        f0, f1 = make_pair(h, w, seed=base_seed + s, shift=shift)
        if p:
            f1 = perturb_brightness(f1, h, w, seed=12345 + p)
        pairs.append((f0, f1))
    return pairs
