"""Seeded synthetic frame pairs (numpy/scipy only, no reference code needed).

The generator is the one SURVEY.md §8(d) fixes for configs 1-5: a Gaussian-filtered
uniform-noise texture, frame 2 = cubic-spline sub-pixel shift of frame 1, both on
the 8-bit lattice k/255 that the reference's image loader produces
(reference utils.py:39-42).  The brightness perturbation and the joint normalisation are
byte-for-byte restatements of the reference's dataset tools (bin/create_lum_dataset.py,
bin/normalize_image.py); tests/golden/lum.npz holds outputs of the tools themselves.
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


def _trunc8(a):
    """`np.uint8(255*np.clip(f, 0, 1))` then the loader's `/255`: what a PNG round trip through the reference's
    dataset tools does to an image (truncation, not rounding; bin/create_lum_dataset.py:57, utils.py:39-42)."""
    return np.uint8(255 * np.clip(a, 0, 1)).astype(np.float64) / 255


def perturb_brightness(f, h, w, seed):
    """Byte-for-byte restatement of the reference's illumination tool (bin/create_lum_dataset.py:23-57) for
    `random.seed(seed)`: two random rectangles, then two random discs, each adding uniform(-0.25, 0.25), then
    clip and truncate to 8 bits.  The `random` calls are made in the tool's order (L_x, L_y, r_x, r_y, v per
    rectangle; R, c_x, c_y, v per disc); the pixel loops are array slices adding the same value once per pixel,
    so the float results are the tool's.  Checked against the tool itself in tests/golden/lum.npz."""
    import random
    rnd = random.Random(seed)                            # same Mersenne stream as random.seed(seed)
    img = np.array(f, dtype=np.float64).reshape(h, w).copy()
    for _ in range(2):
        L_x = rnd.randint(10, w - 1); L_y = rnd.randint(10, h - 1)
        r_x = rnd.randint(int(L_x / 2), int(w - L_x / 2)); r_y = rnd.randint(int(L_y / 2), int(h - L_y / 2))
        v = rnd.uniform(-0.25, 0.25)
        img[int(r_y - L_y / 2):int(r_y + L_y / 2), int(r_x - L_x / 2):int(r_x + L_x / 2)] += v
    jj, ii = np.mgrid[0:h, 0:w]
    for _ in range(2):
        R = rnd.randint(10, min(w, h)) / 2
        c_x = rnd.randint(int(R), int(w - R)); c_y = rnd.randint(int(R), int(h - R))
        v = rnd.uniform(-0.25, 0.25)
        img[(ii - c_x) ** 2 + (jj - c_y) ** 2 < R ** 2] += v
    return _trunc8(img).ravel()


def normalize_pair(f0, f1):
    """Joint mass / peak normalisation of the reference's bin/normalize_image.py:20-29: each frame is scaled to
    unit mass, both are divided by the larger peak, clipped and truncated to 8 bits (what the saved PNGs hold)."""
    a = np.asarray(f0, dtype=np.float64); b = np.asarray(f1, dtype=np.float64)
    a = a / np.sum(a); b = b / np.sum(b)
    scale = max(np.max(a), np.max(b))
    return _trunc8(a / scale), _trunc8(b / scale)


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


CONFIG3_PARAMS = dict(r=1.0, convergence_tol=0.1, reg_epsilon=1e-3, max_it=100)     # reference CLI defaults (main.py:38-42)


def config3_pairs(sequences=None, perturbations=range(8)):
    """BASELINE.json config 3: 8 Middlebury-shape sequences x 8 perturbations = 64 pairs.
    Sequence s: texture seed s, sub-pixel shift (0.4 + 0.05 s, 0.7 - 0.05 s); perturbation p = 0 is the clean
    pair, p > 0 applies the reference's illumination tool (perturb_brightness, random.seed(12345 + p)) to frame 2.
    Returns [(name, h, w, f0, f1)] in sequence-major order; the one definition tools, tests and goldens share."""
    out = []
    for s, (name, (h, w)) in enumerate(MIDDLEBURY_SHAPES.items()):
        if sequences is not None and s not in sequences and name not in sequences:
            continue
        f0, f1 = make_pair(h, w, seed=s, shift=(0.4 + 0.05 * s, 0.7 - 0.05 * s))
        for p in perturbations:
            out.append((f"{name}/{p}", h, w, f0, f1 if p == 0 else perturb_brightness(f1, h, w, seed=12345 + p)))
    return out
