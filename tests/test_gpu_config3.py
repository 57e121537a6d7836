"""GPU: parity on every configuration that carries a published number.

  * the 64 pairs of BASELINE.json config 3 (8 Middlebury-shape sequences x 8 illumination perturbations,
    foto_b200.synth.config3_pairs) against the CPU oracle's results (tests/golden/config3_oracle.npz, written by
    tests/golden/make_config3_oracle.py): outer-iteration count and every CG iteration count equal, u / v / m within 1e-9 relative
    on the stored sample and on whole-field sums;
  * the 32 pairs bench.py times (8 ranks x 4 seeds; tests/golden/bench_seeds_oracle.npz), same bar;
  * one pair per config-3 shape that config 1 does not cover, against goldens recorded from the UNMODIFIED reference
    (make_golden.py --only config3ref): 480x640 runs the large on-chip CG variant (x in global memory), 380x420 a
    third tile plan.
"""
import numpy as np
import pytest

from conftest import load_golden, relerr, epe_max

import foto_b200
from foto_b200 import synth

pytestmark = pytest.mark.gpu

SEQS = list(synth.MIDDLEBURY_SHAPES)


def _golden_or_skip(name):
    try:
        return load_golden(name)
    except FileNotFoundError:
        pytest.skip(f"tests/golden/{name}.npz not generated yet (tests/golden/make_config3_oracle.py)")


def _stats(a):
    return np.array([a.sum(), np.abs(a).sum(), np.abs(a).max(), np.sqrt((a * a).sum())])


FLIPS = []          # (key, outer iteration, cuda count, oracle count) of every Poisson solve whose CG count differs
SOLVES = [0]
PAIRS = [0]
FLIP_TOL = 5e-6     # relative, per flow component, for a pair with a flipped count (see _compare)


def _compare(g, key, f0, f1, h, w, kw):
    """One pair against its golden.  The reference's inner CG stops at rtol 1e-6, so its iteration count -- and with it
    the last ~1e-6 of phi -- depends on the rounding of its dot products: re-ordering the summation inside
    scipy's own recurrence (tools/cg_count_sensitivity.py) moves ||r_k||^2 by 1e-5..1e-4 relative after a few
    hundred iterations, so about 1 solve in 300 ends one iteration earlier or later in ANY implementation that is
    not bit-identical to scipy + its BLAS (the C oracle and scipy disagree with each other at that rate too).
    Such a flip is accepted here when it is +-1 iteration, recorded in FLIPS, and the pair is then held to FLIP_TOL
    and to the 1e-6 px endpoint-error bound; every other pair must have identical counts and meet 1e-9."""
    u, v, m, info = foto_b200.solve(f0, f1, 4, w, h, **kw)
    ref_cg = g[f"{key}/cg_iters"]
    assert info["n_outer"] == int(g[f"{key}/n_outer"]), key
    diff = info["cg_iters"].astype(int) - ref_cg.astype(int)
    assert np.abs(diff).max() <= 1, (key, info["cg_iters"], ref_cg)
    flipped = [(key, int(i), int(info["cg_iters"][i]), int(ref_cg[i])) for i in np.nonzero(diff)[0]]
    FLIPS.extend(flipped); SOLVES[0] += len(ref_cg); PAIRS[0] += 1
    tol = FLIP_TOL if flipped else 1e-9
    np.testing.assert_allclose(info["crit"], g[f"{key}/crit"], rtol=1e-5 if flipped else 1e-7, err_msg=key)
    sub = np.arange(0, h * w, int(g["sub_stride"]))
    for name, full in (("u", u), ("v", v), ("m", m)):
        assert relerr(full[sub], g[f"{key}/{name}"]) < tol, (key, name, flipped)
        ref = g[f"{key}/{name}_stats"]
        np.testing.assert_allclose(_stats(full), ref, rtol=tol, atol=tol * ref[1], err_msg=f"{key} {name}")
    assert epe_max(u[sub], v[sub], g[f"{key}/u"], g[f"{key}/v"]) < 1e-6, key


@pytest.mark.parametrize("seq", SEQS)
def test_config3_sequence_vs_oracle(seq):
    """All 8 perturbations of one sequence (the perturbed pairs run 50-100 outer iterations x ~700 CG iterations:
    a single flipped CG count would move the answer by 1e-6)."""
    g = _golden_or_skip("config3_oracle")
    kw = dict(synth.CONFIG3_PARAMS)
    foto_b200.set_default_cg_variant(-1)
    n = 0
    for name, h, w, f0, f1 in synth.config3_pairs(sequences=[seq]):
        _compare(g, name, f0, f1, h, w, kw)
        n += 1
    assert n == 8


@pytest.mark.parametrize("rank", range(8))
def test_bench_seed_pairs_vs_oracle(rank):
    import bench
    g, g16 = _golden_or_skip("bench_seeds_oracle"), _golden_or_skip("bench_seeds_oracle_b16")      # pairs 0-3 / 4-15 of every rank
    foto_b200.set_default_cg_variant(-1)
    for i in range(16):
        f0, f1 = synth.make_pair(bench.H, bench.W, seed=bench.pair_seed(rank, i))
        _compare(g if i < 4 else g16, f"rank{rank}/pair{i}", f0, f1, bench.H, bench.W, dict(bench.PARAMS))


def test_bench_seed_pairs_streaming_kernel_vs_oracle():
    """The textbook-recurrence streaming kernel on two of the bench pairs (the on-chip kernel covers all 128 above)."""
    import bench
    g = _golden_or_skip("bench_seeds_oracle")
    foto_b200.set_default_cg_variant(0)
    try:
        for rank, i in ((0, 1), (5, 2)):
            f0, f1 = synth.make_pair(bench.H, bench.W, seed=bench.pair_seed(rank, i))
            _compare(g, f"rank{rank}/pair{i}", f0, f1, bench.H, bench.W, dict(bench.PARAMS))
    finally:
        foto_b200.set_default_cg_variant(-1)


@pytest.mark.parametrize("name,variant", [("foto_480x640_grove2", -1), ("foto_480x640_grove2", 0),
                                          ("foto_380x420_venus", -1), ("foto_380x420_venus", 0)])
def test_config3_shapes_vs_unmodified_reference(name, variant):
    """480x640x4 takes cg_fused_kernel<384,4,6,.,XG> (x in global memory) under auto: compared here with the reference
    itself, not with another CUDA kernel."""
    g = _golden_or_skip(name)
    h, w, Nt = map(int, g["dims"]); r, tol, eps, max_it = g["params"]
    f0 = g["f0_u8"].astype(np.float64).ravel() / 255; f1 = g["f1_u8"].astype(np.float64).ravel() / 255
    foto_b200.set_default_cg_variant(variant)
    try:
        u, v, m, info = foto_b200.solve(f0, f1, Nt, w, h, r=r, convergence_tol=tol, reg_epsilon=eps, max_it=int(max_it))
    finally:
        foto_b200.set_default_cg_variant(-1)
    assert info["n_outer"] == len(g["crit"])
    np.testing.assert_array_equal(info["cg_iters"], g["cg_iters"])
    np.testing.assert_allclose(info["crit"], g["crit"], rtol=1e-7)
    sub = g["sub"]
    for comp, full in (("u", u), ("v", v), ("m", m)):
        assert relerr(full[sub], g[comp]) < 1e-9, comp
        st = _stats(full)
        np.testing.assert_allclose(st, g[comp + "_stats"], rtol=1e-9, atol=1e-9 * g[comp + "_stats"][1])
    assert epe_max(u[sub], v[sub], g["u"], g["v"]) < 1e-6


def test_auto_kernel_at_480x640_is_the_large_onchip_variant():
    import torch
    h, w = 480, 640
    f0, f1 = synth.make_pair(h, w, seed=3)
    ctx = foto_b200.Context(0)
    d0 = torch.from_numpy(f0).cuda(); d1 = torch.from_numpy(f1).cuda()
    o = [torch.empty(h * w, dtype=torch.float64, device="cuda") for _ in range(3)]
    ctx.solve_dev(d0.data_ptr(), d1.data_ptr(), 4, w, h, *[t.data_ptr() for t in o], max_it=1)
    assert ctx.stats()["cg_variant"] == 3
    ctx.close()


def test_zz_cg_count_flips_are_rare():
    """Runs last in this file: over every pair compared above, flipped CG counts must stay at the level the
    reference's own sensitivity predicts (well under 2 % of the Poisson solves) and at least 90 % of the pairs must
    have met the 1e-9 bar with identical counts."""
    if PAIRS[0] == 0:
        pytest.skip("no pair was compared in this session")
    print(f"\n{len(FLIPS)} flipped CG counts in {SOLVES[0]} Poisson solves of {PAIRS[0]} pairs: {FLIPS}")
    assert len(FLIPS) <= max(1, 0.02 * SOLVES[0])
    assert len({f[0] for f in FLIPS}) <= max(1, 0.1 * PAIRS[0])
