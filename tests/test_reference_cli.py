"""GPU + reference checkout: the UNMODIFIED reference main.py (main.py:23-25,84-146) run as a script against the shim
directory, both algorithms, with its own argparse, prints, benchmark block and .flo writer.

The reference is not part of this repository: the test looks for it at $FOTO_REFERENCE (default /root/reference,
which exists in the build container only) and is skipped where neither it nor a GPU is present.  To run it on a GPU
box, ship a scratch copy (git-ignored) and point FOTO_REFERENCE at it -- profiles/r2_reference_cli_b200.log is the
record of such a run."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import PKG, load_golden

from foto_b200 import synth

pytestmark = pytest.mark.gpu

REF = os.environ.get("FOTO_REFERENCE", "/root/reference")
needs_ref = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "main.py")),
                               reason="the reference checkout is only in the build container (set FOTO_REFERENCE)")


def _run_main(tmp_path, f0, f1, h, w, *flags):
    from PIL import Image
    for name, f in (("f0.png", f0), ("f1.png", f1)):
        Image.fromarray(np.uint8(np.round(255 * f)).reshape(h, w), "L").save(str(tmp_path / name))
    argv = ["main.py", str(tmp_path / "f0.png"), str(tmp_path / "f1.png"), *flags, "--out", str(tmp_path / "o.flo"),
            "--save-benchmark", str(tmp_path / "bench.txt"), "--save-reconstruction", str(tmp_path / "rec.png"),
            "--save-lum", str(tmp_path / "lum.png")]
    code = ("import runpy, sys; sys.path.insert(0, %r); sys.argv = %r; runpy.run_path(%r, run_name='__main__')"
            % (os.path.join(PKG, "shim"), argv, os.path.join(REF, "main.py")))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, cwd=REF)
    assert out.returncode == 0, out.stdout[-1500:] + out.stderr[-2500:]
    raw = np.fromfile(str(tmp_path / "o.flo"), np.float32)
    hdr = np.fromfile(str(tmp_path / "o.flo"), np.int32, count=3)
    assert raw[0] == np.float32(202021.25) and hdr[1] == w and hdr[2] == h
    print(out.stdout[-1200:])
    return raw[3:].reshape(-1, 2), out.stdout, tmp_path


def _flo_bytes(u, v):
    return np.stack([u, v], axis=1).astype(np.float32)


@needs_ref
def test_unmodified_main_foto_against_shim(tmp_path):
    g = load_golden("foto_24x32")
    h, w, Nt = map(int, g["dims"])
    f0 = g["f0_u8"].astype(np.float64).ravel() / 255; f1 = g["f1_u8"].astype(np.float64).ravel() / 255
    flo, stdout, _ = _run_main(tmp_path, f0, f1, h, w, "--algo=foto")
    # the per-iteration line of benamou_brenier.py:252, as many times as the reference iterated
    assert sum(1 for l in stdout.splitlines() if l.endswith(f"/100)")) == len(g["crit"])
    want = _flo_bytes(g["u"], g["v"])                     # what utils.saveFlo writes for the reference's own flow
    assert np.mean(flo != want) < 1e-3 and np.max(np.abs(flo - want)) <= 2e-7 * max(1.0, np.abs(want).max())
    assert os.path.getsize(tmp_path / "rec.png") > 0 and os.path.getsize(tmp_path / "lum.png") > 0
    assert "IE" in open(tmp_path / "bench.txt").read()


@needs_ref
def test_unmodified_main_gn_against_shim(tmp_path):
    g = load_golden("gn_24x32")
    h, w = map(int, g["dims"]); alpha, lam = g["params"]
    f0 = g["f0_u8"].astype(np.float64).ravel() / 255; f1 = g["f1_u8"].astype(np.float64).ravel() / 255
    flo, stdout, _ = _run_main(tmp_path, f0, f1, h, w, "--algo=GN", f"--alpha={alpha}", f"--lambda={lam}")   # run.sh:103 spelling
    want = _flo_bytes(g["u"], g["v"])
    assert np.mean(flo != want) < 1e-3 and np.max(np.abs(flo - want)) <= 2e-7 * max(1.0, np.abs(want).max())


@needs_ref
def test_unmodified_main_foto_flo_equals_oracle_float32(tmp_path, oracle):
    """97x146, CLI defaults: the .flo the unmodified CLI writes through the CUDA path against the float32 cast of the
    CPU oracle's flow -- byte for byte except where a 1e-12 difference straddles a float32 rounding boundary."""
    h, w = 97, 146
    f0, f1 = synth.make_pair(h, w, seed=0)
    flo, stdout, _ = _run_main(tmp_path, f0, f1, h, w, "--algo=foto")
    uo, vo, mo = oracle.solve(f0, f1, 4, w, h, r=1.0, convergence_tol=0.1, reg_epsilon=1e-3, max_it=100)
    want = _flo_bytes(uo, vo)
    n_diff = int(np.sum(flo != want))
    print(f".flo payload: {flo.size} float32 values, {n_diff} differ from the oracle's")
    assert n_diff <= max(2, flo.size // 2000) and np.max(np.abs(flo - want)) <= 2e-7 * max(1.0, np.abs(want).max())
