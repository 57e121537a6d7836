"""GPU: the device-resident ingest / egress rows of SURVEY.md section 8(f) -- 8-bit ingest, .flo packing, EE/AE/IE
metrics and the warp on device buffers (no host round trip between the solver and its consumers), against goldens
recorded from the reference's own utils.py (tests/golden/metrics.npz, warp.npz) and against the host-buffer API."""
import os

import numpy as np
import pytest
import torch

from conftest import load_golden

import foto_b200
from foto_b200 import synth

pytestmark = pytest.mark.gpu


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def test_ingest_u8_is_the_loader_division():
    """utils.openGrayscaleImage: uint8 / 255 in float64 (utils.py:42) -- all 256 levels, bit for bit."""
    ctx = foto_b200.Context(0)
    a = np.arange(256, dtype=np.uint8).repeat(3)
    d_in = _dev(a); d_out = torch.empty(a.size, dtype=torch.float64, device="cuda")
    ctx.ingest_u8(d_in.data_ptr(), a.size, d_out.data_ptr())
    torch.cuda.synchronize()
    np.testing.assert_array_equal(d_out.cpu().numpy(), a / 255)
    ctx.close()


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_device_metrics_vs_reference_goldens(tag):
    """EE / AE with the reference's > 50 px and NaN filters (utils.py:294-338), inputs and sums on the device."""
    g = load_golden("metrics")
    u, v, ug, vg = (g[f"{tag}/{k}"] for k in ("u", "v", "ug", "vg"))
    ctx = foto_b200.Context(0)
    d = [_dev(x) for x in (u, v, ug, vg)]
    out = torch.empty(6, dtype=torch.float64, device="cuda")
    ctx.flow_metrics_dev(*[t.data_ptr() for t in d], u.size, out.data_ptr())
    torch.cuda.synchronize()
    s = out.cpu().numpy()
    aee = s[0] / s[2]; aae = s[3] / s[5]
    sdee = np.sqrt(max(s[1] / s[2] - aee * aee, 0.0)); sdae = np.sqrt(max(s[4] / s[5] - aae * aae, 0.0))
    np.testing.assert_allclose([aee, sdee], g[f"{tag}/EE"], rtol=1e-10, atol=1e-9)
    np.testing.assert_allclose([aae, sdae], g[f"{tag}/AE"], rtol=1e-10, atol=2e-8)     # sqrt of a difference of sums near 0
    ctx.close()


@pytest.mark.parametrize("tag", ["a", "b", "d"])
def test_warp_dev_bit_exact_and_ie(tag):
    """utils.apply_opticalflow on device buffers (bit-exact vs the reference golden) + IE's sum of squares."""
    g = load_golden("warp")
    h, w = map(int, g[f"{tag}/dims"])
    f1, u, v, m = (g[f"{tag}/{k}"] for k in ("f1", "u", "v", "m"))
    igt = np.random.default_rng(3).random(h * w)
    ctx = foto_b200.Context(0)
    d = [_dev(x) for x in (f1, u, v, m, igt)]
    out = torch.empty(h * w, dtype=torch.float64, device="cuda"); ie = torch.empty(1, dtype=torch.float64, device="cuda")
    ctx.warp_dev(d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), w, h, d[3].data_ptr(), out.data_ptr(), d[4].data_ptr(), ie.data_ptr())
    torch.cuda.synchronize()
    rec = out.cpu().numpy()
    np.testing.assert_array_equal(rec, g[f"{tag}/out_m"])
    ref_ie = np.sqrt(np.sum((255 * rec - 255 * igt) ** 2) / (w * h))                  # utils.IE, utils.py:354
    assert abs(np.sqrt(float(ie.item()) / (w * h)) - ref_ie) < 1e-10 * max(ref_ie, 1.0)
    ctx.warp_dev(d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), w, h, None, out.data_ptr())
    torch.cuda.synchronize()
    np.testing.assert_array_equal(out.cpu().numpy(), g[f"{tag}/out_m0"])
    ctx.close()


def test_pack_flo_dev_matches_host_packing():
    rng = np.random.default_rng(9)
    n = 97 * 146
    u, v = rng.standard_normal(n) * 3, rng.standard_normal(n) * 3
    ctx = foto_b200.Context(0)
    du, dv = _dev(u), _dev(v)
    out = torch.empty(2 * n, dtype=torch.float32, device="cuda")
    ctx.pack_flo_dev(du.data_ptr(), dv.data_ptr(), n, out.data_ptr())
    torch.cuda.synchronize()
    np.testing.assert_array_equal(out.cpu().numpy(), np.stack([u, v], axis=1).astype(np.float32).ravel())
    ctx.close()


def test_solve_batch_u8_writes_the_reference_flo_bytes(tmp_path):
    """8-bit frames in, .flo payload out: the file equals utils.saveFlo(u, v) of the float64 host-buffer solve, byte
    for byte, and m is the same array."""
    h, w, Nt = 48, 64, 4
    pairs = synth.make_batch(3, h, w, base_seed=40)
    f0s = np.stack([np.round(p[0] * 255).astype(np.uint8) for p in pairs])
    f1s = np.stack([np.round(p[1] * 255).astype(np.uint8) for p in pairs])
    kw = dict(r=1.0, convergence_tol=0.1, reg_epsilon=1e-3, max_it=5)
    flo, ms, outer = foto_b200.solve_batch_u8(f0s, f1s, Nt, w, h, devices=list(range(foto_b200.device_count())), **kw)
    for i in range(3):
        u, v, m, info = foto_b200.solve(f0s[i] / 255, f1s[i] / 255, Nt, w, h, **kw)
        assert outer[i] == info["n_outer"]
        np.testing.assert_array_equal(ms[i], m)
        a = tmp_path / f"a{i}.flo"; b = tmp_path / f"b{i}.flo"
        foto_b200.save_flo_payload(w, h, flo[i], a)
        with open(b, "wb") as f:                                                      # utils.saveFlo, utils.py:285-292
            np.array([202021.25], dtype=np.float32).tofile(f); np.array([w, h], dtype=np.int32).tofile(f)
            np.stack([u, v], axis=1).astype(np.float32).tofile(f)
        assert a.read_bytes() == b.read_bytes()
