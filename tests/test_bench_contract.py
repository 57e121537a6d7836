"""bench.py prints exactly one JSON line on stdout with the keys the driver reads (both arms)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
        "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline"}


def _run(*flags):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *flags], capture_output=True, text=True,
                       timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines                       # nothing but the JSON line on stdout
    return json.loads(lines[0])


def test_reference_arm_line():
    """--impl reference: the oracle port on the host cores, bounded sample, same metric/unit/config keys."""
    d = _run("--impl", "reference", "--steps", "1", "--warmup", "0")
    assert BASE <= set(d) and d["impl"] == "reference"
    assert d["metric"] == "FOTO frame-pairs/s at 388x584" and d["unit"] == "pairs/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["n_gpus"] == 1 and d["dtype"] == "f64" and d["steps"] >= 1
    assert d["config"]["workload"] == "foto_388x584_nt4_cli_defaults"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["e2e"]["value"] == d["value"]


@pytest.mark.gpu
def test_b200_arm_line():
    d = _run("--steps", "1", "--warmup", "1", "--pairs-per-gpu", "1")
    assert BASE | {"roofline", "gpu_launches", "clocks", "per_rank", "onchip"} <= set(d)
    assert d["value"] > 1 and d["e2e"]["value"] > 1
    assert d["e2e"]["h2d_bytes_per_step"] == 2 * 388 * 584 * 8 and d["e2e"]["d2h_bytes_per_step"] == 3 * 388 * 584 * 8
    rf = d["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic", "kernel"} <= set(rf)
    # the HBM statement is made on the streaming kernels at HD size: a fraction of the HBM peak, never above it
    assert rf["bound"] == "hbm" and rf["unit"] == "GB/s" and abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-12
    assert 0.3 < rf["frac"] < 1.0 and rf["kernel"] in ("k_rhs", "cg_stream_kernel", "k_prox_dual_tma", "k_prox_dual")
    assert all(0.3 < rf["streaming_hd"][k]["frac"] < 1.0 for k in ("K1_rhs", "K2a_cg_stream", "K3_prox_dual"))
    assert d["onchip"]["kernel"] == "cg_fused_kernel" and d["onchip"]["share_of_step"] > 0.8 and d["gpu_launches"] > 0
    assert len(d["per_rank"]) == 1 and d["per_rank"][0]["pairs"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] > 0
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
