"""CPU: the C-ABI library builds, loads and exports every symbol include/foto_b200.h declares;
host-only helpers work; compute entry points fail loudly (no CPU fallback) without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT, load_golden, gpu_available

import foto_b200


def _declared():
    text = open(os.path.join(ROOT, "include", "foto_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(foto_[A-Za-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_header_symbols():
    foto_b200.build()
    lib = ctypes.CDLL(foto_b200.library_path())
    declared = _declared()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in foto_b200.h but not exported"
    assert sorted(foto_b200.EXPORTS) == declared


def test_version_and_error_string():
    assert foto_b200.lib().foto_version() >= 100
    assert isinstance(foto_b200.last_error(), str)


def test_tri_coeffs_host_helper_matches_reference_dense():
    g = load_golden("operators")
    for k in [k for k in g.files if k.startswith("dense/")]:
        _, kind, bc, n, h = k.split("/")
        lo, di, up = foto_b200.tri_coeffs(kind, int(n), float(h), bc)
        dense = np.diag(di) + np.diag(lo[1:], -1) + np.diag(up[:-1], 1)
        np.testing.assert_array_equal(dense, g[k], err_msg=k)


def test_bad_boundary_condition_is_not_implemented():
    with pytest.raises(NotImplementedError):
        foto_b200.tri_coeffs("lap1d", 5, 1.0, "X")


@pytest.mark.skipif(gpu_available(), reason="only meaningful on a box without a GPU")
def test_no_cpu_fallback_without_gpu():
    with pytest.raises(foto_b200.FotoError):
        foto_b200.stepB(np.zeros(3 * 8), 2, 2, 2)
    with pytest.raises(foto_b200.FotoError):
        foto_b200.solve(np.ones(16), np.ones(16), 4, 4, 4)
    with pytest.raises(foto_b200.FotoError):
        foto_b200.gn_solve(np.ones(16), np.ones(16), 4, 4, 0.1, 0.2)


def test_product_path_does_not_import_oracle():
    pkg = os.path.join(ROOT, "optical-flow-optimal-transport_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text and "foto_oracle" not in text, f
