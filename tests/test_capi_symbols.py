# -*- coding: utf-8 -*-
"""The C-ABI library loads without a GPU and exports every symbol include/adi_b200.h
declares; the Python glue fails loudly (no fallback) when no device is present."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "adi_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(adi_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    from adi_thermal_fields_b200 import _build, _capi
    _build.build()
    return _capi.load()


def test_header_symbols_all_exported(lib):
    from adi_thermal_fields_b200 import _capi
    names = _declared()
    assert len(names) >= 20
    assert set(names) == set(_capi.SYMBOLS)
    for n in names:
        assert hasattr(lib, n), n


def test_version_and_error_string(lib):
    assert b"sm_100a" in lib.adi_version()
    assert isinstance(lib.adi_last_error(), bytes)


def test_no_silent_fallback_without_gpu(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    assert lib.adi_ctx_create(0, C.byref(h)) != 0
    assert lib.adi_last_error()
    from adi_thermal_fields_b200 import adi3d_gpu_coeff as g
    import numpy as np
    with pytest.raises(RuntimeError):
        g.Grid3D(2, 2, 2, 1e-3, np.ones((2, 2, 2), bool))


def test_null_context_is_rejected(lib):
    assert lib.adi_cart_bind(None, 4, 4, 4, 1e-3) != 0
    assert lib.adi_cart_step(None, None, None, 0.1, 0.5, 1.0, 0.0, None) != 0


def test_null_context_is_rejected_by_the_newer_entry_points(lib):
    """Output path, solve-first z pass, options: argument checks come before any CUDA call."""
    n = C.c_ulonglong(0)
    assert lib.adi_text_capacity(10) >= 10 * 15
    assert lib.adi_text_format(None, None, 0, 4, 4, 4, 0, 4, 0, None, 0, C.byref(n), None) != 0
    assert lib.adi_text_write(None, b"/nonexistent/x.vtk", 0, b"", 0, None, 0, 4, 4, 4, 0, C.byref(n), None) != 0
    assert lib.adi_probe_open(None, 4, 1024) != 0
    assert lib.adi_probe_fetch(None, 0, None, 0, None, 1) != 0
    assert lib.adi_cart_zsweep_solve0(None, None, None, 0.1, 0.5, 1.0, 0.0, None) != 0
    assert lib.adi_cart_zsweep_apply(None, None, None, None, None, None, None, None, 32, None) != 0
    assert lib.adi_get_option(None, b"sparse_coeff") == -1
    assert lib.adi_set_option(None, b"sparse_coeff", 0) != 0


def test_tools_do_not_import_oracle():
    """Developer tools measure and drive; anything that checks against the oracle lives under tests/."""
    for f in os.listdir(os.path.join(ROOT, "tools")):
        if f.endswith(".py") and f != "gen_golden.py":
            txt = open(os.path.join(ROOT, "tools", f)).read()
            assert "import oracle" not in txt and "from oracle" not in txt, f


def test_product_does_not_import_oracle():
    """Only tests/, smoke() and bench.py's CPU legs may touch oracle/ (checker, never shipped)."""
    pkg = os.path.join(ROOT, "adi_thermal_fields_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dp, f), errors="replace").read()
                assert "import oracle" not in txt and "from oracle" not in txt, f
                assert "liboracle" not in txt, f
