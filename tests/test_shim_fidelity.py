# -*- coding: utf-8 -*-
"""Fidelity of the `cupy` shim (adi_thermal_fields_b200/devarray.py): the REFERENCE's own CuPy
algorithm (adi3d_gpu_coeff.py, unmodified, imported from /root/reference when that tree is present)
executed on the shim's arrays must reproduce the reference's Numba path.  This is what an unchanged
driver does when the reference's adi3d_gpu_coeff.py shadows the drop-in (script directory first on
sys.path), and it exercises every array operation the drivers rely on (slicing assignment, boolean
masks, NumPy type promotion, where).  Skipped where the reference tree is absent (GPU box)."""
import importlib.util
import os
import sys
import tempfile

import numpy as np
import pytest
import torch

import cases

REF = os.environ.get("ADI_REFERENCE_TREE", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "adi3d_gpu_coeff.py")),
                                reason="reference tree not present")


@pytest.fixture(scope="module")
def mods():
    os.environ.setdefault("NUMBA_CACHE_DIR", tempfile.mkdtemp(prefix="numba_cache_"))
    sys.dont_write_bytecode = True
    from adi_thermal_fields_b200 import devarray
    old = devarray._FORCE_DEVICE
    devarray._FORCE_DEVICE = torch.device("cpu")
    saved = sys.modules.get("cupy")
    sys.modules["cupy"] = devarray
    try:
        def load(name):
            spec = importlib.util.spec_from_file_location("_ref_" + name, os.path.join(REF, name + ".py"))
            m = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(m)
            return m
        yield load("adi3d_gpu_coeff"), load("adi3d_numba_coeff"), devarray
    finally:
        devarray._FORCE_DEVICE = old
        if saved is None:
            sys.modules.pop("cupy", None)
        else:
            sys.modules["cupy"] = saved


@pytest.mark.parametrize("name", ["full_robin6", "holes_combined", "cyl_backend_10steps", "track_mixed",
                                  "random_neumann_fields", "thin_combined"])
def test_reference_cupy_algorithm_on_shim(name, mods):
    gpu, cpu, cp = mods
    c = cases.build_cart_case(name)
    nx, ny, nz = c["shape"]
    T0 = np.nan_to_num(c["T0"], nan=20.0)   # the CuPy algorithm multiplies by 0-couplings: no NaN in void cells
    cg, cm = cpu.Grid3D(nx, ny, nz, c["dx"], c["mask"]), cpu.Material(c["rho"], c["cp"], c["k"])
    gg, gm = gpu.Grid3D(nx, ny, nz, c["dx"], c["mask"]), gpu.Material(c["rho"], c["cp"], c["k"])
    cpk = cpu.precompute_coeff_packs_unified(cg, cm, **c["bcs"])
    gpk = gpu.precompute_coeff_packs_unified(gg, gm, **c["bcs"])
    Tc, Tg = T0.copy(), cp.asarray(T0)
    for _ in range(c["nsteps"]):
        Tc = cpu.adi_step_numba_coeff(Tc, cg, cm, cpu.Params(c["dt"], c["theta"]), cpk, Tinf=c["Tinf"])
        Tg = gpu.adi_step_gpu_coeff(Tg, gg, gm, gpu.Params(c["dt"], c["theta"]), gpk, Tinf=c["Tinf"])
    assert cases.rel_l2(cp.asnumpy(Tg), Tc, c["mask"]) <= 1e-13
