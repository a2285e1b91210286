# -*- coding: utf-8 -*-
"""GPU parity of the voxel_bc_correction kernels (k_voxel_project / k_voxel_correct through the C ABI and
the reference's build_corrected_robin_fields interface) against golden outputs of the unmodified
reference.  The scatter uses fp64 atomics, so sums into one voxel are compared to 1e-12 relative; which
voxels receive area, and every fallback entry, must agree exactly."""
import os

import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["ellipsoid", "coarse"])
@pytest.mark.parametrize("fallback", [True, False])
def test_corrected_fields_match_reference(name, fallback, golden_dir):
    from adi_thermal_fields_b200 import _capi, devarray as cp, voxel_bc_correction as vb
    c = cases.build_voxel_bc_case(name)
    g = np.load(os.path.join(golden_dir, f"voxel_bc_{name}.npz"))
    n0 = _capi.load().adi_launch_count(_capi.context(0))
    robin, scale = vb.build_corrected_robin_fields(c["mesh"], c["mask"], c["origin"], c["dx"], c["base_h"],
                                                   fallback_to_base=fallback, max_subdiv=c["max_subdiv"])
    assert _capi.load().adi_launch_count(_capi.context(0)) - n0 == 2      # both kernels ran
    assert list(robin) == list(c["base_h"]) and list(scale) == list(c["base_h"])
    for f in robin:
        want = g[("robin_" if fallback else "robin_nofallback_") + f]
        got = cp.asnumpy(robin[f])
        assert np.array_equal(got != 0.0, want != 0.0)
        assert np.allclose(got, want, rtol=1e-12, atol=0.0)
        if fallback:
            assert np.allclose(cp.asnumpy(scale[f]), g["scale_" + f], rtol=1e-12, atol=0.0)


def test_fields_feed_the_pack_builder(golden_dir):
    """The corrected fields go straight into precompute_coeff_packs_unified (dict face -> 3-D field)
    and one ADI step, as in quick_compare_robin_end_robin_corrected.py:174-207."""
    from adi_thermal_fields_b200 import adi3d_gpu_coeff as ga, devarray as cp, voxel_bc_correction as vb
    from oracle import cart, voxel_bc
    c = cases.build_voxel_bc_case("coarse")
    robin, _ = vb.build_corrected_robin_fields(c["mesh"], c["mask"], c["origin"], c["dx"], c["base_h"],
                                               max_subdiv=c["max_subdiv"])
    nx, ny, nz = c["shape"]
    grid, mat = ga.Grid3D(nx, ny, nz, c["dx"], c["mask"]), ga.Material(cases.RHO, cases.CP, cases.K)
    packs = ga.precompute_coeff_packs_unified(grid, mat, robin_h=robin)
    T0 = 20.0 + 500.0 * cases.splitmix_uniform(77, c["shape"])
    kappa = cases.K / (cases.RHO * cases.CP)
    dt = 2.0 * c["dx"] ** 2 / kappa
    out = cp.asnumpy(ga.adi_step_gpu_coeff(cp.asarray(T0), grid, mat, ga.Params(dt, 0.5), packs, Tinf=20.0))
    hrobin, _ = voxel_bc.build_corrected_robin_fields(c["mesh"], c["mask"], c["origin"], c["dx"], c["base_h"],
                                                      True, c["max_subdiv"])
    hg, hm = cart.Grid3D(nx, ny, nz, c["dx"], c["mask"]), cart.Material(cases.RHO, cases.CP, cases.K)
    ref = cart.adi_step_numba_coeff(T0, hg, hm, cart.Params(dt, 0.5),
                                    cart.precompute_coeff_packs_unified(hg, hm, robin_h=hrobin), Tinf=20.0)
    assert cases.rel_l2(out, ref, c["mask"]) <= 1e-12


def test_bad_face_raises():
    from adi_thermal_fields_b200 import voxel_bc_correction as vb
    c = cases.build_voxel_bc_case("coarse")
    with pytest.raises(ValueError):
        vb.build_corrected_robin_fields(c["mesh"], c["mask"], c["origin"], c["dx"], {"w+": 1.0})
