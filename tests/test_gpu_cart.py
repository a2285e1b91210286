# -*- coding: utf-8 -*-
"""GPU parity of the Cartesian ADI step (sm_100a kernels through the C ABI / the
reference's adi3d_gpu_coeff interface) against golden vectors of the unmodified reference.

Bar (north_star): relative L2 <= 1e-12 per step on active cells; cells outside the mask
bit-identical to the input (adi3d_gpu_coeff.py:229).  Nothing here reads /root/reference."""
import os

import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu
TOL = 1e-12


@pytest.fixture(scope="module")
def g():
    from adi_thermal_fields_b200 import adi3d_gpu_coeff as mod
    return mod


@pytest.fixture(scope="module")
def cp():
    from adi_thermal_fields_b200 import devarray
    return devarray


def _run(g, cp, c, host_bcs=True):
    nx, ny, nz = c["shape"]
    grid = g.Grid3D(nx, ny, nz, c["dx"], c["mask"])
    mat = g.Material(c["rho"], c["cp"], c["k"])
    prm = g.Params(c["dt"], c["theta"])
    packs = g.precompute_coeff_packs_unified(grid, mat, robin_Tinf=c["Tinf"], **c["bcs"])
    T = cp.asarray(c["T0"])
    trace = []
    for _ in range(c["nsteps"]):
        T = g.adi_step_gpu_coeff(T, grid, mat, prm, packs, Tinf=c["Tinf"])
        trace.append(cp.asnumpy(T[nx // 2, ny // 2, :]))
    return cp.asnumpy(T), np.array(trace), packs


@pytest.mark.parametrize("name", sorted(cases.CART_CASES))
def test_step_matches_reference(name, g, cp, golden_dir):
    c = cases.build_cart_case(name)
    gold = np.load(os.path.join(golden_dir, f"cart_{name}.npz"))
    T, trace, _ = _run(g, cp, c)
    m = c["mask"]
    assert cases.rel_l2(T, gold["T_out"], m) <= TOL
    assert np.array_equal(T[~m], c["T0"][~m], equal_nan=True)
    if m[c["shape"][0] // 2, c["shape"][1] // 2].any():
        assert cases.rel_l2(trace, gold["trace"], np.isfinite(gold["trace"])) <= TOL


@pytest.mark.parametrize("name", sorted(n for n, v in cases.CART_CASES.items() if v[3] != 1.0))
def test_fused_explicit_stage_variant(name, g, cp, golden_dir):
    """Engine option fuse=1: the explicit stage (adi3d_numba_coeff.py:298) applied while the x
    sweep loads its lines, instead of as its own streaming pass."""
    c = cases.build_cart_case(name)
    gold = np.load(os.path.join(golden_dir, f"cart_{name}.npz"))
    g.set_option("fuse", 1)
    try:
        T, _, _ = _run(g, cp, c)
    finally:
        g.set_option("fuse", 0)
    m = c["mask"]
    assert cases.rel_l2(T, gold["T_out"], m) <= TOL
    assert np.array_equal(T[~m], c["T0"][~m], equal_nan=True)


@pytest.mark.parametrize("name", ["holes_combined", "track_mixed", "random_neumann_fields"])
def test_device_pack_builder_bit_exact(name, g, cp, golden_dir):
    """precompute_coeff_packs_unified on the device (kernel k_build_packs) vs the reference."""
    c = cases.build_cart_case(name)
    gold = np.load(os.path.join(golden_dir, f"packs_{name}.npz"))
    _, _, packs = _run(g, cp, c)
    for a, ax in enumerate("xyz"):
        assert np.array_equal(cp.asnumpy(packs[a].coeff), gold[f"coeff_{ax}"])
        assert np.array_equal(cp.asnumpy(packs[a].qflux), gold[f"q_{ax}"])
    assert np.array_equal(cp.asnumpy(packs[0].dir_mask), gold["dir_mask"])
    assert np.array_equal(cp.asnumpy(packs[0].dir_val), gold["dir_val"])


@pytest.mark.parametrize("face", ["x-", "x+", "y-", "y+", "z-", "z+"])
def test_exposed_mask(face, g, cp):
    from oracle import cart
    m = cases.make_mask("thin", cases.SHAPE_A, 3) | cases.make_mask("cyl_holes", cases.SHAPE_A, 5)
    assert np.array_equal(cp.asnumpy(g.exposed_mask(m, face)), cart.exposed_mask(m, face))


def test_exposed_mask_bad_face(g):
    with pytest.raises(ValueError):
        g.exposed_mask(np.ones((3, 3, 3), bool), "w+")


@pytest.mark.parametrize("name", ["full_robin6", "cyl_robin6", "track_robin6", "B_full_robin6"])
def test_scalar_robin_lazy_packs(name, g, cp, golden_dir):
    """Scalar Robin packs stay symbolic (no dense coeff array) and still match; reading
    pack.coeff afterwards materialises the reference's dense array."""
    from oracle import cart
    c = cases.build_cart_case(name)
    gold = np.load(os.path.join(golden_dir, f"cart_{name}.npz"))
    T, _, packs = _run(g, cp, c)
    assert all(p._coeff is None for p in packs)
    assert cases.rel_l2(T, gold["T_out"], c["mask"]) <= TOL
    nx, ny, nz = c["shape"]
    ref = cart.precompute_coeff_packs_unified(cart.Grid3D(nx, ny, nz, c["dx"], c["mask"]),
                                              cart.Material(c["rho"], c["cp"], c["k"]), **c["bcs"])
    for a in range(3):
        assert np.array_equal(cp.asnumpy(packs[a].coeff), ref[a].coeff)
        assert not cp.asnumpy(packs[a].qflux).any()


def test_user_built_packs_and_oracle(g, cp):
    """AxisCoeffPack objects built by the caller from host arrays (different Dirichlet masks
    per axis), checked against the oracle on the same inputs."""
    from oracle import cart
    shape = (21, 18, 35)
    mask = cases.make_mask("random", shape, 9) | cases.make_mask("cyl", shape, 9)
    T0 = 20.0 + 500.0 * cases.splitmix_uniform(901, shape)
    kappa = cases.K / (cases.RHO * cases.CP)
    dt = 1.7 * cases.DX ** 2 / kappa
    hp, dp_ = [], []
    for a in range(3):
        coeff = 2.0 * cases.splitmix_uniform(910 + a, shape)
        dm = cases.splitmix_uniform(920 + a, shape) < 0.05
        dv = 100.0 * cases.splitmix_uniform(930 + a, shape)
        q = 50.0 * (cases.splitmix_uniform(940 + a, shape) - 0.5)
        hp.append(cart.AxisCoeffPack(coeff, dm, dv, q))
        dp_.append(g.AxisCoeffPack(coeff, dm, dv, q))
    nx, ny, nz = shape
    ref = cart.adi_step_numba_coeff(T0, cart.Grid3D(nx, ny, nz, cases.DX, mask),
                                    cart.Material(cases.RHO, cases.CP, cases.K), cart.Params(dt, 0.5),
                                    hp, Tinf=15.0)
    out = g.adi_step_gpu_coeff(cp.asarray(T0), g.Grid3D(nx, ny, nz, cases.DX, mask),
                               g.Material(cases.RHO, cases.CP, cases.K), g.Params(dt, 0.5), dp_, Tinf=15.0)
    assert cases.rel_l2(cp.asnumpy(out), ref, mask) <= TOL


def test_host_array_entry_point(g, golden_dir):
    """adi_cart_step_host: NumPy in, NumPy out (H2D + steps + D2H inside the C ABI)."""
    c = cases.build_cart_case("cyl_backend_10steps")
    gold = np.load(os.path.join(golden_dir, "cart_cyl_backend_10steps.npz"))
    nx, ny, nz = c["shape"]
    grid = g.Grid3D(nx, ny, nz, c["dx"], c["mask"])
    mat = g.Material(c["rho"], c["cp"], c["k"])
    packs = g.precompute_coeff_packs_unified(grid, mat, **c["bcs"])
    T = g.adi_step_host(c["T0"], grid, mat, g.Params(c["dt"], c["theta"]), packs, Tinf=c["Tinf"],
                        nsteps=c["nsteps"])
    assert isinstance(T, np.ndarray)
    assert cases.rel_l2(T, gold["T_out"], c["mask"]) <= TOL


def test_input_not_modified_and_new_array(g, cp):
    c = cases.build_cart_case("cyl_robin6")
    nx, ny, nz = c["shape"]
    grid = g.Grid3D(nx, ny, nz, c["dx"], c["mask"])
    mat = g.Material(c["rho"], c["cp"], c["k"])
    packs = g.precompute_coeff_packs_unified(grid, mat, **c["bcs"])
    T = cp.asarray(c["T0"])
    out = g.adi_step_gpu_coeff(T, grid, mat, g.Params(c["dt"], 0.5), packs, Tinf=20.0)
    assert out is not T and out._t.data_ptr() != T._t.data_ptr()
    assert np.array_equal(cp.asnumpy(T), c["T0"])


def test_mask_mutation_and_rebinding_are_seen(g, cp):
    """Layer birth: grid.mask mutated in place on the device (quick_compare_layer_birth_robin_v3.py:
    272-277) and rebound to a host array (waam_from_stl_v7_mm.py:494-495)."""
    from oracle import cart
    shape = (16, 17, 40)
    nx, ny, nz = shape
    mask = np.zeros(shape, bool)
    mask[:, :, :20] = cases.build_cyl_mask(nx, ny, 20, cases.DX, 0.4 * nx * cases.DX)
    T0 = np.full(shape, 20.0)
    kappa = cases.K / (cases.RHO * cases.CP)
    dt = 2.0 * cases.DX ** 2 / kappa
    bcs = dict(robin_h={f: 30.0 for f in cart.FACES})
    mat_h = cart.Material(cases.RHO, cases.CP, cases.K)
    grid_d = g.Grid3D(nx, ny, nz, cases.DX, mask)
    mat_d = g.Material(cases.RHO, cases.CP, cases.K)
    Td = cp.asarray(T0)
    Th = T0.copy()
    mh = mask.copy()
    for birth in range(3):
        k0, k1 = 20 + 5 * birth, 25 + 5 * birth
        born = np.zeros(shape, bool)
        born[:, :, k0:k1] = mask[:, :, :1]
        # host side (oracle)
        Th[born] = 1000.0
        mh |= born
        grid_h = cart.Grid3D(nx, ny, nz, cases.DX, mh)
        ph = cart.precompute_coeff_packs_unified(grid_h, mat_h, **bcs)
        # device side, alternating the two mutation styles
        if birth % 2 == 0:
            bd = cp.asarray(born)
            Td[bd] = 1000.0
            if not isinstance(grid_d.mask, cp.ndarray):   # rebound to a host array last birth
                grid_d.mask = cp.asarray(grid_d.mask)
            grid_d.mask[bd] = True
        else:
            idx = np.where(born)
            Td[idx] = 1000.0
            grid_d.mask = mh.copy()
        pd = g.precompute_coeff_packs_unified(grid_d, mat_d, **bcs)
        for _ in range(2):
            Th = cart.adi_step_numba_coeff(Th, grid_h, mat_h, cart.Params(dt, 0.5), ph, Tinf=20.0)
            Td = g.adi_step_gpu_coeff(Td, grid_d, mat_d, g.Params(dt, 0.5), pd, Tinf=20.0)
        assert cases.rel_l2(cp.asnumpy(Td), Th, mh) <= TOL
    # scalar-Robin packs are symbolic (the kernel derives the coefficients from the mask bound at step time); the
    # reference freezes them at precompute time, so stepping with packs older than the mask is refused
    grid_d.mask = cp.asarray(mh)
    grid_d.mask[:, :, 36:38] = True
    with pytest.raises(RuntimeError):
        g.adi_step_gpu_coeff(Td, grid_d, mat_d, g.Params(dt, 0.5), pd, Tinf=20.0)


def test_shape_errors(g, cp):
    with pytest.raises(AssertionError):
        g.Grid3D(4, 4, 4, 1e-3, np.ones((4, 4, 5), bool))
    grid = g.Grid3D(4, 4, 4, 1e-3, np.ones((4, 4, 4), bool))
    mat = g.Material(1.0, 1.0, 1.0)
    packs = g.precompute_coeff_packs_unified(grid, mat)
    with pytest.raises(ValueError):
        g.adi_step_gpu_coeff(cp.zeros((4, 4, 5)), grid, mat, g.Params(1e-3), packs)
    # a float32 field is promoted, as the reference does on its first step (waam_from_stl_v7_mm.py --precision float32)
    T32 = cp.asarray(np.linspace(20.0, 900.0, 64, dtype=np.float32).reshape(4, 4, 4))
    o32 = g.adi_step_gpu_coeff(T32, grid, mat, g.Params(1e-3), packs)
    o64 = g.adi_step_gpu_coeff(cp.asarray(cp.asnumpy(T32).astype(np.float64)), grid, mat, g.Params(1e-3), packs)
    assert o32.dtype == np.float64 and np.array_equal(cp.asnumpy(o32), cp.asnumpy(o64))
    with pytest.raises(ValueError):
        g.AxisCoeffPack(np.zeros((4, 4, 4)), np.zeros((4, 4, 5), bool), np.zeros((4, 4, 4)))
    with pytest.raises(ValueError):
        g.precompute_coeff_packs_unified(grid, mat, dir_mask=np.zeros((4, 4, 5), bool), dir_value=1.0)


# ---- size-independent properties at the benchmark's full size (512^3) -------------------
@pytest.fixture(scope="module")
def big(g, cp):
    import torch
    n = 512
    mask = torch.zeros((n, n, n), dtype=torch.bool, device="cuda")
    mask[:, :, :504] = True
    mask[:16, :256, 504:] = True            # single_track_on_plate.py:113-114,159
    grid = g.Grid3D.__new__(g.Grid3D)
    grid.nx = grid.ny = grid.nz = n
    grid.dx = 1e-3
    grid.mask = cp.ndarray(mask)
    mat = g.Material(7800.0, 500.0, 25.0)
    return grid, mat


def test_full_size_heat_conservation_and_void(g, cp, big):
    """Insulated body (no Robin/Neumann/Dirichlet): every directional operator has zero column
    sums, so the ADI step conserves the sum over active cells; void cells pass through."""
    import torch
    grid, mat = big
    packs = g.precompute_coeff_packs_unified(grid, mat)
    gen = torch.Generator(device="cuda").manual_seed(0)
    T = torch.rand(grid.mask.shape, dtype=torch.float64, device="cuda", generator=gen) * 1380.0 + 20.0
    out = g.adi_step_gpu_coeff(cp.ndarray(T), grid, mat, g.Params(0.02, 0.5), packs, Tinf=20.0)._t
    m = grid.mask._t
    s0, s1 = T[m].sum().item(), out[m].sum().item()
    assert abs(s1 - s0) <= 1e-12 * abs(s0)
    assert torch.equal(out[~m], T[~m])


def test_full_size_linearity_and_ambient_fixed_point(g, cp, big):
    import torch
    grid, mat = big
    h = {f: 10.0 for f in ("x-", "x+", "y-", "y+", "z-", "z+")}
    packs = g.precompute_coeff_packs_unified(grid, mat, robin_h=h)
    prm = g.Params(0.02, 0.5)
    m = grid.mask._t
    # a field at the ambient temperature stays there
    T = torch.full(m.shape, 20.0, dtype=torch.float64, device="cuda")
    out = g.adi_step_gpu_coeff(cp.ndarray(T), grid, mat, prm, packs, Tinf=20.0)._t
    assert (out[m] - 20.0).abs().max().item() <= 1e-11
    # linear in T when Tinf = 0
    gen = torch.Generator(device="cuda").manual_seed(1)
    A = torch.rand(m.shape, dtype=torch.float64, device="cuda", generator=gen)
    B = torch.rand(m.shape, dtype=torch.float64, device="cuda", generator=gen)
    step = lambda X: g.adi_step_gpu_coeff(cp.ndarray(X), grid, mat, prm, packs, Tinf=0.0)._t
    lhs = step(2.0 * A - 3.0 * B)
    rhs = 2.0 * step(A) - 3.0 * step(B)
    err = ((lhs - rhs)[m].norm() / rhs[m].norm()).item()
    assert err <= 1e-13


def _oracle_case(shape, seed, theta=0.5, cfl=1.3, holes=True):
    """Seeded case checked against the oracle directly (no golden file): dense per-face Robin h,
    a Neumann face, Dirichlet plane, random holes, NaN in the void."""
    from oracle import cart
    nx, ny, nz = shape
    mask = np.ones(shape, bool)
    if holes:
        mask &= cases.splitmix_uniform(seed, shape) > 0.15
    T0 = 20.0 + 900.0 * cases.splitmix_uniform(seed + 1, shape)
    T0[~mask] = np.nan
    kappa = cases.K / (cases.RHO * cases.CP)
    dt = cfl * cases.DX ** 2 / kappa
    h = {f: 400.0 * cases.splitmix_uniform(seed + 2 + i, shape) for i, f in enumerate(cart.FACES)}
    dm = cases.splitmix_uniform(seed + 20, shape) < 0.02
    bcs = dict(robin_h=h, neumann={"z-": 2.0e5, "x+": 1.0e5}, dir_mask=dm, dir_value=333.0)
    return dict(shape=shape, mask=mask, T0=T0, dt=dt, theta=theta, bcs=bcs, kappa=kappa)


def _both(g, cp, c, nsteps=2, Tinf=20.0):
    from oracle import cart
    nx, ny, nz = c["shape"]
    gh = cart.Grid3D(nx, ny, nz, cases.DX, c["mask"])
    mh = cart.Material(cases.RHO, cases.CP, cases.K)
    ph = cart.precompute_coeff_packs_unified(gh, mh, **c["bcs"])
    gd = g.Grid3D(nx, ny, nz, cases.DX, c["mask"])
    md = g.Material(cases.RHO, cases.CP, cases.K)
    pd = g.precompute_coeff_packs_unified(gd, md, **c["bcs"])
    Th, Td = c["T0"], cp.asarray(c["T0"])
    for _ in range(nsteps):
        Th = cart.adi_step_numba_coeff(Th, gh, mh, cart.Params(c["dt"], c["theta"]), ph, Tinf=Tinf)
        Td = g.adi_step_gpu_coeff(Td, gd, md, g.Params(c["dt"], c["theta"]), pd, Tinf=Tinf)
    Td = cp.asnumpy(Td)
    m = c["mask"]
    assert cases.rel_l2(Td, Th, m) <= TOL
    assert np.array_equal(Td[~m], c["T0"][~m], equal_nan=True)


@pytest.mark.parametrize("shape", [(1100, 3, 6), (5, 1300, 4), (3, 4, 2100), (600, 7, 520), (2, 2050, 34)])
def test_long_lines_select_the_M32_variants(shape, g, cp):
    """Lines of 513..1024 cells run the M=32 / 256-thread kernels, 1025..4096 the 512-thread ones."""
    _both(g, cp, _oracle_case(shape, seed=4000 + shape[0]))


@pytest.mark.parametrize("opts", [dict(m=32), dict(m=32, kt=16), dict(kt=4), dict(lt=2), dict(m=32, lt=4)])
@pytest.mark.parametrize("shape", [(40, 67, 130), (96, 33, 64)])
def test_kernel_variant_options_agree(shape, opts, g, cp):
    """Every tunable launch shape (adi_set_option) gives the same answer."""
    for k, v in opts.items():
        g.set_option(k, v)
    try:
        _both(g, cp, _oracle_case(shape, seed=5000 + shape[2]), nsteps=1)
        _both(g, cp, _oracle_case(shape, seed=5100 + shape[2], theta=1.0, cfl=40.0), nsteps=1)
    finally:
        for k in opts:
            g.set_option(k, 0)


def test_ragged_and_tiny_grids(g, cp):
    for shape in [(1, 1, 1), (2, 1, 3), (1, 17, 1), (16, 16, 16), (17, 33, 15), (31, 2, 47)]:
        _both(g, cp, _oracle_case(shape, seed=6000 + sum(shape), holes=shape != (1, 1, 1)), nsteps=1)


@pytest.mark.parametrize("shape", [(6, 7, 1000), (700, 5, 16), (5, 1024, 24), (5, 1100, 8), (2100, 3, 8), (3, 4, 1500)],
                         ids=lambda s: "x".join(map(str, s)))
@pytest.mark.parametrize("bk,theta", [("robin6", 0.5), ("combined", 0.5), ("robin_dict3d", 1.0)])
def test_long_lines(shape, bk, theta, g, cp):
    """Lines of 513..1024 cells (16-cell chunks, 512-thread blocks) and longer (32-cell chunks),
    as in BASELINE configs 4 and 5, against the oracle on the same inputs."""
    from oracle import cart
    mask = cases.make_mask("cyl_holes" if min(shape[:2]) > 4 else "random", shape, 21)
    bcs = cases.make_bcs(bk, shape, mask, 21, 20.0)
    T0 = 20.0 + 1380.0 * cases.splitmix_uniform(22, shape)
    T0[~mask] = np.nan
    kappa = cases.K / (cases.RHO * cases.CP)
    dt = 2.0 * cases.DX ** 2 / kappa
    nx, ny, nz = shape
    hg, hm = cart.Grid3D(nx, ny, nz, cases.DX, mask), cart.Material(cases.RHO, cases.CP, cases.K)
    ref = cart.adi_step_numba_coeff(T0, hg, hm, cart.Params(dt, theta),
                                    cart.precompute_coeff_packs_unified(hg, hm, **bcs), Tinf=20.0)
    grid, mat = g.Grid3D(nx, ny, nz, cases.DX, mask), g.Material(cases.RHO, cases.CP, cases.K)
    packs = g.precompute_coeff_packs_unified(grid, mat, **bcs)
    for fuse in (0, 1):
        g.set_option("fuse", fuse)
        try:
            out = cp.asnumpy(g.adi_step_gpu_coeff(cp.asarray(T0), grid, mat, g.Params(dt, theta), packs, Tinf=20.0))
        finally:
            g.set_option("fuse", 0)
        assert cases.rel_l2(out, ref, mask) <= TOL
        assert np.array_equal(out[~mask], T0[~mask], equal_nan=True)


def test_pipelined_host_fields(g, golden_dir):
    """adi_cart_step_host_async: five independent host fields through two staging slots / two streams."""
    c = cases.build_cart_case("holes_combined")
    nx, ny, nz = c["shape"]
    grid = g.Grid3D(nx, ny, nz, c["dx"], c["mask"])
    mat = g.Material(c["rho"], c["cp"], c["k"])
    packs = g.precompute_coeff_packs_unified(grid, mat, **c["bcs"])
    prm = g.Params(c["dt"], c["theta"])
    fields = []
    for i in range(5):
        f = c["T0"].copy()
        f[c["mask"]] += 3.0 * i
        fields.append(f)
    outs = g.adi_step_host_pipelined(fields, grid, mat, prm, packs, Tinf=c["Tinf"])
    for f, o in zip(fields, outs):
        ref = g.adi_step_host(f, grid, mat, prm, packs, Tinf=c["Tinf"])
        assert np.array_equal(o, ref, equal_nan=True)


# ---- surface-only coefficient fields: the sweeps read them at exposed cells only -----------------
@pytest.mark.parametrize("shape", [(40, 67, 130), (96, 33, 64), (600, 7, 48), (5, 1024, 24), (33, 18, 37), (20, 20, 515)])
def test_sparse_coefficient_reads_agree(shape, g, cp):
    """Packs from precompute_coeff_packs_unified vanish away from the surface (adi3d_numba_coeff.py:93-99);
    the engine verifies that (k_check_sparse) and skips the interior coefficient reads.  Same bits as
    with the dense reads (option sparse_coeff=0), and parity with the oracle either way."""
    c = _oracle_case(shape, seed=7000 + shape[0])
    nx, ny, nz = shape
    gd = g.Grid3D(nx, ny, nz, cases.DX, c["mask"])
    md = g.Material(cases.RHO, cases.CP, cases.K)
    pd = g.precompute_coeff_packs_unified(gd, md, **c["bcs"])
    outs = []
    try:
        for sparse in (1, 0):
            g.set_option("sparse_coeff", sparse)
            T = cp.asarray(c["T0"])
            for _ in range(2):
                T = g.adi_step_gpu_coeff(T, gd, md, g.Params(c["dt"], c["theta"]), pd, Tinf=20.0)
            outs.append(cp.asnumpy(T))
    finally:
        g.set_option("sparse_coeff", 1)
    # same rows either way; the sweeps that skip the interior coefficient reads may also take the tabulated
    # uniform-run factors, so the two agree to rounding, not bit for bit
    assert cases.rel_l2(outs[0], outs[1], c["mask"]) <= 1e-13
    assert np.array_equal(outs[0][~c["mask"]], outs[1][~c["mask"]], equal_nan=True)
    _both(g, cp, c)


def test_coefficient_fields_dense_on_one_axis_only(g, cp):
    """A caller-made x pack with non-zero coefficients in the interior next to surface-only y / z packs:
    the x sweep must keep its dense reads."""
    from oracle import cart
    shape = (37, 29, 48)
    nx, ny, nz = shape
    c = _oracle_case(shape, seed=7100)
    gh = cart.Grid3D(nx, ny, nz, cases.DX, c["mask"])
    mh = cart.Material(cases.RHO, cases.CP, cases.K)
    ph = list(cart.precompute_coeff_packs_unified(gh, mh, **c["bcs"]))
    dense = 3.0 * cases.splitmix_uniform(7101, shape)
    ph[0] = cart.AxisCoeffPack(dense, ph[0].dir_mask, ph[0].dir_val, ph[0].qflux)
    gd = g.Grid3D(nx, ny, nz, cases.DX, c["mask"])
    md = g.Material(cases.RHO, cases.CP, cases.K)
    pd = list(g.precompute_coeff_packs_unified(gd, md, **c["bcs"]))
    pd[0] = g.AxisCoeffPack(cp.asarray(dense), pd[0].dir_mask, pd[0].dir_val, pd[0].qflux)
    Th = cart.adi_step_numba_coeff(c["T0"], gh, mh, cart.Params(c["dt"], 0.5), ph, Tinf=20.0)
    Td = cp.asnumpy(g.adi_step_gpu_coeff(cp.asarray(c["T0"]), gd, md, g.Params(c["dt"], 0.5), pd, Tinf=20.0))
    assert cases.rel_l2(Td, Th, c["mask"]) <= TOL


def test_coefficient_edited_in_place_is_reexamined(g, cp):
    """After a step, the caller writes a coefficient into an interior cell of the bound z pack: the
    next step must use it (the pack is re-bound and re-examined, not assumed surface-only)."""
    from oracle import cart
    shape = (24, 25, 40)
    nx, ny, nz = shape
    c = _oracle_case(shape, seed=7200, holes=False)
    c["bcs"] = dict(robin_h=c["bcs"]["robin_h"])
    gh = cart.Grid3D(nx, ny, nz, cases.DX, c["mask"])
    mh = cart.Material(cases.RHO, cases.CP, cases.K)
    ph = cart.precompute_coeff_packs_unified(gh, mh, **c["bcs"])
    gd = g.Grid3D(nx, ny, nz, cases.DX, c["mask"])
    md = g.Material(cases.RHO, cases.CP, cases.K)
    pd = g.precompute_coeff_packs_unified(gd, md, **c["bcs"])
    prm_h, prm_d = cart.Params(c["dt"], 0.5), g.Params(c["dt"], 0.5)
    Th = cart.adi_step_numba_coeff(c["T0"], gh, mh, prm_h, ph, Tinf=20.0)
    Td = g.adi_step_gpu_coeff(cp.asarray(c["T0"]), gd, md, prm_d, pd, Tinf=20.0)
    ph[2].coeff[10:14, 11, 17:23] = 7.5
    pd[2].coeff[10:14, 11, 17:23] = 7.5
    Th = cart.adi_step_numba_coeff(Th, gh, mh, prm_h, ph, Tinf=20.0)
    Td = g.adi_step_gpu_coeff(Td, gd, md, prm_d, pd, Tinf=20.0)
    assert cases.rel_l2(cp.asnumpy(Td), Th, c["mask"]) <= TOL
    assert abs(cp.asnumpy(Td)[12, 11, 20] - Th[12, 11, 20]) <= 1e-9 * abs(Th[12, 11, 20])


# ---- second-generation x / y sweeps: uniform chunks, transposed codes, warp-level reduced solve ------------
def _uniform_case(shape, mask_kind, bk, theta, cfl, seed):
    mask = cases.make_mask(mask_kind, shape, seed)
    bcs = cases.make_bcs(bk, shape, mask, seed, 20.0)
    T0 = 20.0 + 1380.0 * cases.splitmix_uniform(seed + 1, shape)
    T0[~mask] = np.nan
    kappa = cases.K / (cases.RHO * cases.CP)
    return dict(shape=shape, mask=mask, T0=T0, dt=cfl * cases.DX ** 2 / kappa, theta=theta, bcs=bcs, kappa=kappa)


@pytest.mark.parametrize("opts", [dict(), dict(uni=0), dict(tw=1), dict(xy2=0), dict(m=16), dict(m=32), dict(kt=4), dict(m=16, kt=2),
                                  dict(m=16, occ=3), dict(m=16, occ=4, tw=1), dict(remap=1), dict(remap=1, tw=1), dict(wide=1),
                                  dict(m=16, wide=1, tw=1), dict(lt=1), dict(lt=4), dict(sparse_coeff=0), dict(zt=0), dict(zt=0, uni=0), dict(bulk=0), dict(bulk=0, uni=0), dict(tiles=0), dict(hyb=0), dict(xyp=1)],
                         ids=lambda o: "-".join(f"{k}{v}" for k, v in o.items()) or "default")
@pytest.mark.parametrize("shape,mask_kind", [((70, 40, 37), "full"), ((40, 70, 130), "plate_track"), ((96, 50, 64), "cyl_holes"),
                                             ((600, 7, 48), "full"), ((5, 1100, 24), "full"), ((2050, 3, 10), "full"),
                                             ((1030, 2, 9), "random"), ((6, 5, 1000), "full"), ((3, 4, 1500), "full"),
                                             ((20, 9, 515), "plate_track")],
                         ids=lambda v: "x".join(map(str, v)) if isinstance(v, tuple) else v)
def test_uniform_chunk_paths(shape, mask_kind, opts, g, cp):
    """Chunks whose cells all have both neighbours along the swept axis take tabulated elimination factors
    (adi_core.h UniConst); every launch shape / option gives the oracle's answer: dense per-face h and scalar h,
    theta 0.5 and 1, cfl from 0.128 to 3000 (slowly decaying couplings)."""
    restore = {k: int(g.get_option(k)) for k in opts}
    for k, v in opts.items():
        g.set_option(k, v)
    try:
        for bk, theta, cfl in [("robin_dict3d", 0.5, 0.128), ("robin6", 0.5, 3000.0), ("robin_dict3d", 1.0, 40.0)]:
            _both(g, cp, _uniform_case(shape, mask_kind, bk, theta, cfl, seed=8000 + shape[0]), nsteps=2)
    finally:
        for k, v in restore.items():
            g.set_option(k, v)


def test_active_tile_lists(g, cp):
    """A part under construction: most sweep tiles hold no active cell and are not launched at all (option tiles);
    births move the boundary.  Same answer as the oracle, void cells untouched, and the lists really are short."""
    from oracle import cart
    shape = (160, 144, 200)
    nx, ny, nz = shape
    ax = [(np.arange(n) + 0.5) / n - 0.5 for n in shape]
    X, Y, Z = np.meshgrid(*ax, indexing="ij")
    full = (X / 0.3) ** 2 + (Y / 0.35) ** 2 + (Z / 0.4) ** 2 <= 1.0
    h = {f: 40.0 * (0.3 + cases.splitmix_uniform(90 + i, shape)) for i, f in enumerate(cart.FACES)}
    kappa = cases.K / (cases.RHO * cases.CP)
    dt = 50.0 * cases.DX ** 2 / kappa
    act = np.zeros(shape, bool)
    Th = np.full(shape, 20.0)
    Td = cp.asarray(Th)
    grid = g.Grid3D(nx, ny, nz, cases.DX, act)
    mat, hm = g.Material(cases.RHO, cases.CP, cases.K), cart.Material(cases.RHO, cases.CP, cases.K)
    hd = {f: cp.asarray(v) for f, v in h.items()}
    for opt in (1, 0):
        g.set_option("tiles", opt)
        try:
            for k0 in (20, 36, 52):
                born = np.zeros(shape, bool)
                born[:, :, k0:k0 + 16] = full[:, :, k0:k0 + 16] & ~act[:, :, k0:k0 + 16]
                act |= born
                Th[born] = 1000.0
                Td[cp.asarray(born)] = 1000.0
                grid.mask = act.copy()
                pd = g.precompute_coeff_packs_unified(grid, mat, robin_h=hd)
                hg = cart.Grid3D(nx, ny, nz, cases.DX, act)
                ph = cart.precompute_coeff_packs_unified(hg, hm, robin_h=h)
                for _ in range(2):
                    Th = cart.adi_step_numba_coeff(Th, hg, hm, cart.Params(dt, 0.5), ph, Tinf=20.0)
                    Td = g.adi_step_gpu_coeff(Td, grid, mat, g.Params(dt, 0.5), pd, Tinf=20.0)
                out = cp.asnumpy(Td)
                assert cases.rel_l2(out, Th, act) <= 6 * TOL
                assert np.array_equal(out[~act], Th[~act])
                if opt:
                    assert 0 < g.get_option("tiles_active") < 0.5 * g.get_option("tiles_total")
        finally:
            g.set_option("tiles", 1)
        act[:] = False
        Th[:] = 20.0
        Td = cp.asarray(Th)


@pytest.mark.parametrize("opts", [dict(), dict(hyb=0), dict(lt=2), dict(lt=16), dict(m=16), dict(bulk=0)],
                         ids=lambda o: "-".join(f"{k}{v}" for k, v in o.items()) or "default")
@pytest.mark.parametrize("shape", [(24, 40, 256), (16, 12, 515), (9, 33, 160), (12, 20, 96), (10, 17, 128), (7, 9, 64)],
                         ids=lambda s: "x".join(map(str, s)))
def test_z_sweep_surface_chunks(shape, opts, g, cp):
    """z lines that cross the top surface of a part: the chunk under the surface has a uniform lead and a general tail
    (adi_core.h chunk_forward_hybrid).  Terraced surface heights (every lead length occurs, different ones inside one
    warp), a floating block above a gap, NaN in the void; dense per-face h, scalar h, and a Neumann + Dirichlet set
    (which keeps the general rows) -- against the oracle; void cells keep their bits."""
    nx, ny, nz = shape
    k = np.arange(nz)[None, None, :]
    height = (nz // 3 + (7 * np.arange(nx)[:, None] + 3 * (np.arange(ny)[None, :] // 5)) % (nz - nz // 3 - 1))[:, :, None]
    mask = k < height
    mask |= (k >= height + 3) & (k < height + 9) & ((np.arange(nx) % 4 == 1)[:, None, None])   # floating blocks
    mask[0, :, :] = True
    mask[:, 1, 5:11] = False                                                                   # a gap low in the line
    restore = {kk: int(g.get_option(kk)) for kk in opts}
    for kk, v in opts.items():
        g.set_option(kk, v)
    try:
        for bk, theta, cfl in [("robin_dict3d", 0.5, 0.9), ("robin6", 1.0, 300.0), ("combined", 0.5, 5.0)]:
            bcs = cases.make_bcs(bk, shape, mask, 31, 20.0)
            T0 = 20.0 + 1380.0 * cases.splitmix_uniform(32, shape)
            T0[~mask] = np.nan
            kappa = cases.K / (cases.RHO * cases.CP)
            _both(g, cp, dict(shape=shape, mask=mask, T0=T0, dt=cfl * cases.DX ** 2 / kappa, theta=theta, bcs=bcs,
                              kappa=kappa), nsteps=2)
    finally:
        for kk, v in restore.items():
            g.set_option(kk, v)


@pytest.mark.parametrize("opts", [dict(), dict(ukt=16), dict(ukt=8), dict(xyu=0), dict(xyp=1), dict(xyp=1, tiles=0), dict(tiles=0)],
                         ids=lambda o: "-".join(f"{k}{v}" for k, v in o.items()) or "default")
@pytest.mark.parametrize("shape,mask_kind", [((1152, 40, 38), "plate_track"), ((40, 1152, 38), "cyl_holes"), ((2048, 21, 20), "full"),
                                             ((6, 2048, 70), "plate_track"), ((1536, 30, 10), "random"), ((1100, 7, 9), "full")],
                         ids=lambda v: "x".join(map(str, v)) if isinstance(v, tuple) else v)
def test_long_lines_persistent_blocks(shape, mask_kind, opts, g, cp):
    """x / y lines of 1025..2048 cells: the all-uniform tiles run on k_sweep_xyu (half of every chunk in shared
    memory, two blocks per SM), the other active tiles on k_sweep_xy; option xyp: persistent blocks whose tiles arrive
    as TMA tensor copies (adi_sweep_xyp.cuh).  More tiles than blocks, uniform and general warps mixed, ragged z
    tiles, void tiles (in place: skipped, with and without the tile lists), dense per-face h and scalar h."""
    restore = {k: int(g.get_option(k)) for k in opts}
    for k, v in opts.items():
        g.set_option(k, v)
    try:
        used0 = g.get_option("xyp_used")
        uni0 = g.get_option("xyu_used")
        for bk, theta, cfl in [("robin_dict3d", 0.5, 0.7), ("robin6", 1.0, 500.0)]:
            _both(g, cp, _uniform_case(shape, mask_kind, bk, theta, cfl, seed=9000 + shape[0]), nsteps=2)
        n_long = max(shape[0], shape[1])
        if opts.get("xyu", 1) and not opts.get("xyp") and opts.get("tiles", 1) and n_long % 32 == 0 and mask_kind in ("full", "plate_track"):
            assert g.get_option("xyu_used") > uni0, "the all-uniform tiles did not go to k_sweep_xyu"
        if opts.get("xyu") == 0 or opts.get("tiles") == 0:
            assert g.get_option("xyu_used") == uni0
        if opts.get("xyp") == 1 and n_long % 128 == 0 and shape[2] % 2 == 0:
            assert g.get_option("xyp_layout") >= 0, "the driver refused the tensor map"
            assert g.get_option("xyp_used") > used0, "the persistent kernel did not run"
    finally:
        for k, v in restore.items():
            g.set_option(k, v)


@pytest.mark.parametrize("shape", [(24, 27, 32), (19, 38, 64), (40, 132, 48), (130, 20, 144)])
@pytest.mark.parametrize("mk,bk", [("full", "robin_dict3d"), ("cyl_holes", "combined"), ("thin", "dir_scalar_interior"),
                                   ("random", "combined"), ("plate_track", "robin_mixed"), ("empty", "robin6")])
def test_mask_kernels_word_forms(shape, mk, bk, g, cp):
    """The per-mask-change kernels in word form (adi_mask_core.h: neighbour code by 16 cells, transposed codes by
    128 x 128 byte tiles, pack builder by 4 cells; nz % 16 == 0 here so that all three run) against the oracle and,
    bit for bit, against the one-cell-per-thread forms (option maskv=0)."""
    from oracle import cart
    nx, ny, nz = shape
    seed = 4242 + nx
    mask = cases.make_mask(mk, shape, seed)
    bcs = cases.make_bcs(bk, shape, mask, seed, 20.0)
    T0 = 20.0 + 1380.0 * cases.splitmix_uniform(seed + 1, shape)
    T0[~mask] = np.nan
    kappa = cases.K / (cases.RHO * cases.CP)
    dt = 2.0 * cases.DX ** 2 / kappa
    hg, hm = cart.Grid3D(nx, ny, nz, cases.DX, mask), cart.Material(cases.RHO, cases.CP, cases.K)
    hp = cart.precompute_coeff_packs_unified(hg, hm, **bcs)
    ref = cart.adi_step_numba_coeff(T0, hg, hm, cart.Params(dt, 0.5), hp, Tinf=20.0)
    outs = {}
    try:
        for v in (1, 0):
            g.set_option("maskv", v)
            grid = g.Grid3D(nx, ny, nz, cases.DX, mask)
            mat = g.Material(cases.RHO, cases.CP, cases.K)
            packs = g.precompute_coeff_packs_unified(grid, mat, **bcs)
            dense = packs[0]._coeff is not None
            out = cp.asnumpy(g.adi_step_gpu_coeff(cp.asarray(T0), grid, mat, g.Params(dt, 0.5), packs, Tinf=20.0))
            used = g.get_option("maskv_used")
            assert (used & 3) == (3 if v else 0), used
            if dense:
                assert (used & 4) == (4 if v else 0), used
            for a in range(3):
                assert np.array_equal(cp.asnumpy(packs[a].coeff).view(np.uint64), hp[a].coeff.view(np.uint64))
                assert np.array_equal(cp.asnumpy(packs[a].qflux).view(np.uint64), hp[a].qflux.view(np.uint64))
            outs[v] = out
    finally:
        g.set_option("maskv", 1)
    assert cases.rel_l2(outs[1], ref, mask) <= TOL
    assert np.array_equal(outs[1][~mask], T0[~mask], equal_nan=True)
    assert np.array_equal(outs[1].view(np.uint64), outs[0].view(np.uint64))


@pytest.mark.parametrize("top", [0, 1, 150, 289, 500])
@pytest.mark.parametrize("bk", ["robin_dict3d", "robin6", "combined", "neumann_fields"])
def test_z_sweep_stops_at_the_top_of_the_part(top, bk, g, cp):
    """A part under construction (layers born bottom-up along z): the z sweep solves only the cells below the top of the
    part (option ztrim; the top comes out of the code build) -- same answer as the oracle and as the untrimmed sweep,
    void cells (NaN) untouched."""
    from oracle import cart
    shape = (20, 24, 512)
    nx, ny, nz = shape
    seed = 777 + top
    mask = cases.make_mask("cyl_holes", shape, seed)
    mask[:, :, top:] = False
    if top > 40:
        mask[3:9, 5:7, top - 40:top - 20] = False     # a cavity: interior exposed faces
    bcs = cases.make_bcs(bk, shape, mask, seed, 20.0)
    T0 = 20.0 + 1380.0 * cases.splitmix_uniform(seed + 1, shape)
    T0[~mask] = np.nan
    kappa = cases.K / (cases.RHO * cases.CP)
    dt = 50.0 * cases.DX ** 2 / kappa
    hg, hm = cart.Grid3D(nx, ny, nz, cases.DX, mask), cart.Material(cases.RHO, cases.CP, cases.K)
    ref = cart.adi_step_numba_coeff(T0, hg, hm, cart.Params(dt, 0.5),
                                    cart.precompute_coeff_packs_unified(hg, hm, **bcs), Tinf=20.0)
    outs = {}
    try:
        for v in (1, 0):
            g.set_option("ztrim", v)
            grid = g.Grid3D(nx, ny, nz, cases.DX, mask)
            mat = g.Material(cases.RHO, cases.CP, cases.K)
            packs = g.precompute_coeff_packs_unified(grid, mat, **bcs)
            used0 = g.get_option("ztrim_used")
            out = g.adi_step_gpu_coeff(cp.asarray(T0), grid, mat, g.Params(dt, 0.5), packs, Tinf=20.0)
            out = cp.asnumpy(g.adi_step_gpu_coeff(out, grid, mat, g.Params(dt, 0.5), packs, Tinf=20.0))
            if v:
                assert g.get_option("ztop") == top
            trimmed = g.get_option("ztrim_used") - used0
            # the trimmed length is ztop rounded up to 32-cell chunks, at least 256: 500 -> 512 = nz, nothing to trim
            if not v or top > 480:
                assert trimmed == 0, trimmed
            elif top > 0 and bk in ("robin_dict3d", "robin6"):    # (operand sets that stay on k_sweep_z are not trimmed)
                assert trimmed == 2, trimmed
            outs[v] = out
    finally:
        g.set_option("ztrim", 1)
    ref = cart.adi_step_numba_coeff(ref, hg, hm, cart.Params(dt, 0.5),
                                    cart.precompute_coeff_packs_unified(hg, hm, **bcs), Tinf=20.0)
    for v in (1, 0):
        assert cases.rel_l2(outs[v], ref, mask) <= 2 * TOL
        assert np.array_equal(outs[v][~mask], T0[~mask], equal_nan=True)
    assert cases.rel_l2(outs[1], outs[0], mask) <= 1e-14


@pytest.mark.parametrize("xr", [(100, 300), (40, 500), (0, 10), (505, 512), (0, 512)])
@pytest.mark.parametrize("shape,bk", [((512, 20, 64), "robin_dict3d"), ((512, 20, 64), "combined"), ((512, 20, 64), "robin6"),
                                      ((512, 6, 512), "robin_dict3d")])
def test_x_sweep_runs_on_the_x_extent_of_the_part(xr, shape, bk, g, cp):
    """The in-place x sweep covers only the x planes that hold an active cell (whole 32-cell chunks, at least 256
    planes; the extent comes out of the code build) -- same answer as the oracle and as the full sweep."""
    from oracle import cart
    nx, ny, nz = shape
    seed = 900 + xr[0]
    mask = cases.make_mask("random", shape, seed)
    mask[:xr[0]] = False
    mask[xr[1]:] = False
    if nz == 512:
        mask[:, :, 300:] = False                      # z and x trimmed together
    bcs = cases.make_bcs(bk, shape, mask, seed, 20.0)
    T0 = 20.0 + 1380.0 * cases.splitmix_uniform(seed + 1, shape)
    T0[~mask] = np.nan
    kappa = cases.K / (cases.RHO * cases.CP)
    dt = 50.0 * cases.DX ** 2 / kappa
    hg, hm = cart.Grid3D(nx, ny, nz, cases.DX, mask), cart.Material(cases.RHO, cases.CP, cases.K)
    hp = cart.precompute_coeff_packs_unified(hg, hm, **bcs)
    ref = cart.adi_step_numba_coeff(T0, hg, hm, cart.Params(dt, 0.5), hp, Tinf=20.0)
    ref = cart.adi_step_numba_coeff(ref, hg, hm, cart.Params(dt, 0.5), hp, Tinf=20.0)
    planes = np.nonzero(mask.any(axis=(1, 2)))[0]
    outs = {}
    try:
        for v in (1, 0):
            g.set_option("ztrim", v)
            grid = g.Grid3D(nx, ny, nz, cases.DX, mask)
            mat = g.Material(cases.RHO, cases.CP, cases.K)
            packs = g.precompute_coeff_packs_unified(grid, mat, **bcs)
            used0 = g.get_option("xtrim_used")
            out = g.adi_step_gpu_coeff(cp.asarray(T0), grid, mat, g.Params(dt, 0.5), packs, Tinf=20.0)
            outs[v] = cp.asnumpy(g.adi_step_gpu_coeff(out, grid, mat, g.Params(dt, 0.5), packs, Tinf=20.0))
            trimmed = g.get_option("xtrim_used") - used0
            if v and planes.size:
                assert (g.get_option("xlo"), g.get_option("xhi")) == (int(planes[0]), int(planes[-1]) + 1)
                lo = planes[0] // 32 * 32
                hi = min(nx, (planes[-1] + 32) // 32 * 32)
                assert trimmed == (2 if max(hi - lo, 256) < nx else 0), trimmed
            if not v:
                assert trimmed == 0
    finally:
        g.set_option("ztrim", 1)
    for v in (1, 0):
        assert cases.rel_l2(outs[v], ref, mask) <= 2 * TOL
        assert np.array_equal(outs[v][~mask], T0[~mask], equal_nan=True)
    assert cases.rel_l2(outs[1], outs[0], mask) <= 1e-14
