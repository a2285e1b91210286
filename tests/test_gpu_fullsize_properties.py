# -*- coding: utf-8 -*-
"""Full-size checks (BASELINE sizes: Cartesian 512^3, cylindrical 256 x 1024 x 512) through size-independent
properties of the step -- the oracle cannot run these sizes in seconds, so the CUDA path is checked against
itself in ways a wrong kernel would break:
  linearity        step(a*T1 + b*T2) = a*step(T1) + b*step(T2)            (Tinf = 0, no flux: the step is linear)
  constants        a uniform field with no Robin / flux / Dirichlet is a fixed point (rows sum to one)
  mirror symmetry  stepping the x- (y-, z-) mirrored problem gives the mirrored result
  rotation         the cylindrical step commutes with a shift in phi (coefficients do not depend on phi)
  slab consistency two z-slabs (halo planes + interface exchange) reproduce the undivided step
  oracle slab      a 512 x 24 x 512 sub-problem (full-length x and z lines) against the oracle
Tolerances: 1e-12 relative L2 (north_star), void cells bit-identical."""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TOL = 1e-12
N = 512
RHO, CP, K, DX, DT = 7800.0, 500.0, 25.0, 1.0e-3, 0.02


def rel(a, b, m=None):
    import torch
    if m is not None:
        a, b = a[m], b[m]
    return float(torch.linalg.vector_norm(a - b) / torch.linalg.vector_norm(b))


@pytest.fixture(scope="module")
def plate():
    """single_track_on_plate at 512^3: plate + track mask, dense per-face h (bench.py's workload)."""
    import torch
    from adi_thermal_fields_b200 import adi3d_gpu_coeff as g, devarray as cp
    dev = torch.device("cuda", 0)
    mask = torch.zeros((N, N, N), dtype=torch.bool, device=dev)
    nzp = N - N // 64
    mask[:, :, :nzp] = True
    mask[: N // 32, : N // 2, nzp:] = True
    grid = g.Grid3D.__new__(g.Grid3D)
    grid.nx = grid.ny = grid.nz = N
    grid.dx = DX
    grid.mask = cp.ndarray(mask)
    mat = g.Material(RHO, CP, K)
    gen = torch.Generator(device=dev).manual_seed(7)
    h = {f: cp.ndarray(10.0 * (0.3 + torch.rand((N, N, N), dtype=torch.float64, device=dev, generator=gen))) for f in g.FACES}
    packs = g.precompute_coeff_packs_unified(grid, mat, robin_h=h)
    del h
    T1 = 20.0 + 1380.0 * torch.rand((N, N, N), dtype=torch.float64, device=dev, generator=gen)
    T2 = 500.0 * torch.rand((N, N, N), dtype=torch.float64, device=dev, generator=gen)
    yield dict(g=g, cp=cp, grid=grid, mat=mat, packs=packs, mask=mask, T1=T1, T2=T2, dev=dev)
    torch.cuda.empty_cache()


def step(p, T, Tinf=0.0, theta=0.5, packs=None, grid=None):
    g, cp = p["g"], p["cp"]
    return g.adi_step_gpu_coeff(cp.ndarray(T), grid or p["grid"], p["mat"], g.Params(DT, theta), packs or p["packs"], Tinf=Tinf)._t


def test_cartesian_linearity_512(plate):
    p = plate
    a, b = 0.75, -1.25
    lhs = step(p, a * p["T1"] + b * p["T2"])
    rhs = a * step(p, p["T1"]) + b * step(p, p["T2"])
    assert rel(lhs, rhs, p["mask"]) <= TOL
    assert bool((lhs[~p["mask"]] == (a * p["T1"] + b * p["T2"])[~p["mask"]]).all())   # void cells untouched


def test_cartesian_constant_is_fixed_point_512(plate):
    import torch
    p = plate
    g = p["g"]
    packs0 = g.precompute_coeff_packs_unified(p["grid"], p["mat"])       # no Robin, no flux, no Dirichlet
    T = torch.full((N, N, N), 345.678, dtype=torch.float64, device=p["dev"])
    out = step(p, T, packs=packs0)
    assert float((out - T).abs().max()) <= 1e-10 * 345.678
    # with Robin cooling towards the field's own value it stays a fixed point too
    out = step(p, T, Tinf=345.678)
    assert float((out - T).abs().max()) <= 1e-10 * 345.678


@pytest.mark.parametrize("axis", [0, 1, 2])
def test_cartesian_mirror_symmetry_512(plate, axis):
    import torch
    p = plate
    g, cp = p["g"], p["cp"]
    ref = step(p, p["T1"], Tinf=20.0)
    # mirrored problem: mask, fields and the two faces of the mirrored axis swap
    gm = g.Grid3D.__new__(g.Grid3D)
    gm.nx = gm.ny = gm.nz = N
    gm.dx = DX
    gm.mask = cp.ndarray(torch.flip(p["mask"], dims=(axis,)).contiguous())
    pk = []
    for a, pa in enumerate(p["packs"]):
        pk.append(g.AxisCoeffPack(cp.ndarray(torch.flip(pa.coeff._t, dims=(axis,)).contiguous()), pa.dir_mask, pa.dir_val))
    out = step(p, torch.flip(p["T1"], dims=(axis,)).contiguous(), Tinf=20.0, packs=pk, grid=gm)
    assert rel(torch.flip(out, dims=(axis,)), ref, p["mask"]) <= TOL


def test_cartesian_two_slabs_match_undivided_512(plate):
    import torch
    from adi_thermal_fields_b200 import slab
    p = plate
    ref = step(p, p["T1"], Tinf=20.0)
    half = N // 2

    class Mat:
        rho, cp, k = RHO, CP, K

    class Prm:
        dt, theta = DT, 0.5

    def rank_fn(v):
        z0, z1 = v.rank * half, (v.rank + 1) * half
        grid = slab.SlabGrid3D(N, N, half, DX, p["mask"][:, :, z0:z1].contiguous(), v)
        packs = slab.SlabPacks([(pa.coeff._t[:, :, z0:z1].contiguous(), None, None, None) for pa in p["packs"]], None)
        T = p["T1"][:, :, z0:z1].contiguous()
        return z0, z1, slab.adi_step_gpu_coeff(T, grid, Mat, Prm, packs, Tinf=20.0)

    parts = slab.LocalComm(2).run(rank_fn)
    out = torch.empty_like(ref)
    for z0, z1, t in parts:
        out[:, :, z0:z1] = t
    assert rel(out, ref, p["mask"]) <= TOL
    assert bool((out[~p["mask"]] == p["T1"][~p["mask"]]).all())


def test_cartesian_full_length_lines_against_oracle(plate):
    """512 x 24 x 512 sub-problem (x and z lines at full BASELINE length) against the oracle."""
    import torch
    from oracle import cart
    p = plate
    g, cp = p["g"], p["cp"]
    ny = 24
    mask = p["mask"][:, :ny, :].contiguous()
    T0 = p["T1"][:, :ny, :].contiguous()
    coeff = [pa.coeff._t[:, :ny, :].contiguous() for pa in p["packs"]]
    grid = g.Grid3D(N, ny, N, DX, cp.ndarray(mask))
    dm = torch.zeros_like(mask)
    packs = [g.AxisCoeffPack(cp.ndarray(c), cp.ndarray(dm), cp.ndarray(torch.zeros_like(T0))) for c in coeff]
    out = g.adi_step_gpu_coeff(cp.ndarray(T0), grid, p["mat"], g.Params(DT, 0.5), packs, Tinf=20.0)._t.cpu().numpy()
    hm = mask.cpu().numpy()
    hp = [cart.AxisCoeffPack(c.cpu().numpy(), np.zeros(hm.shape, bool), np.zeros(hm.shape)) for c in coeff]
    cart.set_threads(max(1, cart.max_threads()))
    ref = cart.adi_step_numba_coeff(T0.cpu().numpy(), cart.Grid3D(N, ny, N, DX, hm), cart.Material(RHO, CP, K),
                                    cart.Params(DT, 0.5), hp, Tinf=20.0)
    num = np.sqrt(((out - ref)[hm] ** 2).sum())
    den = np.sqrt((ref[hm] ** 2).sum())
    assert num / den <= TOL
    assert np.array_equal(out[~hm], T0.cpu().numpy()[~hm])


@pytest.fixture(scope="module")
def cylgrid():
    import torch
    from adi_thermal_fields_b200 import adi3d_cyl_phi_v3 as gc
    nr, nphi, nz = 256, 1024, 512
    R = 0.02
    dr = R / nr
    grid = gc.GridCyl(nr, nphi, nz, dr, 2 * math.pi / nphi, dr, R)
    mat = gc.Material(7800.0, 490.0, 54.0)
    dt = min(dr * dr, (R * grid.dphi) ** 2) / mat.alpha
    gen = torch.Generator(device="cuda").manual_seed(11)
    T = 20.0 + 980.0 * torch.rand((nr, nphi, nz), dtype=torch.float64, device="cuda", generator=gen)
    yield dict(gc=gc, grid=grid, mat=mat, prm=gc.Params(dt, 1.0, "be"), T=T)
    torch.cuda.empty_cache()


def test_cylindrical_phi_shift_invariance_full_size(cylgrid):
    import torch
    c = cylgrid
    gc = c["gc"]
    rob, zbc = gc.RobinR(500.0, 20.0), gc.ZBC("neumann0", "robin", h_top=500.0, T_inf_top=20.0)
    ref = gc.adi_step_device(c["T"], c["grid"], c["mat"], c["prm"], rob, zbc)
    for s in (1, 37, 512):
        out = gc.adi_step_device(torch.roll(c["T"], s, dims=1).contiguous(), c["grid"], c["mat"], c["prm"], rob, zbc)
        assert rel(torch.roll(out, -s, dims=1), ref) <= TOL


def test_cylindrical_linearity_and_constants_full_size(cylgrid):
    import torch
    c = cylgrid
    gc = c["gc"]
    rob0, zbc0 = gc.RobinR(0.0, 0.0), gc.ZBC("neumann0", "neumann0")
    T2 = torch.flip(c["T"], dims=(2,)).contiguous() * 0.5
    a, b = -0.6, 1.7
    lhs = gc.adi_step_device(a * c["T"] + b * T2, c["grid"], c["mat"], c["prm"], rob0, zbc0)
    rhs = a * gc.adi_step_device(c["T"], c["grid"], c["mat"], c["prm"], rob0, zbc0) + \
        b * gc.adi_step_device(T2, c["grid"], c["mat"], c["prm"], rob0, zbc0)
    assert rel(lhs, rhs) <= TOL
    const = torch.full_like(c["T"], 123.456)
    out = gc.adi_step_device(const, c["grid"], c["mat"], c["prm"], rob0, zbc0)     # insulated: a fixed point
    assert float((out - const).abs().max()) <= 1e-9 * 123.456
    rob, zbc = gc.RobinR(500.0, 123.456), gc.ZBC("robin", "robin", h_bot=80.0, h_top=500.0, T_inf_bot=123.456, T_inf_top=123.456)
    out = gc.adi_step_device(const, c["grid"], c["mat"], c["prm"], rob, zbc)       # ambient = field: fixed point
    assert float((out - const).abs().max()) <= 1e-9 * 123.456


def test_cylindrical_two_slabs_match_undivided_full_size(cylgrid):
    import torch
    from adi_thermal_fields_b200 import slab
    c = cylgrid
    gc = c["gc"]
    rob, zbc = gc.RobinR(500.0, 20.0), gc.ZBC("neumann0", "robin", h_top=500.0, T_inf_top=20.0)
    ref = gc.adi_step_device(c["T"], c["grid"], c["mat"], c["prm"], rob, zbc)
    g = c["grid"]
    half = g.nz // 2

    def rank_fn(v):
        z0, z1 = v.rank * half, (v.rank + 1) * half
        sg = slab.SlabGridCyl(g.nr, g.nphi, half, g.dr, g.dphi, g.dz, g.R, v, nz_per_rank=[half, half])
        return z0, z1, slab.adi_step_cyl(c["T"][:, :, z0:z1].contiguous(), sg, c["mat"], c["prm"], rob, zbc)

    parts = slab.LocalComm(2).run(rank_fn)
    out = torch.empty_like(ref)
    for z0, z1, t in parts:
        out[:, :, z0:z1] = t
    assert rel(out, ref) <= TOL
