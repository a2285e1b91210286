# -*- coding: utf-8 -*-
"""Deterministic input builders shared by the golden-vector generator
(tools/gen_golden.py, run in the build container where /root/reference is
importable) and by the parity tests (which run where it is NOT).

Inputs are produced by a splitmix64 hash of the cell index, i.e. by integer
arithmetic only, so they are bit-identical on every NumPy version; only the
reference OUTPUTS are stored under tests/golden/.

The case matrix follows SURVEY.md section 8(c): masks {full, voxel cylinder,
cylinder with holes, plate+track, thin features}, BC sets {Robin scalar,
Robin dict of 3-D arrays, Neumann, Dirichlet top, Dirichlet both, combined},
theta in {0.5, 1}, cfl in {0.128, 2, 3000}, Tinf in {0, 20}, NaN-poisoned
void cells.
"""
from __future__ import annotations

import math

import numpy as np

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix_uniform(seed: int, shape) -> np.ndarray:
    """U[0,1) doubles from splitmix64(seed*2^32 + flat index); integer math only."""
    n = int(np.prod(shape))
    with np.errstate(over="ignore"):
        z = (np.arange(n, dtype=np.uint64) + np.uint64(seed) * np.uint64(0x100000000)
             + np.uint64(0x9E3779B97F4A7C15))
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    u = (z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
    return u.reshape(shape)


# ----------------------------------------------------------------------------
# Cartesian cases (oracle: adi3d_numba_coeff.adi_step_numba_coeff)
# ----------------------------------------------------------------------------
RHO, CP, K = 7800.0, 500.0, 25.0
DX = 1.0e-3
SHAPE_A = (24, 27, 29)   # non-cubic: catches axis mix-ups; lines span 2 chunks of 16
SHAPE_B = (19, 38, 70)   # longer lines: 3 and 5 chunks, partial last chunk


def build_cyl_mask(nx, ny, nz, dx, R):
    """Voxel cylinder as in quick_compare_neumann_robin_backend.py:100-105."""
    cx = nx / 2.0
    cy = ny / 2.0
    xs = (np.arange(nx) + 0.5 - cx) * dx
    ys = (np.arange(ny) + 0.5 - cy) * dx
    X, Y = np.meshgrid(xs, ys, indexing="ij")
    mask2d = np.sqrt(X ** 2 + Y ** 2) <= R + 1e-12
    return np.repeat(mask2d[:, :, None], nz, axis=2)


def make_mask(kind: str, shape, seed: int) -> np.ndarray:
    nx, ny, nz = shape
    if kind == "full":
        return np.ones(shape, dtype=bool)
    if kind == "cyl":
        return build_cyl_mask(nx, ny, nz, DX, 0.42 * min(nx, ny) * DX)
    if kind == "cyl_holes":
        m = build_cyl_mask(nx, ny, nz, DX, 0.42 * min(nx, ny) * DX)
        holes = splitmix_uniform(seed + 101, shape) < 0.20
        holes[:, :, : nz // 2] = False
        return m & ~holes
    if kind == "plate_track":
        # single_track_on_plate.py:113-114,159 -- plate + a partially deposited track
        nzp = nz - max(2, nz // 8)
        m = np.zeros(shape, dtype=bool)
        m[:, :, :nzp] = True
        m[: max(1, nx // 6), : ny // 2, nzp:] = True
        return m
    if kind == "thin":
        # single-cell-wide features, isolated cells, 1-cell gaps
        m = np.zeros(shape, dtype=bool)
        m[nx // 2, :, :] = True            # a one-cell-thick plate
        m[:, ny // 3, nz // 2] = True      # a one-cell rod along x
        m[1, 1, 1] = True                  # an isolated cell
        m[::2, -1, ::3] = True             # checkered edge
        m[3:9, 3:9, 3:9] = True            # a block ...
        m[5, 5, :] = False                 # ... with a needle hole through it
        m[0, 0, :] = True                  # full line on the domain edge
        return m
    if kind == "random":
        return splitmix_uniform(seed + 77, shape) < 0.55
    if kind == "empty":
        return np.zeros(shape, dtype=bool)
    raise ValueError(kind)


def make_bcs(kind: str, shape, mask, seed: int, Tinf: float):
    """Returns kwargs for precompute_coeff_packs_unified (dir_mask, dir_value, neumann, robin_h)."""
    nx, ny, nz = shape
    faces = ("x-", "x+", "y-", "y+", "z-", "z+")
    kw = dict(dir_mask=None, dir_value=None, neumann=None, robin_h=None)
    if kind == "none":
        return kw
    if kind == "robin6":
        kw["robin_h"] = {f: 10.0 for f in faces}
        return kw
    if kind == "robin_scalar":
        kw["robin_h"] = 35.0
        return kw
    if kind == "robin_sides":
        kw["robin_h"] = {f: 120.0 for f in faces[:4]}
        return kw
    if kind == "robin_dict3d":
        kw["robin_h"] = {f: 500.0 * splitmix_uniform(seed + 11 + i, shape)
                         for i, f in enumerate(faces)}
        return kw
    if kind == "robin_mixed":
        # dict with scalars on some faces and 3-D fields on others, one face missing
        kw["robin_h"] = {"x-": 40.0, "y+": 300.0 * splitmix_uniform(seed + 21, shape),
                         "z-": 5.0, "z+": 80.0 * splitmix_uniform(seed + 22, shape)}
        return kw
    if kind == "robin_array":
        kw["robin_h"] = 200.0 * splitmix_uniform(seed + 31, shape)
        return kw
    if kind == "neumann_zm":
        kw["neumann"] = {"z-": 2.0e6}
        kw["robin_h"] = {f: 25.0 for f in faces[:4]}
        return kw
    if kind == "neumann_fields":
        kw["neumann"] = {"z-": 2.0e6 * splitmix_uniform(seed + 41, shape),
                         "x+": -3.0e5, "y-": None}
        return kw
    if kind == "dir_top":
        dm = np.zeros(shape, dtype=bool)
        dm[:, :, nz - 1] = mask[:, :, nz - 1]
        kw["dir_mask"] = dm
        kw["dir_value"] = np.full(shape, Tinf, dtype=float)
        kw["robin_h"] = {f: 15.0 for f in faces[:4]}
        return kw
    if kind == "dir_both":
        # quick_compare_dirichlet_robin.py:129-135 pattern: bottom hot, top ambient
        dm = np.zeros(shape, dtype=bool)
        dm[:, :, 0] = mask[:, :, 0]
        dm[:, :, nz - 1] = mask[:, :, nz - 1]
        dv = np.full(shape, Tinf, dtype=float)
        dv[:, :, 0] = 400.0
        kw["dir_mask"] = dm
        kw["dir_value"] = dv
        kw["robin_h"] = {f: 60.0 for f in faces[:4]}
        return kw
    if kind == "dir_scalar_interior":
        # scalar dir_value, Dirichlet cells scattered in the interior (and on void cells)
        kw["dir_mask"] = splitmix_uniform(seed + 51, shape) < 0.07
        kw["dir_value"] = 333.0
        return kw
    if kind == "combined":
        dm = np.zeros(shape, dtype=bool)
        dm[:, :, nz - 1] = mask[:, :, nz - 1]
        dm |= splitmix_uniform(seed + 52, shape) < 0.02
        kw["dir_mask"] = dm
        kw["dir_value"] = Tinf + 50.0 * splitmix_uniform(seed + 53, shape)
        kw["neumann"] = {"z-": 2.0e6, "x-": 1.0e5 * splitmix_uniform(seed + 54, shape)}
        kw["robin_h"] = {f: 500.0 * splitmix_uniform(seed + 60 + i, shape)
                         for i, f in enumerate(faces)}
        return kw
    raise ValueError(kind)


# name: (shape, mask_kind, bc_kind, theta, cfl, Tinf, nan_void, nsteps)
CART_CASES = {
    "full_robin6":            (SHAPE_A, "full",        "robin6",        0.5, 0.128, 20.0, False, 1),
    "full_none_be":           (SHAPE_A, "full",        "none",          1.0, 2.0,   0.0,  False, 1),
    "full_dict3d_cfl3000":    (SHAPE_A, "full",        "robin_dict3d",  0.5, 3000., 20.0, False, 1),
    "cyl_robin6":             (SHAPE_A, "cyl",         "robin6",        0.5, 0.128, 20.0, False, 1),
    "cyl_neumann":            (SHAPE_A, "cyl",         "neumann_zm",    0.5, 2.0,   20.0, False, 1),
    "cyl_dirtop":             (SHAPE_A, "cyl",         "dir_top",       0.5, 2.0,   20.0, False, 1),
    "cyl_dirboth_be":         (SHAPE_A, "cyl",         "dir_both",      1.0, 3000., 20.0, False, 1),
    "holes_dict3d":           (SHAPE_A, "cyl_holes",   "robin_dict3d",  0.5, 2.0,   20.0, True,  1),
    "holes_combined":         (SHAPE_A, "cyl_holes",   "combined",      0.5, 0.128, 20.0, True,  1),
    "holes_combined_cfl3000": (SHAPE_A, "cyl_holes",   "combined",      1.0, 3000., 0.0,  True,  1),
    "track_robin6":           (SHAPE_A, "plate_track", "robin6",        0.5, 0.128, 20.0, False, 1),
    "track_mixed":            (SHAPE_A, "plate_track", "robin_mixed",   0.5, 2.0,   20.0, True,  1),
    "thin_robin_scalar":      (SHAPE_A, "thin",        "robin_scalar",  0.5, 2.0,   20.0, True,  1),
    "thin_combined":          (SHAPE_A, "thin",        "combined",      0.5, 3000., 20.0, True,  1),
    "random_array":           (SHAPE_A, "random",      "robin_array",   0.5, 2.0,   0.0,  True,  1),
    "random_dirint":          (SHAPE_A, "random",      "dir_scalar_interior", 0.5, 0.128, 20.0, False, 1),
    "random_neumann_fields":  (SHAPE_A, "random",      "neumann_fields", 1.0, 2.0,  20.0, False, 1),
    "empty_robin6":           (SHAPE_A, "empty",       "robin6",        0.5, 2.0,   20.0, False, 1),
    "B_full_robin6":          (SHAPE_B, "full",        "robin6",        0.5, 0.128, 20.0, False, 1),
    "B_holes_combined":       (SHAPE_B, "cyl_holes",   "combined",      0.5, 3000., 20.0, True,  1),
    "B_track_dict3d":         (SHAPE_B, "plate_track", "robin_dict3d",  0.5, 2.0,   20.0, False, 1),
    "B_random_sides_be":      (SHAPE_B, "random",      "robin_sides",   1.0, 2.0,   20.0, True,  1),
    # multi-step: the quick_compare_neumann_robin_backend.py default BC set at reduced size
    "cyl_backend_10steps":    (SHAPE_A, "cyl",         "backend_default", 0.5, 0.5, 20.0, False, 10),
}


def build_cart_case(name: str) -> dict:
    shape, mk, bk, theta, cfl, Tinf, nan_void, nsteps = CART_CASES[name]
    seed = 1000 + sorted(CART_CASES).index(name) * 7
    mask = make_mask(mk, shape, seed)
    if bk == "backend_default":
        # quick_compare_neumann_robin_backend.py:120-131
        nx, ny, nz = shape
        dm = np.zeros(shape, dtype=bool)
        dm[:, :, nz - 1] = mask[:, :, nz - 1]
        bcs = dict(dir_mask=dm, dir_value=np.full(shape, Tinf, dtype=float),
                   neumann={"z-": 2.0e6},
                   robin_h={"x-": 150.0, "x+": 150.0, "y-": 150.0, "y+": 150.0})
        T0 = np.full(shape, Tinf, dtype=float)
    else:
        bcs = make_bcs(bk, shape, mask, seed, Tinf)
        T0 = 20.0 + 1380.0 * splitmix_uniform(seed + 1, shape)
    if nan_void:
        T0 = T0.copy()
        T0[~mask] = np.nan
    kappa = K / (RHO * CP)
    dt = cfl * DX * DX / kappa
    return dict(name=name, shape=shape, dx=DX, rho=RHO, cp=CP, k=K, mask=mask, T0=T0,
                theta=theta, dt=dt, Tinf=Tinf, nsteps=nsteps, bcs=bcs)


# ----------------------------------------------------------------------------
# Cylindrical cases (oracle: adi3d_cyl_phi_v3.adi_step, scheme "be";
# masked wrapper: quick_spiral_deposition_gif_v5.adi_step_masked)
# ----------------------------------------------------------------------------
C_RHO, C_CP, C_K = 7800.0, 490.0, 54.0

# name: (nr, nphi, nz, R, cfl, kind_bot, kind_top, h_r, source, masked)
CYL_CASES = {
    "spiral_grid":     (6, 36, 7,    0.032, 1.0,  "neumann0", "robin",     400.0, False, False),
    "mid_default":     (16, 32, 40,  0.02,  1.0,  "neumann0", "robin",     500.0, False, False),
    "mid_source":      (16, 32, 40,  0.02,  4.0,  "neumann0", "robin",     500.0, True,  False),
    "mid_dd":          (16, 32, 40,  0.02,  1.0,  "dirichlet", "dirichlet", 500.0, False, False),
    "mid_rr":          (16, 32, 40,  0.02,  10.0, "robin",    "robin",     250.0, False, False),
    "mid_nn_h0":       (16, 32, 40,  0.02,  1.0,  "neumann0", "neumann0",  0.0,   False, False),
    "mid_dn":          (16, 32, 40,  0.02,  0.3,  "dirichlet", "neumann0", 500.0, True,  False),
    "mid_rd":          (16, 32, 40,  0.02,  1.0,  "robin",    "dirichlet", 500.0, False, False),
    "mid_nd":          (16, 32, 40,  0.02,  1.0,  "neumann0", "dirichlet", 500.0, False, False),
    "mid_dr":          (16, 32, 40,  0.02,  1.0,  "dirichlet", "robin",    500.0, False, False),
    "mid_rn":          (16, 32, 40,  0.02,  1.0,  "robin",    "neumann0",  500.0, False, False),
    "nphi1":           (12, 1, 33,   0.02,  1.0,  "neumann0", "robin",     500.0, False, False),
    "odd_sizes":       (13, 37, 21,  0.02,  2.0,  "robin",    "robin",     300.0, True,  False),
    "big":             (32, 128, 64, 0.02,  1.0,  "neumann0", "robin",     500.0, False, False),
    "big_cfl50":       (24, 64, 32,  0.02,  50.0, "neumann0", "robin",     500.0, False, False),
    "masked_mid":      (16, 32, 40,  0.02,  1.0,  "neumann0", "robin",     500.0, False, True),
    "masked_spiral":   (6, 36, 7,    0.032, 1.0,  "neumann0", "robin",     400.0, False, True),
}


# round-2 additions (own seeds: the seeds of CYL_CASES depend on that dict's sorted order, which must not move):
# r lines at BASELINE configs[2] length (nr = 256: P = 16 chunks, Robin row at the far end), cfl 1 and 50
CYL_CASES_R2 = {
    "long_r":          ((256, 8, 6,  0.02,  1.0,  "neumann0", "robin",     500.0, False, False), 9001),
    "long_r_cfl50":    ((256, 8, 6,  0.02,  50.0, "robin",    "robin",     500.0, True,  False), 9014),
}
ALL_CYL_CASES = sorted(CYL_CASES) + sorted(CYL_CASES_R2)
# a slab of BASELINE configs[2] itself: 256 x 1024 x 16 (golden stored as a sub-sample + SHA-256 of the full array)
CYL_C3_SLICE = ((256, 1024, 16, 0.02, 1.0, "neumann0", "robin", 500.0, False, False), 9027)


def build_cyl_case(name: str) -> dict:
    if name in CYL_CASES_R2 or name == "c3_slice":
        (nr, nphi, nz, R, cfl, kb, kt, h_r, source, masked), seed = CYL_CASES_R2[name] if name in CYL_CASES_R2 else CYL_C3_SLICE
    else:
        nr, nphi, nz, R, cfl, kb, kt, h_r, source, masked = CYL_CASES[name]
        seed = 5000 + sorted(CYL_CASES).index(name) * 13
    dr = R / nr
    dz = dr
    dphi = (2.0 * math.pi) / max(nphi, 1)
    alpha = C_K / (C_RHO * C_CP)
    # quick_compare_layer_birth_robin_cyl_v3.py:126-127
    dt = cfl * min(dr * dr, dz * dz, (R * dphi) ** 2 if nphi > 1 else 1e9) / max(alpha, 1e-16)
    shape = (nr, nphi, nz)
    T0 = 20.0 + 980.0 * splitmix_uniform(seed + 1, shape)
    S = (5.0e9 * splitmix_uniform(seed + 2, shape)) if source else None
    active = None
    if masked:
        active = splitmix_uniform(seed + 3, shape) < 0.6
        active[:, :, : nz // 3] = True
    zbc = dict(kind_bot=kb, kind_top=kt, h_bot=180.0, h_top=500.0,
               T_inf_bot=35.0, T_inf_top=20.0, T_bot=250.0, T_top=60.0)
    return dict(name=name, nr=nr, nphi=nphi, nz=nz, dr=dr, dphi=dphi, dz=dz, R=R,
                rho=C_RHO, cp=C_CP, k=C_K, dt=dt, h_r=h_r, Tinf_r=20.0, zbc=zbc,
                T0=T0, S=S, active=active, T_void=20.0, T_inner=25.0)


def rel_l2(a: np.ndarray, b: np.ndarray, where=None) -> float:
    """Relative L2 error of a against b over `where` (all cells if None)."""
    if where is not None:
        a = a[where]
        b = b[where]
    if a.size == 0:
        return 0.0
    den = float(np.sqrt(np.sum(b.astype(np.float64) ** 2)))
    num = float(np.sqrt(np.sum((a.astype(np.float64) - b.astype(np.float64)) ** 2)))
    return num / den if den > 0 else num


# ----------------------------------------------------------------------------
# voxel_bc_correction cases: a duck-typed triangle mesh (the four attributes the reference
# reads: triangles, face_normals, area_faces, triangles_center) of an ellipsoid, and the voxel
# mask of the same ellipsoid.
# ----------------------------------------------------------------------------
class TriMesh:
    def __init__(self, triangles):
        self.triangles = np.ascontiguousarray(triangles, dtype=np.float64)
        e1 = self.triangles[:, 1] - self.triangles[:, 0]
        e2 = self.triangles[:, 2] - self.triangles[:, 0]
        n = np.cross(e1, e2)
        nrm = np.sqrt((n * n).sum(axis=1))
        self.area_faces = 0.5 * nrm
        self.face_normals = n / np.maximum(nrm, 1e-300)[:, None]
        self.triangles_center = self.triangles.mean(axis=1)


def ellipsoid_mesh(center, radii, nu=14, nv=24):
    """Latitude-longitude triangulation, outward normals."""
    cx, cy, cz = center
    a, b, c = radii

    def pt(i, j):
        th = math.pi * i / nu
        ph = 2.0 * math.pi * (j % nv) / nv
        return (cx + a * math.sin(th) * math.cos(ph), cy + b * math.sin(th) * math.sin(ph), cz + c * math.cos(th))

    tris = []
    for i in range(nu):
        for j in range(nv):
            p00, p01, p10, p11 = pt(i, j), pt(i, j + 1), pt(i + 1, j), pt(i + 1, j + 1)
            if i > 0:
                tris.append((p00, p10, p01))
            if i < nu - 1:
                tris.append((p01, p10, p11))
    return TriMesh(np.array(tris))


def build_voxel_bc_case(name="ellipsoid"):
    shape = (22, 26, 30)
    dx = 1.0e-3
    origin = (-2.0e-3, 1.0e-3, 0.5e-3)
    center = (origin[0] + 0.5 * shape[0] * dx, origin[1] + 0.5 * shape[1] * dx, origin[2] + 0.5 * shape[2] * dx)
    radii = (0.40 * shape[0] * dx, 0.42 * shape[1] * dx, 0.45 * shape[2] * dx)
    xs = [origin[d] + (np.arange(shape[d]) + 0.5) * dx for d in range(3)]
    X, Y, Z = np.meshgrid(*xs, indexing="ij")
    mask = ((X - center[0]) / radii[0]) ** 2 + ((Y - center[1]) / radii[1]) ** 2 + ((Z - center[2]) / radii[2]) ** 2 <= 1.0
    mesh = ellipsoid_mesh(center, radii, nu=10 if name == "coarse" else 22, nv=16 if name == "coarse" else 40)
    base_h = {"x-": 40.0, "x+": 55.0, "y-": 40.0, "y+": 0.0, "z-": 25.0, "z+": 80.0}
    return dict(name=name, shape=shape, dx=dx, origin=origin, mask=mask, mesh=mesh, base_h=base_h,
                max_subdiv=6 if name != "coarse" else 4)


# ---- ASCII VTK output (vtk_writer.py, waam_from_stl_v7_mm.py:186-215) -------------------------
def vtk_text_cases():
    """name -> dict(T, dx, origin, mask, field_name): small seeded fields that exercise every
    branch of '.6e' / '.6g' formatting (negative values, exact ties, zeros, tiny and huge
    magnitudes, 3-digit exponents, non-finite entries) and ragged line ends (N % 9 != 0)."""
    out = {}
    rng = np.random.default_rng(2024)
    T = 20.0 + 1380.0 * rng.random((7, 5, 4))
    m = rng.random((7, 5, 4)) < 0.7
    out["plate"] = dict(T=T, dx=1e-3, origin=(0.0, 0.0, 0.0), mask=m, field_name="Temperature")
    T = (rng.random((6, 3, 5)) - 0.4) * 10.0 ** rng.integers(-12, 13, size=(6, 3, 5))
    T[0, 0, 0] = 0.0; T[1, 0, 0] = -0.0; T[2, 0, 0] = 1234567.5; T[3, 0, 0] = 123456.5
    T[4, 0, 0] = 9.9999995; T[5, 0, 0] = 999999.5; T[0, 1, 0] = 1e-100; T[1, 1, 0] = -3.5e120
    T[2, 1, 0] = 1e5; T[3, 1, 0] = 1e-5; T[4, 1, 0] = 100.0; T[5, 1, 0] = 0.0001
    out["wide_range"] = dict(T=T, dx=0.25, origin=(-1.5, 2.0, 1e-7), mask=None, field_name="Temperature")
    T = rng.normal(0.0, 50.0, (10, 1, 1))
    out["line10"] = dict(T=T, dx=2.0, origin=(1.0, 2.0, 3.0), mask=np.ones((10, 1, 1), bool), field_name="T_C")
    T = rng.random((3, 4, 3)) * 1e3
    T[1, 1, 1] = np.nan; T[2, 2, 2] = np.inf; T[0, 3, 1] = -np.inf
    out["nonfinite"] = dict(T=T, dx=0.5, origin=(0.0, 0.0, 0.0), mask=None, field_name="Temperature")
    T = np.round(rng.random((9, 9, 2)) * 4096.0) / 16.0            # short dyadic values
    out["dyadic"] = dict(T=T.astype(np.float32), dx=1.0, origin=(0.0, 0.0, 0.0), mask=rng.random((9, 9, 2)) < 0.5,
                         field_name="Temperature")
    return out


def vtk_text_big_case():
    """A seeded field with multi-piece rows (nx > 256) whose reference files are pinned by digest only."""
    rng = np.random.default_rng(777)
    T = 20.0 + 1380.0 * rng.random((300, 5, 7))
    T[::7, :, 1] *= -1.0
    T[5:9, 2, 3] = 0.0
    T[::11, 1, :] = np.round(T[::11, 1, :] * 8.0) / 8.0
    m = rng.random((300, 5, 7)) < 0.6
    return dict(T=T, dx=2.5e-4, origin=(0.1, -0.2, 0.3), mask=m, field_name="Temperature")
