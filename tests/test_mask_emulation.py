# -*- coding: utf-8 -*-
"""Word forms of the per-mask-change kernels (csrc/adi_mask_core.h: k_build_code_v, k_transpose_code_v,
k_build_packs_v) run on the CPU through csrc/host_emulation.cpp -- the same per-thread source the sm_100a kernels are
compiled from -- against the one-cell-per-thread forms, NumPy and the oracle's precompute_coeff_packs_unified
(adi3d_gpu_coeff.py:31-110).  Integer / byte work and per-face products: bit-exact."""
import ctypes as C

import numpy as np
import pytest

import emu
from oracle import cart

BP = C.POINTER(C.c_uint8)
DP = C.POINTER(C.c_double)


def _bp(a):
    return a.ctypes.data_as(BP) if a is not None else C.cast(None, BP)


def _dp(a):
    return a.ctypes.data_as(DP) if a is not None else C.cast(None, DP)


def _mask(shape, kind, rng):
    nx, ny, nz = shape
    if kind == "full":
        return np.ones(shape, dtype=np.uint8)
    if kind == "void":
        return np.zeros(shape, dtype=np.uint8)
    if kind == "random":
        return (rng.random(shape) < 0.6).astype(np.uint8)
    if kind == "bytes":      # any non-zero byte counts as active, as `if (mask[idx])` does
        m = rng.integers(0, 256, shape).astype(np.uint8)
        m[rng.random(shape) < 0.4] = 0
        return m
    if kind == "layers":     # a part under construction: the lower planes of a blob, the rest void
        i, j, k = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
        blob = ((i - nx / 2) ** 2 / (0.4 * nx) ** 2 + (j - ny / 2) ** 2 / (0.45 * ny) ** 2) <= 1.0
        return (blob & (k < (2 * nz) // 3)).astype(np.uint8)
    raise KeyError(kind)


def _code_ref(mask, dirm, mlo, mhi):
    L = emu.lib()
    nx, ny, nz = mask.shape
    code = np.empty(mask.shape, dtype=np.uint8)
    L.emu_build_code_halo.argtypes = [BP, BP, BP, C.c_int, C.c_int, C.c_int, BP, BP]
    L.emu_build_code_halo.restype = None
    L.emu_build_code_halo(_bp(mask), _bp(dirm), _bp(code), nx, ny, nz, _bp(mlo), _bp(mhi))
    return code


@pytest.mark.parametrize("shape", [(5, 6, 16), (4, 3, 48), (1, 1, 32), (3, 1, 16), (7, 9, 64)])
@pytest.mark.parametrize("kind", ["full", "void", "random", "bytes", "layers"])
@pytest.mark.parametrize("halo", [False, True])
@pytest.mark.parametrize("with_dir", [False, True])
def test_code_word_form_matches_the_cell_form(shape, kind, halo, with_dir):
    rng = np.random.default_rng(hash((shape, kind, halo, with_dir)) & 0xffff)
    mask = _mask(shape, kind, rng)
    nx, ny, nz = shape
    dirm = (rng.random(shape) < 0.2).astype(np.uint8) if with_dir else None
    mlo = (rng.random((nx, ny)) < 0.5).astype(np.uint8) if halo else None
    mhi = (rng.random((nx, ny)) < 0.5).astype(np.uint8) if halo else None
    ref = _code_ref(mask, dirm, mlo, mhi)
    L = emu.lib()
    L.emu_build_code_v.argtypes = [BP, BP, BP, C.c_int, C.c_int, C.c_int, BP, BP]
    L.emu_build_code_v.restype = C.c_int
    out = np.full(shape, 0xAA, dtype=np.uint8)
    top = L.emu_build_code_v(_bp(mask), _bp(dirm), _bp(out), nx, ny, nz, _bp(mlo), _bp(mhi))
    assert np.array_equal(out, ref)
    planes = np.nonzero(mask.any(axis=(0, 1)))[0]
    assert top == (int(planes[-1]) + 1 if planes.size else 0)      # top of the part, for the trimmed z sweep


@pytest.mark.parametrize("shape", [(5, 6, 16), (130, 3, 132), (33, 2, 4), (256, 2, 128), (100, 5, 260), (1, 1, 4)])
@pytest.mark.parametrize("axis", [0, 1])
def test_transposed_code_word_form(shape, axis):
    rng = np.random.default_rng(7)
    nx, ny, nz = shape
    code = rng.integers(0, 256, shape).astype(np.uint8)
    n, batch = (nx, ny) if axis == 0 else (ny, nx)
    npad = (n + 31) // 32 * 32
    snx = ny * nz
    sb, sr = (nz, snx) if axis == 0 else (snx, nz)
    dst = np.zeros((batch, nz, npad), dtype=np.uint8)    # cudaMemset of the allocation
    L = emu.lib()
    L.emu_transpose_code_v.argtypes = [BP, BP, C.c_int, C.c_int, C.c_int, C.c_int, C.c_size_t, C.c_size_t]
    L.emu_transpose_code_v.restype = None
    L.emu_transpose_code_v(_bp(code), _bp(dst), n, nz, npad, batch, sb, sr)
    # dst[b][c][r] = code[r, b, c] (x lines) / code[b, r, c] (y lines); padding columns zero
    want = np.zeros_like(dst)
    want[:, :, :n] = code.transpose(1, 2, 0) if axis == 0 else code.transpose(0, 2, 1)
    assert np.array_equal(dst, want)


def _packs_v(mask, dx, rho, cp, h, q, mlo=None, mhi=None, want_q=True, nc=2):
    """h / q: six entries, each None, a scalar or a dense field."""
    nx, ny, nz = mask.shape
    kinds = lambda v: 0 if v is None else (1 if np.isscalar(v) else 2)
    hk = (C.c_int * 6)(*[kinds(v) for v in h])
    qk = (C.c_int * 6)(*[kinds(v) for v in q])
    hs = (C.c_double * 6)(*[float(v) if np.isscalar(v) else 0.0 for v in h])
    qs = (C.c_double * 6)(*[float(v) if np.isscalar(v) else 0.0 for v in q])
    hf = (DP * 6)(*[_dp(v) if isinstance(v, np.ndarray) else C.cast(None, DP) for v in h])
    qf = (DP * 6)(*[_dp(v) if isinstance(v, np.ndarray) else C.cast(None, DP) for v in q])
    coeff = [np.full(mask.shape, np.nan) for _ in range(3)]
    qout = [np.full(mask.shape, np.nan) for _ in range(3)] if want_q else [None] * 3
    L = emu.lib()
    L.emu_build_packs_v.argtypes = [C.c_int, BP, C.c_int, C.c_int, C.c_int, BP, BP, C.c_double, C.c_double, C.c_double,
                                    C.POINTER(C.c_int), DP, C.POINTER(DP), C.POINTER(C.c_int), DP, C.POINTER(DP),
                                    C.POINTER(DP), C.POINTER(DP)]
    L.emu_build_packs_v.restype = None
    L.emu_build_packs_v(nc, _bp(mask), nx, ny, nz, _bp(mlo), _bp(mhi), dx, rho, cp, hk, hs, hf, qk, qs, qf,
                        (DP * 3)(*[_dp(a) for a in coeff]), (DP * 3)(*[_dp(a) for a in qout]))
    return coeff, qout


@pytest.mark.parametrize("shape,nc", [((5, 6, 8), 2), ((5, 6, 8), 4), ((4, 3, 12), 4), ((4, 3, 10), 2), ((1, 1, 4), 4),
                                      ((1, 1, 2), 2), ((6, 7, 32), 2), ((6, 7, 32), 4)])
@pytest.mark.parametrize("kind", ["full", "random", "layers", "void"])
@pytest.mark.parametrize("bc", ["scalar", "fields", "mixed"])
def test_pack_builder_word_form_bit_exact_vs_oracle(shape, nc, kind, bc):
    rng = np.random.default_rng(11)
    mask = _mask(shape, kind, rng)
    dx, rho, cp = 1.3e-3, 7800.0, 490.0
    fld = lambda lo, hi: np.ascontiguousarray(lo + (hi - lo) * rng.random(shape))
    if bc == "scalar":
        h = [25.0, 40.0, 0.0, 12.5, 300.0, 7.0]
        q = [None, 1.0e4, None, None, -2.0e3, None]
    elif bc == "fields":
        h = [fld(5, 500) for _ in range(6)]
        q = [fld(-1e4, 1e4) for _ in range(6)]
    else:
        h = [fld(5, 500), 40.0, None, fld(1, 2), 300.0, fld(0, 1)]
        q = [None, fld(-1e4, 1e4), 5.0e3, None, None, fld(0, 10)]
    coeff, qout = _packs_v(mask, dx, rho, cp, h, q, nc=nc)
    grid = cart.Grid3D(*shape, dx, mask.astype(bool))
    mat = cart.Material(rho, cp, 54.0)
    robin = {f: (0.0 if v is None else v) for f, v in zip(cart.FACES, h)}
    neumann = {f: v for f, v in zip(cart.FACES, q) if v is not None}
    packs = cart.precompute_coeff_packs_unified(grid, mat, robin_h=robin, neumann=neumann)
    for ax in range(3):
        assert np.array_equal(coeff[ax].view(np.uint64), packs[ax].coeff.view(np.uint64)), f"coeff axis {ax}"
        assert np.array_equal(qout[ax].view(np.uint64), packs[ax].qflux.view(np.uint64)), f"qflux axis {ax}"


def test_pack_builder_word_form_slab_halo_planes():
    """z-slab of a larger grid: the adjacent ranks' mask planes decide exposure at the slab's ends."""
    rng = np.random.default_rng(5)
    shape = (5, 4, 24)
    mask = _mask(shape, "random", rng)
    dx, rho, cp = 2.0e-3, 7800.0, 490.0
    h = [10.0, 20.0, 30.0, 40.0, 50.0, 60.0]
    grid = cart.Grid3D(*shape, dx, mask.astype(bool))
    packs = cart.precompute_coeff_packs_unified(grid, cart.Material(rho, cp, 54.0),
                                                robin_h=dict(zip(cart.FACES, h)))
    for z0, z1 in ((0, 8), (8, 16), (16, 24)):
        sub = np.ascontiguousarray(mask[:, :, z0:z1])
        mlo = np.ascontiguousarray(mask[:, :, z0 - 1]) if z0 > 0 else None
        mhi = np.ascontiguousarray(mask[:, :, z1]) if z1 < shape[2] else None
        for nc in (2, 4):
            coeff, _ = _packs_v(sub, dx, rho, cp, h, [None] * 6, mlo, mhi, want_q=False, nc=nc)
            for ax in range(3):
                assert np.array_equal(coeff[ax].view(np.uint64),
                                      np.ascontiguousarray(packs[ax].coeff[:, :, z0:z1]).view(np.uint64))
