# -*- coding: utf-8 -*-
"""Test-only backend for adi_thermal_fields_b200.slab: the slab phases run through the host build of
the kernels' per-thread code (csrc/host_emulation.cpp, emu_cart_slab) on CPU torch tensors, so the
exchange logic of the N>1 path (halo planes, interface all-gather, inter-rank solve) can be
tested with gloo / in-process ranks where there is no GPU.  Never used by the product path."""
import ctypes as C

import numpy as np
import torch

import emu


class HostBackend:
    def __init__(self, variant=0):
        self.L = emu.lib()
        self.variant = variant
        self.launches = 0

    def empty(self, shape, dtype):
        return torch.zeros(shape, dtype=dtype)

    def asarray(self, x, dtype):
        t = x if isinstance(x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(x))
        return t.to(dtype=dtype).contiguous().clone()

    def bind(self, nx, ny, nz, dx, mask, rank, world):
        self.nx, self.ny, self.nz, self.dx = nx, ny, nz, dx
        self.mask, self.rank, self.world = mask, rank, world
        self.mlo = self.mhi = None
        self.packs, self.fc = [(None,) * 4] * 3, None

    def mark_mask_changed(self, mask):
        self.mask = mask

    def pack_planes(self, field, lo, hi):
        lo.copy_(field[:, :, 0])
        hi.copy_(field[:, :, -1])

    def set_mask_halo(self, lo, hi):
        self.mlo, self.mhi = lo, hi

    def build_packs(self, rho, cp, hk, hs, hf, qk, qs, qf, shape):
        """NumPy statement of precompute_coeff_packs_unified (adi3d_numba_coeff.py:57-118) with the
        z neighbours across slab boundaries taken from the halo planes."""
        m = self.mask.numpy().astype(bool)
        nx, ny, nz = m.shape
        lo = self.mlo.numpy().astype(bool) if self.mlo is not None else np.zeros((nx, ny), bool)
        hi = self.mhi.numpy().astype(bool) if self.mhi is not None else np.zeros((nx, ny), bool)
        pad = np.zeros((nx + 2, ny + 2, nz + 2), bool)
        pad[1:-1, 1:-1, 1:-1] = m
        pad[1:-1, 1:-1, 0] = lo
        pad[1:-1, 1:-1, -1] = hi
        nb = [pad[:-2, 1:-1, 1:-1], pad[2:, 1:-1, 1:-1], pad[1:-1, :-2, 1:-1], pad[1:-1, 2:, 1:-1],
              pad[1:-1, 1:-1, :-2], pad[1:-1, 1:-1, 2:]]
        A, Ccell = self.dx * self.dx, rho * cp * self.dx ** 3
        dense_h = any(k == 2 for k in hk)
        coeffs, qo = [None] * 3, [None] * 3
        for a in range(3):
            c = np.zeros(shape)
            q = np.zeros(shape)
            for s in range(2):
                f = 2 * a + s
                ex = m & ~nb[f]
                if hk[f]:
                    h = hf[f].numpy() if hk[f] == 2 else np.full(shape, hs[f])
                    c[ex] += h[ex] * A / Ccell
                if qk[f]:
                    v = qf[f].numpy() if qk[f] == 2 else np.full(shape, qs[f])
                    q[ex] += v[ex] * A / Ccell
            if dense_h:
                coeffs[a] = torch.from_numpy(c)
            if qk[2 * a] or qk[2 * a + 1]:
                qo[a] = torch.from_numpy(q)
        return coeffs, qo

    def set_packs(self, packs, face_coeff):
        self.packs, self.fc = packs, face_coeff

    def _call(self, Tin, Tout, Tlo, Thi, phase, dyn, stat, dyn_all, stat_all, dt, theta, kappa, Tinf):
        dp, bp = C.POINTER(C.c_double), C.POINTER(C.c_uint8)

        def p(t, ty):
            return C.cast(t.data_ptr(), C.POINTER(ty)) if t is not None else C.cast(None, C.POINTER(ty))

        def arr3(idx, ty):
            return (C.POINTER(ty) * 3)(*[p(self.packs[a][idx], ty) for a in range(3)])

        fc = None if self.fc is None else torch.tensor(self.fc, dtype=torch.float64)
        self.L.emu_cart_slab.argtypes = [dp, dp, bp, C.c_int, C.c_int, C.c_int] + [C.c_double] * 5 + \
            [C.POINTER(dp), C.POINTER(bp), C.POINTER(dp), C.POINTER(dp), dp, C.c_int, bp, bp, dp, dp, C.c_int,
             dp, dp, dp, dp, C.c_int, C.c_int]
        rc = self.L.emu_cart_slab(p(Tin, C.c_double), p(Tout, C.c_double), p(self.mask, C.c_uint8), self.nx, self.ny,
                                  self.nz, self.dx, dt, theta, kappa, Tinf, arr3(0, C.c_double), arr3(1, C.c_uint8),
                                  arr3(2, C.c_double), arr3(3, C.c_double), p(fc, C.c_double), self.variant,
                                  p(self.mlo, C.c_uint8), p(self.mhi, C.c_uint8), p(Tlo, C.c_double),
                                  p(Thi, C.c_double), phase, p(dyn, C.c_double), p(stat, C.c_double),
                                  p(dyn_all, C.c_double), p(stat_all, C.c_double), self.rank, self.world)
        assert rc == 0, rc
        self.launches += 1

    def step_xy(self, Tin, Tout, Tlo, Thi, dt, theta, kappa, Tinf):
        self._call(Tin, Tout, Tlo, Thi, 0, None, None, None, None, dt, theta, kappa, Tinf)

    def zsweep_reduce(self, T, dyn, stat, dt, theta, kappa, Tinf):
        self._call(T, T, None, None, 1 if stat is not None else 3, dyn, stat, None, None, dt, theta, kappa, Tinf)

    def zsweep_finish(self, T, dyn_all, stat_all, dt, theta, kappa, Tinf):
        self._call(T, T, None, None, 2, None, None, dyn_all, stat_all, dt, theta, kappa, Tinf)

    # solve-first z pass (adi_cart_zsweep_spike / _solve0 / _apply): the kernels' per-thread code for the
    # solves (emu_cart_slab phases 4-7), NumPy for k_spike_pack / k_spike_apply
    def zsweep_spikes(self, shape, kmax, threshold, dt, theta, kappa):
        nz = shape[2]
        out = []
        for end in (0, 1):
            F = torch.zeros(shape, dtype=torch.float64)
            self._call(F, F, None, None, 5 + end, None, None, None, None, dt, theta, kappa, 0.0)
            f = F.numpy().reshape(-1, nz)
            big = np.abs(f) > threshold
            first, last = np.argmax(big, axis=1), nz - 1 - np.argmax(big[:, ::-1], axis=1)
            far = np.where(big.any(axis=1), (last + 1) if end == 0 else (nz - first), 0)
            if far.size and far.max() > kmax:
                return None
            comp = f[:, :kmax] if end == 0 else f[:, nz - kmax:]
            out += [torch.from_numpy(np.ascontiguousarray(comp)), torch.from_numpy(np.minimum(far, kmax).astype(np.int32))]
        return out[0], out[2], out[1], out[3]

    def zsweep_solve0(self, T, dyn, dt, theta, kappa, Tinf):
        self._call(T, T, None, None, 4, dyn, None, None, None, dt, theta, kappa, Tinf)

    def zsweep_apply(self, T, dyn_all, stat_all, spikes, kmax):
        vC, wC, Kv, Kw = (x.numpy() for x in spikes)
        nl, nz = self.nx * self.ny, self.nz
        ghosts = torch.zeros((2, nl), dtype=torch.float64)
        self._call(T, T, None, None, 7, ghosts, None, dyn_all, stat_all, 0.0, 0.0, 0.0, 0.0)
        L, R = ghosts.numpy()
        j = np.arange(kmax)
        corr = np.zeros((nl, nz))
        corr[:, :kmax] += np.where(j[None, :] < Kv[:, None], L[:, None] * vC, 0.0)
        corr[:, nz - kmax:] += np.where(j[None, :] >= kmax - Kw[:, None], R[:, None] * wC, 0.0)
        t = T.numpy().reshape(nl, nz)           # a view: the update lands in T
        hit = corr != 0.0
        t[hit] += corr[hit]
        self.launches += 1

    def launch_count(self):
        return self.launches
