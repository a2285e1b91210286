# -*- coding: utf-8 -*-
"""NumPy restatement of the reference's cylindrical (r, phi, z) backward-Euler ADI
step -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Restates adi3d_cyl_phi_v3.py (scheme "be", :338-350) and the activation-mask
wrapper quick_spiral_deposition_gif_v5.py:31-70.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline leg may import it.

Parity status: PINNED against the unmodified reference (tools/gen_golden.py ->
tests/golden/cyl_*.npz, spiral_sim.npz, cyl_birth.npz; tests/test_oracle_golden.py).

Third-party arithmetic on this path: the periodic phi solve of the reference is
np.fft.rfft / irfft (NumPy's bundled pocketfft; NumPy 2.3.5 in the build container,
unpinned by the reference -- it ships no requirements file).  It is an exact solver of
the circulant system (I - fac*D2_periodic) x = rhs; this file restates it the same way
(phi_solve_spectral) and also provides the direct cyclic-tridiagonal solution
(phi_solve_cyclic, Sherman-Morrison) that the CUDA kernel implements, so the two can be
compared on the CPU.  The reference's own `_cyclic_thomas_batch_np` (:92-123) is dead code
and wrong and is NOT restated.

The reference builds (M, n) coefficient matrices whose rows are all equal; here the
coefficients are 1-D tables along the swept axis (same arithmetic per entry).
"""
from __future__ import annotations

import numpy as np


class GridCyl:  # adi3d_cyl_phi_v3.py:33-43 (+ R_in accepted and ignored)
    def __init__(self, nr, nphi, nz, dr, dphi, dz, R, R_in=0.0):
        self.nr, self.nphi, self.nz = int(nr), int(nphi), int(nz)
        self.dr, self.dphi, self.dz = float(dr), float(dphi), float(dz)
        self.R = float(R)
        self.R_in = float(R_in)
        self.r = (np.arange(self.nr, dtype=np.float64) + 0.5) * self.dr
        self.r_imh = self.r - 0.5 * self.dr
        self.r_iph = self.r + 0.5 * self.dr
        self.r_outer_face = self.r_iph[-1]


class Material:  # :45-50
    def __init__(self, rho, cp, k):
        self.rho, self.cp, self.k = float(rho), float(cp), float(k)

    @property
    def alpha(self):
        return self.k / (self.rho * self.cp)


class Params:  # :52-54
    def __init__(self, dt, theta=0.5, scheme="be"):
        self.dt, self.theta, self.scheme = float(dt), float(theta), str(scheme).lower()


class RobinR:  # :56-58
    def __init__(self, h, T_inf):
        self.h, self.T_inf = float(h), float(T_inf)


class ZBC:  # :60-68
    def __init__(self, kind_bot="neumann0", kind_top="robin", h_bot=0.0, h_top=0.0,
                 T_inf_bot=20.0, T_inf_top=20.0, T_bot=20.0, T_top=20.0):
        self.kind_bot, self.kind_top = kind_bot, kind_top
        self.h_bot, self.h_top = float(h_bot), float(h_top)
        self.T_inf_bot, self.T_inf_top = float(T_inf_bot), float(T_inf_top)
        self.T_bot, self.T_top = float(T_bot), float(T_top)


def r_tables(grid, mat, dt, robin_r, theta=1.0):
    """adi3d_cyl_phi_v3.py:155-201  build_coeff_r: returns 1-D (a, b, c) along r and the
    scalar added to the last RHS entry (Robin ambient term, :200-201)."""
    nr, dr = grid.nr, grid.dr
    r_i = np.maximum(grid.r, 1e-15)
    r_imh = np.maximum(grid.r_imh, 1e-15)
    r_iph = grid.r_iph
    fac = theta * mat.alpha * dt
    a = np.zeros(nr)
    b = np.zeros(nr)
    c = np.zeros(nr)
    ai = -fac * (r_imh[1:-1] / (r_i[1:-1] * dr * dr))  # :175-180
    ci = -fac * (r_iph[1:-1] / (r_i[1:-1] * dr * dr))
    a[1:-1] = ai
    b[1:-1] = 1.0 - (ai + ci)
    c[1:-1] = ci
    a[0] = 0.0  # axis row :183-186
    c0 = -fac * (r_iph[0] / (r_i[0] * dr * dr))
    b[0] = 1.0 - c0
    c[0] = c0
    h = float(robin_r.h)  # outer Robin row :189-196
    aN = -fac * (r_imh[-1] / (r_i[-1] * dr * dr))
    bN = 1.0 + fac * (r_imh[-1] / (r_i[-1] * dr * dr))
    add = 0.0
    if h != 0.0:
        bN += fac * (r_iph[-1] * (h / mat.k)) / (r_i[-1] * dr)
        add = fac * (r_iph[-1] * (h / mat.k)) / (r_i[-1] * dr) * robin_r.T_inf
    a[-1] = aN
    b[-1] = bN
    c[-1] = 0.0
    return a, b, c, add


def z_tables(grid, mat, dt, zbc, theta=1.0):
    """adi3d_cyl_phi_v3.py:255-298  build_coeff_z: 1-D (a, b, c), plus for each end
    ('set', value) for Dirichlet or ('add', value) otherwise."""
    nz, dz = grid.nz, grid.dz
    fac = theta * mat.alpha * dt / (dz * dz)
    a = np.zeros(nz)
    b = np.zeros(nz)
    c = np.zeros(nz)
    a[1:-1] = -fac
    b[1:-1] = 1.0 + 2.0 * fac
    c[1:-1] = -fac
    bot = ("add", 0.0)
    top = ("add", 0.0)
    if zbc.kind_bot == "neumann0":
        a[0] = 0.0; b[0] = 1.0 + fac; c[0] = -fac
    elif zbc.kind_bot == "dirichlet":
        a[0] = 0.0; b[0] = 1.0; c[0] = 0.0
        bot = ("set", zbc.T_bot)
    elif zbc.kind_bot == "robin":
        beta = zbc.h_bot / mat.k
        a[0] = 0.0; b[0] = 1.0 + fac * (1.0 + beta * dz); c[0] = -fac
        bot = ("add", (theta * mat.alpha * dt) * (beta / dz) * zbc.T_inf_bot)
    else:
        raise ValueError("unknown zbc.kind_bot")
    if zbc.kind_top == "neumann0":
        a[-1] = -fac; b[-1] = 1.0 + fac; c[-1] = 0.0
    elif zbc.kind_top == "dirichlet":
        a[-1] = 0.0; b[-1] = 1.0; c[-1] = 0.0
        top = ("set", zbc.T_top)
    elif zbc.kind_top == "robin":
        beta = zbc.h_top / mat.k
        a[-1] = -fac; b[-1] = 1.0 + fac * (1.0 + beta * dz); c[-1] = 0.0
        top = ("add", (theta * mat.alpha * dt) * (beta / dz) * zbc.T_inf_top)
    else:
        raise ValueError("unknown zbc.kind_top")
    return a, b, c, bot, top


def thomas_axis(a, b, c, d, axis):
    """adi3d_cyl_phi_v3.py:71-87  _thomas_batch_np (normalised Thomas, c' and d' form),
    applied along `axis` of d with line-independent 1-D coefficient tables."""
    d = np.moveaxis(d, axis, 0)
    n = d.shape[0]
    cp = np.empty(n)
    dp = np.empty_like(d)
    x = np.empty_like(d)
    cp[0] = c[0] / b[0]
    dp[0] = d[0] / b[0]
    for i in range(1, n):
        denom = b[i] - a[i] * cp[i - 1]
        cp[i] = c[i] / denom if i < n - 1 else 0.0
        dp[i] = (d[i] - a[i] * dp[i - 1]) / denom
    x[n - 1] = dp[n - 1]
    for i in range(n - 2, -1, -1):
        x[i] = dp[i] - cp[i] * x[i + 1]
    return np.moveaxis(x, 0, axis)


def phi_fac(grid, mat, dt, theta=1.0):
    """adi3d_cyl_phi_v3.py:311-317: fac_i = theta*alpha*dt/(r_i^2 dphi^2), fac_0 = 0 (axis ring)."""
    fac = np.zeros(grid.nr)
    r = grid.r
    for ir in range(1, grid.nr):
        fac[ir] = theta * mat.alpha * dt / (r[ir] * r[ir] * grid.dphi * grid.dphi)
    return fac


def phi_solve_spectral(Tin, grid, mat, theta, dt):
    """adi3d_cyl_phi_v3.py:302-329: rfft along phi, divide by
    lambda_k = 1 + 2 fac_i (1 - cos(2 pi k / nphi)), irfft."""
    nr, nphi, nz = Tin.shape
    if nphi == 1:
        return Tin.copy()
    fac = phi_fac(grid, mat, dt, theta)
    k = np.arange(nphi // 2 + 1, dtype=np.float64)
    cosk = np.cos(2.0 * np.pi * k / float(nphi))
    lam = 1.0 + 2.0 * fac[:, None] * (1.0 - cosk[None, :])
    F = np.fft.rfft(Tin, axis=1)
    F /= lam[:, :, None]
    return np.fft.irfft(F, n=nphi, axis=1)


def phi_solve_cyclic(Tin, grid, mat, theta, dt):
    """The same periodic system solved directly: rows (-f, 1+2f, -f) with wrap-around,
    by Sherman-Morrison on the tridiagonal part (u=(-f,0..0,-f)^T... written out below).
    This is what the CUDA kernel computes; tests compare it with phi_solve_spectral."""
    nr, nphi, nz = Tin.shape
    if nphi == 1:
        return Tin.copy()
    fac = phi_fac(grid, mat, dt, theta)
    out = np.empty_like(Tin)
    for ir in range(nr):
        f = fac[ir]
        if f == 0.0 or nphi == 2:
            if f == 0.0:
                out[ir] = Tin[ir]
            else:  # nphi == 2: both neighbours are the same cell -> 2x2 system
                b, o = 1.0 + 2.0 * f, -2.0 * f
                det = b * b - o * o
                out[ir, 0] = (b * Tin[ir, 0] - o * Tin[ir, 1]) / det
                out[ir, 1] = (b * Tin[ir, 1] - o * Tin[ir, 0]) / det
            continue
        # A = B + u v^T with B tridiagonal, B[0,0] = b - g, B[n-1,n-1] = b - w*w/g,
        # u = (g, 0, .., 0, w)^T, v = (1, 0, .., 0, w/g)^T, w = -f (the wrap coupling), g = -b.
        n = nphi
        bdiag = 1.0 + 2.0 * f
        w = -f
        g = -bdiag
        a = np.full(n, -f); a[0] = 0.0
        c = np.full(n, -f); c[-1] = 0.0
        b = np.full(n, bdiag); b[0] = bdiag - g; b[-1] = bdiag - w * w / g
        u = np.zeros(n); u[0] = g; u[-1] = w
        y = thomas_axis(a, b, c, Tin[ir], 0)
        q = thomas_axis(a, b, c, u[:, None], 0)[:, 0]
        vy = y[0] + (w / g) * y[-1]
        vq = q[0] + (w / g) * q[-1]
        out[ir] = y - q[:, None] * (vy / (1.0 + vq))
    return out


def adi_step(Tn, grid, mat, prm, robin_r, zbc, S=None, theta=None, phi_solver=phi_solve_spectral):
    """adi3d_cyl_phi_v3.py:332-350, scheme 'be': r-, phi-, z-implicit solves with theta=1.
    (scheme 'douglas' is non-deterministic in the reference -- np.empty_like reads at
    :149-151, SURVEY.md F4 -- and is deliberately not restated.)"""
    if prm.scheme == "douglas":
        raise NotImplementedError("only scheme='be' is pinned (SURVEY.md F4)")
    dt = prm.dt
    Tn = np.asarray(Tn, dtype=np.float64)
    R0 = Tn + (dt * (S / (mat.rho * mat.cp)) if S is not None else 0.0)  # :339
    a, b, c, add = r_tables(grid, mat, dt, robin_r, 1.0)
    rhs = R0.copy()
    if add != 0.0 or robin_r.h != 0.0:
        rhs[-1] += add
    TR = thomas_axis(a, b, c, rhs, 0)  # :341-344
    Tphi = phi_solver(TR, grid, mat, 1.0, dt)  # :346
    a, b, c, bot, top = z_tables(grid, mat, dt, zbc, 1.0)
    d = Tphi.copy()
    if bot[0] == "set":
        d[:, :, 0] = bot[1]
    else:
        d[:, :, 0] += bot[1]
    if top[0] == "set":
        d[:, :, -1] = top[1]
    else:
        d[:, :, -1] += top[1]
    return thomas_axis(a, b, c, d, 2)  # :348-350


def adi_step_masked(Tn, grid, mat, prm, robin_outer, zbc, active, robin_inner=None,
                    robin_void=None, phi_solver=phi_solve_spectral):
    """quick_spiral_deposition_gif_v5.py:31-70: void cells are clamped to robin_void.T_inf
    before and after the step; inactive axis-ring cells are set to robin_inner.T_inf."""
    robin_inner = robin_inner or robin_outer
    robin_void = robin_void or robin_outer
    Tw = np.array(Tn, dtype=np.float64, copy=True)
    void = ~np.asarray(active, dtype=bool)
    Tw[void] = float(robin_void.T_inf)
    out = adi_step(Tw, grid, mat, prm, robin_outer, zbc, phi_solver=phi_solver)
    out[void] = float(robin_void.T_inf)
    out[0, void[0]] = float(robin_inner.T_inf)
    return out


def build_grid_annular(R_out, wall_thickness, height, z_back, nr, nphi, dz_override=None):
    """quick_spiral_deposition_gif_v5.py:74-80 (with a GridCyl that accepts R_in)."""
    import math
    R_in = max(0.0, R_out - wall_thickness)
    dr = (R_out - R_in) / float(nr)
    dz = dr if (dz_override is None or dz_override <= 0.0) else float(dz_override)
    nz = int(round((z_back + height) / dz))
    dphi = (2.0 * math.pi) / max(1, nphi)
    return GridCyl(nr, nphi, nz, dr, dphi, dz, R_out, R_in=R_in), R_in, R_out, dz
