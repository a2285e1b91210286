/*
 * oracle/adi_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C, CPU restatement of the reference's Cartesian ADI heat step
 * (Matemusi/ADI_thermal_fields, adi3d_numba_coeff.py).  It exists so that the
 * CUDA path can be checked against the reference's algorithm on a box where
 * the reference itself (Python + Numba) is not present.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library; the product (adi_thermal_fields_b200/) never does.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks every function
 * here against outputs of the unmodified reference run in the build container
 * (tools/gen_golden.py -> tests/golden/cart_*.npz, packs_*.npz).
 *
 * Each function cites the reference lines it restates.  Arithmetic is kept in
 * the reference's evaluation order and the file is compiled with
 * -ffp-contract=off, so results agree with Numba's (no FMA contraction there).
 *
 * Layout: C-order (nx,ny,nz) -> z contiguous; masks are 1 byte per cell.
 * Lines are independent, so the sweeps may run under OpenMP; the reference is
 * serial (no prange), which is why the thread count is a caller's choice
 * (oracle_set_threads) and is reported by the benchmark as `cores`.
 */
#include <stdint.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define IDX(i, j, k) (((size_t)(i) * ny + (size_t)(j)) * nz + (size_t)(k))

static int g_threads = 1;

void oracle_set_threads(int n) { g_threads = n < 1 ? 1 : n; }

int oracle_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* adi3d_numba_coeff.py:38-55  exposed_mask(mask, face)
 * face: 0 'x-', 1 'x+', 2 'y-', 3 'y+', 4 'z-', 5 'z+'.
 * A cell is exposed on a face iff it is active and its neighbour across that
 * face is void or lies outside the domain. */
int oracle_exposed_mask(const uint8_t *mask, int nx, int ny, int nz, int face, uint8_t *exp)
{
    if (face < 0 || face > 5) return -1; /* ValueError("bad face") :54 */
    const int ax = face >> 1, plus = face & 1;
#pragma omp parallel for collapse(2) num_threads(g_threads)
    for (int i = 0; i < nx; ++i)
        for (int j = 0; j < ny; ++j)
            for (int k = 0; k < nz; ++k) {
                int ii = i, jj = j, kk = k, inside;
                if (ax == 0) { ii += plus ? 1 : -1; inside = ii >= 0 && ii < nx; }
                else if (ax == 1) { jj += plus ? 1 : -1; inside = jj >= 0 && jj < ny; }
                else { kk += plus ? 1 : -1; inside = kk >= 0 && kk < nz; }
                uint8_t nb = inside ? mask[IDX(ii, jj, kk)] : 0;
                exp[IDX(i, j, k)] = (uint8_t)(mask[IDX(i, j, k)] && !nb);
            }
    return 0;
}

/* adi3d_numba_coeff.py:57-118  precompute_coeff_packs_unified
 * h_kind[f] / q_kind[f]: 0 = absent (None), 1 = scalar h_scalar[f], 2 = field h_field[f].
 * Robin (:93-99): coeff_axis[exp] += h[exp]*A/Ccell, faces visited '-' then '+'.
 * Neumann (:104-114): S[exp] = q[exp]*A/Ccell summed per axis in dict order; the
 * caller passes faces in that order through q_order[0..nq).  Outputs are zeroed here. */
int oracle_precompute_packs(const uint8_t *mask, int nx, int ny, int nz, double dx,
                            double rho, double cp,
                            const int *h_kind, const double *h_scalar, const double *const *h_field,
                            int nq, const int *q_order,
                            const int *q_kind, const double *q_scalar, const double *const *q_field,
                            double *coeff_x, double *coeff_y, double *coeff_z,
                            double *q_x, double *q_y, double *q_z)
{
    const size_t n = (size_t)nx * ny * nz;
    const double A = dx * dx, V = pow(dx, 3.0); /* dx**3: Python float power = C pow() */
    const double Ccell = rho * cp * V;
    double *coeff[3] = {coeff_x, coeff_y, coeff_z};
    double *qq[3] = {q_x, q_y, q_z};
    for (int a = 0; a < 3; ++a) {
        memset(coeff[a], 0, n * sizeof(double));
        memset(qq[a], 0, n * sizeof(double));
    }
    uint8_t *exp = (uint8_t *)malloc(n ? n : 1);
    if (!exp) return -2;
    for (int f = 0; f < 6; ++f) {
        if (h_kind[f] == 0) continue;
        oracle_exposed_mask(mask, nx, ny, nz, f, exp);
        double *c = coeff[f >> 1];
        const double *hf = h_kind[f] == 2 ? h_field[f] : NULL;
        const double hs = h_scalar[f];
#pragma omp parallel for num_threads(g_threads)
        for (size_t p = 0; p < n; ++p)
            if (exp[p]) c[p] += ((hf ? hf[p] : hs) * A / Ccell);
    }
    for (int t = 0; t < nq; ++t) {
        const int f = q_order[t];
        if (f < 0 || f > 5) { free(exp); return -1; }
        if (q_kind[f] == 0) continue; /* "if qv is None: continue" :106 */
        oracle_exposed_mask(mask, nx, ny, nz, f, exp);
        double *q = qq[f >> 1];
        const double *qf = q_kind[f] == 2 ? q_field[f] : NULL;
        const double qs = q_scalar[f];
#pragma omp parallel for num_threads(g_threads)
        for (size_t p = 0; p < n; ++p) {
            /* S zero off the exposed set; q_axis += S   (:110-114) */
            double S = exp[p] ? ((qf ? qf[p] : qs) * A / Ccell) : 0.0;
            q[p] += S;
        }
    }
    free(exp);
    return 0;
}

/* adi3d_numba_coeff.py:240-288  lap1D_x / lap1D_y / lap1D_z
 * out = ((sum of ACTIVE axis neighbours, '-' first) - cnt*T) / dx^2 on active cells, 0 elsewhere. */
void oracle_lap1d(const double *T, const uint8_t *mask, int nx, int ny, int nz, double dx,
                  int axis, double *out)
{
    const double invdx2 = 1.0 / (dx * dx);
#pragma omp parallel for collapse(2) num_threads(g_threads)
    for (int i = 0; i < nx; ++i)
        for (int j = 0; j < ny; ++j)
            for (int k = 0; k < nz; ++k) {
                const size_t p = IDX(i, j, k);
                if (!mask[p]) { out[p] = 0.0; continue; }
                double s = 0.0, c = 0.0;
                int lo, hi;
                size_t pm, pp;
                if (axis == 0) { lo = i - 1 >= 0; hi = i + 1 < nx; pm = IDX(i - 1, j, k); pp = IDX(i + 1, j, k); }
                else if (axis == 1) { lo = j - 1 >= 0; hi = j + 1 < ny; pm = IDX(i, j - 1, k); pp = IDX(i, j + 1, k); }
                else { lo = k - 1 >= 0; hi = k + 1 < nz; pm = p - 1; pp = p + 1; }
                if (lo && mask[pm]) { s += T[pm]; c += 1.0; }
                if (hi && mask[pp]) { s += T[pp]; c += 1.0; }
                out[p] = (s - c * T[p]) * invdx2;
            }
}

/* adi3d_numba_coeff.py:121-130  thomas_solve(a,b,c,d,n) -- in-place forward
 * elimination with multiplier m=a[i]/b[i-1], then back substitution with divides. */
static void thomas_solve(const double *a, double *b, const double *c, double *d, double *x, int n)
{
    for (int i = 1; i < n; ++i) {
        double m = a[i] / b[i - 1];
        b[i] = b[i] - m * c[i - 1];
        d[i] = d[i] - m * d[i - 1];
    }
    x[n - 1] = d[n - 1] / b[n - 1];
    for (int i = n - 2; i >= 0; --i) x[i] = (d[i] - c[i] * x[i + 1]) / b[i];
}

/* adi3d_numba_coeff.py:133-237  sweep_axis0 / sweep_axis1 / sweep_axis2
 * out = prev.copy(); per line along `axis`: compress the active cells (:138-148),
 * rows a=-theta*gam*[prev neighbour active], c likewise, b=1+theta*gam*nnb+dt*coeff (:151-155),
 * d=out+dt*qflux+dt*coeff*Tinf (:162); Dirichlet rows a=c=0,b=1,d=dir_val (:157-158);
 * thomas_solve; scatter (:165-166).  Void cells keep prev. */
int oracle_sweep_axis(int axis, const double *prev, const uint8_t *mask, const double *coeff,
                      const uint8_t *dir_mask, const double *dir_val, const double *qflux,
                      int nx, int ny, int nz, double theta, double gam, double dt, double Tinf,
                      double *out)
{
    const size_t ncell = (size_t)nx * ny * nz;
    if (out != prev) memcpy(out, prev, ncell * sizeof(double));
    const int n = axis == 0 ? nx : (axis == 1 ? ny : nz);
    const int n1 = axis == 0 ? ny : nx;             /* outer loop extent */
    const int n2 = axis == 2 ? ny : nz;             /* inner loop extent */
    const size_t stride = axis == 0 ? (size_t)ny * nz : (axis == 1 ? (size_t)nz : 1);
    int fail = 0;
#pragma omp parallel num_threads(g_threads)
    {
        double *buf = (double *)malloc((size_t)(n > 0 ? n : 1) * 5 * sizeof(double));
        int *idx = (int *)malloc((size_t)(n > 0 ? n : 1) * sizeof(int));
        if (!buf || !idx) {
#pragma omp atomic write
            fail = 1;
        } else {
            double *a = buf, *b = buf + n, *c = buf + 2 * n, *d = buf + 3 * n, *x = buf + 4 * n;
#pragma omp for collapse(2)
            for (int u = 0; u < n1; ++u)
                for (int v = 0; v < n2; ++v) {
                    size_t base;
                    if (axis == 0) base = IDX(0, u, v);
                    else if (axis == 1) base = IDX(u, 0, v);
                    else base = IDX(u, v, 0);
                    int cnt = 0;
                    for (int t = 0; t < n; ++t) {
                        const size_t p = base + (size_t)t * stride;
                        if (!mask[p]) continue;
                        idx[cnt] = t;
                        int nnb = 0;
                        double lo = 0.0, hi = 0.0;
                        if (t - 1 >= 0 && mask[p - stride]) { lo = -theta * gam; nnb += 1; }
                        if (t + 1 < n && mask[p + stride]) { hi = -theta * gam; nnb += 1; }
                        double diag = 1.0 + theta * gam * nnb + dt * coeff[p];
                        if (dir_mask[p]) {
                            a[cnt] = 0.0; c[cnt] = 0.0; b[cnt] = 1.0; d[cnt] = dir_val[p];
                        } else {
                            a[cnt] = lo; b[cnt] = diag; c[cnt] = hi;
                            d[cnt] = out[p] + dt * qflux[p] + dt * coeff[p] * Tinf;
                        }
                        ++cnt;
                    }
                    if (cnt == 0) continue;
                    thomas_solve(a, b, c, d, x, cnt);
                    for (int m = 0; m < cnt; ++m) out[base + (size_t)idx[m] * stride] = x[m];
                }
        }
        free(buf);
        free(idx);
    }
    return fail ? -2 : 0;
}

/* adi3d_numba_coeff.py:290-302  adi_step_numba_coeff
 * kappa=k/(rho*cp); gam=kappa*dt/dx^2; R0 = Tn + dt*kappa*(1-theta)*(Lx+Ly+Lz) (:298);
 * U=sweep0(R0,packx); V=sweep1(U,packy); W=sweep2(V,packz).  Tn is not modified.
 * work: caller scratch of 3*ncell doubles (or NULL -> malloc here). */
int oracle_adi_step(const double *Tn, const uint8_t *mask, int nx, int ny, int nz, double dx,
                    double rho, double cp, double k, double dt, double theta, double Tinf,
                    const double *coeff_x, const uint8_t *dirm_x, const double *dirv_x, const double *q_x,
                    const double *coeff_y, const uint8_t *dirm_y, const double *dirv_y, const double *q_y,
                    const double *coeff_z, const uint8_t *dirm_z, const double *dirv_z, const double *q_z,
                    double *Tout, double *work)
{
    const size_t n = (size_t)nx * ny * nz;
    double *own = NULL;
    if (!work) {
        own = (double *)malloc((n ? n : 1) * 3 * sizeof(double));
        if (!own) return -2;
        work = own;
    }
    double *Lx = work, *Ly = work + n, *Lz = work + 2 * n;
    const double kappa = k / (rho * cp);
    const double gam = kappa * dt / (dx * dx);
    oracle_lap1d(Tn, mask, nx, ny, nz, dx, 0, Lx);
    oracle_lap1d(Tn, mask, nx, ny, nz, dx, 1, Ly);
    oracle_lap1d(Tn, mask, nx, ny, nz, dx, 2, Lz);
    const double s = dt * kappa * (1.0 - theta);
#pragma omp parallel for num_threads(g_threads)
    for (size_t p = 0; p < n; ++p) Lx[p] = Tn[p] + s * ((Lx[p] + Ly[p]) + Lz[p]); /* R0 */
    int rc = oracle_sweep_axis(0, Lx, mask, coeff_x, dirm_x, dirv_x, q_x, nx, ny, nz, theta, gam, dt, Tinf, Ly);
    if (!rc) rc = oracle_sweep_axis(1, Ly, mask, coeff_y, dirm_y, dirv_y, q_y, nx, ny, nz, theta, gam, dt, Tinf, Lz);
    if (!rc) rc = oracle_sweep_axis(2, Lz, mask, coeff_z, dirm_z, dirv_z, q_z, nx, ny, nz, theta, gam, dt, Tinf, Tout);
    free(own);
    return rc;
}
