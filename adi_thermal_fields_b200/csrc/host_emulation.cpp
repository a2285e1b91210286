// host_emulation.cpp -- runs the per-thread code of adi_core.h (the same source the
// sm_100a kernels are compiled from) on the CPU, one "thread" after another with the
// shared-memory exchange replaced by plain arrays.  Built only by tests/ (g++), so that
// the partitioned solve can be checked against the oracle without a GPU.  Not shipped in
// libadi_b200.so and never used by the product path.
#include <stdint.h>
#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "adi_core.h"

using namespace adi;

namespace {

int g_uniform = 1;   // take the tabulated uniform-chunk path (adi_core.h) where it applies, as the kernels do

template <int M>
struct HostOps {
    const double *coeff, *qp, *dvp;  // offset to the chunk's first cell; may be null
    size_t stride;
    int nv;
    double f[2 * M];                 // factor store (shared memory on the device)
    double coef(int e) const { return (coeff && e < nv) ? coeff[(size_t)e * stride] : 0.0; }
    double q(int e) const { return (qp && e < nv) ? qp[(size_t)e * stride] : 0.0; }
    double dirv(int e) const { return (dvp && e < nv) ? dvp[(size_t)e * stride] : 0.0; }
    void put2(int e, double la, double u) { f[2 * e] = la; f[2 * e + 1] = u; }
    double la(int e) const { return f[2 * e]; }
    double u(int e) const { return f[2 * e + 1]; }
    void put1(int e, double v) { f[e] = v; }
    double rinv(int e) const { return f[e]; }
};

template <int M, int NS, int CMODE, bool EXTRA>
void sweep_line(double *T, const uint8_t *code, const double *coeff, const double *q,
                const double *dirv, size_t base, size_t stride, int n, unsigned LO, unsigned HI,
                const SweepConst &k, int zmode = 0, Iface *iface = nullptr, double Lg = 0.0, double Rg = 0.0)
{
    const int P = (n + M - 1) / M;
    std::vector<Chunk<M>> ch(P);
    std::vector<HostOps<M>> ops(P);
    std::vector<First> fi(P);
    std::vector<Red> red(P), nxt(P);
    std::vector<char> uni(P, 0);
    std::vector<UniHead> hd(P);
    UniConst uc;
    uni_const_build(uc, k.g);
    for (int p = 0; p < P; ++p) {
        const size_t idx0 = base + (size_t)p * M * stride;
        ops[p].coeff = coeff ? coeff + idx0 : nullptr;
        ops[p].qp = q ? q + idx0 : nullptr;
        ops[p].dvp = dirv ? dirv + idx0 : nullptr;
        ops[p].stride = stride;
        ops[p].nv = std::min(std::max(n - p * M, 0), M);
        for (int e = 0; e < M; ++e) {
            const bool ok = e < ops[p].nv;
            const size_t idx = idx0 + (size_t)e * stride;
            ch[p].set_code(e, ok ? code[idx] : 0u);
            ch[p].T[e] = (ok && (code[idx] & CB_SELF)) ? T[idx] : 0.0;   // load rule
        }
        // the kernels take the uniform paths when the coefficient field is known to vanish away from the
        // surface (CMODE 1 by construction; CMODE 2 after k_check_sparse): here the values themselves are checked
        int v = 0;   // 0: general, 1: OFF 0, 2: OFF 1
        if (g_uniform && !EXTRA && ops[p].nv == M) {
            if (chunk_uniform<M, 0>(ch[p], LO, HI)) v = 1;
            else if (chunk_uniform<M, 1>(ch[p], LO, HI)) v = 2;
            if (v && CMODE == 2)
                for (int e = v - 1; e < M - 1; ++e)
                    if (ops[p].coef(e) != 0.0) v = 0;
        }
        uni[p] = (char)v;
        if (v) {
            const Row sep = make_row<CMODE, EXTRA>(ch[p].code(M - 1), LO, HI, ch[p].T[M - 1], CMODE == 2 ? ops[p].coef(M - 1) : 0.0,
                                                   0.0, 0.0, k);
            const Row head = make_row<CMODE, EXTRA>(ch[p].code(0), LO, HI, ch[p].T[0], CMODE == 2 ? ops[p].coef(0) : 0.0,
                                                    0.0, 0.0, k);
            fi[p] = v == 1 ? chunk_forward_uniform<M, 0>(ch[p], uc, sep, head, hd[p])
                           : chunk_forward_uniform<M, 1>(ch[p], uc, sep, head, hd[p]);
        } else {
            fi[p] = chunk_forward<M, CMODE, EXTRA, NS>(ch[p], ops[p], LO, HI, k);
        }
    }
    if (zmode == 1) {  // z-slab pass 1: separators as affine functions of the ghosts (solve_reduced3)
        std::vector<Red3> r3(P), n3(P);
        for (int p = 0; p < P; ++p)
            r3[p] = reduced_row3(chunk_reduced_row(ch[p], p + 1 < P ? fi[p + 1] : ghost_first()), p, P);
        for (int s = 1; s < P; s <<= 1) {
            for (int p = 0; p < P; ++p) {
                Red3 lo, hi;
                lo.A = lo.C = lo.D = lo.DL = lo.DR = 0.0;
                hi = lo;
                if (p - s >= 0) lo = r3[p - s];
                if (p + s < P) hi = r3[p + s];
                n3[p] = pcr_step3(r3[p], lo, hi);
            }
            r3.swap(n3);
        }
        iface->yf = fma(fi[0].W, r3[0].D, fi[0].Y);
        iface->vf = fma(fi[0].W, r3[0].DL, fi[0].V);
        iface->wf = fi[0].W * r3[0].DR;
        iface->yl = r3[P - 1].D; iface->vl = r3[P - 1].DL; iface->wl = r3[P - 1].DR;
        return;
    }
    for (int p = 0; p < P; ++p) {
        First nx;
        nx.Y = nx.V = nx.W = 0.0;
        if (p + 1 < P) nx = fi[p + 1];
        else if (zmode == 2 || zmode == 3) nx = ghost_first();
        red[p] = chunk_reduced_row(ch[p], nx);
        if (zmode == 2) {
            if (p == 0) red[p].D = fma(-red[p].A, Lg, red[p].D);
            if (p == P - 1) red[p].D = fma(-red[p].C, Rg, red[p].D);
        }
    }
    for (int s = 1; s < P; s <<= 1) {
        for (int p = 0; p < P; ++p) {
            Red lo, hi;
            lo.A = lo.C = lo.D = 0.0;
            hi.A = hi.C = hi.D = 0.0;
            if (p - s >= 0) lo = red[p - s];
            if (p + s < P) hi = red[p + s];
            nxt[p] = pcr_step(red[p], lo, hi);
        }
        red.swap(nxt);
    }
    if (zmode == 3) {  // right-hand-side part of the interface relation only (both ghosts zero)
        iface->yf = fma(fi[0].W, red[0].D, fi[0].Y);
        iface->yl = red[P - 1].D;
        return;
    }
    for (int p = 0; p < P; ++p) {
        const double Slp = p > 0 ? red[p - 1].D : (zmode == 2 ? Lg : 0.0);
        if (uni[p] == 1) chunk_backward_uniform<M, 0>(ch[p], uc, hd[p], Slp, red[p].D);
        else if (uni[p] == 2) chunk_backward_uniform<M, 1>(ch[p], uc, hd[p], Slp, red[p].D);
        else chunk_backward<M, EXTRA, NS>(ch[p], ops[p], LO, HI, k.g, Slp, red[p].D);
        for (int e = 0; e < ops[p].nv; ++e)
            if (ch[p].active(e)) T[base + ((size_t)p * M + e) * stride] = ch[p].T[e];
    }
}

template <int M, int NS>
void sweep_line_any(bool dense, bool extra, double *T, const uint8_t *code, const double *coeff,
                    const double *q, const double *dirv, size_t base, size_t stride, int n, unsigned LO,
                    unsigned HI, const SweepConst &k, int zmode = 0, Iface *iface = nullptr, double Lg = 0.0,
                    double Rg = 0.0)
{
    if (dense) {
        if (extra) sweep_line<M, NS, 2, true>(T, code, coeff, q, dirv, base, stride, n, LO, HI, k, zmode, iface, Lg, Rg);
        else sweep_line<M, NS, 2, false>(T, code, coeff, nullptr, nullptr, base, stride, n, LO, HI, k, zmode, iface, Lg, Rg);
    } else {
        if (extra) sweep_line<M, NS, 1, true>(T, code, nullptr, q, dirv, base, stride, n, LO, HI, k, zmode, iface, Lg, Rg);
        else sweep_line<M, NS, 1, false>(T, code, nullptr, nullptr, nullptr, base, stride, n, LO, HI, k, zmode, iface, Lg, Rg);
    }
}

}  // namespace

extern "C" {

void emu_set_uniform(int on) { g_uniform = on; }

// code as built by k_build_code (adi_cart.cuh)
void emu_build_code(const uint8_t *mask, const uint8_t *dirm, uint8_t *code, int nx, int ny, int nz)
{
    const size_t snx = (size_t)ny * nz;
    for (int i = 0; i < nx; ++i)
        for (int j = 0; j < ny; ++j)
            for (int k = 0; k < nz; ++k) {
                const size_t idx = ((size_t)i * ny + j) * nz + k;
                unsigned c = 0;
                if (mask[idx]) {
                    c = CB_SELF;
                    if (dirm && dirm[idx]) c |= CB_DIR;
                    if (i > 0 && mask[idx - snx]) c |= CB_XM;
                    if (i + 1 < nx && mask[idx + snx]) c |= CB_XP;
                    if (j > 0 && mask[idx - nz]) c |= CB_YM;
                    if (j + 1 < ny && mask[idx + nz]) c |= CB_YP;
                    if (k > 0 && mask[idx - 1]) c |= CB_ZM;
                    if (k + 1 < nz && mask[idx + 1]) c |= CB_ZP;
                }
                code[idx] = (uint8_t)c;
            }
}

// the same with the mask planes of the adjacent z slabs (adi_cart_set_mask_halo)
void emu_build_code_halo(const uint8_t *mask, const uint8_t *dirm, uint8_t *code, int nx, int ny, int nz,
                         const uint8_t *mlo, const uint8_t *mhi)
{
    emu_build_code(mask, dirm, code, nx, ny, nz);
    if (!nz) return;
    for (size_t ij = 0; ij < (size_t)nx * ny; ++ij) {
        if (mlo && mlo[ij] && mask[ij * nz]) code[ij * nz] |= CB_ZM;
        if (mhi && mhi[ij] && mask[ij * nz + nz - 1]) code[ij * nz + nz - 1] |= CB_ZP;
    }
}

// z-slab phases with the operand conventions of adi_cart_step_xy / adi_cart_zsweep_reduce /
// adi_cart_zsweep_finish (adi_b200.h).  phase 0: explicit stage + x + y sweeps, Tin -> Tout;
// phase 1: interface relations of the z lines of Tout -> iface_dyn[2][nx*ny], iface_stat[4][nx*ny];
// phase 3: the right-hand-side part iface_dyn only;
// phase 2: inter-rank solve from dyn_all[nranks][2][nx*ny], stat_all[nranks][4][nx*ny] + local finish in place.
// Solve-first form (adi_cart_zsweep_solve0 / _spike / _apply):
// phase 4: local finish in place with both ghosts at zero; iface_dyn = (first, last) value of every segment;
// phase 5 / 6: the homogeneous system (no flux, Dirichlet value 0, ambient 0) with ghost 1 at the lower /
//              upper end, in place on the caller's zeroed field: the unit-ghost response;
// phase 7: inter-rank solve only: iface_dyn[2][nx*ny] = (L, R) ghosts of this rank.
int emu_cart_slab(const double *Tin, double *Tout, const uint8_t *mask, int nx, int ny, int nz,
                  double dx, double dt, double theta, double kappa, double Tinf,
                  const double *const coeff[3], const uint8_t *const dirm[3],
                  const double *const dirv[3], const double *const q[3], const double *face_coeff,
                  int variant, const uint8_t *mlo, const uint8_t *mhi, const double *Tlo, const double *Thi,
                  int phase, double *iface_dyn, double *iface_stat, const double *dyn_all, const double *stat_all,
                  int rank, int nranks)
{
    const size_t n = (size_t)nx * ny * nz, nlines = (size_t)nx * ny;
    std::vector<uint8_t> code(n ? n : 1);
    SweepConst k;
    const double gam = kappa * dt / (dx * dx);
    k.g = theta * gam;
    k.dt = dt;
    k.Tinf = Tinf;
    k.beta = dt * kappa * (1.0 - theta);
    k.invdx2 = 1.0 / (dx * dx);
    const size_t snx = (size_t)ny * nz;
    const int M = variant == 0 ? 16 : 32;
    if (phase == 0) {
        emu_build_code_halo(mask, dirm[0], code.data(), nx, ny, nz, mlo, mhi);
        for (size_t idx = 0; idx < n; ++idx) {
            const unsigned c = code[idx];
            const size_t ij = idx / nz;
            const int kz = (int)(idx % nz);
            double v[6] = {0, 0, 0, 0, 0, 0};
            if (c & CB_XM) v[0] = Tin[idx - snx];
            if (c & CB_XP) v[1] = Tin[idx + snx];
            if (c & CB_YM) v[2] = Tin[idx - nz];
            if (c & CB_YP) v[3] = Tin[idx + nz];
            if (c & CB_ZM) v[4] = kz > 0 ? Tin[idx - 1] : Tlo[ij];
            if (c & CB_ZP) v[5] = kz + 1 < nz ? Tin[idx + 1] : Thi[ij];
            Tout[idx] = (k.beta != 0.0 && (c & CB_SELF))
                            ? explicit_r0(c, Tin[idx], v[0], v[1], v[2], v[3], v[4], v[5], k)
                            : Tin[idx];
        }
    }
    if (phase == 7) {
        for (size_t line = 0; line < nlines; ++line) {
            double Lg = 0.0, Rg = 0.0;
            iface_solve([&](int r) {
                const double *dd = dyn_all + (size_t)r * 2 * nlines + line;
                const double *qq = stat_all + (size_t)r * 4 * nlines + line;
                Iface w;
                w.yf = dd[0]; w.yl = dd[nlines];
                w.vf = qq[0]; w.wf = qq[nlines]; w.vl = qq[2 * nlines]; w.wl = qq[3 * nlines];
                return w;
            }, nranks, rank, &Lg, &Rg);
            iface_dyn[line] = Lg; iface_dyn[nlines + line] = Rg;
        }
        return 0;
    }
    const bool homogeneous = phase == 5 || phase == 6;
    if (homogeneous) k.Tinf = 0.0;
    const int a0 = phase == 0 ? 0 : 2, a1 = phase == 0 ? 1 : 2;
    for (int axis = a0; axis <= a1; ++axis) {
        if (axis > 0 || phase != 0) emu_build_code_halo(mask, dirm[axis], code.data(), nx, ny, nz, mlo, mhi);
        k.h_lo = face_coeff ? face_coeff[2 * axis] : 0.0;
        k.h_hi = face_coeff ? face_coeff[2 * axis + 1] : 0.0;
        const int len = axis == 0 ? nx : (axis == 1 ? ny : nz);
        const size_t stride = axis == 0 ? snx : (axis == 1 ? (size_t)nz : 1);
        const unsigned LO = axis == 0 ? CB_XM : (axis == 1 ? CB_YM : CB_ZM);
        const unsigned HI = axis == 0 ? CB_XP : (axis == 1 ? CB_YP : CB_ZP);
        const int n1 = axis == 0 ? ny : nx, n2 = axis == 2 ? ny : nz;
        const bool dense = coeff[axis] != nullptr;
        const bool extra = q[axis] != nullptr || dirm[axis] != nullptr;
        if (axis == 2 && len % M != 0) return -2;
        for (int u = 0; u < n1; ++u)
            for (int v = 0; v < n2; ++v) {
                size_t base;
                if (axis == 0) base = (size_t)u * nz + v;
                else if (axis == 1) base = (size_t)u * snx + v;
                else base = ((size_t)u * ny + v) * nz;
                int zmode = 0;
                Iface f;
                double Lg = 0.0, Rg = 0.0;
                const size_t line = (size_t)u * ny + v;
                if (axis == 2) {
                    zmode = phase >= 4 ? 2 : phase;
                    if (phase == 5) Lg = 1.0;
                    if (phase == 6) Rg = 1.0;
                    if (phase == 2)
                        iface_solve([&](int r) {
                            const double *dd = dyn_all + (size_t)r * 2 * nlines + line;
                            const double *qq = stat_all + (size_t)r * 4 * nlines + line;
                            Iface w;
                            w.yf = dd[0]; w.yl = dd[nlines];
                            w.vf = qq[0]; w.wf = qq[nlines]; w.vl = qq[2 * nlines]; w.wl = qq[3 * nlines];
                            return w;
                        }, nranks, rank, &Lg, &Rg);
                }
                const double *qa = homogeneous ? nullptr : q[axis], *dva = homogeneous ? nullptr : dirv[axis];
                if (variant == 0)
                    sweep_line_any<16, 2>(dense, extra, Tout, code.data(), coeff[axis], qa, dva, base,
                                          stride, len, LO, HI, k, zmode, &f, Lg, Rg);
                else
                    sweep_line_any<32, 1>(dense, extra, Tout, code.data(), coeff[axis], qa, dva, base,
                                          stride, len, LO, HI, k, zmode, &f, Lg, Rg);
                if (axis == 2 && phase == 4) {  // void end cells count as 0 (load rule of the kernels)
                    iface_dyn[line] = (code[base] & CB_SELF) ? Tout[base] : 0.0;
                    iface_dyn[nlines + line] = (code[base + nz - 1] & CB_SELF) ? Tout[base + nz - 1] : 0.0;
                }
                if (axis == 2 && (phase == 1 || phase == 3)) {
                    iface_dyn[line] = f.yf; iface_dyn[nlines + line] = f.yl;
                    if (phase == 1) {
                        iface_stat[line] = f.vf; iface_stat[nlines + line] = f.wf;
                        iface_stat[2 * nlines + line] = f.vl; iface_stat[3 * nlines + line] = f.wl;
                    }
                }
            }
    }
    return 0;
}

// One full step with the operand conventions of adi_cart_step (adi_b200.h).
// coeff[a]/q[a]/dirm[a]/dirv[a] may be NULL; face_coeff != NULL selects the scalar Robin mode.
// variant: 0 = M 16 / two factors per cell, 1 = M 32 / one factor per cell (adi_launch.h)
int emu_cart_step(const double *Tin, double *Tout, const uint8_t *mask, int nx, int ny, int nz,
                  double dx, double dt, double theta, double kappa, double Tinf,
                  const double *const coeff[3], const uint8_t *const dirm[3],
                  const double *const dirv[3], const double *const q[3], const double *face_coeff,
                  int variant)
{
    const size_t n = (size_t)nx * ny * nz;
    std::vector<uint8_t> code(n ? n : 1);
    SweepConst k;
    const double gam = kappa * dt / (dx * dx);
    k.g = theta * gam;
    k.dt = dt;
    k.Tinf = Tinf;
    k.beta = dt * kappa * (1.0 - theta);
    k.invdx2 = 1.0 / (dx * dx);
    const size_t snx = (size_t)ny * nz;
    // explicit stage fused in front of the x sweep
    emu_build_code(mask, dirm[0], code.data(), nx, ny, nz);
    for (size_t idx = 0; idx < n; ++idx) {
        const unsigned c = code[idx];
        double v[6] = {0, 0, 0, 0, 0, 0};
        if (c & CB_XM) v[0] = Tin[idx - snx];
        if (c & CB_XP) v[1] = Tin[idx + snx];
        if (c & CB_YM) v[2] = Tin[idx - nz];
        if (c & CB_YP) v[3] = Tin[idx + nz];
        if (c & CB_ZM) v[4] = Tin[idx - 1];
        if (c & CB_ZP) v[5] = Tin[idx + 1];
        Tout[idx] = (k.beta != 0.0 && (c & CB_SELF))
                        ? explicit_r0(c, Tin[idx], v[0], v[1], v[2], v[3], v[4], v[5], k)
                        : Tin[idx];
    }
    for (int axis = 0; axis < 3; ++axis) {
        if (axis > 0) emu_build_code(mask, dirm[axis], code.data(), nx, ny, nz);
        k.h_lo = face_coeff ? face_coeff[2 * axis] : 0.0;
        k.h_hi = face_coeff ? face_coeff[2 * axis + 1] : 0.0;
        const int len = axis == 0 ? nx : (axis == 1 ? ny : nz);
        const size_t stride = axis == 0 ? snx : (axis == 1 ? (size_t)nz : 1);
        const unsigned LO = axis == 0 ? CB_XM : (axis == 1 ? CB_YM : CB_ZM);
        const unsigned HI = axis == 0 ? CB_XP : (axis == 1 ? CB_YP : CB_ZP);
        const int n1 = axis == 0 ? ny : nx, n2 = axis == 2 ? ny : nz;
        const bool dense = coeff[axis] != nullptr;
        const bool extra = q[axis] != nullptr || dirm[axis] != nullptr;
        for (int u = 0; u < n1; ++u)
            for (int v = 0; v < n2; ++v) {
                size_t base;
                if (axis == 0) base = (size_t)u * nz + v;
                else if (axis == 1) base = (size_t)u * snx + v;
                else base = ((size_t)u * ny + v) * nz;
                if (variant == 0)
                    sweep_line_any<16, 2>(dense, extra, Tout, code.data(), coeff[axis], q[axis], dirv[axis], base,
                                          stride, len, LO, HI, k);
                else
                    sweep_line_any<32, 1>(dense, extra, Tout, code.data(), coeff[axis], q[axis], dirv[axis], base,
                                          stride, len, LO, HI, k);
            }
    }
    return 0;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------
// Cylindrical path: the table-driven solve of adi_tab_core.h, one "thread" after another.
// ---------------------------------------------------------------------------------------
#include "adi_tab_core.h"

namespace {

// zm 0: whole line.  zm 1: z-slab pass 1 -- y[0] = yf, y[1] = yl of the segment, nothing stored.
// zm 2: z-slab pass 2 with the ghost values Lg, Rg.
template <int M>
void tab_line(double *T, size_t base, size_t stride, const TabGeom &g, const int *geom, const double *blob,
              int set_first, double val_first, int set_last, double val_last, int zm = 0, double *y = nullptr,
              double Lg = 0.0, double Rg = 0.0)
{
    const int P = g.P;
    const int *cbase = geom, *end = geom + P, *len = geom + 2 * P;
    std::vector<double> d((size_t)P * M), Y(P), Yl(P), D(P), Dn(P);
    for (int p = 0; p < P; ++p) {
        double (&dd)[M] = *reinterpret_cast<double (*)[M]>(&d[(size_t)p * M]);
        for (int e = 0; e < M; ++e) {
            const int i = end[p] - (M - 1 - e);
            double v = 0.0;
            if (e >= M - len[p]) {
                v = T[base + (size_t)i * stride];
                if (i == 0) v = set_first ? val_first : v + val_first;
                if (i == g.n - 1) v = set_last ? val_last : v + val_last;
            }
            dd[e] = v;
        }
        Y[p] = tab_forward<M>(dd, blob + g.o_f + 2 * cbase[p], blob + g.o_alpha + cbase[p], &Yl[p]);
    }
    for (int p = 0; p < P; ++p)
        D[p] = tab_reduced_rhs(blob[g.o_t0 + p], blob[g.o_t1 + p], blob[g.o_t2 + p], d[(size_t)p * M + M - 1], Yl[p],
                               Y[tab_hi(p, 1, P, g.cyclic)]);
    for (int l = 0; l < g.levels; ++l) {
        const int s = 1 << l;
        const double *R = blob + g.o_lvl + (size_t)l * 3 * P;
        for (int p = 0; p < P; ++p)
            Dn[p] = tab_level(R[p], R[P + p], R[2 * P + p], D[p], D[tab_lo(p, s, P, g.cyclic)],
                              D[tab_hi(p, s, P, g.cyclic)]);
        D.swap(Dn);
    }
    if (zm == 1) {
        y[0] = fma(blob[g.o_misc], D[0], Y[0]);     // x_first = Y_0 + W_0*S_0 (+ ghost terms)
        y[1] = D[P - 1];
        return;
    }
    if (zm == 2)
        for (int p = 0; p < P; ++p) D[p] = fma(blob[g.o_dr + p], Rg, fma(blob[g.o_dl + p], Lg, D[p]));
    for (int p = 0; p < P; ++p) {
        double (&dd)[M] = *reinterpret_cast<double (*)[M]>(&d[(size_t)p * M]);
        tab_backward<M>(dd, blob + g.o_b + 2 * cbase[p], (zm == 2 && p == 0) ? Lg : D[tab_lo(p, 1, P, g.cyclic)], D[p]);
        for (int e = 0; e < M; ++e) {
            const int i = end[p] - (M - 1 - e);
            if (e >= M - len[p]) T[base + (size_t)i * stride] = dd[e];
        }
    }
}

void tab_line_any(int M, double *T, size_t base, size_t stride, const TabGeom &g, const int *geom,
                  const double *blob, int sf, double vf, int sl, double vl, int zm = 0, double *y = nullptr,
                  double Lg = 0.0, double Rg = 0.0)
{
    if (M == 16) tab_line<16>(T, base, stride, g, geom, blob, sf, vf, sl, vl, zm, y, Lg, Rg);
    else if (M == 8) tab_line<8>(T, base, stride, g, geom, blob, sf, vf, sl, vl, zm, y, Lg, Rg);
    else if (M == 4) tab_line<4>(T, base, stride, g, geom, blob, sf, vf, sl, vl, zm, y, Lg, Rg);
    else tab_line<32>(T, base, stride, g, geom, blob, sf, vf, sl, vl, zm, y, Lg, Rg);
}

}  // namespace

extern "C" {

// prm: dt, rho, cp, k, h_r, Tinf_r, h_bot, h_top, Tinf_bot, Tinf_top, T_bot, T_top, T_void, T_inner
int emu_cyl_step_slab(const double *Tin, double *Tout, int nr, int nphi, int nz, double dr, double dphi, double dz,
                      const double *prm, int kind_bot, int kind_top, const uint8_t *active, const double *S, int M,
                      int nslab);

int emu_cyl_step(const double *Tin, double *Tout, int nr, int nphi, int nz, double dr, double dphi, double dz,
                 const double *prm, int kind_bot, int kind_top, const uint8_t *active, const double *S, int M)
{
    return emu_cyl_step_slab(Tin, Tout, nr, nphi, nz, dr, dphi, dz, prm, kind_bot, kind_top, active, S, M, 1);
}

// nslab > 1: the z sweep runs as nslab segments per line (the multi-GPU z-slab algorithm, all "ranks" here)
int emu_cyl_step_slab(const double *Tin, double *Tout, int nr, int nphi, int nz, double dr, double dphi, double dz,
                      const double *prm, int kind_bot, int kind_top, const uint8_t *active, const double *S, int M,
                      int nslab)
{
    const double dt = prm[0], rho = prm[1], cp = prm[2], k = prm[3], h_r = prm[4], Tinf_r = prm[5];
    const double alpha = k / (rho * cp);
    const size_t ncell = (size_t)nr * nphi * nz;
    for (size_t g = 0; g < ncell; ++g) {
        double v = Tin[g];
        if (active && !active[g]) v = prm[12];
        if (S) v = v + dt * (S[g] / (rho * cp));
        Tout[g] = v;
    }
    {
        const TabGeom g = tab_geom(nr, M, false, false);
        std::vector<int> geom(3 * g.P);
        std::vector<double> blob(g.ndbl), a(nr), b(nr), c(nr);
        const double add = cyl_rows_r(nr, dr, alpha, k, dt, h_r, Tinf_r, a.data(), b.data(), c.data());
        tab_partition(g, false, geom.data(), geom.data() + g.P, geom.data() + 2 * g.P);
        tab_build(g, geom.data(), geom.data() + g.P, geom.data() + 2 * g.P, a.data(), b.data(), c.data(), blob.data());
        for (int j = 0; j < nphi; ++j)
            for (int kz = 0; kz < nz; ++kz)
                tab_line_any(M, Tout, (size_t)j * nz + kz, (size_t)nphi * nz, g, geom.data(), blob.data(), 0, 0.0, 0,
                             h_r != 0.0 ? add : 0.0);
    }
    if (nphi > 1) {
        const TabGeom g = tab_geom(nphi, M, true, true);
        std::vector<int> geom(3 * g.P);
        std::vector<double> blob(g.ndbl), a(nphi), b(nphi), c(nphi);
        tab_partition(g, true, geom.data(), geom.data() + g.P, geom.data() + 2 * g.P);
        for (int ir = 0; ir < nr; ++ir) {
            const double f = cyl_fac_phi(ir, dr, dphi, alpha, dt);
            for (int j = 0; j < nphi; ++j) { a[j] = -f; b[j] = 1.0 + 2.0 * f; c[j] = -f; }
            tab_build(g, geom.data(), geom.data() + g.P, geom.data() + 2 * g.P, a.data(), b.data(), c.data(), blob.data());
            for (int kz = 0; kz < nz; ++kz)
                tab_line_any(M, Tout, (size_t)ir * nphi * nz + kz, (size_t)nz, g, geom.data(), blob.data(), 0, 0.0, 0, 0.0);
        }
    }
    if (nslab <= 1) {
        const TabGeom g = tab_geom(nz, M, false, false);
        std::vector<int> geom(3 * g.P);
        std::vector<double> blob(g.ndbl), a(nz), b(nz), c(nz);
        ZEnd bot, top;
        if (cyl_rows_z(nz, dz, alpha, k, dt, kind_bot, kind_top, prm[6], prm[7], prm[8], prm[9], prm[10], prm[11],
                       a.data(), b.data(), c.data(), &bot, &top))
            return -1;
        tab_partition(g, false, geom.data(), geom.data() + g.P, geom.data() + 2 * g.P);
        tab_build(g, geom.data(), geom.data() + g.P, geom.data() + 2 * g.P, a.data(), b.data(), c.data(), blob.data());
        for (size_t line = 0; line < (size_t)nr * nphi; ++line)
            tab_line_any(M, Tout, line * nz, 1, g, geom.data(), blob.data(), bot.set, bot.val, top.set, top.val);
    } else {
        // z-slab decomposition: nslab segments per line, each with its own tables (ghost couplings at the
        // inner ends), pass 1 -> inter-segment solve -> pass 2, as the ranks of the multi-GPU path do
        std::vector<TabSet> ts(nslab);
        std::vector<ZEnd> bots(nslab), tops(nslab);
        std::vector<int> z0(nslab + 1, 0);
        for (int q = 0; q < nslab; ++q) z0[q + 1] = z0[q] + nz / nslab + (q < nz % nslab ? 1 : 0);
        std::vector<Iface> cst(nslab);
        for (int q = 0; q < nslab; ++q) {
            const int nl = z0[q + 1] - z0[q];
            std::vector<double> a(nl), b(nl), c(nl);
            if (cyl_rows_z_segment(nl, q == 0, q == nslab - 1, dz, alpha, k, dt, kind_bot, kind_top, prm[6], prm[7], prm[8],
                                   prm[9], prm[10], prm[11], a.data(), b.data(), c.data(), &bots[q], &tops[q]))
                return -1;
            ts[q] = tab_make(nl, M, false, a.data(), b.data(), c.data(), q > 0, q < nslab - 1);
            const TabGeom &g = ts[q].g;
            const double *bl = ts[q].blob.data();
            cst[q].yf = cst[q].yl = 0.0;
            cst[q].vf = bl[g.o_misc + 1] + bl[g.o_misc] * bl[g.o_dl];      // V_0 + W_0*dl_0
            cst[q].wf = bl[g.o_misc] * bl[g.o_dr];
            cst[q].vl = bl[g.o_dl + g.P - 1];
            cst[q].wl = bl[g.o_dr + g.P - 1];
        }
        for (size_t line = 0; line < (size_t)nr * nphi; ++line) {
            std::vector<Iface> rel(cst);
            for (int q = 0; q < nslab; ++q) {
                double y[2];
                tab_line_any(M, Tout, line * nz + z0[q], 1, ts[q].g, ts[q].geom.data(), ts[q].blob.data(), bots[q].set,
                             bots[q].val, tops[q].set, tops[q].val, 1, y);
                rel[q].yf = y[0]; rel[q].yl = y[1];
            }
            for (int q = 0; q < nslab; ++q) {
                double Lg, Rg;
                iface_solve([&](int r) { return rel[r]; }, nslab, q, &Lg, &Rg);
                tab_line_any(M, Tout, line * nz + z0[q], 1, ts[q].g, ts[q].geom.data(), ts[q].blob.data(), bots[q].set,
                             bots[q].val, tops[q].set, tops[q].val, 2, nullptr, Lg, Rg);
            }
        }
    }
    if (active)
        for (size_t g = 0; g < ncell; ++g)
            if (!active[g]) Tout[g] = (g / ((size_t)nphi * nz) == 0) ? prm[13] : prm[12];
    return 0;
}

}  // extern "C"

// ---- ASCII output path: adi_fmt_core.h on the CPU ---------------------------------------------
#include "adi_fmt_core.h"

static const double h_pow10[][2] = {ADI_POW10_TABLE};

extern "C" {

// formats n values; out has 16 bytes per value (NUL padded), lens[i] = text length
void emu_format_values(const double *v, long n, int fmt, char *out, int *lens)
{
    for (long i = 0; i < n; ++i) {
        char *o = out + 16 * i;
        memset(o, 0, 16);
        lens[i] = fmt == adifmt::FMT_E6 ? adifmt::format_value<adifmt::FMT_E6>(v[i], &h_pow10[0][0], o)
                                        : adifmt::format_value<adifmt::FMT_G6>(v[i], &h_pow10[0][0], o);
    }
}

// exact comparison on its own (tests drive it over the full exponent range)
int emu_exact_cmp_half(double a, unsigned n, int q) { return adifmt::exact_cmp_half(a, n, q); }

// the whole data section of a field, in the kernel's order; returns the byte count
long emu_text_field(const void *src, int dtype, int nx, int ny, int nz, int fmt, char *out)
{
    const uint64_t N = (uint64_t)nx * ny * nz;
    char *p = out;
    for (int k = 0; k < nz; ++k)
        for (int j = 0; j < ny; ++j)
            for (int i = 0; i < nx; ++i) {
                const size_t c = ((size_t)i * ny + j) * nz + k;
                const double v = dtype == 0 ? ((const double *)src)[c]
                                 : dtype == 1 ? (double)((const float *)src)[c]
                                              : (double)((const uint8_t *)src)[c];
                p += fmt == adifmt::FMT_E6 ? adifmt::format_value<adifmt::FMT_E6>(v, &h_pow10[0][0], p)
                                           : adifmt::format_value<adifmt::FMT_G6>(v, &h_pow10[0][0], p);
                *p++ = adifmt::separator(fmt, ((uint64_t)k * ny + j) * nx + i, i, nx, N);
            }
    return (long)(p - out);
}

}  // extern "C"

// ---- word forms of the per-mask-change kernels (adi_mask_core.h), "thread" after "thread" ----------------------
#include "adi_mask_core.h"

extern "C" {

// k_build_code_v: nz % 16 == 0
// returns z + 1 of the highest active cell (0: none), as the kernel's reduction does
int emu_build_code_v(const uint8_t *mask, const uint8_t *dirm, uint8_t *code, int nx, int ny, int nz,
                     const uint8_t *mlo, const uint8_t *mhi)
{
    const size_t n16 = (size_t)nx * ny * nz / 16;
    int top = 0;
    for (size_t t = 0; t < n16; ++t) top = std::max(top, build_code16(mask, dirm, code, t * 16, nx, ny, nz, mlo, mhi));
    return top;
}

// k_transpose_code_v with its launch geometry: grid (ceil(nz/128), ceil(n/128), batch), 256 threads, two phases
// around the barrier.  dst holds batch*nz*npad bytes (pre-filled by the caller, as cudaMemset does).
void emu_transpose_code_v(const uint8_t *src, uint8_t *dst, int n, int nz, int npad, int batch, size_t sb, size_t sr)
{
    TrArgs a;
    a.src = src; a.dst = dst; a.n = n; a.nz = nz; a.npad = npad; a.sb = sb; a.sr = sr;
    std::vector<uint32_t> S(128 * 32);
    for (int b = 0; b < batch; ++b)
        for (int by = 0; by < (n + 127) / 128; ++by)
            for (int bx = 0; bx < (nz + 127) / 128; ++bx) {
                std::fill(S.begin(), S.end(), 0xdeadbeefu);
                for (int tid = 0; tid < 256; ++tid) tr_load(a, S.data(), tid, bx * 128, by * 128, b);
                for (int tid = 0; tid < 256; ++tid) tr_store(a, S.data(), tid, bx * 128, by * 128, b);
            }
}

// k_build_packs_v<nc>: nz % nc == 0, nc = 2 | 4.  kinds: 0 none, 1 scalar, 2 field (adi_cart_build_packs).
void emu_build_packs_v(int nc, const uint8_t *mask, int nx, int ny, int nz, const uint8_t *mlo, const uint8_t *mhi, double dx,
                       double rho, double cp, const int *h_kind, const double *h_scalar, const double *const *h_field,
                       const int *q_kind, const double *q_scalar, const double *const *q_field, double *const *coeff,
                       double *const *qout)
{
    PackArgs a;
    a.mask = mask; a.nx = nx; a.ny = ny; a.nz = nz; a.mlo = mlo; a.mhi = mhi;
    a.A = dx * dx;
    a.Ccell = rho * cp * std::pow(dx, 3.0);
    for (int f = 0; f < 6; ++f) {
        a.h_kind[f] = h_kind[f]; a.h_scalar[f] = h_scalar[f]; a.h_field[f] = h_field[f];
        a.q_kind[f] = q_kind[f]; a.q_scalar[f] = q_scalar[f]; a.q_field[f] = q_field[f];
    }
    for (int ax = 0; ax < 3; ++ax) { a.coeff[ax] = coeff[ax]; a.qout[ax] = qout[ax]; }
    const size_t nw = (size_t)nx * ny * nz / nc;
    for (size_t t = 0; t < nw; ++t) {
        if (nc == 4) build_packs_cells<4, false>(a, t * 4, ldcells<4>(mask + t * 4));
        else build_packs_cells<2, false>(a, t * 2, ldcells<2>(mask + t * 2));
    }
}

}  // extern "C"
