// adi_tab_core.h -- table-driven partitioned tridiagonal line solve (cylindrical path).
//
// The cylindrical backward-Euler step (adi3d_cyl_phi_v3.py:332-350) solves, along r, phi and z,
// tridiagonal systems whose coefficients do not depend on the line: build_coeff_r (:155-202)
// depends on the radial index only, build_coeff_z (:255-298) on the z index only, and the
// periodic phi operator (:302-329) on the ring only.  Everything that involves the matrix is
// therefore computed ONCE on the host, in double precision with true divisions, and the
// kernels stream only right-hand sides:
//
//   a line of n cells is cut into P chunks of at most M cells (balanced lengths, the cells
//   right-aligned in the M register slots of a thread, so the chunk's last cell -- its
//   separator S_p -- always sits in slot M-1);
//   phase 1  forward elimination of the interior with the tabulated factors
//            (dp_e = la_e*dp_{e-1} + d_e*rinv_e) and one running sum Y = sum alpha_e*dp_e
//            (first interior cell as a function of the separators: x_0 = Y + V*S_{p-1} + W*S_p);
//   phase 2  reduced system over the P separators: D_p = t0*d_s + t1*Yl + t2*Y_{p+1}, then
//            ceil(log2 P) parallel-cyclic-reduction levels D <- r*D + a*D_lo + c*D_hi whose
//            coefficients are tabulated per level; for the periodic phi lines P is a power of
//            two, neighbours wrap around, and the last level folds the two (identical)
//            neighbours -- no Sherman-Morrison correction is needed;
//   phase 3  back substitution x_e = u_e*x_{e+1} + (dp_e + v_e*S_{p-1}).
//
// Per cell: 1 mul + 4 fma and five table reads; each cell is read once and written once.
//
// The same functions run on the CPU in tests/ (csrc/host_emulation.cpp) against the oracle.
#pragma once
#include <stdint.h>

#include <vector>

#if !defined(__CUDACC__)
#include <cmath>
using std::fma;
#endif

#ifndef ADI_HD
#if defined(__CUDACC__)
#define ADI_HD __host__ __device__ __forceinline__
#else
#define ADI_HD inline
#endif
#endif

namespace adi {

// Geometry and blob layout of one table set (identical for every ring of the phi sweep).
struct TabGeom {
    int n, M, P, levels, cyclic, nvar;
    // offsets into the double blob (all even: the pair tables are read as 16-byte words)
    int o_f;      // cell table, forward pass:  {rinv, la} pairs, 2*nvar*M doubles
    int o_alpha;  // cell table, forward pass:  alpha, nvar*M doubles
    int o_b;      // cell table, backward pass: {u, v} pairs, 2*nvar*M doubles
    int o_t0, o_t1, o_t2;                 // reduced right-hand side assembly, P each
    int o_lvl;                            // levels * 3 * P: [level][r|a|c][p]
    int o_dl, o_dr;                       // z-slab segments: response of separator p to the ghost values, P each
    int o_misc;                           // [0] W of chunk 0, [1] V of chunk 0 (first-cell relation), 2 doubles
    int ndbl;                             // doubles per blob
    // int blob: cbase[P] (offset of the chunk's cell tables), end[P] (index of the
    // separator along the line), len[P]
};

inline int tab_pick_P(int n, int M, bool cyclic)
{
    int P = (n + M - 1) / M;
    if (P < 1) P = 1;
    if (cyclic) {
        int q = 2;
        while (q < P) q <<= 1;
        P = q;
    }
    return P;
}

inline TabGeom tab_geom_nvar(int n, int M, bool cyclic, int nvar);

inline TabGeom tab_geom(int n, int M, bool cyclic, bool uniform)
{
    const int P = tab_pick_P(n, M, cyclic);
    // uniform (translation-invariant) lines: chunks of equal length share their cell tables
    return tab_geom_nvar(n, M, cyclic, uniform ? ((n % P) ? 2 : 1) : P);
}

inline TabGeom tab_geom_nvar(int n, int M, bool cyclic, int nvar)
{
    TabGeom g;
    g.n = n; g.M = M; g.cyclic = cyclic ? 1 : 0;
    g.P = tab_pick_P(n, M, cyclic);
    g.levels = 0;
    if (cyclic) {
        while ((1 << g.levels) < g.P) g.levels++;
    } else {
        while ((1 << g.levels) < g.P) g.levels++;
    }
    g.nvar = nvar;
    int o = 0;
    g.o_f = o; o += 2 * g.nvar * M;
    g.o_b = o; o += 2 * g.nvar * M;
    g.o_alpha = o; o += g.nvar * M;
    g.o_t0 = o; o += g.P;
    g.o_t1 = o; o += g.P;
    g.o_t2 = o; o += g.P;
    g.o_lvl = o; o += g.levels * 3 * g.P;
    g.o_dl = o; o += g.P;
    g.o_dr = o; o += g.P;
    g.o_misc = o; o += 2;
    g.ndbl = (o + 1) & ~1;   // even, so that consecutive blobs stay 16-byte aligned
    return g;
}

// Chunk partition: balanced lengths, every chunk holds at least one cell (needs n >= P).
inline void tab_partition(const TabGeom &g, bool uniform, int *cbase, int *end, int *len)
{
    const int q = g.n / g.P, r = g.n % g.P;
    int start = 0;
    for (int p = 0; p < g.P; ++p) {
        len[p] = q + (p < r ? 1 : 0);
        end[p] = start + len[p] - 1;
        start += len[p];
        cbase[p] = uniform ? ((r && len[p] == q) ? g.M : 0) : p * g.M;
    }
}

// Host: fill one blob from the line's rows  a[i]*x[i-1] + b[i]*x[i] + c[i]*x[i+1] = d[i].
// Non-cyclic: a[0] and c[n-1] are ignored.  Cyclic: a[0] couples to x[n-1], c[n-1] to x[0].
// ghost_lo / ghost_hi (non-cyclic lines only): the line is a SEGMENT of a longer line (z-slab
// decomposition); a[0] then couples to the ghost value L before the segment and c[n-1] to the ghost R
// after it.  The tables solve for L = R = 0 and o_dl / o_dr hold the separators' response to the ghosts:
// S_p = D_p + dl_p*L + dr_p*R.
inline void tab_build(const TabGeom &g, const int *cbase, const int *end, const int *len,
                      const double *a, const double *b, const double *c, double *blob,
                      bool ghost_lo = false, bool ghost_hi = false)
{
    const int M = g.M, P = g.P, n = g.n;
    for (int i = 0; i < g.ndbl; ++i) blob[i] = 0.0;
    for (int v = 0; v < g.nvar * M; ++v) blob[g.o_f + 2 * v] = 1.0;   // rinv of padding slots
    std::vector<double> V(P), W(P), Vl(P), Wl(P), A(P), C(P), An(P), Cn(P);
    for (int p = 0; p < P; ++p) {
        const int e0 = M - len[p];
        double cp_prev = 0.0, vprev = 1.0, alpha = 1.0, Vsum = 0.0, ulast = 0.0;
        for (int e = e0; e < M - 1; ++e) {
            const int i = end[p] - (M - 1 - e);
            const double ai = (!g.cyclic && i == 0 && !ghost_lo) ? 0.0 : a[i];
            const double den = b[i] - ai * cp_prev;
            const double cpe = c[i] / den;
            const double la = -ai / den;
            const double uc = -cpe;
            const double v = la * vprev;
            const int t = cbase[p] + e;
            blob[g.o_f + 2 * t] = 1.0 / den;
            blob[g.o_f + 2 * t + 1] = la;
            blob[g.o_alpha + t] = alpha;
            blob[g.o_b + 2 * t] = uc;
            blob[g.o_b + 2 * t + 1] = v;
            Vsum += alpha * v;
            alpha *= uc;
            cp_prev = cpe; vprev = v; ulast = uc;
        }
        const bool interior = len[p] > 1;
        V[p] = interior ? Vsum : 0.0;
        W[p] = interior ? alpha : 1.0;
        Vl[p] = vprev;   // 1 when the chunk has no interior: its left neighbour is S_{p-1} itself
        Wl[p] = ulast;
    }
    for (int p = 0; p < P; ++p) {
        const int i = end[p];
        const bool has_next = g.cyclic || p + 1 < P;
        const bool to_ghost = !has_next && ghost_hi;      // the cell after the last separator is the ghost R
        const int nx = (p + 1) % P;
        const double aS = (!g.cyclic && i == 0 && !ghost_lo) ? 0.0 : a[i];
        const double cS = ((has_next && (g.cyclic || i < n - 1)) || to_ghost) ? c[i] : 0.0;
        const double Vn = to_ghost ? 0.0 : V[nx], Wn = to_ghost ? 1.0 : W[nx];
        const double B = b[i] + aS * Wl[p] + cS * Vn;
        A[p] = aS * Vl[p] / B;
        C[p] = cS * Wn / B;
        blob[g.o_t0 + p] = 1.0 / B;
        blob[g.o_t1 + p] = -aS / B;
        blob[g.o_t2 + p] = to_ghost ? 0.0 : -cS / B;   // the ghost cell contributes no right-hand side of its own
    }
    // ghost couplings move to right-hand-side columns (the PCR below treats rows beyond the ends as zero)
    std::vector<double> DL(P, 0.0), DR(P, 0.0), DLn(P), DRn(P);
    if (!g.cyclic) {
        DL[0] = -A[0]; A[0] = 0.0;
        DR[P - 1] = -C[P - 1]; C[P - 1] = 0.0;
    }
    blob[g.o_misc] = W[0];
    blob[g.o_misc + 1] = V[0];
    for (int l = 0; l < g.levels; ++l) {
        const int s = 1 << l;
        double *R = blob + g.o_lvl + (size_t)l * 3 * P;
        for (int p = 0; p < P; ++p) {
            double r, ca, cc, an = 0.0, cn = 0.0;
            if (g.cyclic) {
                const int lo = (p - s + P) % P, hi = (p + s) % P;
                if (2 * s == P) {  // lo == hi: the two neighbours are the same unknown
                    const double gp = A[p] + C[p], gj = A[lo] + C[lo];
                    const double Bn = 1.0 - gp * gj;
                    r = 1.0 / Bn; ca = -gp / Bn; cc = 0.0;
                } else {
                    const double Bn = 1.0 - A[p] * C[lo] - C[p] * A[hi];
                    r = 1.0 / Bn; ca = -A[p] / Bn; cc = -C[p] / Bn;
                    an = -A[p] * A[lo] / Bn; cn = -C[p] * C[hi] / Bn;
                }
            } else {
                const int lo = p - s, hi = p + s;
                const double Clo = lo >= 0 ? C[lo] : 0.0, Alo = lo >= 0 ? A[lo] : 0.0;
                const double Ahi = hi < P ? A[hi] : 0.0, Chi = hi < P ? C[hi] : 0.0;
                const double Ap = lo >= 0 ? A[p] : 0.0, Cp = hi < P ? C[p] : 0.0;
                const double Bn = 1.0 - Ap * Clo - Cp * Ahi;
                r = 1.0 / Bn; ca = -Ap / Bn; cc = -Cp / Bn;
                an = -Ap * Alo / Bn; cn = -Cp * Chi / Bn;
            }
            R[p] = r; R[P + p] = ca; R[2 * P + p] = cc;
            An[p] = an; Cn[p] = cn;
            if (!g.cyclic) {
                const int lo = p - s >= 0 ? p - s : 0, hi = p + s < P ? p + s : P - 1;   // coefficients are 0 where clamped
                DLn[p] = r * DL[p] + ca * DL[lo] + cc * DL[hi];
                DRn[p] = r * DR[p] + ca * DR[lo] + cc * DR[hi];
            }
        }
        A.swap(An); C.swap(Cn);
        if (!g.cyclic) { DL.swap(DLn); DR.swap(DRn); }
    }
    for (int p = 0; p < P; ++p) {
        blob[g.o_dl + p] = g.cyclic ? 0.0 : DL[p];
        blob[g.o_dr + p] = g.cyclic ? 0.0 : DR[p];
    }
}

// Table reads: through the read-only (L1) path on the device -- every block of an SM walks the
// same few KB of tables while the field data streams past L1 (cp.async.cg / plain stores).
ADI_HD double tab_ld(const double *p)
{
#if defined(__CUDA_ARCH__)
    return __ldg(p);
#else
    return *p;
#endif
}
struct TabPair {
    double a, b;
};
ADI_HD TabPair tab_ld2(const double *p)   // p is 16-byte aligned
{
    TabPair r;
#if defined(__CUDA_ARCH__)
    const double2 t = __ldg(reinterpret_cast<const double2 *>(p));
    r.a = t.x; r.b = t.y;
#else
    r.a = p[0]; r.b = p[1];
#endif
    return r;
}

// One table set (host).  tab_make builds the tables of a line whose rows are arbitrary and then
// merges chunks whose cell tables are bit-identical (e.g. all interior chunks of a z line), so
// the set stays small enough to live in L1.
struct TabSet {
    TabGeom g;
    std::vector<double> blob;
    std::vector<int> geom;  // cbase[P], end[P], len[P]
};

inline TabSet tab_make(int n, int M, bool cyclic, const double *a, const double *b, const double *c,
                       bool ghost_lo = false, bool ghost_hi = false)
{
    TabSet full;
    full.g = tab_geom(n, M, cyclic, false);
    const int P = full.g.P;
    full.geom.resize(3 * P);
    full.blob.resize(full.g.ndbl);
    int *cb = full.geom.data(), *en = cb + P, *ln = cb + 2 * P;
    tab_partition(full.g, false, cb, en, ln);
    tab_build(full.g, cb, en, ln, a, b, c, full.blob.data(), ghost_lo, ghost_hi);
    // per chunk: 2M + 2M + M table doubles
    auto same_chunk = [&](int p, int q) {
        for (int e = 0; e < 2 * M; ++e)
            if (full.blob[full.g.o_f + 2 * p * M + e] != full.blob[full.g.o_f + 2 * q * M + e] ||
                full.blob[full.g.o_b + 2 * p * M + e] != full.blob[full.g.o_b + 2 * q * M + e]) return false;
        for (int e = 0; e < M; ++e)
            if (full.blob[full.g.o_alpha + p * M + e] != full.blob[full.g.o_alpha + q * M + e]) return false;
        return true;
    };
    std::vector<int> var(P), rep;
    for (int p = 0; p < P; ++p) {
        int found = -1;
        for (size_t v = 0; v < rep.size() && found < 0; ++v)
            if (same_chunk(p, rep[v])) found = (int)v;
        if (found < 0) { rep.push_back(p); found = (int)rep.size() - 1; }
        var[p] = found;
    }
    TabSet out;
    out.g = tab_geom_nvar(n, M, cyclic, (int)rep.size());
    out.geom = full.geom;
    out.blob.assign(out.g.ndbl, 0.0);
    for (size_t v = 0; v < rep.size(); ++v) {
        for (int e = 0; e < 2 * M; ++e) {
            out.blob[out.g.o_f + 2 * v * M + e] = full.blob[full.g.o_f + 2 * rep[v] * M + e];
            out.blob[out.g.o_b + 2 * v * M + e] = full.blob[full.g.o_b + 2 * rep[v] * M + e];
        }
        for (int e = 0; e < M; ++e) out.blob[out.g.o_alpha + v * M + e] = full.blob[full.g.o_alpha + rep[v] * M + e];
    }
    for (int p = 0; p < P; ++p) out.geom[p] = var[p] * M;
    for (int i = 0; i < 3 * P; ++i) out.blob[out.g.o_t0 + i] = full.blob[full.g.o_t0 + i];
    for (int i = 0; i < out.g.levels * 3 * P; ++i) out.blob[out.g.o_lvl + i] = full.blob[full.g.o_lvl + i];
    for (int i = 0; i < 2 * P + 2; ++i) out.blob[out.g.o_dl + i] = full.blob[full.g.o_dl + i];   // dl | dr | misc
    return out;
}

// Neighbour chunk indices of the reduced system (clamped where the coefficient is zero).
ADI_HD int tab_lo(int p, int s, int P, int cyclic) { return cyclic ? ((p - s) & (P - 1)) : (p - s >= 0 ? p - s : 0); }
ADI_HD int tab_hi(int p, int s, int P, int cyclic) { return cyclic ? ((p + s) & (P - 1)) : (p + s < P ? p + s : P - 1); }

// Phase 1.  f / alpha already point at the chunk's tables ({rinv, la} pairs; alpha).  d[e] holds the
// right-hand side (0 in padding slots) and is replaced by dp_e; returns Y, *Yl = dp of the last
// interior cell.
template <int M>
ADI_HD double tab_forward(double (&d)[M], const double *f, const double *alpha, double *Yl)
{
    double dp = 0.0, Y = 0.0;
#pragma unroll
    for (int e = 0; e < M - 1; ++e) {
        const TabPair t = tab_ld2(f + 2 * e);          // rinv, la
        dp = fma(t.b, dp, d[e] * t.a);
        d[e] = dp;
        Y = fma(tab_ld(alpha + e), dp, Y);
    }
    *Yl = dp;
    return Y;
}

// Reduced right-hand side of this chunk's separator row.
ADI_HD double tab_reduced_rhs(double t0, double t1, double t2, double ds, double Yl, double Ynext)
{
    return fma(t2, Ynext, fma(t1, Yl, t0 * ds));
}

// One PCR level on the right-hand side.
ADI_HD double tab_level(double r, double a, double c, double D, double Dlo, double Dhi)
{
    return fma(c, Dhi, fma(a, Dlo, r * D));
}

// Phase 3.  b points at the chunk's {u, v} pairs.  Leaves the solution in d (padding slots: 0).
template <int M>
ADI_HD void tab_backward(double (&d)[M], const double *b, double Sl, double S)
{
    double xn = S;
    d[M - 1] = S;
#pragma unroll
    for (int e = M - 2; e >= 0; --e) {
        const TabPair t = tab_ld2(b + 2 * e);          // u, v
        const double x = fma(t.a, xn, fma(t.b, Sl, d[e]));
        d[e] = x;
        xn = x;
    }
}

// ---- coefficient rows of the reference (host) --------------------------------------------

// build_coeff_r  adi3d_cyl_phi_v3.py:155-202 with theta = 1 (scheme "be", :341).
// Returns the amount added to the right-hand side of the outer cell (:200-201).
inline double cyl_rows_r(int nr, double dr, double alpha, double k, double dt, double h, double Tinf,
                         double *a, double *b, double *c)
{
    auto r_i = [&](int i) { double r = ((double)i + 0.5) * dr; return r > 1e-15 ? r : 1e-15; };
    auto r_imh = [&](int i) { double r = ((double)i + 0.5) * dr - 0.5 * dr; return r > 1e-15 ? r : 1e-15; };
    auto r_iph = [&](int i) { return ((double)i + 0.5) * dr + 0.5 * dr; };
    const double fac = 1.0 * alpha * dt;
    for (int i = 0; i < nr; ++i) a[i] = b[i] = c[i] = 0.0;
    for (int i = 1; i < nr - 1; ++i) {  // :175-180
        const double ai = -fac * (r_imh(i) / (r_i(i) * dr * dr));
        const double ci = -fac * (r_iph(i) / (r_i(i) * dr * dr));
        a[i] = ai; b[i] = 1.0 - (ai + ci); c[i] = ci;
    }
    {  // axis row :183-186
        const double c0 = -fac * (r_iph(0) / (r_i(0) * dr * dr));
        a[0] = 0.0; b[0] = 1.0 - c0; c[0] = c0;
    }
    const int N = nr - 1;  // outer Robin row :189-196
    const double aN = -fac * (r_imh(N) / (r_i(N) * dr * dr));
    double bN = 1.0 + fac * (r_imh(N) / (r_i(N) * dr * dr));
    double add = 0.0;
    if (h != 0.0) {
        bN += fac * (r_iph(N) * (h / k)) / (r_i(N) * dr);
        add = fac * (r_iph(N) * (h / k)) / (r_i(N) * dr) * Tinf;
    }
    a[N] = aN; b[N] = bN; c[N] = 0.0;
    return add;
}

// fac_i of phi_solve_spectral  :311-317 (theta = 1): rows are (-f, 1+2f, -f), periodic.
inline double cyl_fac_phi(int ir, double dr, double dphi, double alpha, double dt)
{
    if (ir == 0) return 0.0;
    const double r = ((double)ir + 0.5) * dr;
    return 1.0 * alpha * dt / (r * r * dphi * dphi);
}

struct ZEnd {
    int set;       // 1: Dirichlet (right-hand side := val), 0: right-hand side += val
    double val;
};

// build_coeff_z  :255-298 with theta = 1.  kind: 0 neumann0, 1 dirichlet, 2 robin.
inline int cyl_rows_z(int nz, double dz, double alpha, double k, double dt, int kind_bot, int kind_top,
                      double h_bot, double h_top, double Tinf_bot, double Tinf_top, double T_bot,
                      double T_top, double *a, double *b, double *c, ZEnd *bot, ZEnd *top)
{
    const double fac = 1.0 * alpha * dt / (dz * dz);
    for (int i = 0; i < nz; ++i) { a[i] = b[i] = c[i] = 0.0; }
    for (int i = 1; i < nz - 1; ++i) { a[i] = -fac; b[i] = 1.0 + 2.0 * fac; c[i] = -fac; }
    bot->set = 0; bot->val = 0.0;
    top->set = 0; top->val = 0.0;
    if (kind_bot == 0) { a[0] = 0.0; b[0] = 1.0 + fac; c[0] = -fac; }
    else if (kind_bot == 1) { a[0] = 0.0; b[0] = 1.0; c[0] = 0.0; bot->set = 1; bot->val = T_bot; }
    else if (kind_bot == 2) {
        const double beta = h_bot / k;
        a[0] = 0.0; b[0] = 1.0 + fac * (1.0 + beta * dz); c[0] = -fac;
        bot->val = (1.0 * alpha * dt) * (beta / dz) * Tinf_bot;
    } else return -1;
    const int N = nz - 1;
    if (kind_top == 0) { a[N] = -fac; b[N] = 1.0 + fac; c[N] = 0.0; }
    else if (kind_top == 1) { a[N] = 0.0; b[N] = 1.0; c[N] = 0.0; top->set = 1; top->val = T_top; }
    else if (kind_top == 2) {
        const double beta = h_top / k;
        a[N] = -fac; b[N] = 1.0 + fac * (1.0 + beta * dz); c[N] = 0.0;
        top->val = (1.0 * alpha * dt) * (beta / dz) * Tinf_top;
    } else return -1;
    return 0;
}

// Rows of the z sweep for the segment [z0, z0 + nz_loc) of a line of nz_glob cells: the boundary rows of
// build_coeff_z only on the segments that hold the global first / last cell, interior rows elsewhere
// (their end couplings then reach the adjacent segment's cell -- tab_build's ghosts).
inline int cyl_rows_z_segment(int nz_loc, bool first, bool last, double dz, double alpha, double k, double dt,
                              int kind_bot, int kind_top, double h_bot, double h_top, double Tinf_bot, double Tinf_top,
                              double T_bot, double T_top, double *a, double *b, double *c, ZEnd *bot, ZEnd *top)
{
    const int rc = cyl_rows_z(nz_loc, dz, alpha, k, dt, kind_bot, kind_top, h_bot, h_top, Tinf_bot, Tinf_top, T_bot,
                              T_top, a, b, c, bot, top);
    if (rc) return rc;
    const double fac = 1.0 * alpha * dt / (dz * dz);
    // order matters for a one-cell segment: the reference writes the bottom row first, then the top row
    if (!first) { a[0] = -fac; b[0] = 1.0 + 2.0 * fac; c[0] = -fac; bot->set = 0; bot->val = 0.0; }
    if (!last) { a[nz_loc - 1] = -fac; b[nz_loc - 1] = 1.0 + 2.0 * fac; c[nz_loc - 1] = -fac; top->set = 0; top->val = 0.0; }
    if (!first && last && nz_loc == 1) { /* top row written by cyl_rows_z stays, with its a = -fac */ }
    return 0;
}

}  // namespace adi
