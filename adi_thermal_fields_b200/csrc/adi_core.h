// adi_core.h -- per-thread building blocks of the partitioned tridiagonal line solve.
//
// One ADI sweep solves, for every grid line along the swept axis, the tridiagonal system
// the reference assembles in sweep_axis0/1/2 (adi3d_numba_coeff.py:133-237; un-compressed
// form: adi3d_gpu_coeff.py:154-191).  A line of n cells is cut into P = ceil(n/M) chunks of
// M cells; one THREAD owns one chunk and keeps it in registers:
//
//   phase 1  forward elimination over the chunk's M-1 interior cells with the coupling to
//            the previous chunk kept symbolic (spike v'), then a backward recurrence that
//            yields the first interior cell as  x_0 = Y0 - V0*S_{p-1} - W0*S_p ;
//   phase 2  the chunk's LAST cell is its separator S_p; substituting the neighbours'
//            relations into its row gives one row of a P-unknown tridiagonal system,
//            which the P threads of the line solve together by parallel cyclic reduction
//            (ceil(log2 P) steps, exchanged through shared memory);
//   phase 3  back substitution inside the chunk with S_{p-1}, S_p known.
//
// Each cell is therefore read once and written once; nothing is spilled to HBM.
// Void cells (mask false) are identity rows that are never coupled to anything and come
// out bit-identical to the input (the reference never touches them); no value of a void
// cell ever enters arithmetic that reaches an active cell (they may hold NaN).
//
// The functions are __host__ __device__ so that tests/ can run the very same code on the
// CPU (csrc/host_emulation.cpp) against the oracle.
#pragma once
#include <stdint.h>
#if !defined(__CUDACC__)
#include <cmath>
using std::fma;
#endif

#if defined(__CUDACC__)
#define ADI_HD __host__ __device__ __forceinline__
#else
#define ADI_HD inline
#endif

namespace adi {

// Per-cell neighbour code, built once per mask change (kernel build_code):
// bit0 = cell active; bits 1..6 = neighbour across x-,x+,y-,y+,z-,z+ exists and is active;
// bit7 = Dirichlet cell of this axis' pack (dir_mask & mask).
enum : unsigned {
    CB_SELF = 1u, CB_XM = 2u, CB_XP = 4u, CB_YM = 8u, CB_YP = 16u, CB_ZM = 32u, CB_ZP = 64u,
    CB_DIR = 128u
};

ADI_HD double frcp(double x)
{
#if defined(__CUDA_ARCH__) && !defined(ADI_EXACT_DIV)
    // MUFU.RCP64H seed (>= 20 good bits) + two Newton steps in residual form: ~1 ulp,
    // branch-free.  Pivots here are >= 1 (diagonally dominant M-matrix rows), never denormal.
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    return r;
#else
    return 1.0 / x;
#endif
}

ADI_HD double sel(bool c, double a, double b) { return c ? a : b; }

// Scalars of one sweep (adi3d_numba_coeff.py:291-292,299-301).
struct SweepConst {
    double g;       // theta*gam, gam = kappa*dt/dx^2
    double dt;      // params.dt
    double Tinf;    // ambient of the Robin term
    double h_lo;    // CMODE 1: h*A/Ccell of the '-' face of this axis (on-the-fly Robin)
    double h_hi;    // CMODE 1: same for the '+' face
    double beta;    // dt*kappa*(1-theta), explicit stage (:298)
    double invdx2;  // 1/dx^2 (:243)
};

// Row of the tridiagonal system at one cell (adi3d_numba_coeff.py:146-163).
// Off-diagonals are -g or 0, so they are carried as the non-negative numbers aa=-a, cc=-c.
struct Row {
    double aa, cc, b, d;
    bool active;
};

// lo/hi: bit masks of the '-' / '+' neighbour along the swept axis.
// CMODE 0: no Robin term; 1: scalar per face, derived from the code; 2: dense coeff value cval.
template <int CMODE, bool EXTRA>
ADI_HD Row make_row(unsigned code, unsigned lo, unsigned hi, double Tval, double cval, double qval,
                    double dirval, const SweepConst &k)
{
    Row r;
    r.active = (code & CB_SELF) != 0;
    const bool L = r.active && (code & lo);
    const bool R = r.active && (code & hi);
    double c = 0.0;
    if (CMODE == 1) {
        // exposed on a face <=> active and no active neighbour across it (:38-55);
        // coeff = (0 + h_lo*A/Ccell) + h_hi*A/Ccell in the reference's face order (:93-99)
        c = sel(r.active && !(code & lo), k.h_lo, 0.0);
        c = c + sel(r.active && !(code & hi), k.h_hi, 0.0);
    } else if (CMODE == 2) {
        c = sel(r.active, cval, 0.0);
    }
    const double q = EXTRA ? sel(r.active, qval, 0.0) : 0.0;
    r.aa = sel(L, k.g, 0.0);
    r.cc = sel(R, k.g, 0.0);
    const double nn = r.aa + r.cc;                 // theta*gam*nnb
    const double dtc = k.dt * c;
    r.b = (1.0 + nn) + dtc;                        // :155
    r.d = fma(k.dt, q, Tval) + dtc * k.Tinf;       // :162
    if (EXTRA) {
        if (r.active && (code & CB_DIR)) {         // :157-158
            r.aa = 0.0; r.cc = 0.0; r.b = 1.0; r.d = dirval;
        }
    }
    return r;
}

// Relation handed to the previous chunk: x_first = Y - V*S_{p-1} - W*S_p  (all finite).
struct First {
    double Y, V, W;
};

// Normalised reduced row: A*S_{p-s} + S_p + C*S_{p+s} = D.
struct Red {
    double A, C, D;
};

// Chunk state held in registers across the phases.  T[e] holds, in turn, the input value,
// the eliminated right-hand side d'_e and finally the solution; Cc[e] holds coeff then c'_e;
// Vp[e] holds the spike v'_e.  code[e] are the neighbour codes.
template <int M>
struct Chunk {
    double T[M];
    double Cc[M];
    double Vp[M];
    unsigned code[M];
    // separator row pieces kept from phase 1 to phase 2
    double s_aa, s_cc, s_b, s_d;
    double Yl, Vl, Wl;
    bool s_active;
};

// Phase 1.  Q[e]/DV[e] (Neumann flux, Dirichlet value) are only read when EXTRA.
template <int M, int CMODE, bool EXTRA>
ADI_HD First chunk_forward(Chunk<M> &ch, const double *Q, const double *DV, unsigned lo, unsigned hi,
                           const SweepConst &k)
{
    double cprev = 0.0, dprev = 0.0, vprev = 0.0;
#pragma unroll
    for (int e = 0; e < M - 1; ++e) {
        const Row r = make_row<CMODE, EXTRA>(ch.code[e], lo, hi, ch.T[e], ch.Cc[e],
                                             EXTRA ? Q[e] : 0.0, EXTRA ? DV[e] : 0.0, k);
        double den, vp;
        if (e == 0) {
            den = r.b;                       // the coupling aa to S_{p-1} stays symbolic
        } else {
            den = fma(r.aa, cprev, r.b);     // b - a*c'_{e-1}
        }
        const double rinv = frcp(den);
        const double cp = -r.cc * rinv;      // c'_e = c/den  (<= 0)
        const double dp = (e == 0 ? r.d : fma(r.aa, dprev, r.d)) * rinv;
        if (e == 0) vp = -r.aa * rinv;       // v'_0 = a_0/den
        else vp = r.aa * vprev * rinv;       // v'_e = -a_e v'_{e-1}/den
        ch.Cc[e] = cp;
        ch.Vp[e] = vp;
        if (r.active) ch.T[e] = dp;          // a void cell keeps its input bits
        cprev = cp;
        vprev = vp;
        dprev = sel(r.active, dp, 0.0);      // never let a void value travel
    }
    {
        const int e = M - 1;
        const Row r = make_row<CMODE, EXTRA>(ch.code[e], lo, hi, ch.T[e], ch.Cc[e],
                                             EXTRA ? Q[e] : 0.0, EXTRA ? DV[e] : 0.0, k);
        ch.s_aa = r.aa; ch.s_cc = r.cc; ch.s_b = r.b; ch.s_d = sel(r.active, r.d, 0.0);
        ch.s_active = r.active;
    }
    // backward recurrence for (Y,V,W) of the first interior cell
    const bool actl = (ch.code[M - 2] & CB_SELF) != 0;
    double Y = sel(actl, ch.T[M - 2], 0.0), V = ch.Vp[M - 2], W = ch.Cc[M - 2];
    ch.Yl = Y; ch.Vl = V; ch.Wl = W;
#pragma unroll
    for (int e = M - 3; e >= 0; --e) {
        const bool act = (ch.code[e] & CB_SELF) != 0;
        const double cp = ch.Cc[e];
        Y = sel(act, fma(-cp, Y, ch.T[e]), 0.0);
        V = fma(-cp, V, ch.Vp[e]);
        W = -cp * W;
    }
    First f;
    f.Y = Y; f.V = V; f.W = W;
    return f;
}

// Phase 2a: the separator row of this chunk given the next chunk's First relation
// (zeros when there is no next chunk).  Returns the normalised reduced row.
template <int M>
ADI_HD Red chunk_reduced_row(const Chunk<M> &ch, const First &nx)
{
    // a_s x_{M-2} + b_s S_p + c_s x_0^{(p+1)} = d_s,   a_s=-s_aa, c_s=-s_cc,
    // x_{M-2} = Yl - Vl S_{p-1} - Wl S_p,   x_0^{(p+1)} = nx.Y - nx.V S_p - nx.W S_{p+1}
    const double A = ch.s_aa * ch.Vl;
    const double B = fma(ch.s_aa, ch.Wl, ch.s_b) + ch.s_cc * nx.V;
    const double C = ch.s_cc * nx.W;
    const double D = fma(ch.s_aa, ch.Yl, ch.s_d) + ch.s_cc * nx.Y;
    const double rB = frcp(B);
    Red r;
    r.A = A * rB; r.C = C * rB; r.D = D * rB;
    return r;
}

// Phase 2b: one parallel-cyclic-reduction step; lo/hi are the rows at distance s
// (all-zero rows beyond the ends).
ADI_HD Red pcr_step(const Red &me, const Red &lo, const Red &hi)
{
    const double B = fma(-me.C, hi.A, fma(-me.A, lo.C, 1.0));
    const double D = fma(-me.C, hi.D, fma(-me.A, lo.D, me.D));
    const double A = -me.A * lo.A;
    const double C = -me.C * hi.C;
    const double rB = frcp(B);
    Red r;
    r.A = A * rB; r.C = C * rB; r.D = D * rB;
    return r;
}

// Phase 3: Sl = S_{p-1} (0 for the first chunk), S = S_p.  Leaves the solution in ch.T.
template <int M>
ADI_HD void chunk_backward(Chunk<M> &ch, double Sl, double S)
{
    double xn = sel(ch.s_active, S, 0.0);
    if (ch.s_active) ch.T[M - 1] = S;
#pragma unroll
    for (int e = M - 2; e >= 0; --e) {
        const bool act = (ch.code[e] & CB_SELF) != 0;
        double x = fma(-ch.Vp[e], Sl, ch.T[e]);
        x = fma(-ch.Cc[e], xn, x);
        if (act) ch.T[e] = x;
        xn = sel(act, x, 0.0);
    }
}

// Explicit stage (adi3d_numba_coeff.py:240-288,298) for one cell:
// R0 = T + beta*((Lx+Ly)+Lz), L* = ((sum of active neighbours, '-' first) - cnt*T)/dx^2.
// The caller supplies neighbour values (ignored where the code bit is clear).
ADI_HD double explicit_r0(unsigned code, double T, double xm, double xp, double ym, double yp,
                          double zm, double zp, const SweepConst &k)
{
    if (!(code & CB_SELF)) return T;
    double s, c, L[3];
    s = 0.0; c = 0.0;
    if (code & CB_XM) { s += xm; c += 1.0; }
    if (code & CB_XP) { s += xp; c += 1.0; }
    L[0] = (s - c * T) * k.invdx2;
    s = 0.0; c = 0.0;
    if (code & CB_YM) { s += ym; c += 1.0; }
    if (code & CB_YP) { s += yp; c += 1.0; }
    L[1] = (s - c * T) * k.invdx2;
    s = 0.0; c = 0.0;
    if (code & CB_ZM) { s += zm; c += 1.0; }
    if (code & CB_ZP) { s += zp; c += 1.0; }
    L[2] = (s - c * T) * k.invdx2;
    return T + k.beta * ((L[0] + L[1]) + L[2]);
}

}  // namespace adi
