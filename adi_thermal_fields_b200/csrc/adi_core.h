// adi_core.h -- per-thread building blocks of the partitioned tridiagonal line solve.
//
// One ADI sweep solves, for every grid line along the swept axis, the tridiagonal system
// the reference assembles in sweep_axis0/1/2 (adi3d_numba_coeff.py:133-237; un-compressed
// form: adi3d_gpu_coeff.py:154-191).  A line of n cells is cut into P = ceil(n/M) chunks of
// M cells; one THREAD owns one chunk.  The last cell of a chunk is its separator S_p, the
// M-1 cells before it are its interior.
//
//   phase 1  one forward pass over the interior: the row (a,b,c,d) of each cell is built
//            from the neighbour code and the operands, eliminated against its left
//            neighbour (Thomas, normalised form), and three running sums give the first
//            interior cell as an affine function of the two separators around the chunk,
//                x_0 = Y + V*S_{p-1} + W*S_p ,
//            likewise the last one, x_{M-2} = Yl + Vl*S_{p-1} + Wl*S_p.
//            Kept per cell: the scaled right-hand side d_e/den_e (registers, in place of
//            T) and the two normalised couplings aa_e/den_e, cc_e/den_e (shared memory);
//            for long lines only 1/den_e is kept and the couplings are rebuilt from the code.
//   phase 2  substituting the neighbours' relations into the separator rows gives a
//            P-unknown tridiagonal system per line, solved by the P threads of the line
//            with parallel cyclic reduction (ceil(log2 P) steps);
//   phase 3  with S_{p-1}, S_p known: forward elimination of d once more (now with the
//            true left boundary value) and back substitution.
//
// Each cell is read once and written once; nothing is spilled to HBM.
// Void cells (mask false) are identity rows that are never coupled to anything; their
// values are never loaded into the arithmetic (they may hold NaN) and never stored, so they
// come out bit-identical to the input (the reference never touches them).
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

// Per-cell neighbour code, built once per mask change (kernel k_build_code):
// bit0 = cell active; bits 1..6 = neighbour across x-,x+,y-,y+,z-,z+ exists and is active;
// bit7 = Dirichlet cell of this axis' pack (dir_mask & mask).  A void cell has code 0, so a
// set neighbour or Dirichlet bit implies an active cell.
enum : unsigned {
    CB_SELF = 1u, CB_XM = 2u, CB_XP = 4u, CB_YM = 8u, CB_YP = 16u, CB_ZM = 32u, CB_ZP = 64u,
    CB_DIR = 128u
};

ADI_HD double frcp(double x)
{
#if defined(__CUDA_ARCH__) && !defined(ADI_EXACT_DIV)
    // MUFU.RCP64H seed (>= 20 good bits) + one cubically convergent step
    // r' = r + r*(e + e^2), e = 1 - x*r  (|1 - x*r'| = |e|^3 <= 2^-60): ~1 ulp, branch-free.
    // Pivots here are >= 1 (diagonally dominant M-matrix rows), never denormal or huge.
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    const double e = fma(-x, r, 1.0);
    const double t = fma(e, e, e);
    return fma(r, t, r);
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

// Couplings of a row to its '-' / '+' neighbour along the swept axis: present when the cell
// and that neighbour are active and the cell is not a Dirichlet cell (:151-158; the
// neighbours of a Dirichlet cell still couple TO it).
template <bool EXTRA>
ADI_HD bool couples(unsigned code, unsigned nb)
{
    return EXTRA ? ((code & (nb | CB_DIR)) == nb) : ((code & nb) != 0);
}

// lo/hi: bit masks of the '-' / '+' neighbour along the swept axis.
// CMODE 0: no Robin term; 1: scalar per face, derived from the code; 2: dense coeff value cval.
template <int CMODE, bool EXTRA>
ADI_HD Row make_row(unsigned code, unsigned lo, unsigned hi, double Tval, double cval, double qval,
                    double dirval, const SweepConst &k)
{
    Row r;
    r.active = (code & CB_SELF) != 0;
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
    r.aa = sel((code & lo) != 0, k.g, 0.0);
    r.cc = sel((code & hi) != 0, k.g, 0.0);
    const double nn = r.aa + r.cc;                 // theta*gam*nnb
    const double dtc = k.dt * c;
    r.b = (1.0 + nn) + dtc;                        // :155
    r.d = fma(k.dt, q, Tval) + dtc * k.Tinf;       // :162
    if (EXTRA) {
        if (code & CB_DIR) {                       // :157-158
            r.aa = 0.0; r.cc = 0.0; r.b = 1.0; r.d = dirval;
        }
    }
    return r;
}

// Affine relation of a chunk's first interior cell: x_0 = Y + V*S_{p-1} + W*S_p.
struct First {
    double Y, V, W;
};

// Normalised reduced row: A*S_{p-s} + S_p + C*S_{p+s} = D.
struct Red {
    double A, C, D;
};

// Chunk state held in registers across the phases.  T[e] holds, in turn, the input value
// (0 for a void cell -- see load rule below), the row's scaled right-hand side and finally the
// solution.  cw packs the neighbour codes, four cells per word.
//
// Load rule: the caller puts 0.0 into T[e] of void cells (and of padding cells beyond the
// line end, which carry code 0).  Void rows are then exact identity rows with zero
// right-hand side and zero couplings, so no select is needed anywhere downstream; the
// caller never stores T[e] of a void cell (the output keeps / is given the input bits).
template <int M>
struct Chunk {
    double T[M];
    unsigned cw[(M + 3) / 4];
    // separator row and last-interior relation, kept from phase 1 to phase 2
    double s_aa, s_cc, s_b, s_d;
    double Yl, Vl, Wl;

    ADI_HD unsigned code(int e) const { return (cw[e >> 2] >> (8 * (e & 3))) & 0xffu; }
    ADI_HD bool active(int e) const { return (cw[e >> 2] >> (8 * (e & 3))) & 1u; }
    ADI_HD void set_code(int e, unsigned c)
    {
        if ((e & 3) == 0) cw[e >> 2] = c;
        else cw[e >> 2] |= c << (8 * (e & 3));
    }
};

// Operand access of one chunk.  OPS provides
//   double coef(int e)            dense Robin coefficient of cell e (CMODE 2)
//   double q(int e), dirv(int e)  Neumann flux / Dirichlet value (EXTRA)
//   per-cell factor store between phase 1 and phase 3:
//     NS == 2:  put2(e, la, u), la(e), u(e)     la = aa/den, u = cc/den
//     NS == 1:  put1(e, rinv), rinv(e)          rinv = 1/den (long lines: half the storage)
// with e a compile-time constant after unrolling.

// A chunk is "solid" when all its M cells are active, none is a Dirichlet cell and every link
// between two of its cells exists; only the outward links of its first and last cell may be
// missing.  Solid chunks (all of them in the bulk of a part) skip the per-cell decoding of the
// neighbour code: the same row arithmetic with the code bits known at compile time.
template <int M>
ADI_HD bool chunk_solid(const Chunk<M> &ch, unsigned lo, unsigned hi)
{
    static_assert(M % 4 == 0, "codes are packed four per word");
    const unsigned need = (CB_SELF | lo | hi) * 0x01010101u, care = (CB_SELF | lo | hi | CB_DIR) * 0x01010101u;
    bool ok = true;
#pragma unroll
    for (int w = 0; w < M / 4; ++w) {
        unsigned c = ch.cw[w];
        if (w == 0) c |= lo;                               // cell 0 may lack its '-' link
        if (w == M / 4 - 1) c |= hi << 24;                 // cell M-1 may lack its '+' link
        ok = ok && ((c & care) == need);
    }
    return ok;
}

// Code of cell e of a solid chunk: a compile-time constant except for the two end cells.
template <int M>
ADI_HD unsigned solid_code(const Chunk<M> &ch, int e, unsigned lo, unsigned hi)
{
    if (e == 0) return CB_SELF | hi | (ch.code(0) & lo);
    if (e == M - 1) return CB_SELF | lo | (ch.code(M - 1) & hi);
    return CB_SELF | lo | hi;
}

// Phase 1.  SOLID: the chunk passed chunk_solid (same arithmetic, selects folded away).
// ENDS_ONLY (with SOLID, CMODE 2): the coefficient field is known to vanish on cells that have both
// neighbours, i.e. everywhere in a solid chunk except possibly its two end cells.
template <int M, int CMODE, bool EXTRA, int NS, bool SOLID = false, bool ENDS_ONLY = false, class OPS>
ADI_HD First chunk_forward(Chunk<M> &ch, OPS &ops, unsigned lo, unsigned hi, const SweepConst &k)
{
    double uprev = 0.0, dprev = 0.0, vprev = 1.0, alpha = 1.0;
    First f;
    f.Y = 0.0; f.V = 0.0; f.W = 0.0;
#pragma unroll
    for (int e = 0; e < M - 1; ++e) {
        const Row r = make_row<CMODE, EXTRA>(SOLID ? solid_code<M>(ch, e, lo, hi) : ch.code(e), lo, hi, ch.T[e],
                                             CMODE == 2 ? ((ENDS_ONLY && e != 0) ? 0.0 : ops.coef(e)) : 0.0,
                                             EXTRA ? ops.q(e) : 0.0, EXTRA ? ops.dirv(e) : 0.0, k);
        // e == 0: the coupling aa to S_{p-1} stays symbolic (uprev = dprev = 0, vprev = 1)
        const double den = fma(-r.aa, uprev, r.b);      // b - a*c'_{e-1}
        const double rinv = frcp(den);
        const double u = r.cc * rinv;                   // -c'_e  (>= 0)
        const double la = r.aa * rinv;
        const double ds = r.d * rinv;
        const double dp = fma(la, dprev, ds);           // d'_e with S_{p-1} = 0
        const double vp = la * vprev;                   // coefficient of S_{p-1} in x_e
        if (NS == 2) { ops.put2(e, la, u); ch.T[e] = ds; }
        else { ops.put1(e, rinv); ch.T[e] = r.d; }
        f.Y = fma(alpha, dp, f.Y);
        f.V = fma(alpha, vp, f.V);
        alpha = alpha * u;
        uprev = u; dprev = dp; vprev = vp;
    }
    f.W = alpha;
    ch.Yl = dprev; ch.Vl = vprev; ch.Wl = uprev;        // x_{M-2} = Yl + Vl*S_{p-1} + Wl*S_p
    {
        const int e = M - 1;
        const Row r = make_row<CMODE, EXTRA>(SOLID ? solid_code<M>(ch, e, lo, hi) : ch.code(e), lo, hi, ch.T[e],
                                             CMODE == 2 ? ops.coef(e) : 0.0,
                                             EXTRA ? ops.q(e) : 0.0, EXTRA ? ops.dirv(e) : 0.0, k);
        ch.s_aa = r.aa; ch.s_cc = r.cc; ch.s_b = r.b; ch.s_d = r.d;
    }
    return f;
}

// ---- uniform chunks ---------------------------------------------------------------------
// A cell is "uniform" when it is active, has BOTH neighbours along the swept axis (so it is exposed on
// neither face of this axis: the Robin coefficient is zero, :93-99), is no Dirichlet cell and carries no
// flux term: its row is (-g, 1+2g, -g | T).  This is the bulk of every part.  The elimination factors of a
// run of uniform cells do not depend on the data or on where the run sits -- only on g -- so they are
// tabulated once per launch on the host (UniConst, passed in the kernel parameters = constant bank
// operands) and phase 1 / phase 3 shrink to three / two fused multiply-adds per cell with no reciprocal,
// no code decoding and no factor store.  Two chunk shapes take this path:
//   OFF 0   interior cells 0..M-2 uniform;
//   OFF 1   cell 0 any active, non-Dirichlet cell coupled to cell 1 (typically the exposed first cell of a
//           line), cells 1..M-2 uniform: cell 0 is eliminated by hand against the tabulated run.
// The separator (cell M-1) is a general row in both (typically the exposed last cell of a line).
constexpr int UNI_MAX = 32;
struct UniConst {
    double rinv[UNI_MAX];  // 1/den_e
    double u[UNI_MAX];     // cc/den_e = aa/den_e  (la == u in a uniform run)
    double vp[UNI_MAX];    // coefficient of the value left of the run in x_e after the forward pass
    double al[UNI_MAX];    // u_0 ... u_{e-1}: weight of d'_e in the relation of the run's first cell
    double Vn[UNI_MAX];    // sum_{e<n} al_e*vp_e: coefficient of the left value in the first cell of a run of n cells
};

// Host side (same recurrences as chunk_forward on uniform rows).
inline void uni_const_build(UniConst &uc, double g)
{
    const double b = (1.0 + (g + g)) + 0.0;      // make_row: (1 + nn) + dt*c with c = 0
    double uprev = 0.0, vprev = 1.0, alpha = 1.0, V = 0.0;
    for (int e = 0; e < UNI_MAX; ++e) {
        uc.Vn[e] = V;
        const double den = fma(-g, uprev, b);
        const double rinv = 1.0 / den;
        const double u = g * rinv;
        const double vp = u * vprev;
        uc.rinv[e] = rinv; uc.u[e] = u; uc.vp[e] = vp; uc.al[e] = alpha;
        V = fma(alpha, vp, V);
        alpha = alpha * u;
        uprev = u; vprev = vp;
    }
}

// Cells OFF..M-2 uniform (and, OFF 1, cell 0 active, not Dirichlet, coupled to cell 1).
template <int M, int OFF>
ADI_HD bool chunk_uniform(const Chunk<M> &ch, unsigned lo, unsigned hi)
{
    static_assert(M % 4 == 0 && M >= 8, "codes are packed four per word");
    const unsigned need1 = CB_SELF | lo | hi, care1 = CB_SELF | lo | hi | CB_DIR;
    bool ok = true;
#pragma unroll
    for (int w = 0; w < M / 4; ++w) {
        unsigned need = 0u, care = 0u;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int cell = 4 * w + b;
            if (cell >= OFF && cell <= M - 2) { need |= need1 << (8 * b); care |= care1 << (8 * b); }
        }
        ok = ok && ((ch.cw[w] & care) == need);
    }
    if (OFF == 1) ok = ok && ((ch.cw[0] & (CB_SELF | hi | CB_DIR)) == (CB_SELF | hi));
    return ok;
}

// Relation of the hand-eliminated first cell (OFF 1): x_0 = al + bl*S_{p-1} + br*S_p.
struct UniHead {
    double al, bl, br;
};

// Phase 1.  sep: row of the separator (cell M-1); head: row of cell 0 (OFF 1 only).  Leaves d'_e
// (forward-eliminated right-hand side with the value left of the run = 0) in ch.T[e] of the run.
template <int M, int OFF>
ADI_HD First chunk_forward_uniform(Chunk<M> &ch, const UniConst &uc, const Row &sep, const Row &head, UniHead &hd)
{
    static_assert(M - 1 <= UNI_MAX, "UniConst is too short for this chunk length");
    constexpr int N = M - 1 - OFF;               // cells in the run
    double dprev = 0.0, Y = 0.0;
#pragma unroll
    for (int j = 0; j < N; ++j) {
        const double dp = fma(uc.u[j], dprev, ch.T[OFF + j] * uc.rinv[j]);
        ch.T[OFF + j] = dp;
        Y = fma(uc.al[j], dp, Y);
        dprev = dp;
    }
    // run: x_first = Y + V*left + W*S_p,  x_last = dprev + Vl*left + Wl*S_p
    const double V = uc.Vn[N], W = uc.al[N], Vl = uc.vp[N - 1], Wl = uc.u[N - 1];
    First f;
    if (OFF == 0) {
        f.Y = Y; f.V = V; f.W = W;
        ch.Yl = dprev; ch.Vl = Vl; ch.Wl = Wl;
        hd.al = hd.bl = hd.br = 0.0;
    } else {
        // -aa0*S_{p-1} + b0*x0 - cc0*x_first = d0
        const double r = frcp(fma(-head.cc, V, head.b));
        hd.al = fma(head.cc, Y, head.d) * r;
        hd.bl = head.aa * r;
        hd.br = (head.cc * W) * r;
        f.Y = hd.al; f.V = hd.bl; f.W = hd.br;
        ch.Yl = fma(Vl, hd.al, dprev); ch.Vl = Vl * hd.bl; ch.Wl = fma(Vl, hd.br, Wl);
    }
    ch.s_aa = sep.aa; ch.s_cc = sep.cc; ch.s_b = sep.b; ch.s_d = sep.d;
    return f;
}

// Phase 3.
template <int M, int OFF>
ADI_HD void chunk_backward_uniform(Chunk<M> &ch, const UniConst &uc, const UniHead &hd, double Sl, double S)
{
    constexpr int N = M - 1 - OFF;
    double left = Sl;
    if (OFF == 1) {
        left = fma(hd.br, S, fma(hd.bl, Sl, hd.al));
        ch.T[0] = left;
    }
    double xn = S;
    ch.T[M - 1] = S;
#pragma unroll
    for (int j = N - 1; j >= 0; --j) {
        const double x = fma(uc.u[j], xn, fma(uc.vp[j], left, ch.T[OFF + j]));
        ch.T[OFF + j] = x;
        xn = x;
    }
}

// ---- hybrid chunks ----------------------------------------------------------------------
// A chunk whose first R4 cells (R4 a multiple of four, 4 <= R4 <= M-4, the same for every lane of the warp) are
// uniform and whose remaining cells are anything: the chunk that holds the top surface of a part in a z sweep
// (uniform bulk below the surface, the exposed cell, void above).  The run takes the tabulated factors, the
// tail the general row assembly; the hand-over carries the run's last elimination state (all tabulated).
// Storage convention: T[e] = d'_e for e < R4 (as in chunk_forward_uniform), T[e] = d_e and ops.put1(e, 1/den_e)
// for e >= R4 (as in chunk_forward with NS = 1).  The branch per group of four cells is warp-uniform.

// Number of leading uniform cells of a chunk, rounded down to a multiple of four (0 .. M-4).
template <int M>
ADI_HD int chunk_uniform_lead4(const Chunk<M> &ch, unsigned lo, unsigned hi)
{
    const unsigned need = (CB_SELF | lo | hi) * 0x01010101u, care = (CB_SELF | lo | hi | CB_DIR) * 0x01010101u;
    int r = 0;
    bool run = true;
#pragma unroll
    for (int w = 0; w < M / 4 - 1; ++w) {
        run = run && ((ch.cw[w] & care) == need);
        r += run ? 4 : 0;
    }
    return r;
}

template <int M, int CMODE, bool EXTRA, class OPS>
ADI_HD First chunk_forward_hybrid(Chunk<M> &ch, const UniConst &uc, OPS &ops, int R4, unsigned lo, unsigned hi,
                                  const SweepConst &k)
{
    static_assert(M % 4 == 0 && M <= UNI_MAX, "groups of four cells; tables hold UNI_MAX entries");
    double uprev = 0.0, dprev = 0.0, vprev = 1.0, alpha = 1.0;
    First f;
    f.Y = 0.0; f.V = 0.0; f.W = 0.0;
#pragma unroll
    for (int g4 = 0; g4 < M / 4; ++g4) {
        if (4 * g4 < R4) {               // warp-uniform; R4 <= M-4, so never the group of the separator
#pragma unroll
            for (int e = 4 * g4; e < 4 * g4 + 4; ++e) {
                const double dp = fma(uc.u[e], dprev, ch.T[e] * uc.rinv[e]);
                ch.T[e] = dp;
                f.Y = fma(uc.al[e], dp, f.Y);
                dprev = dp;
            }
            uprev = uc.u[4 * g4 + 3]; vprev = uc.vp[4 * g4 + 3];
            alpha = uc.al[(4 * g4 + 4) & (UNI_MAX - 1)]; f.V = uc.Vn[(4 * g4 + 4) & (UNI_MAX - 1)];
        } else {
#pragma unroll
            for (int e = 4 * g4; e < 4 * g4 + 4; ++e) {
                if (e == M - 1) break;
                const Row r = make_row<CMODE, EXTRA>(ch.code(e), lo, hi, ch.T[e], CMODE == 2 ? ops.coef(e) : 0.0,
                                                     EXTRA ? ops.q(e) : 0.0, EXTRA ? ops.dirv(e) : 0.0, k);
                const double den = fma(-r.aa, uprev, r.b);
                const double rinv = frcp(den);
                const double u = r.cc * rinv;
                const double la = r.aa * rinv;
                const double dp = fma(la, dprev, r.d * rinv);
                const double vp = la * vprev;
                ops.put1(e, rinv); ch.T[e] = r.d;
                f.Y = fma(alpha, dp, f.Y);
                f.V = fma(alpha, vp, f.V);
                alpha = alpha * u;
                uprev = u; dprev = dp; vprev = vp;
            }
        }
    }
    f.W = alpha;
    ch.Yl = dprev; ch.Vl = vprev; ch.Wl = uprev;
    {
        const int e = M - 1;
        const Row r = make_row<CMODE, EXTRA>(ch.code(e), lo, hi, ch.T[e], CMODE == 2 ? ops.coef(e) : 0.0,
                                             EXTRA ? ops.q(e) : 0.0, EXTRA ? ops.dirv(e) : 0.0, k);
        ch.s_aa = r.aa; ch.s_cc = r.cc; ch.s_b = r.b; ch.s_d = r.d;
    }
    return f;
}

template <int M, bool EXTRA, class OPS>
ADI_HD void chunk_backward_hybrid(Chunk<M> &ch, const UniConst &uc, OPS &ops, int R4, unsigned lo, unsigned hi,
                                  double g, double Sl, double S)
{
    // tail: forward elimination once more with the true value left of it, d'_{R4-1} + vp_{R4-1}*S_{p-1}
    double dprev = Sl;
#pragma unroll
    for (int g4 = 1; g4 < M / 4; ++g4) {
        if (4 * g4 == R4) dprev = fma(uc.vp[4 * g4 - 1], Sl, ch.T[4 * g4 - 1]);
        if (4 * g4 >= R4) {
#pragma unroll
            for (int e = 4 * g4; e < 4 * g4 + 4; ++e) {
                if (e == M - 1) break;
                const double aa = sel(couples<EXTRA>(ch.code(e), lo), g, 0.0);
                const double dp = fma(aa, dprev, ch.T[e]) * ops.rinv(e);
                ch.T[e] = dp;
                dprev = dp;
            }
        }
    }
    double xn = S;
    ch.T[M - 1] = S;
#pragma unroll
    for (int g4 = M / 4 - 1; g4 >= 0; --g4) {
        if (4 * g4 >= R4) {
#pragma unroll
            for (int e = 4 * g4 + 3; e >= 4 * g4; --e) {
                if (e == M - 1) continue;
                const double u = sel(couples<EXTRA>(ch.code(e), hi), g, 0.0) * ops.rinv(e);
                const double x = fma(u, xn, ch.T[e]);
                ch.T[e] = x;
                xn = x;
            }
        } else {
#pragma unroll
            for (int e = 4 * g4 + 3; e >= 4 * g4; --e) {
                const double x = fma(uc.u[e], xn, fma(uc.vp[e], Sl, ch.T[e]));
                ch.T[e] = x;
                xn = x;
            }
        }
    }
}

// Phase 2a: the separator row of this chunk given the next chunk's First relation
// (zeros when there is no next chunk).  Returns the normalised reduced row.
template <int M>
ADI_HD Red chunk_reduced_row(const Chunk<M> &ch, const First &nx)
{
    // -aa_s x_{M-2} + b_s S_p - cc_s x_0^{(p+1)} = d_s
    // x_{M-2} = Yl + Vl S_{p-1} + Wl S_p,   x_0^{(p+1)} = nx.Y + nx.V S_p + nx.W S_{p+1}
    const double B = fma(-ch.s_cc, nx.V, fma(-ch.s_aa, ch.Wl, ch.s_b));
    const double D = fma(ch.s_cc, nx.Y, fma(ch.s_aa, ch.Yl, ch.s_d));
    const double rB = frcp(B);
    Red r;
    r.A = -(ch.s_aa * ch.Vl) * rB;
    r.C = -(ch.s_cc * nx.W) * rB;
    r.D = D * rB;
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

// ---- z-slab decomposition (multi-GPU z sweep) ---------------------------------------------
// A rank holds the cells [z0, z1) of every z line.  Its segment couples to the last cell of the
// rank below (ghost value L) and to the first cell of the rank above (ghost value R) through
// the ZM bit of its first cell and the ZP bit of its last cell (the neighbour code is built
// with the neighbours' mask planes).  Pass 1 carries the separators as affine functions of
// the two ghosts, S_p = D + DL*L + DR*R, through the same PCR: three right-hand-side columns.
struct Red3 {
    double A, C, D, DL, DR;
};

// Reduced row of chunk p of P with the ghost couplings moved to the right-hand side columns.
ADI_HD Red3 reduced_row3(const Red &r, int p, int P)
{
    Red3 q;
    q.A = r.A; q.C = r.C; q.D = r.D; q.DL = 0.0; q.DR = 0.0;
    if (p == 0) { q.DL = -r.A; q.A = 0.0; }          // A_0 multiplies S_{-1} = L
    if (p == P - 1) { q.DR = -r.C; q.C = 0.0; }      // C_{P-1} multiplies the ghost R
    return q;
}

ADI_HD Red3 pcr_step3(const Red3 &me, const Red3 &lo, const Red3 &hi)
{
    const double B = fma(-me.C, hi.A, fma(-me.A, lo.C, 1.0));
    const double rB = frcp(B);
    Red3 r;
    r.A = (-me.A * lo.A) * rB;
    r.C = (-me.C * hi.C) * rB;
    r.D = fma(-me.C, hi.D, fma(-me.A, lo.D, me.D)) * rB;
    r.DL = fma(-me.C, hi.DL, fma(-me.A, lo.DL, me.DL)) * rB;
    r.DR = fma(-me.C, hi.DR, fma(-me.A, lo.DR, me.DR)) * rB;
    return r;
}

// The `First` relation a ghost cell presents to the last chunk: x = 0 + 0*S_p + 1*R.
ADI_HD First ghost_first()
{
    First f;
    f.Y = 0.0; f.V = 0.0; f.W = 1.0;
    return f;
}

// Interface relation of a rank's segment of one line:
//   x_first = yf + vf*L + wf*R,   x_last = yl + vl*L + wl*R.
struct Iface {
    double yf, vf, wf, yl, vl, wl;
};

// Inter-rank system of one line (R ranks, 2R unknowns), solved by every rank for its own two
// ghosts.  get(r) returns rank r's Iface.  L = x_last of rank-1, Rg = x_first of rank+1
// (0 beyond the ends: the couplings there are zero).
template <class GET>
ADI_HD void iface_solve(GET get, int nranks, int rank, double *Lout, double *Rout)
{
    // Unknowns per rank r: f_r = x_first, l_r = x_last, with
    //   f_r = yf + vf*l_{r-1} + wf*f_{r+1},   l_r = yl + vl*l_{r-1} + wl*f_{r+1}.
    // From the left, ranks 0..rank-1 give  l_{rank-1} = al + be*f_rank;  from the right, ranks
    // nranks-1..rank+1 give  f_{rank+1} = ga + de*l_rank;  a 2x2 system closes at `rank`.
    // No arrays: one pass from each side (this runs per line inside the sweep kernels).
    double al = 0.0, be = 0.0;
    for (int r = 0; r < rank; ++r) {
        const Iface q = get(r);
        const double rd = frcp(1.0 - q.vf * be);
        const double g = (q.yf + q.vf * al) * rd;      // f_r = g + d*f_{r+1}
        const double d = q.wf * rd;
        const double al2 = q.yl + q.vl * (al + be * g);
        be = q.wl + q.vl * be * d;
        al = al2;
    }
    double ga = 0.0, de = 0.0;
    for (int r = nranks - 1; r > rank; --r) {
        const Iface q = get(r);
        const double rd = frcp(1.0 - q.wl * de);
        const double lc = (q.yl + q.wl * ga) * rd;     // l_r = lc + lv*l_{r-1}
        const double lv = q.vl * rd;
        const double ga2 = q.yf + q.wf * (ga + de * lc);
        de = q.vf + q.wf * de * lv;
        ga = ga2;
    }
    const Iface q = get(rank);
    const double a11 = 1.0 - q.vf * be, a12 = -q.wf * de, b1 = q.yf + q.vf * al + q.wf * ga;
    const double a21 = -q.vl * be, a22 = 1.0 - q.wl * de, b2 = q.yl + q.vl * al + q.wl * ga;
    const double rdet = frcp(a11 * a22 - a12 * a21);
    const double fr = (b1 * a22 - a12 * b2) * rdet;
    const double lr = (a11 * b2 - a21 * b1) * rdet;
    *Lout = al + be * fr;
    *Rout = ga + de * lr;
}

// Phase 3: Sl = S_{p-1} (0 for the first chunk), S = S_p.  Leaves the solution in ch.T
// (0 in void cells, which the caller does not store).
template <int M, bool EXTRA, int NS, class OPS>
ADI_HD void chunk_backward(Chunk<M> &ch, OPS &ops, unsigned lo, unsigned hi, double g, double Sl, double S)
{
    double dprev = Sl;
#pragma unroll
    for (int e = 0; e < M - 1; ++e) {
        double dp;
        if (NS == 2) {
            dp = fma(ops.la(e), dprev, ch.T[e]);
        } else {
            const double aa = sel(couples<EXTRA>(ch.code(e), lo), g, 0.0);
            dp = fma(aa, dprev, ch.T[e]) * ops.rinv(e);
        }
        ch.T[e] = dp;
        dprev = dp;
    }
    double xn = S;   // a void separator solves to exactly 0 (identity row, zero rhs)
    ch.T[M - 1] = S;
#pragma unroll
    for (int e = M - 2; e >= 0; --e) {
        double u;
        if (NS == 2) u = ops.u(e);
        else u = sel(couples<EXTRA>(ch.code(e), hi), g, 0.0) * ops.rinv(e);
        const double x = fma(u, xn, ch.T[e]);
        ch.T[e] = x;
        xn = x;
    }
}

// Number (as a double) of the two neighbour bits `bits` set in `code`.
ADI_HD double nb_count(unsigned code, unsigned bits)
{
#if defined(__CUDA_ARCH__)
    return (double)__popc(code & bits);
#else
    return (double)__builtin_popcount(code & bits);
#endif
}

// Explicit stage (adi3d_numba_coeff.py:240-288,298) for one ACTIVE cell:
// R0 = T + beta*((Lx+Ly)+Lz), L* = ((sum of active neighbours, '-' first) - cnt*T)/dx^2.
// The caller supplies neighbour values that are 0.0 where the code bit is clear.
ADI_HD double explicit_r0(unsigned code, double T, double xm, double xp, double ym, double yp,
                          double zm, double zp, const SweepConst &k)
{
    const double L0 = ((xm + xp) - nb_count(code, CB_XM | CB_XP) * T) * k.invdx2;
    const double L1 = ((ym + yp) - nb_count(code, CB_YM | CB_YP) * T) * k.invdx2;
    const double L2 = ((zm + zp) - nb_count(code, CB_ZM | CB_ZP) * T) * k.invdx2;
    return T + k.beta * ((L0 + L1) + L2);
}

}  // namespace adi
