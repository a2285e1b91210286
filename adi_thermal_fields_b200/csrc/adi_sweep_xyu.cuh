// adi_sweep_xyu.cuh -- K1u: x / y sweeps of long lines (1025..2048 cells), ALL-UNIFORM tiles only.
//
// 64 chunks x 8 lanes = 512 threads; at the 128 registers k_sweep_xy needs (32 cells in registers + the general row
// assembly) that is the whole register file: one block per SM whose load, solve and store phases nothing overlaps
// (3.4 / 4.2 TB/s in the x / y sweep against 4.2 / 5.0 TB/s for 1024-cell lines, which run two blocks per SM with the
// same 64-byte rows).  Most tiles of a large grid need no row assembly at all: every chunk of every lane is a run of
// uniform cells (adi_core.h) -- k_tile_flags marks them when the neighbour code is built.  For those tiles this
// kernel keeps
//   * cells 0..15 of a chunk in registers and cells 16..31 in the thread's own shared-memory column (cp.async
//     straight from global memory; the tabulated elimination streams through them once forwards, once backwards),
//   * no neighbour codes beyond the first and last byte of the chunk, no factor store, no general path,
// which fits 64 registers and 88 KB per block: TWO blocks per SM.  The other active tiles go to k_sweep_xy as before
// (two complementary tile lists).
//
// Reference semantics: adi3d_numba_coeff.py:133-203 (sweep_axis0 / sweep_axis1).
#pragma once
#include "adi_sweep_xy.cuh"

namespace adi {

// smem: colT[16][NTH] doubles | xch[6*NTH] doubles
template <int AXIS, int CMODE, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB) k_sweep_xyu(const SweepArgs a)
{
    constexpr int M = 32, H = 16;
    constexpr bool EXTRA = false;
    extern __shared__ double smem[];
    const int KT = blockDim.x, P = blockDim.y;
    const int kk = threadIdx.x, p = threadIdx.y;
    const int NTH = KT * P;
    const int tid = p * KT + kk;
    const unsigned t = (unsigned)a.tiles[blockIdx.x];       // always launched from the list of all-uniform tiles
    const unsigned by = t / (unsigned)a.tiles_nx, bx = t - by * (unsigned)a.tiles_nx;
    const int k = bx * KT + kk;                             // < nz: an all-uniform tile has all its lanes
    const int n = (AXIS == 0) ? a.nx : a.ny;
    const unsigned sl = (AXIS == 0) ? (unsigned)a.ny * (unsigned)a.nz : (unsigned)a.nz;
    constexpr unsigned LO = (AXIS == 0) ? CB_XM : CB_YM;
    constexpr unsigned HI = (AXIS == 0) ? CB_XP : CB_YP;
    const int t0 = p * M;
    const size_t idx0 = ((AXIS == 0) ? (size_t)by * a.nz : (size_t)by * a.ny * a.nz) + (size_t)k + (size_t)t0 * sl;
    double *col = smem + tid;                               // cell 16 + j of this chunk: col[j * NTH]
    double *xch = smem + (size_t)H * NTH;
    const unsigned sl8 = sl * 8u, nth8 = (unsigned)NTH * 8u;
    const unsigned scol = smem_u32(col);
    const char *tb = reinterpret_cast<const char *>(a.in + idx0);

    // upper half straight into shared memory, lower half into registers, the chunk's first and last code
#pragma unroll
    for (int j = 0; j < H; ++j) cp_async8(scol + j * nth8, tb + (size_t)((unsigned)(H + j) * sl8));
    double T[H];
#pragma unroll
    for (int e = 0; e < H; ++e) T[e] = ldg_f64(tb + (size_t)((unsigned)e * sl8));
    const uint8_t *cb = a.codeT + ((size_t)by * a.nz + k) * (size_t)a.npad + t0;
    const unsigned c0 = ldg_u8(cb), cs = ldg_u8(cb + (M - 1));
    // surface-only coefficient field: in an all-uniform tile only the two ends of the line are exposed
    double ce0 = 0.0, ce1 = 0.0;
    if (CMODE == 2) {
        if (t0 == 0) ce0 = ldg_f64(a.coeff + idx0);
        if (t0 + M == n) ce1 = ldg_f64(reinterpret_cast<const char *>(a.coeff + idx0) + (size_t)((unsigned)(M - 1) * sl8));
    }
    cp_async_wait_all();

    auto Tat = [&](int e) -> double { return e < H ? T[e < H ? e : 0] : col[(e - H) * NTH]; };
    auto Tput = [&](int e, double v) { if (e < H) T[e < H ? e : 0] = v; else col[(e - H) * NTH] = v; };

    // the first chunk's cell 0 is the exposed first cell of the line: its warp (chunks 0..3) eliminates it by hand
    // (OFF 1, as in k_sweep_xy); every other warp's cell 0 is one more uniform cell
    const bool off1 = (tid >> 5) == 0;
    const Row sep = make_row<CMODE, EXTRA>(cs, LO, HI, Tat(M - 1), (t0 + M == n) ? ce1 : 0.0, 0.0, 0.0, a.k);
    Chunk<4> ch;    // carries the separator row and the last-interior relation into the reduced solve
    First f;
    UniHead hd;
    hd.al = hd.bl = hd.br = 0.0;
    const UniConst &uc = a.uc;
    if (!off1) {
        constexpr int N = M - 1;
        double dprev = 0.0, Y = 0.0;
#pragma unroll
        for (int j = 0; j < N; ++j) {
            const double dp = fma(uc.u[j], dprev, Tat(j) * uc.rinv[j]);
            Tput(j, dp);
            Y = fma(uc.al[j], dp, Y);
            dprev = dp;
        }
        f.Y = Y; f.V = uc.Vn[N]; f.W = uc.al[N];
        ch.Yl = dprev; ch.Vl = uc.vp[N - 1]; ch.Wl = uc.u[N - 1];
    } else {
        constexpr int N = M - 2;
        double dprev = 0.0, Y = 0.0;
#pragma unroll
        for (int j = 0; j < N; ++j) {
            const double dp = fma(uc.u[j], dprev, Tat(1 + j) * uc.rinv[j]);
            Tput(1 + j, dp);
            Y = fma(uc.al[j], dp, Y);
            dprev = dp;
        }
        const double V = uc.Vn[N], W = uc.al[N], Vl = uc.vp[N - 1], Wl = uc.u[N - 1];
        const Row head = make_row<CMODE, EXTRA>(c0, LO, HI, T[0], t0 == 0 ? ce0 : 0.0, 0.0, 0.0, a.k);
        const double r = frcp(fma(-head.cc, V, head.b));
        hd.al = fma(head.cc, Y, head.d) * r;
        hd.bl = head.aa * r;
        hd.br = (head.cc * W) * r;
        f.Y = hd.al; f.V = hd.bl; f.W = hd.br;
        ch.Yl = fma(Vl, hd.al, dprev); ch.Vl = Vl * hd.bl; ch.Wl = fma(Vl, hd.br, Wl);
    }
    ch.s_aa = sep.aa; ch.s_cc = sep.cc; ch.s_b = sep.b; ch.s_d = sep.d;

    double Sl;
    const double S = solve_reduced<4>(ch, f, xch, NTH, tid, KT, p, P, &Sl);

    double *op = a.out + idx0;
    op[(M - 1) * sl] = S;
    double xn = S;
    if (!off1) {
#pragma unroll
        for (int j = M - 2; j >= 0; --j) {
            const double x = fma(uc.u[j], xn, fma(uc.vp[j], Sl, Tat(j)));
            op[(size_t)j * sl] = x;
            xn = x;
        }
    } else {
        const double left = fma(hd.br, S, fma(hd.bl, Sl, hd.al));
#pragma unroll
        for (int j = M - 3; j >= 0; --j) {
            const double x = fma(uc.u[j], xn, fma(uc.vp[j], left, Tat(1 + j)));
            op[(size_t)(1 + j) * sl] = x;
            xn = x;
        }
        op[0] = left;
    }
}

}  // namespace adi
