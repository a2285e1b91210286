// adi_sweep_xy.cuh -- K1x: x / y sweeps (strided axes), second generation.
//
// Same tiling as K1 (k_sweep_strided): blockDim = (KT lanes along z, P chunks of M cells), one thread
// owns one chunk of one line in registers.  What is new:
//
//  * neighbour codes come from a per-axis TRANSPOSED copy of the code array (line axis fastest, lines padded
//    to a multiple of 32 with code 0): the M codes of a chunk are M contiguous bytes -- one 16-byte load
//    per 16 cells instead of 16 strided byte loads, and no bounds logic (padding cells are void cells);
//  * UNIFORM chunks (adi_core.h: all cells active with both neighbours along the axis, no Dirichlet / flux
//    operand, coefficient field known to vanish there) take tabulated elimination factors from the kernel
//    parameters: 3 + 2 fused multiply-adds per cell, no reciprocal, no factor store in shared memory.
//    Decided per warp; every other warp runs the general row assembly of K1;
//  * the reduced system (one separator per chunk) is solved by WARPS: each chunk publishes seven numbers
//    in a transposed, padded shared-memory table, one warp per line does the parallel cyclic reduction
//    with shuffles (1 or 2 rows per lane: up to 64 chunks per line) -- two block barriers per tile
//    instead of 2 + ceil(log2 P).
//
// Reference semantics: adi3d_numba_coeff.py:133-203 (sweep_axis0 / sweep_axis1).
#pragma once
#include "adi_cart.cuh"

namespace adi {

__device__ __forceinline__ uint4 ldg_u128(const void *p)
{
    uint4 v;
    asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}

// Row i+s / i-s of a PR-rows-per-lane distribution (row i = j*32 + lane).  Out-of-range rows are masked by
// the caller.
template <int PR>
__device__ __forceinline__ void rows_shift_up(const double (&v)[PR], int s, int lane, double (&out)[PR])
{
    // out[j] = row (j*32 + lane) + s
    const int sj = s >> 5, sl = s & 31;
    const int src = (lane + sl) & 31;
    const bool wrap = lane + sl >= 32;
#pragma unroll
    for (int j = 0; j < PR; ++j) {
        const int ja = j + sj, jb = j + sj + 1;
        double a = 0.0, b = 0.0;
        if (sl == 0) {
            a = ja < PR ? v[ja < PR ? ja : 0] : 0.0;
        } else {
            // both candidates are fetched with the same source lane; one of them is kept
            const double va = ja < PR ? v[ja < PR ? ja : 0] : 0.0;
            const double vb = jb < PR ? v[jb < PR ? jb : 0] : 0.0;
            a = __shfl_sync(0xffffffffu, va, src);
            b = __shfl_sync(0xffffffffu, vb, src);
        }
        out[j] = (sl != 0 && wrap) ? b : a;
    }
}

template <int PR>
__device__ __forceinline__ void rows_shift_dn(const double (&v)[PR], int s, int lane, double (&out)[PR])
{
    // out[j] = row (j*32 + lane) - s
    const int sj = s >> 5, sl = s & 31;
    const int src = (lane - sl) & 31;
    const bool wrap = lane - sl < 0;
#pragma unroll
    for (int j = 0; j < PR; ++j) {
        const int ja = j - sj, jb = j - sj - 1;
        double a = 0.0, b = 0.0;
        if (sl == 0) {
            a = ja >= 0 ? v[ja >= 0 ? ja : 0] : 0.0;
        } else {
            const double va = ja >= 0 ? v[ja >= 0 ? ja : 0] : 0.0;
            const double vb = jb >= 0 ? v[jb >= 0 ? jb : 0] : 0.0;
            a = __shfl_sync(0xffffffffu, va, src);
            b = __shfl_sync(0xffffffffu, vb, src);
        }
        out[j] = (sl != 0 && wrap) ? b : a;
    }
}

// Reduced-system solve, transposed exchange + warp PCR.  xch: 7 planes of P*(KT+1) doubles.
// Needs NTH >= 32 and P <= 32*PR; partial warps (NTH % 32 != 0) take no part in the PCR.
template <int M, int PR>
__device__ __forceinline__ double solve_reduced_tw(const Chunk<M> &ch, const First &f, double *xch, int KT, int P,
                                                   int kk, int p, int tid, int NTH, double *Sl)
{
    const int LD = KT + 1, PL = P * LD;
    const int o = p * LD + kk;
    xch[o] = f.Y;
    xch[PL + o] = f.V;
    xch[2 * PL + o] = f.W;
    xch[3 * PL + o] = fma(-ch.s_aa, ch.Wl, ch.s_b);     // diagonal without the next chunk's term
    xch[4 * PL + o] = fma(ch.s_aa, ch.Yl, ch.s_d);      // right-hand side without the next chunk's term
    xch[5 * PL + o] = -(ch.s_aa * ch.Vl);               // coefficient of S_{p-1}
    xch[6 * PL + o] = ch.s_cc;
    __syncthreads();
    const int lane = tid & 31, warp = tid >> 5, nwarps = NTH >> 5;
    if (warp < nwarps) {
        int W2 = 32;
        if (PR == 1) while (W2 > 1 && (W2 >> 1) >= P) W2 >>= 1;    // smallest power of two >= P
        const int LPP = 32 / W2;                                  // lines per warp pass
        const int q0 = lane & (W2 - 1), sub = lane / W2;
        for (int l0 = warp * LPP; l0 < KT; l0 += nwarps * LPP) {
            const int line = l0 + sub;
            double Y[PR], V[PR], W[PR], Bp[PR], Dp[PR], Ap[PR], cc[PR];
            bool ok[PR];
#pragma unroll
            for (int j = 0; j < PR; ++j) {
                const int q = q0 + 32 * j;
                ok[j] = line < KT && q < P;
                const int oo = ok[j] ? q * LD + line : 0;
                Y[j] = xch[oo]; V[j] = xch[PL + oo]; W[j] = xch[2 * PL + oo];
                Bp[j] = xch[3 * PL + oo]; Dp[j] = xch[4 * PL + oo]; Ap[j] = xch[5 * PL + oo]; cc[j] = xch[6 * PL + oo];
            }
            double nY[PR], nV[PR], nW[PR];
            if (PR == 1) {
                nY[0] = __shfl_down_sync(0xffffffffu, Y[0], 1, W2);
                nV[0] = __shfl_down_sync(0xffffffffu, V[0], 1, W2);
                nW[0] = __shfl_down_sync(0xffffffffu, W[0], 1, W2);
            } else {
                rows_shift_up<PR>(Y, 1, lane, nY);
                rows_shift_up<PR>(V, 1, lane, nV);
                rows_shift_up<PR>(W, 1, lane, nW);
            }
            double rA[PR], rC[PR], rD[PR];
#pragma unroll
            for (int j = 0; j < PR; ++j) {
                const int q = q0 + 32 * j;
                const bool nxt = q + 1 < P;
                const double B = nxt ? fma(-cc[j], nV[j], Bp[j]) : Bp[j];
                const double D = nxt ? fma(cc[j], nY[j], Dp[j]) : Dp[j];
                const double rB = frcp(ok[j] ? B : 1.0);
                rA[j] = ok[j] ? Ap[j] * rB : 0.0;
                rC[j] = (ok[j] && nxt) ? -(cc[j] * nW[j]) * rB : 0.0;
                rD[j] = ok[j] ? D * rB : 0.0;
            }
            for (int s = 1; s < P; s <<= 1) {
                double lA[PR], lC[PR], lD[PR], hA[PR], hC[PR], hD[PR];
                if (PR == 1) {
                    lA[0] = __shfl_up_sync(0xffffffffu, rA[0], s, W2);
                    lC[0] = __shfl_up_sync(0xffffffffu, rC[0], s, W2);
                    lD[0] = __shfl_up_sync(0xffffffffu, rD[0], s, W2);
                    hA[0] = __shfl_down_sync(0xffffffffu, rA[0], s, W2);
                    hC[0] = __shfl_down_sync(0xffffffffu, rC[0], s, W2);
                    hD[0] = __shfl_down_sync(0xffffffffu, rD[0], s, W2);
                } else {
                    rows_shift_dn<PR>(rA, s, lane, lA); rows_shift_dn<PR>(rC, s, lane, lC); rows_shift_dn<PR>(rD, s, lane, lD);
                    rows_shift_up<PR>(rA, s, lane, hA); rows_shift_up<PR>(rC, s, lane, hC); rows_shift_up<PR>(rD, s, lane, hD);
                }
#pragma unroll
                for (int j = 0; j < PR; ++j) {
                    const int q = q0 + 32 * j;
                    Red me, lo, hi;
                    me.A = rA[j]; me.C = rC[j]; me.D = rD[j];
                    const bool hl = q - s >= 0, hh = q + s < P;
                    lo.A = hl ? lA[j] : 0.0; lo.C = hl ? lC[j] : 0.0; lo.D = hl ? lD[j] : 0.0;
                    hi.A = hh ? hA[j] : 0.0; hi.C = hh ? hC[j] : 0.0; hi.D = hh ? hD[j] : 0.0;
                    const Red r = pcr_step(me, lo, hi);
                    rA[j] = r.A; rC[j] = r.C; rD[j] = r.D;
                }
            }
#pragma unroll
            for (int j = 0; j < PR; ++j)
                if (ok[j]) xch[(q0 + 32 * j) * LD + line] = rD[j];
        }
    }
    __syncthreads();
    *Sl = p > 0 ? xch[o - LD] : 0.0;
    return xch[o];
}

template <int AXIS, int M, int NS, int CMODE, bool EXTRA, int PR, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB) k_sweep_xy(const SweepArgs a)
{
    static_assert(M % 16 == 0, "codes are loaded sixteen at a time");
    extern __shared__ double smem[];
    const int KT = blockDim.x, P = blockDim.y;
    const int kk = threadIdx.x;
    // chunk of this thread: with `remap` the two ends of the line share a warp (threadIdx.y = 0, 1, 2, 3 ->
    // chunks 0, P-1, 1, P-2, ...), so that a part spanning the whole line has ONE warp per block on the
    // general path (its end chunks are exposed) instead of two
    const int p = a.remap ? ((threadIdx.y & 1) ? P - 1 - (int)(threadIdx.y >> 1) : (int)(threadIdx.y >> 1)) : (int)threadIdx.y;
    const int NTH = KT * P;
    const int tid = threadIdx.y * KT + kk;
    unsigned bx = blockIdx.x, by = blockIdx.y;
    if (a.tiles) {   // active-tile list: the block's tile comes from the list
        const unsigned t = (unsigned)a.tiles[blockIdx.x >> a.tsplit];
        by = t / (unsigned)a.tiles_nx; bx = t - by * (unsigned)a.tiles_nx;
        if (a.tsplit) bx = 2u * bx + (blockIdx.x & 1u);     // the list counts tiles of 2*KT lanes: two blocks per entry
    }
    const int k = bx * KT + kk;
    const int n = (AXIS == 0) ? a.nx : a.ny;
    const unsigned sl = (AXIS == 0) ? (unsigned)a.ny * (unsigned)a.nz : (unsigned)a.nz;
    constexpr unsigned LO = (AXIS == 0) ? CB_XM : CB_YM;
    constexpr unsigned HI = (AXIS == 0) ? CB_XP : CB_YP;
    const int t0 = p * M;                                   // < n: the launcher uses P = ceil(n / M)
    const bool lane_ok = k < a.nz;
    const int nv = lane_ok ? min(n - t0, M) : 0;
    const int kc = min(k, a.nz - 1);
    const size_t idx0 = ((AXIS == 0) ? (size_t)by * a.nz : (size_t)by * a.ny * a.nz) +
                        (size_t)kc + (size_t)t0 * sl;
    double *col = smem + tid;
    double *xch = smem + (size_t)NS * M * NTH;              // behind the factor slots
    const double *tp = a.in + idx0;

    Chunk<M> ch;
    const int nvm1 = max(nv - 1, 0);
    const unsigned sl8 = sl * 8u;
    const char *tb = reinterpret_cast<const char *>(tp);
    {
        const uint8_t *cb = a.codeT + ((size_t)by * a.nz + kc) * (size_t)a.npad + t0;
#pragma unroll
        for (int w = 0; w < M / 16; ++w) {
            const uint4 v = ldg_u128(cb + 16 * w);
            ch.cw[4 * w] = lane_ok ? v.x : 0u; ch.cw[4 * w + 1] = lane_ok ? v.y : 0u;
            ch.cw[4 * w + 2] = lane_ok ? v.z : 0u; ch.cw[4 * w + 3] = lane_ok ? v.w : 0u;
        }
    }
#pragma unroll
    for (int e = 0; e < M; ++e) ch.T[e] = ldg_f64(tb + (size_t)((unsigned)min(e, nvm1) * sl8));
    const unsigned scol = smem_u32(col);
    const unsigned nth8 = (unsigned)NTH * 8u;
    const char *cf = reinterpret_cast<const char *>(a.coeff + (CMODE == 2 ? idx0 : 0));
    const int el = n - 1 - t0;                              // slot of the line's last cell, if it is in this chunk
    if (CMODE == 2) {
        if (!a.sparse) {
#pragma unroll
            for (int e = 0; e < M; ++e) cp_async8(scol + e * nth8, cf + (size_t)((unsigned)min(e, nvm1) * sl8));
        } else if (nv > 0) {
            // surface-only coefficient field: the two ends of the line (always exposed when active) are requested
            // now, cells next to an interior void once the codes have arrived
            if (t0 == 0) cp_async8(scol, cf);
            if (el < M) cp_async8(scol + (unsigned)el * nth8, cf + (size_t)((unsigned)el * sl8));
        }
    }
    // The barrier keeps every load above it in flight together (one memory round trip per thread) and tells
    // whether the tile holds an active cell at all: in place, a tile of void cells has nothing to solve.
    bool any = false;
#pragma unroll
    for (int w = 0; w < M / 4; ++w) any = any || (ch.cw[w] & 0x01010101u) != 0u;
    const bool live = __syncthreads_or(any);
    cp_async_wait_all();
    if (a.in == a.out && !live) return;

    // coefficient of an ACTIVE cell e of this chunk when only exposed cells carry one (CMODE 1 / sparse CMODE 2 are
    // the modes of the uniform paths; make_row derives the CMODE 1 value from the code itself)
    auto exposed_coef = [&](int e, unsigned c) -> double {
        if (CMODE != 2) return 0.0;
        if ((c & (LO | HI)) == (LO | HI)) return 0.0;
        const int cell = t0 + e;
        if (cell == 0 || cell == n - 1) return col[e * NTH];
        return ldg_f64(cf + (size_t)((unsigned)e * sl8));
    };

    // 0: general rows; 1: cells 0..M-2 uniform; 2: cell 0 general, cells 1..M-2 uniform (adi_core.h)
    int path = 0;
    if (a.uni) {
        if (__all_sync(0xffffffffu, chunk_uniform<M, 1>(ch, LO, HI)))
            path = __all_sync(0xffffffffu, chunk_uniform<M, 0>(ch, LO, HI)) ? 1 : 2;
    }
    bool solid = path != 0;
    StridedOps<M, true> ops;
    ops.coeff = nullptr; ops.qp = nullptr; ops.dvp = nullptr;
    ops.sl = sl; ops.nv = nv; ops.col = col; ops.NTH = NTH;
    First f;
    f.Y = f.V = f.W = 0.0;
    UniHead hd;
    hd.al = hd.bl = hd.br = 0.0;
    if (a.dbg == 1) {
    } else if (path != 0) {
        const unsigned cs = ch.code(M - 1), c0 = ch.code(0);
        const Row sep = make_row<CMODE, EXTRA>(cs, LO, HI, ch.T[M - 1], exposed_coef(M - 1, cs), 0.0, 0.0, a.k);
        if (path == 1) {
            f = chunk_forward_uniform<M, 0>(ch, a.uc, sep, sep, hd);
        } else {
            const Row head = make_row<CMODE, EXTRA>(c0, LO, HI, ch.T[0], exposed_coef(0, c0), 0.0, 0.0, a.k);
            f = chunk_forward_uniform<M, 1>(ch, a.uc, sep, head, hd);
        }
    } else {
        solid = NS == 2 && __all_sync(0xffffffffu, nv == M && chunk_solid<M>(ch, LO, HI));
        if (!solid) {
#pragma unroll
            for (int e = 0; e < M; ++e) ch.T[e] = ch.active(e) ? ch.T[e] : 0.0;  // load rule (adi_core.h)
        }
        const bool ends_only = CMODE == 2 && a.sparse && solid;
        if (CMODE == 2 && a.sparse) {
#pragma unroll
            for (int e = 0; e < M; ++e) {
                if (solid && e != 0 && e != M - 1) continue;
                const unsigned c = ch.code(e);  // 0 beyond the chunk's valid cells
                const int cell = t0 + e;
                if (e < nv && (cell == 0 || cell == n - 1)) continue;  // staged above
                col[e * NTH] = ((c & CB_SELF) && (c & (LO | HI)) != (LO | HI)) ? ldg_f64(cf + (size_t)((unsigned)e * sl8)) : 0.0;
            }
        }
        ops.qp = (EXTRA && a.q) ? a.q + idx0 : nullptr;
        ops.dvp = (EXTRA && a.dirv) ? a.dirv + idx0 : nullptr;
        if (ends_only) f = chunk_forward<M, CMODE, EXTRA, NS, true, CMODE == 2>(ch, ops, LO, HI, a.k);
        else if (solid) f = chunk_forward<M, CMODE, EXTRA, NS, true>(ch, ops, LO, HI, a.k);
        else f = chunk_forward<M, CMODE, EXTRA, NS, false>(ch, ops, LO, HI, a.k);
    }
    if (a.dbg != 1) {   // dbg 1: data movement only (tuning aid; wrong results)
        double Sl, S;
        if (a.tw) S = solve_reduced_tw<M, PR>(ch, f, xch, KT, P, kk, p, tid, NTH, &Sl);
        else S = solve_reduced<M>(ch, f, xch, NTH, p * KT + kk, KT, p, P, &Sl);
        if (path == 1) chunk_backward_uniform<M, 0>(ch, a.uc, hd, Sl, S);
        else if (path == 2) chunk_backward_uniform<M, 1>(ch, a.uc, hd, Sl, S);
        else chunk_backward<M, EXTRA, NS>(ch, ops, LO, HI, a.k.g, Sl, S);
    }

    double *op = a.out + idx0;
    if (solid) {
#pragma unroll
        for (int e = 0; e < M; ++e) op[e * sl] = ch.T[e];
    } else if (a.in == a.out) {
#pragma unroll
        for (int e = 0; e < M; ++e)
            if (e < nv && ch.active(e)) op[e * sl] = ch.T[e];
    } else {
#pragma unroll
        for (int e = 0; e < M; ++e)
            if (e < nv) op[e * sl] = ch.active(e) ? ch.T[e] : tp[e * sl];
    }
}

}  // namespace adi
