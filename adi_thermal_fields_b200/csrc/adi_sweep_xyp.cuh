// adi_sweep_xyp.cuh -- K1p: x / y sweeps of long lines (1025..2048 cells), persistent blocks with a one-tile
// prefetch.
//
// A line of up to 2048 cells needs 64 chunks of 32 cells: 64 x 8 lanes = 512 threads at 128 registers is the
// whole register file of an SM, so k_sweep_xy runs ONE block per SM and its load, solve and store phases cannot
// hide behind another block's (measured on 2048 x 2048 x 128: x 2.69 / y 2.18 ms against 2.08 / 1.69 ms for the
// same loads and stores without the solve, and 1.28 ms at the copy peak).  Here the block stays on its SM and
// walks over tiles (grid = number of SMs); while it solves tile t, the field values and neighbour codes of
// tile t + gridDim.x travel into shared memory with cp.async:
//
//   * every thread prefetches exactly the 32 cells (+ 32 code bytes, + the line-end coefficients) it will own in
//     the next tile, into its OWN shared-memory column (col[e * NTH], the layout of the factor slots of
//     k_sweep_xy) -- no other thread ever touches those slots, so the hand-over needs no barrier, only the
//     thread's own cp.async.wait_all at the top of the next tile;
//   * warps on the tabulated uniform paths (the bulk) issue the prefetch right after reading their chunk out of
//     the column, i.e. before the elimination; warps on the general path keep 1/den in the same column (as
//     k_sweep_xy does) and issue it after their back substitution;
//   * results go straight from registers to global memory (fire-and-forget stores), so the stores of tile t
//     overlap the wait for tile t + gridDim.x.
//
// Only launched for sweeps that may take the uniform paths (no flux / Dirichlet operand, coefficient field
// scalar or verified surface-only); everything else keeps k_sweep_xy.
//
// Reference semantics: adi3d_numba_coeff.py:133-203 (sweep_axis0 / sweep_axis1).
#pragma once
#include "adi_sweep_xy.cuh"

namespace adi {

__device__ __forceinline__ void cp_async16u(unsigned dst_smem, const void *src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_smem), "l"(src) : "memory");
}

// smem: col[M][NTH] doubles | xch[6*NTH] doubles | codes [2][NTH] uint4 | cend [2][NTH] doubles (line-end coefficients)
template <int AXIS, int CMODE>
__global__ void __launch_bounds__(512, 1) k_sweep_xyp(const SweepArgs a, const int ntiles)
{
    constexpr int M = 32, NS = 1;
    constexpr bool EXTRA = false;
    extern __shared__ double smem[];
    const int KT = blockDim.x, P = blockDim.y;
    const int kk = threadIdx.x, p = threadIdx.y;
    const int NTH = KT * P;
    const int tid = p * KT + kk;
    const int n = (AXIS == 0) ? a.nx : a.ny;
    const unsigned sl = (AXIS == 0) ? (unsigned)a.ny * (unsigned)a.nz : (unsigned)a.nz;
    constexpr unsigned LO = (AXIS == 0) ? CB_XM : CB_YM;
    constexpr unsigned HI = (AXIS == 0) ? CB_XP : CB_YP;
    const int t0 = p * M;                                   // < n: the launcher uses P = ceil(n / M)
    const int nti = (a.nz + KT - 1) / KT;                   // tiles per index of the other strided axis
    const int el = n - 1 - t0;                              // slot of the line's last cell, if it is in this chunk
    const unsigned sl8 = sl * 8u;
    const unsigned nth8 = (unsigned)NTH * 8u;

    double *col = smem + tid;
    double *xch = smem + (size_t)M * NTH;
    uint4 *cslot = reinterpret_cast<uint4 *>(xch + (size_t)6 * NTH) + tid;      // word w at cslot[w * NTH]
    double *cend = reinterpret_cast<double *>(reinterpret_cast<uint4 *>(xch + (size_t)6 * NTH) + (size_t)2 * NTH) + tid;
    const unsigned scol = smem_u32(col), scode = smem_u32(cslot), send = smem_u32(cend);

    // position of tile t for this thread: first cell of its chunk, lane validity
    struct Pos {
        size_t idx0;
        size_t crow;      // offset of the chunk's codes in codeT
        bool lane_ok;
    };
    auto locate = [&](int t) -> Pos {
        const unsigned id = a.tiles ? (unsigned)a.tiles[t] : (unsigned)t;
        const unsigned by = id / (unsigned)nti, bx = id - by * (unsigned)nti;
        const int k = (int)bx * KT + kk;
        const int kc = min(k, a.nz - 1);
        Pos q;
        q.lane_ok = k < a.nz;
        q.idx0 = ((AXIS == 0) ? (size_t)by * a.nz : (size_t)by * a.ny * a.nz) + (size_t)kc + (size_t)t0 * sl;
        q.crow = ((size_t)by * a.nz + kc) * (size_t)a.npad + t0;
        return q;
    };
    // cells beyond the line's end re-read the chunk's last valid cell; their codes are the padding zeros of codeT
    const int nvl = min(n - t0, M) - 1;
    auto prefetch = [&](int t) {
        const Pos q = locate(t);
        const uint8_t *cb = a.codeT + q.crow;
        cp_async16u(scode, cb);
        cp_async16u(scode + (unsigned)NTH * 16u, cb + 16);
        const char *tb = reinterpret_cast<const char *>(a.in + q.idx0);
#pragma unroll
        for (int e = 0; e < M; ++e) cp_async8(scol + e * nth8, tb + (size_t)((unsigned)min(e, nvl) * sl8));
        if (CMODE == 2) {
            // surface-only coefficient field: the two ends of the line are always exposed when active
            const char *cf = reinterpret_cast<const char *>(a.coeff + q.idx0);
            if (t0 == 0) cp_async8(send, cf);
            if (el < M) cp_async8(send + nth8, cf + (size_t)((unsigned)el * sl8));
        }
    };

    int t = blockIdx.x;
    if (t < ntiles) prefetch(t);
    for (; t < ntiles; t += gridDim.x) {
        const Pos q = locate(t);
        const int tn = t + gridDim.x;
        const int nv = q.lane_ok ? nvl + 1 : 0;
        cp_async_wait_all();                                // this thread's own copies of tile t have landed
        Chunk<M> ch;
        {
            const uint4 v0 = cslot[0], v1 = cslot[NTH];
            ch.cw[0] = q.lane_ok ? v0.x : 0u; ch.cw[1] = q.lane_ok ? v0.y : 0u;
            ch.cw[2] = q.lane_ok ? v0.z : 0u; ch.cw[3] = q.lane_ok ? v0.w : 0u;
            ch.cw[4] = q.lane_ok ? v1.x : 0u; ch.cw[5] = q.lane_ok ? v1.y : 0u;
            ch.cw[6] = q.lane_ok ? v1.z : 0u; ch.cw[7] = q.lane_ok ? v1.w : 0u;
        }
#pragma unroll
        for (int e = 0; e < M; ++e) ch.T[e] = col[e * NTH];
        double ce0 = 0.0, ce1 = 0.0;
        if (CMODE == 2) { ce0 = cend[0]; ce1 = cend[NTH]; }
        bool any = false;
#pragma unroll
        for (int w = 0; w < M / 4; ++w) any = any || (ch.cw[w] & 0x01010101u) != 0u;
        // (also the barrier that separates this tile's use of xch from the previous tile's)
        const bool live = __syncthreads_or(any);
        if (a.in == a.out && !live) {                       // in place, a tile of void cells: nothing to solve
            if (tn < ntiles) prefetch(tn);
            continue;
        }
        const char *cf = reinterpret_cast<const char *>(a.coeff + (CMODE == 2 ? q.idx0 : 0));
        auto exposed_coef = [&](int e, unsigned c) -> double {
            if (CMODE != 2) return 0.0;
            if ((c & (LO | HI)) == (LO | HI)) return 0.0;
            const int cell = t0 + e;
            if (cell == 0) return ce0;
            if (cell == n - 1) return ce1;
            return ldg_f64(cf + (size_t)((unsigned)e * sl8));
        };

        // 0: general rows; 1: cells 0..M-2 uniform; 2: cell 0 general, cells 1..M-2 uniform (adi_core.h)
        int path = 0;
        if (a.uni) {
            if (__all_sync(0xffffffffu, chunk_uniform<M, 1>(ch, LO, HI)))
                path = __all_sync(0xffffffffu, chunk_uniform<M, 0>(ch, LO, HI)) ? 1 : 2;
        }
        StridedOps<M, true> ops;
        ops.coeff = nullptr; ops.qp = nullptr; ops.dvp = nullptr;
        ops.sl = sl; ops.nv = nv; ops.col = col; ops.NTH = NTH;
        First f;
        f.Y = f.V = f.W = 0.0;
        UniHead hd;
        hd.al = hd.bl = hd.br = 0.0;
        if (path != 0) {
            // the column is free again: the next tile starts travelling while this one is solved
            if (tn < ntiles) prefetch(tn);
            const unsigned cs = ch.code(M - 1), c0 = ch.code(0);
            const Row sep = make_row<CMODE, EXTRA>(cs, LO, HI, ch.T[M - 1], exposed_coef(M - 1, cs), 0.0, 0.0, a.k);
            if (path == 1) {
                f = chunk_forward_uniform<M, 0>(ch, a.uc, sep, sep, hd);
            } else {
                const Row head = make_row<CMODE, EXTRA>(c0, LO, HI, ch.T[0], exposed_coef(0, c0), 0.0, 0.0, a.k);
                f = chunk_forward_uniform<M, 1>(ch, a.uc, sep, head, hd);
            }
        } else {
#pragma unroll
            for (int e = 0; e < M; ++e) ch.T[e] = ch.active(e) ? ch.T[e] : 0.0;  // load rule (adi_core.h)
            if (CMODE == 2) {
#pragma unroll
                for (int e = 0; e < M; ++e) {
                    const unsigned c = ch.code(e);  // 0 beyond the chunk's valid cells
                    col[e * NTH] = (c & CB_SELF) ? exposed_coef(e, c) : 0.0;
                }
            }
            f = chunk_forward<M, CMODE, EXTRA, NS, false>(ch, ops, LO, HI, a.k);
        }
        double Sl;
        const double S = solve_reduced<M>(ch, f, xch, NTH, tid, KT, p, P, &Sl);
        if (path == 1) chunk_backward_uniform<M, 0>(ch, a.uc, hd, Sl, S);
        else if (path == 2) chunk_backward_uniform<M, 1>(ch, a.uc, hd, Sl, S);
        else {
            chunk_backward<M, EXTRA, NS>(ch, ops, LO, HI, a.k.g, Sl, S);
            if (tn < ntiles) prefetch(tn);                  // the factors are no longer needed
        }

        double *op = a.out + q.idx0;
        if (path != 0) {
#pragma unroll
            for (int e = 0; e < M; ++e) op[e * sl] = ch.T[e];
        } else if (a.in == a.out) {
#pragma unroll
            for (int e = 0; e < M; ++e)
                if (e < nv && ch.active(e)) op[e * sl] = ch.T[e];
        } else {
            const double *tp = a.in + q.idx0;
#pragma unroll
            for (int e = 0; e < M; ++e)
                if (e < nv) op[e * sl] = ch.active(e) ? ch.T[e] : tp[e * sl];
        }
    }
}

}  // namespace adi
