// adi_sweep_xyp.cuh -- K1p: x / y sweeps of long lines (1025..2048 cells), persistent blocks, tiles prefetched
// by the TMA engine.
//
// A line of up to 2048 cells needs 64 chunks of 32 cells: 64 x 8 lanes = 512 threads at 128 registers is the
// whole register file of an SM, so k_sweep_xy runs ONE block per SM, and its 512 x 32 eight-byte loads per tile
// are 2048 separate 64-byte row requests that fill the load/store unit's queue (ncu r02n: "LG throttle" 10.1 /
// 6.7 stall cycles per issued instruction in the x / y sweep, 2.72 / 2.16 ms on 2048 x 2048 x 128 where the copy
// peak allows 1.28 ms).  Here
//
//   * the block stays on its SM and walks over tiles (grid = number of SMs);
//   * the field values of a tile travel as TENSOR copies (cp.async.bulk.tensor.4d, SASS UTMALDG) described by a
//     4-D tensor map over (z, chunk, cell in chunk, other strided axis): one instruction per WARP fetches the
//     box (8 lanes, the warp's 4 chunks, 32 cells) = 8 KB that the warp's 32 threads own, laid out in shared
//     memory as [cell][chunk][lane] -- the 32 values a warp reads together are 256 consecutive bytes;
//   * every warp prefetches its own box of the NEXT tile as soon as it has pulled the current one into
//     registers (warps on the general path keep 1/den in the same slots and prefetch after their back
//     substitution): the hand-over is warp-local -- a proxy fence, __syncwarp and the warp's own mbarrier --
//     and the load of tile t + gridDim.x overlaps the solve and the stores of tile t;
//   * neighbour codes (32 bytes per thread from the transposed code array) and the line-end coefficients follow
//     with cp.async into thread-owned slots; results go from registers to global memory (plain stores: they
//     need no registers to come back, so the load/store unit only ever queues stores).
//
// Only launched for sweeps that may take the uniform paths (no flux / Dirichlet operand, coefficient field
// scalar or verified surface-only) on lines whose length is a multiple of 128; everything else keeps k_sweep_xy.
//
// Reference semantics: adi3d_numba_coeff.py:133-203 (sweep_axis0 / sweep_axis1).
#pragma once
#include <cuda.h>

#include "adi_sweep_xy.cuh"

namespace adi {

__device__ __forceinline__ void cp_async16u(unsigned dst_smem, const void *src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_smem), "l"(src) : "memory");
}
__device__ __forceinline__ void xyp_mbar_init(unsigned bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void xyp_mbar_expect_tx(unsigned bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void xyp_mbar_wait(unsigned bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void xyp_tma_load4(unsigned dst_smem, const CUtensorMap *tm, unsigned bar, int c0, int c1, int c2, int c3)
{
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst_smem), "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}

constexpr int XYP_BOX_BYTES = 8 * 4 * 32 * 8;   // one warp's box: 8 lanes x 4 chunks x 32 cells of fp64

// smem: tile [NW][1024] doubles (8 KB per warp, 1 KB aligned) | xch[6*NTH] | codes [2][NTH] uint4 | cend [2][NTH] | mbarrier [NW]
// lay 0: tensor dimensions (z, chunk, cell, other): shared-memory slot of cell e = e*32 + c*8 + lane   (se 32, sc 8)
// lay 1: dimensions in memory order (host fallback when the driver refuses lay 0): slot = c*256 + e*8 + lane (se 8, sc 256)
template <int AXIS, int CMODE>
__global__ void __launch_bounds__(512, 1) k_sweep_xyp(const SweepArgs a, const __grid_constant__ CUtensorMap tm,
                                                      const int ntiles, const int lay, const int seq)
{
    constexpr int M = 32, NS = 1;
    constexpr bool EXTRA = false;
    extern __shared__ __align__(1024) double smem[];
    const int KT = 8, P = blockDim.y;
    const int kk = threadIdx.x, p = threadIdx.y;
    const int NTH = KT * P, NW = NTH >> 5;
    const int tid = p * KT + kk;
    const int warp = tid >> 5, lane = tid & 31, c = p & 3;
    const int n = (AXIS == 0) ? a.nx : a.ny;
    const unsigned sl = (AXIS == 0) ? (unsigned)a.ny * (unsigned)a.nz : (unsigned)a.nz;
    constexpr unsigned LO = (AXIS == 0) ? CB_XM : CB_YM;
    constexpr unsigned HI = (AXIS == 0) ? CB_XP : CB_YP;
    const int t0 = p * M;
    const int nti = (a.nz + KT - 1) / KT;                   // tiles per index of the other strided axis
    const int el = n - 1 - t0;                              // slot of the line's last cell, if it is in this chunk
    const unsigned sl8 = sl * 8u;
    const int se = lay ? 8 : 32, sc = lay ? 256 : 8;

    double *wtile = smem + (size_t)warp * 1024;
    double *col = wtile + c * sc + kk;                      // slot of cell e: col[e * se]
    double *xch = smem + (size_t)NW * 1024;
    uint4 *cslot = reinterpret_cast<uint4 *>(xch + (size_t)6 * NTH) + tid;      // word w at cslot[w * NTH]
    double *cend = reinterpret_cast<double *>(reinterpret_cast<uint4 *>(xch + (size_t)6 * NTH) + (size_t)2 * NTH) + tid;
    uint64_t *bars = reinterpret_cast<uint64_t *>(reinterpret_cast<double *>(reinterpret_cast<uint4 *>(xch + (size_t)6 * NTH) + (size_t)2 * NTH) + (size_t)2 * NTH);
    const unsigned swt = smem_u32(wtile), scode = smem_u32(cslot), send = smem_u32(cend), sbar = smem_u32(bars + warp);
    const unsigned nth8 = (unsigned)NTH * 8u;

    if (tid == 0) {
        for (int w = 0; w < NW; ++w) xyp_mbar_init(smem_u32(bars + w), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    struct Pos {
        size_t idx0;
        size_t crow;      // offset of the chunk's codes in codeT
        int bx, by;
        bool lane_ok;
    };
    auto locate = [&](int t) -> Pos {
        const unsigned id = a.tiles ? (unsigned)a.tiles[t] : (unsigned)t;
        const unsigned by = id / (unsigned)nti, bx = id - by * (unsigned)nti;
        const int k = (int)bx * KT + kk;
        const int kc = min(k, a.nz - 1);
        Pos q;
        q.bx = (int)bx; q.by = (int)by;
        q.lane_ok = k < a.nz;
        q.idx0 = ((AXIS == 0) ? (size_t)by * a.nz : (size_t)by * a.ny * a.nz) + (size_t)kc + (size_t)t0 * sl;
        q.crow = ((size_t)by * a.nz + kc) * (size_t)a.npad + t0;
        return q;
    };
    // Called by all lanes of a warp together, after their last access to the warp's box.
    auto prefetch = [&](int t) {
        const Pos q = locate(t);
        const uint8_t *cb = a.codeT + q.crow;
        cp_async16u(scode, cb);
        cp_async16u(scode + (unsigned)NTH * 16u, cb + 16);
        if (CMODE == 2) {
            // surface-only coefficient field: the two ends of the line are always exposed when active
            const char *cf = reinterpret_cast<const char *>(a.coeff + q.idx0);
            if (t0 == 0) cp_async8(send, cf);
            if (el < M) cp_async8(send + nth8, cf + (size_t)((unsigned)el * sl8));
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic accesses to the box come first
        __syncwarp();
        if (lane == 0) {
            xyp_mbar_expect_tx(sbar, (unsigned)XYP_BOX_BYTES);
            const int z0 = q.bx * KT, ch0 = 4 * warp;
            if (lay == 0) xyp_tma_load4(swt, &tm, sbar, z0, ch0, 0, q.by);
            else if (AXIS == 1) xyp_tma_load4(swt, &tm, sbar, z0, 0, ch0, q.by);
            else xyp_tma_load4(swt, &tm, sbar, z0, q.by, 0, ch0);
        }
    };

    // seq: every block walks over a CONTIGUOUS range of tiles (consecutive z tiles of the same lines: the 64-byte rows
    // of neighbouring tiles share 128-byte lines and DRAM pages, and the tensor map asks L2 to fetch 256 bytes at a
    // time); otherwise tiles are dealt round-robin
    unsigned parity = 0;
    const int per = (ntiles + (int)gridDim.x - 1) / (int)gridDim.x;
    const int tstep = seq ? 1 : (int)gridDim.x;
    const int tend = seq ? min(ntiles, ((int)blockIdx.x + 1) * per) : ntiles;
    int t = seq ? (int)blockIdx.x * per : (int)blockIdx.x;
    if (t < tend) prefetch(t);
    for (; t < tend; t += tstep, parity ^= 1u) {
        const Pos q = locate(t);
        const int tn = t + tstep;
        const int nv = q.lane_ok ? M : 0;
        cp_async_wait_all();                                // this thread's codes / line-end coefficients
        xyp_mbar_wait(sbar, parity);                        // the warp's box
        Chunk<M> ch;
        {
            const uint4 v0 = cslot[0], v1 = cslot[NTH];
            ch.cw[0] = q.lane_ok ? v0.x : 0u; ch.cw[1] = q.lane_ok ? v0.y : 0u;
            ch.cw[2] = q.lane_ok ? v0.z : 0u; ch.cw[3] = q.lane_ok ? v0.w : 0u;
            ch.cw[4] = q.lane_ok ? v1.x : 0u; ch.cw[5] = q.lane_ok ? v1.y : 0u;
            ch.cw[6] = q.lane_ok ? v1.z : 0u; ch.cw[7] = q.lane_ok ? v1.w : 0u;
        }
#pragma unroll
        for (int e = 0; e < M; ++e) ch.T[e] = col[e * se];
        double ce0 = 0.0, ce1 = 0.0;
        if (CMODE == 2) { ce0 = cend[0]; ce1 = cend[NTH]; }
        bool any = false;
#pragma unroll
        for (int w = 0; w < M / 4; ++w) any = any || (ch.cw[w] & 0x01010101u) != 0u;
        // (also the barrier that separates this tile's use of xch from the previous tile's)
        const bool live = __syncthreads_or(any);
        if (a.in == a.out && !live) {                       // in place, a tile of void cells: nothing to solve
            if (tn < tend) prefetch(tn);
            continue;
        }
        const char *cf = reinterpret_cast<const char *>(a.coeff + (CMODE == 2 ? q.idx0 : 0));
        auto exposed_coef = [&](int e, unsigned cd) -> double {
            if (CMODE != 2) return 0.0;
            if ((cd & (LO | HI)) == (LO | HI)) return 0.0;
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
        ops.sl = sl; ops.nv = nv; ops.col = col; ops.NTH = se;
        First f;
        f.Y = f.V = f.W = 0.0;
        UniHead hd;
        hd.al = hd.bl = hd.br = 0.0;
        if (path != 0) {
            // the box is free again: the next tile starts travelling while this one is solved
            if (tn < tend) prefetch(tn);
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
                    const unsigned cd = ch.code(e);
                    col[e * se] = (cd & CB_SELF) ? exposed_coef(e, cd) : 0.0;
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
            if (tn < tend) prefetch(tn);                  // the factors are no longer needed
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

// Host: tensor map of the field `base` for the boxes above.  lay 0 = (z, chunk, cell, other); returns the layout
// that the driver accepted in *lay (1 = dimensions in memory order), or an error.
inline int xyp_tensor_map(CUtensorMap *tm, int *lay, int axis, const double *base, int nx, int ny, int nz, int promo)
{
    typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static encode_fn encode = nullptr;
    if (!encode) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) {
            set_error("adi_cart_step: cuTensorMapEncodeTiled is not available");
            return ADI_ESTATE;
        }
        encode = (encode_fn)fn;
    }
    const cuuint64_t n = axis == 0 ? nx : ny, other = axis == 0 ? ny : nx;
    const cuuint64_t s_cell = (axis == 0 ? (cuuint64_t)ny * nz : (cuuint64_t)nz) * 8ull;
    const cuuint64_t s_other = (axis == 0 ? (cuuint64_t)nz : (cuuint64_t)ny * nz) * 8ull;
    const cuuint32_t es[4] = {1, 1, 1, 1};
    for (int l = *lay; l < 2; ++l) {
        cuuint64_t dim[4], str[3];
        cuuint32_t box[4];
        dim[0] = (cuuint64_t)nz; box[0] = 8;
        if (l == 0) {
            dim[1] = n / 32; str[0] = 32 * s_cell; box[1] = 4;
            dim[2] = 32;     str[1] = s_cell;      box[2] = 32;
            dim[3] = other;  str[2] = s_other;     box[3] = 1;
        } else if (axis == 1) {
            dim[1] = 32;     str[0] = s_cell;      box[1] = 32;
            dim[2] = n / 32; str[1] = 32 * s_cell; box[2] = 4;
            dim[3] = other;  str[2] = s_other;     box[3] = 1;
        } else {
            dim[1] = other;  str[0] = s_other;     box[1] = 1;
            dim[2] = 32;     str[1] = s_cell;      box[2] = 32;
            dim[3] = n / 32; str[2] = 32 * s_cell; box[3] = 4;
        }
        const CUresult r = encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, const_cast<double *>(base), dim, str, box, es,
                                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                  promo == 256 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : promo == 128 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : promo == 64 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B : CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r == CUDA_SUCCESS) { *lay = l; return ADI_OK; }
    }
    set_error("adi_cart_step: cuTensorMapEncodeTiled refused the sweep's tensor map");
    return ADI_ESTATE;
}

}  // namespace adi
