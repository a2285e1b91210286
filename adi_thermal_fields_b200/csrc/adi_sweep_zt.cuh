// adi_sweep_zt.cuh -- K3t: z sweep (contiguous axis), second generation.
//
// The first-generation z sweep (k_sweep_z) gives every line to one warp, lanes along the line.  Whether a
// chunk may take the tabulated "uniform run" path (adi_core.h) is decided per warp, so there a single exposed
// cell inside a line -- the top of a plate -- sends the whole line down the general path.  Here the tile is
// KT consecutive lines x P chunks and the lanes of a warp run ACROSS the lines (threadIdx.x = line), exactly as
// in the x / y sweeps (adi_sweep_xy.cuh): neighbouring lines look alike, so the warps holding the bulk of a
// part are uniform and only the warps that hold the surface chunks assemble general rows.
//
//  * the tile (KT lines of nz contiguous cells) travels as BULK ASYNCHRONOUS COPIES (cp.async.bulk, the TMA
//    engine's 1-D form, SASS UBLKCP): one instruction per line and direction, completion on an mbarrier on the
//    way in, bulk-group wait on the way out -- the per-16-byte staging loops of k_sweep_z cost more instructions
//    than the solve (measured: 347 M warp instructions per 512^3 sweep with cp.async pieces against 134 M of the
//    x / y sweeps).  Lines whose length is no multiple of 16 fall back to cp.async pieces / scalar copies.
//    Shared-memory pitch RL+2 doubles: the per-thread 16-byte chunk reads of a quarter warp fall into eight
//    different bank groups;
//  * a dense coefficient field is read at exposed cells only (the field must have been verified surface-only,
//    k_check_sparse; otherwise the launcher keeps k_sweep_z): the two end cells of every line with the tile,
//    cells next to an interior void once the codes have arrived;
//  * the general path keeps one factor per cell (1/den) in the thread's own slots of the staged tile;
//  * reduced system: shared-memory PCR across the P chunks of a line (adi_cart.cuh solve_reduced*), which
//    also carries the z-slab modes (ZMODE as in k_sweep_z).
//
// Reference semantics: adi3d_numba_coeff.py:205-237 (sweep_axis2).
#pragma once
#include "adi_cart.cuh"

namespace adi {

// ---- bulk asynchronous copies (the TMA engine's 1-D form: SASS UBLKCP) and the mbarrier they complete on ----
__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// global -> shared, `bytes` a multiple of 16, both addresses 16-byte aligned; completes on `bar`
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src, unsigned bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// shared -> global
__device__ __forceinline__ void bulk_s2g(void *dst, const void *src_smem, unsigned bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src_smem)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit_wait_read()
{
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// The slots of VOID cells are never written: they keep the bits the cell has in global memory, so that a whole
// line can be copied out again (void rows are identity rows: 1/den = 1).
template <int M>
struct ZtOps {
    double *slot;            // this chunk's M slots of the staged tile: coefficient, then 1/den (active cells)
    const double *qp, *dvp;  // global, contiguous along the line; may be null
    const Chunk<M> *ch;
    int nv;
    __device__ __forceinline__ double coef(int e) const { return slot[e]; }   // make_row ignores it for void cells
    __device__ __forceinline__ double q(int e) const { return (qp && e < nv) ? qp[e] : 0.0; }
    __device__ __forceinline__ double dirv(int e) const { return (dvp && e < nv) ? dvp[e] : 0.0; }
    __device__ __forceinline__ void put1(int e, double v) { if (ch->active(e)) slot[e] = v; }
    __device__ __forceinline__ double rinv(int e) const { return ch->active(e) ? slot[e] : 1.0; }
    // two-factor interface (unused: NS = 1)
    __device__ __forceinline__ void put2(int, double, double) {}
    __device__ __forceinline__ double la(int) const { return 0.0; }
    __device__ __forceinline__ double u(int) const { return 0.0; }
};

// smem: sT[KT][RL+2] doubles | xch[(ZMODE 1 ? 10 : 6) * NTH] | sEnd[KT][2] | mbarrier (16 B) | sCode[KT][RL+16] bytes
// ZF: the sweep solves only the first nz cells of lines whose operand arrays are a.zfull apart (launch_sweep_zt); its
// own instantiation, so that the common kernel keeps one stride
template <int M, int CMODE, bool EXTRA, int MAXT, int MINB, int ZMODE, bool ZF = false>
__global__ void __launch_bounds__(MAXT, MINB) k_sweep_zt(const SweepArgs a, const int vec)
{
    static_assert(M % 16 == 0, "codes are read sixteen at a time");
    constexpr int NS = 1;
    extern __shared__ double smem[];
    const int KT = blockDim.x, P = blockDim.y;
    const int kk = threadIdx.x, p = threadIdx.y;
    const int NTH = KT * P, tid = p * KT + kk;
    const int nz = a.nz;
    const size_t nlines = (size_t)a.nx * a.ny;
    const size_t L0 = (size_t)(a.tiles ? a.tiles[blockIdx.x] : (int)blockIdx.x) * KT;   // active-tile list, if any
    const int RL = P * M;                 // padded line length
    const int pitch = RL + 2;             // doubles
    const int cpitch = RL + 16;           // bytes
    double *sT = smem;
    double *xch = sT + (size_t)KT * pitch;
    double *sEnd = xch + (size_t)(ZMODE == 1 ? 10 : 6) * NTH;
    uint64_t *bar = reinterpret_cast<uint64_t *>(sEnd + 2 * KT);       // one mbarrier (+ 8 bytes of padding)
    uint8_t *sCode = reinterpret_cast<uint8_t *>(sEnd + 2 * KT + 2);
    const bool sparse = CMODE == 2;       // this kernel reads a dense coefficient field at exposed cells only
    const int nval = (int)min((size_t)KT, nlines - L0);                // lines of this tile that exist
    // line stride of the field (pitched buffers of the cylindrical path) and of the code array (0: one code line
    // shared by every z line -- the cylindrical z sweep, whose rows do not depend on the line)
    const size_t zs = a.zpitch ? (size_t)a.zpitch : (size_t)nz;
    // line stride of the operand arrays (trimmed sweeps: the full nz).  Spelled out at every use: one hoisted 64-bit
    // value cost the common kernel 52 bytes of spill stores (ptxas) and 5 % at 512^3.
#define ADI_ZT_OPSTRIDE (ZF ? (size_t)a.zfull : (size_t)nz)
    const size_t cs = a.code_line ? 0 : ADI_ZT_OPSTRIDE;
    // vec 2: whole lines travel as bulk asynchronous copies (one instruction per line and direction);
    // vec 1: 16-byte cp.async pieces / vector stores; vec 0: scalar loads and stores
    const bool bulk = vec == 2;

    // ---- stage in ----
    if (bulk) {
        if (tid == 0) {
            mbar_init(bar, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        // cells beyond the line's end and lines beyond the grid: code 0, T 0 (generic stores, disjoint from the copies)
        for (int i = tid; i < KT * (RL - nz); i += NTH) {
            const int l = i / (RL - nz), z = nz + (i - l * (RL - nz));
            sT[(size_t)l * pitch + z] = 0.0;
            sCode[(size_t)l * cpitch + z] = 0;
        }
        for (int i = tid; i < (KT - nval) * nz; i += NTH) {
            const int l = nval + i / nz, z = i % nz;
            sT[(size_t)l * pitch + z] = 0.0;
            sCode[(size_t)l * cpitch + z] = 0;
        }
        __syncthreads();
        if (tid < 32) {
            if (tid == 0) mbar_expect_tx(bar, (unsigned)nval * (unsigned)nz * 9u);
            __syncwarp();
            for (int l = tid; l < nval; l += 32) {
                bulk_g2s(sT + (size_t)l * pitch, a.in + (L0 + l) * zs, (unsigned)nz * 8u, bar);
                bulk_g2s(sCode + (size_t)l * cpitch, a.code + (L0 + l) * cs, (unsigned)nz, bar);
            }
        }
    } else if (vec) {
        const int ppl = RL / 2;
        for (int l = 0; l < KT; ++l) {
            const size_t line = L0 + l;
            const bool lok = line < nlines;
            const double *src = a.in + (lok ? line * zs : 0);
            for (int pr = tid; pr < ppl; pr += NTH) {
                const int z = 2 * pr;
                const bool ok = lok && z < nz;
                cp_async16(sT + (size_t)l * pitch + z, src + (ok ? z : 0), ok ? 16 : 0);
            }
        }
    } else {
        for (int l = 0; l < KT; ++l) {
            const size_t line = L0 + l;
            const bool lok = line < nlines;
            for (int z = tid; z < RL; z += NTH)
                sT[(size_t)l * pitch + z] = (lok && z < nz) ? a.in[line * zs + z] : 0.0;
        }
    }
    if (bulk) {
    } else if (vec && (nz & 15) == 0) {
        const int cpl = RL / 16;
        for (int i = tid; i < KT * cpl; i += NTH) {
            const int l = i / cpl, c16 = i - l * cpl;
            const size_t line = L0 + l;
            const bool ok = line < nlines && 16 * c16 < nz;
            cp_async16(sCode + (size_t)l * cpitch + 16 * c16, a.code + (ok ? line * cs + 16 * c16 : 0), ok ? 16 : 0);
        }
    } else {
        for (int l = 0; l < KT; ++l) {
            const size_t line = L0 + l;
            const bool lok = line < nlines;
            for (int z = tid; z < RL; z += NTH)
                sCode[(size_t)l * cpitch + z] = (lok && z < nz) ? a.code[line * cs + z] : (uint8_t)0;
        }
    }
    if (sparse) {
        // the two ends of every line (always exposed when active) are requested with the tile
        for (int i = tid; i < 2 * KT; i += NTH) {
            const size_t line = min(L0 + (size_t)(i >> 1), nlines - 1);
            cp_async8(smem_u32(sEnd + i), a.coeff + line * ADI_ZT_OPSTRIDE + ((i & 1) ? nz - 1 : 0));
        }
    }
    cp_async_wait_all();
    if (bulk) mbar_wait(bar, 0);
    __syncthreads();

    // ---- chunk to registers ----
    Chunk<M> ch;
    double *slot = sT + (size_t)kk * pitch + p * M;
    {
        const uint4 *cb = reinterpret_cast<const uint4 *>(sCode + (size_t)kk * cpitch + p * M);
#pragma unroll
        for (int w = 0; w < M / 16; ++w) {
            const uint4 v = cb[w];
            ch.cw[4 * w] = v.x; ch.cw[4 * w + 1] = v.y; ch.cw[4 * w + 2] = v.z; ch.cw[4 * w + 3] = v.w;
        }
#pragma unroll
        for (int j = 0; j < M / 2; ++j) {
            const double2 v = *reinterpret_cast<const double2 *>(slot + 2 * j);
            ch.T[2 * j] = v.x; ch.T[2 * j + 1] = v.y;
        }
    }
    if ((ZMODE == 0 || ZMODE == 2) && a.in == a.out) {  // (ZMODE 4 must still emit its relation)
        // tiles without an active cell (the void around a part that is still being built): nothing to solve,
        // nothing to write (the sweep is in place)
        bool any = false;
#pragma unroll
        for (int w = 0; w < M / 4; ++w) any = any || (ch.cw[w] & 0x01010101u) != 0u;
        if (!__syncthreads_or(any)) return;
    }
    const size_t line = L0 + kk;
    const size_t gline = min(line, nlines - 1) * ADI_ZT_OPSTRIDE;
    const int nv = line < nlines ? min(max(nz - p * M, 0), M) : 0;

    // coefficient of ACTIVE cell e of this chunk (exposed cells only carry one)
    auto exposed_coef = [&](int e, unsigned c) -> double {
        if (CMODE != 2) return 0.0;
        if ((c & (CB_ZM | CB_ZP)) == (CB_ZM | CB_ZP)) return 0.0;
        const int cell = p * M + e;
        if (cell == 0) return sEnd[2 * kk];
        if (cell == nz - 1) return sEnd[2 * kk + 1];
        return ldg_f64(a.coeff + gline + cell);
    };
    // coefficient of cell e into its slot (CMODE 2; active cells only: void slots keep the cell's bits).
    // Exposed interior cells are fetched asynchronously (cp.async, 8 bytes) -- the caller waits before reading.
    auto stage_coef = [&](int e) {
        const unsigned c = ch.code(e);   // 0 beyond the line's end
        if (!(c & CB_SELF)) return;
        const int cell = p * M + e;
        if ((c & (CB_ZM | CB_ZP)) == (CB_ZM | CB_ZP)) slot[e] = 0.0;
        else if (cell == 0) slot[e] = sEnd[2 * kk];
        else if (cell == nz - 1) slot[e] = sEnd[2 * kk + 1];
        else cp_async8(smem_u32(slot + e), a.coeff + gline + cell);
    };

    // 0: general rows; 1: cells 0..M-2 uniform; 2: cell 0 general, cells 1..M-2 uniform; 3: the first R4 cells
    // uniform, general rows behind them (adi_core.h)
    int path = 0, R4 = 0;
    if (a.uni) {
        if (__all_sync(0xffffffffu, chunk_uniform<M, 1>(ch, CB_ZM, CB_ZP))) {
            path = __all_sync(0xffffffffu, chunk_uniform<M, 0>(ch, CB_ZM, CB_ZP)) ? 1 : 2;
        } else if (a.uni > 1) {
            R4 = __reduce_min_sync(0xffffffffu, chunk_uniform_lead4<M>(ch, CB_ZM, CB_ZP));
            if (R4 >= 8) path = 3;
        }
    }
    ZtOps<M> ops;
    ops.slot = slot; ops.qp = nullptr; ops.dvp = nullptr; ops.nv = nv; ops.ch = &ch;
    First f;
    UniHead hd;
    hd.al = hd.bl = hd.br = 0.0;
    if (path == 1 || path == 2) {
        const unsigned cs = ch.code(M - 1), c0 = ch.code(0);
        const Row sep = make_row<CMODE, EXTRA>(cs, CB_ZM, CB_ZP, ch.T[M - 1], exposed_coef(M - 1, cs), 0.0, 0.0, a.k);
        if (path == 1) {
            f = chunk_forward_uniform<M, 0>(ch, a.uc, sep, sep, hd);
        } else {
            const Row head = make_row<CMODE, EXTRA>(c0, CB_ZM, CB_ZP, ch.T[0], exposed_coef(0, c0), 0.0, 0.0, a.k);
            f = chunk_forward_uniform<M, 1>(ch, a.uc, sep, head, hd);
        }
    } else if (path == 3) {
        // the coefficients of the tail are requested first and arrive while the run is eliminated
        if (CMODE == 2) {
#pragma unroll
            for (int g4 = 2; g4 < M / 4; ++g4) {
                if (4 * g4 >= R4) {
#pragma unroll
                    for (int e = 4 * g4; e < 4 * g4 + 4; ++e) stage_coef(e);
                }
            }
        }
#pragma unroll
        for (int g4 = 2; g4 < M / 4; ++g4) {
            if (4 * g4 >= R4) {
#pragma unroll
                for (int e = 4 * g4; e < 4 * g4 + 4; ++e) ch.T[e] = ch.active(e) ? ch.T[e] : 0.0;  // load rule
            }
        }
        ops.qp = (EXTRA && a.q) ? a.q + gline + min(p * M, nz - 1) : nullptr;
        ops.dvp = (EXTRA && a.dirv) ? a.dirv + gline + min(p * M, nz - 1) : nullptr;
        if (CMODE == 2) cp_async_wait_all();
        f = chunk_forward_hybrid<M, CMODE, EXTRA>(ch, a.uc, ops, R4, CB_ZM, CB_ZP, a.k);
    } else {
        if (CMODE == 2) {
#pragma unroll
            for (int e = 0; e < M; ++e) stage_coef(e);
        }
#pragma unroll
        for (int e = 0; e < M; ++e) ch.T[e] = ch.active(e) ? ch.T[e] : 0.0;  // load rule (adi_core.h)
        ops.qp = (EXTRA && a.q) ? a.q + gline + min(p * M, nz - 1) : nullptr;
        ops.dvp = (EXTRA && a.dirv) ? a.dirv + gline + min(p * M, nz - 1) : nullptr;
        if (CMODE == 2) cp_async_wait_all();
        f = chunk_forward<M, CMODE, EXTRA, NS, false>(ch, ops, CB_ZM, CB_ZP, a.k);
    }
    if (ZMODE == 1) {
        const Red3 r = solve_reduced3<M>(ch, f, xch, NTH, tid, KT, p, P);
        if (line < nlines) {
            // x_first = f.Y + f.V*L + f.W*S_0,  x_last = S_{P-1}   (adi_core.h, Iface)
            if (p == 0) {
                a.iface_dyn[line] = fma(f.W, r.D, f.Y);
                a.iface_stat[line] = fma(f.W, r.DL, f.V);
                a.iface_stat[nlines + line] = f.W * r.DR;
            }
            if (p == P - 1) {
                a.iface_dyn[nlines + line] = r.D;
                a.iface_stat[2 * nlines + line] = r.DL;
                a.iface_stat[3 * nlines + line] = r.DR;
            }
        }
        return;
    }
    if (ZMODE == 3) {
        double Sl0;
        const double S0 = solve_reduced<M, true>(ch, f, xch, NTH, tid, KT, p, P, &Sl0, 0.0, 0.0);
        if (line < nlines) {
            if (p == 0) a.iface_dyn[line] = fma(f.W, S0, f.Y);
            if (p == P - 1) a.iface_dyn[nlines + line] = S0;
        }
        return;
    }
    double Sl, Lg = 0.0, Rg = 0.0;
    if (ZMODE == 2) {
        const size_t lc = min(line, nlines - 1);
        if (p == 0) Lg = a.ghost[lc];
        if (p == P - 1) Rg = a.ghost[nlines + lc];
    }
    constexpr bool GHOSTS = (ZMODE == 2 || ZMODE == 4);  // 4: both ghosts at zero
    const double S = solve_reduced<M, GHOSTS>(ch, f, xch, NTH, tid, KT, p, P, &Sl, Lg, Rg);
    if (path == 1) chunk_backward_uniform<M, 0>(ch, a.uc, hd, Sl, S);
    else if (path == 2) chunk_backward_uniform<M, 1>(ch, a.uc, hd, Sl, S);
    else if (path == 3) chunk_backward_hybrid<M, EXTRA>(ch, a.uc, ops, R4, CB_ZM, CB_ZP, a.k.g, Sl, S);
    else chunk_backward<M, EXTRA, NS>(ch, ops, CB_ZM, CB_ZP, a.k.g, Sl, S);
    if (ZMODE == 4) {
        if (line < nlines) {  // void cells hold 0 (load rule), as in the relation of ZMODE 3
            if (p == 0) a.iface_dyn[line] = ch.T[0];
            if (p == P - 1) a.iface_dyn[nlines + line] = ch.T[M - 1];
        }
    }

    // ---- results back through the thread's own slots, then coalesced copy-out ----
    // The sweep runs in place, so void cells must keep the bits they have in global memory.  Bulk mode stores whole
    // lines: uniform chunks hold no void cell, and the general paths never write the slot of a void cell.
    if (path == 1 || path == 2) {
#pragma unroll
        for (int j = 0; j < M / 2; ++j)
            *reinterpret_cast<double2 *>(slot + 2 * j) = make_double2(ch.T[2 * j], ch.T[2 * j + 1]);
    } else {
        // the slots of void cells still hold the cells' own bits
#pragma unroll
        for (int e = 0; e < M; ++e)
            if (ch.active(e)) slot[e] = ch.T[e];
    }
    if (bulk) {
        fence_proxy_async();       // generic-proxy writes above -> visible to the bulk copy engine
        __syncthreads();
        if (tid < 32) {
            for (int l = tid; l < nval; l += 32)
                bulk_s2g(a.out + (L0 + l) * zs, sT + (size_t)l * pitch, (unsigned)nz * 8u);
            bulk_commit_wait_read();   // shared memory must outlive the reads of the copy engine
        }
        return;
    }
    __syncthreads();
    const bool inplace = (a.in == a.out);
    if (vec) {
        const int ppl = RL / 2;
        for (int l = 0; l < KT; ++l) {
            const size_t ln = L0 + l;
            if (ln >= nlines) break;
            for (int pr = tid; pr < ppl; pr += NTH) {
                const int z = 2 * pr;
                if (z >= nz) break;
                const double2 v = *reinterpret_cast<const double2 *>(sT + (size_t)l * pitch + z);
                const unsigned cc = *reinterpret_cast<const unsigned short *>(sCode + (size_t)l * cpitch + z);
                const size_t g = ln * zs + z;
                const bool a0 = cc & 1u, a1 = (cc >> 8) & 1u;
                if (a0 && a1) {
                    *reinterpret_cast<double2 *>(a.out + g) = v;
                } else if (inplace) {
                    if (a0) a.out[g] = v.x;
                    if (a1) a.out[g + 1] = v.y;
                } else {
                    a.out[g] = a0 ? v.x : a.in[g];
                    a.out[g + 1] = a1 ? v.y : a.in[g + 1];
                }
            }
        }
    } else {
        for (int l = 0; l < KT; ++l) {
            const size_t ln = L0 + l;
            if (ln >= nlines) break;
            for (int z = tid; z < nz; z += NTH) {
                const size_t g = ln * zs + z;
                const bool act = sCode[(size_t)l * cpitch + z] & 1u;
                if (act) a.out[g] = sT[(size_t)l * pitch + z];
                else if (!inplace) a.out[g] = a.in[g];
            }
        }
    }
}
#undef ADI_ZT_OPSTRIDE

}  // namespace adi
