// adi_sweep_strided.inl -- launcher body of the strided sweeps; included by adi_sweep_x.cu
// (ADI_AXIS 0, with the fused explicit stage) and adi_sweep_y.cu (ADI_AXIS 1).
#include <stdint.h>

#include "adi_launch.h"
#include "adi_sweep_xyp.cuh"
#include "adi_sweep_xyu.cuh"

namespace adi {

static int launch_strided_axis(adi_ctx *ctx, const SweepArgs &a, bool dense, bool extra, bool expl,
                               cudaStream_t st)
{
    constexpr int AXIS = ADI_AXIS;
    const int n = AXIS == 0 ? a.nx : a.ny;
    const int other = AXIS == 0 ? a.ny : a.nx;
    if (ctx->opt_xy2 && ctx->npadT[AXIS] > 0 && n <= 2048 && !(AXIS == 0 && expl) && other <= 65535) {
        // second-generation sweeps (adi_sweep_xy.cuh).  Shapes (measured on B200, profiles/r02c_xy_probe.txt):
        //   n <= 128          M 16 (two factors per cell), 256 threads, 2 blocks / SM
        //   128 < n <= 1024   M 32 (one factor per cell), <= 32 chunks, 256 threads, 2 blocks / SM; up to 512 cells
        //                     that is 16 lanes per row (128-byte rows: x 0.40 / y 0.40 ms at 512^3 against
        //                     0.51 / 0.46 ms for M 16 with 64-byte rows at 3 blocks / SM)
        //   1024 < n <= 2048  M 32, <= 64 chunks, 512 threads, 1 block / SM
        // options: m=16 / m=32 force the chunk length (n <= 512), occ=3|4 more resident blocks (M 16), wide=1
        // 512-thread blocks (twice the lanes per row), kt=N lanes per row
        int M = n <= 128 ? 16 : 32;
        if (n <= 512 && (ctx->opt_m == 16 || ctx->opt_m == 32)) M = (int)ctx->opt_m;
        const int P = (n + M - 1) / M;
        const bool wide = P <= 32 && ctx->opt_wide;
        const bool half = P > 32 && ctx->opt_lb == 256;   // long lines in 256-thread blocks (4 lanes per row), 2 blocks / SM
        const int PR = P > 32 ? 2 : 1, maxt = ((P > 32 && !half) || wide) ? 512 : 256;
        int KT = 32;
        while (KT > 1 && KT * P > maxt) KT >>= 1;
        if (ctx->opt_kt > 0) {
            int w = 1;
            while (2 * w <= ctx->opt_kt && 2 * w <= 32 && 2 * w * P <= maxt) w <<= 1;
            KT = w;
        }
        const int NTH = KT * P;
        SweepArgs b = a;
        b.codeT = ctx->codeT[AXIS] + a.line0;
        b.npad = ctx->npadT[AXIS];
        b.tw = (ctx->opt_tw && NTH >= 32) ? 1 : 0;
        b.remap = ctx->opt_remap ? 1 : 0;
        b.dbg = (int)ctx->opt_dbg;
        b.uni = (ctx->opt_uni && !extra && (!dense || a.sparse)) ? 1 : 0;
        uni_const_build(b.uc, a.k.g);
        // blocks per SM (M 16 only): 2 = 128 registers, two factors per cell in shared memory; 3 / 4 = 80 / 64
        // registers with one factor per cell (the general path recomputes the couplings from the code)
        const int occ = (M == 16 && (ctx->opt_occ == 3 || ctx->opt_occ == 4)) ? (int)ctx->opt_occ : 2;
        const int NS = (M == 16 && occ == 2 && !wide) ? 2 : 1;
        const bool big = NTH > 256;   // 512-thread launch bounds
        const size_t xch = std::max<size_t>((size_t)7 * P * (KT + 1), (size_t)6 * NTH);
        const size_t smem = ((size_t)NS * M * NTH + xch) * sizeof(double);
        if ((unsigned long long)M * 8ull * (AXIS == 0 ? (unsigned long long)a.ny * a.nz : (unsigned long long)a.nz) >= (1ull << 32)) {
            set_error("adi_cart_step: grid too large for 32-bit in-chunk byte offsets (chunk length x line stride x 8 >= 4 GiB)");
            return ADI_EINVAL;
        }
        dim3 block(KT, P), grid((a.nz + KT - 1) / KT, other);
        // Long lines whose sweep may take the uniform paths: the ALL-UNIFORM tiles (k_tile_flags) go to k_sweep_xyu and
        // only the other active tiles to k_sweep_xy.  16 lanes per tile (1024 threads, one block per SM, 128-byte rows)
        // where the rows of a tile lie >= 4 MB apart (x sweep of 2048 x 2048 x 1024: 17.3 against 25.2 ms; x 256: 4.09
        // against 5.11 ms), otherwise 8 lanes (512 threads, two blocks per SM: y sweep 16.0 against 19.2 ms; r02y / r02z).
        // Lines of 513..1024 cells: twice the lanes of k_sweep_xy at half the registers -- 512 threads, two blocks per SM,
        // 128-byte rows (1024 x 1024 x 256: x 1.09 -> 0.88, y 0.92 -> 0.84 ms).  Not for 512-cell lines, where
        // k_sweep_xy already has 128-byte rows (0.37 -> 0.41 ms; option xyu=2 forces it).
        const bool long_u = P > 32 && !half && KT == 8;
        const bool mid_u = P >= (ctx->opt_xyu >= 2 ? 16 : 17) && P <= 32 && M == 32 && !wide && !ctx->opt_kt && !ctx->opt_m && KT * P == 256;
        const bool want_u = (long_u || mid_u) && b.uni && ctx->opt_xyu && !ctx->opt_xyp && a.in == a.out && n % 32 == 0;
        bool listed = false;
        if (want_u) {
            const unsigned long long row_stride = (AXIS == 0 ? (unsigned long long)a.ny * a.nz : (unsigned long long)a.nz) * 8ull;
            const int KTU = mid_u ? 2 * KT
                                  : (ctx->opt_ukt == 16 ? 16 : (ctx->opt_ukt == 8 ? 8 : (row_stride >= (4ull << 20) ? 16 : 8)));
            const int *lu = nullptr, *lg = nullptr;
            int nu = 0, ng = 0, tnx2 = 0;
            int rc = ensure_tiles_split(ctx, AXIS, KTU, st, &lu, &nu, &lg, &ng, &tnx2);
            if (rc) return rc;
            if (lu && lg) {
                listed = true;
                if (nu > 0) {
                    SweepArgs u = b;
                    u.tiles = lu; u.tiles_nx = tnx2;
                    const size_t smu = ((size_t)16 * KTU * P + (size_t)6 * KTU * P) * sizeof(double);
                    const dim3 ublock(KTU, P);
                    if (KTU * P > 512) {
                        if (dense) rc = launch(k_sweep_xyu<AXIS, 2, 1024, 1>, dim3((unsigned)nu), ublock, smu, st, ctx, u);
                        else rc = launch(k_sweep_xyu<AXIS, 1, 1024, 1>, dim3((unsigned)nu), ublock, smu, st, ctx, u);
                    } else {
                        if (dense) rc = launch(k_sweep_xyu<AXIS, 2, 512, 2>, dim3((unsigned)nu), ublock, smu, st, ctx, u);
                        else rc = launch(k_sweep_xyu<AXIS, 1, 512, 2>, dim3((unsigned)nu), ublock, smu, st, ctx, u);
                    }
                    if (rc) return rc;
                    ctx->xyu_used++;
                }
                if (ng == 0) return ADI_OK;
                // the remaining tiles: k_sweep_xy, 8 lanes per block -- a 16-lane tile of the list is two blocks
                b.tiles = lg; b.tiles_nx = tnx2;
                b.tsplit = KTU == 2 * KT ? 1 : 0;
                grid = dim3((unsigned)ng << b.tsplit, 1);
            }
        }
        if (a.in == a.out && !listed) {
            // in place, nothing to do for void tiles: launch only the tiles that hold an active cell
            const int *list = nullptr;
            int nact = 0, tnx = 0;
            int rc = ensure_tiles(ctx, AXIS, KT, st, &list, &nact, &tnx);
            if (rc) return rc;
            if (list) {
                if (nact == 0) return ADI_OK;
                b.tiles = list; b.tiles_nx = tnx;
                grid = dim3((unsigned)nact, 1);
            }
        }
        if (P > 32 && !half && b.uni && ctx->opt_xyp && !ctx->opt_kt && KT == 8 && n % 128 == 0 && a.nz % 2 == 0 &&
            ((uintptr_t)a.in & 15) == 0 && ctx->xyp_state >= 0) {
            // long lines, uniform paths allowed: persistent blocks, tiles prefetched by the TMA engine (adi_sweep_xyp.cuh)
            if (ctx->sm_count <= 0) {
                int sms = 0;
                ADI_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device));
                ctx->sm_count = sms > 0 ? sms : 148;
            }
            const long long nt = b.tiles ? (long long)grid.x : (long long)grid.x * grid.y;
            CUtensorMap tm;
            int lay = ctx->xyp_state;
            if (nt <= 0x7fffffffll && xyp_tensor_map(&tm, &lay, AXIS, a.in, a.nx, a.ny, a.nz, (int)ctx->opt_promo) == ADI_OK) {
                ctx->xyp_state = lay;
                ctx->xyp_used++;
                const int NW = NTH / 32;
                const size_t smp = (size_t)NW * 8192 + (size_t)6 * NTH * sizeof(double) + (size_t)2 * NTH * 16 +
                                   (size_t)2 * NTH * sizeof(double) + (size_t)NW * 8;
                const dim3 pgrid((unsigned)std::min<long long>(nt, ctx->sm_count));
                if (dense) return launch(k_sweep_xyp<AXIS, 2>, pgrid, block, smp, st, ctx, b, tm, (int)nt, lay, (int)ctx->opt_seq);
                return launch(k_sweep_xyp<AXIS, 1>, pgrid, block, smp, st, ctx, b, tm, (int)nt, lay, (int)ctx->opt_seq);
            }
            if (nt <= 0x7fffffffll) ctx->xyp_state = -1;   // the driver refused the tensor map: k_sweep_xy from now on
        }
#define ADI_GO2(M_, NS_, PR_, MAXT, MINB)                                                                            \
        {                                                                                                        \
            if (dense) {                                                                                         \
                if (extra) return launch(k_sweep_xy<AXIS, M_, NS_, 2, true, PR_, MAXT, MINB>, grid, block, smem, st, ctx, b);  \
                return launch(k_sweep_xy<AXIS, M_, NS_, 2, false, PR_, MAXT, MINB>, grid, block, smem, st, ctx, b);            \
            }                                                                                                    \
            if (extra) return launch(k_sweep_xy<AXIS, M_, NS_, 1, true, PR_, MAXT, MINB>, grid, block, smem, st, ctx, b);      \
            return launch(k_sweep_xy<AXIS, M_, NS_, 1, false, PR_, MAXT, MINB>, grid, block, smem, st, ctx, b);                \
        }
        if (M == 16 && wide) ADI_GO2(16, 1, 1, 512, 2)
        else if (M == 16 && occ == 2) ADI_GO2(16, 2, 1, 256, 2)
        else if (M == 16 && occ == 3) ADI_GO2(16, 1, 1, 256, 3)
        else if (M == 16) ADI_GO2(16, 1, 1, 256, 4)
        else if (PR == 1 && !big) ADI_GO2(32, 1, 1, 256, 2)
        else if (PR == 1) ADI_GO2(32, 1, 1, 512, 1)
        else if (half) ADI_GO2(32, 1, 2, 256, 2)
        else ADI_GO2(32, 1, 2, 512, 1)
#undef ADI_GO2
    }
    if (n > 1024 && n <= 4096 && !(AXIS == 0 && expl) && ctx->opt_m != 32) {
        // long lines: a cluster of 2 or 4 CTAs shares each line (K1c)
        // 2 or 4 CTAs of <= 64 chunks x 8 lanes (512 threads, one CTA per SM); measured against 4 / 8 CTAs
        // of 256 threads (two per SM): 2.58 vs 2.84 ms on 2048 x 2048 x 64
        const int CS = n <= 2048 ? 2 : 4, M = 16;
        const int Ptot = (n + M - 1) / M, P = (Ptot + CS - 1) / CS;   // chunks per CTA (<= 64)
        int KT = 8;
        while (KT > 1 && KT * P > 512) KT >>= 1;
        dim3 block(KT, P), grid((a.nz + KT - 1) / KT, other, CS);
        const size_t smem = ((size_t)(2 * M + 10) * KT * P + 8 * KT) * sizeof(double);
        if ((unsigned long long)M * 8ull * (AXIS == 0 ? (unsigned long long)a.ny * a.nz : (unsigned long long)a.nz) >= (1ull << 32)) {
            set_error("adi_cart_step: grid too large for 32-bit in-chunk byte offsets (chunk length x line stride x 8 >= 4 GiB)");
            return ADI_EINVAL;
        }
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 1; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = CS;
        cfg.attrs = attr; cfg.numAttrs = 1;
#define ADI_CGO(CM, EX)                                                                                       \
        {                                                                                                     \
            auto kern = k_sweep_strided_cl<AXIS, CM, EX, 512, 1>;                                             \
            ADI_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));     \
            ADI_CUDA(cudaLaunchKernelEx(&cfg, kern, a));                                                      \
        }
        if (dense) { if (extra) ADI_CGO(2, true) else ADI_CGO(2, false) }
        else { if (extra) ADI_CGO(1, true) else ADI_CGO(1, false) }
#undef ADI_CGO
        ctx->launches++;
        ADI_CUDA(cudaGetLastError());
        return ADI_OK;
    }
    Shape s;
    int rc = pick_shape(ctx, n, ctx->opt_kt, &s);
    if (rc) return rc;
    dim3 block(s.W, s.P), grid((a.nz + s.W - 1) / s.W, other);
    const size_t nth = (size_t)s.W * s.P;
    // columns: NS factor slots + exchange buffer; the staged explicit stage adds a third slot (which
    // the exchange buffer then reuses) and the z halo [2][M][P]
    const size_t smem = ((AXIS == 0 && expl && s.NS == 2) ? (size_t)3 * s.M * nth + (size_t)2 * s.M * s.P
                                                        : (size_t)(s.NS * s.M + 6) * nth) * sizeof(double);
    if ((unsigned long long)s.M * 8ull * (AXIS == 0 ? (unsigned long long)a.ny * a.nz : (unsigned long long)a.nz) >= (1ull << 32)) {
        set_error("adi_cart_step: grid too large for 32-bit in-chunk byte offsets (chunk length x line stride x 8 >= 4 GiB)");
        return ADI_EINVAL;
    }
#define ADI_GO(M, NS, MAXT, MINB)                                                                           \
    {                                                                                                   \
        if (dense) {                                                                                    \
            if (extra) return launch(k_sweep_strided<AXIS, M, NS, 2, true, XP, MAXT, MINB>, grid, block, smem, st, ctx, a); \
            return launch(k_sweep_strided<AXIS, M, NS, 2, false, XP, MAXT, MINB>, grid, block, smem, st, ctx, a);           \
        }                                                                                               \
        if (extra) return launch(k_sweep_strided<AXIS, M, NS, 1, true, XP, MAXT, MINB>, grid, block, smem, st, ctx, a);     \
        return launch(k_sweep_strided<AXIS, M, NS, 1, false, XP, MAXT, MINB>, grid, block, smem, st, ctx, a);               \
    }
    if (AXIS == 0 && expl) {
        constexpr bool XP = (AXIS == 0);
        ADI_FOR_VARIANT(s.var, ADI_GO)
    } else {
        constexpr bool XP = false;
        ADI_FOR_VARIANT(s.var, ADI_GO)
    }
#undef ADI_GO
    return ADI_OK;
}

}  // namespace adi
