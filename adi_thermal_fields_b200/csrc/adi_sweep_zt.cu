// adi_sweep_zt.cu -- launcher of the second-generation z sweep (adi_sweep_zt.cuh).
#include <stdint.h>

#include "adi_launch.h"
#include "adi_sweep_zt.cuh"

namespace adi {

template <int ZMODE>
static int launch_zt_mode(adi_ctx *ctx, const SweepArgs &a, bool dense, bool extra, int M, int P, int KT, size_t smem,
                          cudaStream_t st)
{
    SweepArgs b = a;
    b.uni = (ctx->opt_uni && !extra) ? (ctx->opt_hyb ? 2 : 1) : 0;   // 2: chunks with a uniform lead take the hybrid path
    uni_const_build(b.uc, a.k.g);
    const size_t nlines = (size_t)a.nx * a.ny;
    dim3 block(KT, P), grid((unsigned)((nlines + KT - 1) / KT));
    if (!a.line_batch && (ZMODE == 0 || ZMODE == 2) && a.in == a.out) {
        // in place, nothing to emit for void lines: launch only the tiles that hold an active cell
        const int *list = nullptr;
        int nact = 0, tnx = 0;
        int rc = ensure_tiles(ctx, 2, KT, st, &list, &nact, &tnx);
        if (rc) return rc;
        if (list) {
            if (nact == 0) return ADI_OK;
            b.tiles = list; b.tiles_nx = tnx;
            grid = dim3((unsigned)nact);
        }
    }
    int vec = ((a.nz & 1) == 0 && (a.zpitch & 1) == 0 && (((uintptr_t)a.in | (uintptr_t)a.out | (uintptr_t)a.code) & 15) == 0) ? 1 : 0;
    if (vec && (a.nz & 15) == 0 && a.in == a.out && ctx->opt_bulk) vec = 2;   // whole lines as bulk asynchronous copies
    const bool big = KT * P > 256;
#define ADI_GOZ(M_, MAXT, MINB, ZF_)                                                                                   \
    {                                                                                                                  \
        if (dense) {                                                                                                   \
            if (extra) return launch(k_sweep_zt<M_, 2, true, MAXT, MINB, ZMODE, ZF_>, grid, block, smem, st, ctx, b, vec);  \
            return launch(k_sweep_zt<M_, 2, false, MAXT, MINB, ZMODE, ZF_>, grid, block, smem, st, ctx, b, vec);            \
        }                                                                                                              \
        if (extra) return launch(k_sweep_zt<M_, 1, true, MAXT, MINB, ZMODE, ZF_>, grid, block, smem, st, ctx, b, vec);      \
        return launch(k_sweep_zt<M_, 1, false, MAXT, MINB, ZMODE, ZF_>, grid, block, smem, st, ctx, b, vec);                \
    }
    if constexpr (ZMODE == 0) {
        if (a.zfull) {   // trimmed lines (always 32-cell chunks: nz >= 256)
            if (M != 32) { set_error("adi_cart_step: trimmed z sweep needs 32-cell chunks"); return ADI_EINVAL; }
            if (big) ADI_GOZ(32, 512, 1, true) else ADI_GOZ(32, 256, 2, true)
        }
    }
    if (M == 16) { if (big) ADI_GOZ(16, 512, 1, false) else ADI_GOZ(16, 256, 2, false) }
    else { if (big) ADI_GOZ(32, 512, 1, false) else ADI_GOZ(32, 256, 2, false) }
#undef ADI_GOZ
    return ADI_OK;
}

// *used = 0: this sweep is not one for k_sweep_zt (the caller falls back to k_sweep_z).
// Shapes: nz <= 128: 16-cell chunks; longer lines: 32-cell chunks (the z-slab modes need nz % chunk == 0 and take
// 16-cell chunks when nz is no multiple of 32); up to 32 chunks per line in 256-thread blocks (2 per SM), up to 64
// in 512-thread blocks; KT lines per tile = threads / chunks, halved until the tile fits.
int launch_sweep_zt(adi_ctx *ctx, const SweepArgs &a0, bool dense, bool extra, int zmode, cudaStream_t st, int *used)
{
    *used = 0;
    if (!ctx->opt_zt || a0.nz > 2048) return ADI_OK;
    // A part under construction (waam_from_stl_v7_mm.py:487-550: layers are born bottom-up along z): every cell above
    // the top of the part is void -- an identity row coupled to nothing -- so the single-GPU, in-place sweep solves only
    // the first ztop cells of each line (rounded up to whole 32-cell chunks, at least 256) and leaves the rest where the
    // explicit stage put them.  The top comes from the code build (k_build_code_v); operand arrays keep their stride.
    SweepArgs a = a0;
    if (zmode == 0 && ctx->opt_ztrim && !a.line_batch && !a.code_line && a.zpitch == 0 && a.zfull == 0 && a.in == a.out &&
        a.nz >= 512 && a.nz % 32 == 0 && a.nz == ctx->nz && a.code == ctx->code[2]) {
        int rc = part_extent(ctx, st);
        if (rc) return rc;
        if (ctx->ztop >= 0) {
            const int ne = std::max(256, (ctx->ztop + 31) / 32 * 32);
            if (ne < a.nz) {
                a.zfull = a.nz; a.zpitch = a.nz; a.nz = ne;
                ctx->ztrim_used++;
            }
        }
    }
    if (dense && !a.sparse) return ADI_OK;     // a dense coefficient field that must be read everywhere: k_sweep_z stages it
    // 32-cell chunks from 64 cells up: on 2048 x 2048 x 128 (the slab of configs[4] at N = 8) 4 chunks x 16 lines take
    // 1.31 ms = 6.9 TB/s against 2.12 ms for k_sweep_z and 2.6 ms for 16-cell chunks (r02x)
    int M = a.nz < 64 ? 16 : 32;
    if (ctx->opt_m == 16 || ctx->opt_zm == 16 || (zmode != 0 && a.nz % 32 != 0)) M = 16;
    // short lines in 16-cell chunks stay with k_sweep_z (one-warp tiles, shuffle PCR): 0.69 against 0.89 ms at nz = 128
    if (a.nz <= 128 && M == 16 && ctx->opt_zt != 2) return ADI_OK;
    if (zmode != 0 && a.nz % M != 0) return ADI_OK;   // (k_sweep_z reports the error)
    const int P = (a.nz + M - 1) / M;
    if (P > 64) return ADI_OK;
    const int maxt = P > 32 ? 512 : 256;
    // 128-thread tiles by default (four resident per SM overlap their load / solve / store phases better than two
    // 256-thread ones: 0.59 against 0.66 ms at 512^3 with one general warp per tile, r02g)
    int KT = 32;
    while (KT > 1 && KT * P > (P > 32 ? 512 : 128)) KT >>= 1;
    if (P <= 4) KT = std::min(KT, 16);     // short lines: 64-thread tiles (1.31 against 1.39 ms, r02x)
    if (ctx->opt_lt > 0) {
        int w = 1;
        while (2 * w <= ctx->opt_lt && 2 * w <= 32 && 2 * w * P <= maxt) w <<= 1;
        KT = w;
    }
    const size_t RL = (size_t)P * M;
    auto bytes = [&](int kt) {
        return ((size_t)kt * (RL + 2) + (size_t)(zmode == 1 ? 10 : 6) * kt * P + 2 * (size_t)kt + 2) * sizeof(double) + (size_t)kt * (RL + 16);
    };
    const size_t limit = maxt == 256 ? 113 * 1024 : 226 * 1024;
    while (KT > 1 && bytes(KT) > limit) KT >>= 1;
    if (bytes(KT) > 226 * 1024) return ADI_OK;
    *used = 1;
    const size_t smem = bytes(KT);
    if (zmode == 1) return launch_zt_mode<1>(ctx, a, dense, extra, M, P, KT, smem, st);
    if (zmode == 2) return launch_zt_mode<2>(ctx, a, dense, extra, M, P, KT, smem, st);
    if (zmode == 3) return launch_zt_mode<3>(ctx, a, dense, extra, M, P, KT, smem, st);
    if (zmode == 4) return launch_zt_mode<4>(ctx, a, dense, extra, M, P, KT, smem, st);
    return launch_zt_mode<0>(ctx, a, dense, extra, M, P, KT, smem, st);
}

}  // namespace adi
