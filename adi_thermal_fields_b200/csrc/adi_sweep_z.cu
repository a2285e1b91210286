// adi_sweep_z.cu -- z sweep (contiguous axis), in place (adi3d_numba_coeff.py:301).
#include <stdint.h>

#include "adi_launch.h"

namespace adi {

template <int ZMODE>
static int launch_sweep_z_mode(adi_ctx *ctx, const SweepArgs &a, bool dense, bool extra, cudaStream_t st)
{
    Shape s;
    int rc = pick_shape(ctx, a.nz, -1, &s);  // -1: z sweep (no lane-count option)
    if (rc) return rc;
    if (ZMODE != 0 && a.nz % s.M != 0) {
        set_error("adi_cart_zsweep_*: the local z extent must be a multiple of the chunk length (16; 32 for local nz > 1024)");
        return ADI_EINVAL;
    }
    // lines per block: fill the block, but keep the staged tiles small enough for two
    // resident blocks per SM where the line length allows it
    const size_t line_bytes = (size_t)s.P * s.M * (8 * (1 + s.NS) + 1);  // T + factor tiles + codes
    const size_t budget = s.var == VAR_32L ? 220 * 1024 : 28 * 1024;  // one 1024-cell line is 25.6 KB  // small blocks overlap best (measured)
    int LT = s.W;
    while (LT > 1 && LT * line_bytes > budget) LT >>= 1;
    if (s.var == VAR_16) {
        // measured on B200 (dense operands, 512 x 512 lines): one-warp tiles win for short lines -- many
        // resident blocks keep the loads, solves and stores of different tiles overlapped:
        // nz 32: 0.31 -> 0.22 ms (LT 8), 64: 0.17 -> 0.135 (8), 128: 0.22 -> 0.18 (4), 256: 0.35 -> 0.31 (2),
        // 384: 0.61 -> 0.57 (4)
        if (s.P <= 16) LT = std::min(8, std::max(1, 32 / s.P));
        else if (s.P <= 24) LT = 4;
    }
    if (ctx->opt_lt > 0) LT = (int)std::min<long>(ctx->opt_lt, s.W);
    const size_t smem = LT * line_bytes;
    if (smem > 227 * 1024) {
        set_error("adi_cart_step: z tile does not fit shared memory");
        return ADI_EINVAL;
    }
    const size_t nlines = (size_t)a.nx * a.ny;
    dim3 block(s.P, LT), grid((unsigned)((nlines + LT - 1) / LT));
    SweepArgs b = a;
    b.uni = (ctx->opt_uni && !extra && !dense && s.M <= UNI_MAX) ? 1 : 0;
    uni_const_build(b.uc, a.k.g);
    const int vec = ((a.nz & 1) == 0 && (((uintptr_t)a.in | (uintptr_t)a.out | (uintptr_t)a.coeff |
                                          (uintptr_t)a.code) & 15) == 0) ? 1 : 0;
#define ADI_GO(M, NS, MAXT, MINB)                                                                          \
    {                                                                                                  \
        if (dense) {                                                                                   \
            if (extra) return launch(k_sweep_z<M, NS, 2, true, MAXT, MINB, ZMODE>, grid, block, smem, st, ctx, b, vec);  \
            return launch(k_sweep_z<M, NS, 2, false, MAXT, MINB, ZMODE>, grid, block, smem, st, ctx, b, vec);            \
        }                                                                                              \
        if (extra) return launch(k_sweep_z<M, NS, 1, true, MAXT, MINB, ZMODE>, grid, block, smem, st, ctx, b, vec);      \
        return launch(k_sweep_z<M, NS, 1, false, MAXT, MINB, ZMODE>, grid, block, smem, st, ctx, b, vec);                \
    }
    ADI_FOR_VARIANT(s.var, ADI_GO)
#undef ADI_GO
    return ADI_OK;
}

int launch_sweep_zt(adi_ctx *ctx, const SweepArgs &a, bool dense, bool extra, int zmode, cudaStream_t st, int *used);

int launch_sweep_z(adi_ctx *ctx, const SweepArgs &a, bool dense, bool extra, int zmode, cudaStream_t st)
{
    int used = 0;
    const int rc = launch_sweep_zt(ctx, a, dense, extra, zmode, st, &used);   // second generation (adi_sweep_zt.cuh)
    if (rc || used) return rc;
    if (zmode == 1) return launch_sweep_z_mode<1>(ctx, a, dense, extra, st);
    if (zmode == 2) return launch_sweep_z_mode<2>(ctx, a, dense, extra, st);
    if (zmode == 3) return launch_sweep_z_mode<3>(ctx, a, dense, extra, st);
    if (zmode == 4) return launch_sweep_z_mode<4>(ctx, a, dense, extra, st);
    return launch_sweep_z_mode<0>(ctx, a, dense, extra, st);
}

}  // namespace adi
