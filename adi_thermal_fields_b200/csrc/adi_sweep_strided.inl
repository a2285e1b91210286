// adi_sweep_strided.inl -- launcher body of the strided sweeps; included by adi_sweep_x.cu
// (ADI_AXIS 0, with the fused explicit stage) and adi_sweep_y.cu (ADI_AXIS 1).
#include "adi_launch.h"

namespace adi {

static int launch_strided_axis(adi_ctx *ctx, const SweepArgs &a, bool dense, bool extra, bool expl,
                               cudaStream_t st)
{
    constexpr int AXIS = ADI_AXIS;
    const int n = AXIS == 0 ? a.nx : a.ny;
    const int other = AXIS == 0 ? a.ny : a.nx;
    Shape s;
    int rc = pick_shape(ctx, n, ctx->opt_kt, &s);
    if (rc) return rc;
    dim3 block(s.W, s.P), grid((a.nz + s.W - 1) / s.W, other);
    const size_t nth = (size_t)s.W * s.P;
    // columns: NS factor slots + exchange buffer; the staged explicit stage adds a third slot (which
    // the exchange buffer then reuses) and the z halo [2][M][P]
    const size_t smem = ((AXIS == 0 && expl && s.NS == 2) ? (size_t)3 * s.M * nth + (size_t)2 * s.M * s.P
                                                        : (size_t)(s.NS * s.M + 6) * nth) * sizeof(double);
    if ((unsigned long long)s.M * (AXIS == 0 ? (unsigned long long)a.ny * a.nz : (unsigned long long)a.nz) >= (1ull << 32)) {
        set_error("adi_cart_step: grid too large for 32-bit in-chunk offsets");
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
