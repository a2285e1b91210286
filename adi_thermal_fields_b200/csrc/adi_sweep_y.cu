// adi_sweep_y.cu -- y sweep (stride nz), in place (adi3d_numba_coeff.py:300).
#define ADI_AXIS 1
#include "adi_sweep_strided.inl"

namespace adi {
int launch_sweep_y(adi_ctx *ctx, const SweepArgs &a, bool dense, bool extra, cudaStream_t st)
{
    return launch_strided_axis(ctx, a, dense, extra, false, st);
}
}  // namespace adi
