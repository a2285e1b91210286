// adi_sweep_x.cu -- x sweep (stride ny*nz) with the explicit stage fused (adi3d_numba_coeff.py:298-299).
#define ADI_AXIS 0
#include "adi_sweep_strided.inl"

namespace adi {
int launch_sweep_x(adi_ctx *ctx, const SweepArgs &a, bool dense, bool extra, bool expl, cudaStream_t st)
{
    return launch_strided_axis(ctx, a, dense, extra, expl, st);
}
}  // namespace adi
