// adi_launch.h -- variant selection and launch helpers shared by the sweep translation units.
#pragma once
#include <algorithm>

#include "adi_cart.cuh"
#include "adi_ctx.h"

namespace adi {

// Kernel variants (chunk length M, block-size bound, resident blocks per SM):
//   V16  M=16, two factors per cell in smem, <=256 threads, 2 blocks/SM   lines up to 512 cells
//   V16L M=16, two factors per cell in smem, <=512 threads, 1 block/SM    lines up to 1024 cells
//   V32  M=32, one factor per cell in smem,  <=256 threads, 2 blocks/SM   (option m=32 only)
//   V32L M=32, one factor per cell in smem,  <=512 threads, 1 block/SM    lines up to 4096 cells
// All run at 128 registers per thread; the M=32 variants spill part of their chunk.
enum { VAR_16 = 0, VAR_32 = 1, VAR_32L = 2, VAR_16L = 3 };

struct Shape {
    int var, M, NS, P, W;  // W: lines per block (KT for strided sweeps, LT for z)
};

inline int pick_shape(adi_ctx *ctx, int n, long opt_w, Shape *s)
{
    if (n > 4096) {
        set_error("adi_cart_step: line too long for the register-resident sweep (n > 4096)");
        return ADI_EINVAL;
    }
    int var = n <= 512 ? VAR_16 : (n <= 1024 ? VAR_16L : VAR_32L);
    if (ctx->opt_wide && opt_w >= 0 && n <= 512) var = VAR_16L;  // 512-thread blocks: twice the lanes per row
    if (ctx->opt_m == 32) var = n <= 1024 ? VAR_32 : VAR_32L;
    const int M = (var == VAR_16 || var == VAR_16L) ? 16 : 32;
    const int maxt = (var == VAR_32L || var == VAR_16L) ? 512 : 256;
    const int P = (n + M - 1) / M;
    int W = 32;
    while (W > 1 && W * P > maxt) W >>= 1;
    if (opt_w > 0) {
        int w = 1;
        while (2 * w <= opt_w && 2 * w <= 32 && 2 * w * P <= maxt) w <<= 1;  // power of two
        W = w;
    }
    if (W < 1) W = 1;
    s->var = var; s->M = M; s->NS = M == 16 ? 2 : 1; s->P = P; s->W = W;
    return ADI_OK;
}

template <typename K, typename... Extra>
int launch(K kern, dim3 grid, dim3 block, size_t smem, cudaStream_t st, adi_ctx *ctx, const SweepArgs &a,
           Extra... extra)
{
    if (smem > 48 * 1024)
        ADI_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, block, smem, st>>>(a, extra...);
    ctx->launches++;
    ADI_CUDA(cudaGetLastError());
    return ADI_OK;
}

template <int AXIS>
int launch_strided_sweep(adi_ctx *ctx, const SweepArgs &a, bool dense, bool extra, bool expl, cudaStream_t st);

}  // namespace adi

// Expands BODY(M, NS, MAXT, MINB) for the variant `var`.
#define ADI_FOR_VARIANT(var, BODY)               \
    switch (var) {                               \
    case adi::VAR_16: BODY(16, 2, 256, 2); break; \
    case adi::VAR_16L: BODY(16, 2, 512, 1); break; \
    case adi::VAR_32: BODY(32, 1, 256, 2); break; \
    default: BODY(32, 1, 512, 1); break;          \
    }
