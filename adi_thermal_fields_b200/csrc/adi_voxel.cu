// adi_voxel.cu -- STL-triangle -> voxel-face projected-area correction of the Robin coefficients
// (voxel_bc_correction.py:53-108 compute_voxel_projected_areas, :110-168 build_corrected_fields).
// This is the producer of the per-face dense h fields that make the ADI coefficients variable
// (SURVEY.md 8f, rank 3).
//
// k_voxel_project  one thread per triangle: subdivision count from the bounding-box span (:72-85),
//                  barycentric sub-triangles (:185-204), centroid -> voxel (:88-100), scatter of
//                  |n_c| * sub_area into the six per-face fields with fp64 atomics (:170-183).
//                  Every floating-point operation is written with the reference's association and
//                  without fused multiply-adds, so a centroid falls into the same voxel as in the
//                  reference; only the ORDER of the additions into one voxel differs (atomics).
// k_voxel_correct  per cell and face: robin = base * proj / dx^2, scale = proj / dx^2, exposed faces the
//                  mesh missed fall back to the base coefficient (:141-166).
#include <stdint.h>

#include <algorithm>

#include "adi_ctx.h"

namespace adi {

struct VoxProjArgs {
    const double *tri;   // [ntri][3][3]
    const double *nrm;   // [ntri][3]
    const double *area;  // [ntri]
    int ntri;
    double ox, oy, oz, dx, eps;
    int max_subdiv;
    const uint8_t *mask;
    int nx, ny, nz;
    double *proj[6];
};

__device__ __forceinline__ double bary1(double c, double a, double b, double v0, double v1, double v2)
{
    return __dadd_rn(__dadd_rn(__dmul_rn(c, v0), __dmul_rn(a, v1)), __dmul_rn(b, v2));  // c*v0 + a*v1 + b*v2
}

__device__ __forceinline__ void vox_scatter(const VoxProjArgs &a, const double p0[3], const double p1[3],
                                            const double p2[3], const double n[3], double sub_area)
{
    const double o[3] = {a.ox, a.oy, a.oz};
    const int shp[3] = {a.nx, a.ny, a.nz};
    long long idx[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const double cen = __ddiv_rn(__dadd_rn(__dadd_rn(p0[d], p1[d]), p2[d]), 3.0);   // np.mean of the 3 vertices
        const double q = floor(__ddiv_rn(__dsub_rn(cen, o[d]), a.dx));
        if (!(q >= 0.0) || !(q < (double)shp[d])) return;
        idx[d] = (long long)q;
    }
    const size_t lin = ((size_t)idx[0] * a.ny + (size_t)idx[1]) * a.nz + (size_t)idx[2];
    if (!a.mask[lin]) return;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const double comp = n[d];
        if (comp > 1e-12) {
            const double v = __dmul_rn(sub_area, comp);
            if (v > 0.0) atomicAdd(a.proj[2 * d + 1] + lin, v);
        } else if (comp < -1e-12) {
            const double v = __dmul_rn(sub_area, -comp);
            if (v > 0.0) atomicAdd(a.proj[2 * d] + lin, v);
        }
    }
}

__global__ void k_voxel_project(const VoxProjArgs a)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.ntri) return;
    const double area = a.area[t];
    if (area <= a.eps) return;
    double v0[3], v1[3], v2[3], n[3];
    double smax = 0.0;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        v0[d] = a.tri[(size_t)t * 9 + d];
        v1[d] = a.tri[(size_t)t * 9 + 3 + d];
        v2[d] = a.tri[(size_t)t * 9 + 6 + d];
        n[d] = a.nrm[(size_t)t * 3 + d];
        const double lo = fmin(v0[d], fmin(v1[d], v2[d])), hi = fmax(v0[d], fmax(v1[d], v2[d]));
        const double span = __ddiv_rn(__dsub_rn(hi, lo), a.dx);
        smax = d == 0 ? span : fmax(smax, span);
    }
    int ns = smax > 1.0 ? (int)ceil(smax) : 1;
    ns = max(1, min(ns, a.max_subdiv));
    if (ns == 1) {
        vox_scatter(a, v0, v1, v2, n, area);
        return;
    }
    const double sub_area = __ddiv_rn(area, (double)(ns * ns));
    const double fn = (double)ns;
    auto bary = [&](int i, int j, double (&p)[3]) {
        const double ca = __ddiv_rn((double)i, fn), cb = __ddiv_rn((double)j, fn);
        const double cc = __dsub_rn(__dsub_rn(1.0, ca), cb);
#pragma unroll
        for (int d = 0; d < 3; ++d) p[d] = bary1(cc, ca, cb, v0[d], v1[d], v2[d]);
    };
    for (int i = 0; i < ns; ++i)
        for (int j = 0; j < ns - i; ++j) {
            double p0[3], p1[3], p2[3];
            bary(i, j, p0); bary(i + 1, j, p1); bary(i, j + 1, p2);
            vox_scatter(a, p0, p1, p2, n, sub_area);
            if (i + j < ns - 1) {
                double p3[3];
                bary(i + 1, j + 1, p3);
                vox_scatter(a, p1, p3, p2, n, sub_area);
            }
        }
}

struct VoxCorrArgs {
    const uint8_t *mask;
    int nx, ny, nz;
    double face_area;
    const double *proj[6];
    int has[6];
    double base[6];
    int fallback;
    double *robin[6];
    double *scale[6];
};

__global__ void k_voxel_correct(const VoxCorrArgs a)
{
    const size_t n = (size_t)a.nx * a.ny * a.nz;
    const size_t snx = (size_t)a.ny * a.nz;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (size_t)gridDim.x * blockDim.x) {
        const int k = (int)(idx % a.nz);
        const size_t ij = idx / a.nz;
        const int j = (int)(ij % a.ny), i = (int)(ij / a.ny);
        const bool act = a.mask[idx] != 0;
        // exposed_mask (adi3d_numba_coeff.py:38-55): active and the neighbour across the face void / outside
        const bool ex[6] = {act && !(i > 0 && a.mask[idx - snx]), act && !(i + 1 < a.nx && a.mask[idx + snx]),
                            act && !(j > 0 && a.mask[idx - a.nz]), act && !(j + 1 < a.ny && a.mask[idx + a.nz]),
                            act && !(k > 0 && a.mask[idx - 1]), act && !(k + 1 < a.nz && a.mask[idx + 1])};
#pragma unroll
        for (int f = 0; f < 6; ++f) {
            if (!a.has[f]) continue;
            double arr = 0.0, scl = 0.0;
            if (a.base[f] != 0.0) {
                const double p = a.proj[f][idx];
                if (p > 0.0) {
                    scl = __ddiv_rn(p, a.face_area);
                    arr = __dmul_rn(a.base[f], scl);
                }
                if (a.fallback && ex[f] && arr <= 0.0) { arr = a.base[f]; scl = 1.0; }
            }
            a.robin[f][idx] = arr;
            if (a.scale[f]) a.scale[f][idx] = scl;
        }
    }
}

}  // namespace adi

using namespace adi;

extern "C" {

int adi_voxel_project(adi_ctx *ctx, const double *d_tri, const double *d_nrm, const double *d_area, int ntri,
                      const double origin[3], double dx, int max_subdiv, double area_eps, const uint8_t *d_mask,
                      int nx, int ny, int nz, double *const d_proj[6], void *stream)
{
    if (!ctx || !origin || !d_mask || !d_proj || ntri < 0 || nx < 1 || ny < 1 || nz < 1 || !(dx > 0.0)) {
        set_error("adi_voxel_project: bad arguments");
        return ADI_EINVAL;
    }
    if (ntri == 0) return ADI_OK;
    if (!d_tri || !d_nrm || !d_area) { set_error("adi_voxel_project: NULL mesh arrays"); return ADI_EINVAL; }
    VoxProjArgs a;
    a.tri = d_tri; a.nrm = d_nrm; a.area = d_area; a.ntri = ntri;
    a.ox = origin[0]; a.oy = origin[1]; a.oz = origin[2]; a.dx = dx; a.eps = area_eps;
    a.max_subdiv = max_subdiv < 1 ? 1 : max_subdiv;
    a.mask = d_mask; a.nx = nx; a.ny = ny; a.nz = nz;
    for (int f = 0; f < 6; ++f) {
        if (!d_proj[f]) { set_error("adi_voxel_project: six output fields are required"); return ADI_EINVAL; }
        a.proj[f] = d_proj[f];
    }
    const int threads = 128;
    k_voxel_project<<<(ntri + threads - 1) / threads, threads, 0, (cudaStream_t)stream>>>(a);
    ctx->launches++;
    ADI_CUDA(cudaGetLastError());
    return ADI_OK;
}

int adi_voxel_correct(adi_ctx *ctx, const uint8_t *d_mask, int nx, int ny, int nz, double dx,
                      const double *const d_proj[6], const int has[6], const double base_h[6], int fallback,
                      double *const d_robin[6], double *const d_scale[6], void *stream)
{
    if (!ctx || !d_mask || !d_proj || !has || !base_h || !d_robin || nx < 1 || ny < 1 || nz < 1 || !(dx > 0.0)) {
        set_error("adi_voxel_correct: bad arguments");
        return ADI_EINVAL;
    }
    VoxCorrArgs a;
    a.mask = d_mask; a.nx = nx; a.ny = ny; a.nz = nz;
    a.face_area = dx * dx;
    a.fallback = fallback;
    for (int f = 0; f < 6; ++f) {
        a.has[f] = has[f];
        a.base[f] = base_h[f];
        a.proj[f] = d_proj[f];
        a.robin[f] = d_robin[f];
        a.scale[f] = d_scale ? d_scale[f] : nullptr;
        if (has[f] && (!d_proj[f] || !d_robin[f])) { set_error("adi_voxel_correct: missing field for a requested face"); return ADI_EINVAL; }
    }
    const size_t n = (size_t)nx * ny * nz;
    const int threads = 256;
    const int blocks = (int)std::min<size_t>((n + threads - 1) / threads, 148 * 32);
    k_voxel_correct<<<blocks, threads, 0, (cudaStream_t)stream>>>(a);
    ctx->launches++;
    ADI_CUDA(cudaGetLastError());
    return ADI_OK;
}

}  // extern "C"
