// adi_cart.cuh -- sm_100a kernels of the Cartesian ADI step
// (adi3d_numba_coeff.py:290-302 / adi3d_gpu_coeff.py:213-230).
//
// K0 k_build_code      mask (+dir_mask) -> 1-byte neighbour code per cell
// K1 k_sweep_strided   x sweep (stride ny*nz, explicit stage fused) and y sweep (stride nz):
//                      lanes run along z, so every load/store of a warp is a run of
//                      contiguous 8-byte cells (64-256 B rows); each thread keeps a
//                      16-cell chunk of its line in registers
// K3 k_sweep_z         z sweep (contiguous axis): the tile of lines is staged through
//                      padded shared memory with coalesced 16-byte accesses, then the same
//                      register-resident chunk solve runs with lanes along the line
// K7 k_build_packs     precompute_coeff_packs_unified on the device
#pragma once
#include <cuda_runtime.h>

#include "adi_core.h"

namespace adi {

struct SweepArgs {
    const double *__restrict__ in;
    double *__restrict__ out;
    const uint8_t *__restrict__ code;
    const double *__restrict__ coeff;  // CMODE 2
    const double *__restrict__ q;      // EXTRA, may be null
    const double *__restrict__ dirv;   // EXTRA, may be null
    int nx, ny, nz;
    SweepConst k;
};

// ------------------------------------------------------------------------------------
// K0: neighbour code.  One thread per cell; the six neighbour bytes come from L1/L2.
// ------------------------------------------------------------------------------------
__global__ void k_build_code(const uint8_t *__restrict__ mask, const uint8_t *__restrict__ dirm,
                             uint8_t *__restrict__ code, int nx, int ny, int nz)
{
    const size_t n = (size_t)nx * ny * nz;
    const size_t snx = (size_t)ny * nz;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n;
         idx += (size_t)gridDim.x * blockDim.x) {
        const int k = (int)(idx % nz);
        const size_t ij = idx / nz;
        const int j = (int)(ij % ny);
        const int i = (int)(ij / ny);
        unsigned c = 0;
        if (mask[idx]) {
            c = CB_SELF;
            if (dirm && dirm[idx]) c |= CB_DIR;
        }
        if (i > 0 && mask[idx - snx]) c |= CB_XM;
        if (i + 1 < nx && mask[idx + snx]) c |= CB_XP;
        if (j > 0 && mask[idx - nz]) c |= CB_YM;
        if (j + 1 < ny && mask[idx + nz]) c |= CB_YP;
        if (k > 0 && mask[idx - 1]) c |= CB_ZM;
        if (k + 1 < nz && mask[idx + 1]) c |= CB_ZP;
        code[idx] = (uint8_t)c;
    }
}

// ------------------------------------------------------------------------------------
// Reduced-system solve shared by all sweeps: exchange through shared memory.
// red: 6*NTH doubles.  ridx: this thread's slot; rstep: slot distance between consecutive
// chunks of the same line.  Returns S_p in r.D and S_{p-1} in *Sl.
// ------------------------------------------------------------------------------------
template <int M>
__device__ __forceinline__ double solve_reduced(const Chunk<M> &ch, const First &f, double *red,
                                                int NTH, int ridx, int rstep, int p, int P,
                                                double *Sl)
{
    red[ridx] = f.Y;
    red[NTH + ridx] = f.V;
    red[2 * NTH + ridx] = f.W;
    __syncthreads();
    First nx;
    nx.Y = 0.0; nx.V = 0.0; nx.W = 0.0;
    if (p + 1 < P) {
        nx.Y = red[ridx + rstep];
        nx.V = red[NTH + ridx + rstep];
        nx.W = red[2 * NTH + ridx + rstep];
    }
    Red r = chunk_reduced_row(ch, nx);
    int cur = 1;
    for (int s = 1; s < P; s <<= 1) {
        double *b = red + cur * 3 * NTH;
        b[ridx] = r.A;
        b[NTH + ridx] = r.C;
        b[2 * NTH + ridx] = r.D;
        __syncthreads();
        Red lo, hi;
        lo.A = lo.C = lo.D = 0.0;
        hi.A = hi.C = hi.D = 0.0;
        if (p - s >= 0) {
            const int o = ridx - s * rstep;
            lo.A = b[o]; lo.C = b[NTH + o]; lo.D = b[2 * NTH + o];
        }
        if (p + s < P) {
            const int o = ridx + s * rstep;
            hi.A = b[o]; hi.C = b[NTH + o]; hi.D = b[2 * NTH + o];
        }
        r = pcr_step(r, lo, hi);
        cur ^= 1;
    }
    double *b = red + cur * 3 * NTH;
    b[ridx] = r.D;
    __syncthreads();
    *Sl = (p > 0) ? b[ridx - rstep] : 0.0;
    return r.D;
}

// ------------------------------------------------------------------------------------
// K1: sweeps along the strided axes.  blockDim = (KT lines along z, P chunks);
// grid = (ceil(nz/KT), ny) for AXIS 0 and (ceil(nz/KT), nx) for AXIS 1.
// EXPL (AXIS 0 only): the input is T^n and the explicit stage
// R0 = T + beta*(Lx+Ly+Lz) (adi3d_numba_coeff.py:298) is applied while loading.
// ------------------------------------------------------------------------------------
template <int AXIS, int M, int CMODE, bool EXTRA, bool EXPL>
__global__ void __launch_bounds__(512, 1) k_sweep_strided(const SweepArgs a)
{
    extern __shared__ double red[];
    const int KT = blockDim.x, P = blockDim.y;
    const int kk = threadIdx.x, p = threadIdx.y;
    const int NTH = KT * P;
    const int k = blockIdx.x * KT + kk;
    const bool lane_ok = k < a.nz;
    const int n = (AXIS == 0) ? a.nx : a.ny;
    const size_t sline = (AXIS == 0) ? (size_t)a.ny * a.nz : (size_t)a.nz;
    const size_t base = ((AXIS == 0) ? (size_t)blockIdx.y * a.nz : (size_t)blockIdx.y * a.ny * a.nz) + k;
    constexpr unsigned LO = (AXIS == 0) ? CB_XM : CB_YM;
    constexpr unsigned HI = (AXIS == 0) ? CB_XP : CB_YP;
    const int t0 = p * M;

    Chunk<M> ch;
    double Q[EXTRA ? M : 1], DV[EXTRA ? M : 1];
#pragma unroll
    for (int e = 0; e < M; ++e) {
        const bool ok = lane_ok && (t0 + e) < n;
        const size_t idx = base + (size_t)(t0 + e) * sline;
        ch.code[e] = ok ? (unsigned)a.code[idx] : 0u;
        ch.T[e] = ok ? a.in[idx] : 0.0;
        ch.Cc[e] = (CMODE == 2 && ok) ? a.coeff[idx] : 0.0;
        if (EXTRA) {
            Q[e] = (a.q && ok) ? a.q[idx] : 0.0;
            DV[e] = (a.dirv && ok && (ch.code[e] & CB_DIR)) ? a.dirv[idx] : 0.0;
        }
    }
    if (EXPL) {
        const double *__restrict__ in = a.in;
        const unsigned c0 = ch.code[0], cl = ch.code[M - 1];
        double prev = ((c0 & CB_SELF) && (c0 & CB_XM)) ? in[base + (size_t)(t0 - 1) * sline] : 0.0;
        const double nxt = ((cl & CB_SELF) && (cl & CB_XP)) ? in[base + (size_t)(t0 + M) * sline] : 0.0;
#pragma unroll
        for (int e = 0; e < M; ++e) {
            const unsigned c = ch.code[e];
            const size_t idx = base + (size_t)(t0 + e) * sline;
            const bool act = (c & CB_SELF) != 0;
            const double ym = (act && (c & CB_YM)) ? in[idx - a.nz] : 0.0;
            const double yp = (act && (c & CB_YP)) ? in[idx + a.nz] : 0.0;
            const double zm = (act && (c & CB_ZM)) ? in[idx - 1] : 0.0;
            const double zp = (act && (c & CB_ZP)) ? in[idx + 1] : 0.0;
            const double xp = (e < M - 1) ? ch.T[e + 1] : nxt;
            const double r0 = explicit_r0(c, ch.T[e], prev, xp, ym, yp, zm, zp, a.k);
            prev = ch.T[e];
            ch.T[e] = r0;
        }
    }

    const First f = chunk_forward<M, CMODE, EXTRA>(ch, Q, DV, LO, HI, a.k);
    double Sl;
    const double S = solve_reduced<M>(ch, f, red, NTH, p * KT + kk, KT, p, P, &Sl);
    chunk_backward<M>(ch, Sl, S);

#pragma unroll
    for (int e = 0; e < M; ++e) {
        const bool ok = lane_ok && (t0 + e) < n;
        if (ok) a.out[base + (size_t)(t0 + e) * sline] = ch.T[e];
    }
}

// ------------------------------------------------------------------------------------
// K3: sweep along z (contiguous).  blockDim = (P chunks, LT lines); grid = ceil(nx*ny/LT).
// Shared memory: sT[LT][P*(M+2)] (+ sC when CMODE 2), sCode[LT][P*M], red[6*NTH].
// The +2 padding per chunk makes the per-thread 16-byte reads of a quarter warp hit
// eight different 16-byte bank groups (stride 144 B).
// ------------------------------------------------------------------------------------
template <int M>
__host__ __device__ constexpr int zpad() { return M + 2; }

template <int M, int CMODE, bool EXTRA>
__global__ void __launch_bounds__(512, 1) k_sweep_z(const SweepArgs a)
{
    extern __shared__ double smem[];
    const int P = blockDim.x, LT = blockDim.y;
    const int p = threadIdx.x, ln = threadIdx.y;
    const int NTH = P * LT;
    const int tid = ln * P + p;
    const int nz = a.nz;
    const size_t nlines = (size_t)a.nx * a.ny;
    const size_t L0 = (size_t)blockIdx.x * LT;
    const int LS = P * zpad<M>();  // doubles per staged line
    double *sT = smem;
    double *sC = sT + (size_t)LT * LS;
    double *red = sC + (CMODE == 2 ? (size_t)LT * LS : 0);
    uint8_t *sCode = reinterpret_cast<uint8_t *>(red + 6 * NTH);

    // ---- stage in (coalesced) ----
    const bool vec2 = (nz & 1) == 0;
    for (int l = 0; l < LT; ++l) {
        const size_t line = L0 + l;
        const bool lok = line < nlines;
        const size_t g0 = line * (size_t)nz;
        if (vec2) {
            for (int tv = tid; tv < (nz >> 1); tv += NTH) {
                const int t = tv << 1;
                const int so = l * LS + (t / M) * zpad<M>() + (t % M);
                double2 v = make_double2(0.0, 0.0);
                if (lok) v = *reinterpret_cast<const double2 *>(a.in + g0 + t);
                *reinterpret_cast<double2 *>(sT + so) = v;
                if (CMODE == 2) {
                    double2 c = make_double2(0.0, 0.0);
                    if (lok) c = *reinterpret_cast<const double2 *>(a.coeff + g0 + t);
                    *reinterpret_cast<double2 *>(sC + so) = c;
                }
            }
        } else {
            for (int t = tid; t < nz; t += NTH) {
                const int so = l * LS + (t / M) * zpad<M>() + (t % M);
                sT[so] = lok ? a.in[g0 + t] : 0.0;
                if (CMODE == 2) sC[so] = lok ? a.coeff[g0 + t] : 0.0;
            }
        }
        if ((nz & 15) == 0 && P * M == nz) {
            for (int tv = tid; tv < (nz >> 4); tv += NTH) {
                uint4 v = make_uint4(0u, 0u, 0u, 0u);
                if (lok) v = *reinterpret_cast<const uint4 *>(a.code + g0 + (tv << 4));
                *reinterpret_cast<uint4 *>(sCode + (size_t)l * P * M + (tv << 4)) = v;
            }
        } else {
            for (int t = tid; t < P * M; t += NTH)
                sCode[(size_t)l * P * M + t] = (lok && t < nz) ? a.code[g0 + t] : (uint8_t)0;
        }
    }
    __syncthreads();

    // ---- chunk to registers ----
    Chunk<M> ch;
    double Q[EXTRA ? M : 1], DV[EXTRA ? M : 1];
    const int so0 = ln * LS + p * zpad<M>();
    {
        const uint8_t *cb = sCode + (size_t)ln * P * M + p * M;
#pragma unroll
        for (int e = 0; e < M; e += 4) {
            const unsigned w = *reinterpret_cast<const unsigned *>(cb + e);
            ch.code[e] = w & 0xffu;
            ch.code[e + 1] = (w >> 8) & 0xffu;
            ch.code[e + 2] = (w >> 16) & 0xffu;
            ch.code[e + 3] = w >> 24;
        }
#pragma unroll
        for (int e = 0; e < M; e += 2) {
            const double2 v = *reinterpret_cast<const double2 *>(sT + so0 + e);
            ch.T[e] = v.x;
            ch.T[e + 1] = v.y;
            if (CMODE == 2) {
                const double2 c = *reinterpret_cast<const double2 *>(sC + so0 + e);
                ch.Cc[e] = c.x;
                ch.Cc[e + 1] = c.y;
            } else {
                ch.Cc[e] = 0.0;
                ch.Cc[e + 1] = 0.0;
            }
        }
    }
    if (EXTRA) {
        const size_t line = L0 + ln;
        const size_t g0 = line * (size_t)nz + (size_t)p * M;
#pragma unroll
        for (int e = 0; e < M; ++e) {
            const bool ok = line < nlines && (p * M + e) < nz;
            Q[e] = (a.q && ok) ? a.q[g0 + e] : 0.0;
            DV[e] = (a.dirv && ok && (ch.code[e] & CB_DIR)) ? a.dirv[g0 + e] : 0.0;
        }
    }

    const First f = chunk_forward<M, CMODE, EXTRA>(ch, Q, DV, CB_ZM, CB_ZP, a.k);
    double Sl;
    const double S = solve_reduced<M>(ch, f, red, NTH, tid, 1, p, P, &Sl);
    chunk_backward<M>(ch, Sl, S);

    // ---- results back through shared memory (each thread owns its slots) ----
#pragma unroll
    for (int e = 0; e < M; e += 2)
        *reinterpret_cast<double2 *>(sT + so0 + e) = make_double2(ch.T[e], ch.T[e + 1]);
    __syncthreads();
    for (int l = 0; l < LT; ++l) {
        const size_t line = L0 + l;
        if (line >= nlines) break;
        const size_t g0 = line * (size_t)nz;
        if (vec2) {
            for (int tv = tid; tv < (nz >> 1); tv += NTH) {
                const int t = tv << 1;
                const int so = l * LS + (t / M) * zpad<M>() + (t % M);
                *reinterpret_cast<double2 *>(a.out + g0 + t) = *reinterpret_cast<const double2 *>(sT + so);
            }
        } else {
            for (int t = tid; t < nz; t += NTH)
                a.out[g0 + t] = sT[l * LS + (t / M) * zpad<M>() + (t % M)];
        }
    }
}

// ------------------------------------------------------------------------------------
// K7: exposed_mask / precompute_coeff_packs_unified (adi3d_gpu_coeff.py:31-110).
// ------------------------------------------------------------------------------------
struct PackArgs {
    const uint8_t *mask;
    int nx, ny, nz;
    double A, Ccell;  // dx*dx, rho*cp*dx^3 (adi3d_numba_coeff.py:69-71)
    int h_kind[6];
    double h_scalar[6];
    const double *h_field[6];
    int q_kind[6];
    double q_scalar[6];
    const double *q_field[6];
    double *coeff[3];
    double *qout[3];
};

__device__ __forceinline__ unsigned exposed_bits(const uint8_t *__restrict__ mask, size_t idx, int i,
                                                 int j, int k, int nx, int ny, int nz)
{
    // bit f set <=> cell active and neighbour across face f void or outside (:38-55)
    if (!mask[idx]) return 0u;
    const size_t snx = (size_t)ny * nz;
    unsigned b = 0;
    if (!(i > 0 && mask[idx - snx])) b |= 1u;
    if (!(i + 1 < nx && mask[idx + snx])) b |= 2u;
    if (!(j > 0 && mask[idx - nz])) b |= 4u;
    if (!(j + 1 < ny && mask[idx + nz])) b |= 8u;
    if (!(k > 0 && mask[idx - 1])) b |= 16u;
    if (!(k + 1 < nz && mask[idx + 1])) b |= 32u;
    return b;
}

__global__ void k_build_packs(const PackArgs a)
{
    const size_t n = (size_t)a.nx * a.ny * a.nz;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n;
         idx += (size_t)gridDim.x * blockDim.x) {
        const int k = (int)(idx % a.nz);
        const size_t ij = idx / a.nz;
        const int j = (int)(ij % a.ny);
        const int i = (int)(ij / a.ny);
        const unsigned ex = exposed_bits(a.mask, idx, i, j, k, a.nx, a.ny, a.nz);
#pragma unroll
        for (int ax = 0; ax < 3; ++ax) {
            double c = 0.0, q = 0.0;
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                const int f = 2 * ax + s;
                if (ex & (1u << f)) {
                    if (a.h_kind[f]) {
                        const double h = a.h_kind[f] == 2 ? a.h_field[f][idx] : a.h_scalar[f];
                        c += __ddiv_rn(__dmul_rn(h, a.A), a.Ccell);   // (:99) h*A/Ccell
                    }
                    if (a.q_kind[f]) {
                        const double qv = a.q_kind[f] == 2 ? a.q_field[f][idx] : a.q_scalar[f];
                        q += __ddiv_rn(__dmul_rn(qv, a.A), a.Ccell);  // (:111)
                    }
                }
            }
            if (a.coeff[ax]) a.coeff[ax][idx] = c;
            if (a.qout[ax]) a.qout[ax][idx] = q;
        }
    }
}

__global__ void k_exposed_mask(const uint8_t *__restrict__ mask, uint8_t *__restrict__ out, int face,
                               int nx, int ny, int nz)
{
    const size_t n = (size_t)nx * ny * nz;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n;
         idx += (size_t)gridDim.x * blockDim.x) {
        const int k = (int)(idx % nz);
        const size_t ij = idx / nz;
        const int j = (int)(ij % ny);
        const int i = (int)(ij / ny);
        out[idx] = (uint8_t)((exposed_bits(mask, idx, i, j, k, nx, ny, nz) >> face) & 1u);
    }
}

}  // namespace adi
