// adi_cyl.cu -- cylindrical (r, phi, z) backward-Euler ADI step on sm_100a
// (adi3d_cyl_phi_v3.py:332-350, scheme "be") and its activation-mask wrapper
// (quick_spiral_deposition_gif_v5.py:31-70).
//
// K4 k_cyl_strided   r sweep (stride nphi*nz_pitch) and periodic phi sweep (stride nz_pitch):
//                    lanes run along z, so a warp's loads/stores are runs of contiguous cells;
//                    each thread keeps one chunk of its line in registers
// K6 k_cyl_z         z sweep (contiguous): a tile of lines is staged through padded shared
//                    memory (conflict-free column reads), then the same chunk solve runs with
//                    the threads of a line side by side
// All matrix work is tabulated on the host (adi_tab_core.h); the kernels stream right-hand
// sides.  The r sweep applies the prologue of the step (source term :339, void clamp of
// adi_step_masked :55-57), the z sweep its epilogue (void / axis clamps :61-68).
#include <stdint.h>

#include <algorithm>
#include <cstring>
#include <string>
#include <vector>

#include "adi_cart.cuh"
#include "adi_core.h"
#include "adi_ctx.h"
#include "adi_tab_core.h"

namespace adi {

int launch_sweep_zt(adi_ctx *ctx, const SweepArgs &a, bool dense, bool extra, int zmode, cudaStream_t st, int *used);

struct CylTables {
    // cache key: everything the tables depend on
    adi_cyl_params key;
    bool valid = false;
    int nr = 0, nphi = 0, nz = 0, M = 0;
    TabGeom gr, gp, gz;
    double *d_blob = nullptr;  // [r blob | z blob | nr phi blobs]
    int *d_geom = nullptr;     // [r ints | z ints | phi ints]
    size_t blob_cap = 0, geom_cap = 0;
    size_t off_r = 0, off_z = 0, off_p = 0;   // blob offsets (doubles)
    size_t goff_r = 0, goff_z = 0, goff_p = 0;  // geom offsets (ints)
    double add_r = 0.0;
    ZEnd bot, top;
    // z-slab decomposition
    int slab_rank = 0, slab_nranks = 1;
    std::vector<int> slab_nz;          // local nz of every rank
    double *d_G = nullptr;             // [2][nranks][2] ghost map
    double *d_ghost = nullptr;         // [2][nlines]
    size_t ghost_lines = 0;
    // the z rows as a Cartesian sweep (k_sweep_zt): one line of neighbour codes shared by all z lines, the scalars of
    // its rows; zt_ok = the boundary rows can be written that way (no Dirichlet end, one ambient temperature)
    uint8_t *d_zcode = nullptr;
    size_t zcode_cap = 0;
    bool zt_ok = false;
    SweepConst zk;
};

void cyl_release(adi_ctx *ctx)
{
    if (!ctx->cyl) return;
    if (ctx->cyl->d_blob) cudaFree(ctx->cyl->d_blob);
    if (ctx->cyl->d_geom) cudaFree(ctx->cyl->d_geom);
    if (ctx->cyl->d_G) cudaFree(ctx->cyl->d_G);
    if (ctx->cyl->d_ghost) cudaFree(ctx->cyl->d_ghost);
    if (ctx->cyl->d_zcode) cudaFree(ctx->cyl->d_zcode);
    delete ctx->cyl;
    ctx->cyl = nullptr;
}

struct CylArgs {
    const double *in;       // may alias out (phi and z sweeps run in place)
    double *out;
    const double *blob;     // table blob of blockIdx.y == 0
    const int *geom;
    long long blob_stride;  // doubles between the blobs of consecutive blockIdx.y (phi: per ring)
    TabGeom g;
    int nz;                 // valid cells along z (lanes of the strided sweeps / line length of z)
    long long cell_stride;  // between consecutive cells of a strided line
    long long outer_stride; // between consecutive blockIdx.y planes (strided) / lines (z)
    long long nlines;       // z sweep: number of lines
    int nphi;               // z sweep: line / nphi = ring index
    // prologue (r sweep)
    const uint8_t *active;  // NULL: unmasked
    double T_void, T_inner;
    const double *S;        // NULL: no source
    double dt, rho_cp;
    // right-hand side boundary terms: first / last cell of the line
    int set_first, set_last;
    double val_first, val_last;
    // z-slab decomposition (z sweep): pass 1 writes y[2][nlines] = (yf, yl), pass 2 reads ghost[2][nlines]
    double *y;
    const double *ghost;
};

__device__ __forceinline__ void cyl_cp_async8(double *dst_smem, const double *src)
{
    const unsigned s = (unsigned)__cvta_generic_to_shared(dst_smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(src) : "memory");
}
__device__ __forceinline__ void cyl_cp_async16(void *dst_smem, const void *src)
{
    const unsigned s = (unsigned)__cvta_generic_to_shared(dst_smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(src) : "memory");
}
__device__ __forceinline__ void cyl_cp_async_wait() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// Reduced system over the P chunks of a line.  ex: 3*NTH doubles; slot(q) = index of chunk q of
// this thread's line.  tab: this line's table blob (global, read through L1).
// Returns S_p, *Sl = S_{p-1}.
// SM: the tables have been staged in shared memory (plain loads) -- otherwise global memory through L1 (__ldg)
template <bool SM>
__device__ __forceinline__ double tld(const double *p) { return SM ? *p : tab_ld(p); }
template <bool SM>
__device__ __forceinline__ TabPair tld2(const double *p)
{
    if (!SM) return tab_ld2(p);
    const double2 t = *reinterpret_cast<const double2 *>(p);
    TabPair r;
    r.a = t.x; r.b = t.y;
    return r;
}

template <bool SM = false, class SLOT>
__device__ __forceinline__ double cyl_reduced(const TabGeom &g, const double *__restrict__ tab, double *ex,
                                              int NTH, int p, double ds, double Y, double Yl, SLOT slot,
                                              double *Sl)
{
    const int P = g.P, cyc = g.cyclic;
    const double t0 = tld<SM>(tab + g.o_t0 + p), t1 = tld<SM>(tab + g.o_t1 + p), t2 = tld<SM>(tab + g.o_t2 + p);
    double *sY = ex + 2 * NTH;
    sY[slot(p)] = Y;
    __syncthreads();
    const double Ynext = sY[slot(tab_hi(p, 1, P, cyc))];
    double D = tab_reduced_rhs(t0, t1, t2, ds, Yl, Ynext);
    int cur = 0;
    for (int l = 0; l < g.levels; ++l) {
        const int s = 1 << l;
        const double *R = tab + g.o_lvl + l * 3 * P;
        const double cr = tld<SM>(R + p), ca = tld<SM>(R + P + p), cc = tld<SM>(R + 2 * P + p);
        double *b = ex + cur * NTH;
        b[slot(p)] = D;
        __syncthreads();
        const double Dlo = b[slot(tab_lo(p, s, P, cyc))];
        const double Dhi = b[slot(tab_hi(p, s, P, cyc))];
        D = tab_level(cr, ca, cc, D, Dlo, Dhi);
        cur ^= 1;
    }
    double *b = ex + cur * NTH;
    b[slot(p)] = D;
    __syncthreads();
    *Sl = b[slot(tab_lo(p, 1, P, cyc))];
    return D;
}

// tab_forward / tab_backward (adi_tab_core.h) on tables staged in shared memory
template <int M>
__device__ __forceinline__ double tab_forward_sm(double (&d)[M], const double *f, const double *alpha, double *Yl)
{
    double dp = 0.0, Y = 0.0;
#pragma unroll
    for (int e = 0; e < M - 1; ++e) {
        const TabPair t = tld2<true>(f + 2 * e);
        dp = fma(t.b, dp, d[e] * t.a);
        d[e] = dp;
        Y = fma(alpha[e], dp, Y);
    }
    *Yl = dp;
    return Y;
}
template <int M>
__device__ __forceinline__ void tab_backward_sm(double (&d)[M], const double *b, double Sl, double S)
{
    double xn = S;
    d[M - 1] = S;
#pragma unroll
    for (int e = M - 2; e >= 0; --e) {
        const TabPair t = tld2<true>(b + 2 * e);
        const double x = fma(t.a, xn, fma(t.b, Sl, d[e]));
        d[e] = x;
        xn = x;
    }
}

// ------------------------------------------------------------------------------------
// K4: strided sweeps.  blockDim = (KT lanes along z, P chunks); grid = (ceil(nz/KT), nouter).
// PRO: r sweep of the step -- applies the void clamp and the source term while loading.
// BCL: the last cell of the line takes a boundary term (outer Robin row of the r sweep).
// Shared memory: 3*NTH doubles (reduced-system exchange) only.
// ------------------------------------------------------------------------------------
// SM: the block first copies the line's tables (everything in front of o_dl: cell tables, reduced-system assembly, PCR
// levels -- a few KB) into shared memory and then walks over `nzt` consecutive z tiles with them: ncu (r03c) had the
// kernel waiting on its table reads (five L1 loads per cell, long-scoreboard 10.7 / 11.4 per issued instruction, in
// the back substitution too, where nothing else is loaded).  (The z sweep keeps its tables in L1: they are as large as
// eight of its tiles, and staging them measured slower -- 0.52 against 0.48 ms.)
template <int M, bool PRO, bool BCL, bool SM>
__global__ void __launch_bounds__(256, (M <= 16 ? 3 : 2)) k_cyl_strided(const CylArgs a, const int nzt)
{
    extern __shared__ double smem[];
    const int KT = blockDim.x, P = blockDim.y;
    const int kk = threadIdx.x, p = threadIdx.y;
    const int NTH = KT * P;
    const double *__restrict__ gtab = a.blob + (size_t)blockIdx.y * a.blob_stride;
    const double *__restrict__ tab = gtab;
    if (SM) {
        double *stab = smem + 3 * NTH;
        const int tid = p * KT + kk;
        for (int i = tid; 2 * i < a.g.o_dl; i += NTH)           // (o_dl is even, the blob 16-byte aligned)
            reinterpret_cast<double2 *>(stab)[i] = __ldg(reinterpret_cast<const double2 *>(gtab) + i);
        tab = stab;
        __syncthreads();
    }
    const int cb = __ldg(a.geom + p), endp = __ldg(a.geom + P + p), len = __ldg(a.geom + 2 * P + p);
    const int ntz = (a.nz + KT - 1) / KT;
  for (int zt = blockIdx.x * nzt; zt < min((int)(blockIdx.x + 1) * nzt, ntz); ++zt) {
    if (zt != (int)blockIdx.x * nzt) __syncthreads();           // the exchange buffer of the previous tile is free

    const int k = zt * KT + kk;
    const bool lane_ok = k < a.nz;
    // slot e holds cell i0 + e of the line; slots e < efirst are padding (never dereferenced)
    const int efirst = lane_ok ? M - len : M;
    // byte offset of slot e: e * cs8 (one 32 x 32 -> 64 bit multiply-add per access)
    const unsigned cs8 = (unsigned)a.cell_stride * 8u;
    const long long first = (long long)blockIdx.y * a.outer_stride + min(k, a.nz - 1) + (long long)(endp - (M - 1)) * a.cell_stride;
    const char *src = reinterpret_cast<const char *>(a.in + first);

    double d[M];
#pragma unroll
    for (int e = 0; e < M; ++e) d[e] = e >= efirst ? *reinterpret_cast<const double *>(src + (size_t)e * cs8) : 0.0;
    if (PRO) {
        if (a.active) {  // T_work[~active] = T_void
            const uint8_t *am = a.active + first;
            const unsigned cs1 = (unsigned)a.cell_stride;
            // unconditional byte loads (padding slots re-read a cell of the chunk), all in flight at once
            unsigned mb[M];
            const int elast = M - 1;
#pragma unroll
            for (int e = 0; e < M; ++e) mb[e] = am[(size_t)min(max(e, efirst), elast) * cs1];
#pragma unroll
            for (int e = 0; e < M; ++e) d[e] = mb[e] ? d[e] : (e >= efirst ? a.T_void : 0.0);
        }
        if (a.S) {       // R0 = Tn + dt*(S/(rho*cp))  :339
            const char *sp = reinterpret_cast<const char *>(a.S + first);
#pragma unroll
            for (int e = 0; e < M; ++e)
                if (e >= efirst)
                    d[e] = __dadd_rn(d[e], __dmul_rn(a.dt, __ddiv_rn(*reinterpret_cast<const double *>(sp + (size_t)e * cs8), a.rho_cp)));
        }
    }
    if (BCL) {  // the last cell of the line is the separator of the last chunk
        if (p == P - 1 && lane_ok) d[M - 1] = a.set_last ? a.val_last : __dadd_rn(d[M - 1], a.val_last);
    }
    double Yl;
    const double Y = SM ? tab_forward_sm<M>(d, tab + a.g.o_f + 2 * cb, tab + a.g.o_alpha + cb, &Yl)
                        : tab_forward<M>(d, tab + a.g.o_f + 2 * cb, tab + a.g.o_alpha + cb, &Yl);
    double Sl;
    const double S = cyl_reduced<SM>(a.g, tab, smem, NTH, p, d[M - 1], Y, Yl, [=](int q) { return q * KT + kk; }, &Sl);
    if (SM) tab_backward_sm<M>(d, tab + a.g.o_b + 2 * cb, Sl, S);
    else tab_backward<M>(d, tab + a.g.o_b + 2 * cb, Sl, S);
    char *dst = reinterpret_cast<char *>(a.out + first);
#pragma unroll
    for (int e = 0; e < M; ++e)
        if (e >= efirst) *reinterpret_cast<double *>(dst + (size_t)e * cs8) = d[e];
  }
}

// ------------------------------------------------------------------------------------
// K6: z sweep.  blockDim = (P chunks, LT lines); grid = ceil(nlines/LT).
// The tile of LT lines is staged with cp.async (every byte of the tile in flight at once) and
// each thread then pulls its chunk out of shared memory.
//   VEC  (nz a multiple of M, 16-byte aligned lines): 16-byte copies; inside each block of 16
//        cells the 16-byte pairs are stored at pair ^ (chunk & 7), so the double2 column reads of
//        8 neighbouring chunks hit 8 different bank groups;
//   else 8-byte copies into a padded line (cell z at z + (z >> 4)): conflict-free 64-bit reads.
// EPI: last sweep of a masked step -- void cells := T_void, void axis cells := T_inner.
// Shared memory: ex[3*NTH] | sT[LT][RL] | (EPI && VEC) sMask[LT][nz]
// ------------------------------------------------------------------------------------
__device__ __forceinline__ int zphys(int z) { return z + (z >> 4); }
template <int M>
__device__ __forceinline__ int zswz(int z)
{
    return (z & ~15) | ((((z >> 1) & 7) ^ ((z / M) & 7)) << 1) | (z & 1);
}

// ZM 0: whole lines.  1: z-slab pass 1 -- the right-hand-side part (yf, yl) of each segment's interface
// relation, nothing stored.  2: z-slab pass 2 -- finishes the segment with the ghost values a.ghost.
template <int M, bool EPI, bool VEC, int ZM = 0>
__global__ void __launch_bounds__(256, (M <= 16 ? 3 : 2)) k_cyl_z(const CylArgs a)
{
    extern __shared__ double smem[];
    const int P = blockDim.x, LT = blockDim.y;
    const int p = threadIdx.x, ln = threadIdx.y;
    const int NTH = P * LT, tid = ln * P + p;
    const int nz = a.nz;
    const int RL = VEC ? nz : zphys(nz - 1) + 1;  // doubles per staged line
    double *ex = smem;
    double *sT = ex + 3 * NTH;
    uint8_t *sMask = reinterpret_cast<uint8_t *>(sT + (size_t)LT * RL);
    const double *__restrict__ tab = a.blob;

    const long long L0 = (long long)blockIdx.x * LT;
    const int nl = (int)min((long long)LT, a.nlines - L0);
    if (VEC) {
        const int ppl = nz >> 1;
        for (int l = 0; l < nl; ++l) {
            const double *src = a.in + (size_t)(L0 + l) * a.outer_stride;
            for (int pl = tid; pl < ppl; pl += NTH) cyl_cp_async16(sT + (size_t)l * RL + zswz<M>(2 * pl), src + 2 * pl);
            if (EPI) {
                const uint8_t *am = a.active + (size_t)(L0 + l) * a.outer_stride;
                for (int c = tid; c < (nz >> 4); c += NTH) cyl_cp_async16(sMask + (size_t)l * nz + 16 * c, am + 16 * c);
            }
        }
    } else {
        for (int l = 0; l < nl; ++l) {
            const double *src = a.in + (size_t)(L0 + l) * a.outer_stride;
            for (int z = tid; z < nz; z += NTH) cyl_cp_async8(sT + (size_t)l * RL + zphys(z), src + z);
        }
    }
    const int cb = __ldg(a.geom + p), endp = __ldg(a.geom + P + p), len = __ldg(a.geom + 2 * P + p);
    cyl_cp_async_wait();
    __syncthreads();
    const bool line_ok = ln < nl;
    double *myT = sT + (size_t)ln * RL;
    const int i0 = endp - (M - 1);
    const int efirst = line_ok ? M - len : M;

    double d[M];
    if (VEC) {  // every chunk is full and starts at p*M
#pragma unroll
        for (int j = 0; j < M / 2; ++j) {
            const double2 v = *reinterpret_cast<const double2 *>(myT + zswz<M>(p * M + 2 * j));
            d[2 * j] = line_ok ? v.x : 0.0; d[2 * j + 1] = line_ok ? v.y : 0.0;
        }
    } else {
#pragma unroll
        for (int e = 0; e < M; ++e) d[e] = e >= efirst ? myT[zphys(i0 + e)] : 0.0;
    }
    if (line_ok) {
        // boundary rows (build_coeff_z :271-296): the first cell of the line sits in slot efirst of
        // chunk 0, the last one is the separator of the last chunk; bottom is applied first
        if (p == 0) {
#pragma unroll
            for (int e = 0; e < M; ++e)
                if (e == efirst) d[e] = a.set_first ? a.val_first : __dadd_rn(d[e], a.val_first);
        }
        if (p == P - 1) d[M - 1] = a.set_last ? a.val_last : __dadd_rn(d[M - 1], a.val_last);
    }
    double Yl;
    const double Y = tab_forward<M>(d, tab + a.g.o_f + 2 * cb, tab + a.g.o_alpha + cb, &Yl);
    double Sl;
    double S = cyl_reduced(a.g, tab, ex, NTH, p, d[M - 1], Y, Yl, [=](int q) { return ln * P + q; }, &Sl);
    if (ZM == 1) {
        if (line_ok) {
            const size_t line = (size_t)(L0 + ln);
            if (p == 0) a.y[line] = fma(tab_ld(tab + a.g.o_misc), S, Y);     // x_first = Y_0 + W_0*S_0 (ghosts at 0)
            if (p == P - 1) a.y[(size_t)a.nlines + line] = S;
        }
        return;
    }
    if (ZM == 2) {
        const size_t line = (size_t)min(L0 + ln, a.nlines - 1);
        const double Lg = a.ghost[line], Rg = a.ghost[(size_t)a.nlines + line];
        S = fma(tab_ld(tab + a.g.o_dr + p), Rg, fma(tab_ld(tab + a.g.o_dl + p), Lg, S));
        Sl = p > 0 ? fma(tab_ld(tab + a.g.o_dr + p - 1), Rg, fma(tab_ld(tab + a.g.o_dl + p - 1), Lg, Sl)) : Lg;
    }
    tab_backward<M>(d, tab + a.g.o_b + 2 * cb, Sl, S);
    if (VEC) {
#pragma unroll
        for (int j = 0; j < M / 2; ++j)
            *reinterpret_cast<double2 *>(myT + zswz<M>(p * M + 2 * j)) = make_double2(d[2 * j], d[2 * j + 1]);
    } else {
#pragma unroll
        for (int e = 0; e < M; ++e)
            if (e >= efirst) myT[zphys(i0 + e)] = d[e];
    }
    __syncthreads();
    for (int l = 0; l < nl; ++l) {
        const long long line = L0 + l;
        double *dst = a.out + (size_t)line * a.outer_stride;
        const double vv = (EPI && line / a.nphi == 0) ? a.T_inner : a.T_void;  // :61-68
        if (VEC) {
            for (int pl = tid; pl < (nz >> 1); pl += NTH) {
                double2 v = *reinterpret_cast<const double2 *>(sT + (size_t)l * RL + zswz<M>(2 * pl));
                if (EPI) {
                    const unsigned m2 = *reinterpret_cast<const unsigned short *>(sMask + (size_t)l * nz + 2 * pl);
                    if (!(m2 & 0xffu)) v.x = vv;
                    if (!(m2 >> 8)) v.y = vv;
                }
                *reinterpret_cast<double2 *>(dst + 2 * pl) = v;
            }
        } else {
            const uint8_t *act = EPI ? a.active + (size_t)line * a.outer_stride : nullptr;
            for (int z = tid; z < nz; z += NTH) {
                double v = sT[(size_t)l * RL + zphys(z)];
                if (EPI && !act[z]) v = vv;
                dst[z] = v;
            }
        }
    }
}

// Ghost values of every line from the gathered right-hand-side parts: the inter-segment system has
// line-independent coefficients, so its solution is a fixed linear map (G, 4*nranks doubles, host-built):
//   L = sum_q G[0][q][0]*yf_q + G[0][q][1]*yl_q,   R likewise with G[1].
__global__ void k_cyl_ghosts(const double *__restrict__ y_all, const double *__restrict__ G, double *__restrict__ ghost,
                             size_t nlines, int nranks)
{
    for (size_t l = (size_t)blockIdx.x * blockDim.x + threadIdx.x; l < nlines; l += (size_t)gridDim.x * blockDim.x) {
        double Lg = 0.0, Rg = 0.0;
        for (int q = 0; q < nranks; ++q) {
            const double yf = y_all[((size_t)q * 2) * nlines + l], yl = y_all[((size_t)q * 2 + 1) * nlines + l];
            Lg = fma(G[q * 2 + 1], yl, fma(G[q * 2], yf, Lg));
            Rg = fma(G[2 * nranks + q * 2 + 1], yl, fma(G[2 * nranks + q * 2], yf, Rg));
        }
        ghost[l] = Lg;
        ghost[nlines + l] = Rg;
    }
}

// ------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------
static bool same_key(const adi_cyl_params &x, const adi_cyl_params &y)
{
    // T_void / T_inner do not enter the tables
    return x.dt == y.dt && x.rho == y.rho && x.cp == y.cp && x.k == y.k && x.h_r == y.h_r &&
           x.Tinf_r == y.Tinf_r && x.kind_bot == y.kind_bot && x.kind_top == y.kind_top &&
           x.h_bot == y.h_bot && x.h_top == y.h_top && x.Tinf_bot == y.Tinf_bot &&
           x.Tinf_top == y.Tinf_top && x.T_bot == y.T_bot && x.T_top == y.T_top;
}

static int pick_M(adi_ctx *ctx, int n)
{
    if (ctx->opt_m == 16 && n <= 1024) return 16;
    if (ctx->opt_m == 32) return 32;
    return n <= 512 ? 16 : 32;
}

static int ensure_tables(adi_ctx *ctx, const adi_cyl_params &prm, cudaStream_t st)
{
    if (!ctx->cyl) ctx->cyl = new CylTables();
    CylTables &T = *ctx->cyl;
    const int nr = ctx->nr, nphi = ctx->nphi, nz = ctx->cnz;
    if (T.valid && T.nr == nr && T.nphi == nphi && T.nz == nz && T.M == (int)ctx->opt_m && same_key(T.key, prm)) return ADI_OK;
    if (prm.kind_bot < 0 || prm.kind_bot > 2) { set_error("unknown zbc.kind_bot"); return ADI_EINVAL; }
    if (prm.kind_top < 0 || prm.kind_top > 2) { set_error("unknown zbc.kind_top"); return ADI_EINVAL; }
    const double alpha = prm.k / (prm.rho * prm.cp);  // Material.alpha :49-50

    const bool phi = nphi > 1;
    TabSet sr, sz;
    {
        std::vector<double> a(nr), b(nr), c(nr);
        T.add_r = cyl_rows_r(nr, ctx->dr, alpha, prm.k, prm.dt, prm.h_r, prm.Tinf_r, a.data(), b.data(), c.data());
        sr = tab_make(nr, pick_M(ctx, nr), false, a.data(), b.data(), c.data());
    }
    std::vector<double> Gmap;
    {
        // z rows: the whole line, or this rank's segment of it with ghost couplings at the inner ends
        const int R = T.slab_nranks, me = T.slab_rank;
        auto seg = [&](int q, int nl, ZEnd *bot, ZEnd *top) {
            std::vector<double> a(nl), b(nl), c(nl);
            cyl_rows_z_segment(nl, q == 0, q == R - 1, ctx->dz, alpha, prm.k, prm.dt, prm.kind_bot, prm.kind_top, prm.h_bot,
                               prm.h_top, prm.Tinf_bot, prm.Tinf_top, prm.T_bot, prm.T_top, a.data(), b.data(), c.data(),
                               bot, top);
            return tab_make(nl, pick_M(ctx, nl), false, a.data(), b.data(), c.data(), q > 0, q < R - 1);
        };
        sz = seg(me, nz, &T.bot, &T.top);
        if (R > 1) {
            // line-independent coefficients of every rank's interface relation, then the ghost map of this
            // rank: the inter-rank solve applied to unit right-hand sides
            std::vector<Iface> cst(R);
            for (int q = 0; q < R; ++q) {
                ZEnd b0, t0;
                const TabSet tq = q == me ? sz : seg(q, T.slab_nz[q], &b0, &t0);
                const double *bl = tq.blob.data();
                cst[q].yf = cst[q].yl = 0.0;
                cst[q].vf = bl[tq.g.o_misc + 1] + bl[tq.g.o_misc] * bl[tq.g.o_dl];
                cst[q].wf = bl[tq.g.o_misc] * bl[tq.g.o_dr];
                cst[q].vl = bl[tq.g.o_dl + tq.g.P - 1];
                cst[q].wl = bl[tq.g.o_dr + tq.g.P - 1];
            }
            Gmap.assign(4 * R, 0.0);
            for (int q = 0; q < R; ++q)
                for (int w = 0; w < 2; ++w) {
                    std::vector<Iface> rel(cst);
                    (w == 0 ? rel[q].yf : rel[q].yl) = 1.0;
                    double Lg, Rg;
                    iface_solve([&](int r) { return rel[r]; }, R, me, &Lg, &Rg);
                    Gmap[q * 2 + w] = Lg;
                    Gmap[2 * R + q * 2 + w] = Rg;
                }
        }
    }
    T.gr = sr.g; T.gz = sz.g;
    if (phi) T.gp = tab_geom(nphi, pick_M(ctx, nphi), true, true);
    else memset(&T.gp, 0, sizeof(T.gp));
    if (T.gr.P > 64 || T.gz.P > 64 || (phi && T.gp.P > 64)) {
        set_error("adi_cyl_step: line too long for the register-resident sweep (n > 2048)");
        return ADI_EINVAL;
    }
    // blobs are padded to 16 bytes apart so that every table starts 8-byte aligned anyway
    T.off_r = 0;
    T.off_z = T.off_r + T.gr.ndbl;
    T.off_p = T.off_z + T.gz.ndbl;
    const size_t ndbl = T.off_p + (phi ? (size_t)nr * T.gp.ndbl : 0);
    T.goff_r = 0;
    T.goff_z = T.goff_r + 3 * T.gr.P;
    T.goff_p = T.goff_z + 3 * T.gz.P;
    const size_t nint = T.goff_p + (phi ? 3 * T.gp.P : 0);

    std::vector<double> blob(ndbl);
    std::vector<int> geom(nint);
    std::copy(sr.blob.begin(), sr.blob.end(), blob.begin() + T.off_r);
    std::copy(sz.blob.begin(), sz.blob.end(), blob.begin() + T.off_z);
    std::copy(sr.geom.begin(), sr.geom.end(), geom.begin() + T.goff_r);
    std::copy(sz.geom.begin(), sz.geom.end(), geom.begin() + T.goff_z);
    if (phi) {
        std::vector<double> a(nphi), b(nphi), c(nphi);
        int *gi = geom.data() + T.goff_p;
        tab_partition(T.gp, true, gi, gi + T.gp.P, gi + 2 * T.gp.P);
        for (int ir = 0; ir < nr; ++ir) {
            const double f = cyl_fac_phi(ir, ctx->dr, ctx->dphi, alpha, prm.dt);
            for (int j = 0; j < nphi; ++j) { a[j] = -f; b[j] = 1.0 + 2.0 * f; c[j] = -f; }
            tab_build(T.gp, gi, gi + T.gp.P, gi + 2 * T.gp.P, a.data(), b.data(), c.data(),
                      blob.data() + T.off_p + (size_t)ir * T.gp.ndbl);
        }
    }
    // the previous tables may still be in use by kernels queued on `st`: stream-ordered frees
    if (T.blob_cap < ndbl) {
        if (T.d_blob) { ADI_CUDA(cudaStreamSynchronize(st)); ADI_CUDA(cudaFree(T.d_blob)); }
        T.d_blob = nullptr;
        ADI_CUDA(cudaMalloc(&T.d_blob, ndbl * sizeof(double)));
        T.blob_cap = ndbl;
    }
    if (T.geom_cap < nint) {
        if (T.d_geom) { ADI_CUDA(cudaStreamSynchronize(st)); ADI_CUDA(cudaFree(T.d_geom)); }
        T.d_geom = nullptr;
        ADI_CUDA(cudaMalloc(&T.d_geom, nint * sizeof(int)));
        T.geom_cap = nint;
    }
    // pageable sources: the copies are staged before the calls return
    ADI_CUDA(cudaMemcpyAsync(T.d_blob, blob.data(), ndbl * sizeof(double), cudaMemcpyHostToDevice, st));
    ADI_CUDA(cudaMemcpyAsync(T.d_geom, geom.data(), nint * sizeof(int), cudaMemcpyHostToDevice, st));
    if (!Gmap.empty()) {
        if (T.d_G) { ADI_CUDA(cudaStreamSynchronize(st)); ADI_CUDA(cudaFree(T.d_G)); }
        T.d_G = nullptr;
        ADI_CUDA(cudaMalloc(&T.d_G, Gmap.size() * sizeof(double)));
        ADI_CUDA(cudaMemcpyAsync(T.d_G, Gmap.data(), Gmap.size() * sizeof(double), cudaMemcpyHostToDevice, st));
    }
    {
        // The z rows (build_coeff_z :255-298, theta = 1) are the rows of a Cartesian z sweep of a full line with
        // g = fac and a Robin term on the exposed end cells: b = (1 + fac) + c, d = T + c*T_inf with c = fac*(h/k)*dz
        // (neumann0: c = 0) -- adi_core.h make_row, CMODE 1 with dt = 1.  Dirichlet ends and two different ambient
        // temperatures do not fit that form (k_cyl_z keeps them).
        const double fac = 1.0 * alpha * prm.dt / (ctx->dz * ctx->dz);
        const bool rb = prm.kind_bot == 2, rt = prm.kind_top == 2;
        T.zt_ok = T.slab_nranks == 1 && prm.kind_bot != 1 && prm.kind_top != 1 && nz >= 2 &&
                  !(rb && rt && prm.Tinf_bot != prm.Tinf_top);
        memset(&T.zk, 0, sizeof(T.zk));
        T.zk.g = fac; T.zk.dt = 1.0;
        T.zk.Tinf = rb ? prm.Tinf_bot : prm.Tinf_top;
        T.zk.h_lo = rb ? fac * (prm.h_bot / prm.k) * ctx->dz : 0.0;
        T.zk.h_hi = rt ? fac * (prm.h_top / prm.k) * ctx->dz : 0.0;
        std::vector<uint8_t> zc((size_t)nz + 16, 0);
        for (int kz = 0; kz < nz; ++kz)
            zc[kz] = (uint8_t)(CB_SELF | CB_XM | CB_XP | CB_YM | CB_YP | (kz > 0 ? CB_ZM : 0u) | (kz + 1 < nz ? CB_ZP : 0u));
        if (T.zcode_cap < zc.size()) {
            if (T.d_zcode) { ADI_CUDA(cudaStreamSynchronize(st)); ADI_CUDA(cudaFree(T.d_zcode)); }
            T.d_zcode = nullptr;
            ADI_CUDA(cudaMalloc(&T.d_zcode, zc.size()));
            T.zcode_cap = zc.size();
        }
        ADI_CUDA(cudaMemcpyAsync(T.d_zcode, zc.data(), zc.size(), cudaMemcpyHostToDevice, st));
        ADI_CUDA(cudaStreamSynchronize(st));
    }
    ADI_CUDA(cudaStreamSynchronize(st));
    T.M = (int)ctx->opt_m;
    T.key = prm; T.nr = nr; T.nphi = nphi; T.nz = nz; T.valid = true;
    return ADI_OK;
}

template <typename K, typename... Extra>
static int launch_cyl(K kern, dim3 grid, dim3 block, size_t smem, cudaStream_t st, adi_ctx *ctx, const CylArgs &a,
                      Extra... extra)
{
    if (smem > 227 * 1024) {
        set_error("adi_cyl_step: tile does not fit shared memory");
        return ADI_EINVAL;
    }
    if (smem > 48 * 1024)
        ADI_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, block, smem, st>>>(a, extra...);
    ctx->launches++;
    ADI_CUDA(cudaGetLastError());
    return ADI_OK;
}

static int lanes_for(adi_ctx *ctx, int P, int nz)
{
    // With the tables in shared memory wide rows pay: 16 lanes (128-byte rows) in 256-thread blocks, 32 lanes for
    // <= 8 chunks (r sweep of 256 cells: 0.474 ms with 8 lanes, 0.381 with 16; r03d).  Tables through L1 (cylsm=0):
    // 8 lanes in blocks of at least 128 threads, small blocks keeping more tiles in flight (0.59 -> 0.51 ms then).
    int KT = 32;
    if (ctx->opt_cylsm) {
        if (P > 8) KT = 16;
    } else {
        while (KT > 8 && KT * P > 128) KT >>= 1;
    }
    while (KT > 1 && KT * P > 256) KT >>= 1;
    if (ctx->opt_kt > 0) {
        int w = 1;
        while (2 * w <= ctx->opt_kt && 2 * w <= 32 && 2 * w * P <= 256) w <<= 1;
        KT = w;
    }
    while (KT > 1 && KT / 2 >= nz) KT >>= 1;
    return KT;
}

static int launch_strided(adi_ctx *ctx, CylArgs &a, bool pro, int nouter, cudaStream_t st)
{
    const int P = a.g.P;
    const int KT = lanes_for(ctx, P, a.nz);
    dim3 block(KT, P), grid((a.nz + KT - 1) / KT, nouter);
    const size_t nth = (size_t)KT * P;
    const size_t smem = 3 * nth * sizeof(double);
    const bool bcl = a.set_last != 0 || a.val_last != 0.0;
    if ((unsigned long long)a.cell_stride * 8ull >= (1ull << 32)) {
        set_error("adi_cyl_step: grid too large for 32-bit line strides");
        return ADI_EINVAL;
    }
    // tables in shared memory, `nzt` z tiles per block (option cylsm, default on; 0 = tables through L1, one tile)
    const bool sm = ctx->opt_cylsm != 0 && (a.g.o_dl % 2) == 0 && (nth % 2) == 0 && (size_t)a.g.o_dl * sizeof(double) <= 64 * 1024;
    const int ntz = (int)grid.x;
    int nzt = 1;
    if (sm) {
        nzt = ctx->opt_cylsm > 1 ? (int)ctx->opt_cylsm : 4;
        while (nzt > 1 && (long long)((ntz + nzt - 1) / nzt) * nouter < 148 * 16) nzt >>= 1;   // keep the GPU full
        grid.x = (unsigned)((ntz + nzt - 1) / nzt);
    }
    const size_t smem2 = smem + (sm ? (size_t)a.g.o_dl * sizeof(double) : 0);
#define ADI_SGO2(MM, PR, BC)                                                                                   \
    {                                                                                                          \
        if (sm) return launch_cyl(k_cyl_strided<MM, PR, BC, true>, grid, block, smem2, st, ctx, a, nzt);      \
        return launch_cyl(k_cyl_strided<MM, PR, BC, false>, grid, block, smem2, st, ctx, a, nzt);             \
    }
#define ADI_SGO(MM)                                \
    {                                              \
        if (pro) {                                 \
            if (bcl) ADI_SGO2(MM, true, true)      \
            ADI_SGO2(MM, true, false)              \
        }                                          \
        if (bcl) ADI_SGO2(MM, false, true)         \
        ADI_SGO2(MM, false, false)                 \
    }
    if (a.g.M == 16) ADI_SGO(16)
    ADI_SGO(32)
#undef ADI_SGO
#undef ADI_SGO2
}

static int launch_z(adi_ctx *ctx, CylArgs &a, bool epi, cudaStream_t st, int zm = 0)
{
    const int P = a.g.P, M = a.g.M;
    // 16-byte path: every chunk full and aligned
    const bool vec = (a.nz % M) == 0 && (a.outer_stride % 2) == 0 &&
                     ((((uintptr_t)a.in | (uintptr_t)a.out) & 15) == 0) &&
                     (!epi || ((a.outer_stride % 16) == 0 && ((uintptr_t)a.active & 15) == 0));
    const int RL = vec ? a.nz : a.nz - 1 + ((a.nz - 1) >> 4) + 1;
    auto bytes = [&](int lt) {
        return ((size_t)3 * lt * P + (size_t)lt * RL) * sizeof(double) + ((vec && epi) ? (size_t)lt * a.nz : 0);
    };
    // small blocks: many resident tiles per SM keep loads, solves and stores of different tiles
    // overlapped
    int LT = 32;
    while (LT > 1 && LT * P > 64) LT >>= 1;
    if (ctx->opt_lt > 0) {
        LT = 1;
        while (2 * LT <= ctx->opt_lt && 2 * LT * P <= 256) LT <<= 1;
    }
    while (LT > 1 && bytes(LT) > 100 * 1024) LT >>= 1;
    while (LT > 1 && (long long)(LT / 2) >= a.nlines) LT >>= 1;
    dim3 block(P, LT), grid((unsigned)((a.nlines + LT - 1) / LT));
    const size_t smem = bytes(LT);
#define ADI_ZGO1(MM, ZMV)                                                                             \
    {                                                                                                 \
        if (vec) {                                                                                    \
            if (epi) return launch_cyl(k_cyl_z<MM, true, true, ZMV>, grid, block, smem, st, ctx, a);  \
            return launch_cyl(k_cyl_z<MM, false, true, ZMV>, grid, block, smem, st, ctx, a);          \
        }                                                                                             \
        if (epi) return launch_cyl(k_cyl_z<MM, true, false, ZMV>, grid, block, smem, st, ctx, a);     \
        return launch_cyl(k_cyl_z<MM, false, false, ZMV>, grid, block, smem, st, ctx, a);             \
    }
#define ADI_ZGO(MM)                     \
    {                                   \
        if (zm == 1) ADI_ZGO1(MM, 1)    \
        if (zm == 2) ADI_ZGO1(MM, 2)    \
        ADI_ZGO1(MM, 0)                 \
    }
    if (M == 16) ADI_ZGO(16)
    ADI_ZGO(32)
#undef ADI_ZGO1
#undef ADI_ZGO
}

static int ensure_aux(adi_ctx *ctx, size_t cells)
{
    if (ctx->stage_aux_cells >= cells && ctx->stage_mask) return ADI_OK;
    if (ctx->stage_mask) cudaFree(ctx->stage_mask);
    if (ctx->stage_src) cudaFree(ctx->stage_src);
    ctx->stage_mask = nullptr; ctx->stage_src = nullptr;
    ADI_CUDA(cudaMalloc(&ctx->stage_mask, std::max<size_t>(cells, 1)));
    ADI_CUDA(cudaMalloc(&ctx->stage_src, std::max<size_t>(cells, 1) * sizeof(double)));
    ctx->stage_aux_cells = cells;
    return ADI_OK;
}

}  // namespace adi

using namespace adi;

extern "C" {

int adi_cyl_bind(adi_ctx *ctx, int nr, int nphi, int nz, int nz_pitch, double dr, double dphi, double dz)
{
    if (!ctx) return ADI_EINVAL;
    if (nr < 1 || nphi < 1 || nz < 1 || nz_pitch < nz || !(dr > 0.0) || !(dz > 0.0) || !(dphi > 0.0)) {
        set_error("adi_cyl_bind: bad grid");
        return ADI_EINVAL;
    }
    ADI_CUDA(cudaSetDevice(ctx->device));
    ctx->nr = nr; ctx->nphi = nphi; ctx->cnz = nz; ctx->nz_pitch = nz_pitch;
    ctx->dr = dr; ctx->dphi = dphi; ctx->dz = dz;
    ctx->cyl_bound = true;
    if (ctx->cyl) { ctx->cyl->valid = false; ctx->cyl->slab_rank = 0; ctx->cyl->slab_nranks = 1; ctx->cyl->slab_nz.clear(); }
    return ADI_OK;
}

// phases: bit 0 = r and phi sweeps (Tin -> Tout), bit 1 = z sweep on Tout with mode zm
// (0 whole lines, 1 z-slab pass 1 -> d_y, 2 z-slab pass 2 with d_y_all)
static int cyl_run(adi_ctx *ctx, const double *d_Tin, double *d_Tout, const adi_cyl_params *p, const uint8_t *d_active,
                   const double *d_S, int phases, int zm, double *d_y, const double *d_y_all, cudaStream_t st)
{
    int rc = ensure_tables(ctx, *p, st);
    if (rc) return rc;
    CylTables &T = *ctx->cyl;
    const int nr = ctx->nr, nphi = ctx->nphi, nz = ctx->cnz;
    const long long pitch = ctx->nz_pitch;

    CylArgs a;
    memset(&a, 0, sizeof(a));
    a.nz = nz;
    a.nphi = nphi;
    a.T_void = p->T_void; a.T_inner = p->T_inner;
    if (phases & 1) {
        rc = prof_mark(ctx, 0, st);
        if (rc) return rc;
        rc = prof_mark(ctx, 1, st);
        if (rc) return rc;

        // r sweep  (:341-344): Tin -> Tout, prologue fused
        a.in = d_Tin; a.out = d_Tout;
        a.blob = T.d_blob + T.off_r; a.geom = T.d_geom + T.goff_r; a.blob_stride = 0; a.g = T.gr;
        a.cell_stride = (long long)nphi * pitch; a.outer_stride = pitch;
        a.active = d_active; a.S = d_S; a.dt = p->dt; a.rho_cp = p->rho * p->cp;
        a.set_first = 0; a.val_first = 0.0;
        a.set_last = 0; a.val_last = (p->h_r != 0.0) ? T.add_r : 0.0;
        rc = launch_strided(ctx, a, d_active != nullptr || d_S != nullptr, nphi, st);
        if (rc) return rc;
        rc = prof_mark(ctx, 2, st);
        if (rc) return rc;

        // phi sweep (:346), in place; nphi == 1 is the identity (:309-310)
        a.in = d_Tout; a.active = nullptr; a.S = nullptr;
        a.set_first = a.set_last = 0; a.val_first = a.val_last = 0.0;
        if (nphi > 1) {
            a.blob = T.d_blob + T.off_p; a.geom = T.d_geom + T.goff_p; a.blob_stride = T.gp.ndbl; a.g = T.gp;
            a.cell_stride = pitch; a.outer_stride = (long long)nphi * pitch;
            rc = launch_strided(ctx, a, false, nr, st);
            if (rc) return rc;
        }
        rc = prof_mark(ctx, 3, st);
        if (rc) return rc;
    }
    if (phases & 2) {
        // z sweep (:348-350), in place, epilogue fused
        const size_t nlines = (size_t)nr * nphi;
        a.in = d_Tout; a.out = d_Tout;
        a.blob = T.d_blob + T.off_z; a.geom = T.d_geom + T.goff_z; a.blob_stride = 0; a.g = T.gz;
        a.cell_stride = 0; a.outer_stride = pitch; a.nlines = (long long)nlines;
        a.active = zm == 1 ? nullptr : d_active;
        a.S = nullptr;
        a.set_first = T.bot.set; a.val_first = T.bot.val;
        a.set_last = T.top.set; a.val_last = T.top.val;
        a.y = d_y;
        if (zm == 2) {
            if (T.ghost_lines < nlines) {
                if (T.d_ghost) { ADI_CUDA(cudaStreamSynchronize(st)); ADI_CUDA(cudaFree(T.d_ghost)); }
                T.d_ghost = nullptr;
                ADI_CUDA(cudaMalloc(&T.d_ghost, 2 * nlines * sizeof(double)));
                T.ghost_lines = nlines;
            }
            const int threads = 128;
            const int blocks = (int)std::min<size_t>((nlines + threads - 1) / threads, 148 * 32);
            k_cyl_ghosts<<<blocks, threads, 0, st>>>(d_y_all, T.d_G, T.d_ghost, nlines, T.slab_nranks);
            ctx->launches++;
            ADI_CUDA(cudaGetLastError());
            a.ghost = T.d_ghost;
        }
        int used = 0;
        if (zm == 0 && !d_active && T.zt_ok && ctx->opt_cylzt && nz >= 64 && nr <= 0x7fff && (size_t)nr * nphi <= 0x7fffffffull) {
            // whole lines, no mask: the Cartesian z kernel (bulk asynchronous copies, tabulated uniform runs from the
            // constant bank) -- 0.48 -> 0.35 ms at 256 x 1024 x 512
            SweepArgs s = {};
            s.in = d_Tout; s.out = d_Tout; s.code = T.d_zcode;
            s.nx = nr; s.ny = nphi; s.nz = nz;
            s.k = T.zk;
            s.line_batch = 1;                               // no tile lists (they belong to the Cartesian grid)
            s.zpitch = pitch != nz ? (int)pitch : 0;
            s.code_line = 1;
            rc = launch_sweep_zt(ctx, s, false, false, 0, st, &used);
            if (rc) return rc;
        }
        if (!used) {
            rc = launch_z(ctx, a, zm != 1 && d_active != nullptr, st, zm);
            if (rc) return rc;
        }
        if (zm != 1) {
            rc = prof_mark(ctx, 4, st);
            if (rc) return rc;
        }
    }
    if (ctx->opt_sync_check) ADI_CUDA(cudaStreamSynchronize(st));
    return ADI_OK;
}

static int cyl_check(adi_ctx *ctx, const adi_cyl_params *p, const char *who)
{
    if (!ctx || !p) { set_error(std::string(who) + ": NULL argument"); return ADI_EINVAL; }
    if (!ctx->cyl_bound) { set_error(std::string(who) + ": adi_cyl_bind has not been called"); return ADI_ESTATE; }
    ADI_CUDA(cudaSetDevice(ctx->device));
    return ADI_OK;
}

int adi_cyl_step(adi_ctx *ctx, const double *d_Tin, double *d_Tout, const adi_cyl_params *p,
                 const uint8_t *d_active, const double *d_S, void *stream)
{
    int rc = cyl_check(ctx, p, "adi_cyl_step");
    if (rc) return rc;
    if (!d_Tin || !d_Tout || d_Tin == d_Tout) {
        set_error("adi_cyl_step: Tin/Tout must be distinct device arrays");
        return ADI_EINVAL;
    }
    if (ctx->cyl && ctx->cyl->slab_nranks > 1) {
        set_error("adi_cyl_step: this context holds a z slab; use adi_cyl_step_rphi + adi_cyl_zsweep_*");
        return ADI_ESTATE;
    }
    return cyl_run(ctx, d_Tin, d_Tout, p, d_active, d_S, 3, 0, nullptr, nullptr, (cudaStream_t)stream);
}

int adi_cyl_set_slab(adi_ctx *ctx, int rank, int nranks, const int *nz_per_rank)
{
    if (!ctx || !ctx->cyl_bound) { set_error("adi_cyl_set_slab: adi_cyl_bind has not been called"); return ADI_ESTATE; }
    if (nranks < 1 || nranks > 16 || rank < 0 || rank >= nranks || (nranks > 1 && !nz_per_rank)) {
        set_error("adi_cyl_set_slab: need 0 <= rank < nranks <= 16 and the local nz of every rank");
        return ADI_EINVAL;
    }
    if (nranks > 1 && nz_per_rank[rank] != ctx->cnz) {
        set_error("adi_cyl_set_slab: nz_per_rank[rank] differs from the bound local nz");
        return ADI_EINVAL;
    }
    if (!ctx->cyl) ctx->cyl = new CylTables();
    CylTables &T = *ctx->cyl;
    T.slab_rank = rank; T.slab_nranks = nranks;
    T.slab_nz.assign(nranks, ctx->cnz);
    for (int q = 0; q < nranks && nz_per_rank; ++q) {
        if (nz_per_rank[q] < 1) { set_error("adi_cyl_set_slab: every rank needs at least one z plane"); return ADI_EINVAL; }
        T.slab_nz[q] = nz_per_rank[q];
    }
    T.valid = false;
    return ADI_OK;
}

int adi_cyl_step_rphi(adi_ctx *ctx, const double *d_Tin, double *d_Tout, const adi_cyl_params *p,
                      const uint8_t *d_active, const double *d_S, void *stream)
{
    int rc = cyl_check(ctx, p, "adi_cyl_step_rphi");
    if (rc) return rc;
    if (!d_Tin || !d_Tout || d_Tin == d_Tout) {
        set_error("adi_cyl_step_rphi: Tin/Tout must be distinct device arrays");
        return ADI_EINVAL;
    }
    return cyl_run(ctx, d_Tin, d_Tout, p, d_active, d_S, 1, 0, nullptr, nullptr, (cudaStream_t)stream);
}

int adi_cyl_zsweep_reduce(adi_ctx *ctx, double *d_T, const adi_cyl_params *p, double *d_y, void *stream)
{
    int rc = cyl_check(ctx, p, "adi_cyl_zsweep_reduce");
    if (rc) return rc;
    if (!d_T || !d_y) { set_error("adi_cyl_zsweep_reduce: NULL argument"); return ADI_EINVAL; }
    return cyl_run(ctx, d_T, d_T, p, nullptr, nullptr, 2, 1, d_y, nullptr, (cudaStream_t)stream);
}

int adi_cyl_zsweep_finish(adi_ctx *ctx, double *d_T, const adi_cyl_params *p, const double *d_y_all,
                          const uint8_t *d_active, void *stream)
{
    int rc = cyl_check(ctx, p, "adi_cyl_zsweep_finish");
    if (rc) return rc;
    if (!d_T || !d_y_all) { set_error("adi_cyl_zsweep_finish: NULL argument"); return ADI_EINVAL; }
    if (!ctx->cyl || ctx->cyl->slab_nranks < 2) { set_error("adi_cyl_zsweep_finish: adi_cyl_set_slab has not been called"); return ADI_ESTATE; }
    return cyl_run(ctx, d_T, d_T, p, d_active, nullptr, 2, 2, nullptr, d_y_all, (cudaStream_t)stream);
}

int adi_cyl_step_host(adi_ctx *ctx, const double *h_Tin, double *h_Tout, int nsteps,
                      const adi_cyl_params *p, const uint8_t *h_active, const double *h_S, void *stream)
{
    if (!ctx || !p || !h_Tin || !h_Tout || nsteps < 1) { set_error("adi_cyl_step_host: bad arguments"); return ADI_EINVAL; }
    if (!ctx->cyl_bound) { set_error("adi_cyl_step_host: adi_cyl_bind has not been called"); return ADI_ESTATE; }
    if (ctx->nz_pitch != ctx->cnz) { set_error("adi_cyl_step_host: host arrays are dense (nz_pitch must equal nz)"); return ADI_EINVAL; }
    cudaStream_t st = (cudaStream_t)stream;
    const size_t ncell = (size_t)ctx->nr * ctx->nphi * ctx->cnz;
    if (ctx->stage_cells < ncell || !ctx->stage[0]) {
        for (int i = 0; i < 2; ++i) {
            if (ctx->stage[i]) cudaFree(ctx->stage[i]);
            ctx->stage[i] = nullptr;
            ADI_CUDA(cudaMalloc(&ctx->stage[i], ncell * sizeof(double)));
        }
        ctx->stage_cells = ncell;
    }
    int rc = ensure_aux(ctx, ncell);
    if (rc) return rc;
    if ((rc = stage_h2d(ctx, ctx->stage[0], h_Tin, ncell * sizeof(double), st))) return rc;
    if (h_active && (rc = stage_h2d(ctx, ctx->stage_mask, h_active, ncell, st))) return rc;
    if (h_S && (rc = stage_h2d(ctx, ctx->stage_src, h_S, ncell * sizeof(double), st))) return rc;
    int cur = 0;
    for (int s = 0; s < nsteps; ++s) {
        rc = adi_cyl_step(ctx, ctx->stage[cur], ctx->stage[cur ^ 1], p, h_active ? ctx->stage_mask : nullptr,
                          h_S ? ctx->stage_src : nullptr, stream);
        if (rc) return rc;
        cur ^= 1;
    }
    return stage_d2h(ctx, h_Tout, ctx->stage[cur], ncell * sizeof(double), st);
}

}  // extern "C"
