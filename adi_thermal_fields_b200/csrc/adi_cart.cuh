// adi_cart.cuh -- sm_100a kernels of the Cartesian ADI step
// (adi3d_numba_coeff.py:290-302 / adi3d_gpu_coeff.py:213-230).
//
// K0 k_build_code      mask (+dir_mask) -> 1-byte neighbour code per cell
// K1 k_sweep_strided   x sweep (stride ny*nz, explicit stage fused) and y sweep (stride nz):
//                      lanes run along z, so every load/store of a warp is a run of
//                      contiguous 8-byte cells (64-256 B rows); each thread keeps an
//                      M-cell chunk of its line in registers and the pivot reciprocals in a
//                      conflict-free shared-memory column
// K3 k_sweep_z         z sweep (contiguous axis): the tile of lines is staged through
//                      XOR-swizzled shared memory with 16-byte cp.async copies, then the same
//                      register-resident chunk solve runs with lanes along the line
// K7 k_build_packs     precompute_coeff_packs_unified on the device
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>

#include <type_traits>

#include "adi_core.h"
#include "adi_mask_core.h"

namespace adi {

struct SweepArgs {
    const double *in;                  // may alias out (y and z sweeps run in place)
    double *out;
    const uint8_t *__restrict__ code;
    const double *__restrict__ coeff;  // CMODE 2
    int sparse;                        // CMODE 2, x / y sweeps: coeff is +0.0 wherever the cell has both neighbours
                                       // along the sweep axis (verified by k_check_sparse): read it at exposed cells only
    const double *__restrict__ q;      // EXTRA, may be null
    const double *__restrict__ dirv;   // EXTRA, may be null
    int nx, ny, nz;
    SweepConst k;
    // z-slab decomposition (all NULL / 0 on a single GPU)
    const double *__restrict__ zlo;    // T plane below this slab (nx*ny), explicit stage
    const double *__restrict__ zhi;    // T plane above this slab
    double *iface_dyn;                 // z sweep pass 1 out: [2][nx*ny] (yf, yl): right-hand-side part
    double *iface_stat;                // z sweep pass 1 out: [4][nx*ny] (vf, wf, vl, wl): matrix part (ZMODE 1)
    const double *__restrict__ ghost;  // z sweep pass 2 in:  [2][nx*ny] (L, R per line)
    // second-generation x / y sweeps (adi_sweep_xy.cuh)
    const uint8_t *__restrict__ codeT; // code transposed for this axis: [other][nz][npad], line axis fastest
    int npad;                          // padded line length of codeT (multiple of 32, padding = code 0)
    int uni;                           // uniform chunks may take the tabulated factors of `uc`
    int tw;                            // reduced system by warps (transposed exchange)
    int remap;                         // chunk order inside a block: ends of the line in the same warp
    int dbg;                           // tuning aid: 1 = loads and stores only
    int halo_defer;                    // explicit stage: the adjacent slabs' T planes are still travelling -- cells that need
                                       // them are left to k_explicit_faces (not stored here)
    // active-tile list (parts under construction: most of the box is void): block b works on tile tiles[b] =
    // by * tiles_nx + bx instead of (blockIdx.x, blockIdx.y); NULL = every tile, addressed by blockIdx
    const int *__restrict__ tiles;
    int tiles_nx;
    int tsplit;                        // k_sweep_xy: 1 = the list counts tiles of twice the block's lanes (two blocks per entry)
    int line_batch;                    // z sweep of a batch of lines (multi-GPU): the pointers are offset, no tile list
    int zpitch;                        // k_sweep_zt: elements between consecutive z lines of in / out (0: nz)
    int code_line;                     // k_sweep_zt: 1 = `code` holds ONE line of nz codes shared by all z lines
    UniConst uc;
    int line0;                         // x sweep over the planes line0 .. line0+nx-1 of the grid only (the x extent of a part
                                       // under construction; the field / operand pointers are already offset): offset of
                                       // the transposed code rows
    int zfull;                         // k_sweep_zt: elements between consecutive z lines of code / coeff / q / dirv (0: nz);
                                       // set when the sweep solves only the first nz cells of longer lines (cells above
                                       // the top of a part under construction are void: launch_sweep_zt)
};

#ifdef ADI_CART_MISC_KERNELS  // defined by adi_cart.cu, the one unit that launches K0/K7
// ------------------------------------------------------------------------------------
// K0: neighbour code.  One thread per cell; the six neighbour bytes come from L1/L2.
// ------------------------------------------------------------------------------------
__global__ void k_build_code(const uint8_t *__restrict__ mask, const uint8_t *__restrict__ dirm,
                             uint8_t *__restrict__ code, int nx, int ny, int nz,
                             const uint8_t *__restrict__ mlo, const uint8_t *__restrict__ mhi)
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
        if (mask[idx]) {  // a void cell keeps code 0
            c = CB_SELF;
            if (dirm && dirm[idx]) c |= CB_DIR;
            if (i > 0 && mask[idx - snx]) c |= CB_XM;
            if (i + 1 < nx && mask[idx + snx]) c |= CB_XP;
            if (j > 0 && mask[idx - nz]) c |= CB_YM;
            if (j + 1 < ny && mask[idx + nz]) c |= CB_YP;
            // across a slab boundary the neighbour is the adjacent rank's mask plane
            if (k > 0 ? mask[idx - 1] : (mlo && mlo[ij])) c |= CB_ZM;
            if (k + 1 < nz ? mask[idx + 1] : (mhi && mhi[ij])) c |= CB_ZP;
        }
        code[idx] = (uint8_t)c;
    }
}

// K0 in word form (adi_mask_core.h): 16 cells of a z line per thread, 16-byte loads and stores; runs of void
// cells (most of the box while a part is being built) cost one load and one store.
__global__ void __launch_bounds__(256) k_build_code_v(const uint8_t *__restrict__ mask, const uint8_t *__restrict__ dirm,
                                                      uint8_t *__restrict__ code, int nx, int ny, int nz,
                                                      const uint8_t *__restrict__ mlo, const uint8_t *__restrict__ mhi,
                                                      int *__restrict__ ztop)
{
    const size_t n16 = (size_t)nx * ny * nz / 16;
    int top = 0, xlo = 0x7fffffff, xhi = -1;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < n16; t += (size_t)gridDim.x * blockDim.x) {
        int xi = -1;
        top = max(top, build_code16(mask, dirm, code, t * 16, nx, ny, nz, mlo, mhi, &xi));
        if (xi >= 0) { xlo = min(xlo, xi); xhi = max(xhi, xi); }
    }
    if (ztop) {   // ztop[0] = z + 1 of the highest active cell of the grid; ztop[1] = nx - lowest, ztop[2] = highest + 1
                  // x plane with an active cell (all three start at 0 and grow by atomicMax)
        top = __reduce_max_sync(0xffffffffu, top);
        xlo = __reduce_min_sync(0xffffffffu, xlo);
        xhi = __reduce_max_sync(0xffffffffu, xhi);
        if ((threadIdx.x & 31) == 0 && top > 0) {
            atomicMax(ztop, top);
            atomicMax(ztop + 1, nx - xlo);
            atomicMax(ztop + 2, xhi + 1);
        }
    }
}

// K0t: per-axis transposed copy of the code array for the x / y sweeps (adi_sweep_xy.cuh):
// dst[(b*nz + c)*npad + r] = src[b*sb + r*sr + c],  r < n (line axis), c < nz, b < batch.
// 32 x 32 byte tiles through shared memory; padding columns (r >= n) keep the zeros of the allocation.
__global__ void __launch_bounds__(256) k_transpose_code(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst,
                                                        int n, int nz, int npad, int batch, size_t sb, size_t sr)
{
    __shared__ uint8_t tile[32][33];
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    const int tx = threadIdx.x, ty = threadIdx.y;
    for (int b = blockIdx.z; b < batch; b += gridDim.z) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = r0 + ty + 8 * i, c = c0 + tx;
            tile[ty + 8 * i][tx] = (r < n && c < nz) ? src[(size_t)b * sb + (size_t)r * sr + c] : (uint8_t)0;
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int c = c0 + ty + 8 * i, r = r0 + tx;
            if (c < nz && r < n) dst[((size_t)b * nz + c) * npad + r] = tile[tx][ty + 8 * i];
        }
        __syncthreads();
    }
}

// K0t in word form (adi_mask_core.h): 128 x 128 byte tiles, 128-byte rows both ways.
__global__ void __launch_bounds__(256) k_transpose_code_v(const TrArgs a, int batch)
{
    __shared__ uint32_t S[128 * 32];
    const int c0 = blockIdx.x * 128, r0 = blockIdx.y * 128;
    for (int b = blockIdx.z; b < batch; b += gridDim.z) {
        tr_load(a, S, threadIdx.x, c0, r0, b);
        __syncthreads();
        tr_store(a, S, threadIdx.x, c0, r0, b);
        __syncthreads();
    }
}

// K0f: which sweep tiles hold an active cell?  The rows of `base` (unit bytes each, contiguous) are grouped into
// tiles of KT consecutive rows inside runs of `inner` rows (x / y sweeps: rows = (other index, z), runs = one `other`
// index, base = the transposed code array; z sweep: rows = z lines, one run).  flags[t] = 1 when any byte of tile
// t = by * ceil(inner / KT) + bx is non-zero.  One block per tile.
// uni_lo / uni_hi (x / y sweeps, unit a multiple of 32 bytes, 16-byte aligned rows; 0 = not asked): bit 1 of the flag
// is set when the tile is ALL UNIFORM -- its KT rows all exist and every 32-byte chunk of every row passes
// chunk_uniform<32, 1> (cells 1..30 active with both neighbours along the axis and no Dirichlet bit, cell 0 active
// and coupled to cell 1): such a tile can go to k_sweep_xyu, which has no general row assembly at all.
__global__ void __launch_bounds__(128) k_tile_flags(const uint8_t *__restrict__ base, size_t unit, int KT, int inner,
                                                    int ntiles, uint8_t *__restrict__ flags, unsigned uni_lo = 0u,
                                                    unsigned uni_hi = 0u, int nline = 0)
{
    const int nti = (inner + KT - 1) / KT;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int by = t / nti, bx = t - by * nti;
        const size_t row0 = (size_t)by * inner + (size_t)bx * KT;
        const int rows = min(KT, inner - bx * KT);
        const size_t len = (size_t)rows * unit;
        const uint8_t *p = base + row0 * unit;
        bool any = false;
        if ((((uintptr_t)p | len) & 15) == 0) {
            const uint4 *q = reinterpret_cast<const uint4 *>(p);
            for (size_t i = threadIdx.x; i < len / 16 && !any; i += blockDim.x) {
                const uint4 v = q[i];
                any = (v.x | v.y | v.z | v.w) != 0u;
            }
        } else {
            for (size_t i = threadIdx.x; i < len && !any; i += blockDim.x) any = p[i] != 0;
        }
        any = __syncthreads_or(any);
        bool uni = false;
        if (uni_hi && any && rows == KT && (nline & 31) == 0 && (unit & 31) == 0 && ((uintptr_t)p & 15) == 0) {
            const unsigned need1 = CB_SELF | uni_lo | uni_hi, care1 = need1 | CB_DIR;
            const unsigned need = need1 * 0x01010101u, care = care1 * 0x01010101u;
            const unsigned care0 = (care & 0xffffff00u) | (CB_SELF | uni_hi | CB_DIR);      // cell 0 of the chunk
            const unsigned need0 = (need & 0xffffff00u) | (CB_SELF | uni_hi);
            const unsigned care7 = care & 0x00ffffffu, need7 = need & 0x00ffffffu;         // cell 31 (separator): any
            const int P = nline / 32;
            uni = true;
            for (int c = threadIdx.x; c < rows * P && uni; c += blockDim.x) {
                const int r = c / P, ch = c - r * P;
                const uint4 *q = reinterpret_cast<const uint4 *>(p + (size_t)r * unit + (size_t)ch * 32);
                const uint4 a = q[0], b = q[1];
                uni = ((a.x & care0) == need0) && ((a.y & care) == need) && ((a.z & care) == need) && ((a.w & care) == need) &&
                      ((b.x & care) == need) && ((b.y & care) == need) && ((b.z & care) == need) && ((b.w & care7) == need7);
            }
            uni = __syncthreads_and(uni);
        }
        if (threadIdx.x == 0) flags[t] = (any ? 1 : 0) | (uni ? 2 : 0);
    }
}

#endif  // ADI_CART_MISC_KERNELS (K0)

// ------------------------------------------------------------------------------------
// Reduced-system solve shared by all sweeps: exchange through shared memory.
// red: 6*NTH doubles.  ridx: this thread's slot; rstep: slot distance between consecutive
// chunks of the same line.  Returns S_p and S_{p-1} (*Sl).  Ends with every thread past
// its last read of red only after the caller's next __syncthreads().
// ------------------------------------------------------------------------------------
// GHOST (z-slab decomposition, pass 2): the line continues on the adjacent ranks; their
// boundary values Lg (= S_{-1}) and Rg (the cell after the last separator) are known.
template <int M, bool GHOST = false>
__device__ __forceinline__ double solve_reduced(const Chunk<M> &ch, const First &f, double *red,
                                                int NTH, int ridx, int rstep, int p, int P,
                                                double *Sl, double Lg = 0.0, double Rg = 0.0)
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
    } else if (GHOST) {
        nx = ghost_first();
    }
    Red r = chunk_reduced_row(ch, nx);
    if (GHOST) {
        if (p == 0) r.D = fma(-r.A, Lg, r.D);        // the PCR below ignores A_0 and C_{P-1}
        if (p == P - 1) r.D = fma(-r.C, Rg, r.D);
    }
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
    *Sl = (p > 0) ? b[ridx - rstep] : (GHOST ? Lg : 0.0);
    return r.D;
}

// The same reduced solve when the P chunks of a line are P consecutive lanes of one warp (z sweep with
// P a power of two <= 32): the PCR rows travel by warp shuffles -- no shared memory, no block barrier.
__device__ __forceinline__ double shfl_up_w(double v, int d, int w) { return __shfl_up_sync(0xffffffffu, v, d, w); }
__device__ __forceinline__ double shfl_dn_w(double v, int d, int w) { return __shfl_down_sync(0xffffffffu, v, d, w); }

template <int M, bool GHOST = false>
__device__ __forceinline__ double solve_reduced_warp(const Chunk<M> &ch, const First &f, int p, int P, double *Sl,
                                                     double Lg = 0.0, double Rg = 0.0)
{
    First nx;
    nx.Y = shfl_dn_w(f.Y, 1, P); nx.V = shfl_dn_w(f.V, 1, P); nx.W = shfl_dn_w(f.W, 1, P);
    if (p + 1 >= P) {
        if (GHOST) nx = ghost_first();
        else { nx.Y = 0.0; nx.V = 0.0; nx.W = 0.0; }
    }
    Red r = chunk_reduced_row(ch, nx);
    if (GHOST) {
        if (p == 0) r.D = fma(-r.A, Lg, r.D);
        if (p == P - 1) r.D = fma(-r.C, Rg, r.D);
    }
    for (int s = 1; s < P; s <<= 1) {
        Red lo, hi;
        lo.A = shfl_up_w(r.A, s, P); lo.C = shfl_up_w(r.C, s, P); lo.D = shfl_up_w(r.D, s, P);
        hi.A = shfl_dn_w(r.A, s, P); hi.C = shfl_dn_w(r.C, s, P); hi.D = shfl_dn_w(r.D, s, P);
        if (p - s < 0) { lo.A = 0.0; lo.C = 0.0; lo.D = 0.0; }
        if (p + s >= P) { hi.A = 0.0; hi.C = 0.0; hi.D = 0.0; }
        r = pcr_step(r, lo, hi);
    }
    const double sl = shfl_up_w(r.D, 1, P);
    *Sl = (p > 0) ? sl : (GHOST ? Lg : 0.0);
    return r.D;
}

// z-slab decomposition, pass 1: the separators as affine functions of the two ghosts.
// red: 10*NTH doubles (two buffers of five columns); the First exchange uses the first 3*NTH.
template <int M>
__device__ __forceinline__ Red3 solve_reduced3(const Chunk<M> &ch, const First &f, double *red, int NTH,
                                               int ridx, int rstep, int p, int P)
{
    red[ridx] = f.Y;
    red[NTH + ridx] = f.V;
    red[2 * NTH + ridx] = f.W;
    __syncthreads();
    First nx = ghost_first();
    if (p + 1 < P) {
        nx.Y = red[ridx + rstep];
        nx.V = red[NTH + ridx + rstep];
        nx.W = red[2 * NTH + ridx + rstep];
    }
    Red3 r = reduced_row3(chunk_reduced_row(ch, nx), p, P);
    __syncthreads();  // everybody has read the First values: the buffers are reused below
    int cur = 0;
    for (int s = 1; s < P; s <<= 1) {
        double *b = red + cur * 5 * NTH;
        b[ridx] = r.A;
        b[NTH + ridx] = r.C;
        b[2 * NTH + ridx] = r.D;
        b[3 * NTH + ridx] = r.DL;
        b[4 * NTH + ridx] = r.DR;
        __syncthreads();
        Red3 lo, hi;
        lo.A = lo.C = lo.D = lo.DL = lo.DR = 0.0;
        hi = lo;
        if (p - s >= 0) {
            const int o = ridx - s * rstep;
            lo.A = b[o]; lo.C = b[NTH + o]; lo.D = b[2 * NTH + o]; lo.DL = b[3 * NTH + o]; lo.DR = b[4 * NTH + o];
        }
        if (p + s < P) {
            const int o = ridx + s * rstep;
            hi.A = b[o]; hi.C = b[NTH + o]; hi.D = b[2 * NTH + o]; hi.DL = b[3 * NTH + o]; hi.DR = b[4 * NTH + o];
        }
        r = pcr_step3(r, lo, hi);
        cur ^= 1;
    }
    return r;
}

// ------------------------------------------------------------------------------------
// K1: sweeps along the strided axes.  blockDim = (KT lines along z, P chunks);
// grid = (ceil(nz/KT), ny) for AXIS 0 and (ceil(nz/KT), nx) for AXIS 1.
// EXPL (AXIS 0 only): the input is T^n and the explicit stage
// R0 = T + beta*(Lx+Ly+Lz) (adi3d_numba_coeff.py:298) is applied while loading.
// Shared memory: rv[M][NTH] pivot reciprocals, red[6*NTH] reduced-system exchange.
// ------------------------------------------------------------------------------------
// ---- small PTX helpers -----------------------------------------------------------------
// Loads are written as volatile asm so that the compiler keeps them, in program order, in
// front of the cp.async group and its wait: every global operand of a chunk is then in
// flight at once (one memory round trip per thread).
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ double ldg_f64(const void *p)
{
    double v;
    asm volatile("ld.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned ldg_u8(const void *p)
{
    unsigned v;
    asm volatile("ld.global.u8 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void cp_async8(unsigned dst_smem, const void *src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst_smem), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async16(void *dst_smem, const void *src, int src_bytes)
{
    const unsigned s = (unsigned)__cvta_generic_to_shared(dst_smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// Shared-memory columns of the strided sweeps: slot s of cell e of thread tid lives at
// col[(s*M + e)*NTH] (col = smem + tid), so a warp access is a run of consecutive doubles.
//   slot 0  dense coefficient (staged with cp.async)  ->  la (NS 2) / rinv (NS 1)
//   slot 1  [explicit stage: y- neighbour row]        ->  u  (NS 2)
//   slot 2  [explicit stage: y+ neighbour row]        ->  reduced-system exchange buffer
template <int M, bool STAGED>
struct StridedOps {
    const double *coeff, *qp, *dvp;  // already offset to the chunk's first cell; may be null
    unsigned sl;  // stride between consecutive cells of the line (elements)
    int nv;       // valid cells of this chunk (0 for an out-of-range lane)
    double *col;
    int NTH;
    __device__ __forceinline__ double coef(int e) const
    {
        if (STAGED) return col[e * NTH];
        return e < nv ? coeff[e * sl] : 0.0;
    }
    __device__ __forceinline__ double q(int e) const { return (qp && e < nv) ? qp[e * sl] : 0.0; }
    __device__ __forceinline__ double dirv(int e) const { return (dvp && e < nv) ? dvp[e * sl] : 0.0; }
    __device__ __forceinline__ void put2(int e, double la, double u) { col[e * NTH] = la; col[(M + e) * NTH] = u; }
    __device__ __forceinline__ double la(int e) const { return col[e * NTH]; }
    __device__ __forceinline__ double u(int e) const { return col[(M + e) * NTH]; }
    __device__ __forceinline__ void put1(int e, double v) { col[e * NTH] = v; }
    __device__ __forceinline__ double rinv(int e) const { return col[e * NTH]; }
};

template <int AXIS, int M, int NS, int CMODE, bool EXTRA, bool EXPL, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB) k_sweep_strided(const SweepArgs a)
{
    // STAGED: every global operand of the chunk is requested up front -- T and codes into
    // registers, the coefficient and (explicit stage) the y/z neighbour values with cp.async
    // into the thread's shared-memory column -- so the thread pays one memory round trip.
    constexpr bool STAGED = (NS == 2);
    extern __shared__ double smem[];
    const int KT = blockDim.x, P = blockDim.y;
    const int kk = threadIdx.x, p = threadIdx.y;
    const int NTH = KT * P;
    const int tid = p * KT + kk;
    const int k = blockIdx.x * KT + kk;
    const int n = (AXIS == 0) ? a.nx : a.ny;
    const unsigned sl = (AXIS == 0) ? (unsigned)a.ny * (unsigned)a.nz : (unsigned)a.nz;
    constexpr unsigned LO = (AXIS == 0) ? CB_XM : CB_YM;
    constexpr unsigned HI = (AXIS == 0) ? CB_XP : CB_YP;
    const int t0 = p * M;
    const int nv = (k < a.nz) ? min(max(n - t0, 0), M) : 0;
    // first cell of the chunk (clamped in-range so that pointer arithmetic stays valid)
    const size_t idx0 = ((AXIS == 0) ? (size_t)blockIdx.y * a.nz : (size_t)blockIdx.y * a.ny * a.nz) +
                        (size_t)min(k, a.nz - 1) + (size_t)min(t0, n - 1) * sl;
    double *col = smem + tid;
    double *red = smem + (size_t)NS * M * NTH;            // behind the factor slots
    double *halo = smem + (size_t)3 * M * NTH;            // [2][M][P], explicit stage only
    const double *tp = a.in + idx0;

    // last lane of the z tile, or the lane holding the last cell of the line
    const bool hi_edge = (k == a.nz - 1) || (kk == KT - 1 && k < a.nz - 1);
    Chunk<M> ch;
    double xprev = 0.0, xnext = 0.0;
    if (STAGED) {
        // Cells beyond the chunk's valid range (line end / z tile end) re-read the last valid
        // cell (or the clamped first one): their code is forced to 0, so the values are
        // never used.  Likewise neighbour rows outside the domain re-read the cell itself.
        const int nvm1 = max(nv - 1, 0);
        const unsigned sl8 = sl * 8u;
        const char *tb = reinterpret_cast<const char *>(tp);
        {
            const uint8_t *cb = a.code + idx0;
#pragma unroll
            for (int e = 0; e < M; ++e) {
                const unsigned cv = ldg_u8(cb + (size_t)((unsigned)min(e, nvm1) * sl));
                ch.set_code(e, e < nv ? cv : 0u);
            }
        }
#pragma unroll
        for (int e = 0; e < M; ++e) ch.T[e] = ldg_f64(tb + (size_t)((unsigned)min(e, nvm1) * sl8));
        const unsigned scol = smem_u32(col);
        const unsigned nth8 = (unsigned)NTH * 8u;
        if (CMODE == 2) {
            const char *cf = reinterpret_cast<const char *>(a.coeff + idx0);
            if (!a.sparse) {
#pragma unroll
                for (int e = 0; e < M; ++e) cp_async8(scol + e * nth8, cf + (size_t)((unsigned)min(e, nvm1) * sl8));
            } else {
                // surface-only coefficient field: the two ends of the line (always exposed when active) are
                // requested now, cells next to an interior void once the codes have arrived (below)
#pragma unroll
                for (int e = 0; e < M; ++e) {
                    const int cell = t0 + e;
                    if (e < nv && (cell == 0 || cell == n - 1)) cp_async8(scol + e * nth8, cf + (size_t)((unsigned)e * sl8));
                }
            }
        }
        if (EXPL) {
            // AXIS 0: blockIdx.y is the y index
            const long long ym_off = blockIdx.y > 0 ? -8ll * a.nz : 0ll;
            const long long yp_off = (int)blockIdx.y + 1 < a.ny ? 8ll * a.nz : 0ll;
#pragma unroll
            for (int e = 0; e < M; ++e) {
                const char *pc = tb + (size_t)((unsigned)min(e, nvm1) * sl8);
                cp_async8(scol + (M + e) * nth8, pc + ym_off);
                cp_async8(scol + (2 * M + e) * nth8, pc + yp_off);
            }
            // z halo of the tile: the cell before its first lane / after its last valid lane.  At a
            // slab boundary (k == 0 / k == nz-1 with an adjacent rank) it comes from the T plane
            // received from that rank; a.zlo / a.zhi are (nx, ny) planes.
            if (kk == 0) {
                const long long off = k > 0 ? -8ll : 0ll;
                const unsigned sh = smem_u32(halo + p);
                const bool pl = (k == 0) && a.zlo;
                const char *zb = reinterpret_cast<const char *>(a.zlo) + ((size_t)min(t0, n - 1) * a.ny + blockIdx.y) * 8;
                const size_t zs = (size_t)a.ny * 8;
#pragma unroll
                for (int e = 0; e < M; ++e)
                    cp_async8(sh + e * (unsigned)P * 8u,
                              pl ? zb + (size_t)min(e, nvm1) * zs : tb + (size_t)((unsigned)min(e, nvm1) * sl8) + off);
            }
            if (hi_edge) {
                const long long off = (k + 1 < a.nz && nv > 0) ? 8ll : 0ll;
                const unsigned sh = smem_u32(halo + M * P + p);
                const bool pl = (k == a.nz - 1) && a.zhi;
                const char *zb = reinterpret_cast<const char *>(a.zhi) + ((size_t)min(t0, n - 1) * a.ny + blockIdx.y) * 8;
                const size_t zs = (size_t)a.ny * 8;
#pragma unroll
                for (int e = 0; e < M; ++e)
                    cp_async8(sh + e * (unsigned)P * 8u,
                              pl ? zb + (size_t)min(e, nvm1) * zs : tb + (size_t)((unsigned)min(e, nvm1) * sl8) + off);
            }
            xprev = ldg_f64(tb - ((nv > 0 && t0 > 0) ? (long long)sl8 : 0ll));
            xnext = ldg_f64(tb + ((nv == M && t0 + M < n) ? (long long)M * sl8 : 0ll));
        }
        // The barrier keeps every load above it (the compiler would otherwise sink them next
        // to their uses, one memory round trip per batch); the wait comes after it.  It also tells
        // whether the tile holds an active cell at all: in place, a tile of void cells (the space around
        // a part that is still being built) has nothing to solve and nothing to write.
        bool any = false;
#pragma unroll
        for (int w = 0; w < (M + 3) / 4; ++w) any = any || (ch.cw[w] & 0x01010101u) != 0u;
        const bool live = __syncthreads_or(any);
        cp_async_wait_all();
        if (!EXPL && a.in == a.out && !live) return;
    } else {
        const uint8_t *cp = a.code + idx0;
#pragma unroll
        for (int e = 0; e < M; ++e) ch.set_code(e, e < nv ? (unsigned)cp[e * sl] : 0u);
#pragma unroll
        for (int e = 0; e < M; ++e) ch.T[e] = e < nv ? tp[e * sl] : 0.0;
    }
    if (!STAGED && !EXPL && a.in == a.out) {
        bool any = false;
#pragma unroll
        for (int w = 0; w < (M + 3) / 4; ++w) any = any || (ch.cw[w] & 0x01010101u) != 0u;
        if (!__syncthreads_or(any)) return;
    }
    // solid tile rows (all chunks of the warp full and without void / Dirichlet cells): the bulk of a part
    const bool solid = !EXPL && NS == 2 && __all_sync(0xffffffffu, nv == M && chunk_solid<M>(ch, LO, HI));
    if (!solid) {
#pragma unroll
        for (int e = 0; e < M; ++e) ch.T[e] = ch.active(e) ? ch.T[e] : 0.0;  // load rule (adi_core.h)
    }
    const bool ends_only = STAGED && CMODE == 2 && a.sparse && solid;
    if (STAGED && CMODE == 2 && a.sparse) {
        // the codes are in: coefficient slots of the cells with an exposed face along this axis are fetched
        // now, the others zeroed; a solid chunk can only be exposed at its two end cells
        const char *cf = reinterpret_cast<const char *>(a.coeff + idx0);
        const unsigned sl8 = sl * 8u;
#pragma unroll
        for (int e = 0; e < M; ++e) {
            if (solid && e != 0 && e != M - 1) continue;
            const unsigned c = ch.code(e);  // 0 beyond the chunk's valid cells
            const int cell = t0 + e;
            if (e < nv && (cell == 0 || cell == n - 1)) continue;  // staged above
            col[e * NTH] = ((c & CB_SELF) && (c & (LO | HI)) != (LO | HI)) ? ldg_f64(cf + (size_t)((unsigned)e * sl8)) : 0.0;
        }
    }

    if (EXPL && STAGED) {
        double prev = (ch.code(0) & CB_XM) ? xprev : 0.0;
        const double nxt = (ch.code(M - 1) & CB_XP) ? xnext : 0.0;
#pragma unroll
        for (int e = 0; e < M; ++e) {
            const unsigned c = ch.code(e);
            const double ymr = col[(M + e) * NTH], ypr = col[(2 * M + e) * NTH];
            const double ym = (c & CB_YM) ? ymr : 0.0;
            const double yp = (c & CB_YP) ? ypr : 0.0;
            // z neighbours: the adjacent lanes hold them (already 0 where void); the first /
            // last lane of the z tile takes the staged halo value instead
            double zm = __shfl_up_sync(0xffffffffu, ch.T[e], 1);
            double zp = __shfl_down_sync(0xffffffffu, ch.T[e], 1);
            if (kk == 0) zm = (c & CB_ZM) ? halo[e * P + p] : 0.0;
            if (hi_edge) zp = (c & CB_ZP) ? halo[(M + e) * P + p] : 0.0;
            // x neighbours come from the chunk registers: already 0 where void; a void cell
            // itself (code 0, T 0) must stay 0 whatever its neighbours hold
            const double xp = (e < M - 1) ? ch.T[e + 1] : nxt;
            const double r0 = explicit_r0(c, ch.T[e], prev, xp, ym, yp, zm, zp, a.k);
            prev = ch.T[e];
            ch.T[e] = (c & CB_SELF) ? r0 : 0.0;
        }
    } else if (EXPL) {
        // neighbour values are only ever loaded where the code says the neighbour is active
        const unsigned c0 = ch.code(0), cl = ch.code(M - 1);
        double prev = (c0 & CB_XM) ? *(tp - sl) : 0.0;
        const double nxt = (cl & CB_XP) ? tp[M * sl] : 0.0;
        const int sy = a.nz;
#pragma unroll
        for (int e = 0; e < M; ++e) {
            const unsigned c = ch.code(e);
            const double *pc = tp + e * sl;
            const double ym = (c & CB_YM) ? *(pc - sy) : 0.0;
            const double yp = (c & CB_YP) ? *(pc + sy) : 0.0;
            // at a slab boundary the z neighbour lives in the plane received from the adjacent rank
            const double zm = (c & CB_ZM) ? (k > 0 ? *(pc - 1) : a.zlo[(size_t)(t0 + e) * a.ny + blockIdx.y]) : 0.0;
            const double zp = (c & CB_ZP) ? (k + 1 < a.nz ? *(pc + 1) : a.zhi[(size_t)(t0 + e) * a.ny + blockIdx.y]) : 0.0;
            const double xp = (e < M - 1) ? ch.T[e + 1] : nxt;
            const double r0 = explicit_r0(c, ch.T[e], prev, xp, ym, yp, zm, zp, a.k);
            prev = ch.T[e];
            ch.T[e] = (c & CB_SELF) ? r0 : 0.0;
        }
    }

    StridedOps<M, STAGED> ops;
    ops.coeff = (CMODE == 2) ? a.coeff + idx0 : nullptr;
    ops.qp = (EXTRA && a.q) ? a.q + idx0 : nullptr;
    ops.dvp = (EXTRA && a.dirv) ? a.dirv + idx0 : nullptr;
    ops.sl = sl;
    ops.nv = nv;
    ops.col = col;
    ops.NTH = NTH;

    First f;
    if (ends_only) f = chunk_forward<M, CMODE, EXTRA, NS, true, CMODE == 2>(ch, ops, LO, HI, a.k);
    else if (solid) f = chunk_forward<M, CMODE, EXTRA, NS, true>(ch, ops, LO, HI, a.k);
    else f = chunk_forward<M, CMODE, EXTRA, NS, false>(ch, ops, LO, HI, a.k);
    if (EXPL && STAGED) __syncthreads();  // slot 2 (other threads' y+ values) becomes the exchange buffer
    double Sl;
    const double S = solve_reduced<M>(ch, f, red, NTH, tid, KT, p, P, &Sl);
    chunk_backward<M, EXTRA, NS>(ch, ops, LO, HI, a.k.g, Sl, S);

    double *op = a.out + idx0;
    if (solid) {
#pragma unroll
        for (int e = 0; e < M; ++e) op[e * sl] = ch.T[e];
    } else if (a.in == a.out) {
        // in place: void cells are simply not written
#pragma unroll
        for (int e = 0; e < M; ++e)
            if (e < nv && ch.active(e)) op[e * sl] = ch.T[e];
    } else {
        // out of place: void cells take the input bits (adi3d_gpu_coeff.py:229)
#pragma unroll
        for (int e = 0; e < M; ++e)
            if (e < nv) op[e * sl] = ch.active(e) ? ch.T[e] : tp[e * sl];
    }
}

// ------------------------------------------------------------------------------------
// K1c: strided sweeps of LONG lines (1025..4096 cells) on a thread-block CLUSTER.
// A line of up to CS*1024 cells is shared by the CS CTAs of a cluster (cluster dims (1,1,CS),
// blockIdx.z = rank in the cluster): CTA c owns chunks [c*P, (c+1)*P) of every line of the tile and
// keeps them in registers exactly like K1 (M = 16, two factors per cell in its own shared memory).
// The separators are solved in two levels (see the kernel body): PCR inside each CTA, then one exchange of
// the CTAs' interface relations through distributed shared memory (cluster.map_shared_rank) -- two
// cluster barriers per tile in all.  blockDim = (KT, P) with KT*P <= 512, one CTA per SM, 2 (lines <= 2048
// cells) or 4 CTAs per line.
// ------------------------------------------------------------------------------------
template <int AXIS, int CMODE, bool EXTRA, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB) k_sweep_strided_cl(const SweepArgs a)
{
    constexpr int M = 16, NS = 2;
    extern __shared__ double smem[];
    const int KT = blockDim.x, P = blockDim.y;
    const int kk = threadIdx.x, p = threadIdx.y;
    const int c = blockIdx.z, CS = gridDim.z;
    const int NTH = KT * P;
    const int tid = p * KT + kk;
    const int k = blockIdx.x * KT + kk;
    const int n = (AXIS == 0) ? a.nx : a.ny;
    const unsigned sl = (AXIS == 0) ? (unsigned)a.ny * (unsigned)a.nz : (unsigned)a.nz;
    constexpr unsigned LO = (AXIS == 0) ? CB_XM : CB_YM;
    constexpr unsigned HI = (AXIS == 0) ? CB_XP : CB_YP;
    const int t0 = (c * P + p) * M;
    const int nv = (k < a.nz) ? min(max(n - t0, 0), M) : 0;
    const size_t idx0 = ((AXIS == 0) ? (size_t)blockIdx.y * a.nz : (size_t)blockIdx.y * a.ny * a.nz) +
                        (size_t)min(k, a.nz - 1) + (size_t)min(t0, n - 1) * sl;
    double *col = smem + tid;
    double *red = smem + (size_t)NS * M * NTH;
    const double *tp = a.in + idx0;

    Chunk<M> ch;
    {
        const int nvm1 = max(nv - 1, 0);
        const unsigned sl8 = sl * 8u;
        const char *tb = reinterpret_cast<const char *>(tp);
        const uint8_t *cb = a.code + idx0;
#pragma unroll
        for (int e = 0; e < M; ++e) {
            const unsigned cv = ldg_u8(cb + (size_t)((unsigned)min(e, nvm1) * sl));
            ch.set_code(e, e < nv ? cv : 0u);
        }
#pragma unroll
        for (int e = 0; e < M; ++e) ch.T[e] = ldg_f64(tb + (size_t)((unsigned)min(e, nvm1) * sl8));
        if (CMODE == 2) {
            const unsigned scol = smem_u32(col);
            const unsigned nth8 = (unsigned)NTH * 8u;
            const char *cf = reinterpret_cast<const char *>(a.coeff + idx0);
#pragma unroll
            for (int e = 0; e < M; ++e) cp_async8(scol + e * nth8, cf + (size_t)((unsigned)min(e, nvm1) * sl8));
        }
        __syncthreads();
        cp_async_wait_all();
    }
    const bool solid = __all_sync(0xffffffffu, nv == M && chunk_solid<M>(ch, LO, HI));
    if (!solid) {
#pragma unroll
        for (int e = 0; e < M; ++e) ch.T[e] = ch.active(e) ? ch.T[e] : 0.0;  // load rule (adi_core.h)
    }
    StridedOps<M, true> ops;
    ops.coeff = (CMODE == 2) ? a.coeff + idx0 : nullptr;
    ops.qp = (EXTRA && a.q) ? a.q + idx0 : nullptr;
    ops.dvp = (EXTRA && a.dirv) ? a.dirv + idx0 : nullptr;
    ops.sl = sl;
    ops.nv = nv;
    ops.col = col;
    ops.NTH = NTH;
    First f;
    if (solid) f = chunk_forward<M, CMODE, EXTRA, NS, true>(ch, ops, LO, HI, a.k);
    else f = chunk_forward<M, CMODE, EXTRA, NS, false>(ch, ops, LO, HI, a.k);
    // Two-level solve: (1) inside the CTA, PCR with three right-hand-side columns gives every separator as
    // an affine function of the CTA's two ghost values (the last cell of the CTA before, the first cell of
    // the CTA after) -- CTA barriers only; (2) the CS interface relations of a line meet through
    // distributed shared memory (ONE cluster barrier), each CTA solves the tiny inter-CTA system for its
    // own ghosts and finishes.  The same scheme links the GPUs of the z-slab decomposition.
    namespace cg = cooperative_groups;
    cg::cluster_group cl = cg::this_cluster();
    const Red3 r3 = solve_reduced3<M>(ch, f, red, NTH, tid, KT, p, P);
    double *sI = red + (size_t)10 * NTH;   // [6][KT] interface relation of this CTA's segment, per lane
    double *sG = sI + 6 * KT;              // [2][KT] ghosts
    if (p == 0) {
        sI[kk] = fma(f.W, r3.D, f.Y);
        sI[KT + kk] = fma(f.W, r3.DL, f.V);
        sI[2 * KT + kk] = f.W * r3.DR;
    }
    if (p == P - 1) {
        sI[3 * KT + kk] = r3.D;
        sI[4 * KT + kk] = r3.DL;
        sI[5 * KT + kk] = r3.DR;
    }
    cl.sync();
    if (p == 0) {
        double Lg, Rg;
        iface_solve([&](int rr) {
            const double *q = (rr == c ? sI : cl.map_shared_rank(sI, rr)) + kk;
            Iface v;
            v.yf = q[0]; v.vf = q[KT]; v.wf = q[2 * KT]; v.yl = q[3 * KT]; v.vl = q[4 * KT]; v.wl = q[5 * KT];
            return v;
        }, CS, c, &Lg, &Rg);
        sG[kk] = Lg;
        sG[KT + kk] = Rg;
    }
    cl.sync();  // the peers have read this CTA's relation; the ghosts are visible to the whole CTA
    const double Lg = sG[kk], Rg = sG[KT + kk];
    const double S = fma(r3.DR, Rg, fma(r3.DL, Lg, r3.D));
    red[tid] = S;
    __syncthreads();
    const double Sl = p > 0 ? red[tid - KT] : Lg;
    chunk_backward<M, EXTRA, NS>(ch, ops, LO, HI, a.k.g, Sl, S);

    double *op = a.out + idx0;
    if (solid) {
#pragma unroll
        for (int e = 0; e < M; ++e) op[e * sl] = ch.T[e];
    } else if (a.in == a.out) {
#pragma unroll
        for (int e = 0; e < M; ++e)
            if (e < nv && ch.active(e)) op[e * sl] = ch.T[e];
    } else {
#pragma unroll
        for (int e = 0; e < M; ++e)
            if (e < nv) op[e * sl] = ch.active(e) ? ch.T[e] : tp[e * sl];
    }
}

// ------------------------------------------------------------------------------------
// K3: sweep along z (contiguous).  blockDim = (P chunks, LT lines); grid = ceil(nx*ny/LT).
// Shared memory: sT[LT][P*M] and sC[LT][P*M] (field / coefficient tiles; sC then holds the
// pivot reciprocals), sCode[LT][P*M].  Inside each chunk the 16-byte pairs are stored at
// position j ^ (p & 7): the per-thread 16-byte reads of a quarter warp (8 consecutive
// chunks, 128 B apart) then hit eight different bank groups, without padding.
// The reduced-system exchange buffer aliases sT while the chunk lives in registers.
// ------------------------------------------------------------------------------------

template <int M>
__device__ __forceinline__ int zslot(int chunk, int pair)
{
    constexpr int SW = (M / 2 >= 8) ? 7 : (M / 2 - 1);
    return chunk * M + 2 * (pair ^ (chunk & SW));
}

template <int M>
struct TileOps {
    double *c;          // this chunk's coefficient slots, then la (NS 2) / rinv (NS 1)
    double *uu;         // this chunk's u slots (NS 2)
    int sw;             // p & SW
    const double *qp, *dvp;
    int nv;
    __device__ __forceinline__ int at(int e) const { return 2 * ((e >> 1) ^ sw) + (e & 1); }
    __device__ __forceinline__ double coef(int e) const { return c[at(e)]; }
    __device__ __forceinline__ double q(int e) const { return (qp && e < nv) ? qp[e] : 0.0; }
    __device__ __forceinline__ double dirv(int e) const { return (dvp && e < nv) ? dvp[e] : 0.0; }
    __device__ __forceinline__ void put2(int e, double la, double u) { c[at(e)] = la; uu[at(e)] = u; }
    __device__ __forceinline__ double la(int e) const { return c[at(e)]; }
    __device__ __forceinline__ double u(int e) const { return uu[at(e)]; }
    __device__ __forceinline__ void put1(int e, double v) { c[at(e)] = v; }
    __device__ __forceinline__ double rinv(int e) const { return c[at(e)]; }
};

// smem: sT[LT][RL] | sC[LT][RL] | (NS 2: sU[LT][RL]) | sCode[LT][RL bytes]
// ZMODE 4: z-slab "solve first" pass -- the segment is solved in place with both ghosts at zero and the
// right-hand-side part of its interface relation (its first and last value) goes to a.iface_dyn; the
// ghost terms are added afterwards by k_spike_apply from the cached unit-ghost responses.
// ZMODE 0: whole line on this GPU.  1: z-slab pass 1 -- writes the interface relation of each
// local line segment to a.iface_dyn / a.iface_stat and leaves the field untouched (3: the
// right-hand-side part a.iface_dyn only).  2: z-slab pass 2 -- finishes the
// segment with the ghost values a.ghost.  ZMODE != 0 needs nz % M == 0 (last cell = a separator).
template <int M, int NS, int CMODE, bool EXTRA, int MAXT, int MINB, int ZMODE = 0>
__global__ void __launch_bounds__(MAXT, MINB) k_sweep_z(const SweepArgs a, const int vec)
{
    static_assert(M % 4 == 0, "chunk length must be a multiple of 4");
    extern __shared__ double smem[];
    const int P = blockDim.x, LT = blockDim.y;
    const int p = threadIdx.x, ln = threadIdx.y;
    const int NTH = P * LT;
    const int tid = ln * P + p;
    const int nz = a.nz;
    const size_t nlines = (size_t)a.nx * a.ny;
    const size_t L0 = (size_t)blockIdx.x * LT;
    const int RL = P * M;  // doubles per staged line
    double *sT = smem;
    double *sC = sT + (size_t)LT * RL;
    double *sU = sC + (size_t)LT * RL;
    uint8_t *sCode = reinterpret_cast<uint8_t *>(sU + (NS == 2 ? (size_t)LT * RL : 0));
    double *red = sT;
    constexpr int SW = (M / 2 >= 8) ? 7 : (M / 2 - 1);

    // ---- stage in ----
    const int ppl = RL / 2;  // 16-byte pairs per staged line
    if (vec) {
        for (int l = 0; l < LT; ++l) {
            const size_t line = L0 + l;
            const bool lok = line < nlines;
            const double *srcT = a.in + (lok ? line * (size_t)nz : 0);
            const double *srcC = (CMODE == 2) ? a.coeff + (lok ? line * (size_t)nz : 0) : nullptr;
            for (int pl = tid; pl < ppl; pl += NTH) {
                const int dst = l * RL + zslot<M>(pl / (M / 2), pl % (M / 2));
                const int z = 2 * pl;
                const bool ok = lok && z < nz;
                cp_async16(sT + dst, srcT + (ok ? z : 0), ok ? 16 : 0);
                if (CMODE == 2) cp_async16(sC + dst, srcC + (ok ? z : 0), ok ? 16 : 0);
            }
        }
    } else {
        for (int l = 0; l < LT; ++l) {
            const size_t line = L0 + l;
            const bool lok = line < nlines;
            for (int z = tid; z < RL; z += NTH) {
                const int dst = l * RL + zslot<M>(z / M, (z % M) >> 1) + (z & 1);
                const bool ok = lok && z < nz;
                sT[dst] = ok ? a.in[line * (size_t)nz + z] : 0.0;
                if (CMODE == 2) sC[dst] = ok ? a.coeff[line * (size_t)nz + z] : 0.0;
            }
        }
    }
    if (vec && (nz & 15) == 0) {
        for (int l = 0; l < LT; ++l) {
            const size_t line = L0 + l;
            const bool lok = line < nlines;
            for (int c16 = tid; c16 < RL / 16; c16 += NTH) {
                const bool ok = lok && 16 * c16 < nz;
                cp_async16(sCode + (size_t)l * RL + 16 * c16, a.code + (ok ? line * (size_t)nz + 16 * c16 : 0),
                           ok ? 16 : 0);
            }
        }
    } else {
        for (int l = 0; l < LT; ++l) {
            const size_t line = L0 + l;
            const bool lok = line < nlines;
            for (int z = tid; z < RL; z += NTH)
                sCode[(size_t)l * RL + z] = (lok && z < nz) ? a.code[line * (size_t)nz + z] : (uint8_t)0;
        }
    }
    cp_async_wait_all();
    __syncthreads();

    // ---- chunk to registers ----
    Chunk<M> ch;
    double *myT = sT + (size_t)ln * RL + p * M;
    {
        const unsigned *cb = reinterpret_cast<const unsigned *>(sCode + (size_t)ln * RL + p * M);
#pragma unroll
        for (int w = 0; w < M / 4; ++w) ch.cw[w] = cb[w];
#pragma unroll
        for (int j = 0; j < M / 2; ++j) {
            const double2 v = *reinterpret_cast<const double2 *>(myT + 2 * (j ^ (p & SW)));
            ch.T[2 * j] = ch.active(2 * j) ? v.x : 0.0;          // load rule (adi_core.h)
            ch.T[2 * j + 1] = ch.active(2 * j + 1) ? v.y : 0.0;
        }
    }
    // a line = P consecutive lanes of one warp: warp-shuffle PCR (block-uniform)
    const bool warp_lines = P <= 32 && (P & (P - 1)) == 0;
    // warps whose chunks are all solid (adi_core.h) take the row arithmetic with the code folded away
    const bool solid = NS == 2 && __all_sync(0xffffffffu, chunk_solid<M>(ch, CB_ZM, CB_ZP));
    if ((ZMODE == 0 || ZMODE == 2) && a.in == a.out) {  // (ZMODE 4 must still emit its relation)
        // lines without an active cell (the void around a part that is still being built): nothing to
        // solve, nothing to write (the sweep is in place); this barrier also frees sT for the exchange buffer
        bool any = false;
#pragma unroll
        for (int w = 0; w < M / 4; ++w) any = any || (ch.cw[w] & 0x01010101u) != 0u;
        if (!__syncthreads_or(any)) return;
    } else {
        __syncthreads();  // sT is reused as the reduced-system exchange buffer from here on
    }

    TileOps<M> ops;
    ops.c = sC + (size_t)ln * RL + p * M;
    ops.uu = sU + (size_t)ln * RL + p * M;
    ops.sw = p & SW;
    {
        const size_t line = min(L0 + ln, nlines - 1);
        const size_t g0 = line * (size_t)nz + (size_t)min(p * M, nz - 1);
        ops.qp = (EXTRA && a.q) ? a.q + g0 : nullptr;
        ops.dvp = (EXTRA && a.dirv) ? a.dirv + g0 : nullptr;
        ops.nv = (L0 + ln < nlines) ? min(max(nz - p * M, 0), M) : 0;
    }

    // 0: general rows; 1: cells 0..M-2 uniform; 2: cell 0 general, cells 1..M-2 uniform (adi_core.h).  Only without
    // a dense coefficient field (scalar Robin / no Robin term): a warp here is ONE line, and a dense field read at
    // exposed cells only would cost the general path a second memory round trip (see ensure_sparse)
    int path = 0;
    if (CMODE != 2 && a.uni) {
        if (__all_sync(0xffffffffu, chunk_uniform<M, 1>(ch, CB_ZM, CB_ZP)))
            path = __all_sync(0xffffffffu, chunk_uniform<M, 0>(ch, CB_ZM, CB_ZP)) ? 1 : 2;
    }
    First f;
    UniHead hd;
    hd.al = hd.bl = hd.br = 0.0;
    if (CMODE != 2 && path != 0) {
        const Row sep = make_row<CMODE, EXTRA>(ch.code(M - 1), CB_ZM, CB_ZP, ch.T[M - 1], 0.0, 0.0, 0.0, a.k);
        if (path == 1) {
            f = chunk_forward_uniform<M, 0>(ch, a.uc, sep, sep, hd);
        } else {
            const Row head = make_row<CMODE, EXTRA>(ch.code(0), CB_ZM, CB_ZP, ch.T[0], 0.0, 0.0, 0.0, a.k);
            f = chunk_forward_uniform<M, 1>(ch, a.uc, sep, head, hd);
        }
    } else {
        if (solid) f = chunk_forward<M, CMODE, EXTRA, NS, true>(ch, ops, CB_ZM, CB_ZP, a.k);
        else f = chunk_forward<M, CMODE, EXTRA, NS, false>(ch, ops, CB_ZM, CB_ZP, a.k);
    }
    if (ZMODE == 1) {
        const Red3 r = solve_reduced3<M>(ch, f, red, NTH, tid, 1, p, P);
        const size_t line = L0 + ln;
        if (line < nlines) {
            // x_first = f.Y + f.V*L + f.W*S_0,  x_last = S_{P-1}   (adi_core.h, Iface)
            if (p == 0) {
                a.iface_dyn[line] = fma(f.W, r.D, f.Y);
                a.iface_stat[line] = fma(f.W, r.DL, f.V);
                a.iface_stat[nlines + line] = f.W * r.DR;
            }
            if (p == P - 1) {
                a.iface_dyn[nlines + line] = r.D;
                a.iface_stat[2 * nlines + line] = r.DL;
                a.iface_stat[3 * nlines + line] = r.DR;
            }
        }
        return;
    }
    if (ZMODE == 3) {
        // right-hand-side part only (the matrix part is cached by the caller while mask, packs,
        // dt and theta stay the same): the plain solve with both ghosts at zero
        double Sl0;
        double S0;
        if (warp_lines) S0 = solve_reduced_warp<M, true>(ch, f, p, P, &Sl0, 0.0, 0.0);
        else S0 = solve_reduced<M, true>(ch, f, red, NTH, tid, 1, p, P, &Sl0, 0.0, 0.0);
        const size_t line = L0 + ln;
        if (line < nlines) {
            if (p == 0) a.iface_dyn[line] = fma(f.W, S0, f.Y);
            if (p == P - 1) a.iface_dyn[nlines + line] = S0;
        }
        return;
    }
    double Sl, Lg = 0.0, Rg = 0.0;
    if (ZMODE == 2) {
        const size_t line = min(L0 + ln, nlines - 1);
        if (p == 0) Lg = a.ghost[line];
        if (p == P - 1) Rg = a.ghost[nlines + line];
    }
    double S;
    constexpr bool GHOSTS = (ZMODE == 2 || ZMODE == 4);  // 4: both ghosts at zero, as in ZMODE 3
    if (warp_lines) S = solve_reduced_warp<M, GHOSTS>(ch, f, p, P, &Sl, Lg, Rg);
    else S = solve_reduced<M, GHOSTS>(ch, f, red, NTH, tid, 1, p, P, &Sl, Lg, Rg);
    if (CMODE != 2 && path == 1) chunk_backward_uniform<M, 0>(ch, a.uc, hd, Sl, S);
    else if (CMODE != 2 && path == 2) chunk_backward_uniform<M, 1>(ch, a.uc, hd, Sl, S);
    else chunk_backward<M, EXTRA, NS>(ch, ops, CB_ZM, CB_ZP, a.k.g, Sl, S);
    if (ZMODE == 4) {
        const size_t line = L0 + ln;
        if (line < nlines) {  // void cells hold 0 (load rule), as in the relation of ZMODE 3
            if (p == 0) a.iface_dyn[line] = ch.T[0];
            if (p == P - 1) a.iface_dyn[nlines + line] = ch.T[M - 1];
        }
    }
    __syncthreads();  // everybody is done reading the exchange buffer

    // ---- results back through shared memory (each thread owns its slots) ----
    // The sweep runs in place, so void cells must keep the bits they have in global memory:
    // the copy-out below only writes cells whose staged code is active.
#pragma unroll
    for (int j = 0; j < M / 2; ++j)
        *reinterpret_cast<double2 *>(myT + 2 * (j ^ (p & SW))) = make_double2(ch.T[2 * j], ch.T[2 * j + 1]);
    __syncthreads();
    const bool inplace = (a.in == a.out);
    if (vec) {
        for (int l = 0; l < LT; ++l) {
            const size_t line = L0 + l;
            if (line >= nlines) break;
            for (int pl = tid; pl < ppl; pl += NTH) {
                const int z = 2 * pl;
                if (z >= nz) break;
                const double2 v = *reinterpret_cast<const double2 *>(sT + l * RL + zslot<M>(pl / (M / 2), pl % (M / 2)));
                const unsigned cc = *reinterpret_cast<const unsigned short *>(sCode + (size_t)l * RL + z);
                const size_t g = line * (size_t)nz + z;
                const bool a0 = cc & 1u, a1 = (cc >> 8) & 1u;
                if (a0 && a1) {
                    *reinterpret_cast<double2 *>(a.out + g) = v;
                } else if (inplace) {
                    if (a0) a.out[g] = v.x;
                    if (a1) a.out[g + 1] = v.y;
                } else {
                    a.out[g] = a0 ? v.x : a.in[g];
                    a.out[g + 1] = a1 ? v.y : a.in[g + 1];
                }
            }
        }
    } else {
        for (int l = 0; l < LT; ++l) {
            const size_t line = L0 + l;
            if (line >= nlines) break;
            for (int z = tid; z < nz; z += NTH) {
                const size_t g = line * (size_t)nz + z;
                const bool act = sCode[(size_t)l * RL + z] & 1u;
                if (act) a.out[g] = sT[l * RL + zslot<M>(z / M, (z % M) >> 1) + (z & 1)];
                else if (!inplace) a.out[g] = a.in[g];
            }
        }
    }
}

#ifdef ADI_CART_MISC_KERNELS
// ------------------------------------------------------------------------------------
// K1e: explicit stage as a streaming kernel, R0 = T + beta*(Lx+Ly+Lz)T  (adi3d_numba_coeff.py:240-298).
// Threads run along z (VEC cells each, 16-byte accesses for VEC 2) and march JT rows in y with
// the y-1 / y / y+1 values of their column in registers; the x-1 / x+1 rows come from the two
// adjacent x planes (L2 hits: the blocks of neighbouring planes run at the same time), the
// z neighbours of the thread's first / last cell from the adjacent lanes' lines (L1 hits) or, at
// a slab boundary, from the T plane received from the adjacent rank.
// Void cells are copied bit for bit (the reference never touches them).
// grid = (ceil(nz/VEC/blockDim.x), ceil(ny/JT), nx).
// ------------------------------------------------------------------------------------
template <int VEC>
__global__ void __launch_bounds__(128) k_explicit(const SweepArgs a, const int JT, const int xfast)
{
    const int z = ((xfast ? blockIdx.z : blockIdx.x) * blockDim.x + threadIdx.x) * VEC;
    if (z >= a.nz) return;
    const int i = xfast ? blockIdx.x : blockIdx.z;
    const int j0 = blockIdx.y * JT, j1 = min(j0 + JT, a.ny);
    const size_t snx = (size_t)a.ny * a.nz;
    size_t idx = ((size_t)i * a.ny + j0) * a.nz + z;
    typedef typename std::conditional<VEC == 2, double2, double>::type V;
    auto ldv = [&](size_t g, double (&v)[VEC]) {
        const V t = *reinterpret_cast<const V *>(a.in + g);
        if (VEC == 2) { v[0] = reinterpret_cast<const double *>(&t)[0]; v[1] = reinterpret_cast<const double *>(&t)[VEC - 1]; }
        else v[0] = reinterpret_cast<const double *>(&t)[0];
    };
    double prev[VEC], cur[VEC], next[VEC], xm[VEC], xp[VEC], r[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) prev[v] = next[v] = xm[v] = xp[v] = 0.0;
    if (j0 > 0) ldv(idx - a.nz, prev);
    ldv(idx, cur);
    // the code of a row is fetched one row ahead: rows without an active cell (the space around a part under
    // construction) are plain copies and skip the loads of their x neighbours
    auto ldc = [&](size_t g) -> unsigned {
        if (VEC == 2) return *reinterpret_cast<const unsigned short *>(a.code + g);
        return a.code[g];
    };
    unsigned cw_next = ldc(idx);
    for (int j = j0; j < j1; ++j) {
        const unsigned cw = cw_next;
        if (j + 1 < a.ny) ldv(idx + a.nz, next);
        if (j + 1 < j1) cw_next = ldc(idx + a.nz);
        if (cw != 0u) {
            if (i > 0) ldv(idx - snx, xm);
            if (i + 1 < a.nx) ldv(idx + snx, xp);
        }
        const unsigned c0 = cw & 0xffu, cl = (cw >> (8 * (VEC - 1))) & 0xffu;
        double zlo_v = 0.0, zhi_v = 0.0;
        // halo_defer: the planes are still on their way -- the cells that need them belong to k_explicit_faces
        bool skip0 = false, skipl = false;
        if (c0 & CB_ZM) {
            if (z > 0) zlo_v = a.in[idx - 1];
            else if (a.halo_defer) skip0 = true;
            else zlo_v = a.zlo[(size_t)i * a.ny + j];
        }
        if (cl & CB_ZP) {
            if (z + VEC < a.nz) zhi_v = a.in[idx + VEC];
            else if (a.halo_defer) skipl = true;
            else zhi_v = a.zhi[(size_t)i * a.ny + j];
        }
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            const unsigned c = (cw >> (8 * v)) & 0xffu;
            const double zm = v == 0 ? zlo_v : ((c & CB_ZM) ? cur[v - 1 < 0 ? 0 : v - 1] : 0.0);
            const double zp = v == VEC - 1 ? zhi_v : ((c & CB_ZP) ? cur[v + 1 < VEC ? v + 1 : v] : 0.0);
            const double r0 = explicit_r0(c, (c & CB_SELF) ? cur[v] : 0.0, (c & CB_XM) ? xm[v] : 0.0,
                                          (c & CB_XP) ? xp[v] : 0.0, (c & CB_YM) ? prev[v] : 0.0,
                                          (c & CB_YP) ? next[v] : 0.0, zm, zp, a.k);
            r[v] = (c & CB_SELF) ? r0 : cur[v];
        }
        if (skip0 || skipl) {
            if (!skip0) a.out[idx] = r[0];
            if (VEC == 2 && !skipl) a.out[idx + VEC - 1] = r[VEC - 1];
        } else if (VEC == 2) {
            *reinterpret_cast<double2 *>(a.out + idx) = make_double2(r[0], r[VEC - 1]);
        } else {
            a.out[idx] = r[0];
        }
#pragma unroll
        for (int v = 0; v < VEC; ++v) { prev[v] = cur[v]; cur[v] = next[v]; }
        idx += a.nz;
    }
}

// K1f: explicit stage of the cells on the two faces of a z slab whose stencil reaches into the adjacent slab
// (first / last z plane, neighbour across the face active): the multi-GPU step runs k_explicit (halo_defer: it leaves
// these cells alone) while the T planes of the adjacent ranks are still travelling; this kernel follows the planes
// on the communication stream -- same arithmetic, same operand order -- so it runs beside k_explicit.  One thread
// per z line.
__global__ void k_explicit_faces(const SweepArgs a)
{
    const size_t nlines = (size_t)a.nx * a.ny;
    const size_t snx = (size_t)a.ny * a.nz;
    for (size_t l = (size_t)blockIdx.x * blockDim.x + threadIdx.x; l < nlines; l += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(l / a.ny), j = (int)(l - (size_t)i * a.ny);
        for (int side = 0; side < 2; ++side) {
            const int k = side == 0 ? 0 : a.nz - 1;
            if (side == 1 && a.nz == 1) break;                 // the single plane was done as side 0
            const size_t idx = l * (size_t)a.nz + k;
            const unsigned c = a.code[idx];
            const bool lo = k == 0 && (c & CB_ZM) && a.zlo, hi = k == a.nz - 1 && (c & CB_ZP) && a.zhi;
            if (!(c & CB_SELF) || !(lo || hi)) continue;
            const double T = a.in[idx];
            const double xm = (c & CB_XM) ? a.in[idx - snx] : 0.0, xp = (c & CB_XP) ? a.in[idx + snx] : 0.0;
            const double ym = (c & CB_YM) ? a.in[idx - a.nz] : 0.0, yp = (c & CB_YP) ? a.in[idx + a.nz] : 0.0;
            double zm = 0.0, zp = 0.0;
            if (c & CB_ZM) zm = k > 0 ? a.in[idx - 1] : (a.zlo ? a.zlo[l] : 0.0);
            if (c & CB_ZP) zp = k + 1 < a.nz ? a.in[idx + 1] : (a.zhi ? a.zhi[l] : 0.0);
            a.out[idx] = explicit_r0(c, T, xm, xp, ym, yp, zm, zp, a.k);
        }
    }
}

// ------------------------------------------------------------------------------------
// K8: z-slab exchange helpers.
// k_pack_zplanes: first and last z plane of a field / mask into contiguous (nx, ny) buffers
// (the messages of the halo exchange).  k_iface_solve: every rank solves, per line, the
// 2*nranks-unknown inter-rank system from the gathered interface relations
// all[rank][6][nlines] and keeps its own two ghost values ghost[2][nlines].
// ------------------------------------------------------------------------------------
template <typename T>
__global__ void k_pack_zplanes(const T *__restrict__ f, T *__restrict__ lo, T *__restrict__ hi, size_t nlines, int nz)
{
    for (size_t l = (size_t)blockIdx.x * blockDim.x + threadIdx.x; l < nlines; l += (size_t)gridDim.x * blockDim.x) {
        if (lo) lo[l] = f[l * nz];
        if (hi) hi[l] = f[l * nz + nz - 1];
    }
}

__global__ void k_iface_solve(const double *__restrict__ dyn, const double *__restrict__ stat, size_t sstride,
                              double *__restrict__ ghost, size_t nlines, int nranks, int rank)
{
    // dyn[rank][2][nlines] = (yf, yl), stat[rank][4][sstride] = (vf, wf, vl, wl) (sstride >= nlines: a batch of
    // lines out of a longer array; `stat` already points at the batch's first line)
    for (size_t l = (size_t)blockIdx.x * blockDim.x + threadIdx.x; l < nlines; l += (size_t)gridDim.x * blockDim.x) {
        double Lg, Rg;
        iface_solve([&](int r) {
            const double *d = dyn + (size_t)r * 2 * nlines + l;
            const double *q = stat + (size_t)r * 4 * sstride + l;
            Iface v;
            v.yf = d[0]; v.yl = d[nlines];
            v.vf = q[0]; v.wf = q[sstride]; v.vl = q[2 * sstride]; v.wl = q[3 * sstride];
            return v;
        }, nranks, rank, &Lg, &Rg);
        ghost[l] = Lg;
        ghost[nlines + l] = Rg;
    }
}

__global__ void k_fill_ghost(double *__restrict__ ghost, size_t nlines, double lo, double hi)
{
    for (size_t l = (size_t)blockIdx.x * blockDim.x + threadIdx.x; l < nlines; l += (size_t)gridDim.x * blockDim.x) {
        ghost[l] = lo;
        ghost[nlines + l] = hi;
    }
}

// Unit-ghost response ("spike") of every local line segment, as left in `field` by a ZMODE 2 solve of the
// homogeneous system with ghost 1 at one end: it decays geometrically away from that end.  One warp per
// line: K[line] = number of cells, counted from the end, up to the last one with |value| > thr;
// compact[line][kmax] = the kmax cells next to the end (end 0: cells 0.., end 1: cells nz-kmax..).
__global__ void k_spike_pack(const double *__restrict__ field, size_t nlines, int nz, int end, int kmax, double thr,
                             double *__restrict__ compact, int *__restrict__ K, int *__restrict__ maxK)
{
    const int lane = threadIdx.x & 31;
    const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const size_t nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
    for (size_t l = warp; l < nlines; l += nwarps) {
        const double *f = field + l * (size_t)nz;
        int far = 0;  // extent of the cells above the threshold, counted from the end
        for (int k0 = 0; k0 < nz; k0 += 32) {
            const int k = k0 + lane;
            const bool big = k < nz && fabs(f[k]) > thr;
            const unsigned m = __ballot_sync(0xffffffffu, big);
            if (m) {
                if (end == 0) far = max(far, k0 + 32 - __clz((int)m));          // last set lane + 1
                else far = max(far, nz - (k0 + __ffs((int)m) - 1));             // nz - first set index
            }
        }
        for (int j = lane; j < kmax; j += 32) {
            const int k = end == 0 ? j : nz - kmax + j;
            compact[l * (size_t)kmax + j] = (k >= 0 && k < nz) ? f[k] : 0.0;
        }
        if (lane == 0) {
            K[l] = min(far, kmax);
            if (far > 0) atomicMax(maxK, far);
        }
    }
}

// Most lines of a part respond alike (the response depends on the rows near the face only: every bulk line has the
// same): lines whose K and compact response equal, bit for bit, those of the canonical line `canon` get SPIKE_CANON
// set in K[line] and k_spike_apply reads the canonical row (cache-resident) instead of their own -- 2 x K x 8 bytes
// per line and step less.
constexpr int SPIKE_CANON = 1 << 30;
__global__ void k_spike_canon(const double *__restrict__ compact, int *__restrict__ K, size_t nlines, int kmax, size_t canon)
{
    const unsigned long long *c = reinterpret_cast<const unsigned long long *>(compact + canon * (size_t)kmax);
    const int kc = K[canon] & (SPIKE_CANON - 1);
    for (size_t l = (size_t)blockIdx.x * blockDim.x + threadIdx.x; l < nlines; l += (size_t)gridDim.x * blockDim.x) {
        if ((K[l] & (SPIKE_CANON - 1)) != kc) continue;
        const unsigned long long *r = reinterpret_cast<const unsigned long long *>(compact + l * (size_t)kmax);
        bool same = true;
        for (int j = 0; j < kmax && same; ++j) same = r[j] == c[j];
        if (same) K[l] |= SPIKE_CANON;
    }
}

// x = y + L*v + R*w on the cells within reach of the two ends of every local line segment (8 lanes per line).
// Cells where the response is exactly 0 (void cells, cells behind a void gap) are not touched.
// FUSED: the line's two ghosts (L, R) are not read from `ghost` but solved here from the gathered interface
// relations (iface_solve by the first lane of the line's eight, shuffled to the others): one launch and one
// round trip through memory less per batch of the overlapped multi-GPU z solve.
template <bool FUSED>
__global__ void k_spike_apply(double *__restrict__ T, const double *__restrict__ ghost, const double *__restrict__ vC,
                              const double *__restrict__ wC, const int *__restrict__ Kv, const int *__restrict__ Kw,
                              size_t nlines, int nz, int kmax, const double *__restrict__ vCanon,
                              const double *__restrict__ wCanon, const double *__restrict__ dyn = nullptr,
                              const double *__restrict__ stat = nullptr, size_t sstride = 0, int nranks = 1, int rank = 0)
{
    const int sub = threadIdx.x & 7;
    const size_t step = ((size_t)gridDim.x * blockDim.x) >> 3;
    const size_t lmax = (nlines + step - 1) / step * step;   // whole groups of eight stay together (shuffles below)
    for (size_t l = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3; l < lmax; l += step) {
        const bool ok = l < nlines;
        double L = 0.0, R = 0.0;
        if (FUSED) {
            if (ok && sub == 0) {
                iface_solve([&](int r) {
                    const double *d = dyn + (size_t)r * 2 * nlines + l;
                    const double *q = stat + (size_t)r * 4 * sstride + l;
                    Iface v;
                    v.yf = d[0]; v.yl = d[nlines];
                    v.vf = q[0]; v.wf = q[sstride]; v.vl = q[2 * sstride]; v.wl = q[3 * sstride];
                    return v;
                }, nranks, rank, &L, &R);
            }
            L = __shfl_sync(0xffffffffu, L, 0, 8);
            R = __shfl_sync(0xffffffffu, R, 0, 8);
        } else if (ok) {
            L = ghost[l]; R = ghost[nlines + l];
        }
        if (!ok) continue;
        const int kvf = Kv[l], kwf = Kw[l];
        const int kv = kvf & (SPIKE_CANON - 1), kw = kwf & (SPIKE_CANON - 1);
        double *t = T + l * (size_t)nz;
        const double *v = (kvf & SPIKE_CANON) ? vCanon : vC + l * (size_t)kmax;
        const double *w = (kwf & SPIKE_CANON) ? wCanon : wC + l * (size_t)kmax;
        const int hi0 = nz - kw;  // first cell within reach of the upper end
        for (int k = sub; k < kv; k += 8) {
            double c = L * v[k];
            if (k >= hi0) c = fma(R, w[k - (nz - kmax)], c);
            if (c != 0.0) t[k] += c;
        }
        for (int k = max(kv, hi0) + sub; k < nz; k += 8) {
            const double c = R * w[k - (nz - kmax)];
            if (c != 0.0) t[k] += c;
        }
    }
}

// ------------------------------------------------------------------------------------
// K7: exposed_mask / precompute_coeff_packs_unified (adi3d_gpu_coeff.py:31-110).
// ------------------------------------------------------------------------------------
// PackArgs: adi_mask_core.h

__device__ __forceinline__ unsigned exposed_bits(const uint8_t *__restrict__ mask, size_t idx, int i,
                                                 int j, int k, int nx, int ny, int nz,
                                                 const uint8_t *__restrict__ mlo = nullptr,
                                                 const uint8_t *__restrict__ mhi = nullptr)
{
    // bit f set <=> cell active and neighbour across face f void or outside (:38-55)
    if (!mask[idx]) return 0u;
    const size_t snx = (size_t)ny * nz;
    unsigned b = 0;
    if (!(i > 0 && mask[idx - snx])) b |= 1u;
    if (!(i + 1 < nx && mask[idx + snx])) b |= 2u;
    if (!(j > 0 && mask[idx - nz])) b |= 4u;
    if (!(j + 1 < ny && mask[idx + nz])) b |= 8u;
    const size_t ij = (size_t)i * ny + j;
    if (!(k > 0 ? mask[idx - 1] : (mlo && mlo[ij]))) b |= 16u;
    if (!(k + 1 < nz ? mask[idx + 1] : (mhi && mhi[ij]))) b |= 32u;
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
        const unsigned ex = exposed_bits(a.mask, idx, i, j, k, a.nx, a.ny, a.nz, a.mlo, a.mhi);
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

// K7 in word form (adi_mask_core.h): NC cells of a z line per thread (NC = 2: 16 bytes per output field and thread).
template <int NC, bool CS, int MINB>
__global__ void __launch_bounds__(256, MINB) k_build_packs_v(const PackArgs a)
{
    const size_t nw = (size_t)a.nx * a.ny * a.nz / NC, stride = (size_t)gridDim.x * blockDim.x;
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t raw = t < nw ? ldcells<NC>(a.mask + t * NC) : 0u;
    while (t < nw) {   // the next iteration's mask bytes are on their way while this one's fields are stored
        const size_t tn = t + stride;
        const uint32_t rawn = tn < nw ? ldcells<NC>(a.mask + tn * NC) : 0u;
        build_packs_cells<NC, CS>(a, t * NC, raw);
        t = tn;
        raw = rawn;
    }
}

__global__ void k_exposed_mask(const uint8_t *__restrict__ mask, uint8_t *__restrict__ out, int face,
                               int nx, int ny, int nz, const uint8_t *__restrict__ mlo,
                               const uint8_t *__restrict__ mhi)
{
    const size_t n = (size_t)nx * ny * nz;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n;
         idx += (size_t)gridDim.x * blockDim.x) {
        const int k = (int)(idx % nz);
        const size_t ij = idx / nz;
        const int j = (int)(ij % ny);
        const int i = (int)(ij / ny);
        out[idx] = (uint8_t)((exposed_bits(mask, idx, i, j, k, nx, ny, nz, mlo, mhi) >> face) & 1u);
    }
}

// K8: is a bound dense coefficient field "surface-only"?  viol[ax] counts the active, non-Dirichlet cells
// with both neighbours along axis ax whose coefficient is not +0.0 (bit pattern): when it is 0 the sweep
// may read the field at exposed cells only and take +0.0 elsewhere -- bit-identical rows.
struct SparseCheckArgs {
    const double *coeff[3];
    const uint8_t *code[3];
    unsigned long long *viol;  // [3]
};

__global__ void k_check_sparse(const SparseCheckArgs a, size_t n)
{
    unsigned long long bad[3] = {0ull, 0ull, 0ull};
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n;
         idx += (size_t)gridDim.x * blockDim.x) {
#pragma unroll
        for (int ax = 0; ax < 3; ++ax) {
            if (!a.coeff[ax]) continue;
            const unsigned c = a.code[ax][idx];
            const unsigned both = (CB_XM | CB_XP) << (2 * ax);
            if ((c & CB_SELF) && !(c & CB_DIR) && (c & both) == both &&
                __double_as_longlong(a.coeff[ax][idx]) != 0ll)
                bad[ax]++;
        }
    }
#pragma unroll
    for (int ax = 0; ax < 3; ++ax)
        if (bad[ax]) atomicAdd(a.viol + ax, bad[ax]);
}
#endif  // ADI_CART_MISC_KERNELS (K7)

}  // namespace adi
