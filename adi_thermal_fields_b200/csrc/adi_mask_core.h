// adi_mask_core.h -- per-thread bodies of the kernels that run once per mask change (a layer birth of the
// deposition loop, waam_from_stl_v7_mm.py:487-550): neighbour code, its transposed copies for the x / y sweeps,
// and precompute_coeff_packs_unified (adi3d_gpu_coeff.py:31-110).  Word-at-a-time forms of the one-cell-per-thread
// kernels k_build_code / k_transpose_code / k_build_packs (adi_cart.cuh), which stay as the path for grids whose
// nz or base addresses do not allow aligned word access.  __host__ __device__, so that tests/ run the very same
// code on the CPU (csrc/host_emulation.cpp) against the one-cell forms and the oracle.
#pragma once
#include <stdint.h>
#include <string.h>

#include "adi_core.h"

namespace adi {

// Word forms of the mask logic (k_build_code_v / k_build_packs_v, adi_cart.cuh): one 32-bit word = 4 cells that
// follow each other along z (little-endian: byte 0 = lowest z).  nzbytes() turns the mask bytes into 0xff / 0x00
// ("cell active") so that any non-zero mask byte counts, as `if (mask[idx])` does.
ADI_HD uint32_t nzbytes(uint32_t w)
{
#if defined(__CUDA_ARCH__)
    return __vcmpne4(w, 0u);
#else
    uint32_t r = 0;
    for (int b = 0; b < 4; ++b)
        if ((w >> (8 * b)) & 0xffu) r |= 0xffu << (8 * b);
    return r;
#endif
}

// z- / z+ neighbours of the 4 cells of s (nzbytes form): the cells shifted by one byte, with the last cell of
// the word below (prev, nzbytes form; 0 at the lower end of the line or 0xff000000 for an active cell of the
// adjacent slab) / the first cell of the word above coming in.
ADI_HD uint32_t zminus4(uint32_t s, uint32_t prev) { return (s << 8) | (prev >> 24); }
ADI_HD uint32_t zplus4(uint32_t s, uint32_t next) { return (s >> 8) | (next << 24); }

// Neighbour codes of 4 cells from nzbytes-form words: s = the cells, xm .. zp = the neighbour across each face
// (0 outside the grid), d = Dirichlet mask of the axis' pack.  Byte b = the code k_build_code gives cell b.
ADI_HD uint32_t code4(uint32_t s, uint32_t xm, uint32_t xp, uint32_t ym, uint32_t yp, uint32_t zm, uint32_t zp,
                      uint32_t d)
{
    const uint32_t K = 0x01010101u;
    return s & ((CB_SELF * K) | (xm & (CB_XM * K)) | (xp & (CB_XP * K)) | (ym & (CB_YM * K)) | (yp & (CB_YP * K)) |
                (zm & (CB_ZM * K)) | (zp & (CB_ZP * K)) | (d & (CB_DIR * K)));
}

// PRMT: byte i of the result = byte (nibble i of sel) of the 8 bytes (x = 0..3, y = 4..7).
ADI_HD uint32_t bperm(uint32_t x, uint32_t y, uint32_t sel)
{
#if defined(__CUDA_ARCH__)
    return __byte_perm(x, y, sel);
#else
    const uint64_t v = ((uint64_t)y << 32) | x;
    uint32_t r = 0;
    for (int i = 0; i < 4; ++i) r |= (uint32_t)((v >> (8 * ((sel >> (4 * i)) & 7u))) & 0xffu) << (8 * i);
    return r;
#endif
}

// 4 x 4 byte transpose: rows w[0..3] (4 bytes each) -> t[q] = (w[0].byte q, w[1].byte q, w[2].byte q, w[3].byte q).
ADI_HD void transpose4x4(const uint32_t w[4], uint32_t t[4])
{
    const uint32_t a = bperm(w[0], w[1], 0x5140), b = bperm(w[0], w[1], 0x7362);   // (w0.0 w1.0 w0.1 w1.1), (w0.2 w1.2 w0.3 w1.3)
    const uint32_t c = bperm(w[2], w[3], 0x5140), e = bperm(w[2], w[3], 0x7362);
    t[0] = bperm(a, c, 0x5410);
    t[1] = bperm(a, c, 0x7632);
    t[2] = bperm(b, e, 0x5410);
    t[3] = bperm(b, e, 0x7632);
}

// 4- and 16-byte accesses (aligned by the launch conditions).
ADI_HD uint32_t ld4(const uint8_t *p)
{
#if defined(__CUDA_ARCH__)
    return *reinterpret_cast<const uint32_t *>(p);
#else
    uint32_t v;
    memcpy(&v, p, 4);
    return v;
#endif
}
ADI_HD void st4(uint8_t *p, uint32_t v)
{
#if defined(__CUDA_ARCH__)
    *reinterpret_cast<uint32_t *>(p) = v;
#else
    memcpy(p, &v, 4);
#endif
}
ADI_HD void ld16(const uint8_t *p, uint32_t w[4])
{
#if defined(__CUDA_ARCH__)
    const uint4 v = *reinterpret_cast<const uint4 *>(p);
    w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
#else
    memcpy(w, p, 16);
#endif
}
ADI_HD void st16(uint8_t *p, const uint32_t w[4])
{
#if defined(__CUDA_ARCH__)
    *reinterpret_cast<uint4 *>(p) = make_uint4(w[0], w[1], w[2], w[3]);
#else
    memcpy(p, w, 16);
#endif
}
ADI_HD double mul_rn(double a, double b)
{
#if defined(__CUDA_ARCH__)
    return __dmul_rn(a, b);
#else
    return a * b;
#endif
}
ADI_HD double div_rn(double a, double b)
{
#if defined(__CUDA_ARCH__)
    return __ddiv_rn(a, b);
#else
    return a / b;
#endif
}

// index + 1 of the highest non-zero byte of a word (0: none)
ADI_HD int top_byte(uint32_t w)
{
#if defined(__CUDA_ARCH__)
    return 4 - (__clz((int)w) >> 3);
#else
    return w ? 4 - (__builtin_clz(w) >> 3) : 0;
#endif
}

// ---- K0 in word form: the neighbour codes of the 16 cells idx .. idx+15 of one z line ------------------------
// (nz % 16 == 0, idx % 16 == 0, 16-byte aligned arrays).  A run without an active cell costs one load and one store.
// Returns z + 1 of the highest active cell of the run (0: none): the kernel reduces it to the top of the part, above
// which the z sweep has nothing to solve (launch_sweep_zt); likewise the x planes that hold an active cell bound the
// x sweep (adi_cart_step).
ADI_HD int build_code16(const uint8_t *mask, const uint8_t *dirm, uint8_t *code, size_t idx, int nx, int ny, int nz,
                         const uint8_t *mlo, const uint8_t *mhi, int *xplane = nullptr)
{
    uint32_t s[4], out[4] = {0u, 0u, 0u, 0u};
    int top = 0;
    ld16(mask + idx, s);
    if ((s[0] | s[1] | s[2] | s[3]) != 0u) {
        const int k = (int)(idx % (size_t)nz);
        const size_t ij = idx / (size_t)nz;
        const int j = (int)(ij % (size_t)ny);
        const int i = (int)(ij / (size_t)ny);
        if (xplane) *xplane = i;     // the run's x index, for the x extent of the part (written only when a cell is active)
        const size_t snx = (size_t)ny * nz;
        uint32_t xm[4] = {0u, 0u, 0u, 0u}, xp[4] = {0u, 0u, 0u, 0u}, ym[4] = {0u, 0u, 0u, 0u}, yp[4] = {0u, 0u, 0u, 0u},
                 d[4] = {0u, 0u, 0u, 0u};
        if (i > 0) ld16(mask + idx - snx, xm);
        if (i + 1 < nx) ld16(mask + idx + snx, xp);
        if (j > 0) ld16(mask + idx - nz, ym);
        if (j + 1 < ny) ld16(mask + idx + nz, yp);
        if (dirm) ld16(dirm + idx, d);
        // across a slab boundary the neighbour is the adjacent rank's mask plane
        const bool lo = k > 0 ? mask[idx - 1] != 0 : (mlo && mlo[ij]);
        const bool hi = k + 16 < nz ? mask[idx + 16] != 0 : (mhi && mhi[ij]);
#pragma unroll
        for (int q = 0; q < 4; ++q) s[q] = nzbytes(s[q]);
        uint32_t prev = lo ? 0xff000000u : 0u;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint32_t next = q < 3 ? s[q + 1] : (hi ? 0xffu : 0u);
            out[q] = code4(s[q], nzbytes(xm[q]), nzbytes(xp[q]), nzbytes(ym[q]), nzbytes(yp[q]), zminus4(s[q], prev),
                           zplus4(s[q], next), nzbytes(d[q]));
            prev = s[q];
            if (s[q]) top = k + 4 * q + top_byte(s[q]);
        }
    }
    st16(code + idx, out);
    return top;
}

// ---- K0t in word form: dst[(b*nz + c)*npad + r] = src[b*sb + r*sr + c] by tiles of 128 (r) x 128 (c) bytes ---
// 256 threads per tile.  tr_load: a thread takes 4 x 4 byte blocks (4 rows r, one word of 4 columns c: a warp reads
// 128 contiguous bytes of a row), transposes them in registers (PRMT) and leaves the words in S[c][r / 4] with the
// word column rotated by c / 4 (no bank conflicts either way); tr_store: a warp writes 128 contiguous bytes of a
// dst row.  Needs nz % 4 == 0 (so sb, sr and every row start are word aligned), npad % 4 == 0, word-aligned arrays.
// Rows r >= n read as code 0, so the padding columns n .. npad-1 are written as zeros.
struct TrArgs {
    const uint8_t *src;
    uint8_t *dst;
    int n, nz, npad;
    size_t sb, sr;
};

ADI_HD void tr_load(const TrArgs &a, uint32_t *S, int tid, int c0, int r0, int b)
{
    const int cb = tid & 31, wp = tid >> 5;
    const int c = c0 + 4 * cb;
#pragma unroll
    for (int it = 0; it < 4; ++it) {
        const int rb = wp + 8 * it;
        uint32_t w[4], t[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int r = r0 + 4 * rb + q;
            w[q] = (r < a.n && c < a.nz) ? ld4(a.src + (size_t)b * a.sb + (size_t)r * a.sr + c) : 0u;
        }
        transpose4x4(w, t);
#pragma unroll
        for (int q = 0; q < 4; ++q) S[(4 * cb + q) * 32 + ((rb + cb) & 31)] = t[q];
    }
}

ADI_HD void tr_store(const TrArgs &a, const uint32_t *S, int tid, int c0, int r0, int b)
{
    const int rb = tid & 31, wp = tid >> 5;
    const int r = r0 + 4 * rb;
#pragma unroll
    for (int it = 0; it < 16; ++it) {
        const int cl = wp + 8 * it;
        const int c = c0 + cl;
        if (c < a.nz && r < a.npad)
            st4(a.dst + ((size_t)b * a.nz + c) * a.npad + r, S[cl * 32 + ((rb + (cl >> 2)) & 31)]);
    }
}

// ---- K7 in word form: precompute_coeff_packs_unified for the NC (2 or 4) cells idx .. idx+NC-1 of one z line ----
// (nz % NC == 0, NC-byte aligned mask, 16-byte aligned outputs).  Same operations per exposed face, in the same
// order, as k_build_packs; cells without an exposed face (almost all of them) cost no field read.  NC = 2 is the
// form the library launches: a thread stores 16 bytes per field, so a warp's store covers whole 32-byte sectors
// (with NC = 4 each of the two 16-byte stores of a thread fills half a sector: twice the L1 -> L2 sector traffic,
// ncu profiles/r04d: 32 sectors per request, l1tex 68 % busy, 7.9 against 4.6 ms at 1024^3).
struct PackArgs {
    const uint8_t *mask;
    int nx, ny, nz;
    const uint8_t *mlo, *mhi;  // mask planes of the adjacent z slabs (NULL: domain boundary)
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

template <int NC>
ADI_HD uint32_t ldcells(const uint8_t *p)   // NC mask bytes in the low bytes of a word
{
    if (NC == 4) return ld4(p);
#if defined(__CUDA_ARCH__)
    return (uint32_t)__ldg(reinterpret_cast<const unsigned short *>(p));
#else
    uint16_t v;
    memcpy(&v, p, 2);
    return v;
#endif
}

template <int NC, bool CS>
ADI_HD void stcells(double *p, const double *v)
{
#if defined(__CUDA_ARCH__)
    if (CS) {
        __stcs(reinterpret_cast<double2 *>(p), make_double2(v[0], v[1]));
        if (NC == 4) __stcs(reinterpret_cast<double2 *>(p) + 1, make_double2(v[2], v[3]));
    } else {
        reinterpret_cast<double2 *>(p)[0] = make_double2(v[0], v[1]);
        if (NC == 4) reinterpret_cast<double2 *>(p)[1] = make_double2(v[2], v[3]);
    }
#else
    for (int b = 0; b < NC; ++b) p[b] = v[b];
#endif
}

// raw: the NC mask bytes of the cells (ldcells<NC>(a.mask + idx)), loaded by the caller one iteration ahead
template <int NC = 2, bool CS = false>
ADI_HD void build_packs_cells(const PackArgs &a, size_t idx, uint32_t raw)
{
    const uint32_t s = nzbytes(raw);
    uint32_t ex[6] = {0u, 0u, 0u, 0u, 0u, 0u};   // byte b of ex[f]: cell b active and exposed on face f (:38-55)
    if (s) {
        const int k = (int)(idx % (size_t)a.nz);
        const size_t ij = idx / (size_t)a.nz;
        const int j = (int)(ij % (size_t)a.ny);
        const int i = (int)(ij / (size_t)a.ny);
        const size_t snx = (size_t)a.ny * a.nz;
        const bool lo = k > 0 ? a.mask[idx - 1] != 0 : (a.mlo && a.mlo[ij]);
        const bool hi = k + NC < a.nz ? a.mask[idx + NC] != 0 : (a.mhi && a.mhi[ij]);
        ex[0] = s & ~(i > 0 ? nzbytes(ldcells<NC>(a.mask + idx - snx)) : 0u);
        ex[1] = s & ~(i + 1 < a.nx ? nzbytes(ldcells<NC>(a.mask + idx + snx)) : 0u);
        ex[2] = s & ~(j > 0 ? nzbytes(ldcells<NC>(a.mask + idx - a.nz)) : 0u);
        ex[3] = s & ~(j + 1 < a.ny ? nzbytes(ldcells<NC>(a.mask + idx + a.nz)) : 0u);
        ex[4] = s & ~zminus4(s, lo ? 0xff000000u : 0u);
        ex[5] = s & ~((s >> 8) | (hi ? (0xffu << (8 * (NC - 1))) : 0u));   // zplus4 with the next cell after byte NC-1
    }
#pragma unroll
    for (int ax = 0; ax < 3; ++ax) {
        double c[NC], q[NC];
#pragma unroll
        for (int b = 0; b < NC; ++b) c[b] = q[b] = 0.0;
        if ((ex[2 * ax] | ex[2 * ax + 1]) != 0u) {
#pragma unroll
            for (int b = 0; b < NC; ++b)
#pragma unroll
                for (int sd = 0; sd < 2; ++sd) {
                    const int f = 2 * ax + sd;
                    if ((ex[f] >> (8 * b)) & 1u) {
                        if (a.h_kind[f]) {
                            const double h = a.h_kind[f] == 2 ? a.h_field[f][idx + b] : a.h_scalar[f];
                            c[b] += div_rn(mul_rn(h, a.A), a.Ccell);   // (:99) h*A/Ccell
                        }
                        if (a.q_kind[f]) {
                            const double qv = a.q_kind[f] == 2 ? a.q_field[f][idx + b] : a.q_scalar[f];
                            q[b] += div_rn(mul_rn(qv, a.A), a.Ccell);  // (:111)
                        }
                    }
                }
        }
        if (a.coeff[ax]) stcells<NC, CS>(a.coeff[ax] + idx, c);
        if (a.qout[ax]) stcells<NC, CS>(a.qout[ax] + idx, q);
    }
}

}  // namespace adi
