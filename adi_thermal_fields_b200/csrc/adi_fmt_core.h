// adi_fmt_core.h -- correctly rounded fp64 -> decimal text, one value per thread, for the ASCII
// output path (reference: vtk_writer.py:4-9 "{float(v):.6e}", waam_from_stl_v7_mm.py:203-206
// "{float(T[i,j,k]):.6g}").  Python formats the EXACT binary value, ties to even; so does this:
//   1. |v| * 10^(P-1-k) in double-double arithmetic (tabulated 10^j, adi_pow10_tab.h) gives the P
//      leading digits n and the remainder r to ~1e-24;
//   2. unless |r - 1/2| < 1e-9 the rounding direction is decided; otherwise (exact ties such as
//      1234567.5, and the one-in-1e9 near ties) an exact big-integer comparison of
//      m*2^e against (2n+1)*10^q/2 decides it.
// Host + device: tests/ run the same code on the CPU (host_emulation.cpp) against Python's own
// formatting.  No libc formatting on either side.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "adi_pow10_tab.h"

#ifdef __CUDACC__
#define FMT_HD __host__ __device__ __forceinline__
#define FMT_HD_NOINLINE __host__ __device__ __noinline__
#else
#define FMT_HD inline
#define FMT_HD_NOINLINE inline
#endif

#ifdef __CUDA_ARCH__
#define FMT_MUL(a, b) __dmul_rn((a), (b))
#define FMT_ADD(a, b) __dadd_rn((a), (b))
#define FMT_FMA(a, b, c) __fma_rn((a), (b), (c))
#define FMT_LD(p) __ldg(p)
#else
#define FMT_MUL(a, b) ((a) * (b))
#define FMT_ADD(a, b) ((a) + (b))
#define FMT_FMA(a, b, c) fma((a), (b), (c))
#define FMT_LD(p) (*(p))
#endif

namespace adifmt {

enum { CLS_FINITE = 0, CLS_ZERO = 1, CLS_INF = 2, CLS_NAN = 3 };
enum { FMT_E6 = 0, FMT_G6 = 1 };  // "%.6e" (7 significant digits) / "%.6g" (6 significant digits)
enum { MAX_TEXT = 14 };           // "-1.234567e-100"

struct Dec {
    uint32_t n;  // P significant digits, 10^(P-1) <= n < 10^P
    int k;       // decimal exponent of the leading digit
    int cls;
    int neg;
};

FMT_HD uint64_t bits_of(double v)
{
    uint64_t b;
    memcpy(&b, &v, 8);
    return b;
}

FMT_HD int clz64(uint64_t x)
{
#ifdef __CUDA_ARCH__
    return __clzll((long long)x);
#else
    return __builtin_clzll(x);
#endif
}

// (ah + al) * (bh + bl) -> (h, l)
FMT_HD void dd_mul(double ah, double al, double bh, double bl, double &h, double &l)
{
    const double p = FMT_MUL(ah, bh);
    double e = FMT_FMA(ah, bh, -p);
    e = FMT_FMA(ah, bl, e);
    e = FMT_FMA(al, bh, e);
    const double s = FMT_ADD(p, e);
    l = FMT_ADD(e, -FMT_ADD(s, -p));
    h = s;
}

// a * 10^j as a double-double; tab[j - ADI_POW10_MIN] = {hi, lo}
FMT_HD void scale10(double a, int j, const double *tab, double &h, double &l)
{
    double ah = a, al = 0.0;
    if (j > ADI_POW10_MAX) {
        const double *t = tab + 2 * (200 - ADI_POW10_MIN);
        dd_mul(ah, al, FMT_LD(t), FMT_LD(t + 1), ah, al);
        j -= 200;
    } else if (j < ADI_POW10_MIN) {
        const double *t = tab + 2 * (-150 - ADI_POW10_MIN);
        dd_mul(ah, al, FMT_LD(t), FMT_LD(t + 1), ah, al);
        j += 150;
    }
    const double *t = tab + 2 * (j - ADI_POW10_MIN);
    dd_mul(ah, al, FMT_LD(t), FMT_LD(t + 1), h, l);
}

// ---- exact comparison (rare path) ----------------------------------------------------------
struct Big {
    uint32_t w[40];
    int len;
};

FMT_HD void big_set(Big &b, uint64_t v)
{
    b.w[0] = (uint32_t)v;
    b.w[1] = (uint32_t)(v >> 32);
    b.len = b.w[1] ? 2 : (b.w[0] ? 1 : 0);
}

FMT_HD void big_mul_small(Big &b, uint32_t f)
{
    uint64_t carry = 0;
    for (int i = 0; i < b.len; ++i) {
        const uint64_t t = (uint64_t)b.w[i] * f + carry;
        b.w[i] = (uint32_t)t;
        carry = t >> 32;
    }
    if (carry) b.w[b.len++] = (uint32_t)carry;
}

FMT_HD void big_mul_pow5(Big &b, int p)
{
    for (; p >= 13; p -= 13) big_mul_small(b, 1220703125u);  // 5^13
    uint32_t f = 1;
    for (; p > 0; --p) f *= 5u;
    if (f > 1) big_mul_small(b, f);
}

FMT_HD int big_bitlen(const Big &b)
{
    if (b.len == 0) return 0;
    return 32 * b.len - (clz64((uint64_t)b.w[b.len - 1]) - 32);
}

FMT_HD void big_shl(Big &b, int s)
{
    const int ws = s >> 5, bs = s & 31;
    if (b.len == 0 || s == 0) return;
    int nl = b.len + ws + 1;
    for (int i = nl - 1; i >= 0; --i) {
        const int src = i - ws;
        uint32_t lo = 0, hi = 0;
        if (src >= 0 && src < b.len) hi = b.w[src];
        if (src - 1 >= 0 && src - 1 < b.len) lo = b.w[src - 1];
        b.w[i] = bs ? ((hi << bs) | (lo >> (32 - bs))) : hi;
    }
    while (nl > 0 && b.w[nl - 1] == 0) --nl;
    b.len = nl;
}

FMT_HD int big_cmp(const Big &x, const Big &y)
{
    if (x.len != y.len) return x.len > y.len ? 1 : -1;
    for (int i = x.len - 1; i >= 0; --i)
        if (x.w[i] != y.w[i]) return x.w[i] > y.w[i] ? 1 : -1;
    return 0;
}

// sign of  a - (n + 1/2) * 10^q   for finite a > 0
FMT_HD_NOINLINE int exact_cmp_half(double a, uint32_t n, int q)
{
    const uint64_t b = bits_of(a);
    const int ef = (int)(b >> 52) & 0x7ff;
    uint64_t m = b & ((1ull << 52) - 1);
    int e;
    if (ef == 0) {
        e = -1074;
    } else {
        m |= 1ull << 52;
        e = ef - 1075;
    }
    Big X, Y;  // compare X * 2^sx with Y, where a*2 = m*2^(e+1), (2n+1)*10^q = (2n+1)*5^q*2^q
    big_set(X, m);
    big_set(Y, 2ull * n + 1ull);
    if (q >= 0) big_mul_pow5(Y, q);
    else big_mul_pow5(X, -q);
    const int sx = e + 1 - q;
    const int bx = big_bitlen(X) + sx, by = big_bitlen(Y);
    if (bx != by) return bx > by ? 1 : -1;
    if (sx >= 0) big_shl(X, sx);
    else big_shl(Y, -sx);
    return big_cmp(X, Y);
}

// ---- rounding to P significant digits --------------------------------------------------------
template <int P>
FMT_HD Dec round_sig(double v, const double *tab)
{
    const double lim_hi = P == 7 ? 1e7 : 1e6;   // 10^P
    const double lim_lo = P == 7 ? 1e6 : 1e5;   // 10^(P-1)
    Dec d;
    // classify from the raw bits: the compiler may turn a sign-masked copy into abs.f64, whose
    // result for a NaN input is unspecified in PTX (a negative NaN kept its sign on sm_100a)
    const uint64_t b = bits_of(v);
    const int ef = (int)(b >> 52) & 0x7ff;
    const uint64_t frac = b & ((1ull << 52) - 1);
    d.neg = (int)(b >> 63);
    d.n = 0;
    d.k = 0;
    if (ef == 0x7ff) {
        d.cls = frac ? CLS_NAN : CLS_INF;
        return d;
    }
    if ((b << 1) == 0) {
        d.cls = CLS_ZERO;
        return d;
    }
    d.cls = CLS_FINITE;
    const double a = fabs(v);
    const int E = ef ? ef - 1023 : (63 - clz64(frac)) - 1074;  // 2^E <= a < 2^(E+1)
    int k = (E * 315653) >> 20;                                // floor(E*log10(2)), may be one low
    double h, l;
    int dir = 0;  // k moves one way only: a product within 1e-24 of a power of ten is accepted as it is
    for (;;) {    // (the rounding below then lands on the same digits from either side)
        scale10(a, P - 1 - k, tab, h, l);
        if (dir >= 0 && (h > lim_hi || (h == lim_hi && l >= 0.0))) {
            ++k;
            dir = 1;
        } else if (dir <= 0 && (h < lim_lo || (h == lim_lo && l < 0.0))) {
            --k;
            dir = -1;
        } else {
            break;
        }
    }
    const double fl = floor(h);
    double r = FMT_ADD(FMT_ADD(h, -fl), l);
    long long n = (long long)fl;
    if (r < 0.0) {
        n -= 1;
        r = FMT_ADD(r, 1.0);
    } else if (r >= 1.0) {
        n += 1;
        r = FMT_ADD(r, -1.0);
    }
    bool up;
    if (fabs(r - 0.5) < 1e-9) {
        const int c = exact_cmp_half(a, (uint32_t)n, k - (P - 1));
        up = c > 0 || (c == 0 && (n & 1));
    } else {
        up = r > 0.5;
    }
    n += up ? 1 : 0;
    if (n >= (long long)lim_hi) {  // 9.999..e(k) rounded up to 10^P
        n /= 10;
        ++k;
    }
    d.n = (uint32_t)n;
    d.k = k;
    return d;
}

FMT_HD int exp_len(int k) { return (k >= 100 || k <= -100) ? 5 : 4; }  // "e+XX" / "e+XXX"

FMT_HD int trailing_zeros6(uint32_t n)  // n has 6 digits, n != 0
{
    int tz = 0;
    while (tz < 5 && n % 10u == 0u) {
        n /= 10u;
        ++tz;
    }
    return tz;
}

// length of the text of one value (without its separator)
template <int FMT>
FMT_HD int text_len(const Dec &d)
{
    if (d.cls == CLS_NAN) return 3;
    if (d.cls == CLS_INF) return 3 + d.neg;
    if (FMT == FMT_E6) {
        if (d.cls == CLS_ZERO) return 12 + d.neg;
        return d.neg + 8 + exp_len(d.k);
    }
    if (d.cls == CLS_ZERO) return 1 + d.neg;
    const int nd = 6 - trailing_zeros6(d.n);  // significant digits kept
    if (d.k < -4 || d.k >= 6) return d.neg + nd + (nd > 1 ? 1 : 0) + exp_len(d.k);
    if (d.k >= 0) {
        const int nf = nd - (d.k + 1);
        return d.neg + (d.k + 1) + (nf > 0 ? 1 + nf : 0);
    }
    return d.neg + 2 + (-d.k - 1) + nd;
}

template <typename CH>
FMT_HD CH *put_exp(CH *p, int k)
{
    *p++ = 'e';
    *p++ = k < 0 ? '-' : '+';
    uint32_t a = (uint32_t)(k < 0 ? -k : k);
    if (a >= 100u) {
        *p++ = (char)('0' + a / 100u);
        a %= 100u;
    }
    *p++ = (char)('0' + a / 10u);
    *p++ = (char)('0' + a % 10u);
    return p;
}

// writes the text of one value at p (text_len bytes); returns the end.  CH = char or volatile-free
// shared-memory char.
template <int FMT, typename CH>
FMT_HD CH *emit(const Dec &d, CH *p)
{
    if (d.cls == CLS_NAN) {
        *p++ = 'n'; *p++ = 'a'; *p++ = 'n';
        return p;
    }
    if (d.neg) *p++ = '-';
    if (d.cls == CLS_INF) {
        *p++ = 'i'; *p++ = 'n'; *p++ = 'f';
        return p;
    }
    if (FMT == FMT_E6) {
        uint32_t n = d.cls == CLS_ZERO ? 0u : d.n;
        const int k = d.cls == CLS_ZERO ? 0 : d.k;
        uint32_t dig[7];
#pragma unroll
        for (int i = 6; i >= 0; --i) {
            dig[i] = n % 10u;
            n /= 10u;
        }
        *p++ = (char)('0' + dig[0]);
        *p++ = '.';
#pragma unroll
        for (int i = 1; i < 7; ++i) *p++ = (char)('0' + dig[i]);
        return put_exp(p, k);
    }
    if (d.cls == CLS_ZERO) {
        *p++ = '0';
        return p;
    }
    uint32_t n = d.n;
    uint32_t dig[6];
#pragma unroll
    for (int i = 5; i >= 0; --i) {
        dig[i] = n % 10u;
        n /= 10u;
    }
    int nd = 6;
    while (nd > 1 && dig[nd - 1] == 0u) --nd;
    const int k = d.k;
    if (k < -4 || k >= 6) {
        *p++ = (char)('0' + dig[0]);
        if (nd > 1) {
            *p++ = '.';
#pragma unroll
            for (int i = 1; i < 6; ++i)
                if (i < nd) *p++ = (char)('0' + dig[i]);
        }
        return put_exp(p, k);
    }
    if (k >= 0) {
#pragma unroll
        for (int i = 0; i < 6; ++i)
            if (i <= k) *p++ = (char)('0' + dig[i]);
        if (nd > k + 1) {
            *p++ = '.';
#pragma unroll
            for (int i = 1; i < 6; ++i)
                if (i > k && i < nd) *p++ = (char)('0' + dig[i]);
        }
        return p;
    }
    *p++ = '0';
    *p++ = '.';
    for (int z = 0; z < -k - 1; ++z) *p++ = '0';
#pragma unroll
    for (int i = 0; i < 6; ++i)
        if (i < nd) *p++ = (char)('0' + dig[i]);
    return p;
}

// Layout of a whole field: values in Fortran order of the C-order (nx,ny,nz) array (x fastest),
// f = (k*ny + j)*nx + i.  FMT_E6: nine values per line (vtk_writer.py:7-9), the last line may be
// shorter; FMT_G6: one line per (k, j) row of nx values (waam_from_stl_v7_mm.py:203-206).
FMT_HD char separator(int fmt, uint64_t f, int i, int nx, uint64_t N)
{
    if (fmt == FMT_E6) return (f % 9u == 8u || f + 1u == N) ? '\n' : ' ';
    return i == nx - 1 ? '\n' : ' ';
}

// one value end to end (host-side helpers and tests)
template <int FMT>
FMT_HD int format_value(double v, const double *tab, char *out)
{
    const Dec d = round_sig<FMT == FMT_E6 ? 7 : 6>(v, tab);
    return (int)(emit<FMT>(d, out) - out);
}

}  // namespace adifmt
