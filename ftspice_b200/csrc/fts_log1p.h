// fts_log1p.h -- ln(1 + u) for u >= 0 (and NaN / +inf), the damping function's ln_1p
// (src/engine/newtons_method.rs:73-77: gamma/k * signum(d) * ln_1p(k * |d|), argument never negative).
//
// One straight-line path of ~40 instructions instead of the two divergent ~60-instruction paths of CUDA's log1p (the
// lanes of a warp sit on both sides of its 0.667 threshold in most Newton iterations):
//     w = 1 + u (rounded), c = u - (w - 1) (what the rounding lost), w = 2^k * m, m in [1, 2)
//     (1/c_i, log c_i) from a 128-entry table on the top 7 mantissa bits; r = m / c_i - 1 by one FMA, |r| <= 2^-8
//     (entry 0 is (1, 0): r = m - 1 in [0, 2^-7), exact, so that small arguments lose nothing)
//     1 + u = 2^k * c_i * (1 + r + c / (2^k c_i))   exactly, hence
//     ln(1 + u) = k ln2 + log c_i + r + rc (1 - r) - r^2/2 + r^3/3 - ... - r^8/8   (r^9/9 dropped: < 2^-59 relative)
// summed with a compensated head (Fast2Sum) so that the result carries one final rounding: measured maximum error
// 0.52 ulp against mpmath on 4 M arguments (tests/test_log1p.py) -- CUDA's and glibc's log1p are ~1 ulp functions.
// Every operation is an IEEE add / multiply / FMA: the device result equals the host result of this same header bit
// for bit (tests/test_gpu_parity.py::test_log1p_device_equals_host_bitwise).  Table: log1p_tab.h (generated).
#ifndef FTS_LOG1P_H
#define FTS_LOG1P_H

#if defined(__CUDACC__)
#define FTS_HD __host__ __device__ __forceinline__
#else
#include <math.h>
#define FTS_HD static inline
#endif
#if defined(__CUDA_ARCH__)
#define FTS_D2U(x) ((unsigned long long)__double_as_longlong(x))
#define FTS_U2D(x) __longlong_as_double((long long)(x))
#else
#if !defined(__CUDACC_RTC__)
#include <string.h>
#endif
FTS_HD unsigned long long fts_d2u_(double x) { unsigned long long u; memcpy(&u, &x, 8); return u; }
FTS_HD double fts_u2d_(unsigned long long u) { double x; memcpy(&x, &u, 8); return x; }
#define FTS_D2U(x) fts_d2u_(x)
#define FTS_U2D(x) fts_u2d_(x)
#endif
#define FTS_FMA(a, b, c) fma(a, b, c)

// T: 128 pairs (1/c_i, log c_i) as doubles (the bit patterns of FTS_LOG1P_TABLE), 16-byte aligned
FTS_HD double fts_log1p_pos(double u, const double *T)
{
    const double w = 1.0 + u;
    const double c = u - (w - 1.0);
    const unsigned long long wb = FTS_D2U(w);
    const unsigned int hi = (unsigned int)(wb >> 32);
    const int i = (int)((hi >> 13) & 127u);
    const double m = FTS_U2D((wb & 0x000fffffffffffffULL) | 0x3ff0000000000000ULL);
#if defined(__CUDA_ARCH__)
    const double2 e = *reinterpret_cast<const double2 *>(T + 2 * i);
    const double invc = e.x, logc = e.y;
#else
    const double invc = T[2 * i], logc = T[2 * i + 1];
#endif
    const double r = FTS_FMA(m, invc, -1.0);
    // 2^-k: exponent field 2046 - E (E = 2047, inf / NaN, is overridden below)
    const double s = FTS_U2D((unsigned long long)(0x7fe00000u - (hi & 0x7ff00000u)) << 32);
    double rc = (c * s) * invc;
    rc = FTS_FMA(-rc, r, rc); // ln(1 + r + rc) = ln(1 + r) + rc / (1 + r) + O(rc^2)
    const double kd = (double)((int)(hi >> 20) - 1023);
    const double a = kd * FTS_U2D(FTS_LN2HI_BITS); // exact: 11 x 42 bits
    const double t1 = a + logc;
    const double h = t1 + r;
    double lo = ((a - t1) + logc) + ((t1 - h) + r); // both roundings recovered (a >= log c_i when k >= 1, t1 >= |r| or t1 = 0)
    lo = FTS_FMA(kd, FTS_U2D(FTS_LN2LO_BITS), lo) + rc;
    double p = FTS_FMA(r, 1.0 / 8.0 * -1.0, 1.0 / 7.0);
    p = FTS_FMA(r, p, -1.0 / 6.0);
    p = FTS_FMA(r, p, 1.0 / 5.0);
    p = FTS_FMA(r, p, -1.0 / 4.0);
    p = FTS_FMA(r, p, 1.0 / 3.0);
    p = FTS_FMA(r, p, -1.0 / 2.0);
    double res = h + FTS_FMA(r * r, p, lo);
    if (!(u < FTS_U2D(0x7ff0000000000000ULL))) res = u + u; // +inf -> +inf, NaN -> NaN
    return res;
}

#endif // FTS_LOG1P_H
