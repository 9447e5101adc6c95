// fts_exp.h -- exp(a) and exp(a) - 1 from one argument reduction: the diode and NPN models evaluate both at every
// junction voltage (src/device/diode/model.rs:12-22, src/device/npn/model.rs:24-47: i = Is * exp_m1(v / Vt),
// g = Is / Vt * exp(v / Vt)).
//     k = round(a * 128 / ln2), r = a - k ln2 / 128 (|r| <= 0.0028, two-term reduction), k = 128 e + j
//     exp(a)     = 2^e 2^(j/128) (1 + q),  q = expm1(r) = r + r^2 (1/2 + r/6 + r^2/24 + r^3/120 + r^4/720)
//     exp(a) - 1 = (s - 1) + s q,          s = 2^e 2^(j/128) = sh + sl (table hi + lo), s - 1 by an exact TwoSum
// Both within 1 ulp of the exact value (measured 0.51 ulp for both, tests/test_exp_pair.py); IEEE adds / multiplies / FMAs
// only, so the device returns the host's bits.  |a| > 690 (results near the overflow / subnormal range), +-inf and NaN
// take the platform's exp / expm1.  Table: exp_tab.h (generated).
#ifndef FTS_EXP_H
#define FTS_EXP_H

#include "fts_log1p.h" // FTS_HD, FTS_D2U, FTS_U2D, FTS_FMA

// T: 128 pairs (hi, lo) of 2^(j/128), 16-byte aligned
FTS_HD void fts_exp_pair(double a, const double *T, double *e_out, double *m_out)
{
    const double magic = 6755399441055744.0; // 1.5 * 2^52
    const double z = a * FTS_U2D(FTS_EXP_INVC_BITS);
    const double zs = z + magic;
    const double kd = zs - magic;
    const int ki = (int)(unsigned int)FTS_D2U(zs);
    const double r1 = FTS_FMA(kd, -FTS_U2D(FTS_EXP_CHI_BITS), a); // exact
    const double t2 = -kd * FTS_U2D(FTS_EXP_CLO_BITS);
    const double r = r1 + t2;
    const double rl = (r1 - r) + t2; // what the rounding of r lost
    const int j = ki & 127;
#if defined(__CUDA_ARCH__)
    const double2 t = *reinterpret_cast<const double2 *>(T + 2 * j);
    const double th = t.x, tl = t.y;
#else
    const double th = T[2 * j], tl = T[2 * j + 1];
#endif
    const double scale = FTS_U2D((unsigned long long)(unsigned int)((ki >> 7) + 1023) << 52);
    const double sh = th * scale, sl = tl * scale;
    double p = FTS_FMA(r, 1.0 / 720.0, 1.0 / 120.0);
    p = FTS_FMA(r, p, 1.0 / 24.0);
    p = FTS_FMA(r, p, 1.0 / 6.0);
    p = FTS_FMA(r, p, 0.5);
    const double q2 = FTS_FMA(r * r, p, rl); // expm1(r) - r
    double e = sh + FTS_FMA(sh, q2 + r, sl);
    // exp(a) - 1 = (sh - 1) + sh r + [sh q2 + sl (1 + r)]: s - 1 = d + dl and sh r = ph + pl exactly, the heads summed exactly too
    const double d = sh - 1.0;
    const double bb = d - sh;
    const double dl = (sh - (d - bb)) + (-1.0 - bb);
    const double ph = sh * r;
    const double pl = FTS_FMA(sh, r, -ph);
    const double hs = d + ph;
    const double hb = hs - d;
    const double hl = (d - (hs - hb)) + (ph - hb);
    double m = hs + ((hl + (dl + pl)) + FTS_FMA(sh, q2, FTS_FMA(sl, r, sl)));
    if (a == 0.0) m = a; // exp_m1(-0) = -0
    if (!(fabs(a) <= 690.0)) {
        e = exp(a);
        m = expm1(a);
    }
    *e_out = e;
    *m_out = m;
}

#endif // FTS_EXP_H
