// ctr_sampler.h — counter-based joint sampler shared by host and device.
// Distribution of getRandomJointValues (robot_i.h:1308-1317, section_i.h:328-337):
//   jv[0] and every alpha = nextafter(2*pi, 0) * u,  phi_s = Length_s * u,  u ~ U[0,1).
// The reference draws u from std::mt19937 sequentially; here u for (configuration i, slot j) is
// a pure function of (seed, i, j) so that any shard of a batch can be produced independently
// on any GPU or on the host, bit-identically.
#pragma once
#include <stdint.h>
#include <math.h>
#include "../../include/ctr_fk.h"

#if defined(__CUDACC__)
#define CTR_SAMPLER_HD __host__ __device__ __forceinline__
#else
#define CTR_SAMPLER_HD inline
#endif

namespace ctrk {

struct SamplerTable {
    int    njv;
    double scale[CTR_FK_MAX_JV];
    // extension-length scale policies of transrotscale.h (applied to the Phi slots only)
    int    policy;                       // CTR_FK_SCALE_*
    double p0, p1;                       // gamma | (scale_min, scale_max)
    int    lut_size;                     // GammaCorrection<Real, LookUpSize>; 0 = exact pow
    int    nsec;
    int    phi_slot[CTR_FK_MAX_SECTIONS];
    double seg_len[CTR_FK_MAX_SECTIONS];
    double total_len;
};

CTR_SAMPLER_HD uint64_t splitmix64(uint64_t z)
{
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

// uniform in [0,1) with 53 random bits
CTR_SAMPLER_HD double sampler_u01(uint64_t seed, uint64_t config, uint32_t slot)
{
    const uint64_t a = splitmix64(seed + 0x9E3779B97F4A7C15ull * config);
    const uint64_t b = splitmix64(a + (uint64_t)slot);
    return (double)(b >> 11) * (1.0 / 9007199254740992.0);
}

inline void make_sampler_table(const ctr_robot_desc* d, SamplerTable* t)
{
    const double two_pi_m = nextafter(2.0 * M_PI, 0.0);
    int j = 0;
    t->scale[j++] = two_pi_m;
    t->total_len = 0.0;
    for (int s = 0; s < d->nsections; ++s) {
        t->phi_slot[s] = j;
        t->seg_len[s] = d->sec[s].length;
        t->total_len += d->sec[s].length;             // m_OverallMaxLength, transrotscale.h:170
        t->scale[j++] = d->sec[s].length;
        for (int k = 0; k < d->sec[s].ntube; ++k) t->scale[j++] = two_pi_m;
    }
    t->njv = j;
    t->nsec = d->nsections;
    t->policy = 0; t->p0 = 1.0; t->p1 = 1.0; t->lut_size = 0;
}

// non-contracted arithmetic so that host and device produce the same bits
CTR_SAMPLER_HD double s_mul(double a, double b)
{
#ifdef __CUDA_ARCH__
    return __dmul_rn(a, b);
#else
    return a * b;
#endif
}
CTR_SAMPLER_HD double s_add(double a, double b)
{
#ifdef __CUDA_ARCH__
    return __dadd_rn(a, b);
#else
    return a + b;
#endif
}
CTR_SAMPLER_HD double s_div(double a, double b)
{
#ifdef __CUDA_ARCH__
    return __ddiv_rn(a, b);
#else
    return a / b;
#endif
}

CTR_SAMPLER_HD double s_fma(double a, double b, double c)
{
#ifdef __CUDA_ARCH__
    return __fma_rn(a, b, c);
#else
    return fma(a, b, c);
#endif
}

// ---- pow(u, g) for 0 <= u < 1 with the SAME bits on host and device ------------------------------------------
// libm's pow differs between glibc and CUDA by up to 2 ulp, which would break the promise at the top of this file
// for the gamma policy.  So the sampler carries its own: only +, *, fma and bit operations, in a fixed order, never
// contracted.  u = 2^e m with m in [sqrt(1/2), sqrt(2)); ln m = 2 s (1 + s^2/3 + s^4/5 + ...), s = (m-1)/(m+1), with s
// and the product in double-double (the exponent of the result is g ln u, up to ~700 in magnitude: its absolute error
// is the relative error of the result); exp by ln 2 reduction and a degree-13 polynomial.  Within 2 ulp of the
// correctly rounded value (tests/test_host.py::test_sampler_pow_accuracy, against mpmath).
struct SDD { double h, l; };
CTR_SAMPLER_HD SDD s_two_sum(double a, double b)
{
    const double t = s_add(a, b), bb = s_add(t, -a);
    SDD r; r.h = t; r.l = s_add(s_add(a, -s_add(t, -bb)), s_add(b, -bb));
    return r;
}
CTR_SAMPLER_HD SDD s_two_prod(double a, double b)
{
    SDD r; r.h = s_mul(a, b); r.l = s_fma(a, b, -r.h);
    return r;
}
CTR_SAMPLER_HD SDD s_dd_add(SDD a, SDD b)
{
    SDD t = s_two_sum(a.h, b.h);
    t.l = s_add(t.l, s_add(a.l, b.l));
    return s_two_sum(t.h, t.l);
}
CTR_SAMPLER_HD SDD s_dd_mul(SDD a, SDD b)
{
    SDD t = s_two_prod(a.h, b.h);
    t.l = s_add(t.l, s_add(s_mul(a.h, b.l), s_mul(a.l, b.h)));
    return s_two_sum(t.h, t.l);
}
CTR_SAMPLER_HD double sampler_pow(double u, double g)
{
    if (g == 0.0) return 1.0;
    if (!(u > 0.0)) return g > 0.0 ? 0.0 : 1.0 / 0.0;
    if (u == 1.0) return 1.0;
    // u = 2^e m, m in [sqrt(1/2), sqrt(2)); subnormal u scaled first
    union { double d; uint64_t b; } v;
    v.d = u;
    int e = 0;
    if ((v.b >> 52) == 0) { v.d = s_mul(u, 18014398509481984.0); e = -54; }        // 2^54
    e += (int)((v.b >> 52) & 0x7ff) - 1023;
    v.b = (v.b & 0x000fffffffffffffull) | 0x3ff0000000000000ull;                    // m in [1, 2)
    double m = v.d;
    if (m > 1.4142135623730951) { m = s_mul(m, 0.5); e += 1; }
    // s = (m - 1) / (m + 1) in double-double: m - 1 is exact, m + 1 as a two-sum
    const double num = s_add(m, -1.0);
    const SDD    den = s_two_sum(m, 1.0);
    const double q1  = s_div(num, den.h);
    const SDD    p1  = s_two_prod(q1, den.h);
    const double rem = s_add(s_add(s_add(num, -p1.h), -p1.l), -s_mul(q1, den.l));
    const SDD    sq  = s_two_sum(q1, s_div(rem, den.h));
    const SDD    s2  = s_dd_mul(sq, sq);
    // tail of the series in plain double: it is < 1e-2 of the leading 1
    const double z = s2.h;
    double P = 1.0 / 23.0;
    P = s_fma(P, z, 1.0 / 21.0); P = s_fma(P, z, 1.0 / 19.0); P = s_fma(P, z, 1.0 / 17.0); P = s_fma(P, z, 1.0 / 15.0);
    P = s_fma(P, z, 1.0 / 13.0); P = s_fma(P, z, 1.0 / 11.0); P = s_fma(P, z, 1.0 / 9.0);  P = s_fma(P, z, 1.0 / 7.0);
    P = s_fma(P, z, 1.0 / 5.0);
    SDD third; third.h = 0.33333333333333331; third.l = 1.8503717077085941e-17;     // 1/3 as a double-double
    SDD Pz; Pz.h = s_mul(P, z); Pz.l = 0.0;
    const SDD tail = s_dd_mul(s2, s_dd_add(third, Pz));                             // s^2/3 + s^4/5 + ...
    SDD one; one.h = 1.0; one.l = 0.0;
    SDD lnm = s_dd_mul(sq, s_dd_add(one, tail));
    lnm.h = s_mul(lnm.h, 2.0); lnm.l = s_mul(lnm.l, 2.0);
    // ln u = e ln 2 + ln m;  e * ln2_hi is exact (ln2_hi has 32 trailing zero bits)
    const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10;
    SDD eln; eln.h = s_mul((double)e, ln2_hi); eln.l = s_mul((double)e, ln2_lo);
    const SDD lnu = s_dd_add(s_two_sum(eln.h, eln.l), lnm);
    // y = g ln u
    SDD gg; gg.h = g; gg.l = 0.0;
    const SDD y = s_dd_mul(lnu, gg);
    if (y.h < -745.2) return 0.0;
    if (y.h > 709.7) return 1.0 / 0.0;
    // exp(y): y = k ln 2 + r
    const double kd = floor(s_fma(y.h, 1.44269504088896338700e+00, 0.5));
    const double rh = s_fma(-kd, ln2_hi, y.h);                                      // exact
    const SDD    r  = s_two_sum(rh, s_add(y.l, -s_mul(kd, ln2_lo)));
    const double x = r.h;
    double Q = 1.0 / 6227020800.0;                                                   // 1/13!
    Q = s_fma(Q, x, 1.0 / 479001600.0); Q = s_fma(Q, x, 1.0 / 39916800.0); Q = s_fma(Q, x, 1.0 / 3628800.0);
    Q = s_fma(Q, x, 1.0 / 362880.0);    Q = s_fma(Q, x, 1.0 / 40320.0);    Q = s_fma(Q, x, 1.0 / 5040.0);
    Q = s_fma(Q, x, 1.0 / 720.0);       Q = s_fma(Q, x, 1.0 / 120.0);      Q = s_fma(Q, x, 1.0 / 24.0);
    Q = s_fma(Q, x, 1.0 / 6.0);         Q = s_fma(Q, x, 0.5);
    // exp(x) = 1 + x + x^2 Q, summed smallest first; then the low part of r: exp(r.h + r.l) = exp(r.h) (1 + r.l)
    const SDD    x2 = s_two_prod(x, x);
    const double sm = s_fma(x2.h, Q, s_mul(x2.l, Q));
    const SDD    t  = s_two_sum(x, sm);
    const SDD    ex = s_two_sum(1.0, t.h);
    double val = s_add(ex.h, s_add(ex.l, t.l));
    val = s_fma(val, r.l, val);
    // scale by 2^k (two steps keep the factors normal; results below the normal range round once more)
    const int k = (int)kd;
    const int k1 = k / 2, k2 = k - k1;
    union { double d; uint64_t b; } f1, f2;
    f1.b = (uint64_t)(k1 + 1023) << 52;
    f2.b = (uint64_t)(k2 + 1023) << 52;
    return s_mul(s_mul(val, f1.d), f2.d);
}

// GammaCorrection::scaleTrans (transrotscale.h:106-122): pow(u, gamma), or its LookUpSize-entry
// linearly interpolated table (table = pow(linspace(0,1), gamma) after one set_scale call).  pow is sampler_pow
// (same bits on host and device; within 2 ulp of the reference's std::pow).
CTR_SAMPLER_HD double scale_gamma(double u, double gamma, int lut)
{
    if (lut <= 1) return sampler_pow(u, gamma);
    const double x  = u * (double)(lut - 1);
    const int    i0 = (int)floor(x);
    const int    i1 = (int)ceil(x);
    const double md = x - (double)i0;
    // table nodes as Eigen 3.3's setLinSpaced(n, 0, 1) produces them: i * (1/(n-1)), the last one exactly 1
    const double st = s_div(1.0, (double)(lut - 1));
    const double f0 = sampler_pow(i0 == lut - 1 ? 1.0 : s_mul((double)i0, st), gamma);
    const double f1 = sampler_pow(i1 == lut - 1 ? 1.0 : s_mul((double)i1, st), gamma);
    return s_add(s_mul(f1, md), s_mul(f0, s_add(1.0, -md)));
}

// One configuration's joint vector under a scale policy.  ProportionalScale
// (transrotscale.h:273-291) is sequential over the sections of one configuration (each draw
// narrows the interval of the next so that the total extension hits a goal length) and takes the
// goal of configuration i from the LAST-section draw of configuration i-1 (:286-288); with a
// counter-based generator that draw is recomputed instead of carried, so configurations stay
// independent.  Configuration 0 takes its goal from an extra slot (the reference uses std::rand).
CTR_SAMPLER_HD void sample_config(const SamplerTable& t, uint64_t seed, uint64_t config, double* jv)
{
    for (int j = 0; j < t.njv; ++j) jv[j] = t.scale[j] * sampler_u01(seed, config, (uint32_t)j);
    if (t.policy == 1) {                       // CTR_FK_SCALE_GAMMA
        for (int s = 0; s < t.nsec; ++s) {
            const int j = t.phi_slot[s];
            jv[j] = t.seg_len[s] * scale_gamma(sampler_u01(seed, config, (uint32_t)j), t.p0, t.lut_size);
        }
    } else if (t.policy == 2) {                // CTR_FK_SCALE_PROPORTIONAL
        const double ug = (config == 0) ? sampler_u01(seed, 0, 63u)
                                        : sampler_u01(seed, config - 1, (uint32_t)t.phi_slot[t.nsec - 1]);
        const double inst = s_add(s_mul(t.p1 - t.p0, ug), t.p0);
        double goal = s_mul(t.total_len, inst), maxrem = t.total_len;
        for (int s = 0; s < t.nsec; ++s) {
            const int    j = t.phi_slot[s];
            const double L = t.seg_len[s];
            const double u = sampler_u01(seed, config, (uint32_t)j);
            const double lo = fmax(s_div(s_add(s_add(goal, -maxrem), L), L), 0.0);
            const double hi = fmin(s_div(goal, L), 1.0);
            const double v  = s_add(s_mul(s_add(hi, -lo), u), lo);
            goal = s_add(goal, -s_mul(v, L));
            maxrem = s_add(maxrem, -L);
            jv[j] = s_mul(L, fmin(fmax(v, 0.0), 1.0));   // Erl::clamp, :290
        }
    }
}

}  // namespace ctrk
