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

// GammaCorrection::scaleTrans (transrotscale.h:106-122): pow(u, gamma), or its LookUpSize-entry
// linearly interpolated table (table = pow(linspace(0,1), gamma) after one set_scale call).
CTR_SAMPLER_HD double scale_gamma(double u, double gamma, int lut)
{
    if (lut <= 1) return pow(u, gamma);
    const double x  = u * (double)(lut - 1);
    const int    i0 = (int)floor(x);
    const int    i1 = (int)ceil(x);
    const double md = x - (double)i0;
    // table nodes as Eigen 3.3's setLinSpaced(n, 0, 1) produces them: i * (1/(n-1)), the last one exactly 1
    const double st = s_div(1.0, (double)(lut - 1));
    const double f0 = pow(i0 == lut - 1 ? 1.0 : s_mul((double)i0, st), gamma);
    const double f1 = pow(i1 == lut - 1 ? 1.0 : s_mul((double)i1, st), gamma);
    return f1 * md + f0 * (1.0 - md);
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
