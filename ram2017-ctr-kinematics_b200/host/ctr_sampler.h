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
    for (int s = 0; s < d->nsections; ++s) {
        t->scale[j++] = d->sec[s].length;
        for (int k = 0; k < d->sec[s].ntube; ++k) t->scale[j++] = two_pi_m;
    }
    t->njv = j;
}

}  // namespace ctrk
