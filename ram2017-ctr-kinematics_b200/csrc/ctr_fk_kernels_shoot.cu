// ctr_fk_kernels_shoot.cu — base-angle shooting solve (reference sweep and RK4 forms), all 14 section layouts.
#include "ctr_fk_kernels_extra.cuh"

namespace ctrk {

// ---- base-angle shooting solve ----------------------------------------------
// Sweeps per configuration vary (2-3 from a nearby guess, ~6 from the naive start x = beta, a handful more where the
// line search has to backtrack).  The kernel is persistent: grid = resident CTAs of the device; every lane owns one
// configuration at a time and, when it finishes, draws the next index from a global ticket counter (one atomicAdd
// per warp and refill round, ranks by ballot).  All lanes of a warp execute one round of the damped Newton iteration
// (ShootState::round: one tangent-linear sweep, uniform trip count in fixed-N mode), so nobody waits for a
// slow neighbour to finish all of its rounds.
template <class L, bool RK4>
__global__ void __launch_bounds__(kBlock)
fk_shoot_kernel(const __grid_constant__ KRobot rb, const __grid_constant__ ShootArgs a)
{
    constexpr int NJ = L::NJ;
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    int64_t cfg = -1;
    ShootState<L, RK4> s;
    {
        double zero[NJ];
#pragma unroll
        for (int j = 0; j < NJ; ++j) zero[j] = 0.0;
        s.start(zero, nullptr);
    }
    bool drained = false;                                   // queue exhausted (warp-uniform)
    for (;;) {
        // ---- refill idle lanes from the ticket queue ---------------------------------------------
        const bool need = (cfg < 0) && !drained;
        const unsigned m = __ballot_sync(full, need);
        if (m) {
            const int leader = __ffs(m) - 1;
            unsigned long long first = 0;
            if (lane == leader) first = atomicAdd(a.ticket, (unsigned long long)__popc(m));
            first = __shfl_sync(full, first, leader);
            if (need) {
                const int64_t idx = (int64_t)first + __popc(m & ((1u << lane) - 1u));
                if (idx < a.B) {
                    cfg = a.order ? (int64_t)a.order[idx] : idx;
                    double qb[NJ], qg[NJ];
                    load_jv<L>(a.jv_base, cfg, qb);
                    if (a.jv_guess) load_jv<L>(a.jv_guess, cfg, qg);
                    s.start(qb, a.jv_guess ? qg : nullptr);
                }
            }
            if ((int64_t)first + __popc(m) >= a.B) drained = true;
        }
        if (__ballot_sync(full, cfg >= 0) == 0) break;       // nothing left for this warp
        // ---- one sweep (one damped-Newton round) on every busy lane --------------------------------
        if (cfg >= 0) {
            const int st = s.round(rb, a.mode, a.step, a.nsamples, a.tol, a.max_iter);
            if (st != ShootState<L, RK4>::kRunning) {
#pragma unroll
                for (int j = 0; j < NJ; ++j) a.jv_tip[cfg * NJ + j] = s.q[j];
                if (a.iters) a.iters[cfg] = s.it;
                if (a.status) a.status[cfg] = st;
                if (a.resid) a.resid[cfg] = (st == 1) ? __longlong_as_double(0x7ff8000000000000ll) : s.res;
                cfg = -1;
            }
        }
    }
}

cudaError_t launch_shoot(int key, const KRobot& rb, const ShootArgs& a, int sm_count, cudaStream_t st)
{
    if (a.B <= 0) return cudaSuccess;
    cudaError_t e = cudaSuccess;
    switch (key) {
#define X(S, A, Bq, C) \
    case S * 1000 + A * 100 + Bq * 10 + C: { \
        int per_sm = 0; \
        e = a.rk4 ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fk_shoot_kernel<Lay<S, A, Bq, C>, true>, kBlock, 0) \
                  : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fk_shoot_kernel<Lay<S, A, Bq, C>, false>, kBlock, 0); \
        if (e != cudaSuccess) return e; \
        if (per_sm < 1) per_sm = 1; \
        int64_t want = (a.B + kBlock - 1) / kBlock; \
        int64_t cap = (int64_t)per_sm * sm_count; \
        if (a.rk4) fk_shoot_kernel<Lay<S, A, Bq, C>, true><<<(unsigned)(want < cap ? want : cap), kBlock, 0, st>>>(rb, a); \
        else       fk_shoot_kernel<Lay<S, A, Bq, C>, false><<<(unsigned)(want < cap ? want : cap), kBlock, 0, st>>>(rb, a); \
        break; }
        CTRK_FOR_EACH_LAYOUT_K(X)
#undef X
        default: return cudaErrorInvalidValue;
    }
    count_launch();
    return cudaGetLastError();
}

}  // namespace ctrk
