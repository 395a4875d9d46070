// ctr_fk_kernels_extra.cuh — callers of the hot path that live in the same kernels, as templates over the
// arithmetic variant (FAST: instantiated in ctr_fk_kernels_extra.cu; STRICT: in ctr_fk_kernels_strict.cu, built with
// -fmad=false):
//   * batched finite-difference Jacobians, every calcJacobian form (robot_i.h:901-1072, :41-56):
//     1 + (T-1) + S solves per configuration, all inside one thread so the joint vector is read
//     once (per-thread routine: jacobian_config in ctr_fk_core.cuh);
//   * batched calcStability / calcDegreeStability (robot_i.h:756-899).
#pragma once

#include "ctr_fk_kernels.cuh"
#include "ctr_fk_internal.h"

namespace ctrk {

template <class L, bool WITH_SAMPLE, bool FAST>
__global__ void __launch_bounds__(kBlock)
fk_jacobian_kernel(const __grid_constant__ KRobot rb, const __grid_constant__ JacArgs a)
{
    constexpr int NJ = L::NJ;
    const int64_t slot = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (slot >= a.B) return;
    // ArcLengthStep form: sample counts differ per configuration; `order` lists every window of 4096
    // configurations longest-sweep-first (bin_local_kernel), so a warp sees near-uniform counts
    const int64_t cfg = a.order ? (int64_t)a.order[slot] : slot;
    double q[NJ];
    load_jv<L>(a.jv, cfg, q);
    jacobian_config<L, WITH_SAMPLE, FAST>(rb, q, a.mode, a.step, a.nsamples, a.i_sample, a.fd_trans, a.fd_rot,
                       a.jac + cfg * (NJ * 6), a.jac_sample ? a.jac_sample + cfg * (NJ * 6) : nullptr,
                       a.tip ? a.tip + cfg * 12 : nullptr, a.sample ? a.sample + cfg * 12 : nullptr,
                       a.n_out ? a.n_out + cfg : nullptr, a.given != 0);
}

// WITH_SAMPLE picks the instantiation family: the iSample forms (two Jacobians, sample product next to the tip product)
// and the tip-only forms are compiled in separate translation units (each takes minutes for 14 layouts)
template <bool FAST, bool WITH_SAMPLE>
cudaError_t launch_jacobian_t(int key, const KRobot& rb, const JacArgs& a, cudaStream_t st)
{
    if (a.B <= 0) return cudaSuccess;
    const unsigned grid = (unsigned)((a.B + kBlock - 1) / kBlock);
    switch (key) {
#define X(S, A, Bq, C) \
    case S * 1000 + A * 100 + Bq * 10 + C: \
        fk_jacobian_kernel<Lay<S, A, Bq, C>, WITH_SAMPLE, FAST><<<grid, kBlock, 0, st>>>(rb, a); \
        break;
        CTRK_FOR_EACH_LAYOUT_K(X)
#undef X
        default: return cudaErrorInvalidValue;
    }
    count_launch();
    return cudaGetLastError();
}

// ---- calcStability / calcDegreeStability (robot_i.h:756-899) -------------------------------------
// For every tube 1..T-1 the tip angle is scanned and the tube's base angle watched, which only
// needs the tip->base sweep (no frames).  The scan variable is advanced and compared exactly like
// the reference (`scan += step; scan <= end`), because that decides how many scan points exist.
template <class L, bool FAST>
__device__ __forceinline__ bool base_angles(const KRobot& rb, const double* q, double arc_step, double* alb)
{
    Ctl<L> c;
    setup_control<L>(rb, q, 0, arc_step, 0, c);
    if (c.status) return false;
    if (FAST) solve_fast<L, false>(rb, c, nullptr, alb, NoSample{});
    else      solve_strict<L, false>(rb, c, nullptr, alb, NoSample{});
    return true;
}

// The number of sweeps per configuration varies wildly (the scan length depends on the tip angle,
// and an unstable configuration stops at the first decreasing step), so the kernel is persistent
// with a ticket queue like the shooting solve: the unit of lock-step work is ONE sweep; a lane
// that has finished its configuration draws the next index from a global counter.
template <class L, bool FAST>
__global__ void __launch_bounds__(kBlock)
fk_stability_kernel(const __grid_constant__ KRobot rb, const __grid_constant__ StabArgs a)
{
    constexpr int T = L::T, NJ = L::NJ;
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const double kPi = 3.14159265358979323846, kTwoPi = 2.0 * 3.14159265358979323846;
    const double dstep = a.angle_step;
    const double qnan = __longlong_as_double(0x7ff8000000000000ll);

    int64_t cfg = -1;
    double  q[NJ];
    int     t = 1;                       // tube being scanned
    double  scan = 0.0, scan_end = 0.0, prev = 0.0, smallest = 0.0;
    bool    have_prev = false, checked = false;
#pragma unroll
    for (int j = 0; j < NJ; ++j) q[j] = 0.0;
    bool drained = false;

    auto start_tube = [&](int tube) {                                   // robot_i.h:782-791
        const int jp = 2 + L::sec_of(tube) + tube;
        double tipa = 0.0;
#pragma unroll
        for (int j = 0; j < NJ; ++j) tipa = (j == jp) ? q[j] : tipa;
        if (tipa <= kPi) { scan = 0.0; scan_end = tipa; }
        else             { scan = tipa; scan_end = kTwoPi; }
        have_prev = false;
    };
    auto finish = [&](int stable_flag) {
        if (a.stable) a.stable[cfg] = stable_flag;
        if (a.degree) {
            double deg;
            if (stable_flag < 0) deg = qnan;
            else if (checked) deg = atan2(dstep, smallest);             // :886, :897
            else deg = a.min_degree + rb.ztol;                          // :898
            a.degree[cfg] = deg;
        }
        cfg = -1;
    };

    for (;;) {
        // ---- refill idle lanes ---------------------------------------------------------------
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
                    cfg = idx;
                    load_jv<L>(a.jv, cfg, q);
                    q[0] = 0.0;                                          // robot_i.h:762
                    checked = !(a.min_degree > rb.ztol);                 // :834
                    smallest = 1.7976931348623157e308;
                    t = 1;
                    if (T > 1) start_tube(1);
                }
            }
            if ((int64_t)first + __popc(m) >= a.B) drained = true;
        }
        if (__ballot_sync(full, cfg >= 0) == 0) break;
        // ---- one sweep on every busy lane -------------------------------------------------------
        if (cfg >= 0) {
            if (T == 1) {                                               // nothing to scan
                finish(1);
            } else {
                const int jp = 2 + L::sec_of(t) + t;
                double qq[NJ], cur[T];
#pragma unroll
                for (int j = 0; j < NJ; ++j) qq[j] = (j == jp) ? scan : q[j];
                if (!base_angles<L, FAST>(rb, qq, a.arc_step, cur)) {
                    finish(-1);
                } else {
                    double cur_t = 0.0;
#pragma unroll
                    for (int u = 0; u < T; ++u) cur_t = (u == t) ? cur[u] : cur_t;
                    bool unstable = false;
                    if (have_prev) {                                    // :802-813 / :870-891
                        const double rel = cur_t - prev;
                        if (rel < smallest) { checked = true; smallest = rel; }
                        unstable = rel < 0.0;
                    }
                    prev = cur_t;
                    have_prev = true;
                    if (unstable) {
                        finish(0);
                    } else {
                        scan = x_add(scan, dstep);
                        if (!(scan <= scan_end)) {                      // this tube is done
                            ++t;
                            if (t >= T) finish(1);
                            else start_tube(t);
                        }
                    }
                }
            }
        }
    }
}

template <bool FAST>
cudaError_t launch_stability_t(int key, const KRobot& rb, const StabArgs& a, int sm_count, cudaStream_t st)
{
    if (a.B <= 0) return cudaSuccess;
    cudaError_t e = cudaSuccess;
    switch (key) {
#define X(S, A, Bq, C) \
    case S * 1000 + A * 100 + Bq * 10 + C: { \
        int per_sm = 0; \
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fk_stability_kernel<Lay<S, A, Bq, C>, FAST>, kBlock, 0); \
        if (e != cudaSuccess) return e; \
        if (per_sm < 1) per_sm = 1; \
        int64_t want = (a.B + kBlock - 1) / kBlock; \
        int64_t cap = (int64_t)per_sm * sm_count; \
        fk_stability_kernel<Lay<S, A, Bq, C>, FAST><<<(unsigned)(want < cap ? want : cap), kBlock, 0, st>>>(rb, a); \
        break; }
        CTRK_FOR_EACH_LAYOUT_K(X)
#undef X
        default: return cudaErrorInvalidValue;
    }
    count_launch();
    return cudaGetLastError();
}

}  // namespace ctrk
