// ctr_fk_kernels.cuh — the solver kernels: one GPU thread carries one configuration's whole
// tip-to-base sweep in registers (see ctr_fk_core.cuh).  Instantiated for FAST=true in
// ctr_fk_kernels_fast.cu and FAST=false in ctr_fk_kernels_strict.cu (the latter built with
// -fmad=false so that its arithmetic matches the CPU oracle operation by operation).
//
// Data layout in HBM
//   jv        [B][NJ] double, configuration-major (the reference's JV vector, back to back)
//   tip       [B][12] double, column-major 3x4 per configuration
//   backbone  [B][maxs][12] (AoS, drop-in for m_CentreLineTransform[k]) or [maxs][12][B] (SoA)
//   radius    [B][maxs] / [maxs][B]
// Per-thread global traffic of the tip kernel is NJ*8 B in + 96 B out (+ a few optional scalars)
// against ~35 k FP64 instructions of work, so the kernel is FP64-pipe bound by three orders of
// magnitude; loads and stores are issued as plain per-thread vectors (a warp's 32 rows are
// contiguous in memory, every fetched sector is fully used).
#pragma once

#include "ctr_fk_launch.h"

namespace ctrk {

template <class L>
__device__ __forceinline__ void load_jv(const double* __restrict__ jv, int64_t cfg, double* q)
{
    const double* src = jv + cfg * L::NJ;
    if ((L::NJ % 2) == 0) {
        const double2* s2 = reinterpret_cast<const double2*>(src);   // rows are 16-B aligned
#pragma unroll
        for (int j = 0; j < L::NJ / 2; ++j) {
            const double2 v = __ldg(s2 + j);
            q[2 * j] = v.x; q[2 * j + 1] = v.y;
        }
    } else {
#pragma unroll
        for (int j = 0; j < L::NJ; ++j) q[j] = __ldg(src + j);
    }
}

__device__ __forceinline__ void store_tf(double* dst, const double* t)
{
    double2* d2 = reinterpret_cast<double2*>(dst);                   // 96-B rows, 16-B aligned
#pragma unroll
    for (int j = 0; j < 6; ++j) d2[j] = make_double2(t[2 * j], t[2 * j + 1]);
}

template <class L>
__device__ __forceinline__ void write_invalid(const FkArgs& a, int64_t cfg)
{
    const double qnan = __longlong_as_double(0x7ff8000000000000ll);
    if (a.tip) {
#pragma unroll
        for (int j = 0; j < 12; ++j) a.tip[cfg * 12 + j] = qnan;
    }
    if (a.n_out) a.n_out[cfg] = 0;
    if (a.step_out) a.step_out[cfg] = qnan;
    if (a.alpha_base) {
#pragma unroll
        for (int t = 0; t < L::T; ++t) a.alpha_base[cfg * L::T + t] = qnan;
    }
    if (a.status) a.status[cfg] = 1;
}

// ---- tip-only kernel ---------------------------------------------------------------------------
// Register budget: 4 CTAs of 128 threads per SM (16 warps) leaves 128 registers per thread.  Measured
// on B200 (tools/tune_tip.cu): capping registers for 20 or 24 warps/SM is 3-5 % slower — the sweep
// is bound by FP64 issue, not by latency that more warps could hide.
template <class L, bool FAST>
__global__ void __launch_bounds__(kBlock, FAST ? 4 : 1)
fk_tip_kernel(const __grid_constant__ KRobot rb, const __grid_constant__ FkArgs a)
{
    const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (i >= a.B) return;
    const int64_t cfg = a.order ? (int64_t)a.order[i] : i;

    double q[L::NJ];
    load_jv<L>(a.jv, cfg, q);
    Ctl<L> c;
    setup_control<L>(rb, q, a.mode, a.step, a.nsamples, c);
    if (c.status) { write_invalid<L>(a, cfg); return; }

    double tip[12];
    double alb[L::T];
    if (FAST) solve_fast<L, true>(rb, c, tip, a.alpha_base ? alb : nullptr, NoSample{});
    else      solve_strict<L, true>(rb, c, a.tip ? tip : nullptr, a.alpha_base ? alb : nullptr, NoSample{});

    if (a.tip) store_tf(a.tip + cfg * 12, tip);
    if (a.radius) {
        if (a.soa) fill_radius<L>(rb, c, a.radius + cfg, a.B);
        else       fill_radius<L>(rb, c, a.radius + cfg * (int64_t)rb.maxs, 1);
    }
    if (a.n_out) a.n_out[cfg] = c.nframes;
    if (a.step_out) a.step_out[cfg] = c.h;
    if (a.alpha_base) {
#pragma unroll
        for (int t = 0; t < L::T; ++t) a.alpha_base[cfg * L::T + t] = alb[t];
    }
    if (a.status) a.status[cfg] = 0;
}

// ---- backbone kernel ---------------------------------------------------------------------------
// Pass 1 (tip -> base): the sweep; every sample's (u_x, u_y, u_z, alpha) is parked in the first
// four doubles of that sample's own slot of the output buffer.  Pass 2 (base -> tip): read them
// back, extend the running product I_k = I_{k-1} * Exp(u_k, h) and overwrite the slot with
// Frame_k = Init * Rz(alpha_k + theta) * I_k (robot_i.h:705-717).  No extra scratch memory.
struct ParkSample {
    double* base;        // slot of sample 0, component 0
    int64_t kstride;     // doubles between consecutive samples
    int64_t cstride;     // doubles between consecutive components of one sample
    __device__ __forceinline__ void operator()(int k, double ux, double uy, double uz, double al) const
    {
        double* p = base + (int64_t)k * kstride;
        p[0] = ux; p[cstride] = uy; p[2 * cstride] = uz; p[3 * cstride] = al;
    }
};

template <class L, bool FAST>
__global__ void __launch_bounds__(kBlock)
fk_backbone_kernel(const __grid_constant__ KRobot rb, const __grid_constant__ FkArgs a)
{
    const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (i >= a.B) return;
    const int64_t cfg = a.order ? (int64_t)a.order[i] : i;

    double q[L::NJ];
    load_jv<L>(a.jv, cfg, q);
    Ctl<L> c;
    setup_control<L>(rb, q, a.mode, a.step, a.nsamples, c);
    if (c.status) { write_invalid<L>(a, cfg); return; }

    ParkSample park;
    if (a.soa) { park.base = a.backbone + cfg; park.kstride = 12 * a.B; park.cstride = a.B; }
    else       { park.base = a.backbone + cfg * 12 * (int64_t)rb.maxs; park.kstride = 12; park.cstride = 1; }

    double alb[L::T];
    if (FAST) solve_fast<L, false>(rb, c, nullptr, alb, park);
    else      solve_strict<L, false>(rb, c, nullptr, alb, park);

    double I[12], F[12];
    tf_set_identity(I);
    const double h = c.h;
    for (int k = 0; k < c.nframes; ++k) {
        double* p = park.base + (int64_t)k * park.kstride;
        const double ux = p[0], uy = p[park.cstride], uz = p[2 * park.cstride], al = p[3 * park.cstride];
        double e[12];
        tf_step_strict(ux, uy, uz, h, rb.ztol, e);
        tf_mul(I, e, I);
        double sn, cn;
        if (FAST) fast_sincos(al + c.theta, sn, cn);
        else      sincos(al + c.theta, &sn, &cn);
        tf_finish(rb.ini, I, sn, cn, F);
        if (a.soa) {
#pragma unroll
            for (int j = 0; j < 12; ++j) p[j * park.cstride] = F[j];
        } else {
            store_tf(p, F);
        }
    }
    if (a.tip) store_tf(a.tip + cfg * 12, F);
    if (a.radius) {
        if (a.soa) fill_radius<L>(rb, c, a.radius + cfg, a.B);
        else       fill_radius<L>(rb, c, a.radius + cfg * (int64_t)rb.maxs, 1);
    }
    if (a.n_out) a.n_out[cfg] = c.nframes;
    if (a.step_out) a.step_out[cfg] = c.h;
    if (a.alpha_base) {
#pragma unroll
        for (int t = 0; t < L::T; ++t) a.alpha_base[cfg * L::T + t] = alb[t];
    }
    if (a.status) a.status[cfg] = 0;
}

#define CTRK_LAUNCH_CASE(KERNEL, FASTV, S, A, Bq, C)                                              \
    case S * 1000 + A * 100 + Bq * 10 + C:                                                        \
        KERNEL<Lay<S, A, Bq, C>, FASTV><<<grid, kBlock, 0, st>>>(rb, a);                          \
        break;

}  // namespace ctrk
