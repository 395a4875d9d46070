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

// ---- backbone kernel, AoS layout, FAST arithmetic: warp-parallel second pass ---------------------
// The thread-per-configuration second pass above stores each frame 24.5 KB away from its warp
// neighbours' frames and reads the parked samples with one dependent load per step (ncu, round 1:
// 12.9 GB of DRAM traffic for 5.3 GB of output, long-scoreboard stalls dominate).  Here:
//   pass 1  one thread per configuration sweeps tip -> base and parks (u_x,u_y,u_z,alpha) of every
//           sample COMPACTLY at the start of the configuration's own output region
//           ([k][4] doubles = n*32 B contiguous, 32-B aligned full-sector stores);
//   pass 2  each warp takes the CTA's configurations one at a time: the n*32 B block is copied to
//           shared memory with cp.async (16 B per lane per request, all requests in flight at
//           once, double-buffered so the next configuration's block streams in while this one is
//           processed); lane l owns the contiguous samples [l*c, (l+1)*c), c = ceil(n/32).  The
//           running product I_k = E_0 ... E_k is a prefix "sum" over SE(3) composition, computed
//           as lane-local products + a warp scan (5 shuffle rounds) + a second lane-local walk that
//           emits Frame_k = Init Rz(alpha_k+theta) I_k.  A warp writes one configuration's n*96 B
//           contiguously.  Transforms travel as (unit quaternion, translation): 7 doubles per
//           shuffle, 37 FP64 instructions per composition.
// Shared memory: kBbWarps warps x 2 buffers x MaxSamples x 32 B (32 KB per CTA at MaxSamples=256).
constexpr int kBbWarps = 2;
constexpr int kBbBufs = 1;              // staging buffers per warp (2 = prefetch the next block)
constexpr int kBbBlock = kBbWarps * 32;
constexpr int kBbMaxSamples = 1024;     // larger robots use the thread-per-configuration kernel

struct BbShared {
    double  h, theta;
    double  end[kMaxSec];
    int64_t cfg;
    int     nframes;
    int     ok;
};

__device__ __forceinline__ QT qt_shfl_up(const QT& t, int off)
{
    QT r;
    r.w = __shfl_up_sync(0xffffffffu, t.w, off);   r.x = __shfl_up_sync(0xffffffffu, t.x, off);
    r.y = __shfl_up_sync(0xffffffffu, t.y, off);   r.z = __shfl_up_sync(0xffffffffu, t.z, off);
    r.px = __shfl_up_sync(0xffffffffu, t.px, off); r.py = __shfl_up_sync(0xffffffffu, t.py, off);
    r.pz = __shfl_up_sync(0xffffffffu, t.pz, off);
    return r;
}

// One 256-bit store (sm_100: STG.E.ENL2.256): a full 32-B sector per request.  These kernels are
// bound by the number of L2 write requests their scattered stores generate, not by bytes
// (ncu: long-scoreboard stalls on registers still owned by in-flight 16-B stores), so every store
// here carries a whole sector.  `p` must be 32-B aligned.
__device__ __forceinline__ void st256(double* p, double a, double b, double c, double d)
{
    asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}
__device__ __forceinline__ void store_tf256(double* dst, const double* t)
{
    st256(dst, t[0], t[1], t[2], t[3]);
    st256(dst + 4, t[4], t[5], t[6], t[7]);
    st256(dst + 8, t[8], t[9], t[10], t[11]);
}

// parks a sample as one 32-B store
struct ParkCompact {
    double* base;
    __device__ __forceinline__ void operator()(int k, double ux, double uy, double uz, double al) const
    {
        st256(base + 4 * (int64_t)k, ux, uy, uz, al);
    }
};

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

template <class L>
__global__ void __launch_bounds__(kBbBlock, 8)
fk_backbone_warp_kernel(const __grid_constant__ KRobot rb, const __grid_constant__ FkArgs a)
{
    extern __shared__ __align__(16) double sh_park[];       // [kBbWarps][2][maxs*4]
    __shared__ BbShared sh[kBbBlock];
    const int64_t i = (int64_t)blockIdx.x * kBbBlock + threadIdx.x;
    BbShared me;
    me.ok = 0; me.nframes = 0; me.cfg = 0; me.h = 0.0; me.theta = 0.0;
#pragma unroll
    for (int s = 0; s < kMaxSec; ++s) me.end[s] = 0.0;

    // ---- pass 1: one thread per configuration, tip -> base sweep, samples parked in place -----
    if (i < a.B) {
        const int64_t cfg = a.order ? (int64_t)a.order[i] : i;
        double q[L::NJ];
        load_jv<L>(a.jv, cfg, q);
        Ctl<L> c;
        setup_control<L>(rb, q, a.mode, a.step, a.nsamples, c);
        if (c.status) {
            write_invalid<L>(a, cfg);
        } else {
            ParkCompact park;
            park.base = a.backbone + cfg * 12 * (int64_t)rb.maxs;
            double alb[L::T];
            solve_fast<L, false>(rb, c, nullptr, alb, park);
            if (a.n_out) a.n_out[cfg] = c.nframes;
            if (a.step_out) a.step_out[cfg] = c.h;
            if (a.alpha_base) {
#pragma unroll
                for (int t = 0; t < L::T; ++t) a.alpha_base[cfg * L::T + t] = alb[t];
            }
            if (a.status) a.status[cfg] = 0;
            me.ok = 1; me.nframes = c.nframes; me.cfg = cfg; me.h = c.h; me.theta = c.theta;
#pragma unroll
            for (int s = 0; s < L::S; ++s) me.end[s] = c.end[s];
        }
    }
    sh[threadIdx.x] = me;
    __syncthreads();        // parked samples (global) and sh[] are visible to the whole CTA

    // ---- pass 2: one warp per configuration ---------------------------------------------------
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* buf0 = sh_park + (size_t)warp * kBbBufs * 4 * rb.maxs;
    auto prefetch = [&](int j, int which) {
        const BbShared& cf = sh[warp * 32 + j];
        if (cf.ok) {
            const char* src = reinterpret_cast<const char*>(a.backbone + cf.cfg * 12 * (int64_t)rb.maxs);
            char* dst = reinterpret_cast<char*>(buf0 + (size_t)which * 4 * rb.maxs);
            const int bytes = cf.nframes * 32;
            for (int o = lane * 16; o < bytes; o += 32 * 16) cp_async16(dst + o, src + o);
        }
        cp_async_commit();
    };
    if (kBbBufs == 2) prefetch(0, 0);
    for (int j = 0; j < 32; ++j) {
        if (kBbBufs == 2) {
            if (j + 1 < 32) prefetch(j + 1, (j + 1) & 1);  // stream the next block in meanwhile
            else            cp_async_commit();
            cp_async_wait<1>();                             // this configuration's block has landed
        } else {
            prefetch(j, 0);
            cp_async_wait<0>();
        }
        __syncwarp();
        const BbShared& cf = sh[warp * 32 + j];
        if (cf.ok) {                                        // warp-uniform
            const int n   = cf.nframes;
            const int cnt = (n + 31) >> 5;
            const int k0  = lane * cnt;
            const int k1  = (k0 + cnt < n) ? k0 + cnt : n;
            const double* pk = buf0 + (size_t)(kBbBufs == 2 ? (j & 1) : 0) * 4 * rb.maxs;
            double* slot0 = a.backbone + cf.cfg * 12 * (int64_t)rb.maxs;
            const StepConsts kc = make_step_consts(cf.h, rb.ztol);

            QT T = qt_identity();
            bool allhot = true;       // every sample of this lane takes the branch-free forms below
            for (int k = k0; k < k1; ++k) {
                const double2 v0 = *reinterpret_cast<const double2*>(pk + 4 * k);
                const double2 v1 = *reinterpret_cast<const double2*>(pk + 4 * k + 2);
                const double n2 = x_fma(v1.x, v1.x, x_fma(v0.y, v0.y, v0.x * v0.x));
                allhot = allhot && le_nonneg(kc.ztol2, n2) && hi_word(n2 * kc.hh2) < 0x3f700000 &&
                         (hi_word(v1.y + cf.theta) & 0x7fffffff) < 0x40900000;
                QT E;
                se3_step_fast_k(v0.x, v0.y, v1.x, kc, E.w, E.x, E.y, E.z, E.px, E.py, E.pz);
                T = qt_mul(T, E);
            }
            const bool hot_walk = __all_sync(0xffffffffu, allhot);
            // inclusive scan over lanes: T_l <- T_0 * ... * T_l
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const QT o = qt_shfl_up(T, off);
                if (lane >= off) T = qt_mul(o, T);
            }
            QT I = qt_shfl_up(T, 1);
            if (lane == 0) I = qt_identity();
            {   // renormalise the carried rotation (|q| drifts by ~n ulp)
                const double inv = rsqrt(I.w * I.w + I.x * I.x + I.y * I.y + I.z * I.z);
                I.w *= inv; I.x *= inv; I.y *= inv; I.z *= inv;
            }
            if (hot_walk) {
                // branch-free body, two samples per trip: the step and the frame trigonometry of
                // sample k+1 overlap the (serial) product and frame assembly of sample k
#pragma unroll 2
                for (int k = k0; k < k1; ++k) {
                    const double2 v0 = *reinterpret_cast<const double2*>(pk + 4 * k);
                    const double2 v1 = *reinterpret_cast<const double2*>(pk + 4 * k + 2);   // (u_z, alpha_k)
                    const double sxy = x_fma(v0.y, v0.y, v0.x * v0.x);
                    const double m   = x_fma(v1.x, v1.x, sxy) * kc.hh2;
                    double qa, qb, qc;
                    se3_series4(m, qa, qb, qc);
                    const double b2 = kc.h * qb, b = kc.hh * qb;
                    QT E;
                    E.w = qa; E.x = b * v0.x; E.y = b * v0.y; E.z = b * v1.x;
                    const double C = kc.h3 * qc, cz = C * v1.x;
                    E.px = x_fma(cz, v0.x, b2 * E.y);
                    E.py = x_fma(cz, v0.y, -(b2 * E.x));
                    E.pz = x_fma(-C, sxy, kc.h);
                    I = qt_mul(I, E);
                    double shf, chf, F[12];
                    sincos_bounded(0.5 * (v1.y + cf.theta), shf, chf);
                    qt_frame(rb, I, shf, chf, F);
                    store_tf256(slot0 + (int64_t)k * 12, F);
                    if (k == n - 1 && a.tip) store_tf(a.tip + cf.cfg * 12, F);
                }
            } else {
                for (int k = k0; k < k1; ++k) {
                    const double2 v0 = *reinterpret_cast<const double2*>(pk + 4 * k);
                    const double2 v1 = *reinterpret_cast<const double2*>(pk + 4 * k + 2);   // (u_z, alpha_k)
                    QT E;
                    se3_step_fast_k(v0.x, v0.y, v1.x, kc, E.w, E.x, E.y, E.z, E.px, E.py, E.pz);
                    I = qt_mul(I, E);
                    double shf, chf, F[12];
                    fast_sincos(0.5 * (v1.y + cf.theta), shf, chf);
                    qt_frame(rb, I, shf, chf, F);
                    store_tf256(slot0 + (int64_t)k * 12, F);
                    if (k == n - 1 && a.tip) store_tf(a.tip + cf.cfg * 12, F);
                }
            }
            if (a.radius) {
                // robot_i.h:1286-1307, lanes striding over the samples of each section's run
                double* rad = a.radius + cf.cfg * (int64_t)rb.maxs;
                int    begin = 0;
                double rsec  = 0.0;
#pragma unroll
                for (int s = 0; s < L::S; ++s) {
                    if (begin >= n) continue;
                    rsec = rb.rad[s];
                    const double qd = x_div(cf.end[s], cf.h);
                    int stop = (qd >= 0.0 && qd < (double)n) ? (int)qd : n;
                    if (stop < begin) stop = begin;
                    for (int k = begin + lane; k < stop; k += 32) rad[k] = rsec;
                    begin = stop;
                }
                __syncwarp();
                if (lane == 0 && n > 0) rad[n - 1] = rsec;
            }
        }
        __syncwarp();       // everyone is done with this buffer before it is refilled
    }
}

// ---- backbone kernel, AoS layout, FAST arithmetic: checkpointed sweep, frames emitted per run ----
// fk_backbone_warp_kernel moves 32 B per sample out to the output region in pass 1 and back in
// pass 2, then needs two walks over the samples and a warp scan to turn them into frames.  Here
// (scheme: ctr_fk_core.cuh, "frame emission from a checkpointed sweep"):
//   pass 1  one thread per configuration runs the tip solver's sweep (suffix products carried,
//           software-pipelined) and saves 32 records — sweep state + suffix product, 128-160 B —
//           at the start of the configuration's own output region; sample N-1 is parked whole;
//   pass 2  each warp takes the CTA's configurations one at a time; lane l repeats the sweep over
//           its run of ceil(n/32) samples from record l and emits Frame_k as sample k appears
//           (I_k = T inv(S_{k+1}), then I_{k-1} = I_k inv(E_k)).  No shared-memory staging, no
//           scan, no second walk; a warp still writes one configuration's n*96 B region by itself,
//           which keeps DRAM pages open (thread-per-configuration stores do not).
// Round trip per configuration: 33 records (4-5 KB) instead of 16 KB of parked samples at N = 256.
constexpr int kCkMinSamples = 128;      // below this the records are no smaller than parked samples
#ifndef CTRK_BB_EXPERIMENT
#define CTRK_BB_EXPERIMENT 0
#endif
#ifndef CTRK_CK_CTAS
#define CTRK_CK_CTAS 6      // resident CTAs per SM the kernel is compiled for: 168 registers, 12 warps/SM (2-3 % faster than 128 registers x 16 warps; tools/tune_bb.cu)
#endif
#ifndef CTRK_CK_WARPS
#define CTRK_CK_WARPS 2
#endif
constexpr int kCkWarps = CTRK_CK_WARPS;
constexpr int kCkBlock = kCkWarps * 32;

struct EmShared {            // what pass 2 needs to know about a configuration
    double  h, theta;
    double  end[kMaxSec];
    double  T[7];            // total product E_0 ... E_{frames-1}
    int64_t cfg;
    int     n, nframes;      // nframes = 0: nothing to do (invalid configuration or beyond the batch)
    int     stop[kMaxSec];   // radius runs (robot_i.h:1286-1307): section s covers samples [stop[s-1], stop[s])
};

__device__ __forceinline__ void ld256cg(const double* p, double& a, double& b, double& c, double& d)
{
    asm volatile("ld.global.cg.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p) : "memory");
}

template <class L>
struct RecStoreGlobal {      // CkptHook store: one record = EmitLayout<L>::kStride doubles
    double* base;
    template <class SW>
    __device__ __forceinline__ void operator()(int slot, const SW& sw) const
    {
        constexpr int R = EmitLayout<L>::kStride;
        double v[R];
#pragma unroll
        for (int j = 0; j < R; ++j) v[j] = 0.0;
        sw.save_state(v);
        sw.suffix_with_pending(v + EmitLayout<L>::kState);
        double* p = base + (int64_t)slot * R;
#pragma unroll
        for (int g = 0; g < R; g += 4) st256(p + g, v[g], v[g + 1], v[g + 2], v[g + 3]);
    }
};
struct HeadParkGlobal {      // on_sample of pass 1: only sample N-1 (the head step) is parked
    double* dst;
    int     nm1;
    __device__ __forceinline__ void operator()(int k, double ux, double uy, double uz, double al) const
    {
        if (k == nm1) st256(dst, ux, uy, uz, al);
    }
};
struct FrameStoreGlobal {    // emit_run store: 3 full sectors per frame
    double* slot0;
    double* tip;             // nullable
    int     ktip;
    __device__ __forceinline__ void operator()(int k, const double* F) const
    {
#if CTRK_BB_EXPERIMENT == 2      /* timing experiment: frames computed, (almost) never stored */
        if (F[0] == 123.456) store_tf256(slot0 + (int64_t)k * 12, F);
#else
        store_tf256(slot0 + (int64_t)k * 12, F);
#endif
        if (k == ktip && tip) store_tf(tip, F);
    }
};

template <class L>
__global__ void __launch_bounds__(kCkBlock, CTRK_CK_CTAS)
fk_backbone_emit_kernel(const __grid_constant__ KRobot rb, const __grid_constant__ FkArgs a)
{
    constexpr int R = EmitLayout<L>::kStride;    // record stride, doubles
    __shared__ EmShared sh[kCkBlock];
    const int64_t i = (int64_t)blockIdx.x * kCkBlock + threadIdx.x;
    EmShared me;
    me.n = 0; me.nframes = 0; me.cfg = 0; me.h = 0.0; me.theta = 0.0;
#pragma unroll
    for (int s = 0; s < kMaxSec; ++s) { me.end[s] = 0.0; me.stop[s] = 0; }
#pragma unroll
    for (int q = 0; q < 7; ++q) me.T[q] = 0.0;

    // ---- pass 1: one thread per configuration, tip -> base sweep, 32 records saved --------------
    if (i < a.B) {
        const int64_t cfg = a.order ? (int64_t)a.order[i] : i;
        double q[L::NJ];
        load_jv<L>(a.jv, cfg, q);
        Ctl<L> c;
        setup_control<L>(rb, q, a.mode, a.step, a.nsamples, c);
        if (c.status) {
            write_invalid<L>(a, cfg);
        } else {
            double* slot0 = a.backbone + cfg * 12 * (int64_t)rb.maxs;
            const SegPlan plan(c.n, c.nframes);
            CkptHook<RecStoreGlobal<L>> hook(plan, RecStoreGlobal<L>{slot0});
            const HeadParkGlobal park{slot0 + 32 * R, c.n - 1};
            FastSweep<L, true, HeadParkGlobal> sw(rb, c, park);
            sw.run_hooked(hook);
            if (a.n_out) a.n_out[cfg] = c.nframes;
            if (a.step_out) a.step_out[cfg] = c.h;
            if (a.alpha_base) {
#pragma unroll
                for (int t = 0; t < L::T; ++t) a.alpha_base[cfg * L::T + t] = sw.al[t];
            }
            if (a.status) a.status[cfg] = 0;
            me.n = c.n; me.nframes = c.nframes; me.cfg = cfg; me.h = c.h; me.theta = c.theta;
            me.T[0] = sw.Qw; me.T[1] = sw.Qx; me.T[2] = sw.Qy; me.T[3] = sw.Qz;
            me.T[4] = sw.Px; me.T[5] = sw.Py; me.T[6] = sw.Pz;
            int begin = 0;
#pragma unroll
            for (int s = 0; s < L::S; ++s) {
                me.end[s] = c.end[s];
                // robot_i.h:1295-1304: section s fills [begin, min(floor(End_s/h), n))
                int stop = begin;
                if (begin < c.nframes) {
                    const double qd = x_div(c.end[s], c.h);
                    stop = (qd >= 0.0 && qd < (double)c.nframes) ? (int)qd : c.nframes;
                    if (stop < begin) stop = begin;
                }
                me.stop[s] = stop;
                begin = stop;
            }
        }
    }
    sh[threadIdx.x] = me;
    __syncthreads();        // records (global) and sh[] are visible to the whole CTA

    // ---- pass 2: one warp per configuration ---------------------------------------------------
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#if CTRK_BB_EXPERIMENT == 1      /* timing experiment: pass 1 only */
    if (a.B > 0) return;
#endif
    for (int j = 0; j < 32; ++j) {
        const EmShared& cf = sh[warp * 32 + j];
        if (cf.nframes == 0) continue;                      // warp-uniform
        const int n = cf.nframes;
        const SegPlan plan(cf.n, n);
        const int k0 = lane * plan.cnt;
        const int k1 = (k0 + plan.cnt < n) ? k0 + plan.cnt : n;
        double* slot0 = a.backbone + cf.cfg * 12 * (int64_t)rb.maxs;
        {
            Ctl<L> c;
            c.theta = cf.theta; c.arc_end = 0.0; c.h = cf.h; c.n = cf.n; c.nframes = n; c.status = 0;
#pragma unroll
            for (int s = 0; s < L::S; ++s) {
                c.end[s] = cf.end[s];
                const double d = x_sub(c.end[s], rb.len[s]);     // as setup_control
                c.start[s] = (0.0 < d) ? d : 0.0;
            }
#pragma unroll
            for (int t = 0; t < L::T; ++t) c.al[t] = 0.0;
            double rec[R], head[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
            for (int g = 0; g < R; ++g) rec[g] = 0.0;
            if (k0 < n) {
                const bool holds_head = (k1 - 1 == cf.n - 1);
                if (holds_head) ld256cg(slot0 + 32 * R, head[0], head[1], head[2], head[3]);
                if ((holds_head ? cf.n - 2 : k1 - 1) >= k0) {
#pragma unroll
                    for (int g = 0; g < R; g += 4) ld256cg(slot0 + (int64_t)lane * R + g, rec[g], rec[g + 1], rec[g + 2], rec[g + 3]);
                }
            }
            __syncwarp();       // every lane holds its record: the frames may now overwrite that region
            QT T;
            T.w = cf.T[0]; T.x = cf.T[1]; T.y = cf.T[2]; T.z = cf.T[3]; T.px = cf.T[4]; T.py = cf.T[5]; T.pz = cf.T[6];
            emit_run<L>(rb, c, plan, lane, rec, head, T,
                        FrameStoreGlobal{slot0, a.tip ? a.tip + cf.cfg * 12 : nullptr, n - 1});
        }
        if (a.radius) {
            double* rad = a.radius + cf.cfg * (int64_t)rb.maxs;
            int    begin = 0;
            double rsec  = 0.0;
#pragma unroll
            for (int s = 0; s < L::S; ++s) {
                if (begin >= n) continue;
                rsec = rb.rad[s];
                const int stop = cf.stop[s];
                for (int k = begin + lane; k < stop; k += 32) rad[k] = rsec;
                begin = stop;
            }
            __syncwarp();
            if (lane == 0) rad[n - 1] = rsec;
        }
    }
}

// ---- backbone kernel, SoA layout [MaxS][12][B], FAST arithmetic ---------------------------------
// With configurations along the fastest axis a thread-per-configuration kernel already writes
// coalesced (a warp stores 32 x 8 B contiguous per matrix entry), so nothing needs to be parked or
// handed to other lanes: the thread sweeps twice.  Sweep 1 is the tip solver (total product T);
// sweep 2 repeats the sweep from the tip and emits Frame_k = Init Rz(alpha_k+theta) I_k with
// I_{ktip} = T and I_{k-1} = I_k inv(E_k), one sample late so the frame arithmetic overlaps the
// next sweep step (FastSweep::run_emit).  No shared memory, no scratch traffic at all: DRAM sees
// the joint values and the output.
struct FrameStoreSoa {
    double* base;            // backbone + cfg
    int64_t B;
    double* tip;             // nullable
    int     ktip;
    __device__ __forceinline__ void operator()(int k, const double* F) const
    {
        double* p = base + (int64_t)k * 12 * B;
#pragma unroll
        for (int j = 0; j < 12; ++j) p[j * B] = F[j];
        if (k == ktip && tip) store_tf(tip, F);
    }
};

template <class L>
__global__ void __launch_bounds__(kBlock, 4)
fk_backbone_soa_kernel(const __grid_constant__ KRobot rb, const __grid_constant__ FkArgs a)
{
    const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (i >= a.B) return;
    const int64_t cfg = a.order ? (int64_t)a.order[i] : i;
    double q[L::NJ];
    load_jv<L>(a.jv, cfg, q);
    Ctl<L> c;
    setup_control<L>(rb, q, a.mode, a.step, a.nsamples, c);
    if (c.status) { write_invalid<L>(a, cfg); return; }

    QT T;
    {
        FastSweep<L, true, NoSample> s1(rb, c, NoSample{});
        s1.run();
        T.w = s1.Qw; T.x = s1.Qx; T.y = s1.Qy; T.z = s1.Qz; T.px = s1.Px; T.py = s1.Py; T.pz = s1.Pz;
        if (a.alpha_base) {
#pragma unroll
            for (int t = 0; t < L::T; ++t) a.alpha_base[cfg * L::T + t] = s1.al[t];
        }
    }
    {
        const FrameStoreSoa fs{a.backbone + cfg, a.B, a.tip ? a.tip + cfg * 12 : nullptr, c.nframes - 1};
        const PendingFrameEmitter<FrameStoreSoa> em{&rb, make_step_consts(c.h, rb.ztol), c.theta, fs, T, c.nframes - 1,
                                                    -1, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
        FastSweep<L, false, PendingFrameEmitter<FrameStoreSoa>> s2(rb, c, em);
        s2.run_emit();
    }
    if (a.radius) fill_radius<L>(rb, c, a.radius + cfg, a.B);
    if (a.n_out) a.n_out[cfg] = c.nframes;
    if (a.step_out) a.step_out[cfg] = c.h;
    if (a.status) a.status[cfg] = 0;
}

#define CTRK_LAUNCH_CASE(KERNEL, FASTV, S, A, Bq, C)                                              \
    case S * 1000 + A * 100 + Bq * 10 + C:                                                        \
        KERNEL<Lay<S, A, Bq, C>, FASTV><<<grid, kBlock, 0, st>>>(rb, a);                          \
        break;

}  // namespace ctrk
