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

// the same row through an NVSwitch multicast mapping: one store, replicated to every GPU of the group by the
// switch (multimem.st; SASS: STG.E.128.STRONG.SYS on the multicast address)
__device__ __forceinline__ void store_tf_multicast(double* dst, const double* t)
{
#pragma unroll
    for (int j = 0; j < 6; ++j) {
        const unsigned long long a = (unsigned long long)__double_as_longlong(t[2 * j]);
        const unsigned long long b = (unsigned long long)__double_as_longlong(t[2 * j + 1]);
        asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 2 * j),
                     "f"(__uint_as_float((unsigned)a)), "f"(__uint_as_float((unsigned)(a >> 32))),
                     "f"(__uint_as_float((unsigned)b)), "f"(__uint_as_float((unsigned)(b >> 32)))
                     : "memory");
    }
}
__device__ __forceinline__ void store_tf_gather(const FkArgs& a, int64_t cfg, const double* t)
{
    const int64_t row = (a.gather_row + cfg) * 12;
    if (a.gather_mc) {
        store_tf_multicast(a.gather[0] + row, t);
        return;
    }
#pragma unroll 1
    for (int p = 0; p < a.ngather; ++p) store_tf(a.gather[p] + row, t);
}

template <class L>
__device__ __forceinline__ void write_invalid(const FkArgs& a, int64_t cfg, int maxs)
{
    const double qnan = __longlong_as_double(0x7ff8000000000000ll);
    // frame 0 and radius 0 of an invalid configuration are NaN (its n_out is 0; the rest of its row is left alone)
    if (a.backbone) {
#pragma unroll
        for (int j = 0; j < 12; ++j) {
            if (a.soa) a.backbone[(int64_t)j * a.B + cfg] = qnan;
            else       a.backbone[cfg * (int64_t)maxs * 12 + j] = qnan;
        }
    }
    if (a.radius) a.radius[a.soa ? cfg : cfg * (int64_t)maxs] = qnan;
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
// Register budget: 4 CTAs of 128 threads per SM (16 warps) leaves 128 registers per thread; layouts of 4 and more
// tubes get 3 CTAs and 168 registers (tip_blocks_per_sm, ctr_fk_launch.h).  Measured on B200 (tools/tune_tip.cu):
// capping registers for 20 or 24 warps/SM is 3-5 % slower — the sweep is bound by FP64 issue, not by latency that
// more warps could hide.
// FINE (FAST only): the form of the sweep for finely sampled launches (FastSweep::run_scaled, FkArgs::fine).
template <class L, bool FAST, bool FINE = false>
__global__ void __launch_bounds__(kBlock, FAST ? tip_blocks_per_sm(L::T, L::S) : 1)
fk_tip_kernel(const __grid_constant__ KRobot rb, const __grid_constant__ FkArgs a)
{
    int64_t i;
    if (a.windows) {
        // `order` is sorted inside windows of kBinWindow entries.  CTAs take the windows rank-major:
        // first every window's 128 longest sweeps, then the next 128 ... so the kernel as a whole
        // still runs longest-first and ends on the shortest sweeps (no tail of long stragglers).
        const int64_t w = blockIdx.x % a.windows, r = blockIdx.x / a.windows;
        i = w * kBinWindow + r * kBlock + threadIdx.x;
    } else {
        i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    }
    if (i >= a.B) return;
    const int64_t cfg = a.order ? (int64_t)a.order[i] : i;

    double q[L::NJ];
    load_jv<L>(a.jv, cfg, q);
    Ctl<L> c;
    setup_control<L>(rb, q, a.mode, a.step, a.nsamples, c);
    if (c.status) { write_invalid<L>(a, cfg, rb.maxs); return; }

    double tip[12];
    double alb[L::T];
    if (FAST) solve_fast<L, true, NoSample, FINE>(rb, c, tip, a.alpha_base ? alb : nullptr, NoSample{});
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

// ---- solve + all-gather in one kernel (ctr_fk_batch_gather) --------------------------------------
// The FAST tip solve; every pose goes to row gather_row + cfg of EVERY destination: the gathered buffers of
// all ranks mapped into this process (NVLink peer stores), or one NVSwitch multicast address (multimem.st,
// replicated by the switch).  The stores of one warp ride NVLink while the other warps keep the FP64 pipe busy.
// Rows of a CTA are consecutive unless step-mode binning permuted them, so the CTA stages its 128 poses in
// shared memory (12 KB) and writes them out as whole lines — 16-byte chunks of consecutive lanes are
// consecutive in memory.  One thread storing its own 96-byte row makes 16-byte NVLink writes: fine with one
// peer, 2.3x slower than no collective at all with seven (8 GPUs: 4.80 ms against 2.04 ms); kept for permuted
// rows and as CTR_FK_FLAG_GATHER_ROWWISE for comparison.
template <class L, bool FINE = false>
__global__ void __launch_bounds__(kBlock, tip_blocks_per_sm(L::T, L::S))
fk_tip_gather_kernel(const __grid_constant__ KRobot rb, const __grid_constant__ FkArgs a)
{
    __shared__ double2 stage[kBlock * 6];
    int64_t i;
    if (a.windows) {
        const int64_t w = blockIdx.x % a.windows, r = blockIdx.x / a.windows;
        i = w * kBinWindow + r * kBlock + threadIdx.x;
    } else {
        i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    }
    const bool    active = i < a.B;
    const int64_t cfg = active ? (a.order ? (int64_t)a.order[i] : i) : 0;
    double tip[12];
#pragma unroll
    for (int j = 0; j < 12; ++j) tip[j] = __longlong_as_double(0x7ff8000000000000ll);
    double q[L::NJ];
    if (!a.order) {
        // The CTA's joint vectors are one contiguous run of kBlock * NJ doubles: fetch it as whole lines
        // (consecutive lanes, consecutive 16-byte chunks) and hand every thread its row through shared memory.
        // For device memory this equals the per-thread loads; when `jv` is mapped host memory (the zero-copy
        // form of ctr_fk_batch_host) every byte crosses PCIe exactly once, in full read requests.
        static_assert(L::NJ * 8 <= 6 * 16, "the pose stage holds a CTA's joint vectors");
        const int64_t first = (int64_t)blockIdx.x * kBlock;
        const int64_t left  = a.B - first;
        const int     ndbl  = (int)(left < kBlock ? left : kBlock) * L::NJ;       // doubles of this CTA
        const double* src   = a.jv + first * L::NJ;                              // 16-byte aligned: kBlock * NJ is even
        double*       sj    = reinterpret_cast<double*>(stage);
        for (int ch = threadIdx.x; 2 * ch + 1 < ndbl; ch += kBlock)
            reinterpret_cast<double2*>(sj)[ch] = __ldg(reinterpret_cast<const double2*>(src) + ch);
        if ((ndbl & 1) && threadIdx.x == 0) sj[ndbl - 1] = __ldg(src + ndbl - 1);
        __syncthreads();
        if (active) {
#pragma unroll
            for (int j = 0; j < L::NJ; ++j) q[j] = sj[threadIdx.x * L::NJ + j];
        }
        __syncthreads();                              // the stage is reused for the poses below
    } else if (active) {
        load_jv<L>(a.jv, cfg, q);
    }
    if (active) {
        Ctl<L> c;
        setup_control<L>(rb, q, a.mode, a.step, a.nsamples, c);
        if (c.status) {
            write_invalid<L>(a, cfg, rb.maxs);                  // a.tip is null here: sample count, step, status only
        } else {
            solve_fast<L, true, NoSample, FINE>(rb, c, tip, nullptr, NoSample{});
            if (a.n_out) a.n_out[cfg] = c.nframes;
            if (a.step_out) a.step_out[cfg] = c.h;
            if (a.status) a.status[cfg] = 0;
        }
    }
    if (a.order || a.gather_rowwise) {                 // uniform over the grid: no thread skips the barrier below
        if (active) store_tf_gather(a, cfg, tip);
        return;
    }
#pragma unroll
    for (int j = 0; j < 6; ++j) stage[threadIdx.x * 6 + j] = make_double2(tip[2 * j], tip[2 * j + 1]);
    __syncthreads();
    const int64_t first = (int64_t)blockIdx.x * kBlock;
    const int64_t left = a.B - first;
    const int     nchunk = (int)(left < kBlock ? left : kBlock) * 6;
    if (a.gather_mc) {
        double* d = a.gather[0] + (a.gather_row + first) * 12;
        for (int ch = threadIdx.x; ch < nchunk; ch += kBlock) {
            const double2 v = stage[ch];
            const unsigned long long x = (unsigned long long)__double_as_longlong(v.x), y = (unsigned long long)__double_as_longlong(v.y);
            asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(d + 2 * ch),
                         "f"(__uint_as_float((unsigned)x)), "f"(__uint_as_float((unsigned)(x >> 32))),
                         "f"(__uint_as_float((unsigned)y)), "f"(__uint_as_float((unsigned)(y >> 32)))
                         : "memory");
        }
        return;
    }
#pragma unroll 1
    for (int p = 0; p < a.ngather; ++p) {
        double2* d = reinterpret_cast<double2*>(a.gather[p] + (a.gather_row + first) * 12);
        for (int ch = threadIdx.x; ch < nchunk; ch += kBlock) d[ch] = stage[ch];
    }
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
    if (c.status) { write_invalid<L>(a, cfg, rb.maxs); return; }

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

// One 256-bit store (sm_100: STG.E.ENL2.256): a full 32-B sector per request.  Scattered stores are
// bound by the number of L2 write requests they generate, not by bytes (ncu on the first backbone
// kernels: long-scoreboard stalls on registers still owned by in-flight 16-B stores), so the frame
// stores carry whole sectors.  `p` must be 32-B aligned.
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

// ---- backbone kernel, FAST arithmetic, every layout: two sweeps per thread -----------------------
// Frames need the prefix products E_0 ... E_k, the sweep produces sample N-1 first.  Earlier
// kernels parked every sample in the output region and rebuilt the frames in a second pass
// (thread-per-configuration: v1; warp-per-configuration with a scan over SE(3) products: v2/v3) or
// saved 32 sweep states and let the lanes of a warp repeat their run of the sweep (v4); see
// DESIGN.md §4.3 and profiles/r01_backbone_*.  What won is the simplest scheme: the thread sweeps
// twice.  Sweep 1 is the tip solver (total product T); sweep 2 repeats the sweep from the tip and
// emits Frame_k = Init Rz(alpha_k+theta) I_k with I_{ktip} = T and I_{k-1} = I_k inv(E_k), one
// sample late so the frame arithmetic overlaps the next sweep step (FastSweep::run_emit).  No
// shared memory, no scratch traffic: DRAM sees the joint values and the output.
//   SoA [MaxS][12][B]: a warp stores 32 x 8 B contiguous per matrix entry.
//   AoS [B][MaxS][12]: a thread stores its frame as three full 32-B sectors; consecutive threads
//   write 96*MaxS bytes apart, which the memory system takes at full speed as long as the
//   configurations in flight are neighbours (hence window-local binning in step mode; a random
//   visiting order is 2.8x slower, tools/tune_bb.cu).
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
template <bool ALIGNED32>
struct FrameStoreAos {       // one configuration's frames back to back: 3 full-sector stores per frame
    double* slot0;
    double* tip;
    int     ktip;
    __device__ __forceinline__ void operator()(int k, const double* F) const
    {
        if (ALIGNED32) store_tf256(slot0 + (int64_t)k * 12, F);
        else           store_tf(slot0 + (int64_t)k * 12, F);
        if (k == ktip && tip) store_tf(tip, F);
    }
};

// LAYOUT: 0 = SoA [MaxS][12][B], 1 = AoS [B][MaxS][12] with a 32-byte aligned buffer, 2 = AoS, 16-byte aligned
template <class L, int LAYOUT>
__global__ void __launch_bounds__(kBlock, backbone_blocks_per_sm(L::T, L::S))
fk_backbone_twosweep_kernel(const __grid_constant__ KRobot rb, const __grid_constant__ FkArgs a)
{
    const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (i >= a.B) return;
    const int64_t cfg = a.order ? (int64_t)a.order[i] : i;
    double q[L::NJ];
    load_jv<L>(a.jv, cfg, q);
    Ctl<L> c;
    setup_control<L>(rb, q, a.mode, a.step, a.nsamples, c);
    if (c.status) { write_invalid<L>(a, cfg, rb.maxs); return; }

    QT T;
    {
        FastSweep<L, true, NoSample> s1(rb, c, NoSample{});
        // the tube angles are an output (alpha_base).  Finely sampled launches run the loop in units of the step; coarse
        // ones the generic loop that form falls back to anyway (no third copy of the sweep in this kernel: with one it
        // ran 3 % slower at N = 256, 6.61 against 6.41 ms per 1M x 256 frames), through its one inlined call site
        s1.template run_scaled<true, true>(/*generic=*/!a.fine);
        T.w = s1.Qw; T.x = s1.Qx; T.y = s1.Qy; T.z = s1.Qz; T.px = s1.Px; T.py = s1.Py; T.pz = s1.Pz;
        if (a.alpha_base) {
#pragma unroll
            for (int t = 0; t < L::T; ++t) a.alpha_base[cfg * L::T + t] = s1.al[t];
        }
    }
    if (LAYOUT == 0) {
        const FrameStoreSoa fs{a.backbone + cfg, a.B, a.tip ? a.tip + cfg * 12 : nullptr, c.nframes - 1};
        FastSweep<L, false, PendingFrameEmitter<FrameStoreSoa>, true> s2(rb, c, make_frame_emitter<L>(rb, c, fs, T));
        s2.run_emit();
        if (a.radius) fill_radius<L>(rb, c, a.radius + cfg, a.B);
    } else {
        using Store = FrameStoreAos<LAYOUT == 1>;
        const Store fs{a.backbone + cfg * 12 * (int64_t)rb.maxs, a.tip ? a.tip + cfg * 12 : nullptr, c.nframes - 1};
        FastSweep<L, false, PendingFrameEmitter<Store>, true> s2(rb, c, make_frame_emitter<L>(rb, c, fs, T));
        s2.run_emit();
        if (a.radius) fill_radius<L>(rb, c, a.radius + cfg * (int64_t)rb.maxs, 1);
    }
    if (a.n_out) a.n_out[cfg] = c.nframes;
    if (a.step_out) a.step_out[cfg] = c.h;
    if (a.status) a.status[cfg] = 0;
}

#define CTRK_LAUNCH_CASE(KERNEL, FASTV, S, A, Bq, C)                                              \
    case S * 1000 + A * 100 + Bq * 10 + C:                                                        \
        KERNEL<Lay<S, A, Bq, C>, FASTV><<<grid, kBlock, 0, st>>>(rb, a);                          \
        break;

}  // namespace ctrk
