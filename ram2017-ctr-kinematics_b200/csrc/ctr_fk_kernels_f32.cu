// ctr_fk_kernels_f32.cu — FP32 variant of the tip solver (see ctr_fk_core_f32.cuh).
#include "ctr_fk_kernels.cuh"
#include "ctr_fk_core_f32x2.cuh"
#include "ctr_fk_internal.h"

namespace ctrk {

template <class L>
__global__ void __launch_bounds__(kBlock)
fk_tip_f32_kernel(const __grid_constant__ KRobot rb, const __grid_constant__ FkArgs a)
{
    const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (i >= a.B) return;
    const int64_t cfg = a.order ? (int64_t)a.order[i] : i;
    double q[L::NJ];
    load_jv<L>(a.jv, cfg, q);
    Ctl<L> c;
    setup_control<L>(rb, q, a.mode, a.step, a.nsamples, c);
    if (c.status) { write_invalid<L>(a, cfg, rb.maxs); return; }
    double tip[12];
    solve_f32<L>(rb, c, tip);
    if (a.tip) store_tf(a.tip + cfg * 12, tip);
    if (a.n_out) a.n_out[cfg] = c.nframes;
    if (a.step_out) a.step_out[cfg] = c.h;
    if (a.status) a.status[cfg] = 0;
}

// Fixed-N mode: two configurations per thread in packed single precision (FFMA2).
template <class L>
__global__ void __launch_bounds__(kBlock)
fk_tip_f32x2_kernel(const __grid_constant__ KRobot rb, const __grid_constant__ FkArgs a)
{
    const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    const int64_t cfg0 = 2 * i, cfg1 = 2 * i + 1;
    if (cfg0 >= a.B) return;
    const bool has1 = cfg1 < a.B;
    double q[L::NJ];
    Ctl<L> c0, c1;
    load_jv<L>(a.jv, cfg0, q);
    setup_control<L>(rb, q, a.mode, a.step, a.nsamples, c0);
    if (has1) {
        load_jv<L>(a.jv, cfg1, q);
        setup_control<L>(rb, q, a.mode, a.step, a.nsamples, c1);
    } else {
        c1 = c0;
    }
    const bool ok0 = (c0.status == 0), ok1 = has1 && (c1.status == 0);
    if (!ok0) write_invalid<L>(a, cfg0, rb.maxs);
    if (has1 && !ok1) write_invalid<L>(a, cfg1, rb.maxs);
    if (!ok0 && !ok1) return;
    // an invalid or absent configuration runs as a copy of its partner and is not stored
    double tip0[12], tip1[12];
    if (!ok0) c0 = c1;
    if (!ok1) c1 = c0;
    solve_f32x2<L>(rb, c0, c1, ok0 ? tip0 : nullptr, ok1 ? tip1 : nullptr);
    if (ok0) {
        if (a.tip) store_tf(a.tip + cfg0 * 12, tip0);
        if (a.n_out) a.n_out[cfg0] = c0.nframes;
        if (a.step_out) a.step_out[cfg0] = c0.h;
        if (a.status) a.status[cfg0] = 0;
    }
    if (ok1) {
        if (a.tip) store_tf(a.tip + cfg1 * 12, tip1);
        if (a.n_out) a.n_out[cfg1] = c1.nframes;
        if (a.step_out) a.step_out[cfg1] = c1.h;
        if (a.status) a.status[cfg1] = 0;
    }
}

cudaError_t launch_tip_f32(int key, const KRobot& rb, const FkArgs& a, cudaStream_t st)
{
    if (a.B <= 0) return cudaSuccess;
    if (a.mode == 1 && !a.order) {      // CTR_FK_MODE_NSAMPLE: every configuration takes the same number of steps
        const int64_t pairs = (a.B + 1) / 2;
        const unsigned gridp = (unsigned)((pairs + kBlock - 1) / kBlock);
        switch (key) {
#define X(S, A, Bq, C) \
    case S * 1000 + A * 100 + Bq * 10 + C: \
        fk_tip_f32x2_kernel<Lay<S, A, Bq, C>><<<gridp, kBlock, 0, st>>>(rb, a); \
        break;
            CTRK_FOR_EACH_LAYOUT_K(X)
#undef X
            default: return cudaErrorInvalidValue;
        }
        count_launch();
        return cudaGetLastError();
    }
    const unsigned grid = (unsigned)((a.B + kBlock - 1) / kBlock);
    switch (key) {
#define X(S, A, Bq, C) \
    case S * 1000 + A * 100 + Bq * 10 + C: \
        fk_tip_f32_kernel<Lay<S, A, Bq, C>><<<grid, kBlock, 0, st>>>(rb, a); \
        break;
        CTRK_FOR_EACH_LAYOUT_K(X)
#undef X
        default: return cudaErrorInvalidValue;
    }
    count_launch();
    return cudaGetLastError();
}

}  // namespace ctrk
