// ctr_fk_kernels_f32.cu — FP32 variant of the tip solver (see ctr_fk_core_f32.cuh).
#include "ctr_fk_kernels.cuh"
#include "ctr_fk_core_f32.cuh"
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
    if (c.status) { write_invalid<L>(a, cfg); return; }
    double tip[12];
    solve_f32<L>(rb, c, tip);
    if (a.tip) store_tf(a.tip + cfg * 12, tip);
    if (a.n_out) a.n_out[cfg] = c.nframes;
    if (a.step_out) a.step_out[cfg] = c.h;
    if (a.status) a.status[cfg] = 0;
}

cudaError_t launch_tip_f32(int key, const KRobot& rb, const FkArgs& a, cudaStream_t st)
{
    if (a.B <= 0) return cudaSuccess;
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
