// ctr_fk_kernels_fast.cu — instantiates the solver kernels with FAST arithmetic (series SE(3) step, quaternion accumulation, custom sin/cos)
// for all 14 section layouts.
#include "ctr_fk_kernels.cuh"
#include "ctr_fk_internal.h"

namespace ctrk {

static inline unsigned grid_for(int64_t B) { return (unsigned)((B + kBlock - 1) / kBlock); }

cudaError_t launch_tip_fast(int key, const KRobot& rb, const FkArgs& a, cudaStream_t st)
{
    if (a.B <= 0) return cudaSuccess;
    const unsigned grid = grid_for(a.B);
    switch (key) {
#define X(S, A, Bq, C) CTRK_LAUNCH_CASE(fk_tip_kernel, true, S, A, Bq, C)
        CTRK_FOR_EACH_LAYOUT_K(X)
#undef X
        default: return cudaErrorInvalidValue;
    }
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_backbone_fast(int key, const KRobot& rb, const FkArgs& a, cudaStream_t st)
{
    if (a.B <= 0) return cudaSuccess;
    const unsigned grid = grid_for(a.B);
    switch (key) {
#define X(S, A, Bq, C) CTRK_LAUNCH_CASE(fk_backbone_kernel, true, S, A, Bq, C)
        CTRK_FOR_EACH_LAYOUT_K(X)
#undef X
        default: return cudaErrorInvalidValue;
    }
    count_launch();
    return cudaGetLastError();
}

}  // namespace ctrk
