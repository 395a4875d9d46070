// ctr_fk_kernels_strict.cu — instantiates the solver kernels with STRICT arithmetic (reference operation order); build this file with -fmad=false
// for all 14 section layouts.
#include "ctr_fk_kernels_extra.cuh"

namespace ctrk {

static inline unsigned grid_for(int64_t B) { return (unsigned)((B + kBlock - 1) / kBlock); }

cudaError_t launch_tip_strict(int key, const KRobot& rb, const FkArgs& a, cudaStream_t st)
{
    if (a.B <= 0) return cudaSuccess;
    const unsigned grid = grid_for(a.B);
    switch (key) {
#define X(S, A, Bq, C) CTRK_LAUNCH_CASE(fk_tip_kernel, false, S, A, Bq, C)
        CTRK_FOR_EACH_LAYOUT_K(X)
#undef X
        default: return cudaErrorInvalidValue;
    }
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_backbone_strict(int key, const KRobot& rb, const FkArgs& a, cudaStream_t st)
{
    if (a.B <= 0) return cudaSuccess;
    const unsigned grid = grid_for(a.B);
    switch (key) {
#define X(S, A, Bq, C) CTRK_LAUNCH_CASE(fk_backbone_kernel, false, S, A, Bq, C)
        CTRK_FOR_EACH_LAYOUT_K(X)
#undef X
        default: return cudaErrorInvalidValue;
    }
    count_launch();
    return cudaGetLastError();
}

// CTR_FK_FLAG_STRICT forms of the Jacobian and stability calls: every solve inside them is the STRICT solver
cudaError_t launch_jacobian_strict(int key, const KRobot& rb, const JacArgs& a, cudaStream_t st) { return a.i_sample >= 0 ? launch_jacobian_t<false, true>(key, rb, a, st) : launch_jacobian_t<false, false>(key, rb, a, st); }
cudaError_t launch_stability_strict(int key, const KRobot& rb, const StabArgs& a, int sm_count, cudaStream_t st)
{
    return launch_stability_t<false>(key, rb, a, sm_count, st);
}

}  // namespace ctrk
