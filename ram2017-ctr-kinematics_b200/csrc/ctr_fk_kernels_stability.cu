// ctr_fk_kernels_stability.cu — calcStability / calcDegreeStability kernels, FAST arithmetic, all 14 section layouts.
#include "ctr_fk_kernels_extra.cuh"

namespace ctrk {

cudaError_t launch_stability(int key, const KRobot& rb, const StabArgs& a, int sm_count, cudaStream_t st)
{
    return launch_stability_t<true>(key, rb, a, sm_count, st);
}

}  // namespace ctrk
