// ctr_fk_kernels_jacobian_sample.cu — the iSample forms of the finite-difference Jacobian (FAST arithmetic), all 14 section
// layouts; the tip-only forms are in ctr_fk_kernels_extra.cu.
#include "ctr_fk_kernels_extra.cuh"

namespace ctrk {

cudaError_t launch_jacobian_sample(int key, const KRobot& rb, const JacArgs& a, cudaStream_t st)
{
    return launch_jacobian_t<true, true>(key, rb, a, st);
}

}  // namespace ctrk
