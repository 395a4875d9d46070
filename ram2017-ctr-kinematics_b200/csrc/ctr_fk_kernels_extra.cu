// ctr_fk_kernels_extra.cu — FAST instantiations of the Jacobian and stability kernels (ctr_fk_kernels_extra.cuh) and
// the base-angle shooting solve.
#include "ctr_fk_kernels_extra.cuh"

// (the stability and shooting kernels are instantiated in ctr_fk_kernels_stability.cu / ctr_fk_kernels_shoot.cu so that the
// three translation units compile in parallel)
namespace ctrk {

cudaError_t launch_jacobian_sample(int key, const KRobot& rb, const JacArgs& a, cudaStream_t st);     // ctr_fk_kernels_jacobian_sample.cu
cudaError_t launch_jacobian(int key, const KRobot& rb, const JacArgs& a, cudaStream_t st)
{
    return a.i_sample >= 0 ? launch_jacobian_sample(key, rb, a, st) : launch_jacobian_t<true, false>(key, rb, a, st);
}
}  // namespace ctrk
