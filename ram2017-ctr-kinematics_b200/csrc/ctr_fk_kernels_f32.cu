// ctr_fk_kernels_f32.cu — FP32 variant of the tip solver (placeholder until the FP32 solver lands).
#include "ctr_fk_kernels.cuh"
#include "ctr_fk_internal.h"

namespace ctrk {
cudaError_t launch_tip_f32(int, const KRobot&, const FkArgs&, cudaStream_t) { return cudaErrorNotSupported; }
}  // namespace ctrk
