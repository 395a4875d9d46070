// ctr_fk_kernels_backbone_aos16.cu — the two-sweep backbone kernel, AoS output in a buffer that is only 16-byte aligned
// (two 128-bit stores per 32 bytes instead of one 256-bit store), all 14 section layouts.
#include "ctr_fk_kernels.cuh"
#include "ctr_fk_internal.h"

namespace ctrk {

static inline unsigned grid_for(int64_t B) { return (unsigned)((B + kBlock - 1) / kBlock); }

cudaError_t launch_backbone_fast_aos16(int key, const KRobot& rb, const FkArgs& a, cudaStream_t st)
{
    if (a.B <= 0) return cudaSuccess;
    const unsigned grid = grid_for(a.B);
    // one kernel for every layout: two sweeps per thread, frames emitted as the second sweep passes
    switch (key) {
#define X(S, A, Bq, C) \
    case S * 1000 + A * 100 + Bq * 10 + C: \
        fk_backbone_twosweep_kernel<Lay<S, A, Bq, C>, 2><<<grid, kBlock, 0, st>>>(rb, a); \
        break;
        CTRK_FOR_EACH_LAYOUT_K(X)
#undef X
        default: return cudaErrorInvalidValue;
    }
    count_launch();
    return cudaGetLastError();
}

}  // namespace ctrk
