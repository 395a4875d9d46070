// ctr_fk_kernels_gather.cu — solve + all-gather in one kernel (fk_tip_gather_kernel, FAST arithmetic) for all 14 section
// layouts.
#include "ctr_fk_kernels.cuh"
#include "ctr_fk_internal.h"

namespace ctrk {

static inline unsigned grid_for(int64_t B) { return (unsigned)((B + kBlock - 1) / kBlock); }

cudaError_t launch_tip_gather(int key, const KRobot& rb, const FkArgs& args, cudaStream_t st)
{
    if (args.B <= 0) return cudaSuccess;
    FkArgs a = args;
    unsigned grid = grid_for(a.B);
    if (a.order) {
        a.windows = (int)((a.B + kBinWindow - 1) / kBinWindow);
        grid = (unsigned)a.windows * (kBinWindow / kBlock);
    }
    switch (key) {
#define X(S, A, Bq, C) \
    case S * 1000 + A * 100 + Bq * 10 + C: \
        if (a.fine) fk_tip_gather_kernel<Lay<S, A, Bq, C>, true><<<grid, kBlock, 0, st>>>(rb, a); \
        else        fk_tip_gather_kernel<Lay<S, A, Bq, C>, false><<<grid, kBlock, 0, st>>>(rb, a); \
        break;
        CTRK_FOR_EACH_LAYOUT_K(X)
#undef X
        default: return cudaErrorInvalidValue;
    }
    count_launch();
    return cudaGetLastError();
}

}  // namespace ctrk
