// ctr_fk_kernels_fast.cu — the tip kernel with FAST arithmetic (series SE(3) step, quaternion accumulation, custom
// sin/cos) for all 14 section layouts.  The gather and backbone kernels are instantiated in their own translation units
// (ctr_fk_kernels_gather.cu, ctr_fk_kernels_backbone.cu) so that the three compile in parallel.
#include "ctr_fk_kernels.cuh"
#include "ctr_fk_internal.h"

namespace ctrk {

static inline unsigned grid_for(int64_t B) { return (unsigned)((B + kBlock - 1) / kBlock); }

cudaError_t launch_tip_fast(int key, const KRobot& rb, const FkArgs& args, cudaStream_t st)
{
    if (args.B <= 0) return cudaSuccess;
    FkArgs a = args;
    unsigned grid = grid_for(a.B);
    if (a.order) {      // window-sorted order: walk the windows rank-major (see fk_tip_kernel)
        a.windows = (int)((a.B + kBinWindow - 1) / kBinWindow);
        grid = (unsigned)a.windows * (kBinWindow / kBlock);
    }
    switch (key) {
#define X(S, A, Bq, C) \
    case S * 1000 + A * 100 + Bq * 10 + C: \
        if (a.fine) fk_tip_kernel<Lay<S, A, Bq, C>, true, true><<<grid, kBlock, 0, st>>>(rb, a); \
        else        fk_tip_kernel<Lay<S, A, Bq, C>, true, false><<<grid, kBlock, 0, st>>>(rb, a); \
        break;
        CTRK_FOR_EACH_LAYOUT_K(X)
#undef X
        default: return cudaErrorInvalidValue;
    }
    count_launch();
    return cudaGetLastError();
}

}  // namespace ctrk
