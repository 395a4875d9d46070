// ctr_fk_kernels_backbone.cu — the two-sweep backbone kernel, AoS output ([B][MaxS][12]) in a 32-byte aligned buffer, all
// 14 section layouts.  The SoA form and the 16-byte aligned AoS form are instantiated in ctr_fk_kernels_backbone_soa.cu /
// ctr_fk_kernels_backbone_aos16.cu (three translation units compile in parallel).
#include "ctr_fk_kernels.cuh"
#include "ctr_fk_internal.h"

namespace ctrk {

static inline unsigned grid_for(int64_t B) { return (unsigned)((B + kBlock - 1) / kBlock); }

cudaError_t launch_backbone_fast(int key, const KRobot& rb, const FkArgs& a, cudaStream_t st)
{
    if (a.B <= 0) return cudaSuccess;
    const unsigned grid = grid_for(a.B);
    // one kernel for every layout: two sweeps per thread, frames emitted as the second sweep passes
    if (a.soa) return launch_backbone_fast_soa(key, rb, a, st);      // ctr_fk_kernels_backbone_soa.cu
    if ((reinterpret_cast<uintptr_t>(a.backbone) & 31) != 0) return launch_backbone_fast_aos16(key, rb, a, st);   // ctr_fk_kernels_backbone_aos16.cu
    switch (key) {
#define X(S, A, Bq, C) \
    case S * 1000 + A * 100 + Bq * 10 + C: \
        fk_backbone_twosweep_kernel<Lay<S, A, Bq, C>, 1><<<grid, kBlock, 0, st>>>(rb, a); \
        break;
        CTRK_FOR_EACH_LAYOUT_K(X)
#undef X
        default: return cudaErrorInvalidValue;
    }
    count_launch();
    return cudaGetLastError();
}

}  // namespace ctrk
