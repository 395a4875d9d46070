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
    // AoS: warp-parallel second pass (needs a 32-B aligned buffer for its 256-bit stores)
    const bool warp_ok = !a.soa && rb.maxs <= kBbMaxSamples && (reinterpret_cast<uintptr_t>(a.backbone) & 31) == 0;
    // long backbones: checkpointed sweep, frames emitted per run (no per-sample parking round trip)
    if (warp_ok && rb.maxs >= kCkMinSamples) {
        const unsigned gridw = (unsigned)((a.B + kCkBlock - 1) / kCkBlock);
        switch (key) {
#define X(S, A, Bq, C) \
    case S * 1000 + A * 100 + Bq * 10 + C: \
        fk_backbone_emit_kernel<Lay<S, A, Bq, C>><<<gridw, kCkBlock, 0, st>>>(rb, a); \
        break;
            CTRK_FOR_EACH_LAYOUT_K(X)
#undef X
            default: return cudaErrorInvalidValue;
        }
        count_launch();
        return cudaGetLastError();
    }
    if (warp_ok) {
        const unsigned gridw = (unsigned)((a.B + kBbBlock - 1) / kBbBlock);
        const size_t smem = sizeof(double) * 4 * (size_t)rb.maxs * kBbBufs * kBbWarps;
        switch (key) {
#define X(S, A, Bq, C) \
    case S * 1000 + A * 100 + Bq * 10 + C: { \
        /* per device and cheap: set on every launch (above the 48 KB default for MaxSamples > 384) */ \
        cudaError_t ea = cudaFuncSetAttribute(fk_backbone_warp_kernel<Lay<S, A, Bq, C>>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        if (ea != cudaSuccess) return ea; \
        fk_backbone_warp_kernel<Lay<S, A, Bq, C>><<<gridw, kBbBlock, smem, st>>>(rb, a); \
        break; }
            CTRK_FOR_EACH_LAYOUT_K(X)
#undef X
            default: return cudaErrorInvalidValue;
        }
        count_launch();
        return cudaGetLastError();
    }
    if (a.soa) {        // two sweeps per thread, frames emitted coalesced across configurations
        switch (key) {
#define X(S, A, Bq, C) \
    case S * 1000 + A * 100 + Bq * 10 + C: \
        fk_backbone_soa_kernel<Lay<S, A, Bq, C>><<<grid, kBlock, 0, st>>>(rb, a); \
        break;
            CTRK_FOR_EACH_LAYOUT_K(X)
#undef X
            default: return cudaErrorInvalidValue;
        }
        count_launch();
        return cudaGetLastError();
    }
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
