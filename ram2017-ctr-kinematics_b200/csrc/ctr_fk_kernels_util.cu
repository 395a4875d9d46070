// ctr_fk_kernels_util.cu — small kernels around the solvers: step-mode binning (sample count,
// histogram, counting sort), the device-side joint sampler and the FP64 peak probe.
#include "ctr_fk_kernels.cuh"
#include "ctr_fk_internal.h"

namespace ctrk {

// ---- step-mode binning -------------------------------------------------------------------------
// In step mode N_c = floor(sum(Phi)/step) differs per configuration (1..MaxSamples).  A warp runs
// as long as its longest sweep, so configurations are counting-sorted by N_c (longest first) and
// the solver threads pick their configuration through `order`.  Only the 4-byte index moves.
// History: one global atomic per configuration serialised on ~256 hot addresses in L2 (213 us per
// pass for 1M configurations); block-private shared-memory histograms fixed that; the batch-wide
// sort was then replaced by window-local sorting (below) for the sake of the backbone kernels.
constexpr int kBinThreads = 256;
constexpr int kBinItems = kBinWindow;                       // one CTA sorts one window
constexpr int kBinPerThread = kBinItems / kBinThreads;

template <class L>
__device__ __forceinline__ int sample_count_of(const KRobot& rb, const double* __restrict__ jv, int64_t i,
                                               int mode, double step, int nsamples)
{
    double q[L::NJ];
    load_jv<L>(jv, i, q);
    // Only N is needed for the ordering (not nframes): skip the position replay.
    double run = 0.0;
    bool bad = false;
#pragma unroll
    for (int s = 0; s < L::S; ++s) {
        const double phi = q[1 + s + L::t0(s)];
        if (!((phi >= 0.0) && (phi <= rb.len[s]))) bad = true;
        run = x_add(run, phi);
    }
    if (bad) return 0;
    if (mode != 0) return nsamples < rb.maxs ? nsamples : rb.maxs;
    const double qd = x_div(run, step);
    int cnt = (qd >= (double)rb.maxs + 1.0) ? rb.maxs + 1 : (int)qd;
    if (cnt < 1) cnt = 1;
    if (cnt > rb.maxs) cnt = rb.maxs;
    return cnt;
}

// One kernel: every CTA counting-sorts its own window of kBinItems consecutive configurations
// (longest sweep first) and writes that window of `order`.  Sorting inside windows instead of
// across the batch keeps the configurations a wave of solver threads touches close together:
// the backbone kernels write 24.6 KB per configuration, and with a batch-wide sort the ~75 k
// regions in flight are scattered over the whole output buffer (measured with a random visiting
// order: 2.8x slower, tools/tune_bb.cu); a window of 4096 configurations costs nothing.  A warp
// still sees near-uniform sample counts (a window holds 16 configurations per possible count at
// MaxSamples = 256).  No global atomics, no temporaries besides `order`.
template <class L>
__global__ void __launch_bounds__(kBinThreads)
bin_local_kernel(const __grid_constant__ KRobot rb, const double* __restrict__ jv, int64_t B, int mode,
                 double step, int nsamples, int32_t* __restrict__ order, int nbins)
{
    extern __shared__ uint32_t sh[];          // [0,nbins): window histogram, then bin offsets
    uint32_t* fill = sh + nbins;              // [nbins, 2 nbins): running rank inside a bin
    for (int b = threadIdx.x; b < 2 * nbins; b += kBinThreads) sh[b] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * kBinItems;
    // the count is cheap (one division): it is evaluated again in the scatter pass instead of
    // being kept in kBinPerThread registers
#pragma unroll 4
    for (int j = 0; j < kBinPerThread; ++j) {
        const int64_t i = base + (int64_t)j * kBinThreads + threadIdx.x;
        if (i < B) atomicAdd(&sh[sample_count_of<L>(rb, jv, i, mode, step, nsamples)], 1u);
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        // exclusive scan from the LARGEST count down, one warp, nbins/32 chunks
        uint32_t carry = 0;
        for (int hi = nbins - 1; hi >= 0; hi -= 32) {
            const int b = hi - (int)threadIdx.x;
            const uint32_t c = (b >= 0) ? sh[b] : 0u;
            uint32_t incl = c;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const uint32_t o = __shfl_up_sync(0xffffffffu, incl, off);
                if ((int)threadIdx.x >= off) incl += o;
            }
            if (b >= 0) sh[b] = carry + incl - c;
            carry += __shfl_sync(0xffffffffu, incl, 31);
        }
    }
    __syncthreads();
#pragma unroll 4
    for (int j = 0; j < kBinPerThread; ++j) {
        const int64_t i = base + (int64_t)j * kBinThreads + threadIdx.x;
        if (i >= B) continue;
        const int cnt = sample_count_of<L>(rb, jv, i, mode, step, nsamples);
        const uint32_t r = atomicAdd(&fill[cnt], 1u);
        order[base + sh[cnt] + r] = (int32_t)i;
    }
}

cudaError_t launch_bin_local(int key, const KRobot& rb, const double* jv, int64_t B, int mode, double step,
                             int nsamples, int32_t* order, cudaStream_t st)
{
    if (B <= 0) return cudaSuccess;
    const unsigned grid = (unsigned)((B + kBinItems - 1) / kBinItems);
    const int nbins = rb.maxs + 2;
    const size_t smem = 2 * sizeof(uint32_t) * (size_t)nbins;
    switch (key) {
#define X(S, A, Bq, C) \
    case S * 1000 + A * 100 + Bq * 10 + C: \
        bin_local_kernel<Lay<S, A, Bq, C>><<<grid, kBinThreads, smem, st>>>(rb, jv, B, mode, step, nsamples, order, nbins); \
        break;
        CTRK_FOR_EACH_LAYOUT_K(X)
#undef X
        default: return cudaErrorInvalidValue;
    }
    count_launch();
    return cudaGetLastError();
}

// ---- joint sampler (robot_i.h:1308-1317 distribution, counter-based) ---------------------------
__global__ void __launch_bounds__(256)
sampler_kernel(const __grid_constant__ SamplerArgs a)
{
    const int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x;   // one thread per joint value
    if (e >= a.B * a.tb.njv) return;
    const int64_t i = e / a.tb.njv;
    const int     j = (int)(e - i * a.tb.njv);
    a.jv[e] = a.tb.scale[j] * sampler_u01(a.seed, (uint64_t)(a.first + i), (uint32_t)j);
}

// scale policies couple the Phi slots of one configuration: one thread per configuration
__global__ void __launch_bounds__(256)
sampler_policy_kernel(const __grid_constant__ SamplerArgs a)
{
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= a.B) return;
    double q[CTR_FK_MAX_JV];
    sample_config(a.tb, a.seed, (uint64_t)(a.first + i), q);
    for (int j = 0; j < a.tb.njv; ++j) a.jv[i * a.tb.njv + j] = q[j];
}

cudaError_t launch_sampler(const SamplerArgs& a, cudaStream_t st)
{
    const int64_t total = a.B * a.tb.njv;
    if (total <= 0) return cudaSuccess;
    if (a.tb.policy == 0) sampler_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(a);
    else                  sampler_policy_kernel<<<(unsigned)((a.B + 255) / 256), 256, 0, st>>>(a);
    count_launch();
    return cudaGetLastError();
}

// ---- FP64 FMA-pipe probe -----------------------------------------------------------------------
// kProbeChains independent DFMA chains per thread; flops = threads * chains * iters * 2.
__global__ void __launch_bounds__(kProbeThreads)
fp64_probe_kernel(int iters, double* sink)
{
    double v[kProbeChains];
#pragma unroll
    for (int j = 0; j < kProbeChains; ++j) v[j] = 1.0 + 1e-9 * (double)(threadIdx.x + j);
    const double a = 1.0000000001, b = 1e-12;
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int j = 0; j < kProbeChains; ++j) v[j] = __fma_rn(v[j], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < kProbeChains; ++j) s += v[j];
    if (s == 12345.678) sink[0] = s;   // never true; keeps the chains alive
}

// ---- FP32 FMA-pipe probe, packed form ------------------------------------------------------------
// The FP32 variant's solver issues fma.rn.f32x2 (SASS FFMA2: one instruction, two lanes of a 64-bit register
// pair), so its roofline denominator is the rate at which independent FFMA2 chains retire on every SM.
// flops = threads * chains * 8 * iters * 4 (two FMAs of two flops per instruction).
__global__ void __launch_bounds__(kProbeThreads)
fp32x2_probe_kernel(int iters, double* sink)
{
    unsigned long long v[kProbeChains];
#pragma unroll
    for (int j = 0; j < kProbeChains; ++j) {
        const float f = 1.0f + 1e-3f * (float)(threadIdx.x + j);
        v[j] = ((unsigned long long)__float_as_uint(f) << 32) | __float_as_uint(f);
    }
    const unsigned long long a = ((unsigned long long)__float_as_uint(1.0000001f) << 32) | __float_as_uint(0.9999999f);
    const unsigned long long b = ((unsigned long long)__float_as_uint(1e-7f) << 32) | __float_as_uint(-1e-7f);
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int j = 0; j < kProbeChains; ++j) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(v[j]) : "l"(a), "l"(b));
    }
    unsigned long long s = 0;
#pragma unroll
    for (int j = 0; j < kProbeChains; ++j) s ^= v[j];
    if (s == 0x123456789abcdefull) sink[0] = 1.0;   // never true; keeps the chains alive
}

cudaError_t launch_fp32x2_probe(int iters, int blocks, double* sink, cudaStream_t st)
{
    fp32x2_probe_kernel<<<blocks, kProbeThreads, 0, st>>>(iters, sink);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_fp64_probe(int iters, int blocks, double* sink, cudaStream_t st)
{
    fp64_probe_kernel<<<blocks, kProbeThreads, 0, st>>>(iters, sink);
    count_launch();
    return cudaGetLastError();
}

}  // namespace ctrk
