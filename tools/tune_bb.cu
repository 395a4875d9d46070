// tune_bb.cu — standalone tuning harness (not part of the product): times the AoS backbone kernel
// of ctr_sec3_tub4 (layout 3,2,1,1) on B configurations x 256 frames, visiting the configurations
// in order, in a random order, and shuffled inside windows of 4096 (what step-mode binning does).
// The parking / warp-scan / checkpoint kernels this file used to compare
// are in the history: git show 4b28d88:tools/tune_bb.cu
//   nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo tools/tune_bb.cu -o tools/tune_bb ; tools/tune_bb 1000000
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>

#include "../ram2017-ctr-kinematics_b200/csrc/ctr_fk_kernels.cuh"
#include "../ram2017-ctr-kinematics_b200/csrc/ctr_fk_krobot.h"
#include "../ram2017-ctr-kinematics_b200/host/ctr_sampler.h"

using namespace ctrk;
namespace ctrk { void set_error(const char*, ...) {} void count_launch(int) {} }

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

template <class F>
float time_it(F launch)
{
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int w = 0; w < 3; ++w) launch();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 7; ++r) {
        CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    return best;
}

int main(int argc, char** argv)
{
    const int64_t B = argc > 1 ? atoll(argv[1]) : 200000;
    // ctr_sec3_tub4: (VC 2 tubes) (CC) (CC) — parameters of tests/data/sec3_tub4.xml
    ctr_robot_desc d = {};
    d.nsections = 3; d.max_samples = 256;
    d.init_trafo[0] = d.init_trafo[4] = d.init_trafo[8] = 1.0;
    d.sec[0].ntube = 2; d.sec[0].length = 0.05;
    d.sec[0].stiffness[0] = d.sec[0].stiffness[1] = 5; d.sec[0].curvature[0] = d.sec[0].curvature[1] = 40.0;
    d.sec[0].radius[0] = 0.0015; d.sec[0].radius[1] = 0.00125; d.sec[0].poisson[0] = d.sec[0].poisson[1] = 0.3;
    d.sec[1].ntube = 1; d.sec[1].length = 0.05; d.sec[1].stiffness[0] = 2; d.sec[1].curvature[0] = 50.0;
    d.sec[1].radius[0] = 0.001; d.sec[1].poisson[0] = 0.3;
    d.sec[2].ntube = 1; d.sec[2].length = 0.05; d.sec[2].stiffness[0] = 1; d.sec[2].curvature[0] = 30.0;
    d.sec[2].radius[0] = 0.00075; d.sec[2].poisson[0] = 0.3;
    KRobot rb;
    if (make_krobot(&d, &rb)) return 1;
    using L = Lay<3, 2, 1, 1>;
    std::vector<double> jv(B * L::NJ);
    SamplerTable tb; make_sampler_table(&d, &tb);
    for (int64_t i = 0; i < B; ++i) for (int j = 0; j < L::NJ; ++j) jv[i * L::NJ + j] = tb.scale[j] * sampler_u01(20173, i, j);
    FkArgs a = {};
    double *djv, *dbb, *drad, *dtip;
    CK(cudaMalloc(&djv, sizeof(double) * B * L::NJ));
    CK(cudaMalloc(&dbb, sizeof(double) * B * 12 * 256)); CK(cudaMalloc(&drad, sizeof(double) * B * 256));
    CK(cudaMalloc(&dtip, sizeof(double) * B * 12));
    CK(cudaMemcpy(djv, jv.data(), sizeof(double) * B * L::NJ, cudaMemcpyHostToDevice));
    a.jv = djv; a.B = B; a.mode = 1; a.nsamples = 256; a.tip = dtip; a.backbone = dbb; a.radius = drad;

    std::vector<double> ref(12 * 256 * 4);
    {   // the shipped kernel: two sweeps per thread, AoS stores
        const unsigned grid = (unsigned)((B + kBlock - 1) / kBlock);
        CK(cudaMemset(dbb, 0, sizeof(double) * B * 12 * 256));
        float ms = time_it([&] { fk_backbone_twosweep_kernel<L, 1><<<grid, kBlock>>>(rb, a); });
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(ref.data(), dbb + (int64_t)12 * 256 * 1234, sizeof(double) * ref.size(), cudaMemcpyDeviceToHost));
        printf("two-sweep, AoS stores, configurations in order       %.3f ms  %.3e cfg/s\n", ms, B / (ms * 1e-3));
    }
    {   // same kernel, configurations visited in random order (what step-mode binning does to locality)
        std::vector<int32_t> perm(B);
        for (int64_t i = 0; i < B; ++i) perm[i] = (int32_t)i;
        uint64_t x = 88172645463325252ull;
        for (int64_t i = B - 1; i > 0; --i) { x ^= x << 13; x ^= x >> 7; x ^= x << 17; std::swap(perm[i], perm[x % (uint64_t)(i + 1)]); }
        int32_t* dperm;
        CK(cudaMalloc(&dperm, sizeof(int32_t) * B));
        CK(cudaMemcpy(dperm, perm.data(), sizeof(int32_t) * B, cudaMemcpyHostToDevice));
        FkArgs b = a; b.order = dperm;
        const unsigned grid = (unsigned)((B + kBlock - 1) / kBlock);
        float ms = time_it([&] { fk_backbone_twosweep_kernel<L, 1><<<grid, kBlock>>>(rb, b); });
        CK(cudaDeviceSynchronize());
        printf("two-sweep, AoS stores, random visiting order       %.3f ms  %.3e cfg/s\n", ms, B / (ms * 1e-3));
        // sorted-within-window order: blocks of 4096 consecutive configurations shuffled internally only
        for (int64_t i = 0; i < B; ++i) perm[i] = (int32_t)i;
        for (int64_t b0 = 0; b0 < B; b0 += 4096) {
            const int64_t e = std::min<int64_t>(B, b0 + 4096);
            for (int64_t i = e - 1; i > b0; --i) { x ^= x << 13; x ^= x >> 7; x ^= x << 17; std::swap(perm[i], perm[b0 + x % (uint64_t)(i - b0 + 1)]); }
        }
        CK(cudaMemcpy(dperm, perm.data(), sizeof(int32_t) * B, cudaMemcpyHostToDevice));
        ms = time_it([&] { fk_backbone_twosweep_kernel<L, 1><<<grid, kBlock>>>(rb, b); });
        CK(cudaDeviceSynchronize());
        printf("two-sweep, AoS stores, shuffled within 4096-blocks   %.3f ms  %.3e cfg/s\n", ms, B / (ms * 1e-3));
    }
    return 0;
}
