// tune_bb.cu — standalone tuning harness (not part of the product): times the AoS backbone kernels
// of ctr_sec3_tub4 (layout 3,2,1,1) on 200 k configurations x 256 frames.
//   nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -DCTRK_CK_CTAS=6 -DCTRK_CK_WARPS=2 \
//        tools/tune_bb.cu -o tools/tune_bb_6_2 ; tools/tune_bb_6_2 1000000
// -DCTRK_CK_CTAS=n        register cap of the emit kernel (resident CTAs per SM)
// -DCTRK_BB_EXPERIMENT=1  pass 1 only;  =2  frames computed but not stored (DESIGN.md §4.3 timings)
#include <cstdio>
#include <cstdlib>
#include <vector>
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

    {   // older kernel: per-sample parking
        const size_t smem = sizeof(double) * 4 * 256 * kBbBufs * kBbWarps;
        CK(cudaFuncSetAttribute(fk_backbone_warp_kernel<L>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const unsigned grid = (unsigned)((B + kBbBlock - 1) / kBbBlock);
        float ms = time_it([&] { fk_backbone_warp_kernel<L><<<grid, kBbBlock, smem>>>(rb, a); });
        CK(cudaDeviceSynchronize());
        printf("parking kernel              %.3f ms  %.3e cfg/s\n", ms, B / (ms * 1e-3));
    }
    std::vector<double> ref(12 * 256 * 4);
    CK(cudaMemcpy(ref.data(), dbb + (int64_t)12 * 256 * 1234, sizeof(double) * ref.size(), cudaMemcpyDeviceToHost));
    CK(cudaMemset(dbb, 0, sizeof(double) * B * 12 * 256));
    {
        const size_t smem = 0;
        cudaFuncAttributes fa;
        CK(cudaFuncGetAttributes(&fa, fk_backbone_emit_kernel<L>));
        int occ = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fk_backbone_emit_kernel<L>, kCkBlock, smem));
        const unsigned grid = (unsigned)((B + kCkBlock - 1) / kCkBlock);
        float ms = time_it([&] { fk_backbone_emit_kernel<L><<<grid, kCkBlock, smem>>>(rb, a); });
        CK(cudaDeviceSynchronize());
        std::vector<double> got(ref.size());
        CK(cudaMemcpy(got.data(), dbb + (int64_t)12 * 256 * 1234, sizeof(double) * got.size(), cudaMemcpyDeviceToHost));
        double md = 0;
        for (size_t i = 0; i < got.size(); ++i) md = fmax(md, fabs(got[i] - ref[i]));
        printf("emit kernel ctas %2d warps/cta %d  regs %3d local %4zu smem %zu occ %2d CTAs (%2d warps/SM)  %.3f ms  %.3e cfg/s  maxdiff %.2e\n",
               CTRK_CK_CTAS, kCkWarps, fa.numRegs, fa.localSizeBytes, fa.sharedSizeBytes, occ, occ * kCkWarps, ms, B / (ms * 1e-3), md);
    }
    return 0;
}
