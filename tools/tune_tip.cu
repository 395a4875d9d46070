// tune_tip.cu — standalone tuning harness (not part of the product): times the FAST tip solver of
// ctr_sec2_tub3 (layout 2,2,1) for several block sizes / register caps on 1M configurations.
//   nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -I.. tools/tune_tip.cu -o tools/tune_tip
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#include "../ram2017-ctr-kinematics_b200/csrc/ctr_fk_kernels.cuh"
#include "../ram2017-ctr-kinematics_b200/csrc/ctr_fk_krobot.h"
#include "../ram2017-ctr-kinematics_b200/host/ctr_sampler.h"

using namespace ctrk;
namespace ctrk { void set_error(const char*, ...) {} void count_launch(int) {} }

template <class L, int BLOCK, int MINB>
__global__ void __launch_bounds__(BLOCK, MINB)
tip_variant(const __grid_constant__ KRobot rb, const __grid_constant__ FkArgs a)
{
    const int64_t i = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
    if (i >= a.B) return;
    double q[L::NJ];
    load_jv<L>(a.jv, i, q);
    Ctl<L> c;
    setup_control<L>(rb, q, a.mode, a.step, a.nsamples, c);
    if (c.status) return;
    double tip[12];
    solve_fast<L, true>(rb, c, tip, nullptr, NoSample{});
    store_tf(a.tip + i * 12, tip);
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

template <int BLOCK, int MINB>
void run(const KRobot& rb, FkArgs a, const char* tag)
{
    using L = Lay<2, 2, 1, 0>;
    cudaFuncAttributes fa;
    CK(cudaFuncGetAttributes(&fa, tip_variant<L, BLOCK, MINB>));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, tip_variant<L, BLOCK, MINB>, BLOCK, 0));
    const unsigned grid = (unsigned)((a.B + BLOCK - 1) / BLOCK);
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int w = 0; w < 3; ++w) tip_variant<L, BLOCK, MINB><<<grid, BLOCK>>>(rb, a);
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        CK(cudaEventRecord(e0));
        tip_variant<L, BLOCK, MINB><<<grid, BLOCK>>>(rb, a);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    printf("%-10s block %3d minb %2d regs %3d local %4zu occ %2d blocks (%2d warps/SM)  %.3f ms  %.3e cfg/s\n", tag, BLOCK, MINB,
           fa.numRegs, fa.localSizeBytes, occ, occ * BLOCK / 32, best, a.B / (best * 1e-3));
}

int main()
{
    ctr_robot_desc d = {};
    d.nsections = 2; d.max_samples = 256;
    d.init_trafo[0] = d.init_trafo[4] = d.init_trafo[8] = 1.0;
    d.sec[0].ntube = 2; d.sec[0].length = 0.06647;
    d.sec[0].stiffness[0] = d.sec[0].stiffness[1] = 5; d.sec[0].curvature[0] = d.sec[0].curvature[1] = 52.2;
    d.sec[0].radius[0] = 0.00125; d.sec[0].radius[1] = 0.001; d.sec[0].poisson[0] = d.sec[0].poisson[1] = 0.3;
    d.sec[1].ntube = 1; d.sec[1].length = 0.0667; d.sec[1].stiffness[0] = 1; d.sec[1].curvature[0] = 52.63;
    d.sec[1].radius[0] = 0.001; d.sec[1].poisson[0] = 0.3;
    KRobot rb;
    if (make_krobot(&d, &rb)) return 1;
    const int64_t B = 1000000;
    std::vector<double> jv(B * 6);
    SamplerTable tb; make_sampler_table(&d, &tb);
    for (int64_t i = 0; i < B; ++i) for (int j = 0; j < 6; ++j) jv[i * 6 + j] = tb.scale[j] * sampler_u01(20171, i, j);
    FkArgs a = {};
    double *djv, *dtip;
    CK(cudaMalloc(&djv, sizeof(double) * B * 6)); CK(cudaMalloc(&dtip, sizeof(double) * B * 12));
    CK(cudaMemcpy(djv, jv.data(), sizeof(double) * B * 6, cudaMemcpyHostToDevice));
    a.jv = djv; a.B = B; a.mode = 1; a.nsamples = 256; a.tip = dtip;
    run<128, 4>(rb, a, "base");
    run<128, 3>(rb, a, "");
    run<64, 7>(rb, a, "");
    run<64, 6>(rb, a, "");
    run<32, 14>(rb, a, "");
    run<32, 13>(rb, a, "");
    run<128, 5>(rb, a, "");
    run<128, 6>(rb, a, "");
    run<64, 8>(rb, a, "");
    run<64, 9>(rb, a, "");
    run<64, 10>(rb, a, "");
    run<64, 12>(rb, a, "");
    run<96, 6>(rb, a, "");
    run<32, 16>(rb, a, "");
    run<32, 19>(rb, a, "");
    run<32, 24>(rb, a, "");
    run<256, 2>(rb, a, "");
    run<256, 3>(rb, a, "");
    std::vector<double> tip(12);
    CK(cudaMemcpy(tip.data(), dtip, 96, cudaMemcpyDeviceToHost));
    printf("tip[0] translation %.15g %.15g %.15g\n", tip[9], tip[10], tip[11]);
    return 0;
}
