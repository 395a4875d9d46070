// tune_ws.cu — experiment (not part of the product): warp-specialised tip solver.
// Producer warps run the tip->base sweep (trig, curvature, torsion) and hand each sample
// (u_x,u_y,u_z,alpha) to a paired consumer warp through a shared-memory ring guarded by
// mbarriers; consumer warps run the SE(3) accumulation.  Each role needs about half the registers
// of the fused kernel, so twice as many warps are resident to hide DFMA latency.
//   nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo tools/tune_ws.cu -o tools/tune_ws
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

__device__ __forceinline__ void mbar_init(uint64_t* b, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(b)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b)
{
    asm volatile("{ .reg .b64 st; mbarrier.arrive.shared::cta.b64 st, [%0]; }" ::"r"((unsigned)__cvta_generic_to_shared(b)) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint64_t* b, unsigned parity)
{
    unsigned ok;
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok) : "r"((unsigned)__cvta_generic_to_shared(b)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, unsigned parity)
{
    int spins = 0;
    while (!mbar_try(b, parity)) { if (++spins > (1 << 24)) __trap(); }   // bounded: never hang the GPU
}

template <int PAIRS, int D>
struct Ring {
    double   data[PAIRS][D][4][32];
    uint64_t full[PAIRS][D], empty[PAIRS][D];
};

template <int PAIRS, int D>
struct RingWriter {
    Ring<PAIRS, D>* r;
    int pair, lane, n;
    __device__ __forceinline__ void operator()(int k, double ux, double uy, double uz, double al) const
    {
        const int i = n - 1 - k;                 // running sample index 0..n-1
        const int s = i % D;
        mbar_wait(&r->empty[pair][s], ((i / D) & 1) ^ 1);
        r->data[pair][s][0][lane] = ux; r->data[pair][s][1][lane] = uy;
        r->data[pair][s][2][lane] = uz; r->data[pair][s][3][lane] = al;
        __syncwarp();
        if (lane == 0) mbar_arrive(&r->full[pair][s]);
    }
};

template <class L, int PAIRS, int D>
__global__ void __launch_bounds__(PAIRS * 64)
tip_ws_kernel(const __grid_constant__ KRobot rb, const __grid_constant__ FkArgs a)
{
    __shared__ Ring<PAIRS, D> ring;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool producer = warp < PAIRS;
    const int pair = producer ? warp : warp - PAIRS;
    if (threadIdx.x < PAIRS * D) { mbar_init(&ring.full[threadIdx.x / D][threadIdx.x % D], 1); mbar_init(&ring.empty[threadIdx.x / D][threadIdx.x % D], 1); }
    __syncthreads();
    const int64_t cfg = ((int64_t)blockIdx.x * PAIRS + pair) * 32 + lane;
    const bool live = cfg < a.B;
    double q[L::NJ];
    if (live) load_jv<L>(a.jv, cfg, q);
    else { for (int j = 0; j < L::NJ; ++j) q[j] = 0.0; }
    Ctl<L> c;
    setup_control<L>(rb, q, a.mode, a.step, a.nsamples, c);      // fixed-N mode: n uniform
    const int n = c.n;
    if (producer) {
        RingWriter<PAIRS, D> w{&ring, pair, lane, n};
        solve_fast<L, false>(rb, c, nullptr, nullptr, w);
    } else {
        const StepConsts kc = make_step_consts(c.h, rb.ztol);
        QT P = qt_identity();
        double a_tip = 0.0;
        const int tipk = c.nframes - 1;
#pragma unroll 1
        for (int k = n - 1; k >= 0; --k) {
            const int i = n - 1 - k, s = i % D;
            mbar_wait(&ring.full[pair][s], (i / D) & 1);
            const double ux = ring.data[pair][s][0][lane], uy = ring.data[pair][s][1][lane];
            const double uz = ring.data[pair][s][2][lane], al = ring.data[pair][s][3][lane];
            __syncwarp();
            if (lane == 0) mbar_arrive(&ring.empty[pair][s]);
            if (k <= tipk) {
                QT E;
                se3_step_fast_k(ux, uy, uz, kc, E.w, E.x, E.y, E.z, E.px, E.py, E.pz);
                P = qt_mul(E, P);
            }
            if (k == tipk) a_tip = al;
        }
        if (live) {
            const double nn = P.w * P.w + P.x * P.x + P.y * P.y + P.z * P.z;
            const double sc = 2.0 / nn;
            double Pm[12], tip[12];
            Pm[0] = 1.0 - sc * (P.y * P.y + P.z * P.z); Pm[1] = sc * (P.x * P.y + P.w * P.z); Pm[2] = sc * (P.x * P.z - P.w * P.y);
            Pm[3] = sc * (P.x * P.y - P.w * P.z); Pm[4] = 1.0 - sc * (P.x * P.x + P.z * P.z); Pm[5] = sc * (P.y * P.z + P.w * P.x);
            Pm[6] = sc * (P.x * P.z + P.w * P.y); Pm[7] = sc * (P.y * P.z - P.w * P.x); Pm[8] = 1.0 - sc * (P.x * P.x + P.y * P.y);
            Pm[9] = P.px; Pm[10] = P.py; Pm[11] = P.pz;
            double sa, ca;
            fast_sincos(a_tip + c.theta, sa, ca);
            tf_finish(rb.ini, Pm, sa, ca, tip);
            store_tf(a.tip + cfg * 12, tip);
        }
    }
}

template <class L>
__global__ void __launch_bounds__(128, 4)
tip_ref_kernel(const __grid_constant__ KRobot rb, const __grid_constant__ FkArgs a)
{
    const int64_t i = (int64_t)blockIdx.x * 128 + threadIdx.x;
    if (i >= a.B) return;
    double q[L::NJ];
    load_jv<L>(a.jv, i, q);
    Ctl<L> c;
    setup_control<L>(rb, q, a.mode, a.step, a.nsamples, c);
    double tip[12];
    solve_fast<L, true>(rb, c, tip, nullptr, NoSample{});
    store_tf(a.tip + i * 12, tip);
}

template <class K>
float time_it(K launch)
{
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int w = 0; w < 3; ++w) launch();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    return best;
}

template <int PAIRS, int D>
void run_ws(const KRobot& rb, FkArgs a, double* ref_tip_host)
{
    using L = Lay<2, 2, 1, 0>;
    cudaFuncAttributes fa;
    CK(cudaFuncGetAttributes(&fa, tip_ws_kernel<L, PAIRS, D>));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, tip_ws_kernel<L, PAIRS, D>, PAIRS * 64, 0));
    const unsigned grid = (unsigned)((a.B + PAIRS * 32 - 1) / (PAIRS * 32));
    float ms = time_it([&] { tip_ws_kernel<L, PAIRS, D><<<grid, PAIRS * 64>>>(rb, a); });
    CK(cudaDeviceSynchronize());
    std::vector<double> tip(12 * 1000);
    CK(cudaMemcpy(tip.data(), a.tip + 12 * 123456, sizeof(double) * 12 * 1000, cudaMemcpyDeviceToHost));
    double md = 0;
    for (int i = 0; i < 12000; ++i) md = fmax(md, fabs(tip[i] - ref_tip_host[i]));
    printf("ws pairs %d depth %2d  regs %3d smem %5zu occ %2d CTAs (%2d warps/SM)  %.3f ms  %.3e cfg/s  maxdiff vs fused %.2e\n", PAIRS, D,
           fa.numRegs, fa.sharedSizeBytes, occ, occ * PAIRS * 2, ms, a.B / (ms * 1e-3), md);
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
    using L = Lay<2, 2, 1, 0>;
    float ms = time_it([&] { tip_ref_kernel<L><<<(unsigned)((B + 127) / 128), 128>>>(rb, a); });
    CK(cudaDeviceSynchronize());
    std::vector<double> ref(12 * 1000);
    CK(cudaMemcpy(ref.data(), dtip + 12 * 123456, sizeof(double) * 12 * 1000, cudaMemcpyDeviceToHost));
    printf("fused kernel: %.3f ms  %.3e cfg/s\n", ms, B / (ms * 1e-3));
    CK(cudaMemset(dtip, 0, sizeof(double) * B * 12));
    run_ws<2, 4>(rb, a, ref.data());
    run_ws<2, 8>(rb, a, ref.data());
    run_ws<4, 8>(rb, a, ref.data());
    run_ws<1, 8>(rb, a, ref.data());
    run_ws<2, 16>(rb, a, ref.data());
    return 0;
}
