// store_probe.cu — experiment (not part of the product): how fast can a warp-per-configuration
// kernel write [B][256][12] doubles with 256-bit stores, depending on which lane writes which frame?
//   blocked:     lane l writes frames l*8 .. l*8+7   (lane stride 768 B)  <- the backbone kernels
//   interleaved: lane l writes frames l, l+32, ...   (lane stride 96 B)
//   linear:      lane l writes 32-B pieces l, l+32, ... of the 24 KB block (lane stride 32 B)
//   nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a tools/store_probe.cu -o tools/store_probe
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ void st256(double* p, double a, double b, double c, double d)
{
    asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}

template <int MODE>
__global__ void __launch_bounds__(64) probe(double* out, long long B, int per_warp)
{
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    for (int j = 0; j < per_warp; ++j) {
        const long long cfg = warp * per_warp + j;
        if (cfg >= B) return;
        double* base = out + cfg * 3072;
        const double v = (double)cfg;
        if (MODE == 0) {
#pragma unroll 2
            for (int i = 0; i < 8; ++i) {
                double* p = base + (lane * 8 + i) * 12;
                st256(p, v, v, v, v); st256(p + 4, v, v, v, v); st256(p + 8, v, v, v, v);
            }
        } else if (MODE == 1) {
#pragma unroll 2
            for (int i = 0; i < 8; ++i) {
                double* p = base + (lane + 32 * i) * 12;
                st256(p, v, v, v, v); st256(p + 4, v, v, v, v); st256(p + 8, v, v, v, v);
            }
        } else {
#pragma unroll 4
            for (int i = 0; i < 24; ++i) st256(base + (lane + 32 * i) * 4, v, v, v, v);
        }
    }
}

template <int MODE>
void run(double* out, long long B, const char* tag)
{
    const int per_warp = 32;
    const long long warps = (B + per_warp - 1) / per_warp;
    const unsigned grid = (unsigned)((warps + 1) / 2);
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int w = 0; w < 2; ++w) probe<MODE><<<grid, 64>>>(out, B, per_warp);
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        CK(cudaEventRecord(e0)); probe<MODE><<<grid, 64>>>(out, B, per_warp); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    printf("%-12s %.3f ms  %.0f GB/s\n", tag, best, B * 3072.0 * 8 / (best * 1e-3) / 1e9);
}

int main()
{
    const long long B = 200000;
    double* out;
    CK(cudaMalloc(&out, sizeof(double) * B * 3072));
    run<0>(out, B, "blocked");
    run<1>(out, B, "interleaved");
    run<2>(out, B, "linear");
    return 0;
}
