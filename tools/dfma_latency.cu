// dfma_latency.cu — standalone probe (not part of the product): dependent-issue latency of DFMA and the FP64
// pipe's issue interval on one SM sub-partition, to read the tip kernel's 78 % FP64-pipe figure against.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/dfma_latency.cu -o tools/dfma_latency && tools/dfma_latency
#include <cstdio>
#include <cuda_runtime.h>

template <int CHAINS>
__global__ void probe(double* out, long long* cycles, int iters, double a, double b)
{
    double x[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) x[c] = a + c + threadIdx.x;
    const long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 16; ++r)
#pragma unroll
            for (int c = 0; c < CHAINS; ++c) x[c] = fma(x[c], a, b);
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += x[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// DFMAs mixed with independent integer instructions: does an ALU instruction hide under the 2-cycle FP64 issue
// interval (pipe-limited: 2 cycles per DFMA) or does it take an issue cycle of its own (2 per DFMA + 1 per ALU)?
template <int CHAINS, int INTS>
__global__ void probe_mixed(double* out, long long* cycles, int iters, double a, double b, unsigned k)
{
    double x[CHAINS];
    unsigned y[INTS > 0 ? INTS : 1];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) x[c] = a + c + threadIdx.x;
#pragma unroll
    for (int c = 0; c < INTS; ++c) y[c] = k + c + threadIdx.x;
    const long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 16; ++r) {
#pragma unroll
            for (int c = 0; c < CHAINS; ++c) x[c] = fma(x[c], a, b);
#pragma unroll
            for (int c = 0; c < INTS; ++c) y[c] = (y[c] ^ k) + (y[c] >> 3);      // LOP3 / IADD3 / SHF, no FP64
        }
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += x[c];
#pragma unroll
    for (int c = 0; c < INTS; ++c) s += (double)y[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int CHAINS, int INTS>
void run_mixed(double* out, long long* cyc)
{
    const int iters = 2000, warps = 16;
    for (int rep = 0; rep < 2; ++rep) { probe_mixed<CHAINS, INTS><<<1, 32 * warps>>>(out, cyc, iters, 0.999999, 1e-9, 12345u); cudaDeviceSynchronize(); }
    long long c;
    cudaMemcpy(&c, cyc, sizeof c, cudaMemcpyDeviceToHost);
    const double per = (double)c / (iters * 16.0) / 4.0;      // cycles per round per warp of a sub-partition (4 warps each)
    printf("16 warps, per round %d DFMA + %d integer statements (%d ALU instructions): %.2f cycles per warp-round  (FP64 pipe alone: %d)\n",
           CHAINS, INTS, 2 * INTS, per, 2 * CHAINS);
}

template <int CHAINS>
void run(int warps, double* out, long long* cyc)
{
    const int iters = 2000;
    probe<CHAINS><<<1, 32 * warps>>>(out, cyc, iters, 0.999999, 1e-9);
    cudaDeviceSynchronize();
    probe<CHAINS><<<1, 32 * warps>>>(out, cyc, iters, 0.999999, 1e-9);
    cudaDeviceSynchronize();
    long long c;
    cudaMemcpy(&c, cyc, sizeof c, cudaMemcpyDeviceToHost);
    const double per = (double)c / (iters * 16.0);      // cycles per round of CHAINS DFMAs per warp
    printf("warps %2d (%.1f per sub-partition)  chains %d : %.2f cycles per dependent step, %.2f cycles per DFMA per sub-partition\n", warps,
           warps / 4.0, CHAINS, per, per / (CHAINS * (warps < 4 ? 1.0 : warps / 4.0)));
}

int main()
{
    double* out; long long* cyc;
    cudaMalloc(&out, sizeof(double) * 2048); cudaMalloc(&cyc, sizeof(long long) * 8);
    run<1>(1, out, cyc);      // pure latency
    run<2>(1, out, cyc);
    run<4>(1, out, cyc);
    run<8>(1, out, cyc);
    run<1>(4, out, cyc);      // one warp per sub-partition
    run<2>(16, out, cyc);     // the tip kernel's shape: 4 warps per sub-partition, ~2 chains each
    run<1>(16, out, cyc);
    run<4>(16, out, cyc);
    run_mixed<4, 0>(out, cyc);
    run_mixed<4, 1>(out, cyc);
    run_mixed<4, 2>(out, cyc);
    run_mixed<4, 4>(out, cyc);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
