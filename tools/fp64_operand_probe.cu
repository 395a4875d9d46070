// fp64_operand_probe.cu — standalone probe (not part of the product): does the FP64 pipe keep its 2-cycle issue
// interval when every DFMA reads three distinct register pairs (the solver's case) rather than one register and two
// shared multiplicands (the peak probe's case)?  One CTA per SM sub-partition group, W warps per sub-partition.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/fp64_operand_probe.cu -o tools/fp64_operand_probe
#include <cstdio>
#include <cuda_runtime.h>

// MODE 0: x = fma(x, a, b)           one register pair + two shared pairs
// MODE 1: x = fma(y, z, x)           three distinct pairs per chain
// MODE 2: x = fma(x, c[k], y)        constant-bank multiplicand
// MODE 3: DMUL/DADD/DFMA mix on distinct pairs: t = y*z; x = x + t; x = fma(y, x, z)
template <int CHAINS, int MODE>
__global__ void probe(double* out, long long* cycles, int iters, double a, double b)
{
    double x[CHAINS], y[CHAINS], z[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) { x[c] = a + c + threadIdx.x; y[c] = 1.0 - 1e-9 * (c + threadIdx.x); z[c] = 1e-9 * (c + 1); }
    const long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int c = 0; c < CHAINS; ++c) {
                if (MODE == 0) x[c] = fma(x[c], a, b);
                if (MODE == 1) x[c] = fma(y[c], z[c], x[c]);
                if (MODE == 2) x[c] = fma(x[c], 0.999999, y[c]);
                if (MODE == 3) { const double t = y[c] * z[c]; x[c] = x[c] + t; x[c] = fma(y[c], x[c], z[c]); }
            }
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += x[c] + y[c] + z[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int CHAINS, int MODE>
void run(int warps_per_smsp, double* out, long long* cyc)
{
    const int iters = 2000, warps = 4 * warps_per_smsp;
    for (int rep = 0; rep < 2; ++rep) { probe<CHAINS, MODE><<<1, 32 * warps>>>(out, cyc, iters, 0.999999, 1e-9); cudaDeviceSynchronize(); }
    long long c;
    cudaMemcpy(&c, cyc, sizeof c, cudaMemcpyDeviceToHost);
    const int per_round = (MODE == 3 ? 3 : 1) * CHAINS;
    const double per = (double)c / (iters * 8.0) / (per_round * warps_per_smsp);
    printf("mode %d  chains %d  warps/sub-partition %d : %.3f cycles per FP64 instruction per sub-partition (2.0 = pipe limit)\n", MODE, CHAINS,
           warps_per_smsp, per);
}

int main()
{
    double* out; long long* cyc;
    cudaMalloc(&out, sizeof(double) * 4096); cudaMalloc(&cyc, sizeof(long long) * 8);
    run<8, 0>(4, out, cyc); run<8, 1>(4, out, cyc); run<8, 2>(4, out, cyc); run<8, 3>(4, out, cyc);
    run<2, 0>(4, out, cyc); run<2, 1>(4, out, cyc); run<2, 2>(4, out, cyc); run<2, 3>(4, out, cyc);
    run<1, 1>(4, out, cyc); run<1, 1>(5, out, cyc); run<1, 1>(6, out, cyc); run<1, 1>(8, out, cyc);
    run<2, 1>(3, out, cyc); run<2, 1>(5, out, cyc); run<2, 1>(6, out, cyc); run<2, 1>(8, out, cyc);
    run<4, 1>(1, out, cyc); run<4, 1>(2, out, cyc); run<4, 1>(3, out, cyc); run<4, 1>(4, out, cyc);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
