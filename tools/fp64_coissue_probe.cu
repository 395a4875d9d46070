// fp64_coissue_probe.cu — standalone probe (not part of the product): with the FP64 pipe saturated (8 independent
// DFMA chains per warp, W warps per sub-partition), do integer / select / move instructions issue in the shadow of
// the DFMAs (2 pipe cycles each), or does every one of them cost an issue cycle of its own?
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/fp64_coissue_probe.cu -o tools/fp64_coissue_probe
// Prints cycles per round of (8 DFMA + K other instructions) per sub-partition: 16.0 + 0 K = shadow, 16 + K = serial.
#include <cstdio>
#include <cuda_runtime.h>

template <int K, int KIND>
__global__ void probe(double* out, long long* cycles, int iters, double a, double b, int s)
{
    double x[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) x[c] = a + c + threadIdx.x;
    int u[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) u[c] = s + c * 7 + threadIdx.x;
    float f[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) f[c] = (float)(s + c);
    const long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                x[c] = fma(x[c], a, b);
                if (c < K) {
                    if (KIND == 0) u[c] = (u[c] ^ s) + i;                       // LOP3 + IADD (counted as 2)
                    if (KIND == 1) u[c] = (u[c] > s) ? u[(c + 1) & 7] : u[c];   // ISETP + SEL
                    if (KIND == 2) f[c] = fmaf(f[c], 0.999f, 1e-3f);            // FFMA
                }
            }
        }
    }
    const long long t1 = clock64();
    double sum = 0;
#pragma unroll
    for (int c = 0; c < 8; ++c) sum += x[c] + u[c] + f[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = sum;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int K, int KIND>
void run(int wps, double* out, long long* cyc)
{
    const int iters = 2000;
    for (int rep = 0; rep < 2; ++rep) { probe<K, KIND><<<1, 128 * wps>>>(out, cyc, iters, 0.999999, 1e-9, 12345); cudaDeviceSynchronize(); }
    long long c;
    cudaMemcpy(&c, cyc, sizeof c, cudaMemcpyDeviceToHost);
    const char* kn[] = {"LOP3+IADD (2 instr each)", "ISETP+SEL (2 instr each)", "FFMA (1 instr each)"};
    printf("kind %-26s K %d per 8 DFMA, %d warps/sub-partition: %.2f cycles per round of 8 DFMA per sub-partition (16.0 = FP64 pipe limit)\n",
           kn[KIND], K, wps, (double)c / (iters * 4.0) / wps);
}

int main()
{
    double* out; long long* cyc;
    cudaMalloc(&out, sizeof(double) * 4096); cudaMalloc(&cyc, sizeof(long long) * 8);
    run<0, 0>(4, out, cyc);
    run<2, 0>(4, out, cyc); run<4, 0>(4, out, cyc); run<8, 0>(4, out, cyc);
    run<2, 1>(4, out, cyc); run<4, 1>(4, out, cyc); run<8, 1>(4, out, cyc);
    run<2, 2>(4, out, cyc); run<4, 2>(4, out, cyc); run<8, 2>(4, out, cyc);
    run<0, 0>(2, out, cyc); run<4, 0>(2, out, cyc); run<8, 0>(2, out, cyc); run<8, 2>(2, out, cyc);
    return 0;
}
