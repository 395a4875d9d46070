// ctr_fk_kernels_extra.cu — callers of the hot path that live in the same kernels:
//   * batched finite-difference tip Jacobian (robot_i.h:917-960, :41-56): 1 + (T-1) + S solves
//     per configuration, all inside one thread so the joint vector is read once.
#include "ctr_fk_kernels.cuh"
#include "ctr_fk_internal.h"

namespace ctrk {

// column = d(lo -> hi)/dq (robot_i.h:41-56): translation difference and the vee of
// (R_hi - R_lo) R_hi^T, both divided by dq.
__device__ __forceinline__ void jac_column(const double* lo, const double* hi, double dq, double* col)
{
    double d[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) d[i] = hi[i] - lo[i];
    // M(i,j) = sum_k D(i,k) R_hi(j,k)
    auto M = [&](int i, int j) { return d[i] * hi[j] + d[3 + i] * hi[3 + j] + d[6 + i] * hi[6 + j]; };
    col[0] = (hi[9] - lo[9]) / dq;
    col[1] = (hi[10] - lo[10]) / dq;
    col[2] = (hi[11] - lo[11]) / dq;
    col[3] = ((M(2, 1) - M(1, 2)) * 0.5) / dq;
    col[4] = ((M(0, 2) - M(2, 0)) * 0.5) / dq;
    col[5] = ((M(1, 0) - M(0, 1)) * 0.5) / dq;
}

template <class L>
__global__ void __launch_bounds__(kBlock)
fk_jacobian_kernel(const __grid_constant__ KRobot rb, const __grid_constant__ JacArgs a)
{
    constexpr int S = L::S, T = L::T, NJ = L::NJ;
    const int64_t cfg = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (cfg >= a.B) return;
    double q[NJ];
    load_jv<L>(a.jv, cfg, q);
    double* jac = a.jac + cfg * (NJ * 6);
    const double qnan = __longlong_as_double(0x7ff8000000000000ll);

    Ctl<L> c;
    setup_control<L>(rb, q, 1, 0.0, a.nsamples, c);
    if (c.status) {
        for (int e = 0; e < NJ * 6; ++e) jac[e] = qnan;
        if (a.tip) for (int e = 0; e < 12; ++e) a.tip[cfg * 12 + e] = qnan;
        return;
    }
    double tip0[12];
    solve_fast<L, true>(rb, c, tip0, nullptr, NoSample{});
    if (a.tip) store_tf(a.tip + cfg * 12, tip0);

    // column 2 (alpha of tube 0) is zero (robot_i.h:928)
#pragma unroll
    for (int e = 0; e < 6; ++e) jac[2 * 6 + e] = 0.0;

    // perturbations 0..T-2: tube angles 1..T-1 (+fd_rot); T-1..T-2+S: section extensions
#pragma unroll 1
    for (int p = 0; p < (T - 1) + S; ++p) {
        int    jp = 0;
        double dq = a.fd_rot;
        if (p < T - 1) {
            const int t = p + 1;
            jp = 2 + L::sec_of(t) + t;
        } else {
            const int s = p - (T - 1);
            jp = L::t0(s) + s + 1;
            const double len = rb.len[s];
            double keep = 0.0;
#pragma unroll
            for (int j = 0; j < NJ; ++j) keep = (j == jp) ? q[j] : keep;
            dq = (keep + a.fd_trans >= len) ? -a.fd_trans : a.fd_trans;   // robot_i.h:947-950
        }
        double qq[NJ];
#pragma unroll
        for (int j = 0; j < NJ; ++j) qq[j] = (j == jp) ? q[j] + dq : q[j];
        Ctl<L> cp;
        setup_control<L>(rb, qq, 1, 0.0, a.nsamples, cp);
        double col[6];
        if (cp.status) {
#pragma unroll
            for (int e = 0; e < 6; ++e) col[e] = qnan;
        } else {
            double tp[12];
            solve_fast<L, true>(rb, cp, tp, nullptr, NoSample{});
            jac_column(tip0, tp, dq, col);
        }
#pragma unroll
        for (int e = 0; e < 6; ++e) jac[jp * 6 + e] = col[e];
    }

    // analytic column 0 (robot_i.h:956-959): rotation about the entry frame's z axis
    const double zx = rb.ini[6], zy = rb.ini[7], zz = rb.ini[8];
    const double rx = tip0[9] - rb.ini[9], ry = tip0[10] - rb.ini[10], rz = tip0[11] - rb.ini[11];
    jac[0] = zy * rz - zz * ry;
    jac[1] = zz * rx - zx * rz;
    jac[2] = zx * ry - zy * rx;
    jac[3] = zx; jac[4] = zy; jac[5] = zz;
}

cudaError_t launch_jacobian(int key, const KRobot& rb, const JacArgs& a, cudaStream_t st)
{
    if (a.B <= 0) return cudaSuccess;
    const unsigned grid = (unsigned)((a.B + kBlock - 1) / kBlock);
    switch (key) {
#define X(S, A, Bq, C) \
    case S * 1000 + A * 100 + Bq * 10 + C: \
        fk_jacobian_kernel<Lay<S, A, Bq, C>><<<grid, kBlock, 0, st>>>(rb, a); \
        break;
        CTRK_FOR_EACH_LAYOUT_K(X)
#undef X
        default: return cudaErrorInvalidValue;
    }
    count_launch();
    return cudaGetLastError();
}

// FP32 variant: implemented in ctr_fk_kernels_f32.cu
}  // namespace ctrk
