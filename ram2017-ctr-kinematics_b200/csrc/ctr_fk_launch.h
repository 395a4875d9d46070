// ctr_fk_launch.h — launcher interface between the C-ABI layer (ctr_fk_api.cu) and the kernel
// translation units.  Kernels are instantiated per section layout (14 layouts: 1-3 sections of
// 1 or 2 tubes) and selected by ctrk::layout_key().
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "ctr_fk_core.cuh"

namespace ctrk {

struct FkArgs {
    const double*  jv;          // [B][NJ]
    int64_t        B;
    int            mode;        // CTR_FK_MODE_*
    double         step;
    int            nsamples;
    const int32_t* order;       // nullable: order[i] = configuration solved by thread i
    double*        tip;         // [B][12]            nullable
    int32_t*       n_out;       // [B]                nullable
    double*        step_out;    // [B]                nullable
    double*        alpha_base;  // [B][T]             nullable
    int32_t*       status;      // [B]                nullable
    double*        backbone;    // [B][maxs][12] or [maxs][12][B]   nullable
    double*        radius;      // [B][maxs]     or [maxs][B]       nullable
    int            soa;         // backbone/radius layout
    int            fine;        // tip / gather kernels: finely sampled launch (fine_sampling): the FINE form of the sweep
    int            windows;     // tip kernels, with `order`: number of binning windows; CTAs then walk the
                                // windows rank-major (every window's longest sweeps first), 0 = linear
    // solve + all-gather in one kernel (ctr_fk_batch_gather): every pose goes to row gather_row + cfg of each
    // destination — peer-mapped gathered buffers (NVLink P2P stores) or one NVSwitch multicast mapping
    double*        gather[8];
    int            ngather;     // 0: plain `tip` output
    int            gather_mc;   // destinations are multicast addresses: multimem.st
    int            gather_rowwise; // every thread stores its own row (CTR_FK_FLAG_GATHER_ROWWISE)
    int64_t        gather_row;
};

struct JacArgs {
    const double* jv;
    int64_t       B;
    int           nsamples;
    double        fd_trans, fd_rot;
    double*       jac;          // [B][NJ][6]
    double*       tip;          // [B][12] nullable
    int           mode;         // CTR_FK_MODE_*: STEP = the ArcLengthStep form (robot_i.h:901-907)
    double        step;
    int           i_sample;     // >= 0: also the Jacobian of that sample's frame (robot_i.h:1012-1072)
    double*       jac_sample;   // [B][NJ][6] nullable
    double*       sample;       // [B][12] nullable
    int32_t*      n_out;        // [B] nullable
    int           given;        // tip / sample are inputs (CTR_FK_FLAG_JAC_GIVEN_POSES)
    const int32_t* order;       // nullable: visiting order (step mode, window-local binning)
};

struct StabArgs {
    const double* jv;
    int64_t       B;
    double        angle_step, arc_step, min_degree;
    int32_t*      stable;       // [B] nullable
    double*       degree;       // [B] nullable
    unsigned long long* ticket; // zero-initialised work counter
};

struct ShootArgs {
    const double*       jv_base;
    int64_t             B;
    int                 mode;
    double              step;
    int                 nsamples;
    double              tol;
    int                 max_iter;
    const int32_t*      order;      // nullable
    double*             jv_tip;     // [B][NJ]
    int32_t*            iters;      // nullable
    int32_t*            status;     // nullable
    unsigned long long* ticket;     // zero-initialised work counter
    const double*       jv_guess;   // [B][NJ] nullable: starting tip angles (warm start)
    double*             resid;      // [B] nullable: residual max-norm at the returned tip angles
    int                 rk4;        // CTR_FK_FLAG_RK4: continuum model integrated by RK4 instead of the reference sweep
};

constexpr int kBlock = 128;     // threads per CTA for every solver kernel
// Resident CTAs per SM the FAST tip solver is compiled for.  The sweep of a 3-tube robot fits 128 registers (4 CTAs,
// 16 warps/SM); from 4 tubes (or 3 sections) on the state no longer does and a 128-register build spills and rematerialises.  The
// kernel's rate does not depend on the warp count between 8 and 16 warps/SM (it is bound by the FP64 pipe, not by
// latency: tools/tune_tip.cu, 1M configurations, N=256: 3 tubes 2.008 / 1.998 ms at 12 / 16 warps; 4 tubes 2.388 /
// 2.494; 5 tubes 2.794 / 2.912; 6 tubes 3.289 / 3.463; 3 tubes in 3 sections 2.074 / 2.089 with an 8-byte spill at 128 registers), so
// those layouts get up to 168 registers and 3 CTAs.  Round 2 (loop in units of the step, two iterations per trip, tools/tune_tip.cu):
// 3 tubes 1.846 / 1.858 ms at 12 / 16 warps, 3 tubes in 3 sections 1.964 / 1.957, 4 tubes 2.341 / 2.340, 5 tubes 2.664 / 2.767 / 2.718
// at 8 / 12 / 16 warps: three-tube layouts go to 3 CTAs as well, 5 and 6 tubes to 2 CTAs (up to 255 registers).
constexpr int tip_blocks_per_sm(int ntubes, int nsections) { return ntubes >= 5 ? 2 : ((ntubes >= 3 || nsections >= 3) ? 3 : 4); }
#ifdef CTRK_BB_MINB      /* tools/tune_bb.cu only */
constexpr int backbone_blocks_per_sm(int, int) { return CTRK_BB_MINB; }
#else
// round 2 (tools/tune_bb.cu, ctr_sec3_tub4, 1M x 256 frames): 7.49 / 6.95 / 6.46 ms at 4 / 3 / 2 CTAs per SM (128 / 168 / 224
// registers) and 6.38 ms with two iterations per trip at 2 CTAs: the emitter's 12-double frame, the walked product and
// the sweep state want ~220 registers; with them the constant loads and register copies of the 168-register build go
constexpr int backbone_blocks_per_sm(int ntubes, int nsections) { return (ntubes >= 4 || nsections >= 3) ? 2 : 3; }
#endif
// A launch is finely sampled when no iteration of its sweeps is expected outside the hot tier: the largest step of
// the launch (the arc-length step itself, or MaxLength / N) times a bound on the torsion a tube can build up,
// (1+nu) kappa^2 L, stays below 0.4 rad — N >= 160 on the shipped robots.  Only speed depends on it (which of two
// forms of the same loop runs, FastSweep::run_scaled); it is a property of the call's arguments and the robot, never
// of the batch, so every chunk and every rank of one logical call takes the same form.
inline bool fine_sampling(const KRobot& rb, int mode, double step, int nsamples)
{
    double len = 0.0, kmax = 0.0, onu = 1.0;
    for (int s = 0; s < rb.nsec; ++s) { len += rb.len[s]; kmax = kmax > (rb.kappa[s] < 0 ? -rb.kappa[s] : rb.kappa[s]) ? kmax : (rb.kappa[s] < 0 ? -rb.kappa[s] : rb.kappa[s]); }
    for (int t = 0; t < kMaxTube; ++t) onu = onu > rb.onu[t] ? onu : rb.onu[t];
    const int n = nsamples < rb.maxs ? nsamples : rb.maxs;
    const double h = (mode == 0) ? step : (n > 0 ? len / (double)n : len);
    return h * onu * kmax * kmax * len <= 0.4;
}
constexpr int kBinWindow = 4096;        // step-mode binning sorts windows of this many consecutive configurations

// tip-only solvers (ctr_fk_kernels_fast.cu / ctr_fk_kernels_strict.cu)
cudaError_t launch_tip_fast(int key, const KRobot& rb, const FkArgs& a, cudaStream_t st);
cudaError_t launch_tip_gather(int key, const KRobot& rb, const FkArgs& a, cudaStream_t st);
cudaError_t launch_tip_strict(int key, const KRobot& rb, const FkArgs& a, cudaStream_t st);
cudaError_t launch_tip_f32(int key, const KRobot& rb, const FkArgs& a, cudaStream_t st);
// backbone solvers (two passes inside one kernel; the output buffer doubles as scratch)
cudaError_t launch_backbone_fast(int key, const KRobot& rb, const FkArgs& a, cudaStream_t st);
cudaError_t launch_backbone_fast_soa(int key, const KRobot& rb, const FkArgs& a, cudaStream_t st);
cudaError_t launch_backbone_fast_aos16(int key, const KRobot& rb, const FkArgs& a, cudaStream_t st);
cudaError_t launch_backbone_strict(int key, const KRobot& rb, const FkArgs& a, cudaStream_t st);
// finite-difference Jacobian (fast arithmetic)
cudaError_t launch_jacobian(int key, const KRobot& rb, const JacArgs& a, cudaStream_t st);
// calcStability / calcDegreeStability scan (fast arithmetic, sweep only)
cudaError_t launch_stability(int key, const KRobot& rb, const StabArgs& a, int sm_count, cudaStream_t st);
cudaError_t launch_jacobian_strict(int key, const KRobot& rb, const JacArgs& a, cudaStream_t st);
cudaError_t launch_stability_strict(int key, const KRobot& rb, const StabArgs& a, int sm_count, cudaStream_t st);
// base-angle shooting solve: persistent kernel with a ticket queue (grid sized to the device)
cudaError_t launch_shoot(int key, const KRobot& rb, const ShootArgs& a, int sm_count, cudaStream_t st);

// step-mode binning: `order` = every window of kBinWindow = 4096 consecutive configurations counting-sorted by
// sample count, longest sweeps first (ctr_fk_kernels_util.cu)
cudaError_t launch_bin_local(int key, const KRobot& rb, const double* jv, int64_t B, int mode, double step,
                             int nsamples, int32_t* order, cudaStream_t st);

}  // namespace ctrk
#include "../host/ctr_sampler.h"
namespace ctrk {
struct SamplerArgs {
    uint64_t     seed;
    int64_t      first;
    int64_t      B;
    SamplerTable tb;
    double*      jv;
};
cudaError_t launch_sampler(const SamplerArgs& a, cudaStream_t st);
cudaError_t launch_fp64_probe(int iters, int blocks, double* sink, cudaStream_t st);
cudaError_t launch_fp32x2_probe(int iters, int blocks, double* sink, cudaStream_t st);
constexpr int kProbeThreads = 256;
constexpr int kProbeChains  = 8;

}  // namespace ctrk

// all section layouts: (sections, tubes in section 0, 1, 2)
#define CTRK_FOR_EACH_LAYOUT_K(X) \
    X(1, 1, 0, 0) X(1, 2, 0, 0) \
    X(2, 1, 1, 0) X(2, 1, 2, 0) X(2, 2, 1, 0) X(2, 2, 2, 0) \
    X(3, 1, 1, 1) X(3, 1, 1, 2) X(3, 1, 2, 1) X(3, 1, 2, 2) \
    X(3, 2, 1, 1) X(3, 2, 1, 2) X(3, 2, 2, 1) X(3, 2, 2, 2)
