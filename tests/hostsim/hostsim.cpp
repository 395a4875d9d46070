// hostsim.cpp — TEST HARNESS ONLY.  Compiles the kernels' per-configuration solvers
// (ram2017-ctr-kinematics_b200/csrc/ctr_fk_core.cuh, the code each GPU thread runs) for the CPU
// so that their logic can be compared with the oracle in a container without a GPU.  Not part of
// the product: libctrfk.so has no CPU path and never links this file.
#include <cstdint>
#include <cstdlib>
#include <cstring>

#include "../../include/ctr_fk.h"
#include "../../ram2017-ctr-kinematics_b200/csrc/ctr_fk_krobot.h"
#include "../../ram2017-ctr-kinematics_b200/csrc/ctr_fk_core_f32x2.cuh"

using namespace ctrk;

namespace {
struct Park {
    double* base;   // [maxs][4]
    void operator()(int k, double ux, double uy, double uz, double al) const
    {
        if (!base) return;
        base[4 * k] = ux; base[4 * k + 1] = uy; base[4 * k + 2] = uz; base[4 * k + 3] = al;
    }
};

template <class L>
void run(const KRobot& rb, const double* jv, int64_t B, int mode, double step, int ns, int fast, double* tip,
         int32_t* n_out, double* step_out, double* alpha_base, double* samples)
{
    if (fast == 3) {        // packed FP32 variant: configurations (2i, 2i+1) as one pair, fixed-N mode
        for (int64_t i = 0; i < B; i += 2) {
            Ctl<L> c0, c1;
            setup_control<L>(rb, jv + i * L::NJ, mode, step, ns, c0);
            const bool has1 = i + 1 < B;
            if (has1) setup_control<L>(rb, jv + (i + 1) * L::NJ, mode, step, ns, c1); else c1 = c0;
            if (n_out) { n_out[i] = c0.nframes; if (has1) n_out[i + 1] = c1.nframes; }
            if (step_out) { step_out[i] = c0.h; if (has1) step_out[i + 1] = c1.h; }
            const bool ok0 = c0.status == 0, ok1 = has1 && c1.status == 0;
            if (!ok0 && !ok1) continue;
            double t0[12], t1[12];
            if (!ok0) c0 = c1;
            if (!ok1) c1 = c0;
            solve_f32x2<L>(rb, c0, c1, ok0 ? t0 : nullptr, ok1 ? t1 : nullptr);
            if (tip && ok0) std::memcpy(tip + 12 * i, t0, sizeof t0);
            if (tip && ok1) std::memcpy(tip + 12 * (i + 1), t1, sizeof t1);
        }
        return;
    }
    for (int64_t i = 0; i < B; ++i) {
        Ctl<L> c;
        setup_control<L>(rb, jv + i * L::NJ, mode, step, ns, c);
        if (n_out) n_out[i] = c.nframes;
        if (step_out) step_out[i] = c.h;
        if (c.status) continue;
        Park park{samples ? samples + i * 4 * (int64_t)rb.maxs : nullptr};
        double t[12], ab[L::T];
        if (fast == 2) {       // FP32 variant (no alpha_base / samples)
            solve_f32<L>(rb, c, t);
            for (int q = 0; q < L::T; ++q) ab[q] = 0.0;
        } else if (fast == 4) {   // nobody looks at the samples: the loop in units of the step (FastSweep::run_scaled<true>)
            solve_fast<L, true>(rb, c, t, ab, NoSample{});
        } else if (fast == 5) {   // ... without the tube angles as an output (run_scaled<false>)
            solve_fast<L, true>(rb, c, t, nullptr, NoSample{});
            for (int q = 0; q < L::T; ++q) ab[q] = 0.0;
        } else if (fast == 6) {   // the FINE form of that loop (what finely sampled launches of the tip kernel run)
            solve_fast<L, true, NoSample, true>(rb, c, t, ab, NoSample{});
        } else if (fast == 7) {
            solve_fast<L, true, NoSample, true>(rb, c, t, nullptr, NoSample{});
            for (int q = 0; q < L::T; ++q) ab[q] = 0.0;
        } else if (fast) solve_fast<L, true>(rb, c, t, ab, park);
        else             solve_strict<L, true>(rb, c, t, ab, park);
        if (tip) std::memcpy(tip + 12 * i, t, sizeof t);
        if (alpha_base) std::memcpy(alpha_base + L::T * i, ab, sizeof ab);
    }
}
}  // namespace

extern "C" int hostsim_fk(const ctr_robot_desc* d, const double* jv, int64_t B, int mode, double step, int ns,
                          int fast, double* tip, int32_t* n_out, double* step_out, double* alpha_base,
                          double* samples)
{
    KRobot rb;
    int rc = make_krobot(d, &rb);
    if (rc) return rc;
    switch (layout_key(d)) {
#define X(S, A, Bq, C) case S * 1000 + A * 100 + Bq * 10 + C: run<Lay<S, A, Bq, C>>(rb, jv, B, mode, step, ns, fast, tip, n_out, step_out, alpha_base, samples); break;
        CTRK_FOR_EACH_LAYOUT(X)
#undef X
        default: return CTR_FK_E_DESC;
    }
    return 0;
}

extern "C" void hostsim_sincos(const double* x, int64_t n, double* s, double* c)
{
    for (int64_t i = 0; i < n; ++i) fast_sincos(x[i], s[i], c[i]);
}

// SE(3) step: strict 3x4 (column-major) and fast (quaternion w,x,y,z + translation)
extern "C" void hostsim_se3(double ux, double uy, double uz, double h, double* strict12, double* fast7)
{
    tf_step_strict(ux, uy, uz, h, 1e-12, strict12);
    se3_step_fast(ux, uy, uz, h, 1e-12, fast7[0], fast7[1], fast7[2], fast7[3], fast7[4], fast7[5], fast7[6]);
}

// Base-angle shooting on the CPU build of the device code (ShootState: damped Newton, one sweep per round):
// jv_base holds base angles in the alpha slots, jv_guess (nullable) the starting tip angles; returns tip-angle
// joint vectors in jv_tip, sweep counts, status and the residual at the returned angles.
namespace {
template <class L, bool RK4>
void shoot_run(const KRobot& rb, const double* jvb, const double* guess, int64_t B, int mode, double step, int ns,
               double tol, int maxit, double* jv_tip, int32_t* iters, int32_t* status, double* resid)
{
    for (int64_t i = 0; i < B; ++i) {
        ShootState<L, RK4> s;
        s.start(jvb + i * L::NJ, guess ? guess + i * L::NJ : nullptr);
        if (const char* e = std::getenv("HOSTSIM_SHOOT_MAX_REJECT")) s.max_reject = std::atoi(e);
        if (const char* e = std::getenv("HOSTSIM_SHOOT_RESTARTS")) s.restarts = !guess && std::atoi(e) != 0;
        int st;
        while ((st = s.round(rb, mode, step, ns, tol, maxit)) == ShootState<L, RK4>::kRunning) {}
        for (int j = 0; j < L::NJ; ++j) jv_tip[i * L::NJ + j] = s.q[j];
        iters[i] = s.it; status[i] = st;
        if (resid) resid[i] = s.res;
    }
}
}  // namespace

extern "C" int hostsim_shoot(const ctr_robot_desc* d, const double* jvb, const double* guess, int64_t B, int mode,
                             double step, int ns, double tol, int maxit, int rk4, double* jv_tip, int32_t* iters,
                             int32_t* status, double* resid)
{
    KRobot rb;
    int rc = make_krobot(d, &rb);
    if (rc) return rc;
    switch (layout_key(d)) {
#define X(S, A, Bq, C) case S * 1000 + A * 100 + Bq * 10 + C: \
        if (rk4) shoot_run<Lay<S, A, Bq, C>, true>(rb, jvb, guess, B, mode, step, ns, tol, maxit, jv_tip, iters, status, resid); \
        else     shoot_run<Lay<S, A, Bq, C>, false>(rb, jvb, guess, B, mode, step, ns, tol, maxit, jv_tip, iters, status, resid); \
        break;
        CTRK_FOR_EACH_LAYOUT(X)
#undef X
        default: return CTR_FK_E_DESC;
    }
    return 0;
}

// tangent-linear sweep of one configuration: base angles and d base / d tip angles
namespace {
template <class L>
int tangent_run(const KRobot& rb, const double* jv, int mode, double step, int ns, int rk4, double* base, double* jac)
{
    Ctl<L> c;
    setup_control<L>(rb, jv, mode, step, ns, c);
    if (c.status) return 1;
    if (rk4) sweep_tangent_rk4<L>(rb, c, base, L::T > 1 ? jac : nullptr);
    else     sweep_tangent<L>(rb, c, base, L::T > 1 ? jac : nullptr);
    return 0;
}
}  // namespace

extern "C" int hostsim_tangent(const ctr_robot_desc* d, const double* jv, int mode, double step, int ns, int rk4,
                               double* base, double* jac)
{
    KRobot rb;
    int rc = make_krobot(d, &rb);
    if (rc) return rc;
    switch (layout_key(d)) {
#define X(S, A, Bq, C) case S * 1000 + A * 100 + Bq * 10 + C: return tangent_run<Lay<S, A, Bq, C>>(rb, jv, mode, step, ns, rk4, base, jac);
        CTRK_FOR_EACH_LAYOUT(X)
#undef X
        default: return CTR_FK_E_DESC;
    }
}

namespace {
struct FrameStoreHost {
    double* frames;          // [maxs][12]
    void operator()(int k, const double* F) const { std::memcpy(frames + 12 * k, F, 12 * sizeof(double)); }
};
}  // namespace

// Frame emission of the backbone kernel: a first sweep for the total product, a second sweep
// emitting frames one sample late (FastSweep::run_emit).
namespace {
template <class L>
void emit_full_all(const KRobot& rb, const double* jv, int64_t B, int mode, double step, int ns, double* frames)
{
    for (int64_t i = 0; i < B; ++i) {
        Ctl<L> c;
        setup_control<L>(rb, jv + i * L::NJ, mode, step, ns, c);
        if (c.status) continue;
        FastSweep<L, true, NoSample> s1(rb, c, NoSample{});
        s1.run();
        QT T;
        T.w = s1.Qw; T.x = s1.Qx; T.y = s1.Qy; T.z = s1.Qz; T.px = s1.Px; T.py = s1.Py; T.pz = s1.Pz;
        FrameStoreHost fs{frames + i * 12 * (int64_t)rb.maxs};
        FastSweep<L, false, PendingFrameEmitter<FrameStoreHost>, true> s2(rb, c, make_frame_emitter<L>(rb, c, fs, T));
        s2.run_emit();
    }
}
}  // namespace

extern "C" int hostsim_emit_full(const ctr_robot_desc* d, const double* jv, int64_t B, int mode, double step, int ns,
                                 double* frames)
{
    KRobot rb;
    int rc = make_krobot(d, &rb);
    if (rc) return rc;
    switch (layout_key(d)) {
#define X(S, A, Bq, C) case S * 1000 + A * 100 + Bq * 10 + C: emit_full_all<Lay<S, A, Bq, C>>(rb, jv, B, mode, step, ns, frames); break;
        CTRK_FOR_EACH_LAYOUT(X)
#undef X
        default: return CTR_FK_E_DESC;
    }
    return 0;
}

// jacobian_config (what one GPU thread of fk_jacobian_kernel runs), B configurations
namespace {
template <class L>
void jacobian_all(const KRobot& rb, const double* jv, int64_t B, int mode, double step, int ns, int ks, double fdt, double fdr,
                  double* jac, double* jacs, double* tip, double* smp, int32_t* n_out, bool given)
{
    for (int64_t i = 0; i < B; ++i)
        if (ks >= 0) jacobian_config<L, true>(rb, jv + i * L::NJ, mode, step, ns, ks, fdt, fdr, jac + i * L::NJ * 6,
                           jacs ? jacs + i * L::NJ * 6 : nullptr, tip ? tip + 12 * i : nullptr, smp ? smp + 12 * i : nullptr,
                           n_out ? n_out + i : nullptr, given);
        else jacobian_config<L, false>(rb, jv + i * L::NJ, mode, step, ns, ks, fdt, fdr, jac + i * L::NJ * 6,
                           jacs ? jacs + i * L::NJ * 6 : nullptr, tip ? tip + 12 * i : nullptr, smp ? smp + 12 * i : nullptr,
                           n_out ? n_out + i : nullptr, given);
}
}  // namespace

extern "C" int hostsim_jacobian(const ctr_robot_desc* d, const double* jv, int64_t B, int mode, double step, int ns, int ks,
                                double fdt, double fdr, double* jac, double* jacs, double* tip, double* smp, int32_t* n_out, int given)
{
    KRobot rb;
    int rc = make_krobot(d, &rb);
    if (rc) return rc;
    switch (layout_key(d)) {
#define X(S, A, Bq, C) case S * 1000 + A * 100 + Bq * 10 + C: jacobian_all<Lay<S, A, Bq, C>>(rb, jv, B, mode, step, ns, ks, fdt, fdr, jac, jacs, tip, smp, n_out, given != 0); break;
        CTRK_FOR_EACH_LAYOUT(X)
#undef X
        default: return CTR_FK_E_DESC;
    }
    return 0;
}

#ifdef CTRK_COUNT_COLD   // analysis build (tools/tier_shares.py): how many sweep iterations left the hot form
extern "C" void hostsim_tier_counters(long long* out3, int reset)
{
    out3[0] = g_iter_total; out3[1] = g_iter_cold; out3[2] = g_iter_cold_delta;
    if (reset) g_iter_total = g_iter_cold = g_iter_cold_delta = 0;
}
// iterations of run_scaled's loop, and how many of them hot
extern "C" void hostsim_block_counters(long long* out2, int reset)
{
    out2[0] = g_iter_block; out2[1] = g_iter_blockhot;
    if (reset) g_iter_block = g_iter_blockhot = 0;
}
#endif
