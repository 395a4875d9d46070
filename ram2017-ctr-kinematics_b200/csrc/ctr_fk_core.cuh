// ctr_fk_core.cuh — per-configuration forward kinematics of a concentric tube robot, written
// once as __host__ __device__ templates.  The CUDA kernels (ctr_fk_kernels.cu) run one of these
// solvers per GPU thread with all state in registers; tests/hostsim compiles the very same code
// for the CPU so that kernel logic can be checked where there is no GPU.  (hostsim is a test
// harness; the product library has no CPU path.)
//
// What is computed (reference: /root/reference/ctr_source):
//   robot_i.h:1123-1138 + section_i.h:201-213   joints -> per-section curved interval [Start,End]
//   robot_i.h:1079-1084 / :1093-1097            sample count N, step h, start position
//   robot_i.h:1101-1122, :1139-1238             tip->base explicit sweep of tube angles/torsion
//   robot_i.h:1241-1285                         SE(3) step, Rz, frame product
//   robot_i.h:705-720 / :734-754                which samples exist; the tip frame
//
// Two arithmetic variants:
//   STRICT  - every expression in the reference's operation order, libm-grade sin/cos/div/sqrt.
//   FAST    - algebraically equivalent: reciprocals of per-section constants hoisted, the
//             SE(3) step as power series in (h|u|/2)^2 (no sqrt, no division, no trig), rotation
//             accumulated as a unit quaternion, sin/cos of the tube angles carried and rotated by
//             the per-step increment, software-pipelined sweep.  Differences to STRICT are
//             O(1e-15) per step (measured end to end: DESIGN.md section 5).
// Also here: the tangent-linear sweep and Newton step of the base-angle shooting solve, the
// quaternion+translation helpers of the warp-parallel backbone pass, the radius fill.
// Both use the SAME control scalars (sum of Phi, N, h, the arc-position sequence and the
// section-interval compares), evaluated with non-contracted IEEE operations, because those decide
// which samples exist (SURVEY.md §7 "bit-critical scalars").
#pragma once

#include <math.h>
#include <stdint.h>

#include <type_traits>

#if defined(__CUDACC__)
#define CTR_HD __host__ __device__ __forceinline__
#else
#define CTR_HD inline
#endif

namespace ctrk {

constexpr int kMaxSec  = 3;
constexpr int kMaxTube = 6;

// Robot tables in the form the solvers consume.  Built on the host by make_krobot() from a
// ctr_robot_desc; passed to kernels by value (lives in the constant bank, uniform access).
struct KRobot {
    double ini[12];          // entry frame, column-major 3x4
    double ksum[kMaxSec];    // sum of tube stiffnesses of the section (left to right)
    double kappa[kMaxSec];   // curvature[0]
    double len[kMaxSec];
    double rad[kMaxSec];     // outer radius
    double share[kMaxSec];   // ksum / ntube                          (section_i.h:189)
    double onu[kMaxTube];    // 1 + nu of the tube as used by the torsion step (incl. the
                             // last-VC-section quirk, section_i.h:273)
    double czc[kMaxTube];    // share / onu                           (section_i.h:248)
    double kg[kMaxSec];      // kappa * share                         (FAST)
    double rk[kMaxSec];      // 1 / (ksum[s] + ... + ksum[S-1])       (FAST)
    double onu_last;         // 1 + nu of the innermost tube          (section_i.h:285)
    double cl;               // onu_last / ksum[S-1]                  (FAST)
    double nczc[kMaxTube];   // -czc[t] * cl: torque balance of the innermost tube in one chain (FAST)
    double kgr[kMaxSec][kMaxSec];   // kg[s] * rk[p], p <= s: curvature weight already divided by the
                                    // stiffness of the active sections p..S-1               (FAST)
    double qini[4];          // unit quaternion (w,x,y,z) of the entry frame's rotation
    double ztol;
    int    maxs;
    int    nsec;
    int    ntube[kMaxSec];
};

template <int S_, int A, int B, int C>
struct Lay {
    static constexpr int S = S_;
    static constexpr int T = A + B + C;
    static constexpr int NJ = 1 + S + T;
    CTR_HD static constexpr int nt(int s) { return s == 0 ? A : (s == 1 ? B : C); }
    CTR_HD static constexpr int t0(int s) { return s == 0 ? 0 : (s == 1 ? A : A + B); }
    CTR_HD static constexpr int sec_of(int t) { return t < A ? 0 : (t < A + B ? 1 : 2); }
};

// ---- exact control arithmetic (never contracted into FMA, never reassociated) ----------------
CTR_HD double x_add(double a, double b)
{
#ifdef __CUDA_ARCH__
    return __dadd_rn(a, b);
#else
    return a + b;
#endif
}
CTR_HD double x_sub(double a, double b)
{
#ifdef __CUDA_ARCH__
    return __dsub_rn(a, b);
#else
    return a - b;
#endif
}
CTR_HD double x_div(double a, double b)
{
#ifdef __CUDA_ARCH__
    return __ddiv_rn(a, b);
#else
    return a / b;
#endif
}
CTR_HD double x_fma(double a, double b, double c)
{
#ifdef __CUDA_ARCH__
    return __fma_rn(a, b, c);
#else
    return fma(a, b, c);
#endif
}

// An empty asm that "modifies" x: the value becomes opaque to the optimiser, which can then neither fold it nor
// rematerialise it later from its operands.
CTR_HD void keep_in_register(double& x)
{
#ifdef __CUDA_ARCH__
    asm volatile("" : "+d"(x));
#else
    (void)x;
#endif
}

// Control block of one configuration.
template <class L>
struct Ctl {
    double theta;
    double arc_end;
    double h;
    double end[L::S];
    double start[L::S];
    double al[L::T];     // tube angles at the tip (joint values)
    int    n;            // N: samples of the backward sweep
    int    nframes;      // m_Samples: frames the forward loop emits (N, or N-1 in step mode when
                         // the position-driven loop drops the tip sample, robot_i.h:705-708)
    int    status;
};

// robot_i.h:1123-1138, :1079-1084, :1093-1097, :705-708
template <class L>
CTR_HD void setup_control(const KRobot& rb, const double* __restrict__ jv, int mode, double step,
                          int nsamples, Ctl<L>& c)
{
    c.theta  = jv[0];
    c.status = 0;
    double run = 0.0;
#pragma unroll
    for (int s = 0; s < L::S; ++s) {
        const double phi = jv[1 + s + L::t0(s)];
        if (!((phi >= 0.0) && (phi <= rb.len[s]))) c.status = 1;
#pragma unroll
        for (int k = 0; k < L::nt(s); ++k) c.al[L::t0(s) + k] = jv[2 + s + L::t0(s) + k];
        run        = x_add(run, phi);
        c.end[s]   = run;
        const double d = x_sub(run, rb.len[s]);
        c.start[s] = (0.0 < d) ? d : 0.0;
    }
    c.arc_end = run;
    if (c.status) { c.n = 0; c.nframes = 0; c.h = step; return; }

    int    n;
    double h = step;
    if (mode == 0) {
        const double q = x_div(run, h);
        // (size_t)q of the reference; q >= 0 here.  Clamp before converting (q can be huge).
        n = (q >= (double)rb.maxs + 1.0) ? rb.maxs + 1 : (int)q;
        if (n < 1) n = 1;
        if (n > rb.maxs) { n = rb.maxs; h = x_div(run, (double)n); }
    } else {
        n = nsamples < rb.maxs ? nsamples : rb.maxs;
        h = x_div(run, (double)n);
    }
    c.n = n;
    c.h = h;
    c.nframes = n;
    if (mode == 0 && n > 1) {
        // Replay the arc-position sequence of the reference: N-1 subtractions down the sweep,
        // then N-1 additions up the forward loop; the tip sample exists iff the last position is
        // still <= ArcEnd.
        double p = run;
        for (int k = 1; k < n; ++k) p = x_sub(p, h);
        for (int k = 1; k < n; ++k) p = x_add(p, h);
        if (!(p <= run)) c.nframes = n - 1;
    }
}

// ---- small SE(3) helpers (column-major 3x4) ---------------------------------------------------
CTR_HD void tf_set_identity(double* a)
{
#pragma unroll
    for (int i = 0; i < 12; ++i) a[i] = 0.0;
    a[0] = a[4] = a[8] = 1.0;
}
// c = a*b ; c may alias b
CTR_HD void tf_mul(const double* a, const double* b, double* c)
{
    double r[12];
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
        for (int i = 0; i < 3; ++i)
            r[3 * j + i] = a[i] * b[3 * j] + a[3 + i] * b[3 * j + 1] + a[6 + i] * b[3 * j + 2];
#pragma unroll
    for (int i = 0; i < 3; ++i)
        r[9 + i] = (a[i] * b[9] + a[3 + i] * b[10] + a[6 + i] * b[11]) + a[9 + i];
#pragma unroll
    for (int i = 0; i < 12; ++i) c[i] = r[i];
}
// o = Init * Rz(sn,cn) * t   (robot_i.h:714-716, :1280-1285)
CTR_HD void tf_finish(const double* ini, const double* t, double sn, double cn, double* o)
{
    double r[12];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        r[3 * j]     = t[3 * j] * cn - t[3 * j + 1] * sn;
        r[3 * j + 1] = t[3 * j + 1] * cn + t[3 * j] * sn;
        r[3 * j + 2] = t[3 * j + 2];
    }
    tf_mul(ini, r, o);
}
// robot_i.h:1241-1269
CTR_HD void tf_step_strict(double ux, double uy, double uz, double h, double ztol, double* e)
{
    const double nrm = sqrt(ux * ux + uy * uy + uz * uz);
    if (!(fabs(nrm) < ztol)) {
        const double wx = ux / nrm, wy = uy / nrm, wz = uz / nrm;
        double sn, cn;
#ifdef __CUDA_ARCH__
        sincos(h * nrm, &sn, &cn);
#else
        sn = sin(h * nrm); cn = cos(h * nrm);
#endif
        const double vm  = cn - 1.0;
        const double r00 = (wy * wy + wz * wz) * vm + 1.0;
        const double r01 = -sn * wz - wx * wy * vm;
        const double r02 = sn * wy - wx * wz * vm;
        const double r10 = sn * wz - wx * wy * vm;
        const double r11 = (wx * wx + wz * wz) * vm + 1.0;
        const double r12 = -sn * wx - wy * wz * vm;
        const double r20 = -sn * wy - wx * wz * vm;
        const double r21 = sn * wx - wy * wz * vm;
        const double r22 = (wx * wx + wy * wy) * vm + 1.0;
        e[0] = r00; e[1] = r10; e[2] = r20;
        e[3] = r01; e[4] = r11; e[5] = r21;
        e[6] = r02; e[7] = r12; e[8] = r22;
        e[9]  = h * wx * wz + (-(wy * (r00 - 1.0)) + (r01 * wx)) / nrm;
        e[10] = h * wy * wz + ((wx * (r11 - 1.0)) - (r10 * wy)) / nrm;
        e[11] = h * wz * wz + ((r21 * wx) - (r20 * wy)) / nrm;
        return;
    }
    tf_set_identity(e);
    e[11] = h * nrm;
}

// ---- STRICT solver ---------------------------------------------------------------------------
// Outputs: tip[12] (nullable), albase[T] (nullable).  Returns nothing; c.nframes/c.h hold
// m_Samples and the returned step.
// OnSample: functor called as on_sample(k, ux, uy, uz, alpha_k) for every sample k in sweep order
// (k = N-1 .. 0); used by the backbone kernels.  Pass NoSample{} when not needed.
struct NoSample {
    CTR_HD void operator()(int, double, double, double, double) const {}
};

template <class L, bool WITH_TIP, class OnSample>
CTR_HD void solve_strict(const KRobot& rb, const Ctl<L>& c, double* tip, double* albase,
                         OnSample on_sample)
{
    constexpr int S = L::S, T = L::T;
    double al[T], uzs[T], uxs[T];
#pragma unroll
    for (int t = 0; t < T; ++t) { al[t] = c.al[t]; uzs[t] = 0.0; uxs[t] = 0.0; }
    double P[12];
    tf_set_identity(P);
    double a_tip = 0.0;
    const double h = c.h;
    double pos = c.arc_end;
    const int tipk = c.nframes - 1;

    for (int k = c.n - 1; k >= 0; --k) {
        double sn[T], cs[T];
#pragma unroll
        for (int t = 0; t < T; ++t) {
#ifdef __CUDA_ARCH__
            sincos(al[t], &sn[t], &cs[t]);
#else
            sn[t] = sin(al[t]); cs[t] = cos(al[t]);
#endif
        }
        double kap[S], kst[S], shr[S];
#pragma unroll
        for (int s = 0; s < S; ++s) {
            const bool in_k = (pos <= c.end[s]);
            const bool in_c = in_k && (pos >= c.start[s]);
            kap[s] = in_c ? rb.kappa[s] : 0.0;
            kst[s] = in_k ? rb.ksum[s] : 0.0;
            shr[s] = in_k ? rb.share[s] : 0.0;
        }
        double cx = 0.0, cy = 0.0, ktot = 0.0;
#pragma unroll
        for (int s = 0; s < S; ++s) {
            ktot += kst[s];
#pragma unroll
            for (int j = 0; j < L::nt(s); ++j) {
                const int t = L::t0(s) + j;
                cx += -sn[t] * kap[s] * shr[s];
                cy += cs[t] * kap[s] * shr[s];
            }
        }
        double uy_in = 0.0;
#pragma unroll
        for (int t = 0; t < T; ++t) {
            uxs[t] = (cs[t] * cx + sn[t] * cy) / ktot;
            if (t == T - 1) uy_in = (-sn[t] * cx + cs[t] * cy) / ktot;
        }
        const double uz_k = uzs[T - 1];      // z slot of sample k (0 at the tip)
        double a_k;
        if (k >= 1) {
            // tube angles with the previous sample's torsion (section_i.h:289-312)
            const double z0 = uzs[0];
            al[0] = 0.0;
#pragma unroll
            for (int t = 1; t < T; ++t) al[t] = al[t] + h * (uzs[t] - z0);
            a_k = al[T - 1];
            // torsion (section_i.h:243-286)
            double cz = 0.0;
            constexpr int sl = S - 1;
            if (L::nt(sl) == 2) {
                constexpr int t = L::t0(sl);
                cz += shr[sl] / rb.onu[t] * uzs[t];
                uzs[t] += h * rb.onu[t] * uxs[t] * kap[sl];
            }
#pragma unroll
            for (int s = S - 2; s >= 0; --s)
#pragma unroll
                for (int j = L::nt(s) - 1; j >= 0; --j) {
                    const int t = L::t0(s) + j;
                    cz += shr[s] / rb.onu[t] * uzs[t];
                    uzs[t] += h * rb.onu[t] * uxs[t] * kap[s];
                }
            uzs[T - 1] = (-cz) * rb.onu_last / kst[sl];
        } else {
            a_k = al[T - 1];
        }
        on_sample(k, uxs[T - 1], uy_in, uz_k, a_k);
        if (WITH_TIP) {
            if (k <= tipk) {
                double e[12];
                tf_step_strict(uxs[T - 1], uy_in, uz_k, h, rb.ztol, e);
                tf_mul(e, P, P);
            }
            if (k == tipk) a_tip = a_k;
        }
        if (k >= 1) pos = x_sub(pos, h);
    }
    if (WITH_TIP && tip) {
        double sn, cn;
        const double a = a_tip + c.theta;
#ifdef __CUDA_ARCH__
        sincos(a, &sn, &cn);
#else
        sn = sin(a); cn = cos(a);
#endif
        tf_finish(rb.ini, P, sn, cn, tip);
    }
    if (albase) {
#pragma unroll
        for (int t = 0; t < T; ++t) albase[t] = al[t];
    }
}

// ---- FAST building blocks ---------------------------------------------------------------------
//
// Throughput notes (B200, measured: profiles/): an FP64 instruction keeps the SM sub-partition's
// FP64 pipe busy for 2 cycles while the scheduler can issue 1 instruction per cycle, so the sweep
// is FP64-pipe bound only while it issues fewer non-FP64 than FP64 instructions.  Hence:
//   * every polynomial coefficient lives in __constant__ memory and is consumed as a c[bank][off]
//     operand of the DFMA itself (a literal double costs two extra MOVs per use);
//   * sign flips, quadrant swaps, interval masks and range guards are integer/select operations on
//     the bit patterns, which run on the otherwise idle ALU pipe;
//   * rarely taken branches (huge angles, coarse steps, zero curvature) are kept out of line.

struct FastConsts {
    double two_over_pi, magic, p1, p2, p3;
    double s[6];      // sin kernel, Horner order (highest degree first)
    double c[6];      // cos kernel
    double qa4[4], qb4[4], qc4[4];   // SE(3) series for m <= 2^-8, Horner order
    double qa8[7], qb8[7], qc8[7];   // SE(3) series for m <= 2^-4
    // run_scaled: 4 (|phi| - sin|phi|) / |phi|^3 - 2/3 as a series in m = (|phi|/2)^2, Horner order
    double qcs4[2];
};

#define CTRK_FAST_CONSTS_INIT                                                                      \
    {                                                                                              \
        6.36619772367581382433e-01, 6755399441055744.0, 1.57079632673412561417e+00,                \
            6.07710050650619224932e-11, 2.02226624879595063154e-21,                                \
            {1.58969099521155010221e-10, -2.50507602534068634195e-08, 2.75573137070700676789e-06,  \
             -1.98412698298579493134e-04, 8.33333333332248946124e-03, -1.66666666666666324348e-01},\
            {-1.13596475577881948265e-11, 2.08757232129817482790e-09, -2.75573143513906633035e-07, \
             2.48015872894767294178e-05, -1.38888888888741095749e-03, 4.16666666666666019037e-02}, \
            {1.0 / 40320.0, -1.0 / 720.0, 1.0 / 24.0, -0.5},                                       \
            {1.0 / 362880.0, -1.0 / 5040.0, 1.0 / 120.0, -1.0 / 6.0},                              \
            {256.0 / 39916800.0, -64.0 / 362880.0, 16.0 / 5040.0, -4.0 / 120.0},                   \
            {-1.0 / 87178291200.0, 1.0 / 479001600.0, -1.0 / 3628800.0, 1.0 / 40320.0,             \
             -1.0 / 720.0, 1.0 / 24.0, -0.5},                                                      \
            {-1.0 / 1307674368000.0, 1.0 / 6227020800.0, -1.0 / 39916800.0, 1.0 / 362880.0,        \
             -1.0 / 5040.0, 1.0 / 120.0, -1.0 / 6.0},                                              \
            {4096.0 / 1307674368000.0, -1024.0 / 6227020800.0, 256.0 / 39916800.0,                 \
             -64.0 / 362880.0, 16.0 / 5040.0, -4.0 / 120.0, 1.0 / 6.0},                            \
            {64.0 / 5040.0, -16.0 / 120.0}                                                         \
    }
// NB: qc8's last entry is the constant term 1/6 (its Horner chain has no separate "+1" step).

#if defined(__CUDACC__)
static __constant__ FastConsts kFastDev = CTRK_FAST_CONSTS_INIT;
#endif
static const FastConsts kFastHost = CTRK_FAST_CONSTS_INIT;

#ifdef __CUDA_ARCH__
#define CTRK_FC kFastDev
#else
#define CTRK_FC kFastHost
#endif

// bit-pattern helpers (integer pipe)
CTR_HD int hi_word(double x)
{
#ifdef __CUDA_ARCH__
    return __double2hiint(x);
#else
    union { double d; uint64_t u; } v; v.d = x;
    return (int)(uint32_t)(v.u >> 32);
#endif
}
CTR_HD int lo_word(double x)
{
#ifdef __CUDA_ARCH__
    return __double2loint(x);
#else
    union { double d; uint64_t u; } v; v.d = x;
    return (int)(uint32_t)v.u;
#endif
}
CTR_HD double xor_sign(double x, int mask_hi)      // flips the sign iff mask_hi == 0x80000000
{
#ifdef __CUDA_ARCH__
    return __hiloint2double(__double2hiint(x) ^ mask_hi, __double2loint(x));
#else
    union { double d; uint64_t u; } v; v.d = x;
    v.u ^= ((uint64_t)(uint32_t)mask_hi) << 32;
    return v.d;
#endif
}
// a <= b for doubles known to be >= +0 (or a negative a): signed compare of the bit patterns
CTR_HD bool le_nonneg(double a, double b)
{
#ifdef __CUDA_ARCH__
    return __double_as_longlong(a) <= __double_as_longlong(b);
#else
    union { double d; int64_t i; } x, y; x.d = a; y.d = b;
    return x.i <= y.i;
#endif
}
CTR_HD double libm_sincos_s(double x, double* c)
{
    double s;
#ifdef __CUDA_ARCH__
    sincos(x, &s, c);
#else
    s = sin(x); *c = cos(x);
#endif
    return s;
}

// sin and cos for |x| < 1024 (the caller guards the range): k = rint(x*2/pi) by the magic-number
// trick, r = x - k*pi/2 with a 2-term Cody-Waite split (k*p1 exact for |k| < 2^20; the dropped
// third term is < 1.4e-18 for |k| <= 652), fdlibm kernel polynomials on |r| <= pi/4, quadrant
// fix-up on the bit patterns.  Error < 1 ulp.  22 FP64 instructions.
CTR_HD void sincos_bounded(double x, double& s_out, double& c_out)
{
    const double kd_m = x_fma(x, CTRK_FC.two_over_pi, CTRK_FC.magic);
    const double kd   = kd_m - CTRK_FC.magic;
    const int    q    = lo_word(kd_m);
    double r = x_fma(-kd, CTRK_FC.p1, x);
    r        = x_fma(-kd, CTRK_FC.p2, r);
    const double z = r * r;
    double ps = x_fma(CTRK_FC.s[0], z, CTRK_FC.s[1]);
    double pc = x_fma(CTRK_FC.c[0], z, CTRK_FC.c[1]);
    ps = x_fma(ps, z, CTRK_FC.s[2]);
    pc = x_fma(pc, z, CTRK_FC.c[2]);
    ps = x_fma(ps, z, CTRK_FC.s[3]);
    pc = x_fma(pc, z, CTRK_FC.c[3]);
    ps = x_fma(ps, z, CTRK_FC.s[4]);
    pc = x_fma(pc, z, CTRK_FC.c[4]);
    ps = x_fma(ps, z, CTRK_FC.s[5]);
    pc = x_fma(pc, z, CTRK_FC.c[5]);
    const double sr = x_fma(r * z, ps, r);
    const double cr = x_fma(z * z, pc, x_fma(-0.5, z, 1.0));
    const bool   sw = (q & 1) != 0;
    const double s1 = sw ? cr : sr;
    const double c1 = sw ? sr : cr;
    s_out = xor_sign(s1, (q & 2) << 30);
    c_out = xor_sign(c1, ((q + 1) & 2) << 30);
}

// General-purpose version (any finite or non-finite x): used once per configuration for the
// final Rz and in the backbone pass.
CTR_HD void fast_sincos(double x, double& s_out, double& c_out)
{
    if ((hi_word(x) & 0x7fffffff) < 0x40900000) {      // |x| < 1024
        sincos_bounded(x, s_out, c_out);
        return;
    }
    s_out = libm_sincos_s(x, &c_out);
}

// Series of the SE(3) step in m = (h|u|/2)^2 (half-angle squared):
//   qa(m) = cos(sqrt m)                    = 1 - m/2 + m^2/24 - ...
//   qb(m) = sin(sqrt m)/sqrt m             = 1 - m/6 + m^2/120 - ...
//   qc(m) = (t - sin t)/t^3, t = 2 sqrt m  = 1/6 - 4m/120 + 16 m^2/5040 - ...
// Tier 1 (m <= 2^-8, i.e. h|u| <= 0.125 rad per step: every step of the benchmark workloads):
// through m^4, truncation < 3e-19.  Tier 2 (m <= 2^-4): through m^7, truncation < 1e-18.
constexpr double kFastM1 = 0.00390625;   // 2^-8 (tested on the bit pattern: hi word < 0x3f700000)
constexpr double kFastM2 = 0.0625;       // 2^-4

CTR_HD void se3_series4(double m, double& qa, double& qb, double& qc)
{
    double a = x_fma(CTRK_FC.qa4[0], m, CTRK_FC.qa4[1]);
    double b = x_fma(CTRK_FC.qb4[0], m, CTRK_FC.qb4[1]);
    double c = x_fma(CTRK_FC.qc4[0], m, CTRK_FC.qc4[1]);
    a = x_fma(a, m, CTRK_FC.qa4[2]);
    b = x_fma(b, m, CTRK_FC.qb4[2]);
    c = x_fma(c, m, CTRK_FC.qc4[2]);
    a = x_fma(a, m, CTRK_FC.qa4[3]);
    b = x_fma(b, m, CTRK_FC.qb4[3]);
    c = x_fma(c, m, CTRK_FC.qc4[3]);
    qa = x_fma(a, m, 1.0);
    qb = x_fma(b, m, 1.0);
    qc = x_fma(c, m, 1.0 / 6.0);
}

CTR_HD void se3_series8(double m, double& qa, double& qb, double& qc)
{
    double a = CTRK_FC.qa8[0], b = CTRK_FC.qb8[0], c = CTRK_FC.qc8[0];
#pragma unroll
    for (int i = 1; i < 7; ++i) {
        a = x_fma(a, m, CTRK_FC.qa8[i]);
        b = x_fma(b, m, CTRK_FC.qb8[i]);
        c = x_fma(c, m, CTRK_FC.qc8[i]);
    }
    qa = x_fma(a, m, 1.0);
    qb = x_fma(b, m, 1.0);
    qc = c;
}

// One SE(3) step as (unit quaternion, translation), FAST form.
//   rotation  exp(h [u]x)  ->  qw = cos(t/2), qv = sin(t/2)/|u| * u,  t = h|u|
//   translation (robot_i.h:1257-1259, exact arc integral):
//       pe = h e3 + B (u x e3) + C (u u_z - |u|^2 e3),  B = (1-cos t)/|u|^2, C = (t-sin t)/|u|^3
struct StepConsts {      // per-configuration constants of the step
    double h, hh, hh2, h3, ztol2;
};
CTR_HD StepConsts make_step_consts(double h, double ztol)
{
    StepConsts k;
    k.h = h; k.hh = 0.5 * h; k.hh2 = k.hh * k.hh; k.h3 = h * h * h; k.ztol2 = ztol * ztol;
    return k;
}

// Cold part: zero curvature (reference quirk), tier 2, closed form.
#if defined(__CUDACC__)
__host__ __device__ __noinline__
#endif
inline void se3_step_cold(double ux, double uy, double uz, double sxy, double n2, double m,
                          double h, double hh, double h3, double ztol2,
                          double* o /* qw qx qy qz px py pz */)
{
    // scalars by value on purpose: taking the address of the caller's state would force the
    // whole sweep state into local memory
    if (n2 < ztol2) {
        // |u| "equals" zero: the reference returns the identity rotation and a translation of
        // h*|u| (~0, not h) along z (robot_i.h:1245, :1264-1266).  Reproduced.
        o[0] = 1.0; o[1] = o[2] = o[3] = 0.0; o[4] = o[5] = 0.0; o[6] = h * sqrt(n2);
        return;
    }
    if (m <= kFastM2) {
        double qa, qb, qc;
        se3_series8(m, qa, qb, qc);
        const double b  = hh * qb;
        const double b2 = 2.0 * b;
        const double C  = h3 * qc;
        const double cz = C * uz;
        o[0] = qa; o[1] = b * ux; o[2] = b * uy; o[3] = b * uz;
        o[4] = x_fma(cz, ux, b2 * o[2]);
        o[5] = x_fma(cz, uy, -(b2 * o[1]));
        o[6] = x_fma(-C, sxy, h);
        return;
    }
    const double nrm = sqrt(n2);
    const double t = h * nrm;
    double ch, ct;
    const double sh = libm_sincos_s(0.5 * t, &ch);
    const double st = libm_sincos_s(t, &ct);
    const double b = sh / nrm;
    o[0] = ch; o[1] = b * ux; o[2] = b * uy; o[3] = b * uz;
    const double Bc = (1.0 - ct) / n2;
    const double C  = (t - st) / (n2 * nrm);
    o[4] = Bc * uy + C * ux * uz;
    o[5] = -Bc * ux + C * uy * uz;
    o[6] = h - C * sxy;
}

CTR_HD void se3_step_fast_k(double ux, double uy, double uz, const StepConsts& k, double& qw,
                            double& qx, double& qy, double& qz, double& px, double& py, double& pz)
{
    const double sxy = x_fma(uy, uy, ux * ux);
    const double n2  = x_fma(uz, uz, sxy);
    const double m   = n2 * k.hh2;
    // hot path: ztol^2 <= n2 and m <= 2^-8, both tested on the bit patterns (values are >= 0)
    if (le_nonneg(k.ztol2, n2) && hi_word(m) < 0x3f700000) {
        double qa, qb, qc;
        se3_series4(m, qa, qb, qc);
        const double b2 = k.h * qb;              // 2 sin(t/2)/|u|
        const double b  = k.hh * qb;             //   sin(t/2)/|u|
        qw = qa;
        qx = b * ux; qy = b * uy; qz = b * uz;
        const double C  = k.h3 * qc;
        const double cz = C * uz;
        px = x_fma(cz, ux, b2 * qy);             // B uy + C ux uz,  B = 2 b^2
        py = x_fma(cz, uy, -(b2 * qx));          // -B ux + C uy uz
        pz = x_fma(-C, sxy, k.h);                // h - C (ux^2+uy^2)
        return;
    }
    double o[7];
    se3_step_cold(ux, uy, uz, sxy, n2, m, k.h, k.hh, k.h3, k.ztol2, o);
    qw = o[0]; qx = o[1]; qy = o[2]; qz = o[3]; px = o[4]; py = o[5]; pz = o[6];
}

CTR_HD void se3_step_fast(double ux, double uy, double uz, double h, double ztol, double& qw,
                          double& qx, double& qy, double& qz, double& px, double& py, double& pz)
{
    const StepConsts k = make_step_consts(h, ztol);
    se3_step_fast_k(ux, uy, uz, k, qw, qx, qy, qz, px, py, pz);
}

// ---- FAST solver ---------------------------------------------------------------------------
// Same recurrence as solve_strict; see the header comment for what differs.
//
// Structure of the sweep (k runs from N-1 at the tip down to 0 at the base):
//   head  k = N-1   tube 0 still carries its joint angle, torsion is zero, and in step mode the
//                   frame of this sample may be the dropped tip sample
//   body  k = N-2..1  the hot loop
//   tail  k = 0     no angle/torsion update
// Two things make the body cheap:
//   * Software pipelining: the SE(3) accumulation of sample k+1 is issued in iteration k, next to
//     the (independent) curvature/torsion arithmetic of sample k, so the FP64 pipe always has a
//     second dependency chain to draw from (16 warps per SM are not enough to hide DFMA latency
//     on their own).  Iteration k leaves its sample in `pend` for iteration k-1.
//   * Incremental trigonometry: tube angles move by d_t = h (z_t - z_0) per step, |d_t| ~ 1e-2,
//     so sin/cos(alpha_t) are carried and rotated by (cos d_t, sin d_t) from a 4-term series
//     instead of being re-evaluated (15 instead of 22 FP64 instructions per tube and step).  The
//     rotation is accurate to 1 ulp per step; over N steps that adds O(N) ulp to quantities whose
//     own rounding history (alpha_t += d_t) is of the same size.
// A body iteration is HOT when every tube has |d_t| < 0.09375 (kRotMaxHi), and the pending sample has
// ztol <= |u| and (h|u|/2)^2 < 2^-8; all benchmark workloads are 100 % hot.  Anything else (coarse
// sampling, zero curvature) takes the generic iteration.
#ifdef CTRK_COUNT_COLD
static long long g_iter_total = 0, g_iter_cold = 0, g_iter_cold_delta = 0, g_iter_block = 0, g_iter_blockhot = 0;
#endif
// Carried sin/cos are rotated by the step increment d with the degree-8/9 Taylor pair while
// |d| < 0.09375 (high word of 1.5 * 2^-4): truncation d^10/10! < 1.5e-17, d^11/11! < 2e-19.
constexpr int kRotMaxHi = 0x3fb80000;

// C++14 stand-in for `if constexpr`: only functors used with EMIT need emit_hot/emit_generic
template <bool EMIT, bool HOT> struct EmitPending { template <class F> static CTR_HD void call(F&) {} };
template <> struct EmitPending<true, true>  { template <class F> static CTR_HD void call(F& f) { f.emit_hot(); } };
template <> struct EmitPending<true, false> { template <class F> static CTR_HD void call(F& f) { f.emit_generic(); } };

// HALF: the innermost tube's angle is carried as the sin/cos of its HALF (rotated by d/2 per step) and the full-angle
// pair the sweep needs follows by the double-angle formulas; the functor receives the half-angle pair of every sample
// through set_half() — what a frame emitter needs for the quaternion of Rz(alpha_k + theta) (PendingFrameEmitter).
struct NoHalf { template <class F> static CTR_HD void set(F&, double, double) {} };
struct PassHalf { template <class F> static CTR_HD void set(F& f, double s, double c) { f.set_half(s, c); } };

template <class L, bool WANT_TIP, class OnSample, bool HALF = false>
struct FastSweep {
    static constexpr int  S = L::S, T = L::T;

    using HalfOut = typename std::conditional<HALF, PassHalf, NoHalf>::type;
    static constexpr bool want_tip = WANT_TIP;   // compile-time: keeps the accumulation of the
                                                 // pending sample in the same basic block as the
                                                 // arithmetic of the current one
    const KRobot& rb;
    const Ctl<L>& c;
    OnSample      on_sample;
    StepConsts    kc;
    double al[T], uzs[T];
    double sn[T], cs[T];               // sin/cos(al[t]) for t >= 1 (tube 0: see step())
    double hko[T > 1 ? T - 1 : 1];     // h * (1+nu_t) * kappa_sec(t)
    double Qw, Qx, Qy, Qz, Px, Py, Pz; // suffix product E_k ... E_tip
    double pux, puy, puz;              // pending sample (not yet accumulated)
    double a_tip, pos;
    double hs, hc;                     // HALF: sin/cos(al[T-1] / 2)

    CTR_HD FastSweep(const KRobot& rb_, const Ctl<L>& c_, const OnSample& os)
        : rb(rb_), c(c_), on_sample(os)
    {
        kc = make_step_consts(c.h, rb.ztol);
#pragma unroll
        for (int t = 0; t < T; ++t) { al[t] = c.al[t]; uzs[t] = 0.0; sn[t] = 0.0; cs[t] = 1.0; }
#pragma unroll
        for (int t = 1; t < T; ++t) fast_sincos(al[t], sn[t], cs[t]);
        hs = 0.0; hc = 1.0;
        if (HALF) fast_sincos(0.5 * al[T - 1], hs, hc);
#pragma unroll
        for (int t = 0; t + 1 < T; ++t) hko[t] = c.h * rb.onu[t] * rb.kappa[L::sec_of(t)];
        Qw = 1.0; Qx = Qy = Qz = 0.0; Px = Py = Pz = 0.0;
        pux = puy = puz = 0.0;
        a_tip = 0.0;
        pos = c.arc_end;
    }
    // (Q,P) <- E (Q,P) for one step given as unit quaternion e and translation pe
    static CTR_HD void compose(double ew, double ex, double ey, double ez, double pex, double pey, double pez,
                               double& Qw, double& Qx, double& Qy, double& Qz, double& Px, double& Py, double& Pz)
    {
        // translation: p <- pe + R(e) p,  R(e) v = v + 2 (w t + q x t),  t = q x v
        const double tx = x_fma(ey, Pz, -(ez * Py));
        const double ty = x_fma(ez, Px, -(ex * Pz));
        const double tz = x_fma(ex, Py, -(ey * Px));
        const double ix = x_fma(ew, tx, x_fma(ey, tz, -(ez * ty)));
        const double iy = x_fma(ew, ty, x_fma(ez, tx, -(ex * tz)));
        const double iz = x_fma(ew, tz, x_fma(ex, ty, -(ey * tx)));
        Px = x_fma(2.0, ix, Px + pex);
        Py = x_fma(2.0, iy, Py + pey);
        Pz = x_fma(2.0, iz, Pz + pez);
        // rotation: Q <- e (x) Q
        const double w2 = x_fma(-ez, Qz, x_fma(-ey, Qy, x_fma(-ex, Qx, ew * Qw)));
        const double x2 = x_fma(-ez, Qy, x_fma(ey, Qz, x_fma(ex, Qw, ew * Qx)));
        const double y2 = x_fma(-ex, Qz, x_fma(ez, Qx, x_fma(ey, Qw, ew * Qy)));
        const double z2 = x_fma(-ey, Qx, x_fma(ex, Qy, x_fma(ez, Qw, ew * Qz)));
        Qw = w2; Qx = x2; Qy = y2; Qz = z2;
    }
    CTR_HD void apply(double ew, double ex, double ey, double ez, double pex, double pey, double pez)
    {
        compose(ew, ex, ey, ez, pex, pey, pez, Qw, Qx, Qy, Qz, Px, Py, Pz);
    }
    CTR_HD void accumulate_generic(double ux, double uy, double uz)
    {
        double ew, ex, ey, ez, px_, py_, pz_;
        se3_step_fast_k(ux, uy, uz, kc, ew, ex, ey, ez, px_, py_, pz_);
        apply(ew, ex, ey, ez, px_, py_, pz_);
    }
    // hot form: tier-1 series, no branches; sxy and m precomputed by the caller
    CTR_HD void accumulate_hot(double ux, double uy, double uz, double sxy, double m)
    {
        // rotation series to full precision; the third-order translation term C = h^3 qc only
        // needs ~1e-9 relative (it contributes ~1e-8 m per step), so qc stops after m^2
        double qa = x_fma(CTRK_FC.qa4[0], m, CTRK_FC.qa4[1]);
        double qb = x_fma(CTRK_FC.qb4[0], m, CTRK_FC.qb4[1]);
        double qc = x_fma(CTRK_FC.qc4[2], m, CTRK_FC.qc4[3]);
        qa = x_fma(qa, m, CTRK_FC.qa4[2]);
        qb = x_fma(qb, m, CTRK_FC.qb4[2]);
        qc = x_fma(qc, m, 1.0 / 6.0);
        qa = x_fma(qa, m, CTRK_FC.qa4[3]);
        qb = x_fma(qb, m, CTRK_FC.qb4[3]);
        qa = x_fma(qa, m, 1.0);
        qb = x_fma(qb, m, 1.0);
        const double b2 = kc.h * qb;
        const double b  = kc.hh * qb;
        const double ex = b * ux, ey = b * uy, ez = b * uz;
        const double C  = kc.h3 * qc;
        const double cz = C * uz;
        apply(qa, ex, ey, ez, x_fma(cz, ux, b2 * ey), x_fma(cz, uy, -(b2 * ex)), x_fma(-C, sxy, kc.h));
    }

    // One sample.  HEAD/TAIL as described above.  HOT: pipelined hot form; `dl[t]` = h (z_t - z_0)
    // (angle increments, valid when !TAIL), (sxy, m) describe the pending sample (valid when HOT).
    // ROT (with HOT): rotate the carried sin/cos by d_t; otherwise re-evaluate them from the new
    // angles with sincos_bounded (the caller has checked |alpha_t| < 1024).
    // EMIT (run_emit): the on_sample functor keeps the previous sample pending and turns it into
    // a frame here, next to this sample's arithmetic, exactly like (A) does for the tip.
    template <bool HEAD, bool TAIL, bool HOT, bool ROT = false, bool EMIT = false>
    CTR_HD void step(int k, const double* dl, double sxy_p, double m_p)
    {
        // ---- (A) accumulate the pending sample k+1 (independent of everything below) -----------
        if (!HEAD && want_tip) {
            if (HOT) accumulate_hot(pux, puy, puz, sxy_p, m_p);
            else     accumulate_generic(pux, puy, puz);
        }
        EmitPending<EMIT && !HEAD, HOT>::call(on_sample);
        // ---- (B) this sample ---------------------------------------------------------------------
        double s0 = 0.0, c0 = 1.0;                      // tube 0 is pinned to 0 after the head step
        if (HEAD) fast_sincos(al[0], s0, c0);
        // section masks on the exact arc position (integer compares of the bit patterns)
        bool in_k[S], in_c[S];
#pragma unroll
        for (int s = 0; s < S; ++s) {
            in_k[s] = (s == S - 1) ? true : le_nonneg(pos, c.end[s]);   // pos <= ArcEnd = End_{S-1}
            in_c[s] = in_k[s] && le_nonneg(c.start[s], pos);
        }
        // Curvature weights kappa_s K_s/n_s, already divided by the stiffness of the sections that
        // reach this position.  End is non-decreasing in s, so the active set is a suffix
        // {p..S-1}; section s is only ever weighted when it is active itself, i.e. p <= s.
        double cx = 0.0, cy = 0.0;
        bool first = true;
#pragma unroll
        for (int s = 0; s < S; ++s) {
            double gr = rb.kgr[s][s];
#pragma unroll
            for (int p2 = s - 1; p2 >= 0; --p2) gr = in_k[p2] ? rb.kgr[s][p2] : gr;
            const double g = in_c[s] ? gr : 0.0;
#pragma unroll
            for (int j = 0; j < L::nt(s); ++j) {
                const int t = L::t0(s) + j;
                const double st = (t == 0) ? s0 : sn[t];
                const double ct = (t == 0) ? c0 : cs[t];
                if (!HEAD && t == 0) { cy = g; first = false; continue; }   // sin 0 = 0, cos 0 = 1
                if (first) { cx = -st * g; cy = ct * g; first = false; }
                else       { cx = x_fma(-st, g, cx); cy = x_fma(ct, g, cy); }
            }
        }
        double uxs[T];
#pragma unroll
        for (int t = 0; t < T; ++t) {
            if (t == 0) uxs[t] = HEAD ? x_fma(c0, cx, s0 * cy) : cx;
            else        uxs[t] = x_fma(cs[t], cx, sn[t] * cy);
        }
        double uy_in;
        if (T == 1) uy_in = HEAD ? x_fma(c0, cy, -(s0 * cx)) : cy;
        else        uy_in = x_fma(cs[T - 1], cy, -(sn[T - 1] * cx));
        const double uz_k = uzs[T - 1];

        double a_k = al[T - 1];
        if (!TAIL) {
            // tube angles with the previous sample's torsion (section_i.h:289-312)
            al[0] = 0.0;
#pragma unroll
            for (int t = 1; t < T; ++t) al[t] = al[t] + dl[t];
            a_k = al[T - 1];
            // torsion (section_i.h:243-286).  A section that does not reach this position yet
            // still has u_z = 0 (its update is masked by in_c, and in_c implies in_k), so the
            // in_k mask on the balance coefficients is redundant and dropped.
            if (T > 1) {
                double zl = 0.0;          // innermost tube: -(sum czc_t z_t) * cl, coefficients pre-multiplied
#pragma unroll
                for (int t = T - 2; t >= 0; --t) {
                    const int s = L::sec_of(t);
                    const double hk = in_c[s] ? hko[t] : 0.0;
                    zl     = (t == T - 2) ? rb.nczc[t] * uzs[t] : x_fma(rb.nczc[t], uzs[t], zl);
                    uzs[t] = x_fma(hk, uxs[t], uzs[t]);
                }
                uzs[T - 1] = zl;
            } else {
                uzs[0] = 0.0;      // -0 * cl, section_i.h:285 with an empty balance
            }
            // sin/cos of the new angles, for the next sample
#pragma unroll
            for (int t = 1; t < T; ++t) {
                if (HALF && t == T - 1 && HOT && ROT) {
                    // half-angle pair rotated by d/2, full-angle pair by the double-angle formulas (3 instead of 12
                    // instructions for the pair the sweep needs, and no separate sin/cos of the frame's half angle)
                    const double e  = 0.5 * dl[t];
                    const double e2 = e * e;
                    double pc = x_fma(CTRK_FC.qa4[0], e2, CTRK_FC.qa4[1]);
                    double ps = x_fma(CTRK_FC.qb4[0], e2, CTRK_FC.qb4[1]);
                    pc = x_fma(pc, e2, CTRK_FC.qa4[2]);
                    ps = x_fma(ps, e2, CTRK_FC.qb4[2]);
                    pc = x_fma(pc, e2, CTRK_FC.qa4[3]);
                    ps = x_fma(ps, e2, CTRK_FC.qb4[3]);
                    const double cd = x_fma(pc, e2, 1.0);
                    const double sd = x_fma(e * e2, ps, e);
                    const double s1 = x_fma(hc, sd, hs * cd);
                    const double c1 = x_fma(-hs, sd, hc * cd);
                    hs = s1; hc = c1;
                    sn[t] = 2.0 * (s1 * c1);
                    cs[t] = x_fma(c1, c1, -(s1 * s1));
                } else if (HOT && ROT) {
                    const double d  = dl[t];
                    const double d2 = d * d;
                    double pc = x_fma(CTRK_FC.qa4[0], d2, CTRK_FC.qa4[1]);
                    double ps = x_fma(CTRK_FC.qb4[0], d2, CTRK_FC.qb4[1]);
                    pc = x_fma(pc, d2, CTRK_FC.qa4[2]);
                    ps = x_fma(ps, d2, CTRK_FC.qb4[2]);
                    pc = x_fma(pc, d2, CTRK_FC.qa4[3]);
                    ps = x_fma(ps, d2, CTRK_FC.qb4[3]);
                    const double cd = x_fma(pc, d2, 1.0);          // cos d
                    const double sd = x_fma(d * d2, ps, d);        // sin d
                    const double s1 = x_fma(cs[t], sd, sn[t] * cd);
                    const double c1 = x_fma(-sn[t], sd, cs[t] * cd);
                    sn[t] = s1; cs[t] = c1;
                } else if (HOT) {
                    sincos_bounded(al[t], sn[t], cs[t]);
                    if (HALF && t == T - 1) sincos_bounded(0.5 * al[t], hs, hc);
                } else {
                    fast_sincos(al[t], sn[t], cs[t]);
                    if (HALF && t == T - 1) fast_sincos(0.5 * al[t], hs, hc);
                }
            }
            if (HALF && T == 1) { hs = 0.0; hc = 1.0; }      // tube 0 is pinned to 0 after the head step
        }
        on_sample(k, uxs[T - 1], uy_in, uz_k, a_k);
        if (HALF) HalfOut::set(on_sample, hs, hc);
        if (want_tip) {
            if (TAIL) {
                accumulate_generic(uxs[T - 1], uy_in, uz_k);
            } else if (HEAD && k > c.nframes - 1) {
                pux = puy = puz = 0.0;                  // dropped tip sample: pending = identity step
            } else {
                pux = uxs[T - 1]; puy = uy_in; puz = uz_k;
            }
            if (k == c.nframes - 1) a_tip = a_k;
        }
        if (!TAIL) pos = x_sub(pos, c.h);
    }

    // top-of-iteration quantities shared by the tier tests and the hot steps; returns the tier (0: hot + rotate,
    // 1: hot + sincos_bounded, 2: generic) — what run() computes inline, for callers that step two sweeps in lock-step
    CTR_HD int classify(double* dl, double& sxy, double& m) const
    {
        int dmax = 0, amax = 0;
        const double z0 = uzs[0];
#pragma unroll
        for (int t = 1; t < T; ++t) {
            dl[t] = c.h * (uzs[t] - z0);
            const int d = hi_word(dl[t]) & 0x7fffffff;
            const int a = hi_word(al[t]) & 0x7fffffff;
            dmax = d > dmax ? d : dmax;
            amax = a > amax ? a : amax;
        }
        sxy = x_fma(puy, puy, pux * pux);
        const double n2 = x_fma(puz, puz, sxy);
        m = n2 * kc.hh2;
        const bool se3_hot = !want_tip || (le_nonneg(kc.ztol2, n2) && hi_word(m) < 0x3f700000);
        if (se3_hot && dmax < kRotMaxHi) return 0;
        if (se3_hot && amax < 0x408f0000 && dmax < 0x40300000) return 1;
        return 2;
    }
    CTR_HD void step_tier(int tier, int k, const double* dl, double sxy, double m)
    {
        if (tier == 0)      step<false, false, true, true>(k, dl, sxy, m);
        else if (tier == 1) step<false, false, true, false>(k, dl, sxy, m);
        else                step<false, false, false>(k, dl, sxy, m);
    }

    // ---- run_scaled(): the sweep of run() with the state in units of the step ------------------------------------
    // A good part of run()'s hot iteration is not arithmetic of the recurrence but the factors h, h/2, (h/2)^2, h^3
    // that turn (u, z) into the step's half rotation vector psi = (h/2) u.  Here the loop keeps the state in those
    // units: w_t = h z_t, curvature weights (h/2) kgr (so x_t, y come out as components of psi), translation P / h.
    // Then d_t = w_t - w_0, the series run in m = |psi|^2 with the tables of run(), the step's quaternion is
    // (qa, qb psi) and its translation needs no h: 8 FP64 instructions fewer per step (113 -> 105 on the 3-tube
    // robot), and the tier test needs no |alpha| reduction.  An iteration that is not hot (rare: coarse sampling,
    // zero curvature) stays in these units too (scaled_step<.., false>); everything only that path needs (1/h, the
    // step constants) is recomputed there, so it does not occupy registers in the hot loop.  The sweep itself needs
    // only sin/cos of the tube angles — they are always advanced by rotating the carried pair — so the angles
    // themselves are accumulated only with KEEP_AL (when they are an output), and the pose does not depend on that.
    // Same arithmetic per step as run() up to the placement of the factor h (FAST: results agree to rounding).
    static constexpr int NG = S * (S + 1) / 2;
#ifndef CTRK_SCALED_UNROLL
#define CTRK_SCALED_UNROLL 2
#endif
    // two iterations per trip: the copies that hand the loop-carried state from one iteration to the next and half
    // of the loop control disappear (tools/tune_tip.cu: 1.89 -> 1.85 ms per 1M on the 3-tube robot, 2.86 -> 2.72 on
    // the 5-tube one)
    static constexpr int kScaledUnroll = CTRK_SCALED_UNROLL;
    // run_emit: two iterations per trip where the kernel has the registers for it (layouts that run at 2 CTAs per SM,
    // backbone_blocks_per_sm): 6.46 -> 6.38 ms per 1M x 256 frames; at 168 registers it is slower (7.39 against 6.95)
#ifdef CTRK_EMIT_UNROLL
    static constexpr int kEmitUnroll = CTRK_EMIT_UNROLL;
#else
    static constexpr int kEmitUnroll = (T >= 4 || S >= 3) ? 2 : 1;
#endif

    // one iteration; dl[t] = w_t - w_0, (sxy, m) = |psi_xy|^2, |psi|^2 of the pending sample (pux, puy, puz = psi).
    // HOT: both series apply (compile time; the flags are ignored).  Otherwise (rare: coarse sampling, zero curvature)
    // se3_ok / rot_ok say which of them still does: the step of a sample outside the series' range comes from the
    // generic step routine (handed u = 2 psi / h, its translation taken in units of h), and an angle increment
    // outside the rotation series' range is applied with sin/cos of the increment itself — so no iteration ever
    // needs the tube angles, and the state never leaves the units of the step.
    template <bool KEEP_AL, bool HOT>
    CTR_HD void scaled_step(const double* hg, const double* hko2, double* w, const double* dl, double sxy, double m,
                            bool se3_ok, bool rot_ok)
    {
        // ---- (A) accumulate the pending sample -------------------------------------------------------------------
        if (want_tip) {
            if (HOT || se3_ok) {
                double qa = x_fma(CTRK_FC.qa4[0], m, CTRK_FC.qa4[1]);
                double qb = x_fma(CTRK_FC.qb4[0], m, CTRK_FC.qb4[1]);
                double qc = x_fma(CTRK_FC.qcs4[0], m, CTRK_FC.qcs4[1]);
                qa = x_fma(qa, m, CTRK_FC.qa4[2]);
                qb = x_fma(qb, m, CTRK_FC.qb4[2]);
                qc = x_fma(qc, m, 2.0 / 3.0);                 // 4 (|phi| - sin|phi|) / |phi|^3
                qa = x_fma(qa, m, CTRK_FC.qa4[3]);
                qb = x_fma(qb, m, CTRK_FC.qb4[3]);
                qa = x_fma(qa, m, 1.0);                       // cos(|phi|/2)
                qb = x_fma(qb, m, 1.0);                       // sin(|phi|/2) / (|phi|/2)
                const double ex = qb * pux, ey = qb * puy, ez = qb * puz;
                const double cz = qc * puz;
                apply(qa, ex, ey, ez, x_fma(cz, pux, qb * ey), x_fma(cz, puy, -(qb * ex)), x_fma(-qc, sxy, 1.0));
            } else {
                const double inv_h = 1.0 / c.h, inv_hh = 2.0 * inv_h;
                const StepConsts kg = make_step_consts(c.h, rb.ztol);
                double ew, ex, ey, ez, px_, py_, pz_;
                se3_step_fast_k(pux * inv_hh, puy * inv_hh, puz * inv_hh, kg, ew, ex, ey, ez, px_, py_, pz_);
                apply(ew, ex, ey, ez, px_ * inv_h, py_ * inv_h, pz_ * inv_h);
            }
        }
        // ---- (B) this sample ------------------------------------------------------------------------------------
        bool in_k[S], in_c[S];
#pragma unroll
        for (int s = 0; s < S; ++s) {
            in_k[s] = (s == S - 1) ? true : le_nonneg(pos, c.end[s]);
            in_c[s] = in_k[s] && le_nonneg(c.start[s], pos);
        }
        double cx = 0.0, cy = 0.0;
        bool first = true;
#pragma unroll
        for (int s = 0; s < S; ++s) {
            double gr = hg[s * (s + 1) / 2 + s];
#pragma unroll
            for (int p2 = s - 1; p2 >= 0; --p2) gr = in_k[p2] ? hg[s * (s + 1) / 2 + p2] : gr;
            const double g = in_c[s] ? gr : 0.0;
#pragma unroll
            for (int j = 0; j < L::nt(s); ++j) {
                const int t = L::t0(s) + j;
                if (t == 0) { cy = g; first = false; continue; }            // tube 0: sin 0 = 0, cos 0 = 1
                if (first) { cx = -sn[t] * g; cy = cs[t] * g; first = false; }
                else       { cx = x_fma(-sn[t], g, cx); cy = x_fma(cs[t], g, cy); }
            }
        }
        double uxs[T];
#pragma unroll
        for (int t = 0; t < T; ++t) uxs[t] = (t == 0) ? cx : x_fma(cs[t], cx, sn[t] * cy);
        const double uy_in = (T == 1) ? cy : x_fma(cs[T - 1], cy, -(sn[T - 1] * cx));
        const double pz_k = 0.5 * w[T - 1];
        if (KEEP_AL) {
#pragma unroll
            for (int t = 1; t < T; ++t) al[t] = al[t] + dl[t];
        }
        if (T > 1) {
            double zl = 0.0;
#pragma unroll
            for (int t = T - 2; t >= 0; --t) {
                const double hk = in_c[L::sec_of(t)] ? hko2[t] : 0.0;
                zl   = (t == T - 2) ? rb.nczc[t] * w[t] : x_fma(rb.nczc[t], w[t], zl);
                w[t] = x_fma(hk, uxs[t], w[t]);
            }
            w[T - 1] = zl;
        } else {
            w[0] = 0.0;
        }
#pragma unroll
        for (int t = 1; t < T; ++t) {
            const double d = dl[t];
            double cd, sd;
            if (HOT || rot_ok) {
                const double d2 = d * d;
                double pc = x_fma(CTRK_FC.qa4[0], d2, CTRK_FC.qa4[1]);
                double ps = x_fma(CTRK_FC.qb4[0], d2, CTRK_FC.qb4[1]);
                pc = x_fma(pc, d2, CTRK_FC.qa4[2]);
                ps = x_fma(ps, d2, CTRK_FC.qb4[2]);
                pc = x_fma(pc, d2, CTRK_FC.qa4[3]);
                ps = x_fma(ps, d2, CTRK_FC.qb4[3]);
                cd = x_fma(pc, d2, 1.0);
                sd = x_fma(d * d2, ps, d);
            } else if ((hi_word(d) & 0x7fffffff) < 0x408f0000) {
                sincos_bounded(d, sd, cd);                    // |d| < 992
            } else {
                fast_sincos(d, sd, cd);
            }
            const double s1 = x_fma(cs[t], sd, sn[t] * cd);
            const double c1 = x_fma(-sn[t], sd, cs[t] * cd);
            sn[t] = s1; cs[t] = c1;
        }
        pux = uxs[T - 1]; puy = uy_in; puz = pz_k;
        pos = x_sub(pos, c.h);
    }

    // FINE: the form for finely sampled sweeps (what the tip and gather kernels run when the launch's step is small,
    // fine_sampling() below).  Its iterations outside the hot tier convert the state back, run run()'s own iteration
    // and convert again — costly, but such iterations do not occur there, and the hot loop comes out 1-4 % faster than
    // with the cold step in units of the step next to it (3 tubes 1.843 against 1.858 ms per 1M configurations, 3
    // tubes in 3 sections 1.940 against 2.017, 5 tubes 2.729 against 2.750: equal instruction counts, a matter of
    // how ptxas allocates and schedules the loop next to a large or a small cold block).  Coarse sampling is the other way round
    // (N = 64: 1.21 against 0.82 ms), hence two forms.  Hot iterations are the same arithmetic in both.
    template <bool KEEP_AL, bool FINE = false>
    CTR_HD void run_scaled(bool generic = false)
    {
        const int n = c.n;
        // short sweeps; a retracted robot (h = 0) has no 1 / h; callers that want run() for coarse launches without a
        // second inlined copy of it (the backbone kernel)
        if (n < 8 || !(c.h > 1e-150) || generic) { run(); return; }
        double dl[T];
#pragma unroll
        for (int t = 0; t < T; ++t) dl[t] = 0.0;
        step<true, false, false>(n - 1, dl, 0.0, 0.0);
        {   // the iteration that can hold the tip sample of a step-mode sweep (k == nframes - 1) stays generic
            double sxy, m;
            const int tier = classify(dl, sxy, m);
            step_tier(tier, n - 2, dl, sxy, m);
        }
        const double h = c.h, hh = 0.5 * c.h;
        double hg[NG], hko2[T > 1 ? T - 1 : 1];
#pragma unroll
        for (int s2 = 0; s2 < S; ++s2)
#pragma unroll
            for (int q = 0; q <= s2; ++q) hg[s2 * (s2 + 1) / 2 + q] = hh * rb.kgr[s2][q];
#pragma unroll
        for (int t = 0; t + 1 < T; ++t) hko2[t] = 2.0 * hko[t];
        double zt2s = (rb.ztol * rb.ztol) * (hh * hh);
        // loop invariants the compiler would otherwise recompute in every iteration (it prefers 13 FP64 instructions
        // per step to 7 live register pairs): made opaque, so they stay in registers
#pragma unroll
        for (int i = 0; i < NG; ++i) keep_in_register(hg[i]);
#pragma unroll
        for (int t = 0; t + 1 < T; ++t) keep_in_register(hko2[t]);
        keep_in_register(zt2s);
        // into units of the step
        double w[T];
        {
            const double inv_h = 1.0 / h;
#pragma unroll
            for (int t = 0; t < T; ++t) w[t] = h * uzs[t];
            pux *= hh; puy *= hh; puz *= hh;
            Px *= inv_h; Py *= inv_h; Pz *= inv_h;
        }
        // |psi|^2 = m is hot when zt2s < m < 2^-8; on the high words: hi(zt2s) < hi(m) < hi(2^-8), one unsigned range
        // compare (a tie of the high words goes to the generic step, which compares exactly)
        const unsigned m_lo = (unsigned)hi_word(zt2s) + 1u, m_span = 0x3f700000u - m_lo;
#pragma unroll kScaledUnroll
        for (int k = n - 3; k >= 1; --k) {
            int dmax = 0;
#pragma unroll
            for (int t = 1; t < T; ++t) {
                dl[t] = w[t] - w[0];
                const int d = hi_word(dl[t]) & 0x7fffffff;
                dmax = d > dmax ? d : dmax;
            }
            const double sxy = x_fma(puy, puy, pux * pux);
            const double m   = x_fma(puz, puz, sxy);
            // hot: every |d_t| < 0.09375, zero tolerance <= |u|, (h|u|/2)^2 < 2^-8
            const bool rot_ok = dmax < kRotMaxHi;
            const bool se3_ok = !want_tip || ((unsigned)hi_word(m) - m_lo) < m_span;
#ifdef CTRK_COUNT_COLD   /* host-side analysis builds only (tests/hostsim) */
            ++g_iter_block; if (rot_ok && se3_ok) ++g_iter_blockhot;
#endif
            if (rot_ok && se3_ok) {
                scaled_step<KEEP_AL, true>(hg, hko2, w, dl, sxy, m, true, true);
            } else if (!FINE) {
                scaled_step<KEEP_AL, false>(hg, hko2, w, dl, sxy, m, se3_ok, rot_ok);
            } else {
                // back to (u, z, P), one iteration of run(), and into units of the step again.  The generic
                // iteration evaluates sin/cos of the tube angles; it gets them modulo 2 pi from the carried pair in
                // BOTH forms of this loop, so the pose does not depend on whether the angles are an output too.
                const double inv_h = 1.0 / h, inv_hh = 2.0 * inv_h;
                kc = make_step_consts(h, rb.ztol);
#pragma unroll
                for (int t = 0; t + 1 < T; ++t) hko[t] = 0.5 * hko2[t];
#pragma unroll
                for (int t = 0; t < T; ++t) uzs[t] = w[t] * inv_h;
                pux *= inv_hh; puy *= inv_hh; puz *= inv_hh;
                Px *= h; Py *= h; Pz *= h;
                double al_out[T];
#pragma unroll
                for (int t = 1; t < T; ++t) { al_out[t] = al[t]; al[t] = atan2(sn[t], cs[t]); }
                double sxy_u, m_u;
                const int tier = classify(dl, sxy_u, m_u);
                step_tier(tier, k, dl, sxy_u, m_u);
                if (KEEP_AL) {
#pragma unroll
                    for (int t = 1; t < T; ++t) al[t] = al_out[t] + dl[t];
                }
#pragma unroll
                for (int t = 0; t < T; ++t) w[t] = h * uzs[t];
                pux *= hh; puy *= hh; puz *= hh;
                Px *= inv_h; Py *= inv_h; Pz *= inv_h;
            }
        }
        {   // back to (u, z, P) for the tail
            const double inv_h = 1.0 / h, inv_hh = 2.0 * inv_h;
            kc = make_step_consts(h, rb.ztol);
#pragma unroll
            for (int t = 0; t + 1 < T; ++t) hko[t] = 0.5 * hko2[t];
#pragma unroll
            for (int t = 0; t < T; ++t) uzs[t] = w[t] * inv_h;
            pux *= inv_hh; puy *= inv_hh; puz *= inv_hh;
            Px *= h; Py *= h; Pz *= h;
        }
        {
            const double z0 = uzs[0];
#pragma unroll
            for (int t = 1; t < T; ++t) dl[t] = c.h * (uzs[t] - z0);    // unused by the tail
        }
        step<false, true, false>(0, dl, 0.0, 0.0);
    }

    // Full sweep for a functor that emits frames one sample late (PendingFrameEmitter; thread-per-
    // configuration backbone kernel): as run(), with the frame of sample k+1 built inside
    // iteration k.  An iteration is hot when both the sweep and the pending frame qualify for
    // their branch-free forms.
    CTR_HD void run_emit()
    {
        static_assert(!WANT_TIP, "the frame product is the functor's");
        const int n = c.n;
        double dl[T];
#pragma unroll
        for (int t = 0; t < T; ++t) dl[t] = 0.0;
        if (n == 1) {
            step<true, true, false>(0, dl, 0.0, 0.0);
        } else {
            step<true, false, false>(n - 1, dl, 0.0, 0.0);
#pragma unroll kEmitUnroll
            for (int k = n - 2; k >= 1; --k) {
                int dmax = 0, amax = 0;
                const double z0 = uzs[0];
#pragma unroll
                for (int t = 1; t < T; ++t) {
                    dl[t] = c.h * (uzs[t] - z0);
                    const int d = hi_word(dl[t]) & 0x7fffffff;
                    const int a = hi_word(al[t]) & 0x7fffffff;
                    dmax = d > dmax ? d : dmax;
                    amax = a > amax ? a : amax;
                }
                const bool eh = on_sample.pending_hot();
                if (eh && dmax < kRotMaxHi)                            step<false, false, true, true, true>(k, dl, 0.0, 0.0);
                else if (eh && amax < 0x408f0000 && dmax < 0x40300000) step<false, false, true, false, true>(k, dl, 0.0, 0.0);
                else                                                   step<false, false, false, false, true>(k, dl, 0.0, 0.0);
            }
            {
                const double z0 = uzs[0];
#pragma unroll
                for (int t = 1; t < T; ++t) dl[t] = c.h * (uzs[t] - z0);    // unused by the tail
            }
            step<false, true, false, false, true>(0, dl, 0.0, 0.0);
        }
        on_sample.emit_generic();          // sample 0
    }

    CTR_HD void run()
    {
        const int n = c.n;
        double dl[T];
#pragma unroll
        for (int t = 0; t < T; ++t) dl[t] = 0.0;         // torsion is zero at the tip
        if (n == 1) {
            step<true, true, false>(0, dl, 0.0, 0.0);
            return;
        }
        step<true, false, false>(n - 1, dl, 0.0, 0.0);
#pragma unroll 1
        for (int k = n - 2; k >= 1; --k) {
            // top-of-iteration quantities shared by the hot tests and the hot steps
            int dmax = 0, amax = 0;
            const double z0 = uzs[0];
#pragma unroll
            for (int t = 1; t < T; ++t) {
                dl[t] = c.h * (uzs[t] - z0);
                const int d = hi_word(dl[t]) & 0x7fffffff;
                const int a = hi_word(al[t]) & 0x7fffffff;
                dmax = d > dmax ? d : dmax;
                amax = a > amax ? a : amax;
            }
            const double sxy = x_fma(puy, puy, pux * pux);
            const double n2  = x_fma(puz, puz, sxy);
            const double m   = n2 * kc.hh2;
            const bool se3_hot = !want_tip || (le_nonneg(kc.ztol2, n2) && hi_word(m) < 0x3f700000);
            const bool rot_ok  = dmax < kRotMaxHi;           // every |d_t| < 0.09375
            const bool sc_ok   = amax < 0x408f0000;          // every |alpha_t| < 992 (+ |d_t| stays < 1024)
#ifdef CTRK_COUNT_COLD   /* host-side analysis builds only (tests/hostsim) */
            ++g_iter_total; if (!(se3_hot && (rot_ok || sc_ok))) ++g_iter_cold;
            if (!rot_ok) ++g_iter_cold_delta;
#endif
            if (se3_hot && rot_ok)                 step<false, false, true, true>(k, dl, sxy, m);
            else if (se3_hot && sc_ok && dmax < 0x40300000) step<false, false, true, false>(k, dl, sxy, m);
            else                                   step<false, false, false>(k, dl, sxy, m);
        }
        {
            const double z0 = uzs[0];
#pragma unroll
            for (int t = 1; t < T; ++t) dl[t] = c.h * (uzs[t] - z0);    // unused by the tail
        }
        step<false, true, false>(0, dl, 0.0, 0.0);
    }
};

// WANT_TIP = false skips the frame product (sample-only sweeps: STRICT backbone kernel, stability).
// quaternion -> rotation (normalised: |Q| drifts by ~N ulp), then Init * Rz * P
template <class L, class Sweep>
CTR_HD void finish_tip(const KRobot& rb, const Ctl<L>& c, const Sweep& sw, double* tip)
{
    const double Qw = sw.Qw, Qx = sw.Qx, Qy = sw.Qy, Qz = sw.Qz;
    const double nn = Qw * Qw + Qx * Qx + Qy * Qy + Qz * Qz;
    const double sc = 2.0 / nn;
    double Pm[12];
    Pm[0] = 1.0 - sc * (Qy * Qy + Qz * Qz);
    Pm[1] = sc * (Qx * Qy + Qw * Qz);
    Pm[2] = sc * (Qx * Qz - Qw * Qy);
    Pm[3] = sc * (Qx * Qy - Qw * Qz);
    Pm[4] = 1.0 - sc * (Qx * Qx + Qz * Qz);
    Pm[5] = sc * (Qy * Qz + Qw * Qx);
    Pm[6] = sc * (Qx * Qz + Qw * Qy);
    Pm[7] = sc * (Qy * Qz - Qw * Qx);
    Pm[8] = 1.0 - sc * (Qx * Qx + Qy * Qy);
    Pm[9] = sw.Px; Pm[10] = sw.Py; Pm[11] = sw.Pz;
    double sa, ca;
    fast_sincos(sw.a_tip + c.theta, sa, ca);
    tf_finish(rb.ini, Pm, sa, ca, tip);
}

// WANT_TIP = false skips the frame product (sample-only sweeps: STRICT backbone kernel, stability).
template <class L, bool WANT_TIP, class OnSample, bool FINE = false>
CTR_HD void solve_fast(const KRobot& rb, const Ctl<L>& c, double* tip, double* albase,
                       OnSample on_sample)
{
    constexpr int T = L::T;
    FastSweep<L, WANT_TIP, OnSample> sw(rb, c, on_sample);
    if (std::is_same<OnSample, NoSample>::value) {
        // nobody looks at the samples: the loop in units of the step; the tube angles themselves only if they are an output
        if (albase) sw.template run_scaled<true, FINE>();
        else        sw.template run_scaled<false, FINE>();
    } else {
        sw.run();
    }
    if (WANT_TIP && tip) finish_tip<L>(rb, c, sw, tip);
    if (albase) {
#pragma unroll
        for (int t = 0; t < T; ++t) albase[t] = sw.al[t];
    }
}

// Two configurations with the SAME sample count solved by one thread in lock-step: two independent dependency chains
// for the compiler to interleave, so a warp always has FP64 work ready while the other chain waits on a result.
// Iterations in which both are hot run the two hot steps in one basic block; anything else steps them one after the
// other.  Results are bit-identical to two solve_fast calls (same arithmetic per configuration).
template <class L>
CTR_HD void solve_fast_pair(const KRobot& rb, const Ctl<L>& c0, const Ctl<L>& c1, double* tip0, double* tip1)
{
    constexpr int T = L::T;
    FastSweep<L, true, NoSample> A(rb, c0, NoSample{}), B(rb, c1, NoSample{});
    const int n = c0.n;                                    // == c1.n
    double dlA[T], dlB[T];
#pragma unroll
    for (int t = 0; t < T; ++t) { dlA[t] = 0.0; dlB[t] = 0.0; }
    if (n == 1) {
        A.template step<true, true, false>(0, dlA, 0.0, 0.0);
        B.template step<true, true, false>(0, dlB, 0.0, 0.0);
    } else {
        A.template step<true, false, false>(n - 1, dlA, 0.0, 0.0);
        B.template step<true, false, false>(n - 1, dlB, 0.0, 0.0);
#pragma unroll 1
        for (int k = n - 2; k >= 1; --k) {
            double sxyA, mA, sxyB, mB;
            const int tA = A.classify(dlA, sxyA, mA);
            const int tB = B.classify(dlB, sxyB, mB);
            if ((tA | tB) == 0) {
                A.template step<false, false, true, true>(k, dlA, sxyA, mA);
                B.template step<false, false, true, true>(k, dlB, sxyB, mB);
            } else {
                A.step_tier(tA, k, dlA, sxyA, mA);
                B.step_tier(tB, k, dlB, sxyB, mB);
            }
        }
        {
            double s_, m_;
            A.classify(dlA, s_, m_); B.classify(dlB, s_, m_);       // dl only; unused by the tail
        }
        A.template step<false, true, false>(0, dlA, 0.0, 0.0);
        B.template step<false, true, false>(0, dlB, 0.0, 0.0);
    }
    finish_tip<L>(rb, c0, A, tip0);
    finish_tip<L>(rb, c1, B, tip1);
}


// ---- tangent-linear sweep: base angles and their Jacobian w.r.t. the tip angles -----------------
// The reference's operator maps tip angles to base angles (m_SampleTubeAlpha after the sweep,
// robot_i.h:810 reads it as such).  A robot actuated at the base needs the inverse: find the tip
// angles whose sweep ends at given base angles (the "shooting solve" of the
// north star).  This routine runs the same discrete sweep as solve_fast and carries, next to the
// state (alpha_t, z_t), its exact derivative with respect to the tip angles of tubes 1..T-1 (the
// variational equations of the discrete recurrence), so Newton's method on the base-angle
// residual converges quadratically to a root of the *reference's own* discretisation.
//   base[t]      = alpha_t at the base, t = 0..T-1
//   jac[r*(T-1)+c] = d base[r+1] / d alpha_tip[c+1]
template <class L>
CTR_HD void sweep_tangent(const KRobot& rb, const Ctl<L>& c, double* base, double* jac)
{
    constexpr int S = L::S, T = L::T, U = (T > 1 ? T - 1 : 1);
    double al[T], uz[T], dal[T][U], duz[T][U];
#pragma unroll
    for (int t = 0; t < T; ++t) {
        al[t] = c.al[t]; uz[t] = 0.0;
#pragma unroll
        for (int j = 0; j < U; ++j) { dal[t][j] = (t == j + 1) ? 1.0 : 0.0; duz[t][j] = 0.0; }
    }
    double hko[U];
#pragma unroll
    for (int t = 0; t + 1 < T; ++t) hko[t] = c.h * rb.onu[t] * rb.kappa[L::sec_of(t)];
    double pos = c.arc_end;
    // sin/cos of the tube angles are carried and rotated by the step increment like in FastSweep
    // (re-evaluated when an increment is large); tube 0 is pinned to 0 after the first step
    double sn[T], cs[T];
#pragma unroll
    for (int t = 0; t < T; ++t) fast_sincos(al[t], sn[t], cs[t]);
#pragma unroll 1
    for (int k = c.n - 1; k >= 1; --k) {          // the last sample (k = 0) does not update the state
        bool in_k[S], in_c[S];
#pragma unroll
        for (int s = 0; s < S; ++s) {
            in_k[s] = (s == S - 1) ? true : le_nonneg(pos, c.end[s]);
            in_c[s] = in_k[s] && le_nonneg(c.start[s], pos);
        }
        double rks = rb.rk[S - 1];
#pragma unroll
        for (int s = S - 2; s >= 0; --s) rks = in_k[s] ? rb.rk[s] : rks;
        double cx = 0.0, cy = 0.0, dcx[U], dcy[U];
#pragma unroll
        for (int j = 0; j < U; ++j) { dcx[j] = 0.0; dcy[j] = 0.0; }
#pragma unroll
        for (int s = 0; s < S; ++s) {
            const double g = in_c[s] ? rb.kg[s] : 0.0;
#pragma unroll
            for (int jt = 0; jt < L::nt(s); ++jt) {
                const int t = L::t0(s) + jt;
                cx = x_fma(-sn[t], g, cx);
                cy = x_fma(cs[t], g, cy);
                if (t >= 1) {
#pragma unroll
                    for (int j = 0; j < U; ++j) {
                        dcx[j] = x_fma(-cs[t] * g, dal[t][j], dcx[j]);
                        dcy[j] = x_fma(-sn[t] * g, dal[t][j], dcy[j]);
                    }
                }
            }
        }
        cx *= rks; cy *= rks;
#pragma unroll
        for (int j = 0; j < U; ++j) { dcx[j] *= rks; dcy[j] *= rks; }
        double ux[T], dux[T][U];
#pragma unroll
        for (int t = 0; t < T; ++t) {
            ux[t] = x_fma(cs[t], cx, sn[t] * cy);
            const double uy = x_fma(cs[t], cy, -(sn[t] * cx));      // d ux / d alpha_t
#pragma unroll
            for (int j = 0; j < U; ++j)
                dux[t][j] = x_fma(uy, (t >= 1 ? dal[t][j] : 0.0), x_fma(cs[t], dcx[j], sn[t] * dcy[j]));
        }
        // angles with the previous sample's torsion
        const double z0 = uz[0];
        al[0] = 0.0; sn[0] = 0.0; cs[0] = 1.0;
#pragma unroll
        for (int t = 1; t < T; ++t) {
            const double d = c.h * (uz[t] - z0);
            al[t] = al[t] + d;
            if ((hi_word(d) & 0x7fffffff) < kRotMaxHi) {           // rotate (as FastSweep::step)
                const double d2 = d * d;
                double pc = x_fma(CTRK_FC.qa4[0], d2, CTRK_FC.qa4[1]);
                double ps = x_fma(CTRK_FC.qb4[0], d2, CTRK_FC.qb4[1]);
                pc = x_fma(pc, d2, CTRK_FC.qa4[2]);
                ps = x_fma(ps, d2, CTRK_FC.qb4[2]);
                pc = x_fma(pc, d2, CTRK_FC.qa4[3]);
                ps = x_fma(ps, d2, CTRK_FC.qb4[3]);
                const double cd = x_fma(pc, d2, 1.0);
                const double sd = x_fma(d * d2, ps, d);
                const double s1 = x_fma(cs[t], sd, sn[t] * cd);
                const double c1 = x_fma(-sn[t], sd, cs[t] * cd);
                sn[t] = s1; cs[t] = c1;
            } else {
                fast_sincos(al[t], sn[t], cs[t]);
            }
#pragma unroll
            for (int j = 0; j < U; ++j) dal[t][j] = x_fma(c.h, duz[t][j] - duz[0][j], dal[t][j]);
        }
        // torsion
        if (T > 1) {
            double cz = 0.0, dcz[U];
#pragma unroll
            for (int j = 0; j < U; ++j) dcz[j] = 0.0;
#pragma unroll
            for (int t = T - 2; t >= 0; --t) {
                const double hk = in_c[L::sec_of(t)] ? hko[t] : 0.0;
                cz = x_fma(rb.czc[t], uz[t], cz);
                uz[t] = x_fma(hk, ux[t], uz[t]);
#pragma unroll
                for (int j = 0; j < U; ++j) {
                    dcz[j]    = x_fma(rb.czc[t], duz[t][j], dcz[j]);
                    duz[t][j] = x_fma(hk, dux[t][j], duz[t][j]);
                }
            }
            uz[T - 1] = -cz * rb.cl;
#pragma unroll
            for (int j = 0; j < U; ++j) duz[T - 1][j] = -dcz[j] * rb.cl;
        }
        pos = x_sub(pos, c.h);
    }
#pragma unroll
    for (int t = 0; t < T; ++t) base[t] = al[t];
    if (jac) {
#pragma unroll
        for (int r = 0; r + 1 < T; ++r)
#pragma unroll
            for (int j = 0; j + 1 < T; ++j) jac[r * (T - 1) + j] = dal[r + 1][j];
    }
}

// Solve the U x U system  J d = f  in place (Gaussian elimination with partial pivoting; U <= 5).
// Returns false when a pivot vanishes (singular Jacobian: a bifurcation point of the robot).
template <int U>
CTR_HD bool solve_small(double* J, double* f)
{
#pragma unroll
    for (int col = 0; col < U; ++col) {
        int piv = col;
        double best = fabs(J[col * U + col]);
#pragma unroll
        for (int r = col + 1; r < U; ++r) {
            const double v = fabs(J[r * U + col]);
            if (v > best) { best = v; piv = r; }
        }
        if (!(best > 1e-300)) return false;
#pragma unroll
        for (int r = col + 1; r < U; ++r) {
            if (r == piv) {
#pragma unroll
                for (int cc = 0; cc < U; ++cc) { const double tmp = J[col * U + cc]; J[col * U + cc] = J[r * U + cc]; J[r * U + cc] = tmp; }
                const double tf = f[col]; f[col] = f[r]; f[r] = tf;
            }
        }
        const double inv = 1.0 / J[col * U + col];
#pragma unroll
        for (int r = col + 1; r < U; ++r) {
            const double m = J[r * U + col] * inv;
#pragma unroll
            for (int cc = col; cc < U; ++cc) J[r * U + cc] = x_fma(-m, J[col * U + cc], J[r * U + cc]);
            f[r] = x_fma(-m, f[col], f[r]);
        }
    }
#pragma unroll
    for (int r = U - 1; r >= 0; --r) {
        double acc = f[r];
#pragma unroll
        for (int cc = r + 1; cc < U; ++cc) acc = x_fma(-J[r * U + cc], f[cc], acc);
        f[r] = acc / J[r * U + r];
    }
    return true;
}

// ---- the same operator on the continuum model, integrated by classical RK4 -----------------------------------
// The reference's sweep is explicit Euler (tip -> base, step h) on the torsional rod equations:
//     d alpha_t / d sigma = z_t - z_0                    (t >= 1; alpha_0 = 0: all angles are relative to tube 0)
//     d z_t / d sigma     = (1 + nu_t) kappa_sec(t) x_t   (t <= T-2, inside the curved interval of its section)
//     z_{T-1}             = -cl sum_{t < T-1} czc_t z_t   (torque balance)
//     x_t = cos alpha_t cx + sin alpha_t cy,  (cx, cy) = sum_t' g_sec(t') (-sin alpha_t', cos alpha_t') / K_active
// with sigma the arc length from the tip.  CTR_FK_FLAG_RK4 integrates this system with the fixed-step fourth-order
// Runge-Kutta scheme over the same N-1 steps (split where the right-hand side jumps, see sweep_tangent_rk4) and carries
// the tangent-linear system (the variational equations, d state / d tip angles) through every stage, so Newton's method
// again sees the exact derivative of the discrete map.
// It is a DIFFERENT discretisation from the reference's: its base angles differ from the reference sweep's by O(h)
// (tests/test_kernel_logic_cpu.py::test_rk4_sweep_converges_to_the_reference_sweep); there is no parity claim for it,
// its checker is oracle/ctr_shoot_rk4_py.py.  Tube 0's own joint angle (which the reference lets act on the tip-most
// sample only, section_i.h:308) is taken as 0 throughout.
template <class L>
struct Rk4Rhs {
    static constexpr int S = L::S, T = L::T, U = (T > 1 ? T - 1 : 1);
    // y = (alpha_1..alpha_{T-1}, z_0..z_{T-2}); Y[i][j] = d y_i / d alpha_tip[j+1]
    static CTR_HD void eval(const KRobot& rb, const Ctl<L>& c, double pos, const double* y, const double (*Y)[U],
                            double* dy, double (*dY)[U])
    {
        bool in_k[S], in_c[S];
        const double p = pos > 0.0 ? pos : 0.0;
#pragma unroll
        for (int s = 0; s < S; ++s) {
            in_k[s] = (s == S - 1) ? true : (p <= c.end[s]);
            in_c[s] = in_k[s] && (c.start[s] <= p);
        }
        double rks = rb.rk[S - 1];
#pragma unroll
        for (int s = S - 2; s >= 0; --s) rks = in_k[s] ? rb.rk[s] : rks;
        double sn[T], cs[T];
        sn[0] = 0.0; cs[0] = 1.0;
#pragma unroll
        for (int t = 1; t < T; ++t) fast_sincos(y[t - 1], sn[t], cs[t]);
        double cx = 0.0, cy = 0.0, dcx[U], dcy[U];
#pragma unroll
        for (int j = 0; j < U; ++j) { dcx[j] = 0.0; dcy[j] = 0.0; }
#pragma unroll
        for (int t = 0; t < T; ++t) {
            const double g = in_c[L::sec_of(t)] ? rb.kg[L::sec_of(t)] * rks : 0.0;
            cx = x_fma(-sn[t], g, cx);
            cy = x_fma(cs[t], g, cy);
            if (t >= 1) {
#pragma unroll
                for (int j = 0; j < U; ++j) {
                    dcx[j] = x_fma(-cs[t] * g, Y[t - 1][j], dcx[j]);
                    dcy[j] = x_fma(-sn[t] * g, Y[t - 1][j], dcy[j]);
                }
            }
        }
        // z of the innermost tube from the balance
        double zi = 0.0, dzi[U];
#pragma unroll
        for (int j = 0; j < U; ++j) dzi[j] = 0.0;
        if (T > 1) {
#pragma unroll
            for (int t = 0; t + 1 < T; ++t) {
                zi = x_fma(rb.czc[t], y[U + t], zi);
#pragma unroll
                for (int j = 0; j < U; ++j) dzi[j] = x_fma(rb.czc[t], Y[U + t][j], dzi[j]);
            }
            zi = -zi * rb.cl;
#pragma unroll
            for (int j = 0; j < U; ++j) dzi[j] = -dzi[j] * rb.cl;
        }
        const double z0 = y[U];
#pragma unroll
        for (int t = 1; t < T; ++t) {
            const double zt = (t == T - 1) ? zi : y[U + t];
            dy[t - 1] = zt - z0;
#pragma unroll
            for (int j = 0; j < U; ++j) dY[t - 1][j] = ((t == T - 1) ? dzi[j] : Y[U + t][j]) - Y[U][j];
        }
#pragma unroll
        for (int t = 0; t + 1 < T; ++t) {
            const double w  = in_c[L::sec_of(t)] ? rb.onu[t] * rb.kappa[L::sec_of(t)] : 0.0;
            const double ux = x_fma(cs[t], cx, sn[t] * cy);
            const double uy = x_fma(cs[t], cy, -(sn[t] * cx));          // d ux / d alpha_t
            dy[U + t] = w * ux;
#pragma unroll
            for (int j = 0; j < U; ++j)
                dY[U + t][j] = w * x_fma(uy, (t >= 1 ? Y[t - 1][j] : 0.0), x_fma(cs[t], dcx[j], sn[t] * dcy[j]));
        }
    }
};

// One classical RK4 step of length hs (towards the base) with the interval masks frozen at arc position pm.
template <class L>
CTR_HD void rk4_substep(const KRobot& rb, const Ctl<L>& c, double pm, double hs, double* y,
                        double (*Y)[(L::T > 1 ? L::T - 1 : 1)])
{
    constexpr int T = L::T, U = (T > 1 ? T - 1 : 1), M = 2 * U;
    double ya[M], Ya[M][U], yt[M], Yt[M][U], dy[M], dY[M][U];
    // stage 1
    Rk4Rhs<L>::eval(rb, c, pm, y, Y, dy, dY);
#pragma unroll
    for (int i = 0; i < M; ++i) {
        ya[i] = dy[i]; yt[i] = x_fma(0.5 * hs, dy[i], y[i]);
#pragma unroll
        for (int j = 0; j < U; ++j) { Ya[i][j] = dY[i][j]; Yt[i][j] = x_fma(0.5 * hs, dY[i][j], Y[i][j]); }
    }
    // stages 2 and 3 (midpoint)
#pragma unroll 1
    for (int stage = 2; stage <= 3; ++stage) {
        const double hn = (stage == 2) ? 0.5 * hs : hs;
        Rk4Rhs<L>::eval(rb, c, pm, yt, Yt, dy, dY);
#pragma unroll
        for (int i = 0; i < M; ++i) {
            ya[i] = x_fma(2.0, dy[i], ya[i]); yt[i] = x_fma(hn, dy[i], y[i]);
#pragma unroll
            for (int j = 0; j < U; ++j) { Ya[i][j] = x_fma(2.0, dY[i][j], Ya[i][j]); Yt[i][j] = x_fma(hn, dY[i][j], Y[i][j]); }
        }
    }
    // stage 4
    Rk4Rhs<L>::eval(rb, c, pm, yt, Yt, dy, dY);
    const double w = hs * (1.0 / 6.0);
#pragma unroll
    for (int i = 0; i < M; ++i) {
        y[i] = x_fma(w, ya[i] + dy[i], y[i]);
#pragma unroll
        for (int j = 0; j < U; ++j) Y[i][j] = x_fma(w, Ya[i][j] + dY[i][j], Y[i][j]);
    }
}

// The right-hand side jumps where a section's curved interval starts or ends (<= 2S arc positions, off the sample
// grid in general).  A Runge-Kutta step across a jump is only first-order accurate, which would leave the whole sweep
// at the accuracy of the reference's Euler sweep; so a step that contains such a position is split there, each part
// integrated with the masks of its own side (taken at its midpoint).  The sweep is then fourth-order in h.
template <class L>
CTR_HD void sweep_tangent_rk4(const KRobot& rb, const Ctl<L>& c, double* base, double* jac)
{
    constexpr int S = L::S, T = L::T, U = (T > 1 ? T - 1 : 1), M = 2 * U;
    double y[M], Y[M][U];
#pragma unroll
    for (int i = 0; i < M; ++i) {
        y[i] = (i < U && i + 1 < T) ? c.al[i + 1] : 0.0;
#pragma unroll
        for (int j = 0; j < U; ++j) Y[i][j] = (i == j && i + 1 < T) ? 1.0 : 0.0;
    }
    if (T > 1) {
        double pos = c.arc_end;
#pragma unroll 1
        for (int k = c.n - 1; k >= 1; --k) {
            const double stop = x_sub(pos, c.h);
            double cur = pos;
#pragma unroll 1
            while (cur > stop) {
                double nb = stop;                       // the nearest switching position strictly inside (stop, cur)
#pragma unroll
                for (int s = 0; s < S; ++s) {
                    if (c.end[s] < cur && c.end[s] > nb) nb = c.end[s];
                    if (c.start[s] < cur && c.start[s] > nb) nb = c.start[s];
                }
                const double hs = cur - nb;
                rk4_substep<L>(rb, c, cur - 0.5 * hs, hs, y, Y);
                cur = nb;
            }
            pos = stop;
        }
    }
    base[0] = 0.0;
#pragma unroll
    for (int t = 1; t < T; ++t) base[t] = y[t - 1];
    if (jac) {
#pragma unroll
        for (int r = 0; r + 1 < T; ++r)
#pragma unroll
            for (int j = 0; j + 1 < T; ++j) jac[r * (T - 1) + j] = Y[r][j];
    }
}

// Per-configuration state of the base-angle shooting solve: damped Newton on  base(x) - beta = 0  for the tip angles x
// of tubes 1..T-1.  round() runs ONE sweep (with its tangent) at the current trial point — the unit of lock-step work
// of the persistent kernel — and returns kRunning or a final CTR_FK_SHOOT_* status:
//   * the trial point x_acc - lam d is accepted when it lowers the residual max-norm (Armijo: below
//     (1 - 1e-4 lam) res_acc); then the Newton direction d = J^-1 f is taken from its own Jacobian (step limiter 1 rad:
//     tube angles are periodic, a longer step is an overshoot) and lam returns to 1;
//   * a rejected trial halves lam (backtracking line search); more than max_reject rejections in a row mean the
//     residual has stopped decreasing (a fold of the base-angle map: no root on this branch);
//   * with a guess from the caller (warm start / branch tracking) that ends the solve: CTR_FK_SHOOT_STALLED after a
//     handful of sweeps — a lane holding such an input gives its slot back instead of running to the sweep cap, and the
//     solve never leaves the branch the caller named;
//   * without a guess the first start is x = beta, and a start that stalls (or meets a singular Jacobian) is followed by
//     another from x = beta + u, u uniform in [-pi, pi)^(T-1) (counter-based, keyed by beta: reproducible wherever the
//     configuration is solved), until a root is found or max_iter sweeps are spent.  These robots have several roots
//     for many inputs, so WHICH root comes back is then arbitrary; measured on the shipped robots (3000 random inputs
//     with a root, N = 128, cap 100): 99.8 / 100 / 100 / 99.3 % solved at a mean of 6.6 / 5.7 / 6.5 / 10.6 sweeps
//     (plain Newton from x = beta: 96 / 97 / 96 / 88 %);
//   * the alpha slots of q return the ACCEPTED estimate with the lowest residual over all starts, whatever the status.
template <class L, bool RK4>
struct ShootState {
    static constexpr int T = L::T, U = (T > 1 ? T - 1 : 1), NJ = L::NJ;
    static constexpr int kRunning = -1, kMaxRejectGuess = 5, kMaxRejectFree = 1;
    double q[NJ], beta[T], d[U], best[U];
    double res, lam, best_res;
    int    it, nrej, max_reject;
    bool   fresh, restarts;

    // jv_base: joint vector with the WANTED base angles in its alpha slots; guess (nullable): joint vector whose alpha
    // slots hold the starting tip angles (warm start: the previous configuration's solution); default x = beta
    CTR_HD void start(const double* jv_base, const double* guess)
    {
#pragma unroll
        for (int j = 0; j < NJ; ++j) q[j] = jv_base[j];
#pragma unroll
        for (int t = 0; t < T; ++t) {
            const int j = 2 + L::sec_of(t) + t;
            beta[t] = jv_base[j];
            if (guess && t >= 1) q[j] = guess[j];
        }
#pragma unroll
        for (int r = 0; r < U; ++r) { d[r] = 0.0; best[r] = q[2 + L::sec_of(r + 1 < T ? r + 1 : 0) + (r + 1 < T ? r + 1 : 0)]; }
        res = 0.0; lam = 1.0; it = 0; nrej = 0; fresh = true;
        best_res = -1.0;                                          // nothing accepted yet
        restarts = (guess == nullptr);                            // a caller who names the branch stays on it
        max_reject = guess ? kMaxRejectGuess : kMaxRejectFree;
    }
    // this start has failed (stall or singular Jacobian): remember it if it is the best so far, then either begin
    // again from another point or finish with `st`
    CTR_HD int fail(int st, int max_iter)
    {
        keep_best();
        if (!restarts || it >= max_iter) { restore_best(); return st; }
        // x = beta + u, u uniform in [-pi, pi)^U from a counter-based generator keyed by the wanted angles, so the
        // result does not depend on which lane or device solves the configuration
        unsigned long long key = 0x9e3779b97f4a7c15ull * (unsigned long long)(it + 1);
#pragma unroll
        for (int t = 1; t < T; ++t) key = shoot_mix(key ^ shoot_bits(beta[t]));
#pragma unroll
        for (int t = 1; t < T; ++t) {
            key = shoot_mix(key + 0x9e3779b97f4a7c15ull);
            const double u = (double)(key >> 11) * (1.0 / 9007199254740992.0);       // [0, 1)
            q[2 + L::sec_of(t) + t] = beta[t] + (u - 0.5) * 6.283185307179586;
        }
        lam = 1.0; nrej = 0; fresh = true;
        return kRunning;
    }
    CTR_HD void keep_best()
    {
        if (!fresh && (best_res < 0.0 || res < best_res)) {
            best_res = res;
#pragma unroll
            for (int r = 0; r + 1 < T; ++r) best[r] = q[2 + L::sec_of(r + 1) + (r + 1)];
        }
    }
    CTR_HD void restore_best()
    {
        if (best_res >= 0.0) {
            res = best_res;
#pragma unroll
            for (int r = 0; r + 1 < T; ++r) q[2 + L::sec_of(r + 1) + (r + 1)] = best[r];
        }
    }
    static CTR_HD unsigned long long shoot_mix(unsigned long long z)              // splitmix64 finaliser
    {
        z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
        z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
        return z ^ (z >> 31);
    }
    static CTR_HD unsigned long long shoot_bits(double x)
    {
#ifdef __CUDA_ARCH__
        return (unsigned long long)__double_as_longlong(x);
#else
        union { double d; unsigned long long u; } v; v.d = x; return v.u;
#endif
    }
    CTR_HD int round(const KRobot& rb, int mode, double step, int nsamples, double tol, int max_iter)
    {
        double qt[NJ];
#pragma unroll
        for (int j = 0; j < NJ; ++j) qt[j] = q[j];
        if (!fresh) {
#pragma unroll
            for (int r = 0; r + 1 < T; ++r) {
                const int j = 2 + L::sec_of(r + 1) + (r + 1);
                qt[j] = q[j] - lam * d[r];
            }
        }
        Ctl<L> c;
        setup_control<L>(rb, qt, mode, step, nsamples, c);
        if (c.status) return 1;                                   // CTR_FK_SHOOT_PHI_RANGE
        double base[T], J[U * U], f[U];
        if (RK4) sweep_tangent_rk4<L>(rb, c, base, T > 1 ? J : nullptr);
        else     sweep_tangent<L>(rb, c, base, T > 1 ? J : nullptr);
        ++it;
        if (T == 1) { res = 0.0; return 0; }
        double r_new = 0.0;
#pragma unroll
        for (int r = 0; r + 1 < T; ++r) { f[r] = base[r + 1] - beta[r + 1]; r_new = fmax(r_new, fabs(f[r])); }
        const bool accept = fresh || (r_new < res * (1.0 - 1e-4 * lam));      // false for a NaN residual
        if (accept) {
#pragma unroll
            for (int j = 0; j < NJ; ++j) q[j] = qt[j];
            res = r_new; nrej = 0; fresh = false;
            if (r_new <= tol) return 0;                           // CTR_FK_SHOOT_OK
            if (!(r_new == r_new) || !solve_small<U>(J, f)) return fail(2, max_iter);     // CTR_FK_SHOOT_SINGULAR
#pragma unroll
            for (int r = 0; r < U; ++r) d[r] = fmin(1.0, fmax(-1.0, f[r]));
            lam = 1.0;
        } else {
            lam *= 0.5;
            if (++nrej > max_reject) return fail(4, max_iter);    // CTR_FK_SHOOT_STALLED (or start again)
        }
        if (it >= max_iter) { keep_best(); restore_best(); return 3; }        // CTR_FK_SHOOT_MAX_ITER
        return kRunning;
    }
};

// ---- rigid transform as (unit quaternion, translation): used by the warp-parallel backbone pass
struct QT {
    double w, x, y, z, px, py, pz;
};
CTR_HD QT qt_identity()
{
    QT t; t.w = 1.0; t.x = t.y = t.z = 0.0; t.px = t.py = t.pz = 0.0;
    return t;
}
// a * b : rotation a.q (x) b.q, translation a.p + R(a.q) b.p
CTR_HD QT qt_mul(const QT& a, const QT& b)
{
    QT r;
    const double tx = x_fma(a.y, b.pz, -(a.z * b.py));
    const double ty = x_fma(a.z, b.px, -(a.x * b.pz));
    const double tz = x_fma(a.x, b.py, -(a.y * b.px));
    const double ix = x_fma(a.w, tx, x_fma(a.y, tz, -(a.z * ty)));
    const double iy = x_fma(a.w, ty, x_fma(a.z, tx, -(a.x * tz)));
    const double iz = x_fma(a.w, tz, x_fma(a.x, ty, -(a.y * tx)));
    r.px = x_fma(2.0, ix, a.px + b.px);
    r.py = x_fma(2.0, iy, a.py + b.py);
    r.pz = x_fma(2.0, iz, a.pz + b.pz);
    r.w = x_fma(-a.z, b.z, x_fma(-a.y, b.y, x_fma(-a.x, b.x, a.w * b.w)));
    r.x = x_fma(-a.z, b.y, x_fma(a.y, b.z, x_fma(a.x, b.w, a.w * b.x)));
    r.y = x_fma(-a.x, b.z, x_fma(a.z, b.x, x_fma(a.y, b.w, a.w * b.y)));
    r.z = x_fma(-a.y, b.x, x_fma(a.x, b.y, x_fma(a.z, b.w, a.w * b.z)));
    return r;
}
// Frame_k = Init * Rz(ang) * I  (robot_i.h:714-716) with the rotations composed as quaternions.
// (sh, ch) = sin/cos(ang/2).  Output column-major 3x4.
CTR_HD void qt_frame(const KRobot& rb, const QT& I, double sh, double ch, double* f)
{
    // q1 = q_init (x) q_z(ang/2)
    const double w1 = x_fma(-rb.qini[3], sh, rb.qini[0] * ch);
    const double x1 = x_fma(rb.qini[2], sh, rb.qini[1] * ch);
    const double y1 = x_fma(-rb.qini[1], sh, rb.qini[2] * ch);
    const double z1 = x_fma(rb.qini[0], sh, rb.qini[3] * ch);
    // q = q1 (x) I.q
    const double w = x_fma(-z1, I.z, x_fma(-y1, I.y, x_fma(-x1, I.x, w1 * I.w)));
    const double x = x_fma(-z1, I.y, x_fma(y1, I.z, x_fma(x1, I.w, w1 * I.x)));
    const double y = x_fma(-x1, I.z, x_fma(z1, I.x, x_fma(y1, I.w, w1 * I.y)));
    const double z = x_fma(-y1, I.x, x_fma(x1, I.y, x_fma(z1, I.w, w1 * I.z)));
    const double x2 = x + x, y2 = y + y, z2 = z + z;
    f[0] = 1.0 - x_fma(y2, y, z2 * z);
    f[1] = x_fma(x2, y, w * z2);
    f[2] = x_fma(x2, z, -(w * y2));
    f[3] = x_fma(x2, y, -(w * z2));
    f[4] = 1.0 - x_fma(x2, x, z2 * z);
    f[5] = x_fma(y2, z, w * x2);
    f[6] = x_fma(x2, z, w * y2);
    f[7] = x_fma(y2, z, -(w * x2));
    f[8] = 1.0 - x_fma(x2, x, y2 * y);
    // translation: p_init + R_init (Rz(ang) I.p);  sin/cos(ang) from the half angle
    const double sa = 2.0 * sh * ch, ca = x_fma(ch, ch, -(sh * sh));
    const double rx = x_fma(ca, I.px, -(sa * I.py));
    const double ry = x_fma(ca, I.py, sa * I.px);
    const double rz = I.pz;
    f[9]  = x_fma(rb.ini[0], rx, x_fma(rb.ini[3], ry, x_fma(rb.ini[6], rz, rb.ini[9])));
    f[10] = x_fma(rb.ini[1], rx, x_fma(rb.ini[4], ry, x_fma(rb.ini[7], rz, rb.ini[10])));
    f[11] = x_fma(rb.ini[2], rx, x_fma(rb.ini[5], ry, x_fma(rb.ini[8], rz, rb.ini[11])));
}

// a * inv(b) for unit-quaternion transforms: rotation a.q (x) conj(b.q), translation a.p - R(r.q) b.p
CTR_HD QT qt_mul_inv(const QT& a, const QT& b)
{
    QT r;
    r.w = x_fma(a.z, b.z, x_fma(a.y, b.y, x_fma(a.x, b.x, a.w * b.w)));
    r.x = x_fma(a.z, b.y, x_fma(-a.y, b.z, x_fma(a.x, b.w, -(a.w * b.x))));
    r.y = x_fma(a.x, b.z, x_fma(-a.z, b.x, x_fma(a.y, b.w, -(a.w * b.y))));
    r.z = x_fma(a.y, b.x, x_fma(-a.x, b.y, x_fma(a.z, b.w, -(a.w * b.z))));
    const double tx = x_fma(r.y, b.pz, -(r.z * b.py));
    const double ty = x_fma(r.z, b.px, -(r.x * b.pz));
    const double tz = x_fma(r.x, b.py, -(r.y * b.px));
    const double ix = x_fma(r.w, tx, x_fma(r.y, tz, -(r.z * ty)));
    const double iy = x_fma(r.w, ty, x_fma(r.z, tx, -(r.x * tz)));
    const double iz = x_fma(r.w, tz, x_fma(r.x, ty, -(r.y * tx)));
    r.px = x_fma(-2.0, ix, a.px - b.px);
    r.py = x_fma(-2.0, iy, a.py - b.py);
    r.pz = x_fma(-2.0, iz, a.pz - b.pz);
    return r;
}

// ---- frame emission during a repeated sweep (backbone kernels) ----------------------------------
// A first sweep with WANT_TIP gives the total product T = E_0 ... E_ktip.  A second sweep from the
// tip then has the frame product of every sample for the price of one inverse step:
//     I_ktip = T,   I_{k-1} = I_k inv(E_k),   Frame_k = Init Rz(alpha_k + theta) I_k.
// on_sample functor for a FULL sweep (FastSweep::run_emit): operator() only records the sample;
// the frame is built one iteration later by emit_hot()/emit_generic(), called from
// FastSweep::step next to the following sample's arithmetic, so the two dependency chains overlap.
// I starts as the total product T = E_0 ... E_ktip (a first sweep with WANT_TIP); emitting sample k
// stores Frame_k = Init Rz(alpha_k + theta) I and leaves I = I inv(E_k) = I_{k-1}.  Samples above
// ktip (the dropped tip sample of step mode) are skipped.
template <class StoreFrame>
struct PendingFrameEmitter {
    const KRobot* rb;
    StepConsts    kc;
    double        theta;
    StoreFrame    store;
    QT            I;
    int           ktip;
    int           pk;                    // pending sample, -1: none
    double        pux, puy, puz, pal;
    double        psxy, pm;              // |u_xy|^2 and (h|u|/2)^2 of the pending sample (pending_hot)
    double        phs, phc;              // sin/cos(alpha_k / 2) of the pending sample (FastSweep<.., HALF>)
    double        ths, thc;              // sin/cos(theta / 2)

    CTR_HD void operator()(int k, double ux, double uy, double uz, double al)
    {
        pk = (k <= ktip) ? k : -1;
        pux = ux; puy = uy; puz = uz; pal = al;
    }
    CTR_HD void set_half(double s, double c) { phs = s; phc = c; }
    // branch-free forms apply: series SE(3) step of the first tier
    CTR_HD bool pending_hot()
    {
        psxy = x_fma(puy, puy, pux * pux);
        const double n2 = x_fma(puz, puz, psxy);
        pm = n2 * kc.hh2;
        return pk >= 0 && le_nonneg(kc.ztol2, n2) && hi_word(pm) < 0x3f700000;
    }
    CTR_HD void emit_hot()               // after pending_hot() returned true
    {
        double F[12];
        // (alpha_k + theta) / 2 from the carried half-angle pair and the constant pair of theta / 2
        const double shf = x_fma(phc, ths, phs * thc);
        const double chf = x_fma(-phs, ths, phc * thc);
        qt_frame(*rb, I, shf, chf, F);
        store(pk, F);
        // step series: rotation to full precision; the third-order translation term only needs ~1e-9 relative
        // (FastSweep::accumulate_hot)
        double qa = x_fma(CTRK_FC.qa4[0], pm, CTRK_FC.qa4[1]);
        double qb = x_fma(CTRK_FC.qb4[0], pm, CTRK_FC.qb4[1]);
        double qc = x_fma(CTRK_FC.qc4[2], pm, CTRK_FC.qc4[3]);
        qa = x_fma(qa, pm, CTRK_FC.qa4[2]);
        qb = x_fma(qb, pm, CTRK_FC.qb4[2]);
        qc = x_fma(qc, pm, 1.0 / 6.0);
        qa = x_fma(qa, pm, CTRK_FC.qa4[3]);
        qb = x_fma(qb, pm, CTRK_FC.qb4[3]);
        qa = x_fma(qa, pm, 1.0);
        qb = x_fma(qb, pm, 1.0);
        const double b2 = kc.h * qb, b = kc.hh * qb;
        QT E;
        E.w = qa; E.x = b * pux; E.y = b * puy; E.z = b * puz;
        const double C = kc.h3 * qc, cz = C * puz;
        E.px = x_fma(cz, pux, b2 * E.y);
        E.py = x_fma(cz, puy, -(b2 * E.x));
        E.pz = x_fma(-C, psxy, kc.h);
        I = qt_mul_inv(I, E);
    }
    CTR_HD void emit_generic()
    {
        if (pk < 0) return;
        double shf, chf, F[12];
        fast_sincos(0.5 * (pal + theta), shf, chf);
        qt_frame(*rb, I, shf, chf, F);
        store(pk, F);
        QT E;
        se3_step_fast_k(pux, puy, puz, kc, E.w, E.x, E.y, E.z, E.px, E.py, E.pz);
        I = qt_mul_inv(I, E);
    }
};

// The emitter of one configuration: I starts as the total product `total` of a first sweep.
template <class L, class StoreFrame>
CTR_HD PendingFrameEmitter<StoreFrame> make_frame_emitter(const KRobot& rb, const Ctl<L>& c, const StoreFrame& store, const QT& total)
{
    PendingFrameEmitter<StoreFrame> em{&rb, make_step_consts(c.h, rb.ztol), c.theta, store, total, c.nframes - 1,
                                       -1, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 1.0, 0.0, 1.0};
    fast_sincos(0.5 * c.theta, em.ths, em.thc);
    return em;
}

// ---- finite-difference Jacobians (robot_i.h:901-1072, :41-56) ------------------------------------
// on_sample functor: the frame product of sample ks, B = E_0 ... E_ks, accumulated next to the solver's own
// (software-pipelined) tip product.  Its chain is independent of the sweep's, so the two overlap.
struct SampleProduct {
    StepConsts kc;
    int        ks;
    QT*        B;        // starts as the identity
    double*    a_s;      // angle of the innermost tube at sample ks
    CTR_HD void operator()(int k, double ux, double uy, double uz, double al) const
    {
        if (k > ks) return;
        if (k == ks) *a_s = al;
        QT E;
        se3_step_fast_k(ux, uy, uz, kc, E.w, E.x, E.y, E.z, E.px, E.py, E.pz);
        *B = qt_mul(E, *B);
    }
};

// column = d(lo -> hi)/dq (robot_i.h:41-56): translation difference and the vee of (R_hi - R_lo) R_hi^T,
// both divided by dq.
CTR_HD void jac_column(const double* lo, const double* hi, double dq, double* col)
{
    double d[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) d[i] = hi[i] - lo[i];
#define CTRK_JM(i, j) (d[i] * hi[j] + d[3 + (i)] * hi[3 + (j)] + d[6 + (i)] * hi[6 + (j)])     /* sum_k D(i,k) R_hi(j,k) */
    col[0] = (hi[9] - lo[9]) / dq;
    col[1] = (hi[10] - lo[10]) / dq;
    col[2] = (hi[11] - lo[11]) / dq;
    col[3] = ((CTRK_JM(2, 1) - CTRK_JM(1, 2)) * 0.5) / dq;
    col[4] = ((CTRK_JM(0, 2) - CTRK_JM(2, 0)) * 0.5) / dq;
    col[5] = ((CTRK_JM(1, 0) - CTRK_JM(0, 1)) * 0.5) / dq;
#undef CTRK_JM
}
// analytic column 0 (robot_i.h:956-959, :1064-1072): rotation about the entry frame's z axis
CTR_HD void jac_column0(const KRobot& rb, const double* f, double* jac)
{
    const double zx = rb.ini[6], zy = rb.ini[7], zz = rb.ini[8];
    const double rx = f[9] - rb.ini[9], ry = f[10] - rb.ini[10], rz = f[11] - rb.ini[11];
    jac[0] = zy * rz - zz * ry;
    jac[1] = zz * rx - zx * rz;
    jac[2] = zx * ry - zy * rx;
    jac[3] = zx; jac[4] = zy; jac[5] = zz;
}

// STRICT counterpart: the 3x4 product in the reference's operation order (robot_i.h:705-717)
struct SampleProductStrict {
    double  h, ztol;
    int     ks;
    double* M;        // [12], starts as the identity
    double* a_s;
    CTR_HD void operator()(int k, double ux, double uy, double uz, double al) const
    {
        if (k > ks) return;
        if (k == ks) *a_s = al;
        double e[12];
        tf_step_strict(ux, uy, uz, h, ztol, e);
        tf_mul(e, M, M);
    }
};

// One solve that yields the tip frame and, when ks >= 0, the frame of sample ks.
template <class L, bool WITH_SAMPLE, bool FAST = true>
CTR_HD void solve_tip_and_sample(const KRobot& rb, const Ctl<L>& c, int ks, double* tip, double* smp)
{
    if (!WITH_SAMPLE || ks < 0) {
        if (FAST) solve_fast<L, true>(rb, c, tip, nullptr, NoSample{});
        else      solve_strict<L, true>(rb, c, tip, nullptr, NoSample{});
        return;
    }
    double as = 0.0;
    if (FAST) {
        QT B = qt_identity();
        SampleProduct f{make_step_consts(c.h, rb.ztol), ks, &B, &as};
        solve_fast<L, true>(rb, c, tip, nullptr, f);
        double sh, ch;
        fast_sincos(0.5 * (as + c.theta), sh, ch);
        qt_frame(rb, B, sh, ch, smp);
    } else {
        double M[12];
        tf_set_identity(M);
        SampleProductStrict f{c.h, rb.ztol, ks, M, &as};
        solve_strict<L, true>(rb, c, tip, nullptr, f);
        const double a = as + c.theta;
        tf_finish(rb.ini, M, sin(a), cos(a), smp);
    }
}

// Every calcJacobian form of the reference for one configuration (what one GPU thread runs):
//   mode NSAMPLE, ks < 0  : calcKinematic(JV, N) + calcJacobian(JV, Tip, N, ..)              robot_i.h:917-960
//   mode STEP,    ks < 0  : calcJacobian(JV, ArcLengthStep, ..): base pose from calcKinematic(JV, step), perturbed
//                           solves in NSamples mode with N = m_Samples of that call (:901-907) — the reference
//                           mixes the two discretisations; reproduced
//   ks >= 0               : additionally the Jacobian of sample ks' frame (:908-915, :1012-1072)
//   given                 : tip_out (and smp_out) are INPUTS — the base poses come from the caller, as in
//                           calcJacobian(JV, TTip, N, ..) and (JV, TTip, TiSample, N, .., iSample, ..) (:917, :1012);
//                           no base solve is run
// jac_tip / jac_smp: [NJ][6]; tip_out / smp_out nullable [12]; returns the status of the base solve
// (non-zero: everything NaN), n_out = m_Samples of the base solve.  ks >= m_Samples: the reference would read a
// stale frame; the sample Jacobian and frame are NaN.
// WITH_SAMPLE = false compiles the sample-frame path out (tip-only forms keep their register budget).
// FAST = false: every solve is the STRICT solver (CTR_FK_FLAG_STRICT; built with -fmad=false).
template <class L, bool WITH_SAMPLE, bool FAST = true>
CTR_HD int jacobian_config(const KRobot& rb, const double* q, int mode, double step, int nsamples, int ks,
                           double fd_trans, double fd_rot, double* jac_tip, double* jac_smp, double* tip_out,
                           double* smp_out, int32_t* n_out, bool given = false)
{
    constexpr int S = L::S, T = L::T, NJ = L::NJ;
    union { long long i; double d; } qn; qn.i = 0x7ff8000000000000ll;
    const double qnan = qn.d;
    Ctl<L> c;
    setup_control<L>(rb, q, mode, step, nsamples, c);
    if (n_out) *n_out = c.status ? 0 : c.nframes;
    if (c.status) {
        for (int e = 0; e < NJ * 6; ++e) { jac_tip[e] = qnan; if (jac_smp) jac_smp[e] = qnan; }
        if (tip_out && !given) for (int e = 0; e < 12; ++e) tip_out[e] = qnan;
        if (smp_out && !given) for (int e = 0; e < 12; ++e) smp_out[e] = qnan;
        return c.status;
    }
    const bool smp_ok = WITH_SAMPLE && ks >= 0 && ks < c.nframes;
    if (!WITH_SAMPLE) { jac_smp = nullptr; smp_out = nullptr; }
    const int  kse = smp_ok ? ks : -1;
    const int  npert = (mode == 0) ? c.nframes : nsamples;          // m_Samples, robot_i.h:906
    double tip0[12], smp0[12];
    if (given) {
        for (int e = 0; e < 12; ++e) { tip0[e] = tip_out[e]; smp0[e] = (WITH_SAMPLE && smp_ok) ? smp_out[e] : qnan; }
    } else {
        solve_tip_and_sample<L, WITH_SAMPLE, FAST>(rb, c, kse, tip0, smp0);
        if (tip_out) for (int e = 0; e < 12; ++e) tip_out[e] = tip0[e];
        if (smp_out) for (int e = 0; e < 12; ++e) smp_out[e] = smp_ok ? smp0[e] : qnan;
    }

    // column 2 (alpha of tube 0) is zero (robot_i.h:928)
#pragma unroll
    for (int e = 0; e < 6; ++e) { jac_tip[2 * 6 + e] = 0.0; if (jac_smp) jac_smp[2 * 6 + e] = smp_ok ? 0.0 : qnan; }

    // perturbations 0..T-2: tube angles 1..T-1 (+fd_rot); T-1..T-2+S: section extensions
#pragma unroll 1
    for (int p = 0; p < (T - 1) + S; ++p) {
        int    jp = 0;
        double dq = fd_rot;
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
            dq = (keep + fd_trans >= len) ? -fd_trans : fd_trans;       // robot_i.h:947-950
        }
        double qq[NJ];
#pragma unroll
        for (int j = 0; j < NJ; ++j) qq[j] = (j == jp) ? q[j] + dq : q[j];
        Ctl<L> cp;
        setup_control<L>(rb, qq, 1, 0.0, npert, cp);
        double col[6], cols[6];
        if (cp.status) {
#pragma unroll
            for (int e = 0; e < 6; ++e) col[e] = cols[e] = qnan;
        } else {
            double tp[12], sp[12];
            solve_tip_and_sample<L, WITH_SAMPLE, FAST>(rb, cp, (kse >= 0 && kse < cp.nframes) ? kse : -1, tp, sp);
            jac_column(tip0, tp, dq, col);
            if (WITH_SAMPLE && kse >= 0 && kse < cp.nframes) jac_column(smp0, sp, dq, cols);
            else
#pragma unroll
                for (int e = 0; e < 6; ++e) cols[e] = qnan;
        }
#pragma unroll
        for (int e = 0; e < 6; ++e) { jac_tip[jp * 6 + e] = col[e]; if (jac_smp) jac_smp[jp * 6 + e] = cols[e]; }
    }
    jac_column0(rb, tip0, jac_tip);
    if (jac_smp) {
        if (smp_ok) jac_column0(rb, smp0, jac_smp);
        else
#pragma unroll
            for (int e = 0; e < 6; ++e) jac_smp[e] = qnan;
    }
    return 0;
}

// robot_i.h:1286-1307: outer radius of the section covering each sample.  Writes radius[0..n)
// with stride `stride` doubles.
template <class L>
CTR_HD void fill_radius(const KRobot& rb, const Ctl<L>& c, double* radius, int64_t stride)
{
    const int n = c.nframes;
    int    begin = 0;
    double rsec  = 0.0;
#pragma unroll
    for (int s = 0; s < L::S; ++s) {
        if (begin >= n) continue;
        rsec = rb.rad[s];
        const double q = x_div(c.end[s], c.h);
        int stop = (q >= 0.0 && q < (double)n) ? (int)q : n;
        if (stop < begin) stop = begin;
        for (int k = begin; k < stop; ++k) radius[(int64_t)k * stride] = rsec;
        begin = stop;
    }
    if (n > 0) radius[(int64_t)(n - 1) * stride] = rsec;
}

}  // namespace ctrk
