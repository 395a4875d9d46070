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
//             SE(3) step as power series in (h|u|)^2 (no sqrt, no division, no trig), rotation
//             accumulated as a unit quaternion, custom sin/cos with 2-term Cody-Waite reduction.
//             Differences to STRICT are O(1e-15) per step.
// Both use the SAME control scalars (sum of Phi, N, h, the arc-position sequence and the
// section-interval compares), evaluated with non-contracted IEEE operations, because those decide
// which samples exist (SURVEY.md §7 "bit-critical scalars").
#pragma once

#include <math.h>
#include <stdint.h>

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

// sin and cos of a moderate argument (|x| < ~1e5): k = rint(x*2/pi), r = x - k*pi/2 by a 3-term
// Cody-Waite split (exact products for |k| < 2^20 or so), minimax-grade Taylor kernels on
// |r| <= pi/4 (degree 13 / 14), quadrant fix-up in integer arithmetic.  Max error ~1 ulp.
// Falls back to libm for huge or non-finite arguments.
CTR_HD void fast_sincos(double x, double& s_out, double& c_out)
{
    if (!(fabs(x) < 1.0e5)) {
#ifdef __CUDA_ARCH__
        sincos(x, &s_out, &c_out);
#else
        s_out = sin(x); c_out = cos(x);
#endif
        return;
    }
    const double two_over_pi = 6.36619772367581382433e-01;
    const double magic = 6755399441055744.0;   // 1.5 * 2^52: round-to-nearest-integer trick
    const double kd_m  = x_fma(x, two_over_pi, magic);
    const double kd    = kd_m - magic;
#ifdef __CUDA_ARCH__
    const int q = __double2loint(kd_m);
#else
    union { double d; uint64_t u; } cvt; cvt.d = kd_m;
    const int q = (int)(uint32_t)cvt.u;
#endif
    // pi/2 = p1 + p2 + p3, p1 and p2 with 33 significant bits
    const double p1 = 1.57079632673412561417e+00;
    const double p2 = 6.07710050630396597660e-11;
    const double p3 = 2.02226624879595063154e-21;
    double r = x_fma(-kd, p1, x);
    r = x_fma(-kd, p2, r);
    r = x_fma(-kd, p3, r);
    const double z = r * r;
    // sin(r) = r + r*z*(S1 + z*(S2 + ... S6))      (fdlibm __kernel_sin coefficients)
    double ps = 1.58969099521155010221e-10;
    ps = x_fma(ps, z, -2.50507602534068634195e-08);
    ps = x_fma(ps, z, 2.75573137070700676789e-06);
    ps = x_fma(ps, z, -1.98412698298579493134e-04);
    ps = x_fma(ps, z, 8.33333333332248946124e-03);
    ps = x_fma(ps, z, -1.66666666666666324348e-01);
    const double sr = x_fma(r * z, ps, r);
    // cos(r) = 1 - z/2 + z*z*(C1 + z*(C2 + ... C6))  (fdlibm __kernel_cos coefficients)
    double pc = -1.13596475577881948265e-11;
    pc = x_fma(pc, z, 2.08757232129817482790e-09);
    pc = x_fma(pc, z, -2.75573143513906633035e-07);
    pc = x_fma(pc, z, 2.48015872894767294178e-05);
    pc = x_fma(pc, z, -1.38888888888741095749e-03);
    pc = x_fma(pc, z, 4.16666666666666019037e-02);
    const double cr = x_fma(z * z, pc, x_fma(-0.5, z, 1.0));
    // quadrant: q&1 swaps, bit 1 of q negates sin, bit 1 of (q+1) negates cos
    const double s1 = (q & 1) ? cr : sr;
    const double c1 = (q & 1) ? sr : cr;
    s_out = (q & 2) ? -s1 : s1;
    c_out = ((q + 1) & 2) ? -c1 : c1;
}

// Series used by the FAST SE(3) step, in m = (h|u|/2)^2 (half-angle squared):
//   qa(m) = cos(sqrt m)            = 1 - m/2 + m^2/24 - ...
//   qb(m) = sin(sqrt m)/sqrt m     = 1 - m/6 + m^2/120 - ...
//   qc(m) = (t - sin t)/t^3, t = 2 sqrt m  = 1/6 - 4m/120 + 16 m^2/5040 - ...
// Truncated after m^7; for m <= kFastM the truncation error is below 1e-18.
constexpr double kFastM = 0.0625;   // |h u| <= 0.5 rad per step

CTR_HD void se3_series(double m, double& qa, double& qb, double& qc)
{
    double a = -1.0 / 87178291200.0;                 // -1/14!
    a = x_fma(a, m, 1.0 / 479001600.0);              //  1/12!
    a = x_fma(a, m, -1.0 / 3628800.0);               // -1/10!
    a = x_fma(a, m, 1.0 / 40320.0);
    a = x_fma(a, m, -1.0 / 720.0);
    a = x_fma(a, m, 1.0 / 24.0);
    a = x_fma(a, m, -0.5);
    qa = x_fma(a, m, 1.0);
    double b = -1.0 / 1307674368000.0;               // -1/15!
    b = x_fma(b, m, 1.0 / 6227020800.0);             //  1/13!
    b = x_fma(b, m, -1.0 / 39916800.0);              // -1/11!
    b = x_fma(b, m, 1.0 / 362880.0);
    b = x_fma(b, m, -1.0 / 5040.0);
    b = x_fma(b, m, 1.0 / 120.0);
    b = x_fma(b, m, -1.0 / 6.0);
    qb = x_fma(b, m, 1.0);
    // (t - sin t)/t^3 = sum_{j>=0} (-1)^j t^(2j)/(2j+3)!,  t^2 = 4m
    double cc = 4096.0 / 1307674368000.0;            // +4^6/15!
    cc = x_fma(cc, m, -1024.0 / 6227020800.0);       // -4^5/13!
    cc = x_fma(cc, m, 256.0 / 39916800.0);           // +4^4/11!
    cc = x_fma(cc, m, -64.0 / 362880.0);             // -4^3/9!
    cc = x_fma(cc, m, 16.0 / 5040.0);                // +4^2/7!
    cc = x_fma(cc, m, -4.0 / 120.0);                 // -4/5!
    qc = x_fma(cc, m, 1.0 / 6.0);
}

// One SE(3) step as (unit quaternion (qw, qv), translation pe), FAST form.
//   rotation  exp(h [u]x)  ->  qw = cos(t/2), qv = sin(t/2)/|u| * u,  t = h|u|
//   translation (robot_i.h:1257-1259, exact arc integral):
//       pe = h e3 + B (u x e3) + C (u u_z - |u|^2 e3),  B = (1-cos t)/|u|^2, C = (t-sin t)/|u|^3
CTR_HD void se3_step_fast(double ux, double uy, double uz, double h, double ztol, double& qw,
                          double& qx, double& qy, double& qz, double& px, double& py, double& pz)
{
    const double sxy = x_fma(uy, uy, ux * ux);
    const double n2  = x_fma(uz, uz, sxy);
    const double hh  = 0.5 * h;
    const double m   = n2 * (hh * hh);
    if (n2 < ztol * ztol) {
        // |u| "equals" zero: the reference returns the identity rotation and a translation of
        // h*|u| (~0, not h) along z (robot_i.h:1245, :1264-1266).  Reproduced.
        qw = 1.0; qx = qy = qz = 0.0; px = py = 0.0; pz = h * sqrt(n2);
        return;
    }
    if (m <= kFastM) {
        double qa, qb, qc;
        se3_series(m, qa, qb, qc);
        const double b  = hh * qb;               // sin(t/2)/|u|
        qw = qa;
        qx = b * ux; qy = b * uy; qz = b * uz;
        const double b2 = 2.0 * b;               // B = 2 b^2
        const double C  = (h * h * h) * qc;
        const double cz = C * uz;
        px = x_fma(cz, ux, b2 * qy);             // B uy + C ux uz
        py = x_fma(cz, uy, -(b2 * qx));          // -B ux + C uy uz
        pz = x_fma(-C, sxy, h);                  // h - C (ux^2+uy^2)
        return;
    }
    // large rotation per step (coarse sampling): closed form
    const double nrm = sqrt(n2);
    double sh, ch, st, ct;
    const double t = h * nrm;
#ifdef __CUDA_ARCH__
    sincos(0.5 * t, &sh, &ch);
    sincos(t, &st, &ct);
#else
    sh = sin(0.5 * t); ch = cos(0.5 * t); st = sin(t); ct = cos(t);
#endif
    const double b = sh / nrm;
    qw = ch; qx = b * ux; qy = b * uy; qz = b * uz;
    const double Bc = (1.0 - ct) / n2;
    const double C  = (t - st) / (n2 * nrm);
    px = Bc * uy + C * ux * uz;
    py = -Bc * ux + C * uy * uz;
    pz = h - C * sxy;
}

// ---- FAST solver ---------------------------------------------------------------------------
// Same recurrence as solve_strict; see the header comment for what differs.  After the first
// (tip) sample tube 0's angle is pinned to 0 (section_i.h:308), so its sin/cos are (0,1) and are
// not evaluated: the tip sample is peeled off the loop and handled by the general code.
template <class L, class OnSample>
CTR_HD void solve_fast(const KRobot& rb, const Ctl<L>& c, double* tip, double* albase,
                       OnSample on_sample)
{
    constexpr int S = L::S, T = L::T;
    double al[T], uzs[T], sn[T], cs[T];
#pragma unroll
    for (int t = 0; t < T; ++t) { al[t] = c.al[t]; uzs[t] = 0.0; }
    // accumulated suffix product  P = E_k ... E_tip  as quaternion + translation
    double Qw = 1.0, Qx = 0.0, Qy = 0.0, Qz = 0.0, Px = 0.0, Py = 0.0, Pz = 0.0;
    double a_tip = 0.0;
    const double h = c.h;
    double pos = c.arc_end;
    const int tipk = c.nframes - 1;

    // per-configuration constants of the torsion step: h * onu[t] * kappa[sec(t)]
    double hko[T > 1 ? T - 1 : 1];
#pragma unroll
    for (int t = 0; t + 1 < T; ++t) hko[t] = h * rb.onu[t] * rb.kappa[L::sec_of(t)];

#pragma unroll 1
    for (int k = c.n - 1; k >= 0; --k) {
        // sin/cos of the tube angles; tube 0 is exactly 0 after the first update
        if (k == c.n - 1) {
            fast_sincos(al[0], sn[0], cs[0]);
        } else {
            sn[0] = 0.0; cs[0] = 1.0;
        }
#pragma unroll
        for (int t = 1; t < T; ++t) fast_sincos(al[t], sn[t], cs[t]);

        // section masks on the exact arc position
        bool in_k[S], in_c[S];
#pragma unroll
        for (int s = 0; s < S; ++s) {
            in_k[s] = (pos <= c.end[s]);
            in_c[s] = in_k[s] && (pos >= c.start[s]);
        }
        // 1 / (sum of stiffness of the sections that reach this position): End is
        // non-decreasing in s, so the active set is always a suffix {s0..S-1}.
        double rks = rb.rk[S - 1];
#pragma unroll
        for (int s = S - 2; s >= 0; --s) rks = in_k[s] ? rb.rk[s] : rks;

        double cx = 0.0, cy = 0.0;
#pragma unroll
        for (int s = 0; s < S; ++s) {
            const double g = in_c[s] ? rb.kg[s] : 0.0;
#pragma unroll
            for (int j = 0; j < L::nt(s); ++j) {
                const int t = L::t0(s) + j;
                cx = x_fma(-sn[t], g, cx);
                cy = x_fma(cs[t], g, cy);
            }
        }
        cx *= rks;
        cy *= rks;
        double uxs[T];
#pragma unroll
        for (int t = 0; t < T; ++t) uxs[t] = x_fma(cs[t], cx, sn[t] * cy);
        const double uy_in = x_fma(cs[T - 1], cy, -(sn[T - 1] * cx));
        const double uz_k  = uzs[T - 1];

        double a_k = al[T - 1];
        if (k >= 1) {
            const double z0 = uzs[0];
            al[0] = 0.0;
#pragma unroll
            for (int t = 1; t < T; ++t) al[t] = x_fma(h, uzs[t] - z0, al[t]);
            a_k = al[T - 1];
            double cz = 0.0;
#pragma unroll
            for (int t = T - 2; t >= 0; --t) {
                const int s = L::sec_of(t);
                const double cc = in_k[s] ? rb.czc[t] : 0.0;
                const double hk = in_c[s] ? hko[t] : 0.0;
                cz     = x_fma(cc, uzs[t], cz);
                uzs[t] = x_fma(hk, uxs[t], uzs[t]);
            }
            uzs[T - 1] = -cz * rb.cl;
        }
        on_sample(k, uxs[T - 1], uy_in, uz_k, a_k);

        if (tip && k <= tipk) {
            double ew, ex, ey, ez, ex_p, ey_p, ez_p;
            se3_step_fast(uxs[T - 1], uy_in, uz_k, h, rb.ztol, ew, ex, ey, ez, ex_p, ey_p, ez_p);
            // translation: P.p <- pe + R(e) P.p   with R(e) v = v + 2 w (q x v) + 2 q x (q x v)
            const double tx = 2.0 * x_fma(ey, Pz, -(ez * Py));
            const double ty = 2.0 * x_fma(ez, Px, -(ex * Pz));
            const double tz = 2.0 * x_fma(ex, Py, -(ey * Px));
            const double nx = Px + ex_p, ny = Py + ey_p, nz = Pz + ez_p;
            Px = x_fma(ew, tx, nx) + x_fma(ey, tz, -(ez * ty));
            Py = x_fma(ew, ty, ny) + x_fma(ez, tx, -(ex * tz));
            Pz = x_fma(ew, tz, nz) + x_fma(ex, ty, -(ey * tx));
            // rotation: Q <- e (x) Q
            const double w2 = x_fma(-ez, Qz, x_fma(-ey, Qy, x_fma(-ex, Qx, ew * Qw)));
            const double x2 = x_fma(-ez, Qy, x_fma(ey, Qz, x_fma(ex, Qw, ew * Qx)));
            const double y2 = x_fma(-ex, Qz, x_fma(ez, Qx, x_fma(ey, Qw, ew * Qy)));
            const double z2 = x_fma(-ey, Qx, x_fma(ex, Qy, x_fma(ez, Qw, ew * Qz)));
            Qw = w2; Qx = x2; Qy = y2; Qz = z2;
            if (k == tipk) a_tip = a_k;
        }
        if (k >= 1) pos = x_sub(pos, h);
    }
    if (tip) {
        // quaternion -> rotation (normalised: |Q| drifts by ~N ulp), then Init * Rz * P
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
        Pm[9] = Px; Pm[10] = Py; Pm[11] = Pz;
        double sa, ca;
        fast_sincos(a_tip + c.theta, sa, ca);
        tf_finish(rb.ini, Pm, sa, ca, tip);
    }
    if (albase) {
#pragma unroll
        for (int t = 0; t < T; ++t) albase[t] = al[t];
    }
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
