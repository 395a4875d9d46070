// ctr_fk_core_f32.cuh — FP32 variant of the tip solver (the reference's `float` instantiation,
// ctr_robot.cpp:724-731, re-done for the GPU).  State arithmetic (tube angles, torsion, curvature,
// the SE(3) product) is single precision; the control scalars — sum of Phi, N, h, the arc-position
// sequence and the section-interval compares — stay FP64 and are shared with the FP64 solvers
// (setup_control), so the sample count and returned step are bit-identical to the FP64 path
// (a 1-ulp difference there moves the tip by a whole step).  Inputs and outputs are
// double.  Accuracy against the FP64 oracle is stated, with its tolerance, in
// tests/test_gpu_parity.py::test_f32_variant.
#pragma once

#include "ctr_fk_core.cuh"

namespace ctrk {

CTR_HD void sincos_f32(float x, float& s, float& c)
{
#ifdef __CUDA_ARCH__
    sincosf(x, &s, &c);
#else
    s = sinf(x); c = cosf(x);
#endif
}

// SE(3) step in single precision: series in m = (h|u|/2)^2 up to m <= 2^-6 (truncation 1e-9,
// below float epsilon), closed form beyond; the reference's zero-curvature quirk is reproduced.
CTR_HD void se3_step_f32(float ux, float uy, float uz, float h, float ztol, float& qw, float& qx,
                         float& qy, float& qz, float& px, float& py, float& pz)
{
    const float sxy = fmaf(uy, uy, ux * ux);
    const float n2  = fmaf(uz, uz, sxy);
    const float hh  = 0.5f * h;
    const float m   = n2 * (hh * hh);
    if (n2 < ztol * ztol) {
        qw = 1.f; qx = qy = qz = 0.f; px = py = 0.f; pz = h * sqrtf(n2);
        return;
    }
    if (m <= 0.015625f) {
        const float qa = fmaf(fmaf(fmaf(-1.f / 720.f, m, 1.f / 24.f), m, -0.5f), m, 1.f);
        const float qb = fmaf(fmaf(fmaf(-1.f / 5040.f, m, 1.f / 120.f), m, -1.f / 6.f), m, 1.f);
        const float qc = fmaf(fmaf(16.f / 5040.f, m, -4.f / 120.f), m, 1.f / 6.f);
        const float b  = hh * qb, b2 = h * qb;
        qw = qa; qx = b * ux; qy = b * uy; qz = b * uz;
        const float C = (h * h * h) * qc, cz = C * uz;
        px = fmaf(cz, ux, b2 * qy);
        py = fmaf(cz, uy, -(b2 * qx));
        pz = fmaf(-C, sxy, h);
        return;
    }
    const float nrm = sqrtf(n2), t = h * nrm;
    float sh, ch, st, ct;
    sincos_f32(0.5f * t, sh, ch);
    sincos_f32(t, st, ct);
    const float b = sh / nrm;
    qw = ch; qx = b * ux; qy = b * uy; qz = b * uz;
    const float Bc = (1.f - ct) / n2, C = (t - st) / (n2 * nrm);
    px = Bc * uy + C * ux * uz;
    py = -Bc * ux + C * uy * uz;
    pz = h - C * sxy;
}

template <class L>
CTR_HD void solve_f32(const KRobot& rb, const Ctl<L>& c, double* tip)
{
    constexpr int S = L::S, T = L::T;
    float al[T], uzs[T];
#pragma unroll
    for (int t = 0; t < T; ++t) { al[t] = (float)c.al[t]; uzs[t] = 0.f; }
    const float h = (float)c.h;
    // The descriptor's zero tolerance is used as is (1e-12 by default).  A float-sized tolerance
    // such as 1e-6 would push the kappa = 1e-6 tip section of the shipped robots into the
    // zero-curvature branch, which drops a whole step of length h.
    const float ztol = (float)rb.ztol;
    float kg[S], rk[S], czc[T > 1 ? T - 1 : 1], hko[T > 1 ? T - 1 : 1];
#pragma unroll
    for (int s = 0; s < S; ++s) { kg[s] = (float)rb.kg[s]; rk[s] = (float)rb.rk[s]; }
#pragma unroll
    for (int t = 0; t + 1 < T; ++t) {
        czc[t] = (float)rb.czc[t];
        hko[t] = (float)(c.h * rb.onu[t] * rb.kappa[L::sec_of(t)]);
    }
    const float cl = (float)rb.cl;
    float Qw = 1.f, Qx = 0.f, Qy = 0.f, Qz = 0.f, Px = 0.f, Py = 0.f, Pz = 0.f, a_tip = 0.f;
    double pos = c.arc_end;
    const int tipk = c.nframes - 1;

#pragma unroll 1
    for (int k = c.n - 1; k >= 0; --k) {
        float sn[T], cs[T];
        if (k == c.n - 1) sincos_f32(al[0], sn[0], cs[0]);
        else { sn[0] = 0.f; cs[0] = 1.f; }
#pragma unroll
        for (int t = 1; t < T; ++t) sincos_f32(al[t], sn[t], cs[t]);
        bool in_k[S], in_c[S];
#pragma unroll
        for (int s = 0; s < S; ++s) {
            in_k[s] = (s == S - 1) ? true : le_nonneg(pos, c.end[s]);
            in_c[s] = in_k[s] && le_nonneg(c.start[s], pos);
        }
        float rks = rk[S - 1];
#pragma unroll
        for (int s = S - 2; s >= 0; --s) rks = in_k[s] ? rk[s] : rks;
        float cx = 0.f, cy = 0.f;
#pragma unroll
        for (int s = 0; s < S; ++s) {
            const float g = in_c[s] ? kg[s] : 0.f;
#pragma unroll
            for (int j = 0; j < L::nt(s); ++j) {
                const int t = L::t0(s) + j;
                cx = fmaf(-sn[t], g, cx);
                cy = fmaf(cs[t], g, cy);
            }
        }
        cx *= rks; cy *= rks;
        float uxs[T];
#pragma unroll
        for (int t = 0; t < T; ++t) uxs[t] = fmaf(cs[t], cx, sn[t] * cy);
        const float uy_in = fmaf(cs[T - 1], cy, -(sn[T - 1] * cx));
        const float uz_k  = uzs[T - 1];
        float a_k = al[T - 1];
        if (k >= 1) {
            const float z0 = uzs[0];
            al[0] = 0.f;
#pragma unroll
            for (int t = 1; t < T; ++t) al[t] = fmaf(h, uzs[t] - z0, al[t]);
            a_k = al[T - 1];
            float cz = 0.f;
#pragma unroll
            for (int t = T - 2; t >= 0; --t) {
                const float hk = in_c[L::sec_of(t)] ? hko[t] : 0.f;
                cz     = fmaf(czc[t], uzs[t], cz);
                uzs[t] = fmaf(hk, uxs[t], uzs[t]);
            }
            uzs[T - 1] = -cz * cl;
        }
        if (k <= tipk) {
            float ew, ex, ey, ez, pex, pey, pez;
            se3_step_f32(uxs[T - 1], uy_in, uz_k, h, ztol, ew, ex, ey, ez, pex, pey, pez);
            const float tx = fmaf(ey, Pz, -(ez * Py));
            const float ty = fmaf(ez, Px, -(ex * Pz));
            const float tz = fmaf(ex, Py, -(ey * Px));
            const float ix = fmaf(ew, tx, fmaf(ey, tz, -(ez * ty)));
            const float iy = fmaf(ew, ty, fmaf(ez, tx, -(ex * tz)));
            const float iz = fmaf(ew, tz, fmaf(ex, ty, -(ey * tx)));
            Px = fmaf(2.f, ix, Px + pex);
            Py = fmaf(2.f, iy, Py + pey);
            Pz = fmaf(2.f, iz, Pz + pez);
            const float w2 = fmaf(-ez, Qz, fmaf(-ey, Qy, fmaf(-ex, Qx, ew * Qw)));
            const float x2 = fmaf(-ez, Qy, fmaf(ey, Qz, fmaf(ex, Qw, ew * Qx)));
            const float y2 = fmaf(-ex, Qz, fmaf(ez, Qx, fmaf(ey, Qw, ew * Qy)));
            const float z2 = fmaf(-ey, Qx, fmaf(ex, Qy, fmaf(ez, Qw, ew * Qz)));
            Qw = w2; Qx = x2; Qy = y2; Qz = z2;
            if (k == tipk) a_tip = a_k;
        }
        if (k >= 1) pos = x_sub(pos, c.h);
    }
    // finish in double: the entry frame and theta are double inputs
    const double qw = Qw, qx = Qx, qy = Qy, qz = Qz;
    const double sc = 2.0 / (qw * qw + qx * qx + qy * qy + qz * qz);
    double Pm[12];
    Pm[0] = 1.0 - sc * (qy * qy + qz * qz);
    Pm[1] = sc * (qx * qy + qw * qz);
    Pm[2] = sc * (qx * qz - qw * qy);
    Pm[3] = sc * (qx * qy - qw * qz);
    Pm[4] = 1.0 - sc * (qx * qx + qz * qz);
    Pm[5] = sc * (qy * qz + qw * qx);
    Pm[6] = sc * (qx * qz + qw * qy);
    Pm[7] = sc * (qy * qz - qw * qx);
    Pm[8] = 1.0 - sc * (qx * qx + qy * qy);
    Pm[9] = Px; Pm[10] = Py; Pm[11] = Pz;
    double sa, ca;
    fast_sincos((double)a_tip + c.theta, sa, ca);
    tf_finish(rb.ini, Pm, sa, ca, tip);
}

}  // namespace ctrk
