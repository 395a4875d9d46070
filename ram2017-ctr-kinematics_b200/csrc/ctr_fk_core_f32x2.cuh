// ctr_fk_core_f32x2.cuh — FP32 variant, two configurations per thread in packed single precision.
// Blackwell's FP32 pipe executes FFMA2 / FMUL2 / FADD2 (PTX fma/mul/add.rn.f32x2): one instruction,
// two independent lanes of a 64-bit register pair.  A thread-per-configuration FP32 sweep is bound
// by instruction issue (one instruction per cycle and sub-partition, FP32 or not); carrying two
// configurations in the halves of packed registers halves the issue slots of the arithmetic, so
// the loop runs at the rate of the FP32 pipe instead.
// Same numerical scheme as solve_f32 (ctr_fk_core_f32.cuh) with the reformulations of the FP64 FAST
// solver: sin/cos of the tube angles carried and rotated by the per-step increment (re-evaluated
// every 128 steps and whenever an increment is large), SE(3) step as a power series, rotation as a
// quaternion.  Control scalars (Phi sums, N, h, the arc-position sequence and the section-interval
// compares) stay FP64 per configuration, bit-identical to every other solver; the section masks
// they produce change at most 2S times per sweep, so the packed constants they select are
// re-derived only when the position crosses the next interval bound.
// Fixed-N mode only (both configurations of a pair must take the same number of steps); step mode
// uses solve_f32.
#pragma once

#include "ctr_fk_core_f32.cuh"

namespace ctrk {

struct F2 { float x, y; };

CTR_HD F2 f2(float a) { F2 r; r.x = a; r.y = a; return r; }
CTR_HD F2 f2(float a, float b) { F2 r; r.x = a; r.y = b; return r; }

#ifdef __CUDA_ARCH__
__device__ __forceinline__ unsigned long long f2_pack(F2 a)
{
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a.x), "f"(a.y));
    return r;
}
__device__ __forceinline__ F2 f2_unpack(unsigned long long r)
{
    F2 a;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a.x), "=f"(a.y) : "l"(r));
    return a;
}
#endif
CTR_HD F2 f2_fma(F2 a, F2 b, F2 c)
{
#ifdef __CUDA_ARCH__
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(f2_pack(a)), "l"(f2_pack(b)), "l"(f2_pack(c)));
    return f2_unpack(d);
#else
    return f2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y));
#endif
}
CTR_HD F2 f2_mul(F2 a, F2 b)
{
#ifdef __CUDA_ARCH__
    unsigned long long d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(f2_pack(a)), "l"(f2_pack(b)));
    return f2_unpack(d);
#else
    return f2(a.x * b.x, a.y * b.y);
#endif
}
CTR_HD F2 f2_add(F2 a, F2 b)
{
#ifdef __CUDA_ARCH__
    unsigned long long d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(f2_pack(a)), "l"(f2_pack(b)));
    return f2_unpack(d);
#else
    return f2(a.x + b.x, a.y + b.y);
#endif
}
CTR_HD F2 f2_sub(F2 a, F2 b)
{
#ifdef __CUDA_ARCH__
    unsigned long long d;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(f2_pack(a)), "l"(f2_pack(b)));
    return f2_unpack(d);
#else
    return f2(a.x - b.x, a.y - b.y);
#endif
}
CTR_HD F2 f2_neg(F2 a) { return f2(-a.x, -a.y); }      // folds into the operand modifier of the consumer
CTR_HD F2 f2_sel(bool px, bool py, F2 a, F2 b) { return f2(px ? a.x : b.x, py ? a.y : b.y); }
CTR_HD int f2_absmax_bits(F2 a)                          // max(|x|, |y|) as an ordered bit pattern
{
#ifdef __CUDA_ARCH__
    const int bx = __float_as_int(a.x) & 0x7fffffff, by = __float_as_int(a.y) & 0x7fffffff;
#else
    int bx, by;
    memcpy(&bx, &a.x, 4); memcpy(&by, &a.y, 4);
    bx &= 0x7fffffff; by &= 0x7fffffff;
#endif
    return bx > by ? bx : by;
}

// Section masks of one configuration at arc position pos, and the largest interval bound the
// position has not passed yet (the next place where a mask can change).
template <class L>
struct MaskState {
    bool   in_k[L::S], in_c[L::S];
    double next;           // masks are constant while pos > next
};
template <class L>
CTR_HD void eval_masks(const Ctl<L>& c, double pos, MaskState<L>& m)
{
    constexpr int S = L::S;
    double nx = -1.0;      // below every position: no further change
#pragma unroll
    for (int s = 0; s < S; ++s) {
        m.in_k[s] = (s == S - 1) ? true : le_nonneg(pos, c.end[s]);       // pos <= End_s
        m.in_c[s] = m.in_k[s] && le_nonneg(c.start[s], pos);               // Start_s <= pos
        if (s < S - 1 && !m.in_k[s]) nx = (c.end[s] > nx) ? c.end[s] : nx; // becomes true at pos <= End_s
        if (le_nonneg(c.start[s], pos)) nx = (c.start[s] > nx) ? c.start[s] : nx;   // becomes false at pos < Start_s
    }
    m.next = nx;
}

// One pair of configurations (A in .x, B in .y), same n.  Software-pipelined like the FP64 FAST
// solver: iteration k accumulates the pending sample k+1 next to the arithmetic of sample k, and
// takes a branch-free HOT body when everything qualifies (carried sin/cos can be rotated, series
// SE(3) step, tube 0 pinned); the head step, every 128th step (sin/cos re-evaluated) and coarse
// sampling take the generic body.
template <class L>
struct F32x2Sweep {
    static constexpr int S = L::S, T = L::T, U = (T > 1 ? T - 1 : 1);
    const KRobot& rb;
    const Ctl<L>& cA;
    const Ctl<L>& cB;
    int   n;
    F2    al[T], uzs[T], sn[T], cs[T];
    F2    h, hh, hh2, h3;
    float ztol, ztol2, ncl;
    float kg[S], rk[S], czc[U], hkA[U], hkB[U];
    F2    Qw, Qx, Qy, Qz, Px, Py, Pz, a_tip;
    F2    pux, puy, puz;             // pending sample
    double posA, posB;
    MaskState<L> mA, mB;
    F2    g[S], rks, hk[U];

    CTR_HD F32x2Sweep(const KRobot& rb_, const Ctl<L>& a, const Ctl<L>& b) : rb(rb_), cA(a), cB(b)
    {
        n = cA.n;
#pragma unroll
        for (int t = 0; t < T; ++t) {
            al[t] = f2((float)cA.al[t], (float)cB.al[t]);
            uzs[t] = f2(0.f);
            sincos_f32(al[t].x, sn[t].x, cs[t].x);
            sincos_f32(al[t].y, sn[t].y, cs[t].y);
        }
        h = f2((float)cA.h, (float)cB.h);
        hh = f2_mul(f2(0.5f), h); hh2 = f2_mul(hh, hh); h3 = f2_mul(f2_mul(h, h), h);
        ztol = (float)rb.ztol; ztol2 = ztol * ztol;         // see solve_f32 on the tolerance
#pragma unroll
        for (int s = 0; s < S; ++s) { kg[s] = (float)rb.kg[s]; rk[s] = (float)rb.rk[s]; }
#pragma unroll
        for (int t = 0; t + 1 < T; ++t) {
            czc[t] = (float)rb.czc[t];
            hkA[t] = (float)(cA.h * rb.onu[t] * rb.kappa[L::sec_of(t)]);
            hkB[t] = (float)(cB.h * rb.onu[t] * rb.kappa[L::sec_of(t)]);
        }
        ncl = -(float)rb.cl;
        Qw = f2(1.f); Qx = Qy = Qz = f2(0.f); Px = Py = Pz = f2(0.f); a_tip = f2(0.f);
        pux = puy = puz = f2(0.f);
        posA = cA.arc_end; posB = cB.arc_end;
    }

    // packed per-section constants from the masks at the current arc positions
    CTR_HD void refresh_constants()
    {
        eval_masks<L>(cA, posA, mA);
        eval_masks<L>(cB, posB, mB);
        float ra = rk[S - 1], rb_ = rk[S - 1];
#pragma unroll
        for (int s = S - 2; s >= 0; --s) { ra = mA.in_k[s] ? rk[s] : ra; rb_ = mB.in_k[s] ? rk[s] : rb_; }
        rks = f2(ra, rb_);
#pragma unroll
        for (int s = 0; s < S; ++s) g[s] = f2(mA.in_c[s] ? kg[s] : 0.f, mB.in_c[s] ? kg[s] : 0.f);
#pragma unroll
        for (int t = 0; t + 1 < T; ++t)
            hk[t] = f2(mA.in_c[L::sec_of(t)] ? hkA[t] : 0.f, mB.in_c[L::sec_of(t)] ? hkB[t] : 0.f);
    }

    // P <- E P
    CTR_HD void apply(F2 ew, F2 ex, F2 ey, F2 ez, F2 pex, F2 pey, F2 pez)
    {
        const F2 tx = f2_fma(ey, Pz, f2_neg(f2_mul(ez, Py)));
        const F2 ty = f2_fma(ez, Px, f2_neg(f2_mul(ex, Pz)));
        const F2 tz = f2_fma(ex, Py, f2_neg(f2_mul(ey, Px)));
        const F2 ix = f2_fma(ew, tx, f2_fma(ey, tz, f2_neg(f2_mul(ez, ty))));
        const F2 iy = f2_fma(ew, ty, f2_fma(ez, tx, f2_neg(f2_mul(ex, tz))));
        const F2 iz = f2_fma(ew, tz, f2_fma(ex, ty, f2_neg(f2_mul(ey, tx))));
        Px = f2_fma(f2(2.f), ix, f2_add(Px, pex));
        Py = f2_fma(f2(2.f), iy, f2_add(Py, pey));
        Pz = f2_fma(f2(2.f), iz, f2_add(Pz, pez));
        const F2 w2 = f2_fma(f2_neg(ez), Qz, f2_fma(f2_neg(ey), Qy, f2_fma(f2_neg(ex), Qx, f2_mul(ew, Qw))));
        const F2 x2 = f2_fma(f2_neg(ez), Qy, f2_fma(ey, Qz, f2_fma(ex, Qw, f2_mul(ew, Qx))));
        const F2 y2 = f2_fma(f2_neg(ex), Qz, f2_fma(ez, Qx, f2_fma(ey, Qw, f2_mul(ew, Qy))));
        const F2 z2 = f2_fma(f2_neg(ey), Qx, f2_fma(ex, Qy, f2_fma(ez, Qw, f2_mul(ew, Qz))));
        Qw = w2; Qx = x2; Qy = y2; Qz = z2;
    }
    // series form of the pending step (both m <= 2^-6, no zero curvature)
    CTR_HD void accumulate_series(F2 sxy, F2 m)
    {
        const F2 qa = f2_fma(f2_fma(f2_fma(f2(-1.f / 720.f), m, f2(1.f / 24.f)), m, f2(-0.5f)), m, f2(1.f));
        const F2 qb = f2_fma(f2_fma(f2_fma(f2(-1.f / 5040.f), m, f2(1.f / 120.f)), m, f2(-1.f / 6.f)), m, f2(1.f));
        const F2 qc = f2_fma(f2_fma(f2(16.f / 5040.f), m, f2(-4.f / 120.f)), m, f2(1.f / 6.f));
        const F2 b  = f2_mul(hh, qb), b2 = f2_mul(h, qb);
        const F2 ex = f2_mul(b, pux), ey = f2_mul(b, puy), ez = f2_mul(b, puz);
        const F2 C = f2_mul(h3, qc), czz = f2_mul(C, puz);
        apply(qa, ex, ey, ez, f2_fma(czz, pux, f2_mul(b2, ey)), f2_fma(czz, puy, f2_neg(f2_mul(b2, ex))),
              f2_fma(f2_neg(C), sxy, h));
    }
    CTR_HD bool series_ok(F2 n2, F2 m) const
    {
        const float n2min = n2.x < n2.y ? n2.x : n2.y;
        return f2_absmax_bits(m) <= 0x3c800000 && n2min >= ztol2;
    }
    CTR_HD void accumulate_generic(F2 sxy, F2 n2, F2 m)
    {
        if (series_ok(n2, m)) { accumulate_series(sxy, m); return; }
        F2 ew, ex, ey, ez, pex, pey, pez;
        se3_step_f32(pux.x, puy.x, puz.x, h.x, ztol, ew.x, ex.x, ey.x, ez.x, pex.x, pey.x, pez.x);
        se3_step_f32(pux.y, puy.y, puz.y, h.y, ztol, ew.y, ex.y, ey.y, ez.y, pex.y, pey.y, pez.y);
        apply(ew, ex, ey, ez, pex, pey, pez);
    }

    // One sample.  HOT: 1 <= k < n-1, every |d_t| < 2^-4, k not a multiple of 128, series_ok(pending).
    template <bool HOT>
    CTR_HD void body(int k, const F2* d, F2 sxy_p, F2 n2_p, F2 m_p)
    {
        const bool head = (k == n - 1);
        // ---- (A) pending sample k+1 ---------------------------------------------------------------
        if (HOT)        accumulate_series(sxy_p, m_p);
        else if (!head) accumulate_generic(sxy_p, n2_p, m_p);
        // ---- (B) this sample: curvature -------------------------------------------------------------
        F2 cx = f2(0.f), cy = f2(0.f);
#pragma unroll
        for (int s = 0; s < S; ++s) {
#pragma unroll
            for (int j = 0; j < L::nt(s); ++j) {
                const int t = L::t0(s) + j;
                if (HOT && t == 0) { cy = g[s]; continue; }          // tube 0 pinned: sin 0 = 0, cos 0 = 1
                cx = f2_fma(f2_neg(sn[t]), g[s], cx);
                cy = f2_fma(cs[t], g[s], cy);
            }
        }
        cx = f2_mul(cx, rks); cy = f2_mul(cy, rks);
        F2 uxs[T];
#pragma unroll
        for (int t = 0; t < T; ++t) uxs[t] = (HOT && t == 0) ? cx : f2_fma(cs[t], cx, f2_mul(sn[t], cy));
        F2 uy_in;
        if (HOT && T == 1) uy_in = cy;
        else               uy_in = f2_fma(cs[T - 1], cy, f2_neg(f2_mul(sn[T - 1], cx)));
        const F2 uz_k = uzs[T - 1];
        F2 a_k = al[T - 1];
        if (HOT || k >= 1) {
            // ---- tube angles, torsion -------------------------------------------------------------
            al[0] = f2(0.f); sn[0] = f2(0.f); cs[0] = f2(1.f);       // section_i.h:308
#pragma unroll
            for (int t = 1; t < T; ++t) al[t] = f2_add(al[t], d[t]);
            a_k = al[T - 1];
            F2 cz = f2(0.f);
#pragma unroll
            for (int t = T - 2; t >= 0; --t) {
                cz     = f2_fma(f2(czc[t]), uzs[t], cz);
                uzs[t] = f2_fma(hk[t], uxs[t], uzs[t]);
            }
            uzs[T - 1] = f2_mul(cz, f2(ncl));
            // ---- sin/cos of the new angles ------------------------------------------------------------
#pragma unroll
            for (int t = 1; t < T; ++t) {
                if (HOT) {
                    const F2 d2 = f2_mul(d[t], d[t]);
                    const F2 cd = f2_fma(f2_fma(f2(1.f / 24.f), d2, f2(-0.5f)), d2, f2(1.f));
                    const F2 sp = f2_fma(f2_fma(f2(1.f / 120.f), d2, f2(-1.f / 6.f)), d2, f2(1.f));
                    const F2 sd = f2_mul(d[t], sp);
                    const F2 s1 = f2_fma(cs[t], sd, f2_mul(sn[t], cd));
                    const F2 c1 = f2_fma(f2_neg(sn[t]), sd, f2_mul(cs[t], cd));
                    sn[t] = s1; cs[t] = c1;
                } else {
                    sincos_f32(al[t].x, sn[t].x, cs[t].x);
                    sincos_f32(al[t].y, sn[t].y, cs[t].y);
                }
            }
        }
        pux = uxs[T - 1]; puy = uy_in; puz = uz_k;
        if (!HOT && head) a_tip = a_k;
    }

    CTR_HD void run()
    {
        bool refresh = true;
#pragma unroll 1
        for (int k = n - 1; k >= 0; --k) {
            if (refresh) refresh_constants();
            F2  d[T];
            int dmax = 0;
            d[0] = f2(0.f);
            const F2 z0 = uzs[0];
#pragma unroll
            for (int t = 1; t < T; ++t) {
                d[t] = f2_mul(h, f2_sub(uzs[t], z0));
                const int b = f2_absmax_bits(d[t]);
                dmax = b > dmax ? b : dmax;
            }
            const F2 sxy = f2_fma(puy, puy, f2_mul(pux, pux));
            const F2 n2  = f2_fma(puz, puz, sxy);
            const F2 m   = f2_mul(n2, hh2);
            const bool hot = k >= 1 && k != n - 1 && dmax < 0x3d800000 && (k & 127) != 0 && series_ok(n2, m);
            if (hot) body<true>(k, d, sxy, n2, m);
            else     body<false>(k, d, sxy, n2, m);
            if (k >= 1) {
                posA = x_sub(posA, cA.h);
                posB = x_sub(posB, cB.h);
                refresh = le_nonneg(posA, mA.next) || le_nonneg(posB, mB.next);
            }
        }
        {   // the last pending sample (k = 0)
            const F2 sxy = f2_fma(puy, puy, f2_mul(pux, pux));
            const F2 n2  = f2_fma(puz, puz, sxy);
            accumulate_generic(sxy, n2, f2_mul(n2, hh2));
        }
    }
};

// tipA / tipB: nullable (an invalid or absent configuration runs as a copy of its partner and is
// not stored).
template <class L>
CTR_HD void solve_f32x2(const KRobot& rb, const Ctl<L>& cA, const Ctl<L>& cB, double* tipA, double* tipB)
{
    F32x2Sweep<L> sw(rb, cA, cB);
    sw.run();
    const F2 Qw = sw.Qw, Qx = sw.Qx, Qy = sw.Qy, Qz = sw.Qz, Px = sw.Px, Py = sw.Py, Pz = sw.Pz, a_tip = sw.a_tip;
    // finish in double: the entry frame and theta are double inputs
#pragma unroll
    for (int e = 0; e < 2; ++e) {
        double* tip = e ? tipB : tipA;
        if (!tip) continue;
        const Ctl<L>& c = e ? cB : cA;
        const double qw = e ? Qw.y : Qw.x, qx = e ? Qx.y : Qx.x, qy = e ? Qy.y : Qy.x, qz = e ? Qz.y : Qz.x;
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
        Pm[9] = e ? Px.y : Px.x; Pm[10] = e ? Py.y : Py.x; Pm[11] = e ? Pz.y : Pz.x;
        double sa, ca;
        fast_sincos((double)(e ? a_tip.y : a_tip.x) + c.theta, sa, ca);
        tf_finish(rb.ini, Pm, sa, ca, tip);
    }
}

}  // namespace ctrk
