/* ctr_oracle.c — CPU oracle: plain-C restatement of the reference FK.  TEST INFRASTRUCTURE ONLY
 * (see ctr_oracle.h: pinned bit for bit against the reference's own sources built with stand-in
 * third-party headers, oracle/_ref; Erl's semantics remain assumed).
 *
 * Build with -ffp-contract=off: the reference is built by CMake without -march (CMakeLists.txt:6,
 * :36-39), i.e. baseline x86-64 without FMA, so every a*b+c below rounds twice.
 *
 * Each block cites the reference lines (under /root/reference/ctr_source) it follows.  The
 * structure is data-driven (runtime loops over sections/tubes of a POD descriptor) where the
 * reference is template-unrolled, but operation order inside every expression and the update
 * order of the sweep are preserved, because they decide the low-order bits.
 */
#include "ctr_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MAXS CTR_FK_MAX_SECTIONS
#define MAXT CTR_FK_MAX_TUBES

typedef struct {
    int    S, T;
    int    n[MAXS];        /* tubes in section                               */
    int    t0[MAXS];       /* index of the section's first tube              */
    double ksum[MAXS];     /* m_Stiffness.static_sum(), matrix.h:178-188     */
    double kappa[MAXS];    /* m_Curvature[0]                                 */
    double len[MAXS];
    double rad[MAXS];      /* getOuterRadius(), section_i.h:151              */
    double nu[MAXS][2];
    double ini[12];        /* column-major 3x4                               */
    int    maxs;
    double ztol;
} model_t;

/* per-configuration sweep state (what the reference keeps in RobotI_ members, robot_i.h:1336-1347) */
typedef struct {
    double theta;
    double al[MAXT];       /* m_SampleTubeAlpha                              */
    double uc[3 * MAXT];   /* m_SampleCurvature: (u_x,u_y,u_z) per tube      */
    double end[MAXS];      /* m_CurvedArcEnd                                 */
    double start[MAXS];    /* m_CurvedArcStart                               */
} state_t;

static int build_model(const ctr_robot_desc* d, model_t* m)
{
    if (!d) return CTR_FK_E_NULL;
    if (d->nsections < 1 || d->nsections > MAXS || d->max_samples < 1) return CTR_FK_E_DESC;
    m->S = d->nsections;
    m->T = 0;
    for (int s = 0; s < m->S; ++s) {
        const ctr_section_desc* sd = &d->sec[s];
        if (sd->ntube < 1 || sd->ntube > 2) return CTR_FK_E_DESC;
        m->n[s]  = sd->ntube;
        m->t0[s] = m->T;
        m->T += sd->ntube;
        double sum = sd->stiffness[0];                 /* left-to-right, matrix.h:181-186 */
        for (int k = 1; k < sd->ntube; ++k) sum += sd->stiffness[k];
        m->ksum[s]  = sum;
        m->kappa[s] = sd->curvature[0];
        m->len[s]   = sd->length;
        m->rad[s]   = sd->radius[0];
        m->nu[s][0] = sd->poisson[0];
        m->nu[s][1] = sd->poisson[1];
    }
    memcpy(m->ini, d->init_trafo, sizeof m->ini);
    m->maxs = d->max_samples;
    m->ztol = (d->zero_tol > 0.0) ? d->zero_tol : 1e-12;
    return 0;
}

/* ---- 3x4 rigid transforms, column-major (Erl::Transform semantics) ---- */

static void tf_identity(double* a)
{
    memset(a, 0, 12 * sizeof(double));
    a[0] = a[4] = a[8] = 1.0;
}

/* c = a * b : R = Ra Rb, p = Ra pb + pa (robot_i.h:712,714) */
static void tf_mul(const double* a, const double* b, double* c)
{
    double r[12];
    for (int j = 0; j < 3; ++j)
        for (int i = 0; i < 3; ++i)
            r[3 * j + i] = a[i] * b[3 * j] + a[3 + i] * b[3 * j + 1] + a[6 + i] * b[3 * j + 2];
    for (int i = 0; i < 3; ++i)
        r[9 + i] = (a[i] * b[9] + a[3 + i] * b[10] + a[6 + i] * b[11]) + a[9 + i];
    memcpy(c, r, sizeof r);
}

/* robot_i.h:1241-1269 — SE(3) exponential of one constant-curvature step */
static void tf_step(const double* u, double h, double ztol, double* e)
{
    double nrm = sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
    if (!(fabs(nrm - 0.0) < ztol)) {                   /* !Erl::equal(norm, 0), :1245 */
        double wx = u[0] / nrm, wy = u[1] / nrm, wz = u[2] / nrm;
        double sn = sin(h * nrm), cn = cos(h * nrm);
        double vm = cn - 1.0;
        /* Rodrigues entries, :1252-1255 (constructor arguments are row-major) */
        double r00 = (wy * wy + wz * wz) * vm + 1.0;
        double r01 = -sn * wz - wx * wy * vm;
        double r02 = sn * wy - wx * wz * vm;
        double r10 = sn * wz - wx * wy * vm;
        double r11 = (wx * wx + wz * wz) * vm + 1.0;
        double r12 = -sn * wx - wy * wz * vm;
        double r20 = -sn * wy - wx * wz * vm;
        double r21 = sn * wx - wy * wz * vm;
        double r22 = (wx * wx + wy * wy) * vm + 1.0;
        /* arc translation, :1257-1259 (W.data() is column-major: [0]=r00 [1]=r10 [2]=r20
           [3]=r01 [4]=r11 [5]=r21) */
        double px = h * wx * wz + (-(wy * (r00 - 1.0)) + (r01 * wx)) / nrm;
        double py = h * wy * wz + ((wx * (r11 - 1.0)) - (r10 * wy)) / nrm;
        double pz = h * wz * wz + ((r21 * wx) - (r20 * wy)) / nrm;
        e[0] = r00; e[1] = r10; e[2] = r20;
        e[3] = r01; e[4] = r11; e[5] = r21;
        e[6] = r02; e[7] = r12; e[8] = r22;
        e[9] = px;  e[10] = py; e[11] = pz;
        return;
    }
    tf_identity(e);                                    /* :1264-1266: translation z = h*norm */
    e[11] = h * nrm;
}

/* robot_i.h:1280-1285 — Rz(angle) * t given sin/cos of the angle */
static void tf_rotz_left(const double* t, double sn, double cn, double* o)
{
    double r[12];
    for (int j = 0; j < 4; ++j) {
        r[3 * j]     = t[3 * j] * cn - t[3 * j + 1] * sn;
        r[3 * j + 1] = t[3 * j + 1] * cn + t[3 * j] * sn;
        r[3 * j + 2] = t[3 * j + 2];
    }
    memcpy(o, r, sizeof r);
}

/* ---- joints -> section arc intervals: robot_i.h:1123-1138, section_i.h:201-213 ---- */
static int set_joints(const model_t* m, const double* jv, state_t* st, double* arc_end)
{
    int bad = 0;
    memset(st->uc, 0, sizeof st->uc);                  /* memset, robot_i.h:1076 */
    memset(st->al, 0, sizeof st->al);
    st->theta = jv[0];
    double run = 0.0;
    for (int s = 0; s < m->S; ++s) {
        double phi = jv[1 + s + m->t0[s]];
        if (!((phi >= 0.0) && (phi <= m->len[s]))) bad = 1;   /* assert, section_i.h:205 */
        for (int k = 0; k < m->n[s]; ++k) st->al[m->t0[s] + k] = jv[2 + s + m->t0[s] + k];
        run += phi;
        st->end[s] = run;
        double d = run - m->len[s];
        st->start[s] = (0.0 < d) ? d : 0.0;            /* std::max<Real>(0, End-Length) */
    }
    *arc_end = run;
    return bad;
}

/* One backward sample: robot_i.h:1139-1201 (with_update=1) / :1203-1238 (with_update=0). */
static void sweep_sample(const model_t* m, state_t* st, double pos, double h, int idx,
                         int with_update, double* curv, double* alpha_samp)
{
    double sn[MAXT], cs[MAXT], kap[MAXS], kst[MAXS];
    const int T = m->T, S = m->S;

    for (int t = 0; t < T; ++t) { sn[t] = sin(st->al[t]); cs[t] = cos(st->al[t]); }  /* :1147-1148 */

    for (int s = 0; s < S; ++s) {                      /* section_i.h:322-326 */
        kap[s] = (double)((pos <= st->end[s]) && (pos >= st->start[s])) * m->kappa[s];
        kst[s] = (double)(pos <= st->end[s]) * m->ksum[s];
    }
    double cx = 0.0, cy = 0.0, ksum = 0.0;             /* reset per sample, robot_i.h:1116-1118 */
    for (int s = 0; s < S; ++s) {                      /* section_i.h:215-226, :187-191 */
        ksum += kst[s];
        double share = kst[s] / (double)m->n[s];
        for (int k = 0; k < m->n[s]; ++k) {
            int t = m->t0[s] + k;
            cx += -sn[t] * kap[s] * share;
            cy += cs[t] * kap[s] * share;
        }
    }
    for (int t = 0; t < T; ++t) {                      /* section_i.h:192-198, :228-242 */
        st->uc[3 * t]     = (cs[t] * cx + sn[t] * cy) / ksum;
        st->uc[3 * t + 1] = (-sn[t] * cx + cs[t] * cy) / ksum;
    }
    curv[3 * idx]     = st->uc[3 * (T - 1)];           /* robot_i.h:1170-1171 / :1235-1236 */
    curv[3 * idx + 1] = st->uc[3 * (T - 1) + 1];

    if (!with_update) {                                /* :1237 */
        alpha_samp[idx] = st->al[T - 1];
        return;
    }

    /* tube angles, explicit step with the PREVIOUS sample's u_z: robot_i.h:1174-1181 */
    st->al[0] = 0.0;                                   /* section_i.h:308 */
    if (m->n[0] == 2)                                  /* section_i.h:309-311 */
        st->al[1] = st->al[1] + h * (st->uc[3 * 1 + 2] - st->uc[2]);
    for (int s = 1; s < S; ++s)                        /* section_i.h:289-300 */
        for (int k = m->n[s] - 1; k >= 0; --k) {
            int t = m->t0[s] + k;
            st->al[t] = st->al[t] + h * (st->uc[3 * t + 2] - st->uc[2]);
        }
    alpha_samp[idx] = st->al[T - 1];

    /* torsion: robot_i.h:1184-1196 */
    double cz = 0.0;
    const int sl = S - 1, tl = m->t0[sl];
    if (m->n[sl] == 2) {                               /* section_i.h:265-278: first tube of a VC
                                                          last section, Poisson ratio of tube [1] */
        double onu = 1.0 + m->nu[sl][1];
        cz += (kst[sl] / 2.0) / onu * st->uc[3 * tl + 2];
        st->uc[3 * tl + 2] += h * onu * st->uc[3 * tl] * kap[sl];
    }
    for (int s = S - 2; s >= 0; --s)                   /* section_i.h:253-264, :243-250 */
        for (int k = m->n[s] - 1; k >= 0; --k) {
            int t = m->t0[s] + k;
            double onu = 1.0 + m->nu[s][k];
            cz += (kst[s] / (double)m->n[s]) / onu * st->uc[3 * t + 2];
            st->uc[3 * t + 2] += h * onu * st->uc[3 * t] * kap[s];
        }
    /* innermost tube from the torque balance: section_i.h:280-286 */
    st->uc[3 * (T - 1) + 2] = (-cz) * (1.0 + m->nu[sl][m->n[sl] - 1]) / kst[sl];
    curv[3 * idx - 1] = st->uc[3 * (T - 1) + 2];       /* lagged z slot, robot_i.h:1199 */
}

/* Tip-to-base sweep: robot_i.h:1101-1122.  Returns final (pos, idx) through pointers. */
static void sweep(const model_t* m, state_t* st, double h, double* pos, long* idx, double* curv,
                  double* alpha_samp)
{
    curv[3 * (*idx) + 2] = 0.0;                        /* tip torsion boundary condition, :1108 */
    for (; (*pos >= h) && (*idx >= 1); *pos -= h, --(*idx))
        sweep_sample(m, st, *pos, h, (int)*idx, 1, curv, alpha_samp);
    sweep_sample(m, st, *pos, h, (int)*idx, 0, curv, alpha_samp);
}

/* robot_i.h:1286-1307.  h == 0 (all Phi = 0 in NSamples mode) makes End/h NaN or inf, whose
 * size_t conversion is undefined in the reference; here it saturates to `count`. */
static void fill_radius(const model_t* m, const state_t* st, double h, long count, double* radius)
{
    long  begin = 0;
    double rsec = 0.0;
    for (int s = 0; s < m->S; ++s) {
        if (begin >= count) continue;
        rsec = m->rad[s];
        double q = st->end[s] / h;
        long stop = (q >= 0.0 && q < (double)count) ? (long)q : count;
        if (stop < begin) stop = begin;
        for (long k = begin; k < stop; ++k) radius[k] = rsec;
        begin = stop;
    }
    radius[count - 1] = rsec;
}

typedef struct {
    double* frames;      /* [maxs][12] */
    double* radius;      /* [maxs]     */
    double* curv;        /* [3*maxs]   */
    double* alpha;       /* [maxs]     */
} scratch_t;

static int scratch_alloc(scratch_t* w, int maxs)
{
    w->frames = (double*)malloc(sizeof(double) * 12 * (size_t)maxs);
    w->radius = (double*)malloc(sizeof(double) * (size_t)maxs);
    w->curv   = (double*)malloc(sizeof(double) * 3 * (size_t)maxs + 8);
    w->alpha  = (double*)malloc(sizeof(double) * (size_t)maxs);
    return (w->frames && w->radius && w->curv && w->alpha) ? 0 : CTR_FK_E_NOMEM;
}
static void scratch_free(scratch_t* w)
{
    free(w->frames); free(w->radius); free(w->curv); free(w->alpha);
}

/* calcKinematic: robot_i.h:690-721 (step mode) and :722-755 (NSamples mode). */
static int fk_one(const model_t* m, scratch_t* w, const double* jv, int mode, double step,
                  int32_t nsamples, double* tip12, int32_t* n_out, double* step_out,
                  double* alpha_base)
{
    state_t st;
    double arc_end;
    int bad = set_joints(m, jv, &st, &arc_end);
    if (bad) {
        if (tip12) for (int i = 0; i < 12; ++i) tip12[i] = NAN;
        if (n_out) *n_out = 0;
        if (step_out) *step_out = NAN;
        if (alpha_base) for (int t = 0; t < m->T; ++t) alpha_base[t] = NAN;
        return CTR_FK_STATUS_PHI_RANGE;
    }

    long   count;
    double h = step;
    if (mode == CTR_FK_MODE_STEP) {                    /* robot_i.h:1079-1081 */
        double q = arc_end / h;
        count = (long)q;
        if (count < 1) count = 1;
        if (count > m->maxs) { count = m->maxs; h = arc_end / (double)count; }
    } else {                                           /* robot_i.h:1093-1094 */
        count = (nsamples < m->maxs) ? nsamples : m->maxs;
        h = arc_end / (double)count;
    }
    double pos = arc_end;                              /* :1083-1084 / :1096-1097 */
    long   idx = count - 1;
    sweep(m, &st, h, &pos, &idx, w->curv, w->alpha);

    double inter[12], e[12], rz[12];
    tf_identity(inter);                                /* :699 / :729 */
    /* The reference fills sin/cos(Alpha+theta) for the first SampleMax samples up front
       (:700-703); when step mode caps N that count can round to N-1 and the tip sample reads
       stale memory.  Not reproducible: here every sample gets its own. */
    for (;;) {
        if (mode == CTR_FK_MODE_STEP) { if (!((pos <= arc_end) && (idx < m->maxs))) break; }
        else                          { if (!((idx < nsamples) && (idx < m->maxs))) break; }
        tf_step(&w->curv[3 * idx], h, m->ztol, e);     /* :712 / :743 */
        tf_mul(inter, e, inter);
        double a = w->alpha[idx] + st.theta;
        tf_rotz_left(inter, sin(a), cos(a), rz);       /* :714-716 */
        tf_mul(m->ini, rz, &w->frames[12 * idx]);
        if (mode == CTR_FK_MODE_STEP) pos += h;
        ++idx;
    }
    fill_radius(m, &st, h, idx, w->radius);            /* :718 */
    if (n_out) *n_out = (int32_t)idx;                  /* m_Samples, :719 */
    if (step_out) *step_out = h;
    if (tip12) memcpy(tip12, &w->frames[12 * (idx - 1)], 12 * sizeof(double));   /* :720 */
    if (alpha_base) memcpy(alpha_base, st.al, sizeof(double) * (size_t)m->T);
    return CTR_FK_STATUS_OK;
}

static int check_args(const ctr_robot_desc* d, const double* jv, int mode, double step,
                      int32_t nsamples)
{
    if (!d || !jv) return CTR_FK_E_NULL;
    if (mode == CTR_FK_MODE_STEP) { if (!(step > 0.0) || !isfinite(step)) return CTR_FK_E_ARG; }
    else if (mode == CTR_FK_MODE_NSAMPLE) { if (nsamples < 1) return CTR_FK_E_ARG; }
    else return CTR_FK_E_ARG;
    return 0;
}

int ctr_oracle_fk(const ctr_robot_desc* desc, const double* jv, int mode, double step,
                  int32_t nsamples, double* tip12, double* backbone, double* radius,
                  int32_t* n_out, double* step_out, double* curv, double* alpha_samp,
                  double* alpha_base)
{
    int rc = check_args(desc, jv, mode, step, nsamples);
    if (rc) return rc;
    model_t m;
    if ((rc = build_model(desc, &m))) return rc;
    scratch_t w;
    if ((rc = scratch_alloc(&w, m.maxs))) { scratch_free(&w); return rc; }
    int32_t n = 0;
    int st = fk_one(&m, &w, jv, mode, step, nsamples, tip12, &n, step_out, alpha_base);
    if (n_out) *n_out = n;
    if (backbone) memcpy(backbone, w.frames, sizeof(double) * 12 * (size_t)n);
    if (radius) memcpy(radius, w.radius, sizeof(double) * (size_t)n);
    if (curv) memcpy(curv, w.curv, sizeof(double) * 3 * (size_t)n);
    if (alpha_samp) memcpy(alpha_samp, w.alpha, sizeof(double) * (size_t)n);
    scratch_free(&w);
    return st;
}

int ctr_oracle_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

int ctr_oracle_fk_batch(const ctr_robot_desc* desc, const double* jv, int64_t B, int mode,
                        double step, int32_t nsamples, int nthreads, double* tip,
                        double* backbone, double* radius, int32_t* n_out, double* step_out,
                        double* alpha_base, int32_t* status)
{
    int rc = check_args(desc, jv, mode, step, nsamples);
    if (rc) return rc;
    if (B < 0) return CTR_FK_E_ARG;
    model_t m;
    if ((rc = build_model(desc, &m))) return rc;
    const int njv = 1 + m.S + m.T;
    if (nthreads < 1) nthreads = 1;
    int fail = 0;
#ifdef _OPENMP
#pragma omp parallel num_threads(nthreads)
#endif
    {
        scratch_t w;
        int ok = (scratch_alloc(&w, m.maxs) == 0);
        if (!ok) {
#ifdef _OPENMP
#pragma omp atomic write
#endif
            fail = 1;
        }
#ifdef _OPENMP
#pragma omp for schedule(static)
#endif
        for (int64_t i = 0; i < B; ++i) {
            if (!ok) continue;
            int32_t n = 0;
            int st = fk_one(&m, &w, jv + i * njv, mode, step, nsamples,
                            tip ? tip + 12 * i : NULL, &n, step_out ? step_out + i : NULL,
                            alpha_base ? alpha_base + (int64_t)m.T * i : NULL);
            if (n_out) n_out[i] = n;
            if (status) status[i] = st;
            if (backbone) memcpy(backbone + (int64_t)12 * m.maxs * i, w.frames, sizeof(double) * 12 * (size_t)n);
            if (radius) memcpy(radius + (int64_t)m.maxs * i, w.radius, sizeof(double) * (size_t)n);
        }
        scratch_free(&w);
    }
    return fail ? CTR_FK_E_NOMEM : 0;
}

/* ---- finite-difference Jacobian: robot_i.h:41-56 and :917-960 ---------------------------- */

/* column = d(lo -> hi)/dq: translation difference and the vee of (R_hi - R_lo) R_hi^T */
static void jac_column(const double* lo, const double* hi, double dq, double* col)
{
    double d[9], mprod[9];
    for (int i = 0; i < 9; ++i) d[i] = hi[i] - lo[i];                 /* :48 */
    for (int i = 0; i < 3; ++i)                                       /* :49: D * R_hi^T */
        for (int j = 0; j < 3; ++j)
            mprod[3 * j + i] = d[i] * hi[j] + d[3 + i] * hi[3 + j] + d[6 + i] * hi[6 + j];
    col[0] = (hi[9] - lo[9]) / dq;                                    /* :50-52 */
    col[1] = (hi[10] - lo[10]) / dq;
    col[2] = (hi[11] - lo[11]) / dq;
    /* M(r,c) = mprod[3c + r] */
    col[3] = ((mprod[3 * 1 + 2] - mprod[3 * 2 + 1]) * 0.5) / dq;      /* (M(2,1)-M(1,2))/2, :53 */
    col[4] = ((mprod[3 * 2 + 0] - mprod[3 * 0 + 2]) * 0.5) / dq;      /* (M(0,2)-M(2,0))/2, :54 */
    col[5] = ((mprod[3 * 0 + 1] - mprod[3 * 1 + 0]) * 0.5) / dq;      /* (M(1,0)-M(0,1))/2, :55 */
}

/* analytic column 0, robot_i.h:956-959 (:1064-1072 for the sample): axis = z column of the entry rotation */
static void jac_column0(const model_t* m, const double* frame, double* jac)
{
    const double* z = &m->ini[6];
    double r[3] = { frame[9] - m->ini[9], frame[10] - m->ini[10], frame[11] - m->ini[11] };
    jac[0] = z[1] * r[2] - z[2] * r[1];
    jac[1] = z[2] * r[0] - z[0] * r[2];
    jac[2] = z[0] * r[1] - z[1] * r[0];
    jac[3] = z[0]; jac[4] = z[1]; jac[5] = z[2];
}

/* All calcJacobian forms of the reference:
 *   mode NSAMPLE, i_sample < 0  : calcKinematic(JV, N) + calcJacobian(JV, Tip, N, ..)            robot_i.h:917-960
 *   mode STEP,    i_sample < 0  : calcJacobian(JV, ArcLengthStep, ..) = calcKinematic(JV, step), then the
 *                                 form above with N = m_Samples of that call                     :901-907
 *   mode NSAMPLE, i_sample >= 0 : calcJacobian(JV, N, .., iSample, JTip, JiSample)               :908-915, :1012-1072
 * The perturbed solves are always NSamples-mode solves (:935, :951), also in the step form.       */
int ctr_oracle_jacobian_ex(const ctr_robot_desc* desc, const double* jv, int mode, double step, int32_t nsamples,
                           int32_t i_sample, double fd_trans, double fd_rot, double* jac_tip, double* jac_sample,
                           double* tip12, double* sample12, int32_t* n_out)
{
    int rc = check_args(desc, jv, mode, step, nsamples);
    if (rc) return rc;
    if (!jac_tip || (i_sample >= 0 && !jac_sample)) return CTR_FK_E_NULL;
    model_t m;
    if ((rc = build_model(desc, &m))) return rc;
    scratch_t w;
    if ((rc = scratch_alloc(&w, m.maxs))) { scratch_free(&w); return rc; }
    const int njv = 1 + m.S + m.T;
    double q[CTR_FK_MAX_JV], tip[12], pert[12], smp[12];
    int32_t n0 = 0;
    memcpy(q, jv, sizeof(double) * (size_t)njv);
    int st = fk_one(&m, &w, q, mode, step, nsamples, tip, &n0, NULL, NULL);
    if (st) { scratch_free(&w); return st; }
    if (i_sample >= n0) { scratch_free(&w); return CTR_FK_E_ARG; }     /* the reference would read a stale frame */
    const int32_t N = (mode == CTR_FK_MODE_STEP) ? n0 : nsamples;        /* m_Samples, :906 */
    if (i_sample >= 0) memcpy(smp, &w.frames[12 * i_sample], sizeof smp);
    memset(jac_tip, 0, sizeof(double) * 6 * (size_t)njv);                /* column 2 = 0, :928 */
    if (i_sample >= 0) memset(jac_sample, 0, sizeof(double) * 6 * (size_t)njv);
    for (int s = 0; s < m.S; ++s)                                        /* tubes 1..T-1, :929-939 */
        for (int k = 0; k < m.n[s]; ++k) {
            int t = m.t0[s] + k;
            if (t == 0) continue;
            int j = 2 + s + t;
            double keep = q[j];
            q[j] = keep + fd_rot;
            fk_one(&m, &w, q, CTR_FK_MODE_NSAMPLE, 0.0, N, pert, NULL, NULL, NULL);
            jac_column(tip, pert, fd_rot, jac_tip + 6 * j);
            if (i_sample >= 0) jac_column(smp, &w.frames[12 * i_sample], fd_rot, jac_sample + 6 * j);
            q[j] = keep;
        }
    for (int s = 0; s < m.S; ++s) {                                      /* sections, :941-954 */
        int j = m.t0[s] + s + 1;
        double keep = q[j], dq;
        q[j] = keep + fd_trans;
        if (q[j] >= m.len[s]) { q[j] = keep - fd_trans; dq = -fd_trans; }
        else dq = fd_trans;
        fk_one(&m, &w, q, CTR_FK_MODE_NSAMPLE, 0.0, N, pert, NULL, NULL, NULL);
        jac_column(tip, pert, dq, jac_tip + 6 * j);
        if (i_sample >= 0) jac_column(smp, &w.frames[12 * i_sample], dq, jac_sample + 6 * j);
        q[j] = keep;
    }
    jac_column0(&m, tip, jac_tip);
    if (i_sample >= 0) jac_column0(&m, smp, jac_sample);
    if (tip12) memcpy(tip12, tip, sizeof tip);
    if (sample12 && i_sample >= 0) memcpy(sample12, smp, sizeof smp);
    if (n_out) *n_out = n0;
    scratch_free(&w);
    return 0;
}

int ctr_oracle_jacobian(const ctr_robot_desc* desc, const double* jv, int32_t nsamples,
                        double fd_trans, double fd_rot, double* jac, double* tip12)
{
    return ctr_oracle_jacobian_ex(desc, jv, CTR_FK_MODE_NSAMPLE, 0.0, nsamples, -1, fd_trans, fd_rot, jac, NULL,
                                  tip12, NULL, NULL);
}

/* ---- calcStability: robot_i.h:756-818 ----------------------------------------------------- */

static int base_angles(const model_t* m, scratch_t* w, const double* jv, double step, double* al)
{
    state_t st;
    double arc_end;
    if (set_joints(m, jv, &st, &arc_end)) return CTR_FK_STATUS_PHI_RANGE;
    double h = step;
    long count = (long)(arc_end / h);
    if (count < 1) count = 1;
    if (count > m->maxs) { count = m->maxs; h = arc_end / (double)count; }
    double pos = arc_end;
    long idx = count - 1;
    sweep(m, &st, h, &pos, &idx, w->curv, w->alpha);
    memcpy(al, st.al, sizeof(double) * (size_t)m->T);
    return 0;
}

int ctr_oracle_stability(const ctr_robot_desc* desc, const double* jv, double angle_step,
                         double arc_step)
{
    int rc = check_args(desc, jv, CTR_FK_MODE_STEP, arc_step, 1);
    if (rc) return rc;
    if (!(angle_step > 0.0)) return CTR_FK_E_ARG;
    model_t m;
    if ((rc = build_model(desc, &m))) return rc;
    scratch_t w;
    if ((rc = scratch_alloc(&w, m.maxs))) { scratch_free(&w); return rc; }
    const int njv = 1 + m.S + m.T;
    double q[CTR_FK_MAX_JV], prev[MAXT], cur[MAXT];
    memcpy(q, jv, sizeof(double) * (size_t)njv);
    q[0] = 0.0;                                                       /* :762 */
    int stable = 1;
    for (int s = 0; s < m.S && stable; ++s)
        for (int k = 0; k < m.n[s] && stable; ++k) {
            int t = m.t0[s] + k;
            if (t == 0) continue;
            int j = 2 + s + t;
            double tipa = q[j], scan, scan_end;
            if (tipa <= M_PI) { scan = 0.0; scan_end = tipa; }         /* :782-791 */
            else              { scan = tipa; scan_end = 2.0 * M_PI; }
            q[j] = scan;                                              /* :795-799 */
            if (base_angles(&m, &w, q, arc_step, prev)) { scratch_free(&w); return CTR_FK_E_ARG; }
            scan += angle_step;
            for (; scan <= scan_end; scan += angle_step) {            /* :802-813 */
                q[j] = scan;
                base_angles(&m, &w, q, arc_step, cur);
                if ((cur[t] - prev[t]) < 0.0) { stable = 0; break; }
                memcpy(prev, cur, sizeof prev);
            }
            q[j] = jv[j];                                             /* :814 */
        }
    scratch_free(&w);
    return stable;
}

/* ---- calcDegreeStability: robot_i.h:819-899 ----------------------------------------------- */
int ctr_oracle_degree_stability(const ctr_robot_desc* desc, const double* jv, double angle_step,
                                double arc_step, double min_degree, double* degree, int* stable_out)
{
    int rc = check_args(desc, jv, CTR_FK_MODE_STEP, arc_step, 1);
    if (rc) return rc;
    if (!(angle_step > 0.0) || !degree) return CTR_FK_E_ARG;
    model_t m;
    if ((rc = build_model(desc, &m))) return rc;
    scratch_t w;
    if ((rc = scratch_alloc(&w, m.maxs))) { scratch_free(&w); return rc; }
    const int njv = 1 + m.S + m.T;
    double q[CTR_FK_MAX_JV], prev[MAXT], cur[MAXT];
    memcpy(q, jv, sizeof(double) * (size_t)njv);
    q[0] = 0.0;                                                       /* :825 */
    int stable = 1;
    int checked = (min_degree > m.ztol) ? 0 : 1;                      /* :834 */
    double smallest = 1.7976931348623157e308, result = 0.0;
    int done = 0;
    for (int s = 0; s < m.S && stable; ++s)
        for (int k = 0; k < m.n[s] && stable; ++k) {
            int t = m.t0[s] + k;
            if (t == 0) continue;
            int j = 2 + s + t;
            double tipa = q[j], scan, scan_end;
            if (tipa <= M_PI) { scan = 0.0; scan_end = tipa; }         /* :850-859 */
            else              { scan = tipa; scan_end = 2.0 * M_PI; }
            q[j] = scan;
            if (base_angles(&m, &w, q, arc_step, prev)) { scratch_free(&w); return CTR_FK_E_ARG; }
            scan += angle_step;
            for (; scan <= scan_end; scan += angle_step) {            /* :870-891 */
                q[j] = scan;
                base_angles(&m, &w, q, arc_step, cur);
                double rel = cur[t] - prev[t];
                if (rel < smallest) {
                    checked = 1;
                    smallest = rel;
                    if (rel < 0.0) { stable = 0; result = atan2(angle_step, rel); done = 1; break; }
                }
                memcpy(prev, cur, sizeof prev);
            }
            q[j] = jv[j];
        }
    if (!done) result = checked ? atan2(angle_step, smallest) : (min_degree + m.ztol);   /* :896-898 */
    *degree = result;
    if (stable_out) *stable_out = stable;
    scratch_free(&w);
    return 0;
}

static int ctr_fk_desc_njoints_local(const ctr_robot_desc* d)
{
    int t = 0;
    for (int s = 0; s < d->nsections; ++s) t += d->sec[s].ntube;
    return 1 + d->nsections + t;
}

/* ---- batch forms of the callers (OpenMP over configurations; each call builds its own model and scratch) --------- */
int ctr_oracle_jacobian_batch(const ctr_robot_desc* desc, const double* jv, int64_t B, int mode, double step,
                              int32_t nsamples, int32_t i_sample, double fd_trans, double fd_rot, int nthreads,
                              double* jac_tip, double* jac_sample, double* tip, int32_t* n_out, int32_t* rc_out)
{
    if (!desc || !jv || !jac_tip || !rc_out) return CTR_FK_E_NULL;
    const int njv = ctr_fk_desc_njoints_local(desc);
    if (nthreads < 1) nthreads = 1;
#pragma omp parallel for schedule(dynamic, 16) num_threads(nthreads)
    for (int64_t i = 0; i < B; ++i) {
        double t12[12], s12[12];
        int32_t n = 0;
        rc_out[i] = ctr_oracle_jacobian_ex(desc, jv + i * njv, mode, step, nsamples, i_sample, fd_trans, fd_rot,
                                           jac_tip + i * 6 * njv, jac_sample ? jac_sample + i * 6 * njv : NULL, t12, s12, &n);
        if (tip) memcpy(tip + 12 * i, t12, sizeof t12);
        if (n_out) n_out[i] = n;
    }
    return 0;
}

/* stable[i]: 1 / 0, < 0 = error code of that configuration; degree[i] as calcDegreeStability. */
int ctr_oracle_stability_batch(const ctr_robot_desc* desc, const double* jv, int64_t B, double angle_step,
                               double arc_step, double min_degree, int nthreads, int32_t* stable, double* degree)
{
    if (!desc || !jv || !stable || !degree) return CTR_FK_E_NULL;
    const int njv = ctr_fk_desc_njoints_local(desc);
    if (nthreads < 1) nthreads = 1;
#pragma omp parallel for schedule(dynamic, 16) num_threads(nthreads)
    for (int64_t i = 0; i < B; ++i) {
        int st = 0;
        double dg = 0.0;
        const int rc = ctr_oracle_degree_stability(desc, jv + i * njv, angle_step, arc_step, min_degree, &dg, &st);
        stable[i] = rc ? (rc < 0 ? rc : -rc) : st;
        degree[i] = rc ? NAN : dg;
    }
    return 0;
}

