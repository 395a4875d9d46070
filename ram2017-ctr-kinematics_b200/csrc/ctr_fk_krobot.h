// ctr_fk_krobot.h — host-side construction of the solver tables (ctrk::KRobot) from the public
// descriptor.  Every derived constant that the STRICT solver consumes is produced by the same
// IEEE operation the reference performs at run time (e.g. share = ksum / ntube,
// section_i.h:189), so hoisting it out of the sweep does not change a single bit.
#pragma once

#include <cmath>

#include "../../include/ctr_fk.h"
#include "ctr_fk_core.cuh"

namespace ctrk {

inline int validate_desc(const ctr_robot_desc* d)
{
    if (!d) return CTR_FK_E_NULL;
    if (d->nsections < 1 || d->nsections > CTR_FK_MAX_SECTIONS) return CTR_FK_E_DESC;
    if (d->max_samples < 1 || d->max_samples > (1 << 20)) return CTR_FK_E_DESC;
    if (!(d->zero_tol >= 0.0)) return CTR_FK_E_DESC;
    for (int i = 0; i < 12; ++i)
        if (!std::isfinite(d->init_trafo[i])) return CTR_FK_E_DESC;
    for (int s = 0; s < d->nsections; ++s) {
        const ctr_section_desc& sd = d->sec[s];
        if (sd.ntube < 1 || sd.ntube > 2) return CTR_FK_E_DESC;
        if (!(sd.length >= 0.0) || !std::isfinite(sd.length)) return CTR_FK_E_DESC;
        double sum = 0.0;
        for (int k = 0; k < sd.ntube; ++k) {
            if (!std::isfinite(sd.stiffness[k]) || !std::isfinite(sd.curvature[k]) ||
                !std::isfinite(sd.radius[k]) || !std::isfinite(sd.poisson[k]))
                return CTR_FK_E_DESC;
            if (!(sd.stiffness[k] > 0.0)) return CTR_FK_E_DESC;
            if (!(1.0 + sd.poisson[k] > 0.0)) return CTR_FK_E_DESC;
            sum += sd.stiffness[k];
        }
        if (!(sum > 0.0)) return CTR_FK_E_DESC;
    }
    return 0;
}

inline int make_krobot(const ctr_robot_desc* d, KRobot* r)
{
    int rc = validate_desc(d);
    if (rc) return rc;
    KRobot z = {};
    *r = z;
    r->nsec = d->nsections;
    r->maxs = d->max_samples;
    r->ztol = d->zero_tol > 0.0 ? d->zero_tol : 1e-12;
    for (int i = 0; i < 12; ++i) r->ini[i] = d->init_trafo[i];
    const int S = d->nsections;
    int t = 0;
    for (int s = 0; s < S; ++s) {
        const ctr_section_desc& sd = d->sec[s];
        r->ntube[s] = sd.ntube;
        double sum = sd.stiffness[0];                     // matrix.h:181-186, left to right
        for (int k = 1; k < sd.ntube; ++k) sum += sd.stiffness[k];
        r->ksum[s]  = sum;
        r->kappa[s] = sd.curvature[0];
        r->len[s]   = sd.length;
        r->rad[s]   = sd.radius[0];
        r->share[s] = sum / (double)sd.ntube;
        r->kg[s]    = sd.curvature[0] * r->share[s];
        for (int k = 0; k < sd.ntube; ++k, ++t) {
            double nu = sd.poisson[k];
            // section_i.h:273: the first tube of a two-tube LAST section is processed with the
            // second tube's Poisson ratio.
            if (s == S - 1 && sd.ntube == 2 && k == 0) nu = sd.poisson[1];
            r->onu[t] = 1.0 + nu;
            r->czc[t] = r->share[s] / r->onu[t];
        }
    }
    {   // rotation matrix (column-major) -> unit quaternion, Shepperd's method
        const double* R = r->ini;
        const double m00 = R[0], m10 = R[1], m20 = R[2], m01 = R[3], m11 = R[4], m21 = R[5], m02 = R[6], m12 = R[7], m22 = R[8];
        const double tr = m00 + m11 + m22;
        double w, x, y, z;
        if (tr > 0.0) {
            const double s4 = 2.0 * std::sqrt(tr + 1.0);
            w = 0.25 * s4; x = (m21 - m12) / s4; y = (m02 - m20) / s4; z = (m10 - m01) / s4;
        } else if (m00 > m11 && m00 > m22) {
            const double s4 = 2.0 * std::sqrt(1.0 + m00 - m11 - m22);
            w = (m21 - m12) / s4; x = 0.25 * s4; y = (m01 + m10) / s4; z = (m02 + m20) / s4;
        } else if (m11 > m22) {
            const double s4 = 2.0 * std::sqrt(1.0 + m11 - m00 - m22);
            w = (m02 - m20) / s4; x = (m01 + m10) / s4; y = 0.25 * s4; z = (m12 + m21) / s4;
        } else {
            const double s4 = 2.0 * std::sqrt(1.0 + m22 - m00 - m11);
            w = (m10 - m01) / s4; x = (m02 + m20) / s4; y = (m12 + m21) / s4; z = 0.25 * s4;
        }
        const double nrm = std::sqrt(w * w + x * x + y * y + z * z);
        r->qini[0] = w / nrm; r->qini[1] = x / nrm; r->qini[2] = y / nrm; r->qini[3] = z / nrm;
    }
    const ctr_section_desc& last = d->sec[S - 1];
    r->onu_last = 1.0 + last.poisson[last.ntube - 1];
    r->cl       = r->onu_last / r->ksum[S - 1];
    for (int tt = 0; tt < kMaxTube; ++tt) r->nczc[tt] = -r->czc[tt] * r->cl;
    for (int s = S - 1; s >= 0; --s) {
        // the reference accumulates front to back over the ACTIVE sections, which are always a
        // suffix {s..S-1} (End is non-decreasing); keep that order for the sum before inverting.
        double fwd = 0.0;
        for (int q = s; q < S; ++q) fwd += r->ksum[q];
        r->rk[s] = 1.0 / fwd;
    }
    for (int s = 0; s < S; ++s)
        for (int q = 0; q <= s; ++q) r->kgr[s][q] = r->kg[s] * r->rk[q];
    return 0;
}

inline int desc_ntubes(const ctr_robot_desc* d)
{
    int t = 0;
    for (int s = 0; s < d->nsections; ++s) t += d->sec[s].ntube;
    return t;
}

// layout key: S*1000 + n0*100 + n1*10 + n2
inline int layout_key(const ctr_robot_desc* d)
{
    int n[3] = {0, 0, 0};
    for (int s = 0; s < d->nsections; ++s) n[s] = d->sec[s].ntube;
    return d->nsections * 1000 + n[0] * 100 + n[1] * 10 + n[2];
}

// Invoke f.template operator()<Lay<...>>() for the layout of the descriptor.
#define CTRK_FOR_EACH_LAYOUT(X) \
    X(1, 1, 0, 0) X(1, 2, 0, 0) \
    X(2, 1, 1, 0) X(2, 1, 2, 0) X(2, 2, 1, 0) X(2, 2, 2, 0) \
    X(3, 1, 1, 1) X(3, 1, 1, 2) X(3, 1, 2, 1) X(3, 1, 2, 2) \
    X(3, 2, 1, 1) X(3, 2, 1, 2) X(3, 2, 2, 1) X(3, 2, 2, 2)

}  // namespace ctrk
