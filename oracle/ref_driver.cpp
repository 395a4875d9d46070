// ref_driver.cpp — TEST INFRASTRUCTURE.  C entry points around the UNMODIFIED reference
// (/root/reference/ctr_source, compiled where it lies against the stand-in headers of oracle/shim).
// Everything below goes through the reference's public class CTR::Robot<double> (ctr_robot.h:76-173):
// no arithmetic of the forward kinematics is written here.  Output: oracle/_ref/libctr_ref.so
// (git-ignored; built by `make -C oracle ref` when /root/reference exists).
// Loaded only by tests/, bench.py's cpu_baseline / --impl reference legs and the golden generator.
#include <cstdint>
#include <cstring>
#include <array>
#include <vector>

#include "ctr_robot.h"
#include "section_i.h"
#include "../include/ctr_fk.h"

#ifdef _OPENMP
#include <omp.h>
#endif

typedef CTR::Robot<double> RefRobot;

extern "C" {

void* ctr_ref_create_xml(const char* path, int64_t max_samples)
{
    RefRobot* r = new RefRobot(size_t(max_samples), std::string(path));
    if (!r->init()) { delete r; return nullptr; }
    return r;
}

// Robot from explicit sections and an explicit (already orthogonal) entry frame, through the
// container constructor ctr_robot.h:96-98.
void* ctr_ref_create_desc(const ctr_robot_desc* d)
{
    typedef CTR::types::SecCC<double> SecCC;
    typedef CTR::types::SecVC<double> SecVC;
    std::vector<std::pair<SecCC, size_t>> cc;
    std::vector<std::pair<SecVC, size_t>> vc;
    for (int s = 0; s < d->nsections; ++s) {
        const ctr_section_desc& q = d->sec[s];
        if (q.ntube == 1) {
            SecCC c({q.stiffness[0]}, {q.curvature[0]}, q.length, {q.radius[0]}, {q.poisson[0]});
            cc.push_back(std::make_pair(c, size_t(s)));
        } else {
            std::array<double, 2> k{{q.stiffness[0], q.stiffness[1]}}, u{{q.curvature[0], q.curvature[1]}},
                r{{q.radius[0], q.radius[1]}}, p{{q.poisson[0], q.poisson[1]}};
            SecVC v(k, u, q.length, r, p);
            vc.push_back(std::make_pair(v, size_t(s)));
        }
    }
    const double* t = d->init_trafo;   // column-major 3x4 -> row-major constructor arguments
    RefRobot::Transform init(t[0], t[3], t[6], t[9], t[1], t[4], t[7], t[10], t[2], t[5], t[8], t[11]);
    RefRobot* r = new RefRobot(size_t(d->max_samples), init, cc, vc);
    if (!r->init()) { delete r; return nullptr; }
    return r;
}

void ctr_ref_destroy(void* h) { delete static_cast<RefRobot*>(h); }

int ctr_ref_info(void* h, int32_t* nsections, int32_t* ntubes, int32_t* njoints, int32_t* max_samples, double* max_length)
{
    RefRobot* r = static_cast<RefRobot*>(h);
    if (!r || !r->init()) return -1;
    if (nsections) *nsections = int32_t(r->getNSections());
    if (ntubes) *ntubes = int32_t(r->getNTubes());
    if (njoints) *njoints = int32_t(r->getNJointValues());
    if (max_samples) *max_samples = int32_t(r->getMaxSamples());
    if (max_length) *max_length = r->getMaxLength();
    return 0;
}

// State left by the last calcKinematic (for XML robots: the construction-time call, ctr_static.h:248-252).
int ctr_ref_last_state(void* h, int32_t* n_out, double* tip12, double* jv_out)
{
    RefRobot* r = static_cast<RefRobot*>(h);
    if (!r || !r->init()) return -1;
    if (n_out) *n_out = int32_t(r->getNSamples());
    if (tip12) r->getTipTransform().getData(tip12, Erl::ColMajor);
    if (jv_out) { RefRobot::VectorJ jv = r->getJointValues(); std::memcpy(jv_out, jv.data(), sizeof(double) * size_t(jv.size())); }
    return 0;
}

static void fk_one(RefRobot& r, const double* jv, int nj, int mode, double step, int32_t nsamples, double* tip,
                   double* backbone, double* radius, int32_t* n_out, double* step_out)
{
    RefRobot::VectorJ q(nj, 1);
    std::memcpy(q.data(), jv, sizeof(double) * size_t(nj));
    double h = 0.0;
    RefRobot::Transform t = (mode == 0) ? r.calcKinematic(q, step, h) : r.calcKinematic(q, size_t(nsamples), h);
    const size_t n = r.getNSamples();
    if (tip) t.getData(tip, Erl::ColMajor);
    if (n_out) *n_out = int32_t(n);
    if (step_out) *step_out = h;
    if (backbone) for (size_t k = 0; k < n; ++k) r.getTransformSample(unsigned(k)).getData(backbone + 12 * k, Erl::ColMajor);
    if (radius) for (size_t k = 0; k < n; ++k) radius[k] = r.getRadiusSample(unsigned(k));
}

// B configurations; one copy of the robot per thread (the reference object is not thread-safe).
// mode 0 = calcKinematic(JV, ArcLengthStep, ret), 1 = calcKinematic(JV, NSamples, ret).
int ctr_ref_fk_batch(void* h, const double* jv, int64_t B, int mode, double step, int32_t nsamples, int nthreads,
                     double* tip, double* backbone, double* radius, int32_t* n_out, double* step_out)
{
    RefRobot* r0 = static_cast<RefRobot*>(h);
    if (!r0 || !r0->init()) return -1;
    const int nj = int(r0->getNJointValues());
    const int64_t maxs = int64_t(r0->getMaxSamples());
    if (nthreads < 1) nthreads = 1;
#pragma omp parallel num_threads(nthreads)
    {
        RefRobot mine(*r0);
#pragma omp for schedule(static)
        for (int64_t i = 0; i < B; ++i)
            fk_one(mine, jv + i * nj, nj, mode, step, nsamples, tip ? tip + 12 * i : nullptr,
                   backbone ? backbone + i * maxs * 12 : nullptr, radius ? radius + i * maxs : nullptr,
                   n_out ? n_out + i : nullptr, step_out ? step_out + i : nullptr);
    }
    return 0;
}

// calcKinematic(JV, N) followed by calcJacobian(JV, Tip, N, dTrans, dRot, J) (robot_i.h:917-960).  jac: 6 x n_jv column-major.
int ctr_ref_jacobian(void* h, const double* jv, int32_t nsamples, double fd_trans, double fd_rot, double* jac, double* tip12)
{
    RefRobot* r = static_cast<RefRobot*>(h);
    if (!r || !r->init()) return -1;
    const int nj = int(r->getNJointValues());
    RefRobot::VectorJ q(nj, 1);
    std::memcpy(q.data(), jv, sizeof(double) * size_t(nj));
    RefRobot::Transform tip = r->calcKinematic(q, size_t(nsamples));
    RefRobot::MatJacobian J(6, nj);
    J.setZero();
    r->calcJacobian(q, tip, size_t(nsamples), fd_trans, fd_rot, J);
    std::memcpy(jac, J.data(), sizeof(double) * 6 * size_t(nj));
    if (tip12) tip.getData(tip12, Erl::ColMajor);
    return 0;
}

int ctr_ref_stability(void* h, const double* jv, double angle_step, double arc_step)
{
    RefRobot* r = static_cast<RefRobot*>(h);
    if (!r || !r->init()) return -1;
    const int nj = int(r->getNJointValues());
    RefRobot::VectorJ q(nj, 1);
    std::memcpy(q.data(), jv, sizeof(double) * size_t(nj));
    return r->calcStability(q, angle_step, arc_step) ? 1 : 0;
}

int ctr_ref_degree_stability(void* h, const double* jv, double angle_step, double arc_step, double min_degree,
                             double* degree, int* stable_out)
{
    RefRobot* r = static_cast<RefRobot*>(h);
    if (!r || !r->init()) return -1;
    const int nj = int(r->getNJointValues());
    RefRobot::VectorJ q(nj, 1);
    std::memcpy(q.data(), jv, sizeof(double) * size_t(nj));
    const double d = r->calcDegreeStability(q, angle_step, arc_step, min_degree);
    if (degree) *degree = d;
    if (stable_out) *stable_out = r->getStable() ? 1 : 0;
    return 0;
}

// ctr_test.cpp:30-32 — a default-constructed std::mt19937 and uniform_real_distribution, B consecutive draws.
int ctr_ref_random_joints(void* h, int64_t B, double* jv)
{
    RefRobot* r = static_cast<RefRobot*>(h);
    if (!r || !r->init()) return -1;
    const int nj = int(r->getNJointValues());
    CT_RNG rng;
    CT_RND<double> rnd;
    RefRobot::VectorJ q(nj, 1);
    for (int64_t i = 0; i < B; ++i) {
        r->getRandomJointValues(rng, rnd, q);
        std::memcpy(jv + i * nj, q.data(), sizeof(double) * size_t(nj));
    }
    return 0;
}

int ctr_ref_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

}  // extern "C"

// ---- sampling scale policies (transrotscale.h), driven with caller-supplied uniform variates ----------
#include "transrotscale.h"

extern "C" {

// GammaCorrection<double, LookUpSize>::scaleTrans (transrotscale.h:106-122) after set_scale(gamma).
// lut_size must be one of the instantiated sizes: 0 (closed form), 16, 64, 256, 1024.
int ctr_ref_gamma_scale(int lut_size, double gamma, const double* u, int64_t n, double* out)
{
#define CTR_REF_GAMMA_CASE(L)                                        \
    case L: {                                                        \
        CTR::GammaCorrection<double, L> g;                           \
        g.set_scale(gamma);                                          \
        for (int64_t i = 0; i < n; ++i) out[i] = g.scaleTrans(u[i]); \
        return 0;                                                    \
    }
    switch (lut_size) {
        CTR_REF_GAMMA_CASE(0)
        CTR_REF_GAMMA_CASE(16)
        CTR_REF_GAMMA_CASE(64)
        CTR_REF_GAMMA_CASE(256)
        CTR_REF_GAMMA_CASE(1024)
        default: return -2;
    }
#undef CTR_REF_GAMMA_CASE
}

// ProportionalScale<double> (transrotscale.h:139-309): initialize(robot), set_scale(p0, p1), then one
// scaleTrans call per section and configuration, u[B][nsections] -> out[B][nsections].  The first
// configuration's goal comes from std::rand() inside set_scale (:186-188); goal_first returns it
// (get_scale()) so that a caller can reproduce it.
int ctr_ref_proportional_scale(void* h, double p0, double p1, const double* u, int64_t B, double* out, double* goal_first)
{
    RefRobot* r = static_cast<RefRobot*>(h);
    if (!r || !r->init()) return -1;
    CTR::ProportionalScale<double> ps;
    if (!ps.initialize(*r)) return -1;
    ps.set_scale(p0, p1);
    if (goal_first) *goal_first = ps.get_scale();
    const int S = int(r->getNSections());
    for (int64_t i = 0; i < B; ++i)
        for (int s = 0; s < S; ++s) out[i * S + s] = ps.scaleTrans(u[i * S + s]);
    return 0;
}

}  // extern "C"

// ---- the other calcJacobian forms ----------------------------------------------------------------------
extern "C" {

// calcJacobian(JV, ArcLengthStep, dTrans, dRot, J) (ctr_robot.h:121, robot_i.h:901-907)
int ctr_ref_jacobian_step(void* h, const double* jv, double step, double fd_trans, double fd_rot, double* jac,
                          double* tip12, int32_t* n_out)
{
    RefRobot* r = static_cast<RefRobot*>(h);
    if (!r || !r->init()) return -1;
    const int nj = int(r->getNJointValues());
    RefRobot::VectorJ q(nj, 1);
    std::memcpy(q.data(), jv, sizeof(double) * size_t(nj));
    if (tip12 || n_out) {       // the base pose of the step form, from a call of its own (the Jacobian call redoes it)
        RefRobot::Transform t = r->calcKinematic(q, step);
        if (tip12) t.getData(tip12, Erl::ColMajor);
        if (n_out) *n_out = int32_t(r->getNSamples());
    }
    RefRobot::MatJacobian J(6, nj);
    J.setZero();
    r->calcJacobian(q, step, fd_trans, fd_rot, J);
    std::memcpy(jac, J.data(), sizeof(double) * 6 * size_t(nj));
    return 0;
}

// calcJacobian(JV, NSamples, dTrans, dRot, iSample, JTip, JiSample) (ctr_robot.h:122, robot_i.h:908-915, :1012-1072)
int ctr_ref_jacobian_sample(void* h, const double* jv, int32_t nsamples, int32_t i_sample, double fd_trans, double fd_rot,
                            double* jac_tip, double* jac_sample)
{
    RefRobot* r = static_cast<RefRobot*>(h);
    if (!r || !r->init()) return -1;
    const int nj = int(r->getNJointValues());
    RefRobot::VectorJ q(nj, 1);
    std::memcpy(q.data(), jv, sizeof(double) * size_t(nj));
    RefRobot::MatJacobian J(6, nj), K(6, nj);
    J.setZero(); K.setZero();
    r->calcJacobian(q, size_t(nsamples), fd_trans, fd_rot, size_t(i_sample), J, K);
    std::memcpy(jac_tip, J.data(), sizeof(double) * 6 * size_t(nj));
    std::memcpy(jac_sample, K.data(), sizeof(double) * 6 * size_t(nj));
    return 0;
}

}  // extern "C"

// ---- the adapter a user of the reference includes (include/ctr_fk_reference_adapter.hpp), exercised on a live
// reference robot (the header calls nothing of libctrfk.so).
#include "../include/ctr_fk_reference_adapter.hpp"
extern "C" int ctr_ref_adapter_desc(void* h, ctr_robot_desc* out)
{
    RefRobot* r = static_cast<RefRobot*>(h);
    if (!r) return -1;
    return ctr_fk_desc_from_reference(*r, out);
}

// ---- the float instantiation of the reference (CT_ROBOT_FLOAT, ctr_robot.cpp:724-731) -----------------------
// Only to put the FP32 variant's error next to the reference's own single-precision build.  Note: real Eigen
// 3.3.3 evaluates Array<float>::sin()/cos() with its SSE psin/pcos approximations; the stand-in uses libm's
// sinf/cosf, which are at least as accurate — so this is a lower bound of the reference's float error.
typedef CTR::Robot<float> RefRobotF;

extern "C" {

void* ctr_ref_create_desc_f32(const ctr_robot_desc* d)
{
    typedef CTR::types::SecCC<float> SecCC;
    typedef CTR::types::SecVC<float> SecVC;
    std::vector<std::pair<SecCC, size_t>> cc;
    std::vector<std::pair<SecVC, size_t>> vc;
    for (int s = 0; s < d->nsections; ++s) {
        const ctr_section_desc& q = d->sec[s];
        if (q.ntube == 1) {
            SecCC c({float(q.stiffness[0])}, {float(q.curvature[0])}, float(q.length), {float(q.radius[0])}, {float(q.poisson[0])});
            cc.push_back(std::make_pair(c, size_t(s)));
        } else {
            std::array<float, 2> k{{float(q.stiffness[0]), float(q.stiffness[1])}}, u{{float(q.curvature[0]), float(q.curvature[1])}},
                r{{float(q.radius[0]), float(q.radius[1])}}, p{{float(q.poisson[0]), float(q.poisson[1])}};
            SecVC v(k, u, float(q.length), r, p);
            vc.push_back(std::make_pair(v, size_t(s)));
        }
    }
    const double* t = d->init_trafo;
    float f[12];
    for (int i = 0; i < 12; ++i) f[i] = (float)t[i];
    RefRobotF::Transform init(f[0], f[3], f[6], f[9], f[1], f[4], f[7], f[10], f[2], f[5], f[8], f[11]);
    RefRobotF* r = new RefRobotF(size_t(d->max_samples), init, cc, vc);
    if (!r->init()) { delete r; return nullptr; }
    return r;
}
void ctr_ref_destroy_f32(void* h) { delete static_cast<RefRobotF*>(h); }

// joint values in double (rounded to float on entry), tip out in double (widened), column-major 3x4
int ctr_ref_fk_batch_f32(void* h, const double* jv, int64_t B, int mode, double step, int32_t nsamples, double* tip, int32_t* n_out)
{
    RefRobotF* r = static_cast<RefRobotF*>(h);
    if (!r || !r->init()) return -1;
    const int nj = int(r->getNJointValues());
    RefRobotF::VectorJ q(nj, 1);
    for (int64_t i = 0; i < B; ++i) {
        for (int j = 0; j < nj; ++j) q[j] = float(jv[i * nj + j]);
        float hret = 0.f, t12[12];
        RefRobotF::Transform t = (mode == 0) ? r->calcKinematic(q, float(step), hret) : r->calcKinematic(q, size_t(nsamples), hret);
        t.getData(t12, Erl::ColMajor);
        for (int e = 0; e < 12; ++e) tip[12 * i + e] = double(t12[e]);
        if (n_out) n_out[i] = int32_t(r->getNSamples());
    }
    return 0;
}

}  // extern "C"
