// ctr_robot.cpp — see ctr_robot.hpp.
#include "ctr_robot.hpp"

#include <cmath>
#include <cstring>
#include <iostream>

namespace ctrb200 {

namespace {
void fool(const char* func)
{
    std::cerr << "FOOL: no robot, FUNC:" << func << std::endl;        // ctr_robot.cpp:103
}
}  // namespace

Transform Transform::Identity()
{
    Transform t;
    for (int i = 0; i < 12; ++i) t.m[i] = 0.0;
    t.m[0] = t.m[4] = t.m[8] = 1.0;
    return t;
}

Transform Transform::operator*(const Transform& b) const
{
    Transform c;
    for (int j = 0; j < 3; ++j)
        for (int i = 0; i < 3; ++i)
            c.m[3 * j + i] = m[i] * b.m[3 * j] + m[3 + i] * b.m[3 * j + 1] + m[6 + i] * b.m[3 * j + 2];
    for (int i = 0; i < 3; ++i) c.m[9 + i] = (m[i] * b.m[9] + m[3 + i] * b.m[10] + m[6 + i] * b.m[11]) + m[9 + i];
    return c;
}

Robot::Robot() : engine_(nullptr), device_(0), nsamples_(0) { std::memset(&desc_, 0, sizeof desc_); }

Robot::Robot(const Robot& o) : desc_(o.desc_), engine_(nullptr), device_(o.device_), nsamples_(0)
{
    if (o.engine_) build(device_);                  // one engine per copy, like one arena per copy
    nsamples_ = o.nsamples_;                        // (ctr_robot.cpp:30-33 -> ctr_static.h:73-85)
    if (engine_) { frames_ = o.frames_; radius_ = o.radius_; alpha_base_ = o.alpha_base_; last_jv_ = o.last_jv_; }
}

Robot& Robot::operator=(const Robot& o)
{
    if (this == &o) return *this;
    destroy();
    desc_ = o.desc_;
    device_ = o.device_;
    if (o.engine_) build(device_);
    if (engine_) { nsamples_ = o.nsamples_; frames_ = o.frames_; radius_ = o.radius_; alpha_base_ = o.alpha_base_; last_jv_ = o.last_jv_; }
    return *this;
}

Robot::Robot(std::size_t maxSample, const Transform& init, const std::vector<ctr_section_desc>& sections, int device)
    : engine_(nullptr), device_(device), nsamples_(0)
{
    std::memset(&desc_, 0, sizeof desc_);
    if (sections.empty() || sections.size() > CTR_FK_MAX_SECTIONS || maxSample == 0) return;
    desc_.nsections = (int32_t)sections.size();
    desc_.max_samples = (int32_t)maxSample;
    init.getData(desc_.init_trafo);
    if (ctr_fk_orthogonalize(desc_.init_trafo, 0.01) != 0) { std::cerr << "ERROR: " << ctr_fk_last_error() << std::endl; return; }
    for (std::size_t s = 0; s < sections.size(); ++s) desc_.sec[s] = sections[s];
    build(device);
}

Robot::Robot(std::size_t maxSample, const std::string& xml, int device) : engine_(nullptr), device_(device), nsamples_(0)
{
    std::memset(&desc_, 0, sizeof desc_);
    double phi[CTR_FK_MAX_SECTIONS];
    if (ctr_fk_desc_from_xml(xml.c_str(), (int32_t)maxSample, &desc_, phi) != 0) {
        std::cerr << "ERROR: " << ctr_fk_last_error() << std::endl;   // ctr_static.h:97-113, :243-247
        return;
    }
    build(device);
    if (!engine_) return;
    // construction-time FK at the XML's phi values, ctr_static.h:248-252
    VectorJ jv(getNJointValues(), 0.0);
    std::size_t t0 = 0;
    for (int s = 0; s < desc_.nsections; ++s) { jv[1 + s + t0] = phi[s]; t0 += (std::size_t)desc_.sec[s].ntube; }
    calcKinematic(jv, getMaxLength() / (double)getMaxSamples());
}

Robot::~Robot() { destroy(); }

void Robot::build(int device)
{
    if (ctr_fk_create(&desc_, device, &engine_) != 0) {
        std::cerr << "ERROR: ctr_fk_create: " << ctr_fk_last_error() << std::endl;
        engine_ = nullptr;
        return;
    }
    frames_.assign((std::size_t)desc_.max_samples * 12, 0.0);
    radius_.assign((std::size_t)desc_.max_samples, 0.0);
    alpha_base_.assign(getNTubes(), 0.0);
}

void Robot::destroy()
{
    if (engine_) ctr_fk_destroy(engine_);
    engine_ = nullptr;
    nsamples_ = 0;
}

Transform Robot::solve_single(const double* jv, int mode, double step, std::size_t ns, double* stepReturn)
{
    Transform tip = Transform::Identity();
    int32_t n = 0;
    double h = 0.0;
    int rc = ctr_fk_single(engine_, jv, mode, step, (int32_t)ns, 0u, tip.m, frames_.data(), radius_.data(), &n, &h,
                           alpha_base_.data());
    if (rc != 0 && rc != CTR_FK_E_RANGE) std::cerr << "ERROR: calcKinematic: " << ctr_fk_last_error() << std::endl;
    nsamples_ = (std::size_t)(n > 0 ? n : 0);
    last_jv_.assign(jv, jv + getNJointValues());
    if (stepReturn) *stepReturn = h;
    return tip;
}

Transform Robot::calcKinematic(const VectorJ& jv, const double& step, double& stepReturn)
{
    if (!engine_) { fool(__func__); return Transform::Identity(); }
    if (jv.size() < getNJointValues()) { std::cerr << "ERROR: calcKinematic: joint vector too short" << std::endl; return Transform::Identity(); }
    return solve_single(jv.data(), CTR_FK_MODE_STEP, step, 0, &stepReturn);
}
Transform Robot::calcKinematic(const VectorJ& jv, const double& step)
{
    double r;
    return calcKinematic(jv, step, r);
}
Transform Robot::calcKinematic(const VectorJ& jv, const std::size_t& ns, double& stepReturn)
{
    if (!engine_) { fool(__func__); return Transform::Identity(); }
    if (jv.size() < getNJointValues()) { std::cerr << "ERROR: calcKinematic: joint vector too short" << std::endl; return Transform::Identity(); }
    return solve_single(jv.data(), CTR_FK_MODE_NSAMPLE, 0.0, ns, &stepReturn);
}
Transform Robot::calcKinematic(const VectorJ& jv, const std::size_t& ns)
{
    double r;
    return calcKinematic(jv, ns, r);
}
void Robot::calcKinematic(double* const jv, const double& step, double* tr)
{
    if (!engine_) { fool(__func__); return; }                          // ctr_robot.cpp:178-184
    solve_single(jv, CTR_FK_MODE_STEP, step, 0, nullptr).getData(tr);   // column-major, :171
}

int Robot::calcKinematicBatch(const double* jv, std::int64_t B, const double& step, double* tip, std::int32_t* nOut,
                              double* stepOut, std::int32_t* status)
{
    if (!engine_) { fool(__func__); return CTR_FK_E_NULL; }
    return ctr_fk_batch_host(engine_, jv, B, CTR_FK_MODE_STEP, step, 0, 0u, tip, nOut, stepOut, status);
}
int Robot::calcKinematicBatch(const double* jv, std::int64_t B, const std::size_t& ns, double* tip, std::int32_t* nOut,
                              double* stepOut, std::int32_t* status)
{
    if (!engine_) { fool(__func__); return CTR_FK_E_NULL; }
    return ctr_fk_batch_host(engine_, jv, B, CTR_FK_MODE_NSAMPLE, 0.0, (int32_t)ns, 0u, tip, nOut, stepOut, status);
}

int Robot::calcKinematicBatch(const double* jv, std::int64_t B, const std::size_t& ns, double* tip, double* backbone,
                              double* radius, std::int32_t* nOut)
{
    if (!engine_) { fool(__func__); return CTR_FK_E_NULL; }
    return ctr_fk_batch_host_backbone(engine_, jv, B, CTR_FK_MODE_NSAMPLE, 0.0, (int32_t)ns, 0u, tip, backbone, radius, nOut,
                                      nullptr, nullptr);
}

std::size_t Robot::getNSections() const { return engine_ ? (std::size_t)desc_.nsections : 0; }
std::size_t Robot::getNTubes() const { return engine_ ? (std::size_t)ctr_fk_desc_ntubes(&desc_) : 0; }
std::size_t Robot::getNJointValues() const { return engine_ ? (std::size_t)ctr_fk_desc_njoints(&desc_) : 0; }
double Robot::getMaxLength() const noexcept { return engine_ ? ctr_fk_desc_max_length(&desc_) : 0.0; }
std::size_t Robot::getMaxSamples() const noexcept { return engine_ ? (std::size_t)desc_.max_samples : 0; }

double Robot::getRadiusSample(unsigned s) const noexcept { return (engine_ && s < nsamples_) ? radius_[s] : 0.0; }
Transform Robot::getTransformSample(unsigned s) const noexcept
{
    Transform t = Transform::Identity();
    if (engine_ && s < nsamples_) std::memcpy(t.m, &frames_[12 * (std::size_t)s], sizeof t.m);
    return t;
}
Vector3 Robot::getTranslationSample(unsigned s) const noexcept { return getTransformSample(s).getTranslation(); }
double Robot::getTipRadius() const noexcept { return nsamples_ ? getRadiusSample((unsigned)nsamples_ - 1) : 0.0; }
Transform Robot::getTipTransform() const noexcept { return nsamples_ ? getTransformSample((unsigned)nsamples_ - 1) : Transform::Identity(); }
Vector3 Robot::getTipTranslation() const noexcept { return getTipTransform().getTranslation(); }

bool Robot::getSection(unsigned i, ctr_section_desc* out) const noexcept
{
    if (!engine_ || !out || i >= (unsigned)desc_.nsections) return false;
    *out = desc_.sec[i];
    return true;
}

void Robot::getRandomJointValues(CT_RNG& rng, CT_RND& rnd, VectorJ& jv) const
{
    if (!engine_) { fool(__func__); return; }
    jv.assign(getNJointValues(), 0.0);
    const double c = std::nextafter(2.0 * M_PI, 0.0);
    jv[0] = c * rnd(rng);                                               // robot_i.h:1310
    std::size_t t0 = 0;
    for (int s = 0; s < desc_.nsections; ++s) {                         // section_i.h:331-336
        jv[1 + s + t0] = desc_.sec[s].length * rnd(rng);
        for (int k = 0; k < desc_.sec[s].ntube; ++k) jv[2 + s + t0 + k] = c * rnd(rng);
        t0 += (std::size_t)desc_.sec[s].ntube;
    }
}

void Robot::getRandomJointValues(CT_RNG& rng, CT_RND& rnd, VectorJ& jv, const TransRotScale& ts) const
{
    if (!engine_) { fool(__func__); return; }
    jv.assign(getNJointValues(), 0.0);
    const double c = std::nextafter(2.0 * M_PI, 0.0);
    jv[0] = c * rnd(rng);
    std::size_t t0 = 0;
    for (int s = 0; s < desc_.nsections; ++s) {                         // section_i.h:341-346
        jv[1 + s + t0] = desc_.sec[s].length * ts.scaleTrans(rnd(rng));
        for (int k = 0; k < desc_.sec[s].ntube; ++k) jv[2 + s + t0 + k] = c * rnd(rng);
        t0 += (std::size_t)desc_.sec[s].ntube;
    }
}

bool ProportionalScale::initialize(const Robot& robot)
{
    if (!robot.init()) return false;
    std::vector<double> lengths;
    for (std::size_t s = 0; s < robot.getNSections(); ++s) {
        ctr_section_desc sd;
        if (!robot.getSection((unsigned)s, &sd)) return false;
        lengths.push_back(sd.length);
    }
    return initialize(lengths);
}

bool Robot::setMaxSample(std::size_t maxSample)
{
    if (maxSample == 0 || !engine_) return false;
    if ((std::size_t)desc_.max_samples == maxSample) return true;
    ctr_robot_desc nd = desc_;
    nd.max_samples = (int32_t)maxSample;
    ctr_fk_handle ne = nullptr;
    if (ctr_fk_create(&nd, device_, &ne) != 0) return false;           // the old robot stays usable, like a failed malloc
    ctr_fk_destroy(engine_);
    engine_ = ne;
    desc_ = nd;
    nsamples_ = 0;
    frames_.assign(maxSample * 12, 0.0);
    radius_.assign(maxSample, 0.0);
    return true;
}

int RobotF::solve(const float* jv, std::int64_t B, int mode, double step, std::size_t ns, float* tip, std::int32_t* nOut,
                  double* stepOut)
{
    if (!r_.init()) { fool(__func__); return CTR_FK_E_NULL; }
    const std::size_t nj = r_.getNJointValues();
    std::vector<double> q((std::size_t)B * nj), t((std::size_t)B * 12), h((std::size_t)B);
    for (std::size_t i = 0; i < q.size(); ++i) q[i] = (double)jv[i];
    int rc = ctr_fk_batch_host(r_.handle(), q.data(), B, mode, step, (int32_t)ns, CTR_FK_FLAG_F32, t.data(), nOut, h.data(), nullptr);
    if (rc) { std::cerr << "ERROR: calcKinematic: " << ctr_fk_last_error() << std::endl; return rc; }
    for (std::size_t i = 0; i < t.size(); ++i) tip[i] = (float)t[i];
    if (stepOut && B > 0) *stepOut = h[0];
    return 0;
}
TransformF RobotF::calcKinematic(const VectorJF& jv, const float& step, float& stepReturn)
{
    TransformF t; for (int i = 0; i < 12; ++i) t.m[i] = (i % 4 == 0 && i < 9) ? 1.f : 0.f;
    if (jv.size() < getNJointValues()) { std::cerr << "ERROR: calcKinematic: joint vector too short" << std::endl; return t; }
    std::int32_t n = 0; double h = 0.0;
    if (solve(jv.data(), 1, CTR_FK_MODE_STEP, (double)step, 0, t.m, &n, &h) == 0) { nsamples_ = (std::size_t)(n > 0 ? n : 0); stepReturn = (float)h; }
    return t;
}
TransformF RobotF::calcKinematic(const VectorJF& jv, const float& step) { float r; return calcKinematic(jv, step, r); }
TransformF RobotF::calcKinematic(const VectorJF& jv, const std::size_t& ns, float& stepReturn)
{
    TransformF t; for (int i = 0; i < 12; ++i) t.m[i] = (i % 4 == 0 && i < 9) ? 1.f : 0.f;
    if (jv.size() < getNJointValues()) { std::cerr << "ERROR: calcKinematic: joint vector too short" << std::endl; return t; }
    std::int32_t n = 0; double h = 0.0;
    if (solve(jv.data(), 1, CTR_FK_MODE_NSAMPLE, 0.0, ns, t.m, &n, &h) == 0) { nsamples_ = (std::size_t)(n > 0 ? n : 0); stepReturn = (float)h; }
    return t;
}
TransformF RobotF::calcKinematic(const VectorJF& jv, const std::size_t& ns) { float r; return calcKinematic(jv, ns, r); }
int RobotF::calcKinematicBatch(const float* jv, std::int64_t B, const std::size_t& ns, float* tip, std::int32_t* nOut)
{
    return solve(jv, B, CTR_FK_MODE_NSAMPLE, 0.0, ns, tip, nOut, nullptr);
}

void Robot::calcJacobian(const VectorJ& jv, const std::size_t& ns, const double& fdTrans, const double& fdRot, MatJacobian& J)
{
    J.cols = (int)getNJointValues();
    J.data.assign((std::size_t)6 * J.cols, 0.0);
    if (!engine_) { fool(__func__); return; }
    // B = 1 through the batched kernel; device staging via the host-batch path is not offered for
    // Jacobians, so use a tiny managed round trip through ctr_fk_jacobian_host.
    int rc = ctr_fk_jacobian_host(engine_, jv.data(), 1, (int32_t)ns, fdTrans, fdRot, J.data.data(), nullptr);
    if (rc) std::cerr << "ERROR: calcJacobian: " << ctr_fk_last_error() << std::endl;
}

// shared tail of the other four forms: one configuration through ctr_fk_jacobian_host_ex
void Robot::jacobianEx(const VectorJ& jv, int mode, double step, std::size_t ns, long iSample, double fdTrans, double fdRot,
                       const Transform* tTip, const Transform* tiSample, MatJacobian& J, MatJacobian* Js, const char* who)
{
    J.cols = (int)getNJointValues();
    J.data.assign((std::size_t)6 * J.cols, 0.0);
    if (Js) { Js->cols = J.cols; Js->data.assign((std::size_t)6 * J.cols, 0.0); }
    if (!engine_) { fool(who); return; }
    double tip[12], smp[12];
    if (tTip) tTip->getData(tip);
    if (tiSample) tiSample->getData(smp);
    int rc = ctr_fk_jacobian_host_ex(engine_, jv.data(), 1, mode, step, (int32_t)ns, (int32_t)iSample, fdTrans, fdRot,
                                     tTip ? CTR_FK_FLAG_JAC_GIVEN_POSES : 0u, J.data.data(), Js ? Js->data.data() : nullptr,
                                     tTip ? tip : nullptr, tiSample ? smp : nullptr, nullptr);
    if (rc) std::cerr << "ERROR: " << who << ": " << ctr_fk_last_error() << std::endl;
}

void Robot::calcJacobian(const VectorJ& jv, const double& step, const double& fdTrans, const double& fdRot, MatJacobian& J)
{ jacobianEx(jv, CTR_FK_MODE_STEP, step, 1, -1, fdTrans, fdRot, nullptr, nullptr, J, nullptr, __func__); }

void Robot::calcJacobian(const VectorJ& jv, const std::size_t& ns, const double& fdTrans, const double& fdRot,
                         const std::size_t& iSample, MatJacobian& J, MatJacobian& Js)
{ jacobianEx(jv, CTR_FK_MODE_NSAMPLE, 0.0, ns, (long)iSample, fdTrans, fdRot, nullptr, nullptr, J, &Js, __func__); }

void Robot::calcJacobian(const VectorJ& jv, const Transform& tTip, const std::size_t& ns, const double& fdTrans,
                         const double& fdRot, MatJacobian& J)
{ jacobianEx(jv, CTR_FK_MODE_NSAMPLE, 0.0, ns, -1, fdTrans, fdRot, &tTip, nullptr, J, nullptr, __func__); }

void Robot::calcJacobian(const VectorJ& jv, const Transform& tTip, const Transform& tiSample, const std::size_t& ns,
                         const double& fdTrans, const double& fdRot, const std::size_t& iSample, MatJacobian& J, MatJacobian& Js)
{ jacobianEx(jv, CTR_FK_MODE_NSAMPLE, 0.0, ns, (long)iSample, fdTrans, fdRot, &tTip, &tiSample, J, &Js, __func__); }

bool Robot::calcStability(const VectorJ& jv, const double& angleStep, const double& arcStep)
{
    if (!engine_) { fool(__func__); return false; }                     // ctr_robot.cpp null-robot branch
    int32_t st = 0;
    int rc = ctr_fk_stability_host(engine_, jv.data(), 1, angleStep, arcStep, -1.0, &st, nullptr);
    if (rc) std::cerr << "ERROR: calcStability: " << ctr_fk_last_error() << std::endl;
    stable_ = (rc == 0 && st == 1);
    return stable_;
}

double Robot::calcDegreeStability(const VectorJ& jv, const double& angleStep, const double& arcStep, const double& minDegree)
{
    if (!engine_) { fool(__func__); return 0.0; }
    int32_t st = 0;
    double deg = 0.0;
    int rc = ctr_fk_stability_host(engine_, jv.data(), 1, angleStep, arcStep, minDegree, &st, &deg);
    if (rc) std::cerr << "ERROR: calcDegreeStability: " << ctr_fk_last_error() << std::endl;
    stable_ = (rc == 0 && st == 1);
    return deg;
}

std::ostream& operator<<(std::ostream& os, const Robot& r)
{
    if (!r.engine_) return os << "FOOL: no robot";
    os << "MaxSamples     : " << r.desc_.max_samples << "\nSamples        : " << r.nsamples_ << "\n";
    for (int s = 0; s < r.desc_.nsections; ++s) {
        const ctr_section_desc& sd = r.desc_.sec[s];
        os << "Section: " << s << "\nNTube      : " << sd.ntube << "\nLength     : " << sd.length << "\n";
        for (int k = 0; k < sd.ntube; ++k)
            os << "  tube " << k << ": stiffness " << sd.stiffness[k] << " curvature " << sd.curvature[k] << " radius "
               << sd.radius[k] << " poisson " << sd.poisson[k] << "\n";
    }
    return os;
}

}  // namespace ctrb200
