// ctr_robot.hpp — C++14 host object with the shape of the reference's CTR::Robot<double>
// (ctr_source/ctr_robot.h:76-173), implemented over the C ABI of include/ctr_fk.h.
// Same method names, argument meaning and error behaviour for the forward-kinematics path:
//   - construction from a ctr_resources XML file runs one FK at the XML's phi values
//     (ctr_static.h:248-252);
//   - a robot that failed to construct answers Identity / 0 / false ("FOOL: no robot",
//     ctr_robot.cpp:96-106) instead of throwing;
//   - calcKinematic keeps the last result, read back with getNSamples / getTransformSample /
//     getRadiusSample / getTip* (robot_i.h:520-533).
// New: calcKinematicBatch, the batched solve this engine exists for.  Every solve — single or
// batched — runs on the GPU; there is no CPU implementation behind this class.
#pragma once

#include <cstddef>
#include <cstdint>
#include <iosfwd>
#include <random>
#include <string>
#include <utility>
#include <vector>

#include "ctr_fk.h"
#include "ctr_transrotscale.hpp"

namespace ctrb200 {

struct Vector3 {
    double v[3];
    double x() const { return v[0]; }
    double y() const { return v[1]; }
    double z() const { return v[2]; }
};

// 3x4 rigid transform, column-major like Erl::Transform::getData(.., ColMajor).
struct Transform {
    double m[12];
    static Transform Identity();
    double operator()(int r, int c) const { return m[3 * c + r]; }
    Vector3 getTranslation() const { return Vector3{{m[9], m[10], m[11]}}; }
    void getData(double* out) const { for (int i = 0; i < 12; ++i) out[i] = m[i]; }
    Transform operator*(const Transform& b) const;
};

typedef std::vector<double> VectorJ;
// 6 x n_jv, column-major (Eigen::Matrix<Real,6,Dynamic,ColMajor>, ctr_robot.h:91)
struct MatJacobian {
    int cols = 0;
    std::vector<double> data;
    double operator()(int r, int c) const { return data[6 * c + r]; }
};

typedef std::mt19937 CT_RNG;                                   // _forwards.h:62-66
typedef std::uniform_real_distribution<double> CT_RND;

class Robot {
public:
    Robot();
    Robot(const Robot& other);
    // sections in base-to-tip order; init is orthogonalised like the XML path
    Robot(std::size_t maxSample, const Transform& initTrafo, const std::vector<ctr_section_desc>& sections,
          int device = 0);
    // ctr_robot.h:96-104: the one-tube (CC) and two-tube (VC) sections as two containers of (section, position)
    // pairs — std::vector, std::deque or std::list like the reference's three constructors, or anything iterable;
    // positions number the sections base to tip and must cover 0..n-1 exactly once (else: no robot)
    template <class C0, class C1,
              class = decltype(std::declval<const C0&>().begin()->second), class = decltype(std::declval<const C1&>().begin()->second)>
    Robot(std::size_t maxSample, const Transform& initTrafo, const C0& cc, const C1& vc, int device = 0)
        : Robot(maxSample, initTrafo, merge_sections(cc, vc), device) {}
    Robot(std::size_t maxSample, const std::string& xmlFilename, int device = 0);
    ~Robot();
    Robot& operator=(const Robot& other);

    bool init() const { return engine_ != nullptr; }
    explicit operator bool() const { return init(); }

    // ctr_robot.h:111-117
    Transform calcKinematic(const VectorJ& jv, const double& arcLengthStep);
    Transform calcKinematic(const VectorJ& jv, const double& arcLengthStep, double& arcLengthStepReturn);
    Transform calcKinematic(const VectorJ& jv, const std::size_t& nSamples);
    Transform calcKinematic(const VectorJ& jv, const std::size_t& nSamples, double& arcLengthStepReturn);
    void      calcKinematic(double* const jv, const double& arcLengthStep, double* tr);

    // Batched solve, HOST pointers: jv [B][n_jv] in, tip [B][12] out (column-major 3x4 each);
    // n_out / step_out / status nullable.  Returns the C-ABI code (0 = OK).
    int calcKinematicBatch(const double* jv, std::int64_t B, const double& arcLengthStep, double* tip,
                           std::int32_t* nOut = nullptr, double* stepOut = nullptr, std::int32_t* status = nullptr);
    int calcKinematicBatch(const double* jv, std::int64_t B, const std::size_t& nSamples, double* tip,
                           std::int32_t* nOut = nullptr, double* stepOut = nullptr, std::int32_t* status = nullptr);

    // ... with every sample's frame and radius: backbone [B][getMaxSamples()][12], radius [B][getMaxSamples()]
    // (what getTransformSample / getRadiusSample return after a single call), radius nullable
    int calcKinematicBatch(const double* jv, std::int64_t B, const std::size_t& nSamples, double* tip, double* backbone,
                           double* radius, std::int32_t* nOut = nullptr);

    // ctr_robot.h:118-119; the last result is kept for getStable() (robot_i.h:535)
    bool   calcStability(const VectorJ& jv, const double& angleDiscretizationStep, const double& arcLengthStep);
    double calcDegreeStability(const VectorJ& jv, const double& angleDiscretizationStep, const double& arcLengthStep,
                               const double& minDegreeStability = -1.);
    bool   getStable() const noexcept { return stable_; }

    // ctr_robot.h:121-126: the four calcJacobian forms of the reference ...
    void calcJacobian(const VectorJ& jv, const double& arcLengthStep, const double& fdTrans, const double& fdRot,
                      MatJacobian& jacobianTip);
    void calcJacobian(const VectorJ& jv, const std::size_t& nSamples, const double& fdTrans, const double& fdRot,
                      const std::size_t& iSample, MatJacobian& jacobianTip, MatJacobian& jacobianiSample);
    void calcJacobian(const VectorJ& jv, const Transform& tTip, const std::size_t& nSamples, const double& fdTrans,
                      const double& fdRot, MatJacobian& jacobianTip);
    void calcJacobian(const VectorJ& jv, const Transform& tTip, const Transform& tiSample, const std::size_t& nSamples,
                      const double& fdTrans, const double& fdRot, const std::size_t& iSample, MatJacobian& jacobianTip,
                      MatJacobian& jacobianiSample);
    // ... and the NSamples form with the tip only (calcKinematic(JV, N) + the third form; not in the reference's façade)
    void calcJacobian(const VectorJ& jv, const std::size_t& nSamples, const double& fdTrans, const double& fdRot,
                      MatJacobian& jacobianTip);

    std::size_t getNSections() const;
    std::size_t getNTubes() const;
    std::size_t getNJointValues() const;
    double      getMaxLength() const noexcept;
    std::size_t getMaxSamples() const noexcept;
    std::size_t getNSamples() const noexcept { return nsamples_; }
    double      getRadiusSample(unsigned s) const noexcept;
    Transform   getTransformSample(unsigned s) const noexcept;
    Vector3     getTranslationSample(unsigned s) const noexcept;
    double      getTipRadius() const noexcept;
    Transform   getTipTransform() const noexcept;
    Vector3     getTipTranslation() const noexcept;
    VectorJ     getJointValues() const noexcept { return last_jv_; }
    void        getJointValues(VectorJ& jv) const noexcept { jv = last_jv_; }          // ctr_robot.h:134
    // ctr_robot.h:149 / ctr_robot.cpp:269: same robot with another sample capacity; the last result is dropped
    // (the reference moves the robot into a fresh arena).  false: no robot, MaxSample 0, or the engine refused
    bool        setMaxSample(std::size_t maxSample);
    const std::vector<double>& getBaseAngles() const noexcept { return alpha_base_; }
    bool        getSection(unsigned i, ctr_section_desc* out) const noexcept;
    const ctr_robot_desc& descriptor() const { return desc_; }
    ctr_fk_handle handle() const { return engine_; }

    void getRandomJointValues(CT_RNG& rng, CT_RND& rnd, VectorJ& jvReturn) const;   // robot_i.h:1308-1317
    // ctr_robot.h:156, section_i.h:338-347: extension lengths through the policy's scaleTrans, angles unscaled
    void getRandomJointValues(CT_RNG& rng, CT_RND& rnd, VectorJ& jvReturn, const TransRotScale& ts) const;

    friend std::ostream& operator<<(std::ostream& os, const Robot& r);

private:
    template <class C0, class C1>
    static std::vector<ctr_section_desc> merge_sections(const C0& cc, const C1& vc)
    {
        std::size_t n = 0;
        for (auto it = cc.begin(); it != cc.end(); ++it) ++n;
        for (auto it = vc.begin(); it != vc.end(); ++it) ++n;
        std::vector<ctr_section_desc> out(n);
        std::vector<char> seen(n, 0);
        bool ok = n > 0;
        for (auto it = cc.begin(); it != cc.end(); ++it) ok = place(out, seen, it->first, (std::size_t)it->second, 1) && ok;
        for (auto it = vc.begin(); it != vc.end(); ++it) ok = place(out, seen, it->first, (std::size_t)it->second, 2) && ok;
        if (!ok) out.clear();
        return out;
    }
    static bool place(std::vector<ctr_section_desc>& out, std::vector<char>& seen, const ctr_section_desc& s,
                      std::size_t pos, int ntube)
    {
        if (pos >= out.size() || seen[pos] || s.ntube != ntube) return false;
        out[pos] = s; seen[pos] = 1;
        return true;
    }
    void build(int device);
    void destroy();
    Transform solve_single(const double* jv, int mode, double step, std::size_t ns, double* stepReturn);
    void jacobianEx(const VectorJ& jv, int mode, double step, std::size_t ns, long iSample, double fdTrans, double fdRot,
                    const Transform* tTip, const Transform* tiSample, MatJacobian& J, MatJacobian* Js, const char* who);

    ctr_robot_desc desc_;
    ctr_fk_handle  engine_;
    int            device_;
    std::size_t    nsamples_;
    bool           stable_ = true;
    std::vector<double> frames_;      // [max_samples][12]
    std::vector<double> radius_;      // [max_samples]
    std::vector<double> alpha_base_;
    VectorJ             last_jv_;
};

// The shape of the reference's float instantiation (CTR::Robot<float>, ctr_robot.cpp:724-731) over ctr_fk_batch_f32:
// joint values and poses in single precision at the interface, FP32 state arithmetic on the device (sample count,
// arc positions and interval tests stay FP64, so getNSamples() equals the double robot's).  Tolerance against the
// double robot: 5e-5 m / 5e-4 rad (tests/test_gpu_parity.py::test_f32_variant).
struct TransformF {
    float m[12];
    float operator()(int r, int c) const { return m[3 * c + r]; }
};
typedef std::vector<float> VectorJF;

class RobotF {
public:
    RobotF() {}
    RobotF(std::size_t maxSample, const std::string& xmlFilename, int device = 0) : r_(maxSample, xmlFilename, device) {}
    explicit RobotF(const Robot& r) : r_(r) {}
    bool init() const { return r_.init(); }
    explicit operator bool() const { return init(); }

    TransformF calcKinematic(const VectorJF& jv, const float& arcLengthStep);
    TransformF calcKinematic(const VectorJF& jv, const float& arcLengthStep, float& arcLengthStepReturn);
    TransformF calcKinematic(const VectorJF& jv, const std::size_t& nSamples);
    TransformF calcKinematic(const VectorJF& jv, const std::size_t& nSamples, float& arcLengthStepReturn);
    // batched, HOST pointers: jv [B][n_jv] float in, tip [B][12] float out
    int calcKinematicBatch(const float* jv, std::int64_t B, const std::size_t& nSamples, float* tip,
                           std::int32_t* nOut = nullptr);

    std::size_t getNSections() const { return r_.getNSections(); }
    std::size_t getNTubes() const { return r_.getNTubes(); }
    std::size_t getNJointValues() const { return r_.getNJointValues(); }
    float       getMaxLength() const noexcept { return (float)r_.getMaxLength(); }
    std::size_t getMaxSamples() const noexcept { return r_.getMaxSamples(); }
    std::size_t getNSamples() const noexcept { return nsamples_; }
    bool        setMaxSample(std::size_t n) { return r_.setMaxSample(n); }
    const Robot& asDouble() const { return r_; }

private:
    int solve(const float* jv, std::int64_t B, int mode, double step, std::size_t ns, float* tip, std::int32_t* nOut,
              double* stepOut);
    Robot       r_;
    std::size_t nsamples_ = 0;
};

}  // namespace ctrb200
