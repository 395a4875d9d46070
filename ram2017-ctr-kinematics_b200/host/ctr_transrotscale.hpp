// ctr_transrotscale.hpp — the scale policies of the reference's joint sampler (ctr_source/transrotscale.h) for the
// host façade: TransRotScale (interface, :28-41), GammaCorrection<LookUpSize> (:43-135) and ProportionalScale
// (:137-310) with the reference's method names, so that code written against
//     Robot::getRandomJointValues(CT_RNG&, CT_RND&, VectorJ&, const TransRotScale&)        (ctr_robot.h:156)
// keeps compiling.  The batched counterpart is ctr_fk_sample_joints_scaled (CTR_FK_SCALE_GAMMA / _PROPORTIONAL),
// whose counter-based draws make every configuration independent; these classes keep the reference's sequential
// semantics (mutable state advanced by every scaleTrans call) for callers that draw one vector at a time.
// tests/test_reference_build.py feeds both implementations the same variates and compares them bit for bit.
#pragma once

#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdlib>
#include <utility>
#include <vector>

namespace ctrb200 {

class Robot;

class TransRotScale {
public:
    virtual ~TransRotScale() {}
    virtual void        set_scale(const double& yfactor) = 0;
    virtual void        add_scale(const double& yfactor) = 0;
    virtual void        next_scale() = 0;
    virtual std::size_t size_scalefactor() const = 0;
    virtual double      get_scale() const = 0;
    virtual long        get_iter_position() const = 0;
    virtual double      scaleTrans(const double& value) const = 0;      // value in [0, 1] -> share of the section length
    virtual double      scaleRot(const double& value) const = 0;
};

// Extension lengths drawn as u^gamma (gamma > 1 favours retracted robots), exactly or through a LookUpSize-entry
// table interpolated linearly.  The table starts as linspace(0, 1) and every set_scale raises its CURRENT entries to
// the new power (transrotscale.h:56-63) — so a second set_scale compounds; kept, since callers of the reference see it.
template <int LookUpSize>
class GammaCorrection : public TransRotScale {
public:
    GammaCorrection() : gamma_(1.0), pos_(0)
    {
        if (LookUpSize > 0) {
            table_.resize((std::size_t)LookUpSize);
            // Eigen 3.3 setLinSpaced(n, 0, 1): node i = i * (1 / (n - 1)), the last node exactly 1
            const double st = LookUpSize > 1 ? 1.0 / (double)(LookUpSize - 1) : 0.0;
            for (int i = 0; i < LookUpSize; ++i) table_[(std::size_t)i] = (i == LookUpSize - 1) ? 1.0 : (double)i * st;
        }
    }
    void set_scale(const double& g) override
    {
        gamma_ = g;
        for (double& v : table_) v = std::pow(v, g);
    }
    void add_scale(const double& g) override
    {
        if (queue_.empty()) { gamma_ = g; pos_ = 0; }
        queue_.push_back(g);
    }
    void        next_scale() override { ++pos_; set_scale(queue_.at(pos_)); }
    std::size_t size_scalefactor() const override { return queue_.size(); }
    double      get_scale() const override { return gamma_; }
    long        get_iter_position() const override { return (long)pos_; }
    double      scaleTrans(const double& u) const override
    {
        if (LookUpSize <= 0) return std::pow(u, gamma_);
        const double x  = u * (double)(LookUpSize - 1);
        const int    lo = (int)std::floor(x), hi = (int)std::ceil(x);
        const double w  = x - (double)lo;
        return table_[(std::size_t)hi] * w + table_[(std::size_t)lo] * (1.0 - w);
    }
    double scaleRot(const double& u) const override { return u - 0.5; }

private:
    double              gamma_;
    std::vector<double> queue_;
    std::size_t         pos_;
    std::vector<double> table_;
};

// Extension lengths whose SUM is a chosen share of the robot's total length: section by section, the draw is mapped
// into the interval that still lets the remaining sections reach the goal (transrotscale.h:273-291).  The share is
// redrawn from [min, max] for every configuration — from the LAST section's variate of the previous one (:286-288).
class ProportionalScale : public TransRotScale {
public:
    ProportionalScale() : lo_(1.0), hi_(1.0), inst_(1.0), total_(0.0), goal_(0.0), room_(0.0), section_(0), pos_(0) {}
    bool initialize(const Robot& robot);                     // section lengths from the robot; false if it has none
    bool initialize(const std::vector<double>& sectionLengths)      // ... or handed in, base to tip
    {
        if (sectionLengths.empty()) return false;
        lengths_ = sectionLengths;
        total_ = 0.0;
        for (double L : lengths_) total_ += L;
        section_ = 0;
        room_ = total_;
        redraw((double)std::rand() / (double)RAND_MAX);
        return true;
    }
    void set_scale(const double& share) override { lo_ = hi_ = share; goal_ = total_ * share; }
    void set_scale(const double& a, const double& b)
    {
        lo_ = std::min(a, b); hi_ = std::max(a, b);
        redraw((double)std::rand() / (double)RAND_MAX);
    }
    void add_scale(const double& share) override { add_scale(share, share); }
    void add_scale(const double& a, const double& b)
    {
        const std::pair<double, double> pr(std::min(a, b), std::max(a, b));
        if (queue_.empty()) { lo_ = pr.first; hi_ = pr.second; pos_ = 0; redraw((double)std::rand() / (double)RAND_MAX); }
        queue_.push_back(pr);
    }
    void next_scale() override
    {
        ++pos_;
        const std::pair<double, double>& pr = queue_.at(pos_);
        lo_ = pr.first; hi_ = pr.second;
        redraw((double)std::rand() / (double)RAND_MAX);
    }
    std::size_t size_scalefactor() const override { return queue_.size(); }
    double      get_scale() const override { return inst_; }
    long        get_iter_position() const override { return (long)pos_; }
    double      scaleTrans(const double& u) const override
    {
        const double L   = lengths_[section_];
        const double vlo = std::max((goal_ - room_ + L) / L, 0.0);
        const double vhi = std::min(goal_ / L, 1.0);
        const double v   = (vhi - vlo) * u + vlo;
        goal_ -= v * L;
        room_ -= L;
        if (++section_ >= lengths_.size()) {                 // configuration complete: next goal from this variate
            section_ = 0;
            room_ = total_;
            inst_ = (hi_ - lo_) * u + lo_;
            goal_ = total_ * inst_;
        }
        return std::min(std::max(v, 0.0), 1.0);
    }
    double scaleRot(const double& u) const override { return u - 0.5; }

private:
    void redraw(double u) { inst_ = (hi_ - lo_) * u + lo_; goal_ = total_ * inst_; }
    double                                 lo_, hi_;
    mutable double                         inst_;
    double                                 total_;
    mutable double                         goal_, room_;
    mutable std::size_t                    section_;
    std::vector<double>                    lengths_;
    std::vector<std::pair<double, double>> queue_;
    std::size_t                            pos_;
};

}  // namespace ctrb200
