// ctr_fk_reference_adapter.hpp — ctr_robot_desc from a LIVE robot object of the reference, using only its
// public interface (ctr_robot.h:76-173).  No change to the reference's sources is needed:
//
//     #include "ctr_robot.h"      // the reference
//     #include "section_i.h"      // the reference (Section getters)
//     #include <ctr_fk.h>
//     #include <ctr_fk_reference_adapter.hpp>
//     CTR::Robot<double> robot(256, "ctr_sec2_tub3.xml");
//     ctr_robot_desc d;  ctr_fk_desc_from_reference(robot, &d);
//     ctr_fk_handle gpu; ctr_fk_create(&d, 0, &gpu);
//
// Sections come from getSection(i, const SecCC**/const SecVC**) (robot_i.h:571-590) and the Section getters
// (section_i.h:150-174).  The entry frame m_InitTrafo is protected (robot_i.h:1335) and has no getter, but it
// is observable: with every joint value zero and one sample, the arc length and the step are 0, the single
// SE(3) step is the identity (robot_i.h:1241-1269 with h = 0) and Rz(0) = I, so the only frame of
// calcKinematic(0, NSamples = 1) is InitTrafo * I — exactly the stored, already orthogonalised entry frame
// (ctr_static.h:202-206).  The call runs on a COPY of the robot (calcKinematic mutates its scratch state).
//
// Header-only and templated on the robot type, so libctrfk.so itself never sees a reference header.
// Checked against the reference build in tests/test_reference_build.py::test_adapter_from_live_reference_robot.
#ifndef CTR_FK_REFERENCE_ADAPTER_HPP
#define CTR_FK_REFERENCE_ADAPTER_HPP

#include <cstddef>
#include <cstring>
#include "ctr_fk.h"

template <class RobotT>
int ctr_fk_desc_from_reference(const RobotT& robot, ctr_robot_desc* out)
{
    typedef typename RobotT::VectorJ   VectorJ;
    typedef typename RobotT::Transform Transform;
    if (!out) return CTR_FK_E_NULL;
    if (!robot.init()) return CTR_FK_E_NULL;                       // "FOOL: no robot"
    std::memset(out, 0, sizeof *out);
    const std::size_t ns = robot.getNSections();
    if (ns < 1 || ns > CTR_FK_MAX_SECTIONS) return CTR_FK_E_DESC;
    out->nsections   = static_cast<int32_t>(ns);
    out->max_samples = static_cast<int32_t>(robot.getMaxSamples());
    out->zero_tol    = 0.0;                                        // default: the tolerance of Erl::equal
    for (std::size_t s = 0; s < ns; ++s) {
        const typename CTR::types::SecCC<double>* cc = nullptr;
        const typename CTR::types::SecVC<double>* vc = nullptr;
        robot.getSection(static_cast<unsigned>(s), &cc);
        robot.getSection(static_cast<unsigned>(s), &vc);
        ctr_section_desc& d = out->sec[s];
        if (cc) {
            d.ntube = 1;
            d.length = cc->getLength();
            d.stiffness[0] = cc->getTubeStiffness(0); d.curvature[0] = cc->getTubeCurvature(0);
            d.radius[0] = cc->getRadius(0);           d.poisson[0] = cc->getPossionRatio(0);
        } else if (vc) {
            d.ntube = 2;
            d.length = vc->getLength();
            for (unsigned k = 0; k < 2; ++k) {
                d.stiffness[k] = vc->getTubeStiffness(k); d.curvature[k] = vc->getTubeCurvature(k);
                d.radius[k] = vc->getRadius(k);           d.poisson[k] = vc->getPossionRatio(k);
            }
        } else {
            return CTR_FK_E_DESC;
        }
    }
    RobotT probe(robot);                                           // ctr_robot.cpp:31-33
    if (!probe.init()) return CTR_FK_E_NOMEM;
    VectorJ zero(static_cast<int>(probe.getNJointValues()), 1);
    for (std::size_t j = 0; j < probe.getNJointValues(); ++j) zero[j] = 0.0;
    const Transform t = probe.calcKinematic(zero, std::size_t(1));
    for (int c = 0; c < 3; ++c)
        for (int r = 0; r < 3; ++r) out->init_trafo[r + 3 * c] = t(r, c) + 0.0;      // + 0.0: -0 -> +0
    const auto p = t.getTranslation();
    out->init_trafo[9] = p.x() + 0.0; out->init_trafo[10] = p.y() + 0.0; out->init_trafo[11] = p.z() + 0.0;
    return 0;                                                      // ctr_fk_create validates the descriptor
}

#endif  // CTR_FK_REFERENCE_ADAPTER_HPP
