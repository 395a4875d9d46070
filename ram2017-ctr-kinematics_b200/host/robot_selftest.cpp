// robot_selftest.cpp — exercises ctrb200::Robot (the C++14 mirror of CTR::Robot<double>,
// ctr_source/ctr_robot.h:76-173) on a GPU: construction paths, null-robot behaviour, single and
// batched solves, getters, copy semantics, Jacobian and stability.  Prints "OK" and exits 0 on
// success; any failed check prints the expression and exits 1.  Driven by
// tests/test_gpu_parity.py::test_cpp_robot_selftest.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "ctr_robot.hpp"

#define CHECK(cond) do { if (!(cond)) { std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); return 1; } } while (0)

static double maxdiff(const double* a, const double* b, int n)
{
    double m = 0.0;
    for (int i = 0; i < n; ++i) m = std::fmax(m, std::fabs(a[i] - b[i]));
    return m;
}

int main(int argc, char** argv)
{
    using namespace ctrb200;
    if (argc < 2) { std::printf("usage: robot_selftest <robot.xml>\n"); return 2; }

    // a default-constructed robot answers like the reference's null robot (ctr_robot.cpp:96-106)
    Robot none;
    CHECK(!none && !none.init());
    CHECK(none.getNSections() == 0 && none.getNJointValues() == 0 && none.getMaxSamples() == 0);
    VectorJ dummy(6, 0.0);
    Transform id = none.calcKinematic(dummy, 1e-3);
    CHECK(maxdiff(id.m, Transform::Identity().m, 12) == 0.0);
    Robot missing(64, std::string("/nonexistent/robot.xml"));
    CHECK(!missing);

    Robot robot(256, std::string(argv[1]));
    CHECK(robot && robot.init());
    CHECK(robot.getNSections() == 2 && robot.getNTubes() == 3 && robot.getNJointValues() == 6);
    CHECK(robot.getMaxSamples() == 256);
    CHECK(std::fabs(robot.getMaxLength() - (0.06647 + 0.0667)) < 1e-15);
    CHECK(robot.getNSamples() == 3);                       // construction-time FK, ctr_static.h:248-252

    // single solve: tip == last sample; both overload families agree with their step return
    VectorJ jv = {0.3, 0.04, 0.5, 1.0, 0.05, 2.0};
    double h = 0.0;
    Transform tip = robot.calcKinematic(jv, std::size_t(256), h);
    CHECK(robot.getNSamples() == 256);
    CHECK(std::fabs(h - 0.09 / 256) < 1e-18);
    CHECK(maxdiff(tip.m, robot.getTipTransform().m, 12) == 0.0);
    CHECK(maxdiff(tip.m, robot.getTransformSample(255).m, 12) == 0.0);
    CHECK(robot.getRadiusSample(0) == 0.00125 && robot.getTipRadius() == 0.001);
    Vector3 p = robot.getTipTranslation();
    CHECK(p.x() == tip.m[9] && p.y() == tip.m[10] && p.z() == tip.m[11]);
    CHECK(maxdiff(robot.getJointValues().data(), jv.data(), 6) == 0.0);
    double raw[12];
    robot.calcKinematic(jv.data(), robot.getMaxLength() / 256.0, raw);     // ctr_robot.cpp:166-186
    Transform tstep = robot.calcKinematic(jv, robot.getMaxLength() / 256.0);
    CHECK(maxdiff(raw, tstep.m, 12) == 0.0 && robot.getNSamples() == 173);
    // frames are rigid: R^T R = I
    for (unsigned k = 0; k < robot.getNSamples(); k += 17) {
        Transform f = robot.getTransformSample(k);
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
                double d = f(0, i) * f(0, j) + f(1, i) * f(1, j) + f(2, i) * f(2, j);
                CHECK(std::fabs(d - (i == j)) < 1e-12);
            }
    }

    // batched solve equals the single solves
    const int B = 1000;
    CT_RNG rng(7); CT_RND rnd;
    std::vector<double> batch(B * 6), tips(B * 12);
    std::vector<int32_t> n(B);
    for (int i = 0; i < B; ++i) { VectorJ q; robot.getRandomJointValues(rng, rnd, q); for (int j = 0; j < 6; ++j) batch[i * 6 + j] = q[j]; }
    CHECK(robot.calcKinematicBatch(batch.data(), B, std::size_t(128), tips.data(), n.data()) == 0);
    for (int i = 0; i < B; i += 111) {
        VectorJ q(batch.begin() + i * 6, batch.begin() + i * 6 + 6);
        Transform t = robot.calcKinematic(q, std::size_t(128));
        CHECK(maxdiff(t.m, &tips[i * 12], 12) < 1e-12 && n[i] == 128);
    }

    // the batch call that also returns every sample's frame and radius
    {
        const int Bb = 300;
        const std::size_t ms = robot.getMaxSamples();
        std::vector<double> bb((std::size_t)Bb * ms * 12), rad((std::size_t)Bb * ms), tp((std::size_t)Bb * 12);
        std::vector<int32_t> nn(Bb);
        CHECK(robot.calcKinematicBatch(batch.data(), Bb, std::size_t(100), tp.data(), bb.data(), rad.data(), nn.data()) == 0);
        for (int i = 0; i < Bb; i += 37) {
            VectorJ q(batch.begin() + i * 6, batch.begin() + i * 6 + 6);
            Transform t = robot.calcKinematic(q, std::size_t(100));
            CHECK(nn[i] == 100 && robot.getNSamples() == 100 && maxdiff(t.m, &tp[i * 12], 12) < 1e-12);
            for (unsigned k = 0; k < 100; k += 9) {
                Transform f = robot.getTransformSample(k);
                CHECK(maxdiff(f.m, &bb[((std::size_t)i * ms + k) * 12], 12) < 1e-12);
                CHECK(robot.getRadiusSample(k) == rad[(std::size_t)i * ms + k]);
            }
        }
    }

    // copies own their engine (one arena per copy in the reference, ctr_robot.cpp:30-33)
    Robot copy(robot);
    CHECK(copy && copy.handle() != robot.handle());
    Transform t2 = copy.calcKinematic(jv, std::size_t(256));
    CHECK(maxdiff(t2.m, tip.m, 12) == 0.0);
    Robot assigned;
    assigned = robot;
    CHECK(assigned && assigned.getNTubes() == 3);

    // Jacobian: column 2 is zero, column 0 is the analytic rigid rotation (robot_i.h:928, :956-959)
    MatJacobian J;
    robot.calcJacobian(jv, std::size_t(128), 1e-5, 1e-4, J);
    CHECK(J.cols == 6);
    for (int r = 0; r < 6; ++r) CHECK(J(r, 2) == 0.0);
    const ctr_robot_desc& d = robot.descriptor();
    CHECK(J(3, 0) == d.init_trafo[6] && J(4, 0) == d.init_trafo[7] && J(5, 0) == d.init_trafo[8]);

    // the reference's other calcJacobian forms (ctr_robot.h:121-126)
    {
        MatJacobian Jt, Js, Jg, Jgs, Jstep;
        robot.calcJacobian(jv, std::size_t(128), 1e-5, 1e-4, std::size_t(127), Jt, Js);      // iSample = tip sample
        for (int c = 0; c < 6; ++c) for (int r = 0; r < 6; ++r) { CHECK(Jt(r, c) == J(r, c)); CHECK(Js(r, c) == Jt(r, c) || std::fabs(Js(r, c) - Jt(r, c)) < 1e-6); }
        robot.calcJacobian(jv, std::size_t(128), 1e-5, 1e-4, std::size_t(40), Jt, Js);
        CHECK(Js(3, 0) == d.init_trafo[6] && Js(0, 2) == 0.0 && std::fabs(Js(0, 1) - Jt(0, 1)) > 1e-9);
        Transform tip = robot.calcKinematic(jv, std::size_t(128));
        Transform smp = robot.getTransformSample(40);
        robot.calcJacobian(jv, tip, std::size_t(128), 1e-5, 1e-4, Jg);                       // base pose handed in
        robot.calcJacobian(jv, tip, smp, std::size_t(128), 1e-5, 1e-4, std::size_t(40), Jg, Jgs);
        for (int c = 0; c < 6; ++c) for (int r = 0; r < 6; ++r) { CHECK(std::fabs(Jg(r, c) - Jt(r, c)) < 1e-6); CHECK(std::fabs(Jgs(r, c) - Js(r, c)) < 1e-6); }
        robot.calcJacobian(jv, robot.getMaxLength() / 256.0, 1e-5, 1e-4, Jstep);             // ArcLengthStep form
        CHECK(Jstep.cols == 6 && Jstep(3, 0) == d.init_trafo[6] && Jstep(2, 2) == 0.0);
    }

    // stability: a barely inserted robot is stable
    VectorJ qs = {0.0, 0.005, 0.0, 0.5, 0.005, 0.7};
    CHECK(robot.calcStability(qs, 0.05, robot.getMaxLength() / 256.0) && robot.getStable());
    double deg = robot.calcDegreeStability(qs, 0.05, robot.getMaxLength() / 256.0);
    CHECK(deg > 0.0 && deg < 1.5707963267948966);

    ctr_section_desc sec;
    CHECK(robot.getSection(0, &sec) && sec.ntube == 2 && !robot.getSection(5, &sec));
    std::printf("OK\n");
    return 0;
}
