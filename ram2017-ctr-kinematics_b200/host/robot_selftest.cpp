// robot_selftest.cpp — exercises ctrb200::Robot (the C++14 mirror of CTR::Robot<double>,
// ctr_source/ctr_robot.h:76-173) on a GPU: construction paths, null-robot behaviour, single and
// batched solves, getters, copy semantics, Jacobian and stability.  Prints "OK" and exits 0 on
// success; any failed check prints the expression and exits 1.  Driven by
// tests/test_gpu_parity.py::test_cpp_robot_selftest.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <deque>
#include <list>
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
        // (the sample form runs the sweep that hands out samples, the tip form the one in units of the step: same
        // arithmetic up to rounding, which the finite difference divides by 1e-5)
        for (int c = 0; c < 6; ++c) for (int r = 0; r < 6; ++r) { CHECK(std::fabs(Jt(r, c) - J(r, c)) < 1e-6); CHECK(std::fabs(Js(r, c) - Jt(r, c)) < 1e-6); }
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

    // getJointValues(VectorJ&) (ctr_robot.h:134)
    {
        Transform tl = robot.calcKinematic(jv, std::size_t(64));
        (void)tl;
        VectorJ got;
        robot.getJointValues(got);
        CHECK(got.size() == 6 && maxdiff(got.data(), jv.data(), 6) == 0.0);
    }

    // container constructors (ctr_robot.h:96-104): CC and VC sections as (section, position) pairs in a vector, a
    // deque and a list build the same robot as the XML file
    {
        ctr_section_desc s0, s1;
        CHECK(robot.getSection(0, &s0) && robot.getSection(1, &s1));
        Transform init; for (int i = 0; i < 12; ++i) init.m[i] = d.init_trafo[i];
        std::vector<std::pair<ctr_section_desc, std::size_t>> ccv = {{s1, 1}}, vcv = {{s0, 0}};
        std::deque<std::pair<ctr_section_desc, std::size_t>> ccd(ccv.begin(), ccv.end()), vcd(vcv.begin(), vcv.end());
        std::list<std::pair<ctr_section_desc, std::size_t>> ccl(ccv.begin(), ccv.end()), vcl(vcv.begin(), vcv.end());
        Robot rv(256, init, ccv, vcv), rd(256, init, ccd, vcd), rl(256, init, ccl, vcl);
        CHECK(rv && rd && rl && rv.getNTubes() == 3 && rl.getNSections() == 2);
        Transform a = rv.calcKinematic(jv, std::size_t(256)), b = rd.calcKinematic(jv, std::size_t(256)),
                  c = rl.calcKinematic(jv, std::size_t(256));
        CHECK(maxdiff(a.m, tip.m, 12) < 1e-12 && maxdiff(b.m, a.m, 12) == 0.0 && maxdiff(c.m, a.m, 12) == 0.0);
        std::vector<std::pair<ctr_section_desc, std::size_t>> bad = {{s1, 0}};          // position 0 taken twice
        Robot rb(256, init, bad, vcv);
        CHECK(!rb);
    }

    // setMaxSample (ctr_robot.h:149): same robot, another sample capacity
    {
        Robot r2(robot);
        CHECK(r2.setMaxSample(64) && r2.getMaxSamples() == 64 && r2.getNSamples() == 0);
        Transform t64 = r2.calcKinematic(jv, std::size_t(256));                          // capped at MaxSamples
        CHECK(r2.getNSamples() == 64);
        Transform w64 = robot.calcKinematic(jv, std::size_t(64));
        CHECK(maxdiff(t64.m, w64.m, 12) < 1e-12);
        CHECK(r2.setMaxSample(512) && r2.getMaxSamples() == 512);
        r2.calcKinematic(jv, std::size_t(512));
        CHECK(r2.getNSamples() == 512);
        CHECK(!r2.setMaxSample(0) && !none.setMaxSample(16));
    }

    // getRandomJointValues with a scale policy (ctr_robot.h:156, section_i.h:338-347)
    {
        GammaCorrection<0> gamma;
        gamma.set_scale(2.0);
        CT_RNG r1(11), r2(11); CT_RND u1, u2;
        VectorJ plain, scaled;
        robot.getRandomJointValues(r1, u1, plain);
        robot.getRandomJointValues(r2, u2, scaled, gamma);
        CHECK(scaled[0] == plain[0] && scaled[2] == plain[2] && scaled[3] == plain[3] && scaled[5] == plain[5]);
        const double L0 = 0.06647, L1 = 0.0667;
        CHECK(std::fabs(scaled[1] - L0 * std::pow(plain[1] / L0, 2.0)) < 1e-17);
        CHECK(std::fabs(scaled[4] - L1 * std::pow(plain[4] / L1, 2.0)) < 1e-17);
        GammaCorrection<64> lut;
        lut.set_scale(2.0);
        CHECK(std::fabs(lut.scaleTrans(0.5) - 0.25) < 1e-3 && lut.scaleTrans(1.0) == 1.0 && lut.scaleRot(0.75) == 0.25);
        gamma.add_scale(2.0); gamma.add_scale(3.0);
        CHECK(gamma.size_scalefactor() == 2 && gamma.get_iter_position() == 0);
        gamma.next_scale();
        CHECK(gamma.get_scale() == 3.0 && gamma.get_iter_position() == 1);

        ProportionalScale prop;
        CHECK(prop.initialize(robot) && !prop.initialize(none));
        prop.set_scale(0.5);                                       // total extension = half the robot's length
        for (int i = 0; i < 20; ++i) {
            VectorJ q;
            robot.getRandomJointValues(r1, u1, q, prop);
            CHECK(q[1] >= 0.0 && q[1] <= L0 && q[4] >= 0.0 && q[4] <= L1);
            CHECK(std::fabs(q[1] + q[4] - 0.5 * (L0 + L1)) < 1e-15);
        }
    }

    // RobotF: the shape of CTR::Robot<float> (ctr_robot.cpp:724-731)
    {
        RobotF rf(256, std::string(argv[1]));
        CHECK(rf && rf.getNJointValues() == 6 && rf.getMaxSamples() == 256);
        VectorJF qf(jv.begin(), jv.end());
        VectorJ qd(qf.begin(), qf.end());                           // the float inputs, exactly
        float hf = 0.f;
        TransformF tf = rf.calcKinematic(qf, std::size_t(256), hf);
        Transform td = robot.calcKinematic(qd, std::size_t(256));
        CHECK(rf.getNSamples() == 256 && std::fabs(hf - 0.09f / 256) < 1e-9f);
        for (int i = 0; i < 9; ++i) CHECK(std::fabs((double)tf.m[i] - td.m[i]) < 5e-4);
        for (int i = 9; i < 12; ++i) CHECK(std::fabs((double)tf.m[i] - td.m[i]) < 5e-5);
        TransformF ts = rf.calcKinematic(qf, 0.13317f / 256.f);
        CHECK(rf.getNSamples() > 100 && rf.getNSamples() <= 256 && ts(2, 3) > 0.f);
        std::vector<float> bf(B * 6), tfb(B * 12);
        for (int i = 0; i < B * 6; ++i) bf[i] = (float)batch[i];
        std::vector<int32_t> nf(B);
        CHECK(rf.calcKinematicBatch(bf.data(), B, std::size_t(128), tfb.data(), nf.data()) == 0);
        for (int i = 0; i < B; i += 97) {
            VectorJ q(bf.begin() + i * 6, bf.begin() + i * 6 + 6);
            Transform t = robot.calcKinematic(q, std::size_t(128));
            CHECK(nf[i] == 128);
            for (int j = 9; j < 12; ++j) CHECK(std::fabs((double)tfb[i * 12 + j] - t.m[j]) < 5e-5);
        }
    }
    std::printf("OK\n");
    return 0;
}
