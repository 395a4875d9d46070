// ctr_test.cpp — smoke binary with the command line of the reference's ctr_source/ctr_test.cpp
// (argv: max samples, XML file).  The reference prints only a random joint vector; this one also
// prints the tip pose of the construction-time FK (ctr_static.h:248-252) and of that joint
// vector, so BASELINE config 1 has something to compare.
#include <cstdio>
#include <iostream>
#include <string>

#include "ctr_robot.hpp"

int main(int argc, char* argv[])
{
    using ctrb200::Robot;
    if (argc > 2) {
        std::size_t samples = std::stoul(argv[1]);
        std::string xml = argv[2];
        Robot robot(samples, xml);
        if (!robot) { std::cerr << "could not construct robot" << std::endl; return 2; }
        std::printf("construction FK: N=%zu tip=(%.17g, %.17g, %.17g)\n", robot.getNSamples(),
                    robot.getTipTranslation().x(), robot.getTipTranslation().y(), robot.getTipTranslation().z());
        ctrb200::CT_RNG rng;          // default seed 5489, ctr_test.cpp:30
        ctrb200::CT_RND rnd;
        ctrb200::VectorJ jv;
        robot.getRandomJointValues(rng, rnd, jv);
        std::cout << "Joint values: ";
        for (double v : jv) std::printf("%.17g ", v);
        std::cout << std::endl;
        ctrb200::Transform tip = robot.calcKinematic(jv, robot.getMaxLength() / (double)robot.getMaxSamples());
        std::printf("FK: N=%zu tip=(%.17g, %.17g, %.17g)\n", robot.getNSamples(), tip.m[9], tip.m[10], tip.m[11]);
        return 0;
    }
    std::cerr << "Not enough input arguments." << std::endl;
    return 1;
}
