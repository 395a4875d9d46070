// facade_scale.cpp — CPU harness for the scale policies of the host facade (host/ctr_transrotscale.hpp): the classes a
// caller of ctrb200::Robot::getRandomJointValues(.., const TransRotScale&) hands in, driven with caller-supplied
// uniform variates so that tests/test_reference_build.py can compare them with the reference's own
// GammaCorrection / ProportionalScale objects (oracle/_ref) bit for bit.  Test infrastructure; no GPU involved.
#include <cstdint>
#include <vector>

#include "../../ram2017-ctr-kinematics_b200/host/ctr_transrotscale.hpp"

using namespace ctrb200;

template <int LUT>
static void gamma_run(double gamma, const double* u, int64_t n, double* out)
{
    GammaCorrection<LUT> g;
    g.set_scale(gamma);
    const TransRotScale& ts = g;                 // through the interface, as Robot::getRandomJointValues calls it
    for (int64_t i = 0; i < n; ++i) out[i] = ts.scaleTrans(u[i]);
}

extern "C" int facade_gamma_scale(int lut, double gamma, const double* u, int64_t n, double* out)
{
    switch (lut) {
        case 0: gamma_run<0>(gamma, u, n, out); return 0;
        case 16: gamma_run<16>(gamma, u, n, out); return 0;
        case 64: gamma_run<64>(gamma, u, n, out); return 0;
        case 256: gamma_run<256>(gamma, u, n, out); return 0;
        case 1024: gamma_run<1024>(gamma, u, n, out); return 0;
        default: return -1;
    }
}

// u [B][S] -> out [B][S]; *first_share = the share drawn for configuration 0 (std::rand inside set_scale)
extern "C" int facade_proportional_scale(const double* lengths, int S, double p0, double p1, const double* u, int64_t B,
                                         double* out, double* first_share)
{
    ProportionalScale ps;
    if (!ps.initialize(std::vector<double>(lengths, lengths + S))) return -1;
    ps.set_scale(p0, p1);
    if (first_share) *first_share = ps.get_scale();
    const TransRotScale& ts = ps;
    for (int64_t i = 0; i < B; ++i)
        for (int s = 0; s < S; ++s) out[i * S + s] = ts.scaleTrans(u[i * S + s]);
    return 0;
}

// the bookkeeping calls of the interface (transrotscale.h:33-38)
extern "C" int facade_scale_queue(double* out6)
{
    GammaCorrection<0> g;
    g.add_scale(2.0); g.add_scale(3.0); g.add_scale(4.0);
    out6[0] = (double)g.size_scalefactor(); out6[1] = (double)g.get_iter_position(); out6[2] = g.get_scale();
    g.next_scale(); g.next_scale();
    out6[3] = (double)g.get_iter_position(); out6[4] = g.get_scale(); out6[5] = g.scaleRot(0.75);
    return 0;
}
