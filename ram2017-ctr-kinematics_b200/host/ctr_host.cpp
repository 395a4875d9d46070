// ctr_host.cpp — host-only entry points of libctrfk (no CUDA): robot descriptors from the
// ctr_resources XML files, entry-frame orthogonalisation, the joint sampler and the sample-count
// helper.  C++14, no third-party dependencies (the reference uses Boost property_tree for the
// same job, ctr_static.h:24-25).
#include <cctype>
#include <cerrno>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <memory>
#include <sstream>
#include <string>
#include <utility>
#include <vector>
#include <atomic>

#include "../../include/ctr_fk.h"
#include "../csrc/ctr_fk_internal.h"
#include "../csrc/ctr_fk_krobot.h"
#include "ctr_sampler.h"

namespace ctrk {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// ---- a very small XML reader: elements, text, comments, declarations; attributes skipped ------
struct XmlNode {
    std::string name;
    std::string text;
    std::vector<std::unique_ptr<XmlNode>> kids;

    const XmlNode* child(const std::string& key) const        // exact-case, first match —
    {                                                         // property_tree get() semantics
        for (const auto& k : kids)
            if (k->name == key) return k.get();
        return nullptr;
    }
};

class XmlReader {
public:
    explicit XmlReader(const std::string& s) : s_(s), i_(0) {}

    bool parse(XmlNode& root, std::string& err)
    {
        root.name = "";
        while (true) {
            skip_misc();
            if (i_ >= s_.size()) break;
            if (s_[i_] != '<') { err = "text outside of the root element"; return false; }
            std::unique_ptr<XmlNode> n(new XmlNode);
            if (!element(*n, err)) return false;
            root.kids.push_back(std::move(n));
        }
        return true;
    }

private:
    void skip_ws() { while (i_ < s_.size() && std::isspace((unsigned char)s_[i_])) ++i_; }
    bool starts(const char* lit) const { return s_.compare(i_, std::strlen(lit), lit) == 0; }
    // whitespace, <!-- comments -->, <?declarations?>, <!DOCTYPE ...>
    void skip_misc()
    {
        for (;;) {
            skip_ws();
            if (starts("<!--")) {
                size_t e = s_.find("-->", i_ + 4);
                i_ = (e == std::string::npos) ? s_.size() : e + 3;
            } else if (starts("<?")) {
                size_t e = s_.find("?>", i_ + 2);
                i_ = (e == std::string::npos) ? s_.size() : e + 2;
            } else if (starts("<!")) {
                size_t e = s_.find('>', i_ + 2);
                i_ = (e == std::string::npos) ? s_.size() : e + 1;
            } else {
                return;
            }
        }
    }
    bool element(XmlNode& n, std::string& err)
    {
        ++i_;  // '<'
        size_t b = i_;
        while (i_ < s_.size() && !std::isspace((unsigned char)s_[i_]) && s_[i_] != '>' && s_[i_] != '/') ++i_;
        n.name = s_.substr(b, i_ - b);
        if (n.name.empty()) { err = "empty tag name"; return false; }
        // skip attributes
        bool self_closed = false;
        while (i_ < s_.size() && s_[i_] != '>') {
            if (s_[i_] == '"' || s_[i_] == '\'') {
                char q = s_[i_++];
                while (i_ < s_.size() && s_[i_] != q) ++i_;
            }
            if (s_[i_] == '/' && i_ + 1 < s_.size() && s_[i_ + 1] == '>') self_closed = true;
            ++i_;
        }
        if (i_ >= s_.size()) { err = "unterminated tag <" + n.name + ">"; return false; }
        ++i_;  // '>'
        if (self_closed) return true;
        for (;;) {
            // text up to the next '<'
            size_t t = s_.find('<', i_);
            if (t == std::string::npos) { err = "missing </" + n.name + ">"; return false; }
            n.text.append(s_, i_, t - i_);
            i_ = t;
            if (starts("<!--")) {
                size_t e = s_.find("-->", i_ + 4);
                if (e == std::string::npos) { err = "unterminated comment"; return false; }
                i_ = e + 3;
                continue;
            }
            if (starts("</")) {
                size_t e = s_.find('>', i_);
                if (e == std::string::npos) { err = "unterminated closing tag"; return false; }
                std::string close = s_.substr(i_ + 2, e - i_ - 2);
                while (!close.empty() && std::isspace((unsigned char)close.back())) close.pop_back();
                if (close != n.name) { err = "mismatched </" + close + "> for <" + n.name + ">"; return false; }
                i_ = e + 1;
                break;
            }
            if (starts("<?") || starts("<!")) { skip_misc(); continue; }
            std::unique_ptr<XmlNode> k(new XmlNode);
            if (!element(*k, err)) return false;
            n.kids.push_back(std::move(k));
        }
        // trim_whitespace
        size_t a = 0, z = n.text.size();
        while (a < z && std::isspace((unsigned char)n.text[a])) ++a;
        while (z > a && std::isspace((unsigned char)n.text[z - 1])) --z;
        n.text = n.text.substr(a, z - a);
        return true;
    }

    const std::string& s_;
    size_t i_;
};

static std::string lower(std::string s)
{
    for (auto& ch : s) ch = (char)std::tolower((unsigned char)ch);
    return s;
}

// get<Real>(key): first exact-case child, whole text must parse as a number
static bool get_real(const XmlNode& n, const char* key, double& out, std::string& err)
{
    const XmlNode* c = n.child(key);
    if (!c) { err = std::string("No such node (") + key + ")"; return false; }
    const char* b = c->text.c_str();
    char* e = nullptr;
    errno = 0;
    double v = std::strtod(b, &e);
    while (e && *e && std::isspace((unsigned char)*e)) ++e;
    if (e == b || (e && *e)) { err = std::string("conversion of <") + key + "> to a number failed"; return false; }
    out = v;
    return true;
}
static bool get_real_or(const XmlNode& n, const char* key, double dflt, double& out, std::string& err)
{
    if (!n.child(key)) { out = dflt; return true; }
    return get_real(n, key, out, err);
}

// 3x3 polar factor by Newton iteration X <- (X + X^-T)/2.  R is column-major 3x3 (first 9
// entries of a transform).  Returns false for a singular or reflecting matrix.
static bool polar_rotation(double* R)
{
    double X[9];
    std::memcpy(X, R, sizeof X);
    for (int it = 0; it < 64; ++it) {
        // inverse-transpose = cofactor matrix / det; cyclic form of the cofactors,
        // X(i,j) = X[i + 3j]
        double cof[9];
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
                const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
                cof[i + 3 * j] = X[i1 + 3 * j1] * X[i2 + 3 * j2] - X[i1 + 3 * j2] * X[i2 + 3 * j1];
            }
        double det = X[0] * cof[0] + X[3] * cof[3] + X[6] * cof[6];
        if (!(det > 1e-12)) return false;
        double delta = 0.0;
        for (int i = 0; i < 9; ++i) {
            double nx = 0.5 * (X[i] + cof[i] / det);
            delta = std::fmax(delta, std::fabs(nx - X[i]));
            X[i] = nx;
        }
        if (delta < 1e-16) break;
    }
    std::memcpy(R, X, sizeof X);
    return true;
}

static double ortho_defect(const double* R)
{
    double worst = 0.0;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double d = R[3 * i] * R[3 * j] + R[3 * i + 1] * R[3 * j + 1] + R[3 * i + 2] * R[3 * j + 2];
            worst = std::fmax(worst, std::fabs(d - (i == j ? 1.0 : 0.0)));
        }
    return worst;
}

// host replica of ctrk::setup_control (same header, same operations)
template <class L>
void count_samples_t(const KRobot& rb, const double* jv, int64_t B, int mode, double step, int32_t ns, int32_t* n)
{
    for (int64_t i = 0; i < B; ++i) {
        Ctl<L> c;
        setup_control<L>(rb, jv + i * L::NJ, mode, step, ns, c);
        n[i] = c.nframes;
    }
}

}  // namespace ctrk

using namespace ctrk;

extern "C" {

const char* ctr_fk_last_error(void) { return g_err; }
const char* ctr_fk_version(void) { return "ctr-fk-b200 0.1 (sm_100a)"; }
int64_t ctr_fk_launch_count(void) { return (int64_t)g_launches.load(); }

int ctr_fk_orthogonalize(double* t, double verify_tol)
{
    if (!t) { set_error("ctr_fk_orthogonalize: null transform"); return CTR_FK_E_NULL; }
    if (verify_tol > 0.0 && !(ortho_defect(t) <= verify_tol)) {
        set_error("InitTrafo: not a transformation matrix (|R^T R - I| = %.3g > %.3g)", ortho_defect(t), verify_tol);
        return CTR_FK_E_ARG;
    }
    if (!polar_rotation(t)) {
        set_error("InitTrafo: rotation part is singular or a reflection");
        return CTR_FK_E_ARG;
    }
    return 0;
}

int ctr_fk_desc_validate(const ctr_robot_desc* d)
{
    int rc = validate_desc(d);
    if (rc) set_error("invalid robot descriptor (code %d)", rc);
    return rc;
}
int ctr_fk_desc_ntubes(const ctr_robot_desc* d) { return validate_desc(d) ? CTR_FK_E_DESC : desc_ntubes(d); }
int ctr_fk_desc_njoints(const ctr_robot_desc* d)
{
    return validate_desc(d) ? CTR_FK_E_DESC : 1 + d->nsections + desc_ntubes(d);
}
double ctr_fk_desc_max_length(const ctr_robot_desc* d)
{
    if (validate_desc(d)) return NAN;
    double len = 0.0;                                 // robot_i.h:537-543
    for (int s = 0; s < d->nsections; ++s) len += d->sec[s].length;
    return len;
}

int ctr_fk_desc_from_xml(const char* path, int32_t max_samples, ctr_robot_desc* out, double* phi_xml)
{
    if (!path || !out) { set_error("ctr_fk_desc_from_xml: null argument"); return CTR_FK_E_NULL; }
    if (max_samples < 1) { set_error("ctr_fk_desc_from_xml: max_samples must be >= 1"); return CTR_FK_E_ARG; }
    std::ifstream f(path, std::ifstream::in);
    if (!f.good()) {                                   // ctr_static.h:97-101
        set_error("concentricTubefromXML: could not read file: %s", path);
        return CTR_FK_E_XML;
    }
    std::stringstream ss;
    ss << f.rdbuf();
    const std::string text = ss.str();
    const bool v2 = text.find("version=\"2.0\"") != std::string::npos;   // ctr_static.h:594-601

    XmlNode doc;
    std::string err;
    XmlReader rd(text);
    if (!rd.parse(doc, err)) { set_error("XML parse error in %s: %s", path, err.c_str()); return CTR_FK_E_XML; }
    const XmlNode* robot = doc.child("robot");
    if (!robot) { set_error("XML %s: no <robot> element", path); return CTR_FK_E_XML; }

    ctr_robot_desc d;
    std::memset(&d, 0, sizeof d);
    d.max_samples = max_samples;
    d.init_trafo[0] = d.init_trafo[4] = d.init_trafo[8] = 1.0;   // Identity, ctr_static.h:116
    double phis[CTR_FK_MAX_SECTIONS] = {0, 0, 0};
    int nsec = 0;

    for (const auto& v : robot->kids) {
        const std::string tag = lower(v->name);
        if (tag == "section") {
            double stiff[2] = {0, 0}, curv[2] = {0, 0}, rad[2] = {0, 0}, nu[2] = {0, 0};
            double phi[2] = {0, 0}, len[2] = {0, 0};
            double sec_len = 0.0, sec_phi = 0.0;
            if (v2) {                                  // ctr_static.h:479-480
                if (!get_real(*v, "length", sec_len, err) || !get_real_or(*v, "phi", 0.0, sec_phi, err)) {
                    set_error("XML %s: %s", path, err.c_str());
                    return CTR_FK_E_XML;
                }
            }
            int nt = 0;
            for (const auto& v2n : v->kids) {
                if (lower(v2n->name) != "tube") continue;
                if (nt > 1) { set_error("XML %s: VIOLATED: maximum of 2 tubes per section", path); return CTR_FK_E_XML; }
                bool ok = get_real(*v2n, "stiffness", stiff[nt], err) && get_real(*v2n, "curvature", curv[nt], err) &&
                          get_real(*v2n, "radius", rad[nt], err) && get_real(*v2n, "poissonRatio", nu[nt], err);
                if (ok && !v2)                         // ctr_static.h:167-168
                    ok = get_real_or(*v2n, "phi", 0.0, phi[nt], err) && get_real(*v2n, "length", len[nt], err);
                if (!ok) { set_error("XML %s: %s", path, err.c_str()); return CTR_FK_E_XML; }
                ++nt;
            }
            if (nt == 0) continue;                     // neither CC nor VC: ignored by the reference
            if (nsec >= CTR_FK_MAX_SECTIONS) {         // ctr_static.h:191-194
                set_error("XML %s: MAX SECTION: only %d sections are supported", path, CTR_FK_MAX_SECTIONS);
                return CTR_FK_E_XML;
            }
            ctr_section_desc& sd = d.sec[nsec];
            sd.ntube  = nt;
            sd.length = v2 ? sec_len : len[0];         // V1: first tube's length, ctr_static.h:174,183
            for (int k = 0; k < nt; ++k) {
                sd.stiffness[k] = stiff[k]; sd.curvature[k] = curv[k];
                sd.radius[k] = rad[k];      sd.poisson[k] = nu[k];
            }
            phis[nsec] = v2 ? sec_phi : phi[0];        // ctr_static.h:179,188
            ++nsec;
        } else if (tag == "entryframe") {
            static const char* k1[12] = {"x00", "x01", "x02", "x03", "x10", "x11", "x12", "x13", "x20", "x21", "x22", "x23"};
            static const char* k2[12] = {"E11", "E12", "E13", "E14", "E21", "E22", "E23", "E24", "E31", "E32", "E33", "E34"};
            double rm[12];                             // row-major r00 r01 r02 tx ...
            for (int i = 0; i < 12; ++i)
                if (!get_real(*v, v2 ? k2[i] : k1[i], rm[i], err)) { set_error("XML %s: %s", path, err.c_str()); return CTR_FK_E_XML; }
            double cm[12];
            for (int r = 0; r < 3; ++r) {
                for (int c = 0; c < 3; ++c) cm[3 * c + r] = rm[4 * r + c];
                cm[9 + r] = rm[4 * r + 3];
            }
            int rc = ctr_fk_orthogonalize(cm, 0.01);   // ctr_static.h:202-206
            if (rc) return CTR_FK_E_XML;
            std::memcpy(d.init_trafo, cm, sizeof cm);
        }
    }
    if (nsec == 0) { set_error("XML %s: robot has no sections", path); return CTR_FK_E_XML; }
    d.nsections = nsec;
    int rc = validate_desc(&d);
    if (rc) { set_error("XML %s: invalid robot parameters (code %d)", path, rc); return CTR_FK_E_XML; }
    *out = d;
    if (phi_xml) for (int s = 0; s < CTR_FK_MAX_SECTIONS; ++s) phi_xml[s] = phis[s];
    return 0;
}

static int check_policy(int policy, double p0, double p1, int32_t lut)
{
    if (policy < CTR_FK_SCALE_NONE || policy > CTR_FK_SCALE_PROPORTIONAL) return CTR_FK_E_ARG;
    if (!std::isfinite(p0) || !std::isfinite(p1) || lut < 0) return CTR_FK_E_ARG;
    if (policy == CTR_FK_SCALE_GAMMA && !(p0 > 0.0)) return CTR_FK_E_ARG;
    if (policy == CTR_FK_SCALE_PROPORTIONAL && !(p0 >= 0.0 && p1 >= p0 && p1 <= 1.0)) return CTR_FK_E_ARG;
    return 0;
}

int ctr_fk_sample_joints_scaled_host(const ctr_robot_desc* d, uint64_t seed, int64_t first, int64_t B, int policy,
                                     double p0, double p1, int32_t lut_size, double* jv)
{
    if (!d || !jv) { set_error("ctr_fk_sample_joints_host: null argument"); return CTR_FK_E_NULL; }
    if (validate_desc(d) || B < 0 || first < 0) { set_error("ctr_fk_sample_joints_host: bad descriptor, B or offset"); return CTR_FK_E_ARG; }
    if (check_policy(policy, p0, p1, lut_size)) { set_error("bad scale policy parameters"); return CTR_FK_E_ARG; }
    SamplerTable tb;
    make_sampler_table(d, &tb);
    tb.policy = policy; tb.p0 = p0; tb.p1 = p1; tb.lut_size = lut_size;
    for (int64_t i = 0; i < B; ++i) sample_config(tb, seed, (uint64_t)(first + i), jv + i * tb.njv);
    return 0;
}

int ctr_fk_sample_joints_host(const ctr_robot_desc* d, uint64_t seed, int64_t first, int64_t B, double* jv)
{
    return ctr_fk_sample_joints_scaled_host(d, seed, first, B, CTR_FK_SCALE_NONE, 1.0, 1.0, 0, jv);
}

int ctr_fk_count_samples_host(const ctr_robot_desc* d, const double* jv, int64_t B, int mode, double step,
                              int32_t nsamples, int32_t* n)
{
    if (!d || !jv || !n) { set_error("ctr_fk_count_samples_host: null argument"); return CTR_FK_E_NULL; }
    KRobot rb;
    int rc = make_krobot(d, &rb);
    if (rc) { set_error("ctr_fk_count_samples_host: invalid descriptor"); return rc; }
    if (!((mode == CTR_FK_MODE_STEP && step > 0.0 && std::isfinite(step)) || (mode == CTR_FK_MODE_NSAMPLE && nsamples >= 1))) {
        set_error("ctr_fk_count_samples_host: bad mode/step/nsamples");
        return CTR_FK_E_ARG;
    }
    switch (layout_key(d)) {
#define X(S, A, Bq, C) case S * 1000 + A * 100 + Bq * 10 + C: count_samples_t<Lay<S, A, Bq, C>>(rb, jv, B, mode, step, nsamples, n); break;
        CTRK_FOR_EACH_LAYOUT(X)
#undef X
        default: set_error("unsupported section layout"); return CTR_FK_E_DESC;
    }
    return 0;
}

}  // extern "C"
