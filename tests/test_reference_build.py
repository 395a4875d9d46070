"""The oracle against THE REFERENCE ITSELF.

oracle/_ref/libctr_ref.so is the unmodified /root/reference/ctr_source (robot_i.h, section_i.h,
ctr_static.h, ctr_robot.cpp ...) compiled where it lies against stand-in headers for its absent
third-party dependencies (oracle/shim: Eigen, Erl, Boost.PropertyTree) and called through its public
class CTR::Robot<double>.  These tests pin the C oracle (oracle/ctr_oracle.c) to it BIT FOR BIT:
sample counts, returned steps, every frame of every sample, radii, Jacobians, stability verdicts.
What stays assumed is listed in oracle/shim/mini_erl.h ([ASSUMED] items: Erl's Transform/Rotmat
constructors and product, Erl::equal's tolerance, nearestOrthogonal).

Where the reference is not available (no /root/reference and no prebuilt library) the module is
skipped; tests/test_oracle.py::test_reference_golden then still checks the committed outputs of
this same library (tests/golden/reference_golden.npz)."""
import os

import numpy as np
import pytest

import util

ref_lib = pytest.importorskip("ref_lib")
pytestmark = pytest.mark.skipif(not ref_lib.available(), reason="reference sources and prebuilt oracle/_ref both absent")

NAMES = ["ctr_sec2_tub3.xml", "ctr_sec3_tub3.xml", "ctr_sec3_tub4.xml", "ctr_sec3_tub5.xml"]


@pytest.fixture(scope="module")
def refs(ctrfk, robots):
    return {n: ref_lib.RefRobot.from_desc(robots[n]) for n in NAMES}


def assert_same_fk(o, r, maxs, tip_defined=None):
    """tip_defined: configurations whose tip sample has its sin/cos computed by the reference (all, except in
    capped step mode)."""
    assert np.array_equal(o["n"], r["n"])
    assert np.array_equal(o["step"], r["step"])
    ok = np.ones(len(o["n"]), bool) if tip_defined is None else tip_defined
    assert np.array_equal(o["tip"][ok], r["tip"][ok])
    k = np.arange(maxs)[None, :]
    live = (k < o["n"][:, None]) & (ok[:, None] | (k < o["n"][:, None] - 1))
    assert np.array_equal(o["backbone"][live], r["backbone"][live])
    assert np.array_equal(o["radius"][k < o["n"][:, None]], r["radius"][k < o["n"][:, None]])


@pytest.mark.parametrize("name", NAMES)
def test_fk_bit_equal_random(ctrfk, oracle, robots, refs, name):
    d, ref = robots[name], refs[name]
    assert (ref.nsections, ref.ntubes, ref.njoints, ref.max_samples) == (d.nsections, d.ntubes, d.njoints, 256)
    assert ref.max_length == d.max_length
    jv = ctrfk.sample_joints_host(d, 20171, 0, 3000)
    for mode, kw in ((ctrfk.MODE_NSAMPLE, dict(nsamples=256)), (ctrfk.MODE_NSAMPLE, dict(nsamples=64)),
                     (ctrfk.MODE_NSAMPLE, dict(nsamples=3)), (ctrfk.MODE_NSAMPLE, dict(nsamples=1)),
                     # (NSamples > MaxSamples is undefined in the reference: topRows(_NSamples), robot_i.h:731-732,
                     #  runs past the MaxSamples rows of m_AlphaInnermostTube into the curvature samples)
                     (ctrfk.MODE_STEP, dict(step=d.max_length / 256.0)), (ctrfk.MODE_STEP, dict(step=d.max_length / 64.0)),
                     (ctrfk.MODE_STEP, dict(step=d.max_length / 1000.0)),       # step mode capped: h = ArcEnd / MaxSamples
                     (ctrfk.MODE_STEP, dict(step=1.0))):                        # one sample
        o = oracle.fk_batch(d, jv, mode, backbone=True, radius=True, **kw)
        r = ref.fk_batch(jv, mode, backbone=True, radius=True, **kw)
        defined = None
        if mode == ctrfk.MODE_STEP and kw["step"] < d.max_length / 256.0:
            # Capped step mode (robot_i.h:1080-1081): SampleMax = size_t(ArcEnd / h') (:700) can come out as
            # N - 1, and then the reference never computes sin/cos of the tip sample's angle — it reads whatever
            # an earlier call left there.  The oracle (and the kernels) always compute them: the one deliberate
            # deviation.  Every other sample is still compared bit for bit.
            arc_end = np.zeros(len(jv))
            for i in util.phi_indices(d):
                arc_end = arc_end + jv[:, i]
            defined = np.maximum(1, np.floor(arc_end / r["step"])) >= r["n"]
        assert_same_fk(o, r, 256, defined)


def test_capped_step_mode_only_differs_in_the_stale_tip_sample(ctrfk, oracle):
    """MaxSamples not a power of two: ArcEnd / (ArcEnd / N) rounds below N for a share of the inputs, the
    reference then leaves the tip sample's sin/cos stale; everything else is bit-equal."""
    d = ctrfk.desc_from_xml(ctrfk.resource("ctr_sec3_tub3.xml"), 100)
    ref = ref_lib.RefRobot.from_desc(d)
    jv = ctrfk.sample_joints_host(d, 8, 0, 4000)
    step = d.max_length / 1000.0
    o = oracle.fk_batch(d, jv, ctrfk.MODE_STEP, step=step, backbone=True, radius=True)
    r = ref.fk_batch(jv, ctrfk.MODE_STEP, step=step, backbone=True, radius=True)
    arc_end = np.zeros(len(jv))
    for i in util.phi_indices(d):
        arc_end = arc_end + jv[:, i]
    defined = np.maximum(1, np.floor(arc_end / r["step"])) >= r["n"]
    assert 0.05 < 1.0 - defined.mean() < 0.95
    assert_same_fk(o, r, 100, defined)


@pytest.mark.parametrize("name", NAMES)
def test_fk_bit_equal_edge_sets(ctrfk, oracle, robots, refs, name):
    d, ref = robots[name], refs[name]
    jv = util.edge_joint_sets(d)
    for mode, kw in ((ctrfk.MODE_NSAMPLE, dict(nsamples=256)), (ctrfk.MODE_NSAMPLE, dict(nsamples=7)),
                     (ctrfk.MODE_STEP, dict(step=d.max_length / 256.0))):
        o = oracle.fk_batch(d, jv, mode, backbone=True, radius=True, **kw)
        r = ref.fk_batch(jv, mode, backbone=True, radius=True, **kw)
        live = np.isfinite(r["tip"]).all(axis=1) & np.isfinite(o["tip"]).all(axis=1)
        # all-phi = 0 in NSamples mode has h = 0/N = 0 and a 0/0 in calcTubeRadius (undefined in the
        # reference, saturated in the oracle): frames are compared, radii only where h > 0
        assert np.array_equal(o["n"], r["n"]) and np.array_equal(o["step"], r["step"])
        assert np.array_equal(o["tip"][live], r["tip"][live]) and live.sum() >= len(jv) - 4
        sane = live & (o["step"] > 0)
        lv = (np.arange(256)[None, :] < o["n"][:, None]) & sane[:, None]
        assert np.array_equal(o["backbone"][lv], r["backbone"][lv])
        assert np.array_equal(o["radius"][lv], r["radius"][lv])


@pytest.mark.parametrize("layout", util.ALL_LAYOUTS, ids=["-".join(map(str, l)) for l in util.ALL_LAYOUTS])
def test_fk_bit_equal_all_layouts(ctrfk, oracle, layout):
    """Every section layout the reference's templates can instantiate (CTR_MAXSECTION = 3), random
    parameters with distinct Poisson ratios per tube (the last-VC-section quirk, section_i.h:273)."""
    rng = np.random.default_rng(1000 + sum(3 ** i * n for i, n in enumerate(layout)))
    d = util.random_robot(ctrfk, rng, layout)
    ref = ref_lib.RefRobot.from_desc(d)
    jv = ctrfk.sample_joints_host(d, 99, 0, 500)
    for mode, kw in ((ctrfk.MODE_NSAMPLE, dict(nsamples=128)), (ctrfk.MODE_STEP, dict(step=d.max_length / 100.0))):
        o = oracle.fk_batch(d, jv, mode, backbone=True, radius=True, **kw)
        r = ref.fk_batch(jv, mode, backbone=True, radius=True, **kw)
        assert_same_fk(o, r, d.max_samples)


def test_step_mode_dropped_tip_sample_agrees(ctrfk, oracle, robots):
    """rounding of the position-driven forward loop can drop the tip sample; the sample
    count of the oracle equals the reference's on every configuration, dropped or not."""
    d = ctrfk.desc_from_xml(ctrfk.resource("ctr_sec3_tub4.xml"), 64)
    ref = ref_lib.RefRobot.from_desc(d)
    jv = ctrfk.sample_joints_host(d, 5, 0, 20000)
    step = d.max_length / 64.0
    o = oracle.fk_batch(d, jv, ctrfk.MODE_STEP, step=step)
    r = ref.fk_batch(jv, ctrfk.MODE_STEP, step=step)
    assert np.array_equal(o["n"], r["n"]) and np.array_equal(o["tip"], r["tip"])
    full = np.maximum(1, np.floor((jv[:, 1] + jv[:, 3] + jv[:, 6]) / step)).astype(int)
    assert (r["n"] < np.minimum(full, 64)).sum() > 20          # the quirk does occur in this sample


@pytest.mark.parametrize("name", NAMES)
def test_jacobian_bit_equal(ctrfk, oracle, robots, refs, name):
    d, ref = robots[name], refs[name]
    jv = ctrfk.sample_joints_host(d, 31, 0, 40)
    jv[0, 1] = d.sec[0].length                  # phi + dphi >= L: backward difference (robot_i.h:948-950)
    for q in jv:
        jo, to = oracle.jacobian(d, q, 128, 1e-5, 1e-4)
        jr, tr = ref.jacobian(q, 128, 1e-5, 1e-4)
        assert np.array_equal(to, tr)
        assert np.array_equal(jo, jr)


@pytest.mark.parametrize("name", NAMES)
def test_stability_bit_equal(ctrfk, oracle, robots, refs, name):
    d, ref = robots[name], refs[name]
    jv = ctrfk.sample_joints_host(d, 77, 0, 60)
    step = d.max_length / 128.0
    verdicts = []
    for q in jv:
        so, sr = oracle.stability(d, q, 0.2, step), ref.stability(q, 0.2, step)
        assert so == sr
        verdicts.append(sr)
        for min_deg in (-1.0, 0.5):
            (do, sto), (dr, str_) = oracle.degree_stability(d, q, 0.2, step, min_deg), ref.degree_stability(q, 0.2, step, min_deg)
            assert sto == str_ and (do == dr or (np.isnan(do) and np.isnan(dr)))
    assert 0 < sum(verdicts) < len(verdicts)       # both verdicts occur


def test_xml_readers_agree(ctrfk, oracle, robots):
    """ctr_fk_desc_from_xml against the reference's own readers (ctr_static.h:88-259 V1, :420-580 V2):
    same structure, same parameters (through identical kinematics), entry frame equal to rounding of
    the assumed nearestOrthogonal, and the construction-time FK call (ctr_static.h:248-252)."""
    for n in NAMES + ["sec3_tub4_v2.xml"]:
        path = ctrfk.resource(n)
        d, phi = ctrfk.desc_from_xml(path, 256, want_phi=True)
        ref = ref_lib.RefRobot.from_xml(path, 256)
        assert (ref.nsections, ref.ntubes, ref.njoints) == (d.nsections, d.ntubes, d.njoints)
        assert ref.max_length == d.max_length
        # state left by the constructor: JV = [0, phi_xml, 0 ...], step = MaxLength / MaxSamples
        n_c, tip_c, jv_c = ref.last_state()
        want = np.zeros(ref.njoints)
        for i, s in enumerate(util.phi_indices(d)):
            want[s] = phi[i]
        assert np.array_equal(jv_c, want)
        oc = oracle.fk_one(d, want, ctrfk.MODE_STEP, step=d.max_length / 256.0)
        assert oc["n"] == n_c == 3 or n != "ctr_sec2_tub3.xml"
        assert oc["n"] == n_c and np.abs(oc["tip"] - tip_c).max() < 1e-15
        jv = ctrfk.sample_joints_host(d, 3, 0, 500)
        o = oracle.fk_batch(d, jv, ctrfk.MODE_NSAMPLE, nsamples=200, backbone=True, radius=True)
        r = ref.fk_batch(jv, ctrfk.MODE_NSAMPLE, nsamples=200, backbone=True, radius=True)
        assert np.array_equal(o["n"], r["n"]) and np.array_equal(o["step"], r["step"]) and np.array_equal(o["radius"], r["radius"])
        assert np.abs(o["backbone"] - r["backbone"]).max() < 1e-15


def test_xml_reader_of_the_reference_on_its_own_resources(ctrfk):
    """In the build container the reference's own ctr_resources/*.xml are read by the reference and by
    ctr_fk_desc_from_xml; the results must equal those of tests/data (generated files, same parameters)."""
    src = "/root/reference/ctr_resources"
    if not os.path.isdir(src):
        pytest.skip("reference resources only exist in the build container")
    for n in NAMES:
        ref = ref_lib.RefRobot.from_xml(os.path.join(src, n), 256)
        d = ctrfk.desc_from_xml(os.path.join(src, n), 256)
        d2 = ctrfk.desc_from_xml(ctrfk.resource(n), 256)
        assert bytes(d) == bytes(d2)
        jv = ctrfk.sample_joints_host(d, 11, 0, 200)
        a = ref.fk_batch(jv, ctrfk.MODE_NSAMPLE, nsamples=256)
        b = ref_lib.RefRobot.from_xml(ctrfk.resource(n), 256).fk_batch(jv, ctrfk.MODE_NSAMPLE, nsamples=256)
        assert np.array_equal(a["tip"], b["tip"])


def test_mt19937_joint_draws(ctrfk, robots, refs):
    """getRandomJointValues with a default std::mt19937 + uniform_real_distribution (ctr_test.cpp:30-32):
    draw order theta, then per section phi and the alphas; scale c = nextafter(2 pi, 0) (robot_i.h:1308-1317,
    section_i.h:328-337).  Replayed here from numpy's MT19937 (same generator, seed 5489) with libstdc++'s
    generate_canonical<double, 53> (two 32-bit draws, low word first)."""
    bg = np.random.MT19937()
    bg._legacy_seeding(5489)
    raw = bg.random_raw(2 * 9 * 50).astype(np.float64)
    u = (raw[0::2] + raw[1::2] * 4294967296.0) / 18446744073709551616.0
    u = np.where(u >= 1.0, np.nextafter(1.0, 0.0), u)
    c = np.nextafter(2 * np.pi, 0.0)
    for n in NAMES:
        d, ref = robots[n], refs[n]
        got = ref.random_joints(50)
        scale = np.full(ref.njoints, c)
        for i, s in enumerate(util.phi_indices(d)):
            scale[s] = d.sec[i].length
        want = (u[:50 * ref.njoints].reshape(50, ref.njoints)) * scale
        assert np.array_equal(got, want)


def sampler_u01(seed, first, B, njv):
    """numpy replica of ctrk::sampler_u01 (host/ctr_sampler.h), checked in tests/test_host.py."""
    def splitmix64(z):
        z = (z + np.uint64(0x9E3779B97F4A7C15))
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))
    with np.errstate(over="ignore"):
        i = (np.arange(B, dtype=np.uint64) + np.uint64(first))[:, None]
        j = np.arange(njv, dtype=np.uint64)[None, :]
        b = splitmix64(splitmix64(np.uint64(seed) + np.uint64(0x9E3779B97F4A7C15) * i) + j)
    return (b >> np.uint64(11)).astype(np.float64) / 9007199254740992.0


def test_scale_policies_against_the_reference_classes(ctrfk, robots, refs):
    """ctr_fk_sample_joints_scaled_host against the reference's own GammaCorrection / ProportionalScale
    (transrotscale.h), fed with the very uniform variates the counter-based sampler draws."""
    name = "ctr_sec3_tub4.xml"
    d, ref = robots[name], refs[name]
    B, seed = 3000, 9
    u = sampler_u01(seed, 0, B, d.njoints)
    phi = util.phi_indices(d)
    lens = np.array([d.sec[s].length for s in range(d.nsections)])
    for lut in (0, 16, 64, 256, 1024):
        got = ctrfk.sample_joints_host(d, seed, 0, B, policy=ctrfk.SCALE_GAMMA, p0=2.5, lut_size=lut)
        want = lens * ref_lib.gamma_scale(lut, 2.5, u[:, phi])          # section_i.h:341: m_Length * scaleTrans(u)
        # the reference calls std::pow; the sampler its own pow (the same bits on host and device, < 1 ulp from the
        # correctly rounded value, ctr_sampler.h) — the two differ by at most 1 ulp of the power: 2 ulp of the product,
        # 4 ulp after the table's interpolation between two such powers (measured: 2 and 3)
        assert (np.abs(got[:, phi] - want) <= (2.0 if lut == 0 else 4.0) * np.spacing(np.abs(want))).all(), lut
        assert (got[:, phi] == want).mean() > 0.85, lut
    # ProportionalScale carries its goal from one configuration to the next (the last-section draw, :286-288);
    # the first goal comes from std::rand() in the reference and from an extra sampler slot here, so configuration 0
    # is compared after handing the reference's first goal to the formula, every later one directly
    p0, p1 = 0.3, 0.6
    got = ctrfk.sample_joints_host(d, seed, 0, B, policy=ctrfk.SCALE_PROPORTIONAL, p0=p0, p1=p1)
    want, first_goal = ref_lib.proportional_scale(ref, p0, p1, u[:, phi])
    assert p0 <= first_goal <= p1
    assert np.array_equal(got[1:, phi], lens * want[1:])


def _load_facade_scale():
    import ctypes as C, subprocess
    d = os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostsim")
    so, src = os.path.join(d, "libfacade_scale.so"), os.path.join(d, "facade_scale.cpp")
    hdr = os.path.join(os.path.dirname(d), "..", "ram2017-ctr-kinematics_b200", "host", "ctr_transrotscale.hpp")
    if (not os.path.exists(so)) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["/usr/bin/g++", "-std=c++14", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", so, src])
    lib = C.CDLL(so)
    lib.facade_gamma_scale.argtypes = [C.c_int, C.c_double, C.c_void_p, C.c_int64, C.c_void_p]
    lib.facade_proportional_scale.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
    lib.facade_scale_queue.argtypes = [C.c_void_p]
    return lib


def test_facade_scale_classes_equal_the_reference_classes(ctrfk, robots, refs):
    """host/ctr_transrotscale.hpp — the GammaCorrection<LookUpSize> / ProportionalScale a caller of
    ctrb200::Robot::getRandomJointValues(.., const TransRotScale&) hands in — against the reference's own classes
    (transrotscale.h, compiled in oracle/_ref), the same uniform variates into both: bit for bit (both call std::pow)."""
    import ctypes as C
    lib = _load_facade_scale()
    name = "ctr_sec3_tub4.xml"
    d, ref = robots[name], refs[name]
    rng = np.random.default_rng(3)
    u = np.ascontiguousarray(rng.random(5000))
    u[:3] = (0.0, 1.0, 0.5)
    for lut in (0, 16, 64, 256, 1024):
        for gamma in (0.5, 2.5):
            got = np.zeros_like(u)
            assert lib.facade_gamma_scale(lut, gamma, u.ctypes.data, u.size, got.ctypes.data) == 0
            want = ref_lib.gamma_scale(lut, gamma, u)
            assert np.array_equal(got, want), (lut, gamma)
    S = d.nsections
    lens = np.array([d.sec[s].length for s in range(S)])
    B = 2000
    us = np.ascontiguousarray(rng.random((B, S)))
    got = np.zeros((B, S)); first = C.c_double(0.0)
    assert lib.facade_proportional_scale(lens.ctypes.data, S, 0.3, 0.6, us.ctypes.data, B, got.ctypes.data, C.addressof(first)) == 0
    want, first_ref = ref_lib.proportional_scale(ref, 0.3, 0.6, us)
    assert 0.3 <= first.value <= 0.6 and 0.3 <= first_ref <= 0.6
    assert np.array_equal(got[1:], want[1:])            # configuration 0 takes its goal from std::rand() in both
    assert ((got >= 0.0) & (got <= 1.0)).all()
    q = np.zeros(6)
    lib.facade_scale_queue(q.ctypes.data)
    assert list(q) == [3.0, 0.0, 2.0, 2.0, 4.0, 0.25]


@pytest.mark.parametrize("name", NAMES)
def test_jacobian_other_forms_bit_equal(ctrfk, oracle, robots, refs, name):
    """The ArcLengthStep form (robot_i.h:901-907: base pose from step mode, perturbed solves in NSamples mode with
    N = m_Samples — the reference mixes the two discretisations, reproduced) and the iSample form (:908-915,
    :1012-1072: Jacobian of an intermediate sample's frame next to the tip's)."""
    d, ref = robots[name], refs[name]
    jv = ctrfk.sample_joints_host(d, 41, 0, 24)
    jv[0, 1] = d.sec[0].length
    for q in jv:
        o = oracle.jacobian_ex(d, q, ctrfk.MODE_STEP, step=d.max_length / 200.0)
        jr, tr, nr = ref.jacobian_step(q, d.max_length / 200.0, 1e-5, 1e-4)
        assert o["n"] == nr and np.array_equal(o["tip"], tr) and np.array_equal(o["jac"], jr)
        for i_sample in (0, 37, 99):
            o = oracle.jacobian_ex(d, q, ctrfk.MODE_NSAMPLE, nsamples=100, i_sample=i_sample)
            jt, js = ref.jacobian_sample(q, 100, i_sample, 1e-5, 1e-4)
            assert np.array_equal(o["jac"], jt) and np.array_equal(o["jac_sample"], js)
        # the tip is the last sample
        assert np.array_equal(o["jac"], o["jac_sample"])


def test_adapter_from_live_reference_robot(ctrfk, oracle, robots):
    """include/ctr_fk_reference_adapter.hpp: the descriptor of a live CTR::Robot<double>, through its public
    interface only (sections from getSection, entry frame observed through one FK call of the all-zero joint
    vector).  Equals the descriptor of ctr_fk_desc_from_xml: sections exactly, entry frame to rounding of the
    orthogonalisation; robots built from sections get their entry frame back bit for bit."""
    for n in NAMES + ["sec3_tub4_v2.xml"]:
        d = ctrfk.desc_from_xml(ctrfk.resource(n), 200)
        a = ref_lib.RefRobot.from_xml(ctrfk.resource(n), 200).adapter_desc(ctrfk.RobotDesc)
        assert (a.nsections, a.max_samples) == (d.nsections, 200)
        assert bytes(a.sec) == bytes(d.sec)
        assert np.abs(np.array(list(a.init_trafo)) - np.array(list(d.init_trafo))).max() < 1e-15
        assert ctrfk.load().ctr_fk_desc_validate(a) == 0
    rng = np.random.default_rng(5)
    for layout in util.ALL_LAYOUTS:
        d = util.random_robot(ctrfk, rng, layout)
        a = ref_lib.RefRobot.from_desc(d).adapter_desc(ctrfk.RobotDesc)
        assert list(a.init_trafo) == list(d.init_trafo)
        for s in range(d.nsections):      # a CC section carries nothing in its [1] slots
            k = d.sec[s].ntube
            for f in ("stiffness", "curvature", "radius", "poisson"):
                assert list(getattr(a.sec[s], f))[:k] == list(getattr(d.sec[s], f))[:k]
            assert a.sec[s].ntube == k and a.sec[s].length == d.sec[s].length


def test_f32_variant_next_to_the_reference_float_build(ctrfk, oracle, hostsim, robots):
    """The FP32 variant (per-thread solvers compiled for the CPU) and the reference's own float instantiation
    (CTR::Robot<float>, with libm sinf/cosf standing in for Eigen's SSE approximations), both against the FP64
    oracle: the variant stays inside its stated tolerance (1e-4 m / 1e-3 rad) and is closer to the FP64 result
    than the reference's float build — whose control scalars (arc positions, interval compares, Erl::equal's
    float tolerance) are single precision too, so single samples change section."""
    for name in util.ROBOT_FILES:
        d = robots[name]
        jv = ctrfk.sample_joints_host(d, util.SEEDS[name], 0, 4000)
        o = oracle.fk_batch(d, jv, ctrfk.MODE_NSAMPLE, nsamples=256)
        r = ref_lib.fk_batch_f32(d, jv, ctrfk.MODE_NSAMPLE, nsamples=256)
        dpr, angr = util.pose_errors(o["tip"], r["tip"])
        for fast in (2, 3):
            s = util.hostsim_fk(hostsim, d, jv, ctrfk.MODE_NSAMPLE, nsamples=256, fast=fast)
            dp, ang = util.pose_errors(o["tip"], s["tip"])
            assert dp.max() <= 1e-4 and ang.max() <= 1e-3
            assert dp.mean() < dpr.mean() and dp.max() < dpr.max()


def test_fk_bit_equal_degenerate_robots(ctrfk, oracle):
    """Robots that reach the branches random parameters never do: zero curvature everywhere (the |u| ~ 0 branch of
    curvature2Transform, robot_i.h:1245/:1264-1266, translation h*|u| instead of h — SURVEY §8g-7), one straight
    section among curved ones, Poisson ratio 0, stiffness ratios of 1e6, a 1 mm section; every layout."""
    rng = np.random.default_rng(77)
    cases = 0
    for layout in util.ALL_LAYOUTS:
        for variant in ("all_straight", "one_straight", "nu_zero", "stiff_ratio", "short_section"):
            d = util.random_robot(ctrfk, rng, layout, max_samples=96)
            for s in range(d.nsections):
                for k in range(d.sec[s].ntube):
                    if variant == "all_straight" or (variant == "one_straight" and s == d.nsections - 1):
                        d.sec[s].curvature[k] = 0.0
                    if variant == "nu_zero":
                        d.sec[s].poisson[k] = 0.0
                    if variant == "stiff_ratio":
                        d.sec[s].stiffness[k] = 1e3 if (s + k) % 2 == 0 else 1e-3
                if variant == "short_section" and s == 0:
                    d.sec[s].length = 1e-3
            ref = ref_lib.RefRobot.from_desc(d)
            jv = ctrfk.sample_joints_host(d, 300 + cases, 0, 120)
            for mode, kw in ((ctrfk.MODE_NSAMPLE, dict(nsamples=96)), (ctrfk.MODE_NSAMPLE, dict(nsamples=5)),
                             (ctrfk.MODE_STEP, dict(step=d.max_length / 50.0))):
                o = oracle.fk_batch(d, jv, mode, backbone=True, radius=True, **kw)
                r = ref.fk_batch(jv, mode, backbone=True, radius=True, **kw)
                assert_same_fk(o, r, d.max_samples)
            if variant == "all_straight":      # the quirk itself: every step contributes h*|u| = 0, the robot has no length
                assert np.abs(r["tip"][:, 9:] - np.array(list(d.init_trafo))[9:]).max() == 0.0
            cases += 1
    assert cases == 70
