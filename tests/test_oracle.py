"""Pins the CPU oracle (oracle/ctr_oracle.c).  The reference has no tests or golden vectors of its
own for this path, so the pins are: (0) outputs of the reference's own code, built from its sources
against stand-in third-party headers (tests/golden/reference_golden.npz, test_reference_golden;
live comparison in tests/test_reference_build.py), (a) SURVEY §9.2 sanity vectors from an
independent transliteration, (b) a second independent restatement in pure Python
(oracle/ctr_oracle_py.py), (c) closed-form known-answer tests K1-K6 of SURVEY §4, (d) frozen
regression vectors of the oracle."""
import copy
import json
import os

import numpy as np
import pytest

import util

HERE = os.path.dirname(os.path.abspath(__file__))
IDENT = [1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0]


def with_identity(ctrfk, d):
    c = ctrfk.RobotDesc.from_buffer_copy(d)
    for i in range(12):
        c.init_trafo[i] = IDENT[i]
    return c


def test_survey_vectors(ctrfk, oracle, robots):
    sv = json.load(open(os.path.join(HERE, "golden", "survey_vectors.json")))
    for case in sv["cases"]:
        d = with_identity(ctrfk, robots[case["xml"]])
        mode = ctrfk.MODE_NSAMPLE if case["mode"] == "F" else ctrfk.MODE_STEP
        r = oracle.fk_one(d, case["jv"], mode, step=d.max_length / 256.0, nsamples=256)
        assert r["n"] == case["n"]
        assert r["step"] == case["h"]
        assert np.abs(r["tip"][9:] - np.array(case["tip"])).max() < 1e-15
    for xml, want in sv["base_angles_F"].items():
        d = with_identity(ctrfk, robots[xml])
        jv = [c["jv"] for c in sv["cases"] if c["xml"] == xml][0]
        r = oracle.fk_one(d, jv, ctrfk.MODE_NSAMPLE, nsamples=256)
        assert np.abs(r["alpha_base"] - np.array(want)).max() < 1e-13
    d = with_identity(ctrfk, robots["ctr_sec2_tub3.xml"])
    r = oracle.fk_one(d, sv["cases"][0]["jv"], ctrfk.MODE_NSAMPLE, nsamples=256)
    assert np.abs(r["tip"][:9] - np.array(sv["tip_rotation_sec2_tub3_F_colmajor"])).max() < 1e-14


@pytest.mark.parametrize("name", util.ROBOT_FILES)
def test_c_oracle_matches_python_restatement(ctrfk, oracle, robots, name):
    import ctr_oracle_py as py
    d = robots[name]
    model = py.model_from_desc(d)
    jv = ctrfk.sample_joints_host(d, 99, 0, 12)
    jv = np.vstack([jv, util.edge_joint_sets(d)[:8]])
    for mode, kw in ((ctrfk.MODE_NSAMPLE, dict(nsamples=64)), (ctrfk.MODE_NSAMPLE, dict(nsamples=256)),
                     (ctrfk.MODE_STEP, dict(step=d.max_length / 256.0)), (ctrfk.MODE_STEP, dict(step=0.004)),
                     (ctrfk.MODE_STEP, dict(step=1e-5))):      # last one: N capped at MaxSamples
        for q in jv:
            a = oracle.fk_one(d, q, mode, **kw)
            b = py.forward_kinematics(model, list(q), mode, **kw)
            assert a["n"] == b["n"]
            assert a["step"] == b["step"]
            fb = np.array([py.frame_to_colmajor(f) for f in b["frames"]])
            assert np.abs(a["backbone"] - fb).max() < 1e-14
            assert np.abs(a["radius"] - np.array(b["radius"])).max() == 0.0
            assert np.abs(a["alpha_base"] - np.array(b["alpha_base"])).max() < 1e-14
            assert np.abs(a["tip"] - fb[-1]).max() < 1e-14


def test_golden_regression(ctrfk, oracle, robots):
    g = np.load(os.path.join(HERE, "golden", "oracle_golden.npz"))
    for name in util.ROBOT_FILES:
        key = name.replace(".xml", "")
        d = robots[name]
        jv = g[key + "/jv"]
        assert np.array_equal(jv, ctrfk.sample_joints_host(d, util.SEEDS[name], 0, jv.shape[0]))
        for tag, mode in (("F", ctrfk.MODE_NSAMPLE), ("S", ctrfk.MODE_STEP)):
            r = oracle.fk_batch(d, jv, mode, step=d.max_length / 256.0, nsamples=256, nthreads=2)
            assert np.array_equal(r["n"], g["%s/%s/n" % (key, tag)])
            assert np.array_equal(r["step"], g["%s/%s/step" % (key, tag)])
            assert np.abs(r["tip"] - g["%s/%s/tip" % (key, tag)]).max() < 1e-13
            assert np.abs(r["alpha_base"] - g["%s/%s/alpha_base" % (key, tag)]).max() < 1e-12


def golden_desc(ctrfk, robots, g, name):
    """Descriptor of the robot with the entry frame exactly as the reference's XML reader produced it."""
    d = ctrfk.RobotDesc.from_buffer_copy(bytes(robots[name]))
    init = g[name.replace(".xml", "") + "/init"]
    assert np.abs(init - np.array(list(d.init_trafo))).max() < 1e-15      # ctr_fk_orthogonalize agrees to rounding
    for i in range(12):
        d.init_trafo[i] = init[i]
    return d


def test_reference_golden(ctrfk, oracle, robots):
    """tests/golden/reference_golden.npz holds outputs of the reference's own code (oracle/_ref, see
    tests/golden/make_reference_golden.py).  The oracle reproduces every stored number bit for bit."""
    g = np.load(os.path.join(HERE, "golden", "reference_golden.npz"))
    for name in util.ROBOT_FILES:
        key = name.replace(".xml", "")
        d = golden_desc(ctrfk, robots, g, name)
        jv = g[key + "/jv"]
        for tag, mode in (("F", ctrfk.MODE_NSAMPLE), ("S", ctrfk.MODE_STEP)):
            r = oracle.fk_batch(d, jv, mode, step=d.max_length / 256.0, nsamples=256, nthreads=2, backbone=True, radius=True)
            assert np.array_equal(r["n"], g["%s/%s/n" % (key, tag)])
            assert np.array_equal(r["step"], g["%s/%s/step" % (key, tag)])
            assert np.array_equal(r["tip"], g["%s/%s/tip" % (key, tag)])
            idx = g["%s/%s/probe_idx" % (key, tag)]
            assert np.array_equal(np.take_along_axis(r["backbone"], idx[:, :, None].astype(np.int64), axis=1),
                                  g["%s/%s/probe_frames" % (key, tag)])
            live = np.arange(256)[None, :] < r["n"][:, None]
            assert np.array_equal(r["radius"].astype(np.float32)[live], g["%s/%s/radius" % (key, tag)][live])
        for i, q in enumerate(jv[:16]):
            assert np.array_equal(oracle.jacobian(d, q, 128, 1e-5, 1e-4)[0], g[key + "/jac"][i])
        for i, q in enumerate(jv[:32]):
            assert oracle.stability(d, q, 0.2, d.max_length / 128.0) == g[key + "/stable"][i]
            assert oracle.degree_stability(d, q, 0.2, d.max_length / 128.0, -1.0)[0] == g[key + "/degree"][i]
        # construction-time call of the XML path (ctr_static.h:248-252)
        c = oracle.fk_one(d, g[key + "/ctor/jv"], ctrfk.MODE_STEP, step=d.max_length / 256.0)
        assert c["n"] == g[key + "/ctor/n"] and np.array_equal(c["tip"], g[key + "/ctor/tip"])


def single_cc_desc(ctrfk, kappa=40.0, length=0.08):
    d = ctrfk.RobotDesc()
    d.nsections = 1
    d.max_samples = 4096
    for i in range(12):
        d.init_trafo[i] = IDENT[i]
    d.sec[0].ntube = 1
    d.sec[0].length = length
    d.sec[0].stiffness[0] = 3.0
    d.sec[0].curvature[0] = kappa
    d.sec[0].radius[0] = 0.001
    d.sec[0].poisson[0] = 0.3
    return d


def test_K1_single_tube_closed_form(ctrfk, oracle):
    """One CC tube: tip = Rz(theta) * arc(kappa, phi) independent of N and of alpha_0.
    With N >= 2 tube 0's angle is pinned to 0 (section_i.h:308) so the bending plane is x-z."""
    kappa, phi = 40.0, 0.05
    d = single_cc_desc(ctrfk, kappa)
    for n in (2, 16, 256, 1024):
        for a0 in (0.0, 1.3):
            r = oracle.fk_one(d, [0.0, phi, a0], ctrfk.MODE_NSAMPLE, nsamples=n)
            # curvature vector of the innermost tube is (0, kappa, 0) in its own frame -> bends in x-z
            want = np.array([(1 - np.cos(kappa * phi)) / kappa, 0.0, np.sin(kappa * phi) / kappa])
            assert np.abs(r["tip"][9:] - want).max() < 1e-15 * n + 1e-16


def rz(t):
    c, s = np.cos(t), np.sin(t)
    return np.array([[c, -s, 0], [s, c, 0], [0, 0, 1.0]])


def to_mat(t12):
    M = np.eye(4)
    M[:3, :3] = np.asarray(t12[:9]).reshape(3, 3).T
    M[:3, 3] = t12[9:]
    return M


@pytest.mark.parametrize("name", util.ROBOT_FILES)
def test_K2_K3_theta_equivariance_and_orthonormality(ctrfk, oracle, robots, name):
    d = robots[name]
    Init = to_mat(list(d.init_trafo))
    jv = ctrfk.sample_joints_host(d, 5, 0, 6)
    for q in jv:
        q0 = q.copy(); q0[0] = 0.0
        a = oracle.fk_one(d, q, ctrfk.MODE_NSAMPLE, nsamples=128)
        b = oracle.fk_one(d, q0, ctrfk.MODE_NSAMPLE, nsamples=128)
        Rz4 = np.eye(4); Rz4[:3, :3] = rz(q[0])
        G = Init @ Rz4 @ np.linalg.inv(Init)
        for k in range(a["n"]):
            A, Bm = to_mat(a["backbone"][k]), to_mat(b["backbone"][k])
            assert np.abs(A - G @ Bm).max() < 5e-15
            R = A[:3, :3]
            assert np.abs(R.T @ R - np.eye(3)).max() < 1e-12


def test_K4_first_order_convergence(ctrfk, oracle, robots):
    """The scheme is O(h): guards against 'fixing' the integrator."""
    d = ctrfk.RobotDesc.from_buffer_copy(robots["ctr_sec2_tub3.xml"])
    d.max_samples = 4096
    q = [0.3, 0.04, 0.5, 1.0, 0.05, 2.0]
    p = {n: oracle.fk_one(d, q, ctrfk.MODE_NSAMPLE, nsamples=n)["tip"][9:] for n in (256, 1024, 4096)}
    e1 = np.linalg.norm(p[256] - p[1024])
    e2 = np.linalg.norm(p[1024] - p[4096])
    assert 1e-4 < e1 < 5e-3          # ~1 mm between N=256 and N=1024 (SURVEY READ-FIRST 2)
    assert 3.0 < e1 / e2 < 9.0          # first order (a second-order scheme would give ~16)


def test_K6_construction_call_and_counts(ctrfk, oracle, robots):
    """ctr_static.h:248-252: JV = [0, phi_xml, 0, ...], step = MaxLength/MaxSamples -> N = 3 for
    ctr_sec2_tub3.xml with MaxSamples = 256."""
    d, phi = ctrfk.desc_from_xml(ctrfk.resource("ctr_sec2_tub3.xml"), 256, want_phi=True)
    assert phi == [0.001, 0.001]
    jv = [0.0, phi[0], 0.0, 0.0, phi[1], 0.0]
    r = oracle.fk_one(d, jv, ctrfk.MODE_STEP, step=d.max_length / 256.0)
    assert r["n"] == 3
    assert abs(r["step"] - (0.06647 + 0.0667) / 256) < 1e-18


def test_edge_cases(ctrfk, oracle, robots):
    d = robots["ctr_sec3_tub4.xml"]
    nj = d.njoints
    # all phi = 0: step mode still yields N = 1 and one step of length h
    q = np.zeros(nj); q[0] = 0.4
    r = oracle.fk_one(d, q, ctrfk.MODE_STEP, step=1e-3)
    assert r["n"] == 1 and r["step"] == 1e-3
    r = oracle.fk_one(d, q, ctrfk.MODE_NSAMPLE, nsamples=16)
    assert r["n"] == 16 and r["step"] == 0.0
    Init = np.array(list(d.init_trafo))
    assert np.abs(r["tip"][9:] - Init[9:]).max() < 1e-18        # zero-length robot: tip at the entry point
    # phi out of range -> status flag, NaN pose
    q[1] = -1e-3
    r = oracle.fk_one(d, q, ctrfk.MODE_NSAMPLE, nsamples=16)
    assert r["status"] == ctrfk.STATUS_PHI_RANGE and np.isnan(r["tip"]).all() and r["n"] == 0
    # bad arguments
    with pytest.raises(RuntimeError):
        oracle.fk_one(d, np.zeros(nj), ctrfk.MODE_STEP, step=0.0)
    with pytest.raises(RuntimeError):
        oracle.fk_one(d, np.zeros(nj), ctrfk.MODE_NSAMPLE, nsamples=0)


def test_step_mode_can_drop_tip_sample(ctrfk, oracle, robots):
    """the position-driven forward loop sometimes emits N-1 frames."""
    d = ctrfk.RobotDesc.from_buffer_copy(robots["ctr_sec3_tub4.xml"])
    d.max_samples = 64
    jv = ctrfk.sample_joints_host(d, 1234, 0, 4000)
    step = d.max_length / 64.0
    r = oracle.fk_batch(d, jv, ctrfk.MODE_STEP, step=step)
    arc = jv[:, 1] + jv[:, 4] + jv[:, 6]
    n_sweep = np.maximum(1, np.floor(arc / step)).astype(int)
    dropped = (r["n"] == n_sweep - 1).sum()
    assert ((r["n"] == n_sweep) | (r["n"] == n_sweep - 1)).all()
    assert 0 < dropped < 200


def test_radius_is_outer_radius_of_covering_section(ctrfk, oracle, robots):
    d = robots["ctr_sec3_tub4.xml"]
    q = [0.1, 0.03, 0.2, 0.3, 0.04, 0.4, 0.01, 0.5]
    r = oracle.fk_one(d, q, ctrfk.MODE_NSAMPLE, nsamples=80)
    h = r["step"]
    rad = r["radius"]
    assert rad[0] == 0.00125 and rad[-1] == 0.0005
    k1, k2 = int(0.03 / h), int(0.07 / h)
    assert (rad[:k1] == 0.00125).all() and (rad[k1:k2] == 0.001).all() and (rad[k2:] == 0.0005).all()


def test_jacobian_structure(ctrfk, oracle, robots):
    d = robots["ctr_sec2_tub3.xml"]
    q = np.array([0.3, 0.04, 0.5, 1.0, 0.05, 2.0])
    J, tip = oracle.jacobian(d, q, 128, 1e-6, 1e-6)
    assert np.all(J[2] == 0.0)                                   # alpha of tube 0 (robot_i.h:928)
    z = np.array(list(d.init_trafo))[6:9]
    assert np.allclose(J[0, 3:], z)
    # finite differences agree with a direct central difference of the oracle
    for j in (1, 3, 4, 5):
        e = np.zeros(6); e[j] = 1e-6
        tp = oracle.fk_one(d, q + e, ctrfk.MODE_NSAMPLE, nsamples=128)["tip"][9:]
        assert np.abs((tp - tip[9:]) / 1e-6 - J[j, :3]).max() < 1e-9


def test_stability_and_degree(ctrfk, oracle, robots):
    """calcStability / calcDegreeStability (robot_i.h:756-899): short, barely curved overlaps are
    stable; long overlaps of strongly precurved tubes lose monotonicity of the base angle."""
    d = robots["ctr_sec2_tub3.xml"]
    step = d.max_length / 256.0
    q_short = np.array([0.0, 0.005, 0.0, 0.5, 0.005, 0.7])
    assert oracle.stability(d, q_short, 0.05, step) == 1
    deg, st = oracle.degree_stability(d, q_short, 0.05, step)
    assert st == 1 and 0.0 < deg < np.pi / 2
    jv = ctrfk.sample_joints_host(d, 4, 0, 300)
    flags = np.array([oracle.stability(d, q, 0.1, step) for q in jv])
    degs = np.array([oracle.degree_stability(d, q, 0.1, step) for q in jv])
    assert np.array_equal(flags, degs[:, 1].astype(int))
    assert 0 < flags.sum() < len(flags)                      # both outcomes occur
    assert (degs[flags == 0, 0] > np.pi / 2).all()           # atan2(step, negative) > pi/2
    assert (degs[flags == 1, 0] <= np.pi / 2).all()
