"""Host-side logic and the C-ABI surface (no GPU needed)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import util

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)


def test_library_exports_every_declared_symbol(ctrfk):
    """The C-ABI library loads and exports every symbol include/ctr_fk.h declares."""
    header = open(os.path.join(REPO, "include", "ctr_fk.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(ctr_fk_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 20
    lib = ctrfk.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert declared == set(ctrfk.SYMBOLS), declared ^ set(ctrfk.SYMBOLS)
    assert b"sm_100a" in lib.ctr_fk_version()


def test_struct_layout_matches_header(ctrfk):
    assert C.sizeof(ctrfk.SectionDesc) == 8 + 8 + 4 * 16
    assert C.sizeof(ctrfk.RobotDesc) == 8 + 12 * 8 + 8 + 3 * C.sizeof(ctrfk.SectionDesc)


def test_xml_v1(ctrfk):
    d, phi = ctrfk.desc_from_xml(ctrfk.resource("ctr_sec3_tub5.xml"), 256, want_phi=True)
    assert d.nsections == 3 and [d.sec[i].ntube for i in range(3)] == [2, 2, 1]
    assert d.njoints == 9 and d.ntubes == 5 and d.max_samples == 256
    assert phi == [0.001, 0.001, 0.001]
    assert d.sec[0].length == 0.06647 and d.sec[1].length == 0.0667 and d.sec[2].length == 0.01265
    assert list(d.sec[0].stiffness) == [5.0, 5.0] and d.sec[2].stiffness[0] == 0.05
    assert d.sec[0].curvature[0] == 52.2 and d.sec[1].curvature[0] == 52.63 and d.sec[2].curvature[0] == 1e-6
    assert list(d.sec[0].radius) == [0.00125, 0.001] and d.sec[2].radius[0] == 0.0005
    assert d.sec[0].poisson[1] == 0.3
    lib = ctrfk.load()
    assert lib.ctr_fk_desc_njoints(C.byref(d)) == 9 and lib.ctr_fk_desc_ntubes(C.byref(d)) == 5
    assert lib.ctr_fk_desc_max_length(C.byref(d)) == 0.06647 + 0.0667 + 0.01265
    # entry frame: orthogonalised (raw one has |R^T R - I| = 1.3e-4), translation untouched
    t = np.array(list(d.init_trafo))
    R = t[:9].reshape(3, 3).T
    assert np.abs(R.T @ R - np.eye(3)).max() < 1e-15 and np.linalg.det(R) > 0.999
    raw = np.array([[-0.69710, -0.59750, -0.39620], [-0.69710, 0.69400, 0.18010], [0.16730, 0.40170, -0.90040]])
    assert np.abs(R - raw).max() < 2e-4
    assert list(t[9:]) == [0.079, 0.077, 0.148]
    # polar factor = U V^T of the SVD
    U, _, Vt = np.linalg.svd(raw)
    assert np.abs(R - U @ Vt).max() < 1e-14


def test_xml_v2_equals_v1(ctrfk):
    a = ctrfk.desc_from_xml(ctrfk.resource("ctr_sec3_tub4.xml"), 128)
    b, phi = ctrfk.desc_from_xml(ctrfk.resource("ctr_sec3_tub4_v2.xml"), 128, want_phi=True)
    assert bytes(a) == bytes(b)
    assert phi == [0.001, 0.001, 0.001]


@pytest.mark.skipif(not os.path.isdir("/root/reference/ctr_resources"), reason="reference tree not present")
def test_generated_xml_equals_reference_resources(ctrfk):
    """Only in the build container: the generated tests/data robots carry exactly the parameters
    of the reference's ctr_resources (read in place, never copied)."""
    for name in util.ROBOT_FILES:
        a = ctrfk.desc_from_xml(ctrfk.resource(name), 256)
        b = ctrfk.desc_from_xml(os.path.join("/root/reference/ctr_resources", name), 256)
        assert bytes(a) == bytes(b), name


def test_xml_errors(ctrfk, tmp_path):
    def expect(text, frag):
        p = tmp_path / "r.xml"
        p.write_text(text)
        with pytest.raises(ctrfk.CtrFkError) as ei:
            ctrfk.desc_from_xml(str(p), 16)
        assert ei.value.code == ctrfk.E_XML
        assert frag in str(ei.value)

    with pytest.raises(ctrfk.CtrFkError) as ei:
        ctrfk.desc_from_xml(str(tmp_path / "missing.xml"), 16)
    assert ei.value.code == ctrfk.E_XML and "could not read file" in str(ei.value)
    tube = "<tube><stiffness>1</stiffness><curvature>2</curvature><radius>0.1</radius><poissonRatio>0.3</poissonRatio><length>0.1</length></tube>"
    expect("<robot><section>%s</section>" % tube, "missing </robot>")
    expect("<robot><section>%s%s%s</section></robot>" % (tube, tube, tube), "maximum of 2 tubes")
    expect("<robot>%s</robot>" % ("<section>%s</section>" % tube * 4), "MAX SECTION")
    expect("<robot><section>%s</section></robot>" % tube.replace("<radius>0.1</radius>", ""), "radius")
    expect("<robot><section>%s</section></robot>" % tube.replace(">2<", ">abc<"), "curvature")
    expect("<robot></robot>", "no sections")
    bad_frame = "<entryFrame>" + "".join("<x%d%d>%s</x%d%d>" % (r, c, "1", r, c) for r in range(3) for c in range(4)) + "</entryFrame>"
    expect("<robot><section>%s</section>%s</robot>" % (tube, bad_frame), "not a transformation matrix")
    # tags are case-insensitive for section/tube/entryframe; <phi> defaults to 0; attributes are skipped
    p = tmp_path / "ok.xml"
    p.write_text('<?xml version="1.0"?><Robot2/><robot a="1"><SECTION><Tube>%s</Tube></SECTION></robot>' % tube[6:-7])
    d, phi = ctrfk.desc_from_xml(str(p), 16, want_phi=True)
    assert d.nsections == 1 and d.sec[0].ntube == 1 and phi == [0.0]


def test_orthogonalize(ctrfk):
    lib = ctrfk.load()
    t = np.array([1, 0, 0, 0, 1, 0, 0, 0, 1, 1, 2, 3], dtype=np.float64)
    assert lib.ctr_fk_orthogonalize(t.ctypes.data, 0.01) == 0
    assert list(t) == [1, 0, 0, 0, 1, 0, 0, 0, 1, 1, 2, 3]
    t[0] = 1.5
    assert lib.ctr_fk_orthogonalize(t.ctypes.data, 0.01) == ctrfk.E_ARG
    t = np.array([1, 0, 0, 0, 1, 0, 0, 0, -1, 0, 0, 0], dtype=np.float64)     # reflection
    assert lib.ctr_fk_orthogonalize(t.ctypes.data, 0.0) == ctrfk.E_ARG


def splitmix64(z):
    z = (z + np.uint64(0x9E3779B97F4A7C15))
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def test_sampler_matches_numpy_replica_and_distribution(ctrfk, robots):
    d = robots["ctr_sec3_tub4.xml"]
    B, seed, first = 5000, 20173, 1000
    jv = ctrfk.sample_joints_host(d, seed, first, B)
    with np.errstate(over="ignore"):
        i = (np.arange(B, dtype=np.uint64) + np.uint64(first))[:, None]
        j = np.arange(d.njoints, dtype=np.uint64)[None, :]
        a = splitmix64(np.uint64(seed) + np.uint64(0x9E3779B97F4A7C15) * i)
        b = splitmix64(a + j)
    u = (b >> np.uint64(11)).astype(np.float64) / 9007199254740992.0
    c = np.nextafter(2 * np.pi, 0.0)
    scale = np.array([c, 0.06647, c, c, 0.0667, c, 0.01265, c])
    assert np.array_equal(jv, u * scale)
    # a shard generated on its own equals the slice of the whole
    assert np.array_equal(ctrfk.sample_joints_host(d, seed, first + 123, 77), jv[123:200])
    assert (u >= 0).all() and (u < 1).all() and abs(u.mean() - 0.5) < 0.01
    assert (jv[:, 0] < 2 * np.pi).all() and (jv[:, 1] < 0.06647).all()


def test_sampler_pow_accuracy(ctrfk, robots):
    """The gamma policy's own pow (ctr_sampler.h::sampler_pow — the same bits on host and device): phi = Length * u^g
    within 0.75 ulp of the power (+ 0.5 ulp of the product) of the exact value, over eight exponents and eighteen decades
    of u.  Checked through the host sampler with the numpy replica of its uniform variates."""
    mp = pytest.importorskip("mpmath")
    mp.mp.prec = 200
    d = robots["ctr_sec2_tub3.xml"]
    B, seed = 1500, 77
    with np.errstate(over="ignore"):
        i = np.arange(B, dtype=np.uint64)[:, None]
        j = np.arange(d.njoints, dtype=np.uint64)[None, :]
        u = (splitmix64(splitmix64(np.uint64(seed) + np.uint64(0x9E3779B97F4A7C15) * i) + j) >> np.uint64(11)).astype(np.float64) / 9007199254740992.0
    worst = 0.0
    for g in (0.01, 0.3, 0.5, 1.0, 1.7, 2.2, 5.0, 30.0):
        jv = ctrfk.sample_joints_host(d, seed, 0, B, policy=ctrfk.SCALE_GAMMA, p0=g, lut_size=0)
        for col, L in ((1, d.sec[0].length), (4, d.sec[1].length)):
            for k in range(0, B, 3):
                exact = mp.mpf(L) * mp.power(mp.mpf(float(u[k, col])), mp.mpf(g))
                if exact < mp.mpf("1e-300"):
                    continue
                err = abs(mp.mpf(float(jv[k, col])) - exact) / mp.mpf(float(np.spacing(float(exact))))
                worst = max(worst, float(err))
    assert worst < 1.5, worst
    # special values: u = 0 is never drawn but must not trap; g = 1 is the identity up to the one rounding of L * u
    jv1 = ctrfk.sample_joints_host(d, seed, 0, B, policy=ctrfk.SCALE_GAMMA, p0=1.0, lut_size=0)
    assert np.abs(jv1[:, 1] - d.sec[0].length * u[:, 1]).max() <= np.spacing(d.sec[0].length)


@pytest.mark.parametrize("name", util.ROBOT_FILES)
def test_count_samples_matches_oracle(ctrfk, oracle, robots, name):
    d = ctrfk.RobotDesc.from_buffer_copy(robots[name])
    d.max_samples = 64
    jv = np.vstack([ctrfk.sample_joints_host(d, 3, 0, 3000), util.edge_joint_sets(d)])
    for mode, kw in ((ctrfk.MODE_STEP, dict(step=d.max_length / 64.0)), (ctrfk.MODE_STEP, dict(step=1e-4)),
                     (ctrfk.MODE_NSAMPLE, dict(nsamples=40)), (ctrfk.MODE_NSAMPLE, dict(nsamples=400))):
        n = ctrfk.count_samples_host(d, jv, mode, **kw)
        r = oracle.fk_batch(d, jv, mode, **kw)
        assert np.array_equal(n, r["n"])


def test_descriptor_validation(ctrfk, robots):
    lib = ctrfk.load()
    d = ctrfk.RobotDesc.from_buffer_copy(robots["ctr_sec2_tub3.xml"])
    assert lib.ctr_fk_desc_validate(C.byref(d)) == 0
    for mut in ("nsec0", "nsec4", "ntube3", "stiff0", "maxs0", "nan"):
        e = ctrfk.RobotDesc.from_buffer_copy(d)
        if mut == "nsec0": e.nsections = 0
        if mut == "nsec4": e.nsections = 4
        if mut == "ntube3": e.sec[0].ntube = 3
        if mut == "stiff0": e.sec[1].stiffness[0] = 0.0
        if mut == "maxs0": e.max_samples = 0
        if mut == "nan": e.init_trafo[3] = float("nan")
        assert lib.ctr_fk_desc_validate(C.byref(e)) == ctrfk.E_DESC, mut
    assert lib.ctr_fk_desc_validate(None) == ctrfk.E_NULL


def test_no_cpu_fallback_without_gpu(ctrfk, robots):
    """The product path fails loudly when there is no CUDA device: no oracle, no CPU path."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(ctrfk.CtrFkError) as ei:
        ctrfk.Engine(robots["ctr_sec2_tub3.xml"], 0)
    assert ei.value.code > 0                     # a cudaError_t
    assert "no CPU path" in str(ei.value)
    # the shared library does not link or mention the oracle
    blob = open(ctrfk.LIB_PATH, "rb").read()
    assert b"ctr_oracle" not in blob and b"libctr_oracle" not in blob


def test_null_handle_and_bad_args(ctrfk):
    lib = ctrfk.load()
    assert lib.ctr_fk_batch(None, None, 1, 0, 1e-3, 0, 0, None, None, None, None, None, None, None, None) == ctrfk.E_NULL
    assert lib.ctr_fk_destroy(None) == 0
    assert lib.ctr_fk_create(None, 0, None) == ctrfk.E_NULL


def test_sampler_scale_policies(ctrfk, robots):
    """transrotscale.h policies on the Phi slots (section_i.h:338-347)."""
    d = robots["ctr_sec3_tub4.xml"]
    B, seed = 4000, 9
    base = ctrfk.sample_joints_host(d, seed, 0, B)
    lens = np.array([0.06647, 0.0667, 0.01265]); phi = [1, 4, 6]
    u = base[:, phi] / lens
    # GammaCorrection without table: phi = L * u^gamma (transrotscale.h:120); other slots untouched
    g = ctrfk.sample_joints_host(d, seed, 0, B, policy=ctrfk.SCALE_GAMMA, p0=2.5)
    assert np.allclose(g[:, phi], lens * u ** 2.5, rtol=1e-14, atol=0)
    rest = [j for j in range(d.njoints) if j not in phi]
    assert np.array_equal(g[:, rest], base[:, rest])
    # ... and with a 64-entry interpolated table (:113-117): close to the exact power, exact at the nodes
    gl = ctrfk.sample_joints_host(d, seed, 0, B, policy=ctrfk.SCALE_GAMMA, p0=2.5, lut_size=64)
    assert np.abs(gl[:, phi] / lens - u ** 2.5).max() < 2e-3
    # ProportionalScale (:273-291): the total extension of configuration i hits
    # MaxLength * (p0 + (p1-p0) * u_goal), u_goal = last-section draw of configuration i-1
    p0, p1 = 0.3, 0.6
    pr = ctrfk.sample_joints_host(d, seed, 0, B, policy=ctrfk.SCALE_PROPORTIONAL, p0=p0, p1=p1)
    total = pr[:, phi].sum(axis=1)
    goal = lens.sum() * (p0 + (p1 - p0) * u[:-1, 2])
    assert np.abs(total[1:] - goal).max() < 1e-15
    assert (pr[:, phi] >= 0).all() and (pr[:, phi] <= lens).all()
    assert total[0] >= lens.sum() * p0 - 1e-15 and total[0] <= lens.sum() * p1 + 1e-15
    # shards are independent
    assert np.array_equal(ctrfk.sample_joints_host(d, seed, 100, 50, policy=ctrfk.SCALE_PROPORTIONAL, p0=p0, p1=p1), pr[100:150])
    for bad in (dict(policy=7), dict(policy=ctrfk.SCALE_GAMMA, p0=-1.0), dict(policy=ctrfk.SCALE_PROPORTIONAL, p0=0.8, p1=0.2)):
        with pytest.raises(ctrfk.CtrFkError):
            ctrfk.sample_joints_host(d, seed, 0, 4, **bad)
