"""GPU parity tests proper: the CUDA path, called through the C ABI (include/ctr_fk.h), against
the CPU oracle on the same seeded inputs.  Gate (BASELINE.json north_star): tip position within
1e-9 m, orientation within 1e-9 rad, sample count and returned step exactly equal."""
import os

import numpy as np
import pytest

import util

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    torch.cuda.set_device(0)
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def engines(ctrfk, robots, dev):
    e = {n: ctrfk.Engine(d, 0) for n, d in robots.items()}
    yield e
    for x in e.values():
        x.close()


def run_gpu(ctrfk, eng, jv_np, mode, flags=0, backbone=False, radius=False, soa=False, **kw):
    B = jv_np.shape[0]
    st = torch.cuda.current_stream().cuda_stream
    jv = torch.from_numpy(np.ascontiguousarray(jv_np)).cuda()
    maxs = eng.desc.max_samples
    out = dict(tip=torch.full((B, 12), -7.0, dtype=torch.float64, device="cuda"),
               n=torch.full((B,), -7, dtype=torch.int32, device="cuda"),
               step=torch.full((B,), -7.0, dtype=torch.float64, device="cuda"),
               alpha_base=torch.full((B, eng.ntubes), -7.0, dtype=torch.float64, device="cuda"),
               status=torch.full((B,), -7, dtype=torch.int32, device="cuda"))
    bb = torch.zeros((maxs, 12, B) if soa else (B, maxs, 12), dtype=torch.float64, device="cuda") if backbone else None
    rd = torch.zeros((maxs, B) if soa else (B, maxs), dtype=torch.float64, device="cuda") if radius else None
    if soa:
        flags |= ctrfk.FLAG_SOA
    before = ctrfk.launch_count()
    eng.batch(jv, B, mode, flags=flags, tip=out["tip"], backbone=bb, radius=rd, n_out=out["n"],
              step_out=out["step"], alpha_base=out["alpha_base"], status=out["status"], stream=st, **kw)
    torch.cuda.synchronize()
    assert B == 0 or ctrfk.launch_count() > before, "no kernel of libctrfk.so was launched"
    res = {k: v.cpu().numpy() for k, v in out.items()}
    if backbone:
        res["backbone"] = (bb.permute(2, 0, 1) if soa else bb).cpu().numpy()
    if radius:
        res["radius"] = (rd.permute(1, 0) if soa else rd).cpu().numpy()
    return res


def assert_parity(o, g, tol_pos=util.TOL_POS, tol_rot=util.TOL_ROT):
    assert np.array_equal(o["n"], g["n"])
    assert np.array_equal(o["step"], g["step"], equal_nan=True)
    ok = o["status"] == 0
    assert np.array_equal(o["status"], g["status"])
    dp, ang = util.pose_errors(o["tip"][ok], g["tip"][ok])
    assert dp.max() <= tol_pos, dp.max()
    assert ang.max() <= tol_rot, ang.max()
    assert np.abs(o["alpha_base"][ok] - g["alpha_base"][ok]).max() <= 1e-9
    assert np.isnan(g["tip"][~ok]).all()
    return dp.max(), ang.max()


MODES = [("F256", 1, dict(nsamples=256)), ("S256", 0, None)]


@pytest.mark.parametrize("name", util.ROBOT_FILES)
def test_tip_parity_1M(ctrfk, oracle, robots, engines, name):
    """SURVEY §8d parity gate: 1e6 seeded configurations per robot and mode, FAST and STRICT
    kernels against the oracle; sample count and returned step bit-exact."""
    d = robots[name]
    B = 1_000_000
    jv = ctrfk.sample_joints_host(d, util.SEEDS[name], 0, B)
    for tag, mode, kw in MODES:
        kw = kw or dict(step=d.max_length / 256.0)
        o = oracle.fk_batch(d, jv, mode, **kw)
        for flags in (0, 1):
            g = run_gpu(ctrfk, engines[name], jv, mode, flags=flags, **kw)
            dp, ang = assert_parity(o, g)
            print("%s %s %s: max tip error %.2e m, %.2e rad" % (name, tag, "strict" if flags else "fast", dp, ang))
            # measured margins are recorded in DESIGN.md §5; both variants sit far inside the gate
            assert dp < (2e-10 if flags == 0 else 1e-12) and ang < 1e-10, (tag, dp, ang)


def test_golden_vectors(ctrfk, robots, engines):
    g = np.load(os.path.join(HERE, "golden", "oracle_golden.npz"))
    for name in util.ROBOT_FILES:
        key = name.replace(".xml", "")
        d = robots[name]
        jv = g[key + "/jv"]
        for tag, mode in (("F", 1), ("S", 0)):
            r = run_gpu(ctrfk, engines[name], jv, mode, step=d.max_length / 256.0, nsamples=256)
            assert np.array_equal(r["n"], g["%s/%s/n" % (key, tag)])
            assert np.array_equal(r["step"], g["%s/%s/step" % (key, tag)])
            dp, ang = util.pose_errors(r["tip"], g["%s/%s/tip" % (key, tag)])
            assert dp.max() <= util.TOL_POS and ang.max() <= util.TOL_ROT
            assert np.abs(r["alpha_base"] - g["%s/%s/alpha_base" % (key, tag)]).max() <= 1e-9


@pytest.mark.parametrize("name", util.ROBOT_FILES)
def test_parity_with_the_reference_build(ctrfk, oracle, robots, engines, name):
    """The CUDA path against the reference's own code running on this box (oracle/_ref/libctr_ref.so, built
    in the container from /root/reference/ctr_source and shipped prebuilt): 200k configurations per robot and
    mode through CTR::Robot<double>::calcKinematic; the oracle must equal it bit for bit here too."""
    import ref_lib
    if not os.path.exists(ref_lib.LIB_PATH):
        pytest.skip("oracle/_ref/libctr_ref.so did not travel to this box")
    d = robots[name]
    ref = ref_lib.RefRobot.from_desc(d)
    B = 200_000
    jv = ctrfk.sample_joints_host(d, util.SEEDS[name] + 100, 0, B)
    for tag, mode, kw in MODES:
        kw = kw or dict(step=d.max_length / 256.0)
        r = ref.fk_batch(jv, mode, **kw)
        o = oracle.fk_batch(d, jv, mode, **kw)
        assert np.array_equal(r["tip"], o["tip"]) and np.array_equal(r["n"], o["n"]) and np.array_equal(r["step"], o["step"])
        g = run_gpu(ctrfk, engines[name], jv, mode, **kw)
        assert np.array_equal(g["n"], r["n"]) and np.array_equal(g["step"], r["step"])
        dp, ang = util.pose_errors(r["tip"], g["tip"])
        print("%s %s vs reference build: %.2e m, %.2e rad" % (name, tag, dp.max(), ang.max()))
        assert dp.max() <= util.TOL_POS and ang.max() <= util.TOL_ROT


def test_reference_golden_gpu(ctrfk, robots, dev):
    """The CUDA path against numbers the REFERENCE ITSELF produced (tests/golden/reference_golden.npz,
    generated by tests/golden/make_reference_golden.py from oracle/_ref): sample count and step exact,
    tip and probed backbone frames within the gate, radius exact, Jacobian and stability verdicts."""
    g = np.load(os.path.join(HERE, "golden", "reference_golden.npz"))
    for name in util.ROBOT_FILES:
        key = name.replace(".xml", "")
        d = ctrfk.RobotDesc.from_buffer_copy(bytes(robots[name]))
        for i in range(12):
            d.init_trafo[i] = g[key + "/init"][i]
        eng = ctrfk.Engine(d, 0)
        jv = g[key + "/jv"]
        for tag, mode in (("F", 1), ("S", 0)):
            for flags in (0, 1):
                r = run_gpu(ctrfk, eng, jv, mode, flags=flags, backbone=True, radius=True, step=d.max_length / 256.0, nsamples=256)
                assert np.array_equal(r["n"], g["%s/%s/n" % (key, tag)])
                assert np.array_equal(r["step"], g["%s/%s/step" % (key, tag)])
                dp, ang = util.pose_errors(r["tip"], g["%s/%s/tip" % (key, tag)])
                assert dp.max() <= util.TOL_POS and ang.max() <= util.TOL_ROT
                idx = g["%s/%s/probe_idx" % (key, tag)].astype(np.int64)
                got = np.take_along_axis(r["backbone"], idx[:, :, None], axis=1)
                dp, ang = util.pose_errors(got.reshape(-1, 12), g["%s/%s/probe_frames" % (key, tag)].reshape(-1, 12))
                assert dp.max() <= util.TOL_POS and ang.max() <= util.TOL_ROT
                live = np.arange(256)[None, :] < r["n"][:, None]
                assert np.array_equal(r["radius"].astype(np.float32)[live], g["%s/%s/radius" % (key, tag)][live])
        jac = np.zeros((16, d.njoints, 6)); tip = np.zeros((16, 12))
        eng.jacobian_host(np.ascontiguousarray(jv[:16]), 16, 128, 1e-5, 1e-4, jac, tip)
        assert np.abs(jac - g[key + "/jac"]).max() < 1e-6
        jq = torch.from_numpy(np.ascontiguousarray(jv[:32])).cuda()
        stable = torch.full((32,), 9, dtype=torch.int32, device="cuda")
        degree = torch.zeros(32, dtype=torch.float64, device="cuda")
        eng.stability(jq, 32, 0.2, d.max_length / 128.0, -1.0, stable, degree, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        assert np.array_equal(stable.cpu().numpy(), g[key + "/stable"].astype(np.int32))
        assert np.abs(degree.cpu().numpy() - g[key + "/degree"]).max() < 1e-7
        eng.close()


@pytest.mark.parametrize("name", util.ROBOT_FILES)
def test_edge_cases(ctrfk, oracle, robots, engines, name):
    d = robots[name]
    jv = util.edge_joint_sets(d)
    bad = jv[:3].copy()
    bad[0, 1] = -1e-6                       # phi < 0
    bad[1, 1] = d.sec[0].length * 1.0001    # phi > Length
    bad[2, 1] = np.nan
    jv = np.vstack([jv, bad])
    for mode, kw in ((1, dict(nsamples=256)), (1, dict(nsamples=1)), (1, dict(nsamples=5)), (1, dict(nsamples=1000)),
                     (0, dict(step=d.max_length / 256.0)), (0, dict(step=0.01)), (0, dict(step=1.0)),
                     (0, dict(step=2e-5))):
        o = oracle.fk_batch(d, jv, mode, **kw)
        for flags in (0, 1):
            g = run_gpu(ctrfk, engines[name], jv, mode, flags=flags, **kw)
            assert_parity(o, g)
            assert list(g["status"][-3:]) == [1, 1, 1] and (g["n"][-3:] == 0).all()


def test_empty_single_and_ragged_batches(ctrfk, oracle, robots, engines):
    name = "ctr_sec2_tub3.xml"
    d = robots[name]
    for B in (0, 1, 31, 127, 128, 129, 1000):
        jv = ctrfk.sample_joints_host(d, 5, 0, B)
        g = run_gpu(ctrfk, engines[name], jv, 1, nsamples=64)
        if B:
            o = oracle.fk_batch(d, jv, 1, nsamples=64)
            assert_parity(o, g)


def test_bad_arguments_are_rejected(ctrfk, robots, engines):
    eng = engines["ctr_sec2_tub3.xml"]
    jv = torch.zeros(4, 6, dtype=torch.float64, device="cuda")
    for kw in (dict(mode=0, step=0.0), dict(mode=0, step=-1.0), dict(mode=0, step=float("nan")),
               dict(mode=1, nsamples=0), dict(mode=7, step=1e-3)):
        mode = kw.pop("mode")
        with pytest.raises(ctrfk.CtrFkError) as ei:
            eng.batch(jv, 4, mode, **kw)
        assert ei.value.code == ctrfk.E_ARG
    with pytest.raises(ctrfk.CtrFkError) as ei:
        eng.batch(None, 4, 1, nsamples=4)
    assert ei.value.code == ctrfk.E_NULL


def test_step_mode_binning_does_not_change_results(ctrfk, oracle, robots, engines):
    """Configurations are counting-sorted by sample count for the launch only: outputs land at
    their original index and are bit-identical with and without binning."""
    name = "ctr_sec3_tub4.xml"
    d = robots[name]
    jv = ctrfk.sample_joints_host(d, 77, 0, 50_000)
    step = d.max_length / 256.0
    a = run_gpu(ctrfk, engines[name], jv, 0, step=step)
    b = run_gpu(ctrfk, engines[name], jv, 0, step=step, flags=ctrfk.FLAG_NO_BINNING)
    for k in ("tip", "n", "step", "alpha_base", "status"):
        assert np.array_equal(a[k], b[k]), k
    o = oracle.fk_batch(d, jv, 0, step=step)
    assert_parity(o, a)


def test_step_mode_dropped_tip_sample(ctrfk, oracle, robots, dev):
    d = ctrfk.RobotDesc.from_buffer_copy(robots["ctr_sec3_tub4.xml"])
    d.max_samples = 64
    eng = ctrfk.Engine(d, 0)
    jv = ctrfk.sample_joints_host(d, 1234, 0, 20_000)
    step = d.max_length / 64.0
    o = oracle.fk_batch(d, jv, 0, step=step)
    arc = jv[:, 1] + jv[:, 4] + jv[:, 6]
    assert (o["n"] == np.maximum(1, np.floor(arc / step)) - 1).sum() > 10      # the quirk occurs
    for flags in (0, 1):
        g = run_gpu(ctrfk, eng, jv, 0, step=step, flags=flags)
        assert_parity(o, g)
    eng.close()


@pytest.mark.parametrize("name", ["ctr_sec3_tub4.xml", "ctr_sec2_tub3.xml"])
@pytest.mark.parametrize("soa", [False, True], ids=["aos", "soa"])
def test_backbone_and_radius_parity(ctrfk, oracle, robots, engines, name, soa):
    d = robots[name]
    B = 3000
    jv = np.vstack([ctrfk.sample_joints_host(d, util.SEEDS[name], 0, B), util.edge_joint_sets(d)])
    for mode, kw in ((1, dict(nsamples=256)), (0, dict(step=d.max_length / 256.0)), (1, dict(nsamples=3))):
        o = oracle.fk_batch(d, jv, mode, backbone=True, radius=True, **kw)
        for flags in (0, 1):
            g = run_gpu(ctrfk, engines[name], jv, mode, flags=flags, backbone=True, radius=True, soa=soa, **kw)
            assert_parity(o, g)
            for i in range(0, jv.shape[0], 37):
                n = o["n"][i]
                dp, ang = util.pose_errors(o["backbone"][i, :n], g["backbone"][i, :n])
                assert dp.max() <= util.TOL_POS and ang.max() <= util.TOL_ROT
                assert np.array_equal(o["radius"][i, :n], g["radius"][i, :n])
                # the tip output is the last frame of the backbone
                assert np.array_equal(g["backbone"][i, n - 1], g["tip"][i])


def test_host_pointer_batch_and_single(ctrfk, oracle, robots, engines):
    name = "ctr_sec3_tub5.xml"
    d = robots[name]
    eng = engines[name]
    B = 300_000                                     # > 2 chunks of the pipelined host path
    jv = ctrfk.sample_joints_host(d, util.SEEDS[name], 0, B)
    tip = np.zeros((B, 12)); n = np.zeros(B, np.int32); st = np.zeros(B); status = np.zeros(B, np.int32)
    eng.batch_host(jv, B, 1, nsamples=128, tip=tip, n_out=n, step_out=st, status=status)
    g = run_gpu(ctrfk, eng, jv, 1, nsamples=128)
    assert np.array_equal(tip, g["tip"]) and np.array_equal(n, g["n"]) and np.array_equal(st, g["step"])
    # pinned memory gives the same answer
    jvp = torch.from_numpy(jv).pin_memory(); tipp = torch.zeros(B, 12, dtype=torch.float64).pin_memory()
    eng.batch_host(jvp, B, 1, nsamples=128, tip=tipp)
    assert np.array_equal(tipp.numpy(), tip)
    o = oracle.fk_batch(d, jv[:20000], 1, nsamples=128)
    dp, ang = util.pose_errors(o["tip"], tip[:20000])
    assert dp.max() <= util.TOL_POS and ang.max() <= util.TOL_ROT
    # single configuration through the same kernels (the reference's one-shot call)
    maxs = d.max_samples
    for q in jv[:5]:
        t1 = np.zeros(12); bb = np.zeros((maxs, 12)); rd = np.zeros(maxs); ab = np.zeros(d.ntubes)
        import ctypes as C
        nn = C.c_int32(0); hh = C.c_double(0)
        rc = eng.single(q, 0, step=d.max_length / 256.0, tip=t1, backbone=bb, radius=rd,
                        n_out=C.addressof(nn), step_out=C.addressof(hh), alpha_base=ab)
        assert rc == 0
        r = oracle.fk_one(d, q, 0, step=d.max_length / 256.0)
        assert nn.value == r["n"] and hh.value == r["step"]
        dp, ang = util.pose_errors(r["backbone"], bb[: r["n"]])
        assert dp.max() <= util.TOL_POS and ang.max() <= util.TOL_ROT
        assert np.array_equal(rd[: r["n"]], r["radius"])
        assert np.abs(ab - r["alpha_base"]).max() < 1e-9
        assert np.array_equal(t1, bb[r["n"] - 1])
    qbad = jv[0].copy(); qbad[1] = -0.5                      # Phi out of range: a code of its own, NaN outputs
    t1 = np.zeros(12)
    assert eng.single(qbad, 1, nsamples=64, tip=t1) == ctrfk.E_RANGE and np.isnan(t1).all()


@pytest.mark.parametrize("name", ["ctr_sec2_tub3.xml", "ctr_sec3_tub3.xml"])      # even and odd joint counts
def test_host_pointer_zero_copy_form(ctrfk, oracle, robots, engines, name):
    """ctr_fk_batch_host with page-locked buffers and CTR_FK_FLAG_HOST_ZEROCOPY: ONE kernel reads the joint vectors from and writes every output to
    host memory directly.  Buffers from ctr_fk_host_alloc (NUMA-local, mapped; one write-combined) and from torch's
    pin_memory(); a batch whose last CTA is partial; an invalid configuration; all outputs.  Bit-equal to the staged
    form (CTR_FK_FLAG_HOST_STAGED) and to the device-pointer call, launches counted: 1 against one per chunk."""
    d = robots[name]
    eng = engines[name]
    B = 200_000 + 77
    jv_np = ctrfk.sample_joints_host(d, util.SEEDS[name] + 5, 0, B)
    jv_np[1234, 1] = -1e-3                                        # Phi out of range: NaN pose, status 1
    jv, p_jv = eng.host_array((B, d.njoints), write_combined=True)
    tip, p_tip = eng.host_array((B, 12))
    n, p_n = eng.host_array((B,), dtype="int32")
    st, p_st = eng.host_array((B,))
    status, p_status = eng.host_array((B,), dtype="int32")
    jv[...] = jv_np
    tip[...] = 7.0
    assert isinstance(eng.host_local_cpus(), str)
    for mode, kw in ((1, dict(nsamples=96)), (0, dict(step=d.max_length / 256.0, flags=ctrfk.FLAG_NO_BINNING))):
        l0 = ctrfk.launch_count()
        kwz = dict(kw); kwz["flags"] = kw.get("flags", 0) | ctrfk.FLAG_HOST_ZEROCOPY
        eng.batch_host(jv, B, mode, tip=tip, n_out=n, step_out=st, status=status, **kwz)
        assert ctrfk.launch_count() - l0 == 1, "zero-copy form is one launch"
        kw2 = dict(kw)                                              # the default: staged
        tip2 = np.zeros((B, 12)); n2 = np.zeros(B, np.int32); st2 = np.zeros(B); status2 = np.zeros(B, np.int32)
        l0 = ctrfk.launch_count()
        eng.batch_host(jv_np, B, mode, tip=tip2, n_out=n2, step_out=st2, status=status2, **kw2)
        assert ctrfk.launch_count() - l0 > 1, "staged form is one launch per chunk"
        assert np.array_equal(tip, tip2, equal_nan=True) and np.array_equal(n, n2)
        assert np.array_equal(st, st2, equal_nan=True) and np.array_equal(status, status2)
        assert status[1234] == 1 and np.isnan(tip[1234]).all() and status.sum() == 1
        g = run_gpu(ctrfk, eng, jv_np, mode, **kw)
        assert np.array_equal(tip, g["tip"], equal_nan=True) and np.array_equal(n, g["n"])
        ok = np.arange(B) != 1234
        o = oracle.fk_batch(d, jv_np[:5000], mode, **{k: v for k, v in kw.items() if k != "flags"})
        dp, ang = util.pose_errors(o["tip"][ok[:5000]], tip[:5000][ok[:5000]])
        assert dp.max() <= util.TOL_POS and ang.max() <= util.TOL_ROT
    # torch-pinned buffers take the same path; a binned step-mode batch and STRICT fall back to the staged form
    jvp = torch.from_numpy(jv_np).pin_memory(); tipp = torch.zeros(B, 12, dtype=torch.float64).pin_memory()
    l0 = ctrfk.launch_count()
    eng.batch_host(jvp, B, 1, nsamples=96, tip=tipp, flags=ctrfk.FLAG_HOST_ZEROCOPY)
    assert ctrfk.launch_count() - l0 == 1
    eng.batch_host(jv, B, 1, nsamples=96, tip=tip, flags=ctrfk.FLAG_HOST_ZEROCOPY)
    assert np.array_equal(tipp.numpy(), tip, equal_nan=True)
    l0 = ctrfk.launch_count()
    eng.batch_host(jvp, B, 0, step=d.max_length / 256.0, tip=tipp, flags=ctrfk.FLAG_HOST_ZEROCOPY)
    assert ctrfk.launch_count() - l0 > 1
    for p in (p_jv, p_tip, p_n, p_st, p_status):
        eng.host_free(p)


@pytest.mark.parametrize("name", util.ROBOT_FILES)
def test_both_forms_of_the_sweep_loop(ctrfk, oracle, robots, name, monkeypatch):
    """The tip and gather kernels run one of two forms of FastSweep::run_scaled, chosen per launch by fine_sampling()
    (ctr_fk_launch.h).  Forced either way (CTRFK_FORCE_FINE, read at ctr_fk_create): with fine sampling every
    iteration is hot and the two forms are the same arithmetic — bit-identical poses; with coarse sampling both stay
    inside the gate against the oracle.  So the choice only ever changes speed."""
    d = robots[name]
    B = 50_000
    jv = ctrfk.sample_joints_host(d, util.SEEDS[name] + 11, 0, B)
    res = {}
    for force in ("0", "1"):
        monkeypatch.setenv("CTRFK_FORCE_FINE", force)
        eng = ctrfk.Engine(d, 0)
        res[force] = {n: run_gpu(ctrfk, eng, jv, 1, nsamples=n) for n in (256, 40)}
        res[force]["step"] = run_gpu(ctrfk, eng, jv, 0, step=d.max_length / 256.0)
        eng.close()
    monkeypatch.delenv("CTRFK_FORCE_FINE")
    for key in (256, "step"):
        # bit-identical wherever every iteration was hot — all but a few configurations per 10 000 (ctr_sec3_tub5's
        # soft inner tubes turn fastest) — and to rounding on the rest
        same = (res["0"][key]["tip"] == res["1"][key]["tip"]).all(axis=1)
        dp, ang = util.pose_errors(res["0"][key]["tip"], res["1"][key]["tip"])
        util.record("sweep_loop_forms", dict(robot=name, case=str(key), identical=float(same.mean()), max_dp=float(dp.max()), max_ang=float(ang.max())))
        assert same.mean() >= 0.999 and dp.max() < 1e-12 and ang.max() < 1e-12, (key, same.mean(), dp.max(), ang.max())
        assert np.array_equal(res["0"][key]["n"], res["1"][key]["n"])
    o = oracle.fk_batch(d, jv, 1, nsamples=40)
    for force in ("0", "1"):
        assert np.array_equal(o["n"], res[force][40]["n"])
        dp, ang = util.pose_errors(o["tip"], res[force][40]["tip"])
        assert dp.max() <= util.TOL_POS and ang.max() <= util.TOL_ROT, (force, dp.max(), ang.max())
        assert np.abs(o["alpha_base"] - res[force][40]["alpha_base"]).max() < 1e-9
    # the automatic choice: N = 256 is a fine launch, N = 40 a coarse one, on every shipped robot
    eng = ctrfk.Engine(d, 0)
    assert np.array_equal(run_gpu(ctrfk, eng, jv, 1, nsamples=256)["tip"], res["1"][256]["tip"])
    assert np.array_equal(run_gpu(ctrfk, eng, jv, 1, nsamples=40)["tip"], res["0"][40]["tip"])
    eng.close()


def test_device_sampler_equals_host_sampler(ctrfk, robots, engines):
    st = torch.cuda.current_stream().cuda_stream
    for name in util.ROBOT_FILES:
        d = robots[name]
        B = 10_000
        jv = torch.empty(B, d.njoints, dtype=torch.float64, device="cuda")
        engines[name].sample_joints(util.SEEDS[name], 12345, B, jv, st)
        torch.cuda.synchronize()
        assert np.array_equal(jv.cpu().numpy(), ctrfk.sample_joints_host(d, util.SEEDS[name], 12345, B))
        # scale policies: bit-identical on host and device — ProportionalScale by non-contracted arithmetic, the
        # gamma policy by the sampler's own pow (ctr_sampler.h::sampler_pow; libm's differs between glibc and CUDA)
        engines[name].sample_joints(5, 77, B, jv, st, policy=ctrfk.SCALE_PROPORTIONAL, p0=0.2, p1=0.9)
        torch.cuda.synchronize()
        assert np.array_equal(jv.cpu().numpy(), ctrfk.sample_joints_host(d, 5, 77, B, policy=ctrfk.SCALE_PROPORTIONAL, p0=0.2, p1=0.9))
        for gamma, lut in ((1.7, 0), (0.45, 0), (2.5, 64), (2.5, 1024)):
            engines[name].sample_joints(5, 77, B, jv, st, policy=ctrfk.SCALE_GAMMA, p0=gamma, lut_size=lut)
            torch.cuda.synchronize()
            assert np.array_equal(jv.cpu().numpy(), ctrfk.sample_joints_host(d, 5, 77, B, policy=ctrfk.SCALE_GAMMA, p0=gamma, lut_size=lut)), (gamma, lut)


# Finite-difference Jacobians: a column is (pose(q + dq e_j) - pose(q)) / dq (robot_i.h:41-56), so an arithmetic
# difference of eps in each pose becomes at most 2 eps / |dq| in the column.  eps = what the two solvers measure against
# the oracle on 1e6 configurations, with a factor ~3 of head room:
#   FAST   position 1e-10 m, orientation 2e-12 rad;   STRICT  position 1e-13 m, orientation 1e-12 rad.
# Rows 0-2 of a column are translations, rows 3-5 the vee of (R_hi - R_lo) R_hi^T / 2.
JAC_EPS = {"fast": (1e-10, 2e-12), "strict": (1e-13, 1e-12)}


def jacobian_tolerance(d, fd_trans, fd_rot, arith):
    """[njv][6] bound on |J_gpu - J_oracle| derived from the finite-difference steps."""
    import util as _u
    eps_p, eps_r = JAC_EPS[arith]
    tol = np.zeros((d.njoints, 6))
    phi = set(_u.phi_indices(d))
    for j in range(d.njoints):
        dq = fd_trans if j in phi else fd_rot
        tol[j, :3] = 2.0 * eps_p / dq
        tol[j, 3:] = 2.0 * eps_r / dq
    tol[0] = (2.0 * eps_p, 2.0 * eps_p, 2.0 * eps_p, 4e-16, 4e-16, 4e-16)      # analytic column 0 (robot_i.h:956-959)
    tol[2] = 0.0                                                               # column 2 is exactly zero (:928)
    return tol


@pytest.mark.parametrize("arith", ["fast", "strict"])
@pytest.mark.parametrize("name", util.ROBOT_FILES)
def test_jacobian_parity(ctrfk, oracle, robots, engines, name, arith):
    """Every Jacobian of 20 000 configurations per robot (NSamples form, robot_i.h:917-960), FAST and STRICT
    (CTR_FK_FLAG_STRICT), entry by entry against the oracle within the bound derived from the finite-difference steps."""
    d = robots[name]
    B = 20_000
    jv = ctrfk.sample_joints_host(d, 9, 0, B)
    jv[0, 1] = d.sec[0].length - 1e-7            # forces the backward difference (robot_i.h:948-950)
    fd_t, fd_r = 1e-5, 1e-4
    flags = ctrfk.FLAG_STRICT if arith == "strict" else 0
    st = torch.cuda.current_stream().cuda_stream
    djv = torch.from_numpy(jv).cuda()
    jac = torch.zeros(B, d.njoints, 6, dtype=torch.float64, device="cuda")
    tip = torch.zeros(B, 12, dtype=torch.float64, device="cuda")
    before = ctrfk.launch_count()
    engines[name].jacobian_ex(djv, B, 1, fd_t, fd_r, jac, nsamples=128, tip=tip, flags=flags, stream=st)
    torch.cuda.synchronize()
    assert ctrfk.launch_count() > before
    J, T = jac.cpu().numpy(), tip.cpu().numpy()
    o = oracle.jacobian_batch(d, jv, 1, nsamples=128, fd_trans=fd_t, fd_rot=fd_r)
    assert (o["rc"] == 0).all()
    dp, ang = util.pose_errors(o["tip"], T)
    assert dp.max() <= util.TOL_POS and ang.max() <= util.TOL_ROT
    tol = jacobian_tolerance(d, fd_t, fd_r, arith)
    err = np.abs(J - o["jac"]).max(axis=0)                    # [njv][6] worst over the batch
    util.record("jacobian_parity", dict(robot=name, arith=arith, configs=B, max_abs_err=float(err.max()),
                                        max_err_over_tol=float((err[tol > 0] / tol[tol > 0]).max()),
                                        tip_pos_err=float(dp.max()), tip_rot_err=float(ang.max())))
    assert (err <= tol).all(), (err, tol)
    assert np.all(J[:, 2] == 0.0)
    if arith == "fast":     # the host-pointer entry gives the same numbers
        Jh = np.zeros((512, d.njoints, 6)); Th = np.zeros((512, 12))
        engines[name].jacobian_host(jv[:512], 512, 128, fd_t, fd_r, Jh, Th)
        assert np.array_equal(Jh, J[:512]) and np.array_equal(Th, T[:512])


@pytest.mark.parametrize("name", util.ROBOT_FILES)
def test_jacobian_other_forms(ctrfk, oracle, robots, engines, name):
    """ctr_fk_jacobian_batch_ex: the ArcLengthStep form (robot_i.h:901-907), the iSample form (:1012-1072) and the
    forms that take the base poses from the caller, against the oracle (itself bit-equal to the reference build)."""
    d, eng = robots[name], engines[name]
    B = 150
    jv_np = ctrfk.sample_joints_host(d, 23, 0, B)
    jv_np[0, 1] = d.sec[0].length - 1e-7
    jv_np[3, 1] = -0.2                                            # invalid -> NaN
    jv = torch.from_numpy(jv_np).cuda()
    st = torch.cuda.current_stream().cuda_stream
    nj = d.njoints
    for mode, kw, ks in ((0, dict(step=d.max_length / 200.0), -1), (1, dict(nsamples=100), 37), (1, dict(nsamples=100), 0)):
        jac = torch.zeros(B, nj, 6, dtype=torch.float64, device="cuda"); jacs = torch.zeros_like(jac)
        tip = torch.zeros(B, 12, dtype=torch.float64, device="cuda"); smp = torch.zeros_like(tip)
        n = torch.zeros(B, dtype=torch.int32, device="cuda")
        before = ctrfk.launch_count()
        eng.jacobian_ex(jv, B, mode, 1e-5, 1e-4, jac, i_sample=ks, jac_sample=jacs if ks >= 0 else None, tip=tip,
                        sample=smp if ks >= 0 else None, n_out=n, stream=st, **kw)
        torch.cuda.synchronize()
        assert ctrfk.launch_count() > before
        J, Js, T, Sm, N = (x.cpu().numpy() for x in (jac, jacs, tip, smp, n))
        assert np.isnan(J[3]).all() and np.isnan(T[3]).all() and N[3] == 0
        for i in list(range(0, B, 9)) + [1]:
            if i == 3:
                continue
            o = oracle.jacobian_ex(d, jv_np[i], mode, i_sample=ks, **kw)
            assert N[i] == o["n"]
            dp, ang = util.pose_errors(o["tip"][None], T[i][None])
            assert dp.max() <= util.TOL_POS and ang.max() <= util.TOL_ROT
            assert np.abs(o["jac"] - J[i]).max() < 1e-6
            if ks >= 0:
                dp, ang = util.pose_errors(o["sample"][None], Sm[i][None])
                assert dp.max() <= util.TOL_POS and ang.max() <= util.TOL_ROT
                assert np.abs(o["jac_sample"] - Js[i]).max() < 1e-6
        if mode == 1:       # base poses handed in (robot_i.h:917, :1012): identical columns, through the host entry
            J2 = np.zeros((B, nj, 6)); Js2 = np.zeros((B, nj, 6))
            eng.jacobian_host_ex(jv_np, B, 1, 1e-5, 1e-4, J2, nsamples=100, i_sample=ks, jac_sample=Js2, tip=T, sample=Sm, given=True)
            ok = np.arange(B) != 3
            assert np.array_equal(J2[ok], J[ok]) and np.array_equal(Js2[ok], Js[ok])
    # ArcLengthStep form on a batch large enough for the window-local binning: same bits with and without it
    Bb = 20_000
    jvb = torch.from_numpy(ctrfk.sample_joints_host(d, 24, 0, Bb)).cuda()
    ja = torch.zeros(Bb, nj, 6, dtype=torch.float64, device="cuda"); jb = torch.zeros_like(ja)
    na = torch.zeros(Bb, dtype=torch.int32, device="cuda"); nb = torch.zeros_like(na)
    eng.jacobian_ex(jvb, Bb, 0, 1e-5, 1e-4, ja, step=d.max_length / 256.0, n_out=na, stream=st)
    eng.jacobian_ex(jvb, Bb, 0, 1e-5, 1e-4, jb, step=d.max_length / 256.0, n_out=nb, flags=ctrfk.FLAG_NO_BINNING, stream=st)
    torch.cuda.synchronize()
    assert torch.equal(ja, jb) and torch.equal(na, nb) and int(na.min()) < int(na.max())
    with pytest.raises(ctrfk.CtrFkError):
        eng.jacobian_ex(jv, B, 1, 1e-5, 1e-4, jac, nsamples=100, i_sample=5, stream=st)       # no jac_sample buffer


def test_fused_gather_kernel_single_device(ctrfk, robots, engines):
    """ctr_fk_batch_gather with the "ranks" emulated on one GPU (two local buffers stand in for two ranks' gathered
    buffers; the multi-process path over NVLink is exercised by bench.py --allgather-fused): every destination gets
    this rank's rows at first_row, bit-equal to ctr_fk_batch, invalid configurations as NaN rows, nothing else touched."""
    name = "ctr_sec3_tub3.xml"
    d, eng = robots[name], engines[name]
    st = torch.cuda.current_stream().cuda_stream
    for mode, kw, B in ((1, dict(nsamples=200), 3000), (0, dict(step=d.max_length / 256.0), 20_000)):
        jv_np = ctrfk.sample_joints_host(d, 61, 0, B)
        jv_np[7, 1] = -1.0
        jv = torch.from_numpy(jv_np).cuda()
        tip = torch.empty(B, 12, dtype=torch.float64, device="cuda")
        n = torch.empty(B, dtype=torch.int32, device="cuda"); n2 = torch.empty_like(n)
        eng.batch(jv, B, mode, tip=tip, n_out=n, stream=st, **kw)
        first, rows = 5, B + 11
        bufs = [torch.full((rows, 12), -3.0, dtype=torch.float64, device="cuda") for _ in range(2)]     # 3000 = 23 CTAs + 56 rows
        before = ctrfk.launch_count()
        eng.batch_gather(jv, B, mode, [b.data_ptr() for b in bufs], first, n_out=n2, stream=st, **kw)
        # the same with every thread storing its own row (the path permuted rows take anyway)
        bufs.append(torch.full((rows, 12), -3.0, dtype=torch.float64, device="cuda"))
        eng.batch_gather(jv, B, mode, [bufs[2].data_ptr()], first, flags=ctrfk.FLAG_GATHER_ROWWISE, stream=st, **kw)
        torch.cuda.synchronize()
        assert ctrfk.launch_count() > before and torch.equal(n, n2)
        for b in bufs:
            assert torch.equal(b[first:first + B], tip) or (torch.isnan(tip[7]).all() and torch.isnan(b[first + 7]).all()
                                                            and torch.equal(torch.nan_to_num(b[first:first + B]), torch.nan_to_num(tip)))
            assert bool((b[:first] == -3.0).all()) and bool((b[first + B:] == -3.0).all())
    with pytest.raises(ctrfk.CtrFkError):
        eng.batch_gather(jv, B, 1, [], 0, nsamples=10, stream=st)
    with pytest.raises(ctrfk.CtrFkError):
        eng.batch_gather(jv, B, 1, [bufs[0].data_ptr()] * 2, 0, nsamples=10, multicast=True, stream=st)


@pytest.mark.parametrize("maxs", [64, 256])
def test_host_backbone_equals_device_path(ctrfk, dev, maxs):
    """ctr_fk_batch_host_backbone (host pointers, frames + radius of every sample back over PCIe, chunked by staging
    slab) against ctr_fk_batch on device buffers: bit-equal frames, radii, poses, counts; several chunks."""
    d = ctrfk.desc_from_xml(ctrfk.resource("ctr_sec3_tub4.xml"), maxs)
    eng = ctrfk.Engine(d, 0)
    B = 45_000 if maxs == 64 else 24_000          # 256 MB slabs: 36 k / 9.6 k configurations per chunk
    for mode, kw in ((1, dict(nsamples=maxs)), (0, dict(step=d.max_length / maxs))):
        jv = ctrfk.sample_joints_host(d, 88, 0, B)
        jv[17, 1] = -0.1
        g = run_gpu(ctrfk, eng, jv, mode, backbone=True, radius=True, **kw)
        bb = np.zeros((B, maxs, 12)); rd = np.zeros((B, maxs)); tip = np.zeros((B, 12))
        n = np.zeros(B, np.int32); st = np.zeros(B); stat = np.zeros(B, np.int32)
        eng.batch_host_backbone(jv, B, mode, bb, tip=tip, radius=rd, n_out=n, step_out=st, status=stat, **kw)
        assert np.array_equal(n, g["n"]) and np.array_equal(stat, g["status"]) and stat[17] == 1
        assert np.array_equal(st, g["step"], equal_nan=True) and np.array_equal(tip, g["tip"], equal_nan=True)
        live = (np.arange(maxs)[None, :] < n[:, None])
        assert np.array_equal(bb[live], g["backbone"][live]) and np.array_equal(rd[live], g["radius"][live])
    with pytest.raises(ctrfk.CtrFkError):
        eng.batch_host_backbone(jv, B, 1, None, nsamples=maxs)
    eng.close()


def test_degenerate_robots(ctrfk, oracle, dev):
    """Zero curvature everywhere (the |u| < zero_tol branch), one straight section, Poisson ratio 0, stiffness ratios of
    1e6, a 1 mm section, over every layout: FAST and STRICT kernels against the oracle (which equals the reference
    build on exactly these robots, tests/test_reference_build.py::test_fk_bit_equal_degenerate_robots)."""
    rng = np.random.default_rng(78)
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
            eng = ctrfk.Engine(d, 0)
            jv = ctrfk.sample_joints_host(d, 500, 0, 256)
            for mode, kw in ((1, dict(nsamples=96)), (0, dict(step=d.max_length / 50.0))):
                o = oracle.fk_batch(d, jv, mode, **kw)
                for flags in (0, 1):
                    g = run_gpu(ctrfk, eng, jv, mode, flags=flags, **kw)
                    assert_parity(o, g)
            eng.close()


# ---- full-size, size-independent properties (BASELINE.json config 2: 1M configurations) --------

def rz_batch(theta):
    c, s = np.cos(theta), np.sin(theta)
    R = np.zeros((theta.size, 3, 3))
    R[:, 0, 0] = c; R[:, 0, 1] = -s; R[:, 1, 0] = s; R[:, 1, 1] = c; R[:, 2, 2] = 1
    return R


def test_full_size_properties_1M(ctrfk, oracle, robots, engines):
    name = "ctr_sec2_tub3.xml"
    d = robots[name]
    eng = engines[name]
    B = 1_000_000
    st = torch.cuda.current_stream().cuda_stream
    jv = torch.empty(B, d.njoints, dtype=torch.float64, device="cuda")
    eng.sample_joints(util.SEEDS[name], 0, B, jv, st)
    tip = torch.empty(B, 12, dtype=torch.float64, device="cuda")
    tip2 = torch.empty_like(tip)
    eng.batch(jv, B, 1, nsamples=256, tip=tip, stream=st)
    eng.batch(jv, B, 1, nsamples=256, tip=tip2, stream=st)
    torch.cuda.synchronize()
    assert torch.equal(tip, tip2)                                    # deterministic
    T = tip.cpu().numpy()
    assert np.isfinite(T).all()
    R = T[:, :9].reshape(-1, 3, 3).transpose(0, 2, 1)
    RtR = np.einsum("bij,bik->bjk", R, R)
    assert np.abs(RtR - np.eye(3)).max() < 1e-12                     # K3: orthonormal frames
    assert np.abs(np.linalg.det(R[::1000]) - 1).max() < 1e-12
    # reach: the tip cannot be farther from the entry point than the inserted arc length
    q = jv.cpu().numpy()
    arc = q[:, 1] + q[:, 4]
    Init = np.array(list(d.init_trafo))
    assert (np.linalg.norm(T[:, 9:] - Init[9:], axis=1) <= arc + 1e-12).all()
    # K2: theta-equivariance  T(theta) = Init Rz(theta) Init^-1 T(0)
    jv0 = jv.clone(); jv0[:, 0] = 0.0
    eng.batch(jv0, B, 1, nsamples=256, tip=tip2, stream=st)
    torch.cuda.synchronize()
    T0 = tip2.cpu().numpy()
    Ri = Init[:9].reshape(3, 3).T; pi = Init[9:]
    G = Ri @ rz_batch(q[:, 0]) @ Ri.T                                 # [B,3,3]
    R0 = T0[:, :9].reshape(-1, 3, 3).transpose(0, 2, 1)
    assert np.abs(G @ R0 - R).max() < 1e-13
    p_want = np.einsum("bij,bj->bi", G, T0[:, 9:] - pi) + pi
    assert np.abs(p_want - T[:, 9:]).max() < 1e-13
    # a seeded slice against the oracle
    sl = slice(500_000, 500_000 + 50_000)
    o = oracle.fk_batch(d, q[sl], 1, nsamples=256)
    dp, ang = util.pose_errors(o["tip"], T[sl])
    assert dp.max() <= util.TOL_POS and ang.max() <= util.TOL_ROT


@pytest.mark.parametrize("soa", [False, True], ids=["aos", "soa"])
def test_backbone_full_size_properties(ctrfk, oracle, robots, engines, soa):
    """BASELINE config 4 at bench size (ctr_sec3_tub4, 200 k configurations x 256 frames, 4.9 GB):
    properties that do not need the oracle on every frame — orthonormal rotations, distance from
    the entry point growing by at most one arc-length step per sample, last frame = tip output, radius from the
    descriptor, AoS and SoA outputs identical — plus a slice against the oracle."""
    name = "ctr_sec3_tub4.xml"
    d, eng = robots[name], engines[name]
    B, N = 200_000, 256
    st = torch.cuda.current_stream().cuda_stream
    jv = torch.empty(B, d.njoints, dtype=torch.float64, device="cuda")
    eng.sample_joints(util.SEEDS[name], 0, B, jv, st)
    tip = torch.empty(B, 12, dtype=torch.float64, device="cuda")
    step_out = torch.empty(B, dtype=torch.float64, device="cuda")
    bb = torch.empty((N, 12, B) if soa else (B, N, 12), dtype=torch.float64, device="cuda")
    rd = torch.empty((N, B) if soa else (B, N), dtype=torch.float64, device="cuda")
    eng.batch(jv, B, 1, nsamples=N, flags=ctrfk.FLAG_SOA if soa else 0, tip=tip, backbone=bb, radius=rd,
              step_out=step_out, stream=st)
    torch.cuda.synchronize()
    F = bb.permute(2, 0, 1) if soa else bb                          # [B, N, 12] view
    assert torch.isfinite(F).all()
    assert torch.equal(F[:, N - 1, :], tip)
    h = step_out
    p_init = torch.tensor(list(d.init_trafo)[9:], dtype=torch.float64, device="cuda")
    worst_orth = 0.0
    for lo in range(0, B, 25_000):                                   # chunks: keep temporaries small
        f = F[lo:lo + 25_000]
        R = f[..., :9].reshape(-1, N, 3, 3)                          # column-major: R[..., j, i] = R_ij
        RtR = torch.einsum("bkji,bkli->bkjl", R, R)
        worst_orth = max(worst_orth, float((RtR - torch.eye(3, dtype=torch.float64, device="cuda")).abs().max()))
        # every frame carries its own twist Rz(alpha_k + theta) (robot_i.h:714), so consecutive
        # positions are not one chord apart; their distance from the entry point is twist-invariant
        # and grows by at most one arc-length step per sample
        dist = (f[..., 9:] - p_init).norm(dim=2)
        assert bool(((dist[:, 1:] - dist[:, :-1]).abs() <= h[lo:lo + 25_000, None] * (1 + 1e-9) + 1e-15).all())
        assert bool((dist[:, 0] <= h[lo:lo + 25_000] * (1 + 1e-9) + 1e-15).all())
    assert worst_orth < 1e-12
    radii = torch.tensor(sorted({d.sec[s].radius[0] for s in range(d.nsections)}), dtype=torch.float64, device="cuda")
    R_ = rd.permute(1, 0) if soa else rd
    assert bool(torch.isin(R_, radii).all())
    sl = slice(123_000, 123_000 + 300)
    q = jv[sl].cpu().numpy()
    o = oracle.fk_batch(d, q, 1, nsamples=N, backbone=True, radius=True)
    got = F[sl].cpu().numpy()
    for i in range(q.shape[0]):
        dp, ang = util.pose_errors(o["backbone"][i], got[i])
        assert dp.max() <= util.TOL_POS and ang.max() <= util.TOL_ROT
    assert np.array_equal(o["radius"], R_[sl].cpu().numpy())


def test_single_tube_closed_form_on_gpu(ctrfk, dev):
    """K1 through the CUDA path: one CC tube -> circular arc, independent of N."""
    d = ctrfk.RobotDesc()
    d.nsections = 1; d.max_samples = 1024
    for i, v in enumerate([1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0]):
        d.init_trafo[i] = v
    d.sec[0].ntube = 1; d.sec[0].length = 0.08; d.sec[0].stiffness[0] = 3.0
    d.sec[0].curvature[0] = 40.0; d.sec[0].radius[0] = 1e-3; d.sec[0].poisson[0] = 0.3
    eng = ctrfk.Engine(d, 0)
    phi = np.linspace(0.001, 0.08, 64)
    jv = np.stack([np.zeros_like(phi), phi, np.full_like(phi, 1.3)], axis=1)
    for ns in (2, 16, 256, 1024):
        g = run_gpu(ctrfk, eng, jv, 1, nsamples=ns)
        want = np.stack([(1 - np.cos(40.0 * phi)) / 40.0, np.zeros_like(phi), np.sin(40.0 * phi) / 40.0], axis=1)
        assert np.abs(g["tip"][:, 9:] - want).max() < 1e-14 * ns + 1e-15
    eng.close()


# FP32 variant (ctr_fk_batch_f32): tolerance against the FP64 oracle 5e-5 m / 5e-4 rad (measured on 100 k random
# configurations per robot and mode: <= 2e-5 m / 2e-4 rad; SURVEY 8d proposed 1e-4 / 1e-3) on random
# inputs; n_out and step_out bit-exact (control scalars stay FP64).  Exact-cancellation inputs
# (all alpha = pi with equal stiffness: |u| = 0 in FP64, ~1e-6 in FP32) land in different branches
# of the reference's zero-curvature test and are excluded — they are a set of measure zero.
TOL_F32_POS, TOL_F32_ROT = 5e-5, 5e-4


@pytest.mark.parametrize("name", util.ROBOT_FILES)
def test_f32_variant(ctrfk, oracle, robots, engines, name):
    d = robots[name]
    eng = engines[name]
    B = 100_000
    jv_np = ctrfk.sample_joints_host(d, util.SEEDS[name], 0, B)
    jv_np[7, 1] = -1.0                                   # one invalid configuration
    jv = torch.from_numpy(jv_np).cuda()
    st = torch.cuda.current_stream().cuda_stream
    for mode, kw in ((1, dict(nsamples=256)), (0, dict(step=d.max_length / 256.0))):
        tip = torch.zeros(B, 12, dtype=torch.float64, device="cuda")
        n = torch.zeros(B, dtype=torch.int32, device="cuda")
        h = torch.zeros(B, dtype=torch.float64, device="cuda")
        status = torch.zeros(B, dtype=torch.int32, device="cuda")
        before = ctrfk.launch_count()
        eng.batch_f32(jv, B, mode, tip=tip, n_out=n, step_out=h, status=status, stream=st, **kw)
        torch.cuda.synchronize()
        assert ctrfk.launch_count() > before
        o = oracle.fk_batch(d, jv_np, mode, **kw)
        assert np.array_equal(o["n"], n.cpu().numpy())
        assert np.array_equal(o["step"], h.cpu().numpy(), equal_nan=True)
        assert np.array_equal(o["status"], status.cpu().numpy())
        ok = o["status"] == 0
        dp, ang = util.pose_errors(o["tip"][ok], tip.cpu().numpy()[ok])
        assert dp.max() < TOL_F32_POS and ang.max() < TOL_F32_ROT, (dp.max(), ang.max())
        assert np.isnan(tip.cpu().numpy()[7]).all()


@pytest.mark.parametrize("arith", ["fast", "strict"])
@pytest.mark.parametrize("name", util.ROBOT_FILES)
def test_stability_parity(ctrfk, oracle, robots, engines, name, arith):
    """calcStability / calcDegreeStability on 100 000 configurations per robot, FAST and STRICT.  The verdict is the
    sign of the smallest base-angle difference `rel` between two scan points (robot_i.h:805, :873): it must equal the
    oracle's wherever |rel| >= 1e-11 (FAST moves rel by ~1e-13; how many configurations sit closer than that to the
    boundary is counted and recorded).  degree = atan2(angle_step, rel): |d degree| <= |d rel| / angle_step."""
    d = robots[name]
    B = 100_000
    jv_np = ctrfk.sample_joints_host(d, 31, 0, B)
    jv_np[5, 1] = -0.5                                      # invalid -> -1 / NaN
    jv = torch.from_numpy(jv_np).cuda()
    stable = torch.full((B,), 9, dtype=torch.int32, device="cuda")
    degree = torch.zeros(B, dtype=torch.float64, device="cuda")
    astep, step = 0.1, d.max_length / 256.0
    before = ctrfk.launch_count()
    engines[name].stability(jv, B, astep, step, -1.0, stable, degree, torch.cuda.current_stream().cuda_stream,
                            flags=ctrfk.FLAG_STRICT if arith == "strict" else 0)
    torch.cuda.synchronize()
    assert ctrfk.launch_count() > before
    st, dg = stable.cpu().numpy(), degree.cpu().numpy()
    assert st[5] == -1 and np.isnan(dg[5])
    want_st, want_dg = oracle.stability_batch(d, jv_np, astep, step, -1.0)
    assert want_st[5] < 0
    ok = np.arange(B) != 5
    # the deciding difference, recovered from the oracle's degree: rel = angle_step / tan(degree)
    with np.errstate(over="ignore", divide="ignore"):       # degree 0 = nothing was compared (rel = DBL_MAX)
        rel = astep * np.cos(want_dg[ok]) / np.sin(want_dg[ok])
    near = np.abs(rel) < 1e-11
    differ = st[ok] != want_st[ok]
    util.record("stability_parity", dict(robot=name, arith=arith, configs=int(ok.sum()), unstable=int((want_st[ok] == 0).sum()),
                                         within_1e11_of_boundary=int(near.sum()), verdicts_differing=int(differ.sum()),
                                         max_degree_err=float(np.abs(dg[ok] - want_dg[ok])[~near].max())))
    assert not (differ & ~near).any(), np.nonzero(differ & ~near)[0][:10]
    tol_deg = (1e-12 if arith == "fast" else 1e-13) / astep * 10.0       # |d rel| / angle_step, x10 head room
    assert np.abs(dg[ok] - want_dg[ok])[~near].max() < tol_deg
    assert 0 < (want_st[ok] == 0).sum() < B
    # single-configuration oracle calls agree with the batch form (and calcStability with calcDegreeStability's flag)
    for i in (0, 1, 2, 77):
        ds, ss = oracle.degree_stability(d, jv_np[i], astep, step, -1.0)
        assert ss == want_st[i] == oracle.stability(d, jv_np[i], astep, step) and ds == want_dg[i]


@pytest.mark.parametrize("name", util.ROBOT_FILES)
@pytest.mark.parametrize("mode", [1, 0], ids=["fixedN", "step"])
def test_base_angle_shooting(ctrfk, oracle, robots, engines, name, mode):
    """ctr_fk_batch_base: persistent Newton shooting kernel.  Checker = the reference sweep (oracle)
    at the tip angles found; the pose must equal the oracle's pose at those tip angles."""
    d = robots[name]
    eng = engines[name]
    B = 30_000
    kw = dict(nsamples=128) if mode == 1 else dict(step=d.max_length / 128.0)
    jv = ctrfk.sample_joints_host(d, 78, 0, B)
    o = oracle.fk_batch(d, jv, mode, **kw)
    jvb_np = util.base_angle_inputs(d, jv, o["alpha_base"])
    jvb_np[11, 1] = -1.0                                    # invalid Phi
    jvb = torch.from_numpy(jvb_np).cuda()
    jt = torch.full((B, d.njoints), -9.0, dtype=torch.float64, device="cuda")
    tip = torch.zeros(B, 12, dtype=torch.float64, device="cuda")
    it = torch.full((B,), -9, dtype=torch.int32, device="cuda")
    st = torch.full((B,), -9, dtype=torch.int32, device="cuda")
    rs = torch.full((B,), -9.0, dtype=torch.float64, device="cuda")
    before = ctrfk.launch_count()
    eng.batch_base(jvb, B, mode, tol=1e-11, max_iter=100, jv_tip=jt, tip=tip, iters=it, status=st, resid=rs,
                   stream=torch.cuda.current_stream().cuda_stream, **kw)
    torch.cuda.synchronize()
    assert ctrfk.launch_count() >= before + 2
    jt_np, tip_np, it_np, st_np = jt.cpu().numpy(), tip.cpu().numpy(), it.cpu().numpy(), st.cpu().numpy()
    rs_np = rs.cpu().numpy()
    assert (st_np >= 0).all() and (it_np >= 0).all() and (jt_np != -9.0).all()     # every ticket was served once
    assert st_np[11] == ctrfk.SHOOT_PHI_RANGE and np.isnan(tip_np[11]).all() and np.isnan(rs_np[11])
    ok = st_np == 0
    valid = st_np != ctrfk.SHOOT_PHI_RANGE
    # naive start x = beta with restarts: >= 98 % of the inputs (all have a root) solved within 100 sweeps; without a
    # guess the solve only ends on a root or on the sweep cap
    assert ok[valid].mean() >= 0.98, ok[valid].mean()
    assert set(np.unique(st_np[valid])) <= {ctrfk.SHOOT_OK, ctrfk.SHOOT_MAX_ITER, ctrfk.SHOOT_STALLED}
    o2 = oracle.fk_batch(d, jt_np, mode, **kw)
    good = ok & (o2["status"] == 0)
    res = np.abs(o2["alpha_base"][good][:, 1:] - jvb_np[good][:, util.alpha_indices(d)[1:]]).max(axis=1)
    assert res.max() < 1e-9                                 # roots confirmed by the reference sweep
    allv = valid & (o2["status"] == 0)
    res_all = np.abs(o2["alpha_base"][allv][:, 1:] - jvb_np[allv][:, util.alpha_indices(d)[1:]]).max(axis=1)
    assert np.abs(res_all - rs_np[allv]).max() < 1e-9       # the reported residual is the reference sweep's, root or not
    dp, ang = util.pose_errors(o2["tip"][good], tip_np[good])
    assert dp.max() <= util.TOL_POS and ang.max() <= util.TOL_ROT
    util.record("shooting_naive_start", dict(robot=name, mode=mode, configs=int(valid.sum()), solved=float(ok[valid].mean()),
                                             mean_sweeps=float(it_np[valid].mean()), p99_sweeps=float(np.percentile(it_np[valid], 99)),
                                             statuses={int(k): int(v) for k, v in zip(*np.unique(st_np, return_counts=True))}))
    assert it_np[ok].max() >= it_np[ok].min() + 5           # divergent iteration counts were absorbed


@pytest.mark.parametrize("name", util.ROBOT_FILES)
def test_base_angle_shooting_warm_start(ctrfk, oracle, robots, engines, name):
    """ctr_fk_batch_base_ex with a guess (branch tracking): from tip angles within 0.02 rad of the ones the base
    angles came from, the solve returns THOSE tip angles — on every configuration away from a fold of the base-angle
    map (smallest singular value of the oracle's own d base / d tip above 0.05) — and never restarts."""
    d = robots[name]
    eng = engines[name]
    B = 20_000
    T = d.ntubes
    idx = util.alpha_indices(d)
    jv = ctrfk.sample_joints_host(d, 79, 0, B)
    o = oracle.fk_batch(d, jv, 1, nsamples=128)
    jvb_np = util.base_angle_inputs(d, jv, o["alpha_base"])
    rng = np.random.default_rng(17)
    guess_np = jv.copy()
    guess_np[:, idx[1:]] += rng.uniform(-0.02, 0.02, size=(B, T - 1))
    jvb = torch.from_numpy(jvb_np).cuda(); guess = torch.from_numpy(guess_np).cuda()
    jt = torch.zeros(B, d.njoints, dtype=torch.float64, device="cuda")
    it = torch.zeros(B, dtype=torch.int32, device="cuda"); st = torch.zeros(B, dtype=torch.int32, device="cuda")
    eng.batch_base(jvb, B, 1, nsamples=128, tol=1e-11, max_iter=40, jv_tip=jt, iters=it, status=st, jv_guess=guess,
                   stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    jt_np, it_np, st_np = jt.cpu().numpy(), it.cpu().numpy(), st.cpu().numpy()
    eps = 1e-6
    J = np.zeros((B, T - 1, T - 1))
    for c in range(T - 1):
        qp = jv.copy(); qm = jv.copy()
        qp[:, idx[c + 1]] += eps; qm[:, idx[c + 1]] -= eps
        J[:, :, c] = (oracle.fk_batch(d, qp, 1, nsamples=128)["alpha_base"][:, 1:]
                      - oracle.fk_batch(d, qm, 1, nsamples=128)["alpha_base"][:, 1:]) / (2 * eps)
    regular = np.linalg.svd(J, compute_uv=False)[:, -1] > 0.05
    assert regular.mean() > 0.5
    solved = (st_np[regular] == 0).mean()
    good = regular & (st_np == 0)
    err = np.abs(jt_np[good][:, idx[1:]] - jv[good][:, idx[1:]]).max(axis=1)
    same = (err < 1e-8).mean()
    util.record("shooting_warm_start", dict(robot=name, configs=B, regular=float(regular.mean()), solved_regular=float(solved),
                                            same_root=float(same), mean_sweeps=float(it_np[good].mean()),
                                            solved_all=float((st_np == 0).mean())))
    assert solved >= 0.99 and same >= 0.99, (solved, same)
    assert it_np[good].mean() < 5.0


@pytest.mark.parametrize("name", ["ctr_sec2_tub3.xml", "ctr_sec3_tub5.xml"])
def test_base_angle_shooting_rk4(ctrfk, oracle, robots, engines, name):
    """CTR_FK_FLAG_RK4: roots of the continuum model integrated by RK4.  Checker: oracle/ctr_shoot_rk4_py.py (its own
    RK4 sweep must end at the wanted base angles for the tip angles the kernel returns); the pose returned is the
    reference FK at those tip angles; and the roots approach the reference-exact ones as O(h)."""
    import ctr_oracle_py as py
    import ctr_shoot_rk4_py as rk
    d = robots[name]
    eng = engines[name]
    model = py.model_from_desc(d)
    idx = util.alpha_indices(d)
    B = 4000
    jv = ctrfk.sample_joints_host(d, 80, 0, B)
    jv[:, idx[0]] = 0.0
    med = []
    for n in (64, 128, 256):
        o = oracle.fk_batch(d, jv, 1, nsamples=n)
        jvb_np = util.base_angle_inputs(d, jv, o["alpha_base"])
        jvb = torch.from_numpy(jvb_np).cuda(); guess = torch.from_numpy(jv).cuda()
        jt = torch.zeros(B, d.njoints, dtype=torch.float64, device="cuda")
        tip = torch.zeros(B, 12, dtype=torch.float64, device="cuda")
        st = torch.zeros(B, dtype=torch.int32, device="cuda")
        eng.batch_base(jvb, B, 1, nsamples=n, tol=1e-11, max_iter=40, jv_tip=jt, tip=tip, status=st, jv_guess=guess,
                       flags=ctrfk.FLAG_RK4, stream=torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        jt_np, st_np, tip_np = jt.cpu().numpy(), st.cpu().numpy(), tip.cpu().numpy()
        ok = st_np == 0
        assert ok.mean() > 0.7
        if n == 64:
            for i in np.flatnonzero(ok)[:10]:
                b, _, _ = rk.base_angles(model, jt_np[i].astype(complex), 1, nsamples=n)
                assert np.abs(np.real(b[1:]) - jvb_np[i, idx[1:]]).max() < 1e-10
            o2 = oracle.fk_batch(d, jt_np[ok], 1, nsamples=n)
            dp, ang = util.pose_errors(o2["tip"], tip_np[ok])
            assert dp.max() <= util.TOL_POS and ang.max() <= util.TOL_ROT
        dist = np.abs(jt_np[:, idx[1:]] - jv[:, idx[1:]]).max(axis=1)
        med.append(float(np.median(dist[ok & (dist < 0.2)])))
    util.record("shooting_rk4_root_distance", dict(robot=name, n=[64, 128, 256], median_distance_to_reference_root=med))
    assert med[0] > med[1] > med[2] and 2.5 < med[0] / med[2] < 6.0, med


def test_cpp_host_binary_config1(ctrfk, oracle, robots, dev):
    """BASELINE config 1 through the C++14 host object (host/ctr_test.cpp, the counterpart of the
    reference's ctr_test.cpp): construction-time FK (N = 3, ctr_static.h:248-252) and the joint
    vector of `std::mt19937{}` + `uniform_real_distribution<double>{}` (ctr_test.cpp:30-32)."""
    import re
    import subprocess
    exe = os.path.join(os.path.dirname(ctrfk.LIB_PATH), "ctr_test")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-s", "-C", os.path.dirname(ctrfk.LIB_PATH), "ctr_test"])
    out = subprocess.run([exe, "256", ctrfk.resource("ctr_sec2_tub3.xml")], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    d = robots["ctr_sec2_tub3.xml"]
    m = re.search(r"construction FK: N=(\d+) tip=\(([^)]*)\)", out.stdout)
    assert m and int(m.group(1)) == 3
    tip0 = np.array([float(x) for x in m.group(2).split(",")])
    r0 = oracle.fk_one(d, [0.0, 0.001, 0.0, 0.0, 0.001, 0.0], ctrfk.MODE_STEP, step=d.max_length / 256.0)
    assert r0["n"] == 3 and np.abs(tip0 - r0["tip"][9:]).max() < util.TOL_POS
    jv = np.array([float(x) for x in re.search(r"Joint values: (.*)", out.stdout).group(1).split()])
    # libstdc++: mt19937 default seed 5489; generate_canonical<double,53> takes two 32-bit draws
    rs = np.random.RandomState(5489)
    raw = rs.randint(0, 2 ** 32, size=12, dtype=np.uint64)
    u = [(float(raw[2 * i]) + float(raw[2 * i + 1]) * 4294967296.0) / 18446744073709551616.0 for i in range(6)]
    c = np.nextafter(2 * np.pi, 0.0)
    want = np.array([c * u[0], 0.06647 * u[1], c * u[2], c * u[3], 0.0667 * u[4], c * u[5]])
    assert np.abs(jv - want).max() < 1e-15
    m = re.search(r"FK: N=(\d+) tip=\(([^)]*)\)", out.stdout.split("Joint values")[1])
    r1 = oracle.fk_one(d, jv, ctrfk.MODE_STEP, step=d.max_length / 256.0)
    assert int(m.group(1)) == r1["n"]
    assert np.abs(np.array([float(x) for x in m.group(2).split(",")]) - r1["tip"][9:]).max() < util.TOL_POS


random_robot, ALL_LAYOUTS = util.random_robot, util.ALL_LAYOUTS


@pytest.mark.parametrize("layout", ALL_LAYOUTS, ids=["-".join(map(str, l)) for l in ALL_LAYOUTS])
def test_all_section_layouts(ctrfk, oracle, dev, layout):
    """Every instantiated layout (1-3 sections of 1 or 2 tubes) with random parameters: tip, backbone,
    radius, Jacobian tip, base angles — FAST and STRICT, both modes."""
    rng = np.random.default_rng(hash(layout) % (2 ** 32))
    d = random_robot(ctrfk, rng, layout)
    eng = ctrfk.Engine(d, 0)
    jv = np.vstack([ctrfk.sample_joints_host(d, 3, 0, 1500), util.edge_joint_sets(d)])
    for mode, kw in ((1, dict(nsamples=128)), (0, dict(step=d.max_length / 100.0)), (1, dict(nsamples=2))):
        o = oracle.fk_batch(d, jv, mode, backbone=True, radius=True, **kw)
        for flags in (0, 1):
            g = run_gpu(ctrfk, eng, jv, mode, flags=flags, backbone=True, radius=True, **kw)
            assert_parity(o, g)
            for i in range(0, jv.shape[0], 97):
                n = o["n"][i]
                dp, ang = util.pose_errors(o["backbone"][i, :n], g["backbone"][i, :n])
                assert dp.max() <= util.TOL_POS and ang.max() <= util.TOL_ROT
                assert np.array_equal(o["radius"][i, :n], g["radius"][i, :n])
            t = run_gpu(ctrfk, eng, jv, mode, flags=flags, **kw)           # tip-only kernel
            assert_parity(o, t)
    eng.close()


def test_large_max_samples(ctrfk, oracle, robots, dev):
    """MaxSamples = 4096 (no binning histogram in shared memory beyond it, thread-per-configuration
    backbone kernel above 1024) and N = 4096 sweeps."""
    d = ctrfk.RobotDesc.from_buffer_copy(robots["ctr_sec3_tub5.xml"])
    d.max_samples = 4096
    eng = ctrfk.Engine(d, 0)
    jv = ctrfk.sample_joints_host(d, 8, 0, 9000)
    for mode, kw in ((1, dict(nsamples=4096)), (0, dict(step=d.max_length / 3000.0))):
        o = oracle.fk_batch(d, jv, mode, **kw)
        g = run_gpu(ctrfk, eng, jv, mode, **kw)
        assert_parity(o, g)
    small = jv[:40]
    o = oracle.fk_batch(d, small, 1, nsamples=2048, backbone=True, radius=True)
    g = run_gpu(ctrfk, eng, small, 1, nsamples=2048, backbone=True, radius=True)
    assert_parity(o, g)
    for i in range(0, 40, 7):
        dp, ang = util.pose_errors(o["backbone"][i, :2048], g["backbone"][i, :2048])
        assert dp.max() <= util.TOL_POS and ang.max() <= util.TOL_ROT
    eng.close()


@pytest.mark.parametrize("max_samples", [32, 64, 127, 128, 160, 1024])
def test_backbone_kernel_variants(ctrfk, oracle, robots, dev, max_samples):
    """AoS backbone across MaxSamples: per-sample parking kernel below 128, checkpointed sweep with
    per-run frame emission from 128 to 1024 (runs of 4, 5 and 32 samples per lane), every frame of
    every configuration; step mode includes dropped tip samples and short sweeps (idle lanes)."""
    for name in ("ctr_sec3_tub5.xml", "ctr_sec2_tub3.xml"):
        d = ctrfk.RobotDesc.from_buffer_copy(robots[name])
        d.max_samples = max_samples
        eng = ctrfk.Engine(d, 0)
        jv = np.vstack([ctrfk.sample_joints_host(d, 77, 0, 700), util.edge_joint_sets(d)])
        for mode, kw in ((1, dict(nsamples=max_samples)), (0, dict(step=d.max_length / max_samples)),
                         (1, dict(nsamples=max_samples - 1)), (1, dict(nsamples=33)), (1, dict(nsamples=1))):
            o = oracle.fk_batch(d, jv, mode, backbone=True, radius=True, **kw)
            g = run_gpu(ctrfk, eng, jv, mode, backbone=True, radius=True, **kw)
            assert_parity(o, g)
            for i in range(jv.shape[0]):
                n = o["n"][i]
                dp, ang = util.pose_errors(o["backbone"][i, :n], g["backbone"][i, :n])
                assert dp.max() <= util.TOL_POS and ang.max() <= util.TOL_ROT, (name, mode, kw, i)
                assert np.array_equal(o["radius"][i, :n], g["radius"][i, :n])
                assert np.array_equal(g["backbone"][i, n - 1], g["tip"][i])
        eng.close()


def test_single_configuration_with_odd_max_samples(ctrfk, oracle, robots, dev):
    """ctr_fk_single keeps its frames behind a radius array of MaxSamples doubles: an odd MaxSamples
    must not misalign them."""
    for maxs in (127, 255, 33):
        d = ctrfk.RobotDesc.from_buffer_copy(robots["ctr_sec3_tub4.xml"])
        d.max_samples = maxs
        eng = ctrfk.Engine(d, 0)
        q = ctrfk.sample_joints_host(d, 3, 0, 1)[0]
        t = np.zeros(12); bb = np.zeros((maxs, 12)); rd = np.zeros(maxs)
        n = eng.single(q, 1, nsamples=maxs, tip=t, backbone=bb, radius=rd)
        o = oracle.fk_batch(d, q[None, :], 1, nsamples=maxs, backbone=True, radius=True)
        nn = int(o["n"][0])
        dp, ang = util.pose_errors(o["backbone"][0, :nn], bb[:nn])
        assert dp.max() <= util.TOL_POS and ang.max() <= util.TOL_ROT
        assert np.array_equal(o["radius"][0, :nn], rd[:nn])
        eng.close()


def test_concurrent_calls_on_one_handle(ctrfk, oracle, robots, engines):
    """include/ctr_fk.h: a handle is immutable after create, any thread may call ctr_fk_batch on it
    concurrently with distinct streams; ctr_fk_batch_host calls are serialised per handle.  Four
    threads hammer one engine (fixed-N, step mode with binning temporaries, host path) and every
    result must equal the one computed alone."""
    import threading
    name = "ctr_sec3_tub3.xml"
    d, eng = robots[name], engines[name]
    B = 20000
    jobs = []
    for w in range(4):
        jv = ctrfk.sample_joints_host(d, 900 + w, 0, B)
        mode, kw = ((1, dict(nsamples=200 + w)), (0, dict(step=d.max_length / (150.0 + w))))[w % 2]
        jobs.append((jv, mode, kw, run_gpu(ctrfk, eng, jv, mode, **kw)))
    errors = []

    def worker(w):
        try:
            jv_np, mode, kw, ref = jobs[w]
            stream = torch.cuda.Stream()
            jv = torch.from_numpy(jv_np).cuda()
            tip = torch.empty((B, 12), dtype=torch.float64, device="cuda")
            n = torch.empty((B,), dtype=torch.int32, device="cuda")
            host_tip = np.empty((B, 12))
            torch.cuda.synchronize()
            for it in range(25):
                tip.fill_(-1.0)
                stream.wait_stream(torch.cuda.current_stream())
                eng.batch(jv, B, mode, tip=tip, n_out=n, stream=stream.cuda_stream, **kw)
                stream.synchronize()
                if not (np.array_equal(tip.cpu().numpy(), ref["tip"]) and np.array_equal(n.cpu().numpy(), ref["n"])):
                    errors.append("thread %d iteration %d: device-pointer result differs" % (w, it))
                    return
                if it % 5 == 0:
                    eng.batch_host(jv_np, B, mode, tip=host_tip, **kw)
                    if not np.array_equal(host_tip, ref["tip"]):
                        errors.append("thread %d iteration %d: host-pointer result differs" % (w, it))
                        return
        except Exception as e:          # noqa: BLE001 - report from the worker
            errors.append("thread %d: %r" % (w, e))

    ts = [threading.Thread(target=worker, args=(w,)) for w in range(4)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errors, errors


def test_alignment_and_odd_sizes(ctrfk, oracle, robots, engines):
    """Odd joint counts and odd batch sizes through the host-pointer paths (staging sub-buffers
    must stay 16-byte aligned) and rejection of misaligned device views."""
    name = "ctr_sec3_tub5.xml"                      # n_jv = 9
    d = robots[name]
    eng = engines[name]
    B = 9001
    jv = ctrfk.sample_joints_host(d, 12, 0, B)
    tip = np.zeros((B, 12)); n = np.zeros(B, np.int32)
    eng.batch_host(jv, B, 1, nsamples=32, tip=tip, n_out=n)
    o = oracle.fk_batch(d, jv, 1, nsamples=32)
    dp, ang = util.pose_errors(o["tip"], tip)
    assert dp.max() <= util.TOL_POS and ang.max() <= util.TOL_ROT and np.array_equal(n, o["n"])
    jac = np.zeros((7, d.njoints, 6)); t7 = np.zeros((7, 12))
    eng.jacobian_host(jv[:7], 7, 32, 1e-5, 1e-4, jac, t7)
    assert np.abs(t7 - tip[:7]).max() < 1e-12
    # a view that starts 8 bytes into an allocation is rejected, not faulted on
    buf = torch.zeros(B * 12 + 1, dtype=torch.float64, device="cuda")
    with pytest.raises(ctrfk.CtrFkError) as ei:
        eng.batch(torch.from_numpy(jv).cuda(), B, 1, nsamples=8, tip=buf[1:])
    assert ei.value.code == ctrfk.E_ARG


def test_no_out_of_bounds_writes(ctrfk, robots, engines):
    """compute-sanitizer is not available on this pool, so every output is placed between guard
    bands filled with a sentinel; after all entry points ran, the bands must be untouched."""
    name = "ctr_sec3_tub4.xml"
    d = robots[name]
    eng = engines[name]
    B, G = 8999, 64                                     # odd batch above the binning threshold
    maxs, T, NJ = d.max_samples, d.ntubes, d.njoints
    st = torch.cuda.current_stream().cuda_stream

    def guarded(numel, dtype):
        buf = torch.full((numel + 2 * G,), -777, dtype=dtype, device="cuda")
        return buf, buf[G:G + numel]

    jv = torch.from_numpy(ctrfk.sample_joints_host(d, 6, 0, B)).cuda()
    bufs = {k: guarded(n, t) for k, (n, t) in dict(
        tip=(B * 12, torch.float64), bb=(B * maxs * 12, torch.float64), rad=(B * maxs, torch.float64),
        n=(B, torch.int32), h=(B, torch.float64), ab=(B * T, torch.float64), stt=(B, torch.int32),
        jac=(B * NJ * 6, torch.float64), jt=(B * NJ, torch.float64), it=(B, torch.int32),
        deg=(B, torch.float64), smp=(B * NJ, torch.float64)).items()}
    v = {k: b[1] for k, b in bufs.items()}
    for mode, kw in ((1, dict(nsamples=256)), (0, dict(step=d.max_length / 256.0)), (1, dict(nsamples=1))):
        for flags in (0, ctrfk.FLAG_STRICT, ctrfk.FLAG_SOA, ctrfk.FLAG_STRICT | ctrfk.FLAG_SOA, ctrfk.FLAG_NO_BINNING):
            eng.batch(jv, B, mode, flags=flags, tip=v["tip"], backbone=v["bb"], radius=v["rad"], n_out=v["n"],
                      step_out=v["h"], alpha_base=v["ab"], status=v["stt"], stream=st, **kw)
            eng.batch(jv, B, mode, flags=flags, tip=v["tip"], radius=v["rad"], n_out=v["n"], step_out=v["h"],
                      alpha_base=v["ab"], status=v["stt"], stream=st, **kw)
        eng.batch_f32(jv, B, mode, tip=v["tip"], n_out=v["n"], step_out=v["h"], status=v["stt"], stream=st, **kw)
        eng.batch_base(jv, B, mode, tol=1e-10, max_iter=5, jv_tip=v["jt"], tip=v["tip"], iters=v["it"],
                       status=v["stt"], stream=st, **kw)
    eng.jacobian(jv, B, 64, 1e-5, 1e-4, v["jac"], v["tip"], st)
    eng.stability(jv, 700, 0.4, d.max_length / 64.0, -1.0, v["stt"], v["deg"], st)
    eng.sample_joints(3, 5, B, v["smp"], st)
    eng.sample_joints(3, 5, B, v["smp"], st, policy=ctrfk.SCALE_PROPORTIONAL, p0=0.1, p1=0.9)
    torch.cuda.synchronize()
    for k, (buf, view) in bufs.items():
        assert bool((buf[:G] == -777).all()) and bool((buf[-G:] == -777).all()), "guard band of %s was written" % k


def test_cpp_robot_selftest(ctrfk, dev):
    """The C++14 host object (host/ctr_robot.hpp) through its own self-test binary."""
    import subprocess
    pkg = os.path.dirname(ctrfk.LIB_PATH)
    exe = os.path.join(pkg, "robot_selftest")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-s", "-C", pkg, "robot_selftest"])
    out = subprocess.run([exe, ctrfk.resource("ctr_sec2_tub3.xml")], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.strip().endswith("OK"), out.stdout + out.stderr
