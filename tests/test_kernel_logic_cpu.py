"""The kernels' per-thread solvers (csrc/ctr_fk_core.cuh), compiled for the CPU by tests/hostsim,
against the oracle.  This checks the kernel LOGIC (control scalars, sweep order, the FAST
reformulation) where there is no GPU; the GPU parity tests proper are in test_gpu_parity.py."""
import numpy as np
import pytest

import util


@pytest.mark.parametrize("name", util.ROBOT_FILES)
@pytest.mark.parametrize("fast", [0, 1])
def test_solvers_match_oracle(ctrfk, oracle, hostsim, robots, name, fast):
    d = robots[name]
    jv = np.vstack([ctrfk.sample_joints_host(d, util.SEEDS[name], 0, 3000), util.edge_joint_sets(d)])
    for mode, kw in ((ctrfk.MODE_NSAMPLE, dict(nsamples=256)), (ctrfk.MODE_STEP, dict(step=d.max_length / 256.0)),
                     (ctrfk.MODE_NSAMPLE, dict(nsamples=1)), (ctrfk.MODE_NSAMPLE, dict(nsamples=7)),
                     (ctrfk.MODE_STEP, dict(step=0.01)), (ctrfk.MODE_STEP, dict(step=2e-5))):
        o = oracle.fk_batch(d, jv, mode, **kw)
        s = util.hostsim_fk(hostsim, d, jv, mode, fast=fast, **kw)
        assert np.array_equal(o["n"], s["n"])
        assert np.array_equal(o["step"], s["step"])
        dp, ang = util.pose_errors(o["tip"], s["tip"])
        # strict: same arithmetic up to the association of the frame product.  fast: the series
        # form of the SE(3) step is MORE accurate than the reference's closed form, which loses
        # 1-cos(h|u|) to cancellation when h|u| < 1e-7 (section with kappa = 1e-6); the gap is
        # largest for very coarse sampling (N = 1: 5e-11 m).  See DESIGN.md "FAST vs STRICT".
        assert dp.max() < (2e-10 if fast else 2e-14), (mode, kw, dp.max())
        assert ang.max() < (1e-11 if fast else 1e-13), (mode, kw, ang.max())
        assert np.abs(o["alpha_base"] - s["alpha_base"]).max() < 1e-10


@pytest.mark.parametrize("name", util.ROBOT_FILES)
@pytest.mark.parametrize("fast", [4, 5, 6, 7], ids=["with_base_angles", "tip_only", "fine_with_base_angles", "fine_tip_only"])
def test_scaled_loop_matches_oracle(ctrfk, oracle, hostsim, robots, name, fast):
    """FastSweep::run_scaled — the loop the tip, gather, Jacobian (tip forms) and stability kernels run when nobody
    looks at the samples: state in units of the step, tube angles not accumulated in the tip-only form.  Same gate as
    the FAST solver; sample counts and steps exact; base angles (fast = 4) as before.  Coarse sampling mixes hot and
    generic iterations (conversions in both directions), N = 7 and N = 1 take run() itself."""
    d = robots[name]
    jv = np.vstack([ctrfk.sample_joints_host(d, util.SEEDS[name], 0, 3000), util.edge_joint_sets(d)])
    for mode, kw in ((ctrfk.MODE_NSAMPLE, dict(nsamples=256)), (ctrfk.MODE_STEP, dict(step=d.max_length / 256.0)),
                     (ctrfk.MODE_NSAMPLE, dict(nsamples=1)), (ctrfk.MODE_NSAMPLE, dict(nsamples=7)),
                     (ctrfk.MODE_NSAMPLE, dict(nsamples=24)), (ctrfk.MODE_NSAMPLE, dict(nsamples=64)),
                     (ctrfk.MODE_STEP, dict(step=0.01)), (ctrfk.MODE_STEP, dict(step=2e-5))):
        o = oracle.fk_batch(d, jv, mode, **kw)
        s = util.hostsim_fk(hostsim, d, jv, mode, fast=fast, **kw)
        assert np.array_equal(o["n"], s["n"]) and np.array_equal(o["step"], s["step"])
        dp, ang = util.pose_errors(o["tip"], s["tip"])
        assert dp.max() < 2e-10 and ang.max() < 1e-11, (mode, kw, dp.max(), ang.max())
        if fast in (4, 6):
            assert np.abs(o["alpha_base"] - s["alpha_base"]).max() < 1e-10
        # ... and against the loop it replaces: the same arithmetic up to the placement of the factor h
        r = util.hostsim_fk(hostsim, d, jv, mode, fast=1, **kw)
        dp, ang = util.pose_errors(r["tip"], s["tip"])
        assert dp.max() < 1e-12 and ang.max() < 1e-12, (mode, kw, dp.max(), ang.max())


def test_scaled_loop_pose_does_not_depend_on_the_base_angle_output(ctrfk, hostsim, robots):
    """Both forms of run_scaled (tube angles accumulated or not) give the same bits: generic iterations take their
    angles modulo 2 pi from the carried sin/cos in both — so ctr_fk_batch (which can return alpha_base) and the
    zero-copy host call (which cannot) agree bit for bit, coarse sampling included."""
    d = robots["ctr_sec3_tub5.xml"]
    jv = ctrfk.sample_joints_host(d, 5, 0, 4000)
    for n in (24, 48, 128, 256):
        for with_al, without in ((4, 5), (6, 7)):              # both forms of the loop (coarse / FINE)
            a = util.hostsim_fk(hostsim, d, jv, ctrfk.MODE_NSAMPLE, nsamples=n, fast=with_al)
            b = util.hostsim_fk(hostsim, d, jv, ctrfk.MODE_NSAMPLE, nsamples=n, fast=without)
            assert np.array_equal(a["tip"], b["tip"])
    # where every iteration is hot (fine sampling) the two forms are the same arithmetic
    a = util.hostsim_fk(hostsim, d, jv, ctrfk.MODE_NSAMPLE, nsamples=256, fast=5)
    b = util.hostsim_fk(hostsim, d, jv, ctrfk.MODE_NSAMPLE, nsamples=256, fast=7)
    assert np.array_equal(a["tip"], b["tip"])


def test_coarse_steps_take_the_closed_form_branch(ctrfk, oracle, hostsim, robots):
    """h|u| > 0.5 rad per step: FAST leaves the series for the closed form."""
    d = robots["ctr_sec2_tub3.xml"]
    jv = ctrfk.sample_joints_host(d, 11, 0, 500)
    for ns in (2, 3, 5):
        o = oracle.fk_batch(d, jv, ctrfk.MODE_NSAMPLE, nsamples=ns)
        s = util.hostsim_fk(hostsim, d, jv, ctrfk.MODE_NSAMPLE, nsamples=ns, fast=1)
        dp, ang = util.pose_errors(o["tip"], s["tip"])
        assert dp.max() < 1e-12 and ang.max() < 1e-12


def test_step_mode_dropped_tip_sample_is_reproduced(ctrfk, oracle, hostsim, robots):
    d = ctrfk.RobotDesc.from_buffer_copy(robots["ctr_sec3_tub4.xml"])
    d.max_samples = 64
    jv = ctrfk.sample_joints_host(d, 1234, 0, 4000)
    step = d.max_length / 64.0
    o = oracle.fk_batch(d, jv, ctrfk.MODE_STEP, step=step)
    for fast in (0, 1):
        s = util.hostsim_fk(hostsim, d, jv, ctrfk.MODE_STEP, step=step, fast=fast)
        assert np.array_equal(o["n"], s["n"])
        dp, ang = util.pose_errors(o["tip"], s["tip"])
        assert dp.max() < (2e-10 if fast else 1e-13) and ang.max() < 1e-11


def test_phi_out_of_range_is_flagged(ctrfk, hostsim, robots):
    d = robots["ctr_sec2_tub3.xml"]
    jv = ctrfk.sample_joints_host(d, 1, 0, 4)
    jv[1, 1] = 0.07      # > Length
    jv[2, 4] = -1e-9
    s = util.hostsim_fk(hostsim, d, jv, ctrfk.MODE_NSAMPLE, nsamples=16)
    assert list(s["n"]) == [16, 0, 0, 16]
    assert np.isnan(s["tip"][1]).all() and np.isfinite(s["tip"][0]).all()


def test_fast_sincos_accuracy(hostsim):
    import mpmath as mp
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.uniform(-50, 50, 20000), rng.uniform(-1e4, 1e4, 5000), np.linspace(-7, 7, 3001),
                        np.array([0.0, np.pi / 4, np.pi / 2, np.pi, 1.5 * np.pi, 2 * np.pi, 1e-300, -1e-8])])
    s = np.empty_like(x); c = np.empty_like(x)
    hostsim.hostsim_sincos(x.ctypes.data, x.size, s.ctypes.data, c.ctypes.data)
    # libm is within 1 ulp; allow 2 ulp total against it, and check a subset against mpmath
    assert np.abs(s - np.sin(x)).max() < 3e-16 and np.abs(c - np.cos(x)).max() < 3e-16
    mp.mp.dps = 40
    for i in range(0, x.size, 997):
        assert abs(mp.sin(mp.mpf(float(x[i]))) - mp.mpf(float(s[i]))) < 2.3e-16
        assert abs(mp.cos(mp.mpf(float(x[i]))) - mp.mpf(float(c[i]))) < 2.3e-16
    # huge / non-finite arguments go to libm
    y = np.array([1e9, -3e7, np.inf, np.nan]); s2 = np.empty(4); c2 = np.empty(4)
    hostsim.hostsim_sincos(y.ctypes.data, 4, s2.ctypes.data, c2.ctypes.data)
    assert s2[0] == np.sin(1e9) and c2[1] == np.cos(-3e7) and np.isnan(s2[2]) and np.isnan(c2[3])


def test_se3_step_series_against_mpmath(hostsim):
    import mpmath as mp
    mp.mp.dps = 40
    rng = np.random.default_rng(1)
    for _ in range(200):
        u = rng.uniform(-60, 60, 3) * rng.choice([1.0, 1e-3, 1e-7])
        h = rng.choice([3.9e-4, 1e-3, 5e-3, 2e-2])
        st = np.zeros(12); f = np.zeros(7)
        hostsim.hostsim_se3(u[0], u[1], u[2], h, st.ctypes.data, f.ctypes.data)
        ux, uy, uz = [mp.mpf(float(v)) for v in u]
        n = mp.sqrt(ux * ux + uy * uy + uz * uz)
        t = mp.mpf(float(h)) * n
        B = (1 - mp.cos(t)) / n ** 2
        Cc = (t - mp.sin(t)) / n ** 3
        p = [B * uy + Cc * ux * uz, -B * ux + Cc * uy * uz, mp.mpf(float(h)) - Cc * (ux * ux + uy * uy)]
        q = [mp.cos(t / 2)] + [mp.sin(t / 2) / n * v for v in (ux, uy, uz)]
        for i in range(3):
            assert abs(mp.mpf(float(f[4 + i])) - p[i]) < 1e-18 + 3e-16 * abs(p[i])
        for i in range(4):
            assert abs(mp.mpf(float(f[i])) - q[i]) < 3e-16


# FP32 variant: state arithmetic in float, control scalars in FP64.  Tolerance against the FP64
# oracle (SURVEY §8d proposal, confirmed by measurement: max 2e-5 m / 2e-4 rad on 20 k random
# configurations per robot): 1e-4 m, 1e-3 rad; sample count and step stay bit-exact.
TOL_F32_POS, TOL_F32_ROT = 1e-4, 1e-3


@pytest.mark.parametrize("name", util.ROBOT_FILES)
def test_f32_solver_logic(ctrfk, oracle, hostsim, robots, name):
    d = robots[name]
    jv = ctrfk.sample_joints_host(d, util.SEEDS[name], 0, 5000)
    for mode, kw in ((ctrfk.MODE_NSAMPLE, dict(nsamples=256)), (ctrfk.MODE_STEP, dict(step=d.max_length / 256.0)),
                     (ctrfk.MODE_NSAMPLE, dict(nsamples=9))):
        o = oracle.fk_batch(d, jv, mode, **kw)
        s = util.hostsim_fk(hostsim, d, jv, mode, fast=2, **kw)
        assert np.array_equal(o["n"], s["n"]) and np.array_equal(o["step"], s["step"])
        dp, ang = util.pose_errors(o["tip"], s["tip"])
        assert dp.max() < TOL_F32_POS and ang.max() < TOL_F32_ROT, (dp.max(), ang.max())
        if mode == ctrfk.MODE_NSAMPLE:
            # packed variant (two configurations per thread, solve_f32x2), odd batch: last pair is half empty
            for jvp in (jv, jv[:4999]):
                op = {k: v[:jvp.shape[0]] for k, v in o.items()}
                p = util.hostsim_fk(hostsim, d, jvp, mode, fast=3, **kw)
                assert np.array_equal(op["n"], p["n"]) and np.array_equal(op["step"], p["step"])
                dp, ang = util.pose_errors(op["tip"], p["tip"])
                assert dp.max() < TOL_F32_POS and ang.max() < TOL_F32_ROT, ("packed", dp.max(), ang.max())


def _hostsim_shoot(hostsim, d, jvb, mode, guess=None, step=0.0, nsamples=0, tol=1e-11, max_iter=60, rk4=0):
    import ctypes as C
    B = jvb.shape[0]
    jt = np.zeros_like(jvb); it = np.zeros(B, np.int32); st = np.zeros(B, np.int32); res = np.zeros(B)
    rc = hostsim.hostsim_shoot(C.addressof(d), jvb.ctypes.data, None if guess is None else guess.ctypes.data, B, mode,
                               step, nsamples, tol, max_iter, rk4, jt.ctypes.data, it.ctypes.data, st.ctypes.data,
                               res.ctypes.data)
    assert rc == 0
    return jt, it, st, res


@pytest.mark.parametrize("name", util.ROBOT_FILES)
def test_shooting_solve_logic(ctrfk, oracle, hostsim, robots, name):
    """Base-angle shooting, CPU build of the device code (ShootState: damped Newton with a
    backtracking line search, one sweep per round).  Its checker is the reference sweep itself: the oracle, run at
    the tip angles found, must end at the wanted base angles.  Inputs are base angles produced by the oracle from
    random tip angles, so a root exists; robots this curved have several roots for many inputs, so from the naive
    start x = beta the root found need not be the one the input came from."""
    d = robots[name]
    B = 400
    jv = ctrfk.sample_joints_host(d, 77, 0, B)
    o = oracle.fk_batch(d, jv, ctrfk.MODE_NSAMPLE, nsamples=128)
    jvb = util.base_angle_inputs(d, jv, o["alpha_base"])
    jt, it, st, res = _hostsim_shoot(hostsim, d, jvb, ctrfk.MODE_NSAMPLE, nsamples=128, max_iter=100)
    ok = st == 0
    assert ok.mean() >= 0.98 and set(np.unique(st)) <= {0, 3, 4}     # restarts: see ShootState
    o2 = oracle.fk_batch(d, jt, ctrfk.MODE_NSAMPLE, nsamples=128)
    r2 = np.abs(o2["alpha_base"][:, 1:] - o["alpha_base"][:, 1:]).max(axis=1)
    assert r2[ok].max() < 1e-10                       # the reference sweep confirms every root
    assert np.abs(r2 - res).max() < 1e-9               # ... and the residual reported for every configuration, root or not
    assert it[ok].max() > it[ok].min() + 3             # sweep counts really diverge
    assert it.mean() < 14.0                            # restarts are paid for by the few configurations that need them
    idx = util.alpha_indices(d)
    keep = [i for i in range(jv.shape[1]) if i not in idx[1:]]
    assert np.array_equal(jt[:, keep], jv[:, keep])    # only alphas 1.. change


@pytest.mark.parametrize("name", util.ROBOT_FILES)
def test_shooting_solve_warm_start_recovers_the_branch(ctrfk, oracle, hostsim, robots, name):
    """Branch tracking: started from a guess near the tip angles the base angles came from (what a controller has
    from the previous cycle), the solve must return THOSE tip angles, not another root of the same base angles.
    Configurations that sit on a fold of the base-angle map (singular Jacobian at the root: the robot is about to
    snap) are excluded by the conditioning of the oracle's own finite-difference Jacobian."""
    d = robots[name]
    B = 300
    T = d.ntubes
    idx = util.alpha_indices(d)
    jv = ctrfk.sample_joints_host(d, 78, 0, B)
    rng = np.random.default_rng(5)
    o = oracle.fk_batch(d, jv, ctrfk.MODE_NSAMPLE, nsamples=128)
    jvb = util.base_angle_inputs(d, jv, o["alpha_base"])
    guess = jv.copy()
    guess[:, idx[1:]] += rng.uniform(-0.02, 0.02, size=(B, T - 1))
    jt, it, st, res = _hostsim_shoot(hostsim, d, jvb, ctrfk.MODE_NSAMPLE, guess=guess, nsamples=128)
    # conditioning of d base / d tip at the true root (central differences of the oracle)
    eps = 1e-6
    smin = np.zeros(B)
    J = np.zeros((B, T - 1, T - 1))
    for c in range(T - 1):
        qp = jv.copy(); qm = jv.copy()
        qp[:, idx[c + 1]] += eps; qm[:, idx[c + 1]] -= eps
        J[:, :, c] = (oracle.fk_batch(d, qp, ctrfk.MODE_NSAMPLE, nsamples=128)["alpha_base"][:, 1:]
                      - oracle.fk_batch(d, qm, ctrfk.MODE_NSAMPLE, nsamples=128)["alpha_base"][:, 1:]) / (2 * eps)
    smin = np.linalg.svd(J, compute_uv=False)[:, -1]
    regular = smin > 0.05                               # a 0.02 rad guess error stays inside Newton's basin
    assert regular.mean() > 0.5
    assert (st[regular] == 0).mean() >= 0.99, np.unique(st[regular], return_counts=True)
    good = regular & (st == 0)
    err = np.abs(jt[good][:, idx[1:]] - jv[good][:, idx[1:]]).max(axis=1)
    assert (err < 1e-8).mean() >= 0.99, np.sort(err)[-5:]
    assert it[good].mean() < 5.0                        # 3-4 sweeps from a nearby guess


@pytest.mark.parametrize("name", util.ROBOT_FILES)
def test_rk4_sweep_matches_python_checker(ctrfk, hostsim, robots, name):
    """CTR_FK_FLAG_RK4's sweep and its variational Jacobian (CPU build of the device code) against
    oracle/ctr_shoot_rk4_py.py: same equations, written independently, derivative by complex step."""
    import ctypes as C
    import ctr_oracle_py as py
    import ctr_shoot_rk4_py as rk
    d = robots[name]
    model = py.model_from_desc(d)
    T = d.ntubes
    jv = ctrfk.sample_joints_host(d, 31, 0, 6)
    for q in jv:
        for mode, kw in ((ctrfk.MODE_NSAMPLE, dict(nsamples=48)), (ctrfk.MODE_STEP, dict(step=d.max_length / 40.0))):
            base = np.zeros(T); jac = np.zeros((T - 1, T - 1))
            rc = hostsim.hostsim_tangent(C.addressof(d), q.ctypes.data, mode, kw.get("step", 0.0), kw.get("nsamples", 0), 1,
                                         base.ctypes.data, jac.ctypes.data)
            assert rc == 0
            b, J = rk.base_angles_and_jacobian(model, q, mode, **kw)
            assert np.abs(base - b).max() < 1e-11, (base, b)
            assert np.abs(jac - J).max() < 1e-9 * max(1.0, np.abs(J).max()), (jac, J)


@pytest.mark.parametrize("name", util.ROBOT_FILES)
def test_rk4_sweep_converges_to_the_reference_sweep(ctrfk, oracle, hostsim, robots, name):
    """What ties the RK4 option to the reference: both integrate the same rod equations over the same arc interval
    [h, ArcEnd], the reference by explicit Euler.  Their base angles differ by O(h): the gap halves when the sample
    count doubles (medians over the configurations: these robots are close to snapping for many inputs, where the
    base angles are very sensitive, so the maximum is dominated by a few of them)."""
    import ctypes as C
    d = robots[name]
    T = d.ntubes
    idx = util.alpha_indices(d)
    jv = ctrfk.sample_joints_host(d, 32, 0, 120)
    jv[:, idx[0]] = 0.0                                  # the RK4 model measures every angle from tube 0
    d2 = ctrfk.RobotDesc.from_buffer_copy(d)
    d2.max_samples = 4096

    def rk4(n):
        out = np.zeros((jv.shape[0], T))
        for i, q in enumerate(jv):
            jac = np.zeros((T - 1, T - 1))
            assert hostsim.hostsim_tangent(C.addressof(d2), q.ctypes.data, ctrfk.MODE_NSAMPLE, 0.0, n, 1,
                                           out[i].ctypes.data, jac.ctypes.data) == 0
        return out

    gaps = []
    for n in (128, 256, 512, 1024):
        e = oracle.fk_batch(d2, jv, ctrfk.MODE_NSAMPLE, nsamples=n)["alpha_base"]
        gaps.append(np.median(np.abs(e - rk4(n)).max(axis=1)))
    for a, b in zip(gaps, gaps[1:]):
        assert 1.5 < a / b < 3.0, gaps                  # first order: Euler's error, not RK4's


@pytest.mark.parametrize("name", ["ctr_sec2_tub3.xml", "ctr_sec3_tub5.xml"])
def test_rk4_sweep_is_fourth_order(ctrfk, hostsim, robots, name):
    """The RK4 sweep's own error, measured against the same arc interval integrated with every step cut into 16
    parts (Python checker): fourth order — because a step is split where a section's curved interval starts or ends;
    stepping across those jumps leaves a first-order error as large as Euler's (checker with split=False)."""
    import ctypes as C
    import ctr_oracle_py as py
    import ctr_shoot_rk4_py as rk
    d = robots[name]
    model = py.model_from_desc(d)
    T = d.ntubes
    idx = util.alpha_indices(d)
    jv = ctrfk.sample_joints_host(d, 33, 0, 10)
    jv[:, idx[0]] = 0.0
    err = {32: [], 64: []}
    nosplit = []
    for q in jv:
        for n in (32, 64):
            base = np.zeros(T); jac = np.zeros((T - 1, T - 1))
            assert hostsim.hostsim_tangent(C.addressof(d), q.ctypes.data, ctrfk.MODE_NSAMPLE, 0.0, n, 1,
                                           base.ctypes.data, jac.ctypes.data) == 0
            fine = np.real(rk.base_angles(model, q.astype(complex), ctrfk.MODE_NSAMPLE, nsamples=n, refine=16)[0])
            err[n].append(np.abs(base - fine).max())
            if n == 64:
                ns = np.real(rk.base_angles(model, q.astype(complex), ctrfk.MODE_NSAMPLE, nsamples=n, split=False)[0])
                nosplit.append(np.abs(ns - fine).max())
    assert max(err[64]) < 2e-5 and np.median(err[64]) < 1e-7, (err[64])
    assert np.median(err[32]) / np.median(err[64]) > 10.0            # 2^4 = 16 in the limit
    assert np.median(nosplit) > 1e4 * np.median(err[64])


@pytest.mark.parametrize("name", ["ctr_sec2_tub3.xml", "ctr_sec3_tub4.xml"])
def test_rk4_shooting_roots_approach_the_reference_roots(ctrfk, oracle, hostsim, robots, name):
    """CTR_FK_FLAG_RK4 end to end on the CPU build: (a) its roots are roots of the Python checker's RK4 sweep;
    (b) started from the reference-exact tip angles they stay on that branch and differ from them by O(h)."""
    import ctr_oracle_py as py
    import ctr_shoot_rk4_py as rk
    d = robots[name]
    idx = util.alpha_indices(d)
    model = py.model_from_desc(d)
    jv = ctrfk.sample_joints_host(d, 79, 0, 150)
    jv[:, idx[0]] = 0.0
    dist = []
    for n in (64, 128, 256):
        o = oracle.fk_batch(d, jv, ctrfk.MODE_NSAMPLE, nsamples=n)
        jvb = util.base_angle_inputs(d, jv, o["alpha_base"])
        jt, it, st, res = _hostsim_shoot(hostsim, d, jvb, ctrfk.MODE_NSAMPLE, guess=jv, nsamples=n, rk4=1)
        ok = st == 0
        assert ok.mean() > 0.7                          # the rest sits near a fold: no RK4 root on the reference's branch
        if n == 64:
            for i in np.flatnonzero(ok)[:8]:
                b, _, _ = rk.base_angles(model, jt[i].astype(complex), ctrfk.MODE_NSAMPLE, nsamples=n)
                assert np.abs(np.real(b[1:]) - jvb[i, idx[1:]]).max() < 1e-10
        dist.append(np.abs(jt[:, idx[1:]] - jv[:, idx[1:]]).max(axis=1))
    both = (dist[0] < 0.2) & (dist[1] < 0.2) & (dist[2] < 0.2)          # same branch at every resolution
    assert both.mean() > 0.6
    m = [np.median(x[both]) for x in dist]
    assert m[0] > m[1] > m[2] and 2.5 < m[0] / m[2] < 6.0, m          # O(h): a factor 4 from N = 64 to N = 256


@pytest.mark.parametrize("name", util.ROBOT_FILES)
def test_tangent_sweep_jacobian(ctrfk, oracle, hostsim, robots, name):
    """The variational Jacobian d(base angles)/d(tip angles) of the shooting solve against central
    differences of the ORACLE's base angles, and its base angles against the oracle's."""
    import ctypes as C
    d = robots[name]
    T = d.ntubes
    idx = util.alpha_indices(d)
    jv = ctrfk.sample_joints_host(d, 21, 0, 25)
    for q in jv:
        for mode, kw in ((ctrfk.MODE_NSAMPLE, dict(nsamples=96)), (ctrfk.MODE_STEP, dict(step=d.max_length / 80.0))):
            base = np.zeros(T); jac = np.zeros((T - 1, T - 1))
            rc = hostsim.hostsim_tangent(C.addressof(d), q.ctypes.data, mode, kw.get("step", 0.0), kw.get("nsamples", 0), 0,
                                         base.ctypes.data, jac.ctypes.data)
            assert rc == 0
            o = oracle.fk_one(d, q, mode, **kw)
            assert np.abs(base - o["alpha_base"]).max() < 1e-11
            eps = 1e-6
            for c in range(T - 1):
                qp = q.copy(); qm = q.copy()
                qp[idx[c + 1]] += eps; qm[idx[c + 1]] -= eps
                fd = (oracle.fk_one(d, qp, mode, **kw)["alpha_base"][1:] - oracle.fk_one(d, qm, mode, **kw)["alpha_base"][1:]) / (2 * eps)
                assert np.abs(jac[:, c] - fd).max() < 1e-6 * max(1.0, np.abs(fd).max()), (c, jac[:, c], fd)


@pytest.mark.parametrize("name", util.ROBOT_FILES)
def test_frame_emission_matches_oracle(ctrfk, oracle, hostsim, robots, name):
    """FastSweep::run_emit + PendingFrameEmitter (the backbone kernel's second sweep): frames
    obtained as I_ktip = T, I_{k-1} = I_k inv(E_k) must be the oracle's frames — every frame, not
    only the tip — including step mode with dropped tip samples and single-sample sweeps."""
    import ctypes as C
    d = robots[name]
    jv = np.vstack([ctrfk.sample_joints_host(d, util.SEEDS[name] + 5, 0, 300), util.edge_joint_sets(d)])
    B = jv.shape[0]
    hostsim.hostsim_emit_full.restype = C.c_int
    hostsim.hostsim_emit_full.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_double, C.c_int, C.c_void_p]
    for mode, step, ns in ((ctrfk.MODE_NSAMPLE, 0.0, 256), (ctrfk.MODE_STEP, d.max_length / 256.0, 0),
                           (ctrfk.MODE_NSAMPLE, 0.0, 1), (ctrfk.MODE_NSAMPLE, 0.0, 2), (ctrfk.MODE_NSAMPLE, 0.0, 31),
                           (ctrfk.MODE_NSAMPLE, 0.0, 33), (ctrfk.MODE_NSAMPLE, 0.0, 64), (ctrfk.MODE_NSAMPLE, 0.0, 65),
                           (ctrfk.MODE_NSAMPLE, 0.0, 225), (ctrfk.MODE_STEP, 0.004, 0), (ctrfk.MODE_STEP, 0.03, 0)):
        o = oracle.fk_batch(d, jv, mode, step=step, nsamples=ns, backbone=True)
        for fn in (hostsim.hostsim_emit_full,):
            got = np.full((B, d.max_samples, 12), np.nan)
            assert fn(C.addressof(d), jv.ctypes.data, B, mode, step, ns, got.ctypes.data) == 0
            worst_p = worst_r = 0.0
            for i in range(B):
                nf = int(o["n"][i])
                assert np.isfinite(got[i, :nf]).all(), (mode, step, ns, i)
                assert np.isnan(got[i, nf:]).all(), (mode, step, ns, i)       # nothing beyond the last frame
                dp, ang = util.pose_errors(o["backbone"][i, :nf], got[i, :nf])
                worst_p = max(worst_p, dp.max()); worst_r = max(worst_r, ang.max())
            assert worst_p < 2e-10 and worst_r < 1e-11, (mode, step, ns, worst_p, worst_r)


@pytest.mark.parametrize("name", util.ROBOT_FILES)
def test_jacobian_config_all_forms(ctrfk, oracle, hostsim, robots, name):
    """jacobian_config (the per-thread routine of fk_jacobian_kernel) for every calcJacobian form of the
    reference: NSamples form, ArcLengthStep form (robot_i.h:901-907), iSample form (:1012-1072)."""
    import ctypes as C
    d = robots[name]
    B = 12
    jv = ctrfk.sample_joints_host(d, 17, 0, B)
    jv[0, 1] = d.sec[0].length
    hostsim.hostsim_jacobian.restype = C.c_int
    hostsim.hostsim_jacobian.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_double, C.c_int, C.c_int, C.c_double,
                                         C.c_double] + [C.c_void_p] * 5 + [C.c_int]
    for mode, kw, ks in ((1, dict(nsamples=100), -1), (0, dict(step=d.max_length / 200.0), -1), (1, dict(nsamples=100), 37),
                         (1, dict(nsamples=100), 0), (1, dict(nsamples=100), 99)):
        jac = np.zeros((B, d.njoints, 6)); jacs = np.zeros((B, d.njoints, 6)); tip = np.zeros((B, 12)); smp = np.zeros((B, 12))
        n = np.zeros(B, np.int32)
        rc = hostsim.hostsim_jacobian(C.addressof(d), jv.ctypes.data, B, mode, kw.get("step", 0.0), kw.get("nsamples", 0), ks, 1e-5,
                                      1e-4, jac.ctypes.data, jacs.ctypes.data if ks >= 0 else None, tip.ctypes.data,
                                      smp.ctypes.data if ks >= 0 else None, n.ctypes.data, 0)
        assert rc == 0
        if mode == 1:       # the forms that take the base poses from the caller (robot_i.h:917, :1012): same columns
            jac2 = np.zeros_like(jac); jacs2 = np.zeros_like(jacs)
            rc = hostsim.hostsim_jacobian(C.addressof(d), jv.ctypes.data, B, mode, 0.0, kw["nsamples"], ks, 1e-5, 1e-4,
                                          jac2.ctypes.data, jacs2.ctypes.data if ks >= 0 else None, tip.ctypes.data,
                                          smp.ctypes.data if ks >= 0 else None, None, 1)
            assert rc == 0 and np.array_equal(jac2, jac) and np.array_equal(jacs2, jacs)
        for i in range(B):
            o = oracle.jacobian_ex(d, jv[i], mode, i_sample=ks, **kw)
            assert n[i] == o["n"]
            dp, ang = util.pose_errors(o["tip"][None], tip[i][None])
            assert dp.max() <= util.TOL_POS and ang.max() <= util.TOL_ROT
            # finite differences divide the 1e-13-level arithmetic differences by the step
            assert np.abs(o["jac"] - jac[i]).max() < 1e-6
            if ks >= 0:
                dp, ang = util.pose_errors(o["sample"][None], smp[i][None])
                assert dp.max() <= util.TOL_POS and ang.max() <= util.TOL_ROT
                assert np.abs(o["jac_sample"] - jacs[i]).max() < 1e-6


def test_fast_solver_on_degenerate_robots(ctrfk, oracle, hostsim):
    """The FAST per-thread solver (CPU build) on the robots of test_fk_bit_equal_degenerate_robots: zero curvature
    (the |u| < zero_tol branch with its h*|u| translation), one straight section, Poisson ratio 0, stiffness ratios
    of 1e6, a 1 mm section — against the oracle, which equals the reference build on exactly these."""
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
            jv = ctrfk.sample_joints_host(d, 500, 0, 60)
            for mode, kw in ((ctrfk.MODE_NSAMPLE, dict(nsamples=96)), (ctrfk.MODE_STEP, dict(step=d.max_length / 50.0))):
                o = oracle.fk_batch(d, jv, mode, **kw)
                for fast in (0, 1):
                    s_ = util.hostsim_fk(hostsim, d, jv, mode, fast=fast, **kw)
                    assert np.array_equal(s_["n"], o["n"]) and np.array_equal(s_["step"], o["step"])
                    dp, ang = util.pose_errors(o["tip"], s_["tip"])
                    assert dp.max() <= util.TOL_POS and ang.max() <= util.TOL_ROT, (layout, variant, fast, dp.max(), ang.max())
