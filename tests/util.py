"""Shared helpers of the test suite."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
ROBOT_FILES = ["ctr_sec2_tub3.xml", "ctr_sec3_tub3.xml", "ctr_sec3_tub4.xml", "ctr_sec3_tub5.xml"]
SEEDS = {"ctr_sec2_tub3.xml": 20171, "ctr_sec3_tub3.xml": 20172, "ctr_sec3_tub4.xml": 20173,
         "ctr_sec3_tub5.xml": 20174}

# Parity gate of BASELINE.json north_star: tip position within 1e-9 m, orientation within 1e-9 rad.
TOL_POS = 1e-9
TOL_ROT = 1e-9


def load_hostsim():
    """CPU build of the kernels' per-thread solvers (tests/hostsim) — test harness only."""
    d = os.path.join(HERE, "hostsim")
    so = os.path.join(d, "libhostsim.so")
    src = os.path.join(d, "hostsim.cpp")
    csrc = os.path.join(REPO, "ram2017-ctr-kinematics_b200", "csrc")
    deps = [src] + [os.path.join(csrc, f) for f in ("ctr_fk_core.cuh", "ctr_fk_core_f32.cuh", "ctr_fk_core_f32x2.cuh", "ctr_fk_krobot.h")]
    if (not os.path.exists(so)) or os.path.getmtime(so) < max(os.path.getmtime(d) for d in deps):
        subprocess.check_call(["/usr/bin/g++", "-std=c++14", "-O2", "-march=x86-64-v3", "-ffp-contract=off",
                               "-fPIC", "-shared", "-Wno-unknown-pragmas", "-o", so, src])
    lib = C.CDLL(so)
    lib.hostsim_fk.restype = C.c_int
    lib.hostsim_fk.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_double, C.c_int, C.c_int,
                               C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.hostsim_sincos.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
    lib.hostsim_se3.argtypes = [C.c_double] * 4 + [C.c_void_p] * 2
    lib.hostsim_shoot.restype = C.c_int
    lib.hostsim_shoot.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_double, C.c_int, C.c_double,
                                  C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.hostsim_tangent.restype = C.c_int
    lib.hostsim_tangent.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    return lib


def hostsim_fk(lib, desc, jv, mode, step=0.0, nsamples=0, fast=1):
    jv = np.ascontiguousarray(jv, dtype=np.float64)
    B = jv.shape[0]
    T = sum(desc.sec[i].ntube for i in range(desc.nsections))
    tip = np.full((B, 12), np.nan)
    n = np.zeros(B, np.int32)
    st = np.zeros(B)
    ab = np.full((B, T), np.nan)
    rc = lib.hostsim_fk(C.addressof(desc), jv.ctypes.data, B, mode, step, nsamples, fast, tip.ctypes.data,
                        n.ctypes.data, st.ctypes.data, ab.ctypes.data, None)
    assert rc == 0
    return dict(tip=tip, n=n, step=st, alpha_base=ab)


def pose_errors(tip_a, tip_b):
    """Per-configuration (position error [m], geodesic rotation angle [rad]) between two [B,12]
    column-major 3x4 batches."""
    a = np.asarray(tip_a).reshape(-1, 12)
    b = np.asarray(tip_b).reshape(-1, 12)
    dp = np.linalg.norm(a[:, 9:] - b[:, 9:], axis=1)
    fro = np.linalg.norm(a[:, :9] - b[:, :9], axis=1)           # |Ra - Rb|_F = 2 sqrt2 sin(angle/2)
    ang = 2.0 * np.arcsin(np.clip(fro / (2.0 * np.sqrt(2.0)), 0.0, 1.0))
    return dp, ang


def edge_joint_sets(desc):
    """Edge cases of SURVEY.md §8d: phi = 0, phi = L, all phi = 0, alpha in {0, pi, 2pi^-}."""
    S = desc.nsections
    nj = 1 + S + sum(desc.sec[i].ntube for i in range(S))
    two_pi_m = np.nextafter(2 * np.pi, 0.0)
    phi_idx, alpha_idx, t0 = [], [], 0
    for s in range(S):
        phi_idx.append(1 + s + t0)
        for k in range(desc.sec[s].ntube):
            alpha_idx.append(2 + s + t0 + k)
        t0 += desc.sec[s].ntube
    lens = [desc.sec[s].length for s in range(S)]
    rows = []
    rng = np.random.default_rng(7)
    for phis in ([0.0] * S, lens, [0.5 * l for l in lens]) + tuple([[lens[i] if i == j else 0.0 for i in range(S)] for j in range(S)]):
        for a in (0.0, np.pi, two_pi_m, None):
            q = np.zeros(nj)
            q[0] = 0.7
            for i, p in zip(phi_idx, phis):
                q[i] = p
            for i in alpha_idx:
                q[i] = a if a is not None else rng.uniform(0, two_pi_m)
            rows.append(q)
    return np.array(rows)


def alpha_indices(desc):
    """Joint-vector index of every tube's angle."""
    idx, t0 = [], 0
    for s in range(desc.nsections):
        for k in range(desc.sec[s].ntube):
            idx.append(2 + s + t0 + k)
        t0 += desc.sec[s].ntube
    return idx


def base_angle_inputs(desc, jv_tip, alpha_base):
    """Joint vectors whose alpha slots carry the BASE angles of the given sweep results."""
    jvb = np.array(jv_tip, copy=True)
    idx = alpha_indices(desc)
    for t in range(1, len(idx)):
        jvb[:, idx[t]] = alpha_base[:, t]
    return jvb


def random_robot(ctrfk, rng, ntubes, max_samples=128):
    """A random but physically sensible robot with the given tubes-per-section layout."""
    d = ctrfk.RobotDesc()
    d.nsections = len(ntubes)
    d.max_samples = max_samples
    A = rng.normal(size=(3, 3))
    Q, _ = np.linalg.qr(A)
    if np.linalg.det(Q) < 0:
        Q[:, 0] = -Q[:, 0]
    t = list(Q.T.reshape(-1)) + list(rng.uniform(-0.1, 0.1, 3))        # column-major rotation + translation
    for i in range(12):
        d.init_trafo[i] = t[i]
    for s, n in enumerate(ntubes):
        d.sec[s].ntube = n
        d.sec[s].length = rng.uniform(0.02, 0.08)
        for k in range(n):
            d.sec[s].stiffness[k] = rng.uniform(0.05, 6.0)
            d.sec[s].curvature[k] = rng.uniform(1.0, 60.0)
            d.sec[s].radius[k] = rng.uniform(3e-4, 2e-3)
            d.sec[s].poisson[k] = rng.uniform(0.2, 0.45)               # distinct per tube: exercises the
    return d                                                           # last-VC-section quirk (section_i.h:273)


ALL_LAYOUTS = [(1,), (2,), (1, 1), (1, 2), (2, 1), (2, 2), (1, 1, 1), (1, 1, 2), (1, 2, 1), (1, 2, 2),
               (2, 1, 1), (2, 1, 2), (2, 2, 1), (2, 2, 2)]


def phi_indices(desc):
    """Joint-vector index of every section's phi."""
    idx, t0 = [], 0
    for s in range(desc.nsections):
        idx.append(1 + s + t0)
        t0 += desc.sec[s].ntube
    return idx


def record(kind, entry):
    """Append a measured figure of a GPU test to gpurun_out/test_measurements.jsonl (kept as a profiles/ artifact):
    the tests assert bounds, this keeps what was actually measured."""
    import json
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "test_measurements.jsonl"), "a") as f:
            f.write(json.dumps(dict(kind=kind, **entry)) + "\n")
    except OSError:
        pass

