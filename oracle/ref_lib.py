"""ctypes binding of oracle/_ref/libctr_ref.so — the UNMODIFIED reference sources
(/root/reference/ctr_source, compiled where they lie) behind a small C driver (oracle/ref_driver.cpp),
built against the stand-in third-party headers of oracle/shim.  TEST INFRASTRUCTURE ONLY: import
from tests/, tests/golden/make_reference_golden.py and bench.py's cpu_baseline / --impl reference
legs.  /root/reference exists only in the build container; on the GPU box the prebuilt .so is used."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_ref", "libctr_ref.so")
REFERENCE_SRC = "/root/reference/ctr_source"
_lib = None
_P = C.c_void_p


def can_build():
    return os.path.isdir(REFERENCE_SRC)


def build():
    subprocess.check_call(["make", "-s", "-j3", "-C", HERE, "ref"])


def available():
    return os.path.exists(LIB_PATH) or can_build()


def load():
    global _lib
    if _lib is None:
        if can_build():
            build()                      # make: no-op when up to date
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("oracle/_ref/libctr_ref.so is missing and /root/reference is not here to build it")
        lib = C.CDLL(LIB_PATH)
        lib.ctr_ref_create_xml.restype = _P
        lib.ctr_ref_create_xml.argtypes = [C.c_char_p, C.c_int64]
        lib.ctr_ref_create_desc.restype = _P
        lib.ctr_ref_create_desc.argtypes = [_P]
        lib.ctr_ref_destroy.argtypes = [_P]
        lib.ctr_ref_info.argtypes = [_P] * 6
        lib.ctr_ref_last_state.argtypes = [_P] * 4
        lib.ctr_ref_fk_batch.argtypes = [_P, _P, C.c_int64, C.c_int, C.c_double, C.c_int32, C.c_int, _P, _P, _P, _P, _P]
        lib.ctr_ref_jacobian.argtypes = [_P, _P, C.c_int32, C.c_double, C.c_double, _P, _P]
        lib.ctr_ref_stability.argtypes = [_P, _P, C.c_double, C.c_double]
        lib.ctr_ref_degree_stability.argtypes = [_P, _P, C.c_double, C.c_double, C.c_double, _P, _P]
        lib.ctr_ref_random_joints.argtypes = [_P, C.c_int64, _P]
        lib.ctr_ref_max_threads.restype = C.c_int
        lib.ctr_ref_create_desc_f32.restype = _P
        lib.ctr_ref_create_desc_f32.argtypes = [_P]
        lib.ctr_ref_destroy_f32.argtypes = [_P]
        lib.ctr_ref_fk_batch_f32.argtypes = [_P, _P, C.c_int64, C.c_int, C.c_double, C.c_int32, _P, _P]
        lib.ctr_ref_adapter_desc.argtypes = [_P, _P]
        lib.ctr_ref_jacobian_step.argtypes = [_P, _P, C.c_double, C.c_double, C.c_double, _P, _P, _P]
        lib.ctr_ref_jacobian_sample.argtypes = [_P, _P, C.c_int32, C.c_int32, C.c_double, C.c_double, _P, _P]
        lib.ctr_ref_gamma_scale.argtypes = [C.c_int, C.c_double, _P, C.c_int64, _P]
        lib.ctr_ref_proportional_scale.argtypes = [_P, C.c_double, C.c_double, _P, C.c_int64, _P, _P]
        _lib = lib
    return _lib


class RefRobot:
    """CTR::Robot<double> of the reference.  from_desc: container constructor (ctr_robot.h:96-98) with the
    descriptor's sections and entry frame; from_xml: Robot(MaxSample, XmlFilename) (ctr_robot.h:105)."""

    def __init__(self, handle):
        if not handle:
            raise RuntimeError("reference robot construction failed")
        self.h = handle
        ns, nt, nj, ms = (C.c_int32() for _ in range(4))
        ml = C.c_double()
        load().ctr_ref_info(self.h, C.addressof(ns), C.addressof(nt), C.addressof(nj), C.addressof(ms), C.addressof(ml))
        self.nsections, self.ntubes, self.njoints, self.max_samples, self.max_length = ns.value, nt.value, nj.value, ms.value, ml.value

    @classmethod
    def from_desc(cls, desc):
        return cls(load().ctr_ref_create_desc(C.addressof(desc)))

    @classmethod
    def from_xml(cls, path, max_samples):
        return cls(load().ctr_ref_create_xml(os.fsencode(path), max_samples))

    def __del__(self):
        if getattr(self, "h", None) and _lib is not None:
            _lib.ctr_ref_destroy(self.h)
            self.h = None

    def last_state(self):
        n = C.c_int32()
        tip = np.zeros(12); jv = np.zeros(self.njoints)
        load().ctr_ref_last_state(self.h, C.addressof(n), tip.ctypes.data, jv.ctypes.data)
        return n.value, tip, jv

    def fk_batch(self, jv, mode, step=0.0, nsamples=0, nthreads=None, backbone=False, radius=False, tip=True):
        jv = np.ascontiguousarray(jv, dtype=np.float64)
        B = jv.shape[0]
        assert jv.shape[1] == self.njoints
        out = dict(n=np.zeros(B, np.int32), step=np.zeros(B))
        if tip:
            out["tip"] = np.full((B, 12), np.nan)
        bb = np.zeros((B, self.max_samples, 12)) if backbone else None
        rd = np.zeros((B, self.max_samples)) if radius else None
        rc = load().ctr_ref_fk_batch(self.h, jv.ctypes.data, B, mode, step, nsamples, nthreads or max_threads(),
                                     out["tip"].ctypes.data if tip else None, bb.ctypes.data if backbone else None,
                                     rd.ctypes.data if radius else None, out["n"].ctypes.data, out["step"].ctypes.data)
        if rc != 0:
            raise RuntimeError("reference error %d" % rc)
        if backbone:
            out["backbone"] = bb
        if radius:
            out["radius"] = rd
        return out

    def jacobian(self, jv, nsamples, fd_trans, fd_rot):
        jv = np.ascontiguousarray(jv, dtype=np.float64)
        jac = np.zeros((self.njoints, 6)); tip = np.zeros(12)
        rc = load().ctr_ref_jacobian(self.h, jv.ctypes.data, nsamples, fd_trans, fd_rot, jac.ctypes.data, tip.ctypes.data)
        if rc != 0:
            raise RuntimeError("reference jacobian error %d" % rc)
        return jac, tip

    def jacobian_step(self, jv, step, fd_trans, fd_rot):
        """calcJacobian(JV, ArcLengthStep, ...) (robot_i.h:901-907); returns (jac, base tip, m_Samples)."""
        jv = np.ascontiguousarray(jv, dtype=np.float64)
        jac = np.zeros((self.njoints, 6)); tip = np.zeros(12); n = C.c_int32(0)
        rc = load().ctr_ref_jacobian_step(self.h, jv.ctypes.data, step, fd_trans, fd_rot, jac.ctypes.data, tip.ctypes.data, C.addressof(n))
        if rc != 0:
            raise RuntimeError("reference jacobian_step error %d" % rc)
        return jac, tip, n.value

    def jacobian_sample(self, jv, nsamples, i_sample, fd_trans, fd_rot):
        """calcJacobian(JV, NSamples, ..., iSample, JTip, JiSample) (robot_i.h:908-915, :1012-1072)."""
        jv = np.ascontiguousarray(jv, dtype=np.float64)
        jac = np.zeros((self.njoints, 6)); js = np.zeros((self.njoints, 6))
        rc = load().ctr_ref_jacobian_sample(self.h, jv.ctypes.data, nsamples, i_sample, fd_trans, fd_rot, jac.ctypes.data, js.ctypes.data)
        if rc != 0:
            raise RuntimeError("reference jacobian_sample error %d" % rc)
        return jac, js

    def stability(self, jv, angle_step, arc_step):
        jv = np.ascontiguousarray(jv, dtype=np.float64)
        return load().ctr_ref_stability(self.h, jv.ctypes.data, angle_step, arc_step)

    def degree_stability(self, jv, angle_step, arc_step, min_degree=-1.0):
        jv = np.ascontiguousarray(jv, dtype=np.float64)
        deg = C.c_double(0.0); st = C.c_int(0)
        rc = load().ctr_ref_degree_stability(self.h, jv.ctypes.data, angle_step, arc_step, min_degree,
                                             C.addressof(deg), C.addressof(st))
        if rc != 0:
            raise RuntimeError("reference degree_stability error %d" % rc)
        return deg.value, st.value

    def adapter_desc(self, desc_type):
        """include/ctr_fk_reference_adapter.hpp applied to this live robot; returns a `desc_type` (ctrfk.RobotDesc)."""
        d = desc_type()
        rc = load().ctr_ref_adapter_desc(self.h, C.addressof(d))
        if rc != 0:
            raise RuntimeError("adapter error %d" % rc)
        return d

    def random_joints(self, B):
        """B consecutive getRandomJointValues draws from a default std::mt19937 (ctr_test.cpp:30-32)."""
        jv = np.zeros((B, self.njoints))
        load().ctr_ref_random_joints(self.h, B, jv.ctypes.data)
        return jv


def max_threads():
    return load().ctr_ref_max_threads()


def fk_batch_f32(desc, jv, mode, step=0.0, nsamples=0):
    """CTR::Robot<float> (the reference's single-precision instantiation, ctr_robot.cpp:724-731; sin/cos from libm's
    sinf/cosf where real Eigen uses its SSE approximations).  Returns dict(tip [B][12] as double, n)."""
    jv = np.ascontiguousarray(jv, dtype=np.float64)
    h = load().ctr_ref_create_desc_f32(C.addressof(desc))
    if not h:
        raise RuntimeError("reference float robot construction failed")
    B = jv.shape[0]
    tip = np.zeros((B, 12)); n = np.zeros(B, np.int32)
    rc = load().ctr_ref_fk_batch_f32(h, jv.ctypes.data, B, mode, step, nsamples, tip.ctypes.data, n.ctypes.data)
    load().ctr_ref_destroy_f32(h)
    if rc != 0:
        raise RuntimeError("reference f32 error %d" % rc)
    return dict(tip=tip, n=n)


def gamma_scale(lut_size, gamma, u):
    """GammaCorrection<double, lut_size>::scaleTrans after set_scale(gamma) (transrotscale.h:106-122)."""
    u = np.ascontiguousarray(u, dtype=np.float64)
    out = np.zeros_like(u)
    rc = load().ctr_ref_gamma_scale(lut_size, gamma, u.ctypes.data, u.size, out.ctypes.data)
    if rc != 0:
        raise RuntimeError("reference gamma_scale error %d (table size not instantiated?)" % rc)
    return out


def proportional_scale(robot, p0, p1, u):
    """ProportionalScale<double> (transrotscale.h:139-309) on u[B][nsections]; returns (scaled, first goal scale)."""
    u = np.ascontiguousarray(u, dtype=np.float64)
    out = np.zeros_like(u)
    g = C.c_double()
    rc = load().ctr_ref_proportional_scale(robot.h, p0, p1, u.ctypes.data, u.shape[0], out.ctypes.data, C.addressof(g))
    if rc != 0:
        raise RuntimeError("reference proportional_scale error %d" % rc)
    return out, g.value
