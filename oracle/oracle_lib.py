"""ctypes binding of the CPU oracle (oracle/libctr_oracle.so).  TEST INFRASTRUCTURE ONLY:
import from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libctr_oracle.so")
_lib = None
_P = C.c_void_p


def build():
    subprocess.check_call(["make", "-s", "-C", HERE])


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        lib = C.CDLL(LIB_PATH)
        lib.ctr_oracle_fk.restype = C.c_int
        lib.ctr_oracle_fk.argtypes = [_P, _P, C.c_int, C.c_double, C.c_int32, _P, _P, _P, _P, _P, _P, _P, _P]
        lib.ctr_oracle_fk_batch.restype = C.c_int
        lib.ctr_oracle_fk_batch.argtypes = [_P, _P, C.c_int64, C.c_int, C.c_double, C.c_int32, C.c_int,
                                            _P, _P, _P, _P, _P, _P, _P]
        lib.ctr_oracle_jacobian.restype = C.c_int
        lib.ctr_oracle_jacobian.argtypes = [_P, _P, C.c_int32, C.c_double, C.c_double, _P, _P]
        lib.ctr_oracle_jacobian_ex.restype = C.c_int
        lib.ctr_oracle_jacobian_ex.argtypes = [_P, _P, C.c_int, C.c_double, C.c_int32, C.c_int32, C.c_double, C.c_double,
                                               _P, _P, _P, _P, _P]
        lib.ctr_oracle_stability.restype = C.c_int
        lib.ctr_oracle_stability.argtypes = [_P, _P, C.c_double, C.c_double]
        lib.ctr_oracle_degree_stability.restype = C.c_int
        lib.ctr_oracle_degree_stability.argtypes = [_P, _P, C.c_double, C.c_double, C.c_double, _P, _P]
        lib.ctr_oracle_max_threads.restype = C.c_int
        lib.ctr_oracle_jacobian_batch.restype = C.c_int
        lib.ctr_oracle_jacobian_batch.argtypes = [_P, _P, C.c_int64, C.c_int, C.c_double, C.c_int32, C.c_int32, C.c_double,
                                                  C.c_double, C.c_int, _P, _P, _P, _P, _P]
        lib.ctr_oracle_stability_batch.restype = C.c_int
        lib.ctr_oracle_stability_batch.argtypes = [_P, _P, C.c_int64, C.c_double, C.c_double, C.c_double, C.c_int, _P, _P]
        _lib = lib
    return _lib


def max_threads():
    return load().ctr_oracle_max_threads()


def fk_batch(desc, jv, mode, step=0.0, nsamples=0, nthreads=None, backbone=False, radius=False):
    """desc: ctypes ctr_robot_desc mirror (any struct with the same layout).  Returns a dict."""
    jv = np.ascontiguousarray(jv, dtype=np.float64)
    B = jv.shape[0]
    T = sum(desc.sec[i].ntube for i in range(desc.nsections))
    maxs = desc.max_samples
    out = dict(tip=np.full((B, 12), np.nan), n=np.zeros(B, np.int32), step=np.zeros(B),
               alpha_base=np.zeros((B, T)), status=np.zeros(B, np.int32))
    bb = np.zeros((B, maxs, 12)) if backbone else None
    rd = np.zeros((B, maxs)) if radius else None
    rc = load().ctr_oracle_fk_batch(C.addressof(desc), jv.ctypes.data, B, mode, step, nsamples,
                                    nthreads or max_threads(), out["tip"].ctypes.data,
                                    bb.ctypes.data if backbone else None, rd.ctypes.data if radius else None,
                                    out["n"].ctypes.data, out["step"].ctypes.data,
                                    out["alpha_base"].ctypes.data, out["status"].ctypes.data)
    if rc != 0:
        raise RuntimeError("oracle error %d" % rc)
    if backbone:
        out["backbone"] = bb
    if radius:
        out["radius"] = rd
    return out


def fk_one(desc, jv, mode, step=0.0, nsamples=0):
    jv = np.ascontiguousarray(jv, dtype=np.float64)
    T = sum(desc.sec[i].ntube for i in range(desc.nsections))
    maxs = desc.max_samples
    tip = np.zeros(12); bb = np.zeros((maxs, 12)); rd = np.zeros(maxs)
    n = C.c_int32(0); h = C.c_double(0.0)
    curv = np.zeros((maxs, 3)); al = np.zeros(maxs); ab = np.zeros(T)
    st = load().ctr_oracle_fk(C.addressof(desc), jv.ctypes.data, mode, step, nsamples, tip.ctypes.data,
                              bb.ctypes.data, rd.ctypes.data, C.addressof(n), C.addressof(h),
                              curv.ctypes.data, al.ctypes.data, ab.ctypes.data)
    if st < 0:
        raise RuntimeError("oracle error %d" % st)
    k = n.value
    return dict(status=st, tip=tip, backbone=bb[:k], radius=rd[:k], n=k, step=h.value, curv=curv[:k],
                alpha=al[:k], alpha_base=ab)


def jacobian(desc, jv, nsamples, fd_trans, fd_rot):
    jv = np.ascontiguousarray(jv, dtype=np.float64)
    njv = jv.shape[0]
    jac = np.zeros((njv, 6)); tip = np.zeros(12)
    rc = load().ctr_oracle_jacobian(C.addressof(desc), jv.ctypes.data, nsamples, fd_trans, fd_rot,
                                    jac.ctypes.data, tip.ctypes.data)
    if rc != 0:
        raise RuntimeError("oracle jacobian error %d" % rc)
    return jac, tip


def jacobian_ex(desc, jv, mode, step=0.0, nsamples=0, i_sample=-1, fd_trans=1e-5, fd_rot=1e-4):
    """All calcJacobian forms; returns dict(jac, jac_sample, tip, sample, n)."""
    jv = np.ascontiguousarray(jv, dtype=np.float64)
    njv = jv.shape[0]
    jac = np.zeros((njv, 6)); js = np.zeros((njv, 6)); tip = np.zeros(12); smp = np.zeros(12); n = C.c_int32(0)
    rc = load().ctr_oracle_jacobian_ex(C.addressof(desc), jv.ctypes.data, mode, step, nsamples, i_sample, fd_trans, fd_rot,
                                       jac.ctypes.data, js.ctypes.data, tip.ctypes.data, smp.ctypes.data, C.addressof(n))
    if rc != 0:
        raise RuntimeError("oracle jacobian_ex error %d" % rc)
    return dict(jac=jac, jac_sample=js if i_sample >= 0 else None, tip=tip, sample=smp if i_sample >= 0 else None, n=n.value)


def stability(desc, jv, angle_step, arc_step):
    jv = np.ascontiguousarray(jv, dtype=np.float64)
    return load().ctr_oracle_stability(C.addressof(desc), jv.ctypes.data, angle_step, arc_step)


def degree_stability(desc, jv, angle_step, arc_step, min_degree=-1.0):
    jv = np.ascontiguousarray(jv, dtype=np.float64)
    deg = C.c_double(0.0); st = C.c_int(0)
    rc = load().ctr_oracle_degree_stability(C.addressof(desc), jv.ctypes.data, angle_step, arc_step, min_degree,
                                            C.addressof(deg), C.addressof(st))
    if rc != 0:
        raise RuntimeError("oracle degree_stability error %d" % rc)
    return deg.value, st.value


def jacobian_batch(desc, jv, mode, step=0.0, nsamples=0, i_sample=-1, fd_trans=1e-5, fd_rot=1e-4, nthreads=None):
    """calcJacobian over B configurations (OpenMP); returns dict(jac [B][njv][6], jac_sample, tip, n, rc)."""
    jv = np.ascontiguousarray(jv, dtype=np.float64)
    B, njv = jv.shape
    jac = np.zeros((B, njv, 6)); js = np.zeros((B, njv, 6)) if i_sample >= 0 else None
    tip = np.zeros((B, 12)); n = np.zeros(B, np.int32); rc = np.zeros(B, np.int32)
    r = load().ctr_oracle_jacobian_batch(C.addressof(desc), jv.ctypes.data, B, mode, step, nsamples, i_sample,
                                         fd_trans, fd_rot, nthreads or max_threads(), jac.ctypes.data,
                                         js.ctypes.data if js is not None else None, tip.ctypes.data, n.ctypes.data, rc.ctypes.data)
    if r != 0:
        raise RuntimeError("oracle jacobian_batch error %d" % r)
    return dict(jac=jac, jac_sample=js, tip=tip, n=n, rc=rc)


def stability_batch(desc, jv, angle_step, arc_step, min_degree=-1.0, nthreads=None):
    """calcDegreeStability over B configurations (OpenMP); returns (stable [B] int32 (< 0: error), degree [B])."""
    jv = np.ascontiguousarray(jv, dtype=np.float64)
    B = jv.shape[0]
    st = np.zeros(B, np.int32); dg = np.zeros(B)
    r = load().ctr_oracle_stability_batch(C.addressof(desc), jv.ctypes.data, B, angle_step, arc_step,
                                          min_degree, nthreads or max_threads(), st.ctypes.data, dg.ctypes.data)
    if r != 0:
        raise RuntimeError("oracle stability_batch error %d" % r)
    return st, dg

