"""ctypes binding of libctrfk.so (include/ctr_fk.h) for tests, bench.py and __graft_entry__.

This is plumbing: it mirrors the C structs, loads the in-tree shared library and passes raw
pointers (torch tensors' data_ptr() on the device, numpy arrays on the host).  There is no Python
implementation of the kinematics here and no fallback: if the library is missing, `load()` raises.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libctrfk.so")
REPO = os.path.dirname(HERE)

MAX_SECTIONS = 3
MAX_TUBES = 6
MODE_STEP, MODE_NSAMPLE = 0, 1
FLAG_JAC_GIVEN_POSES = 0x100
FLAG_GATHER_MULTICAST = 0x200
FLAG_GATHER_ROWWISE = 0x400
FLAG_HOST_STAGED = 0x800
FLAG_RK4 = 0x1000
FLAG_F32 = 0x2000
FLAG_HOST_ZEROCOPY = 0x4000
HOST_WRITE_COMBINED = 1
FLAG_STRICT, FLAG_NO_BINNING, FLAG_SOA = 1, 2, 4
E_NULL, E_DESC, E_ARG, E_XML, E_NOMEM, E_RANGE = -1, -2, -3, -4, -5, -6
STATUS_OK, STATUS_PHI_RANGE = 0, 1
SCALE_NONE, SCALE_GAMMA, SCALE_PROPORTIONAL = 0, 1, 2
SHOOT_OK, SHOOT_PHI_RANGE, SHOOT_SINGULAR, SHOOT_MAX_ITER, SHOOT_STALLED = 0, 1, 2, 3, 4


class SectionDesc(C.Structure):
    _fields_ = [("ntube", C.c_int32), ("reserved", C.c_int32), ("length", C.c_double),
                ("stiffness", C.c_double * 2), ("curvature", C.c_double * 2),
                ("radius", C.c_double * 2), ("poisson", C.c_double * 2)]


class RobotDesc(C.Structure):
    _fields_ = [("nsections", C.c_int32), ("max_samples", C.c_int32),
                ("init_trafo", C.c_double * 12), ("zero_tol", C.c_double),
                ("sec", SectionDesc * MAX_SECTIONS)]

    @property
    def ntubes(self):
        return sum(self.sec[i].ntube for i in range(self.nsections))

    @property
    def njoints(self):
        return 1 + self.nsections + self.ntubes

    @property
    def max_length(self):
        acc = 0.0
        for i in range(self.nsections):
            acc += self.sec[i].length
        return acc


# every symbol include/ctr_fk.h declares: name -> (restype, argtypes)
_P = C.c_void_p
SYMBOLS = {
    "ctr_fk_desc_from_xml": (C.c_int, [C.c_char_p, C.c_int32, C.POINTER(RobotDesc), _P]),
    "ctr_fk_desc_validate": (C.c_int, [C.POINTER(RobotDesc)]),
    "ctr_fk_desc_ntubes": (C.c_int, [C.POINTER(RobotDesc)]),
    "ctr_fk_desc_njoints": (C.c_int, [C.POINTER(RobotDesc)]),
    "ctr_fk_desc_max_length": (C.c_double, [C.POINTER(RobotDesc)]),
    "ctr_fk_orthogonalize": (C.c_int, [_P, C.c_double]),
    "ctr_fk_create": (C.c_int, [C.POINTER(RobotDesc), C.c_int, C.POINTER(_P)]),
    "ctr_fk_destroy": (C.c_int, [_P]),
    "ctr_fk_get_desc": (C.c_int, [_P, C.POINTER(RobotDesc)]),
    "ctr_fk_batch": (C.c_int, [_P, _P, C.c_int64, C.c_int, C.c_double, C.c_int32, C.c_uint32,
                               _P, _P, _P, _P, _P, _P, _P, _P]),
    "ctr_fk_batch_f32": (C.c_int, [_P, _P, C.c_int64, C.c_int, C.c_double, C.c_int32, C.c_uint32,
                                   _P, _P, _P, _P, _P]),
    "ctr_fk_batch_host": (C.c_int, [_P, _P, C.c_int64, C.c_int, C.c_double, C.c_int32, C.c_uint32,
                                    _P, _P, _P, _P]),
    "ctr_fk_single": (C.c_int, [_P, _P, C.c_int, C.c_double, C.c_int32, C.c_uint32,
                                _P, _P, _P, _P, _P, _P]),
    "ctr_fk_batch_host_backbone": (C.c_int, [_P, _P, C.c_int64, C.c_int, C.c_double, C.c_int32, C.c_uint32, _P, _P, _P, _P, _P, _P]),
    "ctr_fk_batch_gather": (C.c_int, [_P, _P, C.c_int64, C.c_int, C.c_double, C.c_int32, C.c_uint32, _P, C.c_int32, C.c_int64,
                                      _P, _P, _P, _P]),
    "ctr_fk_peer_alloc": (C.c_int, [_P, C.c_size_t, C.POINTER(_P), _P]),
    "ctr_fk_peer_open": (C.c_int, [_P, _P, C.POINTER(_P)]),
    "ctr_fk_peer_close": (C.c_int, [_P, _P]),
    "ctr_fk_peer_free": (C.c_int, [_P, _P]),
    "ctr_fk_jacobian_batch": (C.c_int, [_P, _P, C.c_int64, C.c_int32, C.c_double, C.c_double, _P, _P, _P]),
    "ctr_fk_jacobian_host": (C.c_int, [_P, _P, C.c_int64, C.c_int32, C.c_double, C.c_double, _P, _P]),
    "ctr_fk_jacobian_batch_ex": (C.c_int, [_P, _P, C.c_int64, C.c_int, C.c_double, C.c_int32, C.c_int32, C.c_double, C.c_double,
                                           C.c_uint32, _P, _P, _P, _P, _P, _P]),
    "ctr_fk_jacobian_host_ex": (C.c_int, [_P, _P, C.c_int64, C.c_int, C.c_double, C.c_int32, C.c_int32, C.c_double, C.c_double,
                                          C.c_uint32, _P, _P, _P, _P, _P]),
    "ctr_fk_stability_batch": (C.c_int, [_P, _P, C.c_int64, C.c_double, C.c_double, C.c_double, _P, _P, _P]),
    "ctr_fk_stability_host": (C.c_int, [_P, _P, C.c_int64, C.c_double, C.c_double, C.c_double, _P, _P]),
    "ctr_fk_stability_batch_ex": (C.c_int, [_P, _P, C.c_int64, C.c_double, C.c_double, C.c_double, C.c_uint32, _P, _P, _P]),
    "ctr_fk_stability_host_ex": (C.c_int, [_P, _P, C.c_int64, C.c_double, C.c_double, C.c_double, C.c_uint32, _P, _P]),
    "ctr_fk_batch_base": (C.c_int, [_P, _P, C.c_int64, C.c_int, C.c_double, C.c_int32, C.c_double, C.c_int32,
                                    _P, _P, _P, _P, _P]),
    "ctr_fk_batch_base_ex": (C.c_int, [_P, _P, _P, C.c_int64, C.c_int, C.c_double, C.c_int32, C.c_double, C.c_int32,
                                       C.c_uint32, _P, _P, _P, _P, _P, _P]),
    "ctr_fk_sample_joints": (C.c_int, [_P, C.c_uint64, C.c_int64, C.c_int64, _P, _P]),
    "ctr_fk_sample_joints_scaled": (C.c_int, [_P, C.c_uint64, C.c_int64, C.c_int64, C.c_int, C.c_double, C.c_double,
                                              C.c_int32, _P, _P]),
    "ctr_fk_sample_joints_scaled_host": (C.c_int, [C.POINTER(RobotDesc), C.c_uint64, C.c_int64, C.c_int64, C.c_int,
                                                   C.c_double, C.c_double, C.c_int32, _P]),
    "ctr_fk_sample_joints_host": (C.c_int, [C.POINTER(RobotDesc), C.c_uint64, C.c_int64, C.c_int64, _P]),
    "ctr_fk_count_samples_host": (C.c_int, [C.POINTER(RobotDesc), _P, C.c_int64, C.c_int, C.c_double,
                                            C.c_int32, _P]),
    "ctr_fk_host_alloc": (C.c_int, [_P, C.c_size_t, C.c_uint32, C.POINTER(_P)]),
    "ctr_fk_host_free": (C.c_int, [_P, _P]),
    "ctr_fk_host_local_cpus": (C.c_int, [_P, C.c_char_p, C.c_size_t]),
    "ctr_fk_fp64_probe": (C.c_int, [C.c_int, C.c_int32, C.POINTER(C.c_double), _P]),
    "ctr_fk_fp32_probe": (C.c_int, [C.c_int, C.c_int32, C.POINTER(C.c_double), _P]),
    "ctr_fk_launch_count": (C.c_int64, []),
    "ctr_fk_last_error": (C.c_char_p, []),
    "ctr_fk_version": (C.c_char_p, []),
}

_lib = None


def load():
    """Load libctrfk.so (raises if it was not built — there is no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            # not a fallback: build the CUDA library in-tree (nvcc cross-compiles without a GPU)
            import subprocess
            try:
                subprocess.check_call(["make", "-s", "-j", str(min(8, os.cpu_count() or 1)), "-C", HERE, "libctrfk.so"])
            except Exception as e:
                raise RuntimeError(
                    "libctrfk.so is not built and `make -C ram2017-ctr-kinematics_b200` failed (%r); "
                    "this package has no non-CUDA path" % (e,))
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


class CtrFkError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("ctr_fk error %d: %s" % (code, msg))
        self.code = code


def last_error():
    return load().ctr_fk_last_error().decode("utf-8", "replace")


def _check(rc):
    if rc != 0:
        raise CtrFkError(rc, last_error())


def resource(name):
    """Path of a robot XML shipped in tests/data (copies of nothing: generated by
    tests/golden/make_robot_xml.py from the parameter tables in SURVEY.md §9.2)."""
    path = os.path.join(REPO, "tests", "data", name)
    if not os.path.exists(path) and name.startswith("ctr_"):      # BASELINE.json names the robots ctr_*.xml
        path = os.path.join(REPO, "tests", "data", name[len("ctr_"):])
    return path


def desc_from_xml(path, max_samples=256, want_phi=False):
    d = RobotDesc()
    phi = (C.c_double * MAX_SECTIONS)()
    _check(load().ctr_fk_desc_from_xml(os.fsencode(path), max_samples, C.byref(d), C.cast(phi, _P)))
    return (d, list(phi)[: d.nsections]) if want_phi else d


def _ptr(x):
    """Raw pointer of a torch tensor / numpy array / int / None."""
    if x is None:
        return None
    if isinstance(x, int):
        return x
    if hasattr(x, "data_ptr"):
        return x.data_ptr()
    if hasattr(x, "ctypes"):
        return x.ctypes.data
    raise TypeError(type(x))


class Engine:
    """Owns a ctr_fk_handle."""

    def __init__(self, desc, device=0):
        self.desc = desc
        self.handle = _P()
        _check(load().ctr_fk_create(C.byref(desc), device, C.byref(self.handle)))
        self.device = device
        self.njv = desc.njoints
        self.ntubes = desc.ntubes

    def close(self):
        if self.handle:
            load().ctr_fk_destroy(self.handle)
            self.handle = _P()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # device pointers
    def batch(self, jv, B, mode, step=0.0, nsamples=0, flags=0, tip=None, backbone=None, radius=None,
              n_out=None, step_out=None, alpha_base=None, status=None, stream=None):
        _check(load().ctr_fk_batch(self.handle, _ptr(jv), B, mode, step, nsamples, flags, _ptr(tip),
                                   _ptr(backbone), _ptr(radius), _ptr(n_out), _ptr(step_out),
                                   _ptr(alpha_base), _ptr(status), stream))

    def batch_f32(self, jv, B, mode, step=0.0, nsamples=0, flags=0, tip=None, n_out=None, step_out=None,
                  status=None, stream=None):
        _check(load().ctr_fk_batch_f32(self.handle, _ptr(jv), B, mode, step, nsamples, flags, _ptr(tip),
                                       _ptr(n_out), _ptr(step_out), _ptr(status), stream))

    # host pointers
    def batch_host(self, jv, B, mode, step=0.0, nsamples=0, flags=0, tip=None, n_out=None, step_out=None,
                   status=None):
        _check(load().ctr_fk_batch_host(self.handle, _ptr(jv), B, mode, step, nsamples, flags, _ptr(tip),
                                        _ptr(n_out), _ptr(step_out), _ptr(status)))

    def host_alloc(self, nbytes, write_combined=False):
        """Page-locked, device-mapped host memory on the GPU's NUMA node (ctr_fk_host_alloc); returns the address."""
        p = _P()
        _check(load().ctr_fk_host_alloc(self.handle, nbytes, HOST_WRITE_COMBINED if write_combined else 0, C.byref(p)))
        return p.value

    def host_free(self, ptr):
        _check(load().ctr_fk_host_free(self.handle, ptr))

    def host_array(self, shape, dtype="float64", write_combined=False):
        """numpy view of a fresh ctr_fk_host_alloc buffer (freed when the view's base dies or by host_free)."""
        import numpy as np
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        ptr = self.host_alloc(max(n, 16), write_combined)
        buf = (C.c_ubyte * max(n, 16)).from_address(ptr)
        arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
        return arr, ptr

    def host_local_cpus(self):
        buf = C.create_string_buffer(4096)
        _check(load().ctr_fk_host_local_cpus(self.handle, buf, 4096))
        return buf.value.decode()

    def single(self, jv, mode, step=0.0, nsamples=0, flags=0, tip=None, backbone=None, radius=None,
               n_out=None, step_out=None, alpha_base=None):
        rc = load().ctr_fk_single(self.handle, _ptr(jv), mode, step, nsamples, flags, _ptr(tip),
                                  _ptr(backbone), _ptr(radius), _ptr(n_out), _ptr(step_out),
                                  _ptr(alpha_base))
        if rc not in (0, E_RANGE):
            _check(rc)
        return rc

    def batch_host_backbone(self, jv, B, mode, backbone, step=0.0, nsamples=0, flags=0, tip=None, radius=None, n_out=None,
                            step_out=None, status=None):
        _check(load().ctr_fk_batch_host_backbone(self.handle, _ptr(jv), B, mode, step, nsamples, flags, _ptr(tip), _ptr(backbone),
                                                 _ptr(radius), _ptr(n_out), _ptr(step_out), _ptr(status)))

    def batch_gather(self, jv, B, mode, dst_ptrs, first_row, step=0.0, nsamples=0, flags=0, multicast=False, n_out=None,
                     step_out=None, status=None, stream=None):
        """Solve + all-gather in one kernel: dst_ptrs = device addresses (ints) of every rank's gathered buffer as
        mapped in this process, or [multicast address] with multicast=True."""
        arr = (_P * len(dst_ptrs))(*[int(p) for p in dst_ptrs])
        _check(load().ctr_fk_batch_gather(self.handle, _ptr(jv), B, mode, step, nsamples,
                                          flags | (FLAG_GATHER_MULTICAST if multicast else 0), arr, len(dst_ptrs), first_row,
                                          _ptr(n_out), _ptr(step_out), _ptr(status), stream))

    def peer_alloc(self, nbytes):
        """(device pointer, 64-byte IPC handle) of a buffer other ranks can map with peer_open."""
        p = _P()
        hd = (C.c_ubyte * 64)()
        _check(load().ctr_fk_peer_alloc(self.handle, nbytes, C.byref(p), C.cast(hd, _P)))
        return p.value, bytes(hd)

    def peer_open(self, handle64):
        p = _P()
        buf = (C.c_ubyte * 64).from_buffer_copy(handle64)
        _check(load().ctr_fk_peer_open(self.handle, C.cast(buf, _P), C.byref(p)))
        return p.value

    def peer_close(self, ptr):
        _check(load().ctr_fk_peer_close(self.handle, ptr))

    def peer_free(self, ptr):
        _check(load().ctr_fk_peer_free(self.handle, ptr))

    def jacobian(self, jv, B, nsamples, fd_trans, fd_rot, jac, tip=None, stream=None):
        _check(load().ctr_fk_jacobian_batch(self.handle, _ptr(jv), B, nsamples, fd_trans, fd_rot,
                                            _ptr(jac), _ptr(tip), stream))

    def jacobian_host(self, jv, B, nsamples, fd_trans, fd_rot, jac, tip=None):
        _check(load().ctr_fk_jacobian_host(self.handle, _ptr(jv), B, nsamples, fd_trans, fd_rot,
                                           _ptr(jac), _ptr(tip)))

    def jacobian_ex(self, jv, B, mode, fd_trans, fd_rot, jac, step=0.0, nsamples=0, i_sample=-1, jac_sample=None,
                    tip=None, sample=None, n_out=None, given=False, flags=0, stream=None):
        """Every calcJacobian form (include/ctr_fk.h: ctr_fk_jacobian_batch_ex), device pointers."""
        _check(load().ctr_fk_jacobian_batch_ex(self.handle, _ptr(jv), B, mode, step, nsamples, i_sample, fd_trans, fd_rot,
                                               flags | (FLAG_JAC_GIVEN_POSES if given else 0), _ptr(jac), _ptr(jac_sample), _ptr(tip),
                                               _ptr(sample), _ptr(n_out), stream))

    def jacobian_host_ex(self, jv, B, mode, fd_trans, fd_rot, jac, step=0.0, nsamples=0, i_sample=-1, jac_sample=None,
                         tip=None, sample=None, n_out=None, given=False, flags=0):
        _check(load().ctr_fk_jacobian_host_ex(self.handle, _ptr(jv), B, mode, step, nsamples, i_sample, fd_trans, fd_rot,
                                              flags | (FLAG_JAC_GIVEN_POSES if given else 0), _ptr(jac), _ptr(jac_sample), _ptr(tip),
                                              _ptr(sample), _ptr(n_out)))

    def stability(self, jv, B, angle_step, arc_step, min_degree=-1.0, stable=None, degree=None, stream=None, flags=0):
        _check(load().ctr_fk_stability_batch_ex(self.handle, _ptr(jv), B, angle_step, arc_step, min_degree, flags,
                                                _ptr(stable), _ptr(degree), stream))

    def batch_base(self, jv_base, B, mode, step=0.0, nsamples=0, tol=1e-11, max_iter=100, jv_tip=None, tip=None,
                   iters=None, status=None, stream=None, jv_guess=None, resid=None, flags=0):
        """ctr_fk_batch_base_ex; flags: FLAG_RK4.  jv_guess=None starts from x = beta (with restarts)."""
        _check(load().ctr_fk_batch_base_ex(self.handle, _ptr(jv_base), _ptr(jv_guess), B, mode, step, nsamples, tol, max_iter,
                                           flags, _ptr(jv_tip), _ptr(tip), _ptr(iters), _ptr(status), _ptr(resid), stream))

    def sample_joints(self, seed, first, B, jv, stream=None, policy=SCALE_NONE, p0=1.0, p1=1.0, lut_size=0):
        _check(load().ctr_fk_sample_joints_scaled(self.handle, seed, first, B, policy, p0, p1, lut_size,
                                                  _ptr(jv), stream))


def sample_joints_host(desc, seed, first, B, policy=SCALE_NONE, p0=1.0, p1=1.0, lut_size=0):
    import numpy as np
    jv = np.empty((B, desc.njoints), dtype=np.float64)
    _check(load().ctr_fk_sample_joints_scaled_host(C.byref(desc), seed, first, B, policy, p0, p1, lut_size,
                                                   jv.ctypes.data))
    return jv


def count_samples_host(desc, jv, mode, step=0.0, nsamples=0):
    import numpy as np
    jv = np.ascontiguousarray(jv, dtype=np.float64)
    n = np.empty(jv.shape[0], dtype=np.int32)
    _check(load().ctr_fk_count_samples_host(C.byref(desc), jv.ctypes.data, jv.shape[0], mode, step,
                                            nsamples, n.ctypes.data))
    return n


def fp64_probe(device, iters, stream=None):
    flops = C.c_double(0.0)
    _check(load().ctr_fk_fp64_probe(device, iters, C.byref(flops), stream))
    return flops.value


def fp32_probe(device, iters, stream=None):
    flops = C.c_double(0.0)
    _check(load().ctr_fk_fp32_probe(device, iters, C.byref(flops), stream))
    return flops.value


def launch_count():
    return int(load().ctr_fk_launch_count())


# Algorithmic flops per sample-step of the reference algorithm (SURVEY.md §8d / BASELINE.md §3):
# backward 29T + 2S - 9 arithmetic + 2T trig, forward 139 (tip only) / 229 (backbone).
def alg_flops_per_step(desc, backbone=False):
    T, S = desc.ntubes, desc.nsections
    return (29 * T + 2 * S - 9 + 2 * T) + (229 if backbone else 139)


ALG_FLOPS_ONCE = 100


def shard_range(total, world, rank):
    """Contiguous slice [lo, hi) of a batch of `total` configurations owned by `rank` out of
    `world` ranks: ceil(total/world) each, the last ranks possibly short or empty.
    The batch shards with no exchange during the solve; each rank solves its slice alone."""
    per = -(-total // world)
    lo = min(total, rank * per)
    hi = min(total, lo + per)
    return lo, hi
