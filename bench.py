#!/usr/bin/env python
"""bench.py — FK solves/sec (FP64) of the batched concentric-tube-robot forward kinematics.

  python bench.py --gpus N --steps K --warmup W          our CUDA path (one process per GPU)
  python bench.py --impl reference ...                   the reference's own code on the host CPU

A "step" is one pass of the hot path (ctr_fk_batch) over one batch of synthetic joint vectors.
Headline workload = BASELINE.json configs[1]: ctr_sec2_tub3.xml, 1M random joint configurations per GPU,
tip pose out, fixed N=256 samples (SURVEY §8d "Mode F", MaxSamples=256).  Multi-GPU: the batch shards
trivially, each rank solves its own 1M configurations (weak scaling), no data-path collective.

The ONE printed JSON line (rank 0) carries the contract keys for the headline workload plus
  roofline      FP64 FMA pipe (nothing on this path is a dense contraction; tip-only I/O is ~150 B per
                configuration, so neither "tensor" nor "hbm" applies), peak measured live by a DFMA probe
  e2e           the same metric through ctr_fk_batch_host with HOST buffers (copies inside the timed region),
                next to the box's measured PCIe ceiling with all ranks copying at once (`pcie_aggregate`)
  cpu_baseline  the reference's CPU code on a bounded sample (N=1, rank 0)
  configs       BASELINE.json configs[2..4], each timed like the headline with its own roofline and clocks:
                config 3 (ctr_sec3_tub3, 2M per GPU; with more than one rank also with the poses all-gathered:
                NCCL, fused NVLink peer stores, fused NVSwitch multicast), config 4 (ctr_sec3_tub4 backbone +
                radius; HBM and FP64 fractions), config 5 (ctr_sec3_tub5, 12.5M per GPU; FP64 and FP32).
                --configs none skips them.
"""
import argparse
import json
import os
import sys
import threading
import time
import types

REPO = os.path.dirname(os.path.abspath(__file__))
for _p in (os.path.join(REPO, "ram2017-ctr-kinematics_b200"), os.path.join(REPO, "oracle"), os.path.join(REPO, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

METRIC = "FK solves/sec (FP64, 3-tube)"
UNIT = "configurations/s"
SEEDS = {"ctr_sec2_tub3.xml": 20171, "ctr_sec3_tub3.xml": 20172, "ctr_sec3_tub4.xml": 20173, "ctr_sec3_tub5.xml": 20174}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--robot", default="ctr_sec2_tub3.xml")
    ap.add_argument("--batch", type=int, default=1_000_000, help="configurations per GPU per step")
    ap.add_argument("--mode", default="F", choices=["F", "S"], help="F: fixed N samples; S: arc-length step")
    ap.add_argument("--nsamples", type=int, default=256)
    ap.add_argument("--max-samples", type=int, default=256, help="MaxSamples of the robot (install.sh:96 uses 256)")
    ap.add_argument("--strict", action="store_true", help="STRICT arithmetic kernels")
    ap.add_argument("--backbone", action="store_true", help="full backbone + radius output (config 4)")
    ap.add_argument("--soa", action="store_true", help="with --backbone: [MaxS][12][B] output layout (CTR_FK_FLAG_SOA)")
    ap.add_argument("--f32", action="store_true", help="FP32 state arithmetic variant")
    ap.add_argument("--shoot", action="store_true", help="base-angle shooting solve (ctr_fk_batch_base) instead of FK")
    ap.add_argument("--shoot-max-iter", type=int, default=100,
                    help="sweep cap of --shoot (naive start with restarts: 99+ %% solved within 100 sweeps, mean 6-11)")
    ap.add_argument("--shoot-guess", type=float, default=None, metavar="RAD",
                    help="--shoot with a warm start: the guess is the true tip angles + uniform noise of this amplitude")
    ap.add_argument("--shoot-rk4", action="store_true", help="--shoot on the continuum model integrated by RK4 (CTR_FK_FLAG_RK4)")
    ap.add_argument("--allgather", action="store_true", help="time an NCCL all-gather of the tip poses too")
    ap.add_argument("--gather-rowwise", action="store_true", help="--allgather-fused with one thread storing its own row (comparison)")
    ap.add_argument("--gather-replicate", type=int, default=1,
                    help="--allgather-fused ipc/symm: list every destination K times (emulates K times as many peers on a small box)")
    ap.add_argument("--allgather-fused", choices=["ipc", "symm", "multicast", "auto"], default=None,
                    help="solve + all-gather in ONE kernel (ctr_fk_batch_gather): poses stored straight into every rank's "
                         "gathered buffer over NVLink; ipc = buffers shared by CUDA IPC (ctr_fk_peer_*), symm = torch "
                         "symmetric memory, multicast = its NVSwitch multicast address (one multimem.st per 16 bytes), "
                         "auto = multicast when the group has it, else ipc")
    ap.add_argument("--configs", choices=["auto", "none"], default="auto",
                    help="auto: with the default headline workload also time BASELINE configs 3, 4, 5 (`configs` array)")
    ap.add_argument("--config-steps", type=int, default=6, help="timed steps of every `configs` entry")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-pointer end-to-end leg (very large batches)")
    ap.add_argument("--e2e-staged", action="store_true", help="e2e through the chunked copy pipeline (the default form; kept for old command lines)")
    ap.add_argument("--e2e-zerocopy", action="store_true", help="e2e through the zero-copy form (CTR_FK_FLAG_HOST_ZEROCOPY): one kernel reads and writes page-locked host memory")
    ap.add_argument("--e2e-wc", action="store_true", help="e2e: joint vectors in write-combined host memory")
    ap.add_argument("--no-bind", action="store_true", help="do not bind the process to the CPUs local to its GPU")
    ap.add_argument("--cpu-sample", type=int, default=0, help="configurations of the CPU baseline sample (0 = auto)")
    return ap.parse_args()


def workload_of(a, **over):
    """The workload a set of command-line arguments (or a `configs` entry) describes."""
    w = types.SimpleNamespace(
        robot=a.robot, batch=a.batch, mode=a.mode, nsamples=a.nsamples, max_samples=a.max_samples, strict=a.strict,
        backbone=a.backbone, soa=a.soa, f32=a.f32, shoot=a.shoot, shoot_max_iter=a.shoot_max_iter, shoot_guess=a.shoot_guess, shoot_rk4=a.shoot_rk4,
        collective=("nccl" if a.allgather else (a.allgather_fused or "none")),
        gather_rowwise=a.gather_rowwise, gather_replicate=a.gather_replicate, nset=2)
    for k, v in over.items():
        setattr(w, k, v)
    return w


def workload_name(w):
    return "%s, %s configurations/GPU, %s, %s out" % (
        w.robot, "{:,}".format(w.batch).replace(",", "_"),
        ("fixed N=%d samples (calcKinematic(JV, NSamples))" % w.nsamples) if w.mode == "F"
        else "step = MaxLength/%d (calcKinematic(JV, ArcLengthStep))" % w.nsamples,
        ("backbone+radius" + (" (SoA layout)" if getattr(w, "soa", False) else "")) if w.backbone else "tip pose")


class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed region runs."""

    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index):
        self.samples, self.reasons, self.ok, self.max_mhz = [], set(), False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:           # NVML missing: record that, never fake a number
            self.err = repr(e)
        self._stop = threading.Event()
        self._t = None

    def _run(self):
        while not self._stop.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                r = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(self.nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if r & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.002)

    def start(self):
        if self.ok:
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()

    def mark(self):
        """Start of the timed region: forget what was sampled before (the thread is started during
        warm-up because the first NVML queries of a process take tens of milliseconds)."""
        self.samples, self.reasons = [], set()

    def stop(self):
        if self._t:
            self._stop.set()
            self._t.join()
        if not self.ok:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml_unavailable"]}
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


def visible_index(local_rank):
    cvd = os.environ.get("CUDA_VISIBLE_DEVICES")
    if cvd:
        ids = [x.strip() for x in cvd.split(",") if x.strip()]
        try:
            return int(ids[local_rank])
        except Exception:
            return local_rank
    return local_rank


def host_threads():
    """All host cores this process may use.  torchrun exports OMP_NUM_THREADS=1; the CPU arms take
    an explicit thread count (num_threads clause), so that default does not cap them."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def parse_cpulist(s):
    out = set()
    for part in s.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        out.update(range(int(lo), int(hi or lo) + 1))
    return out


# ---------------------------------------------------------------------------------------------------
# CPU arms: the reference's own code (oracle/_ref) or, where that build is absent, the oracle port
# ---------------------------------------------------------------------------------------------------
CPU_ARM_NOTE = {
    "reference": "the reference's own sources (ctr_source/*.h, ctr_robot.cpp) compiled -O3 -march=x86-64-v3 against "
                 "stand-in Eigen/Erl/Boost headers (oracle/_ref/libctr_ref.so), CTR::Robot<double>::calcKinematic, "
                 "one robot per thread, OpenMP over configurations",
    "port": "CPU oracle = step-for-step C restatement of the reference FK (oracle/ctr_oracle.c; "
            "oracle/_ref/libctr_ref.so not present), OpenMP over configurations",
}


def _splitmix64(z):
    import numpy as np
    z = z + np.uint64(0x9E3779B97F4A7C15)
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def sample_joints_numpy(desc, seed, first, B):
    """numpy replica of the counter-based generator of host/ctr_sampler.h (bit-identical:
    tests/test_host.py::test_sampler_matches_numpy_replica_and_distribution) — lets the reference arm
    draw the bench workload's inputs without loading libctrfk.so."""
    import numpy as np
    njv = 1 + desc.nsections + sum(desc.sec[s].ntube for s in range(desc.nsections))
    c = np.nextafter(2 * np.pi, 0.0)
    scale = [c]
    for s in range(desc.nsections):
        scale.append(desc.sec[s].length)
        scale.extend([c] * desc.sec[s].ntube)
    out = np.empty((B, njv), dtype=np.float64)
    with np.errstate(over="ignore"):
        j = np.arange(njv, dtype=np.uint64)[None, :]
        for lo in range(0, B, 1 << 18):
            hi = min(B, lo + (1 << 18))
            i = (np.arange(lo, hi, dtype=np.uint64) + np.uint64(first))[:, None]
            a = _splitmix64(np.uint64(seed) + np.uint64(0x9E3779B97F4A7C15) * i)
            b = _splitmix64(a + j)
            out[lo:hi] = (b >> np.uint64(11)).astype(np.float64) / 9007199254740992.0
    return out * np.array(scale)


class CpuArm:
    """The reference's CPU forward kinematics for one workload, never touching libctrfk.so:
    kind "reference" = CTR::Robot<double> built by its own XML constructor (ctr_robot.h:105) in oracle/_ref;
    kind "port" = the C oracle on a descriptor parsed by the repo's XML reader (only when oracle/_ref is absent)."""

    def __init__(self, w):
        import ctrfk          # module import only: ctypes structs and paths; nothing here calls ctrfk.load()
        import ref_lib
        self.w = w
        xml = ctrfk.resource(w.robot)
        if os.path.exists(ref_lib.LIB_PATH):
            self.kind = "reference"
            self.robot = ref_lib.RefRobot.from_xml(xml, w.max_samples)
            self.desc = self.robot.adapter_desc(ctrfk.RobotDesc)   # through the reference's public interface
            self.max_length = self.robot.max_length
        else:
            self.kind = "port"
            self.desc = ctrfk.desc_from_xml(xml, w.max_samples)    # no reference build here: the repo's XML reader
            self.max_length = self.desc.max_length
        self.mode = 1 if w.mode == "F" else 0

    def inputs(self, seed, first, B):
        return sample_joints_numpy(self.desc, seed, first, B)

    def solve(self, jv, nthreads):
        kw = dict(nsamples=self.w.nsamples, step=self.max_length / self.w.nsamples, nthreads=nthreads,
                  backbone=self.w.backbone, radius=self.w.backbone)
        if self.kind == "reference":
            return self.robot.fk_batch(jv, self.mode, **kw)
        import oracle_lib
        return oracle_lib.fk_batch(self.desc, jv, self.mode, **kw)

    def rate(self, jv, nthreads):
        t0 = time.perf_counter()
        self.solve(jv, nthreads)
        dt = time.perf_counter() - t0
        return jv.shape[0] / dt, dt


def config_keys(w, collective="none"):
    """`config` of a workload — the same dict from both arms (the driver compares them); what is specific to an arm
    (arithmetic, buffers, notes) goes under `arm`."""
    njv = {"ctr_sec2_tub3.xml": 6, "ctr_sec3_tub3.xml": 7, "ctr_sec3_tub4.xml": 8, "ctr_sec3_tub5.xml": 9}.get(w.robot)
    set_mb = None if njv is None else ((njv * 8 + 96) * w.batch + (104.0 * w.max_samples * w.batch if w.backbone else 0)) / 1e6
    return {"workload": workload_name(w), "robot": w.robot, "mode": w.mode, "nsamples": w.nsamples,
            "batch_per_gpu": w.batch,
            "l2": "%s: one step's inputs + outputs are %s MB per GPU against a 126 MB L2; the CUDA arm alternates "
                  "%d such set(s) between steps" % ("no flush needed" if (set_mb or 0) * w.nset > 126 else "NOT L2-cold (small batch)",
                                                    "%.0f" % set_mb if set_mb is not None else "?", w.nset),
            "collective": collective}


def run_reference(a):
    """--impl reference: the reference's CPU algorithm on the box's host cores (all threads), every step one
    full batch of the headline workload; inputs drawn once, outside the timed region.  libctrfk.so is not loaded."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    w = workload_of(a)
    arm = CpuArm(w)
    threads = host_threads()
    seed = SEEDS.get(w.robot, 20170)
    # one full batch per step unless that would run for many minutes: 1M configurations take ~2 s on 16 cores
    probe_rate, _ = arm.rate(arm.inputs(seed, 0, 4096), threads)
    budget_s = 150.0
    per_step_s = budget_s / max(1, a.steps + a.warmup)
    sample = a.cpu_sample or int(min(w.batch, max(8192, probe_rate * per_step_s)))
    jv = arm.inputs(seed, 0, sample)
    for _ in range(a.warmup):
        arm.solve(jv, threads)
    t0 = time.perf_counter()
    for _ in range(a.steps):
        arm.solve(jv, threads)
    dt = time.perf_counter() - t0
    value = sample * a.steps / dt
    cfg = config_keys(w)
    arm_info = {"arithmetic": "reference (IEEE, no FMA contraction)", "note": CPU_ARM_NOTE[arm.kind],
                "sample_per_step": sample,
                "sample_note": "one full batch per step" if sample == w.batch else
                               "bounded sample of the batch per step (a full batch would exceed the time budget on %d cores)" % threads}
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": 1e3 * dt / a.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg, "arm": arm_info,
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": arm.kind,
                             "sample": "%d configurations per step of the bench workload" % sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------------------------------------------
# the CUDA arm
# ---------------------------------------------------------------------------------------------------
class Ctx:
    pass


def barrier(ctx):
    if ctx.world > 1:
        ctx.dist.barrier()
    ctx.torch.cuda.synchronize()


def measure_peaks(ctx):
    """FP64 DFMA and FP32 FFMA2 peak probes (MEASURED_PEAKS.json carries neither)."""
    import ctrfk
    torch = ctx.torch
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    out = {}
    for name, probe in (("fp64", ctrfk.fp64_probe), ("fp32x2", ctrfk.fp32_probe)):
        probe(ctx.local, 200, ctx.stream)
        torch.cuda.synchronize()
        best = 0.0
        for _ in range(3):
            ev0.record()
            flops = probe(ctx.local, 4000, ctx.stream)
            ev1.record()
            torch.cuda.synchronize()
            best = max(best, flops / (ev0.elapsed_time(ev1) * 1e-3) / 1e12)
        out[name] = best
    return out


def hbm_peak():
    try:
        return float(json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
    except Exception:
        return 6650.0, "fallback of B200_PROFILING.md (MEASURED_PEAKS.json absent)"


class _Raw:      # zero-copy torch view of a raw device pointer
    def __init__(self, ptr, rows):
        self.__cuda_array_interface__ = {"shape": (rows, 12), "typestr": "<f8", "data": (int(ptr), False), "version": 2}


def setup_fused(ctx, eng, w, B, mode, step_len, flags, jv0, tip0):
    """Every rank's gathered buffer mapped into every process; one fused step checked against NCCL."""
    import ctrfk
    torch, dist = ctx.torch, ctx.dist
    world, rank = ctx.world, ctx.rank
    if w.f32 or w.backbone or w.shoot or w.strict:
        raise RuntimeError("--allgather-fused applies to the FP64 tip solve")
    nbytes = world * B * 96
    kind = w.collective
    hdl = buf = None
    if kind in ("symm", "multicast", "auto"):
        import torch.distributed._symmetric_memory as symm_mem
        buf = symm_mem.empty((world * B, 12), dtype=torch.float64, device=torch.device("cuda", ctx.local))
        hdl = symm_mem.rendezvous(buf, dist.group.WORLD)
        has_mc = bool(getattr(hdl, "has_multicast_support", False)) and bool(getattr(hdl, "multicast_ptr", 0))
        if kind == "auto":
            kind = "multicast" if has_mc else "ipc"
        elif kind == "multicast" and not has_mc:
            raise RuntimeError("no NVSwitch multicast support reported by torch symmetric memory")
    fused = {"kind": kind, "cleanup": lambda: None}
    if kind == "ipc":
        del hdl, buf
        mine, handle = eng.peer_alloc(nbytes)
        handles = [None] * world
        dist.all_gather_object(handles, handle)
        ptrs = [mine if r == rank else eng.peer_open(handles[r]) for r in range(world)]
        fused.update(ptrs=list(ptrs), mc=False, view=torch.as_tensor(_Raw(mine, world * B), device="cuda"))

        def cleanup(ptrs=ptrs, mine=mine):
            fused.pop("view", None)
            barrier(ctx)
            for r, p in enumerate(ptrs):
                if r != rank:
                    eng.peer_close(p)
            barrier(ctx)
            eng.peer_free(mine)
        fused["cleanup"] = cleanup
    elif kind == "multicast":
        fused.update(ptrs=[int(hdl.multicast_ptr)], mc=True, view=buf, keep=(buf, hdl))
    else:
        fused.update(ptrs=[int(p) for p in hdl.buffer_ptrs], mc=False, view=buf, keep=(buf, hdl))
    if not fused["mc"] and w.gather_replicate > 1:
        fused["ptrs"] = (fused["ptrs"] * w.gather_replicate)[:8]
    fused["flags"] = flags | (ctrfk.FLAG_GATHER_ROWWISE if w.gather_rowwise else 0)
    fused["flag"] = torch.zeros(1, dtype=torch.int32, device="cuda")
    # check once against NCCL: one fused step, a rank barrier, then the same poses gathered by all_gather_into_tensor
    eng.batch_gather(jv0, B, mode, fused["ptrs"], rank * B, step=step_len, nsamples=w.nsamples, flags=fused["flags"],
                     multicast=fused["mc"], stream=ctx.stream)
    dist.all_reduce(fused["flag"])
    eng.batch(jv0, B, mode, step=step_len, nsamples=w.nsamples, flags=flags, tip=tip0, stream=ctx.stream)
    want = torch.empty(world * B, 12, dtype=torch.float64, device="cuda")
    dist.all_gather_into_tensor(want, tip0)
    torch.cuda.synchronize()
    if not torch.equal(fused["view"], want):
        raise RuntimeError("fused all-gather differs from NCCL all_gather_into_tensor")
    del want
    return fused


def run_workload(ctx, w, steps, warmup, keep=False):
    """Warm-up, then `steps` timed passes of one workload; returns the measurement (rank-independent keys) and,
    with keep=True, the live engine and buffers under "_live" (the e2e leg of the headline reuses them)."""
    import numpy as np
    import ctrfk
    torch, dist = ctx.torch, ctx.dist
    rank, world = ctx.rank, ctx.world
    stream = ctx.stream

    desc = ctrfk.desc_from_xml(ctrfk.resource(w.robot), w.max_samples)
    eng = ctrfk.Engine(desc, ctx.local)
    B, njv = w.batch, desc.njoints
    seed = SEEDS.get(w.robot, 20170)
    mode = ctrfk.MODE_NSAMPLE if w.mode == "F" else ctrfk.MODE_STEP
    step_len = desc.max_length / w.nsamples
    flags = ctrfk.FLAG_STRICT if w.strict else 0
    if w.backbone and w.soa:
        flags |= ctrfk.FLAG_SOA

    # alternating input/output sets, so consecutive steps never touch the same lines (each set of the headline
    # workload is 144 MB, above the 126 MB L2 on its own; one set is enough where a single set exceeds L2 many times)
    NSET = w.nset
    jvs, tips, bbs, rds = [], [], [], []
    for s in range(NSET):
        jv = torch.empty(B, njv, dtype=torch.float64, device="cuda")
        eng.sample_joints(seed + 1000 * s, (rank * NSET + s) * B, B, jv, stream)   # rank's own shard
        jvs.append(jv)
        tips.append(torch.empty(B, 12, dtype=torch.float64, device="cuda"))
    if w.backbone:
        for s in range(NSET):
            bbs.append(torch.empty(B, w.max_samples, 12, dtype=torch.float64, device="cuda"))
            rds.append(torch.empty(B, w.max_samples, dtype=torch.float64, device="cuda"))
    nccl_gather = w.collective == "nccl" and world > 1
    gathered = torch.empty(world * B, 12, dtype=torch.float64, device="cuda") if nccl_gather else None
    fused = None
    if w.collective in ("ipc", "symm", "multicast", "auto") and world > 1:
        fused = setup_fused(ctx, eng, w, B, mode, step_len, flags, jvs[0], tips[0])

    shoot = None
    if w.shoot:
        # inputs: base angles reached from the random tip angles (so a root exists for every input)
        import util
        idx = util.alpha_indices(desc)
        shoot = []
        for s in range(NSET):
            ab = torch.empty(B, desc.ntubes, dtype=torch.float64, device="cuda")
            eng.batch(jvs[s], B, mode, step=step_len, nsamples=w.nsamples, alpha_base=ab, stream=stream)
            jb = jvs[s].clone()
            for t in range(1, desc.ntubes):
                jb[:, idx[t]] = ab[:, t]
            guess = None
            if w.shoot_guess is not None:
                guess = jvs[s].clone()
                g = torch.Generator(device="cuda"); g.manual_seed(1234 + s)
                for t in range(1, desc.ntubes):
                    guess[:, idx[t]] += (torch.rand(B, dtype=torch.float64, device="cuda", generator=g) * 2 - 1) * w.shoot_guess
            shoot.append(dict(jvb=jb, jt=torch.empty_like(jb), it=torch.empty(B, dtype=torch.int32, device="cuda"),
                              st=torch.empty(B, dtype=torch.int32, device="cuda"), guess=guess))

    def one_step(i):
        s = i % NSET
        if w.shoot:
            eng.batch_base(shoot[s]["jvb"], B, mode, step=step_len, nsamples=w.nsamples, tol=1e-11, max_iter=w.shoot_max_iter,
                           jv_tip=shoot[s]["jt"], tip=tips[s], iters=shoot[s]["it"], status=shoot[s]["st"], stream=stream,
                           jv_guess=shoot[s]["guess"], flags=ctrfk.FLAG_RK4 if w.shoot_rk4 else 0)
        elif w.f32:
            eng.batch_f32(jvs[s], B, mode, step=step_len, nsamples=w.nsamples, flags=flags, tip=tips[s], stream=stream)
        elif fused is not None:
            # one kernel solves and stores every pose into all ranks' gathered buffers; the one-element all-reduce
            # is the rank barrier after which every rank's buffer is complete (what an all-gather guarantees)
            eng.batch_gather(jvs[s], B, mode, fused["ptrs"], rank * B, step=step_len, nsamples=w.nsamples, flags=fused["flags"],
                             multicast=fused["mc"], stream=stream)
            dist.all_reduce(fused["flag"])
        else:
            eng.batch(jvs[s], B, mode, step=step_len, nsamples=w.nsamples, flags=flags, tip=tips[s],
                      backbone=bbs[s] if w.backbone else None, radius=rds[s] if w.backbone else None, stream=stream)
        if gathered is not None:
            dist.all_gather_into_tensor(gathered, tips[s])

    # ---- warm-up, then the timed region ----------------------------------------------------------
    clocks = ClockSampler(visible_index(ctx.local))
    clocks.start()
    for i in range(max(warmup, 3)):
        one_step(i)
    barrier(ctx)
    launches0 = ctrfk.launch_count()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    clocks.mark()
    evs[0].record()
    for i in range(steps):
        one_step(i)
        evs[i + 1].record()
    barrier(ctx)
    clk = clocks.stop()
    launches = ctrfk.launch_count() - launches0
    elapsed_ms = evs[0].elapsed_time(evs[-1])
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms_max = float(t.item())
    value = world * B * steps / (elapsed_ms_max * 1e-3)

    # ---- roofline of the dominant kernel (the solver launch of each step) ------------------------
    per_step_ms = [evs[i].elapsed_time(evs[i + 1]) for i in range(steps)]
    kern_ms = float(np.mean(per_step_ms))
    fstep = ctrfk.alg_flops_per_step(desc, backbone=w.backbone)
    if w.mode == "F":
        nsum = float(min(w.nsamples, desc.max_samples)) * B
    else:
        ncount = ctrfk.count_samples_host(desc, jvs[0].cpu().numpy(), mode, step=step_len)
        nsum = float(ncount.astype(np.float64).sum())
    alg_flops = nsum * fstep + ctrfk.ALG_FLOPS_ONCE * B
    achieved = alg_flops / (kern_ms * 1e-3) / 1e12
    traffic = None
    prof_json = os.path.join(REPO, "profiles", "dram_traffic.json")
    if os.path.exists(prof_json):
        try:
            ent = json.load(open(prof_json)).get((("backbone_soa" if w.soa else "backbone") if w.backbone else "tip") + "_" + w.mode)
            if ent and ent["robot"] == w.robot and ent["batch"] == B and not (w.strict or w.f32 or w.shoot):
                traffic = ent["bytes"]
        except Exception:
            traffic = None
    peak64 = ctx.peaks["fp64"]
    if w.backbone:
        # config 4: 104 B of frames + radius per sample — HBM stores and the FP64 pipe both matter
        hbm, hbm_src = hbm_peak()
        alg_bytes = B * (njv * 8 + 96) + (12 + 1) * 8 * nsum
        gbs = alg_bytes / (kern_ms * 1e-3) / 1e9
        roofline = {"bound": "hbm", "achieved": gbs, "peak": hbm, "unit": "GB/s", "frac": gbs / hbm,
                    "traffic": traffic, "kernel_ms": kern_ms, "alg_bytes_per_launch": alg_bytes,
                    "peak_source": hbm_src, "mean_samples": nsum / B,
                    "fp64": {"achieved_tflops": achieved, "peak_tflops": peak64, "frac": achieved / peak64 if peak64 > 0 else None,
                             "alg_flops_per_sample_step": fstep}}
    elif w.f32:
        peak32 = ctx.peaks["fp32x2"]
        roofline = {"bound": "fp32", "achieved": achieved, "peak": peak32, "unit": "TFLOP/s",
                    "frac": achieved / peak32 if peak32 > 0 else None, "traffic": traffic, "kernel_ms": kern_ms,
                    "alg_flops_per_launch": alg_flops, "alg_bytes_per_launch": B * (njv * 8 + 96),
                    "peak_source": "measured live: independent packed-FFMA2 chains on all SMs (ctr_fk_fp32_probe; "
                                   "fma.rn.f32x2 is what the FP32 solver issues); nominal 148 SM x 128 lanes x 2 x 1.965 GHz = 74.4",
                    "alg_flops_per_sample_step": fstep, "mean_samples": nsum / B}
    else:
        roofline = {"bound": "fp64", "achieved": achieved, "peak": peak64, "unit": "TFLOP/s",
                    "frac": achieved / peak64 if peak64 > 0 else None, "traffic": traffic,
                    "kernel_ms": kern_ms,
                    "alg_flops_per_launch": alg_flops,
                    "alg_bytes_per_launch": B * (njv * 8 + 96),
                    "peak_source": "measured live: independent-DFMA-chain probe on all SMs (ctr_fk_fp64_probe); "
                                   "MEASURED_PEAKS.json has no FP64 figure; nominal 148 SM x 64 DFMA/clk x 2 x 1.965 GHz = 37.2",
                    "alg_flops_per_sample_step": fstep, "mean_samples": nsum / B,
                    # why this is not an HBM-bound kernel: algorithmic bytes over the same launch time
                    "hbm": {"achieved_GBs": B * (njv * 8 + 96) / (kern_ms * 1e-3) / 1e9,
                            "note": "about 1 % of the measured copy bandwidth (MEASURED_PEAKS.json hbm_gbs)"}}

    if gathered is not None:
        coll = "all_gather(tip): NCCL all_gather_into_tensor after the solve"
    elif fused is not None:
        coll = ("fused into the solver kernel: %s + 1-element all-reduce as rank barrier, checked equal to "
                "all_gather_into_tensor; %d destinations%s" % (
                    "NVSwitch multicast stores (multimem.st)" if fused["mc"] else "NVLink peer stores (%s buffers, CTA-coalesced lines)" % fused["kind"],
                    len(fused["ptrs"]), ", row-wise stores" if w.gather_rowwise else ""))
    else:
        coll = "none"
    cfg = config_keys(w, coll)
    res = {"value": value, "unit": UNIT, "ms_per_step": elapsed_ms_max / steps, "steps": steps, "warmup": max(warmup, 3),
           "dtype": "f32" if w.f32 else "f64", "config": cfg, "arm": {"arithmetic": "strict" if w.strict else "fast"},
           "roofline": roofline, "clocks": clk, "gpu_launches": launches}
    if w.shoot:
        import numpy as np
        it = shoot[0]["it"].cpu().numpy(); stt = shoot[0]["st"].cpu().numpy()
        res["shoot"] = {"converged_frac": float((stt == 0).mean()), "mean_sweeps_converged": float(it[stt == 0].mean()),
                        "max_sweeps": int(it.max()), "sweeps_per_config": float(it.mean()), "sweep_cap": w.shoot_max_iter,
                        "start": "x = beta, restarts" if w.shoot_guess is None else "guess within %g rad of the root" % w.shoot_guess,
                        "integrator": "rk4 (continuum model)" if w.shoot_rk4 else "reference sweep",
                        "statuses": {int(k): int(v) for k, v in zip(*np.unique(stt, return_counts=True))}}
        cfg["workload"] += " [base-angle shooting solve, not the headline FK]"
    if keep:
        res["_live"] = dict(eng=eng, desc=desc, jvs=jvs, tips=tips, mode=mode, step_len=step_len, flags=flags, seed=seed)
    else:
        if fused is not None:
            fused["cleanup"]()
        del jvs, tips, bbs, rds, gathered, fused
        eng.close()
        torch.cuda.empty_cache()
    return res


def run_e2e(ctx, a, w, live):
    """The same metric end to end through ctr_fk_batch_host with HOST buffers, and the box's PCIe ceiling."""
    import numpy as np
    import ctrfk
    torch, dist = ctx.torch, ctx.dist
    world = ctx.world
    eng, desc = live["eng"], live["desc"]
    B, njv = w.batch, desc.njoints
    mode, step_len = live["mode"], live["step_len"]
    flags = live["flags"] | (ctrfk.FLAG_HOST_STAGED if a.e2e_staged else 0) | (ctrfk.FLAG_HOST_ZEROCOPY if a.e2e_zerocopy else 0)

    # this process on the CPUs next to its GPU before any pinned page is touched (multi-socket hosts)
    bound = None
    if not a.no_bind:
        try:
            local_cpus = parse_cpulist(eng.host_local_cpus())
            cur = os.sched_getaffinity(0)
            both = cur & local_cpus
            if both and both != cur:
                os.sched_setaffinity(0, both)
                bound = (cur, sorted(both))
        except Exception:
            bound = None

    # (1) the ceiling: every rank copies one step's bytes both ways at once, all ranks together
    h_jv_t = torch.empty(B, njv, dtype=torch.float64).pin_memory()
    h_tip_t = torch.empty(B, 12, dtype=torch.float64).pin_memory()
    h_jv_t.copy_(live["jvs"][0].cpu())
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
    d_jv, d_tip = live["jvs"][1], live["tips"][1]
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    single = {}
    for name, dst, src in (("h2d_GBs", d_jv, h_jv_t), ("d2h_GBs", h_tip_t, d_tip)):     # one direction, this rank alone is not isolated at N>1
        best = 0.0
        for _ in range(3):
            ev0.record()
            dst.copy_(src, non_blocking=True)
            ev1.record()
            torch.cuda.synchronize()
            best = max(best, src.numel() * 8 / (ev0.elapsed_time(ev1) * 1e-3) / 1e9)
        single[name] = best
    reps = 6

    def both_ways():
        with torch.cuda.stream(s_in):
            d_jv.copy_(h_jv_t, non_blocking=True)
        with torch.cuda.stream(s_out):
            h_tip_t.copy_(d_tip, non_blocking=True)
    def h2d_only():
        with torch.cuda.stream(s_in):
            d_jv.copy_(h_jv_t, non_blocking=True)

    def d2h_only():
        with torch.cuda.stream(s_out):
            h_tip_t.copy_(d_tip, non_blocking=True)

    def all_ranks_rate(fn, nbytes):
        """aggregate GB/s of `fn` run `reps` times by every rank at once (total bytes / slowest rank's time)"""
        fn()
        barrier(ctx)
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return world * nbytes * reps / float(tt.item()) / 1e9
    step_bytes = B * (njv * 8 + 96)
    agg_gbs = all_ranks_rate(both_ways, step_bytes)
    agg_h2d = all_ranks_rate(h2d_only, B * njv * 8)
    agg_d2h = all_ranks_rate(d2h_only, B * 96)
    ceiling = agg_gbs * 1e9 / (njv * 8 + 96)

    # (2) the call itself, host buffers placed by ctr_fk_host_alloc (page-locked, mapped, on the GPU's NUMA node)
    h_jv, p_jv = eng.host_array((B, njv), write_combined=a.e2e_wc)
    h_tip, p_tip = eng.host_array((B, 12))
    h_jv[...] = h_jv_t.numpy()
    e2e_steps = max(3, min(a.steps, 10))
    for _ in range(2):
        eng.batch_host(h_jv, B, mode, step=step_len, nsamples=w.nsamples, flags=flags, tip=h_tip)
    barrier(ctx)
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        eng.batch_host(h_jv, B, mode, step=step_len, nsamples=w.nsamples, flags=flags, tip=h_tip)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    e2e_value = world * B * e2e_steps / float(tt.item())
    # the poses that came back are the poses of the device-resident path
    eng.batch(live["jvs"][0], B, mode, step=step_len, nsamples=w.nsamples, flags=live["flags"], tip=live["tips"][0], stream=ctx.stream)
    torch.cuda.synchronize()
    same = bool(np.array_equal(h_tip, live["tips"][0].cpu().numpy()))
    assert np.isfinite(h_tip).all() and same, "end-to-end poses differ from the device-resident path"
    zero_copy = a.e2e_zerocopy and not (a.e2e_staged or w.strict or (w.mode == "S" and B >= 8192))
    e2e = {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": B * njv * 8, "d2h_bytes_per_step": B * 12 * 8,
           "steps": e2e_steps,
           "api": "ctr_fk_batch_host (host pointers in, tip poses out; " +
                  ("zero-copy form: ONE kernel reads the joint vectors from and writes the poses to page-locked host memory over PCIe"
                   if zero_copy else "staged form: wave-sized chunks copied in, solved and copied out on 4 streams (both copy engines and the SMs busy at once)") + ")",
           "host_buffers": "ctr_fk_host_alloc: page-locked, device-mapped, first-touched on the CPUs local to the GPU" +
                           (", joint vectors write-combined" if a.e2e_wc else ""),
           "cpu_binding": ("process bound to GPU-local CPUs %s" % bound[1]) if bound else "none needed (the GPU's local CPU list covers the process's CPUs, or is unknown)",
           "pcie_measured": single,
           "pcie_aggregate": {"GBs": agg_gbs, "h2d_only_GBs": agg_h2d, "d2h_only_GBs": agg_d2h, "ranks": world, "bytes_per_rank_per_rep": step_bytes, "reps": reps,
                              "how": "after a barrier every rank copies one step's bytes both ways at once (H2D %d MB and D2H %d MB on two "
                                     "streams, pinned memory), %d times; total bytes of all ranks / slowest rank's time" % (
                                         B * njv * 8 // 1000000, B * 96 // 1000000, reps),
                              "ceiling_configurations_per_s": ceiling},
           "frac_of_pcie_ceiling": e2e_value / ceiling if ceiling > 0 else None,
           "frac_of_device_value": None}
    eng.host_free(p_jv)
    eng.host_free(p_tip)
    if bound:
        os.sched_setaffinity(0, bound[0])
    return e2e


def extra_workloads(a, world):
    """BASELINE.json configs[2..4] (per-GPU shard sizes of the 8-GPU statement; the same per-GPU work at every N)."""
    base = dict(mode="F", nsamples=256, max_samples=256, strict=False, soa=False, shoot=False, shoot_max_iter=100, shoot_guess=None, shoot_rk4=False, gather_rowwise=False,
                gather_replicate=1, backbone=False, f32=False, collective="none", nset=2)
    out = []
    c3 = dict(base, robot="ctr_sec3_tub3.xml", batch=2_000_000)
    out.append(("config 3: ctr_sec3_tub3, 16M configurations over 8 GPUs = 2M per GPU", dict(c3)))
    if world > 1:
        out.append(("config 3 + NCCL all-gather of the poses", dict(c3, collective="nccl")))
        out.append(("config 3 + all-gather fused into the solver: NVLink peer stores", dict(c3, collective="ipc")))
        out.append(("config 3 + all-gather fused into the solver: NVSwitch multicast (what `auto` picks when the group has it)",
                    dict(c3, collective="multicast")))
    out.append(("config 4: ctr_sec3_tub4, backbone + radius of every sample, 1M per GPU",
                dict(base, robot="ctr_sec3_tub4.xml", batch=1_000_000, backbone=True, nset=1)))
    c5 = dict(base, robot="ctr_sec3_tub5.xml", batch=12_500_000, nset=1)
    out.append(("config 5: ctr_sec3_tub5, 100M configurations over 8 GPUs = 12.5M per GPU, FP64", dict(c5)))
    out.append(("config 5, FP32 variant (packed FFMA2 solver)", dict(c5, f32=True)))
    return out


def main():
    a = parse()
    if a.impl == "reference":
        return run_reference(a)

    import torch
    import torch.distributed as dist
    import ctrfk

    ctx = Ctx()
    ctx.torch, ctx.dist = torch, dist
    ctx.rank = int(os.environ.get("RANK", "0"))
    ctx.world = int(os.environ.get("WORLD_SIZE", "1"))
    ctx.local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(ctx.local)
    if ctx.world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", ctx.local))
    assert ctx.world == a.gpus or ctx.world == 1, "launch with torchrun --nproc-per-node %d" % a.gpus
    ctx.stream = torch.cuda.current_stream().cuda_stream
    ctx.peaks = measure_peaks(ctx)

    w = workload_of(a)
    res = run_workload(ctx, w, a.steps, a.warmup, keep=True)
    live = res.pop("_live")

    e2e = None
    if not (w.backbone or w.f32 or w.shoot or a.no_e2e):
        e2e = run_e2e(ctx, a, w, live)
        e2e["frac_of_device_value"] = e2e["value"] / res["value"]

    # ---- CPU baseline (rank 0, N=1 only): the reference's code on a bounded sample ------------------
    cpu = None
    if ctx.rank == 0 and ctx.world == 1 and not a.no_cpu_baseline and not w.shoot:
        threads = host_threads()
        arm = CpuArm(w)
        r1, _ = arm.rate(arm.inputs(live["seed"], 0, 2048), 1)
        sample = a.cpu_sample or int(min(w.batch, max(20000, r1 * threads * 10.0)))      # ~10 s of CPU work
        rate, secs = arm.rate(arm.inputs(live["seed"], 0, sample), threads)
        cpu = {"value": rate, "unit": UNIT, "cores": threads, "kind": arm.kind,
               "sample": "%d configurations of the bench workload, %.1f s; %s" % (sample, secs, CPU_ARM_NOTE[arm.kind]),
               "single_thread_value": r1}
    live["eng"].close()
    del live
    torch.cuda.empty_cache()

    line = {"metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": ctx.world, "steps": a.steps, "warmup": res["warmup"],
            "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": res["dtype"], "data": "synthetic", "config": res["config"], "arm": res["arm"], "roofline": res["roofline"],
            "cpu_baseline": cpu, "e2e": e2e, "clocks": res["clocks"], "gpu_launches": res["gpu_launches"]}
    if "shoot" in res:
        line["shoot"] = res["shoot"]

    # ---- BASELINE configs 3, 4, 5 ----------------------------------------------------------------------
    default_headline = (w.robot == "ctr_sec2_tub3.xml" and w.mode == "F" and w.collective == "none" and
                        not (w.backbone or w.f32 or w.shoot or w.strict))
    printed = threading.Event()

    def emit():
        if ctx.rank == 0 and not printed.is_set():
            printed.set()
            print(json.dumps(line), flush=True)

    if a.configs == "auto" and default_headline:
        line["configs"] = []
        # a hang in an optional entry must not cost the headline: after the deadline rank 0 prints what it has
        deadline = threading.Timer(240.0, lambda: (line["configs"].append({"error": "configs leg exceeded 240 s; stopped"}), emit(), os._exit(0)))
        deadline.daemon = True
        deadline.start()
        for name, kw in extra_workloads(a, ctx.world):
            wk = types.SimpleNamespace(**kw)
            try:
                r = run_workload(ctx, wk, a.config_steps, 3)
                r["name"] = name
                r["n_gpus"] = ctx.world
                line["configs"].append(r)
            except (RuntimeError, ImportError, AttributeError) as e:      # e.g. no multicast support in this group
                line["configs"].append({"name": name, "unavailable": repr(e)[:300]})
                torch.cuda.empty_cache()
        deadline.cancel()
    emit()
    if ctx.world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
