#!/usr/bin/env python
"""bench.py — FK solves/sec (FP64) of the batched concentric-tube-robot forward kinematics.

  python bench.py --gpus N --steps K --warmup W          our CUDA path (one process per GPU)
  python bench.py --impl reference ...                   the reference algorithm on the host CPU

A "step" is one pass of the hot path (ctr_fk_batch) over one batch of synthetic joint vectors.
Workload at N=1 = BASELINE.json configs[1]: ctr_sec2_tub3.xml, 1M random joint configurations,
tip pose out, fixed N=256 samples (SURVEY §8d "Mode F", MaxSamples=256).  Multi-GPU: the batch
shards trivially, each rank solves its own 1M configurations (weak scaling), no data-path
collective (an optional NCCL all-gather of the poses can be timed with --allgather).

Printed JSON (one line, rank 0): see the bench contract in the task description; additionally
`roofline` (FP64 FMA pipe — nothing on this path is a dense contraction and tip-only I/O is
~150 B per configuration, so neither "tensor" nor "hbm" applies), `cpu_baseline`, `e2e`,
`clocks`, `gpu_launches`.
"""
import argparse
import json
import os
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
for _p in (os.path.join(REPO, "ram2017-ctr-kinematics_b200"), os.path.join(REPO, "oracle"), os.path.join(REPO, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

METRIC = "FK solves/sec (FP64, 3-tube)"
UNIT = "configurations/s"


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
    ap.add_argument("--shoot-max-iter", type=int, default=24,
                    help="Newton iteration cap of --shoot (99.9 %% of the converging configurations need <= 18)")
    ap.add_argument("--allgather", action="store_true", help="time an NCCL all-gather of the tip poses too")
    ap.add_argument("--gather-rowwise", action="store_true", help="--allgather-fused with one thread storing its own row (comparison)")
    ap.add_argument("--gather-replicate", type=int, default=1,
                    help="--allgather-fused ipc/symm: list every destination K times (emulates K times as many peers on a small box)")
    ap.add_argument("--allgather-fused", choices=["ipc", "symm", "multicast"], default=None,
                    help="solve + all-gather in ONE kernel (ctr_fk_batch_gather): poses stored straight into every rank's "
                         "gathered buffer over NVLink; ipc = buffers shared by CUDA IPC (ctr_fk_peer_*), symm = torch "
                         "symmetric memory, multicast = its NVSwitch multicast address (one multimem.st per 16 bytes)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-pointer end-to-end leg (very large batches)")
    ap.add_argument("--cpu-sample", type=int, default=0, help="configurations of the CPU baseline sample (0 = auto)")
    return ap.parse_args()


def workload_name(a):
    return "%s, %s configurations/GPU, %s, %s out" % (
        a.robot, "{:,}".format(a.batch).replace(",", "_"),
        ("fixed N=%d samples (calcKinematic(JV, NSamples))" % a.nsamples) if a.mode == "F"
        else "step = MaxLength/%d (calcKinematic(JV, ArcLengthStep))" % a.nsamples,
        ("backbone+radius" + (" (SoA layout)" if getattr(a, "soa", False) else "")) if a.backbone else "tip pose")


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
    """All host cores this process may use.  torchrun exports OMP_NUM_THREADS=1; the oracle takes
    an explicit thread count (num_threads clause), so that default does not cap the CPU arm."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


_REF_ROBOTS = {}


def cpu_arm(a):
    """Which CPU implementation the baseline legs time: ("reference", robot) when oracle/_ref/libctr_ref.so
    exists — the reference's own sources compiled against stand-in third-party headers (oracle/Makefile
    target `ref`; prebuilt, it travels to the GPU box) — else ("port", None): the C oracle."""
    import ctrfk
    import ref_lib
    if not os.path.exists(ref_lib.LIB_PATH):
        return "port", None
    key = (a.robot, a.max_samples)
    if key not in _REF_ROBOTS:      # CTR::Robot<double>(MaxSample, XmlFilename), ctr_robot.h:105
        _REF_ROBOTS[key] = ref_lib.RefRobot.from_xml(ctrfk.resource(a.robot), a.max_samples)
    return "reference", _REF_ROBOTS[key]


def cpu_reference_rate(desc, a, nthreads, sample, seed, repeat=1):
    """Times the reference's CPU forward kinematics on `sample` configurations of the bench workload:
    CTR::Robot<double>::calcKinematic through oracle/_ref when that library exists, else the oracle port.
    One robot object per thread, OpenMP over configurations."""
    import ctrfk
    kind, ref = cpu_arm(a)
    jv = ctrfk.sample_joints_host(desc, seed, 0, sample)
    mode = ctrfk.MODE_NSAMPLE if a.mode == "F" else ctrfk.MODE_STEP
    kw = dict(nsamples=a.nsamples, step=desc.max_length / a.nsamples, nthreads=nthreads,
              backbone=a.backbone, radius=a.backbone)
    if kind == "reference":
        call = lambda: ref.fk_batch(jv, mode, **kw)
    else:
        import oracle_lib
        call = lambda: oracle_lib.fk_batch(desc, jv, mode, **kw)
    best = None
    for _ in range(repeat):
        t0 = time.perf_counter()
        call()
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return sample / best, best


CPU_ARM_NOTE = {
    "reference": "the reference's own sources (ctr_source/*.h, ctr_robot.cpp) compiled -O2 against stand-in "
                 "Eigen/Erl/Boost headers (oracle/_ref/libctr_ref.so), CTR::Robot<double>::calcKinematic, "
                 "one robot per thread, OpenMP over configurations",
    "port": "CPU oracle = step-for-step C restatement of the reference FK (oracle/ctr_oracle.c; "
            "oracle/_ref/libctr_ref.so not present), OpenMP over configurations",
}


def run_reference(a):
    """--impl reference: the reference's CPU algorithm on the box's host cores (all threads)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import ctrfk
    desc = ctrfk.desc_from_xml(ctrfk.resource(a.robot), a.max_samples)
    kind, _ = cpu_arm(a)
    threads = host_threads()
    seed = 20171
    # bounded sample per step: ~1 s of CPU work, so K+W steps end within minutes
    probe_rate, _ = cpu_reference_rate(desc, a, threads, 4096, seed)
    sample = a.cpu_sample or int(min(a.batch, max(8192, probe_rate * 1.0)))
    for _ in range(a.warmup):
        cpu_reference_rate(desc, a, threads, sample, seed)
    t0 = time.perf_counter()
    for _ in range(a.steps):
        cpu_reference_rate(desc, a, threads, sample, seed)
    dt = time.perf_counter() - t0
    value = sample * a.steps / dt
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": 1e3 * dt / a.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(a), "robot": a.robot, "mode": a.mode, "nsamples": a.nsamples,
                       "note": CPU_ARM_NOTE[kind]},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind,
                             "sample": "%d configurations per step of the bench workload" % sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


def main():
    a = parse()
    if a.impl == "reference":
        return run_reference(a)

    import numpy as np
    import torch
    import torch.distributed as dist
    import ctrfk

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    assert world == a.gpus or world == 1, "launch with torchrun --nproc-per-node %d" % a.gpus

    desc = ctrfk.desc_from_xml(ctrfk.resource(a.robot), a.max_samples)
    eng = ctrfk.Engine(desc, local)
    B, njv = a.batch, desc.njoints
    seed = {"ctr_sec2_tub3.xml": 20171, "ctr_sec3_tub3.xml": 20172, "ctr_sec3_tub4.xml": 20173,
            "ctr_sec3_tub5.xml": 20174}.get(a.robot, 20170)
    mode = ctrfk.MODE_NSAMPLE if a.mode == "F" else ctrfk.MODE_STEP
    step_len = desc.max_length / a.nsamples
    flags = ctrfk.FLAG_STRICT if a.strict else 0
    if a.backbone and a.soa:
        flags |= ctrfk.FLAG_SOA
    stream = torch.cuda.current_stream().cuda_stream

    # two input/output sets, alternated, so consecutive steps never touch the same lines
    # (each set is 144 MB for the default workload, above the 126 MB L2 on its own)
    NSET = 2
    jvs, tips, bbs, rds = [], [], [], []
    for s in range(NSET):
        jv = torch.empty(B, njv, dtype=torch.float64, device="cuda")
        eng.sample_joints(seed + 1000 * s, (rank * NSET + s) * B, B, jv, stream)   # rank's own shard
        jvs.append(jv)
        tips.append(torch.empty(B, 12, dtype=torch.float64, device="cuda"))
    if a.backbone:
        for s in range(NSET):
            bbs.append(torch.empty(B, a.max_samples, 12, dtype=torch.float64, device="cuda"))
            rds.append(torch.empty(B, a.max_samples, dtype=torch.float64, device="cuda"))
    gathered = torch.empty(world * B, 12, dtype=torch.float64, device="cuda") if (a.allgather and world > 1) else None

    # ---- fused solve + all-gather: every rank's gathered buffer mapped into every process ------------------
    fused = None
    if a.allgather_fused and world > 1:
        if a.f32 or a.backbone or a.shoot or a.strict:
            raise SystemExit("--allgather-fused applies to the FP64 tip solve")
        nbytes = world * B * 96

        class _Raw:      # zero-copy torch view of a raw device pointer
            def __init__(self, ptr, rows):
                self.__cuda_array_interface__ = {"shape": (rows, 12), "typestr": "<f8", "data": (int(ptr), False), "version": 2}

        if a.allgather_fused == "ipc":
            mine, handle = eng.peer_alloc(nbytes)
            handles = [None] * world
            dist.all_gather_object(handles, handle)
            ptrs = [mine if r == rank else eng.peer_open(handles[r]) for r in range(world)]
            fused = dict(ptrs=ptrs, mc=False, view=torch.as_tensor(_Raw(mine, world * B), device="cuda"), keep=(mine, handles))
        else:
            import torch.distributed._symmetric_memory as symm_mem
            buf = symm_mem.empty((world * B, 12), dtype=torch.float64, device=torch.device("cuda", local))
            hdl = symm_mem.rendezvous(buf, dist.group.WORLD)
            if a.allgather_fused == "multicast":
                if not hdl.has_multicast_support or not hdl.multicast_ptr:
                    raise SystemExit("no NVSwitch multicast support reported by torch symmetric memory")
                fused = dict(ptrs=[int(hdl.multicast_ptr)], mc=True, view=buf, keep=(buf, hdl))
            else:
                fused = dict(ptrs=[int(p) for p in hdl.buffer_ptrs], mc=False, view=buf, keep=(buf, hdl))
        if not fused["mc"] and a.gather_replicate > 1:
            fused["ptrs"] = (fused["ptrs"] * a.gather_replicate)[:8]
        fused["flags"] = flags | (ctrfk.FLAG_GATHER_ROWWISE if a.gather_rowwise else 0)
        fused["flag"] = torch.zeros(1, dtype=torch.int32, device="cuda")
        # check once against NCCL: one fused step, a rank barrier, then the same poses gathered by all_gather_into_tensor
        eng.batch_gather(jvs[0], B, mode, fused["ptrs"], rank * B, step=step_len, nsamples=a.nsamples, flags=fused["flags"],
                         multicast=fused["mc"], stream=stream)
        dist.all_reduce(fused["flag"])
        eng.batch(jvs[0], B, mode, step=step_len, nsamples=a.nsamples, flags=flags, tip=tips[0], stream=stream)
        want = torch.empty(world * B, 12, dtype=torch.float64, device="cuda")
        dist.all_gather_into_tensor(want, tips[0])
        torch.cuda.synchronize()
        if not torch.equal(fused["view"], want):
            raise SystemExit("fused all-gather differs from NCCL all_gather_into_tensor")
        del want

    shoot = None
    if a.shoot:
        # inputs: base angles reached from the random tip angles (so a root exists for every input)
        import util
        idx = util.alpha_indices(desc)
        shoot = []
        for s in range(NSET):
            ab = torch.empty(B, desc.ntubes, dtype=torch.float64, device="cuda")
            eng.batch(jvs[s], B, mode, step=step_len, nsamples=a.nsamples, alpha_base=ab, stream=stream)
            jb = jvs[s].clone()
            for t in range(1, desc.ntubes):
                jb[:, idx[t]] = ab[:, t]
            shoot.append(dict(jvb=jb, jt=torch.empty_like(jb), it=torch.empty(B, dtype=torch.int32, device="cuda"),
                              st=torch.empty(B, dtype=torch.int32, device="cuda")))

    def one_step(i):
        s = i % NSET
        if a.shoot:
            eng.batch_base(shoot[s]["jvb"], B, mode, step=step_len, nsamples=a.nsamples, tol=1e-11, max_iter=a.shoot_max_iter,
                           jv_tip=shoot[s]["jt"], tip=tips[s], iters=shoot[s]["it"], status=shoot[s]["st"], stream=stream)
        elif a.f32:
            eng.batch_f32(jvs[s], B, mode, step=step_len, nsamples=a.nsamples, flags=flags, tip=tips[s], stream=stream)
        elif fused is not None:
            # one kernel solves and stores every pose into all ranks' gathered buffers; the one-element all-reduce
            # is the rank barrier after which every rank's buffer is complete (what an all-gather guarantees)
            eng.batch_gather(jvs[s], B, mode, fused["ptrs"], rank * B, step=step_len, nsamples=a.nsamples, flags=fused["flags"],
                             multicast=fused["mc"], stream=stream)
            dist.all_reduce(fused["flag"])
        else:
            eng.batch(jvs[s], B, mode, step=step_len, nsamples=a.nsamples, flags=flags, tip=tips[s],
                      backbone=bbs[s] if a.backbone else None, radius=rds[s] if a.backbone else None, stream=stream)
        if gathered is not None:
            dist.all_gather_into_tensor(gathered, tips[s])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- FP64 peak probe (MEASURED_PEAKS.json carries no FP64 figure) --------------
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ctrfk.fp64_probe(local, 200, stream)
    torch.cuda.synchronize()
    peak_tflops = 0.0
    for _ in range(3):
        ev0.record()
        flops = ctrfk.fp64_probe(local, 4000, stream)
        ev1.record()
        torch.cuda.synchronize()
        peak_tflops = max(peak_tflops, flops / (ev0.elapsed_time(ev1) * 1e-3) / 1e12)

    # ---- warm-up, then the timed region ----------------------------------------------------------
    clocks = ClockSampler(visible_index(local))
    clocks.start()
    for i in range(max(a.warmup, 3)):
        one_step(i)
    barrier()
    launches0 = ctrfk.launch_count()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(a.steps + 1)]
    clocks.mark()
    evs[0].record()
    for i in range(a.steps):
        one_step(i)
        evs[i + 1].record()
    barrier()
    clk = clocks.stop()
    launches = ctrfk.launch_count() - launches0
    elapsed_ms = evs[0].elapsed_time(evs[-1])
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms_max = float(t.item())
    value = world * B * a.steps / (elapsed_ms_max * 1e-3)

    # ---- roofline of the dominant kernel (the solver launch of each step) ------------------------
    per_step_ms = [evs[i].elapsed_time(evs[i + 1]) for i in range(a.steps)]
    kern_ms = float(np.mean(per_step_ms))
    fstep = ctrfk.alg_flops_per_step(desc, backbone=a.backbone)
    if a.mode == "F":
        nsum = float(min(a.nsamples, desc.max_samples)) * B
    else:
        ncount = ctrfk.count_samples_host(desc, jvs[0].cpu().numpy(), mode, step=step_len)
        nsum = float(ncount.astype(np.float64).sum())
    alg_flops = nsum * fstep + ctrfk.ALG_FLOPS_ONCE * B
    achieved = alg_flops / (kern_ms * 1e-3) / 1e12
    traffic = None
    prof_json = os.path.join(REPO, "profiles", "dram_traffic.json")
    if os.path.exists(prof_json):
        try:
            ent = json.load(open(prof_json)).get((("backbone_soa" if a.soa else "backbone") if a.backbone else "tip") + "_" + a.mode)
            if ent and ent["robot"] == a.robot and ent["batch"] == B and not (a.strict or a.f32 or a.shoot):
                traffic = ent["bytes"]
        except Exception:
            traffic = None
    if a.backbone:
        # config 4: 104 B of frames + radius per sample — HBM-store bound
        hbm_peak, hbm_src = 6650.0, "fallback of B200_PROFILING.md (MEASURED_PEAKS.json absent)"
        try:
            hbm_peak = float(json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))["hbm_gbs"])
            hbm_src = "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
        except Exception:
            pass
        alg_bytes = B * (njv * 8 + 96) + (12 + 1) * 8 * nsum
        gbs = alg_bytes / (kern_ms * 1e-3) / 1e9
        roofline = {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak,
                    "traffic": traffic, "kernel_ms": kern_ms, "alg_bytes_per_launch": alg_bytes,
                    "peak_source": hbm_src, "mean_samples": nsum / B,
                    "fp64": {"achieved_tflops": achieved, "peak_tflops": peak_tflops}}
    else:
        roofline = {"bound": "fp64", "achieved": achieved, "peak": peak_tflops, "unit": "TFLOP/s",
                    "frac": achieved / peak_tflops if peak_tflops > 0 else None, "traffic": traffic,
                    "kernel_ms": kern_ms,
                    "alg_flops_per_launch": alg_flops,
                    "alg_bytes_per_launch": B * (njv * 8 + 96),
                    "peak_source": "measured live: independent-DFMA-chain probe on all SMs (ctr_fk_fp64_probe); "
                                   "MEASURED_PEAKS.json has no FP64 figure; nominal 148 SM x 64 DFMA/clk x 2 x 1.965 GHz = 37.2",
                    "alg_flops_per_sample_step": fstep, "mean_samples": nsum / B,
                    # why this is not an HBM-bound kernel: algorithmic bytes over the same launch time
                    "hbm": {"achieved_GBs": B * (njv * 8 + 96) / (kern_ms * 1e-3) / 1e9,
                            "note": "about 1 % of the measured copy bandwidth (MEASURED_PEAKS.json hbm_gbs)"}}

    if a.f32:
        # the FP32 variant runs on the FP32 FMA pipe; no FP32 peak is measured here, so no fraction
        roofline.update({"bound": "fp32", "peak": None, "frac": None,
                         "peak_source": "not measured (FP32 variant is not the headline); achieved counts the "
                                        "reference's algorithmic flops"})

    # ---- end to end through the host-pointer C-ABI call (pinned host buffers) --------------------
    do_e2e = not (a.backbone or a.f32 or a.shoot or a.no_e2e)
    if do_e2e:
        h_jv = torch.empty(B, njv, dtype=torch.float64).pin_memory()
        h_tip = torch.empty(B, 12, dtype=torch.float64).pin_memory()
        h_jv.copy_(jvs[0].cpu())
    # PCIe copy rates with the same pinned buffers: the bound of the end-to-end number
    pcie = {}
    for name, dst, src in ((("h2d_GBs", jvs[1], h_jv), ("d2h_GBs", h_tip, tips[1])) if do_e2e else ()):
        best = 0.0
        for _ in range(3):
            ev0.record()
            dst.copy_(src, non_blocking=True)
            ev1.record()
            torch.cuda.synchronize()
            best = max(best, src.numel() * 8 / (ev0.elapsed_time(ev1) * 1e-3) / 1e9)
        pcie[name] = best
    e2e = None
    if do_e2e:
        e2e_steps = max(3, min(a.steps, 10))
        for _ in range(2):
            eng.batch_host(h_jv, B, mode, step=step_len, nsamples=a.nsamples, flags=flags, tip=h_tip)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            eng.batch_host(h_jv, B, mode, step=step_len, nsamples=a.nsamples, flags=flags, tip=h_tip)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e = {"value": world * B * e2e_steps / float(tt.item()), "unit": UNIT,
               "h2d_bytes_per_step": B * njv * 8, "d2h_bytes_per_step": B * 12 * 8, "steps": e2e_steps,
               "api": "ctr_fk_batch_host (host pointers in, tip poses out; chunked H2D/solve/D2H on 4 streams)",
               "pcie_measured": pcie,
               "bound": "PCIe: %.0f MB out per step at the measured D2H rate caps this at %.3g configurations/s" % (
                   B * 96 / 1e6, B / (B * 96 / 1e9 / pcie["d2h_GBs"]))}
        assert bool(torch.isfinite(h_tip).all())

    # ---- CPU baseline (rank 0, N=1 only): the oracle on a bounded sample ---------------------------
    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline and not a.shoot:
        threads = host_threads()
        kind, _ = cpu_arm(a)
        r1, _ = cpu_reference_rate(desc, a, 1, 2048, seed)
        sample = a.cpu_sample or int(min(B, max(20000, r1 * threads * 10.0)))      # ~10 s of CPU work
        rate, secs = cpu_reference_rate(desc, a, threads, sample, seed)
        cpu = {"value": rate, "unit": UNIT, "cores": threads, "kind": kind,
               "sample": "%d configurations of the bench workload, %.1f s; %s" % (sample, secs, CPU_ARM_NOTE[kind]),
               "single_thread_value": r1}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
                "ms_per_step": elapsed_ms_max / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32" if a.f32 else "f64", "data": "synthetic",
                "config": {"workload": workload_name(a), "robot": a.robot, "mode": a.mode, "nsamples": a.nsamples,
                           "batch_per_gpu": B, "arithmetic": "strict" if a.strict else "fast",
                           "l2": "two alternating input/output sets of %.0f MB each (> 126 MB L2 together and singly)" % ((njv * 8 + 96) * B / 1e6),
                           "collective": ("all_gather(tip)" if gathered is not None else
                                          ("fused into the solver kernel: NVLink peer stores (%s) + 1-element all-reduce as rank barrier, "
                                           "checked equal to all_gather_into_tensor; %d destinations%s" % (a.allgather_fused, len(fused["ptrs"]), ", row-wise stores" if a.gather_rowwise else "")) if fused is not None else "none")},
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "clocks": clk, "gpu_launches": launches}
        if a.shoot:
            it = shoot[0]["it"].cpu().numpy(); stt = shoot[0]["st"].cpu().numpy()
            line["shoot"] = {"converged_frac": float((stt == 0).mean()), "mean_iters": float(it[stt == 0].mean()),
                             "max_iters": int(it.max()), "sweeps_per_config": float(it.mean() + 1)}
            line["config"]["workload"] += " [base-angle shooting solve, not the headline FK]"
        print(json.dumps(line), flush=True)
    eng.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
