#!/usr/bin/env python
"""One launch each of the kernels around the hot path (Jacobian tip-only and iSample forms, stability, base-angle
shooting) on ctr_sec2_tub3 — the command line profiled for profiles/r01_callers_*.md:
   python tools/prof_callers.py                       timings (CUDA events)
   ncu --set full --clock-control none --import-source on -k regex:"jacobian|stability|shoot" -o ... python tools/prof_callers.py --once
"""
import json
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("ram2017-ctr-kinematics_b200", "tests"):
    sys.path.insert(0, os.path.join(REPO, p))
import ctrfk  # noqa: E402

once = "--once" in sys.argv
torch.cuda.set_device(0)
st = torch.cuda.current_stream().cuda_stream
d = ctrfk.desc_from_xml(ctrfk.resource("ctr_sec2_tub3.xml"), 256)
eng = ctrfk.Engine(d, 0)
B = 200_000
jv = torch.empty(B, d.njoints, dtype=torch.float64, device="cuda")
eng.sample_joints(1, 0, B, jv, st)
jac = torch.empty(B, d.njoints, 6, dtype=torch.float64, device="cuda"); jacs = torch.empty_like(jac)
tip = torch.empty(B, 12, dtype=torch.float64, device="cuda"); smp = torch.empty_like(tip)
stable = torch.empty(B, dtype=torch.int32, device="cuda"); deg = torch.empty(B, dtype=torch.float64, device="cuda")
jt = torch.empty_like(jv); it = torch.empty(B, dtype=torch.int32, device="cuda"); stt = torch.empty(B, dtype=torch.int32, device="cuda")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def timed(fn):
    if once:
        fn(); torch.cuda.synchronize()
        return None
    fn(); torch.cuda.synchronize()
    best = 1e30
    for _ in range(3):
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


out = {}
runs = [
    ("jacobian tip-only N=256 (5 solves per configuration)", B, lambda: eng.jacobian(jv, B, 256, 1e-5, 1e-4, jac, tip, st)),
    ("jacobian ArcLengthStep form, step=L/256", B, lambda: eng.jacobian_ex(jv, B, 0, 1e-5, 1e-4, jac, step=d.max_length / 256.0, tip=tip, stream=st)),
    ("jacobian iSample form N=256, iSample=128", B, lambda: eng.jacobian_ex(jv, B, 1, 1e-5, 1e-4, jac, nsamples=256, i_sample=128,
                                                                            jac_sample=jacs, tip=tip, sample=smp, stream=st)),
    ("stability dAlpha=0.1 step=L/256", 50_000, lambda: eng.stability(jv, 50_000, 0.1, d.max_length / 256.0, -1.0, stable, deg, st)),
    ("base-angle shooting N=256", B, lambda: eng.batch_base(jv, B, 1, nsamples=256, max_iter=24, jv_tip=jt, tip=tip, iters=it, status=stt, stream=st)),
]
for name, n, fn in runs:
    ms = timed(fn)
    if ms is not None:
        out[name] = {"ms": ms, "configs_per_s": n / ms * 1e3}
eng.close()
print(json.dumps(out, indent=1))
