#!/usr/bin/env python
"""Throughput of the callers of the hot path that bench.py does not cover (Jacobian, stability).
   python tools/bench_extra.py            (one GPU)"""
import json
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("ram2017-ctr-kinematics_b200", "tests"):
    sys.path.insert(0, os.path.join(REPO, p))
import ctrfk  # noqa: E402

torch.cuda.set_device(0)
st = torch.cuda.current_stream().cuda_stream
out = {}
for name in ("ctr_sec2_tub3.xml", "ctr_sec3_tub5.xml"):
    d = ctrfk.desc_from_xml(ctrfk.resource(name), 256)
    eng = ctrfk.Engine(d, 0)
    B = 1_000_000
    jv = torch.empty(B, d.njoints, dtype=torch.float64, device="cuda")
    eng.sample_joints(1, 0, B, jv, st)
    jac = torch.empty(B, d.njoints, 6, dtype=torch.float64, device="cuda")
    tip = torch.empty(B, 12, dtype=torch.float64, device="cuda")
    stable = torch.empty(B, dtype=torch.int32, device="cuda")
    deg = torch.empty(B, dtype=torch.float64, device="cuda")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed(fn, reps=3):
        fn(); torch.cuda.synchronize()
        best = 1e30
        for _ in range(reps):
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return best

    ms = timed(lambda: eng.jacobian(jv, B, 256, 1e-5, 1e-4, jac, tip, st))
    out[name + " jacobian N=256"] = {"ms": ms, "configs_per_s": B / ms * 1e3, "fk_solves_per_config": d.ntubes + d.nsections}
    Bs = 100_000
    ms = timed(lambda: eng.stability(jv, Bs, 0.1, d.max_length / 256.0, -1.0, stable, deg, st), reps=2)
    out[name + " stability dAlpha=0.1 step=L/256"] = {"ms": ms, "configs_per_s": Bs / ms * 1e3,
                                                       "stable_frac": float((stable[:Bs] == 1).float().mean())}
    eng.close()
print(json.dumps(out, indent=1))
