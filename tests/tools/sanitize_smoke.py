#!/usr/bin/env python
"""Small-batch pass over every kernel of libctrfk.so, meant to run under
`compute-sanitizer --tool memcheck` (one tool per gpurun call, see B200_PROFILING.md)."""
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in ("ram2017-ctr-kinematics_b200", "oracle", "tests"):
    sys.path.insert(0, os.path.join(REPO, p))
import ctrfk      # noqa: E402
import util       # noqa: E402

torch.cuda.set_device(0)
st = torch.cuda.current_stream().cuda_stream
for name in ("ctr_sec2_tub3.xml", "ctr_sec3_tub5.xml"):
    d = ctrfk.desc_from_xml(ctrfk.resource(name), 64)
    eng = ctrfk.Engine(d, 0)
    B = 9001                                   # ragged vs 128-thread CTAs, above the binning threshold
    jv = torch.empty(B, d.njoints, dtype=torch.float64, device="cuda")
    eng.sample_joints(1, 0, B, jv, st)
    eng.sample_joints(1, 0, B, jv, st, policy=ctrfk.SCALE_PROPORTIONAL, p0=0.2, p1=0.8)
    tip = torch.empty(B, 12, dtype=torch.float64, device="cuda")
    n = torch.empty(B, dtype=torch.int32, device="cuda")
    h = torch.empty(B, dtype=torch.float64, device="cuda")
    ab = torch.empty(B, d.ntubes, dtype=torch.float64, device="cuda")
    stt = torch.empty(B, dtype=torch.int32, device="cuda")
    bb = torch.empty(B, 64, 12, dtype=torch.float64, device="cuda")
    rd = torch.empty(B, 64, dtype=torch.float64, device="cuda")
    bbs = torch.empty(64, 12, B, dtype=torch.float64, device="cuda")
    rds = torch.empty(64, B, dtype=torch.float64, device="cuda")
    for mode, kw in ((1, dict(nsamples=64)), (0, dict(step=d.max_length / 64)), (1, dict(nsamples=3))):
        for flags in (0, ctrfk.FLAG_STRICT):
            eng.batch(jv, B, mode, flags=flags, tip=tip, n_out=n, step_out=h, alpha_base=ab, status=stt, stream=st, **kw)
            eng.batch(jv, B, mode, flags=flags, tip=tip, backbone=bb, radius=rd, n_out=n, stream=st, **kw)
            eng.batch(jv, B, mode, flags=flags | ctrfk.FLAG_SOA, tip=tip, backbone=bbs, radius=rds, stream=st, **kw)
        eng.batch_f32(jv, B, mode, tip=tip, n_out=n, step_out=h, status=stt, stream=st, **kw)
        jt = torch.empty_like(jv)
        it = torch.empty(B, dtype=torch.int32, device="cuda")
        eng.batch_base(jv, B, mode, tol=1e-10, max_iter=8, jv_tip=jt, tip=tip, iters=it, status=stt, stream=st, **kw)
    jac = torch.empty(B, d.njoints, 6, dtype=torch.float64, device="cuda")
    eng.jacobian(jv, B, 32, 1e-5, 1e-4, jac, tip, st)
    deg = torch.empty(B, dtype=torch.float64, device="cuda")
    eng.stability(jv[:500], 500, 0.3, d.max_length / 64, -1.0, stt, deg, st)
    hj = jv.cpu().numpy()
    ht = np.zeros((B, 12))
    eng.batch_host(hj, B, 1, nsamples=64, tip=ht)
    t1 = np.zeros(12); b1 = np.zeros((64, 12)); r1 = np.zeros(64)
    eng.single(hj[0], 0, step=d.max_length / 64, tip=t1, backbone=b1, radius=r1)
    torch.cuda.synchronize()
    assert np.isfinite(ht).all() and np.isfinite(t1).all()
    eng.close()
ctrfk.fp64_probe(0, 10, st)
torch.cuda.synchronize()
print("sanitize_smoke OK, launches:", ctrfk.launch_count())
