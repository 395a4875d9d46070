#!/usr/bin/env python
"""BASELINE config 5 in one launch on one GPU: 100M configurations of ctr_sec3_tub5 generated on
the device, solved by one ctr_fk_batch call (FP64) and one ctr_fk_batch_f32 call, slices checked
against the CPU oracle (64-bit indexing, grid sizes beyond 2^16 CTAs).
   python tests/tools/big_batch_check.py [B]"""
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
for p in ("ram2017-ctr-kinematics_b200", "oracle", "tests"):
    sys.path.insert(0, os.path.join(HERE, "..", "..", p))
import ctrfk  # noqa: E402
import oracle_lib  # noqa: E402
import util  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
    ctrfk.load(); oracle_lib.load()
    d = ctrfk.desc_from_xml(ctrfk.resource("ctr_sec3_tub5.xml"), 256)
    eng = ctrfk.Engine(d, 0)
    st = torch.cuda.current_stream().cuda_stream
    jv = torch.empty((B, eng.njv), dtype=torch.float64, device="cuda")
    tip = torch.empty((B, 12), dtype=torch.float64, device="cuda")
    eng.sample_joints(20174, 0, B, jv, stream=st)
    torch.cuda.synchronize()
    for name, call in (("fp64", lambda: eng.batch(jv, B, ctrfk.MODE_NSAMPLE, nsamples=256, tip=tip, stream=st)),
                       ("fp32", lambda: eng.batch_f32(jv, B, ctrfk.MODE_NSAMPLE, nsamples=256, tip=tip, stream=st))):
        tip.fill_(float("nan"))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        call()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        worst_p = worst_r = 0.0
        for lo in (0, B // 3, B // 2 + 12345, B - 2000):
            hi = min(B, lo + 2000)
            q = jv[lo:hi].cpu().numpy()
            o = oracle_lib.fk_batch(d, q, ctrfk.MODE_NSAMPLE, nsamples=256)
            dp, ang = util.pose_errors(o["tip"], tip[lo:hi].cpu().numpy())
            worst_p, worst_r = max(worst_p, dp.max()), max(worst_r, ang.max())
        nan_rows = int(torch.isnan(tip[:, 0]).sum().item())
        print("%s: %d configurations in %.1f ms (%.3e /s), unwritten rows %d, worst slice error %.2e m %.2e rad"
              % (name, B, dt * 1e3, B / dt, nan_rows, worst_p, worst_r))
        tol = (1e-9, 1e-9) if name == "fp64" else (1e-4, 1e-3)
        assert nan_rows == 0 and worst_p <= tol[0] and worst_r <= tol[1]
    eng.close()
    print("OK")


if __name__ == "__main__":
    main()
