#!/usr/bin/env python
"""Latency of the host-pointer entry points for small batches (BASELINE config 1 is a single
configuration): ctr_fk_single and ctr_fk_batch_host, median of 2000 calls after warm-up.
   python tools/latency_probe.py"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "ram2017-ctr-kinematics_b200"))
import ctrfk  # noqa: E402


def med_us(fn, reps=2000, warm=200):
    for _ in range(warm):
        fn()
    t = np.empty(reps)
    for i in range(reps):
        t0 = time.perf_counter_ns()
        fn()
        t[i] = time.perf_counter_ns() - t0
    return float(np.median(t)) / 1e3, float(np.percentile(t, 99)) / 1e3


def main():
    ctrfk.load()
    d = ctrfk.desc_from_xml(ctrfk.resource("ctr_sec2_tub3.xml"), 256)
    eng = ctrfk.Engine(d, 0)
    jv = ctrfk.sample_joints_host(d, 20171, 0, 65536)
    tip = np.zeros((65536, 12))
    q = jv[0].copy()
    t1 = np.zeros(12)
    m, p = med_us(lambda: eng.single(q, ctrfk.MODE_NSAMPLE, nsamples=256, tip=t1))
    print("ctr_fk_single N=256           : median %7.1f us  p99 %7.1f us" % (m, p))
    m, p = med_us(lambda: eng.single(q, ctrfk.MODE_STEP, step=d.max_length / 256.0, tip=t1))
    print("ctr_fk_single step=L/256      : median %7.1f us  p99 %7.1f us" % (m, p))
    for B in (1, 32, 1024, 8192, 65536):
        a, b = jv[:B], tip[:B]
        m, p = med_us(lambda: eng.batch_host(a, B, ctrfk.MODE_NSAMPLE, nsamples=256, tip=b), reps=500 if B > 8192 else 2000)
        print("ctr_fk_batch_host B=%-6d N=256: median %7.1f us  p99 %7.1f us  (%.3g configurations/s)" % (B, m, p, B / (m * 1e-6)))
    eng.close()


if __name__ == "__main__":
    main()
