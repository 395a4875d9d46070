#!/usr/bin/env python
"""Share of sweep iterations that leave the hot form of the FAST solver, per sample count.

Builds tests/hostsim with -DCTRK_COUNT_COLD (the kernels' per-thread solver compiled for the CPU
with counters; analysis only, never part of libctrfk.so) and runs seeded random configurations in
fixed-N mode.  Prints, per N: iterations whose carried sin/cos could not be rotated
(|d_t| >= kRotMaxHi -> middle tier) and iterations that left the hot tiers altogether.
Numbers quoted in DESIGN.md §7 ("Why coarse sampling costs more").
"""
import ctypes as C, os, subprocess, sys, tempfile
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "ram2017-ctr-kinematics_b200")); sys.path.insert(0, os.path.join(REPO, "tests"))
import ctrfk, util  # noqa: E402


def main():
    robot = sys.argv[1] if len(sys.argv) > 1 else "sec2_tub3"
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
    so = os.path.join(tempfile.mkdtemp(), "libhostsim_count.so")
    subprocess.check_call(["/usr/bin/g++", "-std=c++14", "-O2", "-march=x86-64-v3", "-ffp-contract=off", "-fPIC", "-shared",
                           "-Wno-unknown-pragmas", "-DCTRK_COUNT_COLD", "-o", so,
                           os.path.join(REPO, "tests", "hostsim", "hostsim.cpp")])
    lib = C.CDLL(so)
    lib.hostsim_fk.restype = C.c_int
    lib.hostsim_fk.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_double, C.c_int, C.c_int] + [C.c_void_p] * 5
    d = ctrfk.desc_from_xml(os.path.join(REPO, "tests", "data", robot + ".xml"), 1024)
    jv = ctrfk.sample_joints_host(d, 20171, 0, B)
    cnt = (C.c_longlong * 3)()
    print("%6s %12s %10s %10s %14s" % ("N", "iterations", "no-rotate%", "cold%", "configs-hit%"))
    for n in (32, 64, 96, 128, 192, 256, 512):
        lib.hostsim_tier_counters(cnt, 1)
        hit = 0
        for i in range(0, B, 1):
            before = cnt[2]
            util.hostsim_fk(lib, d, jv[i:i + 1], ctrfk.MODE_NSAMPLE, nsamples=n, fast=1)
            lib.hostsim_tier_counters(cnt, 0)
            hit += cnt[2] > before
        tot, cold, nrot = cnt[0], cnt[1], cnt[2]
        print("%6d %12d %10.3f %10.3f %14.2f" % (n, tot, 100.0 * nrot / tot, 100.0 * cold / tot, 100.0 * hit / B))


if __name__ == "__main__":
    main()
