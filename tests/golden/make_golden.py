"""Freezes outputs of the CPU oracle (oracle/ctr_oracle.c) as regression vectors:
tests/golden/oracle_golden.npz.

NOT reference-derived (the reference-derived vectors are tests/golden/reference_golden.npz, made by
make_reference_golden.py from the reference build oracle/_ref).  These vectors pin the oracle
against silent drift and give the GPU box fixed inputs/outputs that do not depend on the oracle being rebuilt.
Inputs: counter-based sampler (include/ctr_fk.h, ctr_fk_sample_joints_host), seeds of SURVEY §8d,
first 256 configurations per robot, MaxSamples = 256; mode F: N = 256; mode S: step = MaxLength/256.

Run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
for p in (os.path.join(REPO, "ram2017-ctr-kinematics_b200"), os.path.join(REPO, "oracle"), os.path.join(REPO, "tests")):
    sys.path.insert(0, p)
import ctrfk          # noqa: E402
import oracle_lib     # noqa: E402
import util           # noqa: E402

N_CFG = 256

if __name__ == "__main__":
    out = {}
    for name in util.ROBOT_FILES:
        d = ctrfk.desc_from_xml(ctrfk.resource(name), 256)
        jv = ctrfk.sample_joints_host(d, util.SEEDS[name], 0, N_CFG)
        key = name.replace(".xml", "")
        out[key + "/jv"] = jv
        for tag, mode in (("F", ctrfk.MODE_NSAMPLE), ("S", ctrfk.MODE_STEP)):
            r = oracle_lib.fk_batch(d, jv, mode, step=d.max_length / 256.0, nsamples=256, nthreads=1)
            out["%s/%s/tip" % (key, tag)] = r["tip"]
            out["%s/%s/n" % (key, tag)] = r["n"]
            out["%s/%s/step" % (key, tag)] = r["step"]
            out["%s/%s/alpha_base" % (key, tag)] = r["alpha_base"]
    np.savez_compressed(os.path.join(HERE, "oracle_golden.npz"), **out)
    print("wrote oracle_golden.npz with", len(out), "arrays")
