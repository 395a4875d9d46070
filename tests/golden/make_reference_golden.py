"""Golden vectors produced by THE REFERENCE ITSELF: tests/golden/reference_golden.npz.

Source of every number in the file: oracle/_ref/libctr_ref.so = the unmodified
/root/reference/ctr_source compiled where it lies (recipe: oracle/Makefile target `ref`) against the
stand-in third-party headers of oracle/shim, called through CTR::Robot<double>
(oracle/ref_driver.cpp).  Can only be regenerated where /root/reference exists (the build
container); the committed file travels to the GPU box, where tests/test_oracle.py::
test_reference_golden and tests/test_gpu_parity.py::test_reference_golden_gpu check the oracle and
the CUDA path against it.

Contents per robot XML (robots built by the reference's own XML reader, so the entry frame is the
reference reader's; it is stored, and fed to oracle and kernels through the descriptor):
  jv            128 joint vectors: 112 draws of the counter-based sampler (seed of SURVEY §8d) + the edge
                set of tests/util.py::edge_joint_sets (phi = 0 / L, alpha = 0 / pi / 2pi^-) minus rows where
                the reference divides 0 by 0
  mt_jv         the first 16 getRandomJointValues draws of a default std::mt19937 (ctr_test.cpp:30-32)
  init          entry frame after the reader's nearestOrthogonal, column-major 3x4
  F/*, S/*      calcKinematic(JV, N=256) and calcKinematic(JV, step=MaxLength/256): n, step, tip,
                backbone frames of 9 probe samples per configuration, radius of all samples
  jac           calcJacobian(JV, Tip, N=128, 1e-5, 1e-4) of the first 16 joint vectors
  stable, degree  calcStability / calcDegreeStability(.., 0.2, MaxLength/128, -1) of the first 32
  ctor/*        state left by the constructor's own FK call (ctr_static.h:248-252)

Run:  python tests/golden/make_reference_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
for p in (os.path.join(REPO, "ram2017-ctr-kinematics_b200"), os.path.join(REPO, "oracle"), os.path.join(REPO, "tests")):
    sys.path.insert(0, p)
import ctrfk          # noqa: E402
import ref_lib        # noqa: E402
import util           # noqa: E402

PROBES = (0, 1, 2, 17, 64, 127, 200, 254, 255)


def main():
    out = {}
    for name in util.ROBOT_FILES:
        key = name.replace(".xml", "")
        path = "/root/reference/ctr_resources/" + name
        ref = ref_lib.RefRobot.from_xml(path, 256)
        d = ctrfk.desc_from_xml(path, 256)
        n_c, tip_c, jv_c = ref.last_state()
        out[key + "/ctor/n"], out[key + "/ctor/tip"], out[key + "/ctor/jv"] = np.int32(n_c), tip_c, jv_c
        # entry frame as the reference's reader left it: FK of the all-zero joint vector with N = 1 has h = 0,
        # every step is the identity and Rz(0) = I, so the only frame IS InitTrafo
        z = ref.fk_batch(np.zeros((1, ref.njoints)), ctrfk.MODE_NSAMPLE, nsamples=1)
        out[key + "/init"] = z["tip"][0]
        edge = util.edge_joint_sets(d)
        edge = edge[[sum(q[i] for i in util.phi_indices(d)) > 0 for q in edge]]
        jv = np.concatenate([ctrfk.sample_joints_host(d, util.SEEDS[name], 0, 128 - len(edge)), edge])
        out[key + "/jv"] = jv
        out[key + "/mt_jv"] = ref.random_joints(16)
        for tag, mode in (("F", ctrfk.MODE_NSAMPLE), ("S", ctrfk.MODE_STEP)):
            r = ref.fk_batch(jv, mode, step=ref.max_length / 256.0, nsamples=256, backbone=True, radius=True)
            out["%s/%s/n" % (key, tag)] = r["n"]
            out["%s/%s/step" % (key, tag)] = r["step"]
            out["%s/%s/tip" % (key, tag)] = r["tip"]
            idx = np.minimum(np.array(PROBES)[None, :], r["n"][:, None] - 1)
            out["%s/%s/probe_idx" % (key, tag)] = idx.astype(np.int32)
            out["%s/%s/probe_frames" % (key, tag)] = np.take_along_axis(r["backbone"], idx[:, :, None], axis=1)
            out["%s/%s/radius" % (key, tag)] = r["radius"].astype(np.float32)     # radii are 0.0005/0.001/0.00125: exact compare after the same cast
        out[key + "/jac"] = np.array([ref.jacobian(q, 128, 1e-5, 1e-4)[0] for q in jv[:16]])
        st, dg = [], []
        for q in jv[:32]:
            st.append(ref.stability(q, 0.2, ref.max_length / 128.0))
            dg.append(ref.degree_stability(q, 0.2, ref.max_length / 128.0, -1.0)[0])
        out[key + "/stable"] = np.array(st, np.int8)
        out[key + "/degree"] = np.array(dg)
    np.savez_compressed(os.path.join(HERE, "reference_golden.npz"), **out)
    print("wrote", os.path.join(HERE, "reference_golden.npz"), len(out), "arrays")


if __name__ == "__main__":
    main()
