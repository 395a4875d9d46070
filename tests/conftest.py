import os
import subprocess
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(REPO, "ram2017-ctr-kinematics_b200")
for p in (REPO, PKG, os.path.join(REPO, "oracle"), os.path.join(REPO, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def ctrfk():
    import ctrfk as m
    if not os.path.exists(m.LIB_PATH):
        subprocess.check_call(["make", "-s", "-j8", "-C", PKG, "libctrfk.so"])
    m.load()
    return m


@pytest.fixture(scope="session")
def oracle():
    import oracle_lib
    oracle_lib.load()
    return oracle_lib


@pytest.fixture(scope="session")
def hostsim():
    import util
    return util.load_hostsim()


@pytest.fixture(scope="session")
def robots(ctrfk):
    names = ["ctr_sec2_tub3.xml", "ctr_sec3_tub3.xml", "ctr_sec3_tub4.xml", "ctr_sec3_tub5.xml"]
    return {n: ctrfk.desc_from_xml(ctrfk.resource(n), 256) for n in names}
