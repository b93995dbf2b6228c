import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "examples"))  # host_circuits (caller-side example package)
sys.path.insert(0, os.path.join(ROOT, "tests"))     # host_backend (test infrastructure)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # a fresh checkout has no built library (it is git-ignored): build it once, in-tree
    lib = os.path.join(ROOT, "renegade_b200", "libb200prover.so")
    if not os.path.exists(lib):
        import subprocess
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "renegade_b200", "csrc"), "-j4", "-s"])


@pytest.fixture(scope="session")
def oracle():
    """The C restatement (test infrastructure)."""
    import oracle_c
    oracle_c.build()
    return oracle_c


@pytest.fixture(scope="session")
def pyoracle():
    import bn254_py
    return bn254_py


@pytest.fixture(scope="session")
def kat():
    import json
    with open(os.path.join(ROOT, "tests", "golden", "kat.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def srs_head():
    with open(os.path.join(ROOT, "tests", "golden", "srs_head.bin"), "rb") as f:
        return f.read()


@pytest.fixture(scope="session")
def srs_2_16():
    """The reference's own SRS, first 2^16 + 3 G1 powers (tests/golden/_large/srs_2_16.bin, cut by
    tools/cut_srs_fixture.py from /root/reference/srs/srs00; git-ignored, ships with the gpurun snapshot)."""
    path = os.path.join(ROOT, "tests", "golden", "_large", "srs_2_16.bin")
    if not os.path.exists(path) and os.path.exists("/root/reference/srs/srs00"):
        import subprocess
        subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "cut_srs_fixture.py")])
    if not os.path.exists(path):
        pytest.skip("large SRS fixture absent (cut it where /root/reference exists: tools/cut_srs_fixture.py)")
    with open(path, "rb") as f:
        return f.read()


@pytest.fixture(scope="session")
def g2_raw():
    """(h, tau * h): the two G2 records of the reference's SRS (tests/golden/srs_g2.bin), 16 x uint64 each."""
    import numpy as np
    raw = np.frombuffer(open(os.path.join(ROOT, "tests", "golden", "srs_g2.bin"), "rb").read(), dtype=np.uint64)
    return raw[:16].copy(), raw[16:32].copy()


@pytest.fixture(scope="session")
def ctx():
    """GPU context; the extension must be present — no skipping to a CPU path."""
    import renegade_b200 as rb
    c = rb.Context(0)
    yield c
    c.close()
