"""GPU: a compiled C host (examples/host_harness.c) drives the library through the C ABI alone —
no Python, no torch in the process — and produces the same proof bytes as the oracle prover."""
import os
import subprocess

import numpy as np
import pytest

from renegade_b200 import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TAU = 0x0123456789abcdef0fedcba9876543210011223344556677


def test_c_host_harness(tmp_path, oracle, pyoracle):
    py = pyoracle
    exe = os.path.join(ROOT, "examples", "host_harness")
    src = os.path.join(ROOT, "examples", "host_harness.c")
    if not os.path.exists(exe) or os.path.getmtime(exe) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-I" + os.path.join(ROOT, "include"), src, "-L" + os.path.join(ROOT, "renegade_b200"),
                               "-lb200prover", "-Wl,-rpath," + os.path.join(ROOT, "renegade_b200"), "-o", exe])
    log_n, n = 10, 1 << 10
    circ = synth.synth_circuit(log_n, num_inputs=4, seed=21, check=True)
    tau = oracle.int_to_limbs(py.to_mont(TAU % py.R, py.R))
    srs = oracle.srs_from_tau(tau, n + 3)
    bl = synth.splitmix_blinders(0xC)
    job = tmp_path / "job.bin"
    with open(job, "wb") as f:
        f.write(np.array([log_n, circ.num_inputs, n + 3], dtype=np.uint64).tobytes())
        for arr in (circ.k, circ.selectors, circ.perm, circ.wires, circ.pub_inputs, bl, srs):
            f.write(np.ascontiguousarray(arr, dtype=np.uint64).tobytes())
    out = tmp_path / "proof.bin"
    r = subprocess.run([exe, str(job), str(out)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    blob = np.frombuffer(open(out, "rb").read(), dtype=np.uint64)
    opk = oracle.plonk_preprocess(log_n, circ.selectors, circ.perm, circ.k, srs)
    rc, oproof, _, olink = oracle.plonk_prove(log_n, circ.num_inputs, circ.k, opk, circ.wires, circ.pub_inputs, bl, srs, True)
    assert rc == 0
    assert (blob[:144] == oproof.to_array()).all()
    assert (blob[144:144 + (n + 2) * 4].reshape(-1, 4) == olink).all()
    vk = blob[144 + (n + 2) * 4:].reshape(18, 8)
    assert (vk[:13] == opk["selector_comms"]).all() and (vk[13:] == opk["sigma_comms"]).all()
