"""CPU: the host-side field inversion of the library (`fe_inv_safegcd`, renegade_b200/csrc/ff.cuh — Bernstein-Yang division
steps, 62 at a time; on the prover's critical path in the affine normalisation of every commitment batch, and ~130 times
per pairing) against the Fermat ladder and the binary Euclid, for Fq and Fr.  The header's host path is plain C++, so the
check is a small program built with g++ (tests/host_inverse_check.cpp)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_safegcd_inverse_matches_fermat_and_euclid(tmp_path):
    exe = tmp_path / "host_inverse_check"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-I" + os.path.join(ROOT, "renegade_b200", "csrc"),
                           "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "host_inverse_check.cpp"), "-o", str(exe)])
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "Fq: mismatches 0" in out.stdout and "Fr: mismatches 0" in out.stdout
