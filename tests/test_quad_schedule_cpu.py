"""The four-stage schedule of `xyzz_add_quad` (renegade_b200/csrc/msm.cu) restated with Python integers: lane q of a quad
computes product q of each stage; the sum it assembles must equal the affine sum of the operands, and stay a valid XYZZ
element (ZZ^3 = ZZZ^2).  This pins the formula split (Y3 as two products, ZZZ3 = ((ZZZ1 ZZZ2) PP) P) that the device code
follows; the device kernels themselves are compared with the oracle in tests/test_gpu_msm.py."""
import random

import pytest


@pytest.fixture(scope="module")
def py():
    import bn254_py
    return bn254_py


def quad_add(A, B, p):
    """Stage by stage, [lane 0, lane 1, lane 2, lane 3]; A, B = (X, Y, ZZ, ZZZ)."""
    m1 = [A[0] * B[2] % p, B[0] * A[2] % p, A[1] * B[3] % p, B[1] * A[3] % p]
    u1, u2, s1, s2 = m1
    pd, rd = (u2 - u1) % p, (s2 - s1) % p
    m2 = [pd * pd % p, rd * rd % p, A[2] * B[2] % p, A[3] * B[3] % p]
    pp = m2[0]
    m3 = [pd * pp % p, u1 * pp % p, m2[2] * pp % p, m2[3] * pp % p]
    ppp, q, r2 = m3[0], m3[1], m2[1]
    x3 = (r2 - ppp - 2 * q) % p
    m4 = [rd * ((q - x3) % p) % p, s1 * ppp % p, None, m3[3] * pd % p]
    return (x3, (m4[0] - m4[1]) % p, m3[2], m4[3]), pd == 0


def test_quad_schedule_is_a_group_addition(py):
    p = py.Q
    rng = random.Random(0xB200)
    g = (1, 2)

    def to_xyzz(P):
        z = rng.randrange(1, p)
        zz, zzz = z * z % p, z * z * z % p
        return (P[0] * zz % p, P[1] * zzz % p, zz, zzz)

    def to_affine(X):
        return (X[0] * pow(X[2], p - 2, p) % p, X[1] * pow(X[3], p - 2, p) % p)

    for _ in range(25):
        P, Qp = py.g1_mul(g, rng.randrange(1, py.R)), py.g1_mul(g, rng.randrange(1, py.R))
        S, p_is_zero = quad_add(to_xyzz(P), to_xyzz(Qp), p)
        assert not p_is_zero
        assert pow(S[2], 3, p) == pow(S[3], 2, p)
        assert to_affine(S) == py.g1_add(P, Qp)
    # equal and opposite operands are what the P = 0 branch hands to the exact one-lane addition
    P = py.g1_mul(g, 12345)
    assert quad_add(to_xyzz(P), to_xyzz(P), p)[1]
    assert quad_add(to_xyzz(P), to_xyzz((P[0], (p - P[1]) % p)), p)[1]
