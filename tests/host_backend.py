"""TEST INFRASTRUCTURE: a CPU backend for renegade_b200/collaborative.py built on the oracle (C restatement) and Python
integers, so that the protocol logic of the collaborative prover is checked without a GPU.  Vectors are lists of
canonical integers wrapped in `HVec`; transforms and commitments go through `oracle_c`."""
import random

import numpy as np

import oracle_c

R = 0x30644E72E131A029B85045B68181585D2833E84879B97091_43E1F593F0000001
RINV = pow(1 << 256, -1, R)


class HVec:
    def __init__(self, vals):
        self.v = list(vals)

    @property
    def shape(self):
        return (len(self.v), 4)

    def __getitem__(self, s):
        assert isinstance(s, slice)
        return HVec(self.v[s])


def _to_mont_arr(vals):
    return oracle_c.ints_to_array([(x << 256) % R for x in vals])


def _from_mont_arr(arr):
    return [x * RINV % R for x in oracle_c.array_to_ints(arr)]


class HostBackend:
    def __init__(self, srs: np.ndarray):
        self.srs = np.ascontiguousarray(srs, dtype=np.uint64).reshape(-1, 8)

    def from_limbs(self, arr):
        return HVec(_from_mont_arr(np.asarray(arr, dtype=np.uint64).reshape(-1, 4)))

    def to_limbs(self, v):
        return _to_mont_arr(v.v)

    def zeros(self, n):
        return HVec([0] * n)

    def concat(self, parts):
        out = []
        for p in parts:
            out += p.v
        return HVec(out)

    def clone(self, v):
        return HVec(v.v)

    def roll(self, v, shift):
        k = shift % len(v.v)
        return HVec(v.v[k:] + v.v[:k])

    def random(self, seed, n):
        rnd = random.Random(seed)
        return HVec([rnd.randrange(R) for _ in range(n)])

    def _bin(self, a, b, f):
        if len(b.v) == 1 and len(a.v) != 1:
            return HVec([f(x, b.v[0]) for x in a.v])
        assert len(a.v) == len(b.v)
        return HVec([f(x, y) for x, y in zip(a.v, b.v)])

    def add(self, a, b):
        return self._bin(a, b, lambda x, y: (x + y) % R)

    def sub(self, a, b):
        return self._bin(a, b, lambda x, y: (x - y) % R)

    def mul(self, a, b):
        return self._bin(a, b, lambda x, y: x * y % R)

    def scalar(self, v):
        return HVec([v % R])

    def batch_inverse(self, a):
        return HVec([pow(x, -1, R) if x else 0 for x in a.v])

    def ntt(self, a, inverse, coset):
        return HVec(_from_mont_arr(oracle_c.ntt(_to_mont_arr(a.v), inverse=inverse, coset=coset)))

    def commit(self, coeffs):
        return oracle_c.msm(self.srs[:len(coeffs.v)], oracle_c.ints_to_array(coeffs.v))

    def poly_eval(self, coeffs, z):
        acc = 0
        for c in reversed(coeffs.v):
            acc = (acc * z + c) % R
        return acc

    def div_linear(self, coeffs, z):
        out, carry = [0] * (len(coeffs.v) - 1), 0
        for i in range(len(coeffs.v) - 1, 0, -1):
            carry = (coeffs.v[i] + carry * z) % R
            out[i - 1] = carry
        return HVec(out)

    def g1_sum(self, points):
        acc, acc_inf = np.zeros(8, dtype=np.uint64), True
        for xy, inf in points:
            acc, acc_inf = oracle_c.g1_add(acc, acc_inf, xy, inf)
        return np.zeros(8, dtype=np.uint64) if acc_inf else acc
