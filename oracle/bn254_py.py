"""TEST INFRASTRUCTURE — pure-Python big-int restatement of the BN254 arithmetic on the
PlonK proving path of renegade-fi/renegade.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline leg may import this; the product path never does.

PARITY STATUS: "parity unpinned" at proof-byte level.  The reference holds no golden
vector, KAT or fixture for MSM / NTT / proofs and its arithmetic lives
in un-vendored crates (ark-ff/ark-ec/ark-poly 0.4.2, mpc-jellyfish @311568a4; Cargo.lock:974,
1046,1199,5024).  What *is* pinned against reference-owned data:
  * the SRS file /root/reference/srs/srs0{0,1,2} (parsed as
    crates/circuits/circuit-types/src/primitives/srs.rs:63-209 does): every point we keep as a
    fixture is on-curve under this file's Fq/Montgomery arithmetic, which pins modulus, R and
    limb layout;
  * the curve/field selection of crates/constants/src/lib.rs:63-89 (Bn254, Fr, G1).
MSM and NTT outputs are canonical group / field elements, so any correct algorithm is
bit-identical with arkworks after affine normalisation.

Each function cites what it restates.  Algorithms follow the *published* definitions
(ark-ec 0.4.2 VariableBaseMSM::msm_bigint, ark-poly 0.4.2 Radix2EvaluationDomain).
"""
from __future__ import annotations

import struct

# ---------------------------------------------------------------------------------------
# Field constants (ark-bn254 0.4.0 FqConfig / FrConfig)
# ---------------------------------------------------------------------------------------
Q = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47  # base field Fq
R = 0x30644E72E131A029B85045B68181585D2833E84879B97091_43E1F593F0000001  # scalar field Fr
MONT_R = 1 << 256
FR_GENERATOR = 5  # ark-bn254 FrConfig::GENERATOR (multiplicative generator, coset shift)
FR_TWO_ADICITY = 28
# 2^28-th primitive root of unity: 5^((r-1)/2^28) (ark-bn254 FrConfig::TWO_ADIC_ROOT_OF_UNITY)
FR_ROOT_2_28 = pow(FR_GENERATOR, (R - 1) >> FR_TWO_ADICITY, R)
CURVE_B = 3  # y^2 = x^3 + 3
G1_GEN = (1, 2)


def to_mont(a: int, p: int) -> int:
    """a -> a*R mod p  (ark-ff MontBackend representation, 4x u64 LE limbs)."""
    return (a << 256) % p


def from_mont(a: int, p: int) -> int:
    """a*R -> a (ark-ff `into_bigint`)."""
    return (a * pow(MONT_R, -1, p)) % p


def mont_mul(a: int, b: int, p: int) -> int:
    """Montgomery product a*b*R^-1 mod p (ark-ff montgomery_backend.rs mul_assign)."""
    return (a * b * pow(MONT_R, -1, p)) % p


def limbs4(a: int) -> list[int]:
    return [(a >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)]


def to_bytes32_le(a: int) -> bytes:
    return a.to_bytes(32, "little")


def from_bytes32_le(b: bytes) -> int:
    return int.from_bytes(b, "little")


# ---------------------------------------------------------------------------------------
# G1 (ark-ec short_weierstrass affine/projective; here plain affine with None = infinity)
# ---------------------------------------------------------------------------------------
def g1_is_on_curve(P) -> bool:
    """G1Affine::is_on_curve, as called by srs.rs:178-179."""
    if P is None:
        return True
    x, y = P
    return (y * y - x * x * x - CURVE_B) % Q == 0


def g1_neg(P):
    if P is None:
        return None
    return (P[0], (-P[1]) % Q)


def g1_add(P, S):
    if P is None:
        return S
    if S is None:
        return P
    x1, y1 = P
    x2, y2 = S
    if x1 == x2:
        if (y1 + y2) % Q == 0:
            return None
        lam = (3 * x1 * x1) * pow(2 * y1, -1, Q) % Q
    else:
        lam = (y2 - y1) * pow(x2 - x1, -1, Q) % Q
    x3 = (lam * lam - x1 - x2) % Q
    y3 = (lam * (x1 - x3) - y1) % Q
    return (x3, y3)


# Jacobian for speed in scalar muls
def _jac_double(P):
    X, Y, Z = P
    if Z == 0:
        return P
    A = X * X % Q
    B = Y * Y % Q
    C = B * B % Q
    D = 2 * ((X + B) * (X + B) - A - C) % Q
    E = 3 * A % Q
    F = E * E % Q
    X3 = (F - 2 * D) % Q
    Y3 = (E * (D - X3) - 8 * C) % Q
    Z3 = 2 * Y * Z % Q
    return (X3, Y3, Z3)


def _jac_add_affine(P, A):
    X1, Y1, Z1 = P
    if A is None:
        return P
    x2, y2 = A
    if Z1 == 0:
        return (x2, y2, 1)
    Z1Z1 = Z1 * Z1 % Q
    U2 = x2 * Z1Z1 % Q
    S2 = y2 * Z1 * Z1Z1 % Q
    if U2 == X1:
        if S2 == Y1:
            return _jac_double(P)
        return (1, 1, 0)
    H = (U2 - X1) % Q
    HH = H * H % Q
    HHH = H * HH % Q
    rr = (S2 - Y1) % Q
    V = X1 * HH % Q
    X3 = (rr * rr - HHH - 2 * V) % Q
    Y3 = (rr * (V - X3) - Y1 * HHH) % Q
    Z3 = Z1 * H % Q
    return (X3, Y3, Z3)


def _jac_to_affine(P):
    X, Y, Z = P
    if Z == 0:
        return None
    zi = pow(Z, -1, Q)
    zi2 = zi * zi % Q
    return (X * zi2 % Q, Y * zi2 * zi % Q)


def g1_mul(P, k: int):
    """Scalar multiplication k*P (double-and-add, MSB first)."""
    k %= R
    if P is None or k == 0:
        return None
    acc = (1, 1, 0)
    for bit in bin(k)[2:]:
        acc = _jac_double(acc)
        if bit == "1":
            acc = _jac_add_affine(acc, P)
    return _jac_to_affine(acc)


def msm_naive(bases, scalars):
    """sum_i scalars[i] * bases[i] — the definition VariableBaseMSM::msm_bigint computes."""
    acc = None
    for P, s in zip(bases, scalars):
        acc = g1_add(acc, g1_mul(P, s))
    return acc


def ark_window_bits(n: int) -> int:
    """ark-ec 0.4.2 scalar_mul/variable_base/mod.rs: c = 3 if n<32 else ln_without_floats(n)+2,
    ln_without_floats(a) = log2(a) * 69 / 100 with ark_std::log2 = ceil(log2)."""
    if n < 32:
        return 3
    lg = (n - 1).bit_length()  # ceil(log2 n) for n >= 2
    return lg * 69 // 100 + 2


def msm_pippenger(bases, scalars):
    """Restatement of ark-ec 0.4.2 msm_bigint: unsigned c-bit windows over 254 bits, one bucket
    set per window, running-sum bucket reduction, windows combined high->low with c doublings.
    (SURVEY.md App. B.)  Uses Jacobian accumulators; output affine."""
    n = min(len(bases), len(scalars))
    c = ark_window_bits(n)
    num_bits = 254
    window_sums = []
    for w_start in range(0, num_bits, c):
        buckets = [(1, 1, 0)] * ((1 << c) - 1)
        for P, s in zip(bases, scalars):
            if s == 0 or P is None:
                continue
            d = (s >> w_start) & ((1 << c) - 1)
            if d:
                buckets[d - 1] = _jac_add_affine(buckets[d - 1], P)
        running = (1, 1, 0)
        res = (1, 1, 0)
        for b in reversed(buckets):
            running = _jac_add_full(running, b)
            res = _jac_add_full(res, running)
        window_sums.append(res)
    total = window_sums[-1]
    for ws in reversed(window_sums[:-1]):
        for _ in range(c):
            total = _jac_double(total)
        total = _jac_add_full(total, ws)
    return _jac_to_affine(total)


def _jac_add_full(P, S):
    X1, Y1, Z1 = P
    X2, Y2, Z2 = S
    if Z1 == 0:
        return S
    if Z2 == 0:
        return P
    Z1Z1 = Z1 * Z1 % Q
    Z2Z2 = Z2 * Z2 % Q
    U1 = X1 * Z2Z2 % Q
    U2 = X2 * Z1Z1 % Q
    S1 = Y1 * Z2 * Z2Z2 % Q
    S2 = Y2 * Z1 * Z1Z1 % Q
    if U1 == U2:
        if S1 == S2:
            return _jac_double(P)
        return (1, 1, 0)
    H = (U2 - U1) % Q
    HH = H * H % Q
    HHH = H * HH % Q
    rr = (S2 - S1) % Q
    V = U1 * HH % Q
    X3 = (rr * rr - HHH - 2 * V) % Q
    Y3 = (rr * (V - X3) - S1 * HHH) % Q
    Z3 = Z1 * Z2 * H % Q
    return (X3, Y3, Z3)


# ---------------------------------------------------------------------------------------
# Radix-2 evaluation domain (ark-poly 0.4.2 Radix2EvaluationDomain<Fr>)
# ---------------------------------------------------------------------------------------
def domain_generator(log_n: int) -> int:
    """group_gen of the size-2^log_n subgroup: TWO_ADIC_ROOT^(2^(28-log_n))."""
    assert 0 <= log_n <= FR_TWO_ADICITY
    return pow(FR_ROOT_2_28, 1 << (FR_TWO_ADICITY - log_n), R)


def dft_naive(x, inverse=False):
    """O(n^2) definition: X[k] = sum_j x[j] w^(jk)  (natural order in and out)."""
    n = len(x)
    log_n = n.bit_length() - 1
    w = domain_generator(log_n)
    if inverse:
        w = pow(w, -1, R)
    out = []
    for k in range(n):
        wk = pow(w, k, R)
        acc, t = 0, 1
        for j in range(n):
            acc = (acc + x[j] * t) % R
            t = t * wk % R
        out.append(acc)
    if inverse:
        ninv = pow(n, -1, R)
        out = [v * ninv % R for v in out]
    return out


def ntt(x, inverse=False):
    """Radix2EvaluationDomain::{fft,ifft}_in_place semantics: natural order in/out,
    ifft scales by n^-1.  Iterative radix-2 DIT after a bit-reversal permutation."""
    n = len(x)
    log_n = n.bit_length() - 1
    assert 1 << log_n == n
    w = domain_generator(log_n)
    if inverse:
        w = pow(w, -1, R)
    a = list(x)
    for i in range(n):  # bit reversal
        j = int(bin(i)[2:].zfill(log_n)[::-1], 2) if log_n else 0
        if i < j:
            a[i], a[j] = a[j], a[i]
    m = 1
    while m < n:
        wm = pow(w, n // (2 * m), R)
        for k in range(0, n, 2 * m):
            t = 1
            for j in range(m):
                u = a[k + j]
                v = a[k + j + m] * t % R
                a[k + j] = (u + v) % R
                a[k + j + m] = (u - v) % R
                t = t * wm % R
        m *= 2
    if inverse:
        ninv = pow(n, -1, R)
        a = [v * ninv % R for v in a]
    return a


def coset_ntt(x, inverse=False, shift=FR_GENERATOR):
    """coset_fft: scale coefficient i by g^i then fft; coset_ifft: ifft then scale by g^-i
    (ark-poly distribute_powers)."""
    n = len(x)
    if not inverse:
        t, y = 1, []
        for v in x:
            y.append(v * t % R)
            t = t * shift % R
        return ntt(y)
    y = ntt(x, inverse=True)
    gi = pow(shift, -1, R)
    t, out = 1, []
    for v in y:
        out.append(v * t % R)
        t = t * gi % R
    return out


# ---------------------------------------------------------------------------------------
# ptau parsing (restates crates/circuits/circuit-types/src/primitives/srs.rs:63-209)
# ---------------------------------------------------------------------------------------
MAX_SRS_POWER = 17
MAX_SRS_DEGREE = (1 << MAX_SRS_POWER) + 2  # srs.rs:44-47


def parse_ptau_g1(data: bytes, count: int = MAX_SRS_DEGREE + 1):
    """Returns (raw_bytes_of_section2_prefix, n_points).  Header: srs.rs:74-92, section 1
    :95-118, section 2 :124-141.  Points are x||y, 4xu64 LE each, already Montgomery
    (srs.rs:201-209)."""
    assert data[:4] == b"ptau"
    version, n_sections = struct.unpack_from("<II", data, 4)
    assert version == 1 and n_sections == 11
    off = 12
    sec, size = struct.unpack_from("<IQ", data, off)
    assert sec == 1
    off += 12
    (mod_bytes,) = struct.unpack_from("<I", data, off)
    modulus = int.from_bytes(data[off + 4 : off + 4 + mod_bytes], "little")
    assert modulus == Q
    power, _ceremony = struct.unpack_from("<II", data, off + 4 + mod_bytes)
    assert power >= MAX_SRS_POWER
    off += size
    sec, size = struct.unpack_from("<IQ", data, off)
    assert sec == 2
    off += 12
    assert count * 64 <= size
    return data[off : off + 64 * count], count


def decode_g1_mont(raw: bytes, i: int):
    """One 64-byte record -> affine (x, y) as canonical integers."""
    xm = from_bytes32_le(raw[64 * i : 64 * i + 32])
    ym = from_bytes32_le(raw[64 * i + 32 : 64 * i + 64])
    return (from_mont(xm, Q), from_mont(ym, Q))


def encode_g1_mont(P) -> bytes:
    """affine canonical -> 64-byte Montgomery record (infinity encoded as all-zero)."""
    if P is None:
        return bytes(64)
    return to_bytes32_le(to_mont(P[0], Q)) + to_bytes32_le(to_mont(P[1], Q))


# ---------------------------------------------------------------------------------------
# Synthetic inputs (SplitMix64 seeds 0xB200 / 0x5CA1A8 / 0x1177)
# ---------------------------------------------------------------------------------------
M64 = 0xFFFFFFFFFFFFFFFF


def splitmix64_stream(seed: int):
    s = seed & M64
    while True:
        s = (s + 0x9E3779B97F4A7C15) & M64
        z = s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M64
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M64
        yield z ^ (z >> 31)


def splitmix_fr(seed: int, n: int):
    """n field elements: 4 consecutive outputs -> 256-bit LE integer, reduced mod r."""
    g = splitmix64_stream(seed)
    out = []
    for _ in range(n):
        v = 0
        for k in range(4):
            v |= next(g) << (64 * k)
        out.append(v % R)
    return out


# ---------------------------------------------------------------------------------------
# Poseidon2 (restates crates/crypto/src/hash/poseidon2.rs:25-209; t = 3, rate 2, capacity 1,
# R_F = 8, R_P = 56, alpha = 5 — constants.rs:15-36).  Round constants are read from the fixture
# tests/golden/poseidon2.json, extracted from the reference's constants.rs; the permutation is
# pinned by the published HorizenLabs known answer for the input (0, 1, 2).
# ---------------------------------------------------------------------------------------
def poseidon2_load_constants(path):
    import json
    with open(path) as f:
        d = json.load(f)
    return [int(v, 16) for v in d["full_round_constants"]], [int(v, 16) for v in d["partial_round_constants"]]


def poseidon2_permute(state, full, partial):
    """Poseidon2Sponge::permute (poseidon2.rs:90-110)."""
    def ext_mds(s):  # poseidon2.rs:146-152
        t = sum(s) % R
        return [(x + t) % R for x in s]

    def int_mds(s):  # poseidon2.rs:187-195
        t = sum(s) % R
        s = [s[0], s[1], 2 * s[2] % R]
        return [(x + t) % R for x in s]
    st = ext_mds(list(state))
    for r in range(4):
        st = ext_mds([pow((x + full[3 * r + i]) % R, 5, R) for i, x in enumerate(st)])
    for r in range(56):
        st = int_mds([pow((st[0] + partial[r]) % R, 5, R), st[1], st[2]])
    for r in range(4, 8):
        st = ext_mds([pow((x + full[3 * r + i]) % R, 5, R) for i, x in enumerate(st)])
    return st


def poseidon2_hash(values, full, partial):
    """Poseidon2Sponge::new().hash(values) = compute_poseidon_hash (mod.rs:12-18): absorb at rate 2
    into state[1..3] (poseidon2.rs:47-59), one squeeze (poseidon2.rs:67-79)."""
    st, nxt = [0, 0, 0], 0
    for x in values:
        if nxt == 2:
            st = poseidon2_permute(st, full, partial)
            nxt = 0
        st[nxt + 1] = (st[nxt + 1] + x) % R
        nxt += 1
    st = poseidon2_permute(st, full, partial)
    return st[1]
