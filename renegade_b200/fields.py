"""Field selection and conversions at the boundary.

`constants::{ScalarField, SystemCurve}` (crates/constants/src/lib.rs:63-89) fix BN254: scalars live in Fr, commitments
in G1 over Fq.  `crypto::fields` (crates/crypto/src/fields.rs:21-163) converts scalars to and from integers with
big-endian byte semantics; this module restates those helpers on Python integers and adds the one conversion the C ABI
needs: the 4 x u64 little-endian Montgomery limbs of ark-ff's in-memory `Fp256` (include/b200prover.h conventions).
"""
from __future__ import annotations

from typing import List

import numpy as np

SCALAR_FIELD_MODULUS = 0x30644E72E131A029B85045B68181585D2833E84879B97091_43E1F593F0000001  # r
BASE_FIELD_MODULUS = 0x30644E72E131A029B85045B68181585D97816A916871CA8D_3C208C16D87CFD47    # q
U256_BYTES = 32
_MASK64 = (1 << 64) - 1


def get_scalar_field_modulus() -> int:
    """fields.rs:21-23."""
    return SCALAR_FIELD_MODULUS


def get_base_field_modulus() -> int:
    """fields.rs:26-28."""
    return BASE_FIELD_MODULUS


# ---- conversions from a scalar (an integer in [0, r)) ------------------------------------------
def scalar_to_bytes_be(a: int) -> bytes:
    """`Scalar::to_bytes_be`: 32 bytes, big-endian."""
    return (a % SCALAR_FIELD_MODULUS).to_bytes(U256_BYTES, "big")


def scalar_to_biguint(a: int) -> int:
    return a % SCALAR_FIELD_MODULUS


def scalar_to_u64(a: int) -> int:
    """fields.rs:76-84: the low 8 bytes, anything above 2^64 - 1 truncated."""
    return int.from_bytes(scalar_to_bytes_be(a)[-8:], "big")


def scalar_to_u128(a: int) -> int:
    """fields.rs:87-95."""
    return int.from_bytes(scalar_to_bytes_be(a)[-16:], "big")


def scalar_to_address(a: int) -> bytes:
    """fields.rs:43-53: the lowest 20 bytes."""
    return scalar_to_bytes_be(a)[-20:]


def scalar_to_u256(a: int) -> int:
    """fields.rs:56-61."""
    return int.from_bytes(scalar_to_bytes_be(a), "big")


# ---- conversions to a scalar ---------------------------------------------------------------------
def biguint_to_scalar(a: int) -> int:
    """fields.rs:118-120 (`Scalar::from(BigUint)` reduces mod r)."""
    if a < 0:
        raise ValueError("biguint_to_scalar takes a non-negative integer")
    return a % SCALAR_FIELD_MODULUS


def bigint_to_scalar(a: int) -> int:
    """fields.rs:106-115: a negative integer maps to the negation of its magnitude."""
    return (-((-a) % SCALAR_FIELD_MODULUS)) % SCALAR_FIELD_MODULUS if a < 0 else a % SCALAR_FIELD_MODULUS


def bigint_to_scalar_bits(a: int, d: int) -> List[int]:
    """fields.rs:128-137: the `d` low bits, little-endian, as 0/1 scalars."""
    mag = abs(a)
    return [(mag >> i) & 1 for i in range(d)]


def bytes_be_to_scalar(b: bytes) -> int:
    """`Scalar::from_be_bytes_mod_order` (used by `address_to_scalar` / `u256_to_scalar`, fields.rs:144-153)."""
    return int.from_bytes(b, "big") % SCALAR_FIELD_MODULUS


def address_to_scalar(address20: bytes) -> int:
    if len(address20) != 20:
        raise ValueError("an address is 20 bytes")
    return bytes_be_to_scalar(address20)


def u256_to_scalar(a: int) -> int:
    return bytes_be_to_scalar(a.to_bytes(U256_BYTES, "big"))


# ---- the ABI representation ----------------------------------------------------------------------
def scalars_to_limbs(values, modulus: int = SCALAR_FIELD_MODULUS) -> np.ndarray:
    """integers -> (len, 4) uint64 little-endian Montgomery limbs (a * 2^256 mod p), ark-ff's `Fp256` in memory."""
    out = np.empty((len(values), 4), dtype=np.uint64)
    for i, v in enumerate(values):
        m = ((v % modulus) << 256) % modulus
        for k in range(4):
            out[i, k] = (m >> (64 * k)) & _MASK64
    return out


def limbs_to_scalars(limbs: np.ndarray, modulus: int = SCALAR_FIELD_MODULUS) -> List[int]:
    """inverse of `scalars_to_limbs`."""
    rinv = pow(1 << 256, -1, modulus)
    arr = np.ascontiguousarray(limbs, dtype=np.uint64).reshape(-1, 4)
    return [sum(int(row[k]) << (64 * k) for k in range(4)) * rinv % modulus for row in arr]
