"""Poseidon2 batch hashing.  CPU: the oracle restatement of
crates/crypto/src/hash/poseidon2.rs reproduces the published known answer with the reference's
constants.  GPU: the device kernels match the oracle bit-exactly."""
import json
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FIX = os.path.join(ROOT, "tests", "golden", "poseidon2.json")


def test_oracle_matches_published_kat(pyoracle):
    py = pyoracle
    full, partial = py.poseidon2_load_constants(FIX)
    assert len(full) == 24 and len(partial) == 56
    kat = json.load(open(FIX))["kat_permutation_0_1_2"]["output"]
    assert py.poseidon2_permute([0, 1, 2], full, partial) == [int(v, 16) for v in kat]


def test_generated_header_matches_fixture(pyoracle):
    """The constants compiled into the product are the reference's (Montgomery form of the fixture)."""
    import re
    py = pyoracle
    full, partial = py.poseidon2_load_constants(FIX)
    txt = open(os.path.join(ROOT, "renegade_b200", "csrc", "poseidon2_constants.h")).read()
    rows = re.findall(r"\{((?:0x[0-9a-f]{8}u(?:, )?){8})\}", txt)
    vals = [sum(int(w.rstrip("u"), 16) << (32 * i) for i, w in enumerate(r.split(", "))) for r in rows]
    assert vals == [py.to_mont(v, py.R) for v in full + partial]


@pytest.mark.gpu
def test_gpu_permutation_and_hash(ctx, oracle, pyoracle):
    from renegade_b200.backend import compute_poseidon_hash_batch, poseidon2_permute_batch
    py = pyoracle
    full, partial = py.poseidon2_load_constants(FIX)
    to_m = lambda vals: oracle.ints_to_array([py.to_mont(v, py.R) for v in vals])
    from_m = lambda arr: [py.from_mont(v, py.R) for v in oracle.array_to_ints(arr)]
    # the published known answer through the device permutation
    kat = [int(v, 16) for v in json.load(open(FIX))["kat_permutation_0_1_2"]["output"]]
    st = to_m([0, 1, 2]).reshape(1, 3, 4)
    assert from_m(poseidon2_permute_batch(ctx, st).reshape(3, 4)) == kat
    # random states
    import random
    rnd = random.Random(5)
    states = [[rnd.randrange(py.R) for _ in range(3)] for _ in range(300)]
    got = poseidon2_permute_batch(ctx, to_m([v for s in states for v in s]).reshape(-1, 3, 4))
    for i, s in enumerate(states[:40]):
        assert from_m(got[i]) == py.poseidon2_permute(s, full, partial)
    # sponge hashes of every length 0..7 (odd lengths leave the rate half full; 0 = hash of nothing)
    for ln in range(0, 8):
        vals = [[rnd.randrange(py.R) for _ in range(ln)] for _ in range(130)]
        arr = to_m([v for row in vals for v in row]).reshape(130, ln, 4) if ln else np.zeros((130, 0, 4), dtype=np.uint64)
        out = compute_poseidon_hash_batch(ctx, arr)
        exp = [py.poseidon2_hash(row, full, partial) for row in vals[:25]]
        assert from_m(out[:25]) == exp, ln


@pytest.mark.gpu
def test_gpu_merkle_roots_and_csprng_streams(ctx, oracle, pyoracle):
    """The batch callers of the witness side: Merkle roots of openings of height 10 (`MERKLE_HEIGHT`,
    constants/src/lib.rs:50) and Poseidon CSPRNG stream values H(seed, i) (darkpool-types/src/csprng.rs:30-75), device
    vs the oracle's Poseidon2 on Python integers."""
    import random
    from renegade_b200.backend import csprng_batch, merkle_root_batch
    py = pyoracle
    full, partial = py.poseidon2_load_constants(FIX)
    H = lambda vals: py.poseidon2_hash(vals, full, partial)
    to_m = lambda vals: oracle.ints_to_array([py.to_mont(v, py.R) for v in vals])
    from_m = lambda arr: [py.from_mont(v, py.R) for v in oracle.array_to_ints(arr)]
    rnd = random.Random(11)
    batch, height = 67, 10
    leaves = [rnd.randrange(py.R) for _ in range(batch)]
    sisters = [[rnd.randrange(py.R) for _ in range(height)] for _ in range(batch)]
    bits = [[rnd.random() < 0.5 for _ in range(height)] for _ in range(batch)]
    got = from_m(merkle_root_batch(ctx, to_m(leaves), to_m([v for row in sisters for v in row]).reshape(batch, height, 4),
                                   np.array(bits, dtype=np.uint8)))
    for i in range(0, batch, 7):
        cur = leaves[i]
        for sis, right in zip(sisters[i], bits[i]):
            cur = H([sis, cur] if right else [cur, sis])
        assert got[i] == cur, i
    # height 0: the root is the leaf hash itself
    assert from_m(merkle_root_batch(ctx, to_m(leaves[:3]), np.zeros((3, 0, 4), dtype=np.uint64), np.zeros((3, 0), dtype=np.uint8))) == leaves[:3]
    seeds = [rnd.randrange(py.R) for _ in range(9)]
    first = [rnd.randrange(0, 1 << 40) for _ in range(9)]
    out = csprng_batch(ctx, to_m(seeds), first, 5)
    for s in (0, 4, 8):
        assert from_m(out[s]) == [H([seeds[s], first[s] + j]) for j in range(5)]
