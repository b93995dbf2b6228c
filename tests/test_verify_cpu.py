"""CPU tests of the product-side verifier (`b200_plonk_verify`, `b200_plonk_verify_link`, `b200_pairing_check`:
renegade_b200/csrc/verify.cu — host code, no device).  It replaces `SingleProverCircuit::verify`
(traits.rs:1003-1019) and `validate_*_link` (proof_linking/intent_only.rs:54-85).

Pinned by the reference's own SRS bytes (tests/golden/srs_head.bin, srs_g2.bin): the pairing must satisfy the
reference's SRS unit test (srs.rs:236-266), and must agree with the independent pure-Python pairing of the oracle.
Proofs come from the oracle prover (test infrastructure) on the real SRS powers."""
import ctypes as C
import os
import random

import numpy as np
import pytest

from renegade_b200 import _lib, synth
from renegade_b200.backend import (B200LinkProof, B200Proof, GroupLayout, PlonkKzgSnark, VerifyingKey, pairing_check,
                                   verify_link_proof)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _g1(srs_head, i):
    return np.frombuffer(srs_head[80 + 64 * i: 80 + 64 * (i + 1)], dtype=np.uint64).copy()


def _neg(pyoracle, xy):
    p = pyoracle.decode_g1_mont(xy.tobytes(), 0)
    return np.frombuffer(pyoracle.encode_g1_mont(pyoracle.g1_neg(p)), dtype=np.uint64).copy()


def test_srs_pairing_ratio_product_pairing(pyoracle, srs_head, g2_raw):
    """srs.rs:236-266 with the product pairing: e(tau^i G, tau H) * e(-tau^(i+1) G, H) == 1."""
    h, tau_h = g2_raw
    rnd = random.Random(2)
    for i in [0] + [rnd.randrange(1, 500) for _ in range(3)]:
        a, b = _g1(srs_head, i), _neg(pyoracle, _g1(srs_head, i + 1))
        assert pairing_check([a, b], [tau_h, h]), i
    assert not pairing_check([_g1(srs_head, 3), _neg(pyoracle, _g1(srs_head, 5))], [tau_h, h])
    # bilinearity: e(77 G, H) * e(-G, 77-fold ... ) is not expressible without G2 arithmetic; use G1 only:
    g = np.frombuffer(pyoracle.encode_g1_mont(pyoracle.G1_GEN), dtype=np.uint64).copy()
    g77 = np.frombuffer(pyoracle.encode_g1_mont(pyoracle.g1_mul(pyoracle.G1_GEN, 77)), dtype=np.uint64).copy()
    # e(77 G, H) == e(G, H)^77  <=>  e(77 G, H) * e(-G, H)^77 == 1: 77 copies of (-G, H)
    assert pairing_check([g77] + [_neg(pyoracle, g)] * 77, [h] * 78)
    # an off-curve point is refused
    bad = g.copy()
    bad[0] ^= np.uint64(1)
    with pytest.raises(_lib.B200Error):
        pairing_check([bad], [h])
    # the identity contributes 1
    assert pairing_check([np.zeros(8, dtype=np.uint64)], [h])


def _prove(oracle, srs_head, log_n, seed, blinder_seed):
    circ = synth.synth_circuit(log_n, num_inputs=4, seed=seed, check=True)
    n = 1 << log_n
    srs = np.frombuffer(srs_head[80:80 + 64 * (n + 3)], dtype=np.uint64).reshape(n + 3, 8).copy()
    pk = oracle.plonk_preprocess(log_n, circ.selectors, circ.perm, circ.k, srs)
    rc, proof, _, link = oracle.plonk_prove(log_n, circ.num_inputs, circ.k, pk, circ.wires, circ.pub_inputs,
                                            synth.splitmix_blinders(blinder_seed), srs, True)
    assert rc == 0
    return circ, srs, pk, proof, link


def test_product_verifier_accepts_and_rejects(oracle, srs_head, g2_raw):
    h, tau_h = g2_raw
    log_n = 7
    circ, srs, pk, proof, _ = _prove(oracle, srs_head, log_n, 31, 8)
    vk = VerifyingKey(log_n, circ.num_inputs, circ.k, pk["selector_comms"], pk["sigma_comms"], h, tau_h)
    p = B200Proof.from_buffer_copy(bytes(proof))
    assert PlonkKzgSnark.verify(vk, circ.pub_inputs, p)
    # agrees with the oracle's verifier + independent Python pairing on the same proof
    import bn254_pairing_py as pr
    g2 = (pr.decode_g2_mont(h.tobytes() + tau_h.tobytes(), 0), pr.decode_g2_mont(h.tobytes() + tau_h.tobytes(), 1))
    assert oracle.plonk_verify_pairing(log_n, circ.num_inputs, circ.k, pk, circ.pub_inputs, proof, *g2)
    # tampering anywhere is rejected: an evaluation, a commitment swapped for another valid point, a public input
    bad = B200Proof.from_buffer_copy(bytes(proof))
    bad.wire_sigma_evals[1][0] ^= 1
    assert not PlonkKzgSnark.verify(vk, circ.pub_inputs, bad)
    bad = B200Proof.from_buffer_copy(bytes(proof))
    for j in range(8):
        bad.wires_poly_comms[0][j], bad.wires_poly_comms[1][j] = bad.wires_poly_comms[1][j], bad.wires_poly_comms[0][j]
    assert not PlonkKzgSnark.verify(vk, circ.pub_inputs, bad)
    bad = B200Proof.from_buffer_copy(bytes(proof))
    bad.opening_proof[0] ^= 1   # off the curve: a rejection, not an error
    assert not PlonkKzgSnark.verify(vk, circ.pub_inputs, bad)
    pi2 = circ.pub_inputs.copy()
    pi2[2, 0] ^= np.uint64(1)
    assert not PlonkKzgSnark.verify(vk, pi2, p)
    # a verifying key of another circuit rejects
    circ2, _, pk2, _, _ = _prove(oracle, srs_head, log_n, 33, 9)
    vk2 = VerifyingKey(log_n, circ2.num_inputs, circ2.k, pk2["selector_comms"], pk2["sigma_comms"], h, tau_h)
    assert not PlonkKzgSnark.verify(vk2, circ.pub_inputs, p)


def test_product_link_verifier(oracle, pyoracle, srs_head, g2_raw):
    h, tau_h = g2_raw
    py = pyoracle
    layout = (5, 6, 5)  # alignment, offset, size
    vals = [(i * 1234567 + 89) % py.R for i in range(layout[2])]
    hints = []
    for log_n, seed in ((6, 3), (7, 4)):
        n = 1 << log_n
        circ = synth.synth_circuit(log_n, num_inputs=2, seed=seed, check=True, link=(layout[0], layout[1], vals))
        srs = np.frombuffer(srs_head[80:80 + 64 * (n + 3)], dtype=np.uint64).reshape(n + 3, 8).copy()
        pk = oracle.plonk_preprocess(log_n, circ.selectors, circ.perm, circ.k, srs)
        rc, proof, _, link = oracle.plonk_prove(log_n, circ.num_inputs, circ.k, pk, circ.wires, circ.pub_inputs,
                                                synth.splitmix_blinders(seed), srs, True)
        assert rc == 0
        hints.append((link, np.array(proof.wires_poly_comms[0], dtype=np.uint64)))
    srs_all = np.frombuffer(srs_head[80:], dtype=np.uint64).reshape(-1, 8)
    rc, lp, _ = oracle.plonk_link(hints[0][0], hints[1][0], hints[0][1], hints[1][1], *layout, srs_all)
    assert rc == 0
    lpp = B200LinkProof.from_buffer_copy(bytes(lp))
    gl = GroupLayout(*layout)
    assert verify_link_proof(hints[0][1], hints[1][1], lpp, gl, h, tau_h)
    assert not verify_link_proof(hints[0][1], hints[1][1], lpp, GroupLayout(layout[0], layout[1] + 1, layout[2]), h, tau_h)
    assert not verify_link_proof(hints[1][1], hints[0][1], lpp, gl, h, tau_h)  # commitments swapped
