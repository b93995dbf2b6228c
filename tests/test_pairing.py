"""Pairing-level acceptance on the REFERENCE's own SRS (unknown tau): the restated verifier's final
check e(A, [tau]_2) == e(B, [1]_2) is evaluated with a pure-Python optimal-ate pairing on the G1
powers and the two G2 points copied from /root/reference/srs (tests/golden/srs_head.bin, srs_g2.bin).
This is the closest available stand-in for `SingleProverCircuit::verify` (traits.rs:1003-1019)."""
import os
import random

import numpy as np
import pytest

from renegade_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def g2(pyoracle):
    import bn254_pairing_py as pr
    raw = open(os.path.join(ROOT, "tests", "golden", "srs_g2.bin"), "rb").read()
    h, tau_h = pr.decode_g2_mont(raw, 0), pr.decode_g2_mont(raw, 1)
    assert pr.g2_is_on_curve(h) and pr.g2_is_on_curve(tau_h)  # srs.rs:193-194
    return h, tau_h


def test_srs_pairing_ratio(pyoracle, srs_head, g2):
    """The reference's own SRS unit test (srs.rs:236-266): for consecutive powers A = tau^i G,
    B = tau^(i+1) G:  e(A, tau H) == e(B, H).  Also pins the pairing restatement (bilinearity)."""
    import bn254_pairing_py as pr
    py = pyoracle
    h, tau_h = g2
    raw = srs_head[80:]
    rnd = random.Random(1)
    for i in [0] + [rnd.randrange(1, 510) for _ in range(3)]:
        a, b = py.decode_g1_mont(raw, i), py.decode_g1_mont(raw, i + 1)
        assert pr.pairing_product_is_one([(a, tau_h), (py.g1_neg(b), h)]), i
    # a non-consecutive pair must fail, and the pairing is bilinear
    a, c = py.decode_g1_mont(raw, 3), py.decode_g1_mont(raw, 5)
    assert not pr.pairing_product_is_one([(a, tau_h), (py.g1_neg(c), h)])
    g = py.G1_GEN
    assert pr.pairing(h, py.g1_mul(g, 77)) == pr.pairing(h, g) ** 77


def _prove_setup(log_n, seed, srs_head):
    circ = synth.synth_circuit(log_n, num_inputs=4, seed=seed, check=True)
    n = 1 << log_n
    srs = np.frombuffer(srs_head[80:80 + 64 * (n + 3)], dtype=np.uint64).reshape(n + 3, 8).copy()
    return circ, srs


def test_oracle_proof_verifies_with_pairing(oracle, srs_head, g2):
    """Oracle prover on the reference's real SRS powers -> accepted by the pairing check; a tampered
    evaluation or public input is rejected."""
    log_n = 7
    circ, srs = _prove_setup(log_n, 31, srs_head)
    pk = oracle.plonk_preprocess(log_n, circ.selectors, circ.perm, circ.k, srs)
    rc, proof, _, _ = oracle.plonk_prove(log_n, circ.num_inputs, circ.k, pk, circ.wires, circ.pub_inputs,
                                         synth.splitmix_blinders(8), srs)
    assert rc == 0
    assert oracle.plonk_verify_pairing(log_n, circ.num_inputs, circ.k, pk, circ.pub_inputs, proof, *g2)
    bad = oracle.PlonkProof.from_buffer_copy(bytes(proof))
    bad.wire_sigma_evals[1][0] ^= 1
    assert not oracle.plonk_verify_pairing(log_n, circ.num_inputs, circ.k, pk, circ.pub_inputs, bad, *g2)
    pi2 = circ.pub_inputs.copy()
    pi2[2, 0] ^= np.uint64(1)
    assert not oracle.plonk_verify_pairing(log_n, circ.num_inputs, circ.k, pk, pi2, proof, *g2)


@pytest.mark.gpu
def test_gpu_proof_on_reference_srs_verifies_with_pairing(ctx, oracle, srs_head, g2):
    """Device prover on the reference's real SRS powers (loaded through the ptau parser) -> the proof
    passes the pairing check, and equals the oracle prover's proof."""
    import renegade_b200 as rb
    from renegade_b200.backend import PlonkKzgSnark
    log_n = 8
    circ, srs = _prove_setup(log_n, 32, srs_head)
    params = rb.parse_ptau_file(ctx, srs_head, count=(1 << log_n) + 3)
    pk = PlonkKzgSnark.preprocess(ctx, params.powers_of_g, log_n, circ.num_inputs, circ.selectors, circ.perm, circ.k)
    bl = synth.splitmix_blinders(9)
    proof, _ = PlonkKzgSnark.prove_with_link_hint(ctx, pk, circ.wires, circ.pub_inputs, bl)
    opk = {"selector_comms": pk.selector_comms, "sigma_comms": pk.sigma_comms}
    op = oracle.PlonkProof.from_buffer_copy(bytes(proof))
    assert oracle.plonk_verify_pairing(log_n, circ.num_inputs, circ.k, opk, circ.pub_inputs, op, *g2)
    full = oracle.plonk_preprocess(log_n, circ.selectors, circ.perm, circ.k, srs)
    rc, oproof, _, _ = oracle.plonk_prove(log_n, circ.num_inputs, circ.k, full, circ.wires, circ.pub_inputs, bl, srs)
    assert rc == 0 and (proof.to_array() == oproof.to_array()).all()


def _link_pairing_ok(py, pr, comm1, comm2, layout, lp_arr, eta_mont, g2):
    """verify_link_proof with a pairing: e(pi, tau H) == e(C1 - C2 - Z_D(eta) Cq + eta pi, H)."""
    h, tau_h = g2
    alignment, offset, size = layout
    eta = py.from_mont(int.from_bytes(eta_mont.tobytes(), "little"), py.R)
    g = py.domain_generator(alignment)
    zd = 1
    for i in range(size):
        zd = zd * (eta - pow(g, offset + i, py.R)) % py.R
    dec = lambda arr: py.decode_g1_mont(np.ascontiguousarray(arr, dtype=np.uint64).tobytes(), 0)
    c1, c2, cq, pi = dec(comm1), dec(comm2), dec(lp_arr[:8]), dec(lp_arr[8:])
    b = py.g1_add(c1, py.g1_neg(c2))
    b = py.g1_add(b, py.g1_neg(py.g1_mul(cq, zd)))
    b = py.g1_add(b, py.g1_mul(pi, eta))
    return pr.pairing_product_is_one([(pi, tau_h), (py.g1_neg(b), h)])


def test_oracle_link_proof_verifies_with_pairing(oracle, pyoracle, srs_head, g2):
    import bn254_pairing_py as pr
    py = pyoracle
    layout = (5, 6, 5)  # alignment, offset, size
    vals = [(i * 1234567 + 89) % py.R for i in range(layout[2])]
    hints = []
    for log_n, seed in ((6, 41), (7, 42)):
        n = 1 << log_n
        circ = synth.synth_circuit(log_n, num_inputs=3, seed=seed, check=True, link=(layout[0], layout[1], vals))
        srs = np.frombuffer(srs_head[80:80 + 64 * (n + 3)], dtype=np.uint64).reshape(n + 3, 8).copy()
        pk = oracle.plonk_preprocess(log_n, circ.selectors, circ.perm, circ.k, srs)
        rc, proof, _, link = oracle.plonk_prove(log_n, circ.num_inputs, circ.k, pk, circ.wires, circ.pub_inputs,
                                                synth.splitmix_blinders(seed), srs, True)
        assert rc == 0
        hints.append((link, np.array(proof.wires_poly_comms[0], dtype=np.uint64)))
    srs_all = np.frombuffer(srs_head[80:], dtype=np.uint64).reshape(-1, 8)
    rc, lp, eta = oracle.plonk_link(hints[0][0], hints[1][0], hints[0][1], hints[1][1], *layout, srs_all)
    assert rc == 0
    assert _link_pairing_ok(py, pr, hints[0][1], hints[1][1], layout, lp.to_array(), eta, g2)
    wrong = (layout[0], layout[1] + 1, layout[2])
    rc, _, _ = oracle.plonk_link(hints[0][0], hints[1][0], hints[0][1], hints[1][1], *wrong, srs_all)
    assert rc == 2   # not linkable on that group: the prover refuses instead of emitting an unverifiable proof
    assert not _link_pairing_ok(py, pr, hints[0][1], hints[1][1], wrong, lp.to_array(), eta, g2)


def test_hard_part_addition_chain_exponent():
    """The final exponentiation of renegade_b200/csrc/verify.cu takes the hard part by the x-addition chain of
    Fuentes-Castaneda et al. (the schedule of `final_exponentiation` there, restated here on EXPONENTS: a product adds,
    a squaring doubles, a conjugation — the inverse in the cyclotomic subgroup — negates, a q^k-Frobenius multiplies by
    q^k, all modulo the subgroup order q^4 - q^2 + 1).  The chain must raise to c (q^4 - q^2 + 1) / r with c prime to r:
    then its result is one exactly when the reduced pairing is."""
    q = 0x30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47
    r = 0x30644e72e131a029b85045b68181585d2833e84879b97091_43e1f593f0000001
    x = 4965661367192848881
    assert x == 0x44e992b44a6909f1            # the constant of fq12_pow_x
    phi = q ** 4 - q ** 2 + 1
    assert phi % r == 0
    h = phi // r
    neg_x = lambda e: -x * e % phi
    frob = lambda e, k: e * pow(q, k, phi) % phi
    t = 1
    y0 = neg_x(t)
    y1 = 2 * y0 % phi
    y2 = 2 * y1 % phi
    y3 = (y2 + y1) % phi
    y4 = neg_x(y3)
    y5 = 2 * y4 % phi
    y6 = neg_x(y5)
    y3, y6 = -y3 % phi, -y6 % phi
    y7 = (y6 + y4) % phi
    y8 = (y7 + y3) % phi
    y9 = (y8 + y1) % phi
    y10 = (y8 + y4) % phi
    y11 = (y10 + t) % phi
    y13 = (frob(y9, 1) + y11) % phi
    y14 = (frob(y8, 2) + y13) % phi
    y15 = (-t + y9) % phi
    e = (frob(y15, 3) + y14) % phi
    assert e % h == 0 and (e // h) % r != 0
