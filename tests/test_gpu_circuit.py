"""GPU: a circuit with the reference's own gate structure — a height-10 Poseidon2 Merkle opening
(`MERKLE_HEIGHT`, crates/constants/src/lib.rs:50; gadgets of circuits-core restated in examples/host_circuits/circuit.py, public
root like the VALID-* statements) — proved on the device: proof bytes identical to the oracle prover's, accepted by the
restated verifier, wrong root rejected."""
import random

import pytest

from host_circuits import circuit as C
from renegade_b200 import synth
from renegade_b200.backend import PlonkKzgSnark

pytestmark = pytest.mark.gpu

TAU = 0x0f1e2d3c4b5a69788796a5b4c3d2e1f00123456789abcdef0fedcba987654321


def test_merkle_opening_circuit_proves_on_device(ctx, oracle, pyoracle):
    py = pyoracle
    rnd = random.Random(0x3E7)
    height = 10
    leaf = [rnd.randrange(C.R) for _ in range(4)]
    opening = C.MerkleOpening([rnd.randrange(C.R) for _ in range(height)], [rnd.random() < 0.5 for _ in range(height)])
    cs, root = C.merkle_membership_circuit(leaf, opening)
    cs.check_circuit_satisfiability([root])
    circ = cs.finalize_for_arithmetization()
    assert circ.log_n == 12 and circ.num_inputs == 1     # 11 hashes x (195 + a few) gates
    tau = oracle.int_to_limbs(py.to_mont(TAU % py.R, py.R))
    srs = oracle.srs_from_tau(tau, circ.n + 3)
    bases = ctx.load_bases(srs)
    pk = PlonkKzgSnark.preprocess(ctx, bases, circ.log_n, circ.num_inputs, circ.selectors, circ.perm, circ.k)
    opk = oracle.plonk_preprocess(circ.log_n, circ.selectors, circ.perm, circ.k, srs)
    assert (pk.selector_comms == opk["selector_comms"]).all() and (pk.sigma_comms == opk["sigma_comms"]).all()
    bl = synth.splitmix_blinders(0x10)
    proof, _ = PlonkKzgSnark.prove_with_link_hint(ctx, pk, circ.wires, circ.pub_inputs, bl)
    rc, oproof, _, _ = oracle.plonk_prove(circ.log_n, circ.num_inputs, circ.k, opk, circ.wires, circ.pub_inputs, bl, srs)
    assert rc == 0 and (proof.to_array() == oproof.to_array()).all()
    assert oracle.plonk_verify_known_tau(circ.log_n, circ.num_inputs, circ.k, opk, circ.pub_inputs,
                                         oracle.PlonkProof.from_buffer_copy(bytes(proof)), tau)
    wrong = synth.to_mont_array([(root + 1) % C.R])
    assert not oracle.plonk_verify_known_tau(circ.log_n, circ.num_inputs, circ.k, opk, wrong,
                                             oracle.PlonkProof.from_buffer_copy(bytes(proof)), tau)


def test_valid_balance_create_proves_on_device(ctx, oracle, pyoracle):
    """BASELINE.json configs[0]: the VALID BALANCE CREATE circuit (restated in examples/host_circuits/valid_balance_create.py,
    n = 2^13, 13 public inputs) — device proof bytes = oracle proof bytes, verifier accepts."""
    from host_circuits import valid_balance_create as vbc
    py = pyoracle
    witness, statement = vbc.create_witness_statement(seed=0xB200)
    cs = vbc.ValidBalanceCreate.build(witness, statement)
    cs.check_circuit_satisfiability(statement.to_scalars())
    circ = cs.finalize_for_arithmetization()
    assert circ.log_n == 13 and circ.num_inputs == 13
    tau = oracle.int_to_limbs(py.to_mont(TAU % py.R, py.R))
    srs = oracle.srs_from_tau(tau, circ.n + 3)
    bases = ctx.load_bases(srs)
    pk = PlonkKzgSnark.preprocess(ctx, bases, circ.log_n, circ.num_inputs, circ.selectors, circ.perm, circ.k)
    opk = oracle.plonk_preprocess(circ.log_n, circ.selectors, circ.perm, circ.k, srs)
    assert (pk.selector_comms == opk["selector_comms"]).all() and (pk.sigma_comms == opk["sigma_comms"]).all()
    bl = synth.splitmix_blinders(0x11)
    proof, _ = PlonkKzgSnark.prove_with_link_hint(ctx, pk, circ.wires, circ.pub_inputs, bl)
    rc, oproof, _, _ = oracle.plonk_prove(circ.log_n, circ.num_inputs, circ.k, opk, circ.wires, circ.pub_inputs, bl, srs)
    assert rc == 0 and (proof.to_array() == oproof.to_array()).all()
    assert oracle.plonk_verify_known_tau(circ.log_n, circ.num_inputs, circ.k, opk, circ.pub_inputs,
                                         oracle.PlonkProof.from_buffer_copy(bytes(proof)), tau)


def test_private_settlement_proves_on_device(ctx, oracle, pyoracle):
    """BASELINE.json configs[3]'s statement at its own size: INTENT AND BALANCE PRIVATE SETTLEMENT restated
    (examples/host_circuits/private_settlement.py; 17 public inputs, four link groups, n = 2^12) — device proof bytes and
    linking hint = the oracle's, verifier accepts."""
    from host_circuits import private_settlement as ps
    py = pyoracle
    parties, statement = ps.create_witness_statement(seed=0xB200)
    cs = ps.IntentAndBalancePrivateSettlementCircuit.build(parties, statement)
    cs.check_circuit_satisfiability(statement.to_scalars())
    circ = cs.finalize_for_arithmetization()
    assert circ.log_n == 12 and circ.num_inputs == 17
    tau = oracle.int_to_limbs(py.to_mont(TAU % py.R, py.R))
    srs = oracle.srs_from_tau(tau, circ.n + 3)
    bases = ctx.load_bases(srs)
    pk = PlonkKzgSnark.preprocess(ctx, bases, circ.log_n, circ.num_inputs, circ.selectors, circ.perm, circ.k)
    opk = oracle.plonk_preprocess(circ.log_n, circ.selectors, circ.perm, circ.k, srs)
    assert (pk.selector_comms == opk["selector_comms"]).all() and (pk.sigma_comms == opk["sigma_comms"]).all()
    bl = synth.splitmix_blinders(0x12)
    proof, hint = PlonkKzgSnark.prove_with_link_hint(ctx, pk, circ.wires, circ.pub_inputs, bl)
    rc, oproof, _, olink = oracle.plonk_prove(circ.log_n, circ.num_inputs, circ.k, opk, circ.wires, circ.pub_inputs, bl, srs, True)
    assert rc == 0 and (proof.to_array() == oproof.to_array()).all()
    assert (hint.linking_wire_poly == olink).all()
    assert oracle.plonk_verify_known_tau(circ.log_n, circ.num_inputs, circ.k, opk, circ.pub_inputs,
                                         oracle.PlonkProof.from_buffer_copy(bytes(proof)), tau)


def test_private_match_bundle_on_device(ctx, oracle, pyoracle):
    """The bundle the reference proves for a private match (native_proof_manager.rs:526-584, 726-782; SURVEY.md §8(d)
    config 4), with the restated circuits: both parties' INTENT AND BALANCE VALIDITY (n = 2^14) and OUTPUT BALANCE
    VALIDITY (n = 2^13) proofs, the PRIVATE SETTLEMENT proof (n = 2^12) and the four link proofs — all on the device,
    each byte-identical to the oracle's, every proof and link accepted by the restated verifiers."""
    from host_circuits import intent_and_balance_validity as val
    from host_circuits import output_balance_validity as obv
    from host_circuits import private_settlement as ps
    from renegade_b200.backend import GroupLayout, link_proofs
    py = pyoracle
    parties, _ = ps.create_witness_statement(seed=41)
    validity = [val.create_witness_statement(seed=50 + i, intent=parties[i].intent, balance=parties[i].input_balance)
                for i in (0, 1)]
    out_validity = [obv.create_witness_statement(60 + i, parties[i].output_balance) for i in (0, 1)]
    parties, statement = ps.create_witness_statement(
        seed=41, linked=[(validity[i][0].new_amount_public_share, validity[i][0].post_match_balance_shares,
                          out_validity[i][0].post_match_balance_shares) for i in (0, 1)])
    settlement_cs = ps.IntentAndBalancePrivateSettlementCircuit.build(parties, statement)
    layouts = settlement_cs.get_circuit_layout()
    circs = [settlement_cs.finalize_for_arithmetization()] + \
            [val.IntentAndBalanceValidityCircuit.build(w, s, layouts).finalize_for_arithmetization() for w, s in validity] + \
            [obv.OutputBalanceValidityCircuit.build(w, s, layouts).finalize_for_arithmetization() for w, s in out_validity]
    assert [c.log_n for c in circs] == [12, 14, 14, 13, 13]
    tau = oracle.int_to_limbs(py.to_mont(TAU % py.R, py.R))
    srs = oracle.srs_from_tau(tau, (1 << 14) + 3)
    bases = ctx.load_bases(srs)
    hints, ohints = [], []
    for i, circ in enumerate(circs):
        pk = PlonkKzgSnark.preprocess(ctx, bases, circ.log_n, circ.num_inputs, circ.selectors, circ.perm, circ.k)
        opk = oracle.plonk_preprocess(circ.log_n, circ.selectors, circ.perm, circ.k, srs[:circ.n + 3])
        bl = synth.splitmix_blinders(0x20 + i)
        proof, hint = PlonkKzgSnark.prove_with_link_hint(ctx, pk, circ.wires, circ.pub_inputs, bl)
        rc, oproof, _, olink = oracle.plonk_prove(circ.log_n, circ.num_inputs, circ.k, opk, circ.wires, circ.pub_inputs, bl,
                                                  srs[:circ.n + 3], True)
        assert rc == 0 and (proof.to_array() == oproof.to_array()).all() and (hint.linking_wire_poly == olink).all()
        assert oracle.plonk_verify_known_tau(circ.log_n, circ.num_inputs, circ.k, opk, circ.pub_inputs,
                                             oracle.PlonkProof.from_buffer_copy(bytes(proof)), tau)
        hints.append(hint)
        pk.free()
    for lay, v in [(layouts[ps.PARTY_LINKS[p]], hints[1 + p]) for p in (0, 1)] + \
                  [(layouts[ps.OUTPUT_LINKS[p]], hints[3 + p]) for p in (0, 1)]:
        s = hints[0]
        lp, eta = link_proofs(ctx, bases, v, s, GroupLayout(lay.alignment, lay.offset, lay.size))
        rc, olp, oeta = oracle.plonk_link(v.linking_wire_poly, s.linking_wire_poly, v.linking_wire_comm, s.linking_wire_comm,
                                          lay.alignment, lay.offset, lay.size, srs)
        assert rc == 0 and (lp.to_array() == olp.to_array()).all() and (eta == oeta).all()
        assert oracle.plonk_link_verify_known_tau(v.linking_wire_comm, s.linking_wire_comm, lay.alignment, lay.offset, lay.size,
                                                  oracle.LinkProof.from_buffer_copy(bytes(lp)), tau)
