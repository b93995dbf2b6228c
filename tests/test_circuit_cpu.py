"""CPU: the host-side constraint system and the reference's Poseidon2 / Merkle gadgets restated on it
(examples/host_circuits/circuit.py).  The native hash is checked against the oracle's Poseidon2 (itself pinned by the published
HorizenLabs known answer with the reference's constants, tests/test_poseidon2.py); the gadget against the native hash
and against the gate counts the reference documents; the arithmetization against the oracle prover and verifier:
a circuit built here proves and verifies on the CPU restatement of the reference algorithm, which is what the device
prover is bit-exact with."""
import os
import random

import numpy as np
import pytest

from host_circuits import circuit as C
from renegade_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TAU = 0x19d2c3b4a5968778695a4b3c2d1e0f00112233445566778899aabbccddeeff01


@pytest.fixture(scope="module")
def p2(pyoracle):
    return pyoracle.poseidon2_load_constants(os.path.join(ROOT, "tests", "golden", "poseidon2.json"))


def test_native_sponge_matches_oracle(pyoracle, p2):
    full, partial = p2
    assert [c for row in C.FULL_ROUND_CONSTANTS for c in row] == full and C.PARTIAL_ROUND_CONSTANTS == partial
    rnd = random.Random(5)
    for ln in range(0, 7):
        xs = [rnd.randrange(C.R) for _ in range(ln)]
        assert C.compute_poseidon_hash(xs) == pyoracle.poseidon2_hash(xs, full, partial), ln
    sp = C.Poseidon2Sponge()
    sp.state = [0, 1, 2]
    sp.permute()
    assert sp.state == pyoracle.poseidon2_permute([0, 1, 2], full, partial)
    # squeezing more than the rate permutes again (poseidon2.rs:67-79)
    sp = C.Poseidon2Sponge()
    sp.absorb_batch([7, 8, 9])
    a, b, c3 = sp.squeeze(), sp.squeeze(), sp.squeeze()
    st = pyoracle.poseidon2_permute(pyoracle.poseidon2_permute([0, 7, 8], full, partial)[:1] +
                                    [(pyoracle.poseidon2_permute([0, 7, 8], full, partial)[1] + 9) % C.R,
                                     pyoracle.poseidon2_permute([0, 7, 8], full, partial)[2]], full, partial)
    assert (a, b) == (st[1], st[2]) and c3 == pyoracle.poseidon2_permute(st, full, partial)[1]


def test_hash_gadget_matches_native_and_gate_count():
    rnd = random.Random(6)
    for ln in (1, 2, 3, 5):
        cs = C.PlonkCircuit()
        base = cs.num_gates                      # the two constant gates pinning zero and one
        xs = [rnd.randrange(C.R) for _ in range(ln)]
        vars_ = [cs.create_variable(x) for x in xs]
        out = C.PoseidonHashGadget(cs.zero()).hash(vars_, cs)
        assert cs.witness(out) == C.compute_poseidon_hash(xs)
        perms = (ln + 1) // 2                    # one permutation per full rate block, the last one at the squeeze
        assert cs.num_gates - base == ln + 195 * perms   # one addition gate per absorbed element; hash.rs:205 -> 195
        cs.check_circuit_satisfiability([])
        # expected output as a public input, constrained squeeze
        cs2 = C.PlonkCircuit()
        exp = cs2.create_public_variable(C.compute_poseidon_hash(xs))
        C.PoseidonHashGadget(cs2.zero()).hash_constrained([cs2.create_variable(x) for x in xs], exp, cs2)
        cs2.check_circuit_satisfiability(cs2.public_input())
        with pytest.raises(C.CircuitError):
            cs2.check_circuit_satisfiability([(cs2.public_input()[0] + 1) % C.R])


def test_basic_gates_and_errors():
    cs = C.PlonkCircuit()
    a, b = cs.create_variable(11), cs.create_variable(C.R - 4)
    assert cs.witness(cs.add(a, b)) == 7 and cs.witness(cs.sub(a, b)) == 15 and cs.witness(cs.mul(a, b)) == C.R - 44
    assert cs.witness(cs.lc([a, b, cs.one(), cs.zero()], [2, 3, 5, 9])) == (22 - 12 + 5) % C.R
    t, f = cs.create_boolean_variable(True), cs.create_boolean_variable(False)
    assert cs.witness(cs.mux(t, a, b)) == 11 and cs.witness(cs.mux(f, a, b)) == C.R - 4
    k = cs.create_constant_variable(99)
    cs.enforce_equal(k, cs.add(cs.create_variable(90), cs.create_variable(9)))
    cs.check_circuit_satisfiability([])
    cs.witness_values[k] = 98                    # breaks its constant gate and the equality
    with pytest.raises(C.CircuitError):
        cs.check_circuit_satisfiability([])
    with pytest.raises(C.CircuitError):
        cs.insert_gate([0, 0, 0], C.AdditionGate())
    with pytest.raises(C.CircuitError):
        cs.witness(10 ** 9)
    bad = C.PlonkCircuit()
    nb = bad.create_variable(2)
    bad.enforce_bool(nb)
    with pytest.raises(C.CircuitError):
        bad.check_circuit_satisfiability([])


def merkle_case(height, leaf_len, seed):
    rnd = random.Random(seed)
    leaf = [rnd.randrange(C.R) for _ in range(leaf_len)]
    opening = C.MerkleOpening([rnd.randrange(C.R) for _ in range(height)], [rnd.random() < 0.5 for _ in range(height)])
    return leaf, opening


def test_merkle_gadget_and_arithmetization(oracle, pyoracle):
    py = pyoracle
    leaf, opening = merkle_case(3, 3, seed=9)
    cs, root = C.merkle_membership_circuit(leaf, opening)
    assert root == C.native_merkle_root(leaf, opening)
    cs.check_circuit_satisfiability([root])
    with pytest.raises(C.CircuitError):
        cs.check_circuit_satisfiability([(root + 1) % C.R])
    circ = cs.finalize_for_arithmetization()
    with pytest.raises(C.CircuitError):
        cs.create_variable(1)                    # closed after finalisation, like upstream
    n = circ.n
    assert circ.num_inputs == 1 and circ.n_gates <= n < 2 * circ.n_gates
    # IO gate first, every row satisfies the gate equation, the permutation is a bijection that preserves values
    assert circ.selectors_int[synth.Q_O][0] == 1 and circ.wires_int[4][0] == root
    for row in range(n):
        pi = circ.pub_inputs_int[row] if row < circ.num_inputs else 0
        assert synth.gate_value([circ.selectors_int[s][row] for s in range(13)],
                                [circ.wires_int[w][row] for w in range(5)], pi) == 0, row
    assert sorted(circ.perm.tolist()) == list(range(5 * n))
    flat = [v for w in circ.wires_int for v in w]
    assert all(flat[i] == flat[int(circ.perm[i])] for i in range(5 * n))
    # the oracle prover proves it and the oracle verifier accepts; a wrong root is rejected
    tau = oracle.int_to_limbs(py.to_mont(TAU % py.R, py.R))
    srs = oracle.srs_from_tau(tau, n + 3)
    opk = oracle.plonk_preprocess(circ.log_n, circ.selectors, circ.perm, circ.k, srs)
    bl = synth.splitmix_blinders(0xC1C)
    rc, proof, _, _ = oracle.plonk_prove(circ.log_n, circ.num_inputs, circ.k, opk, circ.wires, circ.pub_inputs, bl, srs)
    assert rc == 0
    assert oracle.plonk_verify_known_tau(circ.log_n, circ.num_inputs, circ.k, opk, circ.pub_inputs, proof, tau)
    wrong = synth.to_mont_array([(root + 1) % C.R])
    assert not oracle.plonk_verify_known_tau(circ.log_n, circ.num_inputs, circ.k, opk, wrong, proof, tau)
    # a witness that breaks one gate is refused by the prover (WrongQuotientPolyDegree)
    bad = circ.wires.copy()
    bad[4, circ.n_gates - 2] = bad[4, circ.n_gates - 3]
    rc, _, _, _ = oracle.plonk_prove(circ.log_n, circ.num_inputs, circ.k, opk, bad, circ.pub_inputs, bl, srs)
    assert rc != 0


def test_field_helpers(pyoracle, oracle):
    """crypto::fields (fields.rs:21-163) restated: moduli, big-endian truncations, signed conversion, and the ABI's
    Montgomery limbs against the oracle's own conversion."""
    from renegade_b200 import fields as F
    py = pyoracle
    assert F.get_scalar_field_modulus() == py.R and F.get_base_field_modulus() == py.Q
    v = 0x1234_5678_9ABC_DEF0_0FED_CBA9_8765_4321_1122_3344_5566_7788_99AA_BBCC_DDEE_FF00 % py.R
    be = F.scalar_to_bytes_be(v)
    assert len(be) == 32 and int.from_bytes(be, "big") == v
    assert F.scalar_to_u64(v) == v & ((1 << 64) - 1) and F.scalar_to_u128(v) == v & ((1 << 128) - 1)
    assert F.scalar_to_address(v) == be[12:] and F.address_to_scalar(be[12:]) == v & ((1 << 160) - 1)
    assert F.scalar_to_u256(v) == v and F.u256_to_scalar((1 << 256) - 1) == ((1 << 256) - 1) % py.R
    assert F.bigint_to_scalar(-5) == py.R - 5 and F.bigint_to_scalar(5) == 5 and F.bigint_to_scalar(-py.R) == 0
    assert F.biguint_to_scalar(py.R + 3) == 3
    with pytest.raises(ValueError):
        F.biguint_to_scalar(-1)
    assert F.bigint_to_scalar_bits(0b1011, 6) == [1, 1, 0, 1, 0, 0]
    vals = [0, 1, py.R - 1, v]
    limbs = F.scalars_to_limbs(vals)
    assert (limbs == oracle.ints_to_array([py.to_mont(x, py.R) for x in vals])).all()
    assert F.limbs_to_scalars(limbs) == vals
    assert (F.scalars_to_limbs(vals) == synth.to_mont_array(vals)).all()
    q = [2, py.Q - 1]
    assert F.limbs_to_scalars(F.scalars_to_limbs(q, py.Q), py.Q) == q


def test_link_groups_place_shared_values_and_link_on_the_oracle(oracle, pyoracle):
    """Two circuits of different sizes put the same witness values into a link group with the same layout
    (`create_link_group` / `create_variable_with_link_groups`, the way the reference's validity and settlement
    circuits share a balance or an intent); `get_circuit_layout` reports it; the oracle proves both, links the two
    hints (`PlonkKzgSnark::link_proofs`) and its link verifier accepts — and rejects when one circuit holds a
    different value."""
    py = pyoracle
    layout = C.GroupLayout(alignment=6, offset=10)
    shared = [(i * 0x9E3779B97F4A7C15 + 7) % C.R for i in range(5)]

    def build(n_hashes, values):
        cs = C.PlonkCircuit()
        cs.create_link_group("shared_state", layout)
        vars_ = [cs.create_variable_with_link_groups(v, ["shared_state"]) for v in values]
        out = cs.create_public_variable(0)      # placeholder value, fixed below
        acc = vars_
        for _ in range(n_hashes):               # some constraints that consume the shared values
            acc = [C.PoseidonHashGadget(cs.zero()).hash(acc, cs)] + vars_[1:]
        cs.witness_values[out] = cs.witness(acc[0])
        cs.enforce_equal(out, acc[0])
        cs.check_circuit_satisfiability(cs.public_input())
        lay = cs.get_circuit_layout()["shared_state"]
        assert (lay.alignment, lay.offset, lay.size) == (6, 10, 5)
        return cs.finalize_for_arithmetization()

    tau = oracle.int_to_limbs(py.to_mont(TAU % py.R, py.R))
    circs = [build(1, shared), build(3, shared), build(1, shared[:4] + [shared[4] + 1])]
    assert circs[0].log_n != circs[1].log_n
    srs = oracle.srs_from_tau(tau, max(c.n for c in circs) + 3)
    hints = []
    for i, circ in enumerate(circs):
        for j, v in enumerate(shared if i < 2 else shared[:4] + [(shared[4] + 1) % C.R]):
            assert circ.wires_int[0][(layout.offset + j) << (circ.log_n - layout.alignment)] == v
        opk = oracle.plonk_preprocess(circ.log_n, circ.selectors, circ.perm, circ.k, srs[:circ.n + 3])
        rc, proof, _, link = oracle.plonk_prove(circ.log_n, circ.num_inputs, circ.k, opk, circ.wires, circ.pub_inputs,
                                                synth.splitmix_blinders(60 + i), srs[:circ.n + 3], True)
        assert rc == 0
        assert oracle.plonk_verify_known_tau(circ.log_n, circ.num_inputs, circ.k, opk, circ.pub_inputs, proof, tau)
        hints.append((link, proof.to_array()[:8].copy()))
    rc, lp, _ = oracle.plonk_link(hints[0][0], hints[1][0], hints[0][1], hints[1][1], layout.alignment,
                                  layout.offset, len(shared), srs)
    assert rc == 0
    assert oracle.plonk_link_verify_known_tau(hints[0][1], hints[1][1], layout.alignment, layout.offset, len(shared), lp, tau)
    # circuit 2 holds a different value in the group: the prover refuses (a division is not exact) and the
    # honest proof of the (0, 1) pair does not verify against circuit 2's commitment
    rc, _, _ = oracle.plonk_link(hints[0][0], hints[2][0], hints[0][1], hints[2][1], layout.alignment, layout.offset,
                                 len(shared), srs)
    assert rc == 2
    assert not oracle.plonk_link_verify_known_tau(hints[0][1], hints[2][1], layout.alignment, layout.offset,
                                                  len(shared), lp, tau)
    # a layout that collides with the public-input rows is refused
    bad = C.PlonkCircuit()
    bad.create_link_group("g", C.GroupLayout(alignment=3, offset=0))
    bad.create_variable_with_link_groups(5, ["g"])
    bad.create_public_variable(1)
    with pytest.raises(C.CircuitError):
        bad.finalize_for_arithmetization()


def test_bits_and_comparator_gadgets():
    """bits.rs / comparators.rs restated: decomposition, range checks, is-zero, a >= b — satisfiable exactly when the
    statement holds."""
    def sat(build):
        cs = C.PlonkCircuit()
        build(cs)
        try:
            cs.check_circuit_satisfiability([])
            return True
        except C.CircuitError:
            return False

    cs = C.PlonkCircuit()
    v = cs.create_variable(0xDEADBEEF)
    bits = C.ToBitsGadget.to_bits(v, 32, cs)
    assert [cs.witness(b) for b in bits] == C.scalar_to_bits_le(0xDEADBEEF, 32)
    assert cs.witness(C.ToBitsGadget.bit_reconstruct(bits, cs)) == 0xDEADBEEF
    assert cs.witness(cs.lc_sum([v] * 9, list(range(1, 10)))) == 45 * 0xDEADBEEF % C.R
    assert cs.witness(cs.sum([])) == 0
    cs.check_circuit_satisfiability([])
    assert sat(lambda c: C.BitRangeGadget.constrain_bit_range(c.create_variable(2 ** 64 - 1), 64, c))
    assert not sat(lambda c: C.BitRangeGadget.constrain_bit_range(c.create_variable(2 ** 64), 64, c))
    assert not sat(lambda c: C.BitRangeGadget.constrain_bit_range(c.create_variable(C.R - 1), 64, c))
    for val in (0, 5, C.R - 1):
        cs = C.PlonkCircuit()
        z = C.EqZeroGadget.eq_zero_var(cs.create_variable(val), cs)
        assert cs.witness(z) == int(val == 0)
        cs.check_circuit_satisfiability([])
        cs.witness_values[z] ^= 1            # the flag cannot be flipped
        with pytest.raises(C.CircuitError):
            cs.check_circuit_satisfiability([])
    for a, b in ((7, 7), (100, 3), (3, 100), (0, 2 ** 63)):
        cs = C.PlonkCircuit()
        ge = C.GreaterThanEqGadget.greater_than_eq(cs.create_variable(a), cs.create_variable(b), 64, cs)
        assert cs.witness(ge) == int(a >= b)
        cs.check_circuit_satisfiability([])
        assert sat(lambda c: C.GreaterThanEqGadget.constrain_greater_than_eq(c.create_variable(a), c.create_variable(b), 64, c)) == (a >= b)


def test_valid_balance_create_circuit(oracle, pyoracle):
    """BASELINE.json configs[0]: VALID BALANCE CREATE restated (examples/host_circuits/valid_balance_create.py).  The native
    witness / statement satisfy the circuit; every statement field is binding; the domain is 2^13; the oracle
    prover proves it and the oracle verifier accepts."""
    from host_circuits import valid_balance_create as vbc
    py = pyoracle
    witness, statement = vbc.create_witness_statement(seed=7)
    assert statement.recovery_id == witness.initial_recovery_stream.get_ith(witness.initial_recovery_stream.index)
    assert witness.initial_share_stream.index == 8 and len(statement.to_scalars()) == 13
    cs = vbc.ValidBalanceCreate.build(witness, statement)
    pub = statement.to_scalars()
    assert cs.public_input() == pub
    cs.check_circuit_satisfiability(pub)
    for i in range(len(pub)):                     # no statement field is free
        bad = list(pub)
        bad[i] = (bad[i] + 1) % C.R
        with pytest.raises(C.CircuitError):
            cs.check_circuit_satisfiability(bad)
    # a witness whose fee balance is not zero does not satisfy the circuit even with a consistent statement
    w2, s2 = vbc.create_witness_statement(seed=8)
    w2.balance.relayer_fee_balance = 5
    with pytest.raises(C.CircuitError):
        vbc.ValidBalanceCreate.build(w2, s2).check_circuit_satisfiability(s2.to_scalars())
    circ = cs.finalize_for_arithmetization()
    assert circ.log_n == 13 and circ.num_inputs == 13 and 4500 < circ.n_gates < 5200
    tau = oracle.int_to_limbs(py.to_mont(TAU % py.R, py.R))
    srs = oracle.srs_from_tau(tau, circ.n + 3)
    opk = oracle.plonk_preprocess(circ.log_n, circ.selectors, circ.perm, circ.k, srs)
    rc, proof, _, _ = oracle.plonk_prove(circ.log_n, circ.num_inputs, circ.k, opk, circ.wires, circ.pub_inputs,
                                         synth.splitmix_blinders(0xBA1), srs)
    assert rc == 0
    assert oracle.plonk_verify_known_tau(circ.log_n, circ.num_inputs, circ.k, opk, circ.pub_inputs, proof, tau)


def test_private_settlement_circuit(oracle, pyoracle):
    """BASELINE.json configs[3]'s statement, INTENT AND BALANCE PRIVATE SETTLEMENT, restated
    (examples/host_circuits/private_settlement.py): a consistent two-party match satisfies it, each rule it encodes is binding,
    it has 17 public inputs and four link groups, and it is a 2^12-gate circuit; the oracle proves and verifies it."""
    from host_circuits import private_settlement as ps
    py = pyoracle
    parties, statement = ps.create_witness_statement(seed=11)
    build = ps.IntentAndBalancePrivateSettlementCircuit.build
    cs = build(parties, statement)
    pub = statement.to_scalars()
    assert len(pub) == 17 and cs.public_input() == pub
    cs.check_circuit_satisfiability(pub)
    assert {g: l.size for g, l in cs.get_circuit_layout().items()} == {
        ps.PARTY_LINKS[0]: 17, ps.PARTY_LINKS[1]: 17, ps.OUTPUT_LINKS[0]: 11, ps.OUTPUT_LINKS[1]: 11}
    for i in range(14):                                  # share updates are pinned by the statement
        bad = list(pub)
        bad[i] = (bad[i] + 1) % C.R
        with pytest.raises(C.CircuitError):
            cs.check_circuit_satisfiability(bad)

    def broken(mutate):
        p, s = ps.create_witness_statement(seed=11)
        mutate(p, s)
        try:
            build(p, s).check_circuit_satisfiability(s.to_scalars())
            return False
        except C.CircuitError:
            return True
    assert broken(lambda p, s: setattr(p[0].settlement_obligation, "amount_out", p[0].settlement_obligation.amount_out + 1))
    assert broken(lambda p, s: setattr(p[1].intent, "amount_in", p[1].settlement_obligation.amount_in - 1))   # overfill
    assert broken(lambda p, s: setattr(p[0].intent, "min_price", p[0].intent.min_price * 4))                  # bad price
    assert broken(lambda p, s: setattr(p[0].input_balance, "amount", p[0].settlement_obligation.amount_in - 1))
    assert broken(lambda p, s: setattr(p[1].output_balance, "owner", p[1].output_balance.owner ^ 1))
    assert broken(lambda p, s: setattr(p[0].output_balance, "amount", (1 << 100) - 1))                        # overflow
    assert broken(lambda p, s: setattr(s, "protocol_fee", s.protocol_fee + (1 << 40)))                       # fee take changes

    circ = cs.finalize_for_arithmetization()
    assert circ.log_n == 12 and circ.num_inputs == 17 and 2600 < circ.n_gates < 3000
    for gid, lay in cs.get_circuit_layout().items():     # the link values sit where the layout says
        vals = {ps.PARTY_LINKS[0]: parties[0], ps.PARTY_LINKS[1]: parties[1]}.get(gid)
        if vals is not None:
            first = vals.intent.in_token
            assert circ.wires_int[0][lay.offset << (circ.log_n - lay.alignment)] == first
    tau = oracle.int_to_limbs(py.to_mont(TAU % py.R, py.R))
    srs = oracle.srs_from_tau(tau, circ.n + 3)
    opk = oracle.plonk_preprocess(circ.log_n, circ.selectors, circ.perm, circ.k, srs)
    rc, proof, _, _ = oracle.plonk_prove(circ.log_n, circ.num_inputs, circ.k, opk, circ.wires, circ.pub_inputs,
                                         synth.splitmix_blinders(0x5E7), srs)
    assert rc == 0
    assert oracle.plonk_verify_known_tau(circ.log_n, circ.num_inputs, circ.k, opk, circ.pub_inputs, proof, tau)


def test_private_settlement_links_to_a_validity_side_circuit(oracle, pyoracle):
    """The settlement proof is linked to each party's validity proof through the values both circuits place in
    `intent_and_balance_settlement_party0` (native_proof_manager.rs:726-782).  Here the validity side is a stub circuit
    that holds the same 17 values under the same layout; the oracle links the two proofs and its verifier accepts."""
    from host_circuits import private_settlement as ps
    py = pyoracle
    parties, statement = ps.create_witness_statement(seed=12)
    cs = ps.IntentAndBalancePrivateSettlementCircuit.build(parties, statement)
    layout = cs.get_circuit_layout()[ps.PARTY_LINKS[0]]
    p0 = parties[0]
    shared = (p0.intent.to_scalars() + [p0.pre_settlement_amount_public_share] + p0.input_balance.to_scalars() +
              p0.pre_settlement_in_balance_shares)
    assert len(shared) == layout.size == 17
    stub = C.PlonkCircuit()
    stub.create_link_group(ps.PARTY_LINKS[0], C.GroupLayout(layout.alignment, layout.offset))
    vars_ = [stub.create_variable_with_link_groups(v, [ps.PARTY_LINKS[0]]) for v in shared]
    digest = stub.create_public_variable(C.compute_poseidon_hash(shared[:2]))
    C.PoseidonHashGadget(stub.zero()).hash_constrained(vars_[:2], digest, stub)
    stub.check_circuit_satisfiability(stub.public_input())
    tau = oracle.int_to_limbs(py.to_mont(TAU % py.R, py.R))
    hints = []
    circs = [cs.finalize_for_arithmetization(), stub.finalize_for_arithmetization()]
    srs = oracle.srs_from_tau(tau, max(c.n for c in circs) + 3)
    for i, circ in enumerate(circs):
        opk = oracle.plonk_preprocess(circ.log_n, circ.selectors, circ.perm, circ.k, srs[:circ.n + 3])
        rc, proof, _, link = oracle.plonk_prove(circ.log_n, circ.num_inputs, circ.k, opk, circ.wires, circ.pub_inputs,
                                                synth.splitmix_blinders(70 + i), srs[:circ.n + 3], True)
        assert rc == 0
        hints.append((link, proof.to_array()[:8].copy()))
    rc, lp, _ = oracle.plonk_link(hints[0][0], hints[1][0], hints[0][1], hints[1][1], layout.alignment, layout.offset,
                                  layout.size, srs)
    assert rc == 0
    assert oracle.plonk_link_verify_known_tau(hints[0][1], hints[1][1], layout.alignment, layout.offset, layout.size, lp, tau)


def test_intent_and_balance_validity_and_the_settlement_bundle(oracle, pyoracle):
    """INTENT AND BALANCE VALIDITY restated (renegade_b200/intent_and_balance_validity.py): 62 Poseidon2 permutations,
    n = 2^14, 10 public inputs, every statement field binding; OUTPUT BALANCE VALIDITY (output_balance_validity.py): 34
    permutations, n = 2^13, 5 public inputs.  Then the bundle the reference proves for a private match
    (native_proof_manager.rs:526-584, 726-782): each party's two validity proofs, the settlement proof over the same
    intents / balances / shares, and the FOUR link proofs between them — proved, linked and verified on the oracle."""
    from host_circuits import intent_and_balance_validity as val
    from host_circuits import output_balance_validity as obv
    from host_circuits import private_settlement as ps
    py = pyoracle
    # the match first (who trades what), then each party's validity proof over ITS intent and input balance
    parties, _ = ps.create_witness_statement(seed=21)
    validity = [val.create_witness_statement(seed=30 + i, intent=parties[i].intent, balance=parties[i].input_balance)
                for i in (0, 1)]
    out_validity = [obv.create_witness_statement(40 + i, parties[i].output_balance) for i in (0, 1)]
    parties, statement = ps.create_witness_statement(
        seed=21, linked=[(validity[i][0].new_amount_public_share, validity[i][0].post_match_balance_shares,
                          out_validity[i][0].post_match_balance_shares) for i in (0, 1)])
    settlement_cs = ps.IntentAndBalancePrivateSettlementCircuit.build(parties, statement)
    settlement_cs.check_circuit_satisfiability(statement.to_scalars())
    layouts = settlement_cs.get_circuit_layout()

    w0, s0 = validity[0]
    cs0 = val.IntentAndBalanceValidityCircuit.build(w0, s0, layouts)
    pub0 = s0.to_scalars()
    assert len(pub0) == 10
    cs0.check_circuit_satisfiability(pub0)
    for i in range(len(pub0)):
        bad = list(pub0)
        bad[i] = (bad[i] + 1) % C.R
        with pytest.raises(C.CircuitError):
            cs0.check_circuit_satisfiability(bad)
    from collections import Counter
    names = Counter(r.gate.name for r in cs0.rows)
    assert (names["FusedInternalSboxMDSGate"] + names["FusedExternalSboxMDSGate"]) == 62 * 195
    cs1 = val.IntentAndBalanceValidityCircuit.build(*validity[1], layouts)
    cs1.check_circuit_satisfiability(validity[1][1].to_scalars())
    out_cs = [obv.OutputBalanceValidityCircuit.build(w, s, layouts) for w, s in out_validity]
    for ocs, (_, s) in zip(out_cs, out_validity):
        assert len(s.to_scalars()) == 5
        ocs.check_circuit_satisfiability(s.to_scalars())
    onames = Counter(r.gate.name for r in out_cs[0].rows)
    assert (onames["FusedInternalSboxMDSGate"] + onames["FusedExternalSboxMDSGate"]) == 34 * 195

    tau = oracle.int_to_limbs(py.to_mont(TAU % py.R, py.R))
    circs = [settlement_cs.finalize_for_arithmetization(), cs0.finalize_for_arithmetization(), cs1.finalize_for_arithmetization(),
             out_cs[0].finalize_for_arithmetization(), out_cs[1].finalize_for_arithmetization()]
    assert [c.log_n for c in circs] == [12, 14, 14, 13, 13]
    srs = oracle.srs_from_tau(tau, (1 << 14) + 3)
    hints = []
    for i, circ in enumerate(circs):
        opk = oracle.plonk_preprocess(circ.log_n, circ.selectors, circ.perm, circ.k, srs[:circ.n + 3])
        rc, proof, _, link = oracle.plonk_prove(circ.log_n, circ.num_inputs, circ.k, opk, circ.wires, circ.pub_inputs,
                                                synth.splitmix_blinders(90 + i), srs[:circ.n + 3], True)
        assert rc == 0
        assert oracle.plonk_verify_known_tau(circ.log_n, circ.num_inputs, circ.k, opk, circ.pub_inputs, proof, tau)
        hints.append((link, proof.to_array()[:8].copy()))
    for party in (0, 1):                         # each validity proof of party i <-> settlement proof, on ITS group
        for lay, v in ((layouts[ps.PARTY_LINKS[party]], hints[1 + party]), (layouts[ps.OUTPUT_LINKS[party]], hints[3 + party])):
            rc, lp, _ = oracle.plonk_link(v[0], hints[0][0], v[1], hints[0][1], lay.alignment, lay.offset, lay.size, srs)
            assert rc == 0
            assert oracle.plonk_link_verify_known_tau(v[1], hints[0][1], lay.alignment, lay.offset, lay.size, lp, tau)
    # party 1's validity proof does not link on party 0's group: the settlement holds party 0's values there
    lay = layouts[ps.PARTY_LINKS[0]]
    rc, _, _ = oracle.plonk_link(hints[2][0], hints[0][0], hints[2][1], hints[0][1], lay.alignment, lay.offset, lay.size, srs)
    assert rc == 2                               # the prover refuses: the polynomials differ on the group
    rc, lp, _ = oracle.plonk_link(hints[1][0], hints[0][0], hints[1][1], hints[0][1], lay.alignment, lay.offset, lay.size, srs)
    assert rc == 0                               # party 0's honest link proof does not carry over to party 1's commitment
    assert not oracle.plonk_link_verify_known_tau(hints[2][1], hints[0][1], lay.alignment, lay.offset, lay.size, lp, tau)


@pytest.mark.parametrize("which", ["deposit", "withdrawal", "cancellation"])
def test_state_update_circuits(oracle, pyoracle, which):
    """VALID DEPOSIT, VALID WITHDRAWAL and VALID ORDER CANCELLATION restated (examples/host_circuits/state_updates.py): the
    native witness / statement satisfy the circuit, every statement field is binding, the rules the reference tests
    (valid_deposit.rs / valid_withdrawal.rs / valid_order_cancellation.rs `mod test`) are enforced, and the oracle
    prover's proof is accepted by the oracle verifier."""
    from host_circuits import state_updates as su
    py = pyoracle
    make, circuit, n_inputs, gates = {
        "deposit": (su.create_deposit_witness_statement, su.ValidDeposit, 8, (6500, 7200)),
        "withdrawal": (su.create_withdrawal_witness_statement, su.ValidWithdrawal, 8, (6500, 7200)),
        "cancellation": (su.create_cancellation_witness_statement, su.ValidOrderCancellationCircuit, 3, (4200, 4700)),
    }[which]
    witness, statement = make(11)
    cs = circuit.build(witness, statement)
    pub = statement.to_scalars()
    assert cs.public_input() == pub and len(pub) == n_inputs
    cs.check_circuit_satisfiability(pub)
    for i in range(len(pub)):                     # no statement field is free
        bad = list(pub)
        bad[i] = (bad[i] + 1) % C.R
        with pytest.raises(C.CircuitError):
            cs.check_circuit_satisfiability(bad)

    def unsatisfied(w, s):
        with pytest.raises(C.CircuitError):
            circuit.build(w, s).check_circuit_satisfiability(s.to_scalars())

    if which == "deposit":
        # a deposit that overflows the amount range (valid_deposit.rs: test_invalid_deposit__amount_overflow)
        w, s = make(12)
        w.old_balance.inner[su.AMOUNT_IDX] = (1 << C.AMOUNT_BITS) - 1
        unsatisfied(w, s)
        # a deposit of another token / from another owner than the balance's
        w, s = make(13)
        s.deposit.token ^= 1
        unsatisfied(w, s)
    elif which == "withdrawal":
        # outstanding fees block a withdrawal; a zero withdrawal and one above the balance are refused
        w, s = make(12)
        w.old_balance.inner[5] = 1
        unsatisfied(w, s)
        w, s = make(13)
        s.withdrawal.amount = 0
        unsatisfied(w, s)
        w, s = make(14)
        s.withdrawal.amount = w.old_balance.inner[su.AMOUNT_IDX] + 1
        unsatisfied(w, s)
    else:
        # somebody else's intent, a forged opening
        w, s = make(12)
        s.owner ^= 1
        unsatisfied(w, s)
        w, s = make(13)
        w.old_intent_opening.elems[3] ^= 1
        unsatisfied(w, s)
    circ = cs.finalize_for_arithmetization()
    assert circ.log_n == 13 and circ.num_inputs == n_inputs and gates[0] < circ.n_gates < gates[1]
    tau = oracle.int_to_limbs(py.to_mont(TAU % py.R, py.R))
    srs = oracle.srs_from_tau(tau, circ.n + 3)
    opk = oracle.plonk_preprocess(circ.log_n, circ.selectors, circ.perm, circ.k, srs)
    rc, proof, _, _ = oracle.plonk_prove(circ.log_n, circ.num_inputs, circ.k, opk, circ.wires, circ.pub_inputs,
                                         synth.splitmix_blinders(0xBA2), srs)
    assert rc == 0
    assert oracle.plonk_verify_known_tau(circ.log_n, circ.num_inputs, circ.k, opk, circ.pub_inputs, proof, tau)


def test_intent_only_family_and_its_link(oracle, pyoracle):
    """INTENT ONLY VALIDITY / FIRST FILL VALIDITY / PUBLIC SETTLEMENT / BOUNDED SETTLEMENT restated
    (examples/host_circuits/intent_only.py): each satisfied by its native witness / statement with every statement field
    binding and the rules of the reference's tests enforced; the four share ONE link group placed by the public
    settlement circuit; the oracle proves a validity proof and a settlement proof of the same intent, links them
    (proof_linking/intent_only.rs:34-49: validity hint first) and the link verifies — and does not for another intent."""
    from host_circuits import intent_only as io
    py = pyoracle
    rnd = random.Random(77)
    intent = io.Intent(rnd.randrange(1 << 160), rnd.randrange(1 << 160), rnd.randrange(1 << 160),
                       rnd.randrange(1, 1 << 80), rnd.randrange(1 << 40, 1 << 70))
    cases = {
        "validity": (io.IntentOnlyValidityCircuit, io.create_validity_witness_statement(5, intent), 13, 7),
        "first_fill": (io.IntentOnlyFirstFillValidityCircuit, io.create_first_fill_witness_statement(6, intent), 11, 8),
        "public": (io.IntentOnlyPublicSettlementCircuit, io.create_public_settlement_witness_statement(7, intent), 9, 6),
        "bounded": (io.IntentOnlyBoundedSettlementCircuit, io.create_bounded_settlement_witness_statement(8, intent), 9, 9),
    }
    circs, layout = {}, None
    for key, (circuit, (w, s), log_n, n_inputs) in cases.items():
        cs = circuit.build(w, s)
        pub = s.to_scalars()
        assert cs.public_input() == pub and len(pub) == n_inputs
        cs.check_circuit_satisfiability(pub)
        for i in range(len(pub)):
            bad = list(pub)
            bad[i] = (bad[i] + 1) % C.R
            with pytest.raises(C.CircuitError):
                cs.check_circuit_satisfiability(bad)
        lay = cs.get_circuit_layout()[io.INTENT_ONLY_SETTLEMENT_LINK]
        assert (lay.alignment, lay.offset, lay.size) == (9, 16, 5)
        layout = layout or lay
        circs[key] = cs.finalize_for_arithmetization()
        assert circs[key].log_n == log_n, key

    def unsatisfied(circuit, w, s):
        with pytest.raises(C.CircuitError):
            circuit.build(w, s).check_circuit_satisfiability(s.to_scalars())

    # settlement: more than the intent allows, a worse price, the wrong pair
    w, s = io.create_public_settlement_witness_statement(9, intent)
    s.settlement_obligation.amount_in = intent.amount_in + 1
    unsatisfied(io.IntentOnlyPublicSettlementCircuit, w, s)
    w, s = io.create_public_settlement_witness_statement(9, intent)
    s.settlement_obligation.amount_out = ((intent.min_price * s.settlement_obligation.amount_in) >> 63) - 1
    unsatisfied(io.IntentOnlyPublicSettlementCircuit, w, s)
    w, s = io.create_bounded_settlement_witness_statement(9, intent)
    s.bounded_match_result.price = intent.min_price - 1
    unsatisfied(io.IntentOnlyBoundedSettlementCircuit, w, s)
    w, s = io.create_bounded_settlement_witness_statement(9, intent)
    s.bounded_match_result.max_internal_party_amount_in = intent.amount_in + 1
    unsatisfied(io.IntentOnlyBoundedSettlementCircuit, w, s)
    # validity: the linked copy must be the state's intent; first fill: amount / price ranges, the owner
    w, s = io.create_validity_witness_statement(10, intent)
    w.intent = io.Intent(intent.in_token, intent.out_token, intent.owner, intent.min_price, intent.amount_in + 1)
    unsatisfied(io.IntentOnlyValidityCircuit, w, s)
    for field, value in (("amount_in", 1 << C.AMOUNT_BITS), ("min_price", 1 << io.PRICE_BITS)):
        bad_intent = io.Intent(**{**intent.__dict__, field: value})
        unsatisfied(io.IntentOnlyFirstFillValidityCircuit, *io.create_first_fill_witness_statement(11, bad_intent))
    w, s = io.create_first_fill_witness_statement(12, intent)
    s.owner ^= 1
    unsatisfied(io.IntentOnlyFirstFillValidityCircuit, w, s)

    # prove validity (2^13) and public settlement (2^9) on one SRS, link the group, verify the link
    tau = oracle.int_to_limbs(py.to_mont(TAU % py.R, py.R))
    srs = oracle.srs_from_tau(tau, (1 << 13) + 3)

    def prove(c, seed):
        opk = oracle.plonk_preprocess(c.log_n, c.selectors, c.perm, c.k, srs[:c.n + 3])
        rc, proof, _, link = oracle.plonk_prove(c.log_n, c.num_inputs, c.k, opk, c.wires, c.pub_inputs,
                                                synth.splitmix_blinders(seed), srs[:c.n + 3], True)
        assert rc == 0 and oracle.plonk_verify_known_tau(c.log_n, c.num_inputs, c.k, opk, c.pub_inputs, proof, tau)
        return link, proof.to_array()[:8].copy()           # (linking wire polynomial, its commitment)

    validity, settlement = prove(circs["validity"], 0xC0), prove(circs["public"], 0xC1)
    for j, v in enumerate(intent.to_scalars()):          # both circuits hold the intent on the group's rows
        for key in ("validity", "public", "first_fill", "bounded"):
            c = circs[key]
            assert c.wires_int[0][(layout.offset + j) << (c.log_n - layout.alignment)] == v
    rc, lp, _ = oracle.plonk_link(validity[0], settlement[0], validity[1], settlement[1], layout.alignment, layout.offset,
                                  layout.size, srs)
    assert rc == 0
    assert oracle.plonk_link_verify_known_tau(validity[1], settlement[1], layout.alignment, layout.offset, layout.size, lp, tau)
    # a settlement proof about ANOTHER intent does not link: the prover refuses, the honest link proof does not transfer
    other = prove(io.IntentOnlyPublicSettlementCircuit.build(*io.create_public_settlement_witness_statement(13))
                  .finalize_for_arithmetization(), 0xC2)
    rc, _, _ = oracle.plonk_link(validity[0], other[0], validity[1], other[1], layout.alignment, layout.offset, layout.size, srs)
    assert rc == 2
    assert not oracle.plonk_link_verify_known_tau(validity[1], other[1], layout.alignment, layout.offset, layout.size, lp, tau)


def test_public_and_bounded_settlement_circuits(oracle, pyoracle):
    """INTENT AND BALANCE PUBLIC SETTLEMENT / BOUNDED SETTLEMENT restated (examples/host_circuits/public_settlement.py):
    satisfied by party 0 of a consistent match, every statement field binding, the rules the reference's tests exercise
    enforced, the PARTY 0 groups of the private settlement layout inherited — so the validity proof that links to a
    private settlement links to the public one as well (checked on the oracle)."""
    from host_circuits import intent_and_balance_validity as val
    from host_circuits import private_settlement as ps
    from host_circuits import public_settlement as pub
    py = pyoracle
    layouts = ps.IntentAndBalancePrivateSettlementCircuit.build(*ps.create_witness_statement(0)).get_circuit_layout()
    built = {}
    for key, make, circuit, n_inputs in (("public", pub.create_public_witness_statement, pub.IntentAndBalancePublicSettlementCircuit, 14),
                                         ("bounded", pub.create_bounded_witness_statement, pub.IntentAndBalanceBoundedSettlementCircuit, 16)):
        w, s = make(21)
        cs = circuit.build(w, s, layouts)
        pubs = s.to_scalars()
        assert cs.public_input() == pubs and len(pubs) == n_inputs
        cs.check_circuit_satisfiability(pubs)
        for i in range(len(pubs)):
            bad = list(pubs)
            bad[i] = (bad[i] + 1) % C.R
            with pytest.raises(C.CircuitError):
                cs.check_circuit_satisfiability(bad)
        lay = cs.get_circuit_layout()
        assert {g: (l.alignment, l.offset, l.size) for g, l in lay.items()} == \
            {pub.PARTY_LINK: (layouts[pub.PARTY_LINK].alignment, layouts[pub.PARTY_LINK].offset, 17),
             pub.OUTPUT_LINK: (layouts[pub.OUTPUT_LINK].alignment, layouts[pub.OUTPUT_LINK].offset, 11)}
        built[key] = (w, s, cs.finalize_for_arithmetization())
        assert built[key][2].log_n == 12

    def unsatisfied(circuit, w, s):
        with pytest.raises(C.CircuitError):
            circuit.build(w, s, layouts).check_circuit_satisfiability(s.to_scalars())

    P, B = pub.IntentAndBalancePublicSettlementCircuit, pub.IntentAndBalanceBoundedSettlementCircuit
    w, s = pub.create_public_witness_statement(22)          # the balance does not cover the obligation
    w.in_balance.amount = s.settlement_obligation.amount_in - 1
    unsatisfied(P, w, s)
    w, s = pub.create_public_witness_statement(22)          # the receive balance would overflow
    w.out_balance.amount = (1 << C.AMOUNT_BITS) - 1
    unsatisfied(P, w, s)
    w, s = pub.create_public_witness_statement(22)          # someone else's output balance
    w.out_balance.owner ^= 1
    unsatisfied(P, w, s)
    w, s = pub.create_public_witness_statement(22)          # a leaked share that is not the linked one
    s.in_balance_public_shares[2] = (s.in_balance_public_shares[2] + 1) % C.R
    unsatisfied(P, w, s)
    w, s = pub.create_bounded_witness_statement(22)         # the bound exceeds the capitalising balance
    w.in_balance.amount = s.bounded_match_result.max_internal_party_amount_in - 1
    unsatisfied(B, w, s)
    w, s = pub.create_bounded_witness_statement(22)         # a price below the intent's worst case
    s.bounded_match_result.price = w.intent.min_price - 1
    unsatisfied(B, w, s)

    # the party's validity proof links to the PUBLIC settlement proof on the inherited party-0 group
    w, s, circ = built["public"]
    vw, vs = val.create_witness_statement(23, intent=w.intent, balance=w.in_balance)
    w2 = pub.Witness(w.intent, vw.new_amount_public_share, w.in_balance, list(vw.post_match_balance_shares), w.out_balance,
                     w.pre_settlement_out_balance_shares)
    s2 = pub.PublicStatement(s.settlement_obligation, w2.pre_settlement_amount_public_share, list(w2.pre_settlement_in_balance_shares),
                             list(s.out_balance_public_shares), s.relayer_fee_rate, s.protocol_fee_rate, s.relayer_fee_recipient)
    cs2 = P.build(w2, s2, layouts)
    cs2.check_circuit_satisfiability(s2.to_scalars())
    tau = oracle.int_to_limbs(py.to_mont(TAU % py.R, py.R))
    srs = oracle.srs_from_tau(tau, (1 << 14) + 3)

    def prove(c, seed):
        opk = oracle.plonk_preprocess(c.log_n, c.selectors, c.perm, c.k, srs[:c.n + 3])
        rc, proof, _, link = oracle.plonk_prove(c.log_n, c.num_inputs, c.k, opk, c.wires, c.pub_inputs,
                                                synth.splitmix_blinders(seed), srs[:c.n + 3], True)
        assert rc == 0 and oracle.plonk_verify_known_tau(c.log_n, c.num_inputs, c.k, opk, c.pub_inputs, proof, tau)
        return link, proof.to_array()[:8].copy()

    validity = prove(val.IntentAndBalanceValidityCircuit.build(vw, vs, layouts).finalize_for_arithmetization(), 0xD0)
    settlement = prove(cs2.finalize_for_arithmetization(), 0xD1)
    g = layouts[pub.PARTY_LINK]
    rc, lp, _ = oracle.plonk_link(validity[0], settlement[0], validity[1], settlement[1], g.alignment, g.offset, 17, srs)
    assert rc == 0
    assert oracle.plonk_link_verify_known_tau(validity[1], settlement[1], g.alignment, g.offset, 17, lp, tau)


@pytest.mark.parametrize("which", ["note_redemption", "public_protocol_fee", "public_relayer_fee"])
def test_fee_circuits(oracle, pyoracle, which):
    """VALID NOTE REDEMPTION and the two PUBLIC fee payments restated (examples/host_circuits/fees.py): satisfied, every
    statement field binding, the rules of the reference's tests (fees/*.rs `mod test`) enforced; oracle proof + verify."""
    from host_circuits import fees
    py = pyoracle
    make, circuit, n_inputs, log_n = {
        "note_redemption": (fees.create_note_redemption_witness_statement, fees.ValidNoteRedemption, 6, 12),
        "public_protocol_fee": (fees.create_public_protocol_fee_payment_witness_statement, fees.ValidPublicProtocolFeePayment, 9, 13),
        "public_relayer_fee": (fees.create_public_relayer_fee_payment_witness_statement, fees.ValidPublicRelayerFeePayment, 9, 13),
    }[which]
    witness, statement = make(15)
    cs = circuit.build(witness, statement)
    pub = statement.to_scalars()
    assert cs.public_input() == pub and len(pub) == n_inputs
    cs.check_circuit_satisfiability(pub)
    for i in range(len(pub)):
        bad = list(pub)
        bad[i] = (bad[i] + 1) % C.R
        with pytest.raises(C.CircuitError):
            cs.check_circuit_satisfiability(bad)

    def unsatisfied(w, s):
        with pytest.raises(C.CircuitError):
            circuit.build(w, s).check_circuit_satisfiability(s.to_scalars())

    if which == "note_redemption":
        w, s = make(16)                            # a nullifier of another blinder, a forged opening
        s.note.blinder ^= 1
        unsatisfied(w, s)
        w, s = make(16)
        w.note_opening.elems[0] ^= 1
        unsatisfied(w, s)
    else:
        idx = fees.PROTOCOL_FEE_IDX if which == "public_protocol_fee" else fees.RELAYER_FEE_IDX
        w, s = make(16)                            # nothing to pay: a zero fee balance is refused
        w.old_balance.inner[idx] = 0
        unsatisfied(w, s)
        w, s = make(16)                            # a note of another amount / mint
        s.note.amount += 1
        unsatisfied(w, s)
        w, s = make(16)
        s.note.mint ^= 1
        unsatisfied(w, s)
        w, s = make(16)                            # the receiver: free for the protocol fee, bound for the relayer fee
        s.note.receiver ^= 1
        if which == "public_relayer_fee":
            unsatisfied(w, s)
        else:
            circuit.build(w, s).check_circuit_satisfiability(s.to_scalars())
    circ = cs.finalize_for_arithmetization()
    assert circ.log_n == log_n and circ.num_inputs == n_inputs
    tau = oracle.int_to_limbs(py.to_mont(TAU % py.R, py.R))
    srs = oracle.srs_from_tau(tau, circ.n + 3)
    opk = oracle.plonk_preprocess(circ.log_n, circ.selectors, circ.perm, circ.k, srs)
    rc, proof, _, _ = oracle.plonk_prove(circ.log_n, circ.num_inputs, circ.k, opk, circ.wires, circ.pub_inputs,
                                         synth.splitmix_blinders(0xBA3), srs)
    assert rc == 0
    assert oracle.plonk_verify_known_tau(circ.log_n, circ.num_inputs, circ.k, opk, circ.pub_inputs, proof, tau)
