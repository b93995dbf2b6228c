"""CPU tests of the prover service's host logic (renegade_b200/service.py): the JSON codec round-trips every witness /
statement type and the proof / hint / link-proof shapes (api_types.rs:81-130), auth and routing behave like the server
the reference's `ProofServiceClient` expects (prover_service_client.rs:93, 186-205).  Proving itself needs the GPU
(tests/test_gpu_service.py)."""
import numpy as np

from host_circuits import intent_and_balance_validity as val
from host_circuits import output_balance_validity as obv
from host_circuits import private_settlement as ps
from host_circuits import service_routes
from host_circuits import valid_balance_create as vbc
from renegade_b200 import service as sv
from renegade_b200.backend import B200LinkProof, B200Proof, LinkingHint
from renegade_b200.fields import scalars_to_limbs


def test_witness_statement_codec_round_trips():
    import json
    from typing import List
    w, s = vbc.create_witness_statement(3)
    for tp, obj in ((vbc.ValidBalanceCreateWitness, w), (vbc.ValidBalanceCreateStatement, s)):
        back = sv.from_json(tp, json.loads(json.dumps(sv.to_json(obj))))
        assert back == obj
    assert "from" in sv.to_json(s)["deposit"]           # the reference's field name, not Python's `from_`
    parties, st_ = ps.create_witness_statement(5)
    assert sv.from_json(List[ps.PartyWitness], json.loads(json.dumps(sv.to_json(parties)))) == parties
    assert sv.from_json(ps.Statement, sv.to_json(st_)) == st_
    w2, s2 = val.create_witness_statement(7)
    assert sv.from_json(val.Witness, json.loads(json.dumps(sv.to_json(w2)))) == w2
    assert sv.from_json(val.Statement, sv.to_json(s2)) == s2
    w3, s3 = obv.create_witness_statement(9, parties[0].output_balance)
    assert sv.from_json(obv.Witness, sv.to_json(w3)) == w3 and sv.from_json(obv.Statement, sv.to_json(s3)) == s3
    from host_circuits import state_updates as su
    for tp_w, tp_s, make in ((su.BalanceUpdateWitness, su.ValidDepositStatement, su.create_deposit_witness_statement),
                             (su.BalanceUpdateWitness, su.ValidWithdrawalStatement, su.create_withdrawal_witness_statement),
                             (su.ValidOrderCancellationWitness, su.ValidOrderCancellationStatement, su.create_cancellation_witness_statement)):
        w4, s4 = make(11)
        assert sv.from_json(tp_w, json.loads(json.dumps(sv.to_json(w4)))) == w4
        assert sv.from_json(tp_s, json.loads(json.dumps(sv.to_json(s4)))) == s4
    assert set(sv.to_json(su.create_withdrawal_witness_statement(1)[1])["withdrawal"]) == {"to", "token", "amount"}
    # scalars are accepted as decimal, hex and 32-byte big-endian arrays
    v = 0x1234567890abcdef1234567890abcdef
    assert sv.decode_scalar(str(v)) == sv.decode_scalar(hex(v)) == sv.decode_scalar(list(v.to_bytes(32, "big"))) == v


def test_proof_codec_round_trips():
    rnd = np.random.default_rng(1)
    p = B200Proof()
    fq = lambda: [int(x) for x in scalars_to_limbs([int(rnd.integers(1, 2**62)) ** 3 % sv.BASE_FIELD_MODULUS], sv.BASE_FIELD_MODULUS)[0]]
    fr = lambda: [int(x) for x in scalars_to_limbs([int(rnd.integers(1, 2**62)) ** 3])[0]]
    for i in range(5):
        p.wires_poly_comms[i][:] = fq() + fq()
        p.split_quot_poly_comms[i][:] = fq() + fq()
        p.wires_evals[i][:] = fr()
    for i in range(4):
        p.wire_sigma_evals[i][:] = fr()
    p.prod_perm_poly_comm[:] = fq() + fq()
    p.opening_proof[:] = fq() + fq()
    p.shifted_opening_proof[:] = fq() + fq()
    p.perm_next_eval[:] = fr()
    enc = sv.encode_proof(p)
    assert set(enc) == {"wires_poly_comms", "prod_perm_poly_comm", "split_quot_poly_comms", "opening_proof",
                        "shifted_opening_proof", "poly_evals", "plookup_proof"} and enc["plookup_proof"] is None
    assert bytes(sv.decode_proof(enc)) == bytes(p)
    lp = B200LinkProof()
    lp.quotient_commitment[:] = fq() + fq()
    lp.opening_proof[:] = fq() + fq()
    assert bytes(sv.decode_link_proof(sv.encode_link_proof(lp))) == bytes(lp)
    hint = LinkingHint(linking_wire_poly=scalars_to_limbs(list(range(1, 70))), linking_wire_comm=np.array(fq() + fq(), dtype=np.uint64))
    back = sv.decode_link_hint(sv.encode_link_hint(hint))
    assert (back.linking_wire_poly == hint.linking_wire_poly).all() and (back.linking_wire_comm == hint.linking_wire_comm).all()


def test_auth_and_routing_without_a_gpu():
    import threading
    routes = service_routes.routes()
    assert set(routes) <= set(sv.ALL_PATHS) and len(sv.ALL_PATHS) == 20
    service = sv.ProverService(routes, password="hunter2")
    server = service.make_server("127.0.0.1", 0)
    th = threading.Thread(target=server.serve_forever, daemon=True)
    th.start()
    try:
        url = f"http://127.0.0.1:{server.server_address[1]}"
        good, bad = sv.ProofServiceClient(url, "hunter2"), sv.ProofServiceClient(url, "wrong")
        assert bad.send_request("/prove-valid-balance-create", {})[0] == 401
        assert good.send_request("/prove-valid-private-protocol-fee-payment", {})[0] == 501    # a reference path whose circuit is not restated
        assert good.send_request("/prove-something-else", {})[0] == 404
        code, body = good.send_request("/prove-valid-balance-create", {"statement": {}})  # malformed: no witness
        assert code == 400 and "bad request" in body["error"]
        assert service.stats["requests"] == 4 and service.stats["proofs"] == 0
    finally:
        server.shutdown()


def test_every_route_decodes_and_synthesizes_the_keyed_structure():
    """What the service does before the device call, for every registered path: decode the JSON request, synthesize,
    arithmetize.  The circuit STRUCTURE (selectors, copy permutation, public-input count) of a fresh instance must be the
    one the key was preprocessed from (`dummy_instance`, traits.rs:821-855: keys are cached by circuit name) — i.e.
    synthesis is witness-independent — and the instance must satisfy it."""
    import json
    from host_circuits import fees
    from host_circuits import intent_only as io
    from host_circuits import public_settlement as pub
    from host_circuits import state_updates as su
    from host_circuits import statements as st
    routes = service_routes.routes()
    assert len(routes) == 16
    parties, _ = ps.create_witness_statement(31)
    fresh = {
        "/prove-valid-balance-create": vbc.create_witness_statement(31),
        "/prove-valid-deposit": su.create_deposit_witness_statement(31),
        "/prove-valid-withdrawal": su.create_withdrawal_witness_statement(31),
        "/prove-valid-order-cancellation": su.create_cancellation_witness_statement(31),
        "/prove-valid-note-redemption": fees.create_note_redemption_witness_statement(31),
        "/prove-valid-public-protocol-fee-payment": fees.create_public_protocol_fee_payment_witness_statement(31),
        "/prove-valid-public-relayer-fee-payment": fees.create_public_relayer_fee_payment_witness_statement(31),
        "/prove-intent-only-validity": io.create_validity_witness_statement(31),
        "/prove-intent-only-first-fill-validity": io.create_first_fill_witness_statement(31),
        "/prove-intent-only-public-settlement": io.create_public_settlement_witness_statement(31),
        "/prove-intent-only-bounded-settlement": io.create_bounded_settlement_witness_statement(31),
        "/prove-intent-and-balance-validity": val.create_witness_statement(31),
        "/prove-output-balance-validity": obv.create_witness_statement(31, parties[0].output_balance),
        "/prove-intent-and-balance-private-settlement": ps.create_witness_statement(31),
        "/prove-intent-and-balance-public-settlement": pub.create_public_witness_statement(31),
        "/prove-intent-and-balance-bounded-settlement": pub.create_bounded_witness_statement(31),
    }
    assert set(fresh) == set(routes)
    for path, route in routes.items():
        C = route.circuit
        assert C in st.REGISTERED
        w, s = fresh[path]
        body = json.loads(json.dumps({"witness": sv.to_json(w), "statement": sv.to_json(s)}))
        w2, s2 = route.decode_witness(body["witness"]), route.decode_statement(body["statement"])
        layout = C.get_circuit_layout()
        cs = C.synthesize(w2, s2, layout)
        cs.check_circuit_satisfiability(s2.to_scalars())
        circ = cs.finalize_for_arithmetization()
        key_circ = C.synthesize(*C.dummy_instance(), layout).finalize_for_arithmetization()
        assert (circ.log_n, circ.num_inputs) == (key_circ.log_n, key_circ.num_inputs), path
        assert (circ.selectors == key_circ.selectors).all() and (circ.perm == key_circ.perm).all(), path
        assert (C.statement_scalars(s2) == circ.pub_inputs).all(), path
