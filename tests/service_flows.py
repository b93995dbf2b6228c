"""Request flows of the prover service for the later statements, shared by the GPU test (tests/test_gpu_service.py: the
device proves) and the CPU test (tests/test_service_flows_cpu.py: the same service, routes and `SingleProverCircuit`
surface with the oracle prover standing in for the device).  Everything a flow CHECKS is product host code: the JSON codec,
the routing, `verify_singleprover_proof` / `verify_link_proof` (the pairing verifier of csrc/verify.cu)."""
import numpy as np

from host_circuits import fees
from host_circuits import intent_and_balance_validity as val
from host_circuits import intent_only as io
from host_circuits import output_balance_validity as obv
from host_circuits import public_settlement as pub
from host_circuits import state_updates as su
from host_circuits import statements as st
from renegade_b200 import circuit_types as ct
from renegade_b200 import service as sv
from renegade_b200.backend import GroupLayout, verify_link_proof


def plain_proof_paths(client, service, negatives=True):
    """The six `ProofResponse` paths; a statement that does not belong to the witness comes back as a prover error."""
    before = service.stats["proofs"]
    cases = (("/prove-valid-deposit", st.ValidDeposit, su.create_deposit_witness_statement),
             ("/prove-valid-withdrawal", st.ValidWithdrawal, su.create_withdrawal_witness_statement),
             ("/prove-valid-order-cancellation", st.ValidOrderCancellationCircuit, su.create_cancellation_witness_statement),
             ("/prove-valid-note-redemption", st.ValidNoteRedemption, fees.create_note_redemption_witness_statement),
             ("/prove-valid-public-protocol-fee-payment", st.ValidPublicProtocolFeePayment,
              fees.create_public_protocol_fee_payment_witness_statement),
             ("/prove-valid-public-relayer-fee-payment", st.ValidPublicRelayerFeePayment,
              fees.create_public_relayer_fee_payment_witness_statement))
    for i, (path, circuit, make) in enumerate(cases):
        w, s = make(41)
        code, body = client.send_request(path, {"statement": sv.to_json(s), "witness": sv.to_json(w)})
        assert code == 200 and set(body) == {"proof"}, (path, body)
        ct.verify_singleprover_proof(circuit, s, sv.decode_proof(body["proof"]))
        if negatives is True or (negatives == "first" and i == 0):
            _, s_other = make(42)
            code, body = client.send_request(path, {"statement": sv.to_json(s_other), "witness": sv.to_json(w)})
            assert code == 500 and "ProverError" in body["error"], path
    assert service.stats["proofs"] == before + len(cases)


def intent_only_flow(client, service, g2):
    """INTENT ONLY VALIDITY / FIRST FILL VALIDITY (`ProofAndHintResponse`), then PUBLIC / BOUNDED SETTLEMENT with the validity
    proof's hint (`SettlementProofResponse {proof, link_proof}`, api_types.rs:97-104, 279-299)."""
    h, tau_h = g2
    proofs0, links0 = service.stats["proofs"], service.stats["link_proofs"]
    intent = io.create_public_settlement_witness_statement(51)[0].intent
    lay = st.IntentOnlyPublicSettlementCircuit.get_circuit_layout()[io.INTENT_ONLY_SETTLEMENT_LINK]
    group = GroupLayout(lay.alignment, lay.offset, lay.size)
    hints, comms = {}, {}
    for key, path, circuit, (w, s) in (
            ("validity", "/prove-intent-only-validity", st.IntentOnlyValidityCircuit, io.create_validity_witness_statement(52, intent)),
            ("first_fill", "/prove-intent-only-first-fill-validity", st.IntentOnlyFirstFillValidityCircuit,
             io.create_first_fill_witness_statement(53, intent))):
        code, body = client.send_request(path, {"statement": sv.to_json(s), "witness": sv.to_json(w)})
        assert code == 200 and set(body) == {"proof", "link_hint"}, (path, body)
        proof = sv.decode_proof(body["proof"])
        ct.verify_singleprover_proof(circuit, s, proof)
        hints[key], comms[key] = body["link_hint"], sv.decode_link_hint(body["link_hint"]).linking_wire_comm
        assert (comms[key] == np.array(proof.wires_poly_comms[0], dtype=np.uint64)).all()
    for path, circuit, (w, s), hint_key in (
            ("/prove-intent-only-public-settlement", st.IntentOnlyPublicSettlementCircuit,
             io.create_public_settlement_witness_statement(54, intent), "validity"),
            ("/prove-intent-only-bounded-settlement", st.IntentOnlyBoundedSettlementCircuit,
             io.create_bounded_settlement_witness_statement(55, intent), "first_fill")):
        code, body = client.send_request(path, {"statement": sv.to_json(s), "witness": sv.to_json(w),
                                                "validity_link_hint": hints[hint_key]})
        assert code == 200 and set(body) == {"proof", "link_proof"}, (path, body)
        proof = sv.decode_proof(body["proof"])
        ct.verify_singleprover_proof(circuit, s, proof)
        s_comm = np.array(proof.wires_poly_comms[0], dtype=np.uint64)
        assert verify_link_proof(comms[hint_key], s_comm, sv.decode_link_proof(body["link_proof"]), group, h, tau_h)
    # a settlement about ANOTHER intent cannot be linked to this validity proof: the prover refuses the division
    w, s = io.create_public_settlement_witness_statement(56)
    code, body = client.send_request("/prove-intent-only-public-settlement",
                                     {"statement": sv.to_json(s), "witness": sv.to_json(w), "validity_link_hint": hints["validity"]})
    assert code == 500 and "ProverError" in body["error"], body
    assert service.stats["proofs"] == proofs0 + 5 and service.stats["link_proofs"] == links0 + 2


def public_settlement_flow(client, service, g2, which=("public", "bounded")):
    """INTENT AND BALANCE PUBLIC / BOUNDED SETTLEMENT: the party's INTENT AND BALANCE VALIDITY and OUTPUT BALANCE VALIDITY proofs,
    then the settlement request with both hints (`PublicSettlementProofResponse`, api_types.rs:126-137, 238-277)."""
    h, tau_h = g2
    links0 = service.stats["link_proofs"]
    layouts = st.IntentAndBalancePrivateSettlementCircuit.get_circuit_layout()
    cases = {"public": ("/prove-intent-and-balance-public-settlement", st.IntentAndBalancePublicSettlementCircuit,
                        pub.create_public_witness_statement, pub.PublicStatement),
             "bounded": ("/prove-intent-and-balance-bounded-settlement", st.IntentAndBalanceBoundedSettlementCircuit,
                         pub.create_bounded_witness_statement, pub.BoundedStatement)}
    for key in which:
        path, circuit, make, Statement = cases[key]
        w, s = make(61)
        vw, vs = val.create_witness_statement(62, intent=w.intent, balance=w.in_balance)
        ow, os_ = obv.create_witness_statement(63, w.out_balance)
        hints = {}
        for hk, vpath, (w_, s_), C in (("v", "/prove-intent-and-balance-validity", (vw, vs), st.IntentAndBalanceValidityCircuit),
                                      ("o", "/prove-output-balance-validity", (ow, os_), st.OutputBalanceValidityCircuit)):
            code, body = client.send_request(vpath, {"statement": sv.to_json(s_), "witness": sv.to_json(w_)})
            assert code == 200 and set(body) == {"proof", "link_hint"}, (vpath, body)
            ct.verify_singleprover_proof(C, s_, sv.decode_proof(body["proof"]))
            hints[hk] = body["link_hint"]
        # the settlement witness carries the shares the validity proofs produced (what the link proofs check)
        w2 = pub.Witness(w.intent, vw.new_amount_public_share, w.in_balance, list(vw.post_match_balance_shares),
                         w.out_balance, list(ow.post_match_balance_shares))
        fields = dict(s.__dict__)
        fields.update(amount_public_share=w2.pre_settlement_amount_public_share,
                      in_balance_public_shares=list(w2.pre_settlement_in_balance_shares),
                      out_balance_public_shares=list(w2.pre_settlement_out_balance_shares))
        s2 = Statement(**fields)
        code, body = client.send_request(path, {"statement": sv.to_json(s2), "witness": sv.to_json(w2),
                                                "validity_link_hint": hints["v"], "output_balance_link_hint": hints["o"]})
        assert code == 200 and set(body) == {"proof", "validity_link_proof", "output_balance_link_proof"}, (path, body)
        proof = sv.decode_proof(body["proof"])
        ct.verify_singleprover_proof(circuit, s2, proof)
        s_comm = np.array(proof.wires_poly_comms[0], dtype=np.uint64)
        for field, hk, gid in (("validity_link_proof", "v", pub.PARTY_LINK), ("output_balance_link_proof", "o", pub.OUTPUT_LINK)):
            lay = GroupLayout(layouts[gid].alignment, layouts[gid].offset, layouts[gid].size)
            other = sv.decode_link_hint(hints[hk])
            assert verify_link_proof(other.linking_wire_comm, s_comm, sv.decode_link_proof(body[field]), lay, h, tau_h)
    assert service.stats["link_proofs"] == links0 + 2 * len(which)


def private_match_flow(client, service, g2):
    """The private match: both parties' INTENT AND BALANCE VALIDITY and OUTPUT BALANCE VALIDITY proofs
    (`ProofAndHintResponse`), then INTENT AND BALANCE PRIVATE SETTLEMENT carrying the four hints
    (`PrivateSettlementProofResponse`, api_types.rs:106-124, 249-266) — the sequence tests/test_gpu_service.py::
    test_prove_paths_over_http drives on the device (there with the validity requests in flight together)."""
    from host_circuits import private_settlement as ps
    h, tau_h = g2
    links0 = service.stats["link_proofs"]
    parties, _ = ps.create_witness_statement(seed=61)
    validity = [val.create_witness_statement(seed=70 + i, intent=parties[i].intent, balance=parties[i].input_balance) for i in (0, 1)]
    out_validity = [obv.create_witness_statement(80 + i, parties[i].output_balance) for i in (0, 1)]
    parties, statement = ps.create_witness_statement(
        seed=61, linked=[(validity[i][0].new_amount_public_share, validity[i][0].post_match_balance_shares,
                          out_validity[i][0].post_match_balance_shares) for i in (0, 1)])
    hints = {}
    for kind, path, C, insts in (("v", "/prove-intent-and-balance-validity", st.IntentAndBalanceValidityCircuit, validity),
                                 ("o", "/prove-output-balance-validity", st.OutputBalanceValidityCircuit, out_validity)):
        for i, (w_, s_) in enumerate(insts):
            code, body = client.send_request(path, {"statement": sv.to_json(s_), "witness": sv.to_json(w_)})
            assert code == 200 and set(body) == {"proof", "link_hint"}, (path, body)
            ct.verify_singleprover_proof(C, s_, sv.decode_proof(body["proof"]))
            hints[(kind, i)] = body["link_hint"]
    req = {"statement": sv.to_json(statement), "witness": sv.to_json(parties),
           "validity_link_hint_0": hints[("v", 0)], "validity_link_hint_1": hints[("v", 1)],
           "output_balance_link_hint_0": hints[("o", 0)], "output_balance_link_hint_1": hints[("o", 1)]}
    code, body = client.send_request("/prove-intent-and-balance-private-settlement", req)
    assert code == 200 and set(body) == {"proof", "validity_link_proof_0", "validity_link_proof_1",
                                         "output_balance_link_proof_0", "output_balance_link_proof_1"}, body
    S = st.IntentAndBalancePrivateSettlementCircuit
    sp = sv.decode_proof(body["proof"])
    ct.verify_singleprover_proof(S, statement, sp)
    layouts = S.get_circuit_layout()
    s_comm = np.array(sp.wires_poly_comms[0], dtype=np.uint64)
    for field, key, gid in (("validity_link_proof_0", ("v", 0), ps.PARTY_LINKS[0]), ("validity_link_proof_1", ("v", 1), ps.PARTY_LINKS[1]),
                            ("output_balance_link_proof_0", ("o", 0), ps.OUTPUT_LINKS[0]),
                            ("output_balance_link_proof_1", ("o", 1), ps.OUTPUT_LINKS[1])):
        lay = GroupLayout(layouts[gid].alignment, layouts[gid].offset, layouts[gid].size)
        other = sv.decode_link_hint(hints[key])
        assert verify_link_proof(other.linking_wire_comm, s_comm, sv.decode_link_proof(body[field]), lay, h, tau_h)
    # party 1's validity hint in party 0's slot: the group values differ, the prover refuses
    req_bad = dict(req, validity_link_hint_0=hints[("v", 1)])
    code, body = client.send_request("/prove-intent-and-balance-private-settlement", req_bad)
    assert code == 500 and "ProverError" in body["error"], body
    assert service.stats["link_proofs"] == links0 + 4
