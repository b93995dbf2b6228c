"""GPU test of the prover service end to end: a client replays the requests the reference's
`ProofServiceClient` would send (prover_service_client.rs:186-205: POST JSON, basic auth `admin`), the service proves
through the pool on the device, and the returned proofs / link proofs verify with the pairing against the reference's
own SRS."""
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_prove_paths_over_http(srs_2_16, g2_raw):
    import renegade_b200 as rb
    from host_circuits import intent_and_balance_validity as val
    from host_circuits import output_balance_validity as obv
    from host_circuits import private_settlement as ps
    from host_circuits import service_routes
    from host_circuits import statements as st
    from host_circuits import valid_balance_create as vbc
    from renegade_b200 import circuit_types as ct
    from renegade_b200 import service as sv
    from renegade_b200.backend import GroupLayout, ProverPool, verify_link_proof
    h, tau_h = g2_raw
    pool = ProverPool(0, workers=4)
    ctx = pool.context(0)
    params = rb.parse_ptau_file(ctx, srs_2_16, count=(1 << 14) + 3)
    ct.set_system_srs(ctx, params.powers_of_g, h, tau_h, pool=pool)
    service = sv.ProverService(service_routes.routes(), password="pw", pool=pool)
    server = service.make_server("127.0.0.1", 0)
    threading.Thread(target=server.serve_forever, daemon=True).start()
    try:
        client = sv.ProofServiceClient(f"http://127.0.0.1:{server.server_address[1]}", "pw")
        # VALID BALANCE CREATE -> ProofResponse
        w, s = vbc.create_witness_statement(21)
        code, body = client.send_request("/prove-valid-balance-create", {"statement": sv.to_json(s), "witness": sv.to_json(w)})
        assert code == 200 and set(body) == {"proof"}
        ct.verify_singleprover_proof(st.ValidBalanceCreate, s, sv.decode_proof(body["proof"]))
        # an unsatisfied instance comes back as a prover error, and the service keeps serving
        _, s_other = vbc.create_witness_statement(22)
        code, body = client.send_request("/prove-valid-balance-create", {"statement": sv.to_json(s_other), "witness": sv.to_json(w)})
        assert code == 500 and "ProverError" in body["error"]
        # the private match: validity + output-balance proofs (ProofAndHintResponse), then the settlement request carrying
        # their hints (PrivateSettlementProofResponse)
        parties, _ = ps.create_witness_statement(seed=61)
        validity = [val.create_witness_statement(seed=70 + i, intent=parties[i].intent, balance=parties[i].input_balance)
                    for i in (0, 1)]
        out_validity = [obv.create_witness_statement(80 + i, parties[i].output_balance) for i in (0, 1)]
        parties, statement = ps.create_witness_statement(
            seed=61, linked=[(validity[i][0].new_amount_public_share, validity[i][0].post_match_balance_shares,
                              out_validity[i][0].post_match_balance_shares) for i in (0, 1)])
        results = {}

        def ask(key, path, w_, s_):
            results[key] = client.send_request(path, {"statement": sv.to_json(s_), "witness": sv.to_json(w_)})
        threads = [threading.Thread(target=ask, args=(("v", i), "/prove-intent-and-balance-validity", *validity[i])) for i in (0, 1)] + \
                  [threading.Thread(target=ask, args=(("o", i), "/prove-output-balance-validity", *out_validity[i])) for i in (0, 1)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        hints = {}
        for (kind, i), (code, body) in results.items():
            assert code == 200 and set(body) == {"proof", "link_hint"}
            C, (_, s_) = (st.IntentAndBalanceValidityCircuit, validity[i]) if kind == "v" else (st.OutputBalanceValidityCircuit, out_validity[i])
            ct.verify_singleprover_proof(C, s_, sv.decode_proof(body["proof"]))
            hints[(kind, i)] = body["link_hint"]
        req = {"statement": sv.to_json(statement), "witness": sv.to_json(parties),
               "validity_link_hint_0": hints[("v", 0)], "validity_link_hint_1": hints[("v", 1)],
               "output_balance_link_hint_0": hints[("o", 0)], "output_balance_link_hint_1": hints[("o", 1)]}
        code, body = client.send_request("/prove-intent-and-balance-private-settlement", req)
        assert code == 200 and set(body) == {"proof", "validity_link_proof_0", "validity_link_proof_1",
                                             "output_balance_link_proof_0", "output_balance_link_proof_1"}
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
        assert service.stats["proofs"] >= 6 and service.stats["link_proofs"] == 4
    finally:
        server.shutdown()
        ct.clear_key_cache()
        pool.close()


def test_state_update_paths_over_http(srs_2_16, g2_raw):
    """VALID DEPOSIT, VALID WITHDRAWAL, VALID ORDER CANCELLATION, VALID NOTE REDEMPTION and the two PUBLIC fee payments
    (`ProofResponse` paths, prover_service_client.rs:101-147) through the service: the proofs verify with the pairing under the circuits' cached keys; a request whose
    statement does not belong to its witness comes back as a prover error."""
    import renegade_b200 as rb
    from host_circuits import service_routes
    from host_circuits import state_updates as su
    from host_circuits import statements as st
    from renegade_b200 import circuit_types as ct
    from renegade_b200 import service as sv
    from renegade_b200.backend import ProverPool
    h, tau_h = g2_raw
    pool = ProverPool(0, workers=3)
    ctx = pool.context(0)
    params = rb.parse_ptau_file(ctx, srs_2_16, count=(1 << 14) + 3)
    ct.set_system_srs(ctx, params.powers_of_g, h, tau_h, pool=pool)
    service = sv.ProverService(service_routes.routes(), password="pw", pool=pool)
    server = service.make_server("127.0.0.1", 0)
    threading.Thread(target=server.serve_forever, daemon=True).start()
    try:
        client = sv.ProofServiceClient(f"http://127.0.0.1:{server.server_address[1]}", "pw")
        from host_circuits import fees
        for path, circuit, make in (("/prove-valid-deposit", st.ValidDeposit, su.create_deposit_witness_statement),
                                    ("/prove-valid-withdrawal", st.ValidWithdrawal, su.create_withdrawal_witness_statement),
                                    ("/prove-valid-order-cancellation", st.ValidOrderCancellationCircuit,
                                     su.create_cancellation_witness_statement),
                                    ("/prove-valid-note-redemption", st.ValidNoteRedemption, fees.create_note_redemption_witness_statement),
                                    ("/prove-valid-public-protocol-fee-payment", st.ValidPublicProtocolFeePayment,
                                     fees.create_public_protocol_fee_payment_witness_statement),
                                    ("/prove-valid-public-relayer-fee-payment", st.ValidPublicRelayerFeePayment,
                                     fees.create_public_relayer_fee_payment_witness_statement)):
            w, s = make(41)
            code, body = client.send_request(path, {"statement": sv.to_json(s), "witness": sv.to_json(w)})
            assert code == 200 and set(body) == {"proof"}, (path, body)
            ct.verify_singleprover_proof(circuit, s, sv.decode_proof(body["proof"]))
            _, s_other = make(42)
            code, body = client.send_request(path, {"statement": sv.to_json(s_other), "witness": sv.to_json(w)})
            assert code == 500 and "ProverError" in body["error"], path
        assert service.stats["proofs"] == 6
    finally:
        server.shutdown()
        ct.clear_key_cache()
        pool.close()


def test_intent_only_paths_over_http(srs_2_16, g2_raw):
    """The intent-only flow as the relayer drives it (prover_service_client.rs:101-147): INTENT ONLY VALIDITY and FIRST FILL
    VALIDITY (`ProofAndHintResponse`), then INTENT ONLY PUBLIC / BOUNDED SETTLEMENT with the validity proof's hint in the
    request (`SettlementProofResponse {proof, link_proof}`, api_types.rs:97-104, 279-299).  Proofs verify with the
    pairing under the cached keys; the link proofs verify against the two proofs' first wire commitments on the
    group the public settlement circuit places."""
    import renegade_b200 as rb
    from host_circuits import intent_only as io
    from host_circuits import service_routes
    from host_circuits import statements as st
    from renegade_b200 import circuit_types as ct
    from renegade_b200 import service as sv
    from renegade_b200.backend import GroupLayout, ProverPool, verify_link_proof
    h, tau_h = g2_raw
    pool = ProverPool(0, workers=3)
    ctx = pool.context(0)
    params = rb.parse_ptau_file(ctx, srs_2_16, count=(1 << 14) + 3)
    ct.set_system_srs(ctx, params.powers_of_g, h, tau_h, pool=pool)
    service = sv.ProverService(service_routes.routes(), password="pw", pool=pool)
    server = service.make_server("127.0.0.1", 0)
    threading.Thread(target=server.serve_forever, daemon=True).start()
    try:
        client = sv.ProofServiceClient(f"http://127.0.0.1:{server.server_address[1]}", "pw")
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
        assert service.stats["proofs"] == 4 and service.stats["link_proofs"] == 2
    finally:
        server.shutdown()
        ct.clear_key_cache()
        pool.close()


def test_public_settlement_paths_over_http(srs_2_16, g2_raw):
    """INTENT AND BALANCE PUBLIC / BOUNDED SETTLEMENT as the relayer drives them: the party's INTENT AND BALANCE VALIDITY and
    OUTPUT BALANCE VALIDITY proofs first, then the settlement request with both hints
    (`PublicSettlementProofResponse {proof, validity_link_proof, output_balance_link_proof}`, api_types.rs:126-137, 238-277);
    proofs and link proofs verify with the pairing."""
    import renegade_b200 as rb
    from host_circuits import intent_and_balance_validity as val
    from host_circuits import output_balance_validity as obv
    from host_circuits import public_settlement as pub
    from host_circuits import service_routes
    from host_circuits import statements as st
    from renegade_b200 import circuit_types as ct
    from renegade_b200 import service as sv
    from renegade_b200.backend import GroupLayout, ProverPool, verify_link_proof
    h, tau_h = g2_raw
    pool = ProverPool(0, workers=3)
    ctx = pool.context(0)
    params = rb.parse_ptau_file(ctx, srs_2_16, count=(1 << 14) + 3)
    ct.set_system_srs(ctx, params.powers_of_g, h, tau_h, pool=pool)
    service = sv.ProverService(service_routes.routes(), password="pw", pool=pool)
    server = service.make_server("127.0.0.1", 0)
    threading.Thread(target=server.serve_forever, daemon=True).start()
    try:
        client = sv.ProofServiceClient(f"http://127.0.0.1:{server.server_address[1]}", "pw")
        layouts = st.IntentAndBalancePrivateSettlementCircuit.get_circuit_layout()
        for path, circuit, make, Statement in (
                ("/prove-intent-and-balance-public-settlement", st.IntentAndBalancePublicSettlementCircuit,
                 pub.create_public_witness_statement, pub.PublicStatement),
                ("/prove-intent-and-balance-bounded-settlement", st.IntentAndBalanceBoundedSettlementCircuit,
                 pub.create_bounded_witness_statement, pub.BoundedStatement)):
            w, s = make(61)
            vw, vs = val.create_witness_statement(62, intent=w.intent, balance=w.in_balance)
            ow, os_ = obv.create_witness_statement(63, w.out_balance)
            hints = {}
            for key, vpath, (w_, s_), C in (("v", "/prove-intent-and-balance-validity", (vw, vs), st.IntentAndBalanceValidityCircuit),
                                           ("o", "/prove-output-balance-validity", (ow, os_), st.OutputBalanceValidityCircuit)):
                code, body = client.send_request(vpath, {"statement": sv.to_json(s_), "witness": sv.to_json(w_)})
                assert code == 200 and set(body) == {"proof", "link_hint"}, (vpath, body)
                ct.verify_singleprover_proof(C, s_, sv.decode_proof(body["proof"]))
                hints[key] = body["link_hint"]
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
            for field, key, gid in (("validity_link_proof", "v", pub.PARTY_LINK), ("output_balance_link_proof", "o", pub.OUTPUT_LINK)):
                lay = GroupLayout(layouts[gid].alignment, layouts[gid].offset, layouts[gid].size)
                other = sv.decode_link_hint(hints[key])
                assert verify_link_proof(other.linking_wire_comm, s_comm, sv.decode_link_proof(body[field]), lay, h, tau_h)
        assert service.stats["link_proofs"] == 4
    finally:
        server.shutdown()
        ct.clear_key_cache()
        pool.close()
