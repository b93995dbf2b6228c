"""CPU test of the collaborative prover's protocol logic (renegade_b200/collaborative.py) with a host backend built on
the oracle: for 2 and 3 parties holding additive shares of the witness table and of the 17 blinders, the OPENED
collaborative proof is byte-identical to the oracle's single-prover proof of the summed witness with the summed
blinders — every commitment, evaluation and the linking hint; an unsatisfied shared witness is refused.  (Mirrors
`multiprover_prove_and_verify`, circuits-core/src/lib.rs:166-177: prove jointly, open, verify as a normal proof.)"""
import numpy as np
import pytest

from host_backend import HostBackend, R
from renegade_b200 import collaborative as co
from renegade_b200 import synth
from renegade_b200.fields import limbs_to_scalars, scalars_to_limbs

TAU = 0xB200B200B200


@pytest.mark.parametrize("log_n,parties", [(3, 2), (5, 2), (5, 3), (6, 2)])
def test_opened_collaborative_proof_equals_single_prover_proof(oracle, pyoracle, log_n, parties):
    py = pyoracle
    n = 1 << log_n
    circ = synth.synth_circuit(log_n, num_inputs=3 if log_n > 3 else 1, seed=100 + log_n, check=True)
    tau = oracle.int_to_limbs(py.to_mont(TAU % py.R, py.R))
    srs = oracle.srs_from_tau(tau, n + 3)
    bl = synth.splitmix_blinders(40 + log_n)
    opk = oracle.plonk_preprocess(log_n, circ.selectors, circ.perm, circ.k, srs)
    rc, oproof, _, olink = oracle.plonk_prove(log_n, circ.num_inputs, circ.k, opk, circ.wires, circ.pub_inputs, bl, srs, True)
    assert rc == 0
    be = HostBackend(srs)
    pk = co.CollaborativeProvingKey.build(be, log_n, circ.num_inputs, circ.selectors, circ.perm, circ.k)
    assert (pk.selector_comms == opk["selector_comms"]).all() and (pk.sigma_comms == opk["sigma_comms"]).all()
    wire_shares = co.share_table(np.asarray(circ.wires, dtype=np.uint64).reshape(-1, 4), parties, seed=7)
    blinder_shares = co.share_table(np.asarray(bl, dtype=np.uint64).reshape(-1, 4), parties, seed=8)
    # no single share is the witness
    assert all((s != np.asarray(circ.wires, dtype=np.uint64).reshape(-1, 4)).any() for s in wire_shares)
    proof, hint, fab = co.MultiproverPlonkKzgSnark.prove_with_link_hint(be, pk, wire_shares, circ.pub_inputs, blinder_shares)
    assert bytes(proof) == bytes(oproof)
    assert (hint.linking_wire_poly == olink).all()
    assert oracle.plonk_verify_known_tau(log_n, circ.num_inputs, circ.k, opk, circ.pub_inputs,
                                         oracle.PlonkProof.from_buffer_copy(bytes(proof)), tau)
    # the multiplications of rounds 2 and 3 went through Beaver triples: 4 + 2 + 2 + 1 + 1 + log n products of n, 26 of 8n
    assert fab.multiplications >= 26 * 8 * n and fab.opened_elements >= 2 * fab.multiplications


def test_unsatisfied_shared_witness_is_refused(oracle, pyoracle):
    from renegade_b200._lib import B200Error
    py = pyoracle
    log_n, n = 4, 16
    circ = synth.synth_circuit(log_n, num_inputs=2, seed=77, check=True)
    tau = oracle.int_to_limbs(py.to_mont(TAU % py.R, py.R))
    srs = oracle.srs_from_tau(tau, n + 3)
    be = HostBackend(srs)
    pk = co.CollaborativeProvingKey.build(be, log_n, circ.num_inputs, circ.selectors, circ.perm, circ.k)
    wires = np.asarray(circ.wires, dtype=np.uint64).reshape(-1, 4).copy()
    shares = co.share_table(wires, 2, seed=1)
    bad = limbs_to_scalars(shares[1])
    bad[3] = (bad[3] + 1) % R                      # one party's share of one wire value is off by one
    shares[1] = scalars_to_limbs(bad)
    bsh = co.share_table(np.asarray(synth.splitmix_blinders(5), dtype=np.uint64).reshape(-1, 4), 2, seed=2)
    with pytest.raises(B200Error) as err:
        co.MultiproverPlonkKzgSnark.prove_with_link_hint(be, pk, shares, circ.pub_inputs, bsh)
    assert err.value.code == -7
