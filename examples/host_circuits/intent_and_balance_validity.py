"""INTENT AND BALANCE VALIDITY restated on the host-side constraint system — the validity proof each party of a
private settlement brings, linked to the settlement proof through the `intent_and_balance_settlement_party{0,1}` groups.

`circuits-core/src/zk_circuits/validity_proofs/intent_and_balance.rs:53-205` with `state_primitives/state_rotation.rs`,
`commitment.rs` (shared-prefix and partial commitments), `nullifier.rs`, `shares.rs`, `stream_cipher.rs`: the intent and
the balance a trader will use exist in the Merkle state (openings of their commitments), are nullified, and their next
versions — amount / post-match fields re-encrypted with the share streams — are (partially) committed and given recovery
ids.  Witness / statement construction follows the circuit's `test_helpers` (:338-407).  8 public inputs... see
`Statement.to_scalars`; two link groups of 17 values each holding the same variables.

Host-side input generation for tests and benches: the production circuit stays in Rust."""
from __future__ import annotations

import random
from dataclasses import dataclass
from typing import List

from . import circuit as cb
from .circuit import R
from .private_settlement import PARTY_LINKS, Intent
from .valid_balance_create import DarkpoolBalance

MERKLE_HEIGHT = 10                       # crates/constants/src/lib.rs:50
INTENT_PARTIAL_COMMITMENT_SIZE = 4       # IntentShare::NUM_SCALARS - 1 (intent_and_balance.rs:45)
BALANCE_PARTIAL_COMMITMENT_SIZE = 5      # DarkpoolBalance::NUM_SCALARS - PostMatchBalance::NUM_SCALARS (first_fill.rs:53)


@dataclass
class Witness:
    """intent_and_balance.rs:213-232."""
    old_intent: cb.StateWrapper
    old_intent_opening: cb.MerkleOpening
    intent: Intent
    new_amount_public_share: int
    old_balance: cb.StateWrapper
    old_balance_opening: cb.MerkleOpening
    balance: DarkpoolBalance
    post_match_balance_shares: List[int]      # relayer fee balance, protocol fee balance, amount


@dataclass
class Statement:
    """intent_and_balance.rs:240-255, fields in public-input order (a partial commitment is two scalars)."""
    intent_merkle_root: int
    old_intent_nullifier: int
    new_intent_partial_commitment: tuple
    intent_recovery_id: int
    balance_merkle_root: int
    old_balance_nullifier: int
    balance_partial_commitment: tuple
    balance_recovery_id: int

    def to_scalars(self) -> List[int]:
        return [self.intent_merkle_root, self.old_intent_nullifier, *self.new_intent_partial_commitment,
                self.intent_recovery_id, self.balance_merkle_root, self.old_balance_nullifier,
                *self.balance_partial_commitment, self.balance_recovery_id]


def _random_state_wrapper(inner: List[int], rnd: random.Random) -> cb.StateWrapper:
    """A state element some way into its life: both streams seeded at random, the recovery stream already advanced
    (a nullifier needs a previous recovery id)."""
    w = cb.StateWrapper.new(inner, rnd.randrange(R), rnd.randrange(R))
    w.recovery_stream.index = rnd.randrange(1, 1 << 20)
    return w


def _merkle_opening(leaf_hash: int, rnd: random.Random):
    opening = cb.MerkleOpening([rnd.randrange(R) for _ in range(MERKLE_HEIGHT)],
                               [rnd.random() < 0.5 for _ in range(MERKLE_HEIGHT)])
    return cb.native_merkle_root_prehashed(leaf_hash, opening), opening


def create_witness_statement(seed: int = 0, intent: Intent = None, balance: DarkpoolBalance = None):
    """intent_and_balance.rs:356-407 with a seeded RNG (optionally for a given intent and matching balance)."""
    rnd = random.Random(seed)
    addr = lambda: rnd.randrange(1 << 160)
    if intent is None:
        intent = Intent(addr(), addr(), addr(), rnd.randrange(1 << 100), rnd.randrange(1, 1 << 90))
    if balance is None:
        balance = DarkpoolBalance(intent.in_token, intent.owner, addr(), rnd.randrange(R), rnd.randrange(R),
                                  rnd.randrange(1 << 40), rnd.randrange(1 << 40), intent.amount_in + rnd.randrange(1 << 40))
    old_intent = _random_state_wrapper(intent.to_scalars(), rnd)
    intent_root, intent_opening = _merkle_opening(old_intent.compute_commitment(), rnd)
    intent_nullifier = old_intent.compute_nullifier()
    new_intent = old_intent.clone()
    new_amount_public_share = new_intent.share_stream.stream_cipher_encrypt([intent.amount_in])[0]
    new_intent.public_share[4] = new_amount_public_share
    intent_recovery_id = new_intent.compute_recovery_id()
    intent_partial = new_intent.compute_partial_commitment(INTENT_PARTIAL_COMMITMENT_SIZE)

    old_balance = _random_state_wrapper(balance.to_scalars(), rnd)
    balance_nullifier = old_balance.compute_nullifier()
    balance_root, balance_opening = _merkle_opening(old_balance.compute_commitment(), rnd)
    new_balance = old_balance.clone()
    post_match = [balance.relayer_fee_balance, balance.protocol_fee_balance, balance.amount]
    post_match_shares = new_balance.share_stream.stream_cipher_encrypt(post_match)
    new_balance.public_share[5:8] = post_match_shares
    balance_recovery_id = new_balance.compute_recovery_id()
    balance_partial = new_balance.compute_partial_commitment(BALANCE_PARTIAL_COMMITMENT_SIZE)

    witness = Witness(old_intent, intent_opening, intent, new_amount_public_share, old_balance, balance_opening, balance,
                      post_match_shares)
    return witness, Statement(intent_root, intent_nullifier, intent_partial, intent_recovery_id, balance_root,
                              balance_nullifier, balance_partial, balance_recovery_id)


class IntentAndBalanceValidityCircuit:
    @staticmethod
    def name() -> str:
        return f"Intent And Balance Validity ({MERKLE_HEIGHT})"

    @staticmethod
    def build(witness: Witness, statement: Statement, layouts: dict) -> cb.PlonkCircuit:
        """`layouts`: the settlement circuit's `get_circuit_layout()` entries for the two party groups
        (`proof_linking_groups`, :270-279: the validity circuit places its link values where the settlement circuit
        has them, and in BOTH groups, so one validity proof serves either side of a match)."""
        cs = cb.PlonkCircuit()
        groups = list(PARTY_LINKS)
        for gid in groups:
            cs.create_link_group(gid, cb.GroupLayout(layouts[gid].alignment, layouts[gid].offset))
        opening_var = lambda o: cb.MerkleOpeningVar([cs.create_variable(v) for v in o.elems],
                                                    [cs.create_boolean_variable(b) for b in o.indices])
        linked = lambda vals: [cs.create_variable_with_link_groups(v, groups) for v in vals]
        old_intent = cb.StateWrapperVar.create_witness(witness.old_intent, cs)
        old_intent_opening = opening_var(witness.old_intent_opening)
        intent = linked(witness.intent.to_scalars())
        new_amount_public_share = linked([witness.new_amount_public_share])[0]
        old_balance = cb.StateWrapperVar.create_witness(witness.old_balance, cs)
        old_balance_opening = opening_var(witness.old_balance_opening)
        balance = linked(witness.balance.to_scalars())
        post_match_balance_shares = linked(witness.post_match_balance_shares)
        st = [cs.create_public_variable(v) for v in statement.to_scalars()]

        # ---- intent (:82-140) ---------------------------------------------------------------------------
        for a, b in zip(intent, old_intent.inner):
            cs.enforce_equal(a, b)
        old_private = cb.ShareGadget.compute_complementary_shares(old_intent.public_share, old_intent.inner, cs)
        new_intent, new_private = old_intent.clone(), list(old_private)
        pads, ciphertexts = cb.StreamCipherGadget.encrypt([new_intent.inner[4]], new_intent.share_stream, cs)
        new_private[4], new_intent.public_share[4] = pads[0], ciphertexts[0]
        cs.enforce_equal(ciphertexts[0], new_amount_public_share)
        cb.StateElementRotationGadget.rotate_version_with_partial_commitment(
            INTENT_PARTIAL_COMMITMENT_SIZE, old_intent, old_private, old_intent_opening, st[0], st[1],
            new_intent, new_private, (st[2], st[3]), st[4], cs)
        # ---- balance (:144-205) -------------------------------------------------------------------------
        for a, b in zip(balance, old_balance.inner):
            cs.enforce_equal(a, b)
        old_private = cb.ShareGadget.compute_complementary_shares(old_balance.public_share, old_balance.inner, cs)
        new_balance, new_private = old_balance.clone(), list(old_private)
        pads, ciphertexts = cb.StreamCipherGadget.encrypt(new_balance.inner[5:8], new_balance.share_stream, cs)
        new_private[5:8], new_balance.public_share[5:8] = pads, ciphertexts
        for got, exp in zip(ciphertexts, post_match_balance_shares):
            cs.enforce_equal(got, exp)
        cb.StateElementRotationGadget.rotate_version_with_partial_commitment(
            BALANCE_PARTIAL_COMMITMENT_SIZE, old_balance, old_private, old_balance_opening, st[5], st[6],
            new_balance, new_private, (st[7], st[8]), st[9], cs)
        # ---- cross constraints (:70-74): the intent sells what the balance holds, same owner ---------------
        cs.enforce_equal(intent[0], balance[0])
        cs.enforce_equal(intent[2], balance[1])
        return cs
