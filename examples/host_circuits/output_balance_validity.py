"""OUTPUT BALANCE VALIDITY restated on the host-side constraint system — the proof that the balance a party RECEIVES
into exists in the Merkle state and is rotated to its next version, linked to the settlement proof through the
`output_balance_settlement_party{0,1}` groups.

`circuits-core/src/zk_circuits/validity_proofs/output_balance.rs:48-120`: the balance half of INTENT AND BALANCE
VALIDITY on its own (complementary shares, post-match fields re-encrypted, state rotation with a partial commitment).
Witness / statement construction follows the circuit's `test_helpers`.  5 public inputs; two link groups of 11 values.

Host-side input generation for tests and benches: the production circuit stays in Rust."""
from __future__ import annotations

import random
from dataclasses import dataclass
from typing import List

from . import circuit as cb
from .intent_and_balance_validity import BALANCE_PARTIAL_COMMITMENT_SIZE, _merkle_opening, _random_state_wrapper
from .private_settlement import OUTPUT_LINKS
from .valid_balance_create import DarkpoolBalance


@dataclass
class Witness:
    """output_balance.rs:124-133."""
    old_balance: cb.StateWrapper
    balance_opening: cb.MerkleOpening
    balance: DarkpoolBalance
    post_match_balance_shares: List[int]


@dataclass
class Statement:
    """output_balance.rs:141-148."""
    merkle_root: int
    old_balance_nullifier: int
    new_partial_commitment: tuple
    recovery_id: int

    def to_scalars(self) -> List[int]:
        return [self.merkle_root, self.old_balance_nullifier, *self.new_partial_commitment, self.recovery_id]


def create_witness_statement(seed: int, balance: DarkpoolBalance):
    rnd = random.Random(seed)
    old = _random_state_wrapper(balance.to_scalars(), rnd)
    nullifier = old.compute_nullifier()
    root, opening = _merkle_opening(old.compute_commitment(), rnd)
    new = old.clone()
    shares = new.share_stream.stream_cipher_encrypt([balance.relayer_fee_balance, balance.protocol_fee_balance, balance.amount])
    new.public_share[5:8] = shares
    recovery_id = new.compute_recovery_id()
    partial = new.compute_partial_commitment(BALANCE_PARTIAL_COMMITMENT_SIZE)
    return Witness(old, opening, balance, shares), Statement(root, nullifier, partial, recovery_id)


class OutputBalanceValidityCircuit:
    @staticmethod
    def name() -> str:
        return "Output Balance Validity (10)"

    @staticmethod
    def build(witness: Witness, statement: Statement, layouts: dict) -> cb.PlonkCircuit:
        cs = cb.PlonkCircuit()
        groups = list(OUTPUT_LINKS)
        for gid in groups:
            cs.create_link_group(gid, cb.GroupLayout(layouts[gid].alignment, layouts[gid].offset))
        old = cb.StateWrapperVar.create_witness(witness.old_balance, cs)
        opening = cb.MerkleOpeningVar([cs.create_variable(v) for v in witness.balance_opening.elems],
                                      [cs.create_boolean_variable(b) for b in witness.balance_opening.indices])
        balance = [cs.create_variable_with_link_groups(v, groups) for v in witness.balance.to_scalars()]
        post_match = [cs.create_variable_with_link_groups(v, groups) for v in witness.post_match_balance_shares]
        st = [cs.create_public_variable(v) for v in statement.to_scalars()]
        old_private = cb.ShareGadget.compute_complementary_shares(old.public_share, old.inner, cs)
        new, new_private = old.clone(), list(old_private)
        pads, ciphertexts = cb.StreamCipherGadget.encrypt(new.inner[5:8], new.share_stream, cs)
        new_private[5:8], new.public_share[5:8] = pads, ciphertexts
        for got, exp in zip(ciphertexts, post_match):
            cs.enforce_equal(got, exp)
        for a, b in zip(new.inner, balance):
            cs.enforce_equal(a, b)
        cb.StateElementRotationGadget.rotate_version_with_partial_commitment(
            BALANCE_PARTIAL_COMMITMENT_SIZE, old, old_private, opening, st[0], st[1], new, new_private, (st[2], st[3]), st[4], cs)
        return cs
