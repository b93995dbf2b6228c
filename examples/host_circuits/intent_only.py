"""The intent-only family restated on the host-side constraint system — intents capitalised by a PUBLIC balance, settled
in the open; one link group, `intent_only_settlement`, placed by the public settlement circuit and inherited by the rest:

* INTENT ONLY PUBLIC SETTLEMENT   `settlement/intent_only_public_settlement.rs:35-60`   — the obligation respects the intent;
* INTENT ONLY BOUNDED SETTLEMENT  `settlement/intent_only_bounded_settlement.rs:39-58`  — the same for a bounded match result
  (`settlement_lib.rs: BoundedSettlementGadget::verify_intent_constraints`);
* INTENT ONLY VALIDITY            `validity_proofs/intent_only.rs:64-145`                — the intent exists in the Merkle state, is
  nullified, its amount share re-encrypted (leaked at once: the settlement is public), its next version partially committed;
* INTENT ONLY FIRST FILL VALIDITY `validity_proofs/intent_only_first_fill.rs:46-103`    — a new intent: valid amount and price, its
  public shares, first recovery id and the commitment to its private shares.

Witness / statement construction follows each circuit's `test_helpers` with a seeded RNG.
Host-side input generation for tests, the prover service and benches: the production circuits stay in Rust."""
from __future__ import annotations

import random
from dataclasses import dataclass
from typing import List

from . import circuit as cb
from .circuit import R
from .private_settlement import DEFAULT_FP_PRECISION, FixedPointGadget, Intent, SettlementObligation

MERKLE_HEIGHT = 10
INTENT_ONLY_SETTLEMENT_LINK = "intent_only_settlement"          # settlement/mod.rs
PARTIAL_COMMITMENT_SIZE = 4                                      # IntentShare::NUM_SCALARS - 1 (intent_only.rs:48)
PRICE_BITS = DEFAULT_FP_PRECISION + 64                           # circuit-types/src/lib.rs:64
AMOUNT_IN_IDX, OWNER_IDX = 4, 2                                  # positions in Intent's scalar serialisation
# where INTENT ONLY PUBLIC SETTLEMENT places the group (its own domain is 2^9): rows 16 .. 20 of a 2^9 grid
DEFAULT_LAYOUT = cb.GroupLayout(9, 16)


def random_intent(rnd: random.Random) -> Intent:
    addr = lambda: rnd.randrange(1 << 160)
    return Intent(addr(), addr(), addr(), rnd.randrange(1, 1 << 100), rnd.randrange(1, 1 << 90))


def _intent_constraints(intent, obligation, cs) -> None:
    """settlement_lib.rs:45-75 `SettlementGadget::verify_intent_constraints`."""
    in_token, out_token, amount_in, amount_out = obligation
    cs.enforce_equal(in_token, intent[0])
    cs.enforce_equal(out_token, intent[1])
    cb.GreaterThanEqGadget.constrain_greater_than_eq(intent[AMOUNT_IN_IDX], amount_in, cb.AMOUNT_BITS, cs)
    min_output = FixedPointGadget.floor(FixedPointGadget.mul_integer(intent[3], amount_in, cs), cs)
    cb.GreaterThanEqGadget.constrain_greater_than_eq(amount_out, min_output, cb.AMOUNT_BITS, cs)


# ---- INTENT ONLY PUBLIC SETTLEMENT ----------------------------------------------------------------------------------
@dataclass
class PublicSettlementWitness:
    intent: Intent


@dataclass
class PublicSettlementStatement:
    """intent_only_public_settlement.rs:78-100."""
    settlement_obligation: SettlementObligation
    relayer_fee: int
    relayer_fee_recipient: int

    def to_scalars(self) -> List[int]:
        return self.settlement_obligation.to_scalars() + [self.relayer_fee, self.relayer_fee_recipient]


def create_settlement_obligation(intent: Intent, rnd: random.Random) -> SettlementObligation:
    """test_helpers `create_settlement_obligation`: part of the intent's amount at a price the intent accepts."""
    amount_in = rnd.randrange(1, intent.amount_in + 1)
    min_out = (intent.min_price * amount_in) >> DEFAULT_FP_PRECISION
    return SettlementObligation(intent.in_token, intent.out_token, amount_in, min(min_out + rnd.randrange(1 << 20), (1 << cb.AMOUNT_BITS) - 1))


def create_public_settlement_witness_statement(seed: int = 0, intent: Intent = None):
    rnd = random.Random(seed)
    if intent is None:  # keep min_price * amount_in inside the amount range
        intent = Intent(rnd.randrange(1 << 160), rnd.randrange(1 << 160), rnd.randrange(1 << 160),
                        rnd.randrange(1, 1 << (DEFAULT_FP_PRECISION + 20)), rnd.randrange(1, 1 << 70))
    return (PublicSettlementWitness(intent),
            PublicSettlementStatement(create_settlement_obligation(intent, rnd), rnd.randrange(1 << 50), rnd.randrange(1 << 160)))


class IntentOnlyPublicSettlementCircuit:
    @staticmethod
    def name() -> str:
        return "Intent Only Public Settlement"

    @staticmethod
    def build(witness: PublicSettlementWitness, statement: PublicSettlementStatement, layout: cb.GroupLayout = None) -> cb.PlonkCircuit:
        cs = cb.PlonkCircuit()
        cs.create_link_group(INTENT_ONLY_SETTLEMENT_LINK, layout or DEFAULT_LAYOUT)
        intent = [cs.create_variable_with_link_groups(v, [INTENT_ONLY_SETTLEMENT_LINK]) for v in witness.intent.to_scalars()]
        st = [cs.create_public_variable(v) for v in statement.to_scalars()]
        _intent_constraints(intent, st[0:4], cs)  # the fee and its recipient are bound by the transcript only (:84-99)
        return cs


# ---- INTENT ONLY BOUNDED SETTLEMENT ---------------------------------------------------------------------------------
@dataclass
class BoundedMatchResult:
    """darkpool-types/src/bounded_match_result.rs:41-58."""
    internal_party_input_token: int
    internal_party_output_token: int
    min_internal_party_amount_in: int
    max_internal_party_amount_in: int
    price: int
    block_deadline: int

    def to_scalars(self) -> List[int]:
        return [self.internal_party_input_token, self.internal_party_output_token, self.min_internal_party_amount_in,
                self.max_internal_party_amount_in, self.price, self.block_deadline]


@dataclass
class BoundedSettlementStatement:
    """intent_only_bounded_settlement.rs:78-92."""
    bounded_match_result: BoundedMatchResult
    internal_relayer_fee: int
    external_relayer_fee: int
    relayer_fee_recipient: int

    def to_scalars(self) -> List[int]:
        return self.bounded_match_result.to_scalars() + [self.internal_relayer_fee, self.external_relayer_fee,
                                                         self.relayer_fee_recipient]


def create_bounded_settlement_witness_statement(seed: int = 0, intent: Intent = None):
    rnd = random.Random(seed)
    intent = intent or random_intent(rnd)
    hi = rnd.randrange(1, intent.amount_in + 1)
    bmr = BoundedMatchResult(intent.in_token, intent.out_token, rnd.randrange(0, hi + 1), hi,
                             intent.min_price + rnd.randrange(1 << 30), rnd.randrange(1 << 40))
    return (PublicSettlementWitness(intent),
            BoundedSettlementStatement(bmr, rnd.randrange(1 << 50), rnd.randrange(1 << 50), rnd.randrange(1 << 160)))


class IntentOnlyBoundedSettlementCircuit:
    @staticmethod
    def name() -> str:
        return "Intent Only Bounded Settlement"

    @staticmethod
    def build(witness: PublicSettlementWitness, statement: BoundedSettlementStatement, layout: cb.GroupLayout = None) -> cb.PlonkCircuit:
        cs = cb.PlonkCircuit()
        cs.create_link_group(INTENT_ONLY_SETTLEMENT_LINK, layout or DEFAULT_LAYOUT)
        intent = [cs.create_variable_with_link_groups(v, [INTENT_ONLY_SETTLEMENT_LINK]) for v in witness.intent.to_scalars()]
        st = [cs.create_public_variable(v) for v in statement.to_scalars()]
        in_token, out_token, _min_in, max_in, price, _deadline = st[0:6]
        # settlement_lib.rs `BoundedSettlementGadget::verify_intent_constraints`: tokens, upper bound <= intent amount,
        # price >= the intent's worst case (min <= max <= amount is the contracts' check)
        cs.enforce_equal(in_token, intent[0])
        cs.enforce_equal(out_token, intent[1])
        cb.GreaterThanEqGadget.constrain_greater_than_eq(intent[AMOUNT_IN_IDX], max_in, cb.AMOUNT_BITS, cs)
        cb.GreaterThanEqGadget.constrain_greater_than_eq(price, intent[3], PRICE_BITS, cs)
        return cs


# ---- INTENT ONLY VALIDITY -------------------------------------------------------------------------------------------
@dataclass
class ValidityWitness:
    """intent_only.rs:153-166."""
    old_intent: cb.StateWrapper
    old_intent_opening: cb.MerkleOpening
    intent: Intent


@dataclass
class ValidityStatement:
    """intent_only.rs:176-207, fields in public-input order (a partial commitment is two scalars)."""
    owner: int
    merkle_root: int
    old_intent_nullifier: int
    new_amount_public_share: int
    new_intent_partial_commitment: tuple
    recovery_id: int

    def to_scalars(self) -> List[int]:
        return [self.owner, self.merkle_root, self.old_intent_nullifier, self.new_amount_public_share,
                *self.new_intent_partial_commitment, self.recovery_id]


def create_validity_witness_statement(seed: int = 0, intent: Intent = None):
    """intent_only.rs:275-327."""
    rnd = random.Random(seed)
    intent = intent or random_intent(rnd)
    old_intent = cb.StateWrapper.new(intent.to_scalars(), rnd.randrange(R), rnd.randrange(R))
    old_intent.recovery_stream.index = rnd.randrange(1, 1 << 20)
    nullifier = old_intent.compute_nullifier()
    opening = cb.MerkleOpening([rnd.randrange(R) for _ in range(MERKLE_HEIGHT)], [rnd.random() < 0.5 for _ in range(MERKLE_HEIGHT)])
    root = cb.native_merkle_root_prehashed(old_intent.compute_commitment(), opening)
    new_intent = old_intent.clone()
    new_share = new_intent.share_stream.stream_cipher_encrypt([intent.amount_in])[0]
    new_intent.public_share[AMOUNT_IN_IDX] = new_share
    recovery_id = new_intent.compute_recovery_id()
    partial = new_intent.compute_partial_commitment(PARTIAL_COMMITMENT_SIZE)
    return (ValidityWitness(old_intent, opening, intent),
            ValidityStatement(intent.owner, root, nullifier, new_share, partial, recovery_id))


class IntentOnlyValidityCircuit:
    @staticmethod
    def name() -> str:
        return f"Intent Only Validity ({MERKLE_HEIGHT})"

    @staticmethod
    def build(witness: ValidityWitness, statement: ValidityStatement, layout: cb.GroupLayout = None) -> cb.PlonkCircuit:
        cs = cb.PlonkCircuit()
        cs.create_link_group(INTENT_ONLY_SETTLEMENT_LINK, layout or DEFAULT_LAYOUT)
        old_intent = cb.StateWrapperVar.create_witness(witness.old_intent, cs)
        opening = cb.MerkleOpeningVar([cs.create_variable(v) for v in witness.old_intent_opening.elems],
                                      [cs.create_boolean_variable(b) for b in witness.old_intent_opening.indices])
        intent = [cs.create_variable_with_link_groups(v, [INTENT_ONLY_SETTLEMENT_LINK]) for v in witness.intent.to_scalars()]
        st = [cs.create_public_variable(v) for v in statement.to_scalars()]
        owner, merkle_root, nullifier, new_amount_share, pc_private, pc_public, recovery_id = st
        # 1. old private shares, 2. the new intent: same fields, `amount_in` re-encrypted (:70-84, 122-145)
        old_private = cb.ShareGadget.compute_complementary_shares(old_intent.public_share, old_intent.inner, cs)
        new_intent, new_private = old_intent.clone(), list(old_private)
        pads, ciphertexts = cb.StreamCipherGadget.encrypt([new_intent.inner[AMOUNT_IN_IDX]], new_intent.share_stream, cs)
        new_private[AMOUNT_IN_IDX], new_intent.public_share[AMOUNT_IN_IDX] = pads[0], ciphertexts[0]
        # 3. the linked copy of the intent, the leaked share, the leaked owner (:86-96)
        for a, b in zip(new_intent.inner, intent):
            cs.enforce_equal(a, b)
        cs.enforce_equal(new_intent.public_share[AMOUNT_IN_IDX], new_amount_share)
        cs.enforce_equal(new_intent.inner[OWNER_IDX], owner)
        # 4. rotation with a partial commitment to the new version (:98-118)
        cb.StateElementRotationGadget.rotate_version_with_partial_commitment(
            PARTIAL_COMMITMENT_SIZE, old_intent, old_private, opening, merkle_root, nullifier, new_intent, new_private,
            (pc_private, pc_public), recovery_id, cs)
        return cs


# ---- INTENT ONLY FIRST FILL VALIDITY --------------------------------------------------------------------------------
@dataclass
class FirstFillWitness:
    """intent_only_first_fill.rs:111-124."""
    intent: Intent
    initial_intent_share_stream: cb.PoseidonCSPRNG
    initial_intent_recovery_stream: cb.PoseidonCSPRNG
    private_shares: List[int]


@dataclass
class FirstFillStatement:
    """intent_only_first_fill.rs:131-142."""
    owner: int
    intent_private_commitment: int
    recovery_id: int
    intent_public_share: List[int]

    def to_scalars(self) -> List[int]:
        return [self.owner, self.intent_private_commitment, self.recovery_id] + list(self.intent_public_share)


def create_first_fill_witness_statement(seed: int = 0, intent: Intent = None):
    """intent_only_first_fill.rs:222-251: the wrapper as `StateWrapper::new` leaves it (one encryption spent on the
    share stream); the recovery id advances the recovery stream BEFORE the private commitment is taken."""
    rnd = random.Random(seed)
    intent = intent or random_intent(rnd)
    initial = cb.StateWrapper.new(intent.to_scalars(), rnd.randrange(R), rnd.randrange(R))
    after = initial.clone()
    recovery_id = after.compute_recovery_id()
    private_commitment = after.compute_private_commitment()
    witness = FirstFillWitness(intent, cb.PoseidonCSPRNG(initial.share_stream.seed, initial.share_stream.index),
                               cb.PoseidonCSPRNG(initial.recovery_stream.seed, initial.recovery_stream.index),
                               initial.private_shares())
    return witness, FirstFillStatement(intent.owner, private_commitment, recovery_id, list(initial.public_share))


class IntentOnlyFirstFillValidityCircuit:
    @staticmethod
    def name() -> str:
        return "Intent Only First Fill Validity"

    @staticmethod
    def build(witness: FirstFillWitness, statement: FirstFillStatement, layout: cb.GroupLayout = None) -> cb.PlonkCircuit:
        cs = cb.PlonkCircuit()
        cs.create_link_group(INTENT_ONLY_SETTLEMENT_LINK, layout or DEFAULT_LAYOUT)
        intent = [cs.create_variable_with_link_groups(v, [INTENT_ONLY_SETTLEMENT_LINK]) for v in witness.intent.to_scalars()]
        mk = lambda st_: cb.PoseidonCSPRNGVar(cs.create_variable(st_.seed), cs.create_variable(st_.index))
        share_stream, recovery_stream = mk(witness.initial_intent_share_stream), mk(witness.initial_intent_recovery_stream)
        private_shares = [cs.create_variable(v) for v in witness.private_shares]
        st = [cs.create_public_variable(v) for v in statement.to_scalars()]
        owner, private_commitment, recovery_id = st[:3]
        # build_and_validate_intent (:70-103)
        cb.AmountGadget.constrain_valid_amount(intent[AMOUNT_IN_IDX], cs)
        cb.BitRangeGadget.constrain_bit_range(intent[3], PRICE_BITS, cs)
        cs.enforce_equal(intent[OWNER_IDX], owner)
        public_share = cb.ShareGadget.compute_complementary_shares(private_shares, intent, cs)
        for got, exp in zip(public_share, st[3:]):
            cs.enforce_equal(got, exp)
        # recovery id, then the commitment to the private shares with the advanced stream (:53-64; commitment.rs:87-102)
        cs.enforce_equal(cb.RecoveryIdGadget.compute_recovery_id(recovery_stream, cs), recovery_id)
        hasher = cb.PoseidonHashGadget(cs.zero())
        hasher.batch_absorb(private_shares + recovery_stream.to_vars() + share_stream.to_vars(), cs)
        cs.enforce_equal(hasher.squeeze(cs), private_commitment)
        return cs
