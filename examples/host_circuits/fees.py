"""Three of the five fee statements restated on the host-side constraint system (the two PRIVATE fee payments encrypt
their note under an ElGamal key on BabyJubJub — in-circuit curve arithmetic this package does not restate):

* VALID NOTE REDEMPTION             `fees/valid_note_redemption.rs:36-58`            — the (public) note opens to the Merkle
  root and its nullifier is H(commitment, blinder) (`state_gadgets/note.rs`);
* VALID PUBLIC PROTOCOL FEE PAYMENT `fees/valid_public_protocol_fee_payment.rs:45-125` — a balance's non-zero protocol fee
  balance leaves as a public note of the balance's mint; the field is zeroed, re-encrypted and leaked, the balance rotated;
* VALID PUBLIC RELAYER FEE PAYMENT  `fees/valid_public_relayer_fee_payment.rs`       — the same for the relayer fee balance,
  with the note's receiver bound to the balance's relayer fee recipient.

Witness / statement construction follows each circuit's `test_helpers` with a seeded RNG.
Host-side input generation for tests, the prover service and benches: the production circuits stay in Rust."""
from __future__ import annotations

import random
from dataclasses import dataclass
from typing import List

from . import circuit as cb
from .circuit import R
from .state_updates import BalanceUpdateWitness, _merkle_opening, _opening_var, _random_state_wrapper
from .valid_balance_create import DarkpoolBalance

MERKLE_HEIGHT = 10
RELAYER_FEE_IDX, PROTOCOL_FEE_IDX = 5, 6   # positions in DarkpoolBalance's scalar serialisation (balance.rs:48-71)


@dataclass
class Note:
    """darkpool-types/src/note.rs:47-59."""
    mint: int
    amount: int
    receiver: int
    blinder: int

    def to_scalars(self) -> List[int]:
        return [self.mint, self.amount, self.receiver, self.blinder]

    def commitment(self) -> int:
        return cb.compute_poseidon_hash(self.to_scalars())

    def nullifier(self) -> int:
        return cb.compute_poseidon_hash([self.commitment(), self.blinder])


# ---- VALID NOTE REDEMPTION ------------------------------------------------------------------------------------------
@dataclass
class NoteRedemptionWitness:
    note_opening: cb.MerkleOpening


@dataclass
class NoteRedemptionStatement:
    """valid_note_redemption.rs:68-76."""
    note: Note
    note_root: int
    note_nullifier: int

    def to_scalars(self) -> List[int]:
        return self.note.to_scalars() + [self.note_root, self.note_nullifier]


def create_note_redemption_witness_statement(seed: int = 0):
    rnd = random.Random(seed)
    note = Note(rnd.randrange(1 << 160), rnd.randrange(1 << cb.AMOUNT_BITS), rnd.randrange(1 << 160), rnd.randrange(R))
    root, opening = _merkle_opening(note.commitment(), rnd)
    return NoteRedemptionWitness(opening), NoteRedemptionStatement(note, root, note.nullifier())


class ValidNoteRedemption:
    @staticmethod
    def name() -> str:
        return f"Valid Note Redemption ({MERKLE_HEIGHT})"

    @staticmethod
    def build(witness: NoteRedemptionWitness, statement: NoteRedemptionStatement) -> cb.PlonkCircuit:
        cs = cb.PlonkCircuit()
        opening = _opening_var(witness.note_opening, cs)
        st = [cs.create_public_variable(v) for v in statement.to_scalars()]
        note, note_root, note_nullifier = st[0:4], st[4], st[5]
        commitment = cb.PoseidonHashGadget(cs.zero()).hash(note, cs)                       # note.rs compute_note_commitment
        cs.enforce_equal(note_root, cb.PoseidonMerkleHashGadget.compute_root_prehashed(commitment, opening, cs))
        cs.enforce_equal(cb.PoseidonHashGadget(cs.zero()).hash([commitment, note[3]], cs), note_nullifier)
        return cs


# ---- VALID PUBLIC {PROTOCOL, RELAYER} FEE PAYMENT -------------------------------------------------------------------
@dataclass
class PublicFeePaymentStatement:
    """valid_public_protocol_fee_payment.rs:139-160 / valid_public_relayer_fee_payment.rs (same shape)."""
    merkle_root: int
    old_balance_nullifier: int
    new_balance_commitment: int
    recovery_id: int
    new_fee_balance_share: int
    note: Note

    def to_scalars(self) -> List[int]:
        return [self.merkle_root, self.old_balance_nullifier, self.new_balance_commitment, self.recovery_id,
                self.new_fee_balance_share] + self.note.to_scalars()


def _create_fee_payment(seed: int, fee_idx: int, bind_receiver: bool):
    rnd = random.Random(seed)
    addr = lambda: rnd.randrange(1 << 160)
    inner = DarkpoolBalance(addr(), addr(), addr() if bind_receiver else 0, rnd.randrange(R), rnd.randrange(R),
                            rnd.randrange(1, 1 << 60), rnd.randrange(1, 1 << 60), rnd.randrange(1 << 90))
    old = _random_state_wrapper(inner.to_scalars(), rnd)
    nullifier = old.compute_nullifier()
    root, opening = _merkle_opening(old.compute_commitment(), rnd)
    note = Note(inner.mint, old.inner[fee_idx], inner.relayer_fee_recipient if bind_receiver else addr(), rnd.randrange(R))
    new = old.clone()
    new.inner[fee_idx] = 0
    share = new.share_stream.stream_cipher_encrypt([0])[0]
    new.public_share[fee_idx] = share
    recovery_id = new.compute_recovery_id()
    return (BalanceUpdateWitness(old, opening),
            PublicFeePaymentStatement(root, nullifier, new.compute_commitment(), recovery_id, share, note))


def create_public_protocol_fee_payment_witness_statement(seed: int = 0):
    return _create_fee_payment(seed, PROTOCOL_FEE_IDX, bind_receiver=False)


def create_public_relayer_fee_payment_witness_statement(seed: int = 0):
    return _create_fee_payment(seed, RELAYER_FEE_IDX, bind_receiver=True)


def _build_fee_payment(witness: BalanceUpdateWitness, statement: PublicFeePaymentStatement, fee_idx: int,
                       bind_receiver: bool) -> cb.PlonkCircuit:
    cs = cb.PlonkCircuit()
    old_balance = cb.StateWrapperVar.create_witness(witness.old_balance, cs)
    opening = _opening_var(witness.old_balance_opening, cs)
    st = [cs.create_public_variable(v) for v in statement.to_scalars()]
    merkle_root, nullifier, new_commitment, recovery_id, new_fee_share = st[:5]
    note_mint, note_amount, note_receiver, _blinder = st[5:9]
    # verify_note (:67-86): there is a fee to pay, and the note carries exactly it
    cs.enforce_constant(cb.EqZeroGadget.eq_zero_var(old_balance.inner[fee_idx], cs), 0)
    cs.enforce_equal(note_mint, old_balance.inner[0])
    cs.enforce_equal(note_amount, old_balance.inner[fee_idx])
    if bind_receiver:
        cs.enforce_equal(note_receiver, old_balance.inner[2])
    # the state transition (:52-64, 88-113): the fee balance zeroed, re-encrypted, leaked; the element rotated
    old_private = cb.ShareGadget.compute_complementary_shares(old_balance.public_share, old_balance.inner, cs)
    new_balance, new_private = old_balance.clone(), list(old_private)
    new_balance.inner[fee_idx] = cs.zero()
    pads, ciphertexts = cb.StreamCipherGadget.encrypt([new_balance.inner[fee_idx]], new_balance.share_stream, cs)
    new_private[fee_idx], new_balance.public_share[fee_idx] = pads[0], ciphertexts[0]
    cs.enforce_equal(ciphertexts[0], new_fee_share)
    cb.StateElementRotationGadget.rotate_version(old_balance, old_private, opening, merkle_root, nullifier, new_balance,
                                                 new_private, new_commitment, recovery_id, cs)
    return cs


class ValidPublicProtocolFeePayment:
    @staticmethod
    def name() -> str:
        return f"Valid Public Protocol Fee Payment ({MERKLE_HEIGHT})"

    @staticmethod
    def build(witness, statement) -> cb.PlonkCircuit:
        return _build_fee_payment(witness, statement, PROTOCOL_FEE_IDX, bind_receiver=False)


class ValidPublicRelayerFeePayment:
    @staticmethod
    def name() -> str:
        return f"Valid Public Relayer Fee Payment ({MERKLE_HEIGHT})"

    @staticmethod
    def build(witness, statement) -> cb.PlonkCircuit:
        return _build_fee_payment(witness, statement, RELAYER_FEE_IDX, bind_receiver=True)
