"""The three single-element update statements restated on the host-side constraint system:

* VALID DEPOSIT           `circuits-core/src/zk_circuits/valid_deposit.rs:40-129` — a deposit is added to an existing
  balance: the old balance opens to the Merkle root and is nullified, the amount grows by the deposit (both valid
  amounts), only the amount share is re-encrypted and leaked, the new version is committed in full;
* VALID WITHDRAWAL        `valid_withdrawal.rs:36-150` — the mirror image: a non-zero withdrawal not above the balance,
  no outstanding fees, to the balance's owner;
* VALID ORDER CANCELLATION `valid_order_cancellation.rs:43-71` — the intent exists in the Merkle state, its nullifier
  and its owner are the statement's.

Witness / statement construction follows each circuit's `test_helpers` (valid_deposit.rs:231-281,
valid_withdrawal.rs:243-303, valid_order_cancellation.rs:163-196) with a seeded RNG.  Each is one height-10 Merkle
opening plus the commitments of the element(s): 6.8 k gates for the two balance updates, 4.4 k for the cancellation, all on the n = 2^13 domain.

Host-side input generation for tests, the prover service and benches: the production circuits stay in Rust."""
from __future__ import annotations

import random
from dataclasses import dataclass
from typing import List

from . import circuit as cb
from .circuit import R
from .private_settlement import Intent
from .valid_balance_create import DarkpoolBalance, Deposit

MERKLE_HEIGHT = 10  # crates/constants/src/lib.rs:50
AMOUNT_IDX = 7      # position of `amount` in DarkpoolBalance's scalar serialisation (balance.rs:48-71)


@dataclass
class Withdrawal:
    """darkpool-types/src/withdrawal.rs:26-33."""
    to: int
    token: int
    amount: int

    def to_scalars(self) -> List[int]:
        return [self.to, self.token, self.amount]


def _random_state_wrapper(inner: List[int], rnd: random.Random) -> cb.StateWrapper:
    """`create_random_state_wrapper`: random stream seeds, the recovery stream already advanced (a nullifier needs a
    previous recovery id)."""
    w = cb.StateWrapper.new(inner, rnd.randrange(R), rnd.randrange(R))
    w.recovery_stream.index = rnd.randrange(1, 1 << 20)
    return w


def _merkle_opening(leaf_hash: int, rnd: random.Random):
    opening = cb.MerkleOpening([rnd.randrange(R) for _ in range(MERKLE_HEIGHT)],
                               [rnd.random() < 0.5 for _ in range(MERKLE_HEIGHT)])
    return cb.native_merkle_root_prehashed(leaf_hash, opening), opening


def _opening_var(o: cb.MerkleOpening, cs: cb.PlonkCircuit) -> cb.MerkleOpeningVar:
    return cb.MerkleOpeningVar([cs.create_variable(v) for v in o.elems], [cs.create_boolean_variable(b) for b in o.indices])


# ---- VALID DEPOSIT / VALID WITHDRAWAL -------------------------------------------------------------------------------
@dataclass
class BalanceUpdateWitness:
    """valid_deposit.rs:137-142 / valid_withdrawal.rs:158-163."""
    old_balance: cb.StateWrapper
    old_balance_opening: cb.MerkleOpening


@dataclass
class ValidDepositStatement:
    """valid_deposit.rs:153-173, fields in public-input order."""
    deposit: Deposit
    merkle_root: int
    old_balance_nullifier: int
    new_balance_commitment: int
    recovery_id: int
    new_amount_share: int

    def to_scalars(self) -> List[int]:
        return self.deposit.to_scalars() + [self.merkle_root, self.old_balance_nullifier, self.new_balance_commitment,
                                            self.recovery_id, self.new_amount_share]


@dataclass
class ValidWithdrawalStatement:
    """valid_withdrawal.rs:170-183."""
    withdrawal: Withdrawal
    merkle_root: int
    old_balance_nullifier: int
    new_balance_commitment: int
    recovery_id: int
    new_amount_share: int

    def to_scalars(self) -> List[int]:
        return self.withdrawal.to_scalars() + [self.merkle_root, self.old_balance_nullifier, self.new_balance_commitment,
                                               self.recovery_id, self.new_amount_share]


def _rotate_amount(old_balance: cb.StateWrapper, new_amount: int, rnd: random.Random):
    """The part both helpers share: open the old version, re-encrypt the new amount, take the next recovery id, commit."""
    nullifier = old_balance.compute_nullifier()
    root, opening = _merkle_opening(old_balance.compute_commitment(), rnd)
    new_balance = old_balance.clone()
    new_balance.inner[AMOUNT_IDX] = new_amount
    new_share = new_balance.share_stream.stream_cipher_encrypt([new_amount])[0]
    new_balance.public_share[AMOUNT_IDX] = new_share
    recovery_id = new_balance.compute_recovery_id()
    return root, opening, nullifier, new_balance.compute_commitment(), recovery_id, new_share


def create_deposit_witness_statement(seed: int = 0):
    """valid_deposit.rs:231-281: an existing balance of the deposit's mint and owner, fees outstanding or not."""
    rnd = random.Random(seed)
    addr = lambda: rnd.randrange(1 << 160)
    half = 1 << (cb.AMOUNT_BITS - 1)  # old amount + deposit stays a valid amount
    deposit = Deposit(from_=addr(), token=addr(), amount=rnd.randrange(half))
    inner = DarkpoolBalance(deposit.token, deposit.from_, 0, rnd.randrange(R), rnd.randrange(R),
                            rnd.randrange(1 << 40), rnd.randrange(1 << 40), rnd.randrange(half))
    old = _random_state_wrapper(inner.to_scalars(), rnd)
    root, opening, nullifier, commitment, recovery_id, share = _rotate_amount(old, inner.amount + deposit.amount, rnd)
    return BalanceUpdateWitness(old, opening), ValidDepositStatement(deposit, root, nullifier, commitment, recovery_id, share)


def create_withdrawal_witness_statement(seed: int = 0):
    """valid_withdrawal.rs:243-303: a balance with no outstanding fees holding at least the withdrawal."""
    rnd = random.Random(seed)
    addr = lambda: rnd.randrange(1 << 160)
    half = 1 << (cb.AMOUNT_BITS - 1)
    withdrawal = Withdrawal(to=addr(), token=addr(), amount=rnd.randrange(1, half))
    inner = DarkpoolBalance(withdrawal.token, withdrawal.to, addr(), rnd.randrange(R), rnd.randrange(R), 0, 0,
                            withdrawal.amount + rnd.randrange(half))
    old = _random_state_wrapper(inner.to_scalars(), rnd)
    root, opening, nullifier, commitment, recovery_id, share = _rotate_amount(old, inner.amount - withdrawal.amount, rnd)
    return (BalanceUpdateWitness(old, opening),
            ValidWithdrawalStatement(withdrawal, root, nullifier, commitment, recovery_id, share))


def _apply_amount_update(cs: cb.PlonkCircuit, old_balance: cb.StateWrapperVar, opening: cb.MerkleOpeningVar,
                         new_amount: cb.Variable, st_tail) -> None:
    """`create_new_balance` / `build_and_verify_new_balance` + the rotation (valid_deposit.rs:52-78, 101-128;
    valid_withdrawal.rs:50-72, 118-149).  st_tail = (merkle_root, nullifier, new_commitment, recovery_id, new_amount_share)."""
    merkle_root, nullifier, new_commitment, recovery_id, new_amount_share = st_tail
    old_private = cb.ShareGadget.compute_complementary_shares(old_balance.public_share, old_balance.inner, cs)
    new_balance, new_private = old_balance.clone(), list(old_private)
    new_balance.inner[AMOUNT_IDX] = new_amount
    pads, ciphertexts = cb.StreamCipherGadget.encrypt([new_amount], new_balance.share_stream, cs)
    new_private[AMOUNT_IDX], new_balance.public_share[AMOUNT_IDX] = pads[0], ciphertexts[0]
    cs.enforce_equal(ciphertexts[0], new_amount_share)
    cb.StateElementRotationGadget.rotate_version(old_balance, old_private, opening, merkle_root, nullifier,
                                                 new_balance, new_private, new_commitment, recovery_id, cs)


class ValidDeposit:
    @staticmethod
    def name() -> str:
        return f"Valid Deposit ({MERKLE_HEIGHT})"

    @staticmethod
    def build(witness: BalanceUpdateWitness, statement: ValidDepositStatement) -> cb.PlonkCircuit:
        cs = cb.PlonkCircuit()
        old_balance = cb.StateWrapperVar.create_witness(witness.old_balance, cs)
        opening = _opening_var(witness.old_balance_opening, cs)
        st = [cs.create_public_variable(v) for v in statement.to_scalars()]
        dep_from, dep_token, dep_amount = st[:3]
        # validate_deposit (:81-97)
        cb.AmountGadget.constrain_valid_amount(dep_amount, cs)
        cs.enforce_equal(dep_token, old_balance.inner[0])
        cs.enforce_equal(dep_from, old_balance.inner[1])
        # create_new_balance (:101-128): the sum must stay a valid amount
        new_amount = cs.add(old_balance.inner[AMOUNT_IDX], dep_amount)
        cb.AmountGadget.constrain_valid_amount(new_amount, cs)
        _apply_amount_update(cs, old_balance, opening, new_amount, st[3:])
        return cs


class ValidWithdrawal:
    @staticmethod
    def name() -> str:
        return f"Valid Withdrawal ({MERKLE_HEIGHT})"

    @staticmethod
    def build(witness: BalanceUpdateWitness, statement: ValidWithdrawalStatement) -> cb.PlonkCircuit:
        cs = cb.PlonkCircuit()
        old_balance = cb.StateWrapperVar.create_witness(witness.old_balance, cs)
        opening = _opening_var(witness.old_balance_opening, cs)
        st = [cs.create_public_variable(v) for v in statement.to_scalars()]
        wd_to, wd_token, wd_amount = st[:3]
        # validate_withdrawal (:76-101): a non-zero valid amount, of the balance's mint, to its owner, not above the balance
        cb.AmountGadget.constrain_valid_amount(wd_amount, cs)
        cs.enforce_constant(cb.EqZeroGadget.eq_zero_var(wd_amount, cs), 0)
        cs.enforce_equal(wd_token, old_balance.inner[0])
        cs.enforce_equal(wd_to, old_balance.inner[1])
        cb.GreaterThanEqGadget.constrain_greater_than_eq(old_balance.inner[AMOUNT_IDX], wd_amount, cb.AMOUNT_BITS, cs)
        # verify_no_outstanding_fees (:106-114)
        cs.enforce_equal(old_balance.inner[5], cs.zero())
        cs.enforce_equal(old_balance.inner[6], cs.zero())
        new_amount = cs.sub(old_balance.inner[AMOUNT_IDX], wd_amount)
        _apply_amount_update(cs, old_balance, opening, new_amount, st[3:])
        return cs


# ---- VALID ORDER CANCELLATION ---------------------------------------------------------------------------------------
@dataclass
class ValidOrderCancellationWitness:
    """valid_order_cancellation.rs:80-85."""
    old_intent: cb.StateWrapper
    old_intent_opening: cb.MerkleOpening


@dataclass
class ValidOrderCancellationStatement:
    """valid_order_cancellation.rs:98-106."""
    merkle_root: int
    old_intent_nullifier: int
    owner: int

    def to_scalars(self) -> List[int]:
        return [self.merkle_root, self.old_intent_nullifier, self.owner]


def create_cancellation_witness_statement(seed: int = 0, intent: Intent = None):
    """valid_order_cancellation.rs:163-196."""
    rnd = random.Random(seed)
    addr = lambda: rnd.randrange(1 << 160)
    if intent is None:
        intent = Intent(addr(), addr(), addr(), rnd.randrange(1 << 100), rnd.randrange(1, 1 << 90))
    old_intent = _random_state_wrapper(intent.to_scalars(), rnd)
    root, opening = _merkle_opening(old_intent.compute_commitment(), rnd)
    return (ValidOrderCancellationWitness(old_intent, opening),
            ValidOrderCancellationStatement(root, old_intent.compute_nullifier(), intent.owner))


class ValidOrderCancellationCircuit:
    @staticmethod
    def name() -> str:
        return f"Valid Order Cancellation ({MERKLE_HEIGHT})"

    @staticmethod
    def build(witness: ValidOrderCancellationWitness, statement: ValidOrderCancellationStatement) -> cb.PlonkCircuit:
        cs = cb.PlonkCircuit()
        intent = cb.StateWrapperVar.create_witness(witness.old_intent, cs)
        opening = _opening_var(witness.old_intent_opening, cs)
        merkle_root, nullifier, owner = [cs.create_public_variable(v) for v in statement.to_scalars()]
        # 1. the intent exists in the Merkle tree (:49-60)
        private_shares = cb.ShareGadget.compute_complementary_shares(intent.public_share, intent.inner, cs)
        commitment = cb.CommitmentGadget.compute_commitment(private_shares, intent.recovery_stream, intent.share_stream,
                                                            intent.public_share, cs)
        cs.enforce_equal(merkle_root, cb.PoseidonMerkleHashGadget.compute_root_prehashed(commitment, opening, cs))
        # 2. its nullifier, 3. its owner (:63-67; Intent's third scalar is the owner, intent.rs)
        cs.enforce_equal(cb.NullifierGadget.compute_nullifier(intent, cs), nullifier)
        cs.enforce_equal(intent.inner[2], owner)
        return cs
