"""VALID BALANCE CREATE restated on the host-side constraint system (BASELINE.json configs[0]).

`circuits-core/src/zk_circuits/valid_balance_create.rs:36-122` — a deposit creates a new darkpool balance: the deposit
is a valid amount for the balance's mint and owner, the balance is encrypted with its share stream (the public share is
in the statement), its first recovery identifier and its commitment are in the statement.  23 Poseidon2 permutations
plus a 100-bit range check: ≈ 4.7 k gates, domain 2^13.  Types and the witness/statement construction follow
darkpool-types (`balance.rs:48-71`, `deposit.rs:19-23`, `state_wrapper.rs`) and the circuit's own
`test_helpers::create_witness_statement` (valid_balance_create.rs:197-244).

Host-side input generation for tests and benches: the production circuit stays in Rust."""
from __future__ import annotations

import random
from dataclasses import dataclass
from typing import List

from . import circuit as cb
from .circuit import R


@dataclass
class DarkpoolBalance:
    """darkpool-types/src/balance.rs:48-71; `authority` is a BabyJubJub point (two scalars)."""
    mint: int
    owner: int
    relayer_fee_recipient: int
    authority_x: int
    authority_y: int
    relayer_fee_balance: int = 0
    protocol_fee_balance: int = 0
    amount: int = 0

    def to_scalars(self) -> List[int]:
        return [self.mint, self.owner, self.relayer_fee_recipient, self.authority_x, self.authority_y,
                self.relayer_fee_balance, self.protocol_fee_balance, self.amount]


@dataclass
class Deposit:
    """darkpool-types/src/deposit.rs:19-23."""
    from_: int
    token: int
    amount: int

    def to_scalars(self) -> List[int]:
        return [self.from_, self.token, self.amount]


@dataclass
class ValidBalanceCreateWitness:
    initial_share_stream: cb.PoseidonCSPRNG
    initial_recovery_stream: cb.PoseidonCSPRNG
    balance: DarkpoolBalance


@dataclass
class ValidBalanceCreateStatement:
    """valid_balance_create.rs:141-146, fields in public-input order."""
    deposit: Deposit
    balance_commitment: int
    recovery_id: int
    new_balance_share: List[int]

    def to_scalars(self) -> List[int]:
        return self.deposit.to_scalars() + [self.balance_commitment, self.recovery_id] + list(self.new_balance_share)


def create_witness_statement(seed: int = 0):
    """valid_balance_create.rs:197-244 with a seeded RNG: a random deposit, the matching new balance wrapped with
    random stream seeds (`StateWrapper::new` already spends one encryption on the share stream), then a second
    encryption, the first recovery id and the commitment of the updated element."""
    rnd = random.Random(seed)
    addr = lambda: rnd.randrange(1 << 160)
    deposit = Deposit(from_=addr(), token=addr(), amount=rnd.randrange(1 << cb.AMOUNT_BITS))
    inner = DarkpoolBalance(mint=deposit.token, owner=deposit.from_, relayer_fee_recipient=addr(),
                            authority_x=rnd.randrange(R), authority_y=rnd.randrange(R), amount=deposit.amount)
    balance = cb.StateWrapper.new(inner.to_scalars(), rnd.randrange(R), rnd.randrange(R))
    witness = ValidBalanceCreateWitness(cb.PoseidonCSPRNG(balance.share_stream.seed, balance.share_stream.index),
                                        cb.PoseidonCSPRNG(balance.recovery_stream.seed, balance.recovery_stream.index), inner)
    new_balance = balance
    new_balance.public_share = new_balance.share_stream.stream_cipher_encrypt(inner.to_scalars())
    recovery_id = new_balance.compute_recovery_id()
    commitment = new_balance.compute_commitment()
    return witness, ValidBalanceCreateStatement(deposit, commitment, recovery_id, list(new_balance.public_share))


class ValidBalanceCreate:
    """`impl SingleProverCircuit for ValidBalanceCreate` (valid_balance_create.rs:152-169)."""

    @staticmethod
    def name() -> str:
        return "Valid Balance Create"

    @staticmethod
    def build(witness: ValidBalanceCreateWitness, statement: ValidBalanceCreateStatement) -> cb.PlonkCircuit:
        """`SingleProverCircuit::prove` up to the SNARK call (traits.rs:976-991): allocate the witness, allocate the
        statement as public inputs, apply the constraints."""
        cs = cb.PlonkCircuit()
        share_stream = cb.PoseidonCSPRNGVar(cs.create_variable(witness.initial_share_stream.seed),
                                            cs.create_variable(witness.initial_share_stream.index))
        recovery_stream = cb.PoseidonCSPRNGVar(cs.create_variable(witness.initial_recovery_stream.seed),
                                               cs.create_variable(witness.initial_recovery_stream.index))
        balance = [cs.create_variable(v) for v in witness.balance.to_scalars()]
        stmt = [cs.create_public_variable(v) for v in statement.to_scalars()]
        dep_from, dep_token, dep_amount, st_commitment, st_recovery_id = stmt[:5]
        st_share = stmt[5:]
        mint, owner, _recipient, _ax, _ay, relayer_fee, protocol_fee, amount = balance
        # 1. validate the deposit and the new balance (valid_balance_create.rs:64-99)
        cb.AmountGadget.constrain_valid_amount(dep_amount, cs)
        cs.enforce_equal(dep_token, mint)
        cs.enforce_equal(dep_from, owner)
        cs.enforce_equal(amount, dep_amount)
        cs.enforce_equal(relayer_fee, cs.zero())
        cs.enforce_equal(protocol_fee, cs.zero())
        # 2. encrypt the balance with its share stream; the public share is in the statement (:101-122)
        private_share, public_share = cb.StreamCipherGadget.encrypt(balance, share_stream, cs)
        for got, exp in zip(public_share, st_share):
            cs.enforce_equal(got, exp)
        # 3. recovery identifier, 4. commitment to the element with its UPDATED stream states (:48-60)
        cs.enforce_equal(cb.RecoveryIdGadget.compute_recovery_id(recovery_stream, cs), st_recovery_id)
        commitment = cb.CommitmentGadget.compute_commitment(private_share, recovery_stream, share_stream, public_share, cs)
        cs.enforce_equal(commitment, st_commitment)
        return cs
