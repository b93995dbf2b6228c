"""INTENT AND BALANCE PUBLIC SETTLEMENT and INTENT AND BALANCE BOUNDED SETTLEMENT restated on the host-side constraint
system — one party's side of a match settled in the open, linked to that party's validity proofs through the PARTY 0
groups of the private settlement circuit's layout (both circuits inherit them).

`settlement/intent_and_balance_public_settlement.rs:44-92`: fee take from the public fee rates, the obligation respects the
intent, the input balance covers it, the output balance (right mint and owner) does not overflow after fees
(`settlement_lib.rs: verify_intent_and_balance_obligation_constraints`), the leaked pre-update shares and the relayer fee
recipient are the linked witness's.  `settlement/intent_and_balance_bounded_settlement.rs:44-82`: the same against a bounded
match result (`BoundedSettlementGadget`).  Witness / statement construction follows the circuits' `test_helpers`.

Host-side input generation for tests, the prover service and benches: the production circuits stay in Rust."""
from __future__ import annotations

import random
from dataclasses import dataclass
from typing import List

from . import circuit as cb
from . import private_settlement as ps
from .intent_only import PRICE_BITS, BoundedMatchResult
from .private_settlement import DEFAULT_FP_PRECISION, FeeGadget, FixedPointGadget, Intent, SettlementObligation
from .valid_balance_create import DarkpoolBalance

PARTY_LINK, OUTPUT_LINK = ps.PARTY_LINKS[0], ps.OUTPUT_LINKS[0]


@dataclass
class Witness:
    """intent_and_balance_public_settlement.rs:101-126 (the bounded circuit's witness is the same, :88-113)."""
    intent: Intent
    pre_settlement_amount_public_share: int
    in_balance: DarkpoolBalance
    pre_settlement_in_balance_shares: List[int]      # PostMatchBalanceShare: relayer fee, protocol fee, amount
    out_balance: DarkpoolBalance
    pre_settlement_out_balance_shares: List[int]


@dataclass
class PublicStatement:
    """intent_and_balance_public_settlement.rs:133-166."""
    settlement_obligation: SettlementObligation
    amount_public_share: int
    in_balance_public_shares: List[int]
    out_balance_public_shares: List[int]
    relayer_fee_rate: int
    protocol_fee_rate: int
    relayer_fee_recipient: int

    def to_scalars(self) -> List[int]:
        return (self.settlement_obligation.to_scalars() + [self.amount_public_share] + list(self.in_balance_public_shares) +
                list(self.out_balance_public_shares) + [self.relayer_fee_rate, self.protocol_fee_rate, self.relayer_fee_recipient])


@dataclass
class BoundedStatement:
    """intent_and_balance_bounded_settlement.rs:120-150."""
    bounded_match_result: BoundedMatchResult
    amount_public_share: int
    in_balance_public_shares: List[int]
    out_balance_public_shares: List[int]
    internal_relayer_fee: int
    external_relayer_fee: int
    relayer_fee_recipient: int

    def to_scalars(self) -> List[int]:
        return (self.bounded_match_result.to_scalars() + [self.amount_public_share] + list(self.in_balance_public_shares) +
                list(self.out_balance_public_shares) + [self.internal_relayer_fee, self.external_relayer_fee,
                                                        self.relayer_fee_recipient])


def _party(seed: int):
    """Party 0 of a consistent match (private_settlement.create_witness_statement): intent, obligation, balances, shares."""
    parties, statement = ps.create_witness_statement(seed)
    p = parties[0]
    w = Witness(p.intent, p.pre_settlement_amount_public_share, p.input_balance, list(p.pre_settlement_in_balance_shares),
                p.output_balance, list(p.pre_settlement_out_balance_shares))
    return w, p.settlement_obligation, statement.relayer_fee0, statement.protocol_fee


def create_public_witness_statement(seed: int = 0):
    w, obligation, relayer_rate, protocol_rate = _party(seed)
    return w, PublicStatement(obligation, w.pre_settlement_amount_public_share, list(w.pre_settlement_in_balance_shares),
                              list(w.pre_settlement_out_balance_shares), relayer_rate, protocol_rate,
                              w.out_balance.relayer_fee_recipient)


def create_bounded_witness_statement(seed: int = 0):
    w, obligation, relayer_rate, _ = _party(seed)
    rnd = random.Random(seed ^ 0xB0)
    price = (obligation.amount_out << DEFAULT_FP_PRECISION) // obligation.amount_in     # >= the intent's worst case
    bmr = BoundedMatchResult(w.intent.in_token, w.intent.out_token, rnd.randrange(0, obligation.amount_in + 1),
                             obligation.amount_in, price, rnd.randrange(1 << 40))
    return w, BoundedStatement(bmr, w.pre_settlement_amount_public_share, list(w.pre_settlement_in_balance_shares),
                               list(w.pre_settlement_out_balance_shares), relayer_rate, rnd.randrange(1 << 50),
                               w.out_balance.relayer_fee_recipient)


def _allocate(witness: Witness, layouts, cs: cb.PlonkCircuit):
    layouts = layouts or {PARTY_LINK: cb.GroupLayout(12, 32), OUTPUT_LINK: cb.GroupLayout(12, 96)}
    for gid in (PARTY_LINK, OUTPUT_LINK):
        cs.create_link_group(gid, cb.GroupLayout(layouts[gid].alignment, layouts[gid].offset))
    party = lambda vals: [cs.create_variable_with_link_groups(v, [PARTY_LINK]) for v in vals]
    out = lambda vals: [cs.create_variable_with_link_groups(v, [OUTPUT_LINK]) for v in vals]
    return (party(witness.intent.to_scalars()), party([witness.pre_settlement_amount_public_share])[0],
            party(witness.in_balance.to_scalars()), party(witness.pre_settlement_in_balance_shares),
            out(witness.out_balance.to_scalars()), out(witness.pre_settlement_out_balance_shares))


def _leaks(cs, pre_amount, pre_in, pre_out, out_balance, st_amount, st_in, st_out, st_recipient) -> None:
    """The leaked pre-update shares and the fee recipient are the linked witness's (public :66-90, bounded :57-81)."""
    cs.enforce_equal(pre_amount, st_amount)
    for a, b in zip(pre_in, st_in):
        cs.enforce_equal(a, b)
    for a, b in zip(pre_out, st_out):
        cs.enforce_equal(a, b)
    cs.enforce_equal(out_balance[2], st_recipient)


class IntentAndBalancePublicSettlementCircuit:
    @staticmethod
    def name() -> str:
        return "Intent And Balance Public Settlement"

    @staticmethod
    def build(witness: Witness, statement: PublicStatement, layouts=None) -> cb.PlonkCircuit:
        cs = cb.PlonkCircuit()
        intent, pre_amount, in_balance, pre_in, out_balance, pre_out = _allocate(witness, layouts, cs)
        st = [cs.create_public_variable(v) for v in statement.to_scalars()]
        obligation, st_amount, st_in, st_out = st[0:4], st[4], st[5:8], st[8:11]
        relayer_rate, protocol_rate, recipient = st[11], st[12], st[13]
        in_token, out_token, amount_in, amount_out = obligation
        # 1. fee take from the fee rates (:50-55)
        take = FeeGadget.compute_fee_take(amount_out, relayer_rate, protocol_rate, cs)
        # 2. settlement_lib.rs `verify_intent_and_balance_obligation_constraints`: intent, input balance, output balance
        cs.enforce_equal(in_token, intent[0])
        cs.enforce_equal(out_token, intent[1])
        cb.GreaterThanEqGadget.constrain_greater_than_eq(intent[4], amount_in, cb.AMOUNT_BITS, cs)
        min_output = FixedPointGadget.floor(FixedPointGadget.mul_integer(intent[3], amount_in, cs), cs)
        cb.GreaterThanEqGadget.constrain_greater_than_eq(amount_out, min_output, cb.AMOUNT_BITS, cs)
        cb.GreaterThanEqGadget.constrain_greater_than_eq(in_balance[7], amount_in, cb.AMOUNT_BITS, cs)
        cs.enforce_equal(out_balance[0], out_token)
        cs.enforce_equal(out_balance[1], intent[2])
        net_receive = cs.sub(amount_out, FeeGadget.total_fee(take, cs))
        cb.AmountGadget.constrain_valid_amount(cs.add(out_balance[7], net_receive), cs)
        cb.AmountGadget.constrain_valid_amount(cs.add(out_balance[5], take[0]), cs)
        cb.AmountGadget.constrain_valid_amount(cs.add(out_balance[6], take[1]), cs)
        # 3. leaks
        _leaks(cs, pre_amount, pre_in, pre_out, out_balance, st_amount, st_in, st_out, recipient)
        return cs


class IntentAndBalanceBoundedSettlementCircuit:
    @staticmethod
    def name() -> str:
        return "Intent And Balance Bounded Settlement"

    @staticmethod
    def build(witness: Witness, statement: BoundedStatement, layouts=None) -> cb.PlonkCircuit:
        cs = cb.PlonkCircuit()
        intent, pre_amount, in_balance, pre_in, out_balance, pre_out = _allocate(witness, layouts, cs)
        st = [cs.create_public_variable(v) for v in statement.to_scalars()]
        in_token, out_token, _min_in, max_in, price, _deadline = st[0:6]
        st_amount, st_in, st_out, recipient = st[6], st[7:10], st[10:13], st[15]
        # 1. settlement_lib.rs `BoundedSettlementGadget::verify_intent_and_balance_bounded_match_result_constraints`
        cs.enforce_equal(in_token, intent[0])
        cs.enforce_equal(out_token, intent[1])
        cb.GreaterThanEqGadget.constrain_greater_than_eq(intent[4], max_in, cb.AMOUNT_BITS, cs)
        cb.GreaterThanEqGadget.constrain_greater_than_eq(price, intent[3], PRICE_BITS, cs)
        cb.GreaterThanEqGadget.constrain_greater_than_eq(in_balance[7], max_in, cb.AMOUNT_BITS, cs)
        cs.enforce_equal(out_balance[0], out_token)
        cs.enforce_equal(out_balance[1], intent[2])
        max_output = FixedPointGadget.floor(FixedPointGadget.mul_integer(price, max_in, cs), cs)
        cb.AmountGadget.constrain_valid_amount(cs.add(out_balance[7], max_output), cs)
        # 2., 3. leaks
        _leaks(cs, pre_amount, pre_in, pre_out, out_balance, st_amount, st_in, st_out, recipient)
        return cs
