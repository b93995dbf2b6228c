"""INTENT AND BALANCE PRIVATE SETTLEMENT restated on the host-side constraint system — the VALID-MATCH-class
statement BASELINE.json configs[3] names.

`circuits-core/src/zk_circuits/settlement/intent_and_balance_private_settlement.rs:40-140` with the gadgets of
`settlement/settlement_lib.rs:19-160`, `zk_gadgets/state_gadgets/fee.rs`, `zk_gadgets/primitives/fixed_point.rs:87-133`:
two parties' settlement obligations are compatible; each obligation respects its intent (amount, worst-case price),
its input balance, and fits the output balance after fees; the public shares of intent amount and balances are
updated accordingly.  Pure arithmetic — no hashing: range checks dominate.  Four proof-linking groups carry the
intents, balances and pre-settlement shares over from the validity proofs (`#[link_groups = ...]`, :150-180);
17 public inputs.  Types follow darkpool-types (`intent.rs:49-70`, `settlement_obligation.rs:36-47`,
`balance.rs:48-71,145-152`).

Host-side input generation for tests and benches: the production circuit stays in Rust.  Its exact gate count cannot
be compared with the Rust build here (no cargo); the count this restatement gives is what DESIGN.md quotes as the
real size class of the statement."""
from __future__ import annotations

import random
from dataclasses import dataclass
from typing import List

from . import circuit as cb
from .circuit import R
from .valid_balance_create import DarkpoolBalance

DEFAULT_FP_PRECISION = 63      # circuit-types/src/primitives/fixed_point.rs:41
TWO_TO_M = 1 << DEFAULT_FP_PRECISION
PARTY_LINKS = ("intent_and_balance_settlement_party0", "intent_and_balance_settlement_party1")
OUTPUT_LINKS = ("output_balance_settlement_party0", "output_balance_settlement_party1")


@dataclass
class Intent:
    """darkpool-types/src/intent.rs:49-70; `min_price` is a fixed point (repr = price * 2^63), out_token per in_token."""
    in_token: int
    out_token: int
    owner: int
    min_price: int
    amount_in: int

    def to_scalars(self) -> List[int]:
        return [self.in_token, self.out_token, self.owner, self.min_price, self.amount_in]


@dataclass
class SettlementObligation:
    """darkpool-types/src/settlement_obligation.rs:36-47."""
    input_token: int
    output_token: int
    amount_in: int
    amount_out: int

    def to_scalars(self) -> List[int]:
        return [self.input_token, self.output_token, self.amount_in, self.amount_out]


@dataclass
class PartyWitness:
    """One party's half of `IntentAndBalancePrivateSettlementWitness` (:150-180)."""
    settlement_obligation: SettlementObligation
    intent: Intent
    pre_settlement_amount_public_share: int
    input_balance: DarkpoolBalance
    pre_settlement_in_balance_shares: List[int]     # PostMatchBalanceShare: relayer fee, protocol fee, amount
    output_balance: DarkpoolBalance
    pre_settlement_out_balance_shares: List[int]


@dataclass
class PartyStatement:
    new_amount_public_share: int
    new_in_balance_public_shares: List[int]
    new_out_balance_public_shares: List[int]

    def to_scalars(self) -> List[int]:
        return [self.new_amount_public_share] + list(self.new_in_balance_public_shares) + list(self.new_out_balance_public_shares)


@dataclass
class Statement:
    """:186-207, fields in public-input order."""
    party0: PartyStatement
    party1: PartyStatement
    relayer_fee0: int
    relayer_fee1: int
    protocol_fee: int

    def to_scalars(self) -> List[int]:
        return self.party0.to_scalars() + self.party1.to_scalars() + [self.relayer_fee0, self.relayer_fee1, self.protocol_fee]


def fp_floor_mul(fp_repr: int, integer: int) -> int:
    """floor(fixed point * integer): `FixedPoint::floor` drops the 63 fractional bits."""
    return (fp_repr * integer) >> DEFAULT_FP_PRECISION


def create_witness_statement(seed: int = 0, linked=None):
    """A consistent two-party match: party 0 sells token A for B, party 1 the reverse, at a price both intents accept,
    with balances that cover it; fee takes and share updates as the circuit (and the contracts) compute them.
    `linked`: per party, None or (amount public share, [3 input post-match shares][, [3 output post-match shares]]) as its
    validity proof(s) produced them — the values the linked proofs must agree on."""
    rnd = random.Random(seed)
    addr = lambda: rnd.randrange(1 << 160)
    tok_a, tok_b = addr(), addr()
    x, y = rnd.randrange(1, 1 << 60), rnd.randrange(1, 1 << 60)       # party 0 gives x of A, receives y of B
    relayer_fee0, relayer_fee1, protocol_fee = (rnd.randrange(1 << 50) for _ in range(3))   # rates < 2^-13
    parties, stmts = [], []
    for (tin, tout, ain, aout, rfee) in ((tok_a, tok_b, x, y, relayer_fee0), (tok_b, tok_a, y, x, relayer_fee1)):
        owner = addr()
        price = (aout << DEFAULT_FP_PRECISION) // ain                  # exactly acceptable worst-case price
        min_price = rnd.randrange(price // 2, price + 1)
        intent = Intent(tin, tout, owner, min_price, ain + rnd.randrange(1 << 40))
        obligation = SettlementObligation(tin, tout, ain, aout)
        in_bal = DarkpoolBalance(tin, owner, addr(), rnd.randrange(R), rnd.randrange(R), rnd.randrange(1 << 40),
                                 rnd.randrange(1 << 40), ain + rnd.randrange(1 << 40))
        out_bal = DarkpoolBalance(tout, owner, addr(), rnd.randrange(R), rnd.randrange(R), rnd.randrange(1 << 40),
                                  rnd.randrange(1 << 40), rnd.randrange(1 << 60))
        pre_amount_share = rnd.randrange(R)
        pre_in, pre_out = [rnd.randrange(R) for _ in range(3)], [rnd.randrange(R) for _ in range(3)]
        if linked is not None and linked[len(parties)] is not None:
            lk = linked[len(parties)]
            pre_amount_share, pre_in = lk[0], list(lk[1])
            if len(lk) > 2:                                            # the output balance's validity proof as well
                pre_out = list(lk[2])
        relayer_take, protocol_take = fp_floor_mul(rfee, aout), fp_floor_mul(protocol_fee, aout)
        net = aout - relayer_take - protocol_take
        parties.append(PartyWitness(obligation, intent, pre_amount_share, in_bal, pre_in, out_bal, pre_out))
        stmts.append(PartyStatement((pre_amount_share - ain) % R,
                                    [pre_in[0], pre_in[1], (pre_in[2] - ain) % R],
                                    [(pre_out[0] + relayer_take) % R, (pre_out[1] + protocol_take) % R, (pre_out[2] + net) % R]))
    return parties, Statement(stmts[0], stmts[1], relayer_fee0, relayer_fee1, protocol_fee)


# ---- gadgets (fixed_point.rs:87-133, fee.rs:10-45) ----------------------------------------------------------
class FixedPointGadget:
    @staticmethod
    def mul_integer(fp_repr: cb.Variable, integer: cb.Variable, cs: cb.PlonkCircuit) -> cb.Variable:
        return cs.mul(fp_repr, integer)

    @staticmethod
    def floor(fp_repr: cb.Variable, cs: cb.PlonkCircuit) -> cb.Variable:
        floor_var = cs.create_variable(cs.witness(fp_repr) >> DEFAULT_FP_PRECISION)
        # fp - 2^M * integer is a non-negative value of at most M bits (constrain_equal_floor)
        diff = cs.lc([fp_repr, floor_var, cs.zero(), cs.zero()], [1, -TWO_TO_M, 1, 1])
        cb.BitRangeGadget.constrain_bit_range(diff, DEFAULT_FP_PRECISION, cs)
        return floor_var


class FeeGadget:
    @staticmethod
    def compute_fee_take(receive_amount: cb.Variable, relayer_rate: cb.Variable, protocol_rate: cb.Variable, cs):
        relayer_fp = FixedPointGadget.mul_integer(relayer_rate, receive_amount, cs)
        protocol_fp = FixedPointGadget.mul_integer(protocol_rate, receive_amount, cs)
        return FixedPointGadget.floor(relayer_fp, cs), FixedPointGadget.floor(protocol_fp, cs)

    @staticmethod
    def total_fee(fee_take, cs) -> cb.Variable:
        return cs.add(fee_take[0], fee_take[1])


@dataclass
class _PartyVars:
    obligation: List[cb.Variable]
    intent: List[cb.Variable]
    pre_amount_share: cb.Variable
    in_balance: List[cb.Variable]
    pre_in_shares: List[cb.Variable]
    out_balance: List[cb.Variable]
    pre_out_shares: List[cb.Variable]


class IntentAndBalancePrivateSettlementCircuit:
    @staticmethod
    def name() -> str:
        return "Intent And Balance Private Settlement"

    @staticmethod
    def build(parties: List[PartyWitness], statement: Statement, layouts=None) -> cb.PlonkCircuit:
        """Allocate the witness (link groups as the struct annotates them), the statement as public inputs, apply
        the constraints (:40-98).  `layouts`: group id -> GroupLayout; default: consecutive rows of a 2^12 grid."""
        cs = cb.PlonkCircuit()
        layouts = layouts or {PARTY_LINKS[0]: cb.GroupLayout(12, 32), PARTY_LINKS[1]: cb.GroupLayout(12, 64),
                              OUTPUT_LINKS[0]: cb.GroupLayout(12, 96), OUTPUT_LINKS[1]: cb.GroupLayout(12, 128)}
        for gid in PARTY_LINKS + OUTPUT_LINKS:
            cs.create_link_group(gid, layouts[gid])
        pv: List[_PartyVars] = []
        for i, p in enumerate(parties):
            g, og = [PARTY_LINKS[i]], [OUTPUT_LINKS[i]]
            pv.append(_PartyVars(
                obligation=[cs.create_variable(v) for v in p.settlement_obligation.to_scalars()],
                intent=[cs.create_variable_with_link_groups(v, g) for v in p.intent.to_scalars()],
                pre_amount_share=cs.create_variable_with_link_groups(p.pre_settlement_amount_public_share, g),
                in_balance=[cs.create_variable_with_link_groups(v, g) for v in p.input_balance.to_scalars()],
                pre_in_shares=[cs.create_variable_with_link_groups(v, g) for v in p.pre_settlement_in_balance_shares],
                out_balance=[cs.create_variable_with_link_groups(v, og) for v in p.output_balance.to_scalars()],
                pre_out_shares=[cs.create_variable_with_link_groups(v, og) for v in p.pre_settlement_out_balance_shares]))
        stmt = [cs.create_public_variable(v) for v in statement.to_scalars()]
        st_party = [stmt[0:7], stmt[7:14]]
        relayer_fee = [stmt[14], stmt[15]]
        protocol_fee = stmt[16]

        # 1. obligation compatibility (:100-122)
        o0, o1 = pv[0].obligation, pv[1].obligation
        for amount in (o0[2], o0[3], o1[2], o1[3]):
            cb.AmountGadget.constrain_valid_amount(amount, cs)
        cs.enforce_equal(o0[0], o1[1])
        cs.enforce_equal(o0[1], o1[0])
        cs.enforce_equal(o0[2], o1[3])
        cs.enforce_equal(o0[3], o1[2])
        # 2. fee takes from the fee rates (:78-98)
        fee_takes = [FeeGadget.compute_fee_take(pv[i].obligation[3], relayer_fee[i], protocol_fee, cs) for i in (0, 1)]
        for i in (0, 1):
            p, take = pv[i], fee_takes[i]
            in_token, out_token, amount_in, amount_out = p.obligation
            i_in_token, i_out_token, i_owner, i_min_price, i_amount_in = p.intent
            # 3. intent constraints (settlement_lib.rs:38-70)
            cs.enforce_equal(in_token, i_in_token)
            cs.enforce_equal(out_token, i_out_token)
            cb.GreaterThanEqGadget.constrain_greater_than_eq(i_amount_in, amount_in, cb.AMOUNT_BITS, cs)
            min_output = FixedPointGadget.floor(FixedPointGadget.mul_integer(i_min_price, amount_in, cs), cs)
            cb.GreaterThanEqGadget.constrain_greater_than_eq(amount_out, min_output, cb.AMOUNT_BITS, cs)
            # input balance covers the obligation (:71-83)
            cb.GreaterThanEqGadget.constrain_greater_than_eq(p.in_balance[7], amount_in, cb.AMOUNT_BITS, cs)
            # output balance: mint, owner, no overflow after fees (:85-115)
            cs.enforce_equal(p.out_balance[0], out_token)
            cs.enforce_equal(p.out_balance[1], i_owner)
            net_receive = cs.sub(amount_out, FeeGadget.total_fee(take, cs))
            cb.AmountGadget.constrain_valid_amount(cs.add(p.out_balance[7], net_receive), cs)
            cb.AmountGadget.constrain_valid_amount(cs.add(p.out_balance[5], take[0]), cs)
            cb.AmountGadget.constrain_valid_amount(cs.add(p.out_balance[6], take[1]), cs)
        for i in (0, 1):
            p, take, st = pv[i], fee_takes[i], st_party[i]
            amount_in, amount_out = p.obligation[2], p.obligation[3]
            # 4. state updates (:119-170): intent amount share, input balance shares, output balance shares
            cs.enforce_equal(cs.sub(p.pre_amount_share, amount_in), st[0])
            cs.enforce_equal(p.pre_in_shares[0], st[1])
            cs.enforce_equal(p.pre_in_shares[1], st[2])
            cs.enforce_equal(cs.sub(p.pre_in_shares[2], amount_in), st[3])
            net_receive = cs.sub(amount_out, FeeGadget.total_fee(take, cs))
            cs.enforce_equal(cs.add(p.pre_out_shares[0], take[0]), st[4])
            cs.enforce_equal(cs.add(p.pre_out_shares[1], take[1]), st[5])
            cs.enforce_equal(cs.add(p.pre_out_shares[2], net_receive), st[6])
        return cs
