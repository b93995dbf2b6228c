"""The restated statements as `SingleProverCircuit`s (renegade_b200/circuit_types.py), under the reference's
circuit names — the registry the prover service (renegade_b200/service.py) and the tests use.

Reference: the `SingleProverCircuit` impls at circuits-core/src/zk_circuits/valid_balance_create.rs:188-215,
settlement/intent_and_balance_private_settlement.rs:288-330, validity_proofs/intent_and_balance.rs:262-300,
validity_proofs/output_balance.rs (same pattern); the validity circuits take their link-group placement from the
settlement circuit's layout (`proof_linking_groups`, intent_and_balance.rs:270-279)."""
from __future__ import annotations

from renegade_b200.circuit_types import SingleProverCircuit
from renegade_b200.fields import scalars_to_limbs

from . import fees
from . import intent_and_balance_validity as val
from . import intent_only as io
from . import output_balance_validity as obv
from . import private_settlement as ps
from . import public_settlement as pub
from . import state_updates as su
from . import valid_balance_create as vbc


class ValidBalanceCreate(SingleProverCircuit):
    @classmethod
    def name(cls):
        return vbc.ValidBalanceCreate.name()

    @classmethod
    def synthesize(cls, witness, statement, layout):
        return vbc.ValidBalanceCreate.build(witness, statement)

    @classmethod
    def statement_scalars(cls, statement):
        return scalars_to_limbs(statement.to_scalars())

    @classmethod
    def dummy_instance(cls):
        return vbc.create_witness_statement(0)


class IntentAndBalancePrivateSettlementCircuit(SingleProverCircuit):
    """Witness = the two `PartyWitness`es."""

    @classmethod
    def name(cls):
        return ps.IntentAndBalancePrivateSettlementCircuit.name()

    @classmethod
    def proof_linking_groups(cls):
        return [(g, None) for g in ps.PARTY_LINKS + ps.OUTPUT_LINKS]

    @classmethod
    def generate_layout(cls):
        parties, statement = ps.create_witness_statement(0)
        return ps.IntentAndBalancePrivateSettlementCircuit.build(parties, statement).get_circuit_layout()

    @classmethod
    def synthesize(cls, witness, statement, layout):
        return ps.IntentAndBalancePrivateSettlementCircuit.build(witness, statement, layout)

    @classmethod
    def statement_scalars(cls, statement):
        return scalars_to_limbs(statement.to_scalars())

    @classmethod
    def dummy_instance(cls):
        return ps.create_witness_statement(0)


class IntentAndBalanceValidityCircuit(SingleProverCircuit):
    @classmethod
    def name(cls):
        return val.IntentAndBalanceValidityCircuit.name()

    @classmethod
    def proof_linking_groups(cls):
        lay = IntentAndBalancePrivateSettlementCircuit.get_circuit_layout()
        return [(g, lay[g]) for g in ps.PARTY_LINKS]

    @classmethod
    def generate_layout(cls):  # the settlement circuit's placement of the party groups
        return IntentAndBalancePrivateSettlementCircuit.get_circuit_layout()

    @classmethod
    def synthesize(cls, witness, statement, layout):
        return val.IntentAndBalanceValidityCircuit.build(witness, statement, layout)

    @classmethod
    def statement_scalars(cls, statement):
        return scalars_to_limbs(statement.to_scalars())

    @classmethod
    def dummy_instance(cls):
        return val.create_witness_statement(0)


class OutputBalanceValidityCircuit(SingleProverCircuit):
    @classmethod
    def name(cls):
        return obv.OutputBalanceValidityCircuit.name()

    @classmethod
    def proof_linking_groups(cls):
        lay = IntentAndBalancePrivateSettlementCircuit.get_circuit_layout()
        return [(g, lay[g]) for g in ps.OUTPUT_LINKS]

    @classmethod
    def generate_layout(cls):
        return IntentAndBalancePrivateSettlementCircuit.get_circuit_layout()

    @classmethod
    def synthesize(cls, witness, statement, layout):
        return obv.OutputBalanceValidityCircuit.build(witness, statement, layout)

    @classmethod
    def statement_scalars(cls, statement):
        return scalars_to_limbs(statement.to_scalars())

    @classmethod
    def dummy_instance(cls):
        parties, _ = ps.create_witness_statement(0)
        return obv.create_witness_statement(0, parties[0].output_balance)


def _plain(mod_circuit, make_instance):
    """A statement without link groups: name, synthesis, public inputs and a dummy instance."""
    class _C(SingleProverCircuit):
        @classmethod
        def name(cls):
            return mod_circuit.name()

        @classmethod
        def synthesize(cls, witness, statement, layout):
            return mod_circuit.build(witness, statement)

        @classmethod
        def statement_scalars(cls, statement):
            return scalars_to_limbs(statement.to_scalars())

        @classmethod
        def dummy_instance(cls):
            return make_instance(0)
    _C.__name__ = _C.__qualname__ = mod_circuit.__name__
    return _C


# valid_deposit.rs:177-195, valid_withdrawal.rs:189-207, valid_order_cancellation.rs:109-127
ValidDeposit = _plain(su.ValidDeposit, su.create_deposit_witness_statement)
ValidWithdrawal = _plain(su.ValidWithdrawal, su.create_withdrawal_witness_statement)
ValidOrderCancellationCircuit = _plain(su.ValidOrderCancellationCircuit, su.create_cancellation_witness_statement)

# fees/valid_note_redemption.rs:78-96, fees/valid_public_protocol_fee_payment.rs:166-186, fees/valid_public_relayer_fee_payment.rs
ValidNoteRedemption = _plain(fees.ValidNoteRedemption, fees.create_note_redemption_witness_statement)
ValidPublicProtocolFeePayment = _plain(fees.ValidPublicProtocolFeePayment, fees.create_public_protocol_fee_payment_witness_statement)
ValidPublicRelayerFeePayment = _plain(fees.ValidPublicRelayerFeePayment, fees.create_public_relayer_fee_payment_witness_statement)


# ---- one party's side settled in the open: both inherit the PARTY 0 groups of the private settlement circuit's layout
# (intent_and_balance_public_settlement.rs:204-213, intent_and_balance_bounded_settlement.rs:189-198)
def _party0_inheritor(mod_circuit, make_instance):
    class _C(SingleProverCircuit):
        @classmethod
        def name(cls):
            return mod_circuit.name()

        @classmethod
        def proof_linking_groups(cls):
            lay = IntentAndBalancePrivateSettlementCircuit.get_circuit_layout()
            return [(g, lay[g]) for g in (pub.PARTY_LINK, pub.OUTPUT_LINK)]

        @classmethod
        def generate_layout(cls):
            lay = IntentAndBalancePrivateSettlementCircuit.get_circuit_layout()
            return {g: lay[g] for g in (pub.PARTY_LINK, pub.OUTPUT_LINK)}

        @classmethod
        def synthesize(cls, witness, statement, layout):
            return mod_circuit.build(witness, statement, layout)

        @classmethod
        def statement_scalars(cls, statement):
            return scalars_to_limbs(statement.to_scalars())

        @classmethod
        def dummy_instance(cls):
            return make_instance(0)
    _C.__name__ = _C.__qualname__ = mod_circuit.__name__
    return _C


IntentAndBalancePublicSettlementCircuit = _party0_inheritor(pub.IntentAndBalancePublicSettlementCircuit,
                                                            pub.create_public_witness_statement)
IntentAndBalanceBoundedSettlementCircuit = _party0_inheritor(pub.IntentAndBalanceBoundedSettlementCircuit,
                                                             pub.create_bounded_witness_statement)


# ---- the intent-only family: one link group, placed by INTENT ONLY PUBLIC SETTLEMENT and inherited by the other three
# (intent_only_public_settlement.rs:120-123, intent_only_bounded_settlement.rs, intent_only.rs:223-229,
#  intent_only_first_fill.rs:168-174)
class IntentOnlyPublicSettlementCircuit(SingleProverCircuit):
    @classmethod
    def name(cls):
        return io.IntentOnlyPublicSettlementCircuit.name()

    @classmethod
    def proof_linking_groups(cls):
        return [(io.INTENT_ONLY_SETTLEMENT_LINK, None)]

    @classmethod
    def generate_layout(cls):
        return io.IntentOnlyPublicSettlementCircuit.build(*io.create_public_settlement_witness_statement(0)).get_circuit_layout()

    @classmethod
    def synthesize(cls, witness, statement, layout):
        return io.IntentOnlyPublicSettlementCircuit.build(witness, statement, layout[io.INTENT_ONLY_SETTLEMENT_LINK])

    @classmethod
    def statement_scalars(cls, statement):
        return scalars_to_limbs(statement.to_scalars())

    @classmethod
    def dummy_instance(cls):
        return io.create_public_settlement_witness_statement(0)


def _intent_only_inheritor(mod_circuit, make_instance):
    class _C(SingleProverCircuit):
        @classmethod
        def name(cls):
            return mod_circuit.name()

        @classmethod
        def proof_linking_groups(cls):
            lay = IntentOnlyPublicSettlementCircuit.get_circuit_layout()
            return [(io.INTENT_ONLY_SETTLEMENT_LINK, lay[io.INTENT_ONLY_SETTLEMENT_LINK])]

        @classmethod
        def generate_layout(cls):
            return IntentOnlyPublicSettlementCircuit.get_circuit_layout()

        @classmethod
        def synthesize(cls, witness, statement, layout):
            return mod_circuit.build(witness, statement, layout[io.INTENT_ONLY_SETTLEMENT_LINK])

        @classmethod
        def statement_scalars(cls, statement):
            return scalars_to_limbs(statement.to_scalars())

        @classmethod
        def dummy_instance(cls):
            return make_instance(0)
    _C.__name__ = _C.__qualname__ = mod_circuit.__name__
    return _C


IntentOnlyBoundedSettlementCircuit = _intent_only_inheritor(io.IntentOnlyBoundedSettlementCircuit,
                                                            io.create_bounded_settlement_witness_statement)
IntentOnlyValidityCircuit = _intent_only_inheritor(io.IntentOnlyValidityCircuit, io.create_validity_witness_statement)
IntentOnlyFirstFillValidityCircuit = _intent_only_inheritor(io.IntentOnlyFirstFillValidityCircuit,
                                                            io.create_first_fill_witness_statement)

# the circuits `NativeProofManager::preprocess_circuits` registers (native_proof_manager.rs:305-331) that are restated here
REGISTERED = [ValidBalanceCreate, ValidDeposit, ValidWithdrawal, ValidOrderCancellationCircuit,
              IntentOnlyValidityCircuit, IntentOnlyFirstFillValidityCircuit, IntentOnlyPublicSettlementCircuit,
              IntentOnlyBoundedSettlementCircuit,
              IntentAndBalancePrivateSettlementCircuit, IntentAndBalancePublicSettlementCircuit,
              IntentAndBalanceBoundedSettlementCircuit, IntentAndBalanceValidityCircuit, OutputBalanceValidityCircuit,
              ValidNoteRedemption, ValidPublicProtocolFeePayment, ValidPublicRelayerFeePayment]
# not restated: INTENT AND BALANCE FIRST FILL VALIDITY and NEW OUTPUT BALANCE VALIDITY (their constraints call jf-primitives'
# Schnorr `SignatureGadget`, zk_gadgets/primitives/schnorr.rs:20-24), VALID PRIVATE PROTOCOL / RELAYER FEE PAYMENT
# (jf-primitives' `ElGamalEncryptionGadget`, zk_gadgets/primitives/elgamal.rs:41-47) — gadgets of the un-vendored fork


# ---- collaborative counterparts (traits.rs:1103-1154) ---------------------------------------------------------------
from renegade_b200.circuit_types import MultiProverCircuit  # noqa: E402


class IntentAndBalancePrivateSettlementMultiprover(MultiProverCircuit):
    """The VALID-MATCH-class statement proved jointly by the two parties of a match: each holds a share of the wire
    table of `IntentAndBalancePrivateSettlementCircuit`; the opened proof verifies under that circuit's keys."""
    BaseCircuit = IntentAndBalancePrivateSettlementCircuit


class ValidBalanceCreateMultiprover(MultiProverCircuit):
    BaseCircuit = ValidBalanceCreate
