"""`SingleProverCircuit` surface over the B200 proving backend — the host-side mirror of
/root/reference/crates/circuits/circuit-types/src/traits.rs:75-92, 817-1020 and
circuits-core/src/lib.rs:112-142, so that callers (and the parity tests) read like the reference's:

    proof          = singleprover_prove(ValidBalanceCreate, witness, statement)
    proof, hint    = singleprover_prove_with_hint(C, witness, statement)
    verify_singleprover_proof(C, statement, proof)         # raises VerifierError
    singleprover_prove_and_verify(C, witness, statement)

What this module owns (the reference's `circuit-types` side of the boundary):
  * `SYSTEM_SRS` (`set_system_srs`): the G1 powers resident on the device + the two G2 elements of the open key;
  * the proving/verifying key cache keyed by `C.name()` (`setup_preprocessed_keys`, traits.rs:821-855: keys are derived
    once per circuit from a dummy instance of the right topology and shared by every prover thread) and the circuit
    layout cache (traits.rs:914-946);
  * prove = synthesise -> finalize -> draw the 17 blinders from the OS RNG (`thread_rng()`, traits.rs:994) ->
    `PlonkKzgSnark.prove_with_link_hint` (device); verify = statement scalars -> `PlonkKzgSnark.verify` (host pairing);
  * the error split `ProverError::{Circuit, Plonk, Verification}` / `VerifierError::Plonk` (errors.rs:33-58).

What it does NOT own: the constraint system.  mpc-relation's `PlonkCircuit` and the gadgets are upstream / caller code
(Rust in a deployment; the Python restatement under examples/host_circuits for tests and benches).  A circuit class
therefore provides `synthesize(witness, statement, layout)` returning any object with `finalize_for_arithmetization()`
(-> tables in the layout of `renegade_b200.synth.SynthCircuit`) and, for link groups, `get_circuit_layout()`.
"""
from __future__ import annotations

import os
import threading
from dataclasses import dataclass
from typing import Any, Dict, Optional, Tuple

import numpy as np

from . import _lib
from .backend import (B200Proof, Bases, Context, LinkingHint, PlonkKzgSnark, ProvingKey, VerifyingKey)
from .fields import SCALAR_FIELD_MODULUS, scalars_to_limbs

PlonkProof = B200Proof                    # circuit-types/src/lib.rs:72-101
ProofLinkingHint = LinkingHint


class ProverError(Exception):
    """circuit-types/src/errors.rs:33-47.  `kind` is one of "Circuit", "Plonk", "Verification"."""

    def __init__(self, kind: str, detail: Any):
        super().__init__(f"{kind}({detail})")
        self.kind, self.detail = kind, detail


class VerifierError(Exception):
    """circuit-types/src/errors.rs:50-58 (`VerifierError::Plonk`)."""


@dataclass
class SystemSrs:
    """`SYSTEM_SRS` (primitives/srs.rs:30-71): `powers_of_g` on the device, `h` / `beta_h` as 128-byte G2 records."""
    ctx: Context
    powers_of_g: Bases
    g2_h: np.ndarray
    g2_tau_h: np.ndarray
    pool: Any = None   # optional ProverPool: proofs then run `workers` at a time, one request thread per ticket


_SYSTEM_SRS: Optional[SystemSrs] = None
_KEY_LOCK = threading.RLock()
_CIRCUIT_KEY_CACHE: Dict[str, Tuple[ProvingKey, VerifyingKey]] = {}
_CIRCUIT_LAYOUT_CACHE: Dict[str, Any] = {}
_CIRCUIT_STRUCTURE: Dict[str, int] = {}  # name -> fingerprint of the structure its key was preprocessed from (0: unknown)


def set_system_srs(ctx: Context, powers_of_g: Bases, g2_h: np.ndarray, g2_tau_h: np.ndarray, pool=None) -> SystemSrs:
    """Install the process-wide SRS (the reference parses srs/srs00 once, lazily).  Clears the key cache: keys
    belong to the SRS they were derived from.  `pool` (a `ProverPool` on the same device, whose worker-0 context is
    `ctx`): proofs are queued there — the shape of `NativeProofManager`'s rayon pool — instead of serialising on `ctx`."""
    global _SYSTEM_SRS
    with _KEY_LOCK:
        clear_key_cache()
        _SYSTEM_SRS = SystemSrs(ctx, powers_of_g, np.ascontiguousarray(g2_h, dtype=np.uint64).reshape(16),
                                np.ascontiguousarray(g2_tau_h, dtype=np.uint64).reshape(16), pool)
        return _SYSTEM_SRS


def system_srs() -> SystemSrs:
    if _SYSTEM_SRS is None:
        raise ProverError("Plonk", "SYSTEM_SRS is not set (call set_system_srs)")
    return _SYSTEM_SRS


def clear_key_cache() -> None:
    with _KEY_LOCK:
        for pk, _ in _CIRCUIT_KEY_CACHE.values():
            try:
                pk.free()
            except Exception:
                pass
        _CIRCUIT_KEY_CACHE.clear()
        _CIRCUIT_LAYOUT_CACHE.clear()
        _CIRCUIT_STRUCTURE.clear()


def draw_blinders(rng=None) -> np.ndarray:
    """The 17 field elements the prover draws (2 per wire polynomial, 3 for the permutation product, 4 for the quotient
    split), Montgomery form.  rng = None: the OS generator, as `thread_rng()` at traits.rs:994; a `random.Random` makes
    a proof reproducible (the parity tests do that on both sides)."""
    if rng is None:
        vals = [int.from_bytes(os.urandom(48), "little") % SCALAR_FIELD_MODULUS for _ in range(17)]  # 384 bits: negligible bias
    else:
        vals = [rng.randrange(SCALAR_FIELD_MODULUS) for _ in range(17)]
    return scalars_to_limbs(vals)


def setup_preprocessed_keys(circuit: type) -> Tuple[ProvingKey, VerifyingKey]:
    """traits.rs:821-855: keys from the cache, else derived from a dummy instance of the circuit's topology."""
    name = circuit.name()
    with _KEY_LOCK:
        hit = _CIRCUIT_KEY_CACHE.get(name)
        if hit is not None:
            return hit
        srs = system_srs()
        witness, statement = circuit.dummy_instance()
        try:
            cs = circuit.synthesize(witness, statement, circuit.get_circuit_layout())
            circ = cs.finalize_for_arithmetization()
        except Exception as e:  # mpc-relation CircuitError
            raise ProverError("Circuit", e) from e
        if circ.n + 3 > len(srs.powers_of_g):
            raise ProverError("Plonk", f"circuit {name} needs {circ.n + 3} SRS powers, {len(srs.powers_of_g)} loaded")
        try:
            pk = PlonkKzgSnark.preprocess(srs.ctx, srs.powers_of_g, circ.log_n, circ.num_inputs, circ.selectors, circ.perm,
                                          circ.k)
        except _lib.B200Error as e:
            raise ProverError("Plonk", e) from e
        pair = (pk, VerifyingKey.from_proving_key(pk, srs.g2_h, srs.g2_tau_h))
        _CIRCUIT_STRUCTURE[name] = int(getattr(circ, "structure_digest", 0) or 0)
        _CIRCUIT_KEY_CACHE[name] = pair
        return pair


class SingleProverCircuit:
    """traits.rs:868-1020.  Subclasses give `name()`, `synthesize`, `statement_scalars`, `dummy_instance` and, when they
    take part in proof linking, `proof_linking_groups()` / `generate_layout`."""

    # ---- what a circuit provides --------------------------------------------------------------------------
    @classmethod
    def name(cls) -> str:
        raise NotImplementedError

    @classmethod
    def synthesize(cls, witness, statement, layout):
        """`PlonkCircuit::new_turbo_plonk()` + link groups of `layout` + `create_witness` / `create_public_var` +
        `apply_constraints` (traits.rs:976-990): returns the constraint system, not yet finalized."""
        raise NotImplementedError

    @classmethod
    def statement_scalars(cls, statement) -> np.ndarray:
        """`statement.to_scalars()` as (num_inputs, 4) Montgomery limbs (traits.rs:1008)."""
        raise NotImplementedError

    @classmethod
    def dummy_instance(cls):
        """(witness, statement) of the circuit's topology for key set-up (the reference uses zeroed scalars)."""
        raise NotImplementedError

    @classmethod
    def proof_linking_groups(cls):
        """[(group id, GroupLayout or None)] — none by default (traits.rs:903-905)."""
        return []

    @classmethod
    def generate_layout(cls):
        """Layout of the link groups (`cs.generate_layout()`, traits.rs:942); None without link groups."""
        return None

    # ---- keys -----------------------------------------------------------------------------------------------
    @classmethod
    def proving_key(cls) -> ProvingKey:
        return setup_preprocessed_keys(cls)[0]

    @classmethod
    def verifying_key(cls) -> VerifyingKey:
        return setup_preprocessed_keys(cls)[1]

    @classmethod
    def get_circuit_layout(cls):
        """traits.rs:914-946: computed once per circuit name."""
        name = cls.name()
        with _KEY_LOCK:
            if name not in _CIRCUIT_LAYOUT_CACHE:
                _CIRCUIT_LAYOUT_CACHE[name] = cls.generate_layout()
            return _CIRCUIT_LAYOUT_CACHE[name]

    # ---- prove / verify ---------------------------------------------------------------------------------------
    @classmethod
    def prove(cls, witness, statement, rng=None) -> PlonkProof:
        return cls.prove_with_link_hint(witness, statement, rng)[0]

    @classmethod
    def prove_with_link_hint(cls, witness, statement, rng=None) -> Tuple[PlonkProof, ProofLinkingHint]:
        pk = cls.proving_key()
        keyed = _CIRCUIT_STRUCTURE.get(cls.name(), 0)
        try:
            cs = cls.synthesize(witness, statement, cls.get_circuit_layout())
            # the key holds the witness-independent structure; per proof only the value tables are arithmetized
            # (SURVEY 8(f) f4), with a fingerprint of the structure to compare against the key's
            try:
                circ = cs.finalize_for_arithmetization(wires_only=True) if keyed else cs.finalize_for_arithmetization()
            except TypeError:  # a constraint system without the wires-only form
                circ = cs.finalize_for_arithmetization()
        except ProverError:
            raise
        except Exception as e:
            raise ProverError("Circuit", e) from e
        if circ.log_n != pk.log_n or circ.num_inputs != pk.num_inputs:
            raise ProverError("Plonk", f"{cls.name()}: instance shape differs from the preprocessed key")
        if keyed and getattr(circ, "structure_digest", 0) and circ.structure_digest != keyed:
            raise ProverError("Circuit", f"{cls.name()}: the instance's gates / wiring differ from the circuit its key was "
                                         "preprocessed from (synthesis must be witness-independent)")
        srs = system_srs()
        try:
            if srs.pool is None:
                return PlonkKzgSnark.prove_with_link_hint(srs.ctx, pk, circ.wires, circ.pub_inputs, draw_blinders(rng))
            wires = np.ascontiguousarray(circ.wires, dtype=np.uint64)
            tk = srs.pool.submit_prove(pk, wires.ctypes.data, circ.pub_inputs, draw_blinders(rng), with_link_poly=True,
                                       keep=wires)
            proof, link = srs.pool.wait(tk)
            return proof, LinkingHint(linking_wire_poly=link,
                                      linking_wire_comm=np.array(proof.wires_poly_comms[0], dtype=np.uint64))
        except _lib.B200Error as e:  # WrongQuotientPolyDegree et al.
            raise ProverError("Plonk", e) from e

    @classmethod
    def verify(cls, statement, proof: PlonkProof) -> None:
        vk = cls.verifying_key()
        try:
            ok = PlonkKzgSnark.verify(vk, cls.statement_scalars(statement), proof)
        except _lib.B200Error as e:
            raise VerifierError(f"Plonk({e})") from e
        if not ok:
            raise VerifierError("Plonk(WrongProof)")


# ---- circuits-core/src/lib.rs:112-142 ------------------------------------------------------------------------------
def singleprover_prove(circuit: type, witness, statement, rng=None) -> PlonkProof:
    return circuit.prove(witness, statement, rng)


def singleprover_prove_with_hint(circuit: type, witness, statement, rng=None) -> Tuple[PlonkProof, ProofLinkingHint]:
    return circuit.prove_with_link_hint(witness, statement, rng)


def verify_singleprover_proof(circuit: type, statement, proof: PlonkProof) -> None:
    circuit.verify(statement, proof)


def singleprover_prove_and_verify(circuit: type, witness, statement, rng=None) -> None:
    proof = circuit.prove(witness, statement, rng)
    try:
        circuit.verify(statement, proof)
    except VerifierError as e:
        raise ProverError("Verification", e) from e


# ---- circuit-types/src/traits.rs:1103-1154, circuits-core/src/lib.rs:145-177 -------------------------------------------
class MultiProverCircuit:
    """`MultiProverCircuit`: the collaborative counterpart of a `SingleProverCircuit` (the reference pairs them: a
    statement's `MultiProverCircuit` impl names its `BaseCircuit`, and the opened proof verifies under the SAME keys,
    traits.rs:1103-1154).  Subclasses set `BaseCircuit`.

    `prove_with_link_hint(witness_shares, statement)`: every party's share of the 5 x n wire table (and of the 17
    blinders) goes into `collaborative.MultiproverPlonkKzgSnark` on the device; what comes back is the OPENED proof and
    hint (`open_authenticated` in `multiprover_prove_and_verify`, lib.rs:166-177).  Producing the wire-table shares —
    running the constraint system on shared values (`MpcPlonkCircuit`) — is the callers' side, like single-prover
    synthesis; `share_witness_table` stands in for it where a test or an example starts from a clear witness."""
    BaseCircuit: type = None

    @classmethod
    def share_witness_table(cls, witness, statement, parties: int = 2, seed: int = 0):
        from .collaborative import share_table
        base = cls.BaseCircuit
        circ = base.synthesize(witness, statement, base.get_circuit_layout()).finalize_for_arithmetization()
        return circ, share_table(np.asarray(circ.wires, dtype=np.uint64).reshape(-1, 4), parties, seed)

    @classmethod
    def prove_with_link_hint(cls, circ, wire_shares, blinder_shares=None, rng=None):
        """`circ`: the finalized public circuit structure (selectors, permutation, public inputs); `wire_shares[p]`:
        party p's share of its wire table.  Returns (PlonkProof, ProofLinkingHint) — opened."""
        from . import collaborative as co
        srs = system_srs()
        be = co.DeviceBackend(srs.ctx, srs.powers_of_g)
        parties = len(wire_shares)
        if blinder_shares is None:
            blinder_shares = co.share_table(draw_blinders(rng), parties, seed=int.from_bytes(os.urandom(8), "little"))
        try:
            cpk = co.CollaborativeProvingKey.build(be, circ.log_n, circ.num_inputs, circ.selectors, circ.perm, circ.k)
            proof, hint, _ = co.MultiproverPlonkKzgSnark.prove_with_link_hint(be, cpk, wire_shares, circ.pub_inputs, blinder_shares)
        except _lib.B200Error as e:
            raise ProverError("Plonk", e) from e
        return proof, hint

    @classmethod
    def verify(cls, statement, proof: PlonkProof) -> None:
        cls.BaseCircuit.verify(statement, proof)


def multiprover_prove_with_hint(circuit: type, circ, wire_shares, blinder_shares=None, rng=None):
    return circuit.prove_with_link_hint(circ, wire_shares, blinder_shares, rng)


def multiprover_prove_and_verify(circuit: type, witness, statement, parties: int = 2, rng=None) -> None:
    """lib.rs:166-177: prove collaboratively, open, verify under the base circuit's verifying key."""
    circ, shares = circuit.share_witness_table(witness, statement, parties)
    proof, _ = circuit.prove_with_link_hint(circ, shares, rng=rng)
    try:
        circuit.verify(statement, proof)
    except VerifierError as e:
        raise ProverError("Verification", e) from e
