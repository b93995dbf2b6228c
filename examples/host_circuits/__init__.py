"""Caller-side example package (NOT part of the product): a Python restatement of the interface the proving path
consumes from the reference's Rust circuit synthesis — mpc-relation's `PlonkCircuit` with the reference's Poseidon2 /
Merkle / state gadgets (`circuit.py`) — and five of the reference's statements built on it (`valid_balance_create.py`,
`private_settlement.py`, `intent_and_balance_validity.py`, `output_balance_validity.py`).

In a deployment synthesis stays in Rust (`SingleProverCircuit::apply_constraints`, circuit-types/src/traits.rs:976-991)
and hands the finished tables to `libb200prover` through the shim crate (shim/gpu-prover).  This package exists so that
tests, the bench and the prover service can prove circuits with the reference's gate structure; it registers its
statements with `renegade_b200.circuit_types.SingleProverCircuit` (see `statements.py`)."""
