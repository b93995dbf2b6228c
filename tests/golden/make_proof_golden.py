"""Generates tests/golden/proof_kat.json: proof-level known answers of the ORACLE prover (oracle/plonk_oracle.c) for one
seeded circuit — a regression pin of the restated TurboPlonk rounds, transcript layout and blinding order, NOT a vector of
the Rust reference (none exists and none can be produced here: DESIGN.md section 5).  The CPU suite checks that the oracle
still reproduces it; the GPU suite checks the device proof against the same bytes.

    python tests/golden/make_proof_golden.py
"""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.join(HERE, "..", "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import bn254_py as py  # noqa: E402
import oracle_c as oracle  # noqa: E402
from renegade_b200 import synth  # noqa: E402

PARAMS = {"log_n": 5, "num_inputs": 7, "circuit_seed": 105, "blinder_seed": 0xB11D + 5,
          "tau": "0x2f1a6c0b5d3e49788a9bc0d1e2f30415263748596a7b8c9dae0f1f2e3d4c5b6a"}


def compute(params=PARAMS):
    log_n = params["log_n"]
    circ = synth.synth_circuit(log_n, num_inputs=params["num_inputs"], seed=params["circuit_seed"], check=True)
    tau = oracle.int_to_limbs(py.to_mont(int(params["tau"], 16) % py.R, py.R))
    srs = oracle.srs_from_tau(tau, (1 << log_n) + 3)
    opk = oracle.plonk_preprocess(log_n, circ.selectors, circ.perm, circ.k, srs)
    rc, proof, challenges, link = oracle.plonk_prove(log_n, circ.num_inputs, circ.k, opk, circ.wires, circ.pub_inputs,
                                                     synth.splitmix_blinders(params["blinder_seed"]), srs, True)
    assert rc == 0
    return {"params": params, "proof_hex": bytes(proof).hex(), "proof_sha256": hashlib.sha256(bytes(proof)).hexdigest(),
            "challenges_sha256": hashlib.sha256(bytes(challenges)).hexdigest(),
            "vk_sha256": hashlib.sha256(opk["selector_comms"].tobytes() + opk["sigma_comms"].tobytes()).hexdigest(),
            "link_poly_sha256": hashlib.sha256(link.tobytes()).hexdigest()}


if __name__ == "__main__":
    with open(os.path.join(HERE, "proof_kat.json"), "w") as f:
        json.dump(compute(), f, indent=1)
    print("wrote proof_kat.json")
