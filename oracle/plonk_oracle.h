/* TEST INFRASTRUCTURE — CPU restatement of the TurboPlonk/KZG prover the reference calls at
 * crates/circuits/circuit-types/src/traits.rs:850 (`PlonkKzgSnark::preprocess`) and :996
 * (`PlonkKzgSnark::prove_with_link_hint::<_, _, SolidityTranscript>`), and of the link prover
 * called at circuits-core/src/zk_circuits/proof_linking/intent_only.rs:42-47.
 *
 * The algorithm lives in the un-vendored crate mpc-plonk 0.4.0-pre.0
 * (renegade-fi/mpc-jellyfish @311568a4, Cargo.lock:5024-5026): this file restates the published
 * jellyfish TurboPlonk protocol (5 wire columns, 13 selectors
 * q_lc[4] q_mul[2] q_hash[4] q_o q_c q_ecc, quotient on the 8n coset with shift g = 5, split in 5
 * chunks of n+2, blinding (b0 + b1 X) Z_H on wires, degree-2 blinding on z, Keccak-256
 * "SolidityTranscript") as recalled in SURVEY.md App. A.  PARITY UNPINNED: no proof fixture exists
 * in the reference and the fork's transcript byte layout could not be read; the
 * pinned property is acceptance by the verifier below (itself a restatement) on an SRS with known
 * tau, plus algebraic self-checks.  Transcript layout is isolated in orc_transcript_* so it can be
 * swapped when the fork's source is available.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may use this. */
#ifndef PLONK_ORACLE_H
#define PLONK_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define ORC_N_WIRES 5
#define ORC_N_SELECTORS 13
#define ORC_N_BLINDERS 17 /* 2 per wire, 3 for z, 4 for the quotient split */

/* field order mirrors PlonkProofDef / ProofEvaluationsDef
 * (crates/relayer-types/types-proofs/src/rkyv_impls/plonk_proof_def.rs:168-222) */
typedef struct {
    uint64_t wires_poly_comms[5][8];
    uint64_t prod_perm_poly_comm[8];
    uint64_t split_quot_poly_comms[5][8];
    uint64_t opening_proof[8];
    uint64_t shifted_opening_proof[8];
    uint64_t wires_evals[5][4];
    uint64_t wire_sigma_evals[4][4];
    uint64_t perm_next_eval[4];
} orc_plonk_proof;

typedef struct {
    uint64_t beta[4], gamma[4], alpha[4], zeta[4], v[4], u[4];
} orc_plonk_challenges;

void orc_keccak256(const uint8_t* data, size_t len, uint8_t out[32]);

/* PlonkKzgSnark::preprocess.  selectors_evals: 13 x n values over the domain H (Montgomery);
 * perm: 5n entries, perm[i*n + j] = i'*n + j' (the wire position (i,j) maps to under the copy
 * permutation); k: the 5 coset representatives (Montgomery); srs: >= n + 3 affine points.
 * Outputs: coefficient forms and commitments (the ProvingKey / VerifyingKey contents). */
int orc_plonk_preprocess(unsigned log_n, const uint64_t* selectors_evals, const uint64_t* perm,
                         const uint64_t* k, const uint64_t* srs, uint64_t* selector_coeffs,
                         uint64_t* sigma_coeffs, uint64_t* selector_comms, uint64_t* sigma_comms);

/* PlonkKzgSnark::prove_with_link_hint.  wires: 5 x n wire values (Montgomery); pub_inputs: the
 * values of the first num_inputs gates' public-input wire; blinders: the 17 field elements the
 * reference draws from its RNG, in draw order.  link_poly (n + 2 coefficients, may be NULL)
 * receives the blinded wire-0 polynomial of the LinkingHint.
 * Returns 0, or 2 for WrongQuotientPolyDegree (unsatisfied circuit). */
int orc_plonk_prove(unsigned log_n, size_t num_inputs, const uint64_t* k,
                    const uint64_t* selector_coeffs, const uint64_t* sigma_coeffs,
                    const uint64_t* selector_comms, const uint64_t* sigma_comms,
                    const uint64_t* wires, const uint64_t* pub_inputs, const uint64_t* blinders,
                    const uint64_t* srs, orc_plonk_proof* proof, orc_plonk_challenges* challenges,
                    uint64_t* link_poly);

/* PlonkKzgSnark::verify restated for an SRS whose tau is known (Montgomery Fr): the two KZG
 * opening equations are checked in G1 instead of with a pairing.  Returns 1 = accept. */
int orc_plonk_verify_known_tau(unsigned log_n, size_t num_inputs, const uint64_t* k,
                               const uint64_t* selector_comms, const uint64_t* sigma_comms,
                               const uint64_t* pub_inputs, const orc_plonk_proof* proof,
                               const uint64_t* tau);

/* The same verifier for an SRS with UNKNOWN tau (the reference's real SRS): outputs the two G1
 * operands of the pairing check e(A, [tau]_2) == e(B, [1]_2) that `PlonkKzgSnark::verify` performs
 * (traits.rs:1012-1018); the pairing itself is evaluated by oracle/bn254_pairing_py.py. */
int orc_plonk_verify_operands(unsigned log_n, size_t num_inputs, const uint64_t* k,
                              const uint64_t* selector_comms, const uint64_t* sigma_comms,
                              const uint64_t* pub_inputs, const orc_plonk_proof* proof,
                              uint64_t* out_a, int* a_inf, uint64_t* out_b, int* b_inf);

/* ---- proof linking: restates mpc-plonk `PlonkKzgSnark::link_proofs::<SolidityTranscript>` as
 * called at circuits-core/src/zk_circuits/proof_linking/intent_only.rs:42-47:
 * the two proofs' first wire polynomials a1, a2 agree on the link group's sub-domain
 * D = { g^(offset+i) : i < size }, g = generator of the 2^alignment roots of unity
 * (mpc-relation `GroupLayout { offset, size, alignment }`).
 *   q = (a1 - a2) / Z_D, commit q; eta = transcript(comm1, comm2, comm_q);
 *   opening = commit( (a1 - a2 - Z_D(eta) q) / (X - eta) ).
 * field order mirrors `LinkingProof { quotient_commitment, opening_proof }`
 * (types-proofs/src/mocks.rs:28-30). */
typedef struct {
    uint64_t quotient_commitment[8];
    uint64_t opening_proof[8];
} orc_link_proof;

int orc_plonk_link(const uint64_t* a1, size_t len1, const uint64_t* a2, size_t len2,
                   const uint64_t* comm1, const uint64_t* comm2, unsigned alignment, size_t offset,
                   size_t size, const uint64_t* srs, orc_link_proof* proof, uint64_t* eta_out);
/* verify_link_proof for an SRS with known tau: (tau - eta) * opening == comm1 - comm2 - Z_D(eta) * comm_q */
int orc_plonk_link_verify_known_tau(const uint64_t* comm1, const uint64_t* comm2, unsigned alignment,
                                    size_t offset, size_t size, const orc_link_proof* proof,
                                    const uint64_t* tau);

/* srs[i] = tau^i * G, i < n (test SRS with known tau) */
void orc_srs_from_tau(const uint64_t* tau, size_t n, uint64_t* out_xy);

#ifdef __cplusplus
}
#endif
#endif
