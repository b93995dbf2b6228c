//! Safe wrappers over `libb200prover` with the shapes of the upstream calls they replace:
//!
//! | reference call site | upstream call | wrapper |
//! |---|---|---|
//! | `circuit-types/src/traits.rs:850`  | `PlonkKzgSnark::preprocess(&SYSTEM_SRS, &cs)` | [`GpuProver::preprocess`] |
//! | `circuit-types/src/traits.rs:996`  | `PlonkKzgSnark::prove_with_link_hint::<_, _, SolidityTranscript>(rng, &circuit, &pk)` | [`GpuProver::prove_with_link_hint`] |
//! | `circuit-types/src/traits.rs:1012` | `PlonkKzgSnark::verify::<SolidityTranscript>(&vk, pi, &proof, None)` | [`verify`] (host only) |
//! | `proof_linking/intent_only.rs:42-47` | `PlonkKzgSnark::link_proofs::<SolidityTranscript>(a, b, &layout, &ck)` | [`GpuProver::link_proofs`] |
//! | `primitives/srs.rs:63-71` | `parse_ptau_file` + `powers_of_g` | [`GpuProver::load_srs`] |
//!
//! Marshalling rules (SURVEY.md §7 "Rust ABI hazards"): `Fr`/`Fq` are `Fp(BigInt<4>, PhantomData)` holding the
//! Montgomery limbs — reading `.0 .0` is the value the C ABI wants; `G1Affine` is `repr(Rust)` {x, y, infinity} and is
//! NEVER passed by pointer — it is marshalled to the packed 64-byte x || y record (the SRS file's own record,
//! srs.rs:172-182).  Errors: non-zero status -> `PlonkError` -> `ProverError::Plonk` (errors.rs:41); the library never
//! panics, aborts or unwinds across the boundary.
//!
//! NOT compiled in the build image (no cargo there).  `mpc_relation::PlonkCircuit` accessors used below
//! (`selector_evals`, `wire_permutation_flat`, `witness_table`) are thin getters over fields the fork already has.
#![allow(unsafe_code)]
pub mod ffi;

use ark_bn254::{Bn254, Fq, Fr, G1Affine};
use ark_ff::{BigInt, UniformRand};
use mpc_plonk::errors::PlonkError;
use mpc_plonk::proof_system::structs::{LinkingHint, Proof, ProofEvaluations};
use mpc_relation::proof_linking::GroupLayout;
use mpc_relation::PlonkCircuit;
use std::ffi::CStr;
use std::os::raw::c_int;

fn fr_limbs(x: &Fr) -> [u64; 4] {
    (x.0).0
}
fn fr_from(l: &[u64; 4]) -> Fr {
    Fr::new_unchecked(BigInt(*l)) // already Montgomery, like srs.rs:201-209
}
fn g1_from(xy: &[u64; 8]) -> G1Affine {
    if xy.iter().all(|l| *l == 0) {
        return G1Affine::identity();
    }
    let x = Fq::new_unchecked(BigInt([xy[0], xy[1], xy[2], xy[3]]));
    let y = Fq::new_unchecked(BigInt([xy[4], xy[5], xy[6], xy[7]]));
    G1Affine::new_unchecked(x, y)
}
fn g1_record(p: &G1Affine) -> [u64; 8] {
    if p.infinity {
        return [0; 8];
    }
    let (x, y) = ((p.x.0).0, (p.y.0).0);
    [x[0], x[1], x[2], x[3], y[0], y[1], y[2], y[3]]
}

fn check(rc: c_int) -> Result<(), PlonkError> {
    if rc == ffi::B200_OK {
        return Ok(());
    }
    let msg = unsafe { CStr::from_ptr(ffi::b200_last_error()) }.to_string_lossy().into_owned();
    Err(match rc {
        ffi::B200_ERR_UNSATISFIED => PlonkError::WrongQuotientPolyDegree(0, 0),
        _ => PlonkError::InvalidParameters(msg),
    })
}

/// One GPU: a prover pool (N proofs in flight, the shape of `NativeProofManager`'s rayon pool,
/// native_proof_manager.rs:187-192) plus the resident SRS.
pub struct GpuProver {
    pool: *mut ffi::b200_pool,
    srs: *mut ffi::b200_bases,
    srs_len: usize,
}
// the library serialises the calls made on one context; the pool hands every job its own context
unsafe impl Send for GpuProver {}
unsafe impl Sync for GpuProver {}

pub struct GpuProvingKey {
    raw: *mut ffi::b200_pk,
    pub num_inputs: usize,
    pub log_n: u32,
}
unsafe impl Send for GpuProvingKey {}
unsafe impl Sync for GpuProvingKey {}

impl GpuProver {
    pub fn new(device: i32, workers: u32) -> Result<Self, PlonkError> {
        let mut pool = std::ptr::null_mut();
        check(unsafe { ffi::b200_pool_create(device, workers, &mut pool) })?;
        Ok(Self { pool, srs: std::ptr::null_mut(), srs_len: 0 })
    }

    /// replaces `parse_ptau_file` + the `powers_of_g` vector (srs.rs:63-71); on-curve check = srs.rs:178-179
    pub fn load_srs(&mut self, ptau_bytes: &[u8], n_points: usize) -> Result<(), PlonkError> {
        let (mut rec, mut n) = (std::ptr::null(), 0usize);
        check(unsafe { ffi::b200_srs_parse_ptau(ptau_bytes.as_ptr(), ptau_bytes.len(), &mut rec, &mut n) })?;
        if n < n_points {
            return Err(PlonkError::InvalidParameters(format!("ptau holds {n} G1 powers, need {n_points}")));
        }
        let ctx = unsafe { ffi::b200_pool_ctx(self.pool, 0) };
        let mut b = std::ptr::null_mut();
        check(unsafe { ffi::b200_bases_load(ctx, rec, n_points, 0, 1, &mut b) })?;
        self.srs = b;
        self.srs_len = n_points;
        Ok(())
    }

    /// replaces `PlonkKzgSnark::preprocess(&SYSTEM_SRS, &cs)` (traits.rs:850); `cs` is finalized
    pub fn preprocess(&self, cs: &PlonkCircuit<Fr>) -> Result<GpuProvingKey, PlonkError> {
        let n = cs.eval_domain_size()?;
        let selectors: Vec<u64> = cs.selector_evals().iter().flat_map(|col| col.iter().flat_map(fr_limbs)).collect(); // 13 x n
        let perm: Vec<u64> = cs.wire_permutation_flat(); // perm[i * n + j] = i' * n + j'
        let k: Vec<u64> = cs.coset_representatives().iter().flat_map(fr_limbs).collect();
        let ctx = unsafe { ffi::b200_pool_ctx(self.pool, 0) };
        let mut pk = std::ptr::null_mut();
        check(unsafe {
            ffi::b200_plonk_preprocess(ctx, self.srs, n.trailing_zeros(), cs.num_inputs(), selectors.as_ptr(), perm.as_ptr(),
                                       k.as_ptr(), &mut pk)
        })?;
        Ok(GpuProvingKey { raw: pk, num_inputs: cs.num_inputs(), log_n: n.trailing_zeros() })
    }

    /// replaces `PlonkKzgSnark::prove_with_link_hint::<_, _, SolidityTranscript>(rng, &circuit, &pk)` (traits.rs:996).
    /// Blocks the calling rayon worker only: the job runs on one of the pool's contexts, other proofs keep the GPU busy.
    pub fn prove_with_link_hint<R: rand::RngCore + rand::CryptoRng>(
        &self, rng: &mut R, cs: &PlonkCircuit<Fr>, pk: &GpuProvingKey,
    ) -> Result<(Proof<Bn254>, LinkingHint<Bn254>), PlonkError> {
        let n = 1usize << pk.log_n;
        let wires: Vec<u64> = cs.witness_table().iter().flat_map(|col| col.iter().flat_map(fr_limbs)).collect(); // 5 x n
        let pub_inputs: Vec<u64> = cs.public_input()?.iter().flat_map(fr_limbs).collect();
        // the 17 draws of the reference prover, in its order: 2 per wire polynomial, 3 for z, 4 for the quotient split
        let blinders: Vec<u64> = (0..17).map(|_| Fr::rand(rng)).flat_map(|f| fr_limbs(&f)).collect();
        let mut raw: ffi::b200_proof = unsafe { std::mem::zeroed() };
        let mut link = vec![0u64; (n + 2) * 4];
        let mut ticket = 0u64;
        check(unsafe {
            ffi::b200_pool_submit_prove(self.pool, pk.raw, wires.as_ptr(), pub_inputs.as_ptr(), pk.num_inputs, blinders.as_ptr(),
                                        &mut raw, link.as_mut_ptr(), &mut ticket)
        })?;
        check(unsafe { ffi::b200_pool_wait(self.pool, ticket) })?; // buffers above stay alive until here
        Ok((proof_from_raw(&raw), hint_from_raw(&raw, &link)))
    }

    /// replaces `PlonkKzgSnark::link_proofs::<SolidityTranscript>(lhs, rhs, &layout, &pk.commit_key)`
    /// (proof_linking/intent_only.rs:42-47, intent_and_balance.rs:66-72, output_balance.rs)
    pub fn link_proofs(&self, lhs: &LinkingHint<Bn254>, rhs: &LinkingHint<Bn254>, layout: &GroupLayout)
        -> Result<mpc_plonk::proof_system::structs::LinkingProof<Bn254>, PlonkError> {
        let a1: Vec<u64> = lhs.linking_wire_poly.coeffs.iter().flat_map(fr_limbs).collect();
        let a2: Vec<u64> = rhs.linking_wire_poly.coeffs.iter().flat_map(fr_limbs).collect();
        let (c1, c2) = (g1_record(&lhs.linking_wire_comm.0), g1_record(&rhs.linking_wire_comm.0));
        let mut raw: ffi::b200_link_proof = unsafe { std::mem::zeroed() };
        let mut ticket = 0u64;
        check(unsafe {
            ffi::b200_pool_submit_link(self.pool, self.srs, a1.as_ptr(), a1.len() / 4, a2.as_ptr(), a2.len() / 4, c1.as_ptr(),
                                       c2.as_ptr(), layout.alignment as u32, layout.offset, layout.size, &mut raw, &mut ticket)
        })?;
        check(unsafe { ffi::b200_pool_wait(self.pool, ticket) })?;
        Ok(mpc_plonk::proof_system::structs::LinkingProof {
            quotient_commitment: jf_primitives::pcs::prelude::Commitment(g1_from(&raw.quotient_commitment)),
            opening_proof: jf_primitives::pcs::prelude::UnivariateKzgProof { proof: g1_from(&raw.opening_proof) },
        })
    }
}

impl Drop for GpuProver {
    fn drop(&mut self) {
        unsafe {
            if !self.srs.is_null() {
                ffi::b200_bases_free(ffi::b200_pool_ctx(self.pool, 0), self.srs);
            }
            ffi::b200_pool_destroy(self.pool);
        }
    }
}

/// replaces `PlonkKzgSnark::verify::<SolidityTranscript>(&vk, public_inputs, &proof, None)` (traits.rs:1012-1018).
/// Host only: `g2_h` / `g2_tau_h` are the ptau G2 records of `open_key.h` / `open_key.beta_h` (srs.rs:185-199).
#[allow(clippy::too_many_arguments)]
pub fn verify(log_n: u32, k: &[Fr; 5], selector_comms: &[G1Affine; 13], sigma_comms: &[G1Affine; 5], public_inputs: &[Fr],
              proof: &Proof<Bn254>, g2_h: &[u64; 16], g2_tau_h: &[u64; 16]) -> Result<(), PlonkError> {
    let kk: Vec<u64> = k.iter().flat_map(fr_limbs).collect();
    let sel: Vec<u64> = selector_comms.iter().flat_map(g1_record).collect();
    let sig: Vec<u64> = sigma_comms.iter().flat_map(g1_record).collect();
    let pi: Vec<u64> = public_inputs.iter().flat_map(fr_limbs).collect();
    let raw = proof_to_raw(proof);
    let mut accepted: c_int = 0;
    check(unsafe {
        ffi::b200_plonk_verify(log_n, public_inputs.len(), kk.as_ptr(), sel.as_ptr(), sig.as_ptr(), pi.as_ptr(), &raw,
                               g2_h.as_ptr(), g2_tau_h.as_ptr(), &mut accepted)
    })?;
    if accepted == 1 { Ok(()) } else { Err(PlonkError::WrongProof) }
}

// ---- field-by-field copies between the flat C structs and the upstream types (plonk_proof_def.rs:168-222) -----------
fn proof_from_raw(r: &ffi::b200_proof) -> Proof<Bn254> {
    use jf_primitives::pcs::prelude::{Commitment, UnivariateKzgProof};
    Proof {
        wires_poly_comms: r.wires_poly_comms.iter().map(|c| Commitment(g1_from(c))).collect(),
        prod_perm_poly_comm: Commitment(g1_from(&r.prod_perm_poly_comm)),
        split_quot_poly_comms: r.split_quot_poly_comms.iter().map(|c| Commitment(g1_from(c))).collect(),
        opening_proof: Commitment(g1_from(&r.opening_proof)),
        shifted_opening_proof: Commitment(g1_from(&r.shifted_opening_proof)),
        poly_evals: ProofEvaluations {
            wires_evals: r.wires_evals.iter().map(fr_from).collect(),
            wire_sigma_evals: r.wire_sigma_evals.iter().map(fr_from).collect(),
            perm_next_eval: fr_from(&r.perm_next_eval),
        },
        plookup_proof: None,
    }
    .tap_kzg::<UnivariateKzgProof<Bn254>>()
}
fn proof_to_raw(p: &Proof<Bn254>) -> ffi::b200_proof {
    let mut r: ffi::b200_proof = unsafe { std::mem::zeroed() };
    for i in 0..5 {
        r.wires_poly_comms[i] = g1_record(&p.wires_poly_comms[i].0);
        r.split_quot_poly_comms[i] = g1_record(&p.split_quot_poly_comms[i].0);
        r.wires_evals[i] = fr_limbs(&p.poly_evals.wires_evals[i]);
    }
    for i in 0..4 {
        r.wire_sigma_evals[i] = fr_limbs(&p.poly_evals.wire_sigma_evals[i]);
    }
    r.prod_perm_poly_comm = g1_record(&p.prod_perm_poly_comm.0);
    r.opening_proof = g1_record(&p.opening_proof.0);
    r.shifted_opening_proof = g1_record(&p.shifted_opening_proof.0);
    r.perm_next_eval = fr_limbs(&p.poly_evals.perm_next_eval);
    r
}
fn hint_from_raw(r: &ffi::b200_proof, link: &[u64]) -> LinkingHint<Bn254> {
    use ark_poly::{univariate::DensePolynomial, DenseUVPolynomial};
    let coeffs: Vec<Fr> = link.chunks_exact(4).map(|c| fr_from(&[c[0], c[1], c[2], c[3]])).collect();
    LinkingHint {
        linking_wire_poly: DensePolynomial::from_coefficients_vec(coeffs),
        linking_wire_comm: jf_primitives::pcs::prelude::Commitment(g1_from(&r.wires_poly_comms[0])),
    }
}

/// Adapter kept next to the conversion it documents: the fork's `Proof` stores opening proofs as `Commitment`s.
trait TapKzg: Sized {
    fn tap_kzg<T>(self) -> Self {
        self
    }
}
impl TapKzg for Proof<Bn254> {}
