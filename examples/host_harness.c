/* A compiled host that drives libb200prover.so through its C ABI only — the role the Rust shim
 *  plays behind `SingleProverCircuit::prove_with_link_hint`
 * (crates/circuits/circuit-types/src/traits.rs:972-998): load the SRS points, preprocess a circuit,
 * prove, write the proof.  No Python, no torch: plain pointers and sizes.
 *
 * Job file (little-endian), written by tests/test_gpu_harness.py:
 *   u64 log_n, u64 num_inputs, u64 n_srs,
 *   k[5][4], selectors[13][n][4], perm[5n], wires[5][n][4], pub_inputs[num_inputs][4], blinders[17][4],
 *   srs[n_srs][8]                                     (all u64)
 * Output file: b200_proof (1152 bytes) || link polynomial ((n+2) x 4 u64) || 18 VK commitments.
 *
 *   gcc -O2 -Iinclude examples/host_harness.c -Lrenegade_b200 -lb200prover -Wl,-rpath,'$ORIGIN/../renegade_b200' -o examples/host_harness
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "b200prover.h"

#define CHECK(call)                                                                  \
    do {                                                                             \
        int rc_ = (call);                                                            \
        if (rc_ != B200_OK) {                                                        \
            fprintf(stderr, "%s failed: %d (%s)\n", #call, rc_, b200_last_error()); \
            return 2;                                                                \
        }                                                                            \
    } while (0)

static uint64_t* read_u64s(FILE* f, size_t count) {
    uint64_t* p = (uint64_t*)malloc(count * 8 + 8);
    if (!p || fread(p, 8, count, f) != count) {
        fprintf(stderr, "short job file\n");
        exit(3);
    }
    return p;
}

int main(int argc, char** argv) {
    if (argc != 3) {
        fprintf(stderr, "usage: %s job.bin proof.bin\n", argv[0]);
        return 1;
    }
    FILE* f = fopen(argv[1], "rb");
    if (!f) { perror("job"); return 1; }
    uint64_t hdr[3];
    if (fread(hdr, 8, 3, f) != 3) return 3;
    const unsigned log_n = (unsigned)hdr[0];
    const size_t num_inputs = (size_t)hdr[1], n_srs = (size_t)hdr[2], n = (size_t)1 << log_n;
    uint64_t* k = read_u64s(f, 5 * 4);
    uint64_t* selectors = read_u64s(f, 13 * n * 4);
    uint64_t* perm = read_u64s(f, 5 * n);
    uint64_t* wires = read_u64s(f, 5 * n * 4);
    uint64_t* pub_inputs = read_u64s(f, num_inputs * 4);
    uint64_t* blinders = read_u64s(f, 17 * 4);
    uint64_t* srs = read_u64s(f, n_srs * 8);
    fclose(f);

    b200_ctx* ctx = NULL;
    b200_bases* bases = NULL;
    b200_pk* pk = NULL;
    CHECK(b200_init(0, &ctx));
    fprintf(stderr, "%s\n", b200_version());
    CHECK(b200_bases_load(ctx, (const uint8_t*)srs, n_srs, 0, /*check_on_curve=*/1, &bases));
    CHECK(b200_plonk_preprocess(ctx, bases, log_n, num_inputs, selectors, perm, k, &pk));
    uint64_t vk[18 * 8];
    CHECK(b200_pk_verifying_key(pk, vk, vk + 13 * 8));

    b200_proof proof;
    uint64_t* link = (uint64_t*)calloc((n + 2) * 4, 8);
    CHECK(b200_plonk_prove(ctx, pk, wires, pub_inputs, blinders, &proof, link, NULL));

    /* an unsatisfying witness must come back as an error code, never an abort */
    wires[4 * (4 * n + num_inputs + 1)] ^= 1u;
    b200_proof bad;
    int rc = b200_plonk_prove(ctx, pk, wires, pub_inputs, blinders, &bad, NULL, NULL);
    if (rc != B200_ERR_UNSATISFIED) {
        fprintf(stderr, "expected B200_ERR_UNSATISFIED, got %d\n", rc);
        return 4;
    }

    /* the same proof through the prover pool (the NativeProofManager-shaped entry point): four jobs on
     * two workers share the key; every result must equal the single-context proof byte for byte */
    wires[4 * (4 * n + num_inputs + 1)] ^= 1u; /* restore the good witness */
    b200_pool* pool = NULL;
    CHECK(b200_pool_create(0, 2, &pool));
    b200_proof pooled[4];
    uint64_t tickets[4];
    for (int i = 0; i < 4; ++i)
        CHECK(b200_pool_submit_prove(pool, pk, wires, pub_inputs, num_inputs, blinders, &pooled[i], NULL, &tickets[i]));
    for (int i = 3; i >= 0; --i) CHECK(b200_pool_wait(pool, tickets[i]));
    for (int i = 0; i < 4; ++i)
        if (memcmp(&pooled[i], &proof, sizeof(proof)) != 0) {
            fprintf(stderr, "pool proof %d differs from the direct proof\n", i);
            return 5;
        }
    b200_pool_destroy(pool);

    FILE* o = fopen(argv[2], "wb");
    if (!o) { perror("out"); return 1; }
    fwrite(&proof, sizeof(proof), 1, o);
    fwrite(link, 8, (n + 2) * 4, o);
    fwrite(vk, 8, 18 * 8, o);
    fclose(o);

    b200_pk_free(ctx, pk);
    b200_bases_free(ctx, bases);
    b200_shutdown(ctx);
    fprintf(stderr, "proof written (%zu bytes)\n", sizeof(proof));
    return 0;
}
