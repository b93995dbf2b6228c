/* TEST INFRASTRUCTURE — Keccak-256 (original Keccak padding 0x01, as Ethereum / the `sha3`
 * crate's `Keccak256` used by mpc-jellyfish's SolidityTranscript; Cargo.lock:6047-6075).
 * Restated from the Keccak specification (FIPS 202 permutation, rate 1088, capacity 512). */
#include <stdint.h>
#include <string.h>

#include "plonk_oracle.h"

static const uint64_t RC[24] = {
    0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL,
    0x000000000000808bULL, 0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL,
    0x000000000000008aULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000aULL,
    0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL, 0x8000000000008003ULL,
    0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800aULL, 0x800000008000000aULL,
    0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
static const int ROT[25] = {0,  1,  62, 28, 27, 36, 44, 6,  55, 20, 3,  10, 43,
                            25, 39, 41, 45, 15, 21, 8,  18, 2,  61, 56, 14};

static inline uint64_t rol(uint64_t x, int n) { return n ? (x << n) | (x >> (64 - n)) : x; }

static void keccak_f(uint64_t s[25]) {
    for (int round = 0; round < 24; ++round) {
        uint64_t c[5], d[5], b[25];
        for (int x = 0; x < 5; ++x) c[x] = s[x] ^ s[x + 5] ^ s[x + 10] ^ s[x + 15] ^ s[x + 20];
        for (int x = 0; x < 5; ++x) d[x] = c[(x + 4) % 5] ^ rol(c[(x + 1) % 5], 1);
        for (int i = 0; i < 25; ++i) s[i] ^= d[i % 5];
        /* rho + pi: B[y, 2x+3y] = rot(A[x,y]) with index = x + 5y */
        for (int x = 0; x < 5; ++x)
            for (int y = 0; y < 5; ++y) b[y + 5 * ((2 * x + 3 * y) % 5)] = rol(s[x + 5 * y], ROT[x + 5 * y]);
        for (int y = 0; y < 5; ++y)
            for (int x = 0; x < 5; ++x) s[x + 5 * y] = b[x + 5 * y] ^ (~b[(x + 1) % 5 + 5 * y] & b[(x + 2) % 5 + 5 * y]);
        s[0] ^= RC[round];
    }
}

void orc_keccak256(const uint8_t* data, size_t len, uint8_t out[32]) {
    uint64_t s[25];
    uint8_t block[136];
    memset(s, 0, sizeof(s));
    while (len >= 136) {
        for (int i = 0; i < 17; ++i) {
            uint64_t w;
            memcpy(&w, data + 8 * i, 8);
            s[i] ^= w;
        }
        keccak_f(s);
        data += 136;
        len -= 136;
    }
    memset(block, 0, 136);
    memcpy(block, data, len);
    block[len] ^= 0x01;
    block[135] ^= 0x80;
    for (int i = 0; i < 17; ++i) {
        uint64_t w;
        memcpy(&w, block + 8 * i, 8);
        s[i] ^= w;
    }
    keccak_f(s);
    memcpy(out, s, 32);
}
