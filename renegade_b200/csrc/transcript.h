// Fiat–Shamir transcript of the prover: Keccak-256 "SolidityTranscript".
//
// Replaces the `SolidityTranscript` type parameter of
// `PlonkKzgSnark::prove_with_link_hint::<_, _, SolidityTranscript>`
// (/root/reference/crates/circuits/circuit-types/src/traits.rs:996) and of `link_proofs`
// (circuits-core/src/zk_circuits/proof_linking/intent_only.rs:42-47).  The type itself lives in
// the un-vendored mpc-jellyfish fork (plonk/src/transcript/solidity.rs @311568a4); the byte layout
// below is the published jellyfish 0.4 layout as recalled and is kept in this
// one header so it can be aligned with the fork's source when that is available:
//   * append-only byte buffer + 64-byte state;
//   * field elements / curve coordinates appended as 32-byte big-endian canonical integers;
//   * challenge = from_be_bytes_mod_order(state'[..48]),
//     state' = keccak256(state || transcript || 0x00) || keccak256(state || transcript || 0x01).
// Serial, a few KB per proof: runs on the host between device rounds.
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>

#include "ec.cuh"
#include "ff.cuh"

namespace b200 {

class Keccak256 {
  public:
    static void hash(const uint8_t* data, size_t len, uint8_t out[32]) {
        uint64_t st[25];
        std::memset(st, 0, sizeof(st));
        constexpr size_t rate = 136;
        while (len >= rate) {
            absorb_block(st, data);
            data += rate;
            len -= rate;
        }
        uint8_t last[rate];
        std::memset(last, 0, rate);
        std::memcpy(last, data, len);
        last[len] ^= 0x01;  // Keccak (pre-FIPS) padding, as Ethereum's keccak256
        last[rate - 1] ^= 0x80;
        absorb_block(st, last);
        std::memcpy(out, st, 32);
    }

  private:
    static inline uint64_t rotl(uint64_t x, int s) { return s ? (x << s) | (x >> (64 - s)) : x; }
    static void absorb_block(uint64_t st[25], const uint8_t* block) {
        for (int i = 0; i < 17; ++i) {
            uint64_t lane;
            std::memcpy(&lane, block + 8 * i, 8);
            st[i] ^= lane;
        }
        permute(st);
    }
    static void permute(uint64_t a[25]) {
        static const uint64_t round_constants[24] = {
            0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL,
            0x000000000000808bULL, 0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL,
            0x000000000000008aULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000aULL,
            0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL, 0x8000000000008003ULL,
            0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800aULL, 0x800000008000000aULL,
            0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
        // lane (x, y) lives at a[x + 5y]; walk the pi permutation in place with the rho offsets
        static const int pi_lane[24] = {10, 7, 11, 17, 18, 3, 5, 16, 8, 21, 24, 4, 15, 23, 19, 13, 12, 2, 20, 14, 22, 9, 6, 1};
        static const int rho_off[24] = {1, 3, 6, 10, 15, 21, 28, 36, 45, 55, 2, 14, 27, 41, 56, 8, 25, 43, 62, 18, 39, 61, 20, 44};
        for (int r = 0; r < 24; ++r) {
            uint64_t col[5];
            for (int x = 0; x < 5; ++x) col[x] = a[x] ^ a[x + 5] ^ a[x + 10] ^ a[x + 15] ^ a[x + 20];
            for (int x = 0; x < 5; ++x) {
                const uint64_t d = col[(x + 4) % 5] ^ rotl(col[(x + 1) % 5], 1);
                for (int y = 0; y < 25; y += 5) a[x + y] ^= d;
            }
            uint64_t carry = a[1];
            for (int i = 0; i < 24; ++i) {
                const int j = pi_lane[i];
                const uint64_t tmp = a[j];
                a[j] = rotl(carry, rho_off[i]);
                carry = tmp;
            }
            for (int y = 0; y < 25; y += 5) {
                uint64_t row[5];
                for (int x = 0; x < 5; ++x) row[x] = a[y + x];
                for (int x = 0; x < 5; ++x) a[y + x] = row[x] ^ (~row[(x + 1) % 5] & row[(x + 2) % 5]);
            }
            a[0] ^= round_constants[r];
        }
    }
};

class SolidityTranscript {
  public:
    SolidityTranscript() { std::memset(state_, 0, sizeof(state_)); }

    void append_message(const uint8_t* p, size_t n) { buf_.insert(buf_.end(), p, p + n); }
    void append_u32_be(uint32_t v) {
        const uint8_t b[4] = {(uint8_t)(v >> 24), (uint8_t)(v >> 16), (uint8_t)(v >> 8), (uint8_t)v};
        append_message(b, 4);
    }
    void append_u64_be(uint64_t v) {
        uint8_t b[8];
        for (int i = 0; i < 8; ++i) b[i] = (uint8_t)(v >> (56 - 8 * i));
        append_message(b, 8);
    }
    // Montgomery Fr element -> 32 bytes big-endian canonical
    void append_field_elem(const fe& a_mont) {
        uint8_t b[32];
        to_be(fe_from_mont<FrCfg>(a_mont), b);
        append_message(b, 32);
    }
    // affine commitment (Montgomery Fq coordinates); the identity is 64 zero bytes
    void append_commitment(const g1_affine& c) {
        uint8_t b[32];
        to_be(fe_from_mont<FqCfg>(c.x), b);
        append_message(b, 32);
        to_be(fe_from_mont<FqCfg>(c.y), b);
        append_message(b, 32);
    }
    // get_and_append_challenge: returns the challenge in Montgomery form
    fe get_and_append_challenge() {
        std::vector<uint8_t> in(64 + buf_.size() + 1);
        std::memcpy(in.data(), state_, 64);
        if (!buf_.empty()) std::memcpy(in.data() + 64, buf_.data(), buf_.size());
        uint8_t h0[32], h1[32];
        in.back() = 0;
        Keccak256::hash(in.data(), in.size(), h0);
        in.back() = 1;
        Keccak256::hash(in.data(), in.size(), h1);
        std::memcpy(state_, h0, 32);
        std::memcpy(state_ + 32, h1, 32);
        // big-endian 48-byte integer mod r, by Horner in Montgomery form
        fe acc = fe_zero();
        const fe c256 = fe_from_u32<FrCfg>(256);
        for (int i = 0; i < 48; ++i) acc = fe_add<FrCfg>(fe_mul<FrCfg>(acc, c256), fe_from_u32<FrCfg>(state_[i]));
        return acc;
    }

  private:
    static void to_be(const fe& canon, uint8_t out[32]) {
        for (int i = 0; i < 8; ++i)
            for (int j = 0; j < 4; ++j) out[31 - (4 * i + j)] = (uint8_t)(canon.l[i] >> (8 * j));
    }
    std::vector<uint8_t> buf_;
    uint8_t state_[64];
};

}  // namespace b200
