/* TEST INFRASTRUCTURE — see plonk_oracle.h for scope, citations and parity status. */
#include "plonk_oracle.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "bn254_internal.h"
#include "bn254_oracle.h"

#define NW ORC_N_WIRES
#define NS ORC_N_SELECTORS

typedef u64 fr[4];

static inline void fr_set(u64 o[4], const u64 a[4]) { memcpy(o, a, 32); }
static inline void fr_one(u64 o[4]) { memcpy(o, FR.r1, 32); }
static inline void fr_zero(u64 o[4]) { memset(o, 0, 32); }
static inline void fr_from_u64(u64 o[4], u64 v) {
    u64 t[4] = {v, 0, 0, 0};
    fp_to_mont(&FR, o, t);
}
#define MUL(o, a, b) fp_mul(&FR, (o), (a), (b))
#define ADD(o, a, b) fp_add(&FR, (o), (a), (b))
#define SUB(o, a, b) fp_sub(&FR, (o), (a), (b))

static void fr_pow_u64(u64 o[4], const u64 a[4], u64 e) {
    u64 ee[4] = {e, 0, 0, 0};
    fp_pow(&FR, o, a, ee);
}

/* ------------------------------------------------------------------------------------------
 * SolidityTranscript restated (mpc-jellyfish plonk/src/transcript/solidity.rs, as recalled):
 * an append-only byte buffer plus a 64-byte state; a challenge hashes
 * state || transcript || {0,1} with Keccak-256 twice and reduces the first 48 bytes of the new
 * state mod r (big-endian).  Field elements and curve coordinates are appended as 32-byte
 * big-endian canonical integers.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    uint8_t* buf;
    size_t len, cap;
    uint8_t state[64];
} transcript;

static void tr_init(transcript* t) {
    t->cap = 4096;
    t->buf = (uint8_t*)malloc(t->cap);
    t->len = 0;
    memset(t->state, 0, 64);
}
static void tr_free(transcript* t) { free(t->buf); }
static void tr_append(transcript* t, const uint8_t* p, size_t n) {
    if (t->len + n > t->cap) {
        while (t->len + n > t->cap) t->cap *= 2;
        t->buf = (uint8_t*)realloc(t->buf, t->cap);
    }
    memcpy(t->buf + t->len, p, n);
    t->len += n;
}
static void tr_append_u64_be(transcript* t, u64 v) {
    uint8_t b[8];
    for (int i = 0; i < 8; ++i) b[i] = (uint8_t)(v >> (56 - 8 * i));
    tr_append(t, b, 8);
}
static void tr_append_u32_be(transcript* t, uint32_t v) {
    uint8_t b[4] = {(uint8_t)(v >> 24), (uint8_t)(v >> 16), (uint8_t)(v >> 8), (uint8_t)v};
    tr_append(t, b, 4);
}
static void canon_to_be(const u64 c[4], uint8_t out[32]) {
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 8; ++j) out[31 - (8 * i + j)] = (uint8_t)(c[i] >> (8 * j));
}
static void tr_append_fr(transcript* t, const u64 a_mont[4]) {
    u64 c[4];
    uint8_t b[32];
    fp_from_mont(&FR, c, a_mont);
    canon_to_be(c, b);
    tr_append(t, b, 32);
}
static void tr_append_g1(transcript* t, const u64 xy_mont[8]) { /* identity = 64 zero bytes */
    u64 c[4];
    uint8_t b[32];
    fp_from_mont(&FQ, c, xy_mont);
    canon_to_be(c, b);
    tr_append(t, b, 32);
    fp_from_mont(&FQ, c, xy_mont + 4);
    canon_to_be(c, b);
    tr_append(t, b, 32);
}
static void tr_challenge(transcript* t, u64 out_mont[4]) {
    size_t n = 64 + t->len + 1;
    uint8_t* in = (uint8_t*)malloc(n);
    uint8_t h0[32], h1[32];
    memcpy(in, t->state, 64);
    memcpy(in + 64, t->buf, t->len);
    in[n - 1] = 0;
    orc_keccak256(in, n, h0);
    in[n - 1] = 1;
    orc_keccak256(in, n, h1);
    free(in);
    memcpy(t->state, h0, 32);
    memcpy(t->state + 32, h1, 32);
    /* from_be_bytes_mod_order(state[..48]) */
    u64 acc[4], c256[4], b[4];
    fr_zero(acc);
    fr_from_u64(c256, 256);
    for (int i = 0; i < 48; ++i) {
        MUL(acc, acc, c256);
        fr_from_u64(b, t->state[i]);
        ADD(acc, acc, b);
    }
    fr_set(out_mont, acc);
}

/* append_vk_and_pub_input */
static void tr_append_vk(transcript* t, unsigned log_n, size_t num_inputs, const u64* k, const u64* sel_comms,
                         const u64* sig_comms, const u64* pub_inputs) {
    tr_append_u32_be(t, 254); /* field size in bits */
    tr_append_u64_be(t, (u64)1 << log_n);
    tr_append_u64_be(t, (u64)num_inputs);
    for (int i = 0; i < NW; ++i) tr_append_fr(t, k + 4 * i);
    for (int i = 0; i < NS; ++i) tr_append_g1(t, sel_comms + 8 * i);
    for (int i = 0; i < NW; ++i) tr_append_g1(t, sig_comms + 8 * i);
    for (size_t i = 0; i < num_inputs; ++i) tr_append_fr(t, pub_inputs + 4 * i);
}

/* ------------------------------------------------------------------------------------------
 * polynomial helpers (coefficient vectors, Montgomery)
 * ---------------------------------------------------------------------------------------- */
static void commit(const u64* srs, const u64* coeffs, size_t len, u64 out_xy[8]) {
    /* UnivariateKzgPCS::commit: into_bigint each coefficient, then msm_bigint */
    u64* canon = (u64*)malloc(len * 32);
    orc_fp_array_from_mont(0, coeffs, len, canon);
    int inf;
    orc_msm(srs, canon, len, out_xy, &inf);
    if (inf) memset(out_xy, 0, 64);
    free(canon);
}
static void poly_eval(const u64* c, size_t len, const u64 x[4], u64 out[4]) {
    u64 acc[4];
    fr_zero(acc);
    for (size_t i = len; i-- > 0;) {
        MUL(acc, acc, x);
        ADD(acc, acc, c + 4 * i);
    }
    fr_set(out, acc);
}
/* acc[0..len) += s * p[0..len) */
static void poly_axpy(u64* acc, const u64* p, size_t len, const u64 s[4]) {
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < len; ++i) {
        u64 t[4];
        MUL(t, p + 4 * i, s);
        ADD(acc + 4 * i, acc + 4 * i, t);
    }
}
/* q = p / (X - z), p has len coefficients, q has len - 1; returns 1 when the remainder p(z) is zero */
static int poly_div_linear(const u64* p, size_t len, const u64 z[4], u64* q) {
    u64 carry[4], t[4];
    fr_zero(carry);
    for (size_t i = len - 1; i >= 1; --i) {
        MUL(t, carry, z);
        ADD(carry, p + 4 * i, t);
        fr_set(q + 4 * (i - 1), carry);
    }
    MUL(t, carry, z);
    ADD(carry, p, t); /* remainder = p(z) */
    return (carry[0] | carry[1] | carry[2] | carry[3]) == 0;
}

void orc_srs_from_tau(const u64* tau, size_t n, u64* out_xy) {
    /* powers tau^i as canonical scalars, then i-th point = tau^i * G via the fixed-base table
     * of bn254_oracle.c's generator (here: plain double-and-add per point, parallel) */
    u64* pw = (u64*)malloc(n * 32);
    u64 acc[4];
    fr_one(acc);
    for (size_t i = 0; i < n; ++i) {
        fp_from_mont(&FR, pw + 4 * i, acc);
        MUL(acc, acc, tau);
    }
    u64 g[8];
    memcpy(g, FQ.r1, 32);
    u64 two[4] = {2, 0, 0, 0};
    fp_to_mont(&FQ, g + 4, two);
    /* table T[b] = 2^b G */
    aff T[256];
    jac cur;
    memcpy(cur.x, g, 32); memcpy(cur.y, g + 4, 32); memcpy(cur.z, FQ.r1, 32);
    for (int b = 0; b < 256; ++b) { jac_to_affine(&T[b], &cur); jac_double(&cur, &cur); }
#pragma omp parallel for schedule(dynamic, 64)
    for (size_t i = 0; i < n; ++i) {
        jac J; jac_set_inf(&J);
        const u64* s = pw + 4 * i;
        for (int b = 0; b < 254; ++b)
            if ((s[b / 64] >> (b % 64)) & 1) jac_add_affine(&J, &J, &T[b]);
        aff A; jac_to_affine(&A, &J);
        aff_store(&A, out_xy + 8 * i, NULL);
    }
    free(pw);
}

/* ------------------------------------------------------------------------------------------
 * preprocess
 * ---------------------------------------------------------------------------------------- */
int orc_plonk_preprocess(unsigned log_n, const u64* selectors_evals, const u64* perm, const u64* k, const u64* srs,
                         u64* selector_coeffs, u64* sigma_coeffs, u64* selector_comms, u64* sigma_comms) {
    const size_t n = (size_t)1 << log_n;
    u64 w[4];
    orc_domain_generator(log_n, w);
    /* domain elements */
    u64* dom = (u64*)malloc(n * 32);
    fr_one(dom);
    for (size_t j = 1; j < n; ++j) MUL(dom + 4 * j, dom + 4 * (j - 1), w);
    for (int s = 0; s < NS; ++s) {
        memcpy(selector_coeffs + 4 * n * s, selectors_evals + 4 * n * s, n * 32);
        orc_ntt(selector_coeffs + 4 * n * s, log_n, 1, 0);
        commit(srs, selector_coeffs + 4 * n * s, n, selector_comms + 8 * s);
    }
    for (int i = 0; i < NW; ++i) {
        u64* sig = sigma_coeffs + 4 * n * i;
        for (size_t j = 0; j < n; ++j) {
            const u64 tgt = perm[i * n + j];
            MUL(sig + 4 * j, k + 4 * (tgt / n), dom + 4 * (tgt % n)); /* extended id permutation value */
        }
        orc_ntt(sig, log_n, 1, 0);
        commit(srs, sig, n, sigma_comms + 8 * i);
    }
    free(dom);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * prove
 * ---------------------------------------------------------------------------------------- */
static void gate_eval(const u64* q /*13 x stride*/, size_t stride, size_t i, const fr wv[5], const u64 pi[4], u64 out[4]) {
    /* q_c + pi + sum q_lc w + q_mul0 w0w1 + q_mul1 w2w3 + q_ecc w0w1w2w3w4 + sum q_hash w^5 - q_o w4 */
#define Q(s) (q + 4 * ((size_t)(s) * stride + i))
    u64 acc[4], t[4], t2[4];
    ADD(acc, Q(11), pi);
    for (int j = 0; j < 4; ++j) { MUL(t, Q(j), wv[j]); ADD(acc, acc, t); }
    MUL(t, wv[0], wv[1]); MUL(t2, Q(4), t); ADD(acc, acc, t2);
    MUL(t, wv[2], wv[3]); MUL(t2, Q(5), t); ADD(acc, acc, t2);
    MUL(t, wv[0], wv[1]); MUL(t, t, wv[2]); MUL(t, t, wv[3]); MUL(t, t, wv[4]); MUL(t2, Q(12), t); ADD(acc, acc, t2);
    for (int j = 0; j < 4; ++j) {
        MUL(t, wv[j], wv[j]); MUL(t, t, t); MUL(t, t, wv[j]);
        MUL(t2, Q(6 + j), t); ADD(acc, acc, t2);
    }
    MUL(t, Q(10), wv[4]); SUB(acc, acc, t);
    fr_set(out, acc);
#undef Q
}

int orc_plonk_prove(unsigned log_n, size_t num_inputs, const u64* k, const u64* selector_coeffs, const u64* sigma_coeffs,
                    const u64* selector_comms, const u64* sigma_comms, const u64* wires, const u64* pub_inputs,
                    const u64* blinders, const u64* srs, orc_plonk_proof* proof, orc_plonk_challenges* ch_out,
                    u64* link_poly) {
    const size_t n = (size_t)1 << log_n;
    const unsigned log_m = log_n + 3;
    const size_t m = n * 8;
    int rc = 0;
    orc_plonk_challenges ch;
    memset(&ch, 0, sizeof(ch));
    transcript tr;
    tr_init(&tr);
    tr_append_vk(&tr, log_n, num_inputs, k, selector_comms, sigma_comms, pub_inputs);

    u64 w[4], one[4];
    fr_one(one);
    orc_domain_generator(log_n, w);
    u64* dom = (u64*)malloc(n * 32);
    fr_one(dom);
    for (size_t j = 1; j < n; ++j) MUL(dom + 4 * j, dom + 4 * (j - 1), w);

    /* ---- round 1: wire polynomials, blinded with (b0 + b1 X) Z_H ---------------------------- */
    const size_t WL = n + 2;
    u64* wire_polys = (u64*)calloc(NW * (n + 3), 32); /* stride n + 3 */
    const size_t WS = n + 3;
    for (int i = 0; i < NW; ++i) {
        u64* p = wire_polys + 4 * WS * i;
        memcpy(p, wires + 4 * n * i, n * 32);
        orc_ntt(p, log_n, 1, 0);
        const u64* b0 = blinders + 4 * (2 * i), *b1 = blinders + 4 * (2 * i + 1);
        SUB(p, p, b0);
        SUB(p + 4, p + 4, b1);
        ADD(p + 4 * n, p + 4 * n, b0);
        ADD(p + 4 * (n + 1), p + 4 * (n + 1), b1);
        commit(srs, p, WL, proof->wires_poly_comms[i]);
    }
    if (link_poly) memcpy(link_poly, wire_polys, WL * 32);
    /* public input polynomial */
    u64* pi_poly = (u64*)calloc(n, 32);
    memcpy(pi_poly, pub_inputs, num_inputs * 32);
    orc_ntt(pi_poly, log_n, 1, 0);
    for (int i = 0; i < NW; ++i) tr_append_g1(&tr, proof->wires_poly_comms[i]);

    /* ---- round 2: permutation grand product ------------------------------------------------- */
    tr_challenge(&tr, ch.beta);
    tr_challenge(&tr, ch.gamma);
    /* sigma evaluations over H */
    u64* sig_evals = (u64*)malloc(NW * n * 32);
    memcpy(sig_evals, sigma_coeffs, NW * n * 32);
    for (int i = 0; i < NW; ++i) orc_ntt(sig_evals + 4 * n * i, log_n, 0, 0);
    const size_t ZL = n + 3;
    u64* z_poly = (u64*)calloc(ZL, 32);
    {
        u64* num = (u64*)malloc(n * 32);
        u64* den = (u64*)malloc(n * 32);
#pragma omp parallel for schedule(static)
        for (size_t j = 0; j < n; ++j) {
            u64 a[4], b[4];
            fr_one(a); fr_one(b);
            for (int i = 0; i < NW; ++i) {
                u64 t[4], u[4];
                ADD(t, wires + 4 * (n * i + j), ch.gamma);
                MUL(u, k + 4 * i, dom + 4 * j); MUL(u, u, ch.beta); ADD(u, u, t); MUL(a, a, u);
                MUL(u, sig_evals + 4 * (n * i + j), ch.beta); ADD(u, u, t); MUL(b, b, u);
            }
            fr_set(num + 4 * j, a);
            fr_set(den + 4 * j, b);
        }
        fr_one(z_poly);
        for (size_t j = 0; j + 1 < n; ++j) {
            u64 inv[4], t[4];
            fp_inv(&FR, inv, den + 4 * j);
            MUL(t, num + 4 * j, inv);
            MUL(z_poly + 4 * (j + 1), z_poly + 4 * j, t);
        }
        free(num); free(den);
    }
    orc_ntt(z_poly, log_n, 1, 0);
    {
        const u64 *b0 = blinders + 4 * 10, *b1 = blinders + 4 * 11, *b2 = blinders + 4 * 12;
        SUB(z_poly, z_poly, b0); SUB(z_poly + 4, z_poly + 4, b1); SUB(z_poly + 8, z_poly + 8, b2);
        ADD(z_poly + 4 * n, z_poly + 4 * n, b0);
        ADD(z_poly + 4 * (n + 1), z_poly + 4 * (n + 1), b1);
        ADD(z_poly + 4 * (n + 2), z_poly + 4 * (n + 2), b2);
    }
    commit(srs, z_poly, ZL, proof->prod_perm_poly_comm);
    tr_append_g1(&tr, proof->prod_perm_poly_comm);

    /* ---- round 3: quotient on the coset g * H_8n ---------------------------------------------- */
    tr_challenge(&tr, ch.alpha);
    u64* ce = (u64*)calloc((size_t)(NS + 2 * NW + 2) * m, 32); /* 13 selectors, 5 sigmas, 5 wires, z, pi */
    u64 *ce_sel = ce, *ce_sig = ce + 4 * m * NS, *ce_w = ce_sig + 4 * m * NW, *ce_z = ce_w + 4 * m * NW, *ce_pi = ce_z + 4 * m;
    for (int s = 0; s < NS; ++s) { memcpy(ce_sel + 4 * m * s, selector_coeffs + 4 * n * s, n * 32); orc_ntt(ce_sel + 4 * m * s, log_m, 0, 1); }
    for (int i = 0; i < NW; ++i) { memcpy(ce_sig + 4 * m * i, sigma_coeffs + 4 * n * i, n * 32); orc_ntt(ce_sig + 4 * m * i, log_m, 0, 1); }
    for (int i = 0; i < NW; ++i) { memcpy(ce_w + 4 * m * i, wire_polys + 4 * WS * i, WL * 32); orc_ntt(ce_w + 4 * m * i, log_m, 0, 1); }
    memcpy(ce_z, z_poly, ZL * 32); orc_ntt(ce_z, log_m, 0, 1);
    memcpy(ce_pi, pi_poly, n * 32); orc_ntt(ce_pi, log_m, 0, 1);
    u64* quot = (u64*)malloc(m * 32);
    {
        u64 wm[4], g[4], zh_inv[8][4], alpha2[4], nfr[4];
        orc_domain_generator(log_m, wm);
        fr_from_u64(g, 5);
        MUL(alpha2, ch.alpha, ch.alpha);
        fr_from_u64(nfr, (u64)n);
        for (int i = 0; i < 8; ++i) {
            u64 x[4], t[4];
            fr_pow_u64(x, wm, (u64)i); MUL(x, x, g);
            fr_pow_u64(t, x, (u64)n); SUB(t, t, one);
            fp_inv(&FR, zh_inv[i], t);
        }
        u64* pts = (u64*)malloc(m * 32); /* eval points g * wm^i */
        fr_set(pts, g);
        for (size_t i = 1; i < m; ++i) MUL(pts + 4 * i, pts + 4 * (i - 1), wm);
#pragma omp parallel for schedule(static)
        for (size_t i = 0; i < m; ++i) {
            fr wv[5];
            u64 t_circ[4], r1[4], r1b[4], r2[4], t[4], u[4];
            for (int j = 0; j < NW; ++j) fr_set(wv[j], ce_w + 4 * (m * j + i));
            gate_eval(ce_sel, m, i, wv, ce_pi + 4 * i, t_circ);
            const u64* zx = ce_z + 4 * i;
            const u64* zxw = ce_z + 4 * ((i + 8) % m);
            fr_set(r1, zx); fr_set(r1b, zxw);
            for (int j = 0; j < NW; ++j) {
                ADD(t, wv[j], ch.gamma);
                MUL(u, k + 4 * j, pts + 4 * i); MUL(u, u, ch.beta); ADD(u, u, t); MUL(r1, r1, u);
                MUL(u, ce_sig + 4 * (m * j + i), ch.beta); ADD(u, u, t); MUL(r1b, r1b, u);
            }
            SUB(r1, r1, r1b); MUL(r1, r1, ch.alpha);
            /* alpha^2 (z - 1) / (n (x - 1)) */
            SUB(t, pts + 4 * i, one); MUL(t, t, nfr); fp_inv(&FR, t, t);
            SUB(u, zx, one); MUL(u, u, alpha2); MUL(r2, u, t);
            ADD(t, t_circ, r1); MUL(t, t, zh_inv[i % 8]); ADD(t, t, r2);
            fr_set(quot + 4 * i, t);
        }
        free(pts);
    }
    orc_ntt(quot, log_m, 1, 1);
    {
        const size_t deg = NW * (n + 1) + 2; /* expected quotient degree */
        int bad = is_zero4(quot + 4 * deg);
        for (size_t i = deg + 1; i < m; ++i) bad |= !is_zero4(quot + 4 * i);
        if (bad) rc = 2;
    }
    const size_t QS = n + 3; /* stride of the split polynomials */
    u64* split = (u64*)calloc(NW * QS, 32);
    size_t split_len[NW];
    if (rc == 0) {
        const size_t total = NW * (n + 1) + 3; /* coefficients */
        u64 last[4];
        fr_zero(last);
        for (int i = 0; i < NW; ++i) {
            const size_t beg = (size_t)i * (n + 2), end = i < NW - 1 ? beg + n + 2 : total;
            u64* p = split + 4 * QS * i;
            memcpy(p, quot + 4 * beg, (end - beg) * 32);
            split_len[i] = end - beg;
            SUB(p, p, last);
            if (i < NW - 1) {
                const u64* now = blinders + 4 * (13 + i);
                fr_set(p + 4 * (n + 2), now);
                split_len[i] = n + 3;
                fr_set(last, now);
            }
            commit(srs, p, split_len[i], proof->split_quot_poly_comms[i]);
        }
        for (int i = 0; i < NW; ++i) tr_append_g1(&tr, proof->split_quot_poly_comms[i]);

        /* ---- round 4: evaluations --------------------------------------------------------------- */
        tr_challenge(&tr, ch.zeta);
        for (int i = 0; i < NW; ++i) poly_eval(wire_polys + 4 * WS * i, WL, ch.zeta, proof->wires_evals[i]);
        for (int i = 0; i < NW - 1; ++i) poly_eval(sigma_coeffs + 4 * n * i, n, ch.zeta, proof->wire_sigma_evals[i]);
        u64 zw[4];
        MUL(zw, ch.zeta, w);
        poly_eval(z_poly, ZL, zw, proof->perm_next_eval);
        for (int i = 0; i < NW; ++i) tr_append_fr(&tr, proof->wires_evals[i]);
        for (int i = 0; i < NW - 1; ++i) tr_append_fr(&tr, proof->wire_sigma_evals[i]);
        tr_append_fr(&tr, proof->perm_next_eval);

        /* ---- round 5: linearisation polynomial, batched openings ---------------------------------- */
        const size_t LL = n + 3;
        u64* lin = (u64*)calloc(LL, 32);
        const fr* we = (const fr*)proof->wires_evals;
        {
            u64 t[4], t2[4];
            for (int j = 0; j < 4; ++j) poly_axpy(lin, selector_coeffs + 4 * n * j, n, we[j]);
            MUL(t, we[0], we[1]); poly_axpy(lin, selector_coeffs + 4 * n * 4, n, t);
            MUL(t, we[2], we[3]); poly_axpy(lin, selector_coeffs + 4 * n * 5, n, t);
            for (int j = 0; j < 4; ++j) {
                MUL(t, we[j], we[j]); MUL(t, t, t); MUL(t, t, we[j]);
                poly_axpy(lin, selector_coeffs + 4 * n * (6 + j), n, t);
            }
            fr_zero(t2); SUB(t, t2, we[4]); poly_axpy(lin, selector_coeffs + 4 * n * 10, n, t);
            poly_axpy(lin, selector_coeffs + 4 * n * 11, n, one);
            MUL(t, we[0], we[1]); MUL(t, t, we[2]); MUL(t, t, we[3]); MUL(t, t, we[4]);
            poly_axpy(lin, selector_coeffs + 4 * n * 12, n, t);
        }
        u64 vanish[4], l1[4];
        {
            u64 t[4], nfr[4], coeff[4], u[4], alpha2[4];
            fr_pow_u64(vanish, ch.zeta, (u64)n); SUB(vanish, vanish, one);
            fr_from_u64(nfr, (u64)n); SUB(t, ch.zeta, one); MUL(t, t, nfr); fp_inv(&FR, t, t); MUL(l1, vanish, t);
            MUL(alpha2, ch.alpha, ch.alpha);
            /* coefficient of z(X) */
            fr_set(coeff, ch.alpha);
            for (int j = 0; j < NW; ++j) {
                MUL(u, k + 4 * j, ch.zeta); MUL(u, u, ch.beta); ADD(u, u, we[j]); ADD(u, u, ch.gamma);
                MUL(coeff, coeff, u);
            }
            MUL(t, alpha2, l1); ADD(coeff, coeff, t);
            poly_axpy(lin, z_poly, ZL, coeff);
            /* coefficient of sigma_4(X) */
            MUL(coeff, ch.alpha, ch.beta); MUL(coeff, coeff, proof->perm_next_eval);
            for (int j = 0; j < NW - 1; ++j) {
                MUL(u, proof->wire_sigma_evals[j], ch.beta); ADD(u, u, we[j]); ADD(u, u, ch.gamma);
                MUL(coeff, coeff, u);
            }
            fr_zero(t); SUB(coeff, t, coeff);
            poly_axpy(lin, sigma_coeffs + 4 * n * (NW - 1), n, coeff);
            /* - Z_H(zeta) * sum zeta^((n+2) i) t_i(X) */
            u64 zn2[4], c[4];
            ADD(zn2, vanish, one); MUL(zn2, zn2, ch.zeta); MUL(zn2, zn2, ch.zeta);
            fr_zero(t); SUB(c, t, vanish);
            for (int i = 0; i < NW; ++i) {
                poly_axpy(lin, split + 4 * QS * i, split_len[i], c);
                MUL(c, c, zn2);
            }
        }
        tr_challenge(&tr, ch.v);
        /* batch = lin + v w0 + ... + v^5 w4 + v^6 sigma0 + ... + v^9 sigma3 */
        u64* batch = (u64*)calloc(LL, 32);
        {
            u64 c[4];
            fr_one(c);
            poly_axpy(batch, lin, LL, c);
            for (int i = 0; i < NW; ++i) { MUL(c, c, ch.v); poly_axpy(batch, wire_polys + 4 * WS * i, WL, c); }
            for (int i = 0; i < NW - 1; ++i) { MUL(c, c, ch.v); poly_axpy(batch, sigma_coeffs + 4 * n * i, n, c); }
        }
        u64* q = (u64*)calloc(LL, 32);
        poly_div_linear(batch, LL, ch.zeta, q);
        commit(srs, q, LL - 1, proof->opening_proof);
        memset(q, 0, LL * 32);
        poly_div_linear(z_poly, ZL, zw, q);
        commit(srs, q, ZL - 1, proof->shifted_opening_proof);
        free(q); free(batch); free(lin);
        /* the verifier's combiner (appended after the openings) */
        tr_append_g1(&tr, proof->opening_proof);
        tr_append_g1(&tr, proof->shifted_opening_proof);
        tr_challenge(&tr, ch.u);
    }
    if (ch_out) *ch_out = ch;
    free(split); free(quot); free(ce); free(z_poly); free(sig_evals); free(pi_poly); free(wire_polys); free(dom);
    tr_free(&tr);
    return rc;
}

/* ------------------------------------------------------------------------------------------
 * verify (known tau)
 * ---------------------------------------------------------------------------------------- */
static void g1_load(jac* J, const u64 xy[8]) {
    if (is_zero4(xy) && is_zero4(xy + 4)) { jac_set_inf(J); return; }
    memcpy(J->x, xy, 32); memcpy(J->y, xy + 4, 32); memcpy(J->z, FQ.r1, 32);
}
static void g1_scale(jac* out, const u64 xy[8], const u64 s_mont[4]) {
    u64 s[4];
    fp_from_mont(&FR, s, s_mont);
    jac acc; jac_set_inf(&acc);
    if (is_zero4(xy) && is_zero4(xy + 4)) { *out = acc; return; }
    aff A; aff_load(&A, xy, 0);
    for (int i = 255; i >= 0; --i) {
        jac_double(&acc, &acc);
        if ((s[i / 64] >> (i % 64)) & 1) jac_add_affine(&acc, &acc, &A);
    }
    *out = acc;
}
static void g1_acc(jac* acc, const u64 xy[8], const u64 s_mont[4]) {
    jac t;
    g1_scale(&t, xy, s_mont);
    jac_add(acc, acc, &t);
}
static int jac_equal(const jac* a, const jac* b) {
    aff A, B;
    jac_to_affine(&A, a); jac_to_affine(&B, b);
    if (A.inf || B.inf) return A.inf == B.inf;
    return eq4(A.x, B.x) && eq4(A.y, B.y);
}

/* Shared verifier core.  With tau != NULL the two KZG equations are checked in G1 (known-tau SRS).
 * With out_a / out_b != NULL it instead outputs the operands of the batched pairing check
 *     e(A, [tau]_2) == e(B, [1]_2),
 *     A = W_zeta + u W_zeta_omega,
 *     B = zeta W_zeta + u zeta omega W_zeta_omega + (F - E G) + u ([z] - z_omega G),
 * which is the equation jellyfish's verifier hands to the pairing (u drawn after the openings). */
static int verify_impl(unsigned log_n, size_t num_inputs, const u64* k, const u64* selector_comms,
                       const u64* sigma_comms, const u64* pub_inputs, const orc_plonk_proof* proof,
                       const u64* tau, u64* out_a, int* a_inf, u64* out_b, int* b_inf) {
    const size_t n = (size_t)1 << log_n;
    orc_plonk_challenges ch;
    transcript tr;
    tr_init(&tr);
    tr_append_vk(&tr, log_n, num_inputs, k, selector_comms, sigma_comms, pub_inputs);
    for (int i = 0; i < NW; ++i) tr_append_g1(&tr, proof->wires_poly_comms[i]);
    tr_challenge(&tr, ch.beta);
    tr_challenge(&tr, ch.gamma);
    tr_append_g1(&tr, proof->prod_perm_poly_comm);
    tr_challenge(&tr, ch.alpha);
    for (int i = 0; i < NW; ++i) tr_append_g1(&tr, proof->split_quot_poly_comms[i]);
    tr_challenge(&tr, ch.zeta);
    for (int i = 0; i < NW; ++i) tr_append_fr(&tr, proof->wires_evals[i]);
    for (int i = 0; i < NW - 1; ++i) tr_append_fr(&tr, proof->wire_sigma_evals[i]);
    tr_append_fr(&tr, proof->perm_next_eval);
    tr_challenge(&tr, ch.v);
    tr_append_g1(&tr, proof->opening_proof);
    tr_append_g1(&tr, proof->shifted_opening_proof);
    tr_challenge(&tr, ch.u);
    tr_free(&tr);

    u64 one[4], w[4], vanish[4], l1[4], nfr[4], alpha2[4], t[4], u[4];
    fr_one(one);
    orc_domain_generator(log_n, w);
    fr_pow_u64(vanish, ch.zeta, (u64)n); SUB(vanish, vanish, one);
    fr_from_u64(nfr, (u64)n);
    SUB(t, ch.zeta, one); MUL(t, t, nfr); fp_inv(&FR, t, t); MUL(l1, vanish, t);
    MUL(alpha2, ch.alpha, ch.alpha);
    /* PI(zeta) = sum_j pi_j * w^j * Z_H(zeta) / (n (zeta - w^j)) */
    u64 pi_eval[4], wj[4];
    fr_zero(pi_eval); fr_one(wj);
    for (size_t j = 0; j < num_inputs; ++j) {
        SUB(t, ch.zeta, wj); MUL(t, t, nfr); fp_inv(&FR, t, t);
        MUL(t, t, vanish); MUL(t, t, wj); MUL(t, t, pub_inputs + 4 * j);
        ADD(pi_eval, pi_eval, t);
        MUL(wj, wj, w);
    }
    const fr* we = (const fr*)proof->wires_evals;
    const fr* se = (const fr*)proof->wire_sigma_evals;
    /* r0 = PI - alpha^2 L1 - alpha z_w prod_{i<4}(w_i + beta s_i + gamma) (w_4 + gamma) */
    u64 r0[4], prod[4];
    MUL(prod, ch.alpha, proof->perm_next_eval);
    for (int j = 0; j < NW - 1; ++j) { MUL(u, se[j], ch.beta); ADD(u, u, we[j]); ADD(u, u, ch.gamma); MUL(prod, prod, u); }
    u64 prod4[4];
    ADD(u, we[4], ch.gamma); MUL(prod4, prod, u);
    MUL(t, alpha2, l1); SUB(r0, pi_eval, t); SUB(r0, r0, prod4);

    /* D */
    jac D; jac_set_inf(&D);
    for (int j = 0; j < 4; ++j) g1_acc(&D, selector_comms + 8 * j, we[j]);
    MUL(t, we[0], we[1]); g1_acc(&D, selector_comms + 8 * 4, t);
    MUL(t, we[2], we[3]); g1_acc(&D, selector_comms + 8 * 5, t);
    for (int j = 0; j < 4; ++j) { MUL(t, we[j], we[j]); MUL(t, t, t); MUL(t, t, we[j]); g1_acc(&D, selector_comms + 8 * (6 + j), t); }
    fr_zero(u); SUB(t, u, we[4]); g1_acc(&D, selector_comms + 8 * 10, t);
    g1_acc(&D, selector_comms + 8 * 11, one);
    MUL(t, we[0], we[1]); MUL(t, t, we[2]); MUL(t, t, we[3]); MUL(t, t, we[4]); g1_acc(&D, selector_comms + 8 * 12, t);
    u64 coeff[4];
    fr_set(coeff, ch.alpha);
    for (int j = 0; j < NW; ++j) { MUL(u, k + 4 * j, ch.zeta); MUL(u, u, ch.beta); ADD(u, u, we[j]); ADD(u, u, ch.gamma); MUL(coeff, coeff, u); }
    MUL(t, alpha2, l1); ADD(coeff, coeff, t);
    g1_acc(&D, proof->prod_perm_poly_comm, coeff);
    MUL(coeff, prod, ch.beta); fr_zero(u); SUB(coeff, u, coeff);
    g1_acc(&D, sigma_comms + 8 * (NW - 1), coeff);
    u64 zn2[4], c[4];
    ADD(zn2, vanish, one); MUL(zn2, zn2, ch.zeta); MUL(zn2, zn2, ch.zeta);
    fr_zero(u); SUB(c, u, vanish);
    for (int i = 0; i < NW; ++i) { g1_acc(&D, proof->split_quot_poly_comms[i], c); MUL(c, c, zn2); }
    /* F, E */
    jac F = D;
    u64 E[4], vp[4];
    fr_zero(u); SUB(E, u, r0);
    fr_one(vp);
    for (int i = 0; i < NW; ++i) { MUL(vp, vp, ch.v); g1_acc(&F, proof->wires_poly_comms[i], vp); MUL(t, vp, we[i]); ADD(E, E, t); }
    for (int i = 0; i < NW - 1; ++i) { MUL(vp, vp, ch.v); g1_acc(&F, sigma_comms + 8 * i, vp); MUL(t, vp, se[i]); ADD(E, E, t); }
    /* check 1: (tau - zeta) W == F - E G */
    u64 G[8];
    memcpy(G, FQ.r1, 32);
    u64 two[4] = {2, 0, 0, 0};
    fp_to_mont(&FQ, G + 4, two);
    jac lhs, rhs = F, tmp;
    if (!tau) {
        /* pairing operands */
        u64 zw2[4], coef[4];
        MUL(zw2, ch.zeta, w);
        jac A, B = F;
        g1_load(&A, proof->opening_proof);
        g1_acc(&A, proof->shifted_opening_proof, ch.u);
        fr_zero(u); SUB(t, u, E); g1_acc(&B, G, t);                       /* F - E G */
        g1_acc(&B, proof->opening_proof, ch.zeta);                        /* + zeta W */
        MUL(coef, ch.u, zw2); g1_acc(&B, proof->shifted_opening_proof, coef); /* + u zeta omega W' */
        g1_acc(&B, proof->prod_perm_poly_comm, ch.u);                     /* + u [z] */
        MUL(coef, ch.u, proof->perm_next_eval); fr_zero(u); SUB(coef, u, coef);
        g1_acc(&B, G, coef);                                              /* - u z_omega G */
        aff Aa, Ba;
        jac_to_affine(&Aa, &A); jac_to_affine(&Ba, &B);
        aff_store(&Aa, out_a, a_inf); aff_store(&Ba, out_b, b_inf);
        return 1;
    }
    SUB(t, tau, ch.zeta); g1_scale(&lhs, proof->opening_proof, t);
    fr_zero(u); SUB(t, u, E); g1_scale(&tmp, G, t); jac_add(&rhs, &rhs, &tmp);
    int ok = jac_equal(&lhs, &rhs);
    /* check 2: (tau - zeta w) W' == [z] - z_w G */
    u64 zw[4];
    MUL(zw, ch.zeta, w);
    SUB(t, tau, zw); g1_scale(&lhs, proof->shifted_opening_proof, t);
    g1_load(&rhs, proof->prod_perm_poly_comm);
    fr_zero(u); SUB(t, u, proof->perm_next_eval); g1_scale(&tmp, G, t); jac_add(&rhs, &rhs, &tmp);
    ok &= jac_equal(&lhs, &rhs);
    return ok;
}

int orc_plonk_verify_known_tau(unsigned log_n, size_t num_inputs, const u64* k, const u64* selector_comms,
                               const u64* sigma_comms, const u64* pub_inputs, const orc_plonk_proof* proof,
                               const u64* tau) {
    return verify_impl(log_n, num_inputs, k, selector_comms, sigma_comms, pub_inputs, proof, tau, NULL, NULL, NULL, NULL);
}

int orc_plonk_verify_operands(unsigned log_n, size_t num_inputs, const u64* k, const u64* selector_comms,
                              const u64* sigma_comms, const u64* pub_inputs, const orc_plonk_proof* proof,
                              u64* out_a, int* a_inf, u64* out_b, int* b_inf) {
    return verify_impl(log_n, num_inputs, k, selector_comms, sigma_comms, pub_inputs, proof, NULL, out_a, a_inf, out_b, b_inf);
}

/* ------------------------------------------------------------------------------------------
 * proof linking
 * ---------------------------------------------------------------------------------------- */
static void link_challenge(const u64* comm1, const u64* comm2, const u64* comm_q, u64 eta[4]) {
    transcript tr;
    tr_init(&tr);
    tr_append_g1(&tr, comm1);
    tr_append_g1(&tr, comm2);
    tr_append_g1(&tr, comm_q);
    tr_challenge(&tr, eta);
    tr_free(&tr);
}

int orc_plonk_link(const u64* a1, size_t len1, const u64* a2, size_t len2, const u64* comm1, const u64* comm2,
                   unsigned alignment, size_t offset, size_t size, const u64* srs, orc_link_proof* proof, u64* eta_out) {
    const size_t len = len1 > len2 ? len1 : len2;
    if (size == 0 || size >= len) return 1;
    u64* diff = (u64*)calloc(len, 32);
    for (size_t i = 0; i < len; ++i) {
        u64 x[4] = {0, 0, 0, 0}, y[4] = {0, 0, 0, 0};
        if (i < len1) fr_set(x, a1 + 4 * i);
        if (i < len2) fr_set(y, a2 + 4 * i);
        SUB(diff + 4 * i, x, y);
    }
    /* roots of the vanishing polynomial of the link group */
    u64 g[4], root[4];
    orc_domain_generator(alignment, g);
    fr_pow_u64(root, g, (u64)offset);
    /* q = diff / prod (X - root_i): successive synthetic divisions (remainders dropped, like
     * DensePolynomial division) */
    u64* cur = (u64*)malloc(len * 32);
    u64* nxt = (u64*)calloc(len, 32);
    memcpy(cur, diff, len * 32);
    size_t cur_len = len;
    u64 zd_eta_roots[4];
    u64* roots = (u64*)malloc(size * 32);
    int exact = 1; /* the two polynomials must agree on every root of the group */
    for (size_t i = 0; i < size; ++i) {
        fr_set(roots + 4 * i, root);
        exact &= poly_div_linear(cur, cur_len, root, nxt);
        --cur_len;
        u64* t = cur; cur = nxt; nxt = t;
        MUL(root, root, g);
    }
    if (!exact) {
        free(roots); free(cur); free(nxt); free(diff);
        return 2; /* the wire polynomials differ on the link group: no valid link proof exists */
    }
    commit(srs, cur, cur_len, proof->quotient_commitment);
    u64 eta[4];
    link_challenge(comm1, comm2, proof->quotient_commitment, eta);
    if (eta_out) fr_set(eta_out, eta);
    /* identity = diff - Z_D(eta) * q, opened at eta */
    fr_one(zd_eta_roots);
    for (size_t i = 0; i < size; ++i) {
        u64 t[4];
        SUB(t, eta, roots + 4 * i);
        MUL(zd_eta_roots, zd_eta_roots, t);
    }
    u64 neg[4], zero[4];
    fr_zero(zero);
    SUB(neg, zero, zd_eta_roots);
    poly_axpy(diff, cur, cur_len, neg);
    memset(nxt, 0, len * 32);
    poly_div_linear(diff, len, eta, nxt);
    commit(srs, nxt, len - 1, proof->opening_proof);
    free(roots); free(cur); free(nxt); free(diff);
    return 0;
}

int orc_plonk_link_verify_known_tau(const u64* comm1, const u64* comm2, unsigned alignment, size_t offset, size_t size,
                                    const orc_link_proof* proof, const u64* tau) {
    u64 eta[4], g[4], root[4], zd[4], t[4], zero[4];
    link_challenge(comm1, comm2, proof->quotient_commitment, eta);
    orc_domain_generator(alignment, g);
    fr_pow_u64(root, g, (u64)offset);
    fr_one(zd);
    for (size_t i = 0; i < size; ++i) {
        SUB(t, eta, root);
        MUL(zd, zd, t);
        MUL(root, root, g);
    }
    fr_zero(zero);
    jac lhs, rhs, tmp;
    SUB(t, tau, eta);
    g1_scale(&lhs, proof->opening_proof, t);
    g1_load(&rhs, comm1);
    u64 one[4], minus_one[4];
    fr_one(one);
    SUB(minus_one, zero, one);
    g1_scale(&tmp, comm2, minus_one); jac_add(&rhs, &rhs, &tmp);
    SUB(t, zero, zd);
    g1_scale(&tmp, proof->quotient_commitment, t); jac_add(&rhs, &rhs, &tmp);
    return jac_equal(&lhs, &rhs);
}
