/* TEST INFRASTRUCTURE — shared static helpers of the CPU oracle (field + G1 arithmetic).
 * Included by bn254_oracle.c and plonk_oracle.c; see bn254_oracle.h for scope and citations. */
#ifndef BN254_INTERNAL_H
#define BN254_INTERNAL_H
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#if defined(__GNUC__)
#define ORC_UNUSED __attribute__((unused))
#else
#define ORC_UNUSED
#endif

typedef unsigned __int128 u128;
typedef uint64_t u64;

/* ------------------------------------------------------------------------------------------
 * Field parameters (ark-bn254 0.4.0 FrConfig / FqConfig)
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    u64 p[4];    /* modulus */
    u64 r1[4];   /* R mod p  (Montgomery one) */
    u64 r2[4];   /* R^2 mod p */
    u64 inv;     /* -p^-1 mod 2^64 */
} fp_params;

ORC_UNUSED static const fp_params FR = {
    {0x43e1f593f0000001ULL, 0x2833e84879b97091ULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL},
    {0xac96341c4ffffffbULL, 0x36fc76959f60cd29ULL, 0x666ea36f7879462eULL, 0x0e0a77c19a07df2fULL},
    {0x1bb8e645ae216da7ULL, 0x53fe3ab1e35c59e3ULL, 0x8c49833d53bb8085ULL, 0x0216d0b17f4e44a5ULL},
    0xc2e1f593efffffffULL};
ORC_UNUSED static const fp_params FQ = {
    {0x3c208c16d87cfd47ULL, 0x97816a916871ca8dULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL},
    {0xd35d438dc58f0d9dULL, 0x0a78eb28f5c70b3dULL, 0x666ea36f7879462cULL, 0x0e0a77c19a07df2fULL},
    {0xf32cfc5b538afa89ULL, 0xb5e71911d44501fbULL, 0x47ab1eff0a417ff6ULL, 0x06d89f71cab8351fULL},
    0x87d20782e4866389ULL};

ORC_UNUSED static inline const fp_params* params(int which) { return which ? &FQ : &FR; }

ORC_UNUSED static inline int ge4(const u64 a[4], const u64 b[4]) {
    for (int i = 3; i >= 0; --i) {
        if (a[i] != b[i]) return a[i] > b[i];
    }
    return 1;
}
ORC_UNUSED static inline u64 sub4(u64 out[4], const u64 a[4], const u64 b[4]) {
    u64 borrow = 0;
    for (int i = 0; i < 4; ++i) {
        u128 d = (u128)a[i] - b[i] - borrow;
        out[i] = (u64)d;
        borrow = (u64)(d >> 64) & 1;
    }
    return borrow;
}
ORC_UNUSED static inline u64 add4(u64 out[4], const u64 a[4], const u64 b[4]) {
    u64 carry = 0;
    for (int i = 0; i < 4; ++i) {
        u128 s = (u128)a[i] + b[i] + carry;
        out[i] = (u64)s;
        carry = (u64)(s >> 64);
    }
    return carry;
}
ORC_UNUSED static inline int is_zero4(const u64 a[4]) { return (a[0] | a[1] | a[2] | a[3]) == 0; }
ORC_UNUSED static inline int eq4(const u64 a[4], const u64 b[4]) {
    return a[0] == b[0] && a[1] == b[1] && a[2] == b[2] && a[3] == b[3];
}

/* CIOS Montgomery product (the textbook algorithm ark-ff's MontBackend::mul_assign implements) */
ORC_UNUSED static void fp_mul(const fp_params* P, u64 out[4], const u64 a[4], const u64 b[4]) {
    u64 t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; ++i) {
        u64 carry = 0;
        for (int j = 0; j < 4; ++j) {
            u128 s = (u128)a[j] * b[i] + t[j] + carry;
            t[j] = (u64)s;
            carry = (u64)(s >> 64);
        }
        u128 s = (u128)t[4] + carry;
        t[4] = (u64)s;
        t[5] = (u64)(s >> 64);
        u64 m = t[0] * P->inv;
        s = (u128)m * P->p[0] + t[0];
        carry = (u64)(s >> 64);
        for (int j = 1; j < 4; ++j) {
            s = (u128)m * P->p[j] + t[j] + carry;
            t[j - 1] = (u64)s;
            carry = (u64)(s >> 64);
        }
        s = (u128)t[4] + carry;
        t[3] = (u64)s;
        t[4] = t[5] + (u64)(s >> 64);
    }
    if (t[4] || ge4(t, P->p)) sub4(out, t, P->p);
    else memcpy(out, t, 32);
}
ORC_UNUSED static inline void fp_add(const fp_params* P, u64 out[4], const u64 a[4], const u64 b[4]) {
    u64 t[4];
    u64 c = add4(t, a, b);
    if (c || ge4(t, P->p)) sub4(out, t, P->p);
    else memcpy(out, t, 32);
}
ORC_UNUSED static inline void fp_sub(const fp_params* P, u64 out[4], const u64 a[4], const u64 b[4]) {
    u64 t[4];
    if (sub4(t, a, b)) add4(out, t, P->p);
    else memcpy(out, t, 32);
}
ORC_UNUSED static inline void fp_neg(const fp_params* P, u64 out[4], const u64 a[4]) {
    if (is_zero4(a)) memset(out, 0, 32);
    else sub4(out, P->p, a);
}
ORC_UNUSED static inline void fp_dbl(const fp_params* P, u64 out[4], const u64 a[4]) { fp_add(P, out, a, a); }
ORC_UNUSED static inline void fp_sqr(const fp_params* P, u64 out[4], const u64 a[4]) { fp_mul(P, out, a, a); }
ORC_UNUSED static void fp_pow(const fp_params* P, u64 out[4], const u64 a[4], const u64 e[4]) {
    u64 acc[4];
    memcpy(acc, P->r1, 32);
    for (int i = 255; i >= 0; --i) {
        fp_sqr(P, acc, acc);
        if ((e[i / 64] >> (i % 64)) & 1) fp_mul(P, acc, acc, a);
    }
    memcpy(out, acc, 32);
}
/* Inverse by the binary extended Euclidean algorithm (what ark-ff's `inverse` uses: Guajardo-
 * Kumar-Paar-Pelzl binary EEA on the Montgomery residue, then a fix-up by R^2); inverse of 0 is 0.
 * Input/output in Montgomery form: a*R  ->  a^-1 * R. */
ORC_UNUSED static inline void shr1_4(u64 a[4]) {
    a[0] = (a[0] >> 1) | (a[1] << 63);
    a[1] = (a[1] >> 1) | (a[2] << 63);
    a[2] = (a[2] >> 1) | (a[3] << 63);
    a[3] >>= 1;
}
ORC_UNUSED static void fp_inv(const fp_params* P, u64 out[4], const u64 a[4]) {
    if (is_zero4(a)) { memset(out, 0, 32); return; }
    /* invariant: b * a = u (mod p), c * a = v (mod p); a is the Montgomery residue aR, so the
     * result x satisfies x * aR = 1, i.e. x = a^-1 R^-1; multiplying by R^3 (Montgomery) gives a^-1 R */
    u64 u[4], v[4], b[4] = {1, 0, 0, 0}, c[4] = {0, 0, 0, 0};
    const u64 one[4] = {1, 0, 0, 0};
    memcpy(u, a, 32);
    memcpy(v, P->p, 32);
    while (!eq4(u, one) && !eq4(v, one)) {
        while (!(u[0] & 1)) {
            shr1_4(u);
            if (b[0] & 1) { u64 carry = add4(b, b, P->p); shr1_4(b); b[3] |= carry << 63; }
            else shr1_4(b);
        }
        while (!(v[0] & 1)) {
            shr1_4(v);
            if (c[0] & 1) { u64 carry = add4(c, c, P->p); shr1_4(c); c[3] |= carry << 63; }
            else shr1_4(c);
        }
        if (ge4(u, v)) {
            sub4(u, u, v);
            if (sub4(b, b, c)) add4(b, b, P->p);
        } else {
            sub4(v, v, u);
            if (sub4(c, c, b)) add4(c, c, P->p);
        }
    }
    const u64* x = eq4(u, one) ? b : c;
    u64 r3[4];
    fp_mul(P, r3, P->r2, P->r2); /* R^3 */
    fp_mul(P, out, x, r3);
}
ORC_UNUSED static inline void fp_to_mont(const fp_params* P, u64 out[4], const u64 a[4]) { fp_mul(P, out, a, P->r2); }
ORC_UNUSED static inline void fp_from_mont(const fp_params* P, u64 out[4], const u64 a[4]) {
    u64 one[4] = {1, 0, 0, 0};
    fp_mul(P, out, a, one);
}

/* ------------------------------------------------------------------------------------------
 * G1: y^2 = x^3 + 3 over Fq.  Jacobian (X,Y,Z), Z = 0 <=> infinity.  Formulas are the
 * standard EFD dbl-2009-l / madd-2007-bl / add-2007-bl (the same family ark-ec uses).
 * ---------------------------------------------------------------------------------------- */
typedef struct { u64 x[4], y[4], z[4]; } jac;
typedef struct { u64 x[4], y[4]; int inf; } aff;

ORC_UNUSED static inline void jac_set_inf(jac* p) {
    memcpy(p->x, FQ.r1, 32);
    memcpy(p->y, FQ.r1, 32);
    memset(p->z, 0, 32);
}
ORC_UNUSED static inline int jac_is_inf(const jac* p) { return is_zero4(p->z); }

ORC_UNUSED static void jac_double(jac* r, const jac* p) {
    if (jac_is_inf(p)) { *r = *p; return; }
    const fp_params* F = &FQ;
    u64 A[4], B[4], C[4], D[4], E[4], Fv[4], t[4], X3[4], Y3[4], Z3[4];
    fp_sqr(F, A, p->x);
    fp_sqr(F, B, p->y);
    fp_sqr(F, C, B);
    fp_add(F, t, p->x, B);
    fp_sqr(F, t, t);
    fp_sub(F, t, t, A);
    fp_sub(F, t, t, C);
    fp_dbl(F, D, t);
    fp_dbl(F, E, A);
    fp_add(F, E, E, A);
    fp_sqr(F, Fv, E);
    fp_dbl(F, t, D);
    fp_sub(F, X3, Fv, t);
    fp_sub(F, t, D, X3);
    fp_mul(F, Y3, E, t);
    fp_dbl(F, t, C);
    fp_dbl(F, t, t);
    fp_dbl(F, t, t);
    fp_sub(F, Y3, Y3, t);
    fp_mul(F, Z3, p->y, p->z);
    fp_dbl(F, Z3, Z3);
    memcpy(r->x, X3, 32); memcpy(r->y, Y3, 32); memcpy(r->z, Z3, 32);
}

ORC_UNUSED static void jac_add_affine(jac* r, const jac* p, const aff* q) {
    if (q->inf) { *r = *p; return; }
    const fp_params* F = &FQ;
    if (jac_is_inf(p)) {
        memcpy(r->x, q->x, 32); memcpy(r->y, q->y, 32); memcpy(r->z, F->r1, 32);
        return;
    }
    u64 Z1Z1[4], U2[4], S2[4], H[4], HH[4], HHH[4], rr[4], V[4], t[4], X3[4], Y3[4], Z3[4];
    fp_sqr(F, Z1Z1, p->z);
    fp_mul(F, U2, q->x, Z1Z1);
    fp_mul(F, S2, q->y, p->z);
    fp_mul(F, S2, S2, Z1Z1);
    if (eq4(U2, p->x)) {
        if (eq4(S2, p->y)) { jac_double(r, p); return; }
        jac_set_inf(r); return;
    }
    fp_sub(F, H, U2, p->x);
    fp_sqr(F, HH, H);
    fp_mul(F, HHH, H, HH);
    fp_sub(F, rr, S2, p->y);
    fp_mul(F, V, p->x, HH);
    fp_sqr(F, X3, rr);
    fp_sub(F, X3, X3, HHH);
    fp_dbl(F, t, V);
    fp_sub(F, X3, X3, t);
    fp_sub(F, t, V, X3);
    fp_mul(F, Y3, rr, t);
    fp_mul(F, t, p->y, HHH);
    fp_sub(F, Y3, Y3, t);
    fp_mul(F, Z3, p->z, H);
    memcpy(r->x, X3, 32); memcpy(r->y, Y3, 32); memcpy(r->z, Z3, 32);
}

ORC_UNUSED static void jac_add(jac* r, const jac* p, const jac* q) {
    if (jac_is_inf(p)) { *r = *q; return; }
    if (jac_is_inf(q)) { *r = *p; return; }
    const fp_params* F = &FQ;
    u64 Z1Z1[4], Z2Z2[4], U1[4], U2[4], S1[4], S2[4], H[4], HH[4], HHH[4], rr[4], V[4], t[4];
    u64 X3[4], Y3[4], Z3[4];
    fp_sqr(F, Z1Z1, p->z);
    fp_sqr(F, Z2Z2, q->z);
    fp_mul(F, U1, p->x, Z2Z2);
    fp_mul(F, U2, q->x, Z1Z1);
    fp_mul(F, S1, p->y, q->z);
    fp_mul(F, S1, S1, Z2Z2);
    fp_mul(F, S2, q->y, p->z);
    fp_mul(F, S2, S2, Z1Z1);
    if (eq4(U1, U2)) {
        if (eq4(S1, S2)) { jac_double(r, p); return; }
        jac_set_inf(r); return;
    }
    fp_sub(F, H, U2, U1);
    fp_sqr(F, HH, H);
    fp_mul(F, HHH, H, HH);
    fp_sub(F, rr, S2, S1);
    fp_mul(F, V, U1, HH);
    fp_sqr(F, X3, rr);
    fp_sub(F, X3, X3, HHH);
    fp_dbl(F, t, V);
    fp_sub(F, X3, X3, t);
    fp_sub(F, t, V, X3);
    fp_mul(F, Y3, rr, t);
    fp_mul(F, t, S1, HHH);
    fp_sub(F, Y3, Y3, t);
    fp_mul(F, Z3, p->z, q->z);
    fp_mul(F, Z3, Z3, H);
    memcpy(r->x, X3, 32); memcpy(r->y, Y3, 32); memcpy(r->z, Z3, 32);
}

ORC_UNUSED static void jac_to_affine(aff* r, const jac* p) {
    if (jac_is_inf(p)) { memset(r, 0, sizeof(*r)); r->inf = 1; return; }
    const fp_params* F = &FQ;
    u64 zi[4], zi2[4], zi3[4];
    fp_inv(F, zi, p->z);
    fp_sqr(F, zi2, zi);
    fp_mul(F, zi3, zi2, zi);
    fp_mul(F, r->x, p->x, zi2);
    fp_mul(F, r->y, p->y, zi3);
    r->inf = 0;
}

ORC_UNUSED static inline void aff_load(aff* a, const u64 xy[8], int inf) {
    memcpy(a->x, xy, 32); memcpy(a->y, xy + 4, 32); a->inf = inf;
}
ORC_UNUSED static inline void aff_store(const aff* a, u64 xy[8], int* inf) {
    if (a->inf) memset(xy, 0, 64);
    else { memcpy(xy, a->x, 32); memcpy(xy + 4, a->y, 32); }
    if (inf) *inf = a->inf;
}


#endif
