/* TEST INFRASTRUCTURE — see bn254_oracle.h for scope, citations and parity status. */
#include "bn254_oracle.h"

#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
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

static const fp_params FR = {
    {0x43e1f593f0000001ULL, 0x2833e84879b97091ULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL},
    {0xac96341c4ffffffbULL, 0x36fc76959f60cd29ULL, 0x666ea36f7879462eULL, 0x0e0a77c19a07df2fULL},
    {0x1bb8e645ae216da7ULL, 0x53fe3ab1e35c59e3ULL, 0x8c49833d53bb8085ULL, 0x0216d0b17f4e44a5ULL},
    0xc2e1f593efffffffULL};
static const fp_params FQ = {
    {0x3c208c16d87cfd47ULL, 0x97816a916871ca8dULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL},
    {0xd35d438dc58f0d9dULL, 0x0a78eb28f5c70b3dULL, 0x666ea36f7879462cULL, 0x0e0a77c19a07df2fULL},
    {0xf32cfc5b538afa89ULL, 0xb5e71911d44501fbULL, 0x47ab1eff0a417ff6ULL, 0x06d89f71cab8351fULL},
    0x87d20782e4866389ULL};

static inline const fp_params* params(int which) { return which ? &FQ : &FR; }

static inline int ge4(const u64 a[4], const u64 b[4]) {
    for (int i = 3; i >= 0; --i) {
        if (a[i] != b[i]) return a[i] > b[i];
    }
    return 1;
}
static inline u64 sub4(u64 out[4], const u64 a[4], const u64 b[4]) {
    u64 borrow = 0;
    for (int i = 0; i < 4; ++i) {
        u128 d = (u128)a[i] - b[i] - borrow;
        out[i] = (u64)d;
        borrow = (u64)(d >> 64) & 1;
    }
    return borrow;
}
static inline u64 add4(u64 out[4], const u64 a[4], const u64 b[4]) {
    u64 carry = 0;
    for (int i = 0; i < 4; ++i) {
        u128 s = (u128)a[i] + b[i] + carry;
        out[i] = (u64)s;
        carry = (u64)(s >> 64);
    }
    return carry;
}
static inline int is_zero4(const u64 a[4]) { return (a[0] | a[1] | a[2] | a[3]) == 0; }
static inline int eq4(const u64 a[4], const u64 b[4]) {
    return a[0] == b[0] && a[1] == b[1] && a[2] == b[2] && a[3] == b[3];
}

/* CIOS Montgomery product (the textbook algorithm ark-ff's MontBackend::mul_assign implements) */
static void fp_mul(const fp_params* P, u64 out[4], const u64 a[4], const u64 b[4]) {
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
static inline void fp_add(const fp_params* P, u64 out[4], const u64 a[4], const u64 b[4]) {
    u64 t[4];
    u64 c = add4(t, a, b);
    if (c || ge4(t, P->p)) sub4(out, t, P->p);
    else memcpy(out, t, 32);
}
static inline void fp_sub(const fp_params* P, u64 out[4], const u64 a[4], const u64 b[4]) {
    u64 t[4];
    if (sub4(t, a, b)) add4(out, t, P->p);
    else memcpy(out, t, 32);
}
static inline void fp_neg(const fp_params* P, u64 out[4], const u64 a[4]) {
    if (is_zero4(a)) memset(out, 0, 32);
    else sub4(out, P->p, a);
}
static inline void fp_dbl(const fp_params* P, u64 out[4], const u64 a[4]) { fp_add(P, out, a, a); }
static inline void fp_sqr(const fp_params* P, u64 out[4], const u64 a[4]) { fp_mul(P, out, a, a); }
static void fp_pow(const fp_params* P, u64 out[4], const u64 a[4], const u64 e[4]) {
    u64 acc[4];
    memcpy(acc, P->r1, 32);
    for (int i = 255; i >= 0; --i) {
        fp_sqr(P, acc, acc);
        if ((e[i / 64] >> (i % 64)) & 1) fp_mul(P, acc, acc, a);
    }
    memcpy(out, acc, 32);
}
/* Fermat inverse a^(p-2); inverse of 0 is 0 */
static void fp_inv(const fp_params* P, u64 out[4], const u64 a[4]) {
    u64 e[4], two[4] = {2, 0, 0, 0};
    sub4(e, P->p, two);
    fp_pow(P, out, a, e);
}
static inline void fp_to_mont(const fp_params* P, u64 out[4], const u64 a[4]) { fp_mul(P, out, a, P->r2); }
static inline void fp_from_mont(const fp_params* P, u64 out[4], const u64 a[4]) {
    u64 one[4] = {1, 0, 0, 0};
    fp_mul(P, out, a, one);
}

void orc_fp_mul(int w, const u64 a[4], const u64 b[4], u64 o[4]) { fp_mul(params(w), o, a, b); }
void orc_fp_add(int w, const u64 a[4], const u64 b[4], u64 o[4]) { fp_add(params(w), o, a, b); }
void orc_fp_sub(int w, const u64 a[4], const u64 b[4], u64 o[4]) { fp_sub(params(w), o, a, b); }
void orc_fp_inv(int w, const u64 a[4], u64 o[4]) { fp_inv(params(w), o, a); }
void orc_fp_to_mont(int w, const u64 a[4], u64 o[4]) { fp_to_mont(params(w), o, a); }
void orc_fp_from_mont(int w, const u64 a[4], u64 o[4]) { fp_from_mont(params(w), o, a); }
void orc_fp_array_from_mont(int w, const u64* a, size_t n, u64* out) {
    const fp_params* P = params(w);
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; ++i) fp_from_mont(P, out + 4 * i, a + 4 * i);
}
void orc_fp_array_to_mont(int w, const u64* a, size_t n, u64* out) {
    const fp_params* P = params(w);
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; ++i) fp_to_mont(P, out + 4 * i, a + 4 * i);
}

/* ------------------------------------------------------------------------------------------
 * G1: y^2 = x^3 + 3 over Fq.  Jacobian (X,Y,Z), Z = 0 <=> infinity.  Formulas are the
 * standard EFD dbl-2009-l / madd-2007-bl / add-2007-bl (the same family ark-ec uses).
 * ---------------------------------------------------------------------------------------- */
typedef struct { u64 x[4], y[4], z[4]; } jac;
typedef struct { u64 x[4], y[4]; int inf; } aff;

static inline void jac_set_inf(jac* p) {
    memcpy(p->x, FQ.r1, 32);
    memcpy(p->y, FQ.r1, 32);
    memset(p->z, 0, 32);
}
static inline int jac_is_inf(const jac* p) { return is_zero4(p->z); }

static void jac_double(jac* r, const jac* p) {
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

static void jac_add_affine(jac* r, const jac* p, const aff* q) {
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

static void jac_add(jac* r, const jac* p, const jac* q) {
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

static void jac_to_affine(aff* r, const jac* p) {
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

static inline void aff_load(aff* a, const u64 xy[8], int inf) {
    memcpy(a->x, xy, 32); memcpy(a->y, xy + 4, 32); a->inf = inf;
}
static inline void aff_store(const aff* a, u64 xy[8], int* inf) {
    if (a->inf) memset(xy, 0, 64);
    else { memcpy(xy, a->x, 32); memcpy(xy + 4, a->y, 32); }
    if (inf) *inf = a->inf;
}

int orc_g1_on_curve(const u64 xy[8]) {
    const fp_params* F = &FQ;
    u64 lhs[4], rhs[4], three[4] = {3, 0, 0, 0}, b[4];
    fp_to_mont(F, b, three);
    fp_sqr(F, lhs, xy + 4);
    fp_sqr(F, rhs, xy);
    fp_mul(F, rhs, rhs, xy);
    fp_add(F, rhs, rhs, b);
    return eq4(lhs, rhs);
}

void orc_g1_add(const u64 a[8], int a_inf, const u64 b[8], int b_inf, u64 out[8], int* out_inf) {
    aff A, B, Rr; jac J;
    aff_load(&A, a, a_inf); aff_load(&B, b, b_inf);
    jac_set_inf(&J);
    jac_add_affine(&J, &J, &A);
    jac_add_affine(&J, &J, &B);
    jac_to_affine(&Rr, &J);
    aff_store(&Rr, out, out_inf);
}

static void jac_mul(jac* r, const aff* p, const u64 k[4]) {
    jac acc; jac_set_inf(&acc);
    for (int i = 255; i >= 0; --i) {
        jac_double(&acc, &acc);
        if ((k[i / 64] >> (i % 64)) & 1) jac_add_affine(&acc, &acc, p);
    }
    *r = acc;
}

void orc_g1_mul(const u64 p[8], int p_inf, const u64 k[4], u64 out[8], int* out_inf) {
    aff A, Rr; jac J;
    aff_load(&A, p, p_inf);
    jac_mul(&J, &A, k);
    jac_to_affine(&Rr, &J);
    aff_store(&Rr, out, out_inf);
}

/* batch Jacobian -> affine with one inversion (Montgomery's trick) */
static void jac_batch_to_affine(aff* out, const jac* in, size_t n) {
    const fp_params* F = &FQ;
    u64* pre = (u64*)malloc(n * 32);
    u64 acc[4];
    memcpy(acc, F->r1, 32);
    for (size_t i = 0; i < n; ++i) {
        memcpy(pre + 4 * i, acc, 32);
        if (!jac_is_inf(&in[i])) fp_mul(F, acc, acc, in[i].z);
    }
    u64 inv[4];
    fp_inv(F, inv, acc);
    for (size_t i = n; i-- > 0;) {
        if (jac_is_inf(&in[i])) { memset(&out[i], 0, sizeof(aff)); out[i].inf = 1; continue; }
        u64 zi[4], zi2[4], zi3[4];
        fp_mul(F, zi, inv, pre + 4 * i);
        fp_mul(F, inv, inv, in[i].z);
        fp_sqr(F, zi2, zi);
        fp_mul(F, zi3, zi2, zi);
        fp_mul(F, out[i].x, in[i].x, zi2);
        fp_mul(F, out[i].y, in[i].y, zi3);
        out[i].inf = 0;
    }
    free(pre);
}

/* ------------------------------------------------------------------------------------------
 * Synthetic inputs
 * ---------------------------------------------------------------------------------------- */
static inline u64 splitmix_at(u64 seed, u64 j) { /* j-th output (0-based) of SplitMix64(seed) */
    u64 z = seed + (j + 1) * 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
static void splitmix_fr_one(u64 seed, u64 i, u64 out_canon[4]) {
    u64 v[4];
    for (int k = 0; k < 4; ++k) v[k] = splitmix_at(seed, 4 * i + k);
    /* 2^256 / r < 6: reduce by repeated subtraction */
    while (ge4(v, FR.p)) sub4(v, v, FR.p);
    memcpy(out_canon, v, 32);
}
void orc_splitmix_fr(u64 seed, size_t first, size_t n, int montgomery, u64* out) {
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; ++i) {
        u64 v[4];
        splitmix_fr_one(seed, first + i, v);
        if (montgomery) fp_to_mont(&FR, out + 4 * i, v);
        else memcpy(out + 4 * i, v, 32);
    }
}

void orc_g1_known_dlog_bases(u64 seed, size_t first, size_t n, u64* out_xy) {
    /* table T[b] = 2^b * G (affine) so each base costs popcount(a_i) mixed adds */
    aff* T = (aff*)malloc(256 * sizeof(aff));
    jac* TJ = (jac*)malloc(256 * sizeof(jac));
    jac g; memcpy(g.x, FQ.r1, 32);
    u64 two[4] = {2, 0, 0, 0};
    fp_to_mont(&FQ, g.y, two);
    memcpy(g.z, FQ.r1, 32);
    for (int b = 0; b < 256; ++b) { TJ[b] = g; jac_double(&g, &g); }
    jac_batch_to_affine(T, TJ, 256);
    jac* acc = (jac*)malloc(n * sizeof(jac));
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; ++i) {
        u64 a[4];
        splitmix_fr_one(seed, first + i, a);
        jac J; jac_set_inf(&J);
        for (int b = 0; b < 254; ++b)
            if ((a[b / 64] >> (b % 64)) & 1) jac_add_affine(&J, &J, &T[b]);
        acc[i] = J;
    }
    aff* outa = (aff*)malloc(n * sizeof(aff));
    jac_batch_to_affine(outa, acc, n);
    for (size_t i = 0; i < n; ++i) aff_store(&outa[i], out_xy + 8 * i, NULL);
    free(T); free(TJ); free(acc); free(outa);
}

/* ------------------------------------------------------------------------------------------
 * MSM — ark-ec 0.4.2 msm_bigint restated: c = 3 (n<32) else ceil(log2 n)*69/100 + 2; unsigned
 * c-bit digits over num_bits = 254; each window independently (arkworks: rayon over windows,
 * here: OpenMP); buckets 1..2^c-1; running-sum reduction; Horner with c doublings.
 * ---------------------------------------------------------------------------------------- */
int orc_msm_window_bits(size_t n) {
    if (n < 32) return 3;
    int lg = 0;
    while (((size_t)1 << lg) < n) ++lg; /* ark_std::log2 = ceil */
    return lg * 69 / 100 + 2;
}

void orc_msm(const u64* bases, const u64* scalars, size_t n, u64 out_xy[8], int* out_inf) {
    const int c = orc_msm_window_bits(n);
    const int num_bits = 254;
    const int n_win = (num_bits + c - 1) / c;
    jac* wsum = (jac*)malloc(n_win * sizeof(jac));
#pragma omp parallel for schedule(dynamic, 1)
    for (int w = 0; w < n_win; ++w) {
        const int w_start = w * c;
        const size_t nb = ((size_t)1 << c) - 1;
        jac* buckets = (jac*)malloc(nb * sizeof(jac));
        for (size_t b = 0; b < nb; ++b) jac_set_inf(&buckets[b]);
        for (size_t i = 0; i < n; ++i) {
            const u64* s = scalars + 4 * i;
            if (is_zero4(s)) continue;
            /* digit = (s >> w_start) mod 2^c */
            int limb = w_start / 64, sh = w_start % 64;
            u64 d = s[limb] >> sh;
            if (sh + c > 64 && limb + 1 < 4) d |= s[limb + 1] << (64 - sh);
            d &= ((u64)1 << c) - 1;
            if (d == 0) continue;
            aff P; aff_load(&P, bases + 8 * i, 0);
            jac_add_affine(&buckets[d - 1], &buckets[d - 1], &P);
        }
        jac running, res;
        jac_set_inf(&running); jac_set_inf(&res);
        for (size_t b = nb; b-- > 0;) {
            jac_add(&running, &running, &buckets[b]);
            jac_add(&res, &res, &running);
        }
        wsum[w] = res;
        free(buckets);
    }
    jac total = wsum[n_win - 1];
    for (int w = n_win - 2; w >= 0; --w) {
        for (int k = 0; k < c; ++k) jac_double(&total, &total);
        jac_add(&total, &total, &wsum[w]);
    }
    aff Rr; jac_to_affine(&Rr, &total);
    aff_store(&Rr, out_xy, out_inf);
    free(wsum);
}

void orc_msm_naive(const u64* bases, const u64* scalars, size_t n, u64 out_xy[8], int* out_inf) {
    jac total; jac_set_inf(&total);
    for (size_t i = 0; i < n; ++i) {
        aff P; aff_load(&P, bases + 8 * i, 0);
        jac t; jac_mul(&t, &P, scalars + 4 * i);
        jac_add(&total, &total, &t);
    }
    aff Rr; jac_to_affine(&Rr, &total);
    aff_store(&Rr, out_xy, out_inf);
}

/* ------------------------------------------------------------------------------------------
 * NTT — ark-poly 0.4.2 Radix2EvaluationDomain restated (natural order in/out).
 * ---------------------------------------------------------------------------------------- */
/* TWO_ADIC_ROOT_OF_UNITY = 5^((r-1)/2^28), canonical value */
static const u64 ROOT_2_28_CANON[4] = {0x9bd61b6e725b19f0ULL, 0x402d111e41112ed4ULL,
                                       0x00e0a7eb8ef62abcULL, 0x2a3c09f0a58a7e85ULL};

void orc_domain_generator(unsigned log_n, u64 out[4]) {
    u64 w[4];
    fp_to_mont(&FR, w, ROOT_2_28_CANON);
    for (unsigned i = log_n; i < 28; ++i) fp_sqr(&FR, w, w);
    memcpy(out, w, 32);
}

static size_t bitrev(size_t x, unsigned bits) {
    size_t r = 0;
    for (unsigned i = 0; i < bits; ++i) { r = (r << 1) | (x & 1); x >>= 1; }
    return r;
}

static void distribute_powers(u64* data, size_t n, const u64 g[4]) {
    /* data[i] *= g^i; chunked so it parallelises */
    const size_t chunk = 4096;
#pragma omp parallel for schedule(static)
    for (size_t c0 = 0; c0 < n; c0 += chunk) {
        u64 e[4] = {c0, 0, 0, 0}, t[4];
        fp_pow(&FR, t, g, e);
        size_t end = c0 + chunk < n ? c0 + chunk : n;
        for (size_t i = c0; i < end; ++i) {
            fp_mul(&FR, data + 4 * i, data + 4 * i, t);
            fp_mul(&FR, t, t, g);
        }
    }
}

void orc_ntt(u64* data, unsigned log_n, int inverse, int coset) {
    const size_t n = (size_t)1 << log_n;
    u64 g[4], five[4] = {5, 0, 0, 0};
    fp_to_mont(&FR, g, five);
    if (coset && !inverse) distribute_powers(data, n, g);

    u64 w[4];
    orc_domain_generator(log_n, w);
    if (inverse) fp_inv(&FR, w, w);
    /* twiddle table w^k, k < n/2 */
    size_t half = n / 2 ? n / 2 : 1;
    u64* tw = (u64*)malloc(half * 32);
    memcpy(tw, FR.r1, 32);
    for (size_t k = 1; k < half; ++k) fp_mul(&FR, tw + 4 * k, tw + 4 * (k - 1), w);

    for (size_t i = 0; i < n; ++i) {
        size_t j = bitrev(i, log_n);
        if (i < j) {
            u64 t[4];
            memcpy(t, data + 4 * i, 32);
            memcpy(data + 4 * i, data + 4 * j, 32);
            memcpy(data + 4 * j, t, 32);
        }
    }
    for (size_t m = 1; m < n; m <<= 1) {
        const size_t stride = n / (2 * m);
#pragma omp parallel for schedule(static)
        for (size_t b = 0; b < n / 2; ++b) {
            size_t k = (b / m) * 2 * m, j = b % m;
            u64* u = data + 4 * (k + j);
            u64* v = data + 4 * (k + j + m);
            u64 t[4], s[4];
            fp_mul(&FR, t, v, tw + 4 * (j * stride));
            fp_add(&FR, s, u, t);
            fp_sub(&FR, v, u, t);
            memcpy(u, s, 32);
        }
    }
    free(tw);
    if (inverse) {
        u64 nn[4] = {n, 0, 0, 0}, ninv[4];
        fp_to_mont(&FR, nn, nn);
        fp_inv(&FR, ninv, nn);
#pragma omp parallel for schedule(static)
        for (size_t i = 0; i < n; ++i) fp_mul(&FR, data + 4 * i, data + 4 * i, ninv);
        if (coset) {
            u64 gi[4];
            fp_inv(&FR, gi, g);
            distribute_powers(data, n, gi);
        }
    }
}

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
