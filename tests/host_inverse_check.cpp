// Host check of the field inversions of renegade_b200/csrc/ff.cuh (compiled and run by tests/test_host_field_cpu.py): the
// division-step inverse the host side uses (fe_inv_safegcd) against the Fermat ladder and the binary Euclid, on edge values
// (0, 1, small residues, p - 1, p - 2, every power of two), 60 000 seeded random and short residues per field, and
// a * a^-1 == 1.  Prints the mismatch count and the two timings; exit status = number of mismatches.
#include <chrono>
#include <cstdio>
#include <cstdint>
#include <random>
#include "ec.cuh"
using namespace b200;
template <class C> int run(const char* name) {
    std::mt19937_64 rng(12345);
    int bad = 0;
    auto check = [&](const fe& a) {
        fe i1 = fe_inv_safegcd<C>(a), i2 = fe_inv_fermat<C>(a), i3 = fe_inv_euclid<C>(a);
        if (!fe_eq(i1, i2) || !fe_eq(i1, i3)) { ++bad; }
        if (!fe_is_zero(a) && !fe_eq(fe_mul<C>(a, i1), fe_one<C>())) ++bad;
    };
    // edge values (as Montgomery residues of small / large integers, and raw small residues)
    for (uint32_t v : {0u, 1u, 2u, 3u, 5u, 0xffffffffu}) { check(fe_from_u32<C>(v)); fe r = fe_zero(); r.l[0] = v; check(r); }
    fe pm1; for (int i = 0; i < 8; ++i) pm1.l[i] = C::mod(i); pm1.l[0] -= 1; check(pm1);
    fe pm2 = pm1; pm2.l[0] -= 1; check(pm2);
    for (int b = 0; b < 254; ++b) { fe r = fe_zero(); r.l[b >> 5] = 1u << (b & 31); check(r); fe s = r; s.l[0] |= 1; check(s); }
    for (int it = 0; it < 60000; ++it) {
        fe a;
        for (int i = 0; i < 4; ++i) { uint64_t w = rng(); a.l[2*i] = (uint32_t)w; a.l[2*i+1] = (uint32_t)(w >> 32); }
        a.l[7] &= 0x0fffffffu;  // < 2^252 < p
        if (it % 7 == 0) { for (int i = 3 + it % 5; i < 8; ++i) a.l[i] = 0; }  // short values
        check(a);
    }
    printf("%s: mismatches %d\n", name, bad);
    fe x = fe_from_u32<C>(77);
    auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < 100000; ++i) x = fe_inv_safegcd<C>(fe_add<C>(x, fe_one<C>()));
    auto t1 = std::chrono::steady_clock::now();
    for (int i = 0; i < 20000; ++i) x = fe_inv_euclid<C>(fe_add<C>(x, fe_one<C>()));
    auto t2 = std::chrono::steady_clock::now();
    printf("  safegcd %.2f us, euclid %.2f us (%u)\n", std::chrono::duration<double, std::micro>(t1 - t0).count() / 100000,
           std::chrono::duration<double, std::micro>(t2 - t1).count() / 20000, x.l[0]);
    return bad;
}
int main() { return run<FqCfg>("Fq") + run<FrCfg>("Fr"); }
