"""Generates the committed golden fixtures (run once, in the build container where
/root/reference exists; the GPU box only reads the outputs).

  python tests/golden/make_golden.py

Sources of truth:
  * srs_head.bin      — bytes copied verbatim from the reference's SRS file
                        /root/reference/srs/srs00 (section 2, first 512 G1 records) plus the
                        80-byte header, parsed the way srs.rs:63-141 does.  Reference-owned data.
  * kat.json          — known answers computed with the independent big-int restatement
                        oracle/bn254_py.py (Python ints, affine formulas, O(n^2) DFT):
                        field constants, 2G/3G, known-dlog MSMs, edge cases,
                        size-8/16 NTTs.  Values are canonical integers as hex strings.
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import bn254_py as o  # noqa: E402

N_SRS = 512


def hx(v):
    return hex(v)


def pt(P):
    return None if P is None else [hx(P[0]), hx(P[1])]


def main():
    data = open("/root/reference/srs/srs00", "rb").read(80 + 64 * N_SRS + 64)
    raw, n = o.parse_ptau_g1(data + bytes(64 * (o.MAX_SRS_DEGREE + 1)), N_SRS)  # header check only
    with open(os.path.join(HERE, "srs_head.bin"), "wb") as f:
        f.write(data[: 80 + 64 * N_SRS])
    srs_pts = [o.decode_g1_mont(raw, i) for i in range(N_SRS)]
    assert all(o.g1_is_on_curve(p) for p in srs_pts)

    G = o.G1_GEN
    kat = {
        "fr_modulus": hx(o.R), "fq_modulus": hx(o.Q),
        "fr_R": hx((1 << 256) % o.R), "fq_R": hx((1 << 256) % o.Q),
        "fr_R2": hx((1 << 512) % o.R), "fq_R2": hx((1 << 512) % o.Q),
        "fr_root_2_28": hx(o.FR_ROOT_2_28),
        "g1_2G": pt(o.g1_mul(G, 2)), "g1_3G": pt(o.g1_mul(G, 3)),
        "srs_points_0_3": [pt(p) for p in srs_pts[:4]],
    }
    # field products
    import random
    rnd = random.Random(0xB200)
    kat["field_mul"] = []
    for name, p in (("fr", o.R), ("fq", o.Q)):
        for _ in range(8):
            a, b = rnd.randrange(p), rnd.randrange(p)
            kat["field_mul"].append({"field": name, "a": hx(a), "b": hx(b), "ab": hx(a * b % p),
                                     "a_inv": hx(pow(a, -1, p))})
    # known-dlog MSMs: bases a_i*G (SplitMix seed 0xB200), scalars seed 0x5CA1A8
    kat["msm_known_dlog"] = []
    for n in (1, 2, 7, 33, 100):
        a = o.splitmix_fr(0xB200, n)
        s = o.splitmix_fr(0x5CA1A8, n)
        expect = o.g1_mul(G, sum(x * y for x, y in zip(a, s)) % o.R)
        kat["msm_known_dlog"].append({"n": n, "result": pt(expect)})
    # MSM on real SRS points with small scalars, by the naive definition
    s = o.splitmix_fr(0xFEED, 16)
    kat["msm_srs16"] = {"scalars": [hx(v) for v in s], "result": pt(o.msm_naive(srs_pts[:16], s))}
    # edge cases
    P = srs_pts[5]
    kat["edge"] = {
        "P": pt(P),
        "P_plus_P": pt(o.g1_add(P, P)),
        "r_minus_1_times_P": pt(o.g1_mul(P, o.R - 1)),
        "neg_P": pt(o.g1_neg(P)),
    }
    # NTTs by the O(n^2) definition
    kat["ntt"] = []
    for log_n in (1, 3, 4):
        x = o.splitmix_fr(0x1177, 1 << log_n)
        kat["ntt"].append({"log_n": log_n, "x": [hx(v) for v in x],
                           "fft": [hx(v) for v in o.dft_naive(x)],
                           "ifft": [hx(v) for v in o.dft_naive(x, inverse=True)],
                           "coset_fft": [hx(v) for v in o.dft_naive([v * pow(5, i, o.R) % o.R for i, v in enumerate(x)])]})
    # the two G2 points the reference reads (srs.rs:147-157): H and tau*H, first two 128-byte records of
    # ptau section 3 (after the 12-byte section header)
    sec3 = 80 + 64 * ((1 << 18) - 1)  # section 2 holds 2^18 - 1 G1 records
    with open("/root/reference/srs/srs00", "rb") as f:
        f.seek(sec3)
        hdr = f.read(12)
        assert int.from_bytes(hdr[:4], "little") == 3
        g2 = f.read(256)
    with open(os.path.join(HERE, "srs_g2.bin"), "wb") as f:
        f.write(g2)
    with open(os.path.join(HERE, "kat.json"), "w") as f:
        json.dump(kat, f, indent=1)
    print("wrote srs_head.bin (%d bytes) and kat.json" % (80 + 64 * N_SRS))


if __name__ == "__main__":
    main()
