"""GPU parity: Pippenger MSM through the C ABI vs the oracle (bit-exact after affine
normalisation), golden known answers, the reference's SRS points, edge cases, and the
known-discrete-log property at BASELINE.json's 2^20."""
import numpy as np
import pytest

import renegade_b200 as rb

pytestmark = pytest.mark.gpu


@pytest.fixture(params=[False, True], ids=["latency", "throughput"])
def mode(request, ctx):
    """Both MSM tunings (b200_msm_tuning): latency (4 buckets per reduction thread) and the prover's
    throughput mode (16).  Results must be identical."""
    ctx.msm_tuning(request.param)
    yield request.param
    ctx.msm_tuning(False)


def H(s):
    return int(s, 16)


def test_known_dlog_golden(ctx, oracle, pyoracle, kat):
    py = pyoracle
    for rec in kat["msm_known_dlog"]:
        n = rec["n"]
        bases = ctx.load_bases(oracle.known_dlog_bases(0xB200, n), check_on_curve=True)
        scalars = oracle.splitmix_fr(0x5CA1A8, n, montgomery=False)
        out, inf = rb.VariableBaseMSM.msm_bigint(ctx, bases, scalars)
        assert not inf
        assert py.decode_g1_mont(out.tobytes(), 0) == (H(rec["result"][0]), H(rec["result"][1]))


# c = 0: window chosen for throughput, 1: for latency, >= 8: fixed (b200_bases_load)
@pytest.mark.parametrize("n,c", [(1, 0), (31, 0), (31, 1), (32, 8), (1000, 0), (1000, 1), (1000, 9), (4099, 0), (4099, 1),
                                 (4099, 13), (1 << 14, 0), (1 << 14, 1), ((1 << 16) + 3, 0), ((1 << 16) + 3, 1),
                                 (3000, 18), (3000, 20), (3000, 23)])
def test_msm_matches_oracle(ctx, oracle, n, c, mode):
    pts = oracle.known_dlog_bases(0xB200, n)
    bases = ctx.load_bases(pts, window_bits=c)
    s = oracle.splitmix_fr(0x5CA1A8, n, montgomery=False)
    exp, einf = oracle.msm(pts, s)
    out, inf = ctx.msm(bases, s, montgomery=False)
    assert inf == einf and (out == exp).all()
    # Montgomery scalars (what KZG commit is handed) give the same commitment
    out2, inf2 = rb.UnivariateKzgPCS.commit(ctx, bases, oracle.array_to_mont(oracle.FR, s))
    assert inf2 == einf and (out2 == exp).all()
    # sub-range of the bases
    if n > 40:
        off, m = 17, n - 40
        exp3, _ = oracle.msm(pts[off:off + m], s[:m])
        out3, _ = ctx.msm(bases, s[:m], base_off=off)
        assert (out3 == exp3).all()


def test_msm_partial_precompute(ctx, oracle, monkeypatch):
    """More than one physical window (tables not fully precomputed) exercises the host Horner."""
    n = 2000
    pts = oracle.known_dlog_bases(0xB200, n)
    s = oracle.splitmix_fr(0x5CA1A8, n, montgomery=False)
    exp, _ = oracle.msm(pts, s)
    import os
    for phys in ("2", "5"):
        monkeypatch.setenv("B200_MSM_PHYS_WINDOWS", phys)
        bases = ctx.load_bases(pts, window_bits=10)
        assert bases.plan["physical_windows"] == int(phys)
        out, _ = ctx.msm(bases, s)
        assert (out == exp).all()


def test_msm_on_reference_srs_points(ctx, oracle, pyoracle, kat, srs_head, mode):
    py = pyoracle
    params_pts = np.frombuffer(srs_head[80:], dtype=np.uint64).reshape(-1, 8)
    bases = ctx.load_bases(params_pts, check_on_curve=True)  # srs.rs:178-179 on the device
    s16 = oracle.ints_to_array([H(v) for v in kat["msm_srs16"]["scalars"]])
    out, inf = ctx.msm(bases, s16)
    exp = kat["msm_srs16"]["result"]
    assert py.decode_g1_mont(out.tobytes(), 0) == (H(exp[0]), H(exp[1]))
    s = oracle.splitmix_fr(0xFEED, 512, montgomery=False)
    exp_full, _ = oracle.msm(params_pts, s)
    out, _ = ctx.msm(bases, s)
    assert (out == exp_full).all()


def test_on_curve_check_rejects(ctx, srs_head):
    from renegade_b200._lib import B200Error
    pts = np.frombuffer(srs_head[80:80 + 64 * 64], dtype=np.uint64).reshape(-1, 8).copy()
    pts[7, 0] ^= np.uint64(1)
    with pytest.raises(B200Error) as ei:
        ctx.load_bases(pts, check_on_curve=True)
    assert ei.value.code == -5 and "not on curve" in str(ei.value)


def test_msm_edge_cases(ctx, oracle, pyoracle, mode):
    py = pyoracle
    n = 300
    pts = oracle.known_dlog_bases(0xB200, n)
    bases = ctx.load_bases(pts)
    # all-zero scalars -> identity
    out, inf = ctx.msm(bases, np.zeros((n, 4), dtype=np.uint64))
    assert inf and not out.any()
    # empty MSM -> identity
    out, inf = ctx.msm(bases, np.zeros((0, 4), dtype=np.uint64))
    assert inf
    # scalar r-1 everywhere, scalar 1 everywhere
    for v in (py.R - 1, 1, 2, (1 << 253) + 12345):
        s = oracle.ints_to_array([v] * n)
        exp, einf = oracle.msm(pts, s)
        out, inf = ctx.msm(bases, s)
        assert inf == einf and (out == exp).all(), hex(v)
    # duplicate points: same bucket gets P twice (doubling inside the mixed add) and P, -P
    dup = np.repeat(pts[:3], 4, axis=0)
    bd = ctx.load_bases(dup)
    s = oracle.ints_to_array([5, 5, 5, 5, 7, py.R - 7, 9, 9, 1, 1, py.R - 1, py.R - 1])
    exp, einf = oracle.msm(dup, s)
    out, inf = ctx.msm(bd, s)
    assert inf == einf and (out == exp).all()
    # total cancellation -> identity
    s2 = oracle.ints_to_array([11, py.R - 11] + [0] * 10)
    out, inf = ctx.msm(bd, s2)
    assert inf
    # identity among the bases (64 zero bytes) is skipped
    pz = pts[:10].copy()
    pz[4] = 0
    bz = ctx.load_bases(pz)
    s = oracle.splitmix_fr(0x99, 10, montgomery=False)
    keep = [i for i in range(10) if i != 4]
    exp, _ = oracle.msm(pz[keep], s[keep])
    out, _ = ctx.msm(bz, s)
    assert (out == exp).all()


@pytest.mark.parametrize("c", [0, 1, 12, 16])
def test_msm_equal_partial_sums(ctx, oracle, pyoracle, mode, c):
    """One base repeated, scalars 1 + 8 i: every eighth bucket holds the same point, so the reduction trees keep meeting
    EQUAL operands (doubling) — and, with alternating signs, OPPOSITE ones (cancellation to the identity) — at every level,
    including the levels that run four lanes per addition (xyzz_add_quad: the P = 0 paths and identity operands)."""
    py = pyoracle
    n = 256
    one = oracle.known_dlog_bases(0xB200, 1)
    pts = np.repeat(one, n, axis=0)
    bases = ctx.load_bases(pts, window_bits=c)
    same = [1 + 8 * i for i in range(n)]
    alt = [(1 + 8 * i) if i % 2 == 0 else py.R - (1 + 8 * i) for i in range(n)]
    wide = [((1 + 8 * i) << 100) + (1 + 8 * (n - 1 - i)) for i in range(n)]
    for vals in (same, alt, wide):
        s = oracle.ints_to_array(vals)
        exp, einf = oracle.msm(pts, s)
        out, inf = ctx.msm(bases, s)
        assert inf == einf and (out == exp).all()


def test_msm_skewed_scalars(ctx, oracle, pyoracle, mode):
    """Digit distributions that pile points into few buckets: identical scalars (one bucket per
    window holds every point -> block-tree combine path), small scalars (only the low windows),
    and a 0/1/small mix like witness values."""
    py = pyoracle
    n = 20000
    pts = oracle.known_dlog_bases(0xB200, n)
    bases = ctx.load_bases(pts)
    same = oracle.ints_to_array([0x1234567_89abcdef_0fedcba9_87654321_1234567_89abcdef % py.R] * n)
    small = oracle.ints_to_array([(i * 2654435761) % 65521 for i in range(n)])
    mix = oracle.ints_to_array([(0, 1, 1, 2, py.R - 1, 7)[i % 6] for i in range(n)])
    for s in (same, small, mix):
        exp, einf = oracle.msm(pts, s)
        out, inf = ctx.msm(bases, s)
        assert inf == einf and (out == exp).all()


def test_msm_2_20_known_dlog_and_sharding(ctx, oracle, pyoracle, mode):
    """BASELINE.json config 2 at full size.  Bases a_i*G are generated on the device, so
    sum s_i*P_i must equal (sum a_i*s_i mod r)*G — one scalar multiplication checks 2^20
    terms; the same inputs split in two 'ranks' and recombined give the identical point."""
    import torch
    py = pyoracle
    n = 1 << 20
    d_pts = torch.empty((n, 8), dtype=torch.int64, device="cuda")
    d_s = torch.empty((n, 4), dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    ctx.known_dlog_bases_device(0xB200, n, d_pts.data_ptr())
    ctx.splitmix_fr_device(0x5CA1A8, n, d_s.data_ptr(), montgomery=False)
    # device generators agree with the oracle's on a prefix
    assert (d_pts[:64].cpu().numpy().view(np.uint64) == oracle.known_dlog_bases(0xB200, 64)).all()
    assert (d_s[:64].cpu().numpy().view(np.uint64) == oracle.splitmix_fr(0x5CA1A8, 64, False)).all()
    bases = ctx.load_bases_device(d_pts.data_ptr(), n)
    out, inf = ctx.msm_device(bases, d_s.data_ptr(), n, montgomery=False)
    a = oracle.array_to_ints(oracle.splitmix_fr(0xB200, n, False))
    s = oracle.array_to_ints(oracle.splitmix_fr(0x5CA1A8, n, False))
    k = sum(x * y for x, y in zip(a, s)) % py.R
    g = np.frombuffer(py.encode_g1_mont(py.G1_GEN), dtype=np.uint64)
    exp, _ = oracle.g1_mul(g, False, oracle.int_to_limbs(k))
    assert not inf and (out == exp).all()
    # two-way shard + combine (the multi-GPU path on one device)
    from renegade_b200.sharded import combine_partials, shard_range
    recs = np.zeros((2, 9), dtype=np.uint64)
    for r in range(2):
        b, e = shard_range(n, r, 2)
        xy, pinf = ctx.msm_device(bases, d_s.data_ptr() + 32 * b, e - b, montgomery=False, base_off=b)
        recs[r, :8], recs[r, 8] = xy, int(pinf)
    out2, inf2 = combine_partials(recs)
    assert not inf2 and (out2 == exp).all()


def test_msm_batch_via_prover_sizes(ctx, oracle, mode):
    """Batched MSM (several scalar vectors over the same bases in one pass) is exercised through
    the prover; here directly: 5 vectors of 3000 scalars, a stride apart, vs 5 oracle MSMs."""
    import ctypes as C
    import torch
    from renegade_b200 import _lib
    n, batch, stride = 3000, 5, 3008
    pts = oracle.known_dlog_bases(0xB200, n)
    bases = ctx.load_bases(pts)
    s = np.zeros((batch, stride, 4), dtype=np.uint64)
    for b in range(batch):
        s[b, :n] = oracle.splitmix_fr(0x100 + b, n, montgomery=True)
    t = torch.from_numpy(s.view(np.int64)).cuda()
    torch.cuda.synchronize()
    out = np.zeros((batch, 8), dtype=np.uint64)
    inf = (C.c_int * batch)()
    _lib.check(ctx._lib.b200_msm_batch_device(ctx._h, bases._h, 0, C.c_void_p(t.data_ptr()), n, stride, batch, 1,
                                              out.ctypes.data_as(C.c_void_p), inf))
    for b in range(batch):
        exp, _ = oracle.msm(pts, oracle.array_from_mont(oracle.FR, s[b, :n]))
        assert (out[b] == exp).all(), b


def test_parse_ptau_file_on_reference_bytes(ctx, oracle, srs_head):
    """The SRS loader end to end on the reference's own file bytes (first 512 powers of
    srs/srs00): header checks on the host, upload, on-curve assertion on the device, commit."""
    from renegade_b200._lib import B200Error
    params = rb.parse_ptau_file(ctx, srs_head, count=512)
    assert len(params.powers_of_g) == 512 and params.powers_of_g_host.shape == (512, 8)
    coeffs = oracle.splitmix_fr(0xABC, 300, montgomery=True)
    out, inf = rb.UnivariateKzgPCS.commit(ctx, params.powers_of_g, coeffs)
    exp, _ = oracle.msm(params.powers_of_g_host[:300], oracle.array_from_mont(oracle.FR, coeffs))
    assert not inf and (out == exp).all()
    # the full-size request (MAX_SRS_DEGREE + 1 powers) must be refused on a truncated file
    with pytest.raises(B200Error):
        rb.parse_ptau_file(ctx, srs_head)
    # a corrupted record is caught by the device-side on-curve check (srs.rs:179 "point not on curve")
    bad = bytearray(srs_head)
    bad[80 + 64 * 100] ^= 1
    with pytest.raises(B200Error) as ei:
        rb.parse_ptau_file(ctx, bytes(bad), count=512)
    assert ei.value.code == -5
