"""GPU parity: NTT / iNTT / coset variants through the C ABI vs the oracle (bit-exact), the
golden known answers, and size-independent properties at BASELINE.json's 2^20."""
import numpy as np
import pytest

import renegade_b200 as rb

pytestmark = pytest.mark.gpu


def H(s):
    return int(s, 16)


@pytest.mark.parametrize("log_n", [0, 1, 2, 3, 5, 9, 10, 11, 13, 16, 18])
def test_ntt_matches_oracle(ctx, oracle, log_n):
    n = 1 << log_n
    x = oracle.splitmix_fr(0x1177, n, montgomery=True)
    for inverse in (False, True):
        for coset in (False, True):
            got = ctx.ntt(x, inverse=inverse, coset=coset)
            exp = oracle.ntt(x, inverse=inverse, coset=coset)
            assert (got == exp).all(), (log_n, inverse, coset)


def test_ntt_golden(ctx, oracle, pyoracle, kat):
    py = pyoracle
    for rec in kat["ntt"]:
        xm = oracle.ints_to_array([py.to_mont(H(v), py.R) for v in rec["x"]])
        dom = rb.Radix2EvaluationDomain(ctx, len(rec["x"]))
        for key, fn in (("fft", dom.fft), ("ifft", dom.ifft), ("coset_fft", dom.coset_fft)):
            got = [py.from_mont(v, py.R) for v in oracle.array_to_ints(fn(xm))]
            assert got == [H(v) for v in rec[key]], (rec["log_n"], key)


def test_domain_matches_reference_shape(ctx, oracle):
    dom = rb.Radix2EvaluationDomain(ctx, (1 << 16) + 3)  # ark-poly rounds up to the next power of 2
    assert dom.size == 1 << 17 and dom.log_size_of_group == 17
    assert (dom.group_gen == oracle.domain_generator(17)).all()
    # fewer coefficients than the domain: zero-padded like ark-poly
    x = oracle.splitmix_fr(0x77, 100, montgomery=True)
    d2 = rb.Radix2EvaluationDomain(ctx, 100)
    padded = np.concatenate([x, np.zeros((28, 4), dtype=np.uint64)])
    assert (d2.fft(x) == oracle.ntt(padded)).all()


def test_ntt_2_20_full_size(ctx, oracle):
    """BASELINE.json config 3: 2^20 NTT + iNTT round trip, plus exact oracle equality
    (the C oracle does 2^20 in well under a second)."""
    n = 1 << 20
    x = oracle.splitmix_fr(0x1177, n, montgomery=True)
    y = ctx.ntt(x)
    assert (y == oracle.ntt(x)).all()
    assert (ctx.ntt(y, inverse=True) == x).all()
    yc = ctx.ntt(x, coset=True)
    assert (ctx.ntt(yc, inverse=True, coset=True) == x).all()


def test_ntt_properties_2_22(ctx, oracle, pyoracle):
    """Size-independent properties beyond the oracle-comparison sizes: NTT(delta_0) = 1,
    NTT(1) = n*delta_0, linearity, round trip."""
    py = pyoracle
    log_n = 22
    n = 1 << log_n
    one = oracle.int_to_limbs(py.to_mont(1, py.R))
    delta = np.zeros((n, 4), dtype=np.uint64)
    delta[0] = one
    assert (ctx.ntt(delta) == one).all()
    ones = np.tile(one, (n, 1))
    y = ctx.ntt(ones)
    assert oracle.limbs_to_int(y[0]) == py.to_mont(n, py.R) and not y[1:].any()
    x = oracle.splitmix_fr(0x1177, n, montgomery=True)
    fx = ctx.ntt(x)
    assert (ctx.ntt(fx, inverse=True) == x).all()
    # linearity on a sample of outputs: NTT(x + delta) = NTT(x) + 1
    xd = x.copy()
    xd[0] = oracle.fp_binop("orc_fp_add", oracle.FR, x[0], one)
    fxd = ctx.ntt(xd)
    for i in (0, 1, 12345, n - 1):
        assert (fxd[i] == oracle.fp_binop("orc_fp_add", oracle.FR, fx[i], one)).all()


def test_ntt_device_batch(ctx, oracle):
    import torch
    log_n, batch = 12, 5
    n = 1 << log_n
    x = oracle.splitmix_fr(0x2024, n * batch, montgomery=True).reshape(batch, n, 4)
    t = torch.from_numpy(x.view(np.int64).copy()).cuda()
    torch.cuda.synchronize()
    ctx.ntt_device(t.data_ptr(), log_n, inverse=False, coset=True, batch=batch, stride=n)
    got = t.cpu().numpy().view(np.uint64)
    for b in range(batch):
        assert (got[b] == oracle.ntt(x[b], coset=True)).all()
