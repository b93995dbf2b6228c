"""GPU parity: device field arithmetic (IMAD.WIDE Montgomery multiplier, carry-chain add/sub,
Fermat inverse) vs the oracle, bit-exact, through the C ABI."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_selftest_wide_vs_wordserial(ctx):
    # production multiplier vs the word-serial reference multiplier, on the device
    assert ctx.selftest_field(0xC0FFEE, 1 << 20) == 0


@pytest.mark.parametrize("field", [0, 1])
def test_field_ops_vs_oracle(ctx, oracle, pyoracle, field):
    py = pyoracle
    mod = py.R if field == 0 else py.Q
    n = 4096
    a = oracle.splitmix_fr(0xA11CE + field, n, False)  # < r < q: valid residues for both fields
    b = oracle.splitmix_fr(0xB0B + field, n, False)
    # edge operands: 0, 1, p-1
    a[0] = 0
    a[1] = oracle.int_to_limbs(1)
    a[2] = oracle.int_to_limbs(mod - 1)
    b[2] = oracle.int_to_limbs(mod - 1)
    b[3] = 0
    for op, name in ((0, "orc_fp_mul"), (1, "orc_fp_add"), (2, "orc_fp_sub")):
        got = ctx.field_op(field, op, a, b)
        exp = np.stack([oracle.fp_binop(name, field, a[i], b[i]) for i in range(n)])
        assert (got == exp).all(), name
    got = ctx.field_op(field, 3, a[:64], b[:64])
    exp = np.stack([oracle.fp_unop("orc_fp_inv", field, a[i]) for i in range(64)])
    assert (got == exp).all()
