"""GPU parity of the multi-GPU layer behind the C ABI (`b200_multi_*`, csrc/multi.cu): one MSM sharded by point
range, `ncclAllGather` of the per-GPU partial sums, W-term addition on every device.

World size 1 exercises the same code path without the collective on any box; the 2-device cases need a box with at
least two GPUs (`gpurun --gpus 2`) and are skipped otherwise.  Results are compared bit for bit with the oracle and,
at 2^20 points per GPU, with the known-discrete-log closed form."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _device_count():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("world", [1, 2])
def test_multi_msm_vs_oracle(oracle, world):
    if world > _device_count():
        pytest.skip(f"needs {world} GPUs")
    from renegade_b200.sharded import MultiGpu
    m = MultiGpu.single_process(list(range(world)))
    try:
        assert m.world == world and m.local_devices == world
        for n in (1, 5, 4099):           # n < world leaves a shard empty
            pts = oracle.known_dlog_bases(0xB200, n)
            s = oracle.splitmix_fr(0x5CA1A8, n, montgomery=False)
            mb = m.load_bases(pts, check_on_curve=True)
            assert len(mb) == n
            cuts = [mb.shard(i) for i in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            out, inf = m.msm(mb, s)
            exp, einf = oracle.msm(pts, s)
            assert inf == einf and (out == exp).all(), (world, n)
            # Montgomery-form scalars (what `commit` is handed) and a zero vector (identity result)
            out, inf = m.msm(mb, oracle.splitmix_fr(0x5CA1A8, n, montgomery=True), montgomery=True)
            assert inf == einf and (out == exp).all()
            out, inf = m.msm(mb, np.zeros((n, 4), dtype=np.uint64))
            assert inf and not out.any()
            mb.free()
    finally:
        m.close()


@pytest.mark.parametrize("world", [1, 2])
def test_multi_msm_2_20_per_gpu_closed_form(oracle, pyoracle, world):
    """world * 2^20 known-discrete-log bases generated on the devices: the sharded result must equal
    (sum a_i s_i mod r) * G."""
    if world > _device_count():
        pytest.skip(f"needs {world} GPUs")
    import torch
    from renegade_b200.sharded import MultiGpu
    py = pyoracle
    per = 1 << 20
    n = world * per
    m = MultiGpu.single_process(list(range(world)))
    try:
        mb = m.known_dlog_bases(0xB200, n)
        slices, keep = [], []
        for i in range(world):
            b, e = mb.shard(i)
            assert (b, e) == (i * per, (i + 1) * per)
            with torch.cuda.device(i):
                d = torch.empty((per, 4), dtype=torch.int64, device=f"cuda:{i}")
                torch.cuda.synchronize()
                m.ctx(i).splitmix_fr_device(0x5CA1A8, per, d.data_ptr(), montgomery=False, first=b)
            keep.append(d)
            slices.append(d.data_ptr())
        out, inf = m.msm_local(mb, slices, on_device=True)
        a = oracle.array_to_ints(oracle.splitmix_fr(0xB200, n, False))
        s = oracle.array_to_ints(oracle.splitmix_fr(0x5CA1A8, n, False))
        k = sum(x * y for x, y in zip(a, s)) % py.R
        g = np.frombuffer(py.encode_g1_mont(py.G1_GEN), dtype=np.uint64)
        exp, _ = oracle.g1_mul(g, False, oracle.int_to_limbs(k))
        assert not inf and (out == exp).all()
        # host slices give the same point
        hs = [k_.cpu().numpy().view(np.uint64) for k_ in keep]
        out2, inf2 = m.msm_local(mb, [h.ctypes.data for h in hs], on_device=False)
        assert not inf2 and (out2 == exp).all()
        mb.free()
    finally:
        m.close()


def test_nccl_is_bound_at_run_time():
    import ctypes as C
    from renegade_b200 import _lib
    v = C.c_int(0)
    _lib.check(_lib.load().b200_nccl_version(C.byref(v)))
    assert v.value >= 22000
