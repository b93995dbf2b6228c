"""world_size-2 gloo test (CPU) of the multi-GPU MSM's host-visible logic: the library's shard partition
(`b200_shard_range`), one gather of the per-rank partials and their combination (`b200_g1_sum_affine`).
On the CPU box the per-rank partial is produced by the oracle (test infrastructure) and the gather runs over
gloo; on GPUs the partials come from the CUDA path and travel through `ncclAllGather` inside the library
(`b200_multi_msm`, tests/test_gpu_multi.py)."""
import os
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch.distributed as dist
    import oracle_c
    from renegade_b200.sharded import combine_partials, shard_range
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        b, e = shard_range(n, rank, world)
        bases = oracle_c.known_dlog_bases(0xB200, e - b, first=b)
        scalars = oracle_c.splitmix_fr(0x5CA1A8, e - b, False, first=b)
        xy, inf = oracle_c.msm(bases, scalars)
        mine = np.zeros(9, dtype=np.uint64)
        mine[:8] = xy
        mine[8] = 1 if inf else 0
        gathered = [None] * world
        dist.all_gather_object(gathered, mine.tolist())
        out, oinf = combine_partials(np.array(gathered, dtype=np.uint64))
        q.put((rank, out.tolist(), oinf))
    finally:
        dist.destroy_process_group()


def test_sharded_msm_two_ranks(oracle):
    n, world, port = 301, 2, 29731
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    full, finf = oracle.msm(oracle.known_dlog_bases(0xB200, n), oracle.splitmix_fr(0x5CA1A8, n, False))
    for rank, out, oinf in results:
        assert not oinf and (np.array(out, dtype=np.uint64) == full).all(), rank


def test_shard_range_is_a_partition():
    from renegade_b200.sharded import shard_range
    for n in (0, 1, 7, 8, 301, 1 << 20, (1 << 24) + 3):
        for world in (1, 2, 3, 4, 8):
            cuts = [shard_range(n, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in cuts]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
