"""world_size-2 gloo test (CPU) of the multi-GPU MSM host logic: shard ranges, the single
gather collective and the partial-sum combination.  On the CPU box the per-rank partial is
produced by the oracle (test infrastructure); on GPUs it is produced by the CUDA path
(tests/test_gpu_msm.py covers that leg)."""
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
    from renegade_b200.sharded import all_gather_partials, combine_partials, pack_partial, shard_range
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        b, e = shard_range(n, rank, world)
        bases = oracle_c.known_dlog_bases(0xB200, e - b, first=b)
        scalars = oracle_c.splitmix_fr(0x5CA1A8, e - b, False, first=b)
        xy, inf = oracle_c.msm(bases, scalars)
        rec = all_gather_partials(pack_partial(xy, inf))
        out, oinf = combine_partials(None, rec)
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
