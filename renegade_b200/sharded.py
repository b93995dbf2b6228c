"""Multi-GPU MSM: one process per GPU, points sharded by contiguous range, partial sums gathered
with one collective and added (variant B: each rank finishes its own bucket
reduction and contributes a single G1 point).

Elliptic-curve addition is not an NCCL reduction operator, so the "reduce" named in the
north star is an all_gather of 72-byte records (x || y || identity flag) followed by a k-term
G1 sum — bit-identical to the single-GPU result because G1 addition is associative and the
result is normalised to affine.  No other data-path collective exists: NTT stays single-GPU.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from .backend import Context


def shard_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous [begin, end) slice of n items owned by `rank` (sizes differ by at most 1)."""
    base, rem = divmod(n, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def pack_partial(xy: np.ndarray, is_identity: bool) -> torch.Tensor:
    rec = np.zeros(9, dtype=np.uint64)
    rec[:8] = xy
    rec[8] = 1 if is_identity else 0
    return torch.from_numpy(rec.view(np.int64).copy())


def all_gather_partials(partial: torch.Tensor, device: torch.device | None = None) -> np.ndarray:
    """all_gather of every rank's 72-byte partial; returns a (world, 9) uint64 array."""
    world = dist.get_world_size()
    t = partial.to(device) if device is not None else partial
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    return torch.stack(out).cpu().numpy().view(np.uint64)


def combine_partials(ctx_or_none: Context | None, records: np.ndarray):
    """Sum of the gathered partial points (host-side, a handful of additions)."""
    from . import _lib
    import ctypes as C
    lib = _lib.load()
    pts = np.ascontiguousarray(records[:, :8])
    flags = (C.c_int * records.shape[0])(*[int(v) for v in records[:, 8]])
    out = np.zeros(8, dtype=np.uint64)
    inf = C.c_int(0)
    _lib.check(lib.b200_g1_sum_affine(pts.ctypes.data_as(C.c_void_p), flags, records.shape[0],
                                      out.ctypes.data_as(C.c_void_p), C.byref(inf)))
    return out, bool(inf.value)


def sharded_msm_device(ctx: Context, bases_shard, d_scalars_shard: int, n_shard: int,
                       montgomery: bool, device: torch.device):
    """Each rank: local Pippenger over its shard, then gather + sum.  Returns the full result
    on every rank."""
    xy, inf = ctx.msm_device(bases_shard, d_scalars_shard, n_shard, montgomery=montgomery)
    rec = all_gather_partials(pack_partial(xy, inf), device)
    return combine_partials(ctx, rec)
