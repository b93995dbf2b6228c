"""Multi-GPU MSM: points sharded by contiguous range over the GPUs of a box, one collective, bit-identical
result (variant B: every rank finishes its own bucket reduction and contributes a single
G1 element).

The whole data path lives behind the C ABI (`b200_multi_*`, renegade_b200/csrc/multi.cu): local Pippenger on
every device, `ncclAllGather` of the 128-byte XYZZ partials over NVLink, a W-term addition on each device in
rank order.  This module is only the host-side binding: `MultiGpu.single_process(devices)` is the Rust
relayer's shape (one process drives all GPUs), `MultiGpu.from_torch_distributed()` the torchrun shape (one
process per GPU; the NCCL unique id travels over the existing process group).  NTT stays single-GPU.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .backend import Context


def shard_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous [begin, end) slice of n items owned by `rank` (sizes differ by at most 1) — the library's
    own partition (`b200_shard_range`)."""
    b, e = C.c_size_t(0), C.c_size_t(0)
    _lib.load().b200_shard_range(n, rank, world, C.byref(b), C.byref(e))
    return b.value, e.value


class MultiBases:
    def __init__(self, multi: "MultiGpu", handle):
        self._m, self._h = multi, handle

    def __len__(self) -> int:
        return self._m._lib.b200_multi_bases_len(self._h)

    def shard(self, local: int = 0) -> tuple[int, int]:
        b, e = C.c_size_t(0), C.c_size_t(0)
        _lib.check(self._m._lib.b200_multi_bases_shard(self._h, local, C.byref(b), C.byref(e)))
        return b.value, e.value

    def free(self) -> None:
        if self._h:
            self._m._lib.b200_multi_bases_free(self._m._h, self._h)
            self._h = None


class MultiGpu:
    """`b200_multi`: contexts + NCCL communicator(s) for the sharded MSM."""

    def __init__(self, handle):
        self._lib = _lib.load()
        self._h = handle
        self._borrowed = []

    @staticmethod
    def single_process(devices) -> "MultiGpu":
        lib = _lib.load()
        arr = (C.c_int * len(devices))(*devices)
        h = C.c_void_p()
        _lib.check(lib.b200_multi_init(arr, len(devices), C.byref(h)))
        return MultiGpu(h)

    @staticmethod
    def unique_id() -> bytes:
        buf = (C.c_uint8 * 128)()
        _lib.check(_lib.load().b200_nccl_unique_id(buf))
        return bytes(buf)

    @staticmethod
    def for_rank(device: int, rank: int, world: int, unique_id: bytes) -> "MultiGpu":
        lib = _lib.load()
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        h = C.c_void_p()
        _lib.check(lib.b200_multi_init_rank(device, rank, world, buf, C.byref(h)))
        return MultiGpu(h)

    @staticmethod
    def from_torch_distributed(device: int) -> "MultiGpu":
        """One process per GPU: rank 0 makes the NCCL id, the process group broadcasts its 128 bytes."""
        import torch
        import torch.distributed as dist
        rank, world = dist.get_rank(), dist.get_world_size()
        ident = [MultiGpu.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ident, src=0)
        return MultiGpu.for_rank(device, rank, world, ident[0])

    @property
    def world(self) -> int:
        return self._lib.b200_multi_world(self._h)

    @property
    def local_devices(self) -> int:
        return self._lib.b200_multi_local_devices(self._h)

    def rank(self, local: int = 0) -> int:
        return self._lib.b200_multi_rank(self._h, local)

    def ctx(self, local: int = 0) -> Context:
        """Borrowed view of local device `local`'s context (owned by the multi handle)."""
        raw = self._lib.b200_multi_ctx(self._h, local)
        if not raw:
            raise IndexError("local device index out of range")
        c = Context.__new__(Context)
        c._lib, c._h, c.device, c._borrowed = self._lib, C.c_void_p(raw), local, True
        c.close = lambda: None  # the multi handle owns it
        self._borrowed.append(c)
        return c

    def load_bases(self, points: np.ndarray, window_bits: int = 0, check_on_curve: bool = False) -> MultiBases:
        pts = np.ascontiguousarray(points, dtype=np.uint64).reshape(-1, 8)
        h = C.c_void_p()
        _lib.check(self._lib.b200_multi_bases_load(self._h, pts.ctypes.data_as(C.c_void_p), pts.shape[0], window_bits,
                                                   int(check_on_curve), C.byref(h)))
        return MultiBases(self, h)

    def known_dlog_bases(self, seed: int, n: int, window_bits: int = 0) -> MultiBases:
        h = C.c_void_p()
        _lib.check(self._lib.b200_multi_bases_known_dlog(self._h, seed, n, window_bits, C.byref(h)))
        return MultiBases(self, h)

    def msm(self, bases: MultiBases, scalars: np.ndarray, montgomery: bool = False):
        """Full host scalar vector in, (affine x||y, is_identity) out — on every rank."""
        s = np.ascontiguousarray(scalars, dtype=np.uint64).reshape(-1, 4)
        out = np.zeros(8, dtype=np.uint64)
        inf = C.c_int(0)
        _lib.check(self._lib.b200_multi_msm(self._h, bases._h, s.ctypes.data_as(C.c_void_p), s.shape[0], int(montgomery),
                                            out.ctypes.data_as(C.c_void_p), C.byref(inf)))
        return out, bool(inf.value)

    def msm_local(self, bases: MultiBases, scalar_slices, on_device: bool, montgomery: bool = False):
        """`scalar_slices[i]`: address of local device i's scalar slice — on that device (`on_device`) or in host
        memory (copied inside the call)."""
        arr = (C.c_void_p * len(scalar_slices))(*[C.c_void_p(int(p)) if p else None for p in scalar_slices])
        out = np.zeros(8, dtype=np.uint64)
        inf = C.c_int(0)
        _lib.check(self._lib.b200_multi_msm_local(self._h, bases._h, arr, int(on_device), int(montgomery),
                                                  out.ctypes.data_as(C.c_void_p), C.byref(inf)))
        return out, bool(inf.value)

    def close(self) -> None:
        if self._h:
            self._lib.b200_multi_shutdown(self._h)
            self._h = None
            for c in self._borrowed:  # their handles died with the multi handle
                c._h = None


def combine_partials(records: np.ndarray):
    """Host-side sum of affine partial points, (k, 9) uint64 records x || y || identity flag
    (`b200_g1_sum_affine`).  Kept for hosts that gather partials themselves."""
    lib = _lib.load()
    pts = np.ascontiguousarray(records[:, :8])
    flags = (C.c_int * records.shape[0])(*[int(v) for v in records[:, 8]])
    out = np.zeros(8, dtype=np.uint64)
    inf = C.c_int(0)
    _lib.check(lib.b200_g1_sum_affine(pts.ctypes.data_as(C.c_void_p), flags, records.shape[0],
                                      out.ctypes.data_as(C.c_void_p), C.byref(inf)))
    return out, bool(inf.value)
