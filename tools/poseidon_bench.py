"""Poseidon2 batch hashing throughput: 2^20 two-to-one hashes (a Merkle layer), host buffers and device buffers."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np, torch
import renegade_b200 as rb
from renegade_b200 import _lib
ctx = rb.Context(0)
batch, ln = 1 << 20, 2
d_in = torch.empty((batch, ln, 4), dtype=torch.int64, device="cuda")
d_out = torch.empty((batch, 4), dtype=torch.int64, device="cuda")
torch.cuda.synchronize()
ctx.splitmix_fr_device(0x90, batch * ln, d_in.data_ptr(), montgomery=True)
h_in = d_in.cpu().pin_memory(); h_out = torch.empty((batch, 4), dtype=torch.int64).pin_memory()
res = {}
for name, src, dst in (("device_buffers", d_in, d_out), ("host_pinned_buffers", h_in, h_out)):
    for i in range(6):
        t = time.perf_counter()
        _lib.check(ctx._lib.b200_poseidon2_hash_batch(ctx._h, C.c_void_p(src.data_ptr()), batch, ln, C.c_void_p(dst.data_ptr())))
        dt = time.perf_counter() - t
    res[name] = {"ms": round(dt * 1e3, 3), "Mhash_per_s": round(batch / dt / 1e6, 1), "G_fr_mul_per_s": round(batch * 240 / dt / 1e9, 1)}
print(json.dumps({"workload": "2^20 Poseidon2 two-to-one hashes (t=3, 1 permutation each)", **res}))
