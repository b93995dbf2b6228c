"""Device-resident NTT timing (CUDA events inside the library): forward + inverse, sizes given."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import renegade_b200 as rb

sizes = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [16, 19, 20]
batches = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1]
ctx = rb.Context(0)
for lg in sizes:
    for batch in batches:
        n = 1 << lg
        d = torch.empty((batch, n, 4), dtype=torch.int64, device="cuda")
        torch.cuda.synchronize()
        for b in range(batch):
            ctx.splitmix_fr_device(0x1177 + b, n, d[b].data_ptr(), montgomery=True)
        res = {}
        for name, inv, cos in (("fwd", False, False), ("inv", True, False), ("coset_fwd", False, True), ("coset_inv", True, True)):
            ts = []
            for i in range(8):
                ctx.ntt_device(d.data_ptr(), lg, inverse=inv, coset=cos, batch=batch, stride=n)
                if i >= 3:
                    ts.append(ctx.ntt_last_ms())
            res[name] = round(min(ts) * 1e3, 1)
        gbs = batch * n * 64 / (res["fwd"] * 1e-6) / 1e9
        print(json.dumps({"log_n": lg, "batch": batch, "us": res, "fwd_GBps_algorithmic": round(gbs, 1),
                          "fwd_Gmul_per_s": round(batch * n * lg / 2 / (res["fwd"] * 1e-6) / 1e9, 1)}), flush=True)
