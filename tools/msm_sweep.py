"""Window-size / size sweep of the device MSM (device-resident inputs, CUDA-event phase times)."""
import json
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import renegade_b200 as rb

def main():
    sizes = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [16, 20]
    cs = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0]
    ctx = rb.Context(0)
    ctx.msm_timing(True)
    for lg in sizes:
        n = 1 << lg
        d_pts = torch.empty((n, 8), dtype=torch.int64, device="cuda")
        d_s = torch.empty((n, 4), dtype=torch.int64, device="cuda")
        torch.cuda.synchronize()
        ctx.known_dlog_bases_device(0xB200, n, d_pts.data_ptr())
        ctx.splitmix_fr_device(0x5CA1A8, n, d_s.data_ptr(), montgomery=False)
        ref = None
        for c in cs:
            bases = ctx.load_bases_device(d_pts.data_ptr(), n, window_bits=c)
            for _ in range(3):
                out = ctx.msm_device(bases, d_s.data_ptr(), n)
            ph, wall = [], []
            for _ in range(5):
                t0 = time.perf_counter()
                out = ctx.msm_device(bases, d_s.data_ptr(), n)
                wall.append((time.perf_counter() - t0) * 1e3)
                ph.append(ctx.msm_timing(True))
            if ref is None:
                ref = out
            ok = bool((out[0] == ref[0]).all())
            best = min(ph, key=lambda p: p["total"])
            print(json.dumps({"log_n": lg, "plan": bases.plan, "ms": {k: round(v, 4) for k, v in best.items()},
                              "wall_ms_call": round(min(wall), 4),  # the whole call: device phases + read-back + host epilogue (Horner over the bit sums, inversion)
                              "same_result": ok}), flush=True)
            bases.free()

if __name__ == "__main__":
    main()
