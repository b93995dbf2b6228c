"""MSM throughput sweep, FIXED sizes 2^12 ... 2^24 at 1/2/4/8 GPUs (strong scaling), one host
process driving all devices through the C ABI (b200_multi_*: local Pippenger per GPU, ncclAllGather of the partial sums,
on-device addition).  10 warm-up + 30 timed MSMs per point, wall clock around the blocking call (scalars resident on the
devices, result on the host), median and p10/p90; the CPU restatement (oracle, test infrastructure — here as the reported
baseline) is timed beside each size on the box's host cores.  Prints one JSON object per line and a final summary object.

    python tools/msm_sweep_multi.py [--sizes 12,14,...] [--gpus 1,2,4,8] [--cpu-max 24] > gpurun_out/r2_msm_sweep.jsonl
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import torch

from renegade_b200.sharded import MultiGpu

SEED_BASES, SEED_SCALARS = 0xB200, 0x5CA1A8


def pct(xs, p):
    xs = sorted(xs)
    return xs[min(len(xs) - 1, int(round(p * (len(xs) - 1))))]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="12,14,16,18,20,22,24")
    ap.add_argument("--gpus", default="1,2,4,8")
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--iters", type=int, default=30)
    ap.add_argument("--cpu-max", type=int, default=24, help="largest log2 size the CPU baseline is timed at")
    args = ap.parse_args()
    sizes = [int(x) for x in args.sizes.split(",")]
    avail = torch.cuda.device_count()
    gpus = [g for g in (int(x) for x in args.gpus.split(",")) if g <= avail]
    rows, cpu_ms, ref_point = [], {}, {}
    import oracle_c
    oracle_c.build()
    cores = oracle_c.autotune_threads()
    for G in gpus:
        m = MultiGpu.single_process(list(range(G)))
        for lg in sizes:
            n = 1 << lg
            mb = m.known_dlog_bases(SEED_BASES, n)
            keep, slices = [], []
            for i in range(G):
                b, e = mb.shard(i)
                with torch.cuda.device(i):
                    d = torch.empty((max(e - b, 1), 4), dtype=torch.int64, device=f"cuda:{i}")
                    torch.cuda.synchronize()
                    if e > b:
                        m.ctx(i).splitmix_fr_device(SEED_SCALARS, e - b, d.data_ptr(), montgomery=False, first=b)
                keep.append(d)
                slices.append(d.data_ptr() if e > b else 0)
            for _ in range(args.warmup):
                out, inf = m.msm_local(mb, slices, on_device=True)
            ts = []
            for _ in range(args.iters):
                t = time.perf_counter()
                out, inf = m.msm_local(mb, slices, on_device=True)
                ts.append((time.perf_counter() - t) * 1e3)
            if lg not in ref_point:
                ref_point[lg] = out.copy()
            same = bool((out == ref_point[lg]).all())
            # CPU baseline once per size (bases copied from device 0's generator; identical inputs)
            if lg not in cpu_ms and lg <= args.cpu_max:
                ctx0 = m.ctx(0)
                with torch.cuda.device(0):
                    dp = torch.empty((n, 8), dtype=torch.int64, device="cuda:0")
                    ds = torch.empty((n, 4), dtype=torch.int64, device="cuda:0")
                    torch.cuda.synchronize()
                    ctx0.known_dlog_bases_device(SEED_BASES, n, dp.data_ptr())
                    ctx0.splitmix_fr_device(SEED_SCALARS, n, ds.data_ptr(), montgomery=False)
                    hp, hs = dp.cpu().numpy().view(np.uint64), ds.cpu().numpy().view(np.uint64)
                    del dp, ds
                reps = 3 if lg <= 18 else 1
                best = None
                for _ in range(reps):
                    t = time.perf_counter()
                    cxy, cinf = oracle_c.msm(hp, hs)
                    dt = (time.perf_counter() - t) * 1e3
                    best = dt if best is None else min(best, dt)
                cpu_ms[lg] = {"ms": best, "bit_exact_vs_gpu": bool((cxy == out).all() and cinf == inf)}
                del hp, hs
            med = pct(ts, 0.5)
            row = {"log_n": lg, "gpus": G, "ms_median": round(med, 4), "ms_p10": round(pct(ts, 0.1), 4),
                   "ms_p90": round(pct(ts, 0.9), 4), "points_per_s": n / (med * 1e-3),
                   "algorithmic_GBps": n * 96 / (med * 1e-3) / 1e9, "same_point_as_first_config": same,
                   "plan": None, "cpu": cpu_ms.get(lg)}
            import ctypes as C
            plan = (C.c_int * 4)()
            m._lib.b200_multi_bases_plan(mb._h, 0, C.byref(plan))
            row["plan"] = {"window_bits": plan[0], "digits": plan[1], "physical_windows": plan[2], "tables": plan[3]}
            rows.append(row)
            print(json.dumps(row), flush=True)
            mb.free()
            del keep
            torch.cuda.empty_cache()
        m.close()
    base = {r["log_n"]: r["ms_median"] for r in rows if r["gpus"] == gpus[0]}
    summary = {"summary": "strong scaling of a fixed-size MSM, speed-up vs %d GPU(s)" % gpus[0], "cpu_cores": cores,
               "cpu_kind": "port (C + OpenMP restatement of arkworks msm_bigint)",
               "speedup": {str(r["log_n"]): {} for r in rows}}
    for r in rows:
        summary["speedup"][str(r["log_n"])][str(r["gpus"])] = round(base[r["log_n"]] / r["ms_median"], 3)
    summary["gpu_vs_cpu_1gpu"] = {str(lg): round(cpu_ms[lg]["ms"] / base[lg], 1) for lg in cpu_ms if lg in base}
    print(json.dumps(summary), flush=True)


if __name__ == "__main__":
    main()
