#!/usr/bin/env python
"""bench.py — the driver's benchmark contract for the B200-native proving backend.

Workload (BASELINE.json configs[1]): one 2^20-point BN254 G1 Pippenger MSM per GPU, known-dlog
synthetic bases a_i*G (SplitMix64 seed 0xB200) and uniform 254-bit scalars (seed 0x5CA1A8),
SURVEY.md §8(d).  A "step" is one MSM over that batch.  metric = MSM achieved GB/s
= algorithmic bytes (96 B per (point, scalar) pair) / time.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

N > 1 (under torchrun, one rank per GPU): weak scaling — every rank owns 2^20 points of an
N*2^20-point MSM, runs Pippenger on its shard, and one NCCL all_gather of the 72-byte partial
sums + a k-term G1 sum finishes the job (EC addition is not an NCCL reduce op).

--impl reference times the CPU restatement of the reference's arkworks path (oracle/, the
reference itself is Rust and cannot be built here) on the host cores, same metric and config.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LOG_N = 20
N_POINTS = 1 << LOG_N
BYTES_PER_PAIR = 96  # 64 B affine base + 32 B scalar
SEED_BASES, SEED_SCALARS = 0xB200, 0x5CA1A8
KERNELS_PER_MSM = 14  # count, 3x scan, scatter, 3x scan, segfill, accumulate, combine, heavy_combine, reduce, reduce_final
METRIC = "MSM achieved GB/s (algorithmic bytes / time), 2^20-point BN254 G1 Pippenger per GPU"


def measured_peak_hbm():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
        sm, mx, reasons = [], 0.0, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


def run_reference(args, rank, world):
    """CPU arm: the oracle's arkworks-rule Pippenger (OpenMP over windows) on the host cores."""
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_c
    oracle_c.build()
    cores = oracle_c.num_threads()
    # bounded sample: keep the whole run within a few minutes whatever the core count
    t0 = time.perf_counter()
    probe_n = 1 << 14
    pb = oracle_c.known_dlog_bases(SEED_BASES, probe_n)
    ps = oracle_c.splitmix_fr(SEED_SCALARS, probe_n, False)
    t1 = time.perf_counter()
    oracle_c.msm(pb, ps)
    probe = time.perf_counter() - t1
    total_calls = args.steps + args.warmup
    log_sample = LOG_N
    # MSM time grows ~linearly in n at fixed window: scale the probe, cap the run at ~150 s
    while log_sample > 14 and probe * (1 << (log_sample - 14)) * total_calls > 150.0:
        log_sample -= 1
    n = 1 << log_sample
    bases = oracle_c.known_dlog_bases(SEED_BASES, n)
    scalars = oracle_c.splitmix_fr(SEED_SCALARS, n, False)
    for _ in range(args.warmup):
        oracle_c.msm(bases, scalars)
    t = time.perf_counter()
    for _ in range(args.steps):
        oracle_c.msm(bases, scalars)
    dt = (time.perf_counter() - t) / max(args.steps, 1)
    gbs = n * BYTES_PER_PAIR / dt / 1e9
    sample = f"2^{log_sample}-point prefix of the 2^20 workload per step (arkworks window rule c={oracle_c.msm_window_bits(n)})"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": gbs, "unit": "GB/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u32x8 (256-bit Montgomery integers)",
        "data": "synthetic", "points_per_sec": n / dt,
        "config": {"workload": "2^20-point BN254 G1 Pippenger MSM (BASELINE.json configs[1])", "sample": sample},
        "cpu_baseline": {"value": gbs, "unit": "GB/s", "cores": cores, "kind": "port", "sample": sample,
                         "note": "restated CPU baseline (C + OpenMP, arkworks msm_bigint algorithm), not arkworks itself"},
        "e2e": {"value": gbs, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "setup_s": time.perf_counter() - t0,
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--window-bits", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    import renegade_b200 as rb
    from renegade_b200.sharded import all_gather_partials, combine_partials, pack_partial

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    ctx = rb.Context(local_rank)
    n = N_POINTS
    first = rank * n  # this rank's slice of the world*2^20-point MSM
    d_pts = torch.empty((n, 8), dtype=torch.int64, device=dev)
    d_scalars = torch.empty((n, 4), dtype=torch.int64, device=dev)
    torch.cuda.synchronize()
    t_setup = time.perf_counter()
    ctx.known_dlog_bases_device(SEED_BASES, n, d_pts.data_ptr(), first=first)
    ctx.splitmix_fr_device(SEED_SCALARS, n, d_scalars.data_ptr(), montgomery=False, first=first)
    bases = ctx.load_bases_device(d_pts.data_ptr(), n, window_bits=args.window_bits)
    setup_s = time.perf_counter() - t_setup
    plan = bases.plan
    h_scalars = torch.empty((n, 4), dtype=torch.int64).pin_memory()
    h_scalars.copy_(d_scalars)
    torch.cuda.synchronize()
    ctx.msm_timing(True)

    def step_resident():
        xy, inf = ctx.msm_device(bases, d_scalars.data_ptr(), n, montgomery=False)
        if world > 1:
            rec = all_gather_partials(pack_partial(xy, inf), dev)
            xy, inf = combine_partials(ctx, rec)
        return xy, inf

    def step_e2e():
        # public API with HOST buffers: H2D of the scalars and D2H of the result inside the call
        xy, inf = ctx.msm(bases, h_scalars.numpy().view(np.uint64), montgomery=False)
        if world > 1:
            rec = all_gather_partials(pack_partial(xy, inf), dev)
            xy, inf = combine_partials(ctx, rec)
        return xy, inf

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, collect=None):
        barrier()
        t = time.perf_counter()
        for _ in range(steps):
            r = fn()
            if collect is not None:
                collect.append(ctx.msm_timing(True))
        barrier()
        dt = time.perf_counter() - t
        if world > 1:
            tt = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dt = float(tt.item())
        return dt, r

    for _ in range(args.warmup):
        result = step_resident()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    phases = []
    dt, result = timed(step_resident, args.steps, phases)
    clocks = sampler.stop() if rank == 0 else None
    for _ in range(2):
        step_e2e()
    dt_e2e, result_e2e = timed(step_e2e, args.steps)
    assert (result[0] == result_e2e[0]).all() and result[1] == result_e2e[1]

    ms_step = dt / args.steps * 1e3
    total_pairs = n * world
    value = total_pairs * BYTES_PER_PAIR / (dt / args.steps) / 1e9
    e2e_value = total_pairs * BYTES_PER_PAIR / (dt_e2e / args.steps) / 1e9
    acc_ms = sum(p["accumulate"] for p in phases) / len(phases)
    dev_ms = sum(p["total"] for p in phases) / len(phases)
    peak, peak_src = measured_peak_hbm()
    achieved = n * BYTES_PER_PAIR / (acc_ms * 1e-3) / 1e9
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "msm_accumulate_traffic.json")) as f:
            traffic = json.load(f).get("dram_bytes_per_launch")
    except Exception:
        pass

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    out = {
        "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32x8 (256-bit Montgomery integers)", "data": "synthetic",
        "points_per_sec": total_pairs / (dt / args.steps),
        "config": {
            "workload": "2^20-point BN254 G1 Pippenger MSM per GPU (BASELINE.json configs[1]); "
                        "N GPUs = one N*2^20-point MSM sharded by point range + NCCL all_gather of partial sums",
            "points_per_gpu": n, "total_points": total_pairs, "window_bits": plan["window_bits"],
            "digits": plan["digits"], "physical_windows": plan["physical_windows"], "tables": plan["tables"],
            "l2": "inputs larger than L2: window tables %.0f MB + scalars 32 MB per step vs 126 MB L2"
                  % (plan["tables"] * n * 64 / 1e6),
            "timing": "wall clock around K synchronous steps (barrier + cuda sync both sides, max over ranks); "
                      "device_ms_per_step from CUDA events on the library's stream",
        },
        "device_ms_per_step": dev_ms,
        "phases_ms": {k: sum(p[k] for p in phases) / len(phases) for k in ("sort", "accumulate", "reduce")},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "kernel": "msm_accumulate_kernel", "peak_source": peak_src,
                     "note": "the kernel is integer-multiply-pipe bound (≈%d Fq products per point), not HBM bound; "
                             "see DESIGN.md" % (10 * plan["digits"])},
        "e2e": {"value": e2e_value, "unit": "GB/s", "h2d_bytes_per_step": n * 32, "d2h_bytes_per_step": 128 * plan["physical_windows"],
                "ms_per_step": dt_e2e / args.steps * 1e3},
        "gpu_launches": KERNELS_PER_MSM * args.steps,
        "clocks": clocks, "setup_s": setup_s,
    }

    if not args.no_cpu_baseline and world == 1:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import oracle_c  # CPU baseline leg: the only place bench.py executes the oracle
        oracle_c.build()
        log_sample = 18
        ns = 1 << log_sample
        hb = d_pts[:ns].cpu().numpy().view(np.uint64)
        hs = h_scalars[:ns].numpy().view(np.uint64)
        oracle_c.msm(hb[:4096], hs[:4096])
        t = time.perf_counter()
        cxy, cinf = oracle_c.msm(hb, hs)
        cdt = time.perf_counter() - t
        gxy, ginf = ctx.msm_device(bases, d_scalars.data_ptr(), ns, montgomery=False)
        out["cpu_baseline"] = {
            "value": ns * BYTES_PER_PAIR / cdt / 1e9, "unit": "GB/s", "cores": oracle_c.num_threads(), "kind": "port",
            "sample": "one 2^%d-point prefix of the same bases/scalars (%.2f s)" % (log_sample, cdt),
            "points_per_sec": ns / cdt, "bit_exact_vs_gpu": bool((cxy == gxy).all() and cinf == ginf),
            "note": "restated CPU baseline (C + OpenMP over windows, arkworks msm_bigint algorithm); not arkworks itself",
        }
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
