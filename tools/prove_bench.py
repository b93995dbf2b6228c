"""Phase timings of the device prover on a synthetic circuit (default n = 2^16)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import renegade_b200 as rb
from renegade_b200 import synth
from renegade_b200.backend import PlonkKzgSnark, plonk_last_timings, prove_raw

log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
wbits = int(sys.argv[3]) if len(sys.argv) > 3 else 0
n = 1 << log_n
ctx = rb.Context(0)
t = time.time(); circ = synth.synth_circuit(log_n, num_inputs=17, seed=0xB200); t_synth = time.time() - t
d_srs = torch.empty((n + 3, 8), dtype=torch.int64, device="cuda"); torch.cuda.synchronize()
ctx.known_dlog_bases_device(0x7A0, n + 3, d_srs.data_ptr())     # any valid G1 points serve as a timing SRS
t = time.time(); bases = ctx.load_bases_device(d_srs.data_ptr(), n + 3, window_bits=wbits); t_srs = time.time() - t
t = time.time(); pk = PlonkKzgSnark.preprocess(ctx, bases, log_n, circ.num_inputs, circ.selectors, circ.perm, circ.k); t_pre = time.time() - t
bl = synth.splitmix_blinders(1)
hw = torch.from_numpy(circ.wires.view(np.int64)).pin_memory()
dw = hw.cuda(); torch.cuda.synchronize()
for name, ptr in (("host_pinned", hw.data_ptr()), ("device", dw.data_ptr())):
    for _ in range(3): prove_raw(ctx, pk, ptr, circ.pub_inputs, bl)
    ph = []; t = time.time()
    for _ in range(reps):
        prove_raw(ctx, pk, ptr, circ.pub_inputs, bl); ph.append(plonk_last_timings(ctx))
    dt = (time.time() - t) / reps
    avg = {k: round(sum(p[k] for p in ph) / reps, 3) for k in ph[0]}
    print(json.dumps({"log_n": log_n, "wires": name, "ms_per_proof": round(dt * 1e3, 3), "proofs_per_sec": round(1 / dt, 2),
                      "phases_ms": avg, "msm_plan": bases.plan, "setup_s": {"synth": round(t_synth, 2), "srs_tables": round(t_srs, 3), "preprocess": round(t_pre, 3)}}))
