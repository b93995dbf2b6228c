import os, sys, time
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/oracle")
print("nproc", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
for p in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us", "/sys/fs/cgroup/cpu/cpu.cfs_period_us"):
    try: print(p, open(p).read().strip())
    except Exception as e: print(p, "n/a")
import oracle_c
from renegade_b200 import synth
log_n = 13
circ = synth.synth_circuit(log_n, num_inputs=17, seed=1)
srs = oracle_c.known_dlog_bases(7, (1 << log_n) + 3)
t = time.time(); pk = oracle_c.plonk_preprocess(log_n, circ.selectors, circ.perm, circ.k, srs); print("threads", oracle_c.num_threads(), "preprocess", round(time.time() - t, 2))
t = time.time(); oracle_c.plonk_prove(log_n, circ.num_inputs, circ.k, pk, circ.wires, circ.pub_inputs, synth.splitmix_blinders(1), srs); print("prove 2^13", round(time.time() - t, 2))
