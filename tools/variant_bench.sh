#!/bin/bash
# Runs bench.py against alternative builds of the library (renegade_b200/csrc/build/variants/*.so) and prints
# one line per build: proofs/s (device-resident, e2e), single-proof latency, 2^20 MSM ms and its phases.
cd "$(dirname "$0")/.."
for lib in "" renegade_b200/csrc/build/variants/*.so; do
  name=${lib:-default}
  B200_LIB_PATH=${lib:+$PWD/$lib} python bench.py --steps ${STEPS:-150} --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
m=d['msm']
print('$name', 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'lat_ms', round(d['latency_ms_one_proof_in_flight'],3), 'msm20_ms', round(m['ms_per_step'],3), m['device_phases_ms'])"
done
