"""DRAM traffic of one captured launch (`ncu --set full` report) as the JSON `bench.py` reads for `roofline.traffic`.

    python tools/ncu_traffic.py gpurun_out/<T>_prof_msm_accumulate_proof.ncu-rep "<what the launch is>" <algorithmic bytes> \
        > profiles/msm_accumulate_traffic.json
"""
import csv
import io
import json
import subprocess
import sys


def main(path, what, alg_bytes):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, row = rows[0], rows[1], rows[2]
    d, u = dict(zip(hdr, row)), dict(zip(hdr, units))
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tot = 0.0
    for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        tot += float(d[k].replace(",", "")) * scale[u[k]]
    print(json.dumps({"kernel": d["Kernel Name"].split("(")[0].split("::")[-1], "launch": what,
                      "dram_bytes_per_launch": tot, "algorithmic_bytes_per_launch": int(alg_bytes),
                      "traffic_over_algorithmic": tot / float(alg_bytes),
                      "grid": d.get("launch__grid_size"), "duration": d.get("gpu__time_duration.sum") + " " + u.get("gpu__time_duration.sum", ""),
                      "source": path}, indent=1))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3])
