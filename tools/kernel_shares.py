"""Per-proof kernel shares from an ncu launch list (`--metrics gpu__time_duration.sum --csv`) of
tools/prove_bench.py: takes the launches between the last two `k_quotient` launches (one whole proof in
cyclic order), prints launches per proof, total kernel time and each kernel's count / time / share.
Times under ncu are serialised and cold-cache: compare SHARES, not absolutes.

    python tools/kernel_shares.py gpurun_out/launches_proof_final.csv > profiles/<round>_proof_kernel_shares.txt
"""
import collections
import csv
import sys


def short(name: str) -> str:
    name = name.split("(")[0]
    for junk in ("void ", "b200::", "(anonymous namespace)::", "unnamed>::", "<unnamed>::"):
        name = name.replace(junk, "")
    return name.strip()


def main(path: str) -> None:
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 5]
    hdr = next(r for r in rows if "Kernel Name" in r)
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    recs = []
    for r in rows[rows.index(hdr) + 1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[ui].strip(), 1e-3)
        recs.append((short(r[ki]), v * scale))
    marks = [i for i, (n, _) in enumerate(recs) if n.startswith("k_quotient")]
    if len(marks) < 2:
        sys.exit("need at least two proofs in the launch list")
    window = recs[marks[-2]:marks[-1]]
    total = sum(v for _, v in window)
    agg = collections.OrderedDict()
    for n, v in window:
        c, t = agg.get(n, (0, 0.0))
        agg[n] = (c + 1, t + v)
    lg = sys.argv[2] if len(sys.argv) > 2 else "16"
    print(f"# kernel time per proof (n = 2^{lg}), ncu launch list, serialised & cold-cache: compare shares")
    print(f"launches {len(window)}  total {total:.1f} us")
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{n:42s} n={c:3d} {t:9.1f} us {100 * t / total:5.1f}%")


if __name__ == "__main__":
    main(sys.argv[1])
