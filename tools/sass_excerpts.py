"""SASS evidence for the hot kernels (no GPU needed: `cuobjdump -sass` of the built library).

    python tools/sass_excerpts.py > profiles/r2_sass_excerpts.md

Per kernel: instruction count, the mnemonic histogram (top entries), how many IMAD.WIDE.U32 / IMAD.WIDE.U32.X carry-chain
multiply-adds, 128-bit loads / stores, shared-memory and async-copy instructions it has, and the first stretch of one
Montgomery round so the IMAD.WIDE.U32.X chain can be read."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "renegade_b200", "libb200prover.so")
KERNELS = ["ntt_pass_kernelILb0", "ntt_pass_kernelILb1", "msm_accumulate_kernel", "msm_tree_kernel", "msm_blocktree_kernel",
           "msm_scan_kernel", "k_quotient", "field_op_kernelINS_5FqCfg"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    funcs = re.split(r"\n\s*Function : ", sass)[1:]
    print("# SASS excerpts (`cuobjdump -sass renegade_b200/libb200prover.so`, sm_100a, nvcc 12.9)\n")
    print("`IMAD.WIDE.U32.X` = 32x32->64 multiply-add with carry-in/out through a predicate: one per partial product of the "
          "Montgomery multiplier (ff.cuh).  No `UTMALDG` / `UBLKCP` / `LDGSTS` (TMA / async copy) and no tensor-core "
          "(`UTCMMA`, `HMMA`) instructions appear: the kernels are bound by the integer pipe, their loads are 128-bit `LDG.E.128`.\n")
    for want in KERNELS:
        for f in funcs:
            name = f.split("\n", 1)[0].strip()
            if want not in name:
                continue
            ops = re.findall(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", f)
            hist = collections.Counter(ops)
            wide = sum(v for k, v in hist.items() if k.startswith("IMAD.WIDE.U32"))
            widex = sum(v for k, v in hist.items() if k.startswith("IMAD.WIDE.U32.X"))
            print(f"## `{name[:110]}`\n")
            print(f"* instructions: {len(ops)}; `IMAD.WIDE.U32*`: {wide} (of which `.X` carry-chained: {widex}); "
                  f"`LDG.E.128*`: {sum(v for k, v in hist.items() if k.startswith('LDG.E.128'))}; "
                  f"`STG.E.128`: {hist.get('STG.E.128', 0)}; `LDS*`: {sum(v for k, v in hist.items() if k.startswith('LDS'))}; "
                  f"`STS*`: {sum(v for k, v in hist.items() if k.startswith('STS'))}; `BAR*`: {sum(v for k, v in hist.items() if k.startswith('BAR'))}; "
                  f"async copy (`LDGSTS`/`UBLKCP`/`UTMALDG`): {sum(v for k, v in hist.items() if k.startswith(('LDGSTS', 'UBLKCP', 'UTMALDG')))}; "
                  f"local-memory spills (`STL`/`LDL`): {sum(v for k, v in hist.items() if k.startswith(('STL', 'LDL')))}")
            print("* top mnemonics: " + ", ".join(f"`{k}` {v}" for k, v in hist.most_common(10)))
            lines = [ln for ln in f.split("\n") if "IMAD.WIDE.U32.X" in ln]
            if lines:
                i0 = f.split("\n").index(lines[0])
                print("\n```")
                for ln in f.split("\n")[i0:i0 + 14]:
                    m = re.search(r"/\*[0-9a-f]{4}\*/\s+(.*?)\s*;", ln)
                    if m:
                        print("    " + m.group(1))
                print("```\n")
            break


if __name__ == "__main__":
    main()
