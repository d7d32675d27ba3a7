"""SASS opcode census of libb200ldm.so per kernel: which kernels carry tcgen05 MMAs (UTCHMMA /
UTC*MMA), TMEM loads / stores (LDTM / STTM), TMA loads / stores (UTMALDG / UTMASTG), tensor-core barriers (UTCBAR),
the two-wide fp32 instructions (FFMA2 / FADD2) and MUFU.EX2.  Runs without a GPU:

    python scripts/sass_census.py > profiles/r02_sass_opcodes.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "3d-latent-diffusion-model_b200", "b200ldm", "libb200ldm.so")
OPS = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "UTMAPF", "SYNCS", "MUFU.EX2",
       "MUFU.TANH", "FFMA2", "FADD2", "FMNMX3", "HMMA", "STL", "LDL"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = per.setdefault(m.group(1), collections.Counter())
            continue
        if cur is None:
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        cur["_all"] += 1
        for key in OPS:
            if key == "UTCHMMA.2CTA":
                if op.startswith("UTCHMMA") and ".2CTA" in op:
                    cur[key] += 1
            elif op == key or op.startswith(key + "."):
                cur[key] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(per), capture_output=True, text=True).stdout.splitlines()
    tot = collections.Counter()
    print(f"# SASS opcode census of {os.path.relpath(SO, ROOT)} (cuobjdump -sass, sm_100a), {len(per)} kernels")
    print("# columns: " + " ".join(OPS) + " | instructions | kernel")
    for (name, c), dm in zip(per.items(), demangle):
        tot.update(c)
        short = re.sub(r"\(.*", "", dm).replace("b200ldm::", "")
        print(" ".join(f"{c[k]:5d}" for k in OPS) + f" | {c['_all']:6d} | {short}")
    print("# total: " + ", ".join(f"{k} {tot[k]}" for k in OPS if tot[k]))


if __name__ == "__main__":
    sys.exit(main())
