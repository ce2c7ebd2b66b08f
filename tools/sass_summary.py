#!/usr/bin/env python
"""Per-kernel SASS opcode summary of the shipped library (what proves the tcgen05 / TMEM / TMA path):
    python tools/sass_summary.py > profiles/rNN_sass_summary.txt
Counts UTC*MMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UTMALDG / UTMASTG / UBLKCP (TMA), HMMA (legacy mma.sync),
LDGSTS (cp.async) per kernel of anomaly-detection-super-resolution_b200/libadsr_b200.so."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "anomaly-detection-super-resolution_b200", "libadsr_b200.so")
OPS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "HMMA", "LDGSTS", "SYNCS"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    counts = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.search(r"/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m:
            op = m.group(1)
            for o in OPS:
                if op.startswith(o):
                    counts[cur][o] += 1
    names = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
    print(f"# SASS opcode counts per kernel, {os.path.relpath(LIB, ROOT)} (cuobjdump -sass, sm_100a)")
    print("# " + " ".join(f"{o:>8}" for o in OPS) + "  kernel")
    for (mangled, c), name in zip(counts.items(), names):
        short = re.sub(r"^void ", "", name)
        short = re.sub(r"\(anonymous namespace\)::", "", short)
        short = re.sub(r"\((?:adsr::|const|long|int|float|void|unsigned|__nv|CU)[^()]*(\([^()]*\)[^()]*)*\)$", "", short)
        print("  " + " ".join(f"{c[o]:>8}" for o in OPS) + "  " + short)


if __name__ == "__main__":
    sys.exit(main())
