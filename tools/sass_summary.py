#!/usr/bin/env python
"""Per-kernel counts of the SASS mnemonics that prove the Blackwell-native paths (B200_PROFILING.md): tcgen05.mma
(UTCHMMA / UTCIMMA, .2CTA for cta_group::2), tcgen05.ld (LDTM), TMA (UTMALDG tensor copies, UBLKCP bulk copies),
tcgen05.commit (UTCBAR), mbarrier (SYNCS), legacy mma.sync (HMMA / IMMA).

    python tools/sass_summary.py > profiles/r02_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "real-time-disaster-management_b200", "libernet_b200.so")
OPS = ["UTCHMMA", "UTCHMMA.2CTA", "UTCIMMA", "UTCIMMA.2CTA", "LDTM", "UTMALDG", "UBLKCP", "UTCBAR", "SYNCS", "HMMA", "IMMA", "STG", "LDG"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    counts = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        base = op.split(".")[0]
        if base in ("UTCHMMA", "UTCIMMA"):
            counts[cur][base + (".2CTA" if ".2CTA" in op else "")] += 1
        elif base in OPS:
            counts[cur][base] += 1
    names = subprocess.run(["c++filt"] + list(counts), capture_output=True, text=True).stdout.splitlines()
    total = collections.Counter()
    print(f"# {os.path.relpath(LIB, ROOT)}: {len(counts)} kernels; columns = " + " ".join(OPS))
    for name, (mangled, c) in zip(names, counts.items()):
        total.update(c)
        if not any(c[o] for o in OPS[:8] + ["HMMA", "IMMA"]):
            continue
        short = re.sub(r"\(.*", "", name)[:150]
        print(" ".join(f"{c[o]:5d}" for o in OPS), " ", short)
    print("# totals: " + ", ".join(f"{o}={total[o]}" for o in OPS))


if __name__ == "__main__":
    sys.exit(main())
