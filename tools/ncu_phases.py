"""Split a kernel's ncu source page (SASS view) at its BAR.SYNC instructions and print, per phase, executed instructions,
stall samples, shared-memory wavefronts and the opcode mix.  Usage: ncu_phases.py report.ncu-rep kernel-regex"""
import csv
import subprocess
import sys

rep, pat = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{pat}"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
isrc, ie, iss = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
iw, iwi = hdr.index("L1 Wavefronts Shared"), hdr.index("L1 Wavefronts Shared Ideal")
data = []
for r in rows[hi + 1:]:
    if len(r) <= ie or r[0] == "Address" or not r[0].startswith("0x"):
        if r and r[0] == "Kernel Name":
            break          # next kernel instance
        continue
    data.append((r[isrc].strip(), int(r[ie] or 0), int(r[iss] or 0), int(r[iw] or 0), int(r[iwi] or 0)))
tot = sum(d[1] for d in data)
ts = sum(d[2] for d in data)
print("total warp instructions", tot, "stall samples", ts)
acc, cur = [], [0, 0, 0, 0, {}]
for s, e, sm, w, wi in data:
    toks = s.split()
    op = (toks[1] if toks[0].startswith("@") else toks[0]).split(".")[0]
    cur[0] += e; cur[1] += sm; cur[2] += w; cur[3] += wi
    cur[4][op] = cur[4].get(op, 0) + e
    if "BAR.SYNC" in s:
        acc.append(cur); cur = [0, 0, 0, 0, {}]
acc.append(cur)
for i, c in enumerate(acc):
    top = sorted(c[4].items(), key=lambda kv: -kv[1])[:12]
    print(f"phase {i}: inst {c[0]} ({100 * c[0] / max(tot, 1):.1f}%) samples {c[1]} ({100 * c[1] / max(ts, 1):.1f}%) "
          f"smem wavefronts {c[2]} (ideal {c[3]})")
    print("     ", ", ".join(f"{k} {v}" for k, v in top))
