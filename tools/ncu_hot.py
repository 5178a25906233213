#!/usr/bin/env python
"""Where do the stall samples of a kernel fall?  Reads `ncu --page source --csv` (SASS view) of an .ncu-rep and prints
the sample share per address region between barriers/branches plus the top instructions.
usage: python tools/ncu_hot.py rep.ncu-rep kernel_regex"""
import csv, io, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
# first kernel only
start = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
hdr = rows[start[0]]
end = start[1] - 1 if len(start) > 1 else len(rows)
body = [r for r in rows[start[0] + 1:end] if len(r) == len(hdr)]
ci = {h: i for i, h in enumerate(hdr)}
S, E = ci["# Samples"], ci["Instructions Executed"]
tot = sum(int(r[S] or 0) for r in body)
tote = sum(int(r[E] or 0) for r in body)
print("kernel:", rows[start[0] - 1][1][:80], " samples", tot, " warp-instr", tote)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
# regions split at BAR
reg, cur, acc, acce, sacc = [], 0, 0, 0, {}
for i, r in enumerate(body):
    acc += int(r[S] or 0); acce += int(r[E] or 0)
    for h in stalls:
        sacc[h] = sacc.get(h, 0) + int(r[ci[h]] or 0)
    if "BAR.SYNC" in r[1] or i == len(body) - 1:
        reg.append((cur, i, acc, acce, sacc)); cur, acc, acce, sacc = i + 1, 0, 0, {}
for a, b, n, e, sa in reg:
    top = sorted(sa.items(), key=lambda kv: -kv[1])[:3]
    print("  sass[%5d..%5d] samples %5.1f%%  instr %5.1f%%  %s" % (a, b, 100 * n / tot, 100 * e / max(tote, 1), " ".join("%s=%d%%" % (k[6:], 100 * v / max(n, 1)) for k, v in top)))
print("top instructions:")
for r in sorted(body, key=lambda r: -int(r[S] or 0))[:14]:
    idx = body.index(r)
    print("  %5d %5.1f%%  %s" % (idx, 100 * int(r[S] or 0) / tot, r[1].strip()[:70]))
