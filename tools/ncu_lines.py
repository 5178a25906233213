#!/usr/bin/env python
"""Per-source-line instruction and stall-sample shares of one kernel in an .ncu-rep (needs -lineinfo + --import-source).
usage: python tools/ncu_lines.py rep.ncu-rep kernel_regex [top]"""
import csv, io, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + rx],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
fname, hdr, lines, seen_kernel = None, None, [], 0
for r in rows:
    if r and r[0] == "Kernel Name":
        seen_kernel += 1
        if seen_kernel > 1: break
    elif r and r[0] == "File Name": fname = r[1].split("/")[-1]
    elif r and r[0] == "Line No": hdr = r
    elif hdr and r and r[0].isdigit() and len(r) == len(hdr):
        ci = {h: i for i, h in enumerate(hdr)}
        try: lines.append((fname, int(r[0]), r[1].strip(), int(r[ci["# Samples"]] or 0), int(r[ci["Instructions Executed"]] or 0)))
        except ValueError: pass
ts, ti = sum(l[3] for l in lines) or 1, sum(l[4] for l in lines) or 1
print("samples", ts, "warp-instr", ti)
for l in sorted(lines, key=lambda l: -l[4])[:top]:
    print("%5.1f%% instr %5.1f%% samp  %s:%d  %s" % (100 * l[4] / ti, 100 * l[3] / ts, l[0], l[1], l[2][:90]))
