#!/usr/bin/env python
"""Per-kernel time shares from an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import csv, sys
from collections import defaultdict
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value'); bi = hdr.index('Block Size')
agg = defaultdict(list)
for r in rows[1:]:
    try: agg[r[ki][:60] + ' ' + r[bi]].append(float(r[vi].replace(',', '')))
    except Exception: pass
tot = sum(sum(v) for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print('%-82s n=%3d avg %8.1f us share %.1f%%' % (k, len(v), sum(v) / len(v) / 1000, 100 * sum(v) / tot))
