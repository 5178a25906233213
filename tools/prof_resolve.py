"""Resolves the raw samples written by tools/prof_preload.c into a flat per-function profile using `nm` on each object.

    python tools/prof_resolve.py /tmp/prof.txt.<pid> [top_n]
"""
import bisect
import collections
import subprocess
import sys


def symbols(obj):
    syms = {}
    for flags in (["-n", "--defined-only"], ["-n", "-D", "--defined-only"]):  # static table, then the dynamic one (stripped libs)
        out = subprocess.run(["nm"] + flags + [obj], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout.decode()
        for line in out.splitlines():
            parts = line.split()
            if len(parts) == 3 and parts[1] in "tTwWiI":
                syms.setdefault(int(parts[0], 16), parts[2].split("@")[0])
    addrs = sorted(syms)
    return addrs, [syms[a] for a in addrs]


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
    per_obj = collections.defaultdict(list)
    n = 0
    for line in open(path):
        if line.startswith("#"):
            continue
        obj, off = line.split()
        per_obj[obj].append(int(off, 16))
        n += 1
    hist = collections.Counter()
    for obj, offs in per_obj.items():
        addrs, names = symbols(obj) if obj != "?" else ([], [])
        short = obj.rsplit("/", 1)[-1]
        for o in offs:
            i = bisect.bisect_right(addrs, o) - 1
            hist[(names[i] if i >= 0 else "?", short)] += 1
    print("# %d samples (1 ms of thread CPU time each)" % n)
    for (name, obj), c in hist.most_common(top):
        print("%6.2f%% %8d  %s  [%s]" % (100.0 * c / max(n, 1), c, name, obj))


if __name__ == "__main__":
    main()
