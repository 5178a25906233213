#!/usr/bin/env python
"""Regenerates tests/golden/reference_digests.json from the unmodified reference (needs oracle/_ref, i.e. /root/reference):
python tests/golden/make_golden.py"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (os.path.join(ROOT, "svt-av1_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import common as cm  # noqa: E402
import golden_cases as gc  # noqa: E402

assert cm.have_ref(), "build oracle/_ref first (make -C oracle)"
out = {"source": "svt-av1 v0.8.6 C path (oracle/_ref), SHA-256 of the outputs of tests/golden_cases.py", "digests": {}}
for name in gc.CASES:
    out["digests"][name] = gc.digest(name, "ref")
    print(name, out["digests"][name])
json.dump(out, open(os.path.join(HERE, "reference_digests.json"), "w"), indent=1)
