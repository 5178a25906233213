"""The oracle against golden digests of the reference's own outputs (tests/golden/reference_digests.json, written by
tests/golden/make_golden.py from oracle/_ref): runs everywhere, with or without the reference sources."""
import json
import os

import pytest

import golden_cases as gc

GOLDEN = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_digests.json")))["digests"]


@pytest.mark.parametrize("name", sorted(gc.CASES))
def test_oracle_reproduces_reference_digest(name):
    assert name in GOLDEN, "run tests/golden/make_golden.py"
    assert gc.digest(name, "oracle") == GOLDEN[name]
