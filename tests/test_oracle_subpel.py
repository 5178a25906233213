"""The sub-pel refinement oracle (oracle/subpel_oracle.c) pinned against svt_av1_find_best_sub_pixel_tree of the unmodified
reference (oracle/_ref), set up as md_subpel_search does: every block size, the three kernel types, 1 / 2 iterations per
step, with / without 1/8-sample precision, forced stops, every MV cost type, searches cut by the MV limits."""
import ctypes as C

import numpy as np
import pytest

import common as cm
import subpel_cases as sc

needs_ref = pytest.mark.skipif(not cm.have_ref(), reason="oracle/_ref not built")
CONFIGS = [dict(search_type=3, iters=2, allow_hp=1), dict(search_type=3, iters=1, allow_hp=0), dict(search_type=2, iters=2, allow_hp=0),
           dict(search_type=1, iters=2, allow_hp=1), dict(search_type=3, iters=2, allow_hp=1, forced_stop=1),
           dict(search_type=3, iters=2, allow_hp=1, forced_stop=2), dict(search_type=3, iters=2, allow_hp=1, forced_stop=3),
           dict(search_type=3, iters=2, allow_hp=1, cost_type=1), dict(search_type=2, iters=1, allow_hp=1, cost_type=3),
           dict(search_type=3, iters=2, allow_hp=0, cost_type=4), dict(search_type=3, iters=2, allow_hp=1, cost_type=2, epb=1)]


@needs_ref
def test_limits_match_reference():
    rng = np.random.default_rng(3)
    for _ in range(300):
        w, h = int(rng.integers(16, 500)) * 4, int(rng.integers(16, 300)) * 4
        bw, bh = sc.BLOCKS[rng.integers(0, 19)]
        x, y = int(rng.integers(0, (w - bw) // 4 + 1)) * 4, int(rng.integers(0, (h - bh) // 4 + 1)) * 4
        rr, rc = int(rng.integers(-9000, 9001)), int(rng.integers(-9000, 9001))
        mi_cols, mi_rows = 2 * ((w + 7) >> 3), 2 * ((h + 7) >> 3)
        out = (C.c_int16 * 4)()
        cm.refh().refh_subpel_limits(mi_rows, mi_cols, x, y, bw, bh, rr, rc, out)
        assert tuple(out) == sc.limits(mi_rows, mi_cols, x, y, bw, bh, rr, rc)


@needs_ref
@pytest.mark.parametrize("cfg", range(len(CONFIGS)))
def test_subpel_search_matches_reference(cfg):
    w, h = 320, 192
    src, refs = sc.pictures(w, h, 40 + cfg)
    jobs = sc.make_jobs(w, h, len(refs), 132, 50 + cfg)
    p, tabs = sc.params(seed=cfg, **CONFIGS[cfg])
    cm.refh().refh_subpel_search.restype = C.c_int
    a = sc.run_cpu(cm.oracle().orc_subpel_search, p, tabs, src, refs, jobs)
    b = sc.run_cpu(cm.refh().refh_subpel_search, p, tabs, src, refs, jobs)
    for f in ("mv_row", "mv_col", "besterr", "distortion", "sse"):
        np.testing.assert_array_equal(a[f], b[f], f)
    moved = (a["mv_row"] != jobs["start_mv_row"]) | (a["mv_col"] != jobs["start_mv_col"])
    assert moved.any() == (CONFIGS[cfg].get("forced_stop", 0) != 3)
