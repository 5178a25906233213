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


VAR_SIZES = [(4, 4), (4, 8), (4, 16), (8, 4), (8, 8), (8, 16), (8, 32), (16, 4), (16, 8), (16, 16), (16, 32), (16, 64), (32, 8), (32, 16),
             (32, 32), (32, 64), (64, 16), (64, 32), (64, 64), (64, 128), (128, 64), (128, 128)]


@needs_ref
def test_variance_restatement_matches_reference():
    """misc_oracle.variance vs svt_aom_varianceWxH_c (all 22 sizes) and svt_aom_mse16x16_c, random and extreme blocks."""
    import misc_oracle as mo
    lib = cm.ref()
    rng = np.random.default_rng(8)
    for (w, h) in VAR_SIZES:
        fn = getattr(lib, f"svt_aom_variance{w}x{h}_c")
        fn.restype = C.c_uint32
        for kind in ("rand", "extreme", "flat"):
            a = rng.integers(0, 256, (h, w + 7)).astype(np.uint8) if kind == "rand" else np.full((h, w + 7), 255, np.uint8)
            b = rng.integers(0, 256, (h, w + 3)).astype(np.uint8) if kind == "rand" else np.zeros((h, w + 3), np.uint8)
            if kind == "flat":
                b[...] = 250
            sse = C.c_uint32(0)
            want = fn(cm.ptr(a), a.shape[1], cm.ptr(b), b.shape[1], C.byref(sse))
            got = mo.variance(a[:, :w], b[:, :w])
            assert got == (want, sse.value), (w, h, kind)
    lib.svt_aom_mse16x16_c.restype = C.c_uint32
    a, b = rng.integers(0, 256, (16, 16)).astype(np.uint8), rng.integers(0, 256, (16, 16)).astype(np.uint8)
    sse = C.c_uint32(0)
    assert mo.variance(a, b) == (lib.svt_aom_mse16x16_c(cm.ptr(a), 16, cm.ptr(b), 16, C.byref(sse)), sse.value)


@needs_ref
@pytest.mark.parametrize("search", [1, 2, 3])
def test_upsampled_pred_matches_reference(search):
    lib = cm.ref()
    rng = np.random.default_rng(20 + search)
    for (w, h) in [(4, 4), (8, 8), (16, 32), (64, 64), (128, 128), (32, 8)]:
        ref = rng.integers(0, 256, (h + 16, w + 16)).astype(np.uint8)
        for sx in range(8):
            for sy in (0, 3, 7) if sx % 2 else (0, 1, 4):
                a, b = np.zeros((h, w), np.uint8), np.ones((h, w), np.uint8)
                r0 = C.c_void_p(ref.ctypes.data + 8 * ref.shape[1] + 8)
                cm.oracle().orc_upsampled_pred(cm.ptr(a), w, h, sx, sy, r0, ref.shape[1], search)
                lib.svt_aom_upsampled_pred_c(None, None, 0, 0, None, cm.ptr(b), w, h, sx, sy, r0, ref.shape[1], search)
                np.testing.assert_array_equal(a, b, f"{w}x{h} {sx},{sy}")
