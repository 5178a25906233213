"""The CDEF strength decision: the oracle's restatement against the reference's finish_cdef_search on synthetic mse
tables (all four strength tables, skip patterns, ties)."""
import ctypes as C

import numpy as np
import pytest

import common as cm
import svtb200 as sb

needs_ref = pytest.mark.skipif(not cm.have_ref(), reason="oracle/_ref not built")


def decide_case(seed, mi_rows, mi_cols, n_strengths, kind):
    rng = np.random.default_rng(seed)
    nvfb, nhfb = (mi_rows + 15) // 16, (mi_cols + 15) // 16
    nfb = nvfb * nhfb
    rows8, cols8 = (mi_rows + 1) // 2, (mi_cols + 1) // 2
    stride = (cols8 + 15) & ~15
    skip = (rng.random((rows8, stride)) < 0.3).astype(np.uint8)
    for fb in rng.choice(nfb, max(1, nfb // 5), replace=False):  # some filter blocks entirely skipped
        r, c = fb // nhfb, fb % nhfb
        skip[8 * r:8 * r + 8, 8 * c:8 * c + 8] = 1
    mse = np.zeros((2, nfb, 64), np.uint64)
    if kind == "smooth":  # convex-ish in the strength index, minimum position varies per block
        for pl in range(2):
            for fb in range(nfb):
                best = rng.integers(0, n_strengths)
                mse[pl, fb, :n_strengths] = (np.abs(np.arange(n_strengths) - best) * rng.integers(1, 4000) + rng.integers(1000, 90000)).astype(np.uint64)
    elif kind == "ties":
        mse[:, :, :n_strengths] = rng.integers(0, 4, (2, nfb, n_strengths)).astype(np.uint64) * 1000
    else:
        mse[:, :, :n_strengths] = rng.integers(0, 1 << 24, (2, nfb, n_strengths)).astype(np.uint64)
    return skip, stride, mse


@needs_ref
@pytest.mark.parametrize("case", [(0, 3, 68, 120, 43, "smooth"), (1, 3, 68, 120, 20, "rand"), (2, 2, 45, 80, 50, "smooth"), (3, 1, 45, 80, 30, "ties"),
                                  (4, 0, 34, 46, 35, "smooth"), (5, 3, 16, 16, 60, "rand"), (6, 3, 90, 160, 10, "ties"), (7, 3, 270, 480, 43, "smooth")])
def test_decide_restatement_matches_finish_cdef_search(case):
    seed, pick, mi_rows, mi_cols, qidx, kind = case
    orc, refh = cm.oracle(), cm.refh()
    p = sb.CdefDecideParams()
    n = orc.orc_cdef_decide_table(pick, C.byref(p))
    assert n == (64, 32, 20, 10)[pick]
    p.mi_rows, p.mi_cols = mi_rows, mi_cols
    skip, stride, mse = decide_case(seed, mi_rows, mi_cols, n, kind)
    nfb = mse.shape[1]
    bits, nb, lam = C.c_int32(), C.c_int32(), C.c_uint64()
    ys, uvs, fbs = (C.c_int32 * 8)(), (C.c_int32 * 8)(), np.zeros(nfb, np.int8)
    cdef_level = {0: 1, 1: 2, 2: 3, 3: 4}[pick]
    assert refh.refh_cdef_finish(mi_rows, mi_cols, qidx * 4, cdef_level, 8, cm.ptr(skip), stride, cm.ptr(mse), C.byref(bits), C.byref(nb), ys, uvs,
                                 cm.ptr(fbs), C.byref(lam)) == 0
    p.lambda_ = lam.value
    out, got_fb = sb.CdefDecision(), np.zeros(nfb, np.int8)
    orc.orc_cdef_decide(C.byref(p), cm.ptr(mse), cm.ptr(skip), stride, C.byref(out), cm.ptr(got_fb))
    assert (out.cdef_bits, out.nb_cdef_strengths) == (bits.value, nb.value)
    assert list(out.y_strength)[:nb.value] == list(ys)[:nb.value] and list(out.uv_strength)[:nb.value] == list(uvs)[:nb.value]
    np.testing.assert_array_equal(got_fb, fbs)
    assert lam.value > 0


@needs_ref
def test_decide_restatement_all_blocks_skipped_and_single_block():
    """Edge cases of finish_cdef_search: no filter block takes part (every 8x8 skipped) and a one-block picture."""
    orc, refh = cm.oracle(), cm.refh()
    for (mi_rows, mi_cols, all_skip) in ((34, 46, True), (16, 16, False), (9, 13, False)):
        p = sb.CdefDecideParams()
        n = orc.orc_cdef_decide_table(3, C.byref(p))
        p.mi_rows, p.mi_cols = mi_rows, mi_cols
        skip, stride, mse = decide_case(77, mi_rows, mi_cols, n, "rand")
        if all_skip:
            skip[...] = 1
        else:
            skip[...] = 0
        nfb = mse.shape[1]
        bits, nb, lam = C.c_int32(), C.c_int32(), C.c_uint64()
        ys, uvs, fbs = (C.c_int32 * 8)(), (C.c_int32 * 8)(), np.zeros(nfb, np.int8)
        assert refh.refh_cdef_finish(mi_rows, mi_cols, 120, 4, 8, cm.ptr(skip), stride, cm.ptr(mse), C.byref(bits), C.byref(nb), ys, uvs,
                                     cm.ptr(fbs), C.byref(lam)) == 0
        p.lambda_ = lam.value
        out, got_fb = sb.CdefDecision(), np.zeros(nfb, np.int8)
        orc.orc_cdef_decide(C.byref(p), cm.ptr(mse), cm.ptr(skip), stride, C.byref(out), cm.ptr(got_fb))
        assert (out.cdef_bits, out.nb_cdef_strengths) == (bits.value, nb.value)
        k = nb.value
        assert list(out.y_strength)[:k] == list(ys)[:k] and list(out.uv_strength)[:k] == list(uvs)[:k]
        np.testing.assert_array_equal(got_fb, fbs)
        assert out.sb_count == (0 if all_skip else nfb)
