"""Temporal filter: the oracle's restatement against the reference's
svt_av1_apply_temporal_filter_planewise_c / _hbd_c, and its expf against the host libm for EVERY float the filter can
evaluate (the float chain of the reference is reproduced exactly, not approximately)."""
import ctypes as C

import numpy as np
import pytest

import common as cm
import tf_cases as tc

needs_ref = pytest.mark.skipif(not cm.have_ref(), reason="oracle/_ref not built")


def _orc():
    orc = cm.oracle()
    orc.orc_expf.restype = C.c_float
    orc.orc_expf.argtypes = [C.c_float]
    orc.orc_expf_mismatches.restype = C.c_long
    orc.orc_expf_mismatches.argtypes = [C.c_uint32, C.c_uint32, C.POINTER(C.c_long)]
    orc.orc_tf_normalize.restype = C.c_uint64
    return orc


def test_expf_restatement_equals_libm_on_every_float_of_the_filter_range():
    """All 1 090 519 041 floats of [-8, -0] (the filter evaluates expf on [-7, -0]): identical bit patterns and identical
    weights (int)(e * 1000).  ~20 s of CPU on 8 threads."""
    orc = _orc()
    lo, hi = 0x80000000, int(np.float32(-8.0).view(np.uint32))
    w = C.c_long(-1)
    assert orc.orc_expf_mismatches(lo, hi, C.byref(w)) == 0
    assert w.value == 0
    assert orc.orc_expf(0.0) == 1.0


@needs_ref
@pytest.mark.parametrize("kw", tc.CASES)
def test_planewise_restatement_matches_reference(kw):
    orc, refh = _orc(), cm.refh()
    c = tc.make_case(**kw)
    want, got = tc.run_reference(refh, c), tc.run_oracle(orc, c)
    for name, a, b in zip(("y_accum", "y_count", "u_accum", "u_count", "v_accum", "v_count"), got, want):
        np.testing.assert_array_equal(a, b, err_msg=name)
    if kw["amp"] <= 12:  # the mild cases must produce non-zero weights (the extreme ones may legitimately give weight 0)
        assert (got[1] != c.y_cnt0).any()


def test_central_and_normalize():
    orc = _orc()
    rng = np.random.default_rng(5)
    for bd in (8, 10):
        dt = np.uint16 if bd > 8 else np.uint8
        pre = rng.integers(0, 1 << bd, (16, 40)).astype(dt)
        acc, cnt = np.zeros((16, 48), np.uint32), np.zeros((16, 48), np.uint16)
        orc.orc_tf_central(tc.ptr(pre), 40, 32, 16, int(bd > 8), tc.ptr(acc), tc.ptr(cnt), 48)
        assert (acc[:, :32] == 1000 * pre[:, :32].astype(np.uint32)).all() and (cnt[:, :32] == 1000).all() and not acc[:, 32:].any()
        acc[:, :32] += rng.integers(0, 1 << (bd + 9), (16, 32)).astype(np.uint32)
        cnt[:, :32] += rng.integers(0, 900, (16, 32)).astype(np.uint16)
        dst = pre.copy()
        sse = orc.orc_tf_normalize(tc.ptr(dst), 40, 32, 16, int(bd > 8), tc.ptr(acc), tc.ptr(cnt), 48)
        want = (acc[:, :32].astype(np.int64) + (cnt[:, :32] >> 1)) // cnt[:, :32]
        np.testing.assert_array_equal(dst[:, :32], want.astype(dt))
        assert sse == int(((pre[:, :32].astype(np.int64) - want) ** 2).sum())


@needs_ref
def test_planewise_restatement_random_sweep():
    """40 more random blocks (both bit depths, chroma on / off, split and unsplit, small and large motion, noise levels,
    decay controls): every accumulator and counter equal to the reference's."""
    orc, refh = _orc(), cm.refh()
    rng = np.random.default_rng(2024)
    for k in range(40):
        kw = dict(seed=100 + k, bd=int(rng.choice([8, 10])), chroma=int(rng.integers(0, 2)), split=[None, 0, 1][int(rng.integers(0, 3))],
                  big_motion=bool(rng.integers(0, 2)), noise=tuple(float(x) for x in rng.uniform(0, 8, 3)), decay=int(rng.integers(2, 5)),
                  amp=int(rng.choice([1, 3, 8, 20, 50])))
        c = tc.make_case(**kw)
        want, got = tc.run_reference(refh, c), tc.run_oracle(orc, c)
        for name, a, b in zip(("y_accum", "y_count", "u_accum", "u_count", "v_accum", "v_count"), got, want):
            np.testing.assert_array_equal(a, b, err_msg="%s case %r" % (name, kw))
