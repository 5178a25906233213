"""Picture-analysis block statistics: the oracle's restatement against the reference's own
compute_block_mean_compute_variance / compute_chroma_block_mean (one SB at a time, both precision flavours)."""
import ctypes as C

import numpy as np
import pytest

import common as cm

needs_ref = pytest.mark.skipif(not cm.have_ref(), reason="oracle/_ref not built")


def sb_planes(rng, kind):
    if kind == "rand":
        y = rng.integers(0, 256, (96, 136)).astype(np.uint8)
    elif kind == "flat":
        y = np.full((96, 136), 200, np.uint8)
    elif kind == "extreme":
        y = np.where(rng.random((96, 136)) < 0.5, 255, 0).astype(np.uint8)
    else:
        yy, xx = np.mgrid[0:96, 0:136]
        y = np.clip(128 + 90 * np.sin(xx / 6.0) * np.cos(yy / 9.0) + rng.integers(-4, 5, (96, 136)), 0, 255).astype(np.uint8)
    cb = rng.integers(0, 256, (48, 72)).astype(np.uint8)
    cr = rng.integers(0, 256, (48, 72)).astype(np.uint8)
    return y, cb, cr


@needs_ref
def test_sb_statistics_match_reference():
    """BLOCK_MEAN_PREC_SUB only: the reference's FULL flavour calls compute_mean_8x8, an RTCD pointer that is never assigned."""
    full = 0
    orc, refh = cm.oracle(), cm.refh()
    rng = np.random.default_rng(40 + full)
    for kind in ("rand", "flat", "extreme", "smooth", "rand"):
        y, cb, cr = sb_planes(rng, kind)
        oy, ox = int(rng.integers(0, 24)), int(rng.integers(0, 60))
        li, ci = oy * y.shape[1] + ox, (oy // 2) * cb.shape[1] + ox // 2
        w_ym, w_var, w_cb, w_cr = np.zeros(85, np.uint8), np.zeros(85, np.uint16), np.zeros(85, np.uint8), np.zeros(85, np.uint8)
        assert refh.refh_sb_mean_variance(cm.ptr(y), y.shape[1], li, cm.ptr(cb), cm.ptr(cr), cb.shape[1], ci, full, cm.ptr(w_ym), cm.ptr(w_var),
                                          cm.ptr(w_cb), cm.ptr(w_cr)) == 0
        g_ym, g_var, g_cb, g_cr = np.zeros(85, np.uint8), np.zeros(85, np.uint16), np.zeros(21, np.uint8), np.zeros(21, np.uint8)
        orc.orc_sb_mean_variance(C.c_void_p(y.ctypes.data + li), y.shape[1], full, cm.ptr(g_ym), cm.ptr(g_var))
        orc.orc_sb_chroma_mean(C.c_void_p(cb.ctypes.data + ci), cb.shape[1], full, cm.ptr(g_cb))
        orc.orc_sb_chroma_mean(C.c_void_p(cr.ctypes.data + ci), cr.shape[1], full, cm.ptr(g_cr))
        np.testing.assert_array_equal(g_ym, w_ym, err_msg=kind)
        np.testing.assert_array_equal(g_var, w_var, err_msg=kind)
        np.testing.assert_array_equal(g_cb, w_cb[:21], err_msg=kind)
        np.testing.assert_array_equal(g_cr, w_cr[:21], err_msg=kind)
        assert not w_cb[21:].any()  # the reference writes the 64x64 / 32x32 / 16x16 entries only
