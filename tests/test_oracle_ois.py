"""Open-loop intra search (the flavour of presets >= 5: DC_PRED only): the oracle's restatement
against the reference's open_loop_intra_search_mb run over every SB of small pictures, ragged sizes included."""
import numpy as np
import pytest

import common as cm

needs_ref = pytest.mark.skipif(not cm.have_ref(), reason="oracle/_ref not built")


def padded_luma(w, h, seed, pad=80):
    f = cm.synth_yuv(w, h, 1, seed, 8, noise=9)
    return np.ascontiguousarray(np.pad(f.plane(0), pad, mode="edge")), pad


@needs_ref
@pytest.mark.parametrize("geom", [(128, 64), (200, 120), (352, 288), (72, 88), (640, 360), (16, 16), (24, 40), (1920, 1080)])
def test_ois_dc_restatement_matches_reference(geom):
    w, h = geom
    orc, refh = cm.oracle(), cm.refh()
    buf, pad = padded_luma(w, h, 60 + w)
    mbw, mbh = (w + 15) // 16, (h + 15) // 16
    want_cost, want_mode = np.zeros(mbw * mbh, np.int64), np.zeros(mbw * mbh, np.int32)
    assert refh.refh_ois_picture(cm.ptr(buf), buf.shape[1], pad, pad, w, h, 8, cm.ptr(want_cost), cm.ptr(want_mode)) == 0
    assert (want_mode == 0).all() and (want_cost >= 0).all()  # every macroblock with its origin inside the picture is searched, DC_PRED
    got = np.zeros(mbw * mbh, np.int64)
    import ctypes as C
    orc.orc_ois_dc_picture(C.c_void_p(buf.ctypes.data + pad * buf.shape[1] + pad), buf.shape[1], w, h, cm.ptr(got))
    np.testing.assert_array_equal(got, want_cost)
