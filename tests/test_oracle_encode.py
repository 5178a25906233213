"""The per-TU EncDec chain (residual -> fwd txfm -> quant -> inverse -> recon): the oracle composite used by the GPU
tests vs the same chain run through the reference's own RTCD pointers (oracle/_ref harness)."""
import ctypes as C

import numpy as np
import pytest

import common as cm
import svtb200 as sb

needs_ref = pytest.mark.skipif(not cm.have_ref(), reason="oracle/_ref not built")


@needs_ref
@pytest.mark.parametrize("case", [(ts, bd, fp) for ts in range(19) for bd, fp in ((8, 0), (10, 1))] + [(2, 8, 1), (4, 10, 0)])
def test_oracle_chain_matches_reference_chain(case):
    from test_txfm_gpu import make_tus, oracle_encode_tus, quant_plane
    ts, bd, use_fp = case
    rng = np.random.default_rng(1000 + ts)
    W, H = 192, 128
    src = cm.synth_yuv(W, H, 1, 7, bd)
    pred = cm.degrade(src, 11, amp=14)
    p = sb.EncodeParams()
    p.tx_size, p.use_fp = ts, use_fp
    for i in range(3):
        p.q[i] = quant_plane(rng, bd)
    tus = make_tus(rng, ts, W, H, limit=60)
    want_rec, want_q, want_eob = oracle_encode_tus(p, src, pred, tus, use_fp)
    rec = pred.copy()
    n = min(sb.TX_W[ts], 32) * min(sb.TX_H[ts], 32)
    q = np.zeros((len(tus), n), np.int32)
    eob = np.zeros(len(tus), np.uint16)
    arr = (sb.Tu * len(tus))(*tus)
    ss, ps, rs = src.struct(), pred.struct(), rec.struct()
    cm.refh().refh_encode_tus(C.byref(p), C.byref(ss), C.byref(ps), C.byref(rs), arr, len(tus), cm.ptr(q), cm.ptr(eob))
    np.testing.assert_array_equal(eob, want_eob)
    np.testing.assert_array_equal(q, want_q)
    for i in range(3):
        np.testing.assert_array_equal(rec.plane(i), want_rec.plane(i))


@needs_ref
@pytest.mark.parametrize("tx_size", range(19))
def test_oracle_estimate_transform_matches_reference_dispatcher(tx_size):
    """orc_estimate_transform (shape dispatcher restated) vs the reference's av1_estimate_transform for the four
    EB_TRANS_COEFF_SHAPE values: the coefficients the quantiser reads (min(w,32) x min(h,32)) and the reported energy."""
    from test_oracle_txfm import allowed_types, residual_block
    orc, ref = cm.oracle(), cm.refh()
    orc.orc_estimate_transform.restype = C.c_uint64
    ref.refh_estimate_transform.restype = C.c_uint64
    w, h = sb.TX_W[tx_size], sb.TX_H[tx_size]
    n = min(w, 32) * min(h, 32)
    rng = np.random.default_rng(900 + tx_size)
    for bd in (8, 10):
        for shape in range(4):
            for tx_type in allowed_types(tx_size):
                res = residual_block(rng, w, h, bd, "rand")
                a, b = np.zeros(w * h, np.int32), np.zeros(w * h, np.int32)
                stride = res.shape[1] if res.ndim == 2 else w
                ea = orc.orc_estimate_transform(cm.ptr(res), C.c_uint32(stride), cm.ptr(a), tx_size, bd, tx_type, shape)
                eb = ref.refh_estimate_transform(cm.ptr(res), C.c_uint32(stride), cm.ptr(b), tx_size, bd, tx_type, shape)
                np.testing.assert_array_equal(a[:n], b[:n], err_msg=f"bd {bd} shape {shape} type {tx_type}")
                assert ea == eb
                if shape == 3:
                    assert not a[1:n].any()
