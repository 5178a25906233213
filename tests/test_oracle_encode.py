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
