"""Temporal filter on the GPU against the oracle (tests/test_oracle_tf.py pins the oracle to the
reference's svt_av1_apply_temporal_filter_planewise_c / _hbd_c and its expf to the host libm, exhaustively)."""
import ctypes as C

import numpy as np
import pytest

import common as cm
import svtb200 as sb
import tf_cases as tc

pytestmark = pytest.mark.gpu


def _orc():
    orc = cm.oracle()
    orc.orc_expf_checksum.restype = C.c_uint64
    orc.orc_expf_checksum.argtypes = [C.c_uint32, C.c_uint32]
    orc.orc_tf_normalize.restype = C.c_uint64
    return orc


def test_device_expf_equals_the_oracle_expf_on_every_float_of_the_filter_range():
    """The kernel's expf over all 1.09e9 floats of [-8, -0] (position-weighted checksum of the result bit patterns, chunked
    so that a difference is localised) equals the oracle's, which equals the host libm's (test_oracle_tf.py)."""
    lib, orc = sb.load(), _orc()
    lo, hi = 0x80000000, int(np.float32(-8.0).view(np.uint32))
    step = 1 << 26
    for a in range(lo, hi + 1, step):
        b = min(a + step - 1, hi)
        got = C.c_uint64(0)
        sb.check(lib.svt_b200_tf_expf_checksum(a, b, C.byref(got)), lib)
        assert got.value == orc.orc_expf_checksum(a, b), hex(a)


@pytest.mark.parametrize("kw", tc.CASES)
def test_planewise_block_dropin_vs_oracle(kw):
    lib, orc = sb.load(), _orc()
    c = tc.make_case(**kw)
    want = tc.run_oracle(orc, c)
    den, be, df = tc.factors(orc, c)
    got = tc.fresh(c)
    sb.check(lib.svt_b200_tf_planewise_block_host(c.bd, c.chroma, tc.ptr(c.y_src), c.ys, tc.ptr(c.y_pre), c.ps, tc.ptr(c.u_src), tc.ptr(c.v_src),
                                                  c.uvs, tc.ptr(c.u_pre), tc.ptr(c.v_pre), c.ups, c.bw, c.bh, den, be, df,
                                                  *[tc.ptr(a) for a in got]), lib)
    for name, a, b in zip(("y_accum", "y_count", "u_accum", "u_count", "v_accum", "v_count"), got, want):
        np.testing.assert_array_equal(a, b, err_msg=name)


@pytest.mark.parametrize("geom", [(1920, 1080, 8), (640, 360, 10), (3840, 2160, 10)])
def test_picture_level_filter_vs_oracle(geom):
    """One (centre, motion-compensated reference) pair of a whole picture: svt_b200_tf_central + svt_b200_tf_planewise over
    every 32x32 block + svt_b200_tf_normalize, against the oracle applied block by block (all blocks at 360p, a spread of
    blocks at the larger sizes - the device handles all of them; untouched samples must equal the centre frame)."""
    import torch
    import gpu_runner as gr
    lib, orc = sb.load(), _orc()
    w, h, bd = geom
    rng = np.random.default_rng(w + bd)
    src = cm.synth_yuv(w, h, 1, 31, bd)
    pred = cm.degrade(src, 32, amp=6)
    chroma = 1
    xs, ys = range(0, w - 31, 32), range(0, h - 31, 32)
    blocks = [(x, y) for y in ys for x in xs]
    n = len(blocks)
    be = rng.integers(0, 40 * 256, (n, 4)) / 256.0
    dfac = np.maximum(rng.integers(0, 300, (n, 4)) / 100.0, 1.0)
    noise = (C.c_double * 3)(2.0, 1.5, 1.0)
    den = (C.c_double * 3)()
    orc.orc_tf_den(4, noise, den)
    arr = (sb.TfBlock * n)()
    for i, (x, y) in enumerate(blocks):
        arr[i].x, arr[i].y = x, y
        for q in range(4):
            arr[i].block_error[q], arr[i].d_factor[q] = float(be[i, q]), float(dfac[i, q])
    dblocks = torch.from_numpy(np.frombuffer(arr, dtype=np.uint8).copy()).cuda()
    ds, dp = gr.DevYuv(src), gr.DevYuv(pred)
    cw, ch = (w + 1) // 2, (h + 1) // 2
    sy, sc = w + 8, cw + 8
    acc = [torch.zeros(h * sy, dtype=torch.int32, device="cuda"), torch.zeros(ch * sc, dtype=torch.int32, device="cuda"),
           torch.zeros(ch * sc, dtype=torch.int32, device="cuda")]
    cnt = [torch.zeros(h * sy, dtype=torch.int16, device="cuda"), torch.zeros(ch * sc, dtype=torch.int16, device="cuda"),
           torch.zeros(ch * sc, dtype=torch.int16, device="cuda")]
    ta = sb.TfAccum()
    for i in range(3):
        ta.accum[i], ta.count[i] = acc[i].data_ptr(), cnt[i].data_ptr()
    ta.stride_y, ta.stride_c = sy, sc
    p = sb.TfParams()
    for i in range(3):
        p.den[i] = den[i]
    p.chroma, p.block_w, p.block_h = chroma, 32, 32
    ss, ps = ds.struct(), dp.struct()
    sb.check(lib.svt_b200_tf_central(C.byref(ss), C.byref(ta), chroma, None), lib)
    sb.check(lib.svt_b200_tf_planewise(C.byref(p), C.byref(ss), C.byref(ps), C.c_void_p(dblocks.data_ptr()), n, C.byref(ta), None), lib)
    torch.cuda.synchronize()
    g_acc = [a.cpu().numpy().view(np.uint32).reshape(-1, s) for a, s in zip(acc, (sy, sc, sc))]
    g_cnt = [a.cpu().numpy().view(np.uint16).reshape(-1, s) for a, s in zip(cnt, (sy, sc, sc))]
    # oracle: central over the picture, then the chosen blocks
    hbd = int(bd > 8)
    w_acc = [np.zeros((h, sy), np.uint32), np.zeros((ch, sc), np.uint32), np.zeros((ch, sc), np.uint32)]
    w_cnt = [np.zeros((h, sy), np.uint16), np.zeros((ch, sc), np.uint16), np.zeros((ch, sc), np.uint16)]
    for i in range(3):
        pl = np.ascontiguousarray(src.plane(i))
        orc.orc_tf_central(tc.ptr(pl), pl.shape[1], pl.shape[1], pl.shape[0], hbd, tc.ptr(w_acc[i]), tc.ptr(w_cnt[i]), sy if i == 0 else sc)
    pick = range(n) if n <= 300 else sorted(set([0, n - 1, len(xs) - 1] + list(rng.choice(n, 120, replace=False))))
    S, P = [np.ascontiguousarray(src.plane(i)) for i in range(3)], [np.ascontiguousarray(pred.plane(i)) for i in range(3)]
    es = 2 if hbd else 1

    def at(a, y, x):
        return C.c_void_p(a.ctypes.data + (y * a.shape[1] + x) * a.itemsize)
    for i in pick:
        x, y = blocks[i]
        # the oracle uses the PREDICTION's stride for accum / count: run it on block-sized copies, then paste
        ya, yc = np.ascontiguousarray(w_acc[0][y:y + 32, x:x + 32]), np.ascontiguousarray(w_cnt[0][y:y + 32, x:x + 32])
        ua, uc = np.ascontiguousarray(w_acc[1][y // 2:y // 2 + 16, x // 2:x // 2 + 16]), np.ascontiguousarray(w_cnt[1][y // 2:y // 2 + 16, x // 2:x // 2 + 16])
        va, vc = np.ascontiguousarray(w_acc[2][y // 2:y // 2 + 16, x // 2:x // 2 + 16]), np.ascontiguousarray(w_cnt[2][y // 2:y // 2 + 16, x // 2:x // 2 + 16])
        yp = np.ascontiguousarray(P[0][y:y + 32, x:x + 32])
        up, vp = np.ascontiguousarray(P[1][y // 2:y // 2 + 16, x // 2:x // 2 + 16]), np.ascontiguousarray(P[2][y // 2:y // 2 + 16, x // 2:x // 2 + 16])
        orc.orc_tf_planewise(at(S[0], y, x), S[0].shape[1], tc.ptr(yp), 32, at(S[1], y // 2, x // 2), at(S[2], y // 2, x // 2), S[1].shape[1],
                             tc.ptr(up), tc.ptr(vp), 16, 32, 32, 1, 1, den, (C.c_double * 4)(*be[i]), (C.c_double * 4)(*dfac[i]), chroma, bd,
                             tc.ptr(ya), tc.ptr(yc), tc.ptr(ua), tc.ptr(uc), tc.ptr(va), tc.ptr(vc))
        for (ga, gc, wa, wc, yy, xx, sz) in ((g_acc[0], g_cnt[0], ya, yc, y, x, 32), (g_acc[1], g_cnt[1], ua, uc, y // 2, x // 2, 16),
                                           (g_acc[2], g_cnt[2], va, vc, y // 2, x // 2, 16)):
            np.testing.assert_array_equal(ga[yy:yy + sz, xx:xx + sz], wa, err_msg=f"accum block {i}")
            np.testing.assert_array_equal(gc[yy:yy + sz, xx:xx + sz], wc, err_msg=f"count block {i}")
    # normalisation of the device accumulators, against the oracle's formula on the same accumulators
    sse = torch.zeros(2, dtype=torch.int64, device="cuda")
    sb.check(lib.svt_b200_tf_normalize(C.byref(ss), C.byref(ta), chroma, C.c_void_p(sse.data_ptr()), None), lib)
    torch.cuda.synchronize()
    out = ds.download()
    tot = [0, 0]
    for i in range(3):
        pl = np.ascontiguousarray(src.plane(i))
        st = sy if i == 0 else sc
        a, c_ = np.ascontiguousarray(g_acc[i]), np.ascontiguousarray(g_cnt[i])
        tot[1 if i else 0] += orc.orc_tf_normalize(tc.ptr(pl), pl.shape[1], pl.shape[1], pl.shape[0], hbd, tc.ptr(a), tc.ptr(c_), st)
        np.testing.assert_array_equal(out.plane(i), pl, err_msg=f"plane {i}")
    assert [int(v) for v in sse.cpu().numpy()] == tot
    # samples outside every block only saw the centre frame: unchanged
    if w % 32:
        np.testing.assert_array_equal(out.plane(0)[:, (w // 32) * 32:], src.plane(0)[:, (w // 32) * 32:])
