"""BASELINE configs[2] geometry (3840x2160, 10-bit): the EncDec transform-unit kernel and the three in-loop filters on a whole
4K picture, every sample compared with the oracle (parity case, not a bench line)."""
import ctypes as C

import numpy as np
import pytest

import common as cm
import svtb200 as sb

pytestmark = pytest.mark.gpu
W, H, BD = 3840, 2160, 10


def test_dlf_frame_4k_10bit():
    import gpu_runner as gr
    from test_dlf_gpu import flat_mi
    from test_oracle_dlf import dlf_case, dlf_params
    levels = (24, 20, 14, 10)
    mi_rows, mi_cols, part, frame = dlf_case(W, H, BD, 9, levels, 1)
    flat = flat_mi(mi_rows, mi_cols, part, levels)
    p = dlf_params(mi_rows, mi_cols, levels, 1)
    want = frame.copy()
    st = want.struct()
    cm.oracle().orc_dlf_frame(C.byref(p), C.byref(st), flat)
    got = gr.run_gpu_dlf(p, frame, flat)
    for i in range(3):
        np.testing.assert_array_equal(got.plane(i), want.plane(i), f"plane {i}")
    assert (want.plane(0) != frame.plane(0)).any()


def test_cdef_search_and_apply_4k_10bit():
    import gpu_runner as gr
    from test_oracle_cdef import cdef_picture_case
    src, rec, mi_rows, mi_cols, skip = cdef_picture_case(W, H, BD)
    p = sb.CdefSearchParams()
    p.mi_rows, p.mi_cols, p.pri_damping = mi_rows, mi_cols, 5
    cm.oracle().orc_cdef_strength_table(3, C.byref(p))
    nfb = ((mi_rows + 15) // 16) * ((mi_cols + 15) // 16)
    want = np.zeros((2, nfb, 64), np.uint64)
    rs, ss = rec.struct(), src.struct()
    cm.oracle().orc_cdef_search(C.byref(p), C.byref(rs), C.byref(ss), cm.ptr(skip), skip.shape[1], cm.ptr(want))
    got = gr.run_gpu_cdef_search(p, rec, src, skip)
    np.testing.assert_array_equal(got, want)
    pa = sb.CdefApplyParams()
    pa.mi_rows, pa.mi_cols, pa.damping = mi_rows, mi_cols, 5
    for i, (a, b) in enumerate(zip((0, 5, 17, 63, 40, 2, 12, 33), (0, 0, 9, 62, 4, 1, 60, 3))):
        pa.y_strength[i], pa.uv_strength[i] = a, b
    idx = (np.arange(nfb) % 8).astype(np.int8)
    out = rec.copy()
    os_ = out.struct()
    cm.oracle().orc_cdef_apply(C.byref(pa), C.byref(rs), C.byref(os_), cm.ptr(skip), skip.shape[1], cm.ptr(idx))
    got = gr.run_gpu_cdef_apply(pa, rec, skip, idx)
    for i in range(3):
        np.testing.assert_array_equal(got.plane(i), out.plane(i), f"plane {i}")


def test_lr_frame_4k_10bit():
    import gpu_runner as gr
    from test_oracle_lr_frame import lr_case, run_oracle_lr
    cdef, dblk, units = lr_case(W, H, BD, 7, (64, 32, 32), ("mix", "mix", "mix"))
    want = run_oracle_lr(cdef, dblk, units, (64, 32, 32), (3, 3, 3), 0)
    got = gr.run_gpu_lr(cdef, dblk, units, (64, 32, 32), (3, 3, 3), 0)
    for i in range(3):
        np.testing.assert_array_equal(got.plane(i), want.plane(i), f"plane {i}")


def test_encode_tus_4k_10bit():
    """Every 32x32 luma / 16x16 chroma transform unit of the 4K picture on the GPU; a random subset through the oracle."""
    import gpu_runner as gr
    from test_txfm_gpu import make_tus, oracle_encode_tus, quant_plane
    rng = np.random.default_rng(33)
    src = cm.synth_yuv(W, H, 2, 3, BD)
    pred = cm.degrade(src, 5, amp=10)
    p = sb.EncodeParams()
    p.tx_size, p.use_fp = 3, 1
    for i in range(3):
        p.q[i] = quant_plane(rng, BD)
    tus = make_tus(rng, 3, W, H, planes=(0,))
    sub = [tus[i] for i in rng.choice(len(tus), 300, replace=False)]
    want_rec, want_q, want_eob = oracle_encode_tus(p, src, pred, sub, 1)
    got_rec, got_q, got_eob = gr.run_gpu_encode_tus(p, src, pred, tus)
    index = {(t.x, t.y): i for i, t in enumerate(tus)}
    for j, t in enumerate(sub):
        i = index[(t.x, t.y)]
        assert got_eob[i] == want_eob[j]
        np.testing.assert_array_equal(got_q[i], want_q[j])
        np.testing.assert_array_equal(got_rec.plane(0)[t.y:t.y + 32, t.x:t.x + 32], want_rec.plane(0)[t.y:t.y + 32, t.x:t.x + 32])


def test_me_picture_4k_geometry():
    """Open-loop ME at the configs[2] / [4] geometry (3840x2160: 2040 superblocks; ME always runs on the 8-bit luma planes):
    every MeSbResults field vs the oracle."""
    import gpu_runner as gr
    dist = ((1, 2, 3, 4), (1, 2, 3, 4))
    geos, src, refs = cm.make_me_case(W, H, 2, 2, seed=41)
    params = sb.preset8_me_params(W, H, 2, 2, dist, 1, 1)
    got = gr.run_gpu_me(params, src, refs)
    cm.assert_me_equal(got, cm.run_oracle_me(params, src, refs), params, "2160p gpu-vs-oracle")
