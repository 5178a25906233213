"""GPU parity tests for residual / transforms / quantisation (CUDA through the C ABI vs the CPU oracle, bit-exact).
Fixtures mirror test/FwdTxfm2dAsmTest.cc, InvTxfm2dAsmTest.cc, QuantAsmTest.cc, quantize_func_test.cc, ResidualTest.cc."""
import ctypes as C

import numpy as np
import pytest

import common as cm
import svtb200 as sb
from test_oracle_txfm import TX_H, TX_W, allowed_types, quant_tables, residual_block, scan_for

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("tx_size", range(19))
def test_fwd_txfm_dropins(tx_size):
    lib, orc = sb.load(), cm.oracle()
    w, h = TX_W[tx_size], TX_H[tx_size]
    f = getattr(lib, f"svt_av1_fwd_txfm2d_{w}x{h}_cuda")
    rng = np.random.default_rng(tx_size)
    for bd in (8, 10):
        for tx_type in allowed_types(tx_size):
            for mode in ("rand", "max") if tx_type else ("rand", "rand", "max", "min"):
                res = residual_block(rng, w, h, bd, mode)
                want, got = np.zeros(w * h, np.int32), np.zeros(w * h, np.int32)
                orc.orc_fwd_txfm2d(cm.ptr(res), cm.ptr(want), C.c_uint32(w + 5), tx_type, tx_size, bd)
                f(cm.ptr(res), cm.ptr(got), C.c_uint32(w + 5), tx_type, C.c_uint8(bd))
                np.testing.assert_array_equal(got, want, err_msg=f"{w}x{h} type {tx_type} bd {bd} {mode}")


@pytest.mark.parametrize("tx_size", [4, 11, 12, 17, 18])
def test_handle_transform64_dropins(tx_size):
    lib, orc = sb.load(), cm.oracle()
    w, h = TX_W[tx_size], TX_H[tx_size]
    orc.orc_handle_transform64.restype = C.c_uint64
    rng = np.random.default_rng(tx_size)
    a = rng.integers(-(1 << 20), 1 << 20, w * h).astype(np.int32)
    b = a.copy()
    ea = getattr(lib, f"svt_handle_transform{w}x{h}_cuda")(cm.ptr(a))
    eb = orc.orc_handle_transform64(cm.ptr(b), tx_size)
    assert ea == eb
    np.testing.assert_array_equal(a, b)


@pytest.mark.parametrize("tx_size", range(19))
def test_inv_txfm_dropins(tx_size):
    lib, orc = sb.load(), cm.oracle()
    w, h = TX_W[tx_size], TX_H[tx_size]
    iw, ih = min(w, 32), min(h, 32)
    f = getattr(lib, f"svt_av1_inv_txfm2d_add_{w}x{h}_cuda")
    rng = np.random.default_rng(100 + tx_size)
    for bd in (8, 10):
        for tx_type in allowed_types(tx_size):
            for mode in ("fwd", "big"):
                if mode == "fwd":
                    res = residual_block(rng, w, h, bd, "rand")
                    full = np.zeros(w * h, np.int32)
                    orc.orc_fwd_txfm2d(cm.ptr(res), cm.ptr(full), C.c_uint32(w + 5), tx_type, tx_size, bd)
                    coeff = np.ascontiguousarray(full.reshape(h, w)[:ih, :iw]).reshape(-1)
                else:
                    coeff = rng.integers(-(1 << (bd + 9)), 1 << (bd + 9), iw * ih).astype(np.int32)
                pred = rng.integers(0, 1 << bd, (h, w + 3)).astype(np.uint16)
                want, got = np.zeros((h, w + 7), np.uint16), np.zeros((h, w + 7), np.uint16)
                orc.orc_inv_txfm2d_add(cm.ptr(coeff), cm.ptr(pred), w + 3, cm.ptr(want), w + 7, tx_type, tx_size, bd)
                if w == h:
                    f(cm.ptr(coeff), cm.ptr(pred), w + 3, cm.ptr(got), w + 7, tx_type, bd)
                elif (w, h) in ((4, 8), (8, 4), (4, 16), (16, 4)):
                    f(cm.ptr(coeff), cm.ptr(pred), w + 3, cm.ptr(got), w + 7, tx_type, tx_size, bd)
                else:
                    f(cm.ptr(coeff), cm.ptr(pred), w + 3, cm.ptr(got), w + 7, tx_type, tx_size, w * h, bd)
                np.testing.assert_array_equal(got, want, err_msg=f"{w}x{h} type {tx_type} bd {bd} {mode}")


def test_quantizer_dropins():
    lib, orc = sb.load(), cm.oracle()
    rng = np.random.default_rng(5)
    variants = [("svt_aom_quantize_b_cuda", "b", 0), ("svt_aom_highbd_quantize_b_cuda", "b", 1),
                ("svt_av1_quantize_fp_cuda", "fp0", 0), ("svt_av1_quantize_fp_32x32_cuda", "fp1", 0),
                ("svt_av1_quantize_fp_64x64_cuda", "fp2", 0), ("svt_av1_highbd_quantize_fp_cuda", "fph", 1)]
    for name, kind, hbd in variants:
        f = getattr(lib, name)
        for n, ls in ((16, 0), (64, 0), (256, 0), (1024, 1), (1024, 2)):
            if kind.startswith("fp") and kind != "fph":
                ls = int(kind[-1])
            for mode in ("rand", "zero", "big", "small"):
                bd = 10 if hbd else 8
                zbin, rnd, quant, shift, dq = quant_tables(rng, bd, 0)
                amp = {"rand": 1 << (bd + 5), "zero": 1, "big": 1 << 20, "small": 40}[mode]
                coeff = rng.integers(-amp, amp, n).astype(np.int32) if mode != "zero" else np.zeros(n, np.int32)
                scan = scan_for(n, rng)
                q0, d0, e0 = np.full(n, 7, np.int32), np.full(n, 7, np.int32), C.c_uint16(9)
                q1, d1, e1 = np.full(n, 7, np.int32), np.full(n, 7, np.int32), C.c_uint16(9)
                base = [cm.ptr(coeff), C.c_ssize_t(n), cm.ptr(zbin), cm.ptr(rnd), cm.ptr(quant), cm.ptr(shift), cm.ptr(q0),
                        cm.ptr(d0), cm.ptr(dq), C.byref(e0), cm.ptr(scan), cm.ptr(scan)]
                if kind == "b":
                    f(*base, None, None, ls)
                    orc.orc_quantize_b(cm.ptr(coeff), C.c_ssize_t(n), cm.ptr(zbin), cm.ptr(rnd), cm.ptr(quant), cm.ptr(shift),
                                       cm.ptr(q1), cm.ptr(d1), cm.ptr(dq), C.byref(e1), cm.ptr(scan), None, None, ls, hbd)
                else:
                    if kind == "fph":
                        f(*base, C.c_int16(ls))
                    else:
                        f(*base)
                    orc.orc_quantize_fp(cm.ptr(coeff), C.c_ssize_t(n), cm.ptr(rnd), cm.ptr(quant), cm.ptr(q1), cm.ptr(d1),
                                        cm.ptr(dq), C.byref(e1), cm.ptr(scan), ls, hbd)
                np.testing.assert_array_equal(q0, q1, err_msg=f"{name} n={n} {mode}")
                np.testing.assert_array_equal(d0, d1, err_msg=f"{name} n={n} {mode}")
                assert e0.value == e1.value, (name, n, mode)


def test_residual_dropins():
    lib, orc = sb.load(), cm.oracle()
    rng = np.random.default_rng(3)
    for (w, h) in ((4, 4), (8, 8), (16, 16), (32, 32), (64, 64), (16, 4), (64, 16), (8, 32)):
        for hbd in (0, 1):
            dt, mx = (np.uint16, 1024) if hbd else (np.uint8, 256)
            a = rng.integers(0, mx, (h, w + 4)).astype(dt)
            b = rng.integers(0, mx, (h, w + 6)).astype(dt)
            r0, r1 = np.zeros((h, w + 2), np.int16), np.zeros((h, w + 2), np.int16)
            (lib.svt_residual_kernel16bit_cuda if hbd else lib.svt_residual_kernel8bit_cuda)(
                cm.ptr(a), C.c_uint32(w + 4), cm.ptr(b), C.c_uint32(w + 6), cm.ptr(r0), C.c_uint32(w + 2), C.c_uint32(w), C.c_uint32(h))
            orc.orc_residual(cm.ptr(a), C.c_uint32(w + 4), cm.ptr(b), C.c_uint32(w + 6), cm.ptr(r1), C.c_uint32(w + 2),
                             C.c_uint32(w), C.c_uint32(h), hbd)
            np.testing.assert_array_equal(r0, r1)


def quant_plane(rng, bd):
    q = sb.QuantPlane()
    zbin, rnd, quant, shift, dq = quant_tables(rng, bd, 0)
    for i in range(2):
        q.zbin[i], q.round[i], q.quant[i], q.quant_shift[i], q.dequant[i] = int(zbin[i]), int(rnd[i]), int(quant[i]), int(shift[i]), int(dq[i])
        q.round_fp[i], q.quant_fp[i] = int(rnd[i]) // 2 + 1, int(rng.integers(1, 1 << 14))
    return q


def oracle_encode_tus(params, src, pred, tus, use_fp):
    """Per-TU chain of oracle calls = the reference's av1_encode_loop body."""
    orc = cm.oracle()
    ts = params.tx_size
    w, h = TX_W[ts], TX_H[ts]
    iw, ih = min(w, 32), min(h, 32)
    n = iw * ih
    hbd = 1 if src.bd > 8 else 0
    recon = pred.copy()
    qco = np.zeros((len(tus), n), np.int32)
    eobs = np.zeros(len(tus), np.uint16)
    log_scale = (1 if w * h > 256 else 0) + (1 if w * h > 1024 else 0)
    lib = sb.load()
    for i, tu in enumerate(tus):
        sp, pp, rp = src.plane(tu.plane), pred.plane(tu.plane), recon.plane(tu.plane)
        sblk = np.ascontiguousarray(sp[tu.y:tu.y + h, tu.x:tu.x + w])
        pblk = np.ascontiguousarray(pp[tu.y:tu.y + h, tu.x:tu.x + w])
        res = np.zeros((h, w), np.int16)
        orc.orc_residual(cm.ptr(sblk), C.c_uint32(w), cm.ptr(pblk), C.c_uint32(w), cm.ptr(res), C.c_uint32(w), C.c_uint32(w), C.c_uint32(h), hbd)
        coeff = np.zeros(w * h, np.int32)
        orc.orc_fwd_txfm2d(cm.ptr(res), cm.ptr(coeff), C.c_uint32(w), tu.tx_type, ts, src.bd)
        if max(w, h) == 64:
            orc.orc_handle_transform64(cm.ptr(coeff), ts)
        scan = np.zeros(1024, np.int16)
        if cm.have_ref():  # the oracle side takes its scan order from the REFERENCE's av1_scan_orders, not from the product
            assert cm.refh().refh_get_scan(ts, tu.tx_type, cm.ptr(scan)) == n
            mine = np.zeros(1024, np.int16)
            assert lib.svt_b200_get_scan(ts, tu.tx_type, cm.ptr(mine)) == n
            np.testing.assert_array_equal(mine[:n], scan[:n], err_msg="svt_b200_get_scan differs from av1_scan_orders")
        else:
            assert lib.svt_b200_get_scan(ts, tu.tx_type, cm.ptr(scan)) == n
        q, dq, eob = np.zeros(n, np.int32), np.zeros(n, np.int32), C.c_uint16(0)
        qp = params.q[tu.plane]
        if use_fp:
            orc.orc_quantize_fp(cm.ptr(coeff), C.c_ssize_t(n), qp.round_fp, qp.quant_fp, cm.ptr(q), cm.ptr(dq), qp.dequant,
                                C.byref(eob), cm.ptr(scan), log_scale, hbd)
        else:
            orc.orc_quantize_b(cm.ptr(coeff), C.c_ssize_t(n), qp.zbin, qp.round, qp.quant, qp.quant_shift, cm.ptr(q), cm.ptr(dq),
                               qp.dequant, C.byref(eob), cm.ptr(scan), None, None, log_scale, hbd)
        qco[i], eobs[i] = q, eob.value
        p16 = pblk.astype(np.uint16)
        r16 = np.zeros((h, w), np.uint16)
        orc.orc_inv_txfm2d_add(cm.ptr(dq), cm.ptr(p16), w, cm.ptr(r16), w, tu.tx_type, ts, src.bd)
        rp[tu.y:tu.y + h, tu.x:tu.x + w] = r16.astype(rp.dtype)
    return recon, qco, eobs


def make_tus(rng, ts, w_pic, h_pic, planes=(0, 1, 2), limit=None):
    w, h = TX_W[ts], TX_H[ts]
    types = allowed_types(ts)
    tus = []
    for pl in planes:
        pw, ph = (w_pic, h_pic) if pl == 0 else ((w_pic + 1) // 2, (h_pic + 1) // 2)
        for y in range(0, ph - h + 1, h):
            for x in range(0, pw - w + 1, w):
                tus.append(sb.Tu(x, y, pl, int(rng.choice(types))))
    if limit and len(tus) > limit:
        idx = rng.choice(len(tus), limit, replace=False)
        tus = [tus[i] for i in sorted(idx)]
    return tus


@pytest.mark.parametrize("case", [(ts, bd, fp) for ts in range(19) for bd, fp in ((8, 0), (10, 1))] + [(2, 8, 1), (3, 10, 0)])
def test_encode_tus_vs_oracle(case):
    import gpu_runner as gr
    ts, bd, use_fp = case
    rng = np.random.default_rng(1000 + ts)
    W, H = 192, 128
    src = cm.synth_yuv(W, H, 1, 7, bd)
    pred = cm.degrade(src, 11, amp=14)
    p = sb.EncodeParams()
    p.tx_size, p.use_fp = ts, use_fp
    for i in range(3):
        p.q[i] = quant_plane(rng, bd)
    tus = make_tus(rng, ts, W, H, limit=160)
    want_rec, want_q, want_eob = oracle_encode_tus(p, src, pred, tus, use_fp)
    got_rec, got_q, got_eob, got_cul = gr.run_gpu_encode_tus(p, src, pred, tus, with_cul=True)
    np.testing.assert_array_equal(got_eob, want_eob)
    np.testing.assert_array_equal(got_q, want_q)
    # cul_level of av1_quantize_inv_quantize (EbFullLoop.c:1596-1608) from the ORACLE's levels: min(63, sum |q|), then
    # set_dc_sign (|= 64 for a negative DC level, += 128 for a positive one)
    want_cul = np.minimum(63, np.abs(want_q.astype(np.int64)).sum(axis=1))
    want_cul = want_cul + np.where(want_q[:, 0] < 0, 64, np.where(want_q[:, 0] > 0, 128, 0))
    np.testing.assert_array_equal(got_cul, want_cul)
    for i in range(3):
        np.testing.assert_array_equal(got_rec.plane(i), want_rec.plane(i), err_msg=f"plane {i}")


def test_encode_tus_1080p_16x16_and_roundtrip_property():
    """Full BASELINE geometry: every 16x16 luma + 8x8 chroma TU of a 1080p frame; plus the size-independent
    property that with a unit quantiser step the reconstruction error stays within the transform's rounding."""
    import gpu_runner as gr
    rng = np.random.default_rng(77)
    W, H = 1920, 1080
    src = cm.synth_yuv(W, H, 2, 3, 8)
    pred = cm.degrade(src, 5, amp=10)
    p = sb.EncodeParams()
    p.tx_size, p.use_fp = 2, 0
    for i in range(3):
        p.q[i] = quant_plane(rng, 8)
    tus = make_tus(rng, 2, W, H, planes=(0,))
    sub = [tus[i] for i in rng.choice(len(tus), 400, replace=False)]
    want_rec, want_q, want_eob = oracle_encode_tus(p, src, pred, sub, 0)
    got_rec, got_q, got_eob = gr.run_gpu_encode_tus(p, src, pred, tus)
    index = {(t.x, t.y): i for i, t in enumerate(tus)}
    for j, t in enumerate(sub):
        i = index[(t.x, t.y)]
        assert got_eob[i] == want_eob[j]
        np.testing.assert_array_equal(got_q[i], want_q[j])
        np.testing.assert_array_equal(got_rec.plane(0)[t.y:t.y + 16, t.x:t.x + 16], want_rec.plane(0)[t.y:t.y + 16, t.x:t.x + 16])
    # idempotence of the launch
    again_rec, again_q, again_eob = gr.run_gpu_encode_tus(p, src, pred, tus)
    np.testing.assert_array_equal(again_q, got_q)
    np.testing.assert_array_equal(again_rec.plane(0), got_rec.plane(0))


def oracle_encode_tus_ex(qsets, src, pred, tus, use_fp):
    """The chain of oracle_encode_tus for units of mixed size / shape / quantiser set: orc_estimate_transform is the
    restatement of the reference's shape dispatcher (pinned against av1_estimate_transform in test_oracle_encode.py)."""
    orc = cm.oracle()
    orc.orc_estimate_transform.restype = C.c_uint64
    lib = sb.load()
    hbd = 1 if src.bd > 8 else 0
    recon = pred.copy()
    levels, eobs = [], np.zeros(len(tus), np.uint16)
    for i, tu in enumerate(tus):
        ts = tu.tx_size
        w, h = TX_W[ts], TX_H[ts]
        n = min(w, 32) * min(h, 32)
        log_scale = (1 if w * h > 256 else 0) + (1 if w * h > 1024 else 0)
        sp, pp, rp = src.plane(tu.plane), pred.plane(tu.plane), recon.plane(tu.plane)
        sblk = np.ascontiguousarray(sp[tu.y:tu.y + h, tu.x:tu.x + w])
        pblk = np.ascontiguousarray(pp[tu.y:tu.y + h, tu.x:tu.x + w])
        res = np.zeros((h, w), np.int16)
        orc.orc_residual(cm.ptr(sblk), C.c_uint32(w), cm.ptr(pblk), C.c_uint32(w), cm.ptr(res), C.c_uint32(w), C.c_uint32(w), C.c_uint32(h), hbd)
        coeff = np.zeros(w * h, np.int32)
        orc.orc_estimate_transform(cm.ptr(res), C.c_uint32(w), cm.ptr(coeff), ts, src.bd, tu.tx_type, tu.pf_shape)
        scan = np.zeros(1024, np.int16)
        assert lib.svt_b200_get_scan(ts, tu.tx_type, cm.ptr(scan)) == n  # checked against av1_scan_orders in oracle_encode_tus
        q, dq, eob = np.zeros(n, np.int32), np.zeros(n, np.int32), C.c_uint16(0)
        qp = qsets[tu.qset][tu.plane]
        if use_fp:
            orc.orc_quantize_fp(cm.ptr(coeff), C.c_ssize_t(n), qp.round_fp, qp.quant_fp, cm.ptr(q), cm.ptr(dq), qp.dequant,
                                C.byref(eob), cm.ptr(scan), log_scale, hbd)
        else:
            orc.orc_quantize_b(cm.ptr(coeff), C.c_ssize_t(n), qp.zbin, qp.round, qp.quant, qp.quant_shift, cm.ptr(q), cm.ptr(dq),
                               qp.dequant, C.byref(eob), cm.ptr(scan), None, None, log_scale, hbd)
        levels.append(q)
        eobs[i] = eob.value
        p16 = pblk.astype(np.uint16)
        r16 = np.zeros((h, w), np.uint16)
        orc.orc_inv_txfm2d_add(cm.ptr(dq), cm.ptr(p16), w, cm.ptr(r16), w, tu.tx_type, ts, src.bd)
        rp[tu.y:tu.y + h, tu.x:tu.x + w] = r16.astype(rp.dtype)
    return recon, levels, eobs


def mixed_tus(rng, w_pic, h_pic, n_qsets, shapes=(0, 0, 1, 2, 3)):
    """Non-overlapping units of mixed sizes: every 64x64 cell of a plane is tiled with one randomly chosen size."""
    tus = []
    for pl in range(3):
        pw, ph = (w_pic, h_pic) if pl == 0 else ((w_pic + 1) // 2, (h_pic + 1) // 2)
        for cy in range(0, ph - 63, 64):
            for cx in range(0, pw - 63, 64):
                ts = int(rng.integers(0, 19))
                w, h = TX_W[ts], TX_H[ts]
                types = allowed_types(ts)
                cell = [(cx + x, cy + y) for y in range(0, 64, h) for x in range(0, 64, w)]
                for k in rng.choice(len(cell), min(len(cell), 6), replace=False):
                    x, y = cell[int(k)]
                    tus.append(sb.TuEx(x, y, pl, int(rng.choice(types)), ts, int(rng.choice(shapes)), int(rng.integers(0, n_qsets)), 0))
    order = rng.permutation(len(tus))  # the input order is NOT sorted by size: the entry buckets it
    return [tus[int(i)] for i in order]


@pytest.mark.parametrize("bd,use_fp", [(8, 0), (10, 1), (10, 0)])
def test_encode_tus_ex_mixed_sizes_shapes_qsets_vs_oracle(bd, use_fp):
    """the fused entry with the generality of the reference's dispatcher - units of all 19 sizes, the four
    coefficient shapes and several quantiser sets in ONE call, results in input order."""
    import gpu_runner as gr
    rng = np.random.default_rng(4242 + bd + use_fp)
    W, H = 320, 192
    src = cm.synth_yuv(W, H, 1, 9, bd)
    pred = cm.degrade(src, 13, amp=18)
    qsets = [[quant_plane(rng, bd) for _ in range(3)] for _ in range(3)]
    tus = mixed_tus(rng, W, H, len(qsets))
    assert len({t.tx_size for t in tus}) >= 10 and {t.pf_shape for t in tus} == {0, 1, 2, 3}
    want_rec, want_q, want_eob = oracle_encode_tus_ex(qsets, src, pred, tus, use_fp)
    got_rec, got_q, got_eob, got_cul, tail = gr.run_gpu_encode_tus_ex(qsets, use_fp, src, pred, tus)
    np.testing.assert_array_equal(got_eob, want_eob)
    for i, (g, wq) in enumerate(zip(got_q, want_q)):
        np.testing.assert_array_equal(g, wq, err_msg=f"unit {i}: size {tus[i].tx_size} shape {tus[i].pf_shape}")
    assert (tail == 77).all()  # nothing written past qcoeff_offsets[n]
    want_cul = np.array([min(63, int(np.abs(q.astype(np.int64)).sum())) + (64 if q[0] < 0 else 128 if q[0] > 0 else 0) for q in want_q])
    np.testing.assert_array_equal(got_cul, want_cul)
    for i in range(3):
        np.testing.assert_array_equal(got_rec.plane(i), want_rec.plane(i), err_msg=f"plane {i}")


def test_encode_tus_ex_default_shape_equals_the_single_size_entry():
    """Consistency of the two entries: one size, default shape, one quantiser set -> the same levels / eob / recon as
    svt_b200_encode_tus_cul; an empty list and out-of-range fields are handled (no launch / SVT_B200_ERR_ARG)."""
    import gpu_runner as gr
    rng = np.random.default_rng(5)
    W, H = 192, 128
    src = cm.synth_yuv(W, H, 1, 7, 8)
    pred = cm.degrade(src, 11, amp=14)
    p = sb.EncodeParams()
    p.tx_size, p.use_fp = 7, 0
    for i in range(3):
        p.q[i] = quant_plane(rng, 8)
    tus = make_tus(rng, 7, W, H, limit=100)
    rec1, q1, eob1, cul1 = gr.run_gpu_encode_tus(p, src, pred, tus, with_cul=True)
    ex = [sb.TuEx(t.x, t.y, t.plane, t.tx_type, 7, 0, 0, 0) for t in tus]
    rec2, q2, eob2, cul2, _ = gr.run_gpu_encode_tus_ex([[p.q[0], p.q[1], p.q[2]]], 0, src, pred, ex)
    np.testing.assert_array_equal(eob2, eob1)
    np.testing.assert_array_equal(cul2, cul1)
    np.testing.assert_array_equal(np.stack(q2), q1)
    for i in range(3):
        np.testing.assert_array_equal(rec2.plane(i), rec1.plane(i))
    rec3, q3, eob3, _, _ = gr.run_gpu_encode_tus_ex([[p.q[0], p.q[1], p.q[2]]], 0, src, pred, [])
    assert q3 == [] and all((rec3.plane(i) == pred.plane(i)).all() for i in range(3))
    with pytest.raises(Exception):
        gr.run_gpu_encode_tus_ex([[p.q[0], p.q[1], p.q[2]]], 0, src, pred, [sb.TuEx(0, 0, 0, 0, 19, 0, 0, 0)])
    with pytest.raises(Exception):
        gr.run_gpu_encode_tus_ex([[p.q[0], p.q[1], p.q[2]]], 0, src, pred, [sb.TuEx(0, 0, 0, 0, 2, 0, 1, 0)])


@pytest.mark.parametrize("tx_size", range(19))
def test_partial_frequency_dropins(tx_size):
    lib, orc = sb.load(), cm.oracle()
    w, h = TX_W[tx_size], TX_H[tx_size]
    rng = np.random.default_rng(70 + tx_size)
    for shape, sh in (("N2", 1), ("N4", 2)):
        f = getattr(lib, f"svt_av1_fwd_txfm2d_{w}x{h}_{shape}_cuda")
        for tx_type in allowed_types(tx_size)[:6]:
            res = residual_block(rng, w, h, 10, "rand")
            want, got = np.zeros(w * h, np.int32), np.full(w * h, 3, np.int32)
            orc.orc_fwd_txfm2d_pf(cm.ptr(res), cm.ptr(want), C.c_uint32(w + 5), tx_type, tx_size, 10, sh)
            f(cm.ptr(res), cm.ptr(got), C.c_uint32(w + 5), tx_type, C.c_uint8(10))
            np.testing.assert_array_equal(got, want)


@pytest.mark.parametrize("tx_size,tx_class", [(0, 0), (1, 0), (2, 0), (2, 1), (3, 2), (4, 0), (9, 0), (12, 0)])
def test_pack_levels(tx_size, tx_class):
    """svt_b200_pack_levels: the eob levels of every TU in scan order + exclusive offsets, vs numpy."""
    import torch
    lib = sb.load()
    rng = np.random.default_rng(300 + tx_size)
    n = min(TX_W[tx_size], 32) * min(TX_H[tx_size], 32)
    scan = np.zeros(1024, np.int16)
    assert lib.svt_b200_get_scan(tx_size, {0: 0, 1: 10, 2: 11}[tx_class], cm.ptr(scan)) >= 0
    n_tus = 2500
    q = np.zeros((n_tus, n), np.int32)
    eob = np.zeros(n_tus, np.uint16)
    for b in range(n_tus):
        e = int(rng.integers(0, n + 1)) if b % 3 else int(rng.integers(0, 4))
        eob[b] = e
        lv = rng.integers(-300, 301, e)
        if e:
            lv[-1] = lv[-1] or 1
        q[b, scan[:e]] = lv
    dq, de = torch.from_numpy(q).cuda(), torch.from_numpy(eob.view(np.int16)).cuda()
    packed = torch.zeros(n_tus * n, dtype=torch.int32, device="cuda")
    off = torch.zeros(n_tus + 1, dtype=torch.int32, device="cuda")
    tot = torch.zeros(1, dtype=torch.int32, device="cuda")
    sb.check(lib.svt_b200_pack_levels(tx_size, tx_class, C.c_void_p(dq.data_ptr()), C.c_void_p(de.data_ptr()), n_tus,
                                      C.c_void_p(packed.data_ptr()), C.c_void_p(off.data_ptr()), C.c_void_p(tot.data_ptr()), None), lib)
    torch.cuda.synchronize()
    want_off = np.concatenate([[0], np.cumsum(eob.astype(np.int64))])
    np.testing.assert_array_equal(off.cpu().numpy(), want_off)
    assert int(tot.item()) == want_off[-1]
    got = packed.cpu().numpy()
    for b in range(0, n_tus, 7):
        np.testing.assert_array_equal(got[want_off[b]:want_off[b + 1]], q[b, scan[:eob[b]]])


def test_pack_levels_at_chained():
    """svt_b200_pack_levels_at: three transform sizes chained into ONE packed stream (each call starts where the previous
    one ended, the start read on the device), plus the n_tus == 0 pass-through."""
    import torch
    lib = sb.load()
    rng = np.random.default_rng(911)
    packed = torch.zeros(1 << 21, dtype=torch.int32, device="cuda")
    tot = torch.zeros(5, dtype=torch.int32, device="cuda")
    want, base, offs_want, offs_got = [], 0, [], []
    sizes = [(2, 700), (1, 900), (4, 0), (0, 1500)]
    for k, (tx_size, n_tus) in enumerate(sizes):
        n = min(TX_W[tx_size], 32) * min(TX_H[tx_size], 32)
        scan = np.zeros(1024, np.int16)
        assert lib.svt_b200_get_scan(tx_size, 0, cm.ptr(scan)) >= 0
        q = np.zeros((max(n_tus, 1), n), np.int32)
        eob = np.zeros(max(n_tus, 1), np.uint16)
        for b in range(n_tus):
            e = int(rng.integers(0, n + 1)) if b % 4 else 0
            eob[b] = e
            q[b, scan[:e]] = rng.integers(1, 200, e) * rng.choice([-1, 1], e)
            want.append(q[b, scan[:e]])
        dq, de = torch.from_numpy(q).cuda(), torch.from_numpy(eob.view(np.int16)).cuda()
        off = torch.zeros(n_tus + 1, dtype=torch.int32, device="cuda")
        base_p = C.c_void_p(tot.data_ptr() + 4 * (k - 1)) if k else None
        sb.check(lib.svt_b200_pack_levels_at(tx_size, 0, C.c_void_p(dq.data_ptr()), C.c_void_p(de.data_ptr()), n_tus,
                                             C.c_void_p(packed.data_ptr()), C.c_void_p(off.data_ptr()),
                                             C.c_void_p(tot.data_ptr() + 4 * k), base_p, None), lib)
        torch.cuda.synchronize()
        offs_got.append(off.cpu().numpy())
        offs_want.append(base + np.concatenate([[0], np.cumsum(eob[:n_tus].astype(np.int64))]))
        base = int(offs_want[-1][-1])
    for g, w in zip(offs_got, offs_want):
        np.testing.assert_array_equal(g, w)
    np.testing.assert_array_equal(tot.cpu().numpy()[:4], [o[-1] for o in offs_want])
    np.testing.assert_array_equal(packed.cpu().numpy()[:base], np.concatenate(want))


class TxfmParam(C.Structure):  # EbDefinitions.h:779-791 (tx_type / tx_size / tx_set_type are 1-byte packed enums)
    _fields_ = [("tx_type", C.c_uint8), ("tx_size", C.c_uint8), ("lossless", C.c_int32), ("bd", C.c_int32), ("is_hbd", C.c_int32),
                ("tx_set_type", C.c_uint8), ("eob", C.c_int32)]


@pytest.mark.parametrize("tx_size", [0, 1, 2, 3, 4, 5, 8, 11, 13, 18])
def test_inv_txfm_add_lowbd_dropin(tx_size):
    """svt_av1_inv_txfm_add (8-bit wrapper): vs the oracle's inverse on the widened prediction, and vs the reference's own
    svt_av1_inv_txfm_add_c when oracle/_ref is present."""
    lib, orc = sb.load(), cm.oracle()
    w, h = TX_W[tx_size], TX_H[tx_size]
    rng = np.random.default_rng(400 + tx_size)
    for tx_type in allowed_types(tx_size)[:4]:
        coeff = np.zeros(min(w, 32) * min(h, 32), np.int32)
        nz = rng.choice(coeff.size, min(12, coeff.size), replace=False)
        coeff[nz] = rng.integers(-900, 901, nz.size)
        pred = rng.integers(0, 256, (h, w + 5)).astype(np.uint8)
        got = np.zeros((h, w + 9), np.uint8)
        tp = TxfmParam(tx_type, tx_size, 0, 8, 0, 0, int(coeff.size))
        lib.svt_av1_inv_txfm_add_cuda(cm.ptr(coeff), cm.ptr(pred), pred.shape[1], cm.ptr(got), got.shape[1], C.byref(tp))
        p16 = np.ascontiguousarray(pred[:, :w].astype(np.uint16))
        want = np.zeros((h, w), np.uint16)
        orc.orc_inv_txfm2d_add(cm.ptr(coeff), cm.ptr(p16), w, cm.ptr(want), w, tx_type, tx_size, 8)
        np.testing.assert_array_equal(got[:, :w], want.astype(np.uint8))
        if cm.have_ref():
            ref = np.zeros((h, w + 9), np.uint8)
            cm.refh().svt_av1_inv_txfm_add_c(cm.ptr(coeff), cm.ptr(pred), pred.shape[1], cm.ptr(ref), ref.shape[1], C.byref(tp))
            np.testing.assert_array_equal(got[:, :w], ref[:, :w])
