"""Pins the transform / quantisation part of the oracle against the unmodified reference C path (oracle/_ref).
Fixtures follow test/FwdTxfm2dAsmTest.cc, test/InvTxfm2dAsmTest.cc, test/FwdTxfm1dTest.cc, test/QuantAsmTest.cc,
test/quantize_func_test.cc and test/ResidualTest.cc (seeded random + extreme inputs, all sizes x allowed types)."""
import ctypes as C

import numpy as np
import pytest

import common as cm

needs_ref = pytest.mark.skipif(not cm.have_ref(), reason="oracle/_ref not built")

TX_W = [4, 8, 16, 32, 64, 4, 8, 8, 16, 16, 32, 32, 64, 4, 16, 8, 32, 16, 64]
TX_H = [4, 8, 16, 32, 64, 8, 4, 16, 8, 32, 16, 64, 32, 16, 4, 32, 8, 64, 16]
SQUARE_NAMES = {0: "4x4", 1: "8x8", 2: "16x16", 3: "32x32", 4: "64x64"}


def allowed_types(tx_size):
    """is_txfm_allowed (test/FwdTxfm2dAsmTest.cc / AV1 spec): 64 -> DCT only, 32 -> DCT + IDTX, else all 16."""
    m = max(TX_W[tx_size], TX_H[tx_size])
    return [0] if m == 64 else [0, 9] if m == 32 else list(range(16))


def fwd_ref_name(tx_size):
    w, h = TX_W[tx_size], TX_H[tx_size]
    return f"svt_av1_transform_two_d_{w}x{h}_c" if w == h else f"svt_av1_fwd_txfm2d_{w}x{h}_c"


def test_tables_match_reference():
    if not cm.have_ref():
        pytest.skip("oracle/_ref not built")
    ref, orc = cm.ref(), cm.oracle()
    orc.orc_cospi_table.restype = C.POINTER(C.c_int32)
    orc.orc_sinpi_table.restype = C.POINTER(C.c_int32)
    cp = (C.c_int32 * 64 * 7).in_dll(ref, "eb_av1_cospi_arr_data")
    sp = (C.c_int32 * 5 * 7).in_dll(ref, "eb_av1_sinpi_arr_data")
    for bit in range(10, 14):
        assert [cp[bit - 10][i] for i in range(64)] == [orc.orc_cospi_table(bit)[i] for i in range(64)]
        assert [sp[bit - 10][i] for i in range(5)] == [orc.orc_sinpi_table(bit)[i] for i in range(5)]


@needs_ref
def test_1d_transforms_match_reference():
    ref, orc = cm.ref(), cm.oracle()
    rng = np.random.default_rng(0)
    sr = (C.c_int8 * 12)(*([18] * 12))

    def run_ref(name, x, bit):
        out = np.zeros_like(x)
        getattr(ref, name)(cm.ptr(x), cm.ptr(out), C.c_int8(bit), sr)
        return out

    for n in (4, 8, 16, 32, 64):
        for bit in (10, 11, 12, 13):
            for it in range(12):
                x = rng.integers(-(1 << 15), 1 << 15, n).astype(np.int32) if it else np.full(n, 32767, np.int32)
                y = x.copy()
                orc.orc_fdct(cm.ptr(y), 1, n, bit)
                np.testing.assert_array_equal(y, run_ref(f"svt_av1_fdct{n}_new", x, bit))
                y = x.copy()
                orc.orc_idct(cm.ptr(y), 1, n, bit, 18)
                np.testing.assert_array_equal(y, run_ref(f"svt_av1_idct{n}_new", x, bit))
                if n <= 16:
                    y = x.copy()
                    orc.orc_fadst(cm.ptr(y), 1, n, bit)
                    np.testing.assert_array_equal(y, run_ref(f"svt_av1_fadst{n}_new", x, bit))
                    y = x.copy()
                    orc.orc_iadst(cm.ptr(y), 1, n, bit, 18)
                    np.testing.assert_array_equal(y, run_ref(f"svt_av1_iadst{n}_new", x, bit))


def residual_block(rng, w, h, bd, mode):
    lim = (1 << bd) - 1
    if mode == "max":
        return np.full((h, w + 5), lim, np.int16)
    if mode == "min":
        return np.full((h, w + 5), -lim, np.int16)
    return rng.integers(-lim, lim + 1, (h, w + 5)).astype(np.int16)


@needs_ref
@pytest.mark.parametrize("tx_size", range(19))
def test_fwd_txfm2d_matches_reference(tx_size):
    ref, orc = cm.ref(), cm.oracle()
    w, h = TX_W[tx_size], TX_H[tx_size]
    rng = np.random.default_rng(tx_size)
    f = getattr(ref, fwd_ref_name(tx_size))
    for bd in (8, 10):
        for tx_type in allowed_types(tx_size):
            for mode in ("rand", "rand", "max", "min"):
                res = residual_block(rng, w, h, bd, mode)
                want, got = np.zeros(w * h, np.int32), np.zeros(w * h, np.int32)
                f(cm.ptr(res), cm.ptr(want), C.c_uint32(w + 5), tx_type, C.c_uint8(bd))
                orc.orc_fwd_txfm2d(cm.ptr(res), cm.ptr(got), C.c_uint32(w + 5), tx_type, tx_size, bd)
                np.testing.assert_array_equal(got, want, err_msg=f"{w}x{h} type {tx_type} bd {bd} {mode}")


@needs_ref
@pytest.mark.parametrize("tx_size", [4, 11, 12, 17, 18])
def test_handle_transform64_matches_reference(tx_size):
    ref, orc = cm.ref(), cm.oracle()
    w, h = TX_W[tx_size], TX_H[tx_size]
    rng = np.random.default_rng(tx_size)
    f = getattr(ref, f"svt_handle_transform{w}x{h}_c")
    f.restype = C.c_uint64
    orc.orc_handle_transform64.restype = C.c_uint64
    for _ in range(4):
        a = rng.integers(-(1 << 20), 1 << 20, w * h).astype(np.int32)
        b = a.copy()
        ea, eb = f(cm.ptr(a)), orc.orc_handle_transform64(cm.ptr(b), tx_size)
        assert ea == eb
        kw, kh = min(w, 32), min(h, 32)
        np.testing.assert_array_equal(a[: kw * kh], b[: kw * kh])


def call_ref_inv(ref, tx_size, coeff, pred, stride_r, rec, stride_w, tx_type, bd):
    w, h = TX_W[tx_size], TX_H[tx_size]
    f = getattr(ref, f"svt_av1_inv_txfm2d_add_{w}x{h}_c")
    if w == h:
        f(cm.ptr(coeff), cm.ptr(pred), stride_r, cm.ptr(rec), stride_w, tx_type, bd)
    elif (w, h) in ((4, 8), (8, 4), (4, 16), (16, 4)):
        f(cm.ptr(coeff), cm.ptr(pred), stride_r, cm.ptr(rec), stride_w, tx_type, tx_size, bd)
    else:
        f(cm.ptr(coeff), cm.ptr(pred), stride_r, cm.ptr(rec), stride_w, tx_type, tx_size, w * h, bd)


@needs_ref
@pytest.mark.parametrize("tx_size", range(19))
def test_inv_txfm2d_add_matches_reference(tx_size):
    ref, orc = cm.ref(), cm.oracle()
    w, h = TX_W[tx_size], TX_H[tx_size]
    iw, ih = min(w, 32), min(h, 32)
    rng = np.random.default_rng(100 + tx_size)
    for bd in (8, 10):
        for tx_type in allowed_types(tx_size):
            for mode in ("fwd", "fwd", "rand", "big"):
                if mode == "fwd":  # realistic coefficients: forward transform of a random residual
                    res = residual_block(rng, w, h, bd, "rand")
                    full = np.zeros(w * h, np.int32)
                    orc.orc_fwd_txfm2d(cm.ptr(res), cm.ptr(full), C.c_uint32(w + 5), tx_type, tx_size, bd)
                    coeff = np.ascontiguousarray(full.reshape(h, w)[:ih, :iw]).reshape(-1)
                elif mode == "rand":
                    coeff = rng.integers(-(1 << (bd + 3)), 1 << (bd + 3), iw * ih).astype(np.int32)
                else:
                    coeff = rng.integers(-(1 << (bd + 9)), 1 << (bd + 9), iw * ih).astype(np.int32)
                pred = rng.integers(0, 1 << bd, (h, w + 3)).astype(np.uint16)
                want, got = np.zeros((h, w + 7), np.uint16), np.zeros((h, w + 7), np.uint16)
                call_ref_inv(ref, tx_size, coeff, pred, w + 3, want, w + 7, tx_type, bd)
                orc.orc_inv_txfm2d_add(cm.ptr(coeff), cm.ptr(pred), w + 3, cm.ptr(got), w + 7, tx_type, tx_size, bd)
                np.testing.assert_array_equal(got, want, err_msg=f"{w}x{h} type {tx_type} bd {bd} {mode}")


def quant_tables(rng, bd, qidx_like):
    """zbin/round/quant/shift/dequant pairs of plausible magnitude (dc, ac), like test/QuantAsmTest.cc's setup from
    the encoder's Quants tables."""
    dq = np.array([int(rng.integers(4, 1 << (bd - 1))), int(rng.integers(4, 1 << (bd - 1)))], np.int16)
    zbin = ((dq.astype(np.int32) * int(rng.integers(64, 128)) + 64) >> 7).astype(np.int16)
    rnd = ((dq.astype(np.int32) * int(rng.integers(32, 64))) >> 7).astype(np.int16)
    quant = rng.integers(1, 1 << 15, 2).astype(np.int16)
    shift = rng.integers(1, 1 << 15, 2).astype(np.int16)
    return zbin, rnd, quant, shift, dq


def scan_for(n, rng):
    return rng.permutation(n).astype(np.int16)


@needs_ref
@pytest.mark.parametrize("hbd", [0, 1])
def test_quantize_b_matches_reference(hbd):
    ref, orc = cm.ref(), cm.oracle()
    name = "svt_aom_highbd_quantize_b" if hbd else "svt_aom_quantize_b"
    f = C.cast(C.c_void_p.in_dll(ref, name).value, C.CFUNCTYPE(None))
    rng = np.random.default_rng(5 + hbd)
    for n, log_scale in ((16, 0), (64, 0), (256, 0), (1024, 1), (1024, 2), (32, 0), (512, 1)):
        for mode in ("rand", "zero", "big", "small"):
            bd = 10 if hbd else 8
            zbin, rnd, quant, shift, dq = quant_tables(rng, bd, 0)
            amp = {"rand": 1 << (bd + 5), "zero": 1, "big": 1 << 22, "small": 40}[mode]
            coeff = rng.integers(-amp, amp, n).astype(np.int32) if mode != "zero" else np.zeros(n, np.int32)
            scan = scan_for(n, rng)
            outs = []
            for which in (0, 1):
                q, d, eob = np.full(n, 77, np.int32), np.full(n, 77, np.int32), C.c_uint16(9)
                if which == 0:
                    f(cm.ptr(coeff), C.c_ssize_t(n), cm.ptr(zbin), cm.ptr(rnd), cm.ptr(quant), cm.ptr(shift), cm.ptr(q),
                      cm.ptr(d), cm.ptr(dq), C.byref(eob), cm.ptr(scan), cm.ptr(scan), None, None, log_scale)
                else:
                    orc.orc_quantize_b(cm.ptr(coeff), C.c_ssize_t(n), cm.ptr(zbin), cm.ptr(rnd), cm.ptr(quant), cm.ptr(shift),
                                       cm.ptr(q), cm.ptr(d), cm.ptr(dq), C.byref(eob), cm.ptr(scan), None, None, log_scale, hbd)
                outs.append((q, d, eob.value))
            np.testing.assert_array_equal(outs[0][0], outs[1][0])
            np.testing.assert_array_equal(outs[0][1], outs[1][1])
            assert outs[0][2] == outs[1][2]


@needs_ref
@pytest.mark.parametrize("variant", ["fp", "fp_32x32", "fp_64x64", "highbd0", "highbd1", "highbd2"])
def test_quantize_fp_matches_reference(variant):
    ref, orc = cm.ref(), cm.oracle()
    hbd = variant.startswith("highbd")
    log_scale = {"fp": 0, "fp_32x32": 1, "fp_64x64": 2}.get(variant, int(variant[-1]) if hbd else 0)
    name = "svt_av1_highbd_quantize_fp" if hbd else "svt_av1_quantize_" + variant
    f = C.cast(C.c_void_p.in_dll(ref, name).value, C.CFUNCTYPE(None))
    rng = np.random.default_rng(9)
    for n in (16, 64, 256, 1024):
        for mode in ("rand", "zero", "big", "small"):
            bd = 10 if hbd else 8
            zbin, rnd, quant, shift, dq = quant_tables(rng, bd, 0)
            amp = {"rand": 1 << (bd + 5), "zero": 1, "big": 1 << 20, "small": 40}[mode]
            coeff = rng.integers(-amp, amp, n).astype(np.int32) if mode != "zero" else np.zeros(n, np.int32)
            scan = scan_for(n, rng)
            outs = []
            for which in (0, 1):
                q, d, eob = np.full(n, 77, np.int32), np.full(n, 77, np.int32), C.c_uint16(9)
                if which == 0:
                    args = [cm.ptr(coeff), C.c_ssize_t(n), cm.ptr(zbin), cm.ptr(rnd), cm.ptr(quant), cm.ptr(shift), cm.ptr(q),
                            cm.ptr(d), cm.ptr(dq), C.byref(eob), cm.ptr(scan), cm.ptr(scan)]
                    if hbd:
                        args.append(C.c_int16(log_scale))
                    f(*args)
                else:
                    orc.orc_quantize_fp(cm.ptr(coeff), C.c_ssize_t(n), cm.ptr(rnd), cm.ptr(quant), cm.ptr(q), cm.ptr(d),
                                        cm.ptr(dq), C.byref(eob), cm.ptr(scan), log_scale, int(hbd))
                outs.append((q, d, eob.value))
            np.testing.assert_array_equal(outs[0][0], outs[1][0])
            np.testing.assert_array_equal(outs[0][1], outs[1][1])
            assert outs[0][2] == outs[1][2]


@needs_ref
def test_residual_matches_reference():
    ref, orc = cm.ref(), cm.oracle()
    f8 = C.cast(C.c_void_p.in_dll(ref, "svt_residual_kernel8bit").value, C.CFUNCTYPE(None))
    f16 = C.cast(C.c_void_p.in_dll(ref, "svt_residual_kernel16bit").value, C.CFUNCTYPE(None))
    rng = np.random.default_rng(3)
    for (w, h) in ((4, 4), (8, 8), (16, 16), (32, 32), (64, 64), (16, 4), (64, 16), (8, 32)):
        for hbd in (0, 1):
            dt, mx = (np.uint16, 1024) if hbd else (np.uint8, 256)
            a = rng.integers(0, mx, (h, w + 4)).astype(dt)
            b = rng.integers(0, mx, (h, w + 6)).astype(dt)
            r0, r1 = np.zeros((h, w + 2), np.int16), np.zeros((h, w + 2), np.int16)
            (f16 if hbd else f8)(cm.ptr(a), C.c_uint32(w + 4), cm.ptr(b), C.c_uint32(w + 6), cm.ptr(r0), C.c_uint32(w + 2),
                                 C.c_uint32(w), C.c_uint32(h))
            orc.orc_residual(cm.ptr(a), C.c_uint32(w + 4), cm.ptr(b), C.c_uint32(w + 6), cm.ptr(r1), C.c_uint32(w + 2),
                             C.c_uint32(w), C.c_uint32(h), hbd)
            np.testing.assert_array_equal(r0, r1)


@needs_ref
@pytest.mark.parametrize("tx_size", range(19))
def test_partial_frequency_fwd_txfm_matches_reference(tx_size):
    """N2 / N4 shapes: the reference's *_N2_c / *_N4_c functions vs 'full transform + zeroing' (what its own
    test/FwdTxfm2dAsmTest.cc:257-283 asserts for the SIMD versions)."""
    ref, orc = cm.ref(), cm.oracle()
    w, h = TX_W[tx_size], TX_H[tx_size]
    rng = np.random.default_rng(50 + tx_size)
    for shape, sh in (("N2", 1), ("N4", 2)):
        name = (f"av1_transform_two_d_{w}x{h}_{shape}_c" if w == h else f"svt_av1_fwd_txfm2d_{w}x{h}_{shape}_c")
        f = getattr(ref, name)
        for bd in (8, 10):
            for tx_type in allowed_types(tx_size):
                res = residual_block(rng, w, h, bd, "rand")
                want, got = np.full(w * h, 5, np.int32), np.zeros(w * h, np.int32)
                f(cm.ptr(res), cm.ptr(want), C.c_uint32(w + 5), tx_type, C.c_uint8(bd))
                orc.orc_fwd_txfm2d_pf(cm.ptr(res), cm.ptr(got), C.c_uint32(w + 5), tx_type, tx_size, bd, sh)
                np.testing.assert_array_equal(got, want, err_msg=f"{name} type {tx_type} bd {bd}")
