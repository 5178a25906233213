"""Pins the loop-restoration kernels of the oracle against the unmodified reference C path (oracle/_ref).
Fixtures follow test/selfguided_filter_test.cc (:245-480) and test/wiener_convolve_test.cc (:488-515)."""
import ctypes as C

import numpy as np
import pytest

import common as cm

needs_ref = pytest.mark.skipif(not cm.have_ref(), reason="oracle/_ref not built")


class ConvolveParams(C.Structure):
    _fields_ = [("ref", C.c_int32), ("do_average", C.c_int32), ("dst", C.c_void_p), ("dst_stride", C.c_int32),
                ("round_0", C.c_int32), ("round_1", C.c_int32), ("plane", C.c_int32), ("is_compound", C.c_int32),
                ("use_jnt_comp_avg", C.c_int32), ("fwd_offset", C.c_int32), ("bck_offset", C.c_int32),
                ("use_dist_wtd_comp_avg", C.c_int32)]


def lr_image(rng, bd, mode, h=80, w=96):
    mx = (1 << bd) - 1
    if mode == "random":
        return rng.integers(0, mx + 1, (h, w))
    if mode == "flat":
        return np.clip(int(rng.integers(0, mx)) + rng.integers(-1, 2, (h, w)), 0, mx)
    if mode == "max":
        return np.full((h, w), mx)
    base = np.add.outer(np.arange(h) * 3, np.arange(w) * 2) % (mx + 1)
    return np.clip(base + rng.integers(-6, 7, (h, w)) * (1 << (bd - 8)), 0, mx)


def sgr_cases():
    rng = np.random.default_rng(1)
    for bd in (8, 10, 12):
        for mode in ("random", "flat", "smooth", "max"):
            for (w, h) in ((64, 64), (64, 32), (40, 56), (8, 8), (64, 8)):
                yield bd, mode, w, h, int(rng.integers(0, 16)), lr_image(rng, bd, mode)


@needs_ref
def test_selfguided_restoration_matches_reference():
    ref, orc = cm.ref(), cm.oracle()
    f = C.cast(C.c_void_p.in_dll(ref, "svt_av1_selfguided_restoration").value, C.CFUNCTYPE(None))
    for bd, mode, w, h, eps, img in sgr_cases():
        hbd = bd > 8
        a = img.astype(np.uint16 if hbd else np.uint8)
        stride = a.shape[1]
        p = a.ctypes.data + (8 * stride + 8) * a.itemsize
        outs = []
        for which in (0, 1):
            f0, f1 = np.full((h, w + 3), -7, np.int32), np.full((h, w + 3), -7, np.int32)
            if which == 0:
                f(C.c_void_p(p >> 1 if hbd else p), w, h, stride, cm.ptr(f0), cm.ptr(f1), w + 3, eps, bd, int(hbd))
            else:
                orc.orc_selfguided_restoration(C.c_void_p(p), int(hbd), w, h, stride, cm.ptr(f0), cm.ptr(f1), w + 3, eps, bd)
            outs.append((f0, f1))
        np.testing.assert_array_equal(outs[0][0], outs[1][0], err_msg=f"flt0 bd{bd} {mode} {w}x{h} eps{eps}")
        np.testing.assert_array_equal(outs[0][1], outs[1][1], err_msg=f"flt1 bd{bd} {mode} {w}x{h} eps{eps}")


@needs_ref
def test_apply_selfguided_restoration_matches_reference():
    ref, orc = cm.ref(), cm.oracle()
    f = C.cast(C.c_void_p.in_dll(ref, "svt_apply_selfguided_restoration").value, C.CFUNCTYPE(None))
    rng = np.random.default_rng(2)
    tmp = np.zeros(2 * 161 * 161 * 4, np.int32)
    for bd, mode, w, h, eps, img in sgr_cases():
        hbd = bd > 8
        a = img.astype(np.uint16 if hbd else np.uint8)
        stride = a.shape[1]
        p = a.ctypes.data + (8 * stride + 8) * a.itemsize
        xqd = (C.c_int32 * 2)(int(rng.integers(-96, 32)), int(rng.integers(-32, 96)))
        d0, d1 = np.zeros((h, w + 5), a.dtype), np.zeros((h, w + 5), a.dtype)
        f(C.c_void_p(p >> 1 if hbd else p), w, h, stride, eps, xqd, C.c_void_p(d0.ctypes.data >> 1 if hbd else d0.ctypes.data),
          w + 5, cm.ptr(tmp), bd, int(hbd))
        orc.orc_apply_selfguided_restoration(C.c_void_p(p), int(hbd), w, h, stride, eps, xqd, cm.ptr(d1), w + 5, bd)
        np.testing.assert_array_equal(d0, d1, err_msg=f"bd{bd} {mode} {w}x{h} eps{eps}")


def wiener_filter(rng):
    """7-tap symmetric kernel with the reference's constraints (taps sum to 0 around the implicit centre 128)."""
    f = np.zeros(8, np.int16)
    f[0] = f[6] = rng.integers(-5, 11)
    f[1] = f[5] = rng.integers(-23, 9)
    f[2] = f[4] = rng.integers(-17, 47)
    f[3] = -2 * (int(f[0]) + int(f[1]) + int(f[2]))
    return f


def aligned_filter(f):
    """The reference recovers the kernel from a 256-byte aligned table (get_filter_base): place it at an aligned base."""
    raw = np.zeros(8 * 16 + 256, np.int16)
    off = (-raw.ctypes.data) % 256 // 2
    raw[off:off + 8] = f
    return raw, raw.ctypes.data + off * 2


@needs_ref
def test_wiener_convolve_matches_reference():
    ref, orc = cm.ref(), cm.oracle()
    f8 = C.cast(C.c_void_p.in_dll(ref, "svt_av1_wiener_convolve_add_src").value, C.CFUNCTYPE(None))
    f16 = C.cast(C.c_void_p.in_dll(ref, "svt_av1_highbd_wiener_convolve_add_src").value, C.CFUNCTYPE(None))
    rng = np.random.default_rng(3)
    for bd in (8, 10, 12):
        for mode in ("random", "smooth", "max", "flat"):
            for (w, h) in ((64, 64), (32, 16), (16, 64), (8, 8), (64, 24)):
                hbd = bd > 8
                a = lr_image(rng, bd, mode).astype(np.uint16 if hbd else np.uint8)
                stride = a.shape[1]
                p = a.ctypes.data + (8 * stride + 8) * a.itemsize
                fx, fy = wiener_filter(rng), wiener_filter(rng)
                keep_x, px = aligned_filter(fx)
                keep_y, py = aligned_filter(fy)
                cp = ConvolveParams()
                cp.round_0 = 5 if bd == 12 else 3
                cp.round_1 = 14 - cp.round_0
                d0, d1 = np.zeros((h, w + 5), a.dtype), np.zeros((h, w + 5), a.dtype)
                if hbd:
                    f16(C.c_void_p(p >> 1), C.c_ssize_t(stride), C.c_void_p(d0.ctypes.data >> 1), C.c_ssize_t(w + 5), C.c_void_p(px),
                        C.c_void_p(py), w, h, C.byref(cp), bd)
                else:
                    f8(C.c_void_p(p), C.c_ssize_t(stride), cm.ptr(d0), C.c_ssize_t(w + 5), C.c_void_p(px), C.c_void_p(py), w, h, C.byref(cp))
                orc.orc_wiener_convolve_add_src(C.c_void_p(p), int(hbd), C.c_ssize_t(stride), cm.ptr(d1), C.c_ssize_t(w + 5), cm.ptr(fx),
                                                cm.ptr(fy), w, h, cp.round_0, cp.round_1, bd)
                np.testing.assert_array_equal(d0, d1, err_msg=f"bd{bd} {mode} {w}x{h}")
