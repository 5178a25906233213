"""Pins tests/misc_oracle.py (numpy) against the reference C functions (oracle/_ref)."""
import ctypes as C

import numpy as np
import pytest

import common as cm
import misc_oracle as mo

needs_ref = pytest.mark.skipif(not cm.have_ref(), reason="oracle/_ref not built")
SAD_SIZES = [(128, 128), (128, 64), (64, 128), (64, 64), (64, 32), (64, 16), (32, 64), (32, 32), (32, 16), (32, 8), (16, 64),
             (16, 32), (16, 16), (16, 8), (16, 4), (8, 32), (8, 16), (8, 8), (8, 4), (4, 16), (4, 8), (4, 4)]


def fn(name, restype=None):
    return C.cast(C.c_void_p.in_dll(cm.ref(), name).value, C.CFUNCTYPE(restype))


@needs_ref
def test_misc_reductions_match_reference():
    rng = np.random.default_rng(1)
    for (w, h) in SAD_SIZES:
        a = rng.integers(0, 256, (h, w + 7), dtype=np.uint8)
        b = rng.integers(0, 256, (h, w + 3), dtype=np.uint8)
        f = fn(f"svt_aom_sad{w}x{h}", C.c_uint32)
        assert f(cm.ptr(a), w + 7, cm.ptr(b), w + 3) == mo.sad(a[:, :w], b[:, :w])
        f4 = fn(f"svt_aom_sad{w}x{h}x4d")
        refs = [rng.integers(0, 256, (h, w + 3), dtype=np.uint8) for _ in range(4)]
        arr = (C.c_void_p * 4)(*[r.ctypes.data for r in refs])
        out = np.zeros(4, np.uint32)
        f4(cm.ptr(a), w + 7, arr, w + 3, cm.ptr(out))
        assert out.tolist() == [mo.sad(a[:, :w], r[:, :w]) for r in refs]
    for (w, h) in ((8, 8), (16, 4), (32, 32), (64, 64), (40, 24)):
        a = rng.integers(0, 256, (h, w + 5), dtype=np.uint8)
        b = rng.integers(0, 256, (h, w + 9), dtype=np.uint8)
        assert fn("svt_spatial_full_distortion_kernel", C.c_uint64)(cm.ptr(a), C.c_uint32(0), C.c_uint32(w + 5), cm.ptr(b), 0,
                                                                 C.c_uint32(w + 9), C.c_uint32(w), C.c_uint32(h)) == mo.sse(a[:, :w], b[:, :w])
        assert fn("svt_nxm_sad_kernel_sub_sampled", C.c_uint32)(cm.ptr(a), C.c_uint32(w + 5), cm.ptr(b), C.c_uint32(w + 9), C.c_uint32(h),
                                                               C.c_uint32(w)) == mo.sad(a[:, :w], b[:, :w])
        a16 = rng.integers(0, 1024, (h, w + 5)).astype(np.uint16)
        b16 = rng.integers(0, 1024, (h, w + 9)).astype(np.uint16)
        assert fn("sad_16b_kernel", C.c_uint32)(cm.ptr(a16), C.c_uint32(w + 5), cm.ptr(b16), C.c_uint32(w + 9), C.c_uint32(h),
                                               C.c_uint32(w)) == mo.sad(a16[:, :w], b16[:, :w])
        assert fn("svt_full_distortion_kernel16_bits", C.c_uint64)(cm.ptr(a16), C.c_uint32(0), C.c_uint32(w + 5), cm.ptr(b16), 0,
                                                                C.c_uint32(w + 9), C.c_uint32(w), C.c_uint32(h)) == mo.sse(a16[:, :w], b16[:, :w])
        c = rng.integers(-(1 << 17), 1 << 17, (h, w + 2)).astype(np.int32)
        r = rng.integers(-(1 << 17), 1 << 17, (h, w + 4)).astype(np.int32)
        out = np.zeros(2, np.uint64)
        fn("svt_full_distortion_kernel32_bits")(cm.ptr(c), C.c_uint32(w + 2), cm.ptr(r), C.c_uint32(w + 4), cm.ptr(out), C.c_uint32(w), C.c_uint32(h))
        assert tuple(int(x) for x in out) == mo.full_distortion32(c[:, :w], r[:, :w])
        fn("svt_full_distortion_kernel_cbf_zero32_bits")(cm.ptr(c), C.c_uint32(w + 2), cm.ptr(out), C.c_uint32(w), C.c_uint32(h))
        assert int(out[0]) == int(out[1]) == mo.full_distortion32(c[:, :w], c[:, :w])[1]
        d = np.zeros((h, w + 1), np.int16)
        fn("svt_aom_subtract_block")(h, w, cm.ptr(d), C.c_ssize_t(w + 1), cm.ptr(a), C.c_ssize_t(w + 5), cm.ptr(b), C.c_ssize_t(w + 9))
        np.testing.assert_array_equal(d[:, :w], mo.subtract(a[:, :w], b[:, :w]))
    for n in (16, 64, 256, 1024):
        c = rng.integers(-32640, 32641, n).astype(np.int32)
        dq = rng.integers(-32640, 32641, n).astype(np.int32)
        assert cm.ref().svt_aom_satd_c(cm.ptr(c), n) == mo.satd(c)
        ssz = C.c_int64(0)
        f = cm.ref().svt_av1_block_error_c
        f.restype = C.c_int64
        e = f(cm.ptr(c), cm.ptr(dq), C.c_ssize_t(n), C.byref(ssz))
        assert (e, ssz.value) == mo.block_error(c, dq)
