"""GPU parity tests for the small reductions and the SAD family (CUDA drop-ins vs the numpy oracle)."""
import ctypes as C

import numpy as np
import pytest

import common as cm
import misc_oracle as mo
import svtb200 as sb
from test_oracle_misc import SAD_SIZES

pytestmark = pytest.mark.gpu


def test_sad_family_dropins():
    lib = sb.load()
    rng = np.random.default_rng(1)
    for (w, h) in SAD_SIZES:
        a = rng.integers(0, 256, (h, w + 7), dtype=np.uint8)
        refs = [rng.integers(0, 256, (h, w + 3), dtype=np.uint8) for _ in range(4)]
        f = getattr(lib, f"svt_aom_sad{w}x{h}_cuda")
        f.restype = C.c_uint32
        assert f(cm.ptr(a), w + 7, cm.ptr(refs[0]), w + 3) == mo.sad(a[:, :w], refs[0][:, :w])
        if (w * h) % 512 == 0:
            arr = (C.c_void_p * 4)(*[r.ctypes.data for r in refs])
            out = np.zeros(4, np.uint32)
            getattr(lib, f"svt_aom_sad{w}x{h}x4d_cuda")(cm.ptr(a), w + 7, arr, w + 3, cm.ptr(out))
            assert out.tolist() == [mo.sad(a[:, :w], r[:, :w]) for r in refs]


def test_distortion_and_subtract_dropins():
    lib = sb.load()
    for n in ("svt_spatial_full_distortion_kernel_cuda", "svt_full_distortion_kernel16_bits_cuda"):
        getattr(lib, n).restype = C.c_uint64
    lib.svt_av1_block_error_cuda.restype = C.c_int64
    lib.svt_nxm_sad_kernel_sub_sampled_cuda.restype = C.c_uint32
    lib.sad_16b_kernel_cuda.restype = C.c_uint32
    rng = np.random.default_rng(2)
    for (w, h) in ((8, 8), (16, 4), (32, 32), (64, 64), (40, 24)):
        a = rng.integers(0, 256, (h, w + 5), dtype=np.uint8)
        b = rng.integers(0, 256, (h, w + 9), dtype=np.uint8)
        assert lib.svt_spatial_full_distortion_kernel_cuda(cm.ptr(a), C.c_uint32(0), C.c_uint32(w + 5), cm.ptr(b), 0, C.c_uint32(w + 9),
                                                           C.c_uint32(w), C.c_uint32(h)) == mo.sse(a[:, :w], b[:, :w])
        assert lib.svt_nxm_sad_kernel_sub_sampled_cuda(cm.ptr(a), C.c_uint32(w + 5), cm.ptr(b), C.c_uint32(w + 9), C.c_uint32(h),
                                                       C.c_uint32(w)) == mo.sad(a[:, :w], b[:, :w])
        a16 = rng.integers(0, 1024, (h, w + 5)).astype(np.uint16)
        b16 = rng.integers(0, 1024, (h, w + 9)).astype(np.uint16)
        assert lib.sad_16b_kernel_cuda(cm.ptr(a16), C.c_uint32(w + 5), cm.ptr(b16), C.c_uint32(w + 9), C.c_uint32(h), C.c_uint32(w)) == mo.sad(a16[:, :w], b16[:, :w])
        assert lib.svt_full_distortion_kernel16_bits_cuda(cm.ptr(a16), C.c_uint32(0), C.c_uint32(w + 5), cm.ptr(b16), 0, C.c_uint32(w + 9),
                                                          C.c_uint32(w), C.c_uint32(h)) == mo.sse(a16[:, :w], b16[:, :w])
        c = rng.integers(-(1 << 17), 1 << 17, (h, w + 2)).astype(np.int32)
        r = rng.integers(-(1 << 17), 1 << 17, (h, w + 4)).astype(np.int32)
        out = np.zeros(2, np.uint64)
        lib.svt_full_distortion_kernel32_bits_cuda(cm.ptr(c), C.c_uint32(w + 2), cm.ptr(r), C.c_uint32(w + 4), cm.ptr(out), C.c_uint32(w), C.c_uint32(h))
        assert tuple(int(x) for x in out) == mo.full_distortion32(c[:, :w], r[:, :w])
        lib.svt_full_distortion_kernel_cbf_zero32_bits_cuda(cm.ptr(c), C.c_uint32(w + 2), cm.ptr(out), C.c_uint32(w), C.c_uint32(h))
        assert int(out[0]) == int(out[1]) == mo.full_distortion32(c[:, :w], c[:, :w])[1]
        d = np.zeros((h, w + 1), np.int16)
        lib.svt_aom_subtract_block_cuda(h, w, cm.ptr(d), C.c_ssize_t(w + 1), cm.ptr(a), C.c_ssize_t(w + 5), cm.ptr(b), C.c_ssize_t(w + 9))
        np.testing.assert_array_equal(d[:, :w], mo.subtract(a[:, :w], b[:, :w]))
        d[:] = 0
        lib.svt_aom_highbd_subtract_block_cuda(h, w, cm.ptr(d), C.c_ssize_t(w + 1), C.c_void_p(a16.ctypes.data >> 1), C.c_ssize_t(w + 5),
                                               C.c_void_p(b16.ctypes.data >> 1), C.c_ssize_t(w + 9), 10)
        np.testing.assert_array_equal(d[:, :w], mo.subtract(a16[:, :w], b16[:, :w]))
    for n in (16, 64, 256, 1024):
        c = rng.integers(-32640, 32641, n).astype(np.int32)
        dq = rng.integers(-32640, 32641, n).astype(np.int32)
        assert lib.svt_aom_satd_cuda(cm.ptr(c), n) == mo.satd(c)
        ssz = C.c_int64(0)
        e = lib.svt_av1_block_error_cuda(cm.ptr(c), cm.ptr(dq), C.c_ssize_t(n), C.byref(ssz))
        assert (e, ssz.value) == mo.block_error(c, dq)
