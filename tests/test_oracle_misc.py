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


def _cdef_dist_case(rng, bd, bsize, pli):
    bw = 8 if bsize in (2, 3) else 4
    bh = 8 if bsize in (1, 3) else 4
    dt = np.uint16 if bd > 8 else np.uint8
    fb = rng.integers(0, 1 << bd, (64, 64)).astype(dt)
    fb[:32] = (fb[:32].astype(np.int64) // 8 + (1 << (bd - 1))).astype(dt)  # a low-variance half
    cells = [(by, bx) for by in range(64 // bh) for bx in range(64 // bw)]
    pick = [cells[i] for i in sorted(rng.choice(len(cells), int(rng.integers(1, 40)), replace=False))]
    src = np.stack([np.clip(fb[by * bh:(by + 1) * bh, bx * bw:(bx + 1) * bw].astype(np.int64) + rng.integers(-9, 10, (bh, bw)), 0,
                            (1 << bd) - 1).reshape(-1) for by, bx in pick]).astype(dt)
    dl = np.array([(by, bx, 0) for by, bx in pick], np.uint8)
    return fb, src, pick, dl


@needs_ref
@pytest.mark.parametrize("bd", [8, 10, 12])
def test_cdef_dist_restatement_matches_reference(bd):
    """misc_oracle.cdef_dist against compute_cdef_dist_8bit_c / compute_cdef_dist_c of the reference."""
    import ctypes as C
    import misc_oracle as mo
    ref = cm.refh()
    rng = np.random.default_rng(40 + bd)
    for bsize in range(4):
        for pli in (0, 1):
            for _ in range(6):
                fb, src, pick, dl = _cdef_dist_case(rng, bd, bsize, pli)
                f = ref.compute_cdef_dist_c if bd > 8 else ref.compute_cdef_dist_8bit_c
                f.restype = C.c_uint64
                want = f(cm.ptr(fb), 64, cm.ptr(np.ascontiguousarray(src)), cm.ptr(dl), len(pick), bsize, bd - 8, pli)
                assert mo.cdef_dist(fb, 64, src, pick, bsize, bd - 8, pli) == want


def _stats_case(rng, bd, w, h, smooth):
    dt = np.uint16 if bd > 8 else np.uint8
    H, W = h + 16, w + 16
    base = rng.integers(0, 1 << bd, (H, W))
    if smooth:
        yy, xx = np.mgrid[0:H, 0:W]
        base = ((np.sin(xx / 9.0) + np.cos(yy / 7.0) + 2) * ((1 << bd) - 1) / 4 + rng.integers(-3, 4, (H, W))).clip(0, (1 << bd) - 1)
    dgd = base.astype(dt)
    src = np.clip(base + rng.integers(-12, 13, (H, W)), 0, (1 << bd) - 1).astype(dt)
    return dgd, src


@needs_ref
@pytest.mark.parametrize("bd", [8, 10, 12])
def test_compute_stats_restatement_matches_reference(bd):
    ref = cm.refh()
    rng = np.random.default_rng(50 + bd)
    for win in (7, 5):
        for (w, h, smooth) in ((64, 64, 0), (40, 24, 1), (96, 56, 1), (17, 9, 0)):
            dgd, src = _stats_case(rng, bd, w, h, smooth)
            n = win * win
            M, Hm = np.zeros(n, np.int64), np.zeros(n * n, np.int64)
            hs, vs = 8, 8
            if bd == 8:
                ref.svt_av1_compute_stats_c(win, cm.ptr(dgd), cm.ptr(src), hs, hs + w, vs, vs + h, dgd.shape[1], src.shape[1], cm.ptr(M), cm.ptr(Hm))
            else:
                ref.svt_av1_compute_stats_highbd_c(win, C.c_void_p(dgd.ctypes.data >> 1), C.c_void_p(src.ctypes.data >> 1), hs, hs + w, vs, vs + h,
                                                   dgd.shape[1], src.shape[1], cm.ptr(M), cm.ptr(Hm), bd)
            m2, h2 = mo.compute_stats(win, dgd, src, hs, hs + w, vs, vs + h, bd)
            np.testing.assert_array_equal(m2, M)
            np.testing.assert_array_equal(h2.reshape(-1), Hm)


@needs_ref
@pytest.mark.parametrize("bd", [8, 10])
def test_pixel_proj_error_restatement_matches_reference(bd):
    ref = cm.refh()
    rng = np.random.default_rng(60 + bd)
    dt = np.uint16 if bd > 8 else np.uint8
    for r in ((2, 1), (2, 0), (0, 1), (0, 0)):
        for (w, h) in ((64, 64), (33, 17)):
            src = rng.integers(0, 1 << bd, (h, w + 5)).astype(dt)
            dat = np.clip(src.astype(np.int64) + rng.integers(-9, 10, src.shape), 0, (1 << bd) - 1).astype(dt)
            f0 = ((dat.astype(np.int64) << 4) + rng.integers(-200, 201, dat.shape)).astype(np.int32)
            f1 = ((dat.astype(np.int64) << 4) + rng.integers(-200, 201, dat.shape)).astype(np.int32)
            xq = (C.c_int32 * 2)(int(rng.integers(-96, 32)), int(rng.integers(-32, 96)))
            params = (C.c_int32 * 4)(r[0], r[1], 0, 0)
            f = ref.svt_av1_highbd_pixel_proj_error_c if bd > 8 else ref.svt_av1_lowbd_pixel_proj_error_c
            f.restype = C.c_int64
            sp = C.c_void_p(src.ctypes.data >> 1) if bd > 8 else cm.ptr(src)
            dp = C.c_void_p(dat.ctypes.data >> 1) if bd > 8 else cm.ptr(dat)
            want = f(sp, w, h, src.shape[1], dp, dat.shape[1], cm.ptr(f0), f0.shape[1], cm.ptr(f1), f1.shape[1], xq, params)
            got = mo.pixel_proj_error(src[:, :w], dat[:, :w], f0[:, :w], f1[:, :w], (xq[0], xq[1]), r, bd > 8)
            assert got == want


def _proj_case(rng, bd, w, h, flat=False):
    dt = np.uint16 if bd > 8 else np.uint8
    src = rng.integers(0, 1 << bd, (h, w + 5)).astype(dt)
    dat = np.clip(src.astype(np.int64) + rng.integers(-9, 10, src.shape), 0, (1 << bd) - 1).astype(dt)
    if flat:
        f0 = (dat.astype(np.int64) << 4).astype(np.int32)
        f1 = f0.copy()
    else:
        f0 = ((dat.astype(np.int64) << 4) + rng.integers(-200, 201, dat.shape)).astype(np.int32)
        f1 = ((src.astype(np.int64) << 4) + rng.integers(-90, 91, dat.shape)).astype(np.int32)
    return src, dat, f0, f1


@needs_ref
@pytest.mark.parametrize("bd", [8, 10])
def test_get_proj_subspace_restatement_matches_reference(bd):
    ref = cm.refh()
    rng = np.random.default_rng(70 + bd)
    for r in ((2, 1), (2, 0), (0, 1)):
        for (w, h, flat) in ((64, 64, False), (33, 17, False), (48, 40, True), (384, 96, False)):
            src, dat, f0, f1 = _proj_case(rng, bd, w, h, flat)
            xq = (C.c_int32 * 2)(7, 7)
            params = (C.c_int32 * 4)(r[0], r[1], 0, 0)
            sp = C.c_void_p(src.ctypes.data >> 1) if bd > 8 else cm.ptr(src)
            dp = C.c_void_p(dat.ctypes.data >> 1) if bd > 8 else cm.ptr(dat)
            ref.svt_get_proj_subspace_c(sp, w, h, src.shape[1], dp, dat.shape[1], int(bd > 8), cm.ptr(f0), f0.shape[1], cm.ptr(f1), f1.shape[1],
                                        xq, params)
            assert mo.get_proj_subspace(src[:, :w], dat[:, :w], f0[:, :w], f1[:, :w], r) == [xq[0], xq[1]]
