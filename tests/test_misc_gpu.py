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


@pytest.mark.parametrize("bd", [8, 10, 12])
def test_cdef_dist_dropins(bd):
    import ctypes as C
    import misc_oracle as mo
    from test_oracle_misc import _cdef_dist_case
    lib = sb.load()
    rng = np.random.default_rng(90 + bd)
    f = lib.svt_compute_cdef_dist_16bit_cuda if bd > 8 else lib.svt_compute_cdef_dist_8bit_cuda
    f.restype = C.c_uint64
    for bsize in range(4):
        for pli in (0, 1):
            for _ in range(3):
                fb, src, pick, dl = _cdef_dist_case(rng, bd, bsize, pli)
                got = f(cm.ptr(fb), 64, cm.ptr(np.ascontiguousarray(src)), cm.ptr(dl), len(pick), bsize, bd - 8, pli)
                assert got == mo.cdef_dist(fb, 64, src, pick, bsize, bd - 8, pli)


@pytest.mark.parametrize("bd", [8, 10, 12])
def test_compute_stats_dropins(bd):
    import ctypes as C
    import misc_oracle as mo
    from test_oracle_misc import _stats_case
    lib = sb.load()
    rng = np.random.default_rng(150 + bd)
    for win in (7, 5):
        for (w, h, smooth) in ((64, 64, 0), (40, 24, 1), (96, 56, 1), (17, 9, 0), (1, 1, 0)):
            dgd, src = _stats_case(rng, bd, w, h, smooth)
            n = win * win
            M, Hm = np.zeros(n, np.int64), np.zeros(n * n, np.int64)
            hs, vs = 8, 8
            if bd == 8:
                lib.svt_av1_compute_stats_cuda(win, cm.ptr(dgd), cm.ptr(src), hs, hs + w, vs, vs + h, dgd.shape[1], src.shape[1], cm.ptr(M), cm.ptr(Hm))
            else:
                lib.svt_av1_compute_stats_highbd_cuda(win, C.c_void_p(dgd.ctypes.data >> 1), C.c_void_p(src.ctypes.data >> 1), hs, hs + w, vs,
                                                      vs + h, dgd.shape[1], src.shape[1], cm.ptr(M), cm.ptr(Hm), bd)
            m2, h2 = mo.compute_stats(win, dgd, src, hs, hs + w, vs, vs + h, bd)
            np.testing.assert_array_equal(M, m2)
            np.testing.assert_array_equal(Hm, h2.reshape(-1))


@pytest.mark.parametrize("bd", [8, 10, 12])
def test_compute_stats_extreme_values(bd):
    """The int8 limb split of the tensor-core kernel at its range limits: differences of +-(2^bd - 1) (a unit that is
    almost all 0 with a few full-scale samples, and the reverse), a checkerboard (average in the middle), src != dgd."""
    import ctypes as C
    import misc_oracle as mo
    lib = sb.load()
    rng = np.random.default_rng(170 + bd)
    dt = np.uint16 if bd > 8 else np.uint8
    top = (1 << bd) - 1
    w, h = 128, 96  # 2 x 2 tiles of 64 x 64, ragged bottom
    for kind in ("sparse_hi", "sparse_lo", "checker", "rows"):
        H, W = h + 16, w + 16
        if kind == "sparse_hi":
            dgd = np.where(rng.random((H, W)) < 0.002, top, 0)
            src = np.where(rng.random((H, W)) < 0.5, top, 0)
        elif kind == "sparse_lo":
            dgd = np.where(rng.random((H, W)) < 0.002, 0, top)
            src = np.where(rng.random((H, W)) < 0.5, 0, top)
        elif kind == "checker":
            yy, xx = np.mgrid[0:H, 0:W]
            dgd = np.where((yy + xx) & 1, top, 0)
            src = top - dgd
        else:
            yy, xx = np.mgrid[0:H, 0:W]
            dgd = np.where(yy % 3 == 0, top, np.where(yy % 3 == 1, top // 2 + 1, 0))
            src = np.where(xx % 5 < 2, top, 63)
        dgd, src = dgd.astype(dt), src.astype(dt)
        for win in (7, 5):
            n = win * win
            M, Hm = np.zeros(n, np.int64), np.zeros(n * n, np.int64)
            hs, vs = 8, 8
            if bd == 8:
                lib.svt_av1_compute_stats_cuda(win, cm.ptr(dgd), cm.ptr(src), hs, hs + w, vs, vs + h, dgd.shape[1], src.shape[1], cm.ptr(M), cm.ptr(Hm))
            else:
                lib.svt_av1_compute_stats_highbd_cuda(win, C.c_void_p(dgd.ctypes.data >> 1), C.c_void_p(src.ctypes.data >> 1), hs, hs + w, vs,
                                                      vs + h, dgd.shape[1], src.shape[1], cm.ptr(M), cm.ptr(Hm), bd)
            m2, h2 = mo.compute_stats(win, dgd, src, hs, hs + w, vs, vs + h, bd)
            np.testing.assert_array_equal(M, m2, err_msg=f"{kind} win {win}")
            np.testing.assert_array_equal(Hm, h2.reshape(-1), err_msg=f"{kind} win {win}")


@pytest.mark.parametrize("bd", [8, 10])
def test_pixel_proj_error_dropins(bd):
    import ctypes as C
    import misc_oracle as mo
    lib = sb.load()
    rng = np.random.default_rng(160 + bd)
    dt = np.uint16 if bd > 8 else np.uint8
    f = lib.svt_av1_highbd_pixel_proj_error_cuda if bd > 8 else lib.svt_av1_lowbd_pixel_proj_error_cuda
    f.restype = C.c_int64
    for r in ((2, 1), (2, 0), (0, 1), (0, 0)):
        for (w, h) in ((64, 64), (33, 17), (384, 96)):
            src = rng.integers(0, 1 << bd, (h, w + 5)).astype(dt)
            dat = np.clip(src.astype(np.int64) + rng.integers(-9, 10, src.shape), 0, (1 << bd) - 1).astype(dt)
            f0 = ((dat.astype(np.int64) << 4) + rng.integers(-200, 201, dat.shape)).astype(np.int32)
            f1 = ((dat.astype(np.int64) << 4) + rng.integers(-200, 201, dat.shape)).astype(np.int32)
            xq = (C.c_int32 * 2)(int(rng.integers(-96, 32)), int(rng.integers(-32, 96)))
            params = (C.c_int32 * 4)(r[0], r[1], 0, 0)
            sp = C.c_void_p(src.ctypes.data >> 1) if bd > 8 else cm.ptr(src)
            dp = C.c_void_p(dat.ctypes.data >> 1) if bd > 8 else cm.ptr(dat)
            got = f(sp, w, h, src.shape[1], dp, dat.shape[1], cm.ptr(f0), f0.shape[1], cm.ptr(f1), f1.shape[1], xq, params)
            assert got == mo.pixel_proj_error(src[:, :w], dat[:, :w], f0[:, :w], f1[:, :w], (xq[0], xq[1]), r, bd > 8)


@pytest.mark.parametrize("case", [(192, 136, 8, 64), (264, 200, 10, 128), (1920, 1080, 8, 256), (640, 360, 12, 64), (3840, 2160, 10, 256)])
def test_lr_wiener_stats_picture(case):
    """svt_b200_lr_wiener_stats: every restoration unit of every plane in one call, pictures resident on the device,
    against compute_stats on the replicate-extended picture (the reference extends the picture by 3 before the search)."""
    import ctypes as C
    import torch
    import gpu_runner as gr
    import misc_oracle as mo
    w, h, bd, unit = case
    lib = sb.load()
    src = cm.synth_yuv(w, h, 1, 21, bd)
    dgd = cm.degrade(src, 22, amp=8)
    dd, ds = gr.DevYuv(dgd), gr.DevYuv(src)
    for plane in (0, 1):
        win = 7 if plane == 0 else 5
        U = unit if plane == 0 else unit // 2
        pw, ph = (w, h) if plane == 0 else ((w + 1) // 2, (h + 1) // 2)
        rects = []
        y0 = 0
        while y0 < ph:  # foreach_rest_unit_in_tile without the stripe offset (the search uses the same enumeration)
            uh = ph - y0 if ph - y0 < U * 3 // 2 else U
            x0 = 0
            while x0 < pw:
                uw = pw - x0 if pw - x0 < U * 3 // 2 else U
                rects.append((x0, x0 + uw, y0, y0 + uh))
                x0 += uw
            y0 += uh
        if len(rects) > 24:  # bound the numpy work at 1080p / 2160p: a spread of units incl. the corners
            idx = sorted(set([0, len(rects) - 1] + list(np.random.default_rng(3).choice(len(rects), 20 if w < 3000 else 8, replace=False))))
            rects = [rects[i] for i in idx]
        r = torch.tensor(rects, dtype=torch.int32, device="cuda")
        n2 = win * win
        out = torch.zeros(len(rects) * (n2 + n2 * n2), dtype=torch.int64, device="cuda")
        scratch = torch.zeros(len(rects), dtype=torch.int64, device="cuda")
        mw, mh = max(a[1] - a[0] for a in rects), max(a[3] - a[2] for a in rects)
        a, b = dd.struct(), ds.struct()
        sb.check(lib.svt_b200_lr_wiener_stats(C.byref(a), C.byref(b), plane, win, C.c_void_p(r.data_ptr()), len(rects), mw, mh,
                                              C.c_void_p(out.data_ptr()), C.c_void_p(scratch.data_ptr()), None), lib)
        torch.cuda.synchronize()
        got = out.cpu().numpy().reshape(len(rects), n2 + n2 * n2)
        ext = np.pad(dgd.plane(plane), 3, mode="edge")
        sext = np.pad(src.plane(plane), 3, mode="edge")
        for i, (hs, he, vs, ve) in enumerate(rects):
            m2, h2 = mo.compute_stats(win, ext, sext, hs + 3, he + 3, vs + 3, ve + 3, bd)
            np.testing.assert_array_equal(got[i, :n2], m2, err_msg=f"M unit {i}")
            np.testing.assert_array_equal(got[i, n2:], h2.reshape(-1), err_msg=f"H unit {i}")


@pytest.mark.parametrize("bd", [8, 10])
def test_get_proj_subspace_dropin(bd):
    import ctypes as C
    import misc_oracle as mo
    from test_oracle_misc import _proj_case
    lib = sb.load()
    rng = np.random.default_rng(170 + bd)
    for r in ((2, 1), (2, 0), (0, 1)):
        for (w, h, flat) in ((64, 64, False), (33, 17, False), (48, 40, True), (384, 96, False)):
            src, dat, f0, f1 = _proj_case(rng, bd, w, h, flat)
            xq = (C.c_int32 * 2)(7, 7)
            params = (C.c_int32 * 4)(r[0], r[1], 0, 0)
            sp = C.c_void_p(src.ctypes.data >> 1) if bd > 8 else cm.ptr(src)
            dp = C.c_void_p(dat.ctypes.data >> 1) if bd > 8 else cm.ptr(dat)
            lib.svt_get_proj_subspace_cuda(sp, w, h, src.shape[1], dp, dat.shape[1], int(bd > 8), cm.ptr(f0), f0.shape[1], cm.ptr(f1), f1.shape[1],
                                           xq, params)
            assert [xq[0], xq[1]] == mo.get_proj_subspace(src[:, :w], dat[:, :w], f0[:, :w], f1[:, :w], r)


def test_copy_rect8_dropin():
    lib = sb.load()
    rng = np.random.default_rng(5)
    for (v, h) in ((8, 8), (70, 70), (3, 129)):
        src = rng.integers(0, 256, (v, h + 7)).astype(np.uint8)
        dst = np.full((v, h + 3), 9999, np.uint16)
        lib.svt_copy_rect8_8bit_to_16bit_cuda(cm.ptr(dst), dst.shape[1], cm.ptr(src), src.shape[1], v, h)
        np.testing.assert_array_equal(dst[:, :h], src[:, :h])
        assert (dst[:, h:] == 9999).all()
