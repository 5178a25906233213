"""Pins the ME part of the oracle (oracle/me_oracle.c) against the UNMODIFIED reference C path compiled into
oracle/_ref (the reference holds no golden vectors for this path; its tests are C-vs-SIMD on seeded
random data — test/SadTest.cc — so the pin is 'same fixtures, reference _c vs restatement')."""
import ctypes as C

import numpy as np
import pytest

import common as cm
import svtb200 as sb

needs_ref = pytest.mark.skipif(not cm.have_ref(), reason="oracle/_ref not built (no /root/reference here)")

# test/SadTest.cc:90-101 block sizes and :543-554 search areas (width, height)
BLOCKS = [(64, 64), (64, 32), (32, 64), (32, 32), (32, 16), (16, 32), (16, 16), (16, 8), (8, 16), (8, 8), (8, 4),
          (4, 4), (4, 8), (4, 16), (16, 4), (8, 32), (32, 8), (16, 64), (64, 16), (24, 24), (24, 16), (16, 24),
          (48, 48), (48, 16), (16, 48)]
AREAS = [(64, 125), (192, 75), (128, 50), (64, 25), (240, 200), (144, 120), (96, 80), (48, 40), (240, 120),
         (144, 72), (96, 48), (48, 24), (560, 320), (336, 192), (224, 128), (112, 64), (640, 400), (384, 240),
         (256, 160), (128, 80), (480, 120), (288, 72), (192, 48), (96, 24), (160, 60), (96, 36), (64, 24), (32, 12)]


def sad_case(rng, bw, bh, saw, sah, pattern):
    stride = bw + saw + 7
    src = rng.integers(0, 256, (bh, bw + 3), dtype=np.uint8)
    ref = rng.integers(0, 256, (bh + sah, stride), dtype=np.uint8)
    if pattern == "REF_MAX":
        ref[:] = 255
    elif pattern == "SRC_MAX":
        src[:] = 255
    return src, ref, stride


@needs_ref
@pytest.mark.parametrize("pattern", ["RANDOM", "REF_MAX", "SRC_MAX"])
def test_sad_loop_kernel_matches_reference(pattern):
    rng = np.random.default_rng(7)
    fn = C.cast(C.c_void_p.in_dll(cm.ref(), "svt_sad_loop_kernel").value, C.CFUNCTYPE(None))
    for (bw, bh) in BLOCKS[:12]:
        for (saw, sah) in [(a // 8 + 1, b // 8 + 1) for a, b in AREAS[::3]]:
            src, ref, stride = sad_case(rng, bw, bh, saw, sah, pattern)
            res = []
            for f in (fn, cm.oracle().orc_sad_loop_kernel):
                best = C.c_uint64(0)
                x = C.c_int16(-7)
                y = C.c_int16(-9)
                f(cm.ptr(src), C.c_uint32(src.shape[1]), cm.ptr(ref), C.c_uint32(stride), C.c_uint32(bh),
                  C.c_uint32(bw), C.byref(best), C.byref(x), C.byref(y), C.c_uint32(stride), C.c_int16(saw),
                  C.c_int16(sah))
                res.append((best.value, x.value, y.value))
            assert res[0] == res[1], (bw, bh, saw, sah)


@needs_ref
@pytest.mark.parametrize("sub_sad", [0, 1])
def test_ext_all_sad_matches_reference(sub_sad):
    rng = np.random.default_rng(11)
    lib = cm.ref()
    f_all = C.cast(C.c_void_p.in_dll(lib, "svt_ext_all_sad_calculation_8x8_16x16").value, C.CFUNCTYPE(None))
    f_32 = C.cast(C.c_void_p.in_dll(lib, "svt_ext_eight_sad_calculation_32x32_64x64").value, C.CFUNCTYPE(None))
    for it in range(20):
        src = rng.integers(0, 256, (64, 80), dtype=np.uint8)
        ref = rng.integers(0, 256, (64, 96), dtype=np.uint8)
        if it == 0:
            src[:] = 255
            ref[:] = 0
        mv = int(rng.integers(0, 2 ** 32, dtype=np.uint64)) & 0xFFFCFFFC
        outs = []
        for fa, f3 in ((f_all, f_32), (cm.oracle().orc_ext_all_sad_calculation_8x8_16x16,
                                      cm.oracle().orc_ext_eight_sad_calculation_32x32_64x64)):
            st = np.random.default_rng(it)
            b8 = st.integers(0, 9000, 64).astype(np.uint32)
            b16 = st.integers(0, 30000, 16).astype(np.uint32)
            b32 = st.integers(0, 120000, 4).astype(np.uint32)
            b64 = st.integers(0, 500000, 1).astype(np.uint32)
            m8, m16, m32, m64 = (np.zeros(n, np.uint32) for n in (64, 16, 4, 1))
            e16, e8, e32 = np.zeros((16, 8), np.uint32), np.zeros((64, 8), np.uint32), np.zeros((4, 8), np.uint32)
            fa(cm.ptr(src), C.c_uint32(80), cm.ptr(ref), C.c_uint32(96), C.c_uint32(mv), cm.ptr(b8), cm.ptr(b16),
               cm.ptr(m8), cm.ptr(m16), cm.ptr(e16), cm.ptr(e8), C.c_uint8(sub_sad))
            f3(cm.ptr(e16), cm.ptr(b32), cm.ptr(b64), cm.ptr(m32), cm.ptr(m64), C.c_uint32(mv), cm.ptr(e32))
            outs.append([b8, b16, b32, b64, m8, m16, m32, m64, e16, e8, e32])
        for a, b in zip(*outs):
            np.testing.assert_array_equal(a, b)


ME_CASES = [
    # (w, h, n_l0, n_l1, temporal_layer, is_ref, dists)
    (640, 360, 2, 2, 1, 1, ((1, 2, 3, 4), (1, 2, 3, 4))),
    (640, 360, 1, 0, 0, 1, ((4, 8, 12, 16), (1, 2, 3, 4))),   # P picture, base layer
    (320, 192, 4, 3, 0, 1, ((16, 8, 4, 2), (16, 8, 4, 1))),   # base-layer B: list1 HME skipped
    (384, 200, 2, 1, 3, 0, ((1, 3, 3, 4), (1, 2, 3, 4))),     # non-reference picture, ragged height
    (200, 136, 1, 1, 2, 1, ((2, 2, 3, 4), (2, 2, 3, 4))),     # ragged width and height
]


@needs_ref
@pytest.mark.parametrize("case", ME_CASES)
def test_me_picture_matches_reference(case):
    w, h, n0, n1, tl, isref, dist = case
    geos, src, refs = cm.make_me_case(w, h, n0, n1, seed=100 + w)
    params, want = cm.run_ref_me(w, h, 8, n0, n1, dist, tl, isref, geos, src, refs)
    got = cm.run_oracle_me(params, src, refs)
    cm.assert_me_equal(got, want, params, "oracle-vs-reference")
    # the host-side preset table (svtb200.preset8_me_params) must be what the reference derives
    if w * h >= 0x4CE00:
        assert bytes(sb.preset8_me_params(w, h, n0, n1, dist, tl, isref)) == bytes(params)


@needs_ref
def test_me_static_content_exercises_zero_centre_and_sr_shrink():
    w, h = 256, 192
    dist = ((1, 2, 3, 4), (1, 2, 3, 4))
    geos, src, refs = cm.make_me_case(w, h, 2, 2, seed=5, motion=False)
    params, want = cm.run_ref_me(w, h, 8, 2, 2, dist, 1, 1, geos, src, refs)
    got = cm.run_oracle_me(params, src, refs)
    cm.assert_me_equal(got, want, params, "static")


DOWNSAMPLE_SIZES = [(1920, 1080), (640, 360), (322, 182), (176, 144), (70, 66)]


def downsample_case(w, h, seed):
    """Full-resolution padded luma + zeroed 1/4 and 1/16 buffers (a canary value in the padding)."""
    geos = sb.me_geometry(w, h)
    rng = np.random.default_rng(seed)
    img = cm.synth_luma(w, h, 3, seed) if seed % 2 else rng.integers(0, 256, (h, w)).astype(np.uint8)
    full = cm.pad_plane(img, geos[0])
    q = np.full((geos[1].height + 2 * geos[1].origin_y, geos[1].stride), 0xA5, np.uint8)
    s = np.full((geos[2].height + 2 * geos[2].origin_y, geos[2].stride), 0x5A, np.uint8)
    return geos, full, q, s


def run_downsample(fn, geos, full, q, s, filtered):
    planes = sb.MePlanes(full.ctypes.data, q.ctypes.data, s.ctypes.data)
    fn(C.byref(geos[0]), C.byref(geos[1]), C.byref(geos[2]), C.byref(planes), filtered)
    return q, s


@pytest.mark.skipif(not cm.have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("filtered", [1, 0])
@pytest.mark.parametrize("w,h", DOWNSAMPLE_SIZES)
def test_me_downsample_matches_reference(w, h, filtered):
    """orc_me_downsample vs downsample_filtering_input_picture_ime / downsample_decimation_input_picture_ime."""
    for seed in (1, 2):
        geos, full, q0, s0 = downsample_case(w, h, seed)
        qa, sa = run_downsample(cm.oracle().orc_me_downsample, geos, full, q0.copy(), s0.copy(), filtered)
        qb, sbb = run_downsample(cm.refh().refh_me_downsample, geos, full, q0.copy(), s0.copy(), filtered)
        np.testing.assert_array_equal(qa, qb)
        np.testing.assert_array_equal(sa, sbb)
        assert not (qa == 0xA5).all() and not (sa == 0x5A).all()
