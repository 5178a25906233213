"""GPU parity tests for motion estimation: CUDA path (through the C ABI) vs the CPU oracle, bit-exact."""
import ctypes as C

import numpy as np
import pytest

import common as cm
import svtb200 as sb
from test_oracle_me import AREAS, BLOCKS, ME_CASES, sad_case

pytestmark = pytest.mark.gpu


def _call_sad_loop(f, src, ref, stride, bw, bh, saw, sah, sub=False):
    best = C.c_uint64(0)
    x, y = C.c_int16(-7), C.c_int16(-9)
    k = 2 if sub else 1
    f(cm.ptr(src), C.c_uint32(src.shape[1] * k), cm.ptr(ref), C.c_uint32(stride * k), C.c_uint32(bh // k),
      C.c_uint32(bw), C.byref(best), C.byref(x), C.byref(y), C.c_uint32(stride), C.c_int16(saw), C.c_int16(sah))
    return best.value, x.value, y.value


@pytest.mark.parametrize("pattern", ["RANDOM", "REF_MAX", "SRC_MAX"])
def test_sad_loop_kernel_dropin(pattern):
    """test/SadTest.cc sad_LoopTest: every block size x search areas, best SAD and x/y centre."""
    lib = sb.load()
    rng = np.random.default_rng(3)
    for (bw, bh) in BLOCKS:
        for (saw, sah) in [(a // 8 + 1, b // 8 + 1) for a, b in AREAS[::4]] + [(1, 1), (8, 3), (16, 16)]:
            src, ref, stride = sad_case(rng, bw, bh, saw, sah, pattern)
            for sub in (False, True):
                if sub and bh < 2:
                    continue
                a = _call_sad_loop(lib.svt_sad_loop_kernel_cuda, src, ref, stride, bw, bh, saw, sah, sub)
                b = _call_sad_loop(cm.oracle().orc_sad_loop_kernel, src, ref, stride, bw, bh, saw, sah, sub)
                assert a == b, (bw, bh, saw, sah, sub, a, b)


def test_sad_loop_kernel_dropin_large_area():
    lib = sb.load()
    rng = np.random.default_rng(4)
    for (bw, bh, saw, sah) in [(64, 64, 80, 50), (16, 16, 640 // 4, 400 // 4), (32, 32, 200, 3)]:
        src, ref, stride = sad_case(rng, bw, bh, saw, sah, "RANDOM")
        a = _call_sad_loop(lib.svt_sad_loop_kernel_cuda, src, ref, stride, bw, bh, saw, sah)
        b = _call_sad_loop(cm.oracle().orc_sad_loop_kernel, src, ref, stride, bw, bh, saw, sah)
        assert a == b


@pytest.mark.parametrize("sub_sad", [0, 1])
def test_ext_sad_dropins(sub_sad):
    """test/SadTest.cc Allsad_CalculationTest / Extsad_CalculationTest."""
    lib, orc = sb.load(), cm.oracle()
    rng = np.random.default_rng(11)
    for it in range(12):
        src = rng.integers(0, 256, (64, 80), dtype=np.uint8)
        ref = rng.integers(0, 256, (64, 96), dtype=np.uint8)
        if it == 0:
            src[:], ref[:] = 255, 0
        if it == 1:
            src[:], ref[:] = 7, 7
        mv = int(rng.integers(0, 2 ** 32, dtype=np.uint64)) & 0xFFFCFFFC
        outs = []
        for fns in ((lib.svt_ext_all_sad_calculation_8x8_16x16_cuda, lib.svt_ext_eight_sad_calculation_32x32_64x64_cuda,
                     lib.svt_ext_sad_calculation_8x8_16x16_cuda, lib.svt_ext_sad_calculation_32x32_64x64_cuda),
                    (orc.orc_ext_all_sad_calculation_8x8_16x16, orc.orc_ext_eight_sad_calculation_32x32_64x64,
                     orc.orc_ext_sad_calculation_8x8_16x16, orc.orc_ext_sad_calculation_32x32_64x64)):
            st = np.random.default_rng(it)
            b8 = st.integers(0, 9000, 64).astype(np.uint32)
            b16 = st.integers(0, 30000, 16).astype(np.uint32)
            b32 = st.integers(0, 120000, 4).astype(np.uint32)
            b64 = st.integers(0, 500000, 1).astype(np.uint32)
            m8, m16, m32, m64 = (np.zeros(n, np.uint32) for n in (64, 16, 4, 1))
            e16, e8, e32 = np.zeros((16, 8), np.uint32), np.zeros((64, 8), np.uint32), np.zeros((4, 8), np.uint32)
            fns[0](cm.ptr(src), C.c_uint32(80), cm.ptr(ref), C.c_uint32(96), C.c_uint32(mv), cm.ptr(b8), cm.ptr(b16),
                   cm.ptr(m8), cm.ptr(m16), cm.ptr(e16), cm.ptr(e8), C.c_uint8(sub_sad))
            fns[1](cm.ptr(e16), cm.ptr(b32), cm.ptr(b64), cm.ptr(m32), cm.ptr(m64), C.c_uint32(mv), cm.ptr(e32))
            # single-point variants on the first 16x16
            s16, s8, s32 = np.zeros(16, np.uint32), np.zeros(4, np.uint32), np.zeros(4, np.uint32)
            fns[2](cm.ptr(src), C.c_uint32(80), cm.ptr(ref), C.c_uint32(96), cm.ptr(b8), cm.ptr(b16), cm.ptr(m8),
                   cm.ptr(m16), C.c_uint32(mv ^ 0x40004), cm.ptr(s16), cm.ptr(s8), C.c_uint8(sub_sad))
            s16[1:] = st.integers(0, 30000, 15)
            fns[3](cm.ptr(s16), cm.ptr(b32), cm.ptr(b64), cm.ptr(m32), cm.ptr(m64), C.c_uint32(mv ^ 0x40004), cm.ptr(s32))
            outs.append([b8, b16, b32, b64, m8, m16, m32, m64, e16, e8, e32, s16, s8, s32])
        for i, (a, b) in enumerate(zip(*outs)):
            np.testing.assert_array_equal(a, b, err_msg=str(i))


def test_nxm_sad_and_fill_dropins():
    lib, orc = sb.load(), cm.oracle()
    rng = np.random.default_rng(5)
    for (w, h) in [(64, 32), (64, 28), (22, 5), (8, 8), (128, 64)]:
        a = rng.integers(0, 256, (h, w + 9), dtype=np.uint8)
        b = rng.integers(0, 256, (h, w + 5), dtype=np.uint8)
        got = lib.svt_nxm_sad_kernel_cuda(cm.ptr(a), C.c_uint32(w + 9), cm.ptr(b), C.c_uint32(w + 5), C.c_uint32(h), C.c_uint32(w))
        want = orc.orc_nxm_sad(cm.ptr(a), C.c_uint32(w + 9), cm.ptr(b), C.c_uint32(w + 5), C.c_uint32(h), C.c_uint32(w))
        assert got == want
    buf = np.zeros(85, np.uint32)
    lib.svt_initialize_buffer_32bits_cuda(cm.ptr(buf), C.c_uint32(21), C.c_uint32(1), C.c_uint32(128 * 128 * 255))
    assert (buf == 128 * 128 * 255).all()


@pytest.mark.parametrize("case", ME_CASES)
def test_me_picture_vs_oracle(case):
    import gpu_runner as gr
    w, h, n0, n1, tl, isref, dist = case
    geos, src, refs = cm.make_me_case(w, h, n0, n1, seed=100 + w)
    params = sb.preset8_me_params(w, h, n0, n1, dist, tl, isref)
    want = cm.run_oracle_me(params, src, refs)
    got = gr.run_gpu_me(params, src, refs)
    cm.assert_me_equal(got, want, params, "gpu-vs-oracle")


def test_me_picture_static_content():
    import gpu_runner as gr
    w, h = 256, 192
    dist = ((1, 2, 3, 4), (1, 2, 3, 4))
    geos, src, refs = cm.make_me_case(w, h, 2, 2, seed=5, motion=False)
    params = sb.preset8_me_params(w, h, 2, 2, dist, 1, 1)
    cm.assert_me_equal(gr.run_gpu_me(params, src, refs), cm.run_oracle_me(params, src, refs), params, "static")


def test_me_picture_1080p_full_size():
    """BASELINE config 2 geometry (1920x1080, 510 SBs): full bit-exact comparison + idempotence."""
    import gpu_runner as gr
    w, h = 1920, 1080
    dist = ((1, 2, 3, 4), (1, 2, 3, 4))
    geos, src, refs = cm.make_me_case(w, h, 2, 2, seed=77)
    params = sb.preset8_me_params(w, h, 2, 2, dist, 2, 1)
    got = gr.run_gpu_me(params, src, refs)
    again = gr.run_gpu_me(params, src, refs)
    cm.assert_me_equal(got, again, params, "idempotence")
    cm.assert_me_equal(got, cm.run_oracle_me(params, src, refs), params, "1080p gpu-vs-oracle")
    # a picture searched against itself must find SAD 0 for every PU (the MV may be a non-zero tie that comes
    # earlier in raster order on flat/periodic content, exactly as in the reference)
    refs_same = [src] * 8
    same = gr.run_gpu_me(params, src, refs_same)
    assert (same.best_sad[:, 0, 0] == 0).all()


@pytest.mark.parametrize("filtered", [1, 0])
@pytest.mark.parametrize("w,h", [(1920, 1080), (640, 360), (322, 182), (176, 144), (70, 66)])
def test_me_downsample_vs_oracle(w, h, filtered):
    """svt_b200_me_downsample: both HME planes with their padding, every byte of the buffers, vs the oracle."""
    import torch
    from test_oracle_me import downsample_case, run_downsample
    lib = sb.load()
    geos, full, q0, s0 = downsample_case(w, h, 7 + filtered)
    want_q, want_s = run_downsample(cm.oracle().orc_me_downsample, geos, full, q0.copy(), s0.copy(), filtered)
    d_full, d_q, d_s = torch.from_numpy(full).cuda(), torch.from_numpy(q0).cuda(), torch.from_numpy(s0).cuda()
    planes = sb.MePlanes(d_full.data_ptr(), d_q.data_ptr(), d_s.data_ptr())
    sb.check(lib.svt_b200_me_downsample(C.byref(geos[0]), C.byref(geos[1]), C.byref(geos[2]), C.byref(planes), filtered, None), lib)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(d_q.cpu().numpy(), want_q)
    np.testing.assert_array_equal(d_s.cpu().numpy(), want_s)
