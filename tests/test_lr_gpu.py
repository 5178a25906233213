"""GPU parity tests for the loop-restoration kernels (CUDA drop-ins through the C ABI vs the CPU oracle), bit-exact."""
import ctypes as C

import numpy as np
import pytest

import common as cm
import svtb200 as sb
from test_oracle_lr import ConvolveParams, aligned_filter, lr_image, sgr_cases, wiener_filter

pytestmark = pytest.mark.gpu


def test_selfguided_dropins():
    lib, orc = sb.load(), cm.oracle()
    rng = np.random.default_rng(2)
    tmp = np.zeros(16, np.int32)
    for n, (bd, mode, w, h, eps, img) in enumerate(sgr_cases()):
        if n % 2:
            continue
        hbd = bd > 8
        a = img.astype(np.uint16 if hbd else np.uint8)
        stride = a.shape[1]
        p = a.ctypes.data + (8 * stride + 8) * a.itemsize
        f0, f1 = np.full((h, w + 3), -7, np.int32), np.full((h, w + 3), -7, np.int32)
        g0, g1 = f0.copy(), f1.copy()
        lib.svt_av1_selfguided_restoration_cuda(C.c_void_p(p >> 1 if hbd else p), w, h, stride, cm.ptr(f0), cm.ptr(f1), w + 3, eps, bd, int(hbd))
        orc.orc_selfguided_restoration(C.c_void_p(p), int(hbd), w, h, stride, cm.ptr(g0), cm.ptr(g1), w + 3, eps, bd)
        np.testing.assert_array_equal(f0, g0, err_msg=f"flt0 bd{bd} {mode} {w}x{h} eps{eps}")
        np.testing.assert_array_equal(f1, g1, err_msg=f"flt1 bd{bd} {mode} {w}x{h} eps{eps}")
        xqd = (C.c_int32 * 2)(int(rng.integers(-96, 32)), int(rng.integers(-32, 96)))
        d0, d1 = np.zeros((h, w + 5), a.dtype), np.zeros((h, w + 5), a.dtype)
        lib.svt_apply_selfguided_restoration_cuda(C.c_void_p(p >> 1 if hbd else p), w, h, stride, eps, xqd,
                                                  C.c_void_p(d0.ctypes.data >> 1 if hbd else d0.ctypes.data), w + 5, cm.ptr(tmp), bd, int(hbd))
        orc.orc_apply_selfguided_restoration(C.c_void_p(p), int(hbd), w, h, stride, eps, xqd, cm.ptr(d1), w + 5, bd)
        np.testing.assert_array_equal(d0, d1, err_msg=f"apply bd{bd} {mode} {w}x{h} eps{eps}")


def test_wiener_dropins():
    lib, orc = sb.load(), cm.oracle()
    rng = np.random.default_rng(3)
    for bd in (8, 10, 12):
        for mode in ("random", "smooth", "max"):
            for (w, h) in ((64, 64), (32, 16), (16, 64), (8, 8)):
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
                    lib.svt_av1_highbd_wiener_convolve_add_src_cuda(C.c_void_p(p >> 1), C.c_ssize_t(stride), C.c_void_p(d0.ctypes.data >> 1),
                                                                    C.c_ssize_t(w + 5), C.c_void_p(px), C.c_void_p(py), w, h, C.byref(cp), bd)
                else:
                    lib.svt_av1_wiener_convolve_add_src_cuda(C.c_void_p(p), C.c_ssize_t(stride), cm.ptr(d0), C.c_ssize_t(w + 5), C.c_void_p(px),
                                                             C.c_void_p(py), w, h, C.byref(cp))
                orc.orc_wiener_convolve_add_src(C.c_void_p(p), int(hbd), C.c_ssize_t(stride), cm.ptr(d1), C.c_ssize_t(w + 5), cm.ptr(fx),
                                                cm.ptr(fy), w, h, cp.round_0, cp.round_1, bd)
                np.testing.assert_array_equal(d0, d1, err_msg=f"bd{bd} {mode} {w}x{h}")


@pytest.mark.parametrize("case", [(192, 136, 8, 1, (64, 32, 32), ("mix", "mix", "mix"), (3, 3, 3), 0),
                                  (192, 136, 10, 2, (64, 32, 32), ("wiener", "sgr", "mix"), (1, 2, 3), 0),
                                  (264, 200, 8, 4, (64, 64, 32), ("sgr", "wiener", "mix"), (2, 1, 3), 1),
                                  (328, 72, 10, 5, (256, 128, 128), ("mix", "mix", "mix"), (3, 3, 3), 0),
                                  (200, 328, 8, 6, (64, 32, 64), ("mix", "none", "wiener"), (3, 0, 1), 1),
                                  (1920, 1080, 8, 7, (64, 32, 32), ("mix", "mix", "mix"), (3, 3, 3), 0),
                                  (1920, 1080, 10, 8, (256, 128, 128), ("sgr", "wiener", "mix"), (2, 1, 3), 0)])
def test_lr_frame_vs_oracle(case):
    """svt_av1_loop_restoration_filter_frame: the closed-form (plane, stripe, column) CTA decomposition against the
    oracle's in-place save / overwrite / restore walk (itself pinned against the reference), incl. full 1080p."""
    import gpu_runner as gr
    from test_oracle_lr_frame import lr_case, run_oracle_lr
    w, h, bd, seed, unit_sizes, modes, frame_types, optimized = case
    cdef, dblk, units = lr_case(w, h, bd, seed, unit_sizes, modes)
    want = run_oracle_lr(cdef, dblk, units, unit_sizes, frame_types, optimized)
    got = gr.run_gpu_lr(cdef, dblk, units, unit_sizes, frame_types, optimized)
    for i in range(3):
        np.testing.assert_array_equal(got.plane(i), want.plane(i), err_msg=f"plane {i}")
