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


SGR_R = [(2, 1)] * 10 + [(0, 1)] * 4 + [(2, 0)] * 2  # eb_sgr_params radii


@pytest.mark.parametrize("case", [(192, 136, 8, 64, (0, 3, 11, 15)), (200, 104, 10, 128, (5, 12, 14)), (640, 360, 8, 64, tuple(range(16)))])
def test_lr_sgr_search_picture(case):
    """svt_b200_lr_sgr_filter_sums / svt_b200_lr_sgr_proj_error: every unit x parameter set of a plane at once against the
    oracle's self-guided filter applied per processing unit (apply_sgr) + numpy sums / errors."""
    import torch
    import gpu_runner as gr
    import misc_oracle as mo
    w, h, bd, unit, eps = case
    lib, orc = sb.load(), cm.oracle()
    src = cm.synth_yuv(w, h, 1, 31, bd)
    dgd = cm.degrade(src, 32, amp=8)
    dd, ds = gr.DevYuv(dgd), gr.DevYuv(src)
    rng = np.random.default_rng(9)
    for plane in (0, 1):
        U, PU = (unit, 64) if plane == 0 else (unit // 2, 32)
        off = 8 if plane == 0 else 4
        pw, ph = (w, h) if plane == 0 else ((w + 1) // 2, (h + 1) // 2)
        rects, y0 = [], 0
        while y0 < ph:  # foreach_rest_unit_in_tile, with the stripe offset of the vertical limits
            uh = ph - y0 if ph - y0 < U * 3 // 2 else U
            vs, ve = max(0, y0 - off), (y0 + uh - off if y0 + uh < ph else ph)
            x0 = 0
            while x0 < pw:
                uw = pw - x0 if pw - x0 < U * 3 // 2 else U
                rects.append((x0, x0 + uw, vs, ve))
                x0 += uw
            y0 += uh
        n_u, n_e = len(rects), len(eps)
        r = torch.tensor(rects, dtype=torch.int32, device="cuda")
        flt = torch.zeros(n_e * 2 * ph * pw, dtype=torch.int32, device="cuda")
        sums = torch.zeros(n_u * n_e * 5, dtype=torch.int64, device="cuda")
        mw, mh = max(a[1] - a[0] for a in rects), max(a[3] - a[2] for a in rects)
        a, b = dd.struct(), ds.struct()
        ep_arr = (C.c_int32 * n_e)(*eps)
        sb.check(lib.svt_b200_lr_sgr_filter_sums(C.byref(a), C.byref(b), plane, C.c_void_p(r.data_ptr()), n_u, mw, mh, ep_arr, n_e,
                                                 C.c_void_p(flt.data_ptr()), C.c_void_p(sums.data_ptr()), None), lib)
        xq = rng.integers(-60, 90, (n_u, n_e, 2)).astype(np.int32)
        dxq = torch.from_numpy(xq).cuda()
        err = torch.zeros(n_u * n_e, dtype=torch.int64, device="cuda")
        sb.check(lib.svt_b200_lr_sgr_proj_error(C.byref(a), C.byref(b), plane, C.c_void_p(r.data_ptr()), n_u, mw, mh, ep_arr, n_e,
                                                C.c_void_p(flt.data_ptr()), C.c_void_p(dxq.data_ptr()), C.c_void_p(err.data_ptr()), None), lib)
        torch.cuda.synchronize()
        g_flt = flt.cpu().numpy().reshape(n_e, 2, ph, pw)
        g_sums = sums.cpu().numpy().reshape(n_u, n_e, 5)
        g_err = err.cpu().numpy().reshape(n_u, n_e)
        ext = np.ascontiguousarray(np.pad(dgd.plane(plane), 8, mode="edge"))
        sp = src.plane(plane)
        hbd = bd > 8
        check_units = range(n_u) if n_u <= 12 else sorted(set([0, n_u - 1] + list(rng.choice(n_u, 8, replace=False))))
        for ui in check_units:
            hs, he, vs, ve = rects[ui]
            for ei, ep in enumerate(eps):
                f0, f1 = np.zeros((ve - vs, he - hs), np.int32), np.zeros((ve - vs, he - hs), np.int32)
                for i in range(0, ve - vs, PU):      # apply_sgr: processing units tiled from the unit's origin
                    for j in range(0, he - hs, PU):
                        ww, hh = min(PU, he - hs - j), min(PU, ve - vs - i)
                        t0, t1 = np.zeros((hh, ww), np.int32), np.zeros((hh, ww), np.int32)
                        p = ext.ctypes.data + ((vs + i + 8) * ext.shape[1] + hs + j + 8) * ext.itemsize
                        orc.orc_selfguided_restoration(C.c_void_p(p), int(hbd), ww, hh, ext.shape[1], cm.ptr(t0), cm.ptr(t1), ww, ep, bd)
                        f0[i:i + hh, j:j + ww], f1[i:i + hh, j:j + ww] = t0, t1
                r0, r1 = SGR_R[ep]
                if r0:
                    np.testing.assert_array_equal(g_flt[ei, 0, vs:ve, hs:he], f0, err_msg=f"flt0 plane {plane} unit {ui} ep {ep}")
                if r1:
                    np.testing.assert_array_equal(g_flt[ei, 1, vs:ve, hs:he], f1, err_msg=f"flt1 plane {plane} unit {ui} ep {ep}")
                d = dgd.plane(plane)[vs:ve, hs:he].astype(np.int64)
                s = sp[vs:ve, hs:he].astype(np.int64)
                u = d << 4
                a1 = (f0.astype(np.int64) - u) if r0 else np.zeros_like(u)
                a2 = (f1.astype(np.int64) - u) if r1 else np.zeros_like(u)
                sv = (s << 4) - u
                want = [int((a1 * a1).sum()), int((a2 * a2).sum()), int((a1 * a2).sum()), int((a1 * sv).sum()), int((a2 * sv).sum())]
                assert list(g_sums[ui, ei]) == want, (plane, ui, ep)
                assert g_err[ui, ei] == mo.pixel_proj_error(s, d, f0, f1, (int(xq[ui, ei, 0]), int(xq[ui, ei, 1])), (r0, r1), hbd)
