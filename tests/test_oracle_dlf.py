"""Pins the deblocking part of the oracle against the unmodified reference C path (oracle/_ref): the sixteen
svt_aom_[highbd_]lpf_{h,v}_{4,6,8,14} kernels on test/DeblockTest.cc-style random edges, and the whole-frame loop
(which also proves that 'all vertical edges, then all horizontal edges' equals the reference's lagged SB order)."""
import ctypes as C

import numpy as np
import pytest

import common as cm
import svtb200 as sb

needs_ref = pytest.mark.skipif(not cm.have_ref(), reason="oracle/_ref not built")


def edge_cases(seed=0, n=40):
    rng = np.random.default_rng(seed)
    for it in range(n):
        bd = [8, 10, 12][it % 3]
        base = int(rng.integers(0, 1 << bd))
        kind = it % 4
        if kind == 0:
            img = rng.integers(0, 1 << bd, (32, 32))
        elif kind == 1:   # nearly flat: exercises the flat / flat2 paths
            img = np.clip(base + rng.integers(-1, 2, (32, 32)) * (1 << (bd - 8)), 0, (1 << bd) - 1)
        elif kind == 2:   # step edge
            img = np.clip(base + rng.integers(-3, 4, (32, 32)), 0, (1 << bd) - 1)
            img[:, 16:] = np.clip(img[:, 16:] + int(rng.integers(-40, 40)) * (1 << (bd - 8)), 0, (1 << bd) - 1)
            img[16:, :] = np.clip(img[16:, :] + int(rng.integers(-40, 40)) * (1 << (bd - 8)), 0, (1 << bd) - 1)
        else:
            img = np.clip(base + rng.integers(-12, 13, (32, 32)), 0, (1 << bd) - 1)
        lvl = int(rng.integers(0, 64))
        sharp = int(rng.integers(0, 8))
        lim = lvl >> ((sharp > 0) + (sharp > 4))
        if sharp > 0:
            lim = min(lim, 9 - sharp)
        lim = max(lim, 1)
        yield bd, img, 2 * (lvl + 2) + lim, lim, lvl >> 4


@needs_ref
@pytest.mark.parametrize("length", [4, 6, 8, 14])
@pytest.mark.parametrize("direction", ["horizontal", "vertical"])
def test_lpf_kernels_match_reference(length, direction):
    ref, orc = cm.ref(), cm.oracle()
    for bd, img, blimit, limit, thresh in edge_cases(length):
        hbd = bd > 8
        name = f"svt_aom_{'highbd_' if hbd else ''}lpf_{direction}_{length}"
        f = C.cast(C.c_void_p.in_dll(ref, name).value, C.CFUNCTYPE(None))
        a = img.astype(np.uint16 if hbd else np.uint8)
        b = a.copy()
        off = (16 * 32 + 16) * a.itemsize
        bl, li, th = (np.full(16, v, np.uint8) for v in (blimit & 255, limit, thresh))
        if hbd:
            f(C.c_void_p(a.ctypes.data + off), 32, cm.ptr(bl), cm.ptr(li), cm.ptr(th), bd)
        else:
            f(C.c_void_p(a.ctypes.data + off), 32, cm.ptr(bl), cm.ptr(li), cm.ptr(th))
        vert = direction == "vertical"
        orc.orc_lpf_edge(C.c_void_p(b.ctypes.data + off), int(hbd), 1 if vert else 32, 32 if vert else 1, length,
                         blimit & 255, limit, thresh, bd)
        np.testing.assert_array_equal(a, b, err_msg=f"{name} bd{bd}")


def dlf_case(w, h, bd, seed, levels, sharp):
    mi_rows, mi_cols = h // 4, w // 4
    part = cm.random_partition(mi_rows, mi_cols, seed)
    frame = cm.degrade(cm.synth_yuv(w, h, 1, seed, bd), seed, amp=12)
    return mi_rows, mi_cols, part, frame


def run_ref_dlf(mi_rows, mi_cols, part, frame, levels, sharp):
    flat = (sb.DlfMi * (mi_rows * mi_cols))()
    out = frame.copy()
    st = out.struct()
    lv = (C.c_int32 * 4)(*levels)
    sbt, dep, inter, skip = (np.ascontiguousarray(x) for x in part)
    cm.refh().refh_dlf_frame(mi_rows, mi_cols, cm.ptr(sbt), cm.ptr(dep), cm.ptr(inter), cm.ptr(skip), lv, sharp, C.byref(st), flat)
    return out, flat


def dlf_params(mi_rows, mi_cols, levels, sharp):
    p = sb.DlfParams()
    p.mi_rows, p.mi_cols, p.mi_stride, p.sharpness = mi_rows, mi_cols, mi_cols, sharp
    p.filter_level[0], p.filter_level[1], p.filter_level_u, p.filter_level_v = levels
    p.plane_start, p.plane_end = 0, 3
    return p


DLF_CASES = [(192, 136, 8, 1, (20, 24, 12, 9), 0), (192, 136, 10, 2, (33, 17, 40, 25), 3), (264, 72, 8, 3, (63, 63, 63, 63), 7),
             (128, 128, 10, 4, (8, 0, 0, 5), 5), (136, 200, 8, 5, (0, 0, 30, 30), 0), (320, 192, 8, 6, (12, 30, 0, 22), 2)]


@needs_ref
@pytest.mark.parametrize("case", DLF_CASES)
def test_dlf_frame_matches_reference(case):
    w, h, bd, seed, levels, sharp = case
    mi_rows, mi_cols, part, frame = dlf_case(w, h, bd, seed, levels, sharp)
    want, flat = run_ref_dlf(mi_rows, mi_cols, part, frame, levels, sharp)
    got = frame.copy()
    st = got.struct()
    p = dlf_params(mi_rows, mi_cols, levels, sharp)
    cm.oracle().orc_dlf_frame(C.byref(p), C.byref(st), flat)
    for i in range(3):
        np.testing.assert_array_equal(got.plane(i), want.plane(i), err_msg=f"plane {i}")
    if levels[0] >= 20:  # (weak levels on noisy content may legitimately change nothing)
        assert any((want.plane(i) != frame.plane(i)).any() for i in range(3))
