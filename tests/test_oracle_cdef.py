"""Pins the CDEF part of the oracle against the unmodified reference C path (oracle/_ref).  Fixtures follow
test/CdefTest.cc (CDEFBlockTest :342, CDEFFindDirTest :447, compute_cdef_dist :535/:598): seeded random blocks
over all strengths/dampings/directions/bit depths, plus picture-level search/apply."""
import ctypes as C

import numpy as np
import pytest

import common as cm
import svtb200 as sb

needs_ref = pytest.mark.skipif(not cm.have_ref(), reason="oracle/_ref not built")


def rand_tile(rng, bd, mode):
    """CDEF_BSTRIDE(144) x (64+6) uint16 tile like the reference's test (values < 1<<bd, optional VERY_LARGE rim)."""
    t = rng.integers(0, 1 << bd, (70, 144)).astype(np.uint16)
    if mode == "smooth":
        base = int(rng.integers(0, 1 << bd))
        t = np.clip(base + rng.integers(-8, 9, (70, 144)), 0, (1 << bd) - 1).astype(np.uint16)
    if mode == "border":
        t[:3, :] = 16384
        t[:, :8] = 16384
    return t


def filter_block_cases():
    rng = np.random.default_rng(1)
    for bd in (8, 10, 12):
        for mode in ("random", "smooth", "border"):
            for bsize in (3, 0, 1, 2):
                for _ in range(6):
                    cs = bd - 8
                    pri = int(rng.integers(0, 16)) << cs
                    sec = int(rng.choice([0, 1, 2, 4])) << cs
                    yield bd, mode, bsize, pri, sec, int(rng.integers(0, 8)), int(rng.integers(3, 7)) + cs, rand_tile(rng, bd, mode)


@needs_ref
def test_filter_block_matches_reference():
    f_ref = C.cast(C.c_void_p.in_dll(cm.ref(), "svt_cdef_filter_block").value, C.CFUNCTYPE(None))
    for bd, mode, bsize, pri, sec, d, damp, tile in filter_block_cases():
        inp = tile.ctypes.data + (3 * 144 + 8) * 2
        outs = []
        for f in (f_ref, cm.oracle().orc_cdef_filter_block):
            d8, d16 = np.zeros((8, 8), np.uint8), np.zeros((8, 8), np.uint16)
            if bd == 8:
                f(cm.ptr(d8), None, 8, C.c_void_p(inp), pri, sec, d, damp, damp, bsize, bd - 8)
            else:
                f(None, cm.ptr(d16), 8, C.c_void_p(inp), pri, sec, d, damp, damp, bsize, bd - 8)
            outs.append((d8, d16))
        np.testing.assert_array_equal(outs[0][0], outs[1][0])
        np.testing.assert_array_equal(outs[0][1], outs[1][1])


@needs_ref
def test_find_dir_matches_reference():
    f_ref = C.cast(C.c_void_p.in_dll(cm.ref(), "svt_cdef_find_dir").value, C.CFUNCTYPE(C.c_int32))
    rng = np.random.default_rng(2)
    for bd in (8, 10, 12):
        for mode in ("random", "smooth"):
            for _ in range(20):
                tile = rand_tile(rng, bd, mode)
                va, vb = C.c_int32(0), C.c_int32(0)
                a = f_ref(cm.ptr(tile), 144, C.byref(va), bd - 8)
                b = cm.oracle().orc_cdef_find_dir(cm.ptr(tile), 144, C.byref(vb), bd - 8)
                assert (a, va.value) == (b, vb.value)


def cdef_picture_case(w, h, bd, seed=9):
    src = cm.synth_yuv(w, h, 1, seed, bd)
    rec = cm.degrade(src, seed)
    mi_rows, mi_cols = 2 * ((h + 7) // 8), 2 * ((w + 7) // 8)
    skip = cm.skip_map(mi_rows, mi_cols, seed)
    return src, rec, mi_rows, mi_cols, skip


@needs_ref
@pytest.mark.parametrize("case", [(192, 136, 8, 3), (192, 136, 10, 3), (200, 72, 8, 0), (136, 128, 10, 1)])
def test_cdef_search_matches_reference(case):
    w, h, bd, pick = case
    src, rec, mi_rows, mi_cols, skip = cdef_picture_case(w, h, bd)
    nfb = ((mi_rows + 15) // 16) * ((mi_cols + 15) // 16)
    base_q = 172
    want = np.zeros((2, nfb, 64), np.uint64)
    rs, ss = rec.struct(), src.struct()
    cm.refh().refh_cdef_search(mi_rows, mi_cols, base_q, {0: 1, 1: 2, 2: 3, 3: 4}[pick], C.byref(rs), C.byref(ss),
                               cm.ptr(skip), skip.shape[1], cm.ptr(want))
    p = sb.CdefSearchParams()
    p.mi_rows, p.mi_cols, p.pri_damping = mi_rows, mi_cols, 3 + (base_q >> 6)
    cm.oracle().orc_cdef_strength_table(pick, C.byref(p))
    got = np.zeros((2, nfb, 64), np.uint64)
    cm.oracle().orc_cdef_search(C.byref(p), C.byref(rs), C.byref(ss), cm.ptr(skip), skip.shape[1], cm.ptr(got))
    np.testing.assert_array_equal(got, want)
    assert want.any()


@needs_ref
@pytest.mark.parametrize("case", [(192, 136, 8), (192, 136, 10), (264, 72, 8)])
def test_cdef_apply_matches_reference(case):
    w, h, bd = case
    src, rec, mi_rows, mi_cols, skip = cdef_picture_case(w, h, bd, seed=21)
    nfb = ((mi_rows + 15) // 16) * ((mi_cols + 15) // 16)
    rng = np.random.default_rng(4)
    idx = rng.integers(0, 8, nfb).astype(np.int8)
    ys = (C.c_int32 * 8)(0, 5, 17, 63, 40, 2, 12, 33)
    uvs = (C.c_int32 * 8)(0, 0, 9, 62, 4, 1, 60, 3)
    inplace = rec.copy()
    st = inplace.struct()
    cm.refh().refh_cdef_apply(mi_rows, mi_cols, 5, ys, uvs, C.byref(st), cm.ptr(skip), skip.shape[1], cm.ptr(idx))
    p = sb.CdefApplyParams()
    p.mi_rows, p.mi_cols, p.damping = mi_rows, mi_cols, 5
    for i in range(8):
        p.y_strength[i], p.uv_strength[i] = ys[i], uvs[i]
    out = rec.copy()
    rs, os_ = rec.struct(), out.struct()
    cm.oracle().orc_cdef_apply(C.byref(p), C.byref(rs), C.byref(os_), cm.ptr(skip), skip.shape[1], cm.ptr(idx))
    for i in range(3):
        np.testing.assert_array_equal(out.plane(i), inplace.plane(i), err_msg=f"plane {i}")
    assert any((out.plane(i) != rec.plane(i)).any() for i in range(3))
