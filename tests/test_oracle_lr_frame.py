"""Frame-level loop restoration: the oracle (reference-style in-place stripe protocol) against the reference's own
svt_av1_loop_restoration_save_boundary_lines + svt_av1_loop_restoration_filter_frame (oracle/_ref)."""
import ctypes as C

import numpy as np
import pytest

import common as cm
import svtb200 as sb

needs_ref = pytest.mark.skipif(not cm.have_ref(), reason="oracle/_ref not built")


def count_units(unit, size):
    return max((size + (unit >> 1)) // unit, 1)


def random_units(rng, n, mode):
    """mode: 'mix' | 'wiener' | 'sgr' | 'none'.  Coefficients inside the ranges the bitstream can carry."""
    arr = (sb.LrUnit * n)()
    for u in arr:
        t = {"mix": int(rng.integers(0, 3)), "wiener": 1, "sgr": 2, "none": 0}[mode]
        u.restoration_type = t
        for f in (u.vfilter, u.hfilter):
            a, b, c = int(rng.integers(-5, 11)), int(rng.integers(-23, 9)), int(rng.integers(-17, 47))
            f[0], f[1], f[2], f[3], f[4], f[5], f[6], f[7] = a, b, c, -2 * (a + b + c), c, b, a, 0
        u.sgr_ep = int(rng.integers(0, 16))
        u.sgr_xqd[0], u.sgr_xqd[1] = int(rng.integers(-96, 32)), int(rng.integers(-32, 96))
    return arr


def lr_case(w, h, bd, seed, unit_sizes, modes, frame_types=(3, 3, 3)):
    rng = np.random.default_rng(seed)
    src = cm.synth_yuv(w, h, 1, seed, bd)
    cdef = cm.degrade(src, seed + 1, amp=9)
    dblk = cm.degrade(src, seed + 2, amp=13)
    units = []
    for p in range(3):
        pw, ph = (w, h) if p == 0 else ((w + 1) // 2, (h + 1) // 2)
        n = count_units(unit_sizes[p], pw) * count_units(unit_sizes[p], ph)
        units.append(random_units(rng, n, modes[p]))
    return cdef, dblk, units


def params(unit_sizes, frame_types, units_ptrs, optimized):
    p = sb.LrFrameParams()
    for i in range(3):
        p.plane[i].frame_restoration_type = frame_types[i]
        p.plane[i].restoration_unit_size = unit_sizes[i]
        p.plane[i].units = units_ptrs[i]
    p.optimized_lr = optimized
    return p


def run_oracle_lr(cdef, dblk, units, unit_sizes, frame_types, optimized):
    out = cdef.copy()
    for b in out.bufs:
        b[...] = 0
    p = params(unit_sizes, frame_types, [C.addressof(u) for u in units], optimized)
    cs, ds, os_ = cdef.struct(), dblk.struct(), out.struct()
    up = (C.c_void_p * 3)(*[C.addressof(u) for u in units])
    cm.oracle().orc_lr_frame(C.byref(p), C.byref(cs), C.byref(ds), C.byref(os_), up)
    return out


def run_ref_lr(cdef, dblk, units, unit_sizes, frame_types, optimized):
    work = cdef.copy()
    ws, ds = work.struct(), dblk.copy().struct()
    ft = (C.c_int32 * 3)(*frame_types)
    us = (C.c_int32 * 3)(*unit_sizes)
    up = (C.c_void_p * 3)(*[C.addressof(u) for u in units])
    assert cm.refh().refh_lr_frame(C.byref(ws), C.byref(ds), ft, us, up, optimized) == 0
    return work


LR_CASES = [(192, 136, 8, 1, (64, 32, 32), ("mix", "mix", "mix"), (3, 3, 3), 0),
            (192, 136, 10, 2, (64, 32, 32), ("wiener", "sgr", "mix"), (1, 2, 3), 0),
            (264, 200, 8, 3, (128, 64, 64), ("mix", "mix", "none"), (3, 3, 0), 0),
            (264, 200, 8, 4, (64, 64, 32), ("sgr", "wiener", "mix"), (2, 1, 3), 1),
            (328, 72, 10, 5, (256, 128, 128), ("mix", "mix", "mix"), (3, 3, 3), 0),
            (200, 328, 8, 6, (64, 32, 64), ("mix", "none", "wiener"), (3, 0, 1), 1)]


@needs_ref
@pytest.mark.parametrize("case", LR_CASES)
def test_lr_frame_matches_reference(case):
    w, h, bd, seed, unit_sizes, modes, frame_types, optimized = case
    cdef, dblk, units = lr_case(w, h, bd, seed, unit_sizes, modes)
    want = run_ref_lr(cdef, dblk, units, unit_sizes, frame_types, optimized)
    got = run_oracle_lr(cdef, dblk, units, unit_sizes, frame_types, optimized)
    for i in range(3):
        np.testing.assert_array_equal(got.plane(i), want.plane(i), err_msg=f"plane {i}")
    # sanity: an active plane really changes, a RESTORE_NONE plane does not
    for i in range(3):
        if frame_types[i] == 0 or all(u.restoration_type == 0 for u in units[i]):
            np.testing.assert_array_equal(got.plane(i), cdef.plane(i))
        else:
            assert (got.plane(i) != cdef.plane(i)).any()


@needs_ref
def test_unit_counts_match_reference():
    for w, h in ((192, 136), (1920, 1080), (328, 72), (64, 64), (100, 36)):
        for plane in (0, 1):
            for unit in (64, 128, 256):
                if plane and unit == 256:
                    continue
                hu, vu = C.c_int(), C.c_int()
                n = cm.refh().refh_lr_units(w, h, plane, unit, C.byref(hu), C.byref(vu))
                pw, ph = (w, h) if plane == 0 else ((w + 1) // 2, (h + 1) // 2)
                assert (hu.value, vu.value) == (count_units(unit, pw), count_units(unit, ph)) and n == hu.value * vu.value
