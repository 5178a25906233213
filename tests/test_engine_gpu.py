"""The picture engine (svt_b200_engine_*: HOST pictures in, host results out - what the reference's process loops call
through integration/svt_cuda_backend.c) against the CPU oracle, without the encoder: the same seeded cases the device-pointer
entries are tested on, passed as plain numpy buffers."""
import ctypes as C

import numpy as np
import pytest

import common as cm
import svtb200 as sb

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine():
    lib = sb.load()
    e = C.c_void_p()
    sb.check(lib.svt_b200_engine_create(0, C.byref(e)), lib)
    yield lib, e
    lib.svt_b200_engine_destroy(e)


def _host_pic(key, tag, planes, with_decimations=True):
    h = sb.HostMePicture()
    h.key, h.tag, h.full = key, tag, planes[0].ctypes.data
    h.quarter = planes[1].ctypes.data if with_decimations else None
    h.sixteenth = planes[2].ctypes.data if with_decimations else None
    return h


def test_engine_me_picture_residency_and_results():
    """Three 'pictures' in coding order, each searched against the earlier ones: every call equals the oracle, and a picture
    uploaded once (as source) is found resident when it is a reference later (the engine's counters)."""
    lib = sb.load()
    e = C.c_void_p()
    sb.check(lib.svt_b200_engine_create(0, C.byref(e)), lib)
    try:
        w, h = 256, 192
        dist = ((1, 2, 3, 4), (1, 2, 3, 4))
        geos, src, refs = cm.make_me_case(w, h, 2, 2, seed=21)
        pics = [refs[1], refs[0], src, refs[4]]  # distinct contents; keys = their index
        n_sb = ((w + 63) // 64) * ((h + 63) // 64)
        for step, (cur, l0, l1) in enumerate([(1, [0], []), (2, [1, 0], []), (3, [2, 1], [0])]):
            params = sb.preset8_me_params(w, h, len(l0), len(l1), dist, 1, 1)
            want = cm.run_oracle_me(params, pics[cur], [pics[i] for i in l0] + [pics[0]] * (4 - len(l0)) + [pics[i] for i in l1] + [pics[0]] * (4 - len(l1)))
            hs = _host_pic(1000 + cur, cur, pics[cur])
            hr = (sb.HostMePicture * 8)()
            for j, i in enumerate(l0):
                hr[j] = _host_pic(1000 + i, i, pics[i])
            for j, i in enumerate(l1):
                hr[4 + j] = _host_pic(1000 + i, i, pics[i])
            got = cm.MeBuffers(n_sb)
            sb.check(lib.svt_b200_engine_me_picture(e, C.byref(params), C.byref(hs), hr, 1, cm.ptr(got.me_mv), cm.ptr(got.me_cand),
                                                    cm.ptr(got.total_cand), cm.ptr(got.rc)), lib)
            np.testing.assert_array_equal(got.me_mv, want.me_mv, err_msg=f"me_mv step {step}")
            np.testing.assert_array_equal(got.total_cand, want.total_cand)
            np.testing.assert_array_equal(got.rc, want.rc)
            for sbi in range(n_sb):  # candidates beyond total_cand are unspecified
                for pu in range(85):
                    n = int(want.total_cand[sbi, pu])
                    np.testing.assert_array_equal(got.me_cand[sbi, pu * 23:pu * 23 + n], want.me_cand[sbi, pu * 23:pu * 23 + n])
        st = sb.EngineStats()
        sb.check(lib.svt_b200_engine_get_stats(e, C.byref(st)), lib)
        assert st.me_pictures == 3
        assert st.me_plane_uploads == 4 and st.me_plane_hits == 5  # 4 distinct pictures; 1+2+... reference re-uses found resident
    finally:
        lib.svt_b200_engine_destroy(e)


def test_engine_me_device_side_downsample_matches_uploaded_planes():
    """quarter / sixteenth == NULL: the engine derives them with svt_b200_me_downsample (filtered) - same ME result as with the
    host's planes when those are the filtered downsamples of the full plane."""
    from test_oracle_me import downsample_case, run_downsample
    lib = sb.load()
    e = C.c_void_p()
    sb.check(lib.svt_b200_engine_create(0, C.byref(e)), lib)
    try:
        w, h = 192, 128
        dist = ((1, 2, 3, 4), (1, 2, 3, 4))
        pics = []
        for s in (3, 4):
            geos, full, q0, s0 = downsample_case(w, h, s)
            q, s16 = run_downsample(cm.oracle().orc_me_downsample, geos, full, q0.copy(), s0.copy(), 1)
            pics.append((full, q, s16))
        params = sb.preset8_me_params(w, h, 1, 0, dist, 1, 1)
        want = cm.run_oracle_me(params, pics[1], [pics[0]] * 8)
        n_sb = ((w + 63) // 64) * ((h + 63) // 64)
        hs = _host_pic(1, 1, pics[1], with_decimations=False)
        hr = (sb.HostMePicture * 8)()
        hr[0] = _host_pic(2, 0, pics[0], with_decimations=False)
        got = cm.MeBuffers(n_sb)
        sb.check(lib.svt_b200_engine_me_picture(e, C.byref(params), C.byref(hs), hr, 1, cm.ptr(got.me_mv), cm.ptr(got.me_cand),
                                                cm.ptr(got.total_cand), cm.ptr(got.rc)), lib)
        np.testing.assert_array_equal(got.me_mv, want.me_mv)
        np.testing.assert_array_equal(got.rc, want.rc)
    finally:
        lib.svt_b200_engine_destroy(e)


@pytest.mark.parametrize("case", [(320, 192, 8, 6, (12, 30, 0, 22), 2), (192, 136, 10, 2, (33, 17, 40, 25), 3)])
def test_engine_dlf_frame(engine, case):
    from test_dlf_gpu import flat_mi
    from test_oracle_dlf import dlf_case, dlf_params
    lib, _ = engine
    e = C.c_void_p()
    sb.check(lib.svt_b200_engine_create(0, C.byref(e)), lib)  # one geometry per engine
    try:
        w, h, bd, seed, levels, sharp = case
        mi_rows, mi_cols, part, frame = dlf_case(w, h, bd, seed, levels, sharp)
        flat = flat_mi(mi_rows, mi_cols, part, levels)
        p = dlf_params(mi_rows, mi_cols, levels, sharp)
        want = frame.copy()
        st = want.struct()
        cm.oracle().orc_dlf_frame(C.byref(p), C.byref(st), flat)
        got = frame.copy()
        gs = got.struct()
        sb.check(lib.svt_b200_engine_dlf_frame(e, C.byref(p), C.byref(gs), flat), lib)
        for i in range(3):
            np.testing.assert_array_equal(got.plane(i), want.plane(i), err_msg=f"plane {i}")
    finally:
        lib.svt_b200_engine_destroy(e)


@pytest.mark.parametrize("bd,with_dlf", [(8, False), (10, False), (8, True)])
def test_engine_cdef_frame_with_host_decision(bd, with_dlf):
    """search -> callback (the host's strength decision) -> apply, with the deblocking optionally deferred into the same call."""
    from test_dlf_gpu import flat_mi
    from test_oracle_cdef import cdef_picture_case
    from test_oracle_dlf import dlf_params
    lib = sb.load()
    e = C.c_void_p()
    sb.check(lib.svt_b200_engine_create(0, C.byref(e)), lib)
    try:
        w, h = 192, 136
        src, rec, mi_rows, mi_cols, skip = cdef_picture_case(w, h, bd)
        levels = (20, 24, 12, 9)
        pre = rec.copy()  # what the CDEF stages see
        dp, flat = None, None
        if with_dlf:
            part = cm.random_partition(mi_rows, mi_cols, 5)
            flat = flat_mi(mi_rows, mi_cols, part, levels)
            dp = dlf_params(mi_rows, mi_cols, levels, 1)
            ps = pre.struct()
            cm.oracle().orc_dlf_frame(C.byref(dp), C.byref(ps), flat)
        sp = sb.CdefSearchParams()
        sp.mi_rows, sp.mi_cols, sp.pri_damping = mi_rows, mi_cols, 5
        cm.oracle().orc_cdef_strength_table(3, C.byref(sp))
        nfb = ((mi_rows + 15) // 16) * ((mi_cols + 15) // 16)
        want_mse = np.zeros((2, nfb, 64), np.uint64)
        pres, ss = pre.struct(), src.struct()
        cm.oracle().orc_cdef_search(C.byref(sp), C.byref(pres), C.byref(ss), cm.ptr(skip), skip.shape[1], cm.ptr(want_mse))
        ys, uvs = (0, 5, 17, 63, 40, 2, 12, 33), (0, 0, 9, 62, 4, 1, 60, 3)
        idx = (np.arange(nfb) % 8).astype(np.int8)
        pa = sb.CdefApplyParams()
        pa.mi_rows, pa.mi_cols, pa.damping = mi_rows, mi_cols, 5
        for i in range(8):
            pa.y_strength[i], pa.uv_strength[i] = ys[i], uvs[i]
        want = pre.copy()
        ws = want.struct()
        cm.oracle().orc_cdef_apply(C.byref(pa), C.byref(pres), C.byref(ws), cm.ptr(skip), skip.shape[1], cm.ptr(idx))
        seen = {}

        def decide(user, mse, ap, fb_idx):
            seen["mse"] = np.ctypeslib.as_array(mse, shape=(2 * nfb * 64,)).copy().reshape(2, nfb, 64)
            ap.contents.damping = 5
            for i in range(8):
                ap.contents.y_strength[i], ap.contents.uv_strength[i] = ys[i], uvs[i]
            for i in range(nfb):
                fb_idx[i] = int(idx[i])
            return 1
        cb = sb.CDEF_DECIDE_FN(decide)
        got = rec.copy()
        gs = got.struct()
        mse = np.zeros((2, nfb, 64), np.uint64)
        sb.check(lib.svt_b200_engine_dlf_cdef_frame(e, C.byref(dp) if dp else None, flat, C.byref(sp), C.byref(gs), C.byref(ss), cm.ptr(skip),
                                                    skip.shape[1], cm.ptr(mse), cb, None), lib)
        np.testing.assert_array_equal(seen["mse"], want_mse)
        np.testing.assert_array_equal(mse, want_mse)
        for i in range(3):
            np.testing.assert_array_equal(got.plane(i), want.plane(i), err_msg=f"plane {i}")
        # "no apply": the host picture is left untouched
        again = rec.copy()
        ags = again.struct()
        cb0 = sb.CDEF_DECIDE_FN(lambda user, m, ap, fb: 0)
        sb.check(lib.svt_b200_engine_dlf_cdef_frame(e, C.byref(dp) if dp else None, flat, C.byref(sp), C.byref(ags), C.byref(ss), cm.ptr(skip),
                                                    skip.shape[1], cm.ptr(mse), cb0, None), lib)
        for i in range(3):
            np.testing.assert_array_equal(again.plane(i), rec.plane(i))
    finally:
        lib.svt_b200_engine_destroy(e)



@pytest.mark.parametrize("bd,with_dlf,pick", [(8, False, 3), (10, True, 2), (8, True, 3)])
def test_engine_cdef_frame_with_device_decision(bd, with_dlf, pick):
    """deblocking -> search -> svt_b200_cdef_decide (finish_cdef_search on the device) -> apply in ONE engine call: the
    decision, the per-filter-block choice and the filtered picture against the oracle chain."""
    from test_dlf_gpu import flat_mi
    from test_oracle_cdef import cdef_picture_case
    from test_oracle_dlf import dlf_params
    lib, orc = sb.load(), cm.oracle()
    e = C.c_void_p()
    sb.check(lib.svt_b200_engine_create(0, C.byref(e)), lib)
    try:
        w, h = 320, 200
        src, rec, mi_rows, mi_cols, skip = cdef_picture_case(w, h, bd)
        skip[0:8, 0:8] = 1  # one filter block entirely skipped
        levels = (20, 24, 12, 9)
        pre = rec.copy()
        dp, flat = None, None
        if with_dlf:
            part = cm.random_partition(mi_rows, mi_cols, 5)
            flat = flat_mi(mi_rows, mi_cols, part, levels)
            dp = dlf_params(mi_rows, mi_cols, levels, 1)
            ps = pre.struct()
            orc.orc_dlf_frame(C.byref(dp), C.byref(ps), flat)
        sp = sb.CdefSearchParams()
        sp.mi_rows, sp.mi_cols, sp.pri_damping = mi_rows, mi_cols, 5
        orc.orc_cdef_strength_table(pick, C.byref(sp))
        dcp = sb.CdefDecideParams()
        assert lib.svt_b200_cdef_decide_table(pick, C.byref(dcp)) == sp.n_strengths
        ref_tab = sb.CdefDecideParams()
        orc.orc_cdef_decide_table(pick, C.byref(ref_tab))
        assert list(dcp.filter_strength) == list(ref_tab.filter_strength)
        dcp.mi_rows, dcp.mi_cols, dcp.lambda_ = mi_rows, mi_cols, 3500
        nfb = ((mi_rows + 15) // 16) * ((mi_cols + 15) // 16)
        want_mse = np.zeros((2, nfb, 64), np.uint64)
        pres, ss = pre.struct(), src.struct()
        orc.orc_cdef_search(C.byref(sp), C.byref(pres), C.byref(ss), cm.ptr(skip), skip.shape[1], cm.ptr(want_mse))
        want_dec, want_idx = sb.CdefDecision(), np.zeros(nfb, np.int8)
        orc.orc_cdef_decide(C.byref(dcp), cm.ptr(want_mse), cm.ptr(skip), skip.shape[1], C.byref(want_dec), cm.ptr(want_idx))
        pa = sb.CdefApplyParams()
        pa.mi_rows, pa.mi_cols, pa.damping = mi_rows, mi_cols, 5
        for i in range(8):
            pa.y_strength[i], pa.uv_strength[i] = want_dec.y_strength[i], want_dec.uv_strength[i]
        want = pre.copy()
        ws = want.struct()
        orc.orc_cdef_apply(C.byref(pa), C.byref(pres), C.byref(ws), cm.ptr(skip), skip.shape[1], cm.ptr(want_idx))
        got, gs_dec, g_idx = rec.copy(), sb.CdefDecision(), np.full(nfb, 99, np.int8)
        gs = got.struct()
        sb.check(lib.svt_b200_engine_dlf_cdef_frame_dev(e, C.byref(dp) if dp else None, flat, C.byref(sp), C.byref(dcp), 5, 1, C.byref(gs), C.byref(ss),
                                                        cm.ptr(skip), skip.shape[1], C.byref(gs_dec), cm.ptr(g_idx)), lib)
        assert (gs_dec.cdef_bits, gs_dec.nb_cdef_strengths, gs_dec.sb_count) == (want_dec.cdef_bits, want_dec.nb_cdef_strengths, want_dec.sb_count)
        n = want_dec.nb_cdef_strengths
        assert list(gs_dec.y_strength)[:n] == list(want_dec.y_strength)[:n] and list(gs_dec.uv_strength)[:n] == list(want_dec.uv_strength)[:n]
        np.testing.assert_array_equal(g_idx, want_idx)
        assert want_idx[0] == -1 and (want_idx >= 0).any()
        for i in range(3):
            np.testing.assert_array_equal(got.plane(i), want.plane(i), err_msg=f"plane {i}")
        # apply = 0: decision only, the host picture untouched
        again, d2, i2 = rec.copy(), sb.CdefDecision(), np.zeros(nfb, np.int8)
        ags = again.struct()
        sb.check(lib.svt_b200_engine_dlf_cdef_frame_dev(e, C.byref(dp) if dp else None, flat, C.byref(sp), C.byref(dcp), 5, 0, C.byref(ags), C.byref(ss),
                                                        cm.ptr(skip), skip.shape[1], C.byref(d2), cm.ptr(i2)), lib)
        np.testing.assert_array_equal(i2, want_idx)
        for i in range(3):
            np.testing.assert_array_equal(again.plane(i), rec.plane(i))
    finally:
        lib.svt_b200_engine_destroy(e)


@pytest.mark.parametrize("case", [(192, 136, 8, 1, 0, 3, (20, 24, 12, 9)), (128, 128, 10, 3, 0, 3, (40, 40, 33, 20))])
def test_engine_dlf_pick_frame(case):
    """Level search + deblocking as one call: the levels equal the oracle's svt_av1_pick_filter_level, the picture equals the
    oracle's deblocking with those levels (sharpness 0)."""
    from test_dlf_gpu import flat_mi
    from test_oracle_dlf import dlf_params, pick_case, pick_params
    lib = sb.load()
    e = C.c_void_p()
    sb.check(lib.svt_b200_engine_create(0, C.byref(e)), lib)
    try:
        w, h, bd, seed, method, mode, last = case
        mi_rows, mi_cols, part, src, rec = pick_case(w, h, bd, seed)
        flat = flat_mi(mi_rows, mi_cols, part, last)
        p = pick_params(mi_rows, mi_cols, method, mode, last)
        r, t = rec.copy(), rec.copy()
        rs, ss, ts = r.struct(), src.struct(), t.struct()
        want_lv = (C.c_int32 * 4)()
        cm.oracle().orc_pick_filter_level(C.byref(p), C.byref(rs), C.byref(ss), C.byref(ts), flat, want_lv)
        got = rec.copy()
        gs = got.struct()
        lv = (C.c_int32 * 4)()
        sb.check(lib.svt_b200_engine_dlf_pick_frame(e, C.byref(p), C.byref(gs), C.byref(ss), flat, lv), lib)
        assert list(lv) == list(want_lv)
        flat2 = flat_mi(mi_rows, mi_cols, part, tuple(want_lv))
        dp = dlf_params(mi_rows, mi_cols, tuple(want_lv), 0)
        want = rec.copy()
        ws = want.struct()
        cm.oracle().orc_dlf_frame(C.byref(dp), C.byref(ws), flat2)
        for i in range(3):
            np.testing.assert_array_equal(got.plane(i), want.plane(i), err_msg=f"plane {i}")
    finally:
        lib.svt_b200_engine_destroy(e)


@pytest.mark.parametrize("case", [(192, 136, 8, 1, (64, 32, 32), ("mix", "mix", "mix"), (3, 3, 3)),
                                  (264, 200, 10, 3, (128, 64, 64), ("mix", "mix", "none"), (3, 3, 0))])
def test_engine_lr_frame_from_boundary_lines(case):
    """The restoration frame filter fed the way rest_kernel feeds it: the CDEF picture plus the two deblocked lines above /
    below every 64-row stripe (RestorationStripeBoundaries layout), not the whole deblocked picture."""
    from test_oracle_lr_frame import lr_case, run_oracle_lr
    lib = sb.load()
    e = C.c_void_p()
    sb.check(lib.svt_b200_engine_create(0, C.byref(e)), lib)
    try:
        w, h, bd, seed, unit_sizes, modes, ftypes = case
        cdef, dblk, units = lr_case(w, h, bd, seed, unit_sizes, modes)
        want = run_oracle_lr(cdef, dblk, units, unit_sizes, ftypes, 0)
        lines = (sb.HostLrLines * 3)()
        keep = []
        for pl in range(3):
            ss = 1 if pl else 0
            d = dblk.plane(pl)
            ph, pw = d.shape
            SH, off = 64 >> ss, 8 >> ss
            n_stripes = (ph + off + SH - 1) // SH
            stride = pw + 16
            above = np.zeros((2 * n_stripes, stride), d.dtype)
            below = np.zeros((2 * n_stripes, stride), d.dtype)
            for s in range(n_stripes):
                y0, y1 = max(0, s * SH - off), min((s + 1) * SH - off, ph)
                if s > 0:
                    above[2 * s, :pw], above[2 * s + 1, :pw] = d[y0 - 2], d[y0 - 1]
                if y1 < ph:
                    below[2 * s, :pw], below[2 * s + 1, :pw] = d[y1], d[min(y1 + 1, ph - 1)]
            keep += [above, below]
            lines[pl].above, lines[pl].below, lines[pl].stride = above.ctypes.data, below.ctypes.data, stride
        p = sb.LrFrameParams()
        for i in range(3):
            p.plane[i].frame_restoration_type, p.plane[i].restoration_unit_size = ftypes[i], unit_sizes[i]
            p.plane[i].units = C.addressof(units[i])
        p.optimized_lr = 0
        n_units = (C.c_int32 * 3)(*[len(u) for u in units])
        got = cdef.copy()
        gs = got.struct()
        sb.check(lib.svt_b200_engine_lr_frame(e, C.byref(p), n_units, C.byref(gs), lines), lib)
        for i in range(3):
            exp = want.plane(i) if ftypes[i] else cdef.plane(i)
            np.testing.assert_array_equal(got.plane(i), exp, err_msg=f"plane {i}")
    finally:
        lib.svt_b200_engine_destroy(e)
