"""Golden digests: outputs of the UNMODIFIED reference (oracle/_ref) on seeded inputs, reduced to SHA-256 and committed in
tests/golden/reference_digests.json by tests/golden/make_golden.py; test_golden.py checks that the oracle reproduces every
digest, so the oracle stays pinned to the reference even on a machine where oracle/_ref cannot be built.
Each case is a function(which) -> bytes with which in {"ref", "oracle"}."""
import ctypes as C
import hashlib

import numpy as np

import common as cm
import svtb200 as sb


def _frame_bytes(f):
    return b"".join(np.ascontiguousarray(f.plane(i)).tobytes() for i in range(3))


def case_me(which):
    w, h, n0, n1, tl, isref, dist = 320, 192, 2, 1, 1, 1, ((1, 2, 3, 4), (1, 2, 3, 4))
    geos, src, refs = cm.make_me_case(w, h, n0, n1, seed=321)
    if which == "ref":
        params, out = cm.run_ref_me(w, h, 8, n0, n1, dist, tl, isref, geos, src, refs)
    else:
        params = sb.preset8_me_params(w, h, n0, n1, dist, tl, isref)
        out = cm.run_oracle_me(params, src, refs)
    f = out.fields()
    used = [f["best_sad"][:, l, r].tobytes() + f["best_mv"][:, l, r].tobytes() + f["hme"][:, l, r].tobytes()
            for l in range(params.num_lists) for r in range(params.num_refs[l])]
    return b"".join(used) + f["me_mv"].tobytes() + f["me_cand"].tobytes() + f["total_cand"].tobytes() + f["rc"].tobytes()


def case_encode(which):
    from test_txfm_gpu import make_tus, oracle_encode_tus, quant_plane
    out = b""
    for ts, bd, use_fp in ((2, 8, 0), (3, 10, 1), (12, 8, 0), (18, 10, 0)):
        rng = np.random.default_rng(1000 + ts)
        W, H = 192, 128
        src = cm.synth_yuv(W, H, 1, 7, bd)
        pred = cm.degrade(src, 11, amp=14)
        p = sb.EncodeParams()
        p.tx_size, p.use_fp = ts, use_fp
        for i in range(3):
            p.q[i] = quant_plane(rng, bd)
        tus = make_tus(rng, ts, W, H, limit=60)
        if which == "oracle":
            rec, q, eob = oracle_encode_tus(p, src, pred, tus, use_fp)
        else:
            rec = pred.copy()
            n = min(sb.TX_W[ts], 32) * min(sb.TX_H[ts], 32)
            q, eob = np.zeros((len(tus), n), np.int32), np.zeros(len(tus), np.uint16)
            arr = (sb.Tu * len(tus))(*tus)
            ss, ps, rs = src.struct(), pred.struct(), rec.struct()
            cm.refh().refh_encode_tus(C.byref(p), C.byref(ss), C.byref(ps), C.byref(rs), arr, len(tus), cm.ptr(q), cm.ptr(eob))
        out += np.ascontiguousarray(q).tobytes() + np.ascontiguousarray(eob).astype(np.uint16).tobytes() + _frame_bytes(rec)
    return out


def case_dlf(which):
    from test_oracle_dlf import dlf_case, dlf_params, run_ref_dlf
    from test_dlf_gpu import flat_mi
    out = b""
    for (w, h, bd, seed, levels, sharp) in ((192, 136, 8, 1, (20, 24, 12, 9), 0), (192, 136, 10, 2, (33, 17, 40, 25), 3)):
        mi_rows, mi_cols, part, frame = dlf_case(w, h, bd, seed, levels, sharp)
        if which == "ref":
            got, _ = run_ref_dlf(mi_rows, mi_cols, part, frame, levels, sharp)
        else:
            flat = flat_mi(mi_rows, mi_cols, part, levels)
            p = dlf_params(mi_rows, mi_cols, levels, sharp)
            got = frame.copy()
            st = got.struct()
            cm.oracle().orc_dlf_frame(C.byref(p), C.byref(st), flat)
        out += _frame_bytes(got)
    return out


def case_cdef(which):
    from test_oracle_cdef import cdef_picture_case
    out = b""
    for (w, h, bd, pick) in ((192, 136, 8, 3), (136, 128, 10, 1)):
        src, rec, mi_rows, mi_cols, skip = cdef_picture_case(w, h, bd)
        nfb = ((mi_rows + 15) // 16) * ((mi_cols + 15) // 16)
        mse = np.zeros((2, nfb, 64), np.uint64)
        rs, ss = rec.struct(), src.struct()
        if which == "ref":
            cm.refh().refh_cdef_search(mi_rows, mi_cols, 172, {0: 1, 1: 2, 2: 3, 3: 4}[pick], C.byref(rs), C.byref(ss), cm.ptr(skip),
                                       skip.shape[1], cm.ptr(mse))
        else:
            p = sb.CdefSearchParams()
            p.mi_rows, p.mi_cols, p.pri_damping = mi_rows, mi_cols, 3 + (172 >> 6)
            cm.oracle().orc_cdef_strength_table(pick, C.byref(p))
            cm.oracle().orc_cdef_search(C.byref(p), C.byref(rs), C.byref(ss), cm.ptr(skip), skip.shape[1], cm.ptr(mse))
        out += mse.tobytes()
        idx = np.random.default_rng(4).integers(0, 8, nfb).astype(np.int8)
        ys, uvs = (C.c_int32 * 8)(0, 5, 17, 63, 40, 2, 12, 33), (C.c_int32 * 8)(0, 0, 9, 62, 4, 1, 60, 3)
        if which == "ref":
            got = rec.copy()
            st = got.struct()
            cm.refh().refh_cdef_apply(mi_rows, mi_cols, 5, ys, uvs, C.byref(st), cm.ptr(skip), skip.shape[1], cm.ptr(idx))
        else:
            p = sb.CdefApplyParams()
            p.mi_rows, p.mi_cols, p.damping = mi_rows, mi_cols, 5
            for i in range(8):
                p.y_strength[i], p.uv_strength[i] = ys[i], uvs[i]
            got = rec.copy()
            os_ = got.struct()
            cm.oracle().orc_cdef_apply(C.byref(p), C.byref(rs), C.byref(os_), cm.ptr(skip), skip.shape[1], cm.ptr(idx))
        out += _frame_bytes(got)
    return out


def case_lr(which):
    from test_oracle_lr_frame import lr_case, run_oracle_lr, run_ref_lr
    out = b""
    for (w, h, bd, seed, unit_sizes, modes, opt) in ((200, 136, 8, 3, (64, 32, 32), ("mix", "mix", "mix"), 0),
                                                     (192, 144, 10, 4, (128, 64, 64), ("wiener", "sgr", "mix"), 1)):
        cdef, dblk, units = lr_case(w, h, bd, seed, unit_sizes, modes)
        got = (run_ref_lr if which == "ref" else run_oracle_lr)(cdef, dblk, units, unit_sizes, (3, 3, 3), opt)
        out += _frame_bytes(got)
    return out


def case_inter(which):
    import interp_cases as ic
    out = b""
    for (w, h, bd, sb_size, seed) in ((176, 144, 8, 64, 11), (128, 96, 12, 64, 15)):
        refs = [ic.ref_picture(w, h, bd, seed * 10 + i, "texture" if i else "rand") for i in range(3)]
        jobs = ic.make_jobs(w, h, len(refs), seed, sb_size=sb_size)
        fn = cm.refh().refh_inter_predict if which == "ref" else cm.oracle().orc_inter_predict
        got = ic.run_cpu(fn, refs, cm.Yuv(w, h, bd, pad=ic.REF_PAD), jobs)
        out += _frame_bytes(got)
    return out


def case_subpel(which):
    import subpel_cases as sc
    out = b""
    for cfg in (dict(search_type=3, iters=2, allow_hp=1), dict(search_type=2, iters=1, allow_hp=0), dict(search_type=1, iters=2, allow_hp=1, cost_type=1)):
        src, refs = sc.pictures(320, 192, 40)
        jobs = sc.make_jobs(320, 192, len(refs), 88, 50)
        p, tabs = sc.params(seed=1, **cfg)
        if which == "ref":
            cm.refh().refh_subpel_search.restype = C.c_int
        fn = cm.refh().refh_subpel_search if which == "ref" else cm.oracle().orc_subpel_search
        out += sc.run_cpu(fn, p, tabs, src, refs, jobs).tobytes()
    return out


def case_downsample(which):
    from test_oracle_me import downsample_case, run_downsample
    out = b""
    for (w, h, filt) in ((322, 182, 1), (176, 144, 0)):
        geos, full, q, s = downsample_case(w, h, 5)
        fn = cm.refh().refh_me_downsample if which == "ref" else cm.oracle().orc_me_downsample
        q, s = run_downsample(fn, geos, full, q, s, filt)
        out += q.tobytes() + s.tobytes()
    return out


def case_pick_filter_level(which):
    from test_dlf_gpu import flat_mi
    from test_oracle_dlf import PICK_CASES, pick_case, pick_params, run_ref_pick
    out = []
    for (w, h, bd, seed, method, mode, last, only4, deltas) in PICK_CASES[:4]:
        mi_rows, mi_cols, part, src, rec = pick_case(w, h, bd, seed)
        if which == "ref":
            lv, _ = run_ref_pick(mi_rows, mi_cols, part, src, rec, method, mode, last, only4, deltas=deltas)
        else:
            flat = flat_mi(mi_rows, mi_cols, part, last)
            p = pick_params(mi_rows, mi_cols, method, mode, last, only4, deltas=deltas)
            r, t = rec.copy(), rec.copy()
            rs, ss, ts = r.struct(), src.struct(), t.struct()
            got = (C.c_int32 * 4)()
            cm.oracle().orc_pick_filter_level(C.byref(p), C.byref(rs), C.byref(ss), C.byref(ts), flat, got)
            lv = list(got)
        out += [int(x) for x in lv]
    return np.array(out, np.int32).tobytes()


def case_convolve_forms(which):
    """The eight convolve forms x bd 8 / 10 on one block each (the sixteen reference functions by index)."""
    import interp_cases as ic
    from test_oracle_interp import _block, run_oracle_convolve, run_ref_convolve
    rng = np.random.default_rng(77)
    out = b""
    for bd in (8, 10):
        for form in range(8):
            sx, sy, comp = form >> 2 & 1, form >> 1 & 1, form & 1
            r0, r1 = ic.conv_rounds(bd, comp)
            w, h = 16, 8
            src = _block(rng, w, h, bd, "rand")
            conv = rng.integers(0, 1 << (bd + 5), (h, w)).astype(np.uint16)
            dst = np.zeros((h, w), np.uint16 if bd > 8 else np.uint8)
            fn = run_ref_convolve if which == "ref" else run_oracle_convolve
            fn(form, src, w, h, 2, 1, 5 if sx else 0, 11 if sy else 0, bd, r0, r1, comp, 1, 9, 7, conv, dst)
            out += dst.tobytes() + conv.tobytes()
    return out


def case_estimate_transform(which):
    """av1_estimate_transform's shape dispatcher: 19 sizes x 4 shapes x a few types (the coefficients the quantiser reads)."""
    from test_oracle_txfm import allowed_types, residual_block
    orc = cm.oracle()
    out = b""
    for ts in range(19):
        w, h = sb.TX_W[ts], sb.TX_H[ts]
        n = min(w, 32) * min(h, 32)
        rng = np.random.default_rng(900 + ts)
        for shape in range(4):
            for tx_type in allowed_types(ts)[:4]:
                res = residual_block(rng, w, h, 10, "rand")
                a = np.zeros(w * h, np.int32)
                if which == "oracle":
                    orc.orc_estimate_transform(cm.ptr(res), C.c_uint32(res.shape[1]), cm.ptr(a), ts, 10, tx_type, shape)
                else:
                    cm.refh().refh_estimate_transform(cm.ptr(res), C.c_uint32(res.shape[1]), cm.ptr(a), ts, 10, tx_type, shape)
                out += a[:n].tobytes()
    return out


def case_temporal_filter(which):
    import tf_cases as tc
    out = b""
    for kw in tc.CASES:
        c = tc.make_case(**kw)
        res = tc.run_oracle(cm.oracle(), c) if which == "oracle" else tc.run_reference(cm.refh(), c)
        out += b"".join(np.ascontiguousarray(a).tobytes() for a in res)
    return out


def case_picture_statistics(which):
    from test_oracle_pa import sb_planes
    orc = cm.oracle()
    out = b""
    rng = np.random.default_rng(40)
    for kind in ("rand", "flat", "extreme", "smooth", "rand"):
        y, cb, cr = sb_planes(rng, kind)
        oy, ox = int(rng.integers(0, 24)), int(rng.integers(0, 60))
        li, ci = oy * y.shape[1] + ox, (oy // 2) * cb.shape[1] + ox // 2
        ym, var, cbm, crm = np.zeros(85, np.uint8), np.zeros(85, np.uint16), np.zeros(85, np.uint8), np.zeros(85, np.uint8)
        if which == "oracle":
            orc.orc_sb_mean_variance(C.c_void_p(y.ctypes.data + li), y.shape[1], 0, cm.ptr(ym), cm.ptr(var))
            orc.orc_sb_chroma_mean(C.c_void_p(cb.ctypes.data + ci), cb.shape[1], 0, cm.ptr(cbm))
            orc.orc_sb_chroma_mean(C.c_void_p(cr.ctypes.data + ci), cr.shape[1], 0, cm.ptr(crm))
        else:
            cm.refh().refh_sb_mean_variance(cm.ptr(y), y.shape[1], li, cm.ptr(cb), cm.ptr(cr), cb.shape[1], ci, 0, cm.ptr(ym), cm.ptr(var), cm.ptr(cbm),
                                            cm.ptr(crm))
        out += ym.tobytes() + var.tobytes() + cbm[:21].tobytes() + crm[:21].tobytes()
    return out


def case_open_loop_intra(which):
    from test_oracle_ois import padded_luma
    out = b""
    for (w, h) in ((200, 120), (352, 288), (72, 88)):
        buf, pad = padded_luma(w, h, 60 + w)
        mbw, mbh = (w + 15) // 16, (h + 15) // 16
        cost, mode = np.zeros(mbw * mbh, np.int64), np.zeros(mbw * mbh, np.int32)
        if which == "oracle":
            cm.oracle().orc_ois_dc_picture(C.c_void_p(buf.ctypes.data + pad * buf.shape[1] + pad), buf.shape[1], w, h, cm.ptr(cost))
        else:
            cm.refh().refh_ois_picture(cm.ptr(buf), buf.shape[1], pad, pad, w, h, 8, cm.ptr(cost), cm.ptr(mode))
        out += cost.tobytes()
    return out


# full_lambda the reference derives (compute_rdmult_sse) for the harness's picture control set at qindex = 4 * key, 8 bit
GOLDEN_LAMBDAS = {43: 257491, 20: 20078, 50: 554843, 30: 55473, 35: 101038}


def case_cdef_decide(which):
    """finish_cdef_search on synthetic mse tables; the lambda of each case is a fixed number recorded with the reference run
    (it only depends on qindex / bit depth for the zeroed picture control set of the harness)."""
    from test_oracle_cdef_decide import decide_case
    LAMBDAS = GOLDEN_LAMBDAS
    out = b""
    for seed, pick, mi_rows, mi_cols, qidx, kind in ((0, 3, 68, 120, 43, "smooth"), (1, 3, 68, 120, 20, "rand"), (2, 2, 45, 80, 50, "smooth"),
                                                      (3, 1, 45, 80, 30, "ties"), (4, 0, 34, 46, 35, "smooth")):
        p = sb.CdefDecideParams()
        n = cm.oracle().orc_cdef_decide_table(pick, C.byref(p))
        p.mi_rows, p.mi_cols = mi_rows, mi_cols
        skip, stride, mse = decide_case(seed, mi_rows, mi_cols, n, kind)
        nfb = mse.shape[1]
        fbs = np.zeros(nfb, np.int8)
        if which == "oracle":
            p.lambda_ = LAMBDAS[qidx]
            o = sb.CdefDecision()
            cm.oracle().orc_cdef_decide(C.byref(p), cm.ptr(mse), cm.ptr(skip), stride, C.byref(o), cm.ptr(fbs))
            bits, nb, ys, uvs = o.cdef_bits, o.nb_cdef_strengths, list(o.y_strength), list(o.uv_strength)
        else:
            b_, n_, lam = C.c_int32(), C.c_int32(), C.c_uint64()
            y8, uv8 = (C.c_int32 * 8)(), (C.c_int32 * 8)()
            cm.refh().refh_cdef_finish(mi_rows, mi_cols, qidx * 4, {0: 1, 1: 2, 2: 3, 3: 4}[pick], 8, cm.ptr(skip), stride, cm.ptr(mse), C.byref(b_),
                                       C.byref(n_), y8, uv8, cm.ptr(fbs), C.byref(lam))
            assert lam.value == LAMBDAS[qidx], (qidx, lam.value)
            bits, nb, ys, uvs = b_.value, n_.value, list(y8), list(uv8)
        out += bytes([bits, nb]) + np.array(ys[:nb] + uvs[:nb], np.int32).tobytes() + fbs.tobytes()
    return out


CASES = {"me_picture": case_me, "encode_tus": case_encode, "dlf_frame": case_dlf, "cdef_search_apply": case_cdef, "lr_frame": case_lr,
         "inter_predict": case_inter, "subpel_search": case_subpel, "me_downsample": case_downsample,
         "pick_filter_level": case_pick_filter_level, "convolve_forms": case_convolve_forms, "estimate_transform_shapes": case_estimate_transform,
         "temporal_filter_planewise": case_temporal_filter, "picture_statistics": case_picture_statistics, "open_loop_intra_dc": case_open_loop_intra,
         "cdef_decide": case_cdef_decide}


def digest(name, which):
    return hashlib.sha256(CASES[name](which)).hexdigest()
