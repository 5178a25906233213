"""The inter-prediction oracle (oracle/interp_oracle.c) pinned against the unmodified reference C code (oracle/_ref):
its kernel tables, each of the sixteen convolve functions, svt_aom_convolve8_*, and enc_make_inter_predictor on whole
pictures of jobs.  Sizes / phases / filters follow the reference's own convolve tests (test/convolve_2d_test.cc: every
block size, all 16 x 16 phases, every filter pair, bd 8 / 10 / 12, random + extreme inputs)."""
import ctypes as C

import numpy as np
import pytest

import common as cm
import interp_cases as ic
import svtb200 as sb

needs_ref = pytest.mark.skipif(not cm.have_ref(), reason="oracle/_ref not built")
SIZES = [(2, 2), (2, 4), (4, 2), (4, 4), (4, 8), (8, 4), (8, 8), (4, 16), (16, 4), (8, 16), (16, 8), (16, 16), (8, 32), (32, 8),
         (16, 32), (32, 16), (32, 32), (16, 64), (64, 16), (32, 64), (64, 32), (64, 64), (64, 128), (128, 64), (128, 128)]


@needs_ref
def test_kernel_tables_match_reference():
    orc, refh = cm.oracle(), cm.refh()
    for filt in range(4):
        for w in (2, 4, 8, 64):
            for sp in range(16):
                a, b = np.zeros(8, np.int16), np.ones(8, np.int16)
                orc.orc_interp_kernel(filt, w, sp, cm.ptr(a))
                refh.refh_interp_kernel(filt, w, sp, cm.ptr(b))
                np.testing.assert_array_equal(a, b, f"filter {filt} w {w} subpel {sp}")


def _block(rng, w, h, bd, kind):
    mx = (1 << bd) - 1
    shape = (h + 16, w + 16)
    if kind == "rand":
        v = rng.integers(0, mx + 1, shape)
    elif kind == "max":
        v = np.full(shape, mx)
    elif kind == "min":
        v = np.zeros(shape)
    else:
        v = rng.choice([0, mx], shape)
    return np.ascontiguousarray(v.astype(np.uint16 if bd > 8 else np.uint8))


def run_oracle_convolve(which, src, w, h, fx, fy, spx, spy, bd, r0, r1, do_avg, jnt, fwd, bck, conv, dst):
    orc = cm.oracle()
    hbd = int(bd > 8)
    sx, sy, comp = which >> 2 & 1, which >> 1 & 1, which & 1
    kx, ky = np.zeros(8, np.int16), np.zeros(8, np.int16)
    orc.orc_interp_kernel(fx, w, spx, cm.ptr(kx))
    orc.orc_interp_kernel(fy, h, spy, cm.ptr(ky))
    stride = src.shape[1]
    orc.orc_convolve(C.c_void_p(src.ctypes.data + (8 * stride + 8) * src.itemsize), hbd, stride, cm.ptr(dst), dst.shape[1], w, h,
                     cm.ptr(kx) if sx else None, cm.ptr(ky) if sy else None, r0, r1, bd, comp, do_avg, jnt, fwd, bck, cm.ptr(conv),
                     conv.shape[1])


def run_ref_convolve(which, src, w, h, fx, fy, spx, spy, bd, r0, r1, do_avg, jnt, fwd, bck, conv, dst):
    stride = src.shape[1]
    cm.refh().refh_convolve(which, int(bd > 8), C.c_void_p(src.ctypes.data + (8 * stride + 8) * src.itemsize), stride, cm.ptr(dst),
                            dst.shape[1], w, h, fx, fy, spx, spy, r0, r1, do_avg, jnt, fwd, bck, cm.ptr(conv), conv.shape[1], bd)


@needs_ref
@pytest.mark.parametrize("bd", [8, 10, 12])
@pytest.mark.parametrize("which", range(8))
def test_convolve_forms_vs_reference(which, bd):
    """which = sx*4 + sy*2 + compound: convolve[sx][sy][compound] (EbInterPrediction.c:1147-1175)."""
    rng = np.random.default_rng(1000 + which * 16 + bd)
    sx, sy, comp = which >> 2 & 1, which >> 1 & 1, which & 1
    r0, r1 = ic.conv_rounds(bd, comp)
    dt = np.uint16 if bd > 8 else np.uint8
    for n, (w, h) in enumerate(SIZES):
        for kind in ("rand", "extreme", "max", "min"):
            src = _block(rng, w, h, bd, kind)
            for trial in range(6 if kind == "rand" else 2):
                fx, fy = int(rng.integers(0, 4)), int(rng.integers(0, 4))
                spx = int(rng.integers(1, 16)) if sx else 0
                spy = int(rng.integers(1, 16)) if sy else 0
                jnt = int(rng.integers(0, 2))
                fwd, bck = ic.JNT_WEIGHTS[rng.integers(0, 8)]
                for do_avg in ((0, 1) if comp else (0,)):
                    conv0 = rng.integers(0, 1 << (bd + 5), (h, w + 3)).astype(np.uint16) if do_avg else np.zeros((h, w + 3), np.uint16)
                    ca, cb = conv0.copy(), conv0.copy()
                    da, db = np.full((h, w + 5), 7, dt), np.full((h, w + 5), 7, dt)
                    run_oracle_convolve(which, src, w, h, fx, fy, spx, spy, bd, r0, r1, do_avg, jnt, fwd, bck, ca, da)
                    run_ref_convolve(which, src, w, h, fx, fy, spx, spy, bd, r0, r1, do_avg, jnt, fwd, bck, cb, db)
                    np.testing.assert_array_equal(da, db, f"{w}x{h} {kind} f{fx}{fy} sp{spx},{spy} avg{do_avg}")
                    np.testing.assert_array_equal(ca, cb, f"conv {w}x{h} {kind}")
                    if kind == "rand" and w * h >= 16:  # the call did something
                        assert (ca != conv0).any() if comp and not do_avg else (da[:, :w] != 7).any()


@needs_ref
def test_all_phases_2d_vs_reference():
    rng = np.random.default_rng(5)
    w, h, bd = 16, 8, 8
    src = _block(rng, w, h, bd, "rand")
    for comp in (0, 1):
        r0, r1 = ic.conv_rounds(bd, comp)
        for spx in range(16):
            for spy in range(16):
                which = (spx != 0) * 4 + (spy != 0) * 2 + comp
                conv0 = rng.integers(0, 1 << 13, (h, w)).astype(np.uint16)
                ca, cb = conv0.copy(), conv0.copy()
                da, db = np.zeros((h, w), np.uint8), np.ones((h, w), np.uint8)
                run_oracle_convolve(which, src, w, h, 2, 0, spx, spy, bd, r0, r1, comp, 1, 9, 7, ca, da)
                run_ref_convolve(which, src, w, h, 2, 0, spx, spy, bd, r0, r1, comp, 1, 9, 7, cb, db)
                np.testing.assert_array_equal(da, db)


@needs_ref
@pytest.mark.parametrize("vert", [0, 1])
def test_convolve8_vs_reference(vert):
    orc, refh = cm.oracle(), cm.refh()
    rng = np.random.default_rng(77 + vert)
    for (w, h) in [(4, 4), (8, 8), (16, 32), (64, 64), (64, 16)]:
        for step in (16, 16, 24, 32, 11):
            filt, q0 = int(rng.integers(0, 4)), int(rng.integers(0, 16))
            span = ((max(w, h) - 1) * step + q0 >> 4) + 16
            src = rng.integers(0, 256, (span + 16, span + 16)).astype(np.uint8)
            table = np.zeros((16, 8), np.int16)
            for sp in range(16):
                orc.orc_interp_kernel(filt, 8, sp, C.c_void_p(table.ctypes.data + 16 * sp))
            a, b = np.zeros((h, w), np.uint8), np.ones((h, w), np.uint8)
            s0 = C.c_void_p(src.ctypes.data + 8 * src.shape[1] + 8)
            orc.orc_convolve8(s0, C.c_ssize_t(src.shape[1]), cm.ptr(a), C.c_ssize_t(w), cm.ptr(table), q0, step, w, h, vert)
            refh.refh_convolve8(s0, C.c_ssize_t(src.shape[1]), cm.ptr(b), C.c_ssize_t(w), filt, q0, step, w, h, vert)
            np.testing.assert_array_equal(a, b, f"{w}x{h} step {step}")


CASES = [(176, 144, 8, 64, 11), (200, 120, 10, 64, 12), (256, 128, 8, 128, 13), (128, 96, 12, 64, 15)]


@needs_ref
@pytest.mark.parametrize("w,h,bd,sb_size,seed", CASES)
def test_inter_predict_jobs_vs_reference(w, h, bd, sb_size, seed):
    """Whole pictures of jobs (all block shapes, sub8x8 chroma, compound + distance weights, MVs far outside the picture):
    the oracle against enc_make_inter_predictor."""
    refs = [ic.ref_picture(w, h, bd, seed * 10 + i, "texture" if i else "rand") for i in range(3)]
    jobs = ic.make_jobs(w, h, len(refs), seed, sb_size=sb_size)
    assert (jobs["n_refs"] == 2).any() and (jobs["bw"] == 2).any() and len(jobs) > 50
    a = ic.run_cpu(cm.oracle().orc_inter_predict, refs, cm.Yuv(w, h, bd, pad=ic.REF_PAD), jobs)
    b = ic.run_cpu(cm.refh().refh_inter_predict, refs, cm.Yuv(w, h, bd, pad=ic.REF_PAD), jobs)
    for i in range(3):
        np.testing.assert_array_equal(a.bufs[i], b.bufs[i], f"plane {i}")
    assert a.plane(0).any()
