"""CUDA inter prediction (svt-av1_b200/csrc/interp.cu) through the C ABI against the oracle, bit-exact: the sixteen
convolve drop-ins + svt_aom_convolve8_* (same sizes / phases / filters / extreme inputs as test_oracle_interp.py, which
pins the oracle against the reference), and svt_b200_inter_predict on whole pictures of jobs up to 1080p."""
import ctypes as C

import numpy as np
import pytest

import common as cm
import interp_cases as ic
import svtb200 as sb
from test_oracle_interp import SIZES, _block, run_oracle_convolve

pytestmark = pytest.mark.gpu

LO = ["svt_av1_convolve_2d_copy_sr", "svt_av1_jnt_convolve_2d_copy", "svt_av1_convolve_y_sr", "svt_av1_jnt_convolve_y",
      "svt_av1_convolve_x_sr", "svt_av1_jnt_convolve_x", "svt_av1_convolve_2d_sr", "svt_av1_jnt_convolve_2d"]
HI = [n.replace("svt_av1_", "svt_av1_highbd_") for n in LO]


def kernel_table(lib, filt, w):
    t = np.zeros((16, 8), np.int16)
    for sp in range(16):
        assert lib.svt_b200_get_interp_kernel(filt, w, sp, C.c_void_p(t.ctypes.data + 16 * sp)) == 0
    return t


def test_kernel_tables_vs_oracle():
    lib, orc = sb.load(), cm.oracle()
    for filt in range(4):
        for w in (2, 4, 8, 128):
            t = kernel_table(lib, filt, w)
            for sp in range(16):
                want = np.zeros(8, np.int16)
                orc.orc_interp_kernel(filt, w, sp, cm.ptr(want))
                np.testing.assert_array_equal(t[sp], want)


def run_gpu_convolve(lib, which, src, w, h, fx, fy, spx, spy, bd, r0, r1, do_avg, jnt, fwd, bck, conv, dst, tabs):
    hbd = bd > 8
    fn = getattr(lib, (HI if hbd else LO)[which] + "_cuda")
    tx, ty = tabs[(fx, w <= 4)], tabs[(fy, h <= 4)]
    px = sb.InterpFilterParams(tx.ctypes.data, 8, 16, fx)
    py = sb.InterpFilterParams(ty.ctypes.data, 8, 16, fy)
    cp = sb.ConvolveParams(0, do_avg, conv.ctypes.data, conv.shape[1], r0, r1, 0, which & 1, jnt, fwd, bck, jnt)
    stride = src.shape[1]
    args = [C.c_void_p(src.ctypes.data + (8 * stride + 8) * src.itemsize), stride, cm.ptr(dst), dst.shape[1], w, h, C.byref(px),
            C.byref(py), spx, spy, C.byref(cp)]
    if hbd:
        args.append(bd)
    fn(*args)


@pytest.mark.parametrize("bd", [8, 10, 12])
@pytest.mark.parametrize("which", range(8))
def test_convolve_dropins(which, bd):
    lib = sb.load()
    tabs = {(f, small): kernel_table(lib, f, 4 if small else 8) for f in range(4) for small in (False, True)}
    rng = np.random.default_rng(2000 + which * 16 + bd)
    sx, sy, comp = which >> 2 & 1, which >> 1 & 1, which & 1
    r0, r1 = ic.conv_rounds(bd, comp)
    dt = np.uint16 if bd > 8 else np.uint8
    for (w, h) in SIZES:
        for kind in ("rand", "extreme"):
            src = _block(rng, w, h, bd, kind)
            for trial in range(2):
                fx, fy = int(rng.integers(0, 4)), int(rng.integers(0, 4))
                spx = int(rng.integers(0, 16)) if sx else 0  # phase 0 through a filtering form: row 0 of the table
                spy = int(rng.integers(0, 16)) if sy else 0
                jnt = int(rng.integers(0, 2))
                fwd, bck = ic.JNT_WEIGHTS[rng.integers(0, 8)]
                for do_avg in ((0, 1) if comp else (0,)):
                    conv0 = rng.integers(0, 1 << (bd + 5), (h, w + 3)).astype(np.uint16) if do_avg else np.zeros((h, w + 3), np.uint16)
                    ca, cb = conv0.copy(), conv0.copy()
                    da, db = np.full((h, w + 5), 7, dt), np.full((h, w + 5), 7, dt)
                    run_oracle_convolve(which, src, w, h, fx, fy, spx, spy, bd, r0, r1, do_avg, jnt, fwd, bck, ca, da)
                    run_gpu_convolve(lib, which, src, w, h, fx, fy, spx, spy, bd, r0, r1, do_avg, jnt, fwd, bck, cb, db, tabs)
                    np.testing.assert_array_equal(db, da, f"{w}x{h} {kind} f{fx}{fy} sp{spx},{spy} avg{do_avg}")
                    np.testing.assert_array_equal(cb, ca, f"conv {w}x{h} {kind}")


@pytest.mark.parametrize("vert", [0, 1])
def test_convolve8_dropins(vert):
    lib, orc = sb.load(), cm.oracle()
    rng = np.random.default_rng(88 + vert)
    fn = lib.svt_aom_convolve8_vert_cuda if vert else lib.svt_aom_convolve8_horiz_cuda
    raw = np.zeros(16 * 8 + 256, np.int16)  # a 256-byte aligned [16][8] table, as get_filter_base requires
    off = (-raw.ctypes.data % 256) // 2
    table = raw[off:off + 128].reshape(16, 8)
    for (w, h) in [(4, 4), (8, 8), (16, 32), (64, 64), (64, 16), (128, 128)]:
        for step in (16, 16, 24, 32, 11):
            filt, q0 = int(rng.integers(0, 4)), int(rng.integers(0, 16))
            table[...] = kernel_table(lib, filt, 8)
            span = ((max(w, h) - 1) * step + q0 >> 4) + 16
            src = rng.integers(0, 256, (span + 16, span + 16)).astype(np.uint8)
            a, b = np.zeros((h, w), np.uint8), np.ones((h, w), np.uint8)
            s0 = C.c_void_p(src.ctypes.data + 8 * src.shape[1] + 8)
            orc.orc_convolve8(s0, C.c_ssize_t(src.shape[1]), cm.ptr(a), C.c_ssize_t(w), cm.ptr(table), q0, step, w, h, vert)
            f = C.c_void_p(table.ctypes.data + 16 * q0)
            fn(s0, C.c_ssize_t(src.shape[1]), cm.ptr(b), C.c_ssize_t(w), None if vert else f, 0 if vert else step, f if vert else None,
               step if vert else 0, w, h)
            np.testing.assert_array_equal(b, a, f"{w}x{h} step {step}")


CASES = [(176, 144, 8, 64, 21, "texture"), (200, 120, 10, 64, 22, "texture"), (256, 128, 8, 128, 23, "rand"),
         (128, 96, 12, 64, 24, "extreme"), (640, 360, 8, 64, 25, "texture")]


@pytest.mark.parametrize("w,h,bd,sb_size,seed,kind", CASES)
def test_inter_predict_vs_oracle(w, h, bd, sb_size, seed, kind):
    import gpu_runner as gr
    refs = [ic.ref_picture(w, h, bd, seed * 10 + i, kind) for i in range(3)]
    jobs = ic.make_jobs(w, h, len(refs), seed, sb_size=sb_size)
    want = ic.run_cpu(cm.oracle().orc_inter_predict, refs, cm.Yuv(w, h, bd, pad=ic.REF_PAD), jobs)
    got = gr.run_gpu_inter_predict(refs, cm.Yuv(w, h, bd, pad=ic.REF_PAD), jobs)
    for i in range(3):
        np.testing.assert_array_equal(got.bufs[i], want.bufs[i], f"plane {i}")
    assert want.plane(0).any()


@pytest.mark.parametrize("bd", [8, 10])
def test_inter_predict_1080p(bd):
    """BASELINE configs[1] geometry: every sample of the 1080p prediction, 4 reference pictures, vs the oracle."""
    import gpu_runner as gr
    w, h = 1920, 1080
    refs = [ic.ref_picture(w, h, bd, 900 + i) for i in range(4)]
    jobs = ic.make_jobs(w, h, len(refs), 31 + bd)
    want = ic.run_cpu(cm.oracle().orc_inter_predict, refs, cm.Yuv(w, h, bd, pad=ic.REF_PAD), jobs)
    got = gr.run_gpu_inter_predict(refs, cm.Yuv(w, h, bd, pad=ic.REF_PAD), jobs)
    for i in range(3):
        np.testing.assert_array_equal(got.bufs[i], want.bufs[i], f"plane {i}")
    # idempotence of the launch
    again = gr.run_gpu_inter_predict(refs, cm.Yuv(w, h, bd, pad=ic.REF_PAD), jobs)
    for i in range(3):
        np.testing.assert_array_equal(again.bufs[i], got.bufs[i])


@pytest.mark.parametrize("scratch_bytes", [256, 256 + 4 * 300, 256 + 4 * 2000])
def test_inter_predict_small_scratch(scratch_bytes):
    """A scratch too small for the tile list only moves work into the expanding warps: same picture."""
    import gpu_runner as gr
    w, h, bd = 320, 192, 8
    refs = [ic.ref_picture(w, h, bd, 70 + i) for i in range(2)]
    jobs = ic.make_jobs(w, h, len(refs), 5, sb_size=128)
    want = ic.run_cpu(cm.oracle().orc_inter_predict, refs, cm.Yuv(w, h, bd, pad=ic.REF_PAD), jobs)
    got = gr.run_gpu_inter_predict(refs, cm.Yuv(w, h, bd, pad=ic.REF_PAD), jobs, scratch_bytes=scratch_bytes)
    for i in range(3):
        np.testing.assert_array_equal(got.bufs[i], want.bufs[i], f"plane {i}")
