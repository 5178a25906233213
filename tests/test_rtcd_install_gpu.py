"""The drop-in boundary exercised by the reference itself: oracle/rtcd_install.c (the binding INTEGRATION.md section 1 describes,
compiled against the reference's headers) installs every *_cuda symbol into the RTCD function-pointer tables of the UNMODIFIED
reference library; the reference's own C loops then run on the GPU kernels, one host call per block exactly as its pipeline
threads would, and must produce what they produce on the C pointers."""
import ctypes as C
import os

import numpy as np
import pytest

import common as cm
import svtb200 as sb

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not cm.have_ref(), reason="oracle/_ref not built")]
CUDA_SO = os.path.join(cm.ORACLE_DIR, "_ref", "librefcuda.so")


@pytest.fixture()
def installed():
    """Yields a function run(fn) -> (result on the C pointers, result on the CUDA pointers, launches in between)."""
    if not os.path.exists(CUDA_SO):
        pytest.skip("oracle/_ref/librefcuda.so not built")
    cm.refh()  # loads the reference library RTLD_GLOBAL and sets the C pointers
    lib = sb.load()
    rc = C.CDLL(CUDA_SO)

    def run(fn):
        rc.svt_cuda_uninstall_rtcd()
        want = fn()
        n = rc.svt_cuda_install_rtcd()
        assert n == 225, n
        l0 = lib.svt_b200_launch_count()
        try:
            got = fn()
        finally:
            rc.svt_cuda_uninstall_rtcd()
        return want, got, lib.svt_b200_launch_count() - l0

    yield run
    rc.svt_cuda_uninstall_rtcd()


def _planes(f):
    return [f.plane(i).copy() for i in range(3)]


def test_reference_deblocking_frame_on_cuda_pointers(installed):
    from test_oracle_dlf import dlf_case, run_ref_dlf
    mi_rows, mi_cols, part, frame = dlf_case(64, 48, 8, 3, (20, 24, 12, 9), 0)
    want, got, launches = installed(lambda: _planes(run_ref_dlf(mi_rows, mi_cols, part, frame, (20, 24, 12, 9), 0)[0]))
    assert launches > 50
    for a, b in zip(got, want):
        np.testing.assert_array_equal(a, b)


def test_reference_cdef_search_and_apply_on_cuda_pointers(installed):
    from test_oracle_cdef import cdef_picture_case
    src, rec, mi_rows, mi_cols, skip = cdef_picture_case(64, 64, 8)
    nfb = ((mi_rows + 15) // 16) * ((mi_cols + 15) // 16)

    def go():
        mse = np.zeros((2, nfb, 64), np.uint64)
        rs, ss = rec.struct(), src.struct()
        cm.refh().refh_cdef_search(mi_rows, mi_cols, 172, 4, C.byref(rs), C.byref(ss), cm.ptr(skip), skip.shape[1], cm.ptr(mse))
        out = rec.copy()
        st = out.struct()
        idx = np.arange(nfb, dtype=np.int8) % 8
        ys, uvs = (C.c_int32 * 8)(0, 5, 17, 63, 40, 2, 12, 33), (C.c_int32 * 8)(0, 0, 9, 62, 4, 1, 60, 3)
        cm.refh().refh_cdef_apply(mi_rows, mi_cols, 5, ys, uvs, C.byref(st), cm.ptr(skip), skip.shape[1], cm.ptr(idx))
        return [mse] + _planes(out)

    want, got, launches = installed(go)
    assert launches > 50
    for a, b in zip(got, want):
        np.testing.assert_array_equal(a, b)


@pytest.mark.parametrize("case", [(2, 8, 0), (3, 10, 1), (12, 8, 0)])
def test_reference_encdec_chain_on_cuda_pointers(installed, case):
    from test_txfm_gpu import make_tus, quant_plane
    ts, bd, use_fp = case
    rng = np.random.default_rng(500 + ts)
    W, H = 128, 64
    src = cm.synth_yuv(W, H, 1, 7, bd)
    pred = cm.degrade(src, 11, amp=14)
    p = sb.EncodeParams()
    p.tx_size, p.use_fp = ts, use_fp
    for i in range(3):
        p.q[i] = quant_plane(rng, bd)
    tus = make_tus(rng, ts, W, H, limit=24)

    def go():
        rec = pred.copy()
        n = min(sb.TX_W[ts], 32) * min(sb.TX_H[ts], 32)
        q, eob = np.zeros((len(tus), n), np.int32), np.zeros(len(tus), np.uint16)
        arr = (sb.Tu * len(tus))(*tus)
        ss, ps, rs = src.struct(), pred.struct(), rec.struct()
        cm.refh().refh_encode_tus(C.byref(p), C.byref(ss), C.byref(ps), C.byref(rs), arr, len(tus), cm.ptr(q), cm.ptr(eob))
        return [q, eob] + _planes(rec)

    want, got, launches = installed(go)
    assert launches >= 4 * len(tus)
    for a, b in zip(got, want):
        np.testing.assert_array_equal(a, b)


def test_reference_restoration_frame_on_cuda_pointers(installed):
    from test_oracle_lr_frame import lr_case, run_ref_lr
    cdef, dblk, units = lr_case(128, 72, 8, 3, (64, 32, 32), ("mix", "mix", "mix"))
    want, got, launches = installed(lambda: _planes(run_ref_lr(cdef, dblk, units, (64, 32, 32), (3, 3, 3), 0)))
    assert launches > 3
    for a, b in zip(got, want):
        np.testing.assert_array_equal(a, b)


def test_reference_inter_prediction_on_cuda_pointers(installed):
    import interp_cases as ic
    w, h, bd = 96, 64, 8
    refs = [ic.ref_picture(w, h, bd, 300 + i) for i in range(2)]
    jobs = ic.make_jobs(w, h, len(refs), 17)
    want, got, launches = installed(lambda: _planes(ic.run_cpu(cm.refh().refh_inter_predict, refs, cm.Yuv(w, h, bd, pad=ic.REF_PAD), jobs)))
    assert launches >= len(jobs)
    for a, b in zip(got, want):
        np.testing.assert_array_equal(a, b)


def test_reference_subpel_search_on_cuda_pointers(installed):
    import subpel_cases as sc
    src, refs = sc.pictures(128, 96, 61)
    jobs = sc.make_jobs(128, 96, len(refs), 12, 62, blocks=[(8, 8), (16, 16), (32, 16), (4, 8)])
    p, tabs = sc.params(seed=2, search_type=3, iters=2, allow_hp=1)
    cm.refh().refh_subpel_search.restype = C.c_int
    want, got, launches = installed(lambda: sc.run_cpu(cm.refh().refh_subpel_search, p, tabs, src, refs, jobs))
    assert launches > 10 * len(jobs)
    np.testing.assert_array_equal(got, want)


def test_reference_motion_estimate_sb_on_cuda_pointers(installed):
    """The reference's own motion_estimate_sb (HME levels 0-2, integer search, pruning, candidate construction) over every SB
    of a small picture with the SAD pointers (svt_sad_loop_kernel, svt_ext_all_sad_calculation_8x8_16x16,
    svt_ext_eight_sad_calculation_32x32_64x64, svt_ext_sad_calculation_*, svt_nxm_sad_kernel, svt_initialize_buffer_32bits)
    replaced by the CUDA drop-ins: identical MeSbResults (ME was only checked oracle-side)."""
    w, h = 192, 128
    dist = ((1, 2, 3, 4), (1, 2, 3, 4))
    geos, src, refs = cm.make_me_case(w, h, 2, 1, seed=11)

    def run():
        _, out = cm.run_ref_me(w, h, 8, 2, 1, dist, 2, 1, geos, src, refs)
        return out
    want, got, launches = installed(run)
    assert launches > 100
    params = sb.preset8_me_params(w, h, 2, 1, dist, 2, 1)
    cm.assert_me_equal(got, want, params, "reference motion_estimate_sb: CUDA pointers vs C pointers")


@pytest.mark.parametrize("k", [0, 1, 4, 5])
def test_reference_temporal_filter_pointers_on_cuda(installed, k):
    """svt_av1_apply_temporal_filter_planewise / _hbd: the reference's RTCD pointers with the CUDA drop-ins installed give
    the accumulators of its own C functions (MeContext read by the shim in oracle/rtcd_install.c)."""
    import tf_cases as tc
    c = tc.make_case(**tc.CASES[k])
    want, got, launches = installed(lambda: tc.run_reference(cm.refh(), c, via_rtcd=1))
    assert launches >= 1
    for a, b in zip(got, want):
        np.testing.assert_array_equal(a, b)
