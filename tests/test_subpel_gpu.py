"""CUDA sub-pel refinement (svt-av1_b200/csrc/subpel.cu) through the C ABI against the oracle: best MV, cost, distortion
and sse of every job, bit-exact, for the configurations test_oracle_subpel.py pins against the reference, and a 1080p batch."""
import ctypes as C

import numpy as np
import pytest

import common as cm
import subpel_cases as sc
import svtb200 as sb
from test_oracle_subpel import CONFIGS

pytestmark = pytest.mark.gpu


def run_gpu(p, tabs, src, refs, jobs, max_block=None):
    import torch
    import gpu_runner as gr
    lib = sb.load()
    d_src, d_refs = gr.DevYuv(src), [gr.DevYuv(r) for r in refs]
    d_tabs = [torch.from_numpy(t).cuda() for t in tabs]
    q = sb.SubpelParams.from_buffer_copy(p)
    for i in range(2):
        q.mvcost[i] = d_tabs[i].data_ptr() + 4 * sc.MV_MAX
    if max_block:
        q.max_block_w, q.max_block_h = max_block
    d_jobs = torch.from_numpy(np.ascontiguousarray(jobs).view(np.uint8)).cuda()
    d_res = torch.zeros(len(jobs) * 16, dtype=torch.uint8, device="cuda")
    ss = d_src.struct()
    arr = (sb.Frame * len(refs))(*[r.struct() for r in d_refs])
    sb.check(lib.svt_b200_subpel_search(C.byref(q), C.byref(ss), arr, len(refs), C.c_void_p(d_jobs.data_ptr()), len(jobs),
                                        C.c_void_p(d_res.data_ptr()), None), lib)
    torch.cuda.synchronize()
    return d_res.cpu().numpy().view(sb.SUBPEL_RESULT_DTYPE)


@pytest.mark.parametrize("cfg", range(len(CONFIGS)))
def test_subpel_search_vs_oracle(cfg):
    w, h = 320, 192
    src, refs = sc.pictures(w, h, 40 + cfg)
    jobs = sc.make_jobs(w, h, len(refs), 132, 50 + cfg)
    p, tabs = sc.params(seed=cfg, **CONFIGS[cfg])
    want = sc.run_cpu(cm.oracle().orc_subpel_search, p, tabs, src, refs, jobs)
    got = run_gpu(p, tabs, src, refs, jobs)
    for f in ("mv_row", "mv_col", "besterr", "distortion", "sse"):
        np.testing.assert_array_equal(got[f], want[f], f)


def test_subpel_search_1080p_batch():
    """Every 16x16 .. 64x64 block position class of a 1080p picture, 2 references each (4000 searches), small-window launch."""
    w, h = 1920, 1080
    src, refs = sc.pictures(w, h, 77, n_refs=2)
    blocks = [b for b in sc.BLOCKS if max(b) <= 64]
    jobs = sc.make_jobs(w, h, len(refs), 4000, 78, blocks=blocks)
    p, tabs = sc.params(seed=5, search_type=3, iters=2, allow_hp=1)
    want = sc.run_cpu(cm.oracle().orc_subpel_search, p, tabs, src, refs, jobs)
    got = run_gpu(p, tabs, src, refs, jobs, max_block=(64, 64))
    for f in ("mv_row", "mv_col", "besterr", "distortion", "sse"):
        np.testing.assert_array_equal(got[f], want[f], f)
    again = run_gpu(p, tabs, src, refs, jobs)  # default window size (128x128): same answers
    for f in ("mv_row", "mv_col", "besterr"):
        np.testing.assert_array_equal(again[f], want[f], f)


@pytest.mark.parametrize("mx", [8, 16, 32])
def test_subpel_search_small_block_launches(mx):
    """max_block_w / _h select the CTA size (32 / 64 threads for small blocks): same answers."""
    w, h = 256, 128
    src, refs = sc.pictures(w, h, 90 + mx)
    blocks = [b for b in sc.BLOCKS if max(b) <= mx]
    jobs = sc.make_jobs(w, h, len(refs), 150, 91 + mx, blocks=blocks)
    p, tabs = sc.params(seed=mx, search_type=3, iters=2, allow_hp=1)
    want = sc.run_cpu(cm.oracle().orc_subpel_search, p, tabs, src, refs, jobs)
    got = run_gpu(p, tabs, src, refs, jobs, max_block=(mx, mx))
    for f in ("mv_row", "mv_col", "besterr", "distortion", "sse"):
        np.testing.assert_array_equal(got[f], want[f], f)


def test_variance_and_mse_dropins():
    import misc_oracle as mo
    from test_oracle_subpel import VAR_SIZES
    lib = sb.load()
    rng = np.random.default_rng(18)
    for (w, h) in VAR_SIZES:
        fn = getattr(lib, f"svt_aom_variance{w}x{h}_cuda")
        fn.restype = C.c_uint32
        for kind in ("rand", "extreme"):
            a = rng.integers(0, 256, (h, w + 7)).astype(np.uint8) if kind == "rand" else np.full((h, w + 7), 255, np.uint8)
            b = rng.integers(0, 256, (h, w + 3)).astype(np.uint8) if kind == "rand" else np.zeros((h, w + 3), np.uint8)
            sse = C.c_uint32(0)
            got = fn(cm.ptr(a), a.shape[1], cm.ptr(b), b.shape[1], C.byref(sse))
            assert (got, sse.value) == mo.variance(a[:, :w], b[:, :w]), (w, h, kind)
    lib.svt_aom_mse16x16_cuda.restype = C.c_uint32
    a, b = rng.integers(0, 256, (16, 20)).astype(np.uint8), rng.integers(0, 256, (16, 16)).astype(np.uint8)
    sse = C.c_uint32(0)
    got = lib.svt_aom_mse16x16_cuda(cm.ptr(a), 20, cm.ptr(b), 16, C.byref(sse))
    assert (got, sse.value) == mo.variance(a[:, :16], b)


@pytest.mark.parametrize("search", [1, 2, 3])
def test_upsampled_pred_dropin(search):
    lib = sb.load()
    rng = np.random.default_rng(30 + search)
    for (w, h) in [(4, 4), (8, 8), (16, 32), (64, 64), (128, 128), (32, 8)]:
        ref = rng.integers(0, 256, (h + 16, w + 16)).astype(np.uint8)
        for sx, sy in ((0, 0), (3, 0), (0, 5), (1, 1), (7, 4), (4, 7)):
            want, got = np.zeros((h, w), np.uint8), np.ones((h, w), np.uint8)
            r0 = C.c_void_p(ref.ctypes.data + 8 * ref.shape[1] + 8)
            cm.oracle().orc_upsampled_pred(cm.ptr(want), w, h, sx, sy, r0, ref.shape[1], search)
            lib.svt_aom_upsampled_pred_cuda(None, None, 0, 0, None, cm.ptr(got), w, h, sx, sy, r0, ref.shape[1], search)
            np.testing.assert_array_equal(got, want, f"{w}x{h} {sx},{sy}")
