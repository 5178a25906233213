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
