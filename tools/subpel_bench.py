#!/usr/bin/env python
"""One batched sub-pel search launch (every 16x16 block of a 1080p picture x 2 references) for profiling."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "svt-av1_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import gpu_runner as gr  # noqa: E402
import subpel_cases as sc  # noqa: E402
import svtb200 as sb  # noqa: E402

W, H = 1920, 1080
lib = sb.load()
src, refs = sc.pictures(W, H, 77, n_refs=2)
p, tabs = sc.params(seed=5, search_type=3, iters=2, allow_hp=1)
blocks = [(16, 16)] if len(sys.argv) < 2 else [(int(sys.argv[1]), int(sys.argv[1]))]
jobs = sc.make_jobs(W, H, 2, 16000, 78, blocks=blocks)
d_src, d_refs = gr.DevYuv(src), [gr.DevYuv(r) for r in refs]
d_tabs = [torch.from_numpy(t).cuda() for t in tabs]
for i in range(2):
    p.mvcost[i] = d_tabs[i].data_ptr() + 4 * sc.MV_MAX
p.max_block_w, p.max_block_h = blocks[0]
d_jobs = torch.from_numpy(jobs.view(np.uint8)).cuda()
d_res = torch.zeros(len(jobs) * 16, dtype=torch.uint8, device="cuda")
ss = d_src.struct()
arr = (sb.Frame * 2)(*[r.struct() for r in d_refs])
for _ in range(3):
    sb.check(lib.svt_b200_subpel_search(C.byref(p), C.byref(ss), arr, 2, C.c_void_p(d_jobs.data_ptr()), len(jobs), C.c_void_p(d_res.data_ptr()), None), lib)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    sb.check(lib.svt_b200_subpel_search(C.byref(p), C.byref(ss), arr, 2, C.c_void_p(d_jobs.data_ptr()), len(jobs), C.c_void_p(d_res.data_ptr()), None), lib)
e1.record()
torch.cuda.synchronize()
print("%dx%d blocks, %d searches: %.1f us" % (blocks[0][0], blocks[0][1], len(jobs), e0.elapsed_time(e1) / 5 * 1e3))
