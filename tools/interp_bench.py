#!/usr/bin/env python
"""Device time of svt_b200_inter_predict on a 1080p picture (random AV1 partition or a fixed block size)."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "svt-av1_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import common as cm  # noqa: E402
import gpu_runner as gr  # noqa: E402
import interp_cases as ic  # noqa: E402
import svtb200 as sb  # noqa: E402

W, H, BD = 1920, 1080, int(os.environ.get("BD", "8"))
lib = sb.load()
refs = [ic.ref_picture(W, H, BD, 900 + i) for i in range(4)]
jobs = ic.make_jobs(W, H, len(refs), 39)
d_refs = [gr.DevYuv(x) for x in refs]
d_pred = gr.DevYuv(cm.Yuv(W, H, BD, pad=ic.REF_PAD))
arr = (sb.Frame * len(refs))(*[x.struct() for x in d_refs])
ps = d_pred.struct()
lib.svt_b200_inter_predict_scratch_bytes.restype = C.c_size_t
sbytes = lib.svt_b200_inter_predict_scratch_bytes(len(jobs), W, H)
scratch = torch.zeros(sbytes, dtype=torch.uint8, device="cuda")
jobs8 = ic.make_jobs(W, H, len(refs), 40, min_n=8, square_only=True)
for name, jj in (("random partition", jobs), ("squares >= 8x8", jobs8), ("luma >= 16x16 only", jobs[(jobs["plane"] == 0) & (jobs["bw"] >= 16) & (jobs["bh"] >= 16)]),
                 ("<= 8x8 only", jobs[(jobs["bw"] <= 8) & (jobs["bh"] <= 8)])):
    dj = torch.from_numpy(np.ascontiguousarray(jj).view(np.uint8)).cuda()
    fn = lambda: sb.check(lib.svt_b200_inter_predict(arr, len(refs), C.byref(ps), C.c_void_p(dj.data_ptr()), len(jj), C.c_void_p(scratch.data_ptr()), C.c_size_t(sbytes), None), lib)  # noqa: E731
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        fn()
    e1.record()
    torch.cuda.synchronize()
    area = int((jj["bw"].astype(np.int64) * jj["bh"] * jj["n_refs"]).sum())
    print("%-22s %6d jobs  %8.1f us  (%.2f Msamples filtered)" % (name, len(jj), e0.elapsed_time(e1) / 20 * 1e3, area / 1e6))
