#!/usr/bin/env python
"""Device time of the picture-level in-loop filter entries, one picture per launch sequence, at the two bench geometries
(1920x1080 8-bit = configs[1], 3840x2160 10-bit = configs[2]), against the measured HBM peak.

For every entry: a RING of distinct device pictures whose total size exceeds the 126 MB L2 (64 pictures at 1080p, 8 at
2160p); one timed pass = one call per ring picture, captured into a CUDA graph and replayed (so the figure is device
time, not host launch latency), CUDA events on the launch stream around the replay, 5 passes after a warm-up pass;
deblocking works in place, so the ring is restored from a pristine copy (untimed) before every pass.  Reported:
algorithmic bytes / time and - as the ceiling a launch of this SIZE can reach - a plain device
copy of the same pictures through the same ring (read + write of the same bytes).  One JSON object on stdout.

    python tools/kernel_bench.py > gpurun_out/kernel_bench.json
"""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "svt-av1_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import common as cm  # noqa: E402
import gpu_runner as gr  # noqa: E402
import svtb200 as sb  # noqa: E402
from test_dlf_gpu import flat_mi  # noqa: E402
from test_oracle_dlf import dlf_case, dlf_params  # noqa: E402
from test_oracle_cdef import cdef_picture_case  # noqa: E402

lib = sb.load()
try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    PEAK = 6650.0


STREAM = torch.cuda.Stream()
SP = C.c_void_p(STREAM.cuda_stream)


def ring_ms(fn, ring, restore=None, passes=5):
    """fn(i) issues the call on ring picture i on STREAM.  One pass over the ring is captured into a CUDA graph and
    replayed, so the figure is device time (no host launch latency between the calls).  Mean device time per call."""
    with torch.cuda.stream(STREAM):
        for i in range(ring):
            fn(i)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=STREAM):
            for i in range(ring):
                fn(i)
        tot = 0.0
        for _ in range(passes):
            if restore:
                restore()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(STREAM)
            g.replay()
            e1.record(STREAM)
            torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
    return tot / (passes * ring)


def row(name, ms, alg_bytes, copy_ms, note=""):
    return {"entry": name, "ms": round(ms, 5), "algorithmic_bytes": int(alg_bytes), "achieved_gbs": round(alg_bytes / ms / 1e6, 1),
            "frac_of_hbm_peak": round(alg_bytes / ms / 1e6 / PEAK, 4), "same_size_copy_ms": round(copy_ms, 5),
            "frac_of_same_size_copy": round(copy_ms / ms, 3), "note": note}


def geometry(w, h, bd):
    out = {"geometry": "%dx%d %d-bit 4:2:0" % (w, h, bd), "rows": []}
    bps = 2 if bd > 8 else 1
    samples = w * h * 3 // 2
    pic_bytes = samples * bps
    ring = max(4, -(-(160 << 20) // pic_bytes))  # ring of pictures > L2 (126 MB)
    out["ring_pictures"] = ring
    # the ceiling for a launch of this size: device copy of one picture (reads + writes pic_bytes)
    a = [torch.empty(pic_bytes, dtype=torch.uint8, device="cuda") for _ in range(ring)]
    b = [torch.empty(pic_bytes, dtype=torch.uint8, device="cuda") for _ in range(ring)]
    copy_ms = ring_ms(lambda i: b[i].copy_(a[i]), ring)
    out["same_size_copy"] = {"ms": round(copy_ms, 5), "gbs": round(2 * pic_bytes / copy_ms / 1e6, 1),
                             "frac_of_hbm_peak": round(2 * pic_bytes / copy_ms / 1e6 / PEAK, 4)}
    del a, b

    # ---- deblocking --------------------------------------------------------------------------------------------
    levels = (24, 20, 14, 10)
    mi_rows, mi_cols, part, frame = dlf_case(w, h, bd, 9, levels, 1)
    flat = flat_mi(mi_rows, mi_cols, part, levels)
    p = dlf_params(mi_rows, mi_cols, levels, 1)
    pristine = gr.DevYuv(frame.copy())
    dfs = [gr.DevYuv(frame) for _ in range(ring)]
    dmis = [torch.from_numpy(np.frombuffer(flat, dtype=np.uint8).copy()).cuda() for _ in range(ring)]
    sts = [d.struct() for d in dfs]

    def restore():
        for d in dfs:
            for t, t0 in zip(d.t, pristine.t):
                t.copy_(t0)
    call = lambda i: sb.check(lib.svt_b200_dlf_frame(C.byref(p), C.byref(sts[i]), C.c_void_p(dmis[i].data_ptr()), SP), lib)  # noqa: E731
    ms = ring_ms(call, ring, restore)
    alg = 2 * (2 * pic_bytes) + 2 * mi_rows * mi_cols * 16
    out["rows"].append(row("svt_b200_dlf_frame (2 launches: all vertical edges, all horizontal edges)", ms, alg, copy_ms,
                           "2 passes x (read + write) + the 16 B / 4x4 mode-info summary per pass"))
    os.environ["SVT_B200_DLF_LINE_KERNEL"] = "1"
    ms_old = ring_ms(call, ring, restore)
    del os.environ["SVT_B200_DLF_LINE_KERNEL"]
    out["rows"].append(row("  (round-1 kernel: one thread per sample line, scalar accesses)", ms_old, alg, copy_ms))
    del dfs, dmis, pristine

    # ---- CDEF --------------------------------------------------------------------------------------------------
    src, rec, mi_rows, mi_cols, skip = cdef_picture_case(w, h, bd)
    sp = sb.CdefSearchParams()
    sp.mi_rows, sp.mi_cols, sp.pri_damping = mi_rows, mi_cols, 5
    cm.oracle().orc_cdef_strength_table(3, C.byref(sp))
    nfb = ((mi_rows + 15) // 16) * ((mi_cols + 15) // 16)
    drs, dss, dos = [gr.DevYuv(rec) for _ in range(ring)], [gr.DevYuv(src) for _ in range(ring)], [gr.DevYuv(rec) for _ in range(ring)]
    dskip = torch.from_numpy(skip).cuda()
    dmse = torch.zeros(2 * nfb * 64, dtype=torch.int64, device="cuda")
    rss, sss, oss = [d.struct() for d in drs], [d.struct() for d in dss], [d.struct() for d in dos]
    ms = ring_ms(lambda i: sb.check(lib.svt_b200_cdef_search(C.byref(sp), C.byref(rss[i]), C.byref(sss[i]), C.c_void_p(dskip.data_ptr()),
                                                             skip.shape[1], C.c_void_p(dmse.data_ptr()), SP), lib), ring)
    out["rows"].append(row("svt_b200_cdef_search (10 strengths, preset 8 table)", ms, 2 * pic_bytes + nfb * 2 * 64 * 8, copy_ms,
                           "integer-ALU bound: 10 filter evaluations per sample"))
    pa = sb.CdefApplyParams()
    pa.mi_rows, pa.mi_cols, pa.damping = mi_rows, mi_cols, 5
    for k, (x, y) in enumerate(zip((0, 5, 17, 63, 40, 2, 12, 33), (0, 0, 9, 62, 4, 1, 60, 3))):
        pa.y_strength[k], pa.uv_strength[k] = x, y
    didx = torch.from_numpy((np.arange(nfb) % 8).astype(np.int8)).cuda()
    ms = ring_ms(lambda i: sb.check(lib.svt_b200_cdef_apply(C.byref(pa), C.byref(rss[i]), C.byref(oss[i]), C.c_void_p(dskip.data_ptr()),
                                                            skip.shape[1], C.c_void_p(didx.data_ptr()), SP), lib), ring)
    out["rows"].append(row("svt_b200_cdef_apply (8 strength pairs cycled over the filter blocks)", ms, 2 * pic_bytes, copy_ms))
    return out


if __name__ == "__main__":
    res = {"hbm_peak_gbs": PEAK, "gpu": torch.cuda.get_device_name(0),
           "method": "ring of distinct pictures larger than L2, back-to-back calls, CUDA events around a pass, 5 passes after warm-up",
           "geometries": [geometry(1920, 1080, 8), geometry(3840, 2160, 10)]}
    print(json.dumps(res, indent=1))
