#!/usr/bin/env python
"""Device time of the §8 rows that are not part of the preset-8 headline step (loop restoration, deblocking level search),
1080p 8-bit, next to the reference C implementation of the same call on the host (oracle/_ref, single thread).
One JSON object on stdout; run on the GPU box: python tools/bench_extra.py > gpurun_out/extra.json"""
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "svt-av1_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import common as cm  # noqa: E402
import gpu_runner as gr  # noqa: E402
import svtb200 as sb  # noqa: E402
from test_oracle_lr_frame import lr_case, run_ref_lr  # noqa: E402
from test_oracle_dlf import pick_case, pick_params, run_ref_pick  # noqa: E402
from test_dlf_gpu import flat_mi  # noqa: E402

W, H, BD = 1920, 1080, 8
lib = sb.load()
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0


def gpu_ms(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


out = {"geometry": f"{W}x{H} {BD}-bit 4:2:0", "hbm_peak_gbs": peak, "rows": []}
samples = W * H * 3 // 2

# ---- svt_av1_loop_restoration_filter_frame ------------------------------------------------------------------------
unit_sizes, ftypes = (64, 32, 32), (3, 3, 3)
cdef, dblk, units = lr_case(W, H, BD, 7, unit_sizes, ("mix", "mix", "mix"))
dc, dd = gr.DevYuv(cdef), gr.DevYuv(dblk)
do = gr.DevYuv(cdef.copy())
dus = [torch.from_numpy(np.frombuffer(u, dtype=np.uint8).copy()).cuda() for u in units]
p = sb.LrFrameParams()
for i in range(3):
    p.plane[i].frame_restoration_type, p.plane[i].restoration_unit_size, p.plane[i].units = ftypes[i], unit_sizes[i], dus[i].data_ptr()
cs, ds, os_ = dc.struct(), dd.struct(), do.struct()
ms = gpu_ms(lambda: sb.check(lib.svt_b200_lr_frame(C.byref(p), C.byref(cs), C.byref(ds), C.byref(os_), None), lib))
t0 = time.perf_counter()
run_ref_lr(cdef, dblk, units, unit_sizes, ftypes, 0)
ref_ms = (time.perf_counter() - t0) * 1e3
alg = samples * 2
out["rows"].append({"entry": "svt_b200_lr_frame (units 64/32, Wiener/SGR/none mixed)", "gpu_ms": ms, "algorithmic_bytes": alg,
                    "achieved_gbs": alg / ms / 1e6, "hbm_frac": alg / ms / 1e6 / peak, "reference_c_ms_1thread": ref_ms})

# ---- search_wiener's compute_stats for the whole luma plane --------------------------------------------------------
src = cm.synth_yuv(W, H, 1, 21, BD)
dgd = cm.degrade(src, 22, amp=8)
dg, dsrc = gr.DevYuv(dgd), gr.DevYuv(src)
rects = [(x, min(x + 64, W), y, min(y + 64, H)) for y in range(0, H, 64) for x in range(0, W, 64)]
r = torch.tensor(rects, dtype=torch.int32, device="cuda")
stats = torch.zeros(len(rects) * (49 + 49 * 49), dtype=torch.int64, device="cuda")
scr = torch.zeros(len(rects), dtype=torch.int64, device="cuda")
a, b = dg.struct(), dsrc.struct()
ms = gpu_ms(lambda: sb.check(lib.svt_b200_lr_wiener_stats(C.byref(a), C.byref(b), 0, 7, C.c_void_p(r.data_ptr()), len(rects), 64, 64,
                                                          C.c_void_p(stats.data_ptr()), C.c_void_p(scr.data_ptr()), None), lib))
ref = cm.refh()
ext, sext = np.pad(dgd.plane(0), 3, mode="edge"), np.pad(src.plane(0), 3, mode="edge")
M, Hm = np.zeros(49, np.int64), np.zeros(49 * 49, np.int64)
t0 = time.perf_counter()
for (hs, he, vs, ve) in rects[::8]:
    ref.svt_av1_compute_stats_c(7, cm.ptr(ext), cm.ptr(sext), hs + 3, he + 3, vs + 3, ve + 3, ext.shape[1], sext.shape[1], cm.ptr(M), cm.ptr(Hm))
ref_ms = (time.perf_counter() - t0) * 1e3 * 8
macs = W * H * (49 * 50 // 2 + 49)
out["rows"].append({"entry": "svt_b200_lr_wiener_stats (luma, win 7, 510 units)", "gpu_ms": ms, "int_macs": macs, "achieved_tmacs": macs / ms / 1e9,
                    "reference_c_ms_1thread": ref_ms})

# ---- self-guided search: all 16 parameter sets, luma -----------------------------------------------------------------
eps = (C.c_int32 * 16)(*range(16))
flt = torch.zeros(16 * 2 * W * H, dtype=torch.int32, device="cuda")
sums = torch.zeros(len(rects) * 16 * 5, dtype=torch.int64, device="cuda")
ms = gpu_ms(lambda: sb.check(lib.svt_b200_lr_sgr_filter_sums(C.byref(a), C.byref(b), 0, C.c_void_p(r.data_ptr()), len(rects), 64, 64, eps, 16,
                                                             C.c_void_p(flt.data_ptr()), C.c_void_p(sums.data_ptr()), None), lib), n=5)
xq = torch.zeros(len(rects) * 16 * 2, dtype=torch.int32, device="cuda")
err = torch.zeros(len(rects) * 16, dtype=torch.int64, device="cuda")
ms2 = gpu_ms(lambda: sb.check(lib.svt_b200_lr_sgr_proj_error(C.byref(a), C.byref(b), 0, C.c_void_p(r.data_ptr()), len(rects), 64, 64, eps, 16,
                                                             C.c_void_p(flt.data_ptr()), C.c_void_p(xq.data_ptr()), C.c_void_p(err.data_ptr()), None), lib), n=5)
f0, f1 = np.zeros((64, 64), np.int32), np.zeros((64, 64), np.int32)
t0 = time.perf_counter()
for (hs, he, vs, ve) in rects[::32]:
    for ep in range(16):
        ref.svt_av1_selfguided_restoration_c(C.c_void_p(ext.ctypes.data + (vs + 3) * ext.shape[1] + hs + 3), he - hs, ve - vs, ext.shape[1],
                                             cm.ptr(f0), cm.ptr(f1), 64, ep, 8, 0)
ref_ms = (time.perf_counter() - t0) * 1e3 * 32
out["rows"].append({"entry": "svt_b200_lr_sgr_filter_sums (luma, 16 parameter sets, 510 units)", "gpu_ms": ms,
                    "hbm_bytes_written": 16 * 2 * 4 * W * H, "achieved_gbs": 16 * 2 * 4 * W * H / ms / 1e6,
                    "reference_c_ms_1thread (filter only)": ref_ms})
out["rows"].append({"entry": "svt_b200_lr_sgr_proj_error (one hill-climb step, 16 sets x 510 units)", "gpu_ms": ms2,
                    "algorithmic_bytes": 16 * (2 * 4 + 2) * W * H, "achieved_gbs": 16 * (2 * 4 + 2) * W * H / ms2 / 1e6})

# ---- svt_b200_inter_predict: every block of a 1080p picture (random AV1 partition, 35 % compound, 4 references) ------
import interp_cases as ic  # noqa: E402
irefs = [ic.ref_picture(W, H, BD, 900 + i) for i in range(4)]
d_irefs = [gr.DevYuv(x) for x in irefs]
ipred = cm.Yuv(W, H, BD, pad=ic.REF_PAD)
d_ipred = gr.DevYuv(ipred)
iarr = (sb.Frame * len(irefs))(*[x.struct() for x in d_irefs])
ips = d_ipred.struct()
lib.svt_b200_inter_predict_scratch_bytes.restype = C.c_size_t
for what, ijobs in (("every AV1 block shape down to 4x4 / 2x2 chroma", ic.make_jobs(W, H, len(irefs), 39)),
                    ("square blocks 8x8..64x64 (preset-8 like)", ic.make_jobs(W, H, len(irefs), 40, min_n=8, square_only=True))):
    isb = lib.svt_b200_inter_predict_scratch_bytes(len(ijobs), W, H)
    iscr = torch.zeros(isb, dtype=torch.uint8, device="cuda")
    d_ijobs = torch.from_numpy(np.ascontiguousarray(ijobs).view(np.uint8)).cuda()
    ms = gpu_ms(lambda: sb.check(lib.svt_b200_inter_predict(iarr, len(irefs), C.byref(ips), C.c_void_p(d_ijobs.data_ptr()), len(ijobs),
                                                            C.c_void_p(iscr.data_ptr()), C.c_size_t(isb), None), lib))
    t0 = time.perf_counter()
    ic.run_cpu(cm.refh().refh_inter_predict, irefs, ipred, ijobs)
    ref_ms = (time.perf_counter() - t0) * 1e3
    area = int((ijobs["bw"].astype(np.int64) * ijobs["bh"] * ijobs["n_refs"]).sum())
    alg = area + samples  # one read per predicted sample and reference + one write per sample (window overlap not counted)
    out["rows"].append({"entry": "svt_b200_inter_predict, %s: %d jobs, %d compound" % (what, len(ijobs), int((ijobs["n_refs"] == 2).sum())),
                        "gpu_ms": ms, "algorithmic_bytes": alg, "achieved_gbs": alg / ms / 1e6, "hbm_frac": alg / ms / 1e6 / peak,
                        "reference_c_ms_1thread": ref_ms})

# ---- svt_b200_subpel_search: one search per 16x16 block of the picture and 2 references (md_subpel_search's set-up) ----
import subpel_cases as sc  # noqa: E402
from test_subpel_gpu import run_gpu as _subpel_run  # noqa: E402,F401
ssrc, srefs = sc.pictures(W, H, 77, n_refs=2)
sp, stabs = sc.params(seed=5, search_type=3, iters=2, allow_hp=1)
mi_cols, mi_rows = 2 * ((W + 7) >> 3), 2 * ((H + 7) >> 3)
sj = []
srng = np.random.default_rng(9)
for by in range(0, H - 15, 16):
    for bx in range(0, W - 15, 16):
        for r in range(2):
            j = np.zeros((), sb.SUBPEL_JOB_DTYPE)
            fr, fc = int(srng.integers(-8, 9)), int(srng.integers(-8, 9))
            j["blk_x"], j["blk_y"], j["bw"], j["bh"], j["ref"] = bx, by, 16, 16, r
            j["start_mv_row"], j["start_mv_col"], j["ref_mv_row"], j["ref_mv_col"] = fr * 8, fc * 8, fr * 8 + 3, fc * 8 - 5
            j["col_min"], j["col_max"], j["row_min"], j["row_max"] = sc.limits(mi_rows, mi_cols, bx, by, 16, 16, fr * 8 + 3, fc * 8 - 5)
            sj.append(j)
sj = np.array(sj, dtype=sb.SUBPEL_JOB_DTYPE)
d_ssrc, d_srefs = gr.DevYuv(ssrc), [gr.DevYuv(x) for x in srefs]
d_stabs = [torch.from_numpy(t).cuda() for t in stabs]
sq = sb.SubpelParams.from_buffer_copy(sp)
for i in range(2):
    sq.mvcost[i] = d_stabs[i].data_ptr() + 4 * sc.MV_MAX
sq.max_block_w = sq.max_block_h = 16
d_sj = torch.from_numpy(sj.view(np.uint8)).cuda()
d_sr = torch.zeros(len(sj) * 16, dtype=torch.uint8, device="cuda")
sss = d_ssrc.struct()
sarr = (sb.Frame * 2)(*[x.struct() for x in d_srefs])
ms = gpu_ms(lambda: sb.check(lib.svt_b200_subpel_search(C.byref(sq), C.byref(sss), sarr, 2, C.c_void_p(d_sj.data_ptr()), len(sj),
                                                       C.c_void_p(d_sr.data_ptr()), None), lib), n=5)
cm.refh().refh_subpel_search.restype = C.c_int
t0 = time.perf_counter()
sc.run_cpu(cm.refh().refh_subpel_search, sp, stabs, ssrc, srefs, sj[::16])
ref_ms = (time.perf_counter() - t0) * 1e3 * 16
out["rows"].append({"entry": "svt_b200_subpel_search (%d searches: every 16x16 block x 2 references, 8-tap, 1/8 sample, 2 iterations per step)" % len(sj),
                    "gpu_ms": ms, "searches_per_s": len(sj) / ms * 1e3, "reference_c_ms_1thread": ref_ms})

# ---- svt_av1_pick_filter_level (full-image search) -------------------------------------------------------------------
mi_rows, mi_cols, part, psrc, prec = pick_case(W, H, BD, 6)
last = (24, 20, 14, 10)
flat = flat_mi(mi_rows, mi_cols, part, last)
pp = pick_params(mi_rows, mi_cols, 0, 3, last)
t0 = time.perf_counter()
got, _ = gr.run_gpu_pick(pp, prec, psrc, flat)  # includes the uploads of this helper
torch.cuda.synchronize()
dr, dsrc2, dt = gr.DevYuv(prec.copy()), gr.DevYuv(psrc), gr.DevYuv(prec.copy())
dmi = torch.from_numpy(np.frombuffer(flat, dtype=np.uint8).copy()).cuda()
scratch = torch.zeros(1024, dtype=torch.uint8, device="cuda")
rs, ss2, ts = dr.struct(), dsrc2.struct(), dt.struct()
lv = (C.c_int32 * 4)()
t0 = time.perf_counter()
for _ in range(3):
    sb.check(lib.svt_b200_pick_filter_level(C.byref(pp), C.byref(rs), C.byref(ss2), C.byref(ts), C.c_void_p(dmi.data_ptr()),
                                            C.c_void_p(scratch.data_ptr()), lv, None), lib)
wall = (time.perf_counter() - t0) / 3 * 1e3
t0 = time.perf_counter()
want, _ = run_ref_pick(mi_rows, mi_cols, part, psrc, prec, 0, 3, last)
ref_ms = (time.perf_counter() - t0) * 1e3
out["rows"].append({"entry": "svt_b200_pick_filter_level (LPF_PICK_FROM_FULL_IMAGE, loop_filter_mode 3)", "wall_ms_incl_host_bisection": wall,
                    "levels": list(lv), "reference_levels": want, "reference_c_ms_1thread": ref_ms})
print(json.dumps(out, indent=1))
