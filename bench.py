#!/usr/bin/env python
"""bench.py — hot-path throughput of the B200 build vs the reference C path.

A "step" is one pass of the hot path over one mini-GOP of FRAMES_PER_STEP synthetic 1920x1080 8-bit frames
(BASELINE.json configs[1] geometry, preset-8 parameters).  Per frame, in the order the reference's pipeline runs:
  1. open-loop ME        motion_estimation_kernel: HME + full-pel search (+ SB epilogue), 2+2 refs (2 launches)
  2. EncDec final pass   residual -> fwd txfm -> quant/dequant -> inverse txfm -> recon, every TU (3 launches)
  3. deblocking          svt_av1_loop_filter_frame, all planes                                  (2 launches)
  4. CDEF                cdef_seg_search (10 strengths, preset 8) + svt_av1_cdef_frame          (2 launches)
  value : frames/s with all inputs resident in HBM (CUDA events on the launch stream, max over ranks)
  e2e   : same work through the C ABI with HOST (pinned) buffers: H2D of the source/prediction planes, ME planes
          and mode-info summary; D2H of MeSbResults, quantised coefficients + eobs, CDEF mse and the final recon.
  --impl reference : the reference's own C implementation (oracle/_ref, unmodified sources, its RTCD C paths)
          of the same four stages on the host cores, one frame per thread, on a bounded sample.
Multi-GPU (torchrun): independent streams sharded one per rank, no data-path collective ("weak" scaling).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.join(ROOT, "svt-av1_b200"), os.path.join(ROOT, "tests"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

W, H = 1920, 1080
MI_ROWS, MI_COLS = H // 4, W // 4
FRAMES_PER_STEP = 8
N_L0, N_L1 = 2, 2
DIST = ((1, 2, 3, 4), (1, 2, 3, 4))
RING = 4  # distinct mini-GOP input sets cycled between steps so that the working set exceeds the 126 MB L2
QINDEX_LEVELS = (24, 20, 14, 10)  # deblocking levels (Y vert, Y horz, U, V)
BASE_Q_IDX = 172  # qp 43
METRIC = "1080p30 8-bit preset-8 hot-path fps"


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler(threading.Thread):
    def __init__(self, idx):
        super().__init__(daemon=True)
        self.idx, self.rows, self.stop_flag = idx, [], False

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": int(self.rows[0][1]), "reasons": reasons}


# ------------------------------------------------------------------------------------------------------
# workload description shared by both arms
# ------------------------------------------------------------------------------------------------------
def workload_config(frames):
    return {"workload": f"{W}x{H} 8-bit yuv420p preset 8 hot path, {frames} frames/step (BASELINE configs[1] geometry)",
            "stages": ["me(hme+fullpel, %d+%d refs)" % (N_L0, N_L1), "encdec(residual+fwd txfm+quant+inv txfm+recon, all TUs)",
                       "dlf(frame, levels %s)" % (QINDEX_LEVELS,), "cdef(search 10 strengths + apply)"],
            "l2_policy": f"ring of {RING} distinct input sets (> L2) cycled between steps",
            "issue": "4 pictures in flight, one CUDA stream each"}


def make_frames(seed, n):
    """n synthetic 4:2:0 source pictures + a 'prediction' for each (the source with coding-like error)."""
    import common as cm
    src = [cm.synth_yuv(W, H, i, seed, 8) for i in range(n)]
    pred = [cm.degrade(s, seed + i, amp=10) for i, s in enumerate(src)]
    return src, pred


def tu_lists():
    """Transform units of one 1080p frame: 16x16 luma, 8x8 chroma (+ the 8- and 4-row remainders)."""
    import svtb200 as sb
    lists = {2: [], 1: [], 0: []}
    for y in range(0, 1072, 16):
        for x in range(0, W, 16):
            lists[2].append(sb.Tu(x, y, 0, 0))
    for x in range(0, W, 8):
        lists[1].append(sb.Tu(x, 1072, 0, 0))
    for pl in (1, 2):
        for y in range(0, 536, 8):
            for x in range(0, W // 2, 8):
                lists[1].append(sb.Tu(x, y, pl, 0))
        for x in range(0, W // 2, 4):
            lists[0].append(sb.Tu(x, 536, pl, 0))
    return lists


def quant_params(tx_size):
    """Plausible qindex-172 tables (dc/ac) — identical for both arms."""
    import svtb200 as sb
    p = sb.EncodeParams()
    p.tx_size, p.use_fp = tx_size, 0
    for i in range(3):
        q = p.q[i]
        for k, dq in enumerate((88, 104)):
            q.dequant[k] = dq
            q.zbin[k] = (dq * 84 + 64) >> 7
            q.round[k] = (dq * 48) >> 7
            q.quant[k] = ((1 << 16) // dq) - 1 if dq > 2 else 32767
            q.quant_shift[k] = 1 << 14
            q.round_fp[k] = (dq * 48) >> 7
            q.quant_fp[k] = (1 << 16) // dq
    return p


def partition_and_mi(seed):
    import common as cm
    import svtb200 as sb
    from test_dlf_gpu import flat_mi
    part = cm.random_partition(MI_ROWS, MI_COLS, seed, p_split=0.5)
    flat = flat_mi(MI_ROWS, MI_COLS, part, QINDEX_LEVELS)
    skip8 = np.ascontiguousarray((part[3][0::2, 0::2] & part[3][1::2, 0::2] & part[3][0::2, 1::2] & part[3][1::2, 1::2]).astype(np.uint8))
    return part, flat, skip8


def cdef_search_params():
    import svtb200 as sb
    p = sb.CdefSearchParams()
    p.mi_rows, p.mi_cols, p.pri_damping = MI_ROWS, MI_COLS, 3 + (BASE_Q_IDX >> 6)
    sb.load().svt_b200_cdef_strength_table(3, C.byref(p))
    return p


def cdef_apply_params():
    import svtb200 as sb
    p = sb.CdefApplyParams()
    p.mi_rows, p.mi_cols, p.damping = MI_ROWS, MI_COLS, 3 + (BASE_Q_IDX >> 6)
    for i, (a, b) in enumerate(zip((0, 5, 17, 63, 40, 2, 12, 33), (0, 0, 9, 62, 4, 1, 60, 3))):
        p.y_strength[i], p.uv_strength[i] = a, b
    return p


def dlf_params():
    import svtb200 as sb
    p = sb.DlfParams()
    p.mi_rows, p.mi_cols, p.mi_stride, p.sharpness = MI_ROWS, MI_COLS, MI_COLS, 0
    p.filter_level[0], p.filter_level[1], p.filter_level_u, p.filter_level_v = QINDEX_LEVELS
    p.plane_start, p.plane_end = 0, 3
    return p


# ------------------------------------------------------------------------------------------------------
# reference arm (also the cpu_baseline leg)
# ------------------------------------------------------------------------------------------------------
def reference_frames_per_second(n_frames, repeats=1):
    """Runs the four stages with the reference's own C code (oracle/_ref) on n_frames frames, one per thread."""
    import common as cm
    import svtb200 as sb
    from concurrent.futures import ThreadPoolExecutor
    cores = os.cpu_count() or 1
    rh = cm.refh()
    rh.refh_init()
    geos = sb.me_geometry(W, H)
    src, pred = make_frames(1234, n_frames + 4)
    me_pics = [cm.me_planes(np.ascontiguousarray(s.plane(0)), geos) for s in src]
    part, flat, skip8 = partition_and_mi(7)
    sbt, dep, inter, skip = (np.ascontiguousarray(x) for x in part)
    tus = tu_lists()
    nfb = ((MI_ROWS + 15) // 16) * ((MI_COLS + 15) // 16)
    capp = cdef_apply_params()

    def one(i):
        # 1. ME
        f = i + 2
        refs = [me_pics[f - 1], me_pics[f - 2], me_pics[f - 2], me_pics[f - 2], me_pics[f + 1], me_pics[f + 2], me_pics[f + 2], me_pics[f + 2]]
        cm.run_ref_me(W, H, 8, N_L0, N_L1, DIST, 2, 1, geos, me_pics[f], refs)
        # 2. EncDec
        rec = pred[f].copy()
        ss, ps, rs = src[f].struct(), pred[f].struct(), rec.struct()
        for ts, lst in tus.items():
            p = quant_params(ts)
            arr = (sb.Tu * len(lst))(*lst)
            n = min(sb.TX_W[ts], 32) * min(sb.TX_H[ts], 32)
            q = np.zeros(len(lst) * n, np.int32)
            eob = np.zeros(len(lst), np.uint16)
            rh.refh_encode_tus(C.byref(p), C.byref(ss), C.byref(ps), C.byref(rs), arr, len(lst), cm.ptr(q), cm.ptr(eob))
        # 3. deblocking
        lv = (C.c_int32 * 4)(*QINDEX_LEVELS)
        rh.refh_dlf_frame(MI_ROWS, MI_COLS, cm.ptr(sbt), cm.ptr(dep), cm.ptr(inter), cm.ptr(skip), lv, 0, C.byref(rs), None)
        # 4. CDEF search + apply
        mse = np.zeros((2, nfb, 64), np.uint64)
        rh.refh_cdef_search(MI_ROWS, MI_COLS, BASE_Q_IDX, 4, C.byref(rs), C.byref(ss), cm.ptr(skip8), skip8.shape[1], cm.ptr(mse))
        idx = (np.argmin(mse[0, :, :8], axis=1)).astype(np.int8)
        ys = (C.c_int32 * 8)(*capp.y_strength)
        uvs = (C.c_int32 * 8)(*capp.uv_strength)
        rh.refh_cdef_apply(MI_ROWS, MI_COLS, capp.damping, ys, uvs, C.byref(rs), cm.ptr(skip8), skip8.shape[1], cm.ptr(idx))

    times = []
    with ThreadPoolExecutor(min(cores, n_frames)) as ex:
        for _ in range(repeats):
            t0 = time.perf_counter()
            list(ex.map(one, range(n_frames)))
            times.append(time.perf_counter() - t0)
    return n_frames, times, min(cores, n_frames)


def run_reference(args):
    import common as cm
    if not cm.have_ref():
        emit({"impl": "reference", "unavailable": "oracle/_ref not built on this box"})
        return
    cores = os.cpu_count() or 1
    sample = max(1, min(FRAMES_PER_STEP if cores < 16 else 2 * FRAMES_PER_STEP, cores))
    nfr, times, used = reference_frames_per_second(sample, repeats=args.warmup + args.steps)
    times = times[args.warmup:]
    fps = nfr * len(times) / sum(times)
    line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": workload_config(nfr),
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": used, "kind": "reference",
                             "sample": f"{nfr} frames per step of the same 1080p workload, one frame per thread, unmodified "
                                       "reference C paths (-O2, no SIMD: no nasm in the image)"},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


# ------------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------------
class DeviceSet:
    """One mini-GOP worth of inputs: host (pinned) and device copies."""

    def __init__(self, torch, seed, shift):
        import common as cm
        import svtb200 as sb
        self.geos = sb.me_geometry(W, H)
        src, pred = make_frames(seed, FRAMES_PER_STEP + 4)
        if shift:
            for f in src + pred:
                for b in f.bufs:
                    b[...] = np.roll(b, shift, axis=1)
        self.src, self.pred = src, pred
        self.me_host = [[torch.from_numpy(x).pin_memory() for x in cm.me_planes(np.ascontiguousarray(s.plane(0)), self.geos)] for s in src]
        self.me_dev = [[t.cuda() for t in pics] for pics in self.me_host]
        self.src_host = [[torch.from_numpy(b).pin_memory() for b in f.bufs] for f in src]
        self.pred_host = [[torch.from_numpy(b).pin_memory() for b in f.bufs] for f in pred]
        self.src_dev = [[t.cuda() for t in f] for f in self.src_host]
        self.pred_dev = [[t.cuda() for t in f] for f in self.pred_host]


def frame_struct(sb, yuv, tensors):
    p = yuv.pad
    ptrs = [t.data_ptr() + (p * b.shape[1] + p) * b.itemsize for t, b in zip(tensors, yuv.bufs)]
    return sb.Frame(ptrs[0], ptrs[1], ptrs[2], yuv.bufs[0].shape[1], yuv.bufs[1].shape[1], yuv.w, yuv.h, yuv.bd)


def run_b200(args):
    import torch
    import torch.distributed as dist
    import common as cm
    import svtb200 as sb

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = sb.load()
    sb.check(lib.svt_b200_set_device(local), lib)
    stream = torch.cuda.Stream()
    sp = C.c_void_p(stream.cuda_stream)

    me_params = sb.preset8_me_params(W, H, N_L0, N_L1, DIST, 2, 1)
    n_sb = ((W + 63) // 64) * ((H + 63) // 64)
    nfb = n_sb
    import sharding
    sets = [DeviceSet(torch, sharding.stream_seed(1234, rank), 5 * k) for k in range(RING)]
    part, flat, skip8 = partition_and_mi(7)
    h_mi = torch.from_numpy(np.frombuffer(flat, dtype=np.uint8).copy()).pin_memory()
    d_mi = h_mi.cuda()
    h_skip = torch.from_numpy(skip8).pin_memory()
    d_skip = h_skip.cuda()
    tus = tu_lists()
    tu_dev, enc_params, n_coef = {}, {}, {}
    for ts, lst in tus.items():
        arr = (sb.Tu * len(lst))(*lst)
        tu_dev[ts] = torch.from_numpy(np.frombuffer(arr, dtype=np.int32).copy()).cuda()
        enc_params[ts] = quant_params(ts)
        n_coef[ts] = min(sb.TX_W[ts], 32) * min(sb.TX_H[ts], 32)
    csp, cap, dlp = cdef_search_params(), cdef_apply_params(), dlf_params()
    d_idx = torch.from_numpy(np.random.default_rng(3).integers(0, 8, nfb).astype(np.int8)).cuda()
    h_idx = torch.zeros(nfb, dtype=torch.int8).pin_memory()

    F = FRAMES_PER_STEP

    def me_out(pinned=False):
        mk = (lambda n, dt: torch.empty(n, dtype=dt).pin_memory()) if pinned else (lambda n, dt: torch.empty(n, dtype=dt, device="cuda"))
        return {"best_sad": mk(n_sb * 8 * 85, torch.int32), "best_mv": mk(n_sb * 8 * 85, torch.int32),
                "hme": mk(n_sb * 8 * 16, torch.uint8), "me_mv": mk(n_sb * 85 * 7 * 2, torch.int16),
                "me_cand": mk(n_sb * 85 * 23, torch.uint8), "total_cand": mk(n_sb * 85, torch.uint8), "rc": mk(n_sb, torch.int32)}

    d_me = [me_out() for _ in range(F)]
    h_me = [{k: me_out(True)[k] for k in ("me_mv", "me_cand", "total_cand", "rc")} for _ in range(F)]
    me_scratch = [torch.empty(lib.svt_b200_me_scratch_bytes(C.byref(me_params)), dtype=torch.uint8, device="cuda") for _ in range(F)]
    d_q = [{ts: torch.empty(len(tus[ts]) * n_coef[ts], dtype=torch.int32, device="cuda") for ts in tus} for _ in range(F)]
    d_eob = [{ts: torch.empty(len(tus[ts]), dtype=torch.int16, device="cuda") for ts in tus} for _ in range(F)]
    h_q = [{ts: torch.empty(len(tus[ts]) * n_coef[ts], dtype=torch.int32).pin_memory() for ts in tus} for _ in range(F)]
    h_eob = [{ts: torch.empty(len(tus[ts]), dtype=torch.int16).pin_memory() for ts in tus} for _ in range(F)]
    enc_scratch = torch.empty(16384, dtype=torch.uint8, device="cuda")
    d_mse = [torch.empty(2 * nfb * 64, dtype=torch.int64, device="cuda") for _ in range(F)]
    h_mse = [torch.empty(2 * nfb * 64, dtype=torch.int64).pin_memory() for _ in range(F)]
    proto = sets[0].src[0]
    d_rec = [[torch.empty_like(t) for t in sets[0].src_dev[0]] for _ in range(F)]
    d_out = [[torch.empty_like(t) for t in sets[0].src_dev[0]] for _ in range(F)]
    h_out = [[torch.empty(t.shape, dtype=t.dtype).pin_memory() for t in sets[0].src_dev[0]] for _ in range(F)]
    # device staging of the e2e arm
    e_me = [[torch.empty_like(t) for t in pics] for pics in sets[0].me_dev]
    e_src = [[torch.empty_like(t) for t in f] for f in sets[0].src_dev]
    e_pred = [[torch.empty_like(t) for t in f] for f in sets[0].pred_dev]
    e_mi, e_skip = torch.empty_like(d_mi), torch.empty_like(d_skip)

    def planes(tl):
        return sb.MePlanes(tl[0].data_ptr(), tl[1].data_ptr(), tl[2].data_ptr())

    # Pictures are independent once their references are resident, so the mini-GOP is issued the way the reference's
    # picture-level pipeline would: NS pictures in flight, one CUDA stream each (kernel tails, copies and the host-side
    # CDEF strength decision of one picture overlap the kernels of the others).
    NS = 4
    streams = [torch.cuda.Stream() for _ in range(NS)]
    sps = [C.c_void_p(st.cuda_stream) for st in streams]
    copy_stream = torch.cuda.Stream()
    enc_scratch_s = [torch.empty(16384, dtype=torch.uint8, device="cuda") for _ in range(NS)]
    e_idx = [torch.empty(nfb, dtype=torch.int8, device="cuda") for _ in range(F)]
    h_idx_f = [torch.zeros(nfb, dtype=torch.int8).pin_memory() for _ in range(F)]

    def frame_front(i, me_dev, src_dev, pred_dev, mi_dev, skip_dev, q, scratch):
        """ME -> EncDec -> deblocking -> CDEF strength search of picture i on stream q."""
        f = i + 2
        fs, fp, fr = frame_struct(sb, proto, src_dev[f]), frame_struct(sb, proto, pred_dev[f]), frame_struct(sb, proto, d_rec[i])
        r = [me_dev[f - 1], me_dev[f - 2], me_dev[f - 2], me_dev[f - 2], me_dev[f + 1], me_dev[f + 2], me_dev[f + 2], me_dev[f + 2]]
        refs = (sb.MePlanes * 8)(*[planes(x) for x in r])
        s = planes(me_dev[f])
        o = d_me[i]
        outs = sb.MeOutputs(o["best_sad"].data_ptr(), o["best_mv"].data_ptr(), o["hme"].data_ptr(), o["me_mv"].data_ptr(),
                            o["me_cand"].data_ptr(), o["total_cand"].data_ptr(), o["rc"].data_ptr())
        sb.check(lib.svt_b200_me_picture(C.byref(me_params), C.byref(s), refs, C.byref(outs), me_scratch[i].data_ptr(), q), lib)
        for ts in tus:
            sb.check(lib.svt_b200_encode_tus(C.byref(enc_params[ts]), C.byref(fs), C.byref(fp), C.byref(fr),
                                             C.c_void_p(tu_dev[ts].data_ptr()), len(tus[ts]), C.c_void_p(d_q[i][ts].data_ptr()),
                                             C.c_void_p(d_eob[i][ts].data_ptr()), C.c_void_p(scratch.data_ptr()), q), lib)
        sb.check(lib.svt_b200_dlf_frame(C.byref(dlp), C.byref(fr), C.c_void_p(mi_dev.data_ptr()), q), lib)
        sb.check(lib.svt_b200_cdef_search(C.byref(csp), C.byref(fr), C.byref(fs), C.c_void_p(skip_dev.data_ptr()), skip8.shape[1],
                                          C.c_void_p(d_mse[i].data_ptr()), q), lib)

    def frame_back(i, skip_dev, idx_dev, q):
        """CDEF apply of picture i with the per-filter-block strength indices in idx_dev."""
        fr, fo = frame_struct(sb, proto, d_rec[i]), frame_struct(sb, proto, d_out[i])
        sb.check(lib.svt_b200_cdef_apply(C.byref(cap), C.byref(fr), C.byref(fo), C.c_void_p(skip_dev.data_ptr()), skip8.shape[1],
                                         C.c_void_p(idx_dev.data_ptr()), q), lib)

    def fork():
        ev = torch.cuda.Event()
        ev.record(stream)
        for st in streams + [copy_stream]:
            st.wait_event(ev)

    def join():
        for st in streams + [copy_stream]:
            ev = torch.cuda.Event()
            ev.record(st)
            stream.wait_event(ev)

    def hot_path(me_dev, src_dev, pred_dev, mi_dev, skip_dev, idx_dev, only=None):
        """Single-stream issue (per-stage timing)."""
        for i in range(F):
            f = i + 2
            fs, fp, fr = frame_struct(sb, proto, src_dev[f]), frame_struct(sb, proto, pred_dev[f]), frame_struct(sb, proto, d_rec[i])
            fo = frame_struct(sb, proto, d_out[i])
            if only is not None:
                stage_call(only, i, f, me_dev, fs, fp, fr, fo, mi_dev, skip_dev, idx_dev)
            else:
                frame_front(i, me_dev, src_dev, pred_dev, mi_dev, skip_dev, sp, enc_scratch)
                frame_back(i, skip_dev, idx_dev, sp)

    def stage_call(name, i, f, me_dev, fs, fp, fr, fo, mi_dev, skip_dev, idx_dev):
        """One stage of one frame (used by the per-stage roofline timing)."""
        if name == "me":
            r = [me_dev[f - 1], me_dev[f - 2], me_dev[f - 2], me_dev[f - 2], me_dev[f + 1], me_dev[f + 2], me_dev[f + 2], me_dev[f + 2]]
            refs = (sb.MePlanes * 8)(*[planes(x) for x in r])
            s = planes(me_dev[f])
            o = d_me[i]
            outs = sb.MeOutputs(o["best_sad"].data_ptr(), o["best_mv"].data_ptr(), o["hme"].data_ptr(), o["me_mv"].data_ptr(),
                                o["me_cand"].data_ptr(), o["total_cand"].data_ptr(), o["rc"].data_ptr())
            sb.check(lib.svt_b200_me_picture(C.byref(me_params), C.byref(s), refs, C.byref(outs), me_scratch[i].data_ptr(), sp), lib)
        elif name == "encdec":
            for ts in tus:
                sb.check(lib.svt_b200_encode_tus(C.byref(enc_params[ts]), C.byref(fs), C.byref(fp), C.byref(fr),
                                                 C.c_void_p(tu_dev[ts].data_ptr()), len(tus[ts]), C.c_void_p(d_q[i][ts].data_ptr()),
                                                 C.c_void_p(d_eob[i][ts].data_ptr()), C.c_void_p(enc_scratch.data_ptr()), sp), lib)
        elif name == "dlf":
            sb.check(lib.svt_b200_dlf_frame(C.byref(dlp), C.byref(fr), C.c_void_p(mi_dev.data_ptr()), sp), lib)
        elif name == "cdef_search":
            sb.check(lib.svt_b200_cdef_search(C.byref(csp), C.byref(fr), C.byref(fs), C.c_void_p(skip_dev.data_ptr()), skip8.shape[1],
                                              C.c_void_p(d_mse[i].data_ptr()), sp), lib)
        elif name == "cdef_apply":
            sb.check(lib.svt_b200_cdef_apply(C.byref(cap), C.byref(fr), C.byref(fo), C.c_void_p(skip_dev.data_ptr()), skip8.shape[1],
                                             C.c_void_p(idx_dev.data_ptr()), sp), lib)

    def step_resident(k):
        s = sets[k % RING]
        fork()
        for i in range(F):
            q = i % NS
            frame_front(i, s.me_dev, s.src_dev, s.pred_dev, d_mi, d_skip, sps[q], enc_scratch_s[q])
            frame_back(i, d_skip, d_idx, sps[q])
        join()

    def step_e2e(k):
        s = sets[k % RING]
        fork()
        # H2D: the ME planes of the F+4 pictures (shared by neighbouring pictures) on the copy stream, in display order
        me_ready = []
        with torch.cuda.stream(copy_stream):
            e_mi.copy_(h_mi, non_blocking=True)
            e_skip.copy_(h_skip, non_blocking=True)
            for j in range(F + 4):
                for a, b in zip(e_me[j], s.me_host[j]):
                    a.copy_(b, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy_stream)
                me_ready.append(ev)
        mse_ready = []
        for i in range(F):
            q, f = i % NS, i + 2
            with torch.cuda.stream(streams[q]):
                for a, b in zip(e_src[f], s.src_host[f]):
                    a.copy_(b, non_blocking=True)
                for a, b in zip(e_pred[f], s.pred_host[f]):
                    a.copy_(b, non_blocking=True)
            streams[q].wait_event(me_ready[f + 2])
            frame_front(i, e_me, e_src, e_pred, e_mi, e_skip, sps[q], enc_scratch_s[q])
            with torch.cuda.stream(streams[q]):
                h_mse[i].copy_(d_mse[i], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(streams[q])
                mse_ready.append(ev)
                # results that do not depend on the CDEF decision go home while the host decides
                for kk, t in h_me[i].items():
                    t.copy_(d_me[i][kk], non_blocking=True)
                for ts in tus:
                    h_q[i][ts].copy_(d_q[i][ts], non_blocking=True)
                    h_eob[i][ts].copy_(d_eob[i][ts], non_blocking=True)
        for i in range(F):
            q = i % NS
            mse_ready[i].synchronize()
            # stand-in for finish_cdef_search (host side, out of scope §8a): best of the first 8 strengths per block
            m = h_mse[i].numpy().view(np.uint64).reshape(2, nfb, 64)
            h_idx_f[i].numpy()[...] = np.argmin(m[0, :, :8], axis=1).astype(np.int8)
            with torch.cuda.stream(streams[q]):
                e_idx[i].copy_(h_idx_f[i], non_blocking=True)
            frame_back(i, e_skip, e_idx[i], sps[q])
            with torch.cuda.stream(streams[q]):
                for a, b in zip(h_out[i], d_out[i]):
                    a.copy_(b, non_blocking=True)
        join()

    h2d = (sum(t.numel() * t.element_size() for t in sets[0].me_host[0]) * (F + 4) +
           2 * F * sum(t.numel() * t.element_size() for t in sets[0].src_host[0]) + h_mi.numel() + h_skip.numel() + F * nfb)
    d2h = F * (sum(t.numel() * t.element_size() for t in h_me[0].values()) + sum(t.numel() * 4 for t in h_q[0].values()) +
               sum(t.numel() * 2 for t in h_eob[0].values()) + h_mse[0].numel() * 8 + sum(t.numel() * t.element_size() for t in h_out[0]))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for k in range(warmup):
            fn(k)
        barrier()
        l0 = lib.svt_b200_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for k in range(steps):
            fn(warmup + k)
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        launches = lib.svt_b200_launch_count() - l0
        ms = sharding.reduce_max_ms(ms, "cuda")
        return ms, launches

    sampler = ClockSampler(local)
    sampler.start()
    ms, launches = timed(step_resident, args.steps, args.warmup)
    ms_e2e, _ = timed(step_e2e, args.steps, args.warmup)
    sampler.stop_flag = True
    sampler.join(timeout=2)

    pk, pk_kind = peaks()
    stage_ms = stage_breakdown(torch, lib, sb, hot_path, sets[0], d_mi, d_skip, d_idx, stream)
    roof = roofline(stage_ms, pk, pk_kind, n_sb)

    frames = F * args.steps * world
    line = {"metric": METRIC, "value": frames / (ms / 1e3), "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic", "config": workload_config(F), "clocks": sampler.summary(),
            "e2e": {"value": frames / (ms_e2e / 1e3), "unit": "frames/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
            "gpu_launches": int(launches), "roofline": roof, "stage_ms_per_frame": stage_ms}
    if rank == 0:
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline()
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def stage_breakdown(torch, lib, sb, hot_path, s, d_mi, d_skip, d_idx, stream):
    """Per-stage device time per frame: each stage's launches alone, back to back over the mini-GOP's frames, CUDA
    events on the launch stream, warm (3 untimed passes), averaged over 5 passes."""
    out = {}
    for name in (None, "me", "encdec", "dlf", "cdef_search", "cdef_apply"):
        for _ in range(3):
            hot_path(s.me_dev, s.src_dev, s.pred_dev, d_mi, d_skip, d_idx, only=name)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        n = 5
        for _ in range(n):
            hot_path(s.me_dev, s.src_dev, s.pred_dev, d_mi, d_skip, d_idx, only=name)
        e1.record(stream)
        torch.cuda.synchronize()
        out[name or "all_stages"] = e0.elapsed_time(e1) / (n * FRAMES_PER_STEP)
    return out


def roofline(stage_ms, pk, pk_kind, n_sb):
    """Algorithmic HBM bytes per frame of every stage over its measured device time; the top-level
    fields describe the dominant stage, `stages` lists all of them."""
    luma, chroma = W * H, 2 * (W // 2) * (H // 2)
    samples = luma + chroma
    me_planes = (W + 136) * (H + 136) + (W // 2 + 64) * (H // 2 + 64) + (W // 4 + 32) * (H // 4 + 32)
    alg = {
        "me": me_planes * (1 + N_L0 + N_L1) + n_sb * (85 * 7 * 4 + 85 * 23 + 85 + 4) + n_sb * 8 * 85 * 8,
        "encdec": samples * (1 + 1 + 4 + 1),  # src + pred in, qcoeff (int32) + recon out: 7 B/sample
        "dlf": samples * 2 * 2 + (H // 4) * (W // 4) * 16,  # two passes, read + write, + the mi summary
        "cdef_search": samples * 2 + n_sb * 2 * 64 * 8,  # recon + source in, mse table out
        "cdef_apply": samples * 2,
    }
    bound = {"me": "integer ALU / shared memory (VABSDIFF4 issue rate)", "encdec": "shared-memory butterflies, then HBM",
             "dlf": "hbm", "cdef_search": "integer ALU (10 filters per sample)", "cdef_apply": "hbm"}
    stages = []
    for k, b in alg.items():
        ach = b / (stage_ms[k] / 1e3) / 1e9
        stages.append({"stage": k, "ms_per_frame": stage_ms[k], "algorithmic_bytes": int(b), "achieved": ach,
                       "frac": ach / pk["hbm_gbs"], "binding": bound[k]})
    dom = max(stages, key=lambda x: x["ms_per_frame"])
    return {"kernel": "stage '%s' (dominant; its kernels are listed in profiles/)" % dom["stage"], "bound": "hbm",
            "achieved": dom["achieved"], "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": dom["frac"], "traffic": None,
            "peak_source": pk_kind, "stages": stages,
            "note": "ME and the CDEF strength search are integer-ALU/shared-memory bound (SURVEY §8d): their HBM fraction is "
                    "legitimately small; the streaming stages (dlf, cdef_apply, encdec) are the HBM-bound ones"}


def cpu_baseline():
    import common as cm
    if not cm.have_ref():
        return {"value": None, "unit": "frames/s", "cores": 0, "kind": "reference", "sample": "oracle/_ref not built"}
    cores = os.cpu_count() or 1
    sample = max(1, min(FRAMES_PER_STEP if cores < 16 else 2 * FRAMES_PER_STEP, cores))
    nfr, times, used = reference_frames_per_second(sample, repeats=1)
    return {"value": nfr / times[0], "unit": "frames/s", "cores": used, "kind": "reference",
            "sample": f"{nfr} frames of the 1080p workload through the same four stages, one frame per thread, unmodified "
                      "reference C (-O2, no SIMD)"}


_REAL_STDOUT = None


def emit(obj):
    data = (json.dumps(obj) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    # stdout carries exactly ONE line (the JSON result): anything libraries print to fd 1 (e.g. NCCL's version banner)
    # is routed to stderr, and emit() writes the result to the real stdout
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        if int(os.environ.get("RANK", 0)) == 0:
            run_reference(args)
        return
    run_b200(args)


if __name__ == "__main__":
    main()
