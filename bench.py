#!/usr/bin/env python
"""bench.py — hot-path throughput of the B200 build vs the reference C path.

A "step" is one pass of the hot path over one batch (mini-GOP) of FRAMES_PER_STEP synthetic 1080p 8-bit
frames: open-loop ME of every frame against its references [+ the other stages as they land: see STAGES].
  value : frames/s with all inputs resident in HBM (CUDA events on the launch stream, max over ranks)
  e2e   : same work through the C ABI with HOST (pinned) buffers: H2D of every new frame's planes and D2H of
          the per-SB results inside the timed region
  --impl reference : the reference's own C implementation (oracle/_ref, unmodified sources) of the same
          stages on the host cores (all threads), on a bounded sample of the same workload.
Multi-GPU (torchrun): frames/mini-GOPs are sharded one stream per rank, no data-path collective ("weak").
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
FRAMES_PER_STEP = 8
N_L0, N_L1 = 2, 2
DIST = ((1, 2, 3, 4), (1, 2, 3, 4))
RING = 6  # distinct mini-GOP input sets cycled through so that the working set exceeds the 126 MB L2


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


def make_inputs(seed):
    """One mini-GOP worth of host planes: FRAMES_PER_STEP pictures + the 4 reference pictures around them."""
    import common as cm
    import svtb200 as sb
    geos = sb.me_geometry(W, H)
    pics = [cm.me_planes(cm.synth_luma(W, H, n, seed), geos) for n in range(FRAMES_PER_STEP + 4)]
    return geos, pics


def ref_indices(i):
    """References of picture i inside its set: two past, two future (hierarchical-B like)."""
    past = [max(i + 2 - d, 0) for d in (1, 2)]
    fut = [min(i + 2 + d, FRAMES_PER_STEP + 3) for d in (1, 2)]
    return past, fut


# ------------------------------------------------------------------------------------------------------
def run_reference(args):
    """--impl reference: the unmodified reference C path (oracle/_ref) on the host cores."""
    import common as cm
    import svtb200 as sb
    from concurrent.futures import ThreadPoolExecutor
    if not cm.have_ref():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref not built on this box"}))
        return
    cores = os.cpu_count() or 1
    geos, pics = make_inputs(1234)
    sample_frames = max(1, min(FRAMES_PER_STEP, cores))  # bounded sample: one frame per thread

    def one(i):
        past, fut = ref_indices(i)
        refs = [pics[j] for j in past] + [pics[past[-1]]] * 2 + [pics[j] for j in fut] + [pics[fut[-1]]] * 2
        cm.run_ref_me(W, H, 8, N_L0, N_L1, DIST, 2, 1, geos, pics[i + 2], refs)

    cm.refh().refh_init()
    times = []
    with ThreadPoolExecutor(cores) as ex:
        for it in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            list(ex.map(one, range(sample_frames)))
            dt = time.perf_counter() - t0
            if it >= args.warmup:
                times.append(dt)
    total = sum(times)
    fps = sample_frames * len(times) / total
    line = {"impl": "reference", "metric": "1080p30 8-bit preset-8 hot-path fps", "value": fps, "unit": "frames/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": workload_config(sample_frames),
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "reference",
                             "sample": f"{sample_frames} frames per step of the same 1080p workload, one per thread"},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def workload_config(frames):
    return {"workload": f"{W}x{H} 8-bit yuv420p preset 8 hot path, {frames} frames/step (BASELINE configs[1] geometry)",
            "stages": ["me(hme+fullpel, %d+%d refs)" % (N_L0, N_L1)],
            "l2_policy": f"ring of {RING} distinct input sets (> L2) cycled between steps"}


# ------------------------------------------------------------------------------------------------------
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
    params = sb.preset8_me_params(W, H, N_L0, N_L1, DIST, 2, 1)
    n_sb = ((W + 63) // 64) * ((H + 63) // 64)
    stream = torch.cuda.Stream()
    sp = C.c_void_p(stream.cuda_stream)

    # ---- inputs: RING sets, host pinned + device resident ----
    geos, base = make_inputs(1234 + rank)
    host_sets, dev_sets = [], []
    for k in range(RING):
        hs = []
        for pic in base:
            hp = []
            for pl in pic:
                t = torch.from_numpy(np.roll(pl, k * 3, axis=1).copy()).pin_memory()
                hp.append(t)
            hs.append(hp)
        host_sets.append(hs)
        dev_sets.append([[t.cuda() for t in hp] for hp in hs])
    bytes_in_frame = sum(t.numel() for t in host_sets[0][0])

    def out_tensors(pinned=False):
        mk = (lambda n, dt: torch.empty(n, dtype=dt).pin_memory()) if pinned else (lambda n, dt: torch.empty(n, dtype=dt, device="cuda"))
        return {"best_sad": mk(n_sb * 8 * 85, torch.int32), "best_mv": mk(n_sb * 8 * 85, torch.int32),
                "hme": mk(n_sb * 8 * 16, torch.uint8), "me_mv": mk(n_sb * 85 * 7 * 2, torch.int16),
                "me_cand": mk(n_sb * 85 * 23, torch.uint8), "total_cand": mk(n_sb * 85, torch.uint8),
                "rc": mk(n_sb, torch.int32)}

    d_out = [out_tensors() for _ in range(FRAMES_PER_STEP)]
    h_out = [{k: out_tensors(True)[k] for k in ("me_mv", "me_cand", "total_cand", "rc")} for _ in range(FRAMES_PER_STEP)]
    bytes_out_frame = sum(t.numel() * t.element_size() for t in h_out[0].values())
    scratch = [torch.empty(lib.svt_b200_me_scratch_bytes(C.byref(params)), dtype=torch.uint8, device="cuda")
               for _ in range(FRAMES_PER_STEP)]
    # staging area for the e2e arm (device copies of the uploaded frames)
    e2e_dev = [[torch.empty_like(t, device="cuda") for t in hp] for hp in host_sets[0]]

    def planes(tl):
        return sb.MePlanes(tl[0].data_ptr(), tl[1].data_ptr(), tl[2].data_ptr())

    def me_calls(dset):
        for i in range(FRAMES_PER_STEP):
            past, fut = ref_indices(i)
            r = [dset[j] for j in past] + [dset[past[-1]]] * 2 + [dset[j] for j in fut] + [dset[fut[-1]]] * 2
            refs = (sb.MePlanes * 8)(*[planes(x) for x in r])
            s = planes(dset[i + 2])
            o = d_out[i]
            outs = sb.MeOutputs(o["best_sad"].data_ptr(), o["best_mv"].data_ptr(), o["hme"].data_ptr(), o["me_mv"].data_ptr(),
                                o["me_cand"].data_ptr(), o["total_cand"].data_ptr(), o["rc"].data_ptr())
            sb.check(lib.svt_b200_me_picture(C.byref(params), C.byref(s), refs, C.byref(outs), scratch[i].data_ptr(), sp), lib)

    def step_resident(k):
        me_calls(dev_sets[k % RING])

    def step_e2e(k):
        hs = host_sets[k % RING]
        with torch.cuda.stream(stream):
            for j, hp in enumerate(hs):  # H2D of the step's pictures (new frames + the references around them)
                for a, b in zip(e2e_dev[j], hp):
                    a.copy_(b, non_blocking=True)
        me_calls(e2e_dev)
        with torch.cuda.stream(stream):
            for i in range(FRAMES_PER_STEP):
                for kk, t in h_out[i].items():
                    t.copy_(d_out[i][kk].view(t.dtype) if d_out[i][kk].dtype != t.dtype else d_out[i][kk], non_blocking=True)

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
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, launches

    sampler = ClockSampler(local)
    sampler.start()
    ms, launches = timed(step_resident, args.steps, args.warmup)
    ms_e2e, _ = timed(step_e2e, args.steps, args.warmup)
    sampler.stop_flag = True
    sampler.join(timeout=2)

    # ---- roofline of the dominant kernel (full-pel search), measured live with CUDA events ----
    pk, pk_kind = peaks()
    roof = kernel_roofline(lib, sb, params, dev_sets[0], d_out[0], scratch[0], stream, pk, pk_kind)

    frames = FRAMES_PER_STEP * args.steps * world
    value = frames / (ms / 1e3)
    e2e_v = frames / (ms_e2e / 1e3)
    line = {"metric": "1080p30 8-bit preset-8 hot-path fps", "value": value, "unit": "frames/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": workload_config(FRAMES_PER_STEP), "clocks": sampler.summary(),
            "e2e": {"value": e2e_v, "unit": "frames/s", "h2d_bytes_per_step": bytes_in_frame * (FRAMES_PER_STEP + 4),
                    "d2h_bytes_per_step": bytes_out_frame * FRAMES_PER_STEP},
            "gpu_launches": int(launches), "roofline": roof}
    if rank == 0:
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline()
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def kernel_roofline(lib, sb, params, dset, o, scratch, stream, pk, pk_kind):
    """Times whole ME pictures back to back (3 launches each) and attributes the algorithmic bytes of
    DESIGN.md §ME: per SB-ref the 64x64 source block + the (64+saw-1)x(64+sah-1) window, per SB the results."""
    import torch
    past, fut = ref_indices(0)
    r = [dset[j] for j in past] + [dset[past[-1]]] * 2 + [dset[j] for j in fut] + [dset[fut[-1]]] * 2
    refs = (sb.MePlanes * 8)(*[sb.MePlanes(x[0].data_ptr(), x[1].data_ptr(), x[2].data_ptr()) for x in r])
    s = sb.MePlanes(dset[2][0].data_ptr(), dset[2][1].data_ptr(), dset[2][2].data_ptr())
    outs = sb.MeOutputs(o["best_sad"].data_ptr(), o["best_mv"].data_ptr(), o["hme"].data_ptr(), o["me_mv"].data_ptr(),
                        o["me_cand"].data_ptr(), o["total_cand"].data_ptr(), o["rc"].data_ptr())
    sp = C.c_void_p(stream.cuda_stream)
    n = 20
    for _ in range(3):
        lib.svt_b200_me_picture(C.byref(params), C.byref(s), refs, C.byref(outs), scratch.data_ptr(), sp)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(stream)
    for _ in range(n):
        lib.svt_b200_me_picture(C.byref(params), C.byref(s), refs, C.byref(outs), scratch.data_ptr(), sp)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    n_sb = ((W + 63) // 64) * ((H + 63) // 64)
    nref = N_L0 + N_L1
    # algorithmic bytes of one ME picture: source planes once + reference planes once per ref + results
    plane = sum(t.numel() for t in dset[0])
    alg = plane * (1 + nref) + n_sb * (85 * 7 * 4 + 85 * 23 + 85 + 4) + n_sb * 8 * 85 * 8
    ach = alg / (ms / 1e3) / 1e9
    return {"kernel": "me_picture (hme_kernel + fullpel_kernel + finalize_kernel)", "bound": "hbm",
            "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach / pk["hbm_gbs"], "traffic": None,
            "peak_source": pk_kind, "ms_per_launch_group": ms,
            "note": "full-search ME is integer-ALU/shared-memory bound (SURVEY §8d); HBM fraction reported as required"}


def cpu_baseline():
    """Reference C path (oracle/_ref) on the host cores, bounded sample (rank 0, N=1 only)."""
    import common as cm
    from concurrent.futures import ThreadPoolExecutor
    if not cm.have_ref():
        return {"value": None, "unit": "frames/s", "cores": 0, "kind": "reference", "sample": "oracle/_ref not built"}
    cores = os.cpu_count() or 1
    geos, pics = make_inputs(1234)
    nfr = max(1, min(FRAMES_PER_STEP, cores))

    def one(i):
        past, fut = ref_indices(i)
        refs = [pics[j] for j in past] + [pics[past[-1]]] * 2 + [pics[j] for j in fut] + [pics[fut[-1]]] * 2
        cm.run_ref_me(W, H, 8, N_L0, N_L1, DIST, 2, 1, geos, pics[i + 2], refs)

    with ThreadPoolExecutor(cores) as ex:
        t0 = time.perf_counter()
        list(ex.map(one, range(nfr)))
        dt = time.perf_counter() - t0
    return {"value": nfr / dt, "unit": "frames/s", "cores": cores, "kind": "reference",
            "sample": f"{nfr} frames of the 1080p workload, one per thread, unmodified reference C (-O2, no SIMD)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        if int(os.environ.get("RANK", 0)) == 0:
            run_reference(args)
        return
    run_b200(args)


if __name__ == "__main__":
    main()
