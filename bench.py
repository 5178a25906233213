#!/usr/bin/env python
"""bench.py - BASELINE.json's metric: 1080p30 8-bit preset-8 ENCODE fps (configs[1]) of the reference encoder with the
B200 backend bound in, next to the reference's own SIMD (SSE2..AVX2/AVX-512, "minus asm") CPU encoder on the host cores.

What runs:
  * The encoder is the reference's own library built by integration/Makefile with five hooked files: with SVT_CUDA=1 the
    open-loop motion estimation, deblocking and CDEF (search + apply) of every picture run in libsvtav1_b200.so through its
    picture engine (host pictures in, host results out); everything else (mode decision, entropy coding, ...) is the
    reference's AVX2/AVX-512 code.  The bitstream is bit-identical to the CPU encoder's (md5 in the line).
  * A "step" is one mini-GOP of FRAMES_PER_STEP = 16 pictures of the stream; W warm-up steps fill the encoder pipeline
    (look-ahead, temporal filtering of the first pictures), then exactly K steps are timed between output packets.
  * value : frames/s over the K timed steps through the public API (EbSvtAv1Enc.h: send_picture from a clip resident in
    host RAM -> get_packet), integration/enc_bench.c.  e2e : the reference's own application (SvtAv1EncApp, clip file on
    tmpfs -> .ivf), its "Average Speed" over the whole run including pipeline fill - the contract's literal metric
    (SURVEY.md 8d).  The encoder's inputs are host pictures by API contract, so both cross PCIe every picture
    (h2d/d2h bytes from the engine's counters).
  * roofline / extra.hot_path : the device-resident kernel chain (bench_hotpath.py: CUDA events on the launch stream) -
    per-stage kernel times against the measured HBM peak; explains the kernels, not the encoder.
  * --impl reference : the reference's SIMD CPU encoder (oracle/_ref/simd), N concurrent streams on all host cores.
Multi-GPU (torchrun): one independent stream per rank / GPU (BASELINE configs[3]), each encoder process bound to its
GPU's NUMA node; no data-path collective; value = all frames / max over ranks of the timed window ("weak" scaling).
"""
import argparse
import hashlib
import json
import os
import re
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.join(ROOT, "svt-av1_b200"), os.path.join(ROOT, "tools"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

# --config: "1080p" = BASELINE configs[1] (the headline, default); "2160p10" = configs[2] (3840x2160 10-bit preset 6: the
# deblocking level search, CDEF on 16-bit planes and - still on the CPU - loop restoration are on)
CONFIGS = {"1080p": dict(W=1920, H=1080, BITS=8, PRESET=8, QP=43, METRIC="1080p30 8-bit preset-8 encode fps",
                         NAME="BASELINE configs[1]: 1920x1080 8-bit yuv420p, preset 8, CQP qp=43"),
           "2160p10": dict(W=3840, H=2160, BITS=10, PRESET=6, QP=43, METRIC="2160p 10-bit preset-6 encode fps",
                           NAME="BASELINE configs[2]: 3840x2160 10-bit yuv420p10le, preset 6, CQP qp=43")}
W, H, BITS, PRESET, QP = 1920, 1080, 8, 8, 43
CONFIG_NAME = CONFIGS["1080p"]["NAME"]
FRAMES_PER_STEP = 16  # one mini-GOP of the 4-level hierarchical structure presets 6-8 use
METRIC = "1080p30 8-bit preset-8 encode fps"


def select_config(name):
    global W, H, BITS, PRESET, QP, METRIC, CONFIG_NAME
    c = CONFIGS[name]
    W, H, BITS, PRESET, QP, METRIC, CONFIG_NAME = c["W"], c["H"], c["BITS"], c["PRESET"], c["QP"], c["METRIC"], c["NAME"]
B = os.path.join(ROOT, "integration", "_build")
BIN = {"cuda": os.path.join(B, "enc_bench_cuda_simd"), "ref": os.path.join(B, "enc_bench_ref_simd"),
       "app_cuda": os.path.join(B, "SvtAv1EncAppCudaSimd")}
_REAL_STDOUT = None
# N > 1 streams: each stream is confined to its own group of cores (cpu_partition) and is told so with `--lp <group size>`
# (the reference's own switch: thread / picture pools sized for the group) - the configuration of the committed N = 2 / 4
# lines (profiles/bench_r2c_*).  SVTB200_BENCH_LP=0 leaves the encoder's sizing for the whole host in place.
USE_LP = os.environ.get("SVTB200_BENCH_LP", "1") == "1"


def emit(obj):
    data = (json.dumps(obj) + "\n").encode()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, data)


def shm_dir():
    d = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else "/tmp"
    d = os.path.join(d, "svtb200_bench")
    os.makedirs(d, exist_ok=True)
    return d


def clip_path(stream, frames):
    return os.path.join(shm_dir(), "c2_%dx%d_%db_s%d_%df.yuv" % (W, H, BITS, stream, frames))


def make_clip(stream, frames):
    """Synthetic 1080p clip of stream `stream` (SURVEY 8d generator, per-stream seed), cached on tmpfs."""
    import make_yuv
    path = clip_path(stream, frames)
    want = frames * (W * H + 2 * (W // 2) * (H // 2)) * (2 if BITS > 8 else 1)
    if not (os.path.exists(path) and os.path.getsize(path) == want):
        tmp = path + ".tmp%d" % os.getpid()
        make_yuv.write_clip(tmp, W, H, frames, BITS, seed=1234 + 1000 * stream)
        os.replace(tmp, path)
    return path


def md5_file(path):
    h = hashlib.md5()
    with open(path, "rb") as f:
        for blk in iter(lambda: f.read(1 << 22), b""):
            h.update(blk)
    return h.hexdigest()


def gpu_numa_cpus(idx):
    """CPUs local to GPU idx (sysfs local_cpulist of its PCI function), or None."""
    try:
        bus = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(idx)],
                             stdout=subprocess.PIPE, timeout=20).stdout.decode().strip().lower()
        bus = bus[-12:] if len(bus) > 12 else bus  # 00000000:17:00.0 -> 0000:17:00.0
        txt = open("/sys/bus/pci/devices/%s/local_cpulist" % bus).read().strip()
        cpus = set()
        for part in txt.split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        return sorted(cpus) or None
    except Exception:
        return None


def cpu_partition(n, r):
    """CPUs of stream r when n streams share the host: the physical cores (with their hyper-thread siblings) are split
    into n disjoint, equal groups in (socket, core) order - the reference's documented way to run several jobs on one
    server is `--lp <cpus per job>` with unpinned threads (EbAppConfig.c:887-894); the affinity keeps the jobs apart.
    n == 1: None (no restriction beyond the caller's)."""
    if n <= 1:
        return None
    try:
        cores = {}
        base = "/sys/devices/system/cpu"
        for d in os.listdir(base):
            if not re.fullmatch(r"cpu\d+", d):
                continue
            t = os.path.join(base, d, "topology")
            if not os.path.exists(os.path.join(t, "core_id")):
                continue
            key = (int(open(os.path.join(t, "physical_package_id")).read()), int(open(os.path.join(t, "core_id")).read()))
            cores.setdefault(key, []).append(int(d[3:]))
        allowed = os.sched_getaffinity(0)
        groups = [sorted(c for c in v if c in allowed) for _, v in sorted(cores.items())]
        groups = [g for g in groups if g]
        per = max(1, len(groups) // n)
        mine = groups[r * per:(r + 1) * per]
        cpus = sorted(c for g in mine for c in g)
        return cpus or None
    except Exception:
        return None


def start_enc(kind, clip, frames, warm_frames, out=None, env_extra=None, cpus=None):
    env = dict(os.environ)
    for k in list(env):
        if k.startswith("SVT_CUDA"):
            del env[k]
    env.update(env_extra or {})
    if cpus and "ENC_BENCH_LP" not in env and env.get("SVTB200_STREAMS", "1") != "1" and USE_LP:
        env["ENC_BENCH_LP"] = str(len(cpus))  # thread / segment counts sized for this stream's share of the host
    cmd = [BIN[kind], clip, str(W), str(H), str(frames), str(BITS), str(PRESET), str(QP), str(warm_frames)] + ([out] if out else [])

    def pre():
        if cpus:
            try:
                os.sched_setaffinity(0, cpus)
            except OSError:
                pass
    return subprocess.Popen(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, preexec_fn=pre)


def finish_enc(p, what):
    out, err = p.communicate()
    if p.returncode != 0:
        sys.stderr.write(err.decode(errors="replace")[-3000:])
        raise RuntimeError("%s failed (rc %d) - no CPU fallback" % (what, p.returncode))
    res = json.loads([l for l in out.decode().splitlines() if l.startswith("{")][-1])
    res["log"] = [l for l in err.decode(errors="replace").splitlines() if "CUDA profile" in l or "CUDA backend" in l]
    return res


def parse_engine(log):
    """H2D / D2H megabytes and kernel launches of the whole encode from the backend's SVT_CUDA_PROFILE line."""
    for l in log:
        m = re.search(r"H2D ([0-9.]+) MB, D2H ([0-9.]+) MB, pinned ([0-9.]+) MB, (\d+) kernel launches", l)
        if m:
            return float(m.group(1)) * 1e6, float(m.group(2)) * 1e6, int(m.group(4)), l
    return None, None, None, None


def app_average_speed(clip, frames, env_extra, cpus, app=None):
    """The reference's own application on the same clip: file on tmpfs -> .ivf; returns ("Average Speed" fps, ivf path)."""
    ivf = os.path.join(shm_dir(), "out_%d.ivf" % os.getpid())
    env = dict(os.environ)
    env.update(env_extra)
    cmd = [app or BIN["app_cuda"], "-i", clip, "-w", str(W), "-h", str(H), "--fps", "30", "--preset", str(PRESET), "--rc", "0", "-q", str(QP),
           "-n", str(frames), "-b", ivf] + (["--input-depth", str(BITS)] if BITS != 8 else [])
    if cpus and env.get("SVTB200_STREAMS", "1") != "1" and USE_LP:
        cmd += ["--lp", str(len(cpus))]

    def pre():
        if cpus:
            try:
                os.sched_setaffinity(0, cpus)
            except OSError:
                pass
    p = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, preexec_fn=pre)
    txt = p.stdout.decode(errors="replace")
    m = re.search(r"Average Speed:\s+([0-9.]+) fps", txt)
    if p.returncode != 0 or not m:
        sys.stderr.write(txt[-2000:])
        raise RuntimeError("SvtAv1EncApp (CUDA build) failed")
    return float(m.group(1)), ivf


def ivf_payload_md5(path):
    import struct
    d = open(path, "rb").read()
    pos = struct.unpack("<H", d[6:8])[0]
    h = hashlib.md5()
    while pos < len(d):
        n = struct.unpack("<I", d[pos:pos + 4])[0]
        h.update(d[pos + 12:pos + 12 + n])
        pos += 12 + n
    return h.hexdigest()


def workload(frames, world):
    return {"workload": "%s, %d synthetic frames per stream (the BASELINE clip has 300), %d stream(s) one per GPU"
                        % (CONFIG_NAME, frames, world),
            "frames_per_step": FRAMES_PER_STEP, "streams": world,
            "stages_on_gpu": ["open-loop ME (HME + full-pel + candidate construction)",
                              "deblocking (frame level" + (", with the level search svt_av1_pick_filter_level)" if PRESET <= 6 else ")"),
                              "CDEF strength search", "CDEF frame apply"],
            "stages_on_cpu": "mode decision / EncDec, TPL, temporal filtering, global motion, entropy coding" +
                             (", loop restoration (search + apply)" if PRESET <= 6 else "") + " (reference AVX2/AVX-512 code)",
            "l2": "every picture is new data: %.1f MB per picture streamed from the clip (the ring of pictures in flight >> 126 MB L2)"
                  % ((W * H * 3 // 2) * (2 if BITS > 8 else 1) / 1e6),
            "timing": "host monotonic clock between output packets (the encoder pipeline is host-driven); kernel-level times in "
                      "roofline / extra.hot_path are CUDA events on the launch stream"}


def run_reference(args):
    """N concurrent streams through the reference's SIMD CPU encoder on all host cores (rank 0 only)."""
    n = max(1, args.gpus)
    frames, warm = (args.steps + args.warmup) * FRAMES_PER_STEP, args.warmup * FRAMES_PER_STEP
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=min(n, 8)) as ex:
        clips = list(ex.map(lambda s: make_clip(s, frames), range(n)))
    out0 = os.path.join(shm_dir(), "ref_%d.obu" % os.getpid())
    parts = [cpu_partition(n, s) for s in range(n)]
    procs = [start_enc("ref", clips[s], frames, warm, out=out0 if s == 0 else None, cpus=parts[s],
                       env_extra={"SVTB200_STREAMS": str(n)}) for s in range(n)]
    res = [finish_enc(p, "reference encoder (stream %d)" % s) for s, p in enumerate(procs)]
    secs = max(r["seconds_timed"] for r in res)
    timed_frames = (frames - warm) * n
    cores = os.cpu_count() or 1
    val = timed_frames / secs
    sample = ("%d concurrent stream(s) x %d frames (%d warm-up + %d timed) of this line's config.workload through the reference "
              "encoder's own SSE2..AVX2/AVX-512 code paths (asm level picked by its cpuid dispatch), default threading on all "
              "%d logical CPUs%s; built without the 13 nasm files (36 non-hot-path symbols forwarded to C, oracle/simd_asm_shim.c)"
              % (n, frames, warm, frames - warm, cores,
                 "" if n == 1 else " split into %d disjoint core groups of %d CPUs, one per stream, --lp %d unpinned"
                 % (n, len(parts[0] or []), len(parts[0] or []))))
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "frames/s", "n_gpus": n, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": secs * 1e3 / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8" if BITS == 8 else "u16", "data": "synthetic", "config": workload(frames, n),
            "cpu_baseline": {"value": val, "unit": "frames/s", "cores": cores, "kind": "reference-avx2-minus-asm", "sample": sample},
            "e2e": {"value": val, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "bitstream_md5_stream0": md5_file(out0),
            "per_stream_fps": [round((frames - warm) / r["seconds_timed"], 3) for r in res],
            "fps_whole_run": [r["fps_all"] for r in res]}
    os.remove(out0)
    # the like-for-like partner of the CUDA arm's `e2e` (the application's whole-run "Average Speed": initialisation, file
    # input, pipeline fill and drain included), measured with the reference's own SIMD application on stream 0's clip
    ref_app = os.path.join(ROOT, "oracle", "_ref", "app", "SvtAv1EncAppSimd")
    if n == 1 and os.path.exists(ref_app):
        try:
            fps, ivf = app_average_speed(clips[0], frames, {}, None, app=ref_app)
            os.remove(ivf)
            line["app_average_speed_fps"] = fps
        except Exception as e:  # noqa: BLE001
            line["app_average_speed_fps"] = None
            sys.stderr.write("reference application run failed: %s\n" % e)
    emit(line)


def hot_path(args_steps=6):
    """The device-resident kernel chain (bench_hotpath.py) in a child process: roofline + stage times."""
    cmd = [sys.executable, os.path.join(ROOT, "bench_hotpath.py"), "--gpus", "1", "--steps", str(args_steps), "--warmup", "3", "--no-cpu"]
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=1500)
    lines = [l for l in p.stdout.decode().splitlines() if l.startswith("{")]
    if p.returncode != 0 or not lines:
        sys.stderr.write(p.stderr.decode(errors="replace")[-3000:])
        raise RuntimeError("bench_hotpath.py failed")
    return json.loads(lines[-1])


def encoder_roofline(hp_roof):
    """The roofline object for the DOMINANT kernel of the encoder run: shares of GPU time from the committed ncu launch list
    of the encoder (profiles/r2_launches_encoder_1080p.csv: cdef_search_grid_kernel 51 %, fullpel 21 %, hme 17 %, cdef_apply 7 %,
    deblocking 4 %), its device time measured live in this run by the hot-path child with CUDA events (stage times)."""
    import collections
    import csv
    share = collections.Counter()
    try:
        for r in csv.reader(open(os.path.join(ROOT, "profiles", "r2_launches_encoder_1080p.csv"))):
            if len(r) > 14 and r[0].isdigit():
                share[r[4].split("(")[0].split("::")[-1].split("<")[0].strip()] += float(r[14])
    except Exception:
        pass
    stages = hp_roof["stages"]
    tot = sum(share.values())
    for st in stages:
        st["share_of_encoder_gpu_time"] = round(sum(share.get(k, 0.0) for k in st["kernels"]) / tot, 3) if tot else None
    enc_stages = [st for st in stages if st["stage"] != "encdec"]  # the EncDec kernel is not bound into the encoder yet
    dom = max(enc_stages, key=lambda st: st["share_of_encoder_gpu_time"] or 0.0) if tot else max(enc_stages, key=lambda st: st["ms_per_frame"])
    out = dict(hp_roof)
    out.update({"kernel": "%s (stage '%s': %.0f %% of the encoder's GPU time in the committed ncu launch list)"
                          % (" + ".join(dom["kernels"]), dom["stage"], 100 * (dom["share_of_encoder_gpu_time"] or 0)),
                "bound": "hbm", "achieved": dom["achieved"], "frac": dom["frac"], "traffic": dom["traffic"], "alu": dom.get("alu"),
                "binds_on": dom["binding"], "stages": stages,
                "note": "the dominant kernels of this path (CDEF strength search, ME) are integer-ALU bound by their arithmetic (SURVEY 8d): "
                        "`alu` is the meaningful fraction for them, the HBM fraction is reported because the contract asks for it; the "
                        "streaming stage is deblocking (tools/kernel_bench.py: 12 % of HBM peak at 1080p where a plain copy of the same "
                        "picture reaches 26 %, 28 % at 2160p 10-bit where the copy reaches 74 %).  The shares come from the launch list "
                        "committed before the CDEF strength decision moved to the device: cdef_decide_kernel adds one single-CTA launch "
                        "per picture (0.9 ms at 1080p on 1 of 148 SMs, profiles/r2g_kernel_bench.json) - a latency chain, not a "
                        "throughput kernel, so it is not the kernel a roofline describes"})
    return out


def kernel_roofline():
    """configs[2] geometry: tools/kernel_bench.py (CUDA-graph ring, device ms per picture of every filter entry)."""
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "kernel_bench.py")], stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=900)
    if p.returncode != 0:
        sys.stderr.write(p.stderr.decode(errors="replace")[-2000:])
        raise RuntimeError("tools/kernel_bench.py failed")
    kb = json.loads(p.stdout.decode())
    g = [x for x in kb["geometries"] if x["geometry"].startswith("%dx%d" % (W, H))][0]
    rows = [r for r in g["rows"] if not r["entry"].startswith("  ")]
    dom = max(rows, key=lambda r: r["ms"])
    return {"kernel": dom["entry"] + " (dominant picture-level entry at this geometry)", "bound": "hbm", "achieved": dom["achieved_gbs"],
            "peak": kb["hbm_peak_gbs"], "unit": "GB/s", "frac": dom["frac_of_hbm_peak"], "traffic": None,
            "same_size_copy": g["same_size_copy"], "stages": rows,
            "note": "CDEF search is integer-ALU bound (10 filter evaluations per sample); deblocking and the CDEF apply are the streaming "
                    "entries; `same_size_copy` is a plain device copy of one picture through the same launch path (the ceiling for a "
                    "launch of this size)"}


def cpu_baseline_sample():
    """Bounded sample of the same workload on the host cores: the reference's SIMD encoder on a prefix of stream 0."""
    frames, warm = (8 * FRAMES_PER_STEP, 2 * FRAMES_PER_STEP) if W <= 1920 else (3 * FRAMES_PER_STEP, FRAMES_PER_STEP)
    clip = make_clip(0, frames)
    r = finish_enc(start_enc("ref", clip, frames, warm), "reference encoder (cpu_baseline)")
    cores = os.cpu_count() or 1
    return {"value": (frames - warm) / r["seconds_timed"], "unit": "frames/s", "cores": cores, "kind": "reference-avx2-minus-asm",
            "sample": "%d frames (%d warm-up + %d timed) of stream 0 through the reference encoder's SSE2..AVX2/AVX-512 paths, default "
                      "threading on all %d logical CPUs (whole-run average incl. pipeline fill: %.1f fps)"
                      % (frames, warm, frames - warm, cores, r["fps_all"])}


def run_b200(args):
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    frames, warm = (args.steps + args.warmup) * FRAMES_PER_STEP, args.warmup * FRAMES_PER_STEP
    clip = make_clip(rank, frames)
    # one stream: the CPUs local to the GPU; several streams: disjoint core groups (the same split the reference arm uses)
    cpus = gpu_numa_cpus(local) if world == 1 else cpu_partition(world, rank)
    env = {"SVT_CUDA": "1", "SVT_CUDA_DEVICE": str(local), "SVT_CUDA_PROFILE": "1", "SVTB200_STREAMS": str(world)}
    from bench_hotpath import ClockSampler
    sampler = ClockSampler(local)
    sampler.start()
    out0 = os.path.join(shm_dir(), "cuda_%d.obu" % os.getpid())
    if dist is not None:
        dist.barrier()
    r = finish_enc(start_enc("cuda", clip, frames, warm, out=out0, env_extra=env, cpus=cpus), "CUDA-backed encoder")
    secs = r["seconds_timed"]
    if dist is not None:
        import torch
        t = torch.tensor([secs], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        secs_max = float(t.item())
        dist.barrier()
    else:
        secs_max = secs
    # the application-level number (file -> .ivf, whole run), all ranks at once as well
    app_fps, ivf = app_average_speed(clip, frames, env, cpus)
    sampler.stop_flag = True
    sampler.join(timeout=2)
    app_md5 = ivf_payload_md5(ivf)
    os.remove(ivf)
    if dist is not None:
        import torch
        t = torch.tensor([frames / app_fps], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        app_secs_max = float(t.item())
    else:
        app_secs_max = frames / app_fps
    h2d, d2h, launches, eng_line = parse_engine(r["log"])
    md5 = md5_file(out0)
    os.remove(out0)
    if rank == 0:
        timed_frames = (frames - warm) * world
        line = {"metric": METRIC, "value": timed_frames / secs_max, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": secs_max * 1e3 / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "u8" if BITS == 8 else "u16", "data": "synthetic", "config": workload(frames, world),
                "clocks": sampler.summary(),
                "e2e": {"value": frames * world / app_secs_max, "unit": "frames/s",
                        "h2d_bytes_per_step": int(h2d / (args.steps + args.warmup)) if h2d else None,
                        "d2h_bytes_per_step": int(d2h / (args.steps + args.warmup)) if d2h else None,
                        "what": "SvtAv1EncApp (reference application, CUDA-backed library): clip file on tmpfs -> .ivf, 'Average Speed' "
                                "over the whole run incl. initialisation, pipeline fill and drain; value above = the K timed steps through the API "
                                "(the measurement the reference arm's value / e2e is).  Like for like with this number: the reference "
                                "arm's `app_average_speed_fps` (the reference's own SIMD application, same clip)"},
                "gpu_launches": int(launches * args.steps / (args.steps + args.warmup)) if launches else 0,
                "bitstream_md5_stream0": md5, "bitstream_md5_matches_app": md5 == app_md5,
                "encoder": {"fps_whole_run_api": r["fps_all"], "init_s": r["init_s"], "numa_cpus": len(cpus) if cpus else None,
                            "engine": eng_line}}
        if world == 1 and not args.no_hotpath:
            if args.config == "1080p":
                try:
                    hp = hot_path()
                    line["roofline"] = encoder_roofline(hp["roofline"])
                    line["extra"] = {"hot_path": {k: hp[k] for k in ("metric", "value", "unit", "ms_per_step", "e2e", "gpu_launches",
                                                                   "stage_ms_per_frame", "config") if k in hp}}
                except Exception as ex:  # noqa: BLE001 - the encode-fps line must survive a failure of the kernel-chain child
                    sys.stderr.write("hot-path child failed: %s\n" % ex)
                    try:
                        old = json.loads(open(os.path.join(ROOT, "profiles", "r2d_bench_n1.json")).read().strip().splitlines()[-1])
                        line["roofline"] = dict(old["roofline"], live=False,
                                                note_not_live="the kernel-chain child failed in this run: values of profiles/r2d_bench_n1.json")
                    except Exception:  # noqa: BLE001
                        line["roofline"] = None
            else:
                line["roofline"] = kernel_roofline()
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline_sample()
        emit(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)  # 20 x 16 = 320 timed frames: BASELINE configs[1] is a 300-frame clip
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-hotpath", action="store_true", help="skip the device-resident kernel chain (roofline)")
    ap.add_argument("--config", default="1080p", choices=sorted(CONFIGS), help="1080p = BASELINE configs[1] (default), 2160p10 = configs[2]")
    args = ap.parse_args()
    select_config(args.config)
    # stdout carries exactly ONE line (the JSON result): anything libraries print to fd 1 is routed to stderr
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    for k in ("cuda", "ref", "app_cuda"):
        if not os.path.exists(BIN[k]):
            raise SystemExit("bench.py: %s is missing - run `python __graft_entry__.py` (build) first; there is no CPU fallback" % BIN[k])
    if args.impl != "reference":
        # the product library is loaded by this process too (and by every encoder child): a missing GPU is an error
        import ctypes
        lib = ctypes.CDLL(os.path.join(ROOT, "svt-av1_b200", "libsvtav1_b200.so"))
        if lib.svt_b200_device_count() <= 0:
            raise SystemExit("bench.py: no CUDA device - the B200 arm has no CPU fallback")
    if args.impl == "reference":
        if int(os.environ.get("RANK", 0)) == 0:
            run_reference(args)
        return
    run_b200(args)


if __name__ == "__main__":
    main()
